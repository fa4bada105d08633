"""CPU oracle for GeoT's segment-reduction hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this package, and only as the checker or as the timed CPU baseline.
``geot_b200`` (the product) never imports it and has no CPU fallback.

Two layers:

* ``geot_oracle.c`` (ctypes, ``oracle/_build/libgeot_oracle.so``): plain-C restatement of the
  reference's sequential definitions (``/root/reference/csrc/util/check.cuh:77-111``), the reference
  CPU kernel's segment-pointer pass (``csrc/cpu/index_scatter_cpu.cpp:36-75``) and ``coo_to_csr``
  (``geot/match_replace/format_transform.py:5-18``).
* ``torch_*`` functions: the torch formulas the reference's own tests compare against
  (``test/test_index_scatter.py:16-22``, ``test/test_gather_weight_scatter.py:4-11``,
  ``test/test_mh_spmm.py:4-10``) and ``scatter_reduce(include_self=False)`` for mean/max/min.

Parity status: pinned -- see the header of ``geot_oracle.c`` and ``tests/test_oracle.py``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_LIB_PATH = os.path.join(_HERE, "_build", "libgeot_oracle.so")
REF_SEQ_PATH = os.path.join(_HERE, "_ref", "libref_seq.so")
REF_EXT_PATH = os.path.join(_HERE, "_ref", "geot_ref_C.so")

REDUCE_ENUM = {"sum": 0, "mean": 1, "max": 2, "amax": 2, "min": 3, "amin": 3, "prod": 4}

_lib = None


def build(force: bool = False) -> None:
    """Compile the C restatement (and, where /root/reference exists, the reference shim)."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "geot_oracle.c"))
    ):
        subprocess.check_call(["make", "-f", "oracle/Makefile"], cwd=_ROOT, stdout=subprocess.DEVNULL)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.geot_oracle_segment_ptr.restype = ctypes.c_int64
        _lib.geot_oracle_reduce_f32_mt.restype = ctypes.c_int
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def segment_ptr(index: torch.Tensor):
    """(row_index[M], row_offset[M+1]) of a sorted int64 index (index_scatter_cpu.cpp:36-75)."""
    index = index.contiguous().cpu()
    E = index.numel()
    row_index = torch.empty(max(E, 1), dtype=torch.int64)
    row_offset = torch.empty(E + 1, dtype=torch.int64)
    m = lib().geot_oracle_segment_ptr(_p(index), ctypes.c_int64(E), _p(row_index), _p(row_offset))
    return row_index[:m].clone(), row_offset[: m + 1].clone()


def rowptr(index: torch.Tensor, S: int) -> torch.Tensor:
    """CSR rowptr[S+1] of a sorted COO row index (format_transform.py:5-18)."""
    index = index.contiguous().cpu()
    out = torch.empty(S + 1, dtype=torch.int64)
    lib().geot_oracle_rowptr(_p(index), ctypes.c_int64(index.numel()), ctypes.c_int64(S), _p(out))
    return out


def set_threads(n: int = 0) -> int:
    """Sets (n > 0) and returns the OpenMP thread count of the threaded baseline variant."""
    return int(lib().geot_oracle_set_threads(ctypes.c_int(n)))


def segment_reduce(src, src_index, dst_index, weight, reduce="sum", *, S=None, H=1,
                   weight_transposed=False, acc64=False, threads=False):
    """dst[dst_index[e]] (op)= weight[e,h] * src[src_index[e]]   on CPU tensors.

    src: [N, W] or [N, H, F]; weight: None, [E], [E,H] or (weight_transposed) [H,E].
    bf16/fp16 inputs are upcast to fp32, accumulated in fp32 and rounded once at the end
    (mirrors the reference CPU kernel's ``need_acc`` buffer, index_scatter_cpu.cpp:78-86,114-116).
    ``acc64`` accumulates fp32 data in fp64 (the tight-tolerance oracle for sum/mean).
    ``threads`` uses the OpenMP row-parallel variant (sorted dst_index, fp32; timing baseline).
    """
    red = REDUCE_ENUM[reduce]
    out_dtype = src.dtype
    dst_index = dst_index.contiguous().cpu()
    E = dst_index.numel()
    if S is None:
        S = int(dst_index.max()) + 1      # == dst_index[-1] + 1 for sorted input
    shp = list(src.shape)
    N = shp[0]
    W = int(np.prod(shp[1:])) if len(shp) > 1 else 1
    assert W % H == 0
    F = W // H
    comp = torch.float64 if src.dtype == torch.float64 else torch.float32
    src_c = src.detach().cpu().to(comp).contiguous().view(N, W)
    w_c = None
    ws_e, ws_h = 0, 0
    if weight is not None:
        w_c = weight.detach().cpu().to(comp).contiguous()
        if w_c.dim() == 1:
            assert H == 1
            ws_e, ws_h = 1, 0
        elif weight_transposed:
            assert tuple(w_c.shape) == (H, E)
            ws_e, ws_h = 1, E
        else:
            assert tuple(w_c.shape) == (E, H)
            ws_e, ws_h = H, 1
    si = src_index.contiguous().cpu() if src_index is not None else None
    dst = torch.empty(S, W, dtype=comp)
    if comp == torch.float64:
        fn = lib().geot_oracle_reduce_f64
    elif threads:
        fn = lib().geot_oracle_reduce_f32_mt
    elif acc64:
        fn = lib().geot_oracle_reduce_f32_acc64
    else:
        fn = lib().geot_oracle_reduce_f32
    fn(_p(src_c), _p(si), _p(dst_index), _p(w_c), _p(dst), ctypes.c_int64(E), ctypes.c_int64(S),
       ctypes.c_int64(F), ctypes.c_int64(H), ctypes.c_int64(ws_e), ctypes.c_int64(ws_h),
       ctypes.c_int(red))
    return dst.to(out_dtype).view([S] + shp[1:])


# ---- the four ops, reference argument order (geot/*.py) --------------------------------------

def index_scatter(dim, index, src, reduce="sum", sorted=True, **kw):
    """geot.index_scatter (csrc/index_scatter.cpp:26-39): only dim=0 is meaningful in the reference."""
    assert dim == 0
    return segment_reduce(src, None, index, None, reduce, **kw)


def gather_scatter(src_index, dst_index, src, reduce="sum", **kw):
    return segment_reduce(src, src_index, dst_index, None, reduce, **kw)


def gather_weight_scatter(src_index, dst_index, weight, src, reduce="sum", **kw):
    return segment_reduce(src, src_index, dst_index, weight, reduce, **kw)


def mh_spmm(src_index, dst_index, weight, src, reduce="sum", **kw):
    """weight [E,H] or [H,E] (layout by shape, wrapper/mh_spmm_base.h:38-49); src [N,H,F]."""
    E, H = dst_index.numel(), src.shape[1]
    if weight.shape[0] == E:
        return segment_reduce(src, src_index, dst_index, weight, reduce, H=H, **kw)
    if weight.shape[1] == E:
        return segment_reduce(src, src_index, dst_index, weight, reduce, H=H,
                              weight_transposed=True, **kw)
    raise RuntimeError("Invalid weight size")


def sddmm_coo(src_index, dst_index, mat_1, mat_2):
    """out[e] = <mat_1[dst_index[e]], mat_2[src_index[e]]> (geot::sddmm_coo_impl argument order,
    csrc/gather_weight_scatter.cpp:36-44); low precision upcast to fp32, f64 accumulate, rounded once."""
    out_dtype = mat_1.dtype
    comp = torch.float64 if out_dtype == torch.float64 else torch.float32
    a = mat_1.detach().cpu().to(comp).contiguous()
    b = mat_2.detach().cpu().to(comp).contiguous()
    row = dst_index.detach().cpu().long().contiguous()
    col = src_index.detach().cpu().long().contiguous()
    E, F = row.numel(), a.shape[1]
    out = torch.empty(E, dtype=comp)
    fn = lib().geot_oracle_sddmm_f64 if comp == torch.float64 else lib().geot_oracle_sddmm_f32
    fn(_p(a), _p(row), _p(b), _p(col), _p(out), ctypes.c_int64(E), ctypes.c_int64(F))
    return out.to(out_dtype)


def csr_gws(csrptr, csrind, weight, src):
    """geot.csr_gws (csrc/csr_gws.cpp:24-35): fp32, output rows = csrptr.numel() (nrow + 1, last row 0)."""
    ptr = csrptr.detach().cpu().long().contiguous()
    ind = csrind.detach().cpu().long().contiguous()
    val = weight.detach().cpu().float().contiguous()
    x = src.detach().cpu().float().contiguous()
    nrow, F = ptr.numel() - 1, x.shape[1]
    out = torch.empty(nrow + 1, F, dtype=torch.float32)
    lib().geot_oracle_csr_gws_f32(_p(ptr), _p(ind), _p(val), _p(x), _p(out), ctypes.c_int64(nrow), ctypes.c_int64(F))
    return out.to(src.dtype)


def csr_to_coo(csrptr):
    ptr = csrptr.detach().cpu().long().contiguous()
    row = torch.empty(int(ptr[-1]), dtype=torch.int64)
    lib().geot_oracle_csr_to_coo(_p(ptr), ctypes.c_int64(ptr.numel() - 1), _p(row))
    return row


# ---- torch formulas used by the reference's own tests -----------------------------------------

def torch_index_scatter(index, src, reduce="sum", S=None):
    S = int(index[-1]) + 1 if S is None else S
    out = torch.zeros([S] + list(src.shape[1:]), dtype=src.dtype, device=src.device)
    if reduce == "sum":
        return out.index_add_(0, index, src)
    name = {"mean": "mean", "max": "amax", "amax": "amax", "min": "amin", "amin": "amin",
            "prod": "prod"}[reduce]
    idx = index.view([-1] + [1] * (src.dim() - 1)).expand_as(src)
    return out.scatter_reduce_(0, idx, src, name, include_self=False)


def torch_gather_scatter(src_index, dst_index, src, reduce="sum", S=None):
    return torch_index_scatter(dst_index, src.index_select(0, src_index), reduce, S)


def torch_gather_weight_scatter(src_index, dst_index, weight, src, reduce="sum", S=None):
    return torch_index_scatter(dst_index, weight.unsqueeze(-1) * src.index_select(0, src_index),
                               reduce, S)


def torch_mh_spmm(src_index, dst_index, weight_eh, src, reduce="sum", S=None):
    return torch_index_scatter(dst_index, weight_eh.unsqueeze(-1) * src.index_select(0, src_index),
                               reduce, S)


# ---- handles on the real reference (present only where oracle/_ref was built) ----------------

def ref_seq():
    """ctypes handle on the reference's sequential goldens (oracle/ref_seq_shim.cu), or None."""
    if not os.path.exists(REF_SEQ_PATH):
        return None
    return ctypes.CDLL(REF_SEQ_PATH)


def load_ref_extension() -> bool:
    """Load the unmodified reference extension as torch.ops.geot_ref.* (oracle/Makefile.ref)."""
    if not os.path.exists(REF_EXT_PATH):
        return False
    if not hasattr(torch.ops.geot_ref, "index_scatter"):
        torch.ops.load_library(REF_EXT_PATH)
    return True
