"""Pins the CPU oracle (oracle/) -- runs without a GPU.

1. against the committed golden vectors, whose expected outputs come from the reference
   (tests/golden/make_golden.py: the reference's test formulas, its C++ sequential goldens and index
   generator compiled from its sources, and its unmodified CPU kernel's segment counts);
2. against torch.scatter_reduce(include_self=False) for the reductions / dtypes no reference test covers;
3. live against oracle/_ref (the reference built here) when that directory exists.
"""
import ctypes
import os

import numpy as np
import pytest
import torch

import oracle


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    return {k: torch.from_numpy(z[k]) if z[k].ndim else z[k] for k in z.files}


def test_golden_ref_test_index_scatter(golden_dir):
    g = _load(golden_dir, "ref_test_index_scatter.npz")
    out = oracle.index_scatter(0, g["index"], g["src"], "sum")
    assert out.shape == g["expected"].shape
    # the reference's own tolerance (test/test_index_scatter.py:19) and the north-star one
    assert torch.allclose(out, g["expected"], atol=1e-4)
    assert ((out - g["expected"]).abs() <= 1e-5 * g["expected"].abs()).all()


def test_golden_ref_test_gather(golden_dir):
    g = _load(golden_dir, "ref_test_gather.npz")
    gs = oracle.gather_scatter(g["src_index"], g["dst_index"], g["src"])
    gws = oracle.gather_weight_scatter(g["src_index"], g["dst_index"], g["weight"], g["src"])
    assert torch.allclose(gs, g["expected_gs"], atol=1e-4)
    assert torch.allclose(gws, g["expected_gws"], atol=1e-4)
    assert ((gws - g["expected_gws"]).abs() <= 1e-5 * g["expected_gws"].abs().clamp_min(1e-30)).all()


def test_golden_ref_test_mh_spmm(golden_dir):
    g = _load(golden_dir, "ref_test_mh_spmm.npz")
    out = oracle.mh_spmm(g["src_index"], g["dst_index"], g["weight"], g["src"])
    assert torch.allclose(out, g["expected"], atol=1e-4)
    out_t = oracle.mh_spmm(g["src_index"], g["dst_index"], g["weight"].t().contiguous(), g["src"])
    assert torch.equal(out, out_t)


def test_golden_ctest_sequential_is_bit_exact(golden_dir):
    """Same sequential order as check.cuh:77-85 / :101-111 => identical bits."""
    g = _load(golden_dir, "ctest_segreduce.npz")
    out = oracle.index_scatter(0, g["index"], g["src"], "sum")
    assert torch.equal(out, g["expected"])
    # gws_sequencial uses src*weight then += ; gcc may or may not contract it in the reference build,
    # so allow 1 ulp-level differences
    gws = oracle.gather_weight_scatter(g["col"], g["index"], g["weight"], g["feat"])
    assert ((gws - g["expected_gws"]).abs() <= 1e-6 * g["expected_gws"].abs().clamp_min(1e-30)).all()
    assert int(g["dst_len"]) == oracle.segment_ptr(g["index"])[0].numel()


def test_golden_reference_cpu_kernel_segment_counts(golden_dir):
    g = _load(golden_dir, "ref_cpu_counts.npz")
    index, counts = g["index"], g["counts"]
    rows, offs = oracle.segment_ptr(index)
    S = int(index[-1]) + 1
    dense = torch.zeros(S, dtype=torch.int64)
    dense[rows] = offs[1:] - offs[:-1]
    assert torch.equal(dense, counts)
    rp = oracle.rowptr(index, S)
    assert torch.equal(rp[1:] - rp[:-1], counts)
    assert torch.equal(rp, torch.cat([torch.zeros(1, dtype=torch.int64), torch.bincount(index, minlength=S).cumsum(0)]))


@pytest.mark.parametrize("reduce", ["sum", "mean", "max", "min", "prod"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_oracle_matches_torch_restatement(reduce, dtype):
    g = torch.Generator().manual_seed(1)
    E, N, F = 3000, 211, 19
    si = torch.randint(0, N, (E,), generator=g)
    di = torch.randint(0, N, (E,), generator=g).sort().values
    w = (torch.rand(E, generator=g) + 0.5).to(dtype)
    src = (torch.rand(N, F, generator=g) + 0.5).to(dtype)
    got = oracle.gather_weight_scatter(si, di, w, src, reduce)
    exp = oracle.torch_gather_weight_scatter(si, di, w, src, reduce)
    if reduce in ("max", "min"):
        assert torch.equal(got, exp)
    else:
        assert torch.allclose(got, exp, rtol=1e-5 if dtype == torch.float32 else 1e-12, atol=0)
    got = oracle.index_scatter(0, di, src[si], reduce)
    exp = oracle.torch_index_scatter(di, src[si], reduce)
    assert torch.allclose(got, exp, rtol=1e-5, atol=0)


def test_oracle_nan_and_gaps():
    di = torch.tensor([1, 1, 4, 4, 4])
    src = torch.tensor([[1.0], [float("nan")], [2.0], [5.0], [3.0]])
    for red, name in [("max", "amax"), ("min", "amin")]:
        got = oracle.index_scatter(0, di, src, red)
        exp = torch.zeros(5, 1).scatter_reduce_(0, di[:, None], src, name, include_self=False)
        assert torch.equal(torch.isnan(got), torch.isnan(exp))
        assert torch.equal(torch.nan_to_num(got), torch.nan_to_num(exp))
    assert got[0].item() == 0 and got[2].item() == 0 and got[3].item() == 0


def test_oracle_low_precision_rounds_once():
    g = torch.Generator().manual_seed(2)
    E, N, H, F = 2000, 50, 4, 8
    si = torch.randint(0, N, (E,), generator=g)
    di = torch.randint(0, N, (E,), generator=g).sort().values
    w = torch.rand(E, H, generator=g).bfloat16()
    src = torch.rand(N, H, F, generator=g).bfloat16()
    got = oracle.mh_spmm(si, di, w, src)
    exp = oracle.torch_mh_spmm(si, di, w.float(), src.float()).bfloat16()
    assert got.dtype == torch.bfloat16
    assert torch.equal(got, exp)


def test_threaded_variant_is_bit_identical():
    g = torch.Generator().manual_seed(3)
    E, N, F = 20000, 700, 32
    si = torch.randint(0, N, (E,), generator=g)
    di = torch.randint(0, N, (E,), generator=g).sort().values
    w = torch.rand(E, generator=g)
    src = torch.rand(N, F, generator=g)
    for red in ["sum", "mean", "max"]:
        a = oracle.gather_weight_scatter(si, di, w, src, red)
        b = oracle.gather_weight_scatter(si, di, w, src, red, threads=True)
        assert torch.equal(a, b)


# ---- live checks against the reference built here (absent on the GPU box: skipped there) ----------
needs_ref_seq = pytest.mark.skipif(not os.path.exists(oracle.REF_SEQ_PATH), reason="oracle/_ref/libref_seq.so not built")
needs_ref_ext = pytest.mark.skipif(not os.path.exists(oracle.REF_EXT_PATH), reason="oracle/_ref/geot_ref_C.so not built")


@needs_ref_seq
def test_live_reference_sequential_goldens():
    seq = ctypes.CDLL(oracle.REF_SEQ_PATH)
    rs = np.random.RandomState(11)
    nnz, N, keys = 4000, 24, 300
    idx = np.sort(rs.randint(0, keys, size=nnz)).astype(np.int64)
    idx[-1] = keys - 1
    src = rs.rand(nnz, N).astype(np.float32)
    dst = np.zeros((keys, N), dtype=np.float32)
    seq.ref_segment_coo_sequencial_f32(src.ctypes.data_as(ctypes.c_void_p), idx.ctypes.data_as(ctypes.c_void_p),
                                       nnz, N, keys, dst.ctypes.data_as(ctypes.c_void_p))
    got = oracle.index_scatter(0, torch.from_numpy(idx), torch.from_numpy(src), "sum")
    assert torch.equal(got, torch.from_numpy(dst))


@needs_ref_ext
def test_live_reference_cpu_kernel_counts_and_errors():
    assert oracle.load_ref_extension()
    g = torch.Generator().manual_seed(5)
    index = torch.randint(0, 500, (5000,), generator=g).sort().values
    counts = torch.ops.geot_ref.index_scatter(0, index, torch.ones(5000, 1), "sum", True).reshape(-1).long()
    rp = oracle.rowptr(index, int(index[-1]) + 1)
    assert torch.equal(rp[1:] - rp[:-1], counts)
    with pytest.raises(RuntimeError, match="reduce argument must be either sum, prod, mean, amax or amin"):
        torch.ops.geot_ref.index_scatter(0, index, torch.ones(5000, 1), "bogus", True)
