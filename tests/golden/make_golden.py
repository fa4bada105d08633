"""Generates tests/golden/*.npz -- golden input/output vectors for the segment-reduction path.

Run HERE (the build container), where /root/reference exists:   python tests/golden/make_golden.py
The vectors travel to the GPU box; /root/reference does not.

Every expected output is produced by the REFERENCE, not by this repo's code:
  * ref_test_*     the formulas of the reference's own Python tests, with a fixed seed
                   (test/test_index_scatter.py:5-23, test_gather_scatter.py:4-27,
                   test_gather_weight_scatter.py:4-27, test_mh_spmm.py:4-28), evaluated with torch on CPU;
  * ctest_*        the reference's C++ test case `./test -r 1000 -nnz 5000 -min 2 -max 7 -cv 0.5 -N 32`
                   (test/ctest/test.cu:90): index from the reference's generateIndex
                   (csrc/dataloader/dataloader.hpp:21-62), expected output from the reference's
                   sequential goldens (csrc/util/check.cuh:77-111), both compiled from the reference
                   sources where they lie (oracle/ref_seq_shim.cu -> oracle/_ref/libref_seq.so);
  * ref_cpu_counts segment lengths observed through the UNMODIFIED reference CPU kernel
                   (csrc/cpu/index_scatter_cpu.cpp, built by oracle/Makefile.ref): with src = ones its
                   output is the per-row edge count whatever its src-row defect, which
                   pins the segment-pointer pass (:36-75).
"""
import ctypes
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402  (only for the paths of the reference builds)


def ref_spmm(src_index, dst_index, weight, src):
    # test/test_gather_weight_scatter.py:4-11
    sparse_size = int(dst_index[-1]) + 1
    adj = torch.sparse_coo_tensor(torch.stack([dst_index, src_index]), weight, (sparse_size, src.shape[0]))
    adj = adj.coalesce()
    return torch.sparse.mm(adj, src)


def main():
    g = torch.Generator().manual_seed(20240403)
    out = {}

    # --- test/test_index_scatter.py ---------------------------------------------------------------
    E, F = 1000, 32
    src = torch.rand(E, F, generator=g)
    index = torch.randint(0, 10, (E,), generator=g).sort().values
    keys = int(index[-1]) + 1
    ref = torch.zeros(keys, F).scatter_add_(0, index.unsqueeze(-1).expand_as(src), src)
    ref2 = torch.zeros(keys, F).index_add_(0, index, src)
    assert torch.allclose(ref, ref2, atol=1e-4)
    np.savez_compressed(os.path.join(HERE, "ref_test_index_scatter.npz"), src=src.numpy(), index=index.numpy(),
                        expected=ref.numpy())

    # --- test/test_gather_scatter.py, test_gather_weight_scatter.py -------------------------------
    N, E, F = 100, 1000, 32
    src_index = torch.randint(0, N, (E,), generator=g)
    dst_index = torch.randint(0, N, (E,), generator=g).sort().values
    weight = torch.rand(E, generator=g)
    src = torch.rand(N, F, generator=g)
    np.savez_compressed(os.path.join(HERE, "ref_test_gather.npz"), src_index=src_index.numpy(),
                        dst_index=dst_index.numpy(), weight=weight.numpy(), src=src.numpy(),
                        expected_gs=ref_spmm(src_index, dst_index, torch.ones(E), src).numpy(),
                        expected_gws=ref_spmm(src_index, dst_index, weight, src).numpy())

    # --- test/test_mh_spmm.py ----------------------------------------------------------------------
    H = 4
    src_index = torch.randint(0, N, (E,), generator=g)
    dst_index = torch.randint(0, N, (E,), generator=g).sort().values
    weight = torch.rand(E, H, generator=g)
    src3 = torch.rand(N, H, F, generator=g)
    mul = weight.unsqueeze(-1) * src3.index_select(0, src_index)
    exp = torch.zeros_like(src3).index_add(0, dst_index, mul)[: int(dst_index[-1]) + 1]
    np.savez_compressed(os.path.join(HERE, "ref_test_mh_spmm.npz"), src_index=src_index.numpy(),
                        dst_index=dst_index.numpy(), weight=weight.numpy(), src=src3.numpy(), expected=exp.numpy())

    # --- test/ctest/test.cu via the reference's own generator + sequential goldens ----------------
    seq = ctypes.CDLL(oracle.REF_SEQ_PATH)
    rng, nnz, mn, mx, cv, Nf = 1000, 5000, 2, 7, 0.5, 32
    idx = np.zeros(nnz, dtype=np.int64)
    dst_len = seq.ref_generate_index(rng, mn, mx, nnz, ctypes.c_double(cv), idx.ctypes.data_as(ctypes.c_void_p))
    rs = np.random.RandomState(7)
    srcv = (rs.randint(0, 10, size=(nnz, Nf)) / 10.0).astype(np.float32)   # ramArray.cuh:77-81: rand()%10/10
    keys = int(idx[-1]) + 1
    dst = np.zeros((keys, Nf), dtype=np.float32)
    seq.ref_segment_coo_sequencial_f32(srcv.ctypes.data_as(ctypes.c_void_p), idx.ctypes.data_as(ctypes.c_void_p),
                                       nnz, Nf, keys, dst.ctypes.data_as(ctypes.c_void_p))
    col = rs.randint(0, keys, size=nnz).astype(np.int64)
    wv = rs.rand(nnz).astype(np.float32)
    feat = rs.rand(keys, Nf).astype(np.float32)
    dst_gws = np.zeros((keys, Nf), dtype=np.float32)
    seq.ref_gws_sequencial_f32(dst_gws.ctypes.data_as(ctypes.c_void_p), feat.ctypes.data_as(ctypes.c_void_p),
                               idx.ctypes.data_as(ctypes.c_void_p), col.ctypes.data_as(ctypes.c_void_p),
                               wv.ctypes.data_as(ctypes.c_void_p), nnz, Nf, keys)
    np.savez_compressed(os.path.join(HERE, "ctest_segreduce.npz"), index=idx, src=srcv, expected=dst,
                        dst_len=np.int64(dst_len), col=col, weight=wv, feat=feat, expected_gws=dst_gws)

    # --- segment counts through the unmodified reference CPU kernel ----------------------------------
    assert oracle.load_ref_extension()
    lens = torch.randint(0, 6, (400,), generator=g)          # zeros make gaps
    lens[-1] = 3
    index = torch.repeat_interleave(torch.arange(400), lens)
    counts = torch.ops.geot_ref.index_scatter(0, index, torch.ones(index.numel(), 1), "sum", True)
    np.savez_compressed(os.path.join(HERE, "ref_cpu_counts.npz"), index=index.numpy(),
                        counts=counts.numpy().astype(np.int64).reshape(-1))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
