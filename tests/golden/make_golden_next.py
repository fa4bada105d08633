"""Golden vectors for the "next" rows of the hot-path table: csr_gws, coo_to_csr, sddmm_coo.

Run HERE (the build container):   python tests/golden/make_golden_next.py
Expected outputs come from the reference's own test formulas evaluated with torch on CPU, fixed seed:
  * ref_test_csr_gws   test/test_csr_gws.py:6-47 -- `coo_to_csr` helper (bincount -> cumsum -> int32) and
                       `ref_spmm` (torch.sparse.mm of the coalesced COO matrix);
  * ref_test_sddmm     the definition the reference kernel implements (csrc/cuda/sddmm_coo_kernel.cuh:44-71 with
                       the launcher's binding row = dst_index, col = src_index,
                       csrc/cuda/gather_weight_scatter_cuda.cu:46-50), sizes of test/test_sddmm.py scaled down;
                       the reference's own test only times the op, so the GPU suite additionally compares with the
                       reference's compiled kernel (tests/test_gpu_next.py::test_against_reference_cuda_kernels_next).
"""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    g = torch.Generator().manual_seed(20240404)
    # --- test/test_csr_gws.py ------------------------------------------------------------------------
    nrow, nnz, F = 100, 1000, 32
    src_index = torch.randint(0, nrow, (nnz,), generator=g)
    dst_index = torch.randint(0, nrow, (nnz,), generator=g)
    weight = torch.rand(nnz, generator=g)
    dst_sorted, indices = torch.sort(dst_index)
    src_sorted, weight_sorted = src_index[indices], weight[indices]
    csr_row = torch.zeros(nrow + 1)
    csr_row[1:] = torch.cumsum(torch.bincount(dst_sorted, minlength=nrow), 0)
    csr_row = csr_row.int()
    src = torch.rand(nrow, F, generator=g)
    adj = torch.sparse_coo_tensor(torch.stack([dst_sorted, src_sorted]), weight_sorted, (nrow, nrow)).coalesce()
    ref = torch.sparse.mm(adj, src)
    np.savez_compressed(os.path.join(HERE, "ref_test_csr_gws.npz"), rowptr=csr_row.numpy(), colidx=src_sorted.numpy(),
                        val=weight_sorted.numpy(), src=src.numpy(), dst_sorted=dst_sorted.numpy(), expected=ref.numpy())
    # --- sddmm_coo -------------------------------------------------------------------------------------
    nodes, edges, F = 500, 6000, 128
    x1 = torch.rand(nodes, F, generator=g)
    x2 = torch.rand(nodes, F, generator=g)
    si = torch.randint(0, nodes, (edges,), generator=g)
    di = torch.randint(0, nodes, (edges,), generator=g)
    exp = (x1.double().index_select(0, di) * x2.double().index_select(0, si)).sum(-1).float()
    np.savez_compressed(os.path.join(HERE, "ref_test_sddmm.npz"), src_index=si.numpy(), dst_index=di.numpy(),
                        mat_1=x1.numpy(), mat_2=x2.numpy(), expected=exp.numpy())
    print("written")


if __name__ == "__main__":
    main()
