"""``geot.csr_gws`` and ``geot.coo_to_csr`` (reference: ``geot/csr_gws.py:3-37``,
``geot/match_replace/format_transform.py:5-25``) -- the CSR entry point the FX rewriter targets.

``csr_gws(csrptr, csrind, weight, src)``: ``out[r] = sum_{e in [csrptr[r], csrptr[r+1])} weight[e] * src[csrind[e]]``.
Like the reference the output has ``csrptr.shape[0]`` rows (``csrc/csr_gws.cpp:29-31``: nrow + 1, the
last one zero).  The reference runs a dedicated row-caching kernel (``csr_gws_kernel.cuh:13-187``); here the
row index is expanded from ``csrptr`` once per graph (cached in the extension) and the same edge-balanced
segment-reduce kernel as ``gather_weight_scatter`` does the work.
"""
import torch


def csr_gws_impl(csrptr: torch.Tensor, csrind: torch.Tensor, weight: torch.Tensor, src: torch.Tensor) -> torch.Tensor:
    return torch.ops.geot.csr_gws_impl(csrptr, csrind, weight, src)


@torch.library.custom_op("geot::csr_gws", mutates_args=())
def csr_gws(csrptr: torch.Tensor, csrind: torch.Tensor, weight: torch.Tensor, src: torch.Tensor) -> torch.Tensor:
    return csr_gws_impl(csrptr, csrind, weight, src)


@torch.library.register_fake("geot::csr_gws")
def _(csrptr, csrind, weight, src):
    return src.new_empty([csrptr.shape[0], src.shape[1]])


@torch.library.custom_op("geot::coo_to_csr", mutates_args=())
def coo_to_csr(coo_row: torch.Tensor) -> torch.Tensor:
    """int32 CSR row pointer ``[coo_row.max() + 2]`` of a sorted COO row index."""
    return torch.ops.geot.coo_to_csr_impl(coo_row)


@torch.library.register_fake("geot::coo_to_csr")
def _(coo_row):
    ctx = torch.library.get_ctx()
    return coo_row.new_empty([ctx.new_dynamic_size()], dtype=torch.int32)
