"""``geot.gather_weight_scatter`` (reference: ``geot/gather_weight_scatter.py:4-51``)."""
import torch


def gather_weight_scatter_impl(src_index: torch.Tensor, dst_index: torch.Tensor, weight: torch.Tensor,
                               src: torch.Tensor) -> torch.Tensor:
    return torch.ops.geot.gather_weight_scatter_impl(src_index, dst_index, weight, src)


@torch.library.custom_op("geot::gather_weight_scatter", mutates_args=())
def _gather_weight_scatter_op(src_index: torch.Tensor, dst_index: torch.Tensor, weight: torch.Tensor,
                              src: torch.Tensor) -> torch.Tensor:
    return gather_weight_scatter_impl(src_index, dst_index, weight, src)


@torch.library.register_fake("geot::gather_weight_scatter")
def _(src_index, dst_index, weight, src):
    ctx = torch.library.get_ctx()
    dst_node = ctx.new_dynamic_size()
    return src.new_empty([dst_node, src.shape[1]])


def _setup_context(ctx, inputs, output):
    src_index, dst_index, weight, src = inputs
    ctx.save_for_backward(src_index, dst_index, weight, src)


def sddmm_coo_impl(src_index: torch.Tensor, dst_index: torch.Tensor, mat_1: torch.Tensor,
                   mat_2: torch.Tensor) -> torch.Tensor:
    """``out[e] = <mat_1[dst_index[e]], mat_2[src_index[e]]>`` (reference: ``geot/gather_weight_scatter.py:8-12``,
    kernel ``csrc/cuda/sddmm_coo_kernel.cuh``; there fp32 + int32 only)."""
    return torch.ops.geot.sddmm_coo_impl(src_index, dst_index, mat_1, mat_2)


def _backward(ctx, grad):
    from .transpose import transposed_edges
    src_index, dst_index, weight, src = ctx.saved_tensors
    grad = grad.contiguous()
    src_grad = weight_grad = None
    if ctx.needs_input_grad[3]:
        # the same forward kernel on the transposed (src-sorted) edge list, as the reference does
        # (geot/gather_weight_scatter.py:40-46); the sort permutation is cached per graph instead of a
        # torch.sort per call
        t = transposed_edges(src_index, dst_index)
        g = gather_weight_scatter_impl(t.src_index, t.dst_index, weight[t.perm], grad)
        if g.shape[0] < src.shape[0]:      # trailing src rows that no edge reads
            g = torch.cat([g, g.new_zeros(src.shape[0] - g.shape[0], g.shape[1])], 0)
        src_grad = g
    if ctx.needs_input_grad[2]:
        # weight_grad[e] = <grad[dst[e]], src[src[e]]> in the ORIGINAL edge order (dst-sorted, so the kernel
        # keeps the grad row in registers across a segment).  The reference passes the src-sorted lists and
        # returns the gradient permuted (geot/gather_weight_scatter.py:47) -- not reproduced.
        weight_grad = sddmm_coo_impl(src_index, dst_index, grad, src)
    return None, None, weight_grad, src_grad


torch.library.register_autograd("geot::gather_weight_scatter", _backward, setup_context=_setup_context)


def gather_weight_scatter(src_index: torch.Tensor, dst_index: torch.Tensor, weight: torch.Tensor,
                          src: torch.Tensor, reduce: str = "sum") -> torch.Tensor:
    """``out[dst_index[e]] (reduce)= weight[e] * src[src_index[e]]`` -- the GCN aggregation.

    ``dst_index`` sorted; rows = ``dst_index[-1]+1``.  Optional trailing ``reduce`` as the reference's
    callers pass it (``test/test_gather_weight_scatter.py:24``, ``models/conv/spmm.py:14``).
    """
    if reduce == "sum":
        return _gather_weight_scatter_op(src_index, dst_index, weight, src)
    return torch.ops.geot.gather_weight_scatter_reduce(src_index, dst_index, weight, src, reduce)
