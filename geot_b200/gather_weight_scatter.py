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


def _backward(ctx, grad):
    src_index, dst_index, weight, src = ctx.saved_tensors
    grad = grad.contiguous()
    src_grad = weight_grad = None
    if ctx.needs_input_grad[3]:
        _, perm = torch.sort(src_index, stable=True)
        g = gather_weight_scatter_impl(dst_index[perm], src_index[perm], weight[perm], grad)
        if g.shape[0] < src.shape[0]:
            g = torch.cat([g, g.new_zeros(src.shape[0] - g.shape[0], g.shape[1])], 0)
        src_grad = g
    if ctx.needs_input_grad[2]:
        # weight_grad[e] = <grad[dst[e]], src[src[e]]>  (the reference's sddmm_coo,
        # geot/gather_weight_scatter.py:47); SDDMM is outside this round's hot path,
        # so it is expressed with torch ops here.
        weight_grad = (grad.index_select(0, dst_index) * src.index_select(0, src_index)).sum(-1)
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
