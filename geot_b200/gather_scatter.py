"""``geot.gather_scatter`` (reference: ``geot/gather_scatter.py:3-39``).

Registered as ``geot::gather_scatter`` custom op with fake + autograd like the reference; the
backward is the same forward kernel on the transposed (src-sorted) edge list
(``geot/gather_scatter.py:26-37``).
"""
import torch


def gather_scatter_impl(src_index: torch.Tensor, dst_index: torch.Tensor, src: torch.Tensor) -> torch.Tensor:
    return torch.ops.geot.gather_scatter_impl(src_index, dst_index, src)


@torch.library.custom_op("geot::gather_scatter", mutates_args=())
def _gather_scatter_op(src_index: torch.Tensor, dst_index: torch.Tensor, src: torch.Tensor) -> torch.Tensor:
    return gather_scatter_impl(src_index, dst_index, src)


@torch.library.register_fake("geot::gather_scatter")
def _(src_index, dst_index, src):
    ctx = torch.library.get_ctx()
    dst_node = ctx.new_dynamic_size()
    return src.new_empty([dst_node, src.shape[1]])


def _setup_context(ctx, inputs, output):
    src_index, dst_index, src = inputs
    ctx.save_for_backward(src_index, dst_index)
    ctx.n_src = src.shape[0]


def _backward(ctx, grad):
    src_index, dst_index = ctx.saved_tensors
    grad = grad.contiguous()
    # transposed edge list: sorted by src (stable, keeps the dst order inside a src row), cached per graph
    from .transpose import transposed_edges
    t = transposed_edges(src_index, dst_index)
    g = gather_scatter_impl(t.src_index, t.dst_index, grad)
    if g.shape[0] < ctx.n_src:  # trailing src rows that no edge reads
        g = torch.cat([g, g.new_zeros(ctx.n_src - g.shape[0], g.shape[1])], 0)
    return None, None, g


torch.library.register_autograd("geot::gather_scatter", _backward, setup_context=_setup_context)


def gather_scatter(src_index: torch.Tensor, dst_index: torch.Tensor, src: torch.Tensor,
                   reduce: str = "sum") -> torch.Tensor:
    """``out[dst_index[e]] (reduce)= src[src_index[e]]``; ``dst_index`` sorted; rows = ``dst_index[-1]+1``.

    The trailing ``reduce`` is what the reference's tests / models / benchmarks pass
    (``test/test_gather_scatter.py:25``, ``models/conv/spmm.py:8``) although its wrapper at HEAD takes
    none; ``sum`` goes through the differentiable custom op, other reductions through
    ``geot::gather_scatter_reduce`` (forward only).
    """
    if reduce == "sum":
        return _gather_scatter_op(src_index, dst_index, src)
    return torch.ops.geot.gather_scatter_reduce(src_index, dst_index, src, reduce)
