"""``geot.index_scatter`` (reference: ``geot/index_scatter.py:5-8``, op ``csrc/index_scatter.cpp:43-47``)."""
import torch


def _is_index(t: torch.Tensor) -> bool:
    return t.dtype in (torch.int64, torch.int32) and t.dim() == 1


def index_scatter(dim: int, src: torch.Tensor, index: torch.Tensor, reduce: str = "sum",
                  sorted: bool = True) -> torch.Tensor:
    """``out[index[i], ...] (reduce)= src[i, ...]`` along ``dim``; ``out.shape[dim] = index[-1] + 1``.

    The reference wrapper takes ``(dim, src, index)`` (``geot/index_scatter.py:5``) while its README
    (``README.md:17``) and op schema take ``(dim, index, src)``; both orders are accepted -- the 1-D
    integer tensor is the index.  ``index`` must be non-decreasing when ``sorted=True``.
    Unlike the reference CUDA path, ``reduce`` and ``dim`` are honoured
    (sum | mean | max/amax | min/amin | prod).
    """
    if _is_index(src) and not _is_index(index):
        src, index = index, src
    if index.dtype != torch.int64:
        index = index.long()
    return torch.ops.geot.index_scatter(dim, index, src, reduce, sorted)
