"""``geot.index_scatter`` (reference: ``geot/index_scatter.py:5-8``, op ``csrc/index_scatter.cpp:43-47``)."""
import weakref

import torch

# int32 indices are widened once per index tensor, not once per call: the operator caches its plan on the int64
# tensor it is given, and a fresh ``index.long()`` every call would miss that cache every time (a host sync + a
# preprocessing pass per call).  Keyed on the int32 tensor's identity and version; entries die with the tensor.
_widened = {}


def _as_int64(index: torch.Tensor) -> torch.Tensor:
    key = (id(index), index._version, index.data_ptr(), index.numel())
    hit = _widened.get(key)
    if hit is not None and hit[0]() is index:
        return hit[1]
    if len(_widened) >= 8:
        _widened.clear()
    wide = index.long()
    _widened[key] = (weakref.ref(index), wide)
    return wide


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
        index = _as_int64(index)
    return torch.ops.geot.index_scatter(dim, index, src, reduce, sorted)
