"""Transposed (src-sorted) edge list for the backward of ``gather_scatter`` / ``gather_weight_scatter``.

The reference re-sorts the edges by ``src_index`` with ``torch.sort`` on every backward call
(``geot/gather_scatter.py:26-31``, ``geot/gather_weight_scatter.py:40-44``).  The graph does not change
between iterations, so the permutation and the two permuted index tensors are cached per
``(src_index, dst_index)`` pair (storage pointer, length and version counter; small LRU).  An entry
keeps the two index storages alive, so a cached pointer can never be re-used by another tensor.
"""
from collections import OrderedDict
from dataclasses import dataclass

import torch


@dataclass
class TransposedEdges:
    perm: torch.Tensor        # [E] edge ids in src-sorted order (stable: dst order kept inside a src row)
    dst_index: torch.Tensor   # [E] = src_index[perm], non-decreasing: the backward's segment index
    src_index: torch.Tensor   # [E] = dst_index[perm]: the backward's gather index
    keepalive: tuple = ()     # storages of the key tensors


_CACHE: "OrderedDict[tuple, TransposedEdges]" = OrderedDict()
_MAX = 8


def _key(t: torch.Tensor):
    return (t.untyped_storage().data_ptr(), t.storage_offset(), t.numel(), t._version, t.device)


def transposed_edges(src_index: torch.Tensor, dst_index: torch.Tensor) -> TransposedEdges:
    k = (_key(src_index), _key(dst_index))
    hit = _CACHE.get(k)
    if hit is not None:
        _CACHE.move_to_end(k)
        return hit
    sorted_src, perm = torch.sort(src_index, stable=True)
    t = TransposedEdges(perm, sorted_src.contiguous(), dst_index[perm].contiguous(),
                        (src_index.untyped_storage(), dst_index.untyped_storage()))
    _CACHE[k] = t
    while len(_CACHE) > _MAX:
        _CACHE.popitem(last=False)
    return t


def clear_transpose_cache() -> None:
    _CACHE.clear()
