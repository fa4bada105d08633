"""Multi-GPU sharding of the segment-reduction path (new functionality -- the reference
has no distributed code).

One process per GPU (``torch.distributed``, NCCL over NVLink / NVSwitch).  The dst-sorted edge list is
cut into ``world_size`` contiguous dst-row ranges with balanced edge counts; rank g owns rows
``[row_bounds[g], row_bounds[g+1])`` of every node-feature matrix and the edges that point into them,
reduces its own dst slice, and writes only that slice -- there is no reduction collective.  The only
exchange is the all-gather of the src feature rows that gather ops read across shard boundaries.
``index_scatter`` needs no communication at all (edge-aligned data is sharded with the edges).
"""
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.distributed as dist


def shard_bounds_from_rowptr(rowptr: torch.Tensor, parts: int):
    """Edge-balanced row/edge bounds from a CSR rowptr (pure torch; same rule as the device kernel
    behind ``torch.ops.geot.plan_shards``: shard g starts at the segment boundary nearest to
    ``g*E/parts``).  Returns two python lists of ``parts+1`` ints."""
    S = rowptr.numel() - 1
    E = int(rowptr[-1])
    rows, edges = [0], [0]
    for g in range(1, parts):
        target = (E // parts) * g + ((E % parts) * g) // parts
        r = int(torch.searchsorted(rowptr, torch.tensor([target], dtype=rowptr.dtype, device=rowptr.device),
                                   right=False)[0])
        r = min(r, S)
        if r > 0 and target - int(rowptr[r - 1]) < int(rowptr[r]) - target:
            r -= 1
        rows.append(r)
        edges.append(int(rowptr[r]))
    rows.append(S)
    edges.append(E)
    return rows, edges


@dataclass
class GraphShard:
    """Rank-local view of a dst-sorted edge list."""
    rank: int
    world_size: int
    row_bounds: List[int]            # parts+1, dst/src row ownership
    edge_bounds: List[int]           # parts+1
    src_index: Optional[torch.Tensor]  # [E_local] global src row ids (None for index_scatter)
    dst_index: torch.Tensor          # [E_local] LOCAL dst row ids (global - row_bounds[rank])
    weight: Optional[torch.Tensor]   # [E_local] / [E_local, H]

    @property
    def num_local_rows(self) -> int:
        return self.row_bounds[self.rank + 1] - self.row_bounds[self.rank]

    @property
    def num_local_edges(self) -> int:
        return self.edge_bounds[self.rank + 1] - self.edge_bounds[self.rank]

    @property
    def imbalance(self) -> float:
        """max_g |E_g| / (E / G)."""
        e = [self.edge_bounds[g + 1] - self.edge_bounds[g] for g in range(self.world_size)]
        return max(e) / (self.edge_bounds[-1] / self.world_size)


def shard_graph(src_index: Optional[torch.Tensor], dst_index: torch.Tensor, weight: Optional[torch.Tensor],
                rank: int, world_size: int, row_bounds=None, edge_bounds=None) -> GraphShard:
    """Cut the (replicated) global edge list down to this rank's shard."""
    if row_bounds is None:
        if dst_index.is_cuda:
            b = torch.ops.geot.plan_shards(dst_index, world_size)
            row_bounds, edge_bounds = b[0].tolist(), b[1].tolist()
        else:
            S = int(dst_index[-1]) + 1
            deg = torch.bincount(dst_index, minlength=S)
            rowptr = torch.cat([deg.new_zeros(1), deg.cumsum(0)])
            row_bounds, edge_bounds = shard_bounds_from_rowptr(rowptr, world_size)
    e0, e1 = edge_bounds[rank], edge_bounds[rank + 1]
    local_dst = (dst_index[e0:e1] - row_bounds[rank]).contiguous()
    local_src = src_index[e0:e1].contiguous() if src_index is not None else None
    local_w = weight[e0:e1].contiguous() if weight is not None else None
    return GraphShard(rank, world_size, list(row_bounds), list(edge_bounds), local_src, local_dst, local_w)


def all_gather_rows(x_local: torch.Tensor, row_bounds: List[int], group=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """All-gather ragged row shards into the full ``[N, ...]`` matrix (NCCL over NVLink on GPUs).

    Edge-balanced shards own different numbers of rows.  On NCCL every rank's rows land directly in their
    place in ``out`` (``all_gather`` with uneven outputs = one grouped set of broadcasts, no padding and no
    compaction pass); equal shards take the single ``all_gather_into_tensor`` path.  Gloo (CPU tests) has no
    uneven all-gather: shards are padded to the largest one and compacted."""
    world = dist.get_world_size(group)
    sizes = [row_bounds[g + 1] - row_bounds[g] for g in range(world)]
    tail = list(x_local.shape[1:])
    x_local = x_local.contiguous()
    full = out if out is not None else x_local.new_empty([row_bounds[-1]] + tail)
    if all(s == sizes[0] for s in sizes):
        dist.all_gather_into_tensor(full, x_local, group=group)
        return full
    if x_local.is_cuda:
        dist.all_gather([full[row_bounds[g]:row_bounds[g + 1]] for g in range(world)], x_local, group=group)
        return full
    n_max = max(sizes)
    pad = x_local.new_zeros([n_max] + tail)
    pad[: x_local.shape[0]] = x_local
    buf = x_local.new_empty([n_max * world] + tail)
    dist.all_gather_into_tensor(buf, pad, group=group)
    for g in range(world):
        full[row_bounds[g]:row_bounds[g + 1]] = buf[g * n_max: g * n_max + sizes[g]]
    return full


def _pad_rows(out: torch.Tensor, n: int) -> torch.Tensor:
    if out.shape[0] == n:
        return out
    return torch.cat([out, out.new_zeros([n - out.shape[0]] + list(out.shape[1:]))], 0)


def sharded_gather_scatter(shard: GraphShard, x_local: torch.Tensor, reduce: str = "sum", group=None,
                           x_full: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Rank-local rows of ``gather_(weight_)scatter`` on the sharded graph.

    ``x_local``: this rank's rows of the src matrix.  ``x_full`` (optional) skips the all-gather when
    the caller already holds the replicated matrix ("src pre-replicated" measurements)."""
    from .gather_scatter import gather_scatter
    from .gather_weight_scatter import gather_weight_scatter
    if x_full is None:
        x_full = all_gather_rows(x_local, shard.row_bounds, group)
    if shard.num_local_edges == 0:
        return x_full.new_zeros([shard.num_local_rows] + list(x_full.shape[1:]))
    if shard.weight is None:
        out = gather_scatter(shard.src_index, shard.dst_index, x_full, reduce)
    else:
        out = gather_weight_scatter(shard.src_index, shard.dst_index, shard.weight, x_full, reduce)
    return _pad_rows(out, shard.num_local_rows)


def sharded_index_scatter(shard: GraphShard, src_local_edges: torch.Tensor, reduce: str = "sum") -> torch.Tensor:
    """``index_scatter`` on this rank's edge slice; no communication."""
    from .index_scatter import index_scatter
    out = index_scatter(0, src_local_edges, shard.dst_index, reduce, True)
    return _pad_rows(out, shard.num_local_rows)
