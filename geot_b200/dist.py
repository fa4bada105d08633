"""Multi-GPU sharding of the segment-reduction path (new functionality -- the reference
has no distributed code).

One process per GPU (``torch.distributed``, NCCL over NVLink / NVSwitch).  The dst-sorted edge list is
cut into ``world_size`` contiguous dst-row ranges with balanced edge counts; rank g owns rows
``[row_bounds[g], row_bounds[g+1])`` of every node-feature matrix and the edges that point into them,
reduces its own dst slice, and writes only that slice -- there is no reduction collective.  The only
exchange is the all-gather of the src feature rows that gather ops read across shard boundaries.
``index_scatter`` needs no communication at all (edge-aligned data is sharded with the edges).

Two forms of the exchange:

* ``all_gather_rows`` + one reduction (``sharded_gather_scatter``): the whole matrix lands, then the kernel runs.
* ``PipelinedGather``: the all-gather is unrolled into ``world-1`` staggered NCCL send/recv steps (step k: send my
  rows to rank r-k, receive the rows of rank r+k -- every rank sends and receives exactly one shard per step, so
  each step runs at full NVSwitch bandwidth) on a side stream, and the rank's edges are regrouped by the rank that
  owns their src row: the bucket of the rank's own rows is reduced while shard r+1 is in flight, bucket r+k as soon
  as step k has landed.  Buckets write partial results that one combine kernel adds in bucket order (fixed order:
  bit-reproducible), so the exchange costs what exceeds the reduction time instead of adding to it.
* ``PipelinedGather(needed_only=True)``: the same schedule, but step k carries only the rows the receiver's edges
  actually reference (SURVEY 8e "exchanging only the rows actually referenced").  The request lists are exchanged
  once per graph; every call packs the requested rows per peer (one row-gather kernel) and the receiver's bucket k
  reads them from a compact buffer through remapped src ids.  Pays off when a shard references a fraction of a
  peer's rows (sparse, products-like graphs); on dense graphs (Reddit: every shard references almost every row)
  it degenerates to the full exchange plus the pack.
* ``PeerPushGather``: the needed-rows exchange without NCCL on the data path.  Every rank's receive buffer is a
  symmetric-memory allocation mapped into all processes of the box; ONE kernel per call packs the requested rows
  and stores each straight into its slot of the requester's buffer over NVLink (``geot_b200_push_rows``), between
  two cross-GPU barriers on a side stream, while the main stream reduces the edges whose src rows are local; the
  remote edges are reduced when the barrier has passed, and a two-way combine finishes.  Two buckets instead of
  ``world``: 2 reductions + 1 push + 1 combine per call whatever the GPU count.
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


# ---- exchange overlapped with the reduction -------------------------------------------------------

@dataclass
class SrcBuckets:
    """A rank's edges regrouped by the owner of their src row, in processing order: bucket k holds the edges whose
    src row belongs to rank ``(rank + k) % world`` (bucket 0 = the rank's own rows).  The regrouping is stable, so
    every bucket is still sorted by dst."""
    perm: torch.Tensor               # [E_local] position in the shard's edge list of every regrouped edge
    bounds: List[int]                # world+1 offsets of the buckets in the regrouped arrays
    src_index: torch.Tensor          # [E_local] regrouped
    dst_index: torch.Tensor          # [E_local] regrouped (local dst rows)


def bucket_by_src_owner(shard: GraphShard) -> SrcBuckets:
    world, rank = shard.world_size, shard.rank
    cuts = torch.tensor(shard.row_bounds[1:-1], dtype=shard.src_index.dtype, device=shard.src_index.device)
    owner = torch.bucketize(shard.src_index, cuts, right=True)          # rb[g] <= s < rb[g+1]  <=>  owner == g
    key = (owner - rank) % world
    perm = torch.argsort(key, stable=True)
    counts = torch.bincount(key, minlength=world).tolist()
    bounds = [0]
    for c in counts:
        bounds.append(bounds[-1] + int(c))
    return SrcBuckets(perm, bounds, shard.src_index[perm].contiguous(), shard.dst_index[perm].contiguous())


@dataclass
class NeededRows:
    """Per-graph state of the needed-rows exchange (``PipelinedGather(needed_only=True)``)."""
    recv_counts: List[int]           # [world] rows this rank receives from each owner (0 for itself)
    recv_offsets: List[int]          # [world+1] their offsets in the compact receive buffer, in STEP order (rank+1, rank+2, ...)
    send_counts: List[int]           # [world] rows each peer asked this rank for
    send_offsets: List[int]          # [world+1] their offsets in the packed send buffer, in STEP order (rank-1, rank-2, ...)
    send_rows: torch.Tensor          # [sum(send_counts)] LOCAL row ids to pack, grouped by peer in step order
    src_index: torch.Tensor          # [E_local] bucket-ordered src ids: bucket 0 -> local row ids, bucket k -> row ids
                                     #           in the compact receive buffer


def build_needed_rows(shard: GraphShard, buckets: SrcBuckets, group=None) -> NeededRows:
    """One-time request exchange: for every peer the sorted unique rows this rank's edges reference there.

    Collectives: one all-gather of the [world] request counts, then ``world-1`` staggered send/recv steps of the
    request lists (the same schedule the row exchange uses: works on NCCL and gloo alike)."""
    world, rank, rb = shard.world_size, shard.rank, shard.row_bounds
    dev, idt = buckets.src_index.device, buckets.src_index.dtype
    b = buckets.bounds
    src = torch.empty_like(buckets.src_index)
    src[b[0]:b[1]] = buckets.src_index[b[0]:b[1]] - rb[rank]
    need, recv_counts, recv_offsets = {}, [0] * world, [0]
    for k in range(1, world):                       # step order: owner rank+k
        g = (rank + k) % world
        s_k = buckets.src_index[b[k]:b[k + 1]]
        uniq = torch.unique(s_k)                    # sorted
        need[g] = (uniq - rb[g]).contiguous()       # as the owner's local row ids
        recv_counts[g] = int(uniq.numel())
        src[b[k]:b[k + 1]] = recv_offsets[-1] + torch.searchsorted(uniq, s_k)
        recv_offsets.append(recv_offsets[-1] + recv_counts[g])
    recv_offsets.append(recv_offsets[-1])           # world+1 entries (last step has no successor)
    # counts[r][g] = rows rank r needs from rank g
    mine = torch.tensor(recv_counts, dtype=torch.int64, device=dev)
    counts = torch.empty(world * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, mine, group=group)
    counts = counts.view(world, world).cpu()
    send_counts = [int(counts[g][rank]) for g in range(world)]
    send_counts[rank] = 0
    send_offsets, lists = [0], []
    for k in range(1, world):                       # step order: I serve rank-k in step k
        to, frm = (rank - k) % world, (rank + k) % world
        lst = torch.empty(send_counts[to], dtype=idt, device=dev)
        ops = []
        if recv_counts[frm]:
            ops.append(dist.P2POp(dist.isend, need[frm], frm, group))
        if send_counts[to]:
            ops.append(dist.P2POp(dist.irecv, lst, to, group))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        lists.append(lst)
        send_offsets.append(send_offsets[-1] + send_counts[to])
    send_offsets.append(send_offsets[-1])
    send_rows = torch.cat(lists) if lists else torch.empty(0, dtype=idt, device=dev)
    n_local = rb[rank + 1] - rb[rank]
    assert send_rows.numel() == 0 or (int(send_rows.min()) >= 0 and int(send_rows.max()) < n_local)
    return NeededRows(recv_counts, recv_offsets, send_counts, send_offsets, send_rows.contiguous(), src.contiguous())


class PipelinedGather:
    """``gather_(weight_)scatter`` on a dst-row shard with the src-row exchange overlapped (sum / mean).

    Rows may be ``[N, F]`` (weights ``[E]`` or none) or ``[N, H, F]`` with per-head weights ``[E, H]`` (``mh_spmm``).
    ``x_full`` is the caller's [N, ...] replica buffer whose OWN row range already holds this rank's rows (the
    producer writes them there; ``local_rows(x_full)`` is that view).  Every call exchanges the other ranks' rows
    into it while reducing.  ``reducer`` / ``combiner`` / ``permuter`` default to the C-ABI kernels; the gloo tests
    inject CPU stand-ins to check the host logic."""

    def __init__(self, shard: GraphShard, group=None, reducer=None, combiner=None, permuter=None,
                 needed_only: bool = False):
        self.shard, self.group = shard, group
        self.world, self.rank = shard.world_size, shard.rank
        self.buckets = bucket_by_src_owner(shard)
        self.cuda = shard.dst_index.is_cuda
        self._reducer, self._combiner, self._permuter = reducer, combiner, permuter
        self._plans, self._ws, self._parts, self._wperm = {}, None, None, None
        self._rowptr = None
        # needed_only: step k carries only the rows this rank's bucket k references (see the module docstring)
        self.needed = build_needed_rows(shard, self.buckets, group) if needed_only else None
        self._send_buf, self._recv_buf = None, None
        self.comm_stream = torch.cuda.Stream() if self.cuda else None
        self.events = [torch.cuda.Event() for _ in range(self.world)] if self.cuda else None

    def local_rows(self, x_full: torch.Tensor) -> torch.Tensor:
        rb = self.shard.row_bounds
        return x_full[rb[self.rank]:rb[self.rank + 1]]

    # -- default device kernels (C ABI) ---------------------------------------------------------------
    def _reduce_bucket(self, k, x_full, w_b, out):
        b = self.buckets
        e0, e1 = b.bounds[k], b.bounds[k + 1]
        S = self.shard.num_local_rows
        if e1 == e0 or S == 0:
            out.zero_()
            return
        si, di = b.src_index[e0:e1], b.dst_index[e0:e1]
        if self.needed is not None:
            si = self.needed.src_index[e0:e1]      # ids into x_full = the rank's own rows (k == 0) / the compact buffer
        if self._reducer is not None:
            out.copy_(self._reducer(x_full, si, di, w_b, S))
            return
        from . import abi
        if k not in self._plans:
            self._plans[k] = abi.DevicePlan(di, S)
        if self._ws is None:
            # one scratch buffer for all buckets: the partition (hence the scratch size) is not monotonic in E
            W = x_full[0].numel()
            sizes = [self.buckets.bounds[i + 1] - self.buckets.bounds[i] for i in range(len(self.buckets.bounds) - 1)]
            need = lambda n: abi.lib().geot_b200_workspace_bytes(n, W, abi.DTYPE[x_full.dtype], 1)
            self._ws = abi.Workspace(max(sizes, key=need), W, x_full.dtype, x_full.device)
        H = x_full.shape[1] if x_full.dim() == 3 else 1        # [N, H, F] rows with [E, H] weights: mh_spmm
        abi.segment_reduce(x_full, si, di, w_b, "sum", S=S, H=H, plan=self._plans[k], out=out, workspace=self._ws)

    def _combine(self, parts, out, reduce):
        if self._combiner is not None:
            out.copy_(self._combiner(parts, reduce, self.shard.dst_index, self.shard.num_local_rows))
            return
        from . import abi
        rowptr = None
        if reduce == "mean":
            if self._rowptr is None:
                self._rowptr = abi.DevicePlan(self.shard.dst_index, self.shard.num_local_rows).rowptr.clone()
            rowptr = self._rowptr
        abi.combine_partials(parts, out, reduce, rowptr)

    def _permute(self, w):
        if self._permuter is not None:
            return self._permuter(w, self.buckets.perm)
        from . import abi
        if self._wperm is None or self._wperm.shape != w.shape or self._wperm.dtype != w.dtype:
            self._wperm = torch.empty_like(w)
        return abi.permute_edges(w, self.buckets.perm, self._wperm)

    # -- the op ---------------------------------------------------------------------------------------
    def __call__(self, x_full: torch.Tensor, weight: Optional[torch.Tensor] = None, reduce: str = "sum",
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
        assert reduce in ("sum", "mean"), "the pipelined exchange combines partial sums: sum / mean only"
        world, rank, rb = self.world, self.rank, self.shard.row_bounds
        S = self.shard.num_local_rows
        tail = list(x_full.shape[1:])
        if out is None:
            out = x_full.new_empty([S] + tail)
        if self._parts is None or list(self._parts.shape[2:]) != tail or self._parts.dtype != x_full.dtype:
            self._parts = x_full.new_empty([world, S] + tail)
        parts = self._parts
        nd = self.needed
        n_local = rb[rank + 1] - rb[rank]
        # needed_only accepts the rank's own rows [n_local, ...] as well as the [N, ...] buffer (only its own range is read)
        x_mine = x_full if (nd is not None and x_full.shape[0] == n_local) else self.local_rows(x_full)
        if nd is not None:
            n_send, n_recv = nd.send_offsets[-1], nd.recv_offsets[-1]
            if self._send_buf is None or list(self._send_buf.shape[1:]) != tail or self._send_buf.dtype != x_full.dtype:
                self._send_buf = x_full.new_empty([max(n_send, 1)] + tail)
                self._recv_buf = x_full.new_empty([max(n_recv, 1)] + tail)

        # exchange: world-1 staggered send/recv steps on the side stream
        if self.cuda:
            self.comm_stream.wait_stream(torch.cuda.current_stream())   # my rows are final; x_full's old rows were consumed
        if nd is not None and n_send:
            if self.cuda:
                with torch.cuda.stream(self.comm_stream):
                    self._pack(x_mine, nd.send_rows, self._send_buf)
            else:
                self._pack(x_mine, nd.send_rows, self._send_buf)
        for k in range(1, world):
            to, frm = (rank - k) % world, (rank + k) % world
            ops = []
            if nd is None:
                if x_mine.shape[0]:
                    ops.append(dist.P2POp(dist.isend, x_mine, to, self.group))
                if rb[frm + 1] > rb[frm]:
                    ops.append(dist.P2POp(dist.irecv, x_full[rb[frm]:rb[frm + 1]], frm, self.group))
                # (a rank with an empty row range neither sends nor is received from: both sides skip consistently)
            else:
                if nd.send_counts[to]:
                    ops.append(dist.P2POp(dist.isend, self._send_buf[nd.send_offsets[k - 1]:nd.send_offsets[k]], to, self.group))
                if nd.recv_counts[frm]:
                    ops.append(dist.P2POp(dist.irecv, self._recv_buf[nd.recv_offsets[k - 1]:nd.recv_offsets[k]], frm, self.group))
                # (send_counts[to] here == recv_counts[me] on rank `to`: both sides skip an empty step consistently)
            if self.cuda:
                with torch.cuda.stream(self.comm_stream):
                    if ops:
                        for r in dist.batch_isend_irecv(ops):
                            r.wait()
                    self.events[k].record(self.comm_stream)
            elif ops:
                for r in dist.batch_isend_irecv(ops):
                    r.wait()

        # reduction: own bucket first, bucket k once step k has landed
        w_all = self._permute(weight) if weight is not None else None
        b = self.buckets.bounds
        for k in range(world):
            if k > 0 and self.cuda:
                torch.cuda.current_stream().wait_event(self.events[k])
            w_b = w_all[b[k]:b[k + 1]] if w_all is not None else None
            x_k = x_full if nd is None else (x_mine if k == 0 else self._recv_buf)
            self._reduce_bucket(k, x_k, w_b, parts[k])
        self._combine(parts, out, reduce)
        return out

    def aggregate(self, x_local: torch.Tensor, weight: Optional[torch.Tensor] = None, reduce: str = "sum") -> torch.Tensor:
        """The op on this rank's OWN rows ``x_local`` (what a layer produces): the full-exchange form keeps its
        ``[N, ...]`` replica buffer here (one per row shape / dtype) and copies the rows into their range; the
        needed-rows form reads them in place.  Returns a new ``[n_local_rows, ...]`` tensor."""
        if self.needed is not None:
            return self(x_local, weight, reduce)
        tail = list(x_local.shape[1:])
        buf = getattr(self, "_replica", None)
        if buf is None or list(buf.shape[1:]) != tail or buf.dtype != x_local.dtype or buf.device != x_local.device:
            buf = self._replica = x_local.new_empty([self.shard.row_bounds[-1]] + tail)
        self.local_rows(buf).copy_(x_local)
        return self(buf, weight, reduce)

    def _pack(self, x_mine, rows, out):
        """out[i] = x_mine[rows[i]]: the rows the peers asked for, grouped by peer in step order."""
        if self._permuter is not None:
            out[: rows.numel()].copy_(self._permuter(x_mine, rows))
            return
        from . import abi
        abi.permute_edges(x_mine, rows, out)

    def exchanged_rows(self):
        """(rows received per call, rows a full exchange would receive) -- the saving of needed_only."""
        rb = self.shard.row_bounds
        full = rb[-1] - (rb[self.rank + 1] - rb[self.rank])
        return (self.needed.recv_offsets[-1] if self.needed is not None else full), full


@dataclass
class _PeerBuffer:
    """A symmetric-memory receive buffer as the default (CUDA) path of PeerPushGather holds it."""
    handle: object                   # torch _SymmetricMemory: barrier(channel)
    bases: torch.Tensor              # [world] int64 on the device: every peer's mapped base address of the buffer


class PeerPushGather(PipelinedGather):
    """``gather_(weight_)scatter`` on a dst-row shard with the needed src rows PUSHED over peer memory (sum / mean).

    Per graph: the request lists of the needed-rows exchange (``build_needed_rows``), the slot of every requested row
    in its requester's receive buffer (``dest_peer`` / ``dest_row``), and the rank's edges split stably into two
    dst-sorted buckets -- src row local / src row remote -- with src ids that point into ``x_local`` / the receive
    buffer.  Per call (side stream): barrier (every peer has finished reading its buffer), one push kernel, barrier
    (every peer's rows have landed here); (main stream): local bucket, wait, remote bucket, combine.

    ``allocator(shape, dtype, device) -> (buffer, handle)``, ``pusher(x_mine, rows, dest_peer, dest_row, buffer, handle)``
    and ``barrier(handle, channel)`` default to torch symmetric memory + the C-ABI kernel; the gloo tests inject CPU
    stand-ins to check the host logic (slots, buckets, ordering)."""

    def __init__(self, shard: GraphShard, group=None, reducer=None, combiner=None, permuter=None, allocator=None,
                 pusher=None, barrier=None):
        self.shard, self.group = shard, group
        self.world, self.rank = world, rank = shard.world_size, shard.rank
        self.cuda = shard.dst_index.is_cuda
        self._reducer, self._combiner, self._permuter = reducer, combiner, permuter
        self._allocator, self._pusher, self._barrier_fn = allocator, pusher, barrier
        self._plans, self._ws, self._parts, self._wperm = {}, None, None, None
        self._rowptr = None
        self.needed = None                                  # (the base class's NCCL needed-rows state: not used here)
        self._bufs = {}
        dev = shard.dst_index.device
        per_owner = bucket_by_src_owner(shard)
        self.requests = nd = build_needed_rows(shard, per_owner, group)
        E = shard.dst_index.numel()
        # src ids in shard edge order, then the stable two-way split (local first): both halves stay dst-sorted
        compact = torch.empty_like(nd.src_index)
        compact[per_owner.perm] = nd.src_index
        remote = torch.ones(E, dtype=torch.int8, device=dev)
        remote[per_owner.perm[: per_owner.bounds[1]]] = 0
        perm2 = torch.argsort(remote, stable=True)
        self.buckets = SrcBuckets(perm2, [0, per_owner.bounds[1], E], compact[perm2].contiguous(),
                                  shard.dst_index[perm2].contiguous())
        # slots: my send segment k (for rank - k) lands at that rank's receive offset of ITS step k (owner = me)
        mine = torch.tensor(nd.recv_offsets, dtype=torch.int64, device=dev)
        table = torch.empty(world * (world + 1), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(table, mine, group=group)
        table = table.view(world, world + 1).cpu()
        peers, slots = [], []
        for k in range(1, world):
            to = (rank - k) % world
            n = nd.send_counts[to]
            peers.append(torch.full((n,), to, dtype=torch.int32))
            slots.append(int(table[to][k - 1]) + torch.arange(n, dtype=torch.int64))
        self.dest_peer = (torch.cat(peers) if peers else torch.empty(0, dtype=torch.int32)).to(dev)
        self.dest_row = (torch.cat(slots) if slots else torch.empty(0, dtype=torch.int64)).to(dev)
        self.buffer_rows = max(int(table[:, -1].max()), 1)   # same size on every rank (symmetric allocation)
        self.comm_stream = torch.cuda.Stream() if self.cuda else None
        self.event = torch.cuda.Event() if self.cuda else None

    # -- defaults: torch symmetric memory + the C-ABI push kernel ---------------------------------------------
    def _buffer(self, tail, dtype, device):
        key = (tuple(tail), dtype)
        if key not in self._bufs:
            shape = [self.buffer_rows] + list(tail)
            if self._allocator is not None:
                self._bufs[key] = self._allocator(shape, dtype, device)
            else:
                import warnings
                import torch.distributed._symmetric_memory as symm
                pg = self.group if self.group is not None else dist.group.WORLD
                try:        # older torch wants the group registered first; newer ones do it in rendezvous
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        symm.enable_symm_mem_for_group(pg.group_name)
                except Exception:
                    pass
                buf = symm.empty(shape, dtype=dtype, device=device)
                hdl = symm.rendezvous(buf, pg)
                # the peers' mapped base addresses as a device array (what geot_b200_push_rows indexes by dest_peer)
                bases = torch.tensor([int(a) for a in hdl.buffer_ptrs], dtype=torch.int64, device=device)
                self._bufs[key] = (buf, _PeerBuffer(hdl, bases))
        return self._bufs[key]

    def _push(self, x_mine, buf, hdl):
        nd = self.requests
        if self._pusher is not None:       # (a stand-in also plays the receiving side, so it runs even with nothing to send)
            self._pusher(x_mine, nd.send_rows, self.dest_peer, self.dest_row, buf, hdl)
            return
        if nd.send_rows.numel() == 0:
            return
        from . import abi
        abi.push_rows(x_mine, nd.send_rows, self.dest_peer, self.dest_row, hdl.bases.data_ptr())

    def _barrier(self, hdl, channel):
        if self._barrier_fn is not None:
            self._barrier_fn(hdl, channel)
        else:
            hdl.handle.barrier(channel=channel)

    def aggregate(self, x_local, weight=None, reduce="sum"):
        return self(x_local, weight, reduce)

    def exchanged_rows(self):
        rb = self.shard.row_bounds
        return self.requests.recv_offsets[-1], rb[-1] - (rb[self.rank + 1] - rb[self.rank])

    def __call__(self, x: torch.Tensor, weight: Optional[torch.Tensor] = None, reduce: str = "sum",
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
        assert reduce in ("sum", "mean"), "bucket partials are combined by addition: sum / mean only"
        rb, rank = self.shard.row_bounds, self.rank
        S = self.shard.num_local_rows
        x_mine = x if x.shape[0] == S else self.local_rows(x)       # own rows, or the [N, ...] buffer holding them
        x_mine = x_mine.contiguous()
        tail = list(x_mine.shape[1:])
        if out is None:
            out = x_mine.new_empty([S] + tail)
        if self._parts is None or list(self._parts.shape[2:]) != tail or self._parts.dtype != x_mine.dtype:
            self._parts = x_mine.new_empty([2, S] + tail)
        parts = self._parts
        buf, hdl = self._buffer(tail, x_mine.dtype, x_mine.device)

        def exchange():
            self._barrier(hdl, 0)          # every peer is done reading its receive buffer (its previous call)
            self._push(x_mine, buf, hdl)   # my rows into their slots on the peers
            self._barrier(hdl, 1)          # every peer's rows are in my buffer

        if self.cuda:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                exchange()
                self.event.record(self.comm_stream)
        else:
            exchange()

        w_all = self._permute(weight) if weight is not None else None
        b = self.buckets.bounds
        self._reduce_bucket(0, x_mine, w_all[b[0]:b[1]] if w_all is not None else None, parts[0])
        if self.cuda:
            torch.cuda.current_stream().wait_event(self.event)
        self._reduce_bucket(1, buf, w_all[b[1]:b[2]] if w_all is not None else None, parts[1])
        self._combine(parts, out, reduce)
        return out


class SrcBlockedGather(PipelinedGather):
    """Single-GPU experiment: temporal blocking of the src matrix for L2.  The edges are split stably into ``blocks``
    buckets by src row range (each still dst-sorted) and reduced one bucket after the other, so that at any time the
    gathers touch ``1/blocks`` of the src matrix; the bucket partials are combined in bucket order.  Pays only when
    the saved DRAM re-reads exceed the partial-sum traffic (``2*blocks + 1`` passes over the output): high-degree
    graphs whose src matrix is a small multiple of the L2 (Reddit-shape: 119 MB src, degree 492), not products-like
    ones.  Same kernels, same plans per bucket, no communication."""

    def __init__(self, src_index: torch.Tensor, dst_index: torch.Tensor, num_dst_rows: int, num_src_rows: int,
                 blocks: int, reducer=None, combiner=None, permuter=None):
        E = dst_index.numel()
        self.shard = GraphShard(0, 1, [0, num_dst_rows], [0, E], src_index, dst_index, None)
        self.group, self.world, self.rank = None, 1, 0
        self.cuda = dst_index.is_cuda
        self._reducer, self._combiner, self._permuter = reducer, combiner, permuter
        self._plans, self._ws, self._parts, self._wperm = {}, None, None, None
        self._rowptr, self.needed = None, None
        self.blocks = blocks
        rows_per_block = (num_src_rows + blocks - 1) // blocks
        key = torch.div(src_index, rows_per_block, rounding_mode="floor")
        perm = torch.argsort(key, stable=True)
        counts = torch.bincount(key, minlength=blocks).tolist()
        bounds = [0]
        for c in counts:
            bounds.append(bounds[-1] + int(c))
        self.buckets = SrcBuckets(perm, bounds, src_index[perm].contiguous(), dst_index[perm].contiguous())

    def permute_weight(self, weight: torch.Tensor) -> torch.Tensor:
        """Per-edge weights in bucket order (a copy): pass it with ``permuted=True`` when the weights are static."""
        return self._permute(weight).clone()

    def __call__(self, x: torch.Tensor, weight: Optional[torch.Tensor] = None, reduce: str = "sum",
                 out: Optional[torch.Tensor] = None, permuted: bool = False) -> torch.Tensor:
        assert reduce in ("sum", "mean"), "bucket partials are combined by addition: sum / mean only"
        S = self.shard.num_local_rows
        tail = list(x.shape[1:])
        if out is None:
            out = x.new_empty([S] + tail)
        if self._parts is None or list(self._parts.shape[2:]) != tail or self._parts.dtype != x.dtype:
            self._parts = x.new_empty([self.blocks, S] + tail)
        w_all = None if weight is None else (weight if permuted else self._permute(weight))
        b = self.buckets.bounds
        for k in range(self.blocks):
            self._reduce_bucket(k, x, w_all[b[k]:b[k + 1]] if w_all is not None else None, self._parts[k])
        self._combine(self._parts, out, reduce)
        return out

