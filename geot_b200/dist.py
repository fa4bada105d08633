"""Multi-GPU sharding of the segment-reduction path (new functionality -- the reference
has no distributed code).

One process per GPU (``torch.distributed``, NCCL over NVLink / NVSwitch).  The dst-sorted edge list is
cut into ``world_size`` contiguous dst-row ranges with balanced edge counts; rank g owns rows
``[row_bounds[g], row_bounds[g+1])`` of every node-feature matrix and the edges that point into them,
reduces its own dst slice, and writes only that slice -- there is no reduction collective.  The only
exchange is that of the src feature rows that gather ops read across shard boundaries.
``index_scatter`` needs no communication at all (edge-aligned data is sharded with the edges).

Forms of the exchange:

* ``sharded_gather_scatter``: one ragged all-gather, then one reduction (also serves max / min).
* ``BucketedGather`` (sum / mean): the rank's edges are split ONCE per graph, stably, into two dst-sorted buckets --
  src row local / src row remote.  Per call the remote rows travel on a side stream while the main stream reduces the
  local bucket (writing every row of the output, rows without local edges as zeros); when the rows have landed the
  remote bucket is reduced with ``accumulate`` (``dst += partial``, ``geot_b200_segment_reduce_ex``).  The host cost
  is the same whatever the GPU count: 1 exchange + 2 reductions, no combine pass, no per-call weight permutation
  (the kernel reads ``weight[edge_perm[e]]``).  When the rows of the local bucket would be short (products-shape
  shards) the two passes cost more than the overlap buys and the op runs as ONE pass after the exchange over a buffer
  that holds the own and the received rows (``passes``; chosen from the local bucket's degree).  Two transports:

  - ``"allgather"``: one ragged NCCL all-gather into the ``[N, ...]`` replica buffer;
  - ``"push"``: only the rows a peer's edges actually reference, stored by ONE kernel straight into the requesters'
    symmetric-memory buffers over NVLink (``geot_b200_push_rows``: P2P stores, no NCCL on the data path) between two
    cross-GPU barriers.  Pays off when a shard references a fraction of a peer's rows (products-like graphs).
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


# ---- exchange overlapped with the reduction: two edge buckets ---------------------------------------

@dataclass
class SrcBuckets:
    """A rank's edges split stably by where their src row lives: bucket 0 = the rank's own rows, bucket 1 = rows of
    the other ranks.  Both stay sorted by dst.  ``perm[e]`` is the position of bucketed edge e in the shard's edge
    list (int32: what the kernel indexes the caller's weights with)."""
    perm: torch.Tensor               # [E_local] int32
    bounds: List[int]                # [0, n_local_src, E_local]
    src_index: torch.Tensor          # [E_local] int64: bucket 0 -> LOCAL row ids (into the rank's own rows);
                                     #                  bucket 1 -> ids into the buffer the transport fills
    dst_index: torch.Tensor          # [E_local] int64 local dst rows


@dataclass
class NeededRows:
    """Per-graph state of the needed-rows ("push") transport."""
    recv_counts: List[int]           # [world] rows this rank receives from each owner (0 for itself)
    recv_offsets: List[int]          # [world+1] their offsets in the receive buffer, in STEP order (rank+1, rank+2, ...)
    send_counts: List[int]           # [world] rows each peer asked this rank for
    send_rows: torch.Tensor          # [sum(send_counts)] LOCAL row ids to push, grouped by peer in step order
    dest_peer: torch.Tensor          # [sum(send_counts)] int32: the peer every pushed row goes to
    dest_row: torch.Tensor           # [sum(send_counts)] int64: its slot in that peer's receive buffer
    buffer_rows: int                 # rows of the (symmetric: same size everywhere) receive buffer


def split_local_remote(shard: GraphShard, phase_steps=None):
    """(perm [E] int64, bounds): stable split of the shard's edges by where their src row lives -- src-local edges
    first, then the src-remote ones.  ``phase_steps`` (optional, ascending, last = world-1) cuts the remote edges further
    by the exchange step that delivers their row (step k = the rows of rank+k): remote group p holds the owners at
    steps ``(phase_steps[p-1], phase_steps[p]]``.  ``bounds`` = edge offsets of the groups, ``len = groups + 1``."""
    rb, rank, world = shard.row_bounds, shard.rank, shard.world_size
    s = shard.src_index
    if phase_steps is None or len(phase_steps) <= 1 or world <= 2:
        key = ((s < rb[rank]) | (s >= rb[rank + 1])).to(torch.int8)
        groups = 2
    else:
        cuts = torch.tensor(rb[1:-1], dtype=s.dtype, device=s.device)
        step = (torch.bucketize(s, cuts, right=True) - rank) % world            # 0 = local, k = delivered in step k
        key = torch.bucketize(step, torch.tensor([0] + list(phase_steps[:-1]), dtype=step.dtype, device=s.device), right=False).to(torch.int8)
        groups = 1 + len(phase_steps)
    perm = torch.argsort(key, stable=True)
    counts = torch.bincount(key.long(), minlength=groups).tolist()
    bounds = [0]
    for c in counts:
        bounds.append(bounds[-1] + int(c))
    return perm, bounds


def build_needed_rows(shard: GraphShard, remote_src: torch.Tensor, group=None, own_rows_first: bool = False):
    """One-time request exchange of the push transport.  ``remote_src``: the global src ids of the rank's remote-src
    edges.  Returns (NeededRows, compact ids of those edges into the receive buffer).  ``own_rows_first``: every
    rank's buffer starts with its own rows (single-pass form), the received rows follow.

    Collectives: one all-gather of the [world] request counts and offsets, then ``world-1`` staggered send/recv steps
    of the request lists (works on NCCL and gloo alike)."""
    world, rank, rb = shard.world_size, shard.rank, shard.row_bounds
    dev, idt = remote_src.device, remote_src.dtype
    cuts = torch.tensor(rb[1:-1], dtype=idt, device=dev)
    owner = torch.bucketize(remote_src, cuts, right=True) if world > 1 else torch.zeros_like(remote_src)
    compact = torch.empty_like(remote_src)
    need, recv_counts, recv_offsets = {}, [0] * world, [0]
    for k in range(1, world):                       # step order: owner rank+k
        g = (rank + k) % world
        sel = owner == g
        s_k = remote_src[sel]
        uniq = torch.unique(s_k)                    # sorted
        need[g] = (uniq - rb[g]).contiguous()       # as the owner's local row ids
        recv_counts[g] = int(uniq.numel())
        compact[sel] = recv_offsets[-1] + torch.searchsorted(uniq, s_k)
        recv_offsets.append(recv_offsets[-1] + recv_counts[g])
    recv_offsets.append(recv_offsets[-1])           # world+1 entries
    # table[r] = rank r's (recv_counts by owner, recv_offsets by step)
    mine = torch.tensor(recv_counts + recv_offsets, dtype=torch.int64, device=dev)
    table = torch.empty(world * (2 * world + 1), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(table, mine, group=group)
    table = table.view(world, 2 * world + 1).cpu()
    send_counts = [int(table[g][rank]) for g in range(world)]
    send_counts[rank] = 0
    lists, peers, slots = [], [], []
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
        # my rows for `to` land at the offset of ITS step k (owner = me) in its receive buffer
        n = send_counts[to]
        base = (rb[to + 1] - rb[to]) if own_rows_first else 0
        peers.append(torch.full((n,), to, dtype=torch.int32))
        slots.append(base + int(table[to][world + k - 1]) + torch.arange(n, dtype=torch.int64))
    send_rows = torch.cat(lists) if lists else torch.empty(0, dtype=idt, device=dev)
    n_local = rb[rank + 1] - rb[rank]
    assert send_rows.numel() == 0 or (int(send_rows.min()) >= 0 and int(send_rows.max()) < n_local)
    dest_peer = (torch.cat(peers) if peers else torch.empty(0, dtype=torch.int32)).to(dev)
    dest_row = (torch.cat(slots) if slots else torch.empty(0, dtype=torch.int64)).to(dev)
    if own_rows_first:
        compact += n_local
        buffer_rows = max(max(int(table[g][2 * world]) + rb[g + 1] - rb[g] for g in range(world)), 1)
    else:
        buffer_rows = max(int(table[:, 2 * world].max()), 1)
    nd = NeededRows(recv_counts, recv_offsets, send_counts, send_rows.contiguous(), dest_peer, dest_row, buffer_rows)
    return nd, compact


@dataclass
class _PeerBuffer:
    """A symmetric-memory receive buffer as the default (CUDA) push transport holds it."""
    handle: object                   # torch _SymmetricMemory: barrier(channel)
    bases: torch.Tensor              # [world] int64 on the device: every peer's mapped base address of the buffer


class BucketedGather:
    """``gather_(weight_)scatter`` / ``mh_spmm`` on a dst-row shard with the src-row exchange overlapped (sum / mean).

    ``transport``: ``"allgather"`` or ``"push"`` (module docstring).  Rows may be ``[n, F]`` (weights ``[E]`` or none)
    or ``[n, H, F]`` with per-head weights ``[E, H]``.  Call with this rank's OWN rows ``x_local``.

    Injectable stand-ins (the gloo tests check the host logic on CPU; defaults = the C-ABI kernels, NCCL, torch
    symmetric memory):
      ``reducer(x, src_ids, dst_ids, weight, edge_perm, S, out, accumulate, mean_rowptr, reduce)``
      ``allocator(shape, dtype, device) -> (buffer, handle)``; ``pusher(x_local, nd, buffer, handle[, steps=(a, b)])``;
      ``barrier(handle, channel)``."""

    # two passes pay when the rows of the local bucket are long enough that closing every dst row twice (and adding
    # into it) costs less than the overlap buys: Reddit-shape shards (61 local edges per row at 8 GPUs) yes,
    # products-shape ones (3) no -- measured at N = 2, profiles/r02g_n2_exch.txt
    MIN_LOCAL_DEGREE = 16
    PUSH_CTAS = 2 * 148        # grid of a push that runs beside a reduction (geot_b200_push_rows_ex)

    def __init__(self, shard: GraphShard, group=None, transport: str = "allgather", reducer=None, allocator=None,
                 pusher=None, barrier=None, passes: int = 0, phases: int = 0, phase_steps=None):
        assert transport in ("allgather", "push") and passes in (0, 1, 2) and phases >= 0
        assert shard.src_index is not None
        self.shard, self.group, self.transport = shard, group, transport
        self.world, self.rank = shard.world_size, shard.rank
        self.cuda = shard.dst_index.is_cuda
        self._reducer, self._allocator, self._pusher, self._barrier_fn = reducer, allocator, pusher, barrier
        rb, rank = shard.row_bounds, shard.rank
        E = shard.dst_index.numel()
        assert E < 2 ** 31, "edge_perm is int32"
        # push transport, overlapped form: the exchange may run in `phases` rounds (round p delivers the rows of the owners
        # at steps (phase_steps[p-1], phase_steps[p]]), each followed by the reduction of the edges it serves, so that only
        # the first round's transfer is exposed.  Default: 2 rounds from 6 GPUs on (at 8 GPUs a Reddit-shape rank receives
        # 104 MB per call -- 0.16 ms on the wire against 0.06 ms of src-local work to hide it behind), else 1.
        if phases == 0:
            phases = 2 if (transport == "push" and self.world >= 6) else 1
        if transport != "push" or self.world <= 2:
            phases = 1
        phases = min(phases, max(self.world - 1, 1))
        self.phase_steps = [((self.world - 1) * (p + 1)) // phases for p in range(phases)]      # ascending, last = world-1
        if phase_steps is not None and transport == "push" and self.world > 2:                 # explicit round boundaries
            self.phase_steps = sorted(set(int(v) for v in phase_steps if 0 < int(v) < self.world - 1)) + [self.world - 1]
            phases = len(self.phase_steps)
        perm, bounds = split_local_remote(shard, self.phase_steps)
        n_local = bounds[1]
        if passes == 0:
            # the same choice on every rank (the transports are collective): decided on the total local-src edge count
            t = torch.tensor([float(n_local), float(max(shard.num_local_rows, 1))], dtype=torch.float64, device=shard.dst_index.device)
            if self.world > 1:
                dist.all_reduce(t, group=group)
            passes = 2 if t[0].item() / t[1].item() >= self.MIN_LOCAL_DEGREE else 1
        self.passes = passes
        self.phases = phases if passes == 2 else 1
        self.needed = None
        if passes == 1:
            # single pass: the shard's edge list as it is (no permutation), src ids into ONE buffer that holds the own
            # rows and the received ones -- the [N, ...] replica (all-gather) or [own rows | needed rows] (push)
            src = shard.src_index.clone()
            if transport == "push":
                remote = (src < rb[rank]) | (src >= rb[rank + 1])
                self.needed, compact = build_needed_rows(shard, src[remote], group, own_rows_first=True)
                src[~remote] -= rb[rank]
                src[remote] = compact
            self.buckets = SrcBuckets(None, [0, E, E], src.contiguous(), shard.dst_index)
            self.phase_steps = [self.world - 1]
        else:
            src = shard.src_index[perm]
            src[:n_local] -= rb[rank]                      # bucket 0 reads the rank's own rows
            if transport == "push":
                self.needed, compact = build_needed_rows(shard, src[n_local:].clone(), group)
                src[n_local:] = compact
            self.buckets = SrcBuckets(perm.to(torch.int32).contiguous(), bounds, src.contiguous(),
                                      shard.dst_index[perm].contiguous())
        self._plans, self._ws, self._bufs, self._mean_rowptr, self._wperm = {}, {}, {}, None, None
        self._replica = None
        # the exchange runs on a HIGH-priority side stream: its small push grid gets the next free SM slots although the
        # src-local reduction was launched over the whole GPU
        self.comm_stream = torch.cuda.Stream(priority=-1) if self.cuda else None
        self.events = [torch.cuda.Event() for _ in range(self.phases)] if self.cuda else None

    # -- per-graph facts ------------------------------------------------------------------------------
    def exchanged_rows(self):
        """(rows received per call, rows a full exchange would receive)."""
        rb = self.shard.row_bounds
        full = rb[-1] - (rb[self.rank + 1] - rb[self.rank])
        return (self.needed.recv_offsets[-1] if self.needed is not None else full), full

    def main_launches(self) -> int:
        """Reduction launches of one call as set up so far (a bucket that is src-blocked for the L2 takes one per block;
        known after the first call at a given row width)."""
        b = self.buckets.bounds
        n = 0
        for k in range(len(b) - 1):
            if b[k + 1] == b[k]:
                continue
            blocks = [v[1] for key, v in self._ws.items() if key[0] == k]
            n += blocks[-1].n_blocks if (blocks and blocks[-1] is not None) else 1
        return n

    def local_rows(self, x_full: torch.Tensor) -> torch.Tensor:
        rb = self.shard.row_bounds
        return x_full[rb[self.rank]:rb[self.rank + 1]]

    # -- default device pieces (C ABI, NCCL, symmetric memory) ---------------------------------------------
    def _reduce_bucket(self, k, x, weight, out, reduce, accumulate):
        b = self.buckets
        e0, e1 = b.bounds[k], b.bounds[k + 1]
        S = self.shard.num_local_rows
        if S == 0:
            return
        if e1 == e0:
            if not accumulate:
                out.zero_()
            return
        si, di, perm = b.src_index[e0:e1], b.dst_index[e0:e1], (b.perm[e0:e1] if b.perm is not None else None)
        mean_rowptr = None
        if reduce == "mean":
            if self._mean_rowptr is None:
                deg = torch.bincount(self.shard.dst_index, minlength=S)
                self._mean_rowptr = torch.cat([deg.new_zeros(1), deg.cumsum(0)]).contiguous()
            mean_rowptr = self._mean_rowptr
        if self._reducer is not None:
            self._reducer(x, si, di, weight, perm, S, out, accumulate, mean_rowptr, reduce)
            return
        from . import abi
        if weight is not None and weight.dim() == 2:       # [E, H] weights arrive in bucket order (_bucket_order)
            weight, perm = weight[e0:e1], None
        if k not in self._plans:
            self._plans[k] = abi.DevicePlan(di, S)
        W = x[0].numel()
        H = x.shape[1] if x.dim() == 3 else 1
        key = (k, W, x.dtype)                              # scratch / blocking are per (bucket, row width, dtype)
        if key not in self._ws:
            # src-row blocking for the L2 inside the bucket (the remote bucket of a Reddit-shape shard gathers from the
            # whole replica: same reuse, same working set as on one GPU).  The regrouped list carries its own
            # permutation; composed with the bucket's, it indexes the caller's weights directly.
            blocks = None
            nb = 1
            if H == 1 and self.cuda and x.element_size() >= 4:         # (16-bit outputs would be rounded once per pass)
                nb = abi.src_blocks_suggest(e1 - e0, S, self._rows_touched(k, x.shape[0]), W * x.element_size())
            if nb > 1:
                blocks = abi.SrcBlocks(si, di, x.shape[0], nb)
                if perm is not None:
                    blocks.composed_perm = perm[blocks.edge_perm.long()].contiguous()
                    blocks.c.edge_perm = blocks.composed_perm.data_ptr()
            self._ws[key] = (abi.Workspace(e1 - e0, W, x.dtype, x.device, src_blocks=blocks), blocks)
        ws, blocks = self._ws[key]
        if blocks is not None:
            abi.segment_reduce(x, si, di, weight, reduce, S=S, H=H, plan=self._plans[k], out=out, workspace=ws,
                               accumulate=accumulate, mean_rowptr=mean_rowptr, src_blocks=blocks)
            return
        abi.segment_reduce(x, si, di, weight, reduce, S=S, H=H, plan=self._plans[k], out=out, workspace=ws,
                           accumulate=accumulate, edge_perm=perm if weight is not None else None,
                           mean_rowptr=mean_rowptr if self.passes == 2 else None)

    def _rows_touched(self, k, n_rows):
        """Rows of its operand that bucket k can gather from (its working set for the L2): a remote bucket of the push
        transport reads only the rows its exchange round delivered."""
        if k == 0 or self.needed is None or self.passes == 1:
            return n_rows
        lo, hi = self._phase_rows(k - 1)
        return max(hi - lo, 1)

    def _phase_rows(self, p):
        """[lo, hi) rows of the receive buffer that exchange round p delivers."""
        st = [0] + self.phase_steps
        off = self.needed.recv_offsets
        return off[st[p]], off[st[p + 1]]

    def _phase_sends(self, p):
        """[lo, hi) of the send list that exchange round p pushes (the list is grouped by peer in step order)."""
        st = [0] + self.phase_steps
        cnt = [self.needed.send_counts[(self.rank - k) % self.world] for k in range(1, self.world)]
        return sum(cnt[:st[p]]), sum(cnt[:st[p + 1]])

    def _bucket_order(self, weight):
        """Per-head weights [E, H] in bucket order (one permutation pass per call: the kernel's edge_perm serves one
        weight per edge only)."""
        if self.buckets.perm is None:
            return weight
        from . import abi
        if self._wperm is None or self._wperm.shape != weight.shape or self._wperm.dtype != weight.dtype:
            self._wperm = torch.empty_like(weight)
            self._perm64 = self.buckets.perm.long()
        return abi.permute_edges(weight, self._perm64, self._wperm)

    def _buffer(self, tail, dtype, device):
        """The buffer bucket 1 gathers from: the [N, ...] replica (allgather) or the symmetric receive buffer (push)."""
        key = (tuple(tail), dtype)
        if key in self._bufs:
            return self._bufs[key]
        if self.transport == "allgather":
            self._bufs[key] = (torch.empty([self.shard.row_bounds[-1]] + list(tail), dtype=dtype, device=device), None)
        else:
            shape = [self.needed.buffer_rows] + list(tail)
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
                bases = torch.tensor([int(a) for a in hdl.buffer_ptrs], dtype=torch.int64, device=device)
                self._bufs[key] = (buf, _PeerBuffer(hdl, bases))
        return self._bufs[key]

    def _exchange(self, x_local, buf, hdl, events=None):
        """The src-row exchange on the current stream.  ``events`` (overlapped push form): one per exchange round, recorded
        when that round's rows have landed."""
        if self.transport == "allgather":
            all_gather_rows(x_local, self.shard.row_bounds, self.group, out=buf)
            if events is not None:
                events[0].record()
            return
        nd = self.needed
        if self.passes == 1 and x_local.shape[0]:
            buf[: x_local.shape[0]].copy_(x_local)          # the buffer starts with the rank's own rows
        self._barrier(hdl, 0)              # every peer is done reading its receive buffer (its previous call)
        for p in range(self.phases):
            if self._pusher is not None:   # (a stand-in also plays the receiving side: runs even with nothing to send)
                if self.phases == 1:
                    self._pusher(x_local, nd, buf, hdl)
                else:                      # round p = exchange steps (a, b]: to ranks rank-k, from ranks rank+k
                    st = [0] + self.phase_steps
                    self._pusher(x_local, nd, buf, hdl, steps=(st[p], st[p + 1]))
            else:
                lo, hi = self._phase_sends(p) if self.phases > 1 else (0, nd.send_rows.numel())
                if hi > lo:
                    from . import abi
                    # overlapped form: a small grid (2 CTAs per SM) that shares the SMs with the running reduction
                    abi.push_rows(x_local, nd.send_rows[lo:hi], nd.dest_peer[lo:hi], nd.dest_row[lo:hi], hdl.bases.data_ptr(),
                                  max_ctas=self.PUSH_CTAS if self.passes == 2 else 0)
            self._barrier(hdl, 1)          # every peer's rows of this round are in my buffer
            if events is not None:
                events[p].record()

    def _barrier(self, hdl, channel):
        if self._barrier_fn is not None:
            self._barrier_fn(hdl, channel)
        else:
            hdl.handle.barrier(channel=channel)

    # -- the op -------------------------------------------------------------------------------------------
    def __call__(self, x_local: torch.Tensor, weight: Optional[torch.Tensor] = None, reduce: str = "sum",
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
        assert reduce in ("sum", "mean"), "bucket partials add up: sum / mean only (max / min: sharded_gather_scatter)"
        S = self.shard.num_local_rows
        assert x_local.shape[0] == S, "pass this rank's own rows"
        x_local = x_local.contiguous()
        tail = list(x_local.shape[1:])
        if out is None:
            out = x_local.new_empty([S] + tail)
        buf, hdl = self._buffer(tail, x_local.dtype, x_local.device)
        if self.passes == 1:               # nothing to overlap with: exchange, then ONE reduction over the buffer
            self._exchange(x_local, buf, hdl)
            self._reduce_bucket(0, buf, weight, out, reduce, accumulate=False)
            return out
        if self.cuda:
            self.comm_stream.wait_stream(torch.cuda.current_stream())   # x_local is final; buf's old rows were consumed
            with torch.cuda.stream(self.comm_stream):
                self._exchange(x_local, buf, hdl, self.events)
        else:
            self._exchange(x_local, buf, hdl)
        if weight is not None and weight.dim() == 2 and self._reducer is None:
            weight = self._bucket_order(weight)
        self._reduce_bucket(0, x_local, weight, out, reduce, accumulate=False)
        for p in range(self.phases):       # the edges served by exchange round p, as soon as its rows have landed
            if self.cuda:
                torch.cuda.current_stream().wait_event(self.events[p])
            self._reduce_bucket(1 + p, buf, weight, out, reduce, accumulate=True)
        return out

    def aggregate(self, x_local: torch.Tensor, weight: Optional[torch.Tensor] = None, reduce: str = "sum") -> torch.Tensor:
        """The op on this rank's own rows (what a layer produces); returns a new ``[n_local_rows, ...]`` tensor."""
        return self(x_local, weight, reduce)
