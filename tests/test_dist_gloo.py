"""world_size-2, -4 and -8 `gloo` tests of the multi-GPU host logic (geot_b200/dist.py) on CPU.

The reduction kernels need a GPU, so here each rank reduces its shard with the CPU oracle (used as the
checker of the sharding logic): edge-balanced bounds, local index rebasing, ragged all-gather of src
rows, and that the concatenated per-rank results equal the unsharded result.
"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, hub=False):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from geot_b200 import dist as gdist
    g = torch.Generator().manual_seed(0)          # same graph on every rank
    N, E, F = 301, 6000, 12
    deg_w = torch.rand(N, generator=g) ** 3         # skewed in-degrees, some zero-degree rows
    if hub:
        deg_w[N // 2] = 4.0 * float(deg_w.sum())    # one row holds ~80 % of the edges: some rank's row range is empty
    dst = torch.multinomial(deg_w, E, replacement=True, generator=g).sort().values
    dst[-1] = N - 1
    src_index = torch.randint(0, N, (E,), generator=g)
    weight = torch.rand(E, generator=g)
    x = torch.rand(N, F, generator=g)

    shard = gdist.shard_graph(src_index, dst, weight, rank, world)
    rb, eb = shard.row_bounds, shard.edge_bounds
    assert rb[0] == 0 and rb[-1] == N and eb[0] == 0 and eb[-1] == E
    assert all(rb[i] <= rb[i + 1] for i in range(world)) and all(eb[i] <= eb[i + 1] for i in range(world))
    # cuts are at segment boundaries
    for gg in range(1, world):
        if 0 < eb[gg] < E:
            assert dst[eb[gg] - 1] < rb[gg] <= dst[eb[gg]]
    assert hub or shard.imbalance < 1.2
    if hub:
        assert any(rb[i] == rb[i + 1] for i in range(world))          # the case under test: an empty shard

    # ragged all-gather of the row shards reproduces the replicated matrix
    x_local = x[rb[rank]:rb[rank + 1]].clone()
    x_full = gdist.all_gather_rows(x_local, rb)
    assert torch.equal(x_full, x)

    # each rank reduces its own dst slice (checker = CPU oracle), no reduction collective
    for reduce in ["sum", "mean", "max"]:
        if shard.num_local_edges:
            local = oracle.gather_weight_scatter(shard.src_index, shard.dst_index, shard.weight, x_full, reduce,
                                                 S=shard.num_local_rows)
        else:
            local = torch.zeros(shard.num_local_rows, F)
        full = oracle.gather_weight_scatter(src_index, dst, weight, x, reduce)
        assert torch.equal(local, full[rb[rank]:rb[rank + 1]]), reduce
    # exchange overlapped with the reduction (BucketedGather): stable local / remote split, first pass writes every row,
    # second pass accumulates, weights read through edge_perm, mean by the full degree.  CPU stand-in for the device
    # kernel (the oracle as the checker) honouring the same contract as geot_b200_segment_reduce_ex; the host logic
    # under test is the bucketing, the src ids per transport, the slot arithmetic of the push and the pass order.
    def reducer(xf, si, di, w, perm, S, out, accumulate, mean_rowptr, reduce):
        ww = None
        if w is not None:
            ww = w[perm.long()] if perm is not None else w        # (single pass: the shard's own edge order)
        part = oracle.segment_reduce(xf, si, di, ww, "sum", S=S, H=(xf.shape[1] if xf.dim() == 3 else 1))
        if reduce == "mean":
            deg = (mean_rowptr[1:] - mean_rowptr[:-1]).clamp_min(1).to(part.dtype)
            part = part / deg.view([-1] + [1] * (part.dim() - 1))
        if accumulate:
            out += part
        else:
            out.copy_(part)

    def check_buckets(bg):
        b = bg.buckets
        if bg.passes == 1:
            assert b.perm is None and b.bounds == [0, shard.num_local_edges, shard.num_local_edges]
            assert torch.equal(b.dst_index, shard.dst_index)
            return
        assert b.bounds[0] == 0 and b.bounds[-1] == shard.num_local_edges and b.perm.dtype == torch.int32
        assert len(b.bounds) == 2 + bg.phases and b.bounds == sorted(b.bounds)
        glob = shard.src_index[b.perm.long()]
        loc = (glob >= rb[rank]) & (glob < rb[rank + 1])
        assert bool(loc[: b.bounds[1]].all()) and not bool(loc[b.bounds[1]:].any())
        assert torch.equal(b.src_index[: b.bounds[1]], glob[: b.bounds[1]] - rb[rank])
        cuts = torch.tensor(rb[1:-1], dtype=torch.int64)
        step = (torch.bucketize(glob, cuts, right=True) - rank) % world  # exchange step that delivers the edge's row
        st = [0] + bg.phase_steps
        for k in range(len(b.bounds) - 1):
            d_k = b.dst_index[b.bounds[k]:b.bounds[k + 1]]
            assert bool((d_k[1:] >= d_k[:-1]).all())                   # stable split keeps dst order
            if k >= 1:                                                  # remote group k-1 = the owners of its exchange round
                s_k = step[b.bounds[k]:b.bounds[k + 1]]
                assert bool(((s_k > st[k - 1]) & (s_k <= st[k])).all())

    def expect(reduce, weighted):
        full = (oracle.gather_weight_scatter(src_index, dst, weight, x, reduce) if weighted
                else oracle.gather_scatter(src_index, dst, x, reduce))
        return full[rb[rank]:rb[rank + 1]]

    auto = gdist.BucketedGather(shard, transport="allgather", reducer=reducer)
    assert auto.passes in (1, 2)                                        # (chosen from the local bucket's degree, same on every rank)
    for passes in (2, 1):
        bg = gdist.BucketedGather(shard, transport="allgather", reducer=reducer, passes=passes)
        check_buckets(bg)
        if passes == 2:
            assert torch.equal(bg.buckets.src_index[bg.buckets.bounds[1]:], shard.src_index[bg.buckets.perm.long()][bg.buckets.bounds[1]:])
        for reduce in ["sum", "mean"]:
            for weighted in (False, True):
                got = bg(x_local.clone(), shard.weight if weighted else None, reduce)
                assert torch.allclose(got, expect(reduce, weighted), rtol=1e-5, atol=1e-6), ("allgather", passes, reduce, weighted)
        buf, _ = bg._buffer([F], x.dtype, x.device)
        assert torch.equal(buf, x)                                      # the exchange rebuilt the replica

    # push transport: ONE push of the requested rows into slots of the requesters' buffers.  The stand-in pusher ships
    # (slot, row) pairs over gloo and the RECEIVER stores each row where the SENDER's slot says, so the slot arithmetic
    # (dest_peer / dest_row) and the compact src ids are what is under test.
    holder = {}

    def pusher(x_mine, nd, buf, hdl, steps=None):
        ops, inbox = [], {}
        a, b = steps if steps is not None else (0, world - 1)          # exchange steps (a, b] of this round
        for p in range(world):
            if p == rank:
                continue
            m = nd.dest_peer == p
            if not (a < (rank - p) % world <= b):                       # I serve rank-k in step k
                m = m & False
            if int(m.sum()):
                ops.append(dist.P2POp(dist.isend, nd.dest_row[m].contiguous(), p, tag=1))
                ops.append(dist.P2POp(dist.isend, x_mine[nd.send_rows[m]].contiguous(), p, tag=2))
            n = nd.recv_counts[p] if (a < (p - rank) % world <= b) else 0   # rank+k's rows arrive in step k
            if n:
                inbox[p] = (torch.empty(n, dtype=torch.int64), torch.empty([n] + list(x_mine.shape[1:]), dtype=x_mine.dtype))
                ops.append(dist.P2POp(dist.irecv, inbox[p][0], p, tag=1))
                ops.append(dist.P2POp(dist.irecv, inbox[p][1], p, tag=2))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        for p, (slots, data) in inbox.items():
            assert slots.numel() == torch.unique(slots).numel() and int(slots.max()) < buf.shape[0]
            buf[slots] = data

    kw_push = dict(transport="push", reducer=reducer, pusher=pusher,
                   allocator=lambda shape, dtype, device: (torch.full(shape, float("nan"), dtype=dtype), None),
                   barrier=lambda hdl, channel: dist.barrier())
    pp = gdist.BucketedGather(shard, passes=2, **kw_push)
    check_buckets(pp)
    got_rows, full_rows = pp.exchanged_rows()
    assert 0 <= got_rows <= full_rows
    nd = pp.needed
    remote_glob = shard.src_index[pp.buckets.perm.long()][pp.buckets.bounds[1]:]
    assert got_rows == torch.unique(remote_glob).numel() == nd.recv_offsets[-1]
    c = pp.buckets.src_index[pp.buckets.bounds[1]:]
    if c.numel():
        assert int(c.min()) >= 0 and int(c.max()) < nd.recv_offsets[-1] <= nd.buffer_rows
    # the exchange in rounds: every round's rows land in their own range of the buffer and serve their own edge group
    for phases in (1, 2, 3):
        pk = gdist.BucketedGather(shard, passes=2, phases=phases, **kw_push)
        assert pk.phases == (1 if world <= 2 else min(phases, world - 1)) and pk.phase_steps[-1] == world - 1
        check_buckets(pk)
        assert pk.exchanged_rows() == pp.exchanged_rows()
        sends = [pk._phase_sends(p) for p in range(pk.phases)]
        rows = [pk._phase_rows(p) for p in range(pk.phases)]
        assert sends[0][0] == 0 and sends[-1][1] == pk.needed.send_rows.numel() and rows[0][0] == 0 and rows[-1][1] == pk.needed.recv_offsets[-1]
        assert all(sends[i][1] == sends[i + 1][0] and rows[i][1] == rows[i + 1][0] for i in range(pk.phases - 1))
        for k in range(1, pk.phases + 1):                               # a remote group reads only its round's rows
            c_k = pk.buckets.src_index[pk.buckets.bounds[k]:pk.buckets.bounds[k + 1]]
            if c_k.numel():
                assert rows[k - 1][0] <= int(c_k.min()) and int(c_k.max()) < rows[k - 1][1]
        for reduce, weighted in (("sum", True), ("mean", False)):
            for _ in range(2):
                got = pk(x_local.clone(), shard.weight if weighted else None, reduce)
            assert torch.allclose(got, expect(reduce, weighted), rtol=1e-5, atol=1e-6), ("push rounds", phases, reduce, weighted)
    p1 = gdist.BucketedGather(shard, passes=1, **kw_push)               # single pass: [own rows | needed rows] in one buffer
    check_buckets(p1)
    assert p1.exchanged_rows() == pp.exchanged_rows()
    n_loc = shard.num_local_rows
    is_remote = (shard.src_index < rb[rank]) | (shard.src_index >= rb[rank + 1])
    assert torch.equal(p1.buckets.src_index[~is_remote], shard.src_index[~is_remote] - rb[rank])
    if int(is_remote.sum()):
        cr = p1.buckets.src_index[is_remote]
        assert int(cr.min()) >= n_loc and int(cr.max()) < n_loc + p1.needed.recv_offsets[-1] <= p1.needed.buffer_rows
    for obj in (pp, p1):
        for reduce in ["sum", "mean"]:
            for weighted in (False, True):
                for _ in range(2):                                      # twice: the buffer is reused across calls
                    got = obj(x_local.clone(), shard.weight if weighted else None, reduce)
                assert torch.allclose(got, expect(reduce, weighted), rtol=1e-5, atol=1e-6), ("push", obj.passes, reduce, weighted)
    # sparse referencing: when the edges touch few distinct src rows the exchange shrinks accordingly
    few = src_index % 7
    sh2 = gdist.shard_graph(few, dst, None, rank, world, row_bounds=rb, edge_bounds=eb)
    p2 = gdist.BucketedGather(sh2, **kw_push)
    assert p2.exchanged_rows()[0] <= 7
    got = p2(x_local.clone(), None, "sum")
    assert torch.allclose(got, oracle.gather_scatter(few, dst, x, "sum")[rb[rank]:rb[rank + 1]], rtol=1e-5, atol=1e-6)

    # the DEFAULT branches of the push transport (torch symmetric memory + abi.push_rows), with those two modules faked:
    # symm.empty / rendezvous hand out local buffers with a barrier, abi.push_rows delivers over gloo like `pusher`
    import torch.distributed._symmetric_memory as symm
    from geot_b200 import abi

    class FakeHandle:
        live = []

        def __init__(self, buf):
            self.buf, self.buffer_ptrs = buf, [buf.data_ptr() + 4096 * r for r in range(world)]
            FakeHandle.live.append(self)

        def barrier(self, channel=0):
            dist.barrier()

    def fake_push_rows(x_mine, rows, dest_peer, dest_row, bases_ptr, aligned16=True, max_ctas=0):
        h = [h for h in FakeHandle.live if list(h.buf.shape[1:]) == list(x_mine.shape[1:]) and h.buf.dtype == x_mine.dtype][-1]
        pusher(x_mine, holder["pd"].needed, h.buf, None)

    real = (symm.empty, symm.rendezvous, abi.push_rows)
    symm.empty = lambda shape, dtype=None, device=None: torch.full(list(shape), float("nan"), dtype=dtype)
    symm.rendezvous = lambda t, group: FakeHandle(t)
    abi.push_rows = fake_push_rows
    try:
        pd = holder["pd"] = gdist.BucketedGather(shard, transport="push", reducer=reducer, passes=2, phases=1)
        # (a rank with nothing to send skips abi.push_rows on the default path but must still receive in this emulation)
        if pd.needed.send_rows.numel() == 0:
            pd._pusher = pusher
        for reduce in ["sum", "mean"]:
            got = pd(x_local.clone(), shard.weight, reduce)
            assert torch.allclose(got, expect(reduce, True), rtol=1e-5, atol=1e-6), ("push default branches", reduce)
        buf, pb = pd._buffer([F], x_local.dtype, x_local.device)
        assert pb.bases.tolist() == pb.handle.buffer_ptrs and pb.bases.dtype == torch.int64
    finally:
        symm.empty, symm.rendezvous, abi.push_rows = real

    # multi-head rows [N, H, F] with per-head weights [E, H] (mh_spmm) through both transports
    Hh = 3
    xh = torch.rand(N, Hh, 4, generator=g)
    wh = torch.rand(E, Hh, generator=g)
    sh_h = gdist.shard_graph(src_index, dst, wh, rank, world, row_bounds=rb, edge_bounds=eb)
    exp_h = oracle.mh_spmm(src_index, dst, wh, xh)[rb[rank]:rb[rank + 1]]
    xh_local = xh[rb[rank]:rb[rank + 1]].clone()
    for make in (lambda: gdist.BucketedGather(sh_h, transport="allgather", reducer=reducer),
                 lambda: gdist.BucketedGather(sh_h, **kw_push)):
        obj = make()
        got = obj.aggregate(xh_local, sh_h.weight, "sum")
        assert torch.allclose(got, exp_h, rtol=1e-5, atol=1e-6), obj.transport

    # 3-layer GCN / GraphSAGE forward on the shard (BASELINE configs[4] at N > 1): every layer through the bucketed
    # exchange (both transports; hidden width != input width: per-shape buffers), against the plain-torch restatement
    from geot_b200 import gnn
    from tests.helpers import gnn_restatement as restate
    torch.manual_seed(7)                                  # same random weights on every rank
    gcn, sage = gnn.GCN(F, 16, 3), gnn.GraphSAGE(F, 16, 3)
    norm = gnn.gcn_norm(src_index, dst, N, weight)
    sh_gcn = gdist.shard_graph(src_index, dst, norm, rank, world, row_bounds=rb, edge_bounds=eb)
    sh_sage = gdist.shard_graph(src_index, dst, None, rank, world, row_bounds=rb, edge_bounds=eb)
    with torch.no_grad():
        exp_gcn = restate.forward(gcn, x, src_index, dst, norm)[rb[rank]:rb[rank + 1]]
        exp_sage = restate.forward(sage, x, src_index, dst)[rb[rank]:rb[rank + 1]]
        for kw in (dict(transport="allgather", reducer=reducer), kw_push):
            got = gnn.forward_sharded(gcn, x_local, sh_gcn, gather=gdist.BucketedGather(sh_gcn, **kw))
            assert torch.allclose(got, exp_gcn, rtol=1e-4, atol=1e-5), ("gcn", kw["transport"])
            got = gnn.forward_sharded(sage, x_local, sh_sage, gather=gdist.BucketedGather(sh_sage, **kw))
            assert torch.allclose(got, exp_sage, rtol=1e-4, atol=1e-4), ("sage", kw["transport"])
    dist.barrier()
    q.put((rank, shard.num_local_edges))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,hub", [(2, False), (4, False), (4, True), (8, True)])
def test_sharding_gloo(world, hub):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, hub)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    got = sorted(q.get(timeout=5) for _ in range(world))
    assert sum(e for _, e in got) == 6000


def test_shard_bounds_rule():
    from geot_b200 import dist as gdist
    rowptr = torch.tensor([0, 5, 5, 9, 20, 21, 30])
    rows, edges = gdist.shard_bounds_from_rowptr(rowptr, 3)
    assert rows[0] == 0 and rows[-1] == 6 and edges == [int(rowptr[r]) for r in rows]
    # targets 10 and 20: nearest boundaries are 9 (row 3) and 20 (row 4)
    assert edges == [0, 9, 20, 30]
    rows1, edges1 = gdist.shard_bounds_from_rowptr(rowptr, 1)
    assert rows1 == [0, 6] and edges1 == [0, 30]


def test_shard_bounds_properties_randomised():
    """shard_bounds_from_rowptr on random degree sequences (hubs, empty rows, more parts than rows): the bounds cover
    the rows, are monotone, sit on segment boundaries, and every cut is the boundary NEAREST to g*E/parts -- checked
    against a brute-force search over all boundaries."""
    from geot_b200 import dist as gdist
    g = torch.Generator().manual_seed(11)
    for trial in range(200):
        S = int(torch.randint(1, 40, (1,), generator=g))
        deg = torch.randint(0, 6, (S,), generator=g)
        if trial % 3 == 0:
            deg[int(torch.randint(0, S, (1,), generator=g))] += int(torch.randint(20, 200, (1,), generator=g))   # a hub
        if int(deg.sum()) == 0:
            deg[0] = 1
        rowptr = torch.cat([torch.zeros(1, dtype=torch.int64), deg.cumsum(0)])
        E = int(rowptr[-1])
        for parts in (1, 2, 3, 4, 8):
            rows, edges = gdist.shard_bounds_from_rowptr(rowptr, parts)
            assert len(rows) == parts + 1 and rows[0] == 0 and rows[-1] == S and edges[0] == 0 and edges[-1] == E
            assert all(rows[i] <= rows[i + 1] for i in range(parts)) and all(edges[i] <= edges[i + 1] for i in range(parts))
            assert edges == [int(rowptr[r]) for r in rows]
            for gi in range(1, parts):
                target = (E // parts) * gi + ((E % parts) * gi) // parts
                best = int((rowptr - target).abs().min())
                assert abs(edges[gi] - target) == best, (trial, parts, gi, edges, target)
