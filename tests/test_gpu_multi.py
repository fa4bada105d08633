"""Multi-GPU NCCL test of the dst-sharded gather ops (skipped on a one-GPU box; world = min(4, device count)): one
all-gather + one reduction, and the overlapped two-bucket exchange (``dist.BucketedGather``) with both transports --
NCCL all-gather, and the needed-rows push over symmetric memory (``geot_b200_push_rows``) -- against the CPU oracle
on the unsharded graph, on a dense and on a sparsely referencing graph, with bit-reproducibility; per-head weights
(``mh_spmm``); then the sharded 3-layer GCN / GraphSAGE forward (hidden width != input width) through every form."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import oracle
    import geot_b200  # noqa: F401
    from geot_b200 import dist as gdist, gnn
    g = torch.Generator().manual_seed(0)          # same graph on every rank
    N, E = 6000, 400_000
    deg_w = torch.rand(N, generator=g) ** 3
    deg_w[50:90] = 0                               # rows without edges
    dst = torch.multinomial(deg_w, E, replacement=True, generator=g).sort().values
    dst[-1] = N - 1
    dense = torch.randint(0, N, (E,), generator=g)
    sparse = (torch.randint(0, 300, (E,), generator=g) * 20) % N      # 300 distinct src rows
    weight = torch.rand(E, generator=g) + 0.25
    for gi, src_index in enumerate((dense, sparse)):
        for F, dtype in [(128, torch.float32), (64, torch.float32), (256, torch.bfloat16), (6, torch.float32), (3, torch.bfloat16)]:
            x = torch.rand(N, F, generator=g).to(dtype)
            shard = gdist.shard_graph(src_index.to(dev), dst.to(dev), weight.to(dev).to(dtype), rank, world)
            rb = shard.row_bounds
            x_local = x[rb[rank]:rb[rank + 1]].to(dev)
            forms = {"allgather": gdist.BucketedGather(shard, transport="allgather", passes=2),
                     "push": gdist.BucketedGather(shard, transport="push", passes=2, phases=1),
                     "push, exchange in rounds": gdist.BucketedGather(shard, transport="push", passes=2, phases=world - 1),
                     "allgather, one pass": gdist.BucketedGather(shard, transport="allgather", passes=1),
                     "push, one pass": gdist.BucketedGather(shard, transport="push", passes=1)}
            assert gdist.BucketedGather(shard, transport="allgather").passes in (1, 2)
            assert forms["push, exchange in rounds"].phases == max(world - 1, 1) if world > 2 else 1
            got_rows, full_rows = forms["push"].exchanged_rows()
            assert forms["push, one pass"].exchanged_rows() == (got_rows, full_rows)
            assert got_rows <= full_rows and (gi == 0 or got_rows <= 300)
            tol = 1e-5 if dtype == torch.float32 else 2e-2
            for reduce in ("sum", "mean"):
                for weighted in (True, False):
                    if weighted:
                        full = oracle.gather_weight_scatter(src_index, dst, weight.to(dtype), x, reduce, acc64=True)
                    else:
                        full = oracle.gather_scatter(src_index, dst, x, reduce, acc64=True)
                    exp = full[rb[rank]:rb[rank + 1]].double()
                    sh = shard if weighted else gdist.GraphShard(rank, world, shard.row_bounds, shard.edge_bounds,
                                                                shard.src_index, shard.dst_index, None)
                    results = {"one all-gather": gdist.sharded_gather_scatter(sh, x_local, reduce).cpu().double()}
                    for name, bg in forms.items():
                        w = shard.weight if weighted else None
                        out = torch.full((x_local.shape[0], F), 7.0, dtype=dtype, device=dev)    # dirty output buffer
                        b1 = bg(x_local, w, reduce, out=out).clone()
                        b2 = bg(x_local, w, reduce)
                        assert torch.equal(b1, b2), "%s result is not bit-reproducible" % name
                        results[name] = b1.cpu().double()
                    for name, got in results.items():
                        bad = (got - exp).abs() > tol * exp.abs().clamp_min(1e-3 if dtype != torch.float32 else 1e-30)
                        assert not bad.any(), (name, gi, F, dtype, reduce, weighted, int(bad.sum()))
    # max over the shards (no partial sums to add: one all-gather + one reduction), bit-exact
    x = torch.rand(N, 32, generator=g)
    shard = gdist.shard_graph(dense.to(dev), dst.to(dev), None, rank, world)
    rb = shard.row_bounds
    got = gdist.sharded_gather_scatter(shard, x[rb[rank]:rb[rank + 1]].to(dev), "max").cpu()
    assert torch.equal(got, oracle.gather_scatter(dense, dst, x, "max")[rb[rank]:rb[rank + 1]])
    # multi-head rows [N, H, F] with per-head weights [E, H] (mh_spmm), bf16
    Hh = 4
    xh = torch.rand(N, Hh, 32, generator=g).bfloat16()
    wh = torch.rand(E, Hh, generator=g).bfloat16()
    sh_h = gdist.shard_graph(dense.to(dev), dst.to(dev), wh.to(dev), rank, world)
    exp_h = oracle.mh_spmm(dense, dst, wh, xh)[rb[rank]:rb[rank + 1]].float()
    for transport in ("allgather", "push"):
        got = gdist.BucketedGather(sh_h, transport=transport).aggregate(xh[rb[rank]:rb[rank + 1]].to(dev), sh_h.weight, "sum")
        assert torch.allclose(got.cpu().float(), exp_h, rtol=2e-2, atol=2e-2), transport
    # 3-layer GCN / GraphSAGE forward on the shards: widths 48 -> 96 -> 96 (the scratch and replica buffers are per
    # row shape), against the single-GPU forward of the same stack
    torch.manual_seed(7)
    F = 48
    x = torch.rand(N, F, generator=g)
    gcn, sage = gnn.GCN(F, 96, 3).to(dev), gnn.GraphSAGE(F, 96, 3).to(dev)
    si_d, di_d = dense.to(dev), dst.to(dev)
    norm = gnn.gcn_norm(si_d, di_d, N, weight.to(dev))
    sh_gcn = gdist.shard_graph(si_d, di_d, norm, rank, world)
    rb = sh_gcn.row_bounds
    sh_sage = gdist.shard_graph(si_d, di_d, None, rank, world, row_bounds=rb, edge_bounds=sh_gcn.edge_bounds)
    x_local = x[rb[rank]:rb[rank + 1]].to(dev)
    with torch.no_grad():
        exp_gcn = gcn(x.to(dev), si_d, di_d, norm)[rb[rank]:rb[rank + 1]]
        exp_sage = sage(x.to(dev), si_d, di_d)[rb[rank]:rb[rank + 1]]
        for form in (None, "allgather", "push"):
            mk = lambda sh: None if form is None else gdist.BucketedGather(sh, transport=form)
            got = gnn.forward_sharded(gcn, x_local, sh_gcn, gather=mk(sh_gcn))
            assert torch.allclose(got, exp_gcn, rtol=1e-4, atol=1e-5), ("gcn", form)
            got = gnn.forward_sharded(sage, x_local, sh_sage, gather=mk(sh_sage))
            assert torch.allclose(got, exp_sage, rtol=1e-4, atol=1e-3), ("sage", form)
    dist.barrier()
    q.put(rank)
    dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_multi_gpu_exchange_forms_nccl():
    world = min(4, torch.cuda.device_count())
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=900)
    codes = [p.exitcode for p in procs]
    for p in procs:                     # a rank stuck in a collective must not outlive the test
        if p.is_alive():
            p.kill()
    assert all(c == 0 for c in codes), codes
    assert sorted(q.get(timeout=5) for _ in range(world)) == list(range(world))
