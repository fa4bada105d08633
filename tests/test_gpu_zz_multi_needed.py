"""world_size-2 test of the needed-rows exchanges on two GPUs (skipped on a one-GPU box): the NCCL form
(``PipelinedGather(needed_only=True)``: bit-equal to the full pipelined exchange -- same buckets, same summation
order) and the peer-memory form (``PeerPushGather``: symmetric memory + ``geot_b200_push_rows``), against the CPU
oracle on the unsharded graph, on a dense and on a sparsely referencing graph; then the sharded 3-layer GCN /
GraphSAGE forward through every exchange form."""
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
    from geot_b200 import dist as gdist
    g = torch.Generator().manual_seed(0)          # same graph on every rank
    N, E = 6000, 300_000
    deg_w = torch.rand(N, generator=g) ** 3
    dst = torch.multinomial(deg_w, E, replacement=True, generator=g).sort().values
    dst[-1] = N - 1
    dense = torch.randint(0, N, (E,), generator=g)
    sparse = (torch.randint(0, 300, (E,), generator=g) * 20) % N      # 300 distinct src rows
    weight = torch.rand(E, generator=g) + 0.25
    for src_index in (dense, sparse):
        for F, dtype in [(128, torch.float32), (64, torch.float32), (256, torch.bfloat16), (6, torch.float32)]:
            x = torch.rand(N, F, generator=g).to(dtype)
            shard = gdist.shard_graph(src_index.to(dev), dst.to(dev), weight.to(dev).to(dtype), rank, world)
            rb = shard.row_bounds
            x_local = x[rb[rank]:rb[rank + 1]].to(dev)
            pf = gdist.PipelinedGather(shard)
            pn = gdist.PipelinedGather(shard, needed_only=True)
            pp = gdist.PeerPushGather(shard)          # symmetric memory + geot_b200_push_rows (no NCCL on the data path)
            assert pp.exchanged_rows() == pn.exchanged_rows()
            got_rows, full_rows = pn.exchanged_rows()
            assert got_rows <= full_rows and (src_index is dense or got_rows <= 300)
            tol = 1e-5 if dtype == torch.float32 else 2e-2
            for reduce in ("sum", "mean"):
                for weighted in (True, False):
                    w = shard.weight if weighted else None
                    if weighted:
                        full = oracle.gather_weight_scatter(src_index, dst, weight.to(dtype), x, reduce, acc64=True)
                    else:
                        full = oracle.gather_scatter(src_index, dst, x, reduce, acc64=True)
                    exp = full[rb[rank]:rb[rank + 1]].double()
                    x_full = torch.full((N, F), float("nan"), device=dev, dtype=dtype)
                    pf.local_rows(x_full).copy_(x_local)
                    a = pf(x_full, w, reduce).clone()
                    b1 = pn(x_local, w, reduce).clone()
                    b2 = pn(x_local, w, reduce)
                    assert torch.equal(b1, b2), "needed-rows result is not bit-reproducible"
                    assert torch.equal(a, b1), "needed-rows and full exchange differ (same buckets, same order)"
                    bad = (b1.cpu().double() - exp).abs() > tol * exp.abs().clamp_min(1e-3 if dtype != torch.float32 else 1e-30)
                    assert not bad.any(), (F, dtype, reduce, weighted, int(bad.sum()))
                    c1 = pp(x_local, w, reduce).clone()
                    c2 = pp(x_local, w, reduce)
                    assert torch.equal(c1, c2), "peer-push result is not bit-reproducible"
                    bad = (c1.cpu().double() - exp).abs() > tol * exp.abs().clamp_min(1e-3 if dtype != torch.float32 else 1e-30)
                    assert not bad.any(), ("push", F, dtype, reduce, weighted, int(bad.sum()))
    # multi-head rows with per-head weights (mh_spmm, BASELINE configs[3] shape class) through the overlapped forms
    Hh, Fh = 8, 32
    xh = torch.rand(N, Hh, Fh, generator=g).bfloat16()
    wh = torch.rand(E, Hh, generator=g).bfloat16()
    sh_h = gdist.shard_graph(dense.to(dev), dst.to(dev), wh.to(dev), rank, world)
    rbh = sh_h.row_bounds
    exp_h = oracle.mh_spmm(dense, dst, wh, xh)[rbh[rank]:rbh[rank + 1]].double()
    xh_local = xh[rbh[rank]:rbh[rank + 1]].to(dev)
    for obj in (gdist.PipelinedGather(sh_h), gdist.PipelinedGather(sh_h, needed_only=True), gdist.PeerPushGather(sh_h)):
        got = obj.aggregate(xh_local, sh_h.weight, "sum").cpu().double()
        assert not ((got - exp_h).abs() > 2e-2 * exp_h.abs().clamp_min(1e-3)).any(), type(obj).__name__

    # 3-layer GCN / GraphSAGE forward on the shard (BASELINE configs[4] at N > 1): all three exchange forms against
    # the single-GPU forward of the same stack on the unsharded graph
    from geot_b200 import gnn
    torch.manual_seed(7)
    F = 64
    x = torch.rand(N, F, generator=g)
    gcn, sage = gnn.GCN(F, 64, 3).to(dev), gnn.GraphSAGE(F, 64, 3).to(dev)
    si_d, di_d = dense.to(dev), dst.to(dev)
    norm = gnn.gcn_norm(si_d, di_d, N, weight.to(dev))
    sh_gcn = gdist.shard_graph(si_d, di_d, norm, rank, world)
    rb = sh_gcn.row_bounds
    sh_sage = gdist.shard_graph(si_d, di_d, None, rank, world, row_bounds=rb, edge_bounds=sh_gcn.edge_bounds)
    x_local = x[rb[rank]:rb[rank + 1]].to(dev)
    with torch.no_grad():
        exp_gcn = gcn(x.to(dev), si_d, di_d, norm)[rb[rank]:rb[rank + 1]]
        exp_sage = sage(x.to(dev), si_d, di_d)[rb[rank]:rb[rank + 1]]
        for form in ("allgather", "pipeline", "needed", "push"):
            mk = lambda sh: (None if form == "allgather" else gdist.PeerPushGather(sh) if form == "push"
                             else gdist.PipelinedGather(sh, needed_only=(form == "needed")))
            got = gnn.forward_sharded(gcn, x_local, sh_gcn, gather=mk(sh_gcn))
            assert torch.allclose(got, exp_gcn, rtol=1e-4, atol=1e-5), ("gcn", form)
            got = gnn.forward_sharded(sage, x_local, sh_sage, gather=mk(sh_sage))
            assert torch.allclose(got, exp_sage, rtol=1e-4, atol=1e-3), ("sage", form)
    dist.barrier()
    q.put(rank)
    dist.destroy_process_group()


@pytest.mark.skipif(os.environ.get("GEOT_B200_TEST_EXPERIMENTS") != "1",
                    reason="needed-rows / peer-push exchanges were built after the round's GPU budget was spent and have never "
                           "run on hardware: enable with GEOT_B200_TEST_EXPERIMENTS=1 (scripts/gpu_r02_multi.sh does)")
@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_needed_rows_nccl():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
    codes = [p.exitcode for p in procs]
    for p in procs:                     # a rank stuck in a collective must not outlive the test
        if p.is_alive():
            p.kill()
    assert all(c == 0 for c in codes), codes
    assert sorted(q.get(timeout=5) for _ in range(world)) == [0, 1]
