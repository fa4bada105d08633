"""world_size-2 NCCL test of the dst-sharded gather ops on two GPUs (skipped on a one-GPU box): both forms of the
src-row exchange (one all-gather / staggered send-recv steps overlapped with per-owner buckets) against the CPU
oracle on the unsharded graph, and bit-reproducibility of the pipelined form."""
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
    N, E = 5000, 400_000
    deg_w = torch.rand(N, generator=g) ** 3
    dst = torch.multinomial(deg_w, E, replacement=True, generator=g).sort().values
    dst[-1] = N - 1
    src_index = torch.randint(0, N, (E,), generator=g)
    weight = torch.rand(E, generator=g) + 0.25
    for F, dtype in [(128, torch.float32), (64, torch.float32), (256, torch.bfloat16)]:
        x = torch.rand(N, F, generator=g).to(dtype)
        shard = gdist.shard_graph(src_index.to(dev), dst.to(dev), weight.to(dev).to(dtype), rank, world)
        rb = shard.row_bounds
        x_local = x[rb[rank]:rb[rank + 1]].to(dev)
        pg = gdist.PipelinedGather(shard)
        tol = 1e-5 if dtype == torch.float32 else 2e-2
        for reduce in ("sum", "mean"):
            for weighted in (True, False):
                if weighted:
                    full = oracle.gather_weight_scatter(src_index, dst, weight.to(dtype), x, reduce, acc64=True)
                else:
                    full = oracle.gather_scatter(src_index, dst, x, reduce, acc64=True)
                exp = full[rb[rank]:rb[rank + 1]].double()
                sh = shard if weighted else gdist.GraphShard(rank, world, shard.row_bounds, shard.edge_bounds, shard.src_index,
                                                            shard.dst_index, None)
                a = gdist.sharded_gather_scatter(sh, x_local, reduce).cpu().double()
                x_full = torch.full((N, F), float("nan"), device=dev, dtype=dtype)
                pg.local_rows(x_full).copy_(x_local)
                b1 = pg(x_full, shard.weight if weighted else None, reduce).clone()
                assert torch.equal(x_full.cpu(), x), "exchange did not rebuild the replica"
                b2 = pg(x_full, shard.weight if weighted else None, reduce)
                assert torch.equal(b1, b2), "pipelined result is not bit-reproducible"
                for name, got in (("allgather", a), ("pipeline", b1.cpu().double())):
                    bad = (got - exp).abs() > tol * exp.abs().clamp_min(1e-3 if dtype != torch.float32 else 1e-30)
                    assert not bad.any(), (name, F, dtype, reduce, weighted, int(bad.sum()))
    dist.barrier()
    q.put(rank)
    dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_exchange_forms_nccl():
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
