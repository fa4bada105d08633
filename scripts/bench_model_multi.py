"""BASELINE configs[4] at N GPUs: 3-layer GCN / GraphSAGE forward on the synthetic proteins-shape graph
(132,534 nodes, 39.5 M edges, in = hidden = out = 256, fp32), node rows sharded by dst row over the GPUs of one box,
through every form of the src-row exchange (one all-gather per layer / bucketed all-gather / bucketed push).  One process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 \
        scripts/bench_model_multi.py > gpurun_out/<tag>/model_nN.jsonl

Timing: CUDA events on the current stream between barriers, max over ranks, best / median of 10 after 3 warm-ups.
Parity: every rank checks its rows against the single-GPU forward of the same stack (computed once on that rank).
Bench support, not product."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import workloads as wl  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import geot_b200  # noqa: F401
    from geot_b200 import dist as gdist, gnn

    g = wl.power_law_graph("proteins", dev)
    N, E, si, di = g.num_nodes, g.num_edges, g.src_index, g.dst_index
    torch.manual_seed(0)                                   # same weights and features on every rank
    x = wl.features(N, 256, torch.float32, dev)
    norm = gnn.gcn_norm(si, di, N)
    models = {"gcn": gnn.GCN(256, 256, 3).to(dev), "graphsage": gnn.GraphSAGE(256, 256, 3).to(dev)}

    def timed(fn, warmup=3, iters=10):
        for _ in range(warmup):
            fn()
        ts = []
        for _ in range(iters):
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(t.item())
        ts.sort()
        return {"best": round(ts[0], 4), "median": round(ts[len(ts) // 2], 4)}

    with torch.no_grad():
        for name, model in models.items():
            w = norm if name == "gcn" else None
            full = model(x, si, di, norm) if name == "gcn" else model(x, si, di)
            single = timed((lambda: model(x, si, di, norm)) if name == "gcn" else (lambda: model(x, si, di)))
            shard = gdist.shard_graph(si, di, w, rank, world)
            rb = shard.row_bounds
            x_local = x[rb[rank]:rb[rank + 1]].contiguous()
            exp = full[rb[rank]:rb[rank + 1]]
            rec = {"model": "3-layer %s forward, proteins shape, 256-256-256-256, fp32" % name, "N": N, "E": E, "n_gpus": world,
                   "shard_imbalance": round(shard.imbalance, 4), "single_gpu_ms": single}
            for form in ("allgather", "bucket", "push"):
                try:
                    gather = None if form == "allgather" else gdist.BucketedGather(shard, transport="push" if form == "push" else "allgather")
                    got = gnn.forward_sharded(model, x_local, shard, gather=gather)
                    err = float(((got - exp).abs() / exp.abs().clamp_min(1e-3)).max()) if got.numel() else 0.0
                    rec[form + "_ms"] = timed(lambda: gnn.forward_sharded(model, x_local, shard, gather=gather))
                    rec[form + "_max_rel_err_vs_single_gpu"] = err
                    if gather is not None:
                        rec[form + "_rows_received_rank0"] = gather.exchanged_rows()[0]
                except Exception as ex:          # (a failure inside a collective would hang instead: bound the run with timeout)
                    rec[form + "_error"] = repr(ex)[:200]
            if rank == 0:
                print(json.dumps(rec), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
