"""One process per GPU (torchrun): the sharded step of a gather workload timed over the knobs of the push exchange --
rounds (BucketedGather phases) x grid of the overlapped push kernel -- with the workload built once.
    torchrun --nproc-per-node N scripts/exchange_sweep.py [workload] [rounds,rounds,...] [ctas,ctas,...]
Bench support, not product."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "reddit_gws"
# rounds: counts ("1,2,3": equal rounds) or explicit last steps of the rounds joined by '+' ("1+3,2+4": rounds ending at
# exchange steps 1, 3, N-1 and 2, 4, N-1)
rounds = (sys.argv[2] if len(sys.argv) > 2 else "1,2,3").split(",")
ctas = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "296,592").split(",")]
world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
wk = bench.build_workload(name, dev)


def timed(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    bench.barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    bench.barrier(world)
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


ref = None
for k in rounds:
    os.environ.pop("GEOT_B200_EXCHANGE_STEPS", None)
    if "+" in k or k.startswith("s"):
        os.environ["GEOT_B200_EXCHANGE_STEPS"] = k.lstrip("s").replace("+", ",")
        os.environ["GEOT_B200_EXCHANGE_PHASES"] = "0"
    else:
        os.environ["GEOT_B200_EXCHANGE_PHASES"] = k
    r = bench.Runner(wk, world, rank, dev, "push")
    for c in ctas:
        os.environ["GEOT_B200_PUSH_CTAS"] = str(c)
        ms = timed(r.step)
        torch.cuda.synchronize()
        if ref is None:
            ref = r.out.clone()
        same = torch.tensor([0.0 if torch.allclose(r.out, ref, rtol=1e-5, atol=0) else 1.0], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MAX)
        if rank == 0:
            print("%s push N=%d rounds %d (asked %s; last steps %s) push ctas %d: %.4f ms/step, launches/step %d, equal to the first form within 1e-5: %s"
                  % (name, world, r.phases, k, r.bg.phase_steps, c, ms, r.calls_per_step, same.item() == 0), flush=True)
    del r
    torch.cuda.empty_cache()
dist.barrier()
dist.destroy_process_group()
