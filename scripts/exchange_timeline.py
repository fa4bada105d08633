"""Where an N-GPU step of a sharded gather op goes (torchrun, one rank per GPU): every piece of the two-bucket exchange
timed alone with CUDA events (max over ranks), the host's enqueue cost per step, and the same step replayed from a
CUDA graph.
    torchrun --nproc-per-node N scripts/exchange_timeline.py [workload] [exchange]
Bench support, not product."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from geot_b200 import abi

name = sys.argv[1] if len(sys.argv) > 1 else "reddit_gws"
exchange = sys.argv[2] if len(sys.argv) > 2 else "push"
world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
wk = bench.build_workload(name, dev)
r = bench.Runner(wk, world, rank, dev, exchange)
bg = r.bg
x, w, out = r.src_operand(), r.w, r.out


def timed(fn, n=20, warm=3):
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
    lo = t.clone()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    return t.item(), lo.item()


def say(what, v):
    if rank == 0:
        print("%s %s N=%d %-28s max %.4f ms  min %.4f ms" % (name, exchange, world, what, v[0], v[1]), flush=True)


say("full step (eager)", timed(r.step))
# host enqueue cost: wall time to enqueue 20 steps with the GPU drained before and not waited for after
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    r.step()
host_ms = (time.perf_counter() - t0) / 20 * 1e3
torch.cuda.synchronize()
t = torch.tensor([host_ms], device=dev, dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("%s %s N=%d host enqueue per step: %.4f ms (max over ranks)" % (name, exchange, world, t.item()), flush=True)

if bg is not None:
    buf, hdl = bg._buffer(list(x.shape[1:]), x.dtype, x.device)
    say("exchange alone", timed(lambda: bg._exchange(x, buf, hdl)))
    if bg.transport == "push":
        say("two barriers alone", timed(lambda: (bg._barrier(hdl, 0), bg._barrier(hdl, 1))))
        nd = bg.needed
        say("push kernel alone", timed(lambda: abi.push_rows(x, nd.send_rows, nd.dest_peer, nd.dest_row, hdl.bases.data_ptr())))
        if rank == 0:
            print("   rows pushed by rank 0: %d (%.1f MB), received %d" % (nd.send_rows.numel(), nd.send_rows.numel() * x[0].numel() * x.element_size() / 1e6,
                                                                    nd.recv_offsets[-1]), flush=True)
    if bg.passes == 2:
        say("local bucket alone", timed(lambda: bg._reduce_bucket(0, x, w, out, "sum", accumulate=False)))
        for ph in range(bg.phases):
            say("remote bucket %d alone" % ph, timed(lambda: bg._reduce_bucket(1 + ph, buf, w, out, "sum", accumulate=True)))
        if rank == 0:
            print("   rounds %d, bucket edge bounds %s, main-kernel launches per step %d" % (bg.phases, bg.buckets.bounds, bg.main_launches()), flush=True)
    else:
        say("single reduction alone", timed(lambda: bg._reduce_bucket(0, buf, w, out, "sum", accumulate=False)))

# the same step replayed from a CUDA graph (the step's host cost and launch gaps removed)
try:
    r.step(); torch.cuda.synchronize()
    ref = out.clone()
    dist.barrier()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        r.step()
    out.zero_()
    g.replay(); torch.cuda.synchronize()
    same = bool(torch.equal(out, ref))
    say("full step (graph replay)", timed(g.replay))
    ok = torch.tensor([0.0 if same else 1.0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("   graph replay output identical to the eager step on every rank: %s" % (ok.item() == 0), flush=True)
except Exception as ex:
    print("rank %d: graph capture failed: %r" % (rank, ex), flush=True)
dist.barrier()
dist.destroy_process_group()
