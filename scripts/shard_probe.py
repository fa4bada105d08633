"""Times the dst-row shards of a workload one by one on ONE GPU (what each rank of an N-GPU run would do with the src
matrix pre-replicated): finds shards whose time is out of line with their edge count (hub rows, short rows).
    python scripts/shard_probe.py products_gs64 8
Bench support, not product."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from geot_b200 import abi, dist as gdist

name = sys.argv[1] if len(sys.argv) > 1 else "products_gs64"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
wk = bench.build_workload(name, "cuda")
x, F, H = wk["x"], wk["F"], wk["H"]
for rank in range(world):
    sh = gdist.shard_graph(wk["si"], wk["di"], wk["w"], rank, world)
    S, E = sh.num_local_rows, sh.num_local_edges
    plan = abi.DevicePlan(sh.dst_index, S)
    ws = abi.Workspace(E, F * H, wk["dtype"], "cuda")
    out = torch.empty([S] + list(x.shape[1:]), dtype=wk["dtype"], device="cuda")
    f = lambda: abi.segment_reduce(x, sh.src_index, sh.dst_index, sh.weight, "sum", S=S, H=H, plan=plan, out=out, workspace=ws)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    abi.profile_enable(10)
    for _ in range(10): f()
    torch.cuda.synchronize()
    km = abi.profile_read(10); abi.profile_enable(0)
    deg = torch.bincount(sh.dst_index, minlength=S)
    print("%s shard %d/%d: rows %d edges %d max_degree %d has_gaps %d: step %.3f ms main kernel %.3f ms" % (
        name, rank, world, S, E, int(deg.max()), plan.c.has_gaps, e0.elapsed_time(e1) / 10, sum(km) / len(km)), flush=True)
