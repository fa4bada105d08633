"""L2 capacity probe: gather_weight_scatter on the Reddit-shape edge list with F = argv[1] columns, so that the src
matrix is 30 / 60 / 119 MB; run under ncu to read the L2 hit rate per size (does a read-shared matrix get the whole
126 MB L2, or one copy per partition?)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import workloads as wl
from geot_b200 import abi

F = int(sys.argv[1]) if len(sys.argv) > 1 else 128
g = wl.power_law_graph("reddit", "cuda", 1.0)
E, N = g.num_edges, g.num_nodes
x = wl.features(N, F, torch.float32, "cuda")
w = wl.edge_weights(E, None, torch.float32, "cuda")
S = int(g.dst_index[-1]) + 1
plan = abi.DevicePlan(g.dst_index, S)
ws = abi.Workspace(E, F, torch.float32, "cuda")
out = torch.empty(S, F, device="cuda")
f = lambda: abi.segment_reduce(x, g.src_index, g.dst_index, w, "sum", S=S, plan=plan, out=out, workspace=ws)
for _ in range(3): f()
abi.profile_enable(5)
for _ in range(5): f()
torch.cuda.synchronize()
km = abi.profile_read(5)
print("F=%d src %.1f MB: main kernel %.3f ms" % (F, N * F * 4 / 1e6, sum(km) / len(km)))
