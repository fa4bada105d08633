"""Times the register-path kernels of the 2-byte types (max / min, and sum with the ring switched off) on a
Reddit-shape gather: the comparison behind the U0 = 8 (spilling) vs U0 = 4 (spill-free) decision."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import workloads as wl
from geot_b200 import abi

g = wl.power_law_graph("reddit", "cuda", 0.25)
E, N = g.num_edges, g.num_nodes
S = int(g.dst_index[-1]) + 1
plan = abi.DevicePlan(g.dst_index, S)
for dtype in (torch.bfloat16, torch.float16):
    for F in (64, 128, 256):
        x = wl.features(N, F, dtype, "cuda")
        ws = abi.Workspace(E, F, dtype, "cuda")
        out = torch.empty(S, F, device="cuda", dtype=dtype)
        for red, ring in (("max", None), ("min", None), ("sum", "0")):
            if ring is not None:
                os.environ["GEOT_B200_RING"] = ring
            f = lambda: abi.segment_reduce(x, g.src_index, g.dst_index, None, red, S=S, plan=plan, out=out, workspace=ws)
            for _ in range(3): f()
            abi.profile_enable(10)
            for _ in range(10): f()
            torch.cuda.synchronize()
            km = abi.profile_read(10); abi.profile_enable(0)
            os.environ.pop("GEOT_B200_RING", None)
            print("%s F=%d %s%s: main kernel %.3f ms" % (str(dtype)[6:], F, red, " (ring off)" if ring else "", sum(km) / len(km)), flush=True)
