"""Same-process A/B of tuning builds of the library (Makefile VARIANT=...): every library times the same workload
through the C ABI, interleaved ROUNDS times so that clock / power drift hits all of them alike.
  python scripts/tune_libs.py <workload> <lib[,lib...]> [src_blocks[,src_blocks...]]
lib = "default" or the VARIANT suffix (e.g. _mix1 -> geot_b200/lib/libgeot_b200_mix1.so); src_blocks as in
scripts/tune.py (1 = off, 0 = the library's suggestion).  Results are checked against the register path
(GEOT_B200_RING=0) of the default library."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from geot_b200 import abi, LIB_PATH

name = sys.argv[1] if len(sys.argv) > 1 else "reddit_gws"
libs = (sys.argv[2] if len(sys.argv) > 2 else "default").split(",")
blocks_list = [int(b) for b in (sys.argv[3] if len(sys.argv) > 3 else "1").split(",")]
ROUNDS, STEPS = 3, 10


def use(lib):
    os.environ.pop("GEOT_B200_LIB", None)
    if lib != "default":
        os.environ["GEOT_B200_LIB"] = os.path.join(os.path.dirname(LIB_PATH), "libgeot_b200%s.so" % lib)
    abi._lib = None
    return abi.lib()


wk = bench.build_workload(name, "cuda")
E, S, F, H = wk["E"], wk["S"], wk["F"], wk["H"]
w = wk["w"]
layout = abi.W_NONE if w is None else (abi.W_EDGE if w.dim() == 1 else abi.W_EDGE_HEAD)
use("default")
plan = abi.DevicePlan(wk["di"], S)
out = torch.empty([S] + list(wk["x"].shape[1:]), dtype=wk["dtype"], device="cuda")
os.environ["GEOT_B200_RING"] = "0"
_ws = abi.Workspace(E, F * H, wk["dtype"], "cuda")
ref = abi.segment_reduce(wk["x"], wk["si"], wk["di"], w, "sum", S=S, H=H, weight_layout=layout, plan=plan, workspace=_ws).clone()
del _ws, os.environ["GEOT_B200_RING"]

for nb0 in blocks_list:
    use("default")
    nb = nb0 if nb0 > 0 else abi.src_blocks_suggest(E, S, wk["N"], F * H * wk["esize"])
    blocks = abi.SrcBlocks(wk["si"], wk["di"], wk["N"], nb) if (nb > 1 and wk["si"] is not None) else None
    calls = nb if blocks is not None else 1
    ws = abi.Workspace(E, F * H, wk["dtype"], "cuda", src_blocks=blocks)
    res = {lib: [] for lib in libs}
    for rnd in range(ROUNDS):
        for lib in libs:
            use(lib)
            f = lambda: abi.segment_reduce(wk["x"], wk["si"], wk["di"], w, "sum", S=S, H=H, weight_layout=layout, plan=plan, out=out,
                                           workspace=ws, src_blocks=blocks)
            out.fill_(float("nan"))
            for _ in range(3): f()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(STEPS): f()
            e1.record(); torch.cuda.synchronize()
            err = float(((out.float() - ref.float()).abs() / ref.float().abs().clamp_min(1e-20)).max())
            res[lib].append((e0.elapsed_time(e1) / STEPS, err))
    for lib in libs:
        ms = [r[0] for r in res[lib]]
        print("%s lib=%s blocks=%d: step %s ms (best %.3f, %.0f GB/s logical)  maxrel_vs_ring0 %.1e" % (
            name, lib, calls, " / ".join("%.3f" % m for m in ms), min(ms), wk["bytes_logical"] / min(ms) / 1e6,
            max(r[1] for r in res[lib])), flush=True)
