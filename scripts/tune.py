"""Times one workload through the C ABI for the library named by GEOT_B200_LIB over chunk sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from geot_b200 import abi

name = sys.argv[1] if len(sys.argv) > 1 else "reddit_gws"
chunks = [int(c) for c in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["0"])]
wk = bench.build_workload(name, "cuda")
E, S, F, H = wk["E"], wk["S"], wk["F"], wk["H"]
w = wk["w"]
layout = abi.W_NONE if w is None else (abi.W_EDGE if w.dim() == 1 else abi.W_EDGE_HEAD)
plan = abi.DevicePlan(wk["di"], S)
out = torch.empty([S] + list(wk["x"].shape[1:]), dtype=wk["dtype"], device="cuda")
# reference result: the register path (GEOT_B200_RING=0) of the same library
_ring = os.environ.get("GEOT_B200_RING")
os.environ["GEOT_B200_RING"] = "0"
_ws = abi.Workspace(E, F * H, wk["dtype"], "cuda")
ref = abi.segment_reduce(wk["x"], wk["si"], wk["di"], w, "sum", S=S, H=H, weight_layout=layout, plan=plan, workspace=_ws).clone()
del _ws
if _ring is None:
    del os.environ["GEOT_B200_RING"]
else:
    os.environ["GEOT_B200_RING"] = _ring
nb = int(os.environ.get("GEOT_B200_SRC_BLOCKS", "1"))        # n > 1: src-row blocking with n blocks; 0: the library's suggestion
if nb == 0:
    nb = abi.src_blocks_suggest(E, S, wk["N"], F * H * wk["esize"])
blocks = abi.SrcBlocks(wk["si"], wk["di"], wk["N"], nb) if (nb > 1 and wk["si"] is not None) else None
calls = nb if blocks is not None else 1
for c in chunks:
    os.environ["GEOT_B200_CHUNK"] = str(c)
    ws = abi.Workspace(E, F * H, wk["dtype"], "cuda", src_blocks=blocks)
    f = lambda: abi.segment_reduce(wk["x"], wk["si"], wk["di"], w, "sum", S=S, H=H, weight_layout=layout, plan=plan, out=out, workspace=ws,
                                   src_blocks=blocks)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    abi.profile_enable(10 * calls)              # the kernel's own time from a second, instrumented loop
    for _ in range(10): f()
    torch.cuda.synchronize()
    km = abi.profile_read(10 * calls); abi.profile_enable(0)
    ms = e0.elapsed_time(e1) / 10
    err = float(((out.float() - ref.float()).abs() / ref.float().abs().clamp_min(1e-20)).max())
    kms = sum(km) / 10
    print("%s lib=%s chunk=%d blocks=%d: step %.3f ms (%.0f GB/s logical, %.2f Gedge/s)  main kernel(s) %.3f ms  fixup+gaps %.3f ms  maxrel_vs_ring0 %.1e" % (
        name, os.path.basename(os.environ.get("GEOT_B200_LIB", "default")), c, calls, ms, wk["bytes_logical"] / ms / 1e6, E / ms / 1e6,
        kms, ms - kms, err), flush=True)
