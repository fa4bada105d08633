"""A/B of the L2 residency hint (geot_b200_l2_persist) on one workload: same kernel, src matrix with and without the
persisting access-policy window.  Usage: python scripts/l2_persist_ab.py [workload]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from geot_b200 import abi

name = sys.argv[1] if len(sys.argv) > 1 else "reddit_gws"
wk = bench.build_workload(name, "cuda")
E, S, F, H = wk["E"], wk["S"], wk["F"], wk["H"]
w = wk["w"]
layout = abi.W_NONE if w is None else (abi.W_EDGE if w.dim() == 1 else abi.W_EDGE_HEAD)
plan = abi.DevicePlan(wk["di"], S)
out = torch.empty([S] + list(wk["x"].shape[1:]), dtype=wk["dtype"], device="cuda")
ws = abi.Workspace(E, F * H, wk["dtype"], "cuda")
side = torch.cuda.Stream()      # a created stream: the attribute is per stream
f = lambda: abi.segment_reduce(wk["x"], wk["si"], wk["di"], w, "sum", S=S, H=H, weight_layout=layout, plan=plan, out=out, workspace=ws)


def timed(tag):
    with torch.cuda.stream(side):
        for _ in range(3): f()
        abi.profile_enable(10)
        side.synchronize()
        for _ in range(10): f()
        side.synchronize()
        km = abi.profile_read(10); abi.profile_enable(0)
    ms = sum(km) / len(km)
    print("%s %s: main kernel %.3f ms (%.0f GB/s logical)" % (name, tag, ms, wk["bytes_logical"] / ms / 1e6), flush=True)
    return out.clone()


a = timed("no hint")
with torch.cuda.stream(side):
    try:
        print("l2_persist: window %d B, carve-out %d B" % abi.l2_persist(wk["x"]), flush=True)
    except abi.AbiError as e:
        print("l2_persist failed:", e, flush=True)
b = timed("src persisting")
with torch.cuda.stream(side):
    abi.l2_persist_reset()
c = timed("after reset")
print("bit-identical results:", torch.equal(a, b) and torch.equal(a, c))
