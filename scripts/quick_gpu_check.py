"""First-light GPU check: parity vs torch on device for a spread of shapes + a rough timing."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import geot_b200
from geot_b200 import abi

torch.manual_seed(0)
dev = "cuda"

def ref(si, di, w, src, red, S):
    x = src if si is None else src.index_select(0, si)
    if w is not None:
        x = x * (w.unsqueeze(-1) if w.dim() < x.dim() else w)
    out = torch.zeros([S] + list(src.shape[1:]), dtype=x.dtype, device=dev)
    if red == "sum":
        return out.index_add_(0, di, x)
    idx = di.view([-1] + [1] * (x.dim() - 1)).expand_as(x)
    return out.scatter_reduce_(0, idx, x, {"mean": "mean", "max": "amax", "min": "amin", "prod": "prod"}[red], include_self=False)

bad = 0
for (E, N, F) in [(1000, 100, 32), (5000, 1000, 128), (100000, 3000, 64), (50000, 50, 128), (20000, 4000, 7), (30000, 2000, 1), (40000, 100, 256), (40000, 5000, 512), (30000, 300, 48), (30000, 300, 1000)]:
    for red in ["sum", "mean", "max", "min"]:
        si = torch.randint(0, N, (E,), device=dev)
        di = torch.randint(0, N, (E,), device=dev).sort().values
        w = torch.rand(E, device=dev)
        src = torch.rand(N, F, device=dev)
        S = int(di[-1]) + 1
        for name, args in [("gws", (si, di, w)), ("gs", (si, di, None))]:
            got = abi.segment_reduce(src, args[0], args[1], args[2], red, S=S)
            exp = ref(args[0], args[1], args[2], src, red, S)
            err = ((got - exp).abs() / exp.abs().clamp_min(1e-6)).max().item()
            tol = 0 if red in ("max", "min") else 2e-5
            ok = err <= tol
            bad += (not ok)
            if not ok:
                print("MISMATCH", name, E, N, F, red, err)
        srcE = torch.rand(E, F, device=dev)
        got = geot_b200.index_scatter(0, srcE, di, red)
        exp = ref(None, di, None, srcE, red, S)
        err = ((got - exp).abs() / exp.abs().clamp_min(1e-6)).max().item()
        if err > (0 if red in ("max", "min") else 2e-5):
            bad += 1; print("MISMATCH index_scatter", E, N, F, red, err)
print("mismatches:", bad)

# mh_spmm bf16
E, N, H, F = 100000, 5000, 8, 32
si = torch.randint(0, N, (E,), device=dev); di = torch.randint(0, N, (E,), device=dev).sort().values
w = torch.rand(E, H, device=dev).bfloat16(); src = torch.rand(N, H, F, device=dev).bfloat16()
got = geot_b200.mh_spmm(si, di, w, src)
exp = ref(si, di, w.float().unsqueeze(-1), src.float(), "sum", int(di[-1]) + 1).bfloat16()
print("mh_spmm bf16 max rel err", ((got.float() - exp.float()).abs() / exp.float().abs().clamp_min(1e-3)).max().item())
got2 = geot_b200.mh_spmm_transposed(si, di, w, src)
print("mh_spmm_transposed equal:", torch.equal(got, got2))

# rough timing: Reddit-like scaled down (E=20M, N=233K, F=128)
for (E, N, F) in [(20_000_000, 232_965, 128), (1_000_000, 50_000, 64)]:
    si = torch.randint(0, N, (E,), device=dev); di = torch.randint(0, N, (E,), device=dev).sort().values
    w = torch.rand(E, device=dev); src = torch.rand(N, F, device=dev)
    plan = abi.DevicePlan(di)
    ws = abi.Workspace(E, F, torch.float32, dev)
    out = torch.empty(plan.S, F, device=dev)
    for chunk in [0, 16, 32, 64, 128]:
        os.environ["GEOT_B200_CHUNK"] = str(chunk)
        ws = abi.Workspace(E, F, torch.float32, dev)
        for _ in range(3):
            abi.segment_reduce(src, si, di, w, "sum", plan=plan, out=out, workspace=ws)
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(10):
            abi.segment_reduce(src, si, di, w, "sum", plan=plan, out=out, workspace=ws)
        t1.record(); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 10
        gb = (E * (F * 4 + 16 + 4) + plan.S * F * 4) / 1e9
        print(f"gws E={E} N={N} F={F} chunk={chunk}: {ms:.3f} ms  {gb/ms*1e3:.0f} GB/s logical  {E/ms/1e6:.2f} Gedges/s")
    os.environ["GEOT_B200_CHUNK"] = "0"
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10):
        exp = torch.zeros(plan.S, F, device=dev).index_add_(0, di, src.index_select(0, si) * w.unsqueeze(-1))
    t1.record(); torch.cuda.synchronize()
    print(f"  torch index_select*mul+index_add_: {t0.elapsed_time(t1)/10:.3f} ms")
    got = abi.segment_reduce(src, si, di, w, "sum", plan=plan)
    print("  max rel err vs torch:", ((got - exp).abs() / exp.abs().clamp_min(1e-6)).max().item())
