#!/usr/bin/env python
"""bench.py -- headline benchmark of the segment-reduction hot path (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl own|reference] [--workload NAME]

A "step" is one pass of the hot path over one batch of synthetic input.  Default workload (BASELINE.json
configs[1], the configuration the metric is quoted on): gather_weight_scatter, GCN aggregation on a
synthetic Reddit-shape graph (232,965 nodes, 114,615,892 (dst,src)-sorted edges, F = 128, fp32).

  value      effective GB/s = algorithmic ("logical") bytes per step / time, inputs resident in HBM,
             called through the C ABI (libgeot_b200.so) with a cached format_preprocess plan.
  e2e        the same metric through the host-buffer C-ABI entry (geot_b200_segment_reduce_host):
             pinned HOST operands, H2D copies + kernel + D2H of the result inside the timed region.
  roofline   the dominant kernel (segment_reduce_kernel) timed with CUDA events recorded by the library
             around that kernel alone, against the measured HBM copy bandwidth (MEASURED_PEAKS.json).
  cpu_baseline  the CPU restatement of the reference (oracle/, OpenMP over all host cores; torch's
             index_select*mul+index_add_ beside it) on a bounded sample of the same workload.

--impl reference times the reference's CPU side of the path on the host cores (GeoT ships a CPU kernel
for index_scatter only -- numerically wrong, SURVEY 8a A5 -- and none for gather_weight_scatter, so the
arm is the oracle port / torch restatement; where oracle/_ref was built its csrc/cpu kernel is timed too
and labelled).  N > 1: one process per GPU under torchrun; the dst rows are sharded with balanced edge
counts, each rank reduces its own slice, the src rows travel over NCCL each step inside the timed region
(GEOT_B200_EXCHANGE = pipeline [default] | needed | push | allgather | replicated; strong scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (graph, op, F, H, dtype)
    "reddit_gws": ("reddit", "gather_weight_scatter", 128, 1, torch.float32),
    "reddit_index_scatter": ("reddit", "index_scatter", 128, 1, torch.float32),
    "products_gs64": ("products", "gather_scatter", 64, 1, torch.float32),
    "products_gs256": ("products", "gather_scatter", 256, 1, torch.float32),
    "arxiv_mh_spmm": ("arxiv", "mh_spmm", 32, 8, torch.bfloat16),
    "proteins_gws256": ("proteins", "gather_weight_scatter", 256, 1, torch.float32),
    "config1_index_scatter": ("config1", "index_scatter", 64, 1, torch.float32),
}
DTYPE_NAME = {torch.float32: "f32", torch.float64: "f64", torch.bfloat16: "bf16", torch.float16: "f16"}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # clocks under load: the upper half of the samples (idle samples before/after the region excluded)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(name, device, scale=1.0):
    import workloads as wl
    gname, op, F, H, dtype = WORKLOADS[name]
    s = torch.tensor([], dtype=dtype).element_size()
    if gname == "config1":
        E, S = 1_000_000, 50_000
        dst = wl.random_segments(E, S, device)
        g = wl.Graph("config1", S, None, dst, 0, 0.0)
    else:
        g = wl.power_law_graph(gname, device, scale)
    E, N = g.num_edges, g.num_nodes
    w = None
    if op == "index_scatter":
        x = torch.empty(E, F, device=device, dtype=dtype)
        gen = torch.Generator(device=device).manual_seed(1)
        for i in range(0, E, 1 << 24):
            x[i:i + (1 << 24)].uniform_(0, 1, generator=gen) if dtype == torch.float32 else x[i:i + (1 << 24)].copy_(
                torch.rand(min(1 << 24, E - i), F, device=device, generator=gen))
        si = None
    elif op == "mh_spmm":
        x = wl.features(N, (H, F), dtype, device)
        w = wl.edge_weights(E, H, dtype, device)
        si = g.src_index
    else:
        x = wl.features(N, F, dtype, device)
        if op == "gather_weight_scatter":
            w = wl.edge_weights(E, None, dtype, device)
        si = g.src_index
    S = int(g.dst_index[-1]) + 1
    return dict(name=name, graph=g, op=op, F=F, H=H, dtype=dtype, esize=s, x=x, w=w, si=si, di=g.dst_index, E=E, N=N, S=S,
                bytes_logical=wl.bytes_logical(op, E, S, N, F, H, s), bytes_compulsory=wl.bytes_compulsory(op, E, S, N, F, H, s))


def cpu_sample(wk, frac_edges=1.0 / 16, max_edges=8_000_000):
    """Bounded CPU-side sample of the workload: a prefix of the sorted edge list (whole segments)."""
    E = wk["E"]
    n = int(min(E, max(min(E, 1_000_000), min(max_edges, E * frac_edges))))
    di = wk["di"][:n].cpu()
    S = int(di[-1]) + 1
    si = wk["si"][:n].cpu() if wk["si"] is not None else None
    w = wk["w"][:n].cpu() if wk["w"] is not None else None
    x = wk["x"].cpu() if wk["si"] is not None else wk["x"][:n].cpu()
    import workloads as wl
    nbytes = wl.bytes_logical(wk["op"], n, S, x.shape[0], wk["F"], wk["H"], wk["esize"])
    return dict(n=n, S=S, di=di, si=si, w=w, x=x, bytes=nbytes)


def time_cpu(fn, warmup, steps):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return sum(ts) / len(ts), min(ts)


def cpu_arm(wk, warmup, steps):
    """Times the CPU restatements on a bounded sample.  Returns the cpu_baseline object + per-step seconds."""
    import oracle
    sm = cpu_sample(wk)
    H = wk["H"]
    res = {}
    if wk["dtype"] == torch.float32:
        fn = lambda: oracle.segment_reduce(sm["x"], sm["si"], sm["di"], sm["w"], "sum", S=sm["S"], H=H, threads=True)
    else:   # the threaded C variant is fp32-only; low precision goes through the upcasting wrapper
        x32, w32 = sm["x"].float(), (sm["w"].float() if sm["w"] is not None else None)
        fn = lambda: oracle.segment_reduce(x32, sm["si"], sm["di"], w32, "sum", S=sm["S"], H=H, threads=True)
    res["oracle_port_openmp"] = time_cpu(fn, warmup, steps)

    def torch_fn():
        x = sm["x"].float()
        g = x if sm["si"] is None else x.index_select(0, sm["si"])
        if sm["w"] is not None:
            ww = sm["w"].float()
            g = g * (ww.unsqueeze(-1) if ww.dim() < g.dim() else ww)
        return torch.zeros([sm["S"]] + list(x.shape[1:])).index_add_(0, sm["di"], g)
    res["torch_restatement"] = time_cpu(torch_fn, warmup, steps)
    kind = "port"
    if wk["op"] == "index_scatter" and wk["dtype"] == torch.float32 and oracle.load_ref_extension():
        ref = lambda: torch.ops.geot_ref.index_scatter(0, sm["di"], sm["x"], "sum", True)
        res["reference_csrc_cpu_NUMERICALLY_WRONG_A5"] = time_cpu(ref, warmup, steps)
    best = min(res, key=lambda k: res[k][0])
    cores = max(torch.get_num_threads(), 1)
    gbs = {k: sm["bytes"] / v[0] / 1e9 for k, v in res.items()}
    sample = ("first %d of %d edges (%d dst rows) of the workload, sum; GB/s on the sample's logical bytes; avg of %d runs: "
              % (sm["n"], wk["E"], sm["S"], steps) + ", ".join("%s %.2f GB/s" % (k, v) for k, v in gbs.items())
              + "; os.cpu_count=%s torch_threads=%d" % (os.cpu_count(), torch.get_num_threads()))
    obj = {"value": round(gbs[best], 3), "unit": "GB/s", "cores": cores, "kind": kind, "sample": sample,
           "edges_per_s": sm["n"] / res[best][0], "which": best}
    return obj, res[best][0]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.manual_seed(0)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    scale = 1.0 if dev == "cuda" else 1.0 / 16
    wk = build_workload(args.workload, dev, scale)
    obj, sec = cpu_arm(wk, max(1, min(args.warmup, 2)), max(1, min(args.steps, 5)))
    line = {
        "impl": "reference", "metric": metric_name(wk), "value": obj["value"], "unit": "GB/s", "n_gpus": args.gpus,
        "steps": max(1, min(args.steps, 5)), "warmup": max(1, min(args.warmup, 2)), "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": DTYPE_NAME[wk["dtype"]],
        "data": "synthetic", "config": config_of(wk, args.gpus),
        "cpu_baseline": obj,
        "e2e": {"value": obj["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "edges_per_s": obj["edges_per_s"], "gpu_launches": 0,
    }
    print(json.dumps(line))


def metric_name(wk):
    return "%s effective GB/s (logical bytes / time)" % wk["op"]


def config_of(wk, n_gpus, exchange="allgather"):
    g = wk["graph"]
    return {"workload": "%s: %s on synthetic %s-shape graph, %d nodes, %d (dst,src)-sorted edges, F=%d%s, %s" % (
                wk["name"], wk["op"], g.name, wk["N"], wk["E"], wk["F"], (" x H=%d" % wk["H"]) if wk["H"] > 1 else "",
                DTYPE_NAME[wk["dtype"]]),
            "edges": wk["E"], "nodes": wk["N"], "dst_rows": wk["S"], "F": wk["F"], "H": wk["H"],
            "max_degree": g.max_degree, "degree_cv": round(g.degree_cv, 3),
            "bytes_logical_per_step": wk["bytes_logical"], "bytes_compulsory_per_step": wk["bytes_compulsory"],
            "l2": "per-step input streams (%.2f GB) exceed the 126 MB L2; no explicit flush" % (
                (wk["bytes_compulsory"]) / 1e9),
            "parallelism": "1 GPU" if n_gpus == 1 else "dst rows sharded over %d GPUs (edge-balanced); src rows per step: %s" % (
                n_gpus, {"pipeline": "staggered NCCL send/recv steps overlapped with per-owner edge buckets",
                         "needed": "staggered NCCL send/recv of ONLY the rows each bucket references (packed per peer), "
                                   "overlapped with per-owner edge buckets",
                         "push": "ONE kernel stores the referenced rows straight into the requesters' symmetric-memory buffers "
                                 "over NVLink (no NCCL on the data path), overlapped with the local-src edge bucket",
                         "allgather": "one NCCL all-gather, then one reduction",
                         "replicated": "NONE inside the step (src pre-replicated: kernel scaling only, SURVEY 8e)",
                         "none": "no exchange (edge-aligned operands)"}[exchange])}


def run_own(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import geot_b200
    from geot_b200 import abi
    from geot_b200 import dist as gdist

    wk = build_workload(args.workload, dev)
    E, S, N, F, H = wk["E"], wk["S"], wk["N"], wk["F"], wk["H"]
    W = F * H
    x, w, si, di = wk["x"], wk["w"], wk["si"], wk["di"]
    layout = abi.W_NONE if w is None else (abi.W_EDGE if w.dim() == 1 else abi.W_EDGE_HEAD)

    # ---- shard (N > 1) ------------------------------------------------------------------------------
    imbalance = 1.0
    if world > 1:
        shard = gdist.shard_graph(si, di, w, rank, world)
        imbalance = shard.imbalance
        rb = shard.row_bounds
        e0, e1 = shard.edge_bounds[rank], shard.edge_bounds[rank + 1]
        l_di, l_si, l_w = shard.dst_index, shard.src_index, shard.weight
        if wk["op"] == "index_scatter":
            l_x_edges = x[e0:e1].contiguous()
        x_local = x[rb[rank]:rb[rank + 1]].contiguous() if wk["op"] != "index_scatter" else None
        l_S = shard.num_local_rows
        h_src_full = x.cpu() if wk["op"] != "index_scatter" else None      # the host-resident src matrix of the e2e leg
        del x, w, si, di
        torch.cuda.empty_cache()
    else:
        l_di, l_si, l_w, l_S = di, si, w, S
    l_E = l_di.numel()
    plan = abi.DevicePlan(l_di, l_S)
    ws = abi.Workspace(l_E, W, wk["dtype"], dev)
    out = torch.empty([l_S] + list(wk["x"].shape[1:]), dtype=wk["dtype"], device=dev)

    x_full = torch.empty([N] + list(wk["x"].shape[1:]), dtype=wk["dtype"], device=dev) if (world > 1 and wk["op"] != "index_scatter") else None
    # N > 1, gather ops: how the src row shards travel.  "pipeline" (default): staggered NCCL send/recv steps overlapped
    # with the reduction of per-owner edge buckets (geot_b200.dist.PipelinedGather); "allgather": one NCCL all-gather,
    # then one reduction.  Both are inside the timed region.
    exchange = os.environ.get("GEOT_B200_EXCHANGE", "pipeline") if (world > 1 and wk["op"] != "index_scatter") else "none"
    if exchange not in ("pipeline", "needed", "push", "allgather", "replicated", "none"):
        raise SystemExit("GEOT_B200_EXCHANGE must be pipeline, needed, push, allgather or replicated")
    calls_per_step = 1
    pg = None
    if exchange == "replicated":
        # "src pre-replicated" (SURVEY 8e): every rank already holds all src rows, no exchange inside the step.  This is
        # the kernel-scaling number reported BESIDE the default (exchange inside the timed region), never instead of it.
        gdist.all_gather_rows(x_local, rb, out=x_full)
    exchanged = None
    if exchange in ("pipeline", "needed", "push"):
        pg = (gdist.PeerPushGather(shard) if exchange == "push"
              else gdist.PipelinedGather(shard, needed_only=(exchange == "needed")))
        exchanged = pg.exchanged_rows()
        pg.local_rows(x_full).copy_(x_local)
        calls_per_step = 2 if exchange == "push" else world
        del ws
        ws = None

    # opt-in experiment (N = 1, gather ops, one weight per edge at most): temporal blocking of the src matrix for L2
    blocked, w_blocked = None, None
    n_blocks = int(os.environ.get("GEOT_B200_SRC_BLOCKS", "0"))
    if n_blocks > 1 and world == 1 and wk["op"] != "index_scatter" and H == 1:
        blocked = gdist.SrcBlockedGather(l_si, l_di, l_S, N, n_blocks)
        w_blocked = blocked.permute_weight(l_w) if l_w is not None else None      # static weights: permuted once
        calls_per_step = n_blocks

    def step():
        if blocked is not None:
            blocked(wk["x"], w_blocked, "sum", out=out, permuted=True)
            return
        if pg is not None:
            pg(x_full, l_w, "sum", out=out)
            return
        if exchange == "replicated":
            xf = x_full
        elif world > 1 and wk["op"] != "index_scatter":
            xf = gdist.all_gather_rows(x_local, rb, out=x_full)
        elif world > 1:
            xf = l_x_edges
        else:
            xf = wk["x"]
        abi.segment_reduce(xf, l_si, l_di, l_w, "sum", S=l_S, H=H, weight_layout=layout, plan=plan, out=out, workspace=ws)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # opt-in experiment (off by default): pin the src feature matrix in the persisting L2 set-aside
    l2_note = None
    if os.environ.get("GEOT_B200_L2_PERSIST", "0") == "1" and wk["op"] != "index_scatter":
        try:
            win, carve = abi.l2_persist(x_full if x_full is not None else wk["x"])
            l2_note = "src pinned in persisting L2: window %d B, carve-out %d B" % (win, carve)
        except abi.AbiError as e:
            l2_note = "l2_persist unavailable: %s" % e

    for _ in range(max(args.warmup, 3)):
        step()
    abi.profile_enable(args.steps * calls_per_step)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(args.steps):
        step()
    ev[1].record()
    barrier()
    total_ms = ev[0].elapsed_time(ev[1])
    kernel_ms = abi.profile_read(args.steps * calls_per_step)     # pipelined exchange: one main-kernel launch per bucket
    abi.profile_enable(0)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([total_ms, sum(kernel_ms) / args.steps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, kmean = t.tolist()
    else:
        kmean = sum(kernel_ms) / args.steps
    ms_per_step = total_ms / args.steps
    value = wk["bytes_logical"] / (ms_per_step * 1e-3) / 1e9

    # ---- e2e at N > 1: the graph lives in HOST memory, sharded by dst rows; every rank pushes its own shard through
    # the host-buffer entry over its own PCIe link (the src matrix comes from the host on every rank, so this leg needs
    # no GPU-to-GPU exchange at all).  Time = max over ranks between two barriers; bytes = sum over ranks.
    e2e_multi = None
    # (index_scatter's src is edge-aligned: skip the leg when a rank's slice would pin more than 8 GB of host memory)
    e2e_fits = wk["op"] != "index_scatter" or (E * W * wk["esize"]) / world <= 8e9
    if world > 1 and e2e_fits:
        ok, el, moved = 1, 0.0, (0, 0)
        n_e2e = max(3, min(args.steps, 5))
        try:
            tail = list(wk["x"].shape[1:])
            del out, x_full, plan
            if ws is not None:
                del ws
            pg = blocked = None
            torch.cuda.empty_cache()
            if l_E > 0:
                hx = (h_src_full if h_src_full is not None else l_x_edges.cpu()).pin_memory()
                hdi = l_di.cpu().pin_memory()
                hsi = l_si.cpu().pin_memory() if l_si is not None else None
                hw = l_w.cpu().pin_memory() if l_w is not None else None
                hout = torch.empty([l_S] + tail, dtype=wk["dtype"]).pin_memory()
                call = lambda: abi.segment_reduce_host(hx, hsi, hdi, hw, "sum", S=l_S, H=H, weight_layout=layout, out=hout)
                call()
        except Exception as ex:      # local failure only (no collective inside): still take part in the reductions below
            ok = 0
            print("rank %d: e2e leg failed: %r" % (rank, ex), file=sys.stderr)
        barrier()
        t0 = time.perf_counter()
        try:
            if ok and l_E > 0:
                for _ in range(n_e2e):
                    call()
                torch.cuda.synchronize()
                moved = abi.host_last_transfer()
        except Exception as ex:
            ok = 0
            print("rank %d: e2e leg failed: %r" % (rank, ex), file=sys.stderr)
        barrier()
        el = (time.perf_counter() - t0) / n_e2e
        t = torch.tensor([el, float(1 - ok)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        b = torch.tensor([float(moved[0]), float(moved[1])], device=dev, dtype=torch.float64)
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
        if t[1].item() == 0:
            e2e_s = t[0].item()
            e2e_multi = {"value": round(wk["bytes_logical"] / e2e_s / 1e9, 2), "unit": "GB/s",
                         "h2d_bytes_per_step": int(b[0].item()), "d2h_bytes_per_step": int(b[1].item()),
                         "ms_per_step": round(e2e_s * 1e3, 3), "steps": n_e2e, "edges_per_s": E / e2e_s,
                         "api": "geot_b200_segment_reduce_host on every rank's dst-row shard (host-resident graph, pinned host "
                                "operands; H2D + kernels + D2H timed between two barriers, max over ranks; bytes summed over ranks)"}

    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_gbs()
    import workloads as wl
    k_bytes = wl.bytes_logical(wk["op"], l_E, l_S, N, F, H, wk["esize"])      # per launch (per rank)
    achieved = k_bytes / (kmean * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": None, "kernel": "geot::segment_reduce_kernel", "kernel_ms": round(kmean, 4),
                "kernel_share_of_step": round(kmean / ms_per_step, 4), "peak_source": peak_src,
                "frac_of_nominal_8000": round(achieved / 8000.0, 4),
                "note": "achieved = logical bytes per launch / CUDA-event duration of the main kernel alone (events recorded by the "
                        "library around it); for gathers logical bytes include L2-served re-reads of src rows (SURVEY 8d); "
                        "compulsory DRAM bytes per launch = %d" % wl.bytes_compulsory(wk["op"], l_E, l_S, N, F, H, wk["esize"])}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and world == 1:       # the ncu captures are of the single-GPU launch
        try:
            roofline["traffic"] = json.load(open(tp)).get(wk["name"], {}).get("dram_bytes_per_launch")
        except Exception:
            pass
    if roofline["traffic"]:
        # what the DRAM interface itself carried: ncu's bytes per launch over the live launch duration
        roofline["dram_gbs"] = round(roofline["traffic"] / (kmean * 1e-3) / 1e9, 1)
        roofline["dram_frac_of_peak"] = round(roofline["dram_gbs"] / peak, 4)

    # ---- e2e: host buffers through the C-ABI host entry (N == 1) -----------------------------------------
    e2e = None
    cpu_obj = None
    if world == 1:
        hx = wk["x"].cpu().pin_memory(); hdi = l_di.cpu().pin_memory()
        hsi = l_si.cpu().pin_memory() if l_si is not None else None
        hw = l_w.cpu().pin_memory() if l_w is not None else None
        hout = torch.empty(out.shape, dtype=out.dtype).pin_memory()
        del ws, out
        torch.cuda.empty_cache()
        h2d = hx.numel() * hx.element_size() + hdi.numel() * 8 + (hsi.numel() * 8 if hsi is not None else 0) + (
            hw.numel() * hw.element_size() if hw is not None else 0)
        d2h = hout.numel() * hout.element_size()
        n_e2e = max(3, min(args.steps, 5))
        call = lambda: abi.segment_reduce_host(hx, hsi, hdi, hw, "sum", S=l_S, H=H, weight_layout=layout, out=hout)
        call()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            call()
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / n_e2e
        try:        # what the last call really moved (differs from the operand sizes under GEOT_B200_HOST_COMPACT)
            h2d, d2h = abi.host_last_transfer()
        except Exception:
            pass
        e2e = {"value": round(wk["bytes_logical"] / e2e_s / 1e9, 2), "unit": "GB/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": round(e2e_s * 1e3, 3), "steps": n_e2e,
               "transport": {"0": "operands as given", "1": "row pointers instead of dst_index",
                             "2": "int32 src_index", "3": "row pointers + int32 src_index"}.get(
                                 os.environ.get("GEOT_B200_HOST_COMPACT", "0"), "operands as given"),
               "edges_per_s": E / e2e_s,
               "api": "geot_b200_segment_reduce_host (C ABI, pinned host operands; H2D + kernels + D2H timed, host wall clock)"}
        cpu_obj, _ = cpu_arm(wk, 1, 3)
        del hx, hdi, hsi, hw, hout
    elif e2e_multi is not None:
        e2e = e2e_multi
    else:
        e2e = {"value": round(value, 2), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "note": "N > 1: the host-buffer leg was skipped (edge-aligned src too large to pin) or failed on some rank "
                       "(stderr); this repeats the device-resident value"}

    # this library's kernels per step: main + fixup per reduction; pipelined exchange adds the combine and, with
    # weights, the edge permutation (NCCL's own copy kernels are not counted)
    launches_per_step = (2 * calls_per_step + ((1 + (1 if l_w is not None else 0)) if pg is not None else 0)
                         + (1 if exchange in ("needed", "push") else 0))      # + the row pack / push kernel
    line = {
        "metric": metric_name(wk), "value": round(value, 2), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": DTYPE_NAME[wk["dtype"]], "data": "synthetic", "config": config_of(wk, world, exchange),
        "edges_per_s": E / (ms_per_step * 1e-3), "frac_of_measured_hbm": round(value / peak, 4),
        "frac_of_nominal_8000": round(value / 8000.0, 4),
        "roofline": roofline, "cpu_baseline": cpu_obj, "e2e": e2e,
        "gpu_launches": launches_per_step * args.steps, "clocks": clocks, "exchange": exchange,
        "shard_imbalance": round(imbalance, 4),
    }
    if l2_note:
        line["config"]["l2_persist"] = l2_note
    if blocked is not None:
        line["config"]["src_blocks"] = n_blocks
        line["gpu_launches"] = (2 * n_blocks + 1) * args.steps
    if exchanged is not None:
        line["config"]["src_rows_received_per_step_rank0"] = exchanged[0]
        line["config"]["src_rows_full_exchange_rank0"] = exchanged[1]
    print(json.dumps(line))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="reddit_gws", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
