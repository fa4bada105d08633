"""Timings of the "next" rows on the BASELINE shapes: sddmm_coo (weight gradient), csr_gws, backward of
gather_weight_scatter, 3-layer GCN / GraphSAGE forward (config #5).  CUDA events, best / median of 10 after 3 warm-ups;
the reference's CUDA kernels (oracle/_ref) beside ours where they exist.  Bench support, not product.

    python scripts/bench_next.py > gpurun_out/<tag>/bench_next.jsonl
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import geot_b200  # noqa: E402
import oracle  # noqa: E402
import workloads as wl  # noqa: E402
from geot_b200 import gnn  # noqa: E402
from tests.helpers import gnn_restatement as restate  # noqa: E402


def timed(fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return round(ts[0], 4), round(ts[len(ts) // 2], 4)


def main():
    have_ref = oracle.load_ref_extension()
    dev = "cuda"
    for gname, F in (("reddit", 128), ("proteins", 256), ("products", 64)):
        g = wl.power_law_graph(gname, dev)
        N, E, si, di = g.num_nodes, g.num_edges, g.src_index, g.dst_index
        x = wl.features(N, F, torch.float32, dev)
        grad = wl.features(N, F, torch.float32, dev, seed=5)
        w = wl.edge_weights(E, None, torch.float32, dev)
        rec = {"graph": gname, "N": N, "E": E, "F": F}
        # sddmm: logical bytes = E * (2 rows + 16 B indices + 4 B out)
        b, m = timed(lambda: geot_b200.sddmm_coo_impl(si, di, grad, x))
        logical = E * (2 * F * 4 + 16 + 4)
        rec["sddmm_ms"] = {"best": b, "median": m}
        rec["sddmm_GBps_logical"] = round(logical / b / 1e6, 1)
        if have_ref:
            rb, rm = timed(lambda: torch.ops.geot_ref.sddmm_coo_impl(si, di, grad, x), 2, 5)
            rec["sddmm_reference_cuda_ms"] = {"best": rb, "median": rm}
            rec["sddmm_speedup_vs_reference"] = round(rb / b, 2)
        # csr_gws
        rowptr = geot_b200.coo_to_csr(di)
        b, m = timed(lambda: geot_b200.csr_gws(rowptr, si, w, x))
        rec["csr_gws_ms"] = {"best": b, "median": m}
        if have_ref:
            rb, rm = timed(lambda: torch.ops.geot_ref.csr_gws_impl(rowptr, si, w, x), 2, 5)
            rec["csr_gws_reference_cuda_ms"] = {"best": rb, "median": rm}
            rec["csr_gws_speedup_vs_reference"] = round(rb / b, 2)
        # backward of gather_weight_scatter (src grad + weight grad), transposition cached
        xg = x.clone().requires_grad_(True); wg = w.clone().requires_grad_(True)

        def fwd_bwd():
            xg.grad = None; wg.grad = None
            geot_b200.gather_weight_scatter(si, di, wg, xg).backward(grad[: int(di[-1]) + 1])
        b, m = timed(fwd_bwd)
        rec["gws_fwd_bwd_ms"] = {"best": b, "median": m}
        print(json.dumps(rec), flush=True)
        del x, grad, w, xg, wg, rowptr
        if gname == "proteins":      # config #5: 3-layer GCN / GraphSAGE forward, in = hidden = out = 256
            torch.manual_seed(0)
            xin = wl.features(N, 256, torch.float32, dev)
            norm = gnn.gcn_norm(si, di, N)
            gcn = gnn.GCN(256, 256, 3).to(dev)
            sage = gnn.GraphSAGE(256, 256, 3).to(dev)
            with torch.no_grad():
                rec2 = {"model": "3-layer forward, proteins shape, 256-256-256-256, fp32 (TF32 off)", "N": N, "E": E}
                rec2["gcn_forward_ms"] = dict(zip(("best", "median"), timed(lambda: gcn(xin, si, di, norm))))
                rec2["graphsage_forward_ms"] = dict(zip(("best", "median"), timed(lambda: sage(xin, si, di))))
                rec2["gcn_torch_restatement_ms"] = dict(zip(("best", "median"), timed(lambda: restate.forward(gcn, xin, si, di, norm), 1, 3)))
                rec2["aggregation_only_ms"] = dict(zip(("best", "median"), timed(lambda: geot_b200.gather_weight_scatter(si, di, norm, xin))))
                rec2["gemm_only_ms"] = dict(zip(("best", "median"), timed(lambda: gcn.convs[0].lin(xin))))
            print(json.dumps(rec2), flush=True)
            del xin, norm, gcn, sage
        del g, si, di
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
