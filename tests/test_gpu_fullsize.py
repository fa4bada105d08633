"""Parity at BASELINE.json's full sizes, through size-independent properties plus sampled rows against
the CPU oracle (the oracle cannot reduce 114.6 M x 128 in seconds, but it can reduce any chosen rows):

  * counting:    src = 1, weight = 1  =>  out[r, :] == in-degree(r)            exact (degrees < 2^24)
  * selection:   src[i, :] = i, max   =>  out[r, :] == last src id of row r    exact (edges are (dst,src)-sorted)
  * idempotence: two runs are bit-identical (fixed reduction tree, no atomics)
  * scaling:     op(2 x) == 2 op(x) bit for bit (powers of two commute with rounding)
  * sampled rows (hubs included) recomputed by the oracle from their edge slices: 1e-5 relative.
"""
import pytest
import torch

import oracle
import workloads

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import geot_b200

DEV = "cuda"


def _sample_rows(rowptr, k, seed):
    deg = rowptr[1:] - rowptr[:-1]
    g = torch.Generator().manual_seed(seed)
    rows = torch.randint(0, deg.numel(), (k,), generator=g)
    top = torch.topk(deg.cpu(), 3).indices                      # the hubs: rows cut by the most tiles
    return torch.unique(torch.cat([rows, top, torch.tensor([0, deg.numel() - 1])]))


def _check_rows(out, rows, rowptr, src_index, weight, src, reduce, rtol):
    rp = rowptr.cpu()
    for r in rows.tolist():
        b, e = int(rp[r]), int(rp[r + 1])
        if b == e:
            assert out[r].abs().sum().item() == 0
            continue
        si = src_index[b:e].cpu() if src_index is not None else None
        w = weight[b:e].cpu() if weight is not None else None
        if si is not None:
            uniq, inv = torch.unique(si, return_inverse=True)   # only the rows this segment reads
            x = src[uniq.to(src.device)].cpu()
            si = inv
        else:
            x = src[b:e].cpu()
        di = torch.zeros(e - b, dtype=torch.int64)
        exp = oracle.segment_reduce(x, si, di, w, reduce, S=1, acc64=reduce in ("sum", "mean"))[0]
        got = out[r].cpu()
        if reduce in ("max", "min"):
            assert torch.equal(got, exp), r
        else:
            err = ((got.double() - exp.double()).abs() / exp.double().abs().clamp_min(1e-30)).max().item()
            assert err <= rtol, (r, e - b, err)


@pytest.fixture(scope="module")
def reddit():
    g = workloads.power_law_graph("reddit", DEV)
    yield g
    del g
    torch.cuda.empty_cache()


def test_reddit_shape_gather_weight_scatter_f128(reddit):
    g = reddit
    N, E, F = g.num_nodes, g.num_edges, 128
    assert (N, E) == (232_965, 114_615_892)
    plan = geot_b200.format_preprocess(g.dst_index)
    assert plan.num_rows == N and plan.num_segments == N and not plan.has_gaps and plan.max_degree == g.max_degree
    assert int(plan.rowptr[-1]) == E
    deg = (plan.rowptr[1:] - plan.rowptr[:-1])
    # counting
    ones = torch.ones(N, F, device=DEV)
    out = geot_b200.gather_weight_scatter(g.src_index, g.dst_index, torch.ones(E, device=DEV), ones)
    assert torch.equal(out, deg.float().unsqueeze(1).expand(N, F))
    out = geot_b200.gather_scatter(g.src_index, g.dst_index, ones, "mean")
    assert torch.equal(out, ones)
    # selection
    ids = torch.arange(N, device=DEV, dtype=torch.float32).unsqueeze(1).expand(N, F).contiguous()
    out = geot_b200.gather_scatter(g.src_index, g.dst_index, ids, "max")
    assert torch.equal(out[:, 0], g.src_index[plan.rowptr[1:] - 1].float()) and torch.equal(out[:, 0], out[:, F - 1])
    out = geot_b200.gather_scatter(g.src_index, g.dst_index, ids, "min")
    assert torch.equal(out[:, 5], g.src_index[plan.rowptr[:-1]].float())
    del ids, ones
    # the real thing: random weights and features
    w = workloads.edge_weights(E, device=DEV)
    x = workloads.features(N, F, device=DEV)
    out = geot_b200.gather_weight_scatter(g.src_index, g.dst_index, w, x)
    assert torch.equal(out, geot_b200.gather_weight_scatter(g.src_index, g.dst_index, w, x))     # idempotence
    assert torch.equal(geot_b200.gather_weight_scatter(g.src_index, g.dst_index, w, x * 2), out * 2)  # scaling
    rows = _sample_rows(plan.rowptr, 40, seed=0)
    _check_rows(out, rows, plan.rowptr, g.src_index, w, x, "sum", 1e-5)
    out = geot_b200.gather_weight_scatter(g.src_index, g.dst_index, w, x, "max")
    _check_rows(out, rows, plan.rowptr, g.src_index, w, x, "max", 0)


def test_reddit_shape_index_scatter_f128_beyond_int32(reddit):
    """E*F = 1.47e10 elements: the reference's int products overflow here (index_scatter_kernel.cuh:166)."""
    g = reddit
    N, E, F = g.num_nodes, g.num_edges, 128
    plan = geot_b200.format_preprocess(g.dst_index)
    src = torch.empty(E, F, device=DEV)
    gen = torch.Generator(device=DEV).manual_seed(3)
    step = 1 << 24
    for i in range(0, E, step):
        src[i:i + step].uniform_(0, 1, generator=gen)
    out = geot_b200.index_scatter(0, src, g.dst_index, "sum")
    assert out.shape == (N, F)
    rows = _sample_rows(plan.rowptr, 40, seed=1)
    _check_rows(out, rows, plan.rowptr, None, None, src, "sum", 1e-5)
    out = geot_b200.index_scatter(0, src, g.dst_index, "max")
    _check_rows(out, rows, plan.rowptr, None, None, src, "max", 0)
    src.fill_(1.0)
    out = geot_b200.index_scatter(0, src, g.dst_index, "sum")
    assert torch.equal(out[:, 0], (plan.rowptr[1:] - plan.rowptr[:-1]).float()) and torch.equal(out[:, 0], out[:, 127])
    del src, out
    torch.cuda.empty_cache()


@pytest.mark.parametrize("F", [64, 256])
def test_products_shape_gather_scatter(F):
    g = workloads.power_law_graph("products", DEV)
    N, E = g.num_nodes, g.num_edges
    assert (N, E) == (2_449_029, 61_859_140)
    plan = geot_b200.format_preprocess(g.dst_index)
    x = workloads.features(N, F, device=DEV)
    rows = _sample_rows(plan.rowptr, 40, seed=F)
    for reduce in ("sum", "mean"):
        out = geot_b200.gather_scatter(g.src_index, g.dst_index, x, reduce)
        assert out.shape == (N, F)
        _check_rows(out, rows, plan.rowptr, g.src_index, None, x, reduce, 1e-5)
    ones = torch.ones(N, F, device=DEV)
    out = geot_b200.gather_scatter(g.src_index, g.dst_index, ones)
    assert torch.equal(out[:, F - 1], (plan.rowptr[1:] - plan.rowptr[:-1]).float())
    del g, x, out, ones
    torch.cuda.empty_cache()


def test_arxiv_shape_mh_spmm_bf16_full_oracle():
    g = workloads.power_law_graph("arxiv", DEV)
    N, E, H, F = g.num_nodes, g.num_edges, 8, 32
    x = workloads.features(N, (H, F), torch.bfloat16, DEV)
    w = workloads.edge_weights(E, H, torch.bfloat16, DEV)
    out = geot_b200.mh_spmm(g.src_index, g.dst_index, w, x)
    exp = oracle.mh_spmm(g.src_index.cpu(), g.dst_index.cpu(), w.cpu(), x.cpu())
    err = ((out.cpu().float() - exp.float()).abs() / exp.float().abs().clamp_min(1e-3)).max().item()
    assert err <= 1e-2, err
    assert torch.equal(out, geot_b200.mh_spmm_transposed(g.src_index, g.dst_index, w, x))


def test_config1_index_scatter_full_oracle():
    """BASELINE config #1: src [1M, 64] fp32, 50K random-length segments."""
    E, S, F = 1_000_000, 50_000, 64
    index = workloads.random_segments(E, S, DEV)
    src = workloads.features(E, F, device=DEV)
    out = geot_b200.index_scatter(0, src, index, "sum", sorted=True)
    exp = oracle.index_scatter(0, index.cpu(), src.cpu(), "sum", acc64=True)
    assert ((out.cpu().double() - exp.double()).abs() <= 1e-5 * exp.double().abs()).all()
    for red in ("max", "min"):
        out = geot_b200.index_scatter(0, src, index, red)
        assert torch.equal(out.cpu(), oracle.index_scatter(0, index.cpu(), src.cpu(), red))
    plan = geot_b200.format_preprocess(index)
    assert torch.equal(plan.rowptr.cpu(), oracle.rowptr(index.cpu(), S))


def test_proteins_shape_gcn_and_graphsage_forward_full_size():
    """BASELINE configs[4] at its FULL shape (132 534 nodes, 39.6 M edges, 256-256-256-256, fp32): the 3-layer GCN and
    GraphSAGE forward through the operators against the plain-torch restatement of the same stacks with the aggregation
    accumulated in fp64 (tests/helpers/gnn_restatement.py; the reference's models: models/gcn.py:35-60,
    models/graphsage.py:26-64).  Tolerance: 2e-4 of the largest output per layer stack (fp32 GEMMs, TF32 off, between
    the aggregations)."""
    import workloads as wl
    from geot_b200 import gnn
    from tests.helpers import gnn_restatement as restate
    graph = wl.power_law_graph("proteins", DEV)
    N, si, di = graph.num_nodes, graph.src_index, graph.dst_index
    assert N == 132_534 and si.numel() > 39_000_000
    torch.manual_seed(0)
    x = torch.rand(N, 256, device=DEV)
    norm = gnn.gcn_norm(si, di, N)
    gcn = gnn.GCN(256, 256, 3).to(DEV)
    sage = gnn.GraphSAGE(256, 256, 3).to(DEV)
    with torch.no_grad():
        for p in sage.parameters():                      # sum aggregation over ~300 neighbours per layer: keep activations O(1)
            p.mul_(0.05)
        for model, args in ((gcn, (norm,)), (sage, ())):
            out = model(x, si, di, *args)
            ref = restate.forward(model, x, si, di, *args, acc_dtype=torch.float64)
            assert out.shape == (N, 256) and bool(torch.isfinite(out).all())
            scale = float(ref.abs().max())
            err = float((out - ref).abs().max())
            assert err <= 2e-4 * scale, (type(model).__name__, err, scale)
