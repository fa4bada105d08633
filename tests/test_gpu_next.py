"""GPU parity for the "next" rows: sddmm_coo (weight gradient), csr_gws / coo_to_csr (CSR entry
point), the GCN / GraphSAGE forward, and tracing through the operators.  Through the C ABI and the operators,
against the CPU oracle, the golden vectors and -- where it was built -- the reference's own CUDA kernels."""
import os

import numpy as np
import pytest
import torch

import oracle
from tests.helpers import gnn_restatement as restate

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import geot_b200
    from geot_b200 import abi, gnn

DEV = "cuda"


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def _graph(E, N, seed, sort_src=True):
    g = torch.Generator().manual_seed(seed)
    w = torch.rand(N, generator=g) ** 3
    di = torch.multinomial(w / w.sum(), E, replacement=True, generator=g)
    si = torch.randint(0, N, (E,), generator=g)
    key, perm = torch.sort(di * N + si) if sort_src else torch.sort(di)
    return si[perm].contiguous(), di[perm].contiguous(), g


def _rel_ok(got, exp, tol, floor=1e-30):
    got, exp = got.detach().cpu().double(), exp.detach().cpu().double()
    return bool(((got - exp).abs() <= tol * exp.abs().clamp_min(floor)).all())


# ---- sddmm_coo ----------------------------------------------------------------------------------------
def test_sddmm_golden(golden_dir):
    g = _load(golden_dir, "ref_test_sddmm.npz")
    a = [g[k].to(DEV) for k in ("src_index", "dst_index", "mat_1", "mat_2")]
    out = geot_b200.sddmm_coo_impl(*a)
    assert _rel_ok(out, g["expected"], 1e-5)
    out32 = geot_b200.sddmm_coo_impl(a[0].int(), a[1].int(), a[2], a[3])      # the reference narrows to int32
    assert torch.equal(out, out32)


@pytest.mark.parametrize("F", [1, 2, 3, 4, 7, 8, 16, 31, 32, 48, 64, 100, 128, 192, 256, 512, 520, 1000])
def test_sddmm_widths_fp32(F):
    E, N = 5000 + F, 300
    si, di, g = _graph(E, N, seed=F)
    x1 = torch.rand(N, F, generator=g) - 0.25
    x2 = torch.rand(N, F, generator=g) - 0.25
    exp = oracle.sddmm_coo(si, di, x1, x2)
    tol = 1e-5 * (x1[di].abs() * x2[si].abs()).sum(-1).double()    # conditioning-aware: mixed signs cancel
    for row, col in ((di, si), (si, di)):                          # sorted and unsorted row index
        a, b = (x1, x2)
        got = abi.sddmm_coo(a.to(DEV), row.to(DEV), b.to(DEV), col.to(DEV)).cpu().double()
        ref = exp.double() if row is di else oracle.sddmm_coo(di, si, x1, x2).double()
        assert bool(((got - ref).abs() <= tol + 1e-30).all()), F


@pytest.mark.parametrize("dtype", [torch.float64, torch.bfloat16, torch.float16])
def test_sddmm_dtypes(dtype):
    for F in (5, 8, 64, 136, 264, 1024):
        E, N = 4000, 200
        si, di, g = _graph(E, N, seed=F + 1)
        x1 = (torch.rand(N, F, generator=g) + 0.1).to(dtype)
        x2 = (torch.rand(N, F, generator=g) + 0.1).to(dtype)
        got = geot_b200.sddmm_coo_impl(si.to(DEV), di.to(DEV), x1.to(DEV), x2.to(DEV))
        assert got.dtype == dtype
        assert _rel_ok(got, oracle.sddmm_coo(si, di, x1, x2), 1e-12 if dtype == torch.float64 else 1e-2), (dtype, F)


def test_sddmm_edge_cases():
    x = torch.rand(7, 64, device=DEV)
    one = geot_b200.sddmm_coo_impl(torch.tensor([3], device=DEV), torch.tensor([5], device=DEV), x, x)
    assert _rel_ok(one, (x[5] * x[3]).sum().view(1), 1e-6)
    empty = geot_b200.sddmm_coo_impl(torch.zeros(0, dtype=torch.long, device=DEV), torch.zeros(0, dtype=torch.long, device=DEV), x, x)
    assert empty.shape == (0,)
    # one segment spanning many chunks + ragged tail
    E = 70001
    di = torch.full((E,), 2, device=DEV); si = torch.randint(0, 7, (E,), device=DEV)
    got = geot_b200.sddmm_coo_impl(si, di, x, x)
    assert _rel_ok(got, (x[di] * x[si]).sum(-1), 1e-5)
    with pytest.raises(RuntimeError, match="same width"):
        geot_b200.sddmm_coo_impl(si, di, x, torch.rand(7, 32, device=DEV))


def test_backward_uses_sddmm_and_cached_transpose():
    E, N, F = 6000, 150, 48
    si, di, g = _graph(E, N, seed=3)
    di[-1] = N - 1
    si, di = si.to(DEV), di.to(DEV)
    w0 = torch.rand(E, generator=g).to(DEV); x0 = torch.rand(N, F, generator=g).to(DEV); gout = torch.rand(N, F, generator=g).to(DEV)
    for _ in range(2):      # second pass hits the transpose cache
        x = x0.clone().requires_grad_(True); w = w0.clone().requires_grad_(True)
        geot_b200.gather_weight_scatter(si, di, w, x).backward(gout)
        xr = x0.clone().requires_grad_(True); wr = w0.clone().requires_grad_(True)
        torch.zeros(N, F, device=DEV).index_add(0, di, wr.unsqueeze(-1) * xr.index_select(0, si)).backward(gout)
        assert torch.allclose(x.grad, xr.grad, rtol=1e-4, atol=1e-5)
        assert torch.allclose(w.grad, wr.grad, rtol=1e-4, atol=1e-5)
    from geot_b200 import transpose
    assert len(transpose._CACHE) >= 1
    si2 = si.clone(); si2[0] = (si2[0] + 1) % N               # a different graph must not hit the cache
    x = x0.clone().requires_grad_(True)
    geot_b200.gather_scatter(si2, di, x).backward(gout)
    xr = x0.clone().requires_grad_(True)
    torch.zeros(N, F, device=DEV).index_add(0, di, xr.index_select(0, si2)).backward(gout)
    assert torch.allclose(x.grad, xr.grad, rtol=1e-4, atol=1e-5)


# ---- csr_gws / coo_to_csr -------------------------------------------------------------------------------
def test_csr_gws_golden(golden_dir):
    g = _load(golden_dir, "ref_test_csr_gws.npz")
    rowptr, col, val, src = (g[k].to(DEV) for k in ("rowptr", "colidx", "val", "src"))
    out = geot_b200.csr_gws(rowptr, col, val, src)                     # int32 rowptr as the reference's helper makes it
    nrow = rowptr.numel() - 1
    assert out.shape == (nrow + 1, src.shape[1])                       # csrc/csr_gws.cpp:29-31
    assert out[nrow].abs().sum().item() == 0
    assert torch.allclose(out[:nrow].cpu(), g["expected"], atol=1e-4)  # test/test_csr_gws.py:47
    assert _rel_ok(out[:nrow], g["expected"], 1e-5)
    assert torch.equal(out, geot_b200.csr_gws(rowptr.long(), col.int(), val, src))
    assert torch.equal(geot_b200.coo_to_csr(g["dst_sorted"].to(DEV)).cpu(), g["rowptr"][: int(g["dst_sorted"][-1]) + 2])
    assert torch.equal(abi.csr_to_coo(rowptr, col.numel()).cpu(), g["dst_sorted"])


@pytest.mark.parametrize("E,N,F", [(1, 1, 4), (3000, 50, 32), (60000, 4000, 128), (200000, 300, 64)])
def test_csr_gws_vs_oracle(E, N, F):
    si, di, g = _graph(E, N, seed=E)
    nrow = N + 3                                                       # trailing empty rows
    rowptr = oracle.rowptr(di, nrow)
    w = torch.rand(E, generator=g); x = torch.rand(N, F, generator=g)
    out = geot_b200.csr_gws(rowptr.to(DEV), si.to(DEV), w.to(DEV), x.to(DEV))
    assert _rel_ok(out, oracle.csr_gws(rowptr, si, w, x), 1e-5)
    # bit-exact integer work: coo_to_csr (int32, like the reference) and its inverse
    assert torch.equal(geot_b200.coo_to_csr(di.to(DEV)).cpu().long(), oracle.rowptr(di, int(di[-1]) + 1))
    assert geot_b200.coo_to_csr(di.to(DEV)).dtype == torch.int32
    assert torch.equal(abi.csr_to_coo(rowptr.to(DEV), E).cpu(), oracle.csr_to_coo(rowptr))
    assert torch.equal(abi.csr_to_coo(rowptr.int().to(DEV), E).cpu(), di)


# ---- GCN / GraphSAGE forward (BASELINE config #5 shape, scaled) ---------------------------------------------
def test_gcn_and_graphsage_forward_match_torch_restatement():
    import workloads as wl
    graph = wl.power_law_graph("proteins", DEV, scale=1.0 / 64)
    N, si, di = graph.num_nodes, graph.src_index, graph.dst_index
    torch.manual_seed(0)
    x = torch.rand(N, 256, device=DEV)
    norm = gnn.gcn_norm(si, di, N)
    gcn = gnn.GCN(256, 256, 3).to(DEV)
    sage = gnn.GraphSAGE(256, 256, 3).to(DEV)
    with torch.no_grad():
        for p in list(gcn.parameters()) + list(sage.parameters()):     # keep activations O(1) through sum aggregation
            p.mul_(0.05)
        out = gcn(x, si, di, norm)
        ref = restate.forward(gcn, x, si, di, norm, acc_dtype=torch.float64)
        assert out.shape == (N, 256)
        assert torch.allclose(out, ref, rtol=1e-3, atol=1e-4)
        out = sage(x, si, di)
        ref = restate.forward(sage, x, si, di, acc_dtype=torch.float64)
        assert torch.allclose(out, ref, rtol=1e-3, atol=1e-3 * float(ref.abs().max()))
    # the aggregation alone (no GEMM in between) holds the op-level tolerance against the oracle
    h = torch.rand(N, 256, device=DEV)
    agg = geot_b200.gather_weight_scatter(si, di, norm, h)
    assert _rel_ok(agg, oracle.gather_weight_scatter(si.cpu(), di.cpu(), norm.cpu(), h.cpu(), acc64=True), 1e-5)


def test_export_traces_through_the_operators():
    class M(torch.nn.Module):
        def forward(self, si, di, w, x):
            return torch.relu(torch.ops.geot.gather_weight_scatter(si, di, w, x)) + 1.0

    si, di, g = _graph(2000, 64, seed=9)
    di[-1] = 63
    args = (si.to(DEV), di.to(DEV), torch.rand(2000, generator=g).to(DEV), torch.rand(64, 16, generator=g).to(DEV))
    ep = torch.export.export(M(), args)
    assert any("gather_weight_scatter" in str(n.target) for n in ep.graph.nodes)
    assert torch.allclose(ep.module()(*args), M()(*args))


# ---- the reference's own CUDA kernels for these rows ------------------------------------------------------------
@pytest.mark.skipif(not os.path.exists(oracle.REF_EXT_PATH), reason="oracle/_ref/geot_ref_C.so not built")
def test_against_reference_cuda_kernels_next():
    assert oracle.load_ref_extension()
    for (E, N, F) in [(1000, 100, 32), (100000, 3000, 128), (50000, 700, 66), (30000, 500, 7)]:
        si, di, g = _graph(E, N, seed=E + F)
        si, di = si.to(DEV), di.to(DEV)
        x1 = torch.rand(N, F, generator=g).to(DEV); x2 = torch.rand(N, F, generator=g).to(DEV)
        ours = geot_b200.sddmm_coo_impl(si, di, x1, x2)
        ref = torch.ops.geot_ref.sddmm_coo_impl(si, di, x1, x2)
        assert ((ours - ref).abs() <= 2e-5 * ref.abs().clamp_min(1e-30)).all(), (E, N, F)
        w = torch.rand(E, generator=g).to(DEV)
        rowptr = geot_b200.coo_to_csr(di)
        ours = geot_b200.csr_gws(rowptr, si, w, x2)
        ref = torch.ops.geot_ref.csr_gws_impl(rowptr, si, w, x2)
        assert ours.shape == ref.shape
        assert ((ours - ref).abs() <= 2e-5 * ref.abs().clamp_min(1e-30)).all(), (E, N, F)


def test_pattern_transform_numerics():
    """The rewritten program computes what the original torch program computes (test/compile/test_gcn.py prints this diff)."""
    from tests.test_next_host import _mp_models
    GCNLayer, MultiHead, _ = _mp_models()
    g = torch.Generator().manual_seed(2)
    N, E = 300, 5000
    row = torch.randint(0, N, (E,), generator=g)
    col = torch.randint(0, N - 7, (E,), generator=g).sort().values        # the last 7 nodes are isolated
    ei = torch.stack([row, col]).to(DEV)
    x, w = torch.rand(N, 8, generator=g).to(DEV), torch.rand(E, generator=g).to(DEV)
    model = GCNLayer().to(DEV)
    ep = geot_b200.pattern_transform(model, (x, ei, w))
    out, ref = ep.module()(x, ei, w), model(x, ei, w)
    assert out.shape == ref.shape == (N, 16)
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-5)
    xh, wh = torch.rand(N, 4, 8, generator=g).to(DEV), torch.rand(E, 4, generator=g).to(DEV)
    ep = geot_b200.pattern_transform(MultiHead(), (xh, ei[0], ei[1], wh))
    assert torch.allclose(ep.module()(xh, ei[0], ei[1], wh), MultiHead()(xh, ei[0], ei[1], wh), rtol=1e-4, atol=1e-5)
