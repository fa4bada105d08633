"""CPU-side checks for the "next" rows: the oracle's csr_gws / sddmm_coo / csr_to_coo against the
golden vectors made from the reference's test formulas (tests/golden/make_golden_next.py), the operator schemas
of the added entry points, fake (meta) kernels for tracing, and the host logic of the GNN stacks -- no GPU."""
import os

import numpy as np
import pytest
import torch

import geot_b200
import oracle
from geot_b200 import gnn


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def test_oracle_csr_gws_against_reference_test_formula(golden_dir):
    g = _load(golden_dir, "ref_test_csr_gws.npz")
    out = oracle.csr_gws(g["rowptr"], g["colidx"], g["val"], g["src"])
    nrow = g["rowptr"].numel() - 1
    assert out.shape[0] == nrow + 1                      # csrc/csr_gws.cpp:29-31
    assert out[nrow].abs().sum() == 0
    assert torch.allclose(out[:nrow], g["expected"], atol=1e-4)      # test/test_csr_gws.py:47
    assert ((out[:nrow] - g["expected"]).abs() <= 1e-5 * g["expected"].abs().clamp_min(1e-30)).all()
    # rowptr of the reference's helper (bincount -> cumsum) == the oracle's coo_to_csr, and back
    assert torch.equal(oracle.rowptr(g["dst_sorted"], nrow), g["rowptr"].long())
    assert torch.equal(oracle.csr_to_coo(g["rowptr"]), g["dst_sorted"])
    # the CSR and COO statements of the op agree
    coo = oracle.gather_weight_scatter(g["colidx"], g["dst_sorted"], g["val"], g["src"], acc64=True)
    assert torch.allclose(out[: coo.shape[0]], coo, atol=1e-6)


def test_oracle_sddmm_against_golden(golden_dir):
    g = _load(golden_dir, "ref_test_sddmm.npz")
    out = oracle.sddmm_coo(g["src_index"], g["dst_index"], g["mat_1"], g["mat_2"])
    assert ((out - g["expected"]).abs() <= 1e-6 * g["expected"].abs()).all()
    # the weight gradient of gather_weight_scatter is this op (geot/gather_weight_scatter.py:47)
    E = 300
    si, di = g["src_index"][:E], g["dst_index"][:E].sort().values
    w = torch.rand(E, dtype=torch.float64, requires_grad=True)
    x = g["mat_2"][:, :16].double()
    gout = g["mat_1"][:, :16].double()
    S = int(di[-1]) + 1
    torch.zeros(S, 16, dtype=torch.float64).index_add(0, di, w.unsqueeze(-1) * x[si]).backward(gout[:S])
    assert torch.allclose(w.grad, oracle.sddmm_coo(si, di, gout, x), atol=1e-12)


def test_added_operator_schemas():
    s = lambda op: str(op.default._schema)
    # csrc/gather_weight_scatter.cpp:15-16, csrc/csr_gws.cpp:12-13, geot/csr_gws.py:25, format_transform.py:5
    assert s(torch.ops.geot.sddmm_coo_impl) == "geot::sddmm_coo_impl(Tensor src_index, Tensor dst_index, Tensor mat_1, Tensor mat_2) -> Tensor"
    assert s(torch.ops.geot.csr_gws_impl) == "geot::csr_gws_impl(Tensor indptr, Tensor indices, Tensor weight, Tensor src) -> Tensor"
    assert s(torch.ops.geot.csr_gws) == "geot::csr_gws(Tensor csrptr, Tensor csrind, Tensor weight, Tensor src) -> Tensor"
    assert s(torch.ops.geot.coo_to_csr) == "geot::coo_to_csr(Tensor coo_row) -> Tensor"
    for name in ("csr_gws", "coo_to_csr", "sddmm_coo_impl"):
        assert callable(getattr(geot_b200, name))
    with pytest.raises((NotImplementedError, RuntimeError)):        # no CPU path
        geot_b200.csr_gws(torch.tensor([0, 2]), torch.tensor([0, 0]), torch.rand(2), torch.rand(1, 4))
    with pytest.raises((NotImplementedError, RuntimeError)):
        geot_b200.sddmm_coo_impl(torch.tensor([0]), torch.tensor([0]), torch.rand(1, 4), torch.rand(1, 4))


def test_fake_kernels_trace_every_operator():
    """FakeTensor propagation (what torch.export / torch.compile run) through every geot operator."""
    from torch._subclasses.fake_tensor import FakeTensorMode
    from torch.fx.experimental.symbolic_shapes import ShapeEnv
    with FakeTensorMode(shape_env=ShapeEnv(), allow_non_fake_inputs=False) as mode:
        E, N, F, H = 50, 9, 8, 2
        idx = torch.empty(E, dtype=torch.int64)
        x, w = torch.empty(N, F), torch.empty(E)
        out = torch.ops.geot.gather_weight_scatter(idx, idx, w, x)
        assert out.shape[1] == F and out.dtype == x.dtype
        assert torch.ops.geot.gather_scatter(idx, idx, x).shape[1] == F
        assert torch.ops.geot.gather_scatter_impl(idx, idx, x).shape[1] == F
        assert torch.ops.geot.gather_weight_scatter_reduce(idx, idx, w, x, "max").shape[1] == F
        assert torch.ops.geot.index_scatter(0, idx, torch.empty(E, 3, 5), "sum", True).shape[1:] == (3, 5)
        assert torch.ops.geot.mh_spmm(idx, idx, torch.empty(E, H), torch.empty(N, H, F), "sum").shape[1:] == (H, F)
        assert torch.ops.geot.sddmm_coo_impl(idx, idx, x, x).shape == (E,)
        assert torch.ops.geot.csr_gws(torch.empty(N + 1, dtype=torch.int32), idx, w, x).shape == (N + 1, F)
        assert torch.ops.geot.coo_to_csr(idx).dtype == torch.int32


def test_gcn_norm_and_self_loops_match_the_reference_formulas():
    g = torch.Generator().manual_seed(5)
    N, E = 30, 200
    si = torch.randint(0, N, (E,), generator=g)
    di = torch.randint(0, N, (E,), generator=g)
    key, perm = torch.sort(di * N + si)
    si, di = si[perm], di[perm]
    w = gnn.gcn_norm(si, di, N)
    # models/conv/gcnconv.py:51-55 on a dense matrix: deg = row sums of adj_t, D^-1/2 A D^-1/2
    A = torch.zeros(N, N).index_put_((di, si), torch.ones(E), accumulate=True)
    deg = A.sum(1)
    dis = deg.pow(-0.5); dis[dis == float("inf")] = 0
    An = dis.view(-1, 1) * A * dis.view(1, -1)
    dense = torch.zeros(N, N).index_put_((di, si), w, accumulate=True)
    assert torch.allclose(dense, An, atol=1e-6)
    s2, d2, w2 = gnn.add_self_loops(si, di, N, torch.ones(E))
    assert torch.equal(d2, d2.sort().values) and int((s2 == d2).sum()) == N
    A2 = torch.zeros(N, N).index_put_((d2, s2), w2, accumulate=True)
    A_ref = A.clone(); A_ref.fill_diagonal_(1.0)
    assert torch.equal(A2, A_ref)
    m = gnn.GCN(8, 16, 3)
    assert [c.lin.weight.shape for c in m.convs] == [(16, 8), (16, 16), (16, 16)]
    assert len(gnn.GraphSAGE(8, 16, 3, 4).convs) == 3


def _mp_models():
    class GCNLayer(torch.nn.Module):          # PyG-style message passing spelled with aten ops
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(8, 16, bias=False)

        def forward(self, x, edge_index, w):
            row, col = edge_index[0], edge_index[1]
            h = self.lin(x)
            m = w.view(-1, 1) * h.index_select(0, row)
            out = m.new_zeros((x.shape[0], 16)).index_add(0, col, m)
            agg = h.new_zeros((x.shape[0], 16)).index_add(0, col, h.index_select(0, row))
            return torch.relu(out) + agg

    class MultiHead(torch.nn.Module):
        def forward(self, x, row, col, w):
            return torch.zeros_like(x).index_add(0, col, w.unsqueeze(-1) * x.index_select(0, row))

    class NotAMatch(torch.nn.Module):           # accumulates into a non-zero base: must be left alone
        def forward(self, x, row, col):
            return x.index_add(0, col, x.index_select(0, row))

    return GCNLayer, MultiHead, NotAMatch


def test_pattern_transform_rewrites_message_passing_chains():
    """geot/match_replace/match_replace.py:8-32: index_select -> (mul) -> index_add => geot operators."""
    GCNLayer, MultiHead, NotAMatch = _mp_models()
    g = torch.Generator().manual_seed(1)
    N, E = 20, 100
    ei = torch.stack([torch.randint(0, N, (E,), generator=g), torch.randint(0, N, (E,), generator=g).sort().values])
    ep = geot_b200.pattern_transform(GCNLayer(), (torch.rand(N, 8), ei, torch.rand(E)))
    targets = [str(n.target) for n in ep.graph_module.graph.nodes if n.op == "call_function"]
    assert "geot.gather_weight_scatter.default" in targets and "geot.gather_scatter.default" in targets
    assert not any("index_add" in t or "index_select" in t or "new_zeros" in t for t in targets)
    assert targets.count("geot.pad_rows.default") == 2
    ep = geot_b200.pattern_transform(MultiHead(), (torch.rand(N, 4, 8), ei[0], ei[1], torch.rand(E, 4)))
    targets = [str(n.target) for n in ep.graph_module.graph.nodes if n.op == "call_function"]
    assert "geot.mh_spmm.default" in targets and not any("index_add" in t for t in targets)
    ep = geot_b200.pattern_transform(NotAMatch(), (torch.rand(N, 8), ei[0], ei[1]))
    targets = [str(n.target) for n in ep.graph_module.graph.nodes if n.op == "call_function"]
    assert any("index_add" in t for t in targets) and not any("geot" in t for t in targets)
    x = torch.rand(3, 4)
    assert torch.equal(torch.ops.geot.pad_rows(x, 5)[:3], x) and torch.ops.geot.pad_rows(x, 5)[3:].abs().sum() == 0
