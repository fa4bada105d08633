"""GPU parity tests: the CUDA path against the CPU oracle, through the C ABI (geot_b200.abi, ctypes on
libgeot_b200.so) and through the reference-facing operators (torch.ops.geot.* / geot_b200.*).

Tolerances (BASELINE.json north_star): bit-exact for max/min and all index / segment-pointer
preprocessing; |x - ref| <= 1e-5 * |ref| for fp32 sum/mean (summation order differs; the oracle
accumulates in fp64), 1e-2 for bf16/fp16.
"""
import os

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import geot_b200
    from geot_b200 import abi

DEV = "cuda"
RTOL = {torch.float32: 1e-5, torch.float64: 1e-12, torch.bfloat16: 1e-2, torch.float16: 1e-2}


def assert_close(got, exp, dtype, reduce, what=""):
    got, exp = got.detach().cpu(), exp.detach().cpu()
    assert got.shape == exp.shape, (what, got.shape, exp.shape)
    assert got.dtype == exp.dtype
    if reduce in ("max", "min", "amax", "amin"):
        assert torch.equal(torch.isnan(got), torch.isnan(exp)), what
        assert torch.equal(torch.nan_to_num(got.float()), torch.nan_to_num(exp.float())), what  # bit-exact
        return
    g, e = got.double(), exp.double()
    tol = RTOL[dtype]
    floor = 1e-30 if dtype in (torch.float32, torch.float64) else 1e-3
    bad = (g - e).abs() > tol * e.abs().clamp_min(floor)
    assert not bad.any(), "%s: %d bad, max rel err %.3e" % (what, int(bad.sum()), ((g - e).abs() / e.abs().clamp_min(floor)).max())


def make_graph(E, N, seed, skew=0.0, hub=0.0, gaps=False):
    """Sorted dst_index [E] (int64) over N rows + random src_index; optional power-law skew, one hub
    row holding a fraction `hub` of the edges, and empty rows."""
    g = torch.Generator().manual_seed(seed)
    w = torch.rand(N, generator=g) ** (1.0 + 6.0 * skew)
    if gaps:
        w[torch.rand(N, generator=g) < 0.3] = 0
    if hub > 0:
        w[N // 3] = w.sum() * hub / (1 - hub)
    dst = torch.multinomial(w / w.sum(), E, replacement=True, generator=g).sort().values
    src = torch.randint(0, N, (E,), generator=g)
    return src, dst, g


def run_abi(src, si, di, w, reduce, **kw):
    t = lambda x: None if x is None else x.to(DEV)
    return abi.segment_reduce(t(src), t(si), t(di), t(w), reduce, **kw)


# ------------------------------------------------------------------------------------------------
# golden vectors (expected outputs produced by the reference, tests/golden/make_golden.py)
# ------------------------------------------------------------------------------------------------
def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    return {k: torch.from_numpy(z[k]) if z[k].ndim else z[k] for k in z.files}


def test_golden_reference_python_tests(golden_dir):
    g = _load(golden_dir, "ref_test_index_scatter.npz")
    for sorted_flag in (False, True):   # the reference test passes sorted=False on sorted data
        out = geot_b200.index_scatter(0, g["src"].to(DEV), g["index"].to(DEV), "sum", sorted=sorted_flag)
        assert torch.allclose(out.cpu(), g["expected"], atol=1e-4)      # test/test_index_scatter.py:19
        assert_close(out, g["expected"], torch.float32, "sum", "index_scatter")
    g = _load(golden_dir, "ref_test_gather.npz")
    a = [g[k].to(DEV) for k in ("src_index", "dst_index", "weight", "src")]
    out = geot_b200.gather_scatter(a[0], a[1], a[3], "sum")              # reference tests pass reduce
    assert_close(out, g["expected_gs"], torch.float32, "sum", "gather_scatter")
    out = geot_b200.gather_weight_scatter(a[0], a[1], a[2], a[3], "sum")
    assert_close(out, g["expected_gws"], torch.float32, "sum", "gather_weight_scatter")
    g = _load(golden_dir, "ref_test_mh_spmm.npz")
    a = [g[k].to(DEV) for k in ("src_index", "dst_index", "weight", "src")]
    out = geot_b200.mh_spmm_transposed(a[0], a[1], a[2], a[3], "sum")    # test/test_mh_spmm.py:24
    assert_close(out, g["expected"], torch.float32, "sum", "mh_spmm_transposed")
    out = geot_b200.mh_spmm(a[0], a[1], a[2], a[3], "sum")
    assert_close(out, g["expected"], torch.float32, "sum", "mh_spmm")


def test_golden_reference_ctest(golden_dir):
    g = _load(golden_dir, "ctest_segreduce.npz")
    out = run_abi(g["src"], None, g["index"], None, "sum")
    assert_close(out, g["expected"], torch.float32, "sum", "ctest segreduce")
    out = run_abi(g["feat"], g["col"], g["index"], g["weight"], "sum")
    assert_close(out, g["expected_gws"], torch.float32, "sum", "ctest gws")


def test_golden_segment_counts(golden_dir):
    g = _load(golden_dir, "ref_cpu_counts.npz")
    plan = geot_b200.format_preprocess(g["index"].to(DEV))
    assert torch.equal((plan.rowptr[1:] - plan.rowptr[:-1]).cpu(), g["counts"])
    assert plan.has_gaps and plan.is_sorted


# ------------------------------------------------------------------------------------------------
# sweep against the oracle through the C ABI
# ------------------------------------------------------------------------------------------------
WIDTHS = [1, 2, 3, 4, 7, 8, 16, 31, 32, 48, 64, 100, 128, 192, 256, 512, 1000]


@pytest.mark.parametrize("F", WIDTHS)
@pytest.mark.parametrize("reduce", ["sum", "mean", "max", "min"])
def test_abi_vs_oracle_fp32_widths(F, reduce):
    E, N = 6000, 257
    si, di, g = make_graph(E, N, seed=F, skew=0.3, gaps=(F % 2 == 0))
    w = torch.rand(E, generator=g) + 0.25
    # sum / mean: positive data like the reference's tests (torch.rand), so that the 1e-5 bound is on a
    # well-conditioned sum; max / min (bit-exact) also see negative values
    src = torch.rand(N, F, generator=g) - (0.3 if reduce in ("max", "min") else 0.0)
    S = int(di[-1]) + 1
    for name, (a_si, a_w, a_src) in {"index_scatter": (None, None, src[si]), "gather_scatter": (si, None, src),
                                     "gather_weight_scatter": (si, w, src)}.items():
        got = run_abi(a_src, a_si, di, a_w, reduce, S=S)
        exp = oracle.segment_reduce(a_src, a_si, di, a_w, reduce, S=S, acc64=reduce in ("sum", "mean"))
        assert_close(got, exp, torch.float32, reduce, "%s F=%d %s" % (name, F, reduce))


@pytest.mark.parametrize("dtype", [torch.float64, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("reduce", ["sum", "mean", "max", "min", "prod"])
def test_abi_vs_oracle_dtypes(dtype, reduce):
    E, N = 5000, 300
    for F in (5, 8, 64, 136, 264):
        si, di, g = make_graph(E, N, seed=7 + F, skew=0.2)
        lo, hi = (0.9, 1.1) if reduce == "prod" else ((-0.5, 1.0) if reduce in ("max", "min") else (0.1, 1.0))
        w = (torch.rand(E, generator=g) * (hi - lo) + lo).to(dtype)
        src = (torch.rand(N, F, generator=g) * (hi - lo) + lo).to(dtype)
        got = run_abi(src, si, di, w, reduce)
        exp = oracle.segment_reduce(src, si, di, w, reduce)
        if reduce == "prod":
            assert torch.allclose(got.cpu().double(), exp.double(), rtol=RTOL[dtype] * 10 if dtype != torch.float64 else 1e-9, atol=0)
        else:
            assert_close(got, exp, dtype, reduce, "%s F=%d %s" % (dtype, F, reduce))


def test_fp32_prod_and_nan_propagation():
    di = torch.tensor([0, 0, 0, 2, 2, 5])
    src = torch.tensor([[2.0, 1.0], [3.0, float("nan")], [0.5, 4.0], [7.0, -1.0], [float("nan"), -2.0], [1.5, 1.5]])
    for red in ("prod", "max", "min", "sum", "mean"):
        got = run_abi(src, None, di, None, red).cpu()
        exp = oracle.torch_index_scatter(di, src, red)
        assert torch.equal(torch.isnan(got), torch.isnan(exp)), red
        assert torch.equal(torch.nan_to_num(got), torch.nan_to_num(exp)), red
    assert got[1].abs().sum() == 0 and got[3].abs().sum() == 0 and got[4].abs().sum() == 0   # gap rows are 0


@pytest.mark.parametrize("case", ["one_edge", "one_segment", "all_distinct", "hub", "gaps_big", "ragged_tail", "dim_size"])
def test_edge_cases(case):
    g = torch.Generator().manual_seed(99)
    F = 64
    S = None
    if case == "one_edge":
        di = torch.tensor([3]); E = 1
    elif case == "one_segment":            # one row cut by every chunk and tile boundary
        E = 70001; di = torch.full((E,), 2)
    elif case == "all_distinct":
        E = 5003; di = torch.arange(E)
    elif case == "hub":                    # a hub row spanning hundreds of tiles next to tiny rows
        _, di, g = make_graph(200000, 500, seed=5, hub=0.6); E = di.numel()
    elif case == "gaps_big":               # long runs of empty rows
        di = torch.tensor([5, 5, 100000, 100000, 100001, 250000]); E = 6
    elif case == "ragged_tail":
        E = 8 * 64 * 3 + 1; di = torch.randint(0, 40, (E,), generator=g).sort().values
    else:                                  # dim_size larger than index[-1]+1: trailing rows are 0
        E = 1000; di = torch.randint(0, 50, (E,), generator=g).sort().values; S = 80
    N = 97
    si = torch.randint(0, N, (E,), generator=g)
    w = torch.rand(E, generator=g)
    src = torch.rand(N, F, generator=g)
    for reduce in ("sum", "mean", "max"):
        got = run_abi(src, si, di, w, reduce, S=S)
        exp = oracle.segment_reduce(src, si, di, w, reduce, S=S, acc64=reduce != "max")
        assert_close(got, exp, torch.float32, reduce, case)
        # with a plan (no memset when there are no gaps) the result must be identical
        if S is None:
            plan = abi.DevicePlan(di.to(DEV))
            got2 = run_abi(src, si, di, w, reduce, plan=plan)
            assert torch.equal(got, got2), case


@pytest.mark.parametrize("chunk", [8, 16, 32, 64, 128, 256])
def test_partition_size_does_not_change_results_beyond_tolerance(chunk, monkeypatch):
    monkeypatch.setenv("GEOT_B200_CHUNK", str(chunk))
    si, di, g = make_graph(50000, 300, seed=chunk, skew=0.5, hub=0.2)
    w = torch.rand(50000, generator=g)
    for F in (4, 64, 128, 256):
        src = torch.rand(300, F, generator=g)
        got = run_abi(src, si, di, w, "sum")
        exp = oracle.segment_reduce(src, si, di, w, "sum", acc64=True)
        assert_close(got, exp, torch.float32, "sum", "chunk=%d F=%d" % (chunk, F))
        got = run_abi(src, si, di, w, "max")
        assert_close(got, oracle.segment_reduce(src, si, di, w, "max"), torch.float32, "max")


# ring variants (GEOT_B200_RING): 0 = gathered rows straight to registers, 2 / 3 = first-generation cp.async ring,
# 35 / 39 = lean ring depth 3 / 7 (segment_reduce.cuh ShapeOf).  All walk a chunk in the same order, so sums must be
# bit-identical across them, and every one of them is checked against the oracle.
@pytest.mark.parametrize("ring", [2, 3, 35, 39])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_ring_variants_bit_identical_and_vs_oracle(ring, dtype, monkeypatch):
    widths = {torch.float32: (32, 64, 100, 128, 256, 512, 1024), torch.bfloat16: (64, 128, 256, 512), torch.float16: (128, 1024)}[dtype]
    # (E, N, skew, hub, chunk): long rows, short rows (a head in most sub-batches), a hub row cut by many chunks,
    # and edge counts that leave a partial batch / partial chunk at the end of the list
    graphs = [(40000 + 37, 60, 0.2, 0.0, 64), (30000 + 5, 9000, 0.0, 0.0, 32), (65536, 500, 0.6, 0.4, 128), (4099, 40, 0.0, 0.0, 256)]
    for gi, (E, N, skew, hub, chunk) in enumerate(graphs):
        monkeypatch.setenv("GEOT_B200_CHUNK", str(chunk))
        si, di, g = make_graph(E, N, seed=100 * gi + ring, skew=skew, hub=hub, gaps=(gi == 1))
        w = (torch.rand(E, generator=g) + 0.25).to(dtype)
        for F in widths:
            src = torch.rand(N, F, generator=g).to(dtype)
            for name, (a_si, a_w, a_src) in {"index_scatter": (None, None, src[si]), "gather_scatter": (si, None, src),
                                             "gather_weight_scatter": (si, w, src)}.items():
                for reduce in ("sum", "mean"):
                    monkeypatch.setenv("GEOT_B200_RING", "0")
                    base = run_abi(a_src, a_si, di, a_w, reduce)
                    monkeypatch.setenv("GEOT_B200_RING", str(ring))
                    got = run_abi(a_src, a_si, di, a_w, reduce)
                    what = "%s %s ring=%d F=%d graph=%d" % (name, reduce, ring, F, gi)
                    assert torch.equal(got, base), what
                    exp = oracle.segment_reduce(a_src, a_si, di, a_w, reduce, acc64=True)
                    assert_close(got, exp, dtype, reduce, what)


def test_deterministic_bit_reproducible():
    si, di, g = make_graph(300000, 2000, seed=1, skew=0.6, hub=0.1)
    w = torch.rand(300000, generator=g)
    src = torch.rand(2000, 128, generator=g)
    a = run_abi(src, si, di, w, "sum")
    for _ in range(3):
        assert torch.equal(a, run_abi(src, si, di, w, "sum"))


def test_mh_spmm_layouts_and_head_widths():
    E, N = 4000, 150
    for H, F, dtype in [(4, 32, torch.float32), (8, 32, torch.bfloat16), (3, 5, torch.float32), (2, 130, torch.float32),
                        (8, 8, torch.float16), (1, 64, torch.float32)]:
        si, di, g = make_graph(E, N, seed=H * 100 + F)
        w = torch.rand(E, H, generator=g).to(dtype)
        src = torch.rand(N, H, F, generator=g).to(dtype)
        exp = oracle.mh_spmm(si, di, w, src)
        got = geot_b200.mh_spmm(si.to(DEV), di.to(DEV), w.to(DEV), src.to(DEV))
        assert_close(got, exp, dtype, "sum", "mh_spmm H=%d F=%d" % (H, F))
        got_t = geot_b200.mh_spmm_transposed(si.to(DEV), di.to(DEV), w.to(DEV), src.to(DEV))
        assert torch.equal(got, got_t)
        got_abi = run_abi(src, si, di, w.t().contiguous(), "sum", H=H, weight_layout=abi.W_HEAD_EDGE)
        assert torch.equal(got_abi, got)
        for red in ("max", "mean"):
            got = geot_b200.mh_spmm(si.to(DEV), di.to(DEV), w.to(DEV), src.to(DEV), red)
            assert_close(got, oracle.mh_spmm(si, di, w, src, red), dtype, red)


# ------------------------------------------------------------------------------------------------
# format_preprocess: bit-exact integer work
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("E,N,gaps", [(1, 1, False), (1000, 10, False), (50000, 3000, True), (200000, 50, False)])
def test_format_preprocess_bit_exact(E, N, gaps):
    _, di, _ = make_graph(E, N, seed=E, skew=0.4, gaps=gaps)
    S = int(di[-1]) + 1
    plan = geot_b200.format_preprocess(di.to(DEV))
    assert torch.equal(plan.rowptr.cpu(), oracle.rowptr(di, S))                       # == geot::coo_to_csr
    rows, offs = oracle.segment_ptr(di)                                               # == index_scatter_cpu.cpp:36-75
    prow, poff = plan.segment_offsets
    assert torch.equal(prow.cpu(), rows) and torch.equal(poff.cpu(), offs)
    assert (plan.num_edges, plan.num_rows, plan.num_segments) == (E, S, rows.numel())
    assert plan.max_degree == int((offs[1:] - offs[:-1]).max())
    assert plan.is_sorted and plan.has_gaps == (rows.numel() < S)
    assert torch.equal(geot_b200.coo_to_csr(di.to(DEV)).cpu().long(), oracle.rowptr(di, S))   # int32 like the reference
    # the ABI-level plan agrees and the shard rule matches the host restatement
    p2 = abi.DevicePlan(di.to(DEV))
    assert torch.equal(p2.rowptr.cpu(), plan.rowptr.cpu())
    for parts in (1, 2, 3, 8):
        rb, eb = p2.shards(parts)
        rb2, eb2 = geot_b200.dist.shard_bounds_from_rowptr(plan.rowptr.cpu(), parts)
        assert rb == rb2 and eb == eb2
        assert plan.shards(di.to(DEV), parts) == (rb, eb)


def test_unsorted_detection_and_unsorted_path():
    g = torch.Generator().manual_seed(4)
    E, N, F = 20000, 100, 32
    idx = torch.randint(0, N, (E,), generator=g)
    src = torch.rand(E, F, generator=g)
    assert not geot_b200.format_preprocess(idx.to(DEV)).is_sorted
    with pytest.raises(RuntimeError, match="not sorted"):
        geot_b200.index_scatter(0, src.to(DEV), idx.to(DEV), "sum", sorted=True)
    for red in ("sum", "max", "mean"):
        got = geot_b200.index_scatter(0, src.to(DEV), idx.to(DEV), red, sorted=False)
        exp = oracle.index_scatter(0, idx, src, red, acc64=red != "max")
        assert_close(got, exp, torch.float32, red, "unsorted " + red)


@pytest.mark.parametrize("F", [1, 3, 32, 64, 100, 130])
def test_unsorted_sum_vector_atomics_and_deterministic_mode(F):
    """index_scatter(sorted=False), fp32 sum: the vector-atomic path (red.global.add.v4.f32; scalar for rows that are
    not whole 16-byte pieces) against the oracle, with rows that receive no edge; under
    torch.use_deterministic_algorithms the sort-based path runs instead and is bit-reproducible; an index that is in
    fact sorted takes the sorted kernels whatever the flag says (the reference's own test passes sorted=False on a
    sorted index, test/test_index_scatter.py:9-14)."""
    g = torch.Generator().manual_seed(F)
    E, N = 30000, 257
    idx = torch.randint(0, N, (E,), generator=g)
    idx[idx == 7] = 8                                   # row 7 receives nothing
    idx[-1] = 11                                        # the last entry is NOT the largest: S comes from the plan's max
    src = torch.rand(E, F, generator=g)
    exp = oracle.index_scatter(0, idx, src, "sum", acc64=True)
    got = geot_b200.index_scatter(0, src.to(DEV), idx.to(DEV), "sum", sorted=False)
    assert got.shape[0] == int(idx.max()) + 1 and float(got[7].abs().sum()) == 0.0
    assert_close(got, exp, torch.float32, "sum", "unsorted atomics F=%d" % F)
    torch.use_deterministic_algorithms(True)
    try:
        d1 = geot_b200.index_scatter(0, src.to(DEV), idx.to(DEV), "sum", sorted=False)
        d2 = geot_b200.index_scatter(0, src.to(DEV), idx.to(DEV), "sum", sorted=False)
    finally:
        torch.use_deterministic_algorithms(False)
    assert torch.equal(d1, d2)
    assert_close(d1, exp, torch.float32, "sum", "unsorted deterministic F=%d" % F)
    sidx, perm = torch.sort(idx, stable=True)
    ssrc = src[perm].contiguous().to(DEV)
    a = geot_b200.index_scatter(0, ssrc, sidx.to(DEV), "sum", sorted=False)
    b = geot_b200.index_scatter(0, ssrc, sidx.to(DEV), "sum", sorted=True)
    assert torch.equal(a, b)


# ------------------------------------------------------------------------------------------------
# operator-level behaviour (drop-in surface)
# ------------------------------------------------------------------------------------------------
def test_operator_argument_orders_dims_views():
    g = torch.Generator().manual_seed(8)
    E, F = 3000, 24
    idx = torch.randint(0, 77, (E,), generator=g).sort().values.to(DEV)
    src = torch.rand(E, F, generator=g).to(DEV)
    a = geot_b200.index_scatter(0, src, idx)          # wrapper order (geot/index_scatter.py:5)
    b = geot_b200.index_scatter(0, idx, src)          # README / schema order
    assert torch.equal(a, b)
    # N-D src, dim = 0 (index_scatter_base.h:15-17 flattening) and dim = 1 (honoured here)
    src3 = torch.rand(E, 3, 5, generator=g).to(DEV)
    out = geot_b200.index_scatter(0, src3, idx)
    assert_close(out, oracle.torch_index_scatter(idx.cpu(), src3.cpu()), torch.float32, "sum")
    srcT = src.t().contiguous()                       # [F, E]
    out = geot_b200.index_scatter(1, srcT, idx)
    assert torch.equal(out, a.t())
    # non-contiguous / offset views take the element-wise kernels but give the same numbers
    big = torch.rand(E, F + 3, generator=g).to(DEV)
    view = big[:, 1:F + 1]
    assert_close(geot_b200.index_scatter(0, view, idx), oracle.torch_index_scatter(idx.cpu(), view.cpu()), torch.float32, "sum")
    flat = torch.rand(E * F + 1, generator=g).to(DEV)[1:].view(E, F)   # 4-byte aligned base pointer
    assert_close(geot_b200.index_scatter(0, flat, idx), oracle.torch_index_scatter(idx.cpu(), flat.cpu()), torch.float32, "sum")
    # int32 index is accepted by the wrapper
    assert torch.equal(geot_b200.index_scatter(0, src, idx.int()), a)


def test_error_messages_match_the_reference():
    idx = torch.tensor([0, 0, 1], device=DEV)
    x = torch.rand(3, 4, device=DEV)
    with pytest.raises(RuntimeError, match="reduce argument must be either sum, prod, mean, amax or amin, got bogus"):
        geot_b200.index_scatter(0, x, idx, "bogus")                           # reduceutils.h:17-20
    with pytest.raises(RuntimeError, match="dim must be non-negative and less than input dimensions"):
        torch.ops.geot.index_scatter(5, idx, x, "sum", True)                  # index_scatter_cuda.cu:90-91
    with pytest.raises(RuntimeError, match="index length must be equal to src dimension size"):
        torch.ops.geot.index_scatter(0, idx, torch.rand(5, 4, device=DEV), "sum", True)
    with pytest.raises(RuntimeError, match="src must be 2 dimensional"):
        geot_b200.gather_scatter(idx, idx, torch.rand(3, 2, 2, device=DEV))   # gather_scatter_cuda.cu:20
    with pytest.raises(RuntimeError, match="src must be 3 dimensional"):
        geot_b200.mh_spmm(idx, idx, torch.rand(3, 2, device=DEV), x)          # mh_spmm_cuda.cu:29
    with pytest.raises(RuntimeError, match="Invalid weight size"):
        geot_b200.mh_spmm(idx, idx, torch.rand(7, 2, device=DEV), torch.rand(3, 2, 4, device=DEV))  # mh_spmm_base.h:48-49
    with pytest.raises(RuntimeError, match="empty"):
        geot_b200.index_scatter(0, torch.rand(0, 4, device=DEV), torch.zeros(0, dtype=torch.long, device=DEV))
    with pytest.raises(RuntimeError, match="unsupported dtype"):
        geot_b200.index_scatter(0, torch.ones(3, 4, dtype=torch.int32, device=DEV), idx)


def test_plan_cache_tracks_inplace_updates():
    idx = torch.tensor([0, 0, 1, 1, 2], device=DEV)
    x = torch.ones(5, 4, device=DEV)
    assert geot_b200.index_scatter(0, x, idx).shape[0] == 3
    assert geot_b200.index_scatter(0, x, idx).shape[0] == 3          # cache hit
    idx[-1] = 6                                                      # in-place edit bumps the version
    out = geot_b200.index_scatter(0, x, idx)
    assert out.shape[0] == 7 and out[6, 0].item() == 1 and out[3:6].abs().sum().item() == 0
    geot_b200.clear_plan_cache()
    assert geot_b200.index_scatter(0, x, idx).shape[0] == 7


def test_autograd_backward_matches_torch():
    g = torch.Generator().manual_seed(12)
    E, N, F = 4000, 120, 16
    si, di, _ = make_graph(E, N, seed=3)
    di[-1] = N - 1
    si, di = si.to(DEV), di.to(DEV)
    w0 = torch.rand(E, generator=g).to(DEV)
    x0 = torch.rand(N, F, generator=g).to(DEV)
    gout = torch.rand(N, F, generator=g).to(DEV)
    x = x0.clone().requires_grad_(True); w = w0.clone().requires_grad_(True)
    geot_b200.gather_weight_scatter(si, di, w, x).backward(gout)
    xr = x0.clone().requires_grad_(True); wr = w0.clone().requires_grad_(True)
    torch.zeros(N, F, device=DEV).index_add(0, di, wr.unsqueeze(-1) * xr.index_select(0, si)).backward(gout)
    assert torch.allclose(x.grad, xr.grad, rtol=1e-4, atol=1e-5)
    assert torch.allclose(w.grad, wr.grad, rtol=1e-4, atol=1e-5)
    x = x0.clone().requires_grad_(True)
    geot_b200.gather_scatter(si, di, x).backward(gout)
    xr = x0.clone().requires_grad_(True)
    torch.zeros(N, F, device=DEV).index_add(0, di, xr.index_select(0, si)).backward(gout)
    assert torch.allclose(x.grad, xr.grad, rtol=1e-4, atol=1e-5)


def test_non_default_stream_and_host_entry():
    si, di, g = make_graph(30000, 500, seed=21, skew=0.3)
    w = torch.rand(30000, generator=g); src = torch.rand(500, 96, generator=g)
    exp = oracle.segment_reduce(src, si, di, w, "sum", acc64=True)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        got = geot_b200.gather_weight_scatter(si.to(DEV), di.to(DEV), w.to(DEV), src.to(DEV))
    s.synchronize()
    assert_close(got, exp, torch.float32, "sum", "side stream")
    # host-buffer entry of the C ABI: CPU tensors in, CPU tensor out
    out = abi.segment_reduce_host(src, si, di, w, "sum", S=int(di[-1]) + 1)
    assert_close(out, exp, torch.float32, "sum", "host entry")
    out = abi.segment_reduce_host(src.pin_memory(), si.pin_memory(), di.pin_memory(), w.pin_memory(), "max", S=int(di[-1]) + 1)
    assert_close(out, oracle.segment_reduce(src, si, di, w, "max"), torch.float32, "max", "host entry max")


def test_host_entry_pipeline_slices():
    """The host-buffer entry cuts the edge list into slices at segment boundaries (>= 4 slices once
    E >= 4096): hubs, gaps and every op shape must survive the slicing."""
    si, di, g = make_graph(300000, 900, seed=33, skew=0.5, hub=0.3, gaps=True)
    E = di.numel()
    S = int(di[-1]) + 1
    w = torch.rand(E, generator=g)
    src = torch.rand(900, 64, generator=g)
    for red in ("sum", "mean", "min"):
        out = abi.segment_reduce_host(src, si, di, w, red, S=S)
        assert_close(out, oracle.segment_reduce(src, si, di, w, red, S=S, acc64=red != "min"), torch.float32, red, "host gws " + red)
    srcE = torch.rand(E, 24, generator=g)                       # index_scatter: src is sliced with the edges
    out = abi.segment_reduce_host(srcE, None, di, None, "sum", S=S + 5)
    assert_close(out, oracle.segment_reduce(srcE, None, di, None, "sum", S=S + 5, acc64=True), torch.float32, "sum", "host index_scatter")
    H = 4
    wh = torch.rand(E, H, generator=g).bfloat16()
    xh = torch.rand(900, H, 16, generator=g).bfloat16()
    out = abi.segment_reduce_host(xh, si, di, wh, "sum", S=S, H=H)
    assert_close(out, oracle.mh_spmm(si, di, wh, xh), torch.bfloat16, "sum", "host mh_spmm")
    assert abi.lib().geot_b200_host_arena_release() == 0


# ------------------------------------------------------------------------------------------------
# the reference's own CUDA kernels on this GPU (oracle/_ref, built from /root/reference unmodified)
# ------------------------------------------------------------------------------------------------
@pytest.mark.skipif(not os.path.exists(oracle.REF_EXT_PATH), reason="oracle/_ref/geot_ref_C.so not built")
def test_against_reference_cuda_kernels():
    assert oracle.load_ref_extension()
    for (E, N, F) in [(1000, 100, 32), (200000, 5000, 64), (300000, 2000, 128), (50000, 1000, 4)]:
        si, di, g = make_graph(E, N, seed=E + F, skew=0.3)
        si, di = si.to(DEV), di.to(DEV)
        w = torch.rand(E, generator=g).to(DEV)
        src = torch.rand(N, F, generator=g).to(DEV)
        srcE = torch.rand(E, F, generator=g).to(DEV)
        pairs = [
            (geot_b200.index_scatter(0, srcE, di, "sum", True), torch.ops.geot_ref.index_scatter(0, di, srcE, "sum", True)),
            (geot_b200.index_scatter(0, srcE, di, "sum", False), torch.ops.geot_ref.index_scatter(0, di, srcE, "sum", False)),
            (geot_b200.gather_scatter(si, di, src), torch.ops.geot_ref.gather_scatter_impl(si, di, src)),
            (geot_b200.gather_weight_scatter(si, di, w, src), torch.ops.geot_ref.gather_weight_scatter_impl(si, di, w, src)),
        ]
        for ours, ref in pairs:
            # both sides sum in fp32 in different orders: allow 2e-5 between them (each is within 1e-5 of exact)
            assert ours.shape == ref.shape
            assert ((ours - ref).abs() <= 2e-5 * ref.abs().clamp_min(1e-30)).all()
    H, F, E, N = 4, 32, 20000, 700
    si, di, g = make_graph(E, N, seed=77)
    si, di = si.to(DEV), di.to(DEV)
    w = torch.rand(E, H, generator=g).to(DEV); src = torch.rand(N, H, F, generator=g).to(DEV)
    ours, ref = geot_b200.mh_spmm(si, di, w, src), torch.ops.geot_ref.mh_spmm(si, di, w, src, "sum")
    assert ((ours - ref).abs() <= 2e-5 * ref.abs().clamp_min(1e-30)).all()
    ours_t = geot_b200.mh_spmm(si, di, w.t().contiguous(), src)
    ref_t = torch.ops.geot_ref.mh_spmm(si, di, w.t().contiguous(), src, "sum")
    assert ((ours_t - ref_t).abs() <= 2e-5 * ref_t.abs().clamp_min(1e-30)).all()
