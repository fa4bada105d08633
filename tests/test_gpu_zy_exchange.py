"""GPU parity of the device helpers of the multi-GPU path (geot_b200/csrc/exchange.cu) on ONE GPU, through the C ABI:
permute_edges (per-edge weights into bucket order; feature rows packed for a peer -- byte moves, bit-exact) and
combine_partials (bucket partials added in bucket order, mean by degree)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from geot_b200 import abi

DEV = "cuda"
EXPERIMENT = pytest.mark.skipif(os.environ.get("GEOT_B200_TEST_EXPERIMENTS") != "1",
                                reason="opt-in feature built after the round's GPU budget was spent, never run on hardware yet: "
                                       "enable with GEOT_B200_TEST_EXPERIMENTS=1 (scripts/gpu_r02_single.sh does)")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float64])
@pytest.mark.parametrize("width", [1, 2, 3, 4, 8, 12, 32, 64, 65, 128, 256])
def test_permute_edges_records(dtype, width):
    """Records of every size class: 2-byte, 4-byte and 16-byte moves, rows selected with repeats and gaps."""
    g = torch.Generator().manual_seed(width)
    n_in, n_out = 1000 + width, 3000
    x = torch.rand(n_in, width, generator=g).to(dtype)
    if width == 1:
        x = x.view(-1)
    perm = torch.randint(0, n_in, (n_out,), generator=g)
    got = abi.permute_edges(x.to(DEV), perm.to(DEV))
    assert torch.equal(got.cpu(), x[perm])
    # misaligned base pointers (a view one element in) must fall back to the narrower moves
    if width >= 4 and dtype != torch.float64:
        xx = torch.rand(n_in * width + 1, generator=g).to(dtype).to(DEV)
        view = xx[1:].view(n_in, width)
        got = abi.permute_edges(view, perm.to(DEV))
        assert torch.equal(got.cpu(), view.cpu()[perm])


def test_permute_edges_empty():
    x = torch.rand(10, 4, device=DEV)
    got = abi.permute_edges(x, torch.empty(0, dtype=torch.int64, device=DEV))
    assert got.shape == (0, 4)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-6), (torch.float64, 1e-12), (torch.bfloat16, 1e-2), (torch.float16, 2e-3)])
@pytest.mark.parametrize("n_parts", [1, 2, 3, 8])
@pytest.mark.parametrize("W", [1, 7, 64, 128])
def test_combine_partials(dtype, tol, n_parts, W):
    g = torch.Generator().manual_seed(n_parts * 1000 + W)
    S = 777
    parts = (torch.rand(n_parts, S, W, generator=g) + 0.5).to(dtype)
    deg = torch.randint(0, 5, (S,), generator=g)
    rowptr = torch.cat([torch.zeros(1, dtype=torch.int64), deg.cumsum(0)])
    acc = parts.double().sum(0)
    for reduce in ("sum", "mean"):
        out = torch.empty(S, W, dtype=dtype, device=DEV)
        abi.combine_partials(parts.to(DEV), out, reduce, rowptr.to(DEV) if reduce == "mean" else None)
        exp = acc if reduce == "sum" else acc / deg.clamp_min(1).double().unsqueeze(-1)
        assert ((out.cpu().double() - exp).abs() <= tol * exp.abs()).all(), (reduce, n_parts, W)
    if dtype in (torch.float32, torch.float64):      # bucket order is the summation order: bit-exact against it
        seq = parts[0].clone()
        for q in range(1, n_parts):
            seq = seq + parts[q]
        out = torch.empty(S, W, dtype=dtype, device=DEV)
        abi.combine_partials(parts.to(DEV), out, "sum", None)
        assert torch.equal(out.cpu(), seq)


def test_debug_mode_checks_src_index_range(monkeypatch):
    """src_index is unchecked on the fast path (as in the reference); GEOT_B200_DEBUG=1 checks it."""
    import geot_b200
    x = torch.rand(10, 8, device=DEV)
    di = torch.tensor([0, 0, 1, 3], device=DEV)
    good = torch.tensor([1, 2, 3, 9], device=DEV)
    bad = torch.tensor([1, 2, 3, 10], device=DEV)
    monkeypatch.setenv("GEOT_B200_DEBUG", "1")
    out = geot_b200.gather_scatter(good, di, x)
    assert torch.equal(out[3], x[9])
    with pytest.raises(RuntimeError, match="src_index out of range"):
        geot_b200.gather_scatter(bad, di, x)
    with pytest.raises(RuntimeError, match="src_index out of range"):
        geot_b200.gather_weight_scatter(bad - 11, di, torch.rand(4, device=DEV), x)


@EXPERIMENT
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("width", [2, 6, 8, 64, 128, 130])
def test_push_rows_into_peer_buffers(dtype, width):
    """geot_b200_push_rows on one GPU: the "peers" are three local buffers whose base pointers sit in a device array,
    exactly as symmetric memory presents the mapped buffers of the other GPUs.  Byte moves: bit-exact."""
    if width * torch.tensor([], dtype=dtype).element_size() % 4:
        pytest.skip("rows are moved in 4-byte words")
    g = torch.Generator().manual_seed(width)
    n_in, n = 500, 2000
    x = torch.rand(n_in, width, generator=g).to(dtype).to(DEV)
    rows = torch.randint(0, n_in, (n,), generator=g)
    peer = torch.randint(0, 3, (n,), generator=g).to(torch.int32)
    slot = torch.empty(n, dtype=torch.int64)
    for p in range(3):                                           # distinct slots per peer, in random order
        m = peer == p
        slot[m] = torch.randperm(int(m.sum()) + 5, generator=g)[: int(m.sum())]
    bufs = [torch.full((int((peer == p).sum()) + 5, width), -1.0, dtype=dtype, device=DEV) for p in range(3)]
    bases = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=DEV)
    abi.push_rows(x, rows.to(DEV), peer.to(DEV), slot.to(DEV), bases.data_ptr(), aligned16=True)
    torch.cuda.synchronize()
    for p in range(3):
        exp = torch.full(bufs[p].shape, -1.0, dtype=dtype)
        m = peer == p
        exp[slot[m]] = x.cpu()[rows[m]]
        assert torch.equal(bufs[p].cpu(), exp)


@EXPERIMENT
@pytest.mark.parametrize("mask", [1, 2, 3])
def test_host_entry_compact_transport(monkeypatch, mask):
    """GEOT_B200_HOST_COMPACT: row pointers instead of dst_index (1), int32 src_index (2) over the link.  Same slices,
    same kernels, same edge partition => bit-identical to the plain transport; fewer bytes moved."""
    import oracle
    g = torch.Generator().manual_seed(mask)
    E, N, F = 300000, 900, 64
    w_deg = torch.rand(N, generator=g) ** 3
    w_deg[N // 3] = 0.3 * float(w_deg.sum())                    # a hub row
    w_deg[5:40] = 0                                             # a run of empty rows
    di = torch.multinomial(w_deg / w_deg.sum(), E, replacement=True, generator=g).sort().values.contiguous()
    si = torch.randint(0, N, (E,), generator=g)
    S = int(di[-1]) + 1 + 4                                     # trailing empty rows too
    w = torch.rand(E, generator=g)
    x = torch.rand(N, F, generator=g)
    xe = torch.rand(E, 24, generator=g)
    cases = [(x, si, w, "sum", 1), (x, si, None, "mean", 1), (x, si, w, "max", 1), (xe, None, None, "sum", 1)]
    plain, moved = [], []
    monkeypatch.setenv("GEOT_B200_HOST_COMPACT", "0")
    for (src, s_i, ww, red, H) in cases:
        plain.append(abi.segment_reduce_host(src, s_i, di, ww, red, S=S, H=H))
        moved.append(abi.host_last_transfer())
    assert torch.allclose(plain[0], oracle.segment_reduce(x, si, di, w, "sum", S=S, acc64=True), rtol=1e-5, atol=1e-6)
    monkeypatch.setenv("GEOT_B200_HOST_COMPACT", str(mask))
    for threads in ("1", "5"):
        monkeypatch.setenv("GEOT_B200_HOST_THREADS", threads)
        for i, (src, s_i, ww, red, H) in enumerate(cases):
            got = abi.segment_reduce_host(src, s_i, di, ww, red, S=S, H=H)
            assert torch.equal(got, plain[i]), (mask, threads, i)
            h2d, d2h = abi.host_last_transfer()
            assert d2h == moved[i][1]
            if (mask & 1) or s_i is not None:
                assert h2d < moved[i][0], (mask, i, h2d, moved[i][0])
    assert abi.lib().geot_b200_host_arena_release() == 0


@EXPERIMENT
@pytest.mark.parametrize("dtype,F", [(torch.float32, 64), (torch.float32, 7), (torch.bfloat16, 24), (torch.float64, 3)])
def test_zero_only_the_empty_rows(monkeypatch, dtype, F):
    """GEOT_B200_ZERO_EMPTY=1: with a plan, rows without edges are zero-filled one by one instead of a memset of the
    whole output.  Same result bit for bit, on a graph with runs of empty rows, a dirty output buffer and S beyond the
    plan's last row."""
    g = torch.Generator().manual_seed(F)
    N, E = 3000, 20000
    wdeg = torch.rand(N, generator=g) ** 3
    wdeg[100:400] = 0                                            # a run of empty rows
    wdeg[::7] = 0                                                # scattered empty rows
    di = torch.multinomial(wdeg / wdeg.sum(), E, replacement=True, generator=g).sort().values.to(DEV)
    si = torch.randint(0, N, (E,), generator=g).to(DEV)
    x = torch.rand(N, F, generator=g).to(dtype).to(DEV)
    plan = abi.DevicePlan(di)
    assert plan.c.has_gaps == 1
    for S in (plan.S, plan.S + 37):
        outs = []
        for flag in ("0", "1"):
            monkeypatch.setenv("GEOT_B200_ZERO_EMPTY", flag)
            out = torch.full((S, F), 7.0, dtype=dtype, device=DEV)           # dirty: stale values must not survive
            abi.segment_reduce(x, si, di, None, "sum", S=S, plan=plan, out=out)
            outs.append(out.clone())
        assert torch.equal(outs[0], outs[1])
        deg = torch.bincount(di.cpu(), minlength=S)
        assert bool((outs[1].cpu()[deg == 0] == 0).all())


@EXPERIMENT
def test_lean_register_path_bit_identical_and_vs_oracle(monkeypatch):
    """GEOT_B200_RING=96, the lean register path (experiment, fp32 rows of 256 B - 1 KB; other shapes fall back to
    the lean ring): walks a chunk in the same order as every other variant, so sums must be bit-identical to the
    register path (ring 0), and it is checked against the oracle.  Same graphs as the ring-variant test."""
    import oracle
    from test_gpu_parity import assert_close, make_graph, run_abi
    graphs = [(40000 + 37, 60, 0.2, 0.0, 64), (30000 + 5, 9000, 0.0, 0.0, 32), (65536, 500, 0.6, 0.4, 128), (4099, 40, 0.0, 0.0, 256),
              (200000 + 3, 700, 0.3, 0.2, 0)]
    for gi, (E, N, skew, hub, chunk) in enumerate(graphs):
        monkeypatch.setenv("GEOT_B200_CHUNK", str(chunk))
        si, di, g = make_graph(E, N, seed=100 * gi + 96, skew=skew, hub=hub, gaps=(gi == 1))
        w = torch.rand(E, generator=g) + 0.25
        for F in (32, 64, 100, 128, 256, 512):
            src = torch.rand(N, F, generator=g)
            for name, (a_si, a_w, a_src) in {"index_scatter": (None, None, src[si]), "gather_scatter": (si, None, src),
                                             "gather_weight_scatter": (si, w, src)}.items():
                for reduce in ("sum", "mean"):
                    monkeypatch.setenv("GEOT_B200_RING", "0")
                    base = run_abi(a_src, a_si, di, a_w, reduce)
                    monkeypatch.setenv("GEOT_B200_RING", "96")
                    got = run_abi(a_src, a_si, di, a_w, reduce)
                    what = "%s %s ring=96 F=%d graph=%d" % (name, reduce, F, gi)
                    assert torch.equal(got, base), what
                    assert_close(got, oracle.segment_reduce(a_src, a_si, di, a_w, reduce, acc64=True), torch.float32, reduce, what)
