"""GPU parity, on ONE GPU and through the C ABI, of what the multi-GPU path and the host-buffer path are built from:
permute_edges / push_rows (byte moves, bit-exact), the bucketed reduction options of geot_b200_segment_reduce_ex
(accumulate, edge_perm, mean_rowptr), the in-kernel zero-fill of rows without edges, the row-pointer transport of
the host-buffer entry and the resident host graph."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from geot_b200 import abi

DEV = "cuda"

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


@pytest.mark.parametrize("max_ctas", [0, 3])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("width", [2, 3, 6, 8, 64, 128, 130])
def test_push_rows_into_peer_buffers(dtype, width, max_ctas):
    """geot_b200_push_rows on one GPU: the "peers" are three local buffers whose base pointers sit in a device array,
    exactly as symmetric memory presents the mapped buffers of the other GPUs.  Byte moves: bit-exact."""
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
    # max_ctas = 3: the small persistent grid of the overlapped form (every thread loops over several unrolled batches)
    abi.push_rows(x, rows.to(DEV), peer.to(DEV), slot.to(DEV), bases.data_ptr(), aligned16=True, max_ctas=max_ctas)
    torch.cuda.synchronize()
    for p in range(3):
        exp = torch.full(bufs[p].shape, -1.0, dtype=dtype)
        m = peer == p
        exp[slot[m]] = x.cpu()[rows[m]]
        assert torch.equal(bufs[p].cpu(), exp)


@pytest.mark.parametrize("mask", [1])
def test_host_entry_row_pointer_transport(monkeypatch, mask):
    """The host entry's default transport sends each slice's row pointers instead of its dst_index
    (GEOT_B200_HOST_COMPACT=0 sends the operands as given).  Same slices, same kernels, same edge partition =>
    bit-identical results; fewer bytes moved."""
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
            assert h2d < moved[i][0], (mask, i, h2d, moved[i][0])
    assert abi.lib().geot_b200_host_arena_release() == 0


@pytest.mark.parametrize("dtype,F", [(torch.float32, 64), (torch.float32, 7), (torch.float32, 128), (torch.float32, 300),
                                     (torch.bfloat16, 24), (torch.float64, 3)])
@pytest.mark.parametrize("chunk", [0, 8, 64])
def test_rows_without_edges_are_zero_filled_in_kernel(monkeypatch, dtype, F, chunk):
    """No memset of dst (the reference clears all of it first, csrc/gather_scatter.cpp:27-30): the group that sees a
    jump in the sorted index zero-fills the rows in between.  Dirty output buffer, runs of empty rows at the start,
    in the middle (longer than a chunk's rows) and at the end, S beyond the last index, with and without a plan,
    every reduce op."""
    import oracle
    monkeypatch.setenv("GEOT_B200_CHUNK", str(chunk))
    g = torch.Generator().manual_seed(F + chunk)
    N, E = 3000, 20000
    wdeg = torch.rand(N, generator=g) ** 3
    wdeg[:17] = 0                                                # leading empty rows
    wdeg[100:400] = 0                                            # a long run of empty rows
    wdeg[::7] = 0                                                # scattered empty rows
    wdeg[N - 9:] = 0                                             # trailing empty rows (below S)
    di_c = torch.multinomial(wdeg / wdeg.sum(), E, replacement=True, generator=g).sort().values
    si_c = torch.randint(0, N, (E,), generator=g)
    x_c = (torch.rand(N, F, generator=g) + 0.5).to(dtype)
    di, si, x = di_c.to(DEV), si_c.to(DEV), x_c.to(DEV)
    plan = abi.DevicePlan(di)
    assert plan.c.has_gaps == 1 and plan.c.max_row == int(di_c[-1])
    for S in (plan.S, N, N + 37):
        deg = torch.bincount(di_c, minlength=S)
        for reduce in ("sum", "mean", "max", "min"):
            exp = oracle.segment_reduce(x_c, si_c, di_c, None, reduce, S=S)
            for pl in (plan, None):
                out = torch.full((S, F), 7.0, dtype=dtype, device=DEV)       # dirty: stale values must not survive
                abi.segment_reduce(x, si, di, None, reduce, S=S, plan=pl, out=out)
                got = out.cpu()
                assert bool((got[deg == 0] == 0).all()), (S, reduce, pl is None)
                tol = 1e-2 if dtype == torch.bfloat16 else 1e-5
                assert torch.allclose(got.double(), exp.double(), rtol=tol, atol=1e-6), (S, reduce, pl is None)


@pytest.mark.parametrize("dtype,F", [(torch.float32, 128), (torch.float32, 64), (torch.float32, 5), (torch.bfloat16, 64),
                                     (torch.float64, 16)])
@pytest.mark.parametrize("weighted", [False, True])
def test_bucketed_passes_accumulate_perm_mean(monkeypatch, dtype, F, weighted):
    """geot_b200_segment_reduce_ex: the edge list split stably into two dst-sorted buckets (as dist.BucketedGather
    splits it into src-local / src-remote); pass 1 writes every row, pass 2 runs with accumulate, both read the
    caller's weights through edge_perm, mean divides by the degree in the COMPLETE list.  Must equal the one-pass
    result within the sum tolerance, for hub rows cut by many tiles as well."""
    import oracle
    g = torch.Generator().manual_seed(F)
    N, E = 1200, 150000
    wdeg = torch.rand(N, generator=g) ** 4
    wdeg[N // 2] = 0.4 * float(wdeg.sum())                       # hub row: crosses tiles -> fixup path accumulates too
    wdeg[3:30] = 0
    di_c = torch.multinomial(wdeg / wdeg.sum(), E, replacement=True, generator=g).sort().values
    si_c = torch.randint(0, N, (E,), generator=g)
    w_c = (torch.rand(E, generator=g) + 0.25).to(dtype) if weighted else None
    x_c = (torch.rand(N, F, generator=g) + 0.5).to(dtype)
    S = N
    key = (si_c >= N // 3).to(torch.int8)                        # "remote" = src row in the upper two thirds
    perm_c = torch.argsort(key, stable=True)
    n0 = int((key == 0).sum())
    di, si = di_c[perm_c].to(DEV), si_c[perm_c].to(DEV)
    perm32 = perm_c.to(torch.int32).to(DEV)
    x = x_c.to(DEV)
    w = w_c.to(DEV) if weighted else None
    deg = torch.bincount(di_c, minlength=S)
    rowptr = torch.cat([torch.zeros(1, dtype=torch.int64), deg.cumsum(0)]).to(DEV)
    tol = 1e-2 if dtype == torch.bfloat16 else 1e-5
    for chunk in (0, 16):
        monkeypatch.setenv("GEOT_B200_CHUNK", str(chunk))
        for reduce in ("sum", "mean"):
            out = torch.full((S, F), 3.0, dtype=dtype, device=DEV)
            mr = rowptr if reduce == "mean" else None
            for (a, b, acc) in ((0, n0, False), (n0, E, True)):
                plan = abi.DevicePlan(di[a:b], S)
                abi.segment_reduce(x, si[a:b], di[a:b], w, reduce, S=S, plan=plan, out=out, accumulate=acc,
                                   edge_perm=perm32[a:b] if weighted else None, mean_rowptr=mr)
            exp = oracle.segment_reduce(x_c, si_c, di_c, w_c, reduce, S=S, acc64=(dtype == torch.float32))
            assert torch.allclose(out.cpu().double(), exp.double(), rtol=tol, atol=1e-6), (reduce, chunk)
    # what the options do not serve is refused, not mis-computed
    with pytest.raises(abi.AbiError):
        abi.segment_reduce(x, si, di, None, "max", S=S, accumulate=True)
    with pytest.raises(abi.AbiError):
        abi.segment_reduce(x, si, di, None, "mean", S=S, accumulate=True)          # mean needs mean_rowptr


def test_resident_host_graph_matches_stateless_host_entry():
    """geot_b200_host_graph_*: indices uploaded once, per call only src + weights in and dst out.  Same kernels on
    the same slices of the edge list as far as the sums are concerned: equal to the oracle within the sum tolerance,
    bit-exact for max; the bytes per call exclude the index arrays."""
    import oracle
    g = torch.Generator().manual_seed(5)
    E, N, F = 400000, 900, 64
    w_deg = torch.rand(N, generator=g) ** 3
    w_deg[N // 3] = 0.3 * float(w_deg.sum())
    w_deg[5:40] = 0
    di = torch.multinomial(w_deg / w_deg.sum(), E, replacement=True, generator=g).sort().values.contiguous()
    si = torch.randint(0, N, (E,), generator=g)
    S = int(di[-1]) + 1 + 4
    x = torch.rand(N, F, generator=g)
    xe = torch.rand(E, 24, generator=g)
    hg = abi.HostGraph(si, di, S, N)
    for it in range(2):                                          # weights / features change from call to call
        w = torch.rand(E, generator=g)
        x = x + it
        got = hg.reduce(x, w, "sum")
        assert torch.allclose(got, oracle.segment_reduce(x, si, di, w, "sum", S=S, acc64=True), rtol=1e-5, atol=1e-6)
        h2d, d2h, resident = hg.last_transfer()
        assert h2d == x.numel() * 4 + E * 4 and d2h == S * F * 4 and resident == 2 * E * 8
    assert torch.equal(hg.reduce(x, None, "max"), oracle.segment_reduce(x, si, di, None, "max", S=S))
    got = hg.reduce(x, None, "mean")
    assert torch.allclose(got, oracle.segment_reduce(x, si, di, None, "mean", S=S, acc64=True), rtol=1e-5, atol=1e-6)
    wh = torch.rand(E, 4, generator=g).bfloat16()
    xh = torch.rand(N, 4, 16, generator=g).bfloat16()
    got = hg.reduce(xh, wh, "sum", H=4)
    assert torch.allclose(got.float(), oracle.segment_reduce(xh, si, di, wh, "sum", S=S, H=4).float(), rtol=1e-2, atol=1e-2)
    hg.close()
    hi = abi.HostGraph(None, di, S, 0)                           # index_scatter: src rows are edge-aligned
    got = hi.reduce(xe, None, "sum")
    assert torch.allclose(got, oracle.segment_reduce(xe, None, di, None, "sum", S=S, acc64=True), rtol=1e-5, atol=1e-6)
    hi.close()


@pytest.mark.parametrize("n_blocks", [2, 3, 16])
@pytest.mark.parametrize("dtype,F", [(torch.float32, 128), (torch.float32, 20), (torch.bfloat16, 64)])
def test_src_blocked_reduction(monkeypatch, n_blocks, dtype, F):
    """geot_b200_src_blocks_build + segment_reduce_ex(opts.src_blocks): the edge list regrouped stably by src-row
    block (every block dst-sorted, a permutation of the caller's list), reduced block after block into one output.
    Equal to the one-pass result within the sum tolerance; the torch operators take the same path under
    GEOT_B200_SRC_BLOCKS=n and stay within tolerance of the oracle."""
    import oracle
    import geot_b200
    g = torch.Generator().manual_seed(n_blocks * 100 + F)
    N, E = 1500, 120000
    wdeg = torch.rand(N, generator=g) ** 4
    wdeg[N // 2] = 0.3 * float(wdeg.sum())
    wdeg[7:19] = 0
    di_c = torch.multinomial(wdeg / wdeg.sum(), E, replacement=True, generator=g).sort().values
    si_c = torch.randint(0, N, (E,), generator=g)
    key = di_c * N + si_c                                          # (dst, src)-sorted like the benchmark graphs
    key = key.sort().values
    di_c, si_c = key // N, key % N
    w_c = (torch.rand(E, generator=g) + 0.25).to(dtype)
    x_c = (torch.rand(N, F, generator=g) + 0.5).to(dtype)
    S = N
    di, si, w, x = di_c.to(DEV), si_c.to(DEV), w_c.to(DEV), x_c.to(DEV)
    bl = abi.SrcBlocks(si, di, N, n_blocks)
    assert bl.bounds[0] == 0 and bl.bounds[-1] == E and all(bl.bounds[i] <= bl.bounds[i + 1] for i in range(n_blocks))
    perm = bl.edge_perm.cpu().long()
    assert torch.equal(perm.sort().values, torch.arange(E))
    assert torch.equal(bl.src_index.cpu(), si_c[perm]) and torch.equal(bl.dst_index.cpu(), di_c[perm])
    per = (N + n_blocks - 1) // n_blocks
    for b in range(n_blocks):
        s_b, d_b = bl.src_index[bl.bounds[b]:bl.bounds[b + 1]].cpu(), bl.dst_index[bl.bounds[b]:bl.bounds[b + 1]].cpu()
        assert bool(((s_b >= b * per) & (s_b < (b + 1) * per)).all()) and bool((d_b[1:] >= d_b[:-1]).all())
    # a 16-bit output is rounded once per pass (the passes accumulate INTO it): 2^-8 per pass at most; the automatic
    # rule therefore blocks 4- and 8-byte element types only, and a forced count on bf16 is held to that bound
    tol = 2.0 ** -8 * n_blocks if dtype == torch.bfloat16 else 1e-5
    plan = abi.DevicePlan(di, S)
    for reduce in ("sum", "mean"):
        for ww, wc in ((None, None), (w, w_c)):
            out = torch.full((S, F), 5.0, dtype=dtype, device=DEV)
            abi.segment_reduce(x, si, di, ww, reduce, S=S, plan=plan, out=out, src_blocks=bl)
            exp = oracle.segment_reduce(x_c, si_c, di_c, wc, reduce, S=S, acc64=(dtype == torch.float32))
            assert torch.allclose(out.cpu().double(), exp.double(), rtol=tol, atol=1e-6), (reduce, ww is None)
    with pytest.raises(abi.AbiError):
        abi.segment_reduce(x, si, di, None, "max", S=S, plan=plan, src_blocks=bl)
    # the torch operators under a forced block count
    monkeypatch.setenv("GEOT_B200_SRC_BLOCKS", str(n_blocks))
    geot_b200.clear_plan_cache()
    got = geot_b200.gather_weight_scatter(si, di, w, x)
    exp = oracle.gather_weight_scatter(si_c, di_c, w_c, x_c, acc64=(dtype == torch.float32))
    assert torch.allclose(got.cpu().double(), exp.double(), rtol=tol, atol=1e-6)
    got2 = geot_b200.gather_weight_scatter(si, di, w, x)           # cached regrouped list: bit-identical
    assert torch.equal(got, got2)
    got = geot_b200.gather_scatter(si, di, x, "mean")
    assert torch.allclose(got.cpu().double(), oracle.gather_scatter(si_c, di_c, x_c, "mean", acc64=(dtype == torch.float32)).double(),
                          rtol=tol, atol=1e-6)
    assert torch.equal(geot_b200.gather_scatter(si, di, x, "max").cpu(), oracle.gather_scatter(si_c, di_c, x_c, "max"))   # not blocked
    geot_b200.clear_plan_cache()


def test_mean_of_constant_rows_is_the_constant_exactly():
    """mean divides by the segment length through a reciprocal + correction step (div_by_count): when every src row
    holds the same dyadic constants the sums n*c are exact and the mean must return c bit for bit, for every degree
    of the graph (1 .. hub), in-kernel rows and rows finished by the fixup pass alike."""
    import geot_b200
    g = torch.Generator().manual_seed(3)
    N, E, F = 4000, 300000, 32
    wdeg = torch.rand(N, generator=g) ** 6
    wdeg[11] = 0.2 * float(wdeg.sum())
    di = torch.multinomial(wdeg / wdeg.sum(), E, replacement=True, generator=g).sort().values
    di = torch.cat([di, torch.arange(N)]).sort().values           # every row non-empty, many of degree 1
    si = torch.randint(0, N, (di.numel(),), generator=g)
    c = (torch.randint(64, 128, (F,), generator=g).float() / 64.0)  # in [1, 2), 6 fractional bits: n*c exact for n < 2^17
    x = c.unsqueeze(0).expand(N, F).contiguous()
    got = geot_b200.gather_scatter(si.to(DEV), di.to(DEV), x.to(DEV), "mean").cpu()
    assert torch.equal(got, x)
