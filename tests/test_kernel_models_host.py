"""Group-level models of kernel index arithmetic that can be checked without a GPU.

The lean register path of ``segment_reduce_kernel`` (``geot_b200/csrc/kernels/segment_reduce.cuh``, the ``DIRECT`` block;
experiment, GEOT_B200_RING=96) keeps two batches of operands (src row ids, weights) in a circular shared-memory buffer
and two register buffers of U rows (one consumed while the next is in flight).  Its control flow is uniform over the
lanes of a group, so the part that can go wrong silently -- which operand words a prefetch reads (wrap-around into the
next batch's half, the ``pos + c < n_ring`` guard), which half a new batch is parked in, which register buffer a step
consumes -- is modelled here word for word and checked for every shape the launcher instantiates."""
import pytest


def simulate(LPR, U, nfull):
    SB, RING_WORDS = LPR // U, 2 * LPR
    assert SB % 2 == 0                                  # ShapeOf's static_assert for the DIRECT variant
    ids = [None] * RING_WORDS                           # parked operands: the value stands for the edge it belongs to
    n_ring = nfull * LPR

    def batch(bi):
        return [bi * LPR + lane for lane in range(LPR)]

    if nfull > 0:
        ids[0:LPR] = batch(0)
    if nfull > 1:
        ids[LPR:2 * LPR] = batch(1)
    buf = {"va": ids[0:U] if nfull > 0 else None, "vb": None}      # prologue: fetch(ids, va)
    pos = slot = 0
    consumed = []
    for bi in range(nfull):
        has_nn = bi + 2 < nfull
        for s0 in range(0, SB, 2):
            blk = slot + s0 * U
            for t, (cur, nxt) in enumerate((("va", "vb"), ("vb", "va"))):     # step(0, blk, va, vb); step(1, blk, vb, va)
                c = (t + 1) * U
                if pos + c < n_ring:
                    off = (blk + c) & (RING_WORDS - 1)
                    buf[nxt] = list(ids[off:off + U])
                expect = [pos + t * U + u for u in range(U)]
                assert buf[cur] == expect, ("rows", bi, s0, t)
                assert ids[blk + t * U: blk + t * U + U] == expect, ("weights", bi, s0, t)
                consumed += expect
            pos += 2 * U
        if has_nn:
            ids[slot:slot + LPR] = batch(bi + 2)        # this batch's half is free: park batch bi + 2 there
        slot ^= LPR
    assert consumed == list(range(n_ring))


@pytest.mark.parametrize("LPR,U", [(32, 8), (16, 8), (32, 4)])     # (F=128, VPL=1), (F=64, VPL=1), (F=256, VPL=2) fp32
def test_lean_register_path_operand_arithmetic(LPR, U):
    for nfull in range(0, 20):
        simulate(LPR, U, nfull)


def test_compact_host_transport_slice_arithmetic():
    """Model of the host-buffer entry's compact transport (abi.cu, GEOT_B200_HOST_COMPACT=1): the edge list is cut into
    slices at segment boundaries; slice k owns rows [rr0, r1); the host sends the slice's CSR row pointer (the REAL
    library routine, geot_b200_host_row_pointers, runs here), the device expands it to slice-local dst ids and reduces
    into dst + rr0 with S = r1 - rr0.  Checked: the slices tile the rows, the expansion gives dst_index - rr0, and the
    per-slice reductions assemble the unsliced result (the oracle plays the device)."""
    import torch
    import oracle
    from geot_b200 import abi
    g = torch.Generator().manual_seed(5)
    for (E, N, hub, n_slices) in [(5000, 300, 0.0, 4), (20000, 50, 0.5, 4), (9000, 4000, 0.0, 7), (4096, 3, 0.0, 4)]:
        w = torch.rand(N, generator=g) ** 3
        w[N // 4: N // 4 + N // 10] = 0                                    # a run of empty rows
        if hub:
            w[N // 2] = float(w.sum()) * hub / (1 - hub)
        di = torch.multinomial(w / w.sum(), E, replacement=True, generator=g).sort().values.contiguous()
        si = torch.randint(0, N, (E,), generator=g)
        x = torch.rand(N, 6, generator=g)
        S = int(di[-1]) + 1 + 3                                            # trailing empty rows
        # cuts as in geot_b200_segment_reduce_host: equal edge counts, moved right to a segment boundary
        cut = [0]
        for k in range(1, n_slices + 1):
            c = E if k == n_slices else (E // n_slices) * k
            while 0 < c < E and int(di[c]) == int(di[c - 1]):
                c += 1
            if c > cut[-1]:
                cut.append(c)
            if c >= E:
                break
        if cut[-1] != E:
            cut.append(E)
        out = torch.zeros(S, 6)
        covered = 0
        for k in range(len(cut) - 1):
            e0, n = cut[k], cut[k + 1] - cut[k]
            r0 = int(di[e0])
            r1 = int(di[cut[k + 1]]) if k + 2 < len(cut) else S
            rr0 = 0 if k == 0 else r0
            assert rr0 == covered and r1 > rr0                             # the slices tile the rows, in order
            covered = r1
            sl = di[e0:e0 + n].contiguous()
            rp = abi.host_row_pointers(sl, rr0, r1 - rr0, threads=3)
            assert int(rp[0]) == 0 and int(rp[-1]) == n
            local = torch.repeat_interleave(torch.arange(r1 - rr0), rp[1:] - rp[:-1])      # csr_rows_kernel
            assert torch.equal(local, sl - rr0)
            out[rr0:r1] = oracle.segment_reduce(x, si[e0:e0 + n], local, None, "sum", S=r1 - rr0)
        assert covered == S
        assert torch.allclose(out, oracle.segment_reduce(x, si, di, None, "sum", S=S), rtol=1e-5, atol=1e-6)
