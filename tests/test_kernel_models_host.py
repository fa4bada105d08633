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
