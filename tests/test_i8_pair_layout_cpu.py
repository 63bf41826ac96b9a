"""CPU model of the B-operand layouts of the CTA-pair int8-slice kernel (plssvm_b200/csrc/tile_i8_pair.cuh, I8PairLayout2 and the MMA issuer's
enumeration, restated here): with tcgen05.mma.cta_group::2 ONE shared-memory offset serves both CTAs and each CTA supplies half of the N rows of an
instruction — D columns [0, N/2) come from CTA 0's rows, [N/2, N) from CTA 1's.  The test builds the per-CTA B areas exactly as the producer fills them
(wide form: region 1 = plane-sized slots, CTA 1 shifted by HALF planes, region 2 = one piece per partial instruction; narrow form: each CTA's NH / 2 rows
of every plane), replays the instruction list and checks that every accumulator t' receives exactly sum_{p+q = t'+S-1} A_p B_q^T — every digit product
once, none twice, in the accumulator layout the shared epilogue reads.  (The GPU tests pin the kernel itself: bit-identical to the single-CTA kernel.)"""
import numpy as np
import pytest


def pair_layout(S, NH, wide):
    spm = 256 // NH if wide else 1            # B planes one instruction covers
    half = spm // 2
    r1_slots = S - half if wide else S
    slot_rows = NH if wide else NH // 2
    r2_rows = (NH // 2) * (spm * (spm - 1) // 2) if wide else 0
    return dict(spm=spm, half=half, r1_slots=r1_slots, slot_rows=slot_rows, r1_rows=r1_slots * slot_rows, r2_rows=r2_rows)


def r2_offset_rows(L, NH, nsl):
    return L["r1_rows"] + (NH // 2) * (nsl * (nsl - 1) // 2)


def fill_b_area(B, S, NH, wide, rank):
    """Rows (NH-row planes B[q]) of one CTA's B area of a ring stage, in shared-memory order — what the producer of CTA `rank` loads."""
    L = pair_layout(S, NH, wide)
    K = B[0].shape[1]
    area = np.zeros((L["r1_rows"] + L["r2_rows"], K), dtype=np.int64)
    if not wide:
        for q in range(S):  # slot q <- this CTA's NH / 2 rows of plane q
            area[q * (NH // 2):(q + 1) * (NH // 2)] = B[q][rank * (NH // 2):(rank + 1) * (NH // 2)]
        return area
    for j in range(L["r1_slots"]):  # region 1: slot j <- plane j + rank HALF
        area[j * NH:(j + 1) * NH] = B[j + rank * L["half"]]
    for nsl in range(1, L["spm"]):  # region 2: the half-plane granules [rank nsl, (rank + 1) nsl) of the planes S - nsl .. S - 1
        off = r2_offset_rows(L, NH, nsl)
        for i in range(nsl):
            u = rank * nsl + i
            plane, h = S - nsl + (u >> 1), u & 1
            area[off + i * (NH // 2):off + (i + 1) * (NH // 2)] = B[plane][h * (NH // 2):(h + 1) * (NH // 2)]
    return area


def instructions(S, NH, wide):
    """(A plane, row offset of the B operand in each CTA's area, N, first accumulator column) of every instruction of one K step."""
    L = pair_layout(S, NH, wide)
    out = []
    for pp in range(S - 1, -1, -1):
        q_lo, cnt = S - 1 - pp, pp + 1
        c = 0
        while L["spm"] * c < cnt:
            nsl = min(L["spm"], cnt - L["spm"] * c)
            if not wide:
                b_off = (q_lo + c) * (NH // 2)
            elif nsl == L["spm"]:
                b_off = (q_lo + L["spm"] * c) * NH
            else:
                b_off = r2_offset_rows(L, NH, nsl)
            out.append((pp, b_off, nsl * NH, c * L["spm"] * NH))
            c += 1
    return out


@pytest.mark.parametrize("S,NH,wide", [(7, 64, True), (3, 128, True), (4, 128, True), (3, 128, False), (4, 128, False), (7, 64, False)])
def test_pair_b_layout_feeds_every_digit_product_once(S, NH, wide):
    rng = np.random.default_rng(100 * S + NH + int(wide))
    K = 32
    A = [[rng.integers(-128, 128, size=(128, K)) for _ in range(S)] for _ in range(2)]  # A[cta][p]: each CTA's own 128 rows
    B = [rng.integers(-128, 128, size=(NH, K)) for _ in range(S)]                        # the unit's NH columns, shared by the pair
    areas = [fill_b_area(B, S, NH, wide, r) for r in range(2)]
    assert areas[0].shape == areas[1].shape  # one descriptor / one expect_tx byte count for both CTAs
    acc = [np.zeros((128, S * NH), dtype=np.int64) for _ in range(2)]
    written = np.zeros(S * NH, dtype=bool)
    for (pp, b_off, N, d_col) in instructions(S, NH, wide):
        assert N % 16 == 0 and N <= 256 and d_col + N <= 512
        rows = np.concatenate([areas[0][b_off:b_off + N // 2], areas[1][b_off:b_off + N // 2]])  # [CTA 0's half | CTA 1's half]
        assert rows.shape[0] == N
        for cta in range(2):
            acc[cta][:, d_col:d_col + N] += A[cta][pp] @ rows.T
        written[d_col:d_col + N] = True
    assert written.all()
    for cta in range(2):
        for t in range(S):
            want = sum(A[cta][p] @ B[q].T for p in range(S) for q in range(S) if p + q == t + S - 1)
            assert np.array_equal(acc[cta][:, t * NH:(t + 1) * NH], want), (cta, t)


def test_pair_layout_sizes_match_the_documented_ones():
    # bytes of B per 64-feature slab and CTA: fp64 wide 32 KiB (single CTA: 28), fp32 wide 20 KiB, fp32 narrow 12 KiB (single CTA: 24)
    for (S, NH, wide, kib) in ((7, 64, True, 32), (3, 128, True, 20), (3, 128, False, 12)):
        L = pair_layout(S, NH, wide)
        assert (L["r1_rows"] + L["r2_rows"]) * 64 == kib * 1024
    # instructions per K step: fp64 wide 10 (as on a single CTA), fp32 wide 4, fp32 narrow 6 = S (S + 1) / 2
    assert [len(instructions(*a)) for a in ((7, 64, True), (3, 128, True), (3, 128, False))] == [10, 4, 6]
