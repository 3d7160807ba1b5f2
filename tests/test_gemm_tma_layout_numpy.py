"""NumPy restatement of the shared-memory addressing of gemm_tma_kernel (csrc/gemm.cu): TMA boxes land dense with
the 128-byte swizzle, fragment row gq is mapped to tile row sigma8(gq).  Checks, on the CPU, that every fragment load
reads the element it should and that every quarter-warp (the unit a 16-byte shared-memory load is served in) touches
eight distinct 16-byte bank groups; the kernel itself is tested in test_cabi_gpu.py (test_gemm_all_ops)."""
import numpy as np


def sigma8(g):
    return ((g & 1) << 2) | (g >> 1)


def swizzle128(off):
    """physical byte offset of logical offset `off` inside a 1 KB-aligned buffer (CU_TENSOR_MAP_SWIZZLE_128B)"""
    return off ^ (((off >> 7) & 7) << 4)


def landed_tile(kc, bmn, bk=8):
    """phys[byte offset // 16] = (mn, k) of the complex element the TMA boxes put there"""
    phys = {}
    if kc:   # one box {8 k, bmn rows}: logical offset = mn * 128 + k * 16
        for mn in range(bmn):
            for k in range(bk):
                phys[swizzle128(mn * 128 + k * 16) // 16] = (mn, k)
    else:    # bmn / 8 boxes {8 mn, 8 k rows} of 1 KB: box = mn // 8, logical offset inside = k * 128 + (mn % 8) * 16
        for mn in range(bmn):
            for k in range(bk):
                phys[(mn // 8) * 64 + swizzle128(k * 128 + (mn % 8) * 16) // 16] = (mn, k)
    assert len(phys) == bmn * bk
    return phys


def fragment_offset(kc, grp, gq, tq, kk):
    """byte offset the lane (gq, tq) loads for tile row group `grp`, k-step kk -- as written in the kernel"""
    sg, k = sigma8(gq), kk + tq
    ch = (k ^ sg) << 4
    return (grp * 8 + sg) * 128 + ch if kc else grp * 1024 + k * 128 + ch


def test_fragment_loads_read_the_right_element_without_bank_conflicts():
    for kc in (True, False):
        for bmn in (128, 64):
            tile = landed_tile(kc, bmn)
            for grp in range(bmn // 8):
                for kk in (0, 4):
                    for quarter in range(4):
                        banks = set()
                        for lane in range(8 * quarter, 8 * quarter + 8):
                            gq, tq = lane >> 2, lane & 3
                            off = fragment_offset(kc, grp, gq, tq, kk)
                            assert tile[off // 16] == (grp * 8 + sigma8(gq), kk + tq)
                            banks.add((off // 16) % 8)
                        assert len(banks) == 8


def test_sigma8_is_a_permutation_and_matches_the_epilogue():
    assert sorted(sigma8(g) for g in range(8)) == list(range(8))
    # accumulator pair of a lane = fragment columns 2 tq, 2 tq + 1 -> tile columns tq and 4 + tq
    for tq in range(4):
        assert sigma8(2 * tq) == tq and sigma8(2 * tq + 1) == 4 + tq


# ---- float64 tile (gemm_tma_real_kernel): 8-byte elements, half-warps of 16 lanes -------------------------------------

def sigma8r(g):
    return ((g & 3) << 1) | (g >> 2)


def pi16(parity, g):
    return (parity << 2) | ((g & 2) << 2) | ((g & 4) >> 1) | (g & 1)


def landed_tile_real(kc, bmn=128, bk=16):
    """phys[byte offset // 8] = (mn, k) of the double the TMA boxes put there"""
    phys = {}
    for mn in range(bmn):
        for k in range(bk):
            if kc:      # one box {16 k, bmn rows}
                off = swizzle128(mn * 128 + k * 8)
            else:       # boxes {16 mn, 16 k rows} of 2 KB (two 1 KB swizzle atoms)
                off = (mn // 16) * 2048 + swizzle128(k * 128 + (mn % 16) * 8)
            phys[off // 8] = (mn, k)
    assert len(phys) == bmn * bk
    return phys


def frag_off_real(kc, wbase, t, gq, k):
    if kc:
        r = wbase + t * 8 + sigma8r(gq)
        return r * 128 + (((k >> 1) ^ (r & 7)) << 4) + ((k & 1) << 3)
    box, p = (wbase >> 4) + (t >> 1), pi16(t & 1, gq)
    return box * 2048 + k * 128 + (((p >> 1) ^ (k & 7)) << 4) + ((p & 1) << 3)


def frag_pos(kc, t, g):
    return t * 8 + sigma8r(g) if kc else (t >> 1) * 16 + pi16(t & 1, g)


def test_real_fragment_loads_read_the_right_element_without_bank_conflicts():
    for kc in (True, False):
        tile = landed_tile_real(kc)
        for wbase in (0, 32, 64, 96):
            for t in range(4):
                for kk in (0, 4, 8, 12):
                    for half in range(2):
                        slots = set()
                        for lane in range(16 * half, 16 * half + 16):
                            gq, tq = lane >> 2, lane & 3
                            off = frag_off_real(kc, wbase, t, gq, kk + tq)
                            assert tile[off // 8] == (wbase + frag_pos(kc, t, gq), kk + tq)
                            slots.add((off // 8) % 16)
                        assert len(slots) == 16
    for kc in (True, False):   # the positions of a warp tile's fragments are a permutation of its rows
        assert sorted(frag_pos(kc, t, g) for t in range(8) for g in range(8)) == list(range(64))
