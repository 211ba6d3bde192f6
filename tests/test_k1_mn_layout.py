"""CPU: the index arithmetic of K1's MN-major mode (csrc/tica_umma_v2.cuh, "MN-major mode") restated in Python.

The converter's thread map must (1) cover every (feature chunk, row) unit of a 32-row tile exactly once, (2) keep the
16-byte loads of the swizzled raw rows and the 16-byte stores into the swizzled windows bank-conflict free per quarter
warp (a 128-bit shared-memory access is served 8 lanes at a time), and (3) put element (feature m, ring row r) where
the SWIZZLE_128B MN-major descriptor of the UMMAs expects it -- the layout tools/umma_probe7.cu verified on hardware
(profiles/r2z_probe7_mn_major_sw128.log):  block (m // 64) * BLK + r * 128 + (((m % 64) // 8) ^ (r & 7)) * 16 + (m % 8) * 2.
"""
import itertools

import pytest

KT = 32                      # rows of a tile


def thread_units(ct, narrow):
    """(feature block of 32, chunk of 8 features inside it, rows) of converter thread ct (0..255)."""
    cq, fbl, swp = ct & 3, (ct >> 2) & 1, (ct >> 3) & 1
    fbh = 0 if narrow else (ct >> 4) & 1
    rp = (ct >> 4) if narrow else (ct >> 5)
    row_lo = 2 * rp + (fbl ^ swp)
    rows = [row_lo] if narrow else [row_lo, row_lo + 16]
    return 2 * fbh + fbl, cq, rows


def raw_addr(fb, row, chunk16):
    # raw stage: [block of 32 features][row][128 B], 16-byte chunks XOR-swizzled by row & 7 (TMA SWIZZLE_128B)
    return fb * (KT * 128) + row * 128 + ((chunk16 ^ (row & 7)) << 4)


def window_addr(fb, cq, ring_row, block_bytes):
    # what the converter computes: block = fb >> 1 (64 features), chunk inside the 128-byte row = 4 * (fb & 1) + cq
    return (fb >> 1) * block_bytes + ring_row * 128 + (((4 * (fb & 1) + cq) ^ (ring_row & 7)) << 4)


@pytest.mark.parametrize("narrow", [False, True])
def test_units_cover_the_tile_exactly_once(narrow):
    seen = set()
    for ct in range(256):
        fb, cq, rows = thread_units(ct, narrow)
        for r in rows:
            assert (fb, cq, r) not in seen
            seen.add((fb, cq, r))
    n_fb = 2 if narrow else 4
    assert seen == set(itertools.product(range(n_fb), range(4), range(KT)))


@pytest.mark.parametrize("narrow", [False, True])
@pytest.mark.parametrize("stages", [4, 5])
def test_quarter_warps_are_bank_conflict_free(narrow, stages):
    block_bytes = (32 * stages + 32) * 128
    for q0 in range(0, 256, 8):
        for u in range(1 if narrow else 2):
            lanes = [thread_units(ct, narrow) for ct in range(q0, q0 + 8)]
            for half in (0, 1):          # the two 16-byte loads of a thread's 32-byte segment
                groups = {(raw_addr(fb, rows[u], 2 * cq + half) >> 4) & 7 for fb, cq, rows in lanes}
                assert len(groups) == 8, ("load", q0, u, half)
            for slot in range(stages):
                groups = {(window_addr(fb, cq, 32 * slot + rows[u], block_bytes) >> 4) & 7 for fb, cq, rows in lanes}
                assert len(groups) == 8, ("store", q0, u, slot)


@pytest.mark.parametrize("stages", [4, 5])
def test_window_layout_is_the_probed_mn_major_sw128_layout(stages):
    block_bytes = (32 * stages + 32) * 128
    assert block_bytes % 1024 == 0                       # swizzle atoms (8 rows x 128 B) stay aligned per block
    for ct in range(256):
        fb, cq, rows = thread_units(ct, False)
        for r in rows:
            for slot in range(stages):
                ring_row = 32 * slot + r
                base = window_addr(fb, cq, ring_row, block_bytes)
                for e in range(8):                        # the 8 features of the thread's 16-byte chunk
                    m = 32 * fb + 8 * cq + e
                    want = (m // 64) * block_bytes + ring_row * 128 + ((((m % 64) // 8) ^ (ring_row & 7)) << 4) + (m % 8) * 2
                    assert base + 2 * e == want


def test_lagged_rows_never_straddle_the_wrap():
    # tile t in ring slot s reads rows [32 s + 16 ks + lag, + 16) of the SAME buffer: with the mirror of slot 0 behind
    # the last slot they stay inside the window for every lag <= 32
    for stages in (4, 5):
        win_rows = 32 * stages + 32
        for lag in range(1, 33):
            for s in range(stages):
                for ks in (0, 1):
                    first = 32 * s + 16 * ks + lag
                    assert first + 16 <= win_rows
