"""Geometry of the chunked panel layout of the large window (include/cngp.h cngp_large_plan::chunk_blocks,
cngp_large_panel_chunks): a host-only function of libcngp.so - no GPU needed.  The chunks of a panel must tile its payload
(the rows below the diagonal block, 32 k-tiles wide) without gaps, start at fixed absolute row positions, and the first
one must hold the rows of the next diagonal block."""
import ctypes as C

import pytest

from corenav_gp_b200 import _lib as L

NB, TILE, BLK = 256, 8, 16          # block column width, tile edge, row block of the GEMM in tiles


def chunks(plan, k):
    first, count = C.c_int32(), C.c_int32()
    offs = (C.c_int64 * (L.LARGE_MAX_CHUNKS + 1))()
    rc = L.load().cngp_large_panel_chunks(C.byref(plan), k, C.byref(first), C.byref(count), offs)
    assert rc == 0
    return first.value, [offs[i] for i in range(count.value + 1)]


def make_plan(N, world, rank, chunk_blocks):
    p = L.LargePlan()
    assert L.load().cngp_large_make_plan(N, world, rank, C.byref(p)) == 0
    p.chunk_blocks = chunk_blocks
    return p


@pytest.mark.parametrize("N,cs", [(32768, 32), (32768, 18), (5000, 6), (5000, 4), (1300, 2), (2048, 0)])
def test_chunks_tile_the_payload(N, cs):
    p = make_plan(N, 8, 0, cs)
    nt = NB // TILE
    for k in range(int(p.n_blockcols)):
        first, offs = chunks(p, k)
        r0 = (k + 1) * nt
        rows = int(p.row_tiles) - r0
        assert offs[0] == 0 and offs[-1] == nt * rows * 64                 # the whole payload, nothing else
        sizes = [b - a for a, b in zip(offs[:-1], offs[1:])]
        assert all(s > 0 and s % (nt * BLK * 64) == 0 for s in sizes)       # whole 16-tile row blocks
        if cs == 0:
            assert first == 0 and len(sizes) == 1
            continue
        # chunk c covers the absolute row blocks [c cs, (c+1) cs) (the last one takes the remainder), clipped at r0
        n_abs = max(1, (int(p.row_tiles) // BLK) // cs)
        assert len(sizes) <= L.LARGE_MAX_CHUNKS and first + len(sizes) == n_abs
        assert first == min(r0 // BLK // cs, n_abs - 1)
        lo = r0
        for i, s in enumerate(sizes):
            c = first + i
            hi = int(p.row_tiles) if c == n_abs - 1 else (c + 1) * cs * BLK
            assert s == nt * (hi - lo) * 64
            lo = hi
        if k + 1 < p.n_blockcols:
            assert sizes[0] >= nt * nt * 64                                 # the next diagonal block's rows come first


def test_chunks_bad_arguments():
    p = make_plan(2048, 2, 0, 2)
    first, count = C.c_int32(), C.c_int32()
    offs = (C.c_int64 * (L.LARGE_MAX_CHUNKS + 1))()
    lib = L.load()
    assert lib.cngp_large_panel_chunks(C.byref(p), -1, C.byref(first), C.byref(count), offs) != 0
    assert lib.cngp_large_panel_chunks(C.byref(p), int(p.n_blockcols), C.byref(first), C.byref(count), offs) != 0
    q = make_plan(32768, 8, 0, 2)      # 257 row blocks / 2 = 128 chunks: more than the library supports
    assert lib.cngp_large_panel_chunks(C.byref(q), 0, C.byref(first), C.byref(count), offs) != 0
