"""Index math of the generated-operand layouts, restated in Python and checked exhaustively (no GPU): the kernels write their
MMA operands straight into the UMMA shared-memory layouts, so a wrong offset is silent garbage and a poor thread mapping is a
bank conflict on every store.  Layout (csrc/icnn_tc3.cu, icnn_tc3_dP0_kernel): 32-bit MN-major operand tile of 16 k x 128 MN,
SWIZZLE_128B_BASE32B atoms of 4 k-rows x 128 B, LBO = 512 B between 32-wide MN blocks, SBO = 2 KB between groups of 4 k,
the 32-byte chunks of a row XOR-ed with the k-row."""
import numpy as np


def canonical_offset(k, mn):
    """Byte offset of element (k, mn) in the 16 x 128 MN-major SWIZZLE_128B_BASE32B tile."""
    return (k >> 2) * 2048 + (mn >> 5) * 512 + (k & 3) * 128 + ((((mn >> 3) & 3) ^ (k & 3)) << 5) + (mn & 7) * 4


def generator_thread(warp, lane):
    """(sample k of the stage, first of 4 consecutive MN elements, byte offset) as icnn_tc3_dP0_kernel computes them."""
    g4, hf, kr = lane >> 3, (lane >> 2) & 1, lane & 3
    ks = g4 * 4 + kr
    blk, qd = warp >> 2, warp & 3
    off = (g4 * 4 + blk) * 512 + kr * 128 + ((qd ^ kr) << 5) + hf * 16
    return ks, blk * 32 + qd * 8 + hf * 4, off


def test_dp0_generator_offsets_are_the_canonical_layout_and_cover_the_tile_once():
    seen = np.zeros(16 * 128, dtype=np.int32)
    for warp in range(16):
        for lane in range(32):
            k, mn0, off = generator_thread(warp, lane)
            assert off % 16 == 0 and off == canonical_offset(k, mn0)
            for e in range(4):                                  # the float4 store covers 4 consecutive MN elements
                assert canonical_offset(k, mn0 + e) == off + 4 * e
                seen[k * 128 + mn0 + e] += 1
    assert (seen == 1).all()                                    # 512 threads x 4 elements = the whole 16 x 128 tile, once
    offs = sorted(generator_thread(w, l)[2] for w in range(16) for l in range(32))
    assert offs == list(range(0, 8192, 16))                     # ... and the whole 8 KB of it


def test_dp0_generator_stores_are_bank_conflict_free():
    """A 16-byte store is served a quarter warp at a time: its 8 lanes must hit 8 distinct 16-byte bank groups (128 B)."""
    for warp in range(16):
        for quarter in range(4):
            groups = {(generator_thread(warp, quarter * 8 + i)[2] % 128) // 16 for i in range(8)}
            assert len(groups) == 8


def test_k_major_sw64_offsets_cover_the_tile_once():
    """K-major operand tiles of the forward / backward kernels (tc_common.cuh sw64_off): 64-byte rows, 16-byte chunk c of
    row r at r*64 + ((c ^ ((r >> 1) & 3)) << 4); generator thread (rb, c) of icnn_tc3_fwd_kernel writes rows rb + 32 j."""
    sw64 = lambda r, c: r * 64 + ((c ^ ((r >> 1) & 3)) << 4)
    offs = sorted(sw64(r, c) for r in range(128) for c in range(4))
    assert offs == list(range(0, 128 * 64, 16))
    for rb in range(32):                                        # the kernel adds j * 2048 to the offset of (rb, c) ...
        for c in range(4):
            for j in range(4):
                assert sw64(rb + 32 * j, c) == sw64(rb, c) + j * 2048       # ... valid because 32 rows = one swizzle period
    for q in range(4):                                          # quarter warp = 2 rows x 4 chunks: 128 distinct bytes
        lanes = [(2 * q + (i >> 2), i & 3) for i in range(8)]   # t2 & 3 = chunk, (t2 >> 2) = row
        assert len({sw64(r, c) % 128 // 16 for r, c in lanes}) == 8
