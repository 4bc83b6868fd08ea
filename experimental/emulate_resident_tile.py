"""Desk check of the resident-tile index arithmetic of experimental/resident_dense_block.patch (numpy, no GPU).

Hypothesis under test on the hardware (scripts/probe_sbo.sh): TMA SWIZZLE_128B and the K-major UMMA descriptor both XOR the
16-byte piece index with bits [7:9] of the ABSOLUTE shared-memory address.  Given that hypothesis this script checks that
  1. the epilogue's tile writes (own pixels + halo columns pushed into the neighbour strips) reproduce exactly the bytes a TMA
     load of the finished image would have put into every strip's tile, and
  2. the operand window of every (M tile, ky, kx, K step) -- start = ((m*TH + ky)*pitch + kx)*128 + 32*k, 16 groups of 8 rows at
     SBO = pitch*128 -- fetches the channels of the pixels a 3x3 conv needs (zero outside the image).
"""
import numpy as np

TW, TH, MT, PITCH, H, W, C = 8, 16, 2, 10, 32, 32, 64
ROWS = MT * TH + 2
CHUNK = PITCH * ROWS * 128
rng = np.random.default_rng(0)
img = rng.integers(1, 2 ** 15, (H, W, C), dtype=np.uint16)          # stand-in for bf16 bit patterns
n_strips = W // TW


def px(y, x):
    return img[y, x] if 0 <= y < H and 0 <= x < W else np.zeros(C, np.uint16)


def store_row(tile, row, vec64):
    """one 128-byte row, 16-byte pieces XOR (row & 7): what TMA SWIZZLE_128B writes at a 1 KB aligned base"""
    b = vec64.view(np.uint8).reshape(8, 16)
    for j in range(8):
        p = j ^ (row & 7)
        tile[row * 128 + p * 16: row * 128 + p * 16 + 16] = b[j]


def tma_tile(rank):
    t = np.zeros(CHUNK, np.uint8)
    for yb in range(ROWS):
        for xb in range(PITCH):
            store_row(t, yb * PITCH + xb, px(yb - 1, rank * TW + xb - 1).copy())
    return t


def epilogue_tiles():
    """the kernel's write formulas: own pixel at row (yl+1)*pitch + txx+1; edge columns pushed to the neighbours' halo columns"""
    tiles = [np.zeros(CHUNK, np.uint8) for _ in range(n_strips)]

    def put(tile, row, c0, vals16):       # 16 channels = two 16-byte pieces j0, j0+1 at (j ^ (row & 7))
        j0 = (c0 & 63) >> 3
        b = vals16.view(np.uint8).reshape(2, 16)
        for d in range(2):
            p = (j0 + d) ^ (row & 7)
            tile[row * 128 + p * 16: row * 128 + p * 16 + 16] = b[d]

    for rank in range(n_strips):
        for mt in range(MT):
            for m in range(128):
                tyy, txx = m // TW, m % TW
                yl = mt * TH + tyy
                for c0 in range(0, C, 16):
                    v = img[yl, rank * TW + txx, c0:c0 + 16].copy()
                    put(tiles[rank], (yl + 1) * PITCH + txx + 1, c0, v)
                    if txx == 0 and rank > 0:
                        put(tiles[rank - 1], (yl + 1) * PITCH + TW + 1, c0, v)
                    elif txx == TW - 1 and rank + 1 < n_strips:
                        put(tiles[rank + 1], (yl + 1) * PITCH, c0, v)
    return tiles


def umma_fetch(tile, start, sbo, i):
    """row i (0..127) of the K-major SW128 operand window: 32 bytes = one K step, address-based swizzle"""
    addr = start + (i // 8) * sbo + (i % 8) * 128
    out = np.empty(32, np.uint8)
    for d in range(2):
        a = addr + 16 * d
        phys = a ^ (((a >> 7) & 7) << 4)
        out[16 * d:16 * d + 16] = tile[phys:phys + 16]
    return out.view(np.uint16)


def main():
    ep = epilogue_tiles()
    for rank in range(n_strips):
        t = tma_tile(rank)
        # the epilogue never writes the image-border halo (stays zero) nor the top/bottom halo rows: identical everywhere
        assert np.array_equal(t, ep[rank]), f"strip {rank}: epilogue tile differs from the TMA layout"
    print("1. epilogue + halo pushes == TMA layout for all", n_strips, "strips")
    for rank in range(n_strips):
        t = tma_tile(rank)
        for m in range(MT):
            for ky in range(3):
                for kx in range(3):
                    for k in range(4):
                        start = ((m * TH + ky) * PITCH + kx) * 128 + 32 * k
                        for i in range(128):
                            y = m * TH + i // 8 + ky - 1
                            x = rank * TW + i % 8 + kx - 1
                            got = umma_fetch(t, start, PITCH * 128, i)
                            assert np.array_equal(got, px(y, x)[16 * k:16 * k + 16]), (rank, m, ky, kx, k, i)
    print("2. operand windows (SBO = %d B) fetch the right pixels / channels for all taps" % (PITCH * 128))
    # the 32-channel dY slot at tile channel 96 (chunk 1, bytes 64..127): K steps 2, 3 of the same rows
    print("3. a slot at channel offset 32 of a chunk = K steps 2..3: start += 2 * 2 (16-byte units)  [same rows, covered by 2.]")


if __name__ == "__main__":
    main()
