"""NumPy statement of the shared-memory layout of gemm_tma.cu (csrc): where the TMA unit puts element (row, k) of a
64 x 16 operand tile with the 128-byte swizzle, which 16-byte chunk lane (gq, t) of a warp loads for its fragments, and
how many shared-memory wavefronts those loads need.

  k-fast tile   [row][16 k]         byte = row * 128 + (((k >> 1) ^ (row & 7)) << 4) + ((k & 1) << 3)
  row-fast tile [row / 16][k][16]   byte = ((row >> 4) * 16 + k) * 128 + ((((row & 15) >> 1) ^ (k & 7)) << 4) + ((row & 1) << 3)
  (the swizzle XORs the 16-byte chunk index, address bits 4..6, with the 128-byte line index mod 8, bits 7..9)

Lane (gq, t) = (lane >> 2, lane & 3) feeds DMMA step kk (0..3) with k = 8 (kk >> 1) + 4 (t >> 1) + 2 (t & 1) + (kk & 1) and owns
  k-fast:   row(blk, gq) = 8 blk + 4 (gq & 1) + (gq >> 1);   one LDS.128 at chunk (4 p + t) ^ (row & 7) -> steps 2p, 2p + 1
  row-fast: row(blk, gq) = 16 (blk >> 1) + 2 gq + (blk & 1);  one LDS.128 at chunk gq ^ (k & 7) of line (row >> 4) * 16 + k
                                                              -> row blocks 2q, 2q + 1
128-bit shared loads are served per quarter warp (8 lanes): conflict-free means 8 different chunks (mod 8 lines) per quarter."""
import numpy as np


def kmap(kk, t):
    return 8 * (kk >> 1) + 4 * (t >> 1) + 2 * (t & 1) + (kk & 1)


def byte_kfast(row, k):
    return row * 128 + (((k >> 1) ^ (row & 7)) << 4) + ((k & 1) << 3)


def byte_rowfast(row, k):
    return ((row >> 4) * 16 + k) * 128 + ((((row & 15) >> 1) ^ (k & 7)) << 4) + ((row & 1) << 3)


def tile_image(tile, kfast):
    """What the TMA boxes leave in shared memory for a 64 x 16 tile (float64 view of the 8 KB stage half)."""
    img = np.full(64 * 16, np.nan)
    for r in range(64):
        for k in range(16):
            b = byte_kfast(r, k) if kfast else byte_rowfast(r, k)
            img[b // 8] = tile[r, k]
    return img


def warp_fragments(img, kfast, w0):
    """Fragments of one warp (rows w0 .. w0 + 31 of the tile): frag[blk, kk, lane] and the chunk addresses of every load."""
    frag = np.zeros((4, 4, 32))
    loads = []                                   # (list of 32 byte addresses) per LDS.128
    if kfast:
        for blk in range(4):
            for p in range(2):
                addr = []
                for lane in range(32):
                    gq, t = lane >> 2, lane & 3
                    row = w0 + 8 * blk + 4 * (gq & 1) + (gq >> 1)
                    a = row * 128 + (((4 * p + t) ^ (row & 7)) << 4)
                    addr.append(a)
                    frag[blk, 2 * p, lane], frag[blk, 2 * p + 1, lane] = img[a // 8], img[a // 8 + 1]
                loads.append(addr)
    else:
        for q in range(2):
            for kk in range(4):
                addr = []
                for lane in range(32):
                    gq, t = lane >> 2, lane & 3
                    k = kmap(kk, t)
                    a = (((w0 >> 4) + q) * 16 + k) * 128 + ((gq ^ (k & 7)) << 4)
                    addr.append(a)
                    frag[2 * q, kk, lane], frag[2 * q + 1, kk, lane] = img[a // 8], img[a // 8 + 1]
                loads.append(addr)
    return frag, loads


def rowmap(kfast, w0, blk, gq):
    return w0 + 8 * blk + 4 * (gq & 1) + (gq >> 1) if kfast else w0 + 16 * (blk >> 1) + 2 * gq + (blk & 1)


def wavefronts_128(addr):
    """Wavefronts of one LDS.128: per quarter warp, the largest number of lanes that hit the same 16-byte bank group with
    different addresses (8 bank groups of 16 bytes per 128-byte line)."""
    total = 0
    for qw in range(4):
        lanes = addr[8 * qw: 8 * qw + 8]
        groups = {}
        for a in lanes:
            groups.setdefault((a >> 4) & 7, set()).add(a)
        total += max(len(v) for v in groups.values())
    return total
