// FP64 tensor-core GEMM (DMMA m8n8k4), TMA-staged: the contraction kernel of the CTMRG / TRG step shapes.
//
// Same contract as gemm.cu (multi-level operand strides, "permute-on-load", alpha / beta epilogue), different machinery:
//   * every operand is described by ONE tensor map (cuTensorMapEncodeTiled, <= 5 dimensions: the levels of its row group,
//     of the K group and of the batch group); a 64 x 16 operand tile is fetched by 1 .. 4 `cp.async.bulk.tensor` boxes of
//     16 doubles (= 128 bytes, the swizzle span) x RB rows, so the index permutation happens inside the TMA unit and no
//     thread computes a global address in the main loop;
//   * tiles land in shared memory densely with the 128-byte hardware swizzle; the k index a lane feeds to DMMA step kk is
//     permuted (k = 8 (kk / 2) + 4 (t / 2) + 2 (kk % 2) + t % 2) so that the fragment loads of both layouts (k-fast and
//     row-fast) need the minimum of two wavefronts;
//   * one producer warp (one elected lane issues the boxes and arms the `full` mbarrier of the stage with the byte count),
//     four consumer warps (32 x 32 warp tiles, accumulators in registers: sm_100 has no f64 kind for tcgen05 / TMEM),
//     `empty` mbarriers hand the stage back; 4 stages per consumer group;
//   * persistent CTAs (one per SM) with THREE consumer groups, each with its own producer warp and stage ring, working on
//     64 x 64 tiles a third of a tile out of phase: the epilogue (global stores) of one group runs under the main loop of the
//     other (with independent CTAs all tiles of a wave reach their epilogue together and the tensor pipe idles: 58 %
//     pipe-active in ncu); the tile count of the step's shapes (1024 at d = 4, chi = 128) quantises to 148 SMs at 99 %
//     (128 x 128 tiles: 86 %);
//   * split-K without a second launch: partial tiles go to a workspace, the CTA that arrives last at the tile's counter
//     adds them up in split order (deterministic) and writes C.
// Operands the tensor maps cannot describe (odd strides, extents that do not tile, more than 5 dimensions) fall back to
// the cp.async kernel of gemm.cu.
#include "common.h"
#include <cuda.h>
#include <algorithm>
#include <mutex>

namespace tnad {

namespace {

constexpr int TM = 64, TN = 64, TK = 16;
constexpr int NSTAGE = 4;                      // stages per consumer group
constexpr int NGRP = 3;                        // consumer groups per CTA, one producer warp each
constexpr int NCONS = 4;                       // consumer warps per group (2 x 2 warp tiles of 32 x 32)
constexpr int NTHREADS = 32 * NGRP * (NCONS + 1);
constexpr int TILE_BYTES = TM * TK * 8;        // one operand tile of one stage
constexpr int MAXOPS = 4;                      // boxes per operand tile
constexpr int CNT_SLOTS = 1 << 15;             // split-K tile counters per context

struct TmaOp {
  int kfast;               // shared-memory layout: 1 = [row][16 k], 0 = [row / 16][k][16 rows]
  int rb, nops;            // rows per box, boxes per tile
  int nr, rn[MAXL], rd[MAXL];   // row levels: extent, tensor-map dimension
  int nk, kn[MAXL], kd[MAXL];   // K levels
  int nb, bn[MAXL], bd[MAXL];   // batch levels
};

struct TmaGemm {
  int M, N, K, batch, splitk;
  int tm, tn, units;       // tiles along M and N, work units (tiles x splits)
  int dbg_nofetch;         // TNAD_GEMM_NOFETCH=1: only the first ring of stages is fetched (results are wrong; timing aid)
  LvlSet cm, cn, cb;
  double* C;
  double alpha, beta;
  double* ws;
  int* cnt;
  TmaOp a, b;
};

__device__ __forceinline__ long long lvl_off(const LvlSet& L, int i) {
  long long o = 0;
#pragma unroll
  for (int l = 0; l < MAXL; ++l) {
    if (l < L.nl) {
      if (l == L.nl - 1) {
        o += (long long)i * L.s[l];
      } else {
        int q = i / L.n[l];
        o += (long long)(i - q * L.n[l]) * L.s[l];
        i = q;
      }
    }
  }
  return o;
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

__device__ __forceinline__ void mbar_init(unsigned bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// false after about a second of polling: the caller traps instead of hanging the device
__device__ __forceinline__ bool mbar_wait(unsigned bar, unsigned parity) {
  for (unsigned spins = 0; spins < (1u << 22); ++spins) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ void tma_load_5d(unsigned dst, const CUtensorMap* map, unsigned bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// coordinate `q` of a level goes to tensor-map dimension `dim` (register-only scatter: no indexed local array)
__device__ __forceinline__ void put_coord(int dim, int q, int& c0, int& c1, int& c2, int& c3, int& c4) {
  c0 += dim == 0 ? q : 0;
  c1 += dim == 1 ? q : 0;
  c2 += dim == 2 ? q : 0;
  c3 += dim == 3 ? q : 0;
  c4 += dim == 4 ? q : 0;
}

// unit u of the launch -> (tile, split) -> block coordinates
struct Unit {
  int m0, n0, bz, sp, tile, kt0, nkt;
};
__device__ __forceinline__ Unit unit_of(const TmaGemm& g, int u) {
  Unit r;
  const int S = g.splitk > 1 ? g.splitk : 1;
  r.tile = u / S;
  r.sp = u - r.tile * S;
  const int bx = r.tile % g.tm, rest = r.tile / g.tm;
  const int by = rest % g.tn;
  r.bz = rest / g.tn;
  r.m0 = bx * TM;
  r.n0 = by * TN;
  const int nkt_all = (g.K + TK - 1) / TK;
  const int per = (nkt_all + S - 1) / S;
  r.kt0 = r.sp * per;
  r.nkt = max(0, min(per, nkt_all - r.kt0));
  return r;
}

// Persistent CTA (one per SM): NGRP consumer groups of four warps, each with its own ring of NSTAGE stages and its own
// producer warp.  CTA b works on the units b, b + G, b + 2G, ... (G = grid size); group g takes every NGRP-th one of that
// list starting with the g-th, and starts 1 / NGRP of a tile after group g - 1, so that the epilogue (global stores) of one
// group runs under the main loops of the others and the FP64 tensor pipe always has a main loop to serve.
template <bool AKF, bool BKF>
__global__ void __launch_bounds__(NTHREADS, 1)
    gemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                    const __grid_constant__ TmaGemm g) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 1024-byte alignment of the stage buffers (the swizzle pattern is a function of the shared-memory address)
  const unsigned sbase0 = (unsigned)__cvta_generic_to_shared(smem_raw);
  const unsigned sbase = (sbase0 + 1023u) & ~1023u;
  unsigned char* sgen = smem_raw + (sbase - sbase0);
  const unsigned bars = sbase + NGRP * NSTAGE * 2 * TILE_BYTES;   // per group: full[NSTAGE], empty[NSTAGE]; then the stagger barrier
  const unsigned stag = bars + 8 * NGRP * 2 * NSTAGE;           // NGRP - 1 stagger barriers: group g + 1 waits for group g
  __shared__ int s_last[NGRP];
  __shared__ long long s_off[NGRP][TM + TN];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = g.splitk > 1 ? g.splitk : 1;
  const int G = gridDim.x;

  if (tid == 0) {
    for (int q = 0; q < NGRP; ++q)
      for (int s = 0; s < NSTAGE; ++s) {
        mbar_init(bars + 8 * (q * 2 * NSTAGE + s), 1);
        mbar_init(bars + 8 * (q * 2 * NSTAGE + NSTAGE + s), NCONS);
      }
    for (int q = 0; q + 1 < NGRP; ++q) mbar_init(stag + 8 * q, NCONS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // programmatic dependent launch: the grid is launched while its predecessor in the stream drains (the launch latency
  // and the set-up above overlap that tail); nothing before this point touches global memory
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const bool producer = warp >= NGRP * NCONS;
  const int grp = producer ? warp - NGRP * NCONS : warp / NCONS;
  const unsigned ring = sbase + grp * NSTAGE * 2 * TILE_BYTES;
  const unsigned full0 = bars + 8 * (grp * 2 * NSTAGE), empty0 = full0 + 8 * NSTAGE;
  const int u0 = blockIdx.x + grp * G, ustep = NGRP * G;

  if (producer) {
    // ------------------------------------------------ producer ------------------------------------------------
    // Lane l issues box l of the stage (A boxes first, then B boxes).  Everything that needs a division -- the batch
    // index, the row origin of the box, the first k index -- is decomposed over the levels once per unit; per k-tile
    // the K coordinates advance like an odometer.
    const bool isA = lane < g.a.nops;
    const bool act = lane < g.a.nops + g.b.nops;
    const TmaOp& o = isA ? g.a : g.b;
    const CUtensorMap* mp = isA ? &mapA : &mapB;
    const int opi = isA ? lane : lane - g.a.nops;
    const unsigned dst_off = (isA ? 0 : TILE_BYTES) + opi * o.rb * TK * 8;
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    }
    int it = 0;                                      // stages filled so far (all units)
    for (int u = u0; u < g.units; u += ustep) {
      const Unit un = unit_of(g, u);
      int f0 = 0, f1 = 0, f2 = 0, f3 = 0, f4 = 0;    // batch + row part of the coordinates
      int kq[MAXL] = {0, 0, 0, 0};
      if (act) {
        int i = un.bz;
#pragma unroll
        for (int l = 0; l < MAXL; ++l)
          if (l < o.nb) {
            const int q = (l == o.nb - 1) ? i : i % o.bn[l];
            i /= o.bn[l];
            put_coord(o.bd[l], q, f0, f1, f2, f3, f4);
          }
        i = (isA ? un.m0 : un.n0) + opi * o.rb;
#pragma unroll
        for (int l = 0; l < MAXL; ++l)
          if (l < o.nr) {
            const int q = (l == o.nr - 1) ? i : i % o.rn[l];
            i /= o.rn[l];
            put_coord(o.rd[l], q, f0, f1, f2, f3, f4);
          }
        i = un.kt0 * TK;
#pragma unroll
        for (int l = 0; l < MAXL; ++l)
          if (l < o.nk) {
            kq[l] = (l == o.nk - 1) ? i : i % o.kn[l];
            i /= o.kn[l];
          }
      }
      for (int kt = 0; kt < un.nkt; ++kt, ++it) {
        const int s = it % NSTAGE;
        if (it >= NSTAGE && !mbar_wait(empty0 + 8 * s, (unsigned)((it / NSTAGE - 1) & 1))) __trap();
        const unsigned full = full0 + 8 * s;
        const bool fetch = !(g.dbg_nofetch && it >= NSTAGE);
        if (lane == 0) {
          if (fetch) mbar_expect_tx(full, 2 * TILE_BYTES);
          else mbar_arrive(full);                   // measurement aid: consumers run on stale stages
        }
        __syncwarp();
        if (act && fetch) {
          int c0 = f0, c1 = f1, c2 = f2, c3 = f3, c4 = f4;
#pragma unroll
          for (int l = 0; l < MAXL; ++l)
            if (l < o.nk) put_coord(o.kd[l], kq[l], c0, c1, c2, c3, c4);
          tma_load_5d(ring + s * 2 * TILE_BYTES + dst_off, mp, full, c0, c1, c2, c3, c4);
          kq[0] += TK;                               // next k-tile
#pragma unroll
          for (int l = 0; l + 1 < MAXL; ++l)
            if (l + 1 < o.nk && kq[l] >= o.kn[l]) {
              kq[l] = 0;
              kq[l + 1] += 1;
            }
        }
      }
    }
    return;
  }

  // -------------------------------------------------- consumers -------------------------------------------------
  const int cw = warp - grp * NCONS;                 // warp within the group
  const int gq = lane >> 2, t = lane & 3;
  const int wm0 = (cw & 1) * 32, wn0 = (cw >> 1) * 32;
  const unsigned char* ringg = sgen + grp * NSTAGE * 2 * TILE_BYTES;
  // Fragment addressing.  Lane (gq, t) feeds DMMA step kk with k = 8 (kk / 2) + 4 (t / 2) + 2 (t % 2) + kk % 2, and the
  // eight rows of a fragment block are assigned to the lanes so that ONE 16-byte load brings two fragments and the eight
  // lanes of every quarter warp hit eight different 16-byte chunks (the swizzle is chunk ^= line % 8):
  //   k-fast tile  [row][16 k]:        row(blk, gq) = 8 blk + 4 (gq % 2) + gq / 2; the load at chunk (4 p + t) ^ (row % 8) of
  //                                    the row holds k-steps 2p and 2p + 1 of that row
  //   row-fast tile [row / 16][k][16]: row(blk, gq) = 16 (blk / 2) + 2 gq + blk % 2; the load at chunk gq ^ (k % 8) of line
  //                                    (row / 16) * 16 + k holds the rows of blocks 2q and 2q + 1
  // (64-bit shared loads are served per half warp, 128-bit loads per quarter warp: the round-2 layout that was
  // conflict-free over a full warp needed 4 wavefronts per LDS.64, ncu l1tex__data_bank_conflicts = 2 per load.)
  int offa[8], offb[8];
#pragma unroll
  for (int x = 0; x < 8; ++x) {
    {
      const int blk = x >> 1, p = x & 1;                         // k-fast: (row block, pair of k-steps)
      const int ra = wm0 + 8 * blk + 4 * (gq & 1) + (gq >> 1), rb = wn0 + 8 * blk + 4 * (gq & 1) + (gq >> 1);
      const int ka = ra * 128 + (((4 * p + t) ^ (ra & 7)) << 4), kb = rb * 128 + (((4 * p + t) ^ (rb & 7)) << 4);
      const int q = x >> 2, kk = x & 3;                          // row-fast: (pair of row blocks, k-step)
      const int k = 8 * (kk >> 1) + 4 * (t >> 1) + 2 * (t & 1) + (kk & 1);
      const int ma = (((wm0 >> 4) + q) * 16 + k) * 128 + ((gq ^ (k & 7)) << 4);
      const int mb = (((wn0 >> 4) + q) * 16 + k) * 128 + ((gq ^ (k & 7)) << 4);
      offa[x] = AKF ? ka : ma;
      offb[x] = BKF ? kb : mb;
    }
  }
  // tile-local row of (block i, lane group gq) / column of (block j, c = 2 t + e)
  auto rowmap = [&](int i, int gg) { return AKF ? wm0 + 8 * i + 4 * (gg & 1) + (gg >> 1) : wm0 + 16 * (i >> 1) + 2 * gg + (i & 1); };
  auto colmap = [&](int j, int cc) { return BKF ? wn0 + 8 * j + 4 * (cc & 1) + (cc >> 1) : wn0 + 16 * (j >> 1) + 2 * cc + (j & 1); };
  const double alpha = g.alpha, beta = g.beta;
  int it = 0;
  bool first = true;
  for (int u = u0; u < g.units; u += ustep) {
    const Unit un = unit_of(g, u);
    if (first && grp >= 1 && !mbar_wait(stag + 8 * (grp - 1), 0)) __trap();   // 1 / NGRP of a tile behind the group before
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int kt = 0; kt < un.nkt; ++kt, ++it) {
      const int s = it % NSTAGE;
      if (!mbar_wait(full0 + 8 * s, (unsigned)((it / NSTAGE) & 1))) __trap();
      const unsigned char* as = ringg + s * 2 * TILE_BYTES;
      const unsigned char* bs = as + TILE_BYTES;
      // half p of the stage = k-steps 2p, 2p + 1: fragments f[half][block][step]
      double af[2][4][2], bf[2][4][2];
      auto load_half = [&](int p, double (&fa)[4][2], double (&fb)[4][2]) {
        if (AKF) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const double2 v = *reinterpret_cast<const double2*>(as + offa[2 * i + p]);
            fa[i][0] = v.x;
            fa[i][1] = v.y;
          }
        } else {
#pragma unroll
          for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const double2 v = *reinterpret_cast<const double2*>(as + offa[4 * q + 2 * p + h]);
              fa[2 * q][h] = v.x;
              fa[2 * q + 1][h] = v.y;
            }
        }
        if (BKF) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const double2 v = *reinterpret_cast<const double2*>(bs + offb[2 * j + p]);
            fb[j][0] = v.x;
            fb[j][1] = v.y;
          }
        } else {
#pragma unroll
          for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const double2 v = *reinterpret_cast<const double2*>(bs + offb[4 * q + 2 * p + h]);
              fb[2 * q][h] = v.x;
              fb[2 * q + 1][h] = v.y;
            }
        }
      };
      load_half(0, af[0], bf[0]);
      load_half(1, af[1], bf[1]);                  // in flight under the 32 DMMAs of the first half
#pragma unroll
      for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], af[p][i][h], bf[p][j][h]);
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(empty0 + 8 * s);
        if (first && grp + 1 < NGRP && kt == un.nkt / NGRP) mbar_arrive(stag + 8 * grp);
      }
    }
    first = false;

    // ------------------------------------------------ epilogue ------------------------------------------------
    // thread holds C[m0 + rowmap(i, gq)][n0 + colmap(j, 2 t + {0, 1})]
    if (S > 1) {
      // partial tile -> workspace [tile][split][64 x 64, m fastest]; the last arriver sums in split order
      double* wt = g.ws + (long long)un.tile * S * (TM * TN);
      double* w = wt + (long long)un.sp * (TM * TN);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) __stcg(w + rowmap(i, gq) + TM * colmap(j, 2 * t + e), acc[i][j][e]);
      __threadfence();
      asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(32 * NCONS) : "memory");
      if (cw == 0 && lane == 0) {
        const int prev = atomicAdd(g.cnt + un.tile, 1);
        s_last[grp] = prev == S - 1;
        if (prev == S - 1) g.cnt[un.tile] = 0;     // ready for the next launch (stream order)
      }
      asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(32 * NCONS) : "memory");
      const bool last = s_last[grp] != 0;
      asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(32 * NCONS) : "memory");   // s_last may be rewritten by the next unit
      if (!last) continue;
      __threadfence();
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
      // two partial tiles in flight per round trip to L2; the order of the additions is fixed (deterministic)
      for (int s2 = 0; s2 < S; s2 += 2) {
        const double* w2 = wt + (long long)s2 * (TM * TN);
        const bool two = s2 + 1 < S;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          double v0[4][2], v1[4][2];
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int off = rowmap(i, gq) + TM * colmap(j, 2 * t + e);
              v0[j][e] = __ldcg(w2 + off);
              v1[j][e] = two ? __ldcg(w2 + TM * TN + off) : 0.0;
            }
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) acc[i][j][e] = (acc[i][j][e] + v0[j][e]) + v1[j][e];
        }
      }
    }
    // Offsets of the tile's 64 rows and 64 columns of C: one level decomposition per thread (the level sets sit in the
    // constant bank), shared through a table.  Thirteen out-of-line calls per thread that read the level sets through
    // generic pointers cost 6 us per tile here (ncu: long-scoreboard stalls on LD.E), more than the main loop at K = 128.
    asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(32 * NCONS) : "memory");     // the previous tile's table is no longer read
    {
      const int gt = cw * 32 + lane;
      if (gt < TM) {
        const int m = un.m0 + gt;
        s_off[grp][gt] = m < g.M ? lvl_off(g.cm, m) : -1;
      } else {
        const int n = un.n0 + gt - TM;
        s_off[grp][gt] = n < g.N ? lvl_off(g.cn, n) : -1;
      }
    }
    asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(32 * NCONS) : "memory");
    double* gC = g.C + lvl_off(g.cb, un.bz);
    long long coff[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) coff[j][e] = s_off[grp][TM + colmap(j, 2 * t + e)];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long ro = s_off[grp][rowmap(i, gq)];
      if (ro >= 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e)
            if (coff[j][e] >= 0) {
              double* pc = gC + ro + coff[j][e];
              double v = alpha * acc[i][j][e];
              if (beta != 0.0) v += beta * *pc;
              *pc = v;
            }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side: tensor maps
// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  });
  return fn;
}

// One operand: row group `rows` (M of A / N of B), K group `ks`, batch group `bat`; tile of TR rows x 16 k.
bool plan_operand(const double* base, const LvlSet& rows, const LvlSet& ks, const LvlSet& bat, int batch, bool kfast,
                  int TR, CUtensorMap* map, TmaOp* op) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  if (reinterpret_cast<uintptr_t>(base) & 15) return false;
  const int nr = rows.nl, nk = ks.nl, nb = batch > 1 ? bat.nl : 0;
  if (nr + nk + nb > 5) return false;
  memset(op, 0, sizeof(*op));
  op->kfast = kfast ? 1 : 0;
  op->nr = nr;
  op->nk = nk;
  op->nb = nb;
  // the contiguous level comes first, the level it is paired with in the box second
  if ((kfast ? ks : rows).s[0] != 1) return false;
  // every K level but the last must hold whole k-tiles
  if (nk > 1 && ks.n[0] % TK != 0) return false;
  // rows per box
  int rbx;
  if (kfast) {
    if (nr == 1 || rows.n[0] % TR == 0) rbx = TR;
    else if (rows.n[0] < TR && TR % rows.n[0] == 0 && rows.n[0] % 8 == 0) rbx = rows.n[0];
    else return false;
  } else {
    if (nr > 1 && rows.n[0] % 16 != 0) return false;
    rbx = 16;
  }
  // a box of consecutive rows sits inside level 0 (rbx divides n[0]); the higher levels only see the quotient
  op->rb = rbx;
  op->nops = TR / rbx;
  if (op->nops > MAXOPS) return false;
  cuuint64_t gdim[5] = {1, 1, 1, 1, 1};
  cuuint64_t gstr[5] = {8, 16, 16, 16, 16};     // bytes; [0] is implicit
  cuuint32_t box[5] = {1, 1, 1, 1, 1}, estr[5] = {1, 1, 1, 1, 1};
  int nd = 0;
  auto add = [&](long long n, long long s, int b) -> int {
    gdim[nd] = (cuuint64_t)n;
    gstr[nd] = (cuuint64_t)s * 8;
    box[nd] = (cuuint32_t)b;
    return nd++;
  };
  // dimension 0: the contiguous level, 16 elements per box; dimension 1: level 0 of the other group
  if (kfast) {
    op->kd[0] = add(ks.n[0], 1, TK);
    op->rd[0] = add(rows.n[0], rows.s[0], rbx);
  } else {
    op->rd[0] = add(rows.n[0], 1, 16);
    op->kd[0] = add(ks.n[0], ks.s[0], TK);
  }
  // the rest (box extent 1) in ascending stride order
  struct Rest { long long n, s; int kind, l; };
  std::vector<Rest> rest;
  for (int l = 1; l < nr; ++l) rest.push_back({rows.n[l], rows.s[l], 0, l});
  for (int l = 1; l < nk; ++l) rest.push_back({ks.n[l], ks.s[l], 1, l});
  for (int l = 0; l < nb; ++l) rest.push_back({bat.n[l], bat.s[l], 2, l});
  std::stable_sort(rest.begin(), rest.end(), [](const Rest& x, const Rest& y) { return x.s < y.s; });
  for (const Rest& r : rest) {
    const int dim = add(r.n, r.s, 1);
    if (r.kind == 0) op->rd[r.l] = dim;
    else if (r.kind == 1) op->kd[r.l] = dim;
    else op->bd[r.l] = dim;
  }
  for (int l = 0; l < nr; ++l) op->rn[l] = rows.n[l];
  for (int l = 0; l < nk; ++l) op->kn[l] = ks.n[l];
  for (int l = 0; l < nb; ++l) op->bn[l] = bat.n[l];
  for (int i = 1; i < nd; ++i)
    if ((gstr[i] & 15) || gstr[i] == 0 || gstr[i] >= (1ull << 40)) return false;
  for (int i = 0; i < nd; ++i)
    if (gdim[i] == 0 || gdim[i] > 0xffffffffull || box[i] > 256) return false;
  // pad to rank 5 with unit dimensions (one instruction variant in the kernel)
  cuuint64_t pad = 16;
  for (int i = 1; i < nd; ++i) pad = std::max<cuuint64_t>(pad, gstr[i]);
  for (int i = nd; i < 5; ++i) gstr[i] = pad;
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, const_cast<double*>(base), gdim, gstr + 1, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace

// true when the product was launched on the TMA kernel; false = not eligible (the caller falls back to gemm.cu)
bool gemm_tma_try(tnad_ctx* c, const GemmDesc& d) {
  if (opt_i(c, "TNAD_GEMM_TMA", 1) == 0) return false;
  if (d.M < 16 || d.N < 16 || d.K < 16) return false;
  alignas(64) CUtensorMap mapA, mapB;
  TmaGemm g;
  memset(&g, 0, sizeof(g));
  {
    HostTimer ht(c, 3);
    if (!plan_operand(d.A, d.am, d.ak, d.ab, d.batch, d.a_kfast != 0, TM, &mapA, &g.a)) return false;
    if (!plan_operand(d.B, d.bn, d.bk, d.bb, d.batch, d.b_kfast != 0, TN, &mapB, &g.b)) return false;
  }
  g.M = d.M; g.N = d.N; g.K = d.K; g.batch = d.batch;
  g.cm = d.cm; g.cn = d.cn; g.cb = d.cb;
  g.C = d.C; g.alpha = d.alpha; g.beta = d.beta;
  const long long tm = (d.M + TM - 1) / TM, tn = (d.N + TN - 1) / TN;
  const long long tiles = tm * tn * d.batch;
  const int nkt = (d.K + TK - 1) / TK;
  // split-K when the tiles alone leave SMs idle.  Cost of a split count in k-tile times: every SM runs
  // ceil(units / #SMs) units of `per` k-tiles (+1 for the fill), a lone unit per SM keeps one warp per scheduler only
  // (x 1.15), and the CTA that arrives last adds S partial tiles at the end of the kernel (1.4 each).
  int S = 1;
  if (tiles < 2LL * c->num_sms && nkt >= 12 && tiles <= CNT_SLOTS) {
    double best = 1e300;
    for (int s = 1; s <= 32 && nkt / s >= 6; ++s) {
      const int per = (nkt + s - 1) / s;
      const int se = (nkt + per - 1) / per;      // no empty splits
      const long long units = tiles * se;
      const long long rounds = (units + c->num_sms - 1) / c->num_sms;
      const double cost = (double)rounds * (per * (units <= c->num_sms ? 1.15 : 1.0) + 1.0) + (se > 1 ? 1.4 * se : 0.0);
      if (cost < best) {
        best = cost;
        S = se;
      }
    }
  }
  {
    const int force = opt_i(c, "TNAD_GEMM_SPLITK", 0);   // A/B knob: force the split count (clamped so that no split is empty)
    if (force > 0 && tiles <= CNT_SLOTS) {
      const int per = (nkt + force - 1) / force;
      S = (nkt + per - 1) / per;
    }
  }
  if (tiles * S > 0x7fffffffLL) return false;
  Tens ws;
  if (S > 1) {
    int*& cnt = c->gemm_cnt[c->stream == c->stream2 ? 1 : 0];   // concurrent products on the two streams must not share counters
    if (!cnt) {
      TNAD_CUDA(cudaMalloc((void**)&cnt, CNT_SLOTS * sizeof(int)));
      TNAD_CUDA(cudaMemsetAsync(cnt, 0, CNT_SLOTS * sizeof(int), c->stream));
    }
    ws = t_alloc(c, {(int64_t)tiles * S * TM * TN});
    g.ws = ws.p;
    g.cnt = cnt;
  }
  g.splitk = S;
  g.tm = (int)tm;
  g.tn = (int)tn;
  g.units = (int)(tiles * S);
  g.dbg_nofetch = opt_i(c, "TNAD_GEMM_NOFETCH", 0);
  const size_t smem = (size_t)NGRP * NSTAGE * 2 * TILE_BYTES + (NGRP * 2 * NSTAGE + NGRP) * 8 + 1024;
  void (*kern)(CUtensorMap, CUtensorMap, TmaGemm) =
      g.a.kfast ? (g.b.kfast ? gemm_tma_kernel<true, true> : gemm_tma_kernel<true, false>)
                : (g.b.kfast ? gemm_tma_kernel<false, true> : gemm_tma_kernel<false, false>);
  const int variant = (g.a.kfast ? 2 : 0) + (g.b.kfast ? 1 : 0);
  static std::atomic<unsigned long long> attr_devs[4];
  if (!((attr_devs[variant].load(std::memory_order_acquire) >> (c->device & 63)) & 1ULL)) {
    TNAD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_devs[variant].fetch_or(1ULL << (c->device & 63), std::memory_order_release);
  }
  const int grid = (int)std::min<long long>(c->gemm_grid_cap > 0 ? std::min(c->gemm_grid_cap, c->num_sms) : c->num_sms, g.units);
  KTimer kt(c, KF_GEMM);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = c->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  // dependent launch on the main stream only: on the side stream the early-resident CTAs of the next product (one per SM,
  // 192 KB of shared memory) would sit on SMs the main stream's kernels are waiting for
  const int pdl = opt_i(c, "TNAD_GEMM_PDL", 1);
  cfg.numAttrs = (pdl == 2 || (pdl == 1 && c->stream != c->stream2)) ? 1 : 0;
  {
    HostTimer ht(c, 4);
    TNAD_CUDA(cudaLaunchKernelEx(&cfg, kern, mapA, mapB, g));
  }
  c->launches++;
  TNAD_CUDA(cudaGetLastError());
  return true;
}

}  // namespace tnad
