// Backward for D = 32, L = P = 4 with grad_value combined ON CHIP before it leaves the SM
// ("tile" backward).  Replaces, like bwd_d32_kernel, the reference's col2im kernels
// (/root/reference/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:301-403, bilinear part :87-159).
//
// Why: bwd_d32_kernel sends one 128 B reduction per bilinear corner to L2 (22.8 M rows = 2.9 GB at the
// DETR-encoder shape) and is bound by the SM->L2 port at ~23 B/clk/SM (433 us for the reductions alone,
// profiles/r01_microbench.md) although every grad_value row is hit ~64 times.  Shared-memory
// read-modify-write of 128 B rows costs as much as the reduction it replaces (profiles/r01_microbench_smem.md).
// This kernel turns the scatter around instead:
//
//   a corner's contribution is   coef * grad_out[q, m, :]   (coef = bilinear weight * attention weight)
//
// so per destination row r:  grad_value[r] = sum over (q, coef) in list(r) of coef * G[q], i.e. a sparse
// (rows x queries) matrix times the tile's grad_out rows.  One UNIT = (sample b, tile of 64 queries, head m):
//   P1  one thread per bilinear sample: geometry (sample_geom, the same routine as every other kernel),
//       then per valid corner ONE 4-byte shared-memory counter increment (native ATOMS.ADD) on the
//       destination row's slot in a per-level window; the returned rank places the 8-byte entry
//       {coef, sample*4+corner} in a counting sort by destination row.
//   P2  row-centric: 4 lanes own one destination row (8 channels each), 8 rows per warp.  The value row
//       is loaded ONCE, the grad_value row is accumulated in REGISTERS over the row's entries
//       (one broadcast LDS.64 + one 128 B G-row read per contribution, no read-modify-write anywhere)
//       and leaves the SM as ONE red.global.add.v4.f32 per row and unit.  The same pass takes the dot
//       products <G[q], value[r]> the location / weight gradients need and parks them per (sample, corner).
//   P3  one thread per sample: grad_attn_weight and grad_sampling_loc from its four dots (cuh:156-158).
// Rows are visited longest list first (counting sort by list length, lists longer than 32 are cut), so
// the 8 rows of a warp pass have equal trip counts.
//
// Windows: per level min(W,22) x min(H,22) rows positioned at the mean sample position of the unit
// (data-driven: correctness never depends on where the queries are), or the whole pyramid when it has
// <= 2048 rows (ARCTIC: 1045).  Corners outside the window become single-entry rows of their own
// ("leftover" rows: same code path, no combining).  Query tiles: 8x8 pixel blocks of each level when
// Lq == S (encoder self-attention: query i is token i), else 64 consecutive queries.
//
// At the DETR-encoder shape a unit has 4096 contributions on ~1000 distinct rows: the SM->L2 port
// carries a quarter of the bytes, value rows are read once per unit instead of once per corner.
#pragma once

#include "msda_d32.cuh"

namespace msda {

#ifndef MSDA_TILE_EXPERIMENT
#define MSDA_TILE_EXPERIMENT 0
#endif
#ifndef MSDA_TILE_WAVES
#define MSDA_TILE_WAVES 4     // grid = 2 resident CTAs per SM x this many; CTAs stride over the units
#endif

constexpr int kTileQ = 64;                        // queries per unit
constexpr int kTileThreads = 512;
constexpr int kTileWarps = kTileThreads / kWarp;
constexpr int kTileLP = 16;                       // L = P = 4
constexpr int kTileL = 4;
constexpr int kTileSamples = kTileQ * kTileLP;    // 1024: two per thread
constexpr int kTileEntries = kTileSamples * 4;    // 4096 corner contributions at most
constexpr int kWinDim = 20;                       // window edge per level when the pyramid is large
constexpr int kRowCap = 1920;                     // window rows + leftover rows
constexpr int kWholeMaxRows = 1600;               // pyramids up to this many tokens are windowed whole
constexpr int kRowsPerThread = 4;                 // one int4 of counters per thread in the scan (threads < kRowCap / 4)
constexpr int kSplit = 32;                        // longer entry lists are cut into virtual rows
constexpr int kBins = kSplit + 1;
constexpr int kRowListCap = kRowCap + kTileEntries / kSplit;
constexpr int kBatchRows = 8;                     // 4 lanes per destination row
static_assert(kRowCap % 4 == 0 && kRowCap / 4 <= kTileThreads, "the scan reads one int4 of counters per thread");

struct TileUnit {          // decoded once per unit by thread 0
  int m, b, lq, y0, x0, pad[3];
};

struct TileShared {
  float4 G[kTileQ * 8];                 // grad_out rows of the unit's queries (8 KB)
  uint2 entries[kTileEntries];          // {coef, sample * 4 + corner}: window rows' lists from the front, leftovers from the back
  float dbuf[kTileEntries];             // <G[q], value row> per (sample, corner) (16 KB)
  float4 rec[kTileSamples];             // {lh, lw, attention weight, corner mask} per sample (16 KB)
  float2 stage_loc[kTileSamples];       // the NEXT unit's sampling locations / weights, copied asynchronously while
  float stage_aw[kTileSamples];         // this unit is processed (12 KB): their HBM latency never heads a unit
  int off[kRowCap + 8];                 // P1: counts; then exclusive offsets, off[R] = total
  int rowtok[kRowCap + 8];              // token * M of every touched window row ([0, R)) and leftover row ((R, ..))
  unsigned rowlist[kRowListCap];        // first entry << 18 | count << 12 | row, longest first
  int whist[kTileWarps][kBins + 1];     // per-warp histogram of list lengths -> per-warp scatter cursors
  int wsum[kTileWarps];
  LevelTab tab;
  int ww[kTileL], wh[kTileL], wbase[kTileL + 1], fixed[kTileL];
  int tile0[kTileL + 1], tiles_x[kTileL];
  int anchor_sum[kTileL][3];            // sum (w_low + 1), sum (h_low + 1), count of accepted samples per level
  TileUnit unit[2];                     // this unit and the next one (alternating)
  int n_left, n_win, n_vrows, n_tiles, all_fixed, next_pass;
};

__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// unit -> (head, sample, tile origin).  grid mode: tile = an 8x8 pixel block of one level (query i is token i);
// linear mode: 64 consecutive queries.
__device__ __forceinline__ void decode_unit(TileShared& sh, int slot, int grid_mode, unsigned u, unsigned n_tiles,
                                            int M) {
  TileUnit d;
  d.m = (int)(u % (unsigned)M);
  const unsigned um = u / (unsigned)M;
  const int tile = (int)(um % n_tiles);
  d.b = (int)(um / n_tiles);
  if (!grid_mode) {
    d.lq = 0; d.y0 = 0; d.x0 = tile * kTileQ;
  } else {
    d.lq = (tile >= sh.tile0[1]) + (tile >= sh.tile0[2]) + (tile >= sh.tile0[3]);
    const int tt = tile - sh.tile0[d.lq];
    const int ty = tt / sh.tiles_x[d.lq];
    d.y0 = ty * 8;
    d.x0 = (tt - ty * sh.tiles_x[d.lq]) * 8;
  }
  d.pad[0] = d.pad[1] = d.pad[2] = 0;
  sh.unit[slot] = d;
}

__device__ __forceinline__ void cp_async_4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// query index of local slot (ql; pixel dy, dx of the 8x8 block in grid mode) of a unit's tile, or -1
__device__ __forceinline__ int unit_query(const TileShared& sh, const TileUnit& un, int grid_mode, int Lq, int ql, int dy,
                                          int dx) {
  if (!grid_mode) {
    const int q = un.x0 + ql;
    return q < Lq ? q : -1;
  }
  const int y = un.y0 + dy, x = un.x0 + dx, W = sh.tab.W[un.lq];
  return (y < sh.tab.H[un.lq] && x < W) ? sh.tab.start[un.lq] + y * W + x : -1;
}

// Asynchronous copies of a unit's inputs: every thread copies exactly what it will read itself (its two samples),
// plus one 16 B chunk of the tile's grad_out rows (read by others: a barrier follows the wait).
__device__ __forceinline__ void request_samples(TileShared& sh, const D32Params& p, const TileUnit& un, int grid_mode, int t) {
  const unsigned qbase = (unsigned)un.b * (unsigned)p.Lq;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int q = unit_query(sh, un, grid_mode, p.Lq, (t >> 4) + 32 * i, (t >> 7) + 4 * i, (t >> 4) & 7);
    if (q >= 0) {
      const unsigned gs = ((qbase + q) * p.M + un.m) * kTileLP + (t & 15);
      cp_async_8(&sh.stage_loc[t + kTileThreads * i], reinterpret_cast<const float2*>(p.loc) + gs);
      cp_async_4(&sh.stage_aw[t + kTileThreads * i], p.aw + gs);
    }
  }
}
__device__ __forceinline__ void request_grad_rows(TileShared& sh, const D32Params& p, const TileUnit& un, int grid_mode, int t) {
  const int ql = t >> 3;
  const int qq = unit_query(sh, un, grid_mode, p.Lq, ql, ql >> 3, ql & 7);
  const unsigned qbase = (unsigned)un.b * (unsigned)p.Lq;
  if (qq >= 0) cp_async_16(&sh.G[t], reinterpret_cast<const float4*>(p.grad_out) + (((qbase + qq) * p.M + un.m) * 8u + (t & 7)));
  else sh.G[t] = zero4();
}

// One corner contribution served by a single thread (the overflow path of the unit's row table).
template <typename VT, typename GT>
__device__ __noinline__ float lone_contribution(const typename Chunk<VT>::Vec* vr, typename Chunk<GT>::Vec* gr,
                                                const float4* g_row, float coef) {
  float d = 0.f;
  for (int ch = 0; ch < 8; ++ch) {
    const float4 g = g_row[ch];
    d += dot4(g, Chunk<VT>::load(vr + ch));
    Chunk<GT>::red_add(gr + ch, scale4(coef, g));
  }
  return d;
}

// A valid corner outside the unit's window: becomes a single-entry row of its own, its entry goes to the back of
// the entry array right away.  Returns false when the row table is full (the caller then serves the corner alone).
__device__ __noinline__ bool leftover_corner(TileShared& sh, int R, int row_m, float coef, unsigned sc) {
  const int k = atomicAdd(&sh.n_left, 1);
  if (k >= kRowCap - R - 1) return false;
  sh.rowtok[R + 1 + k] = row_m;
  sh.entries[kTileEntries - 1 - k] = make_uint2(__float_as_uint(coef), sc);
  return true;
}

// kFuse: the fused-prologue operator (msda_fused_bwd_f32): `loc` holds raw sampling offsets, `aw` raw attention logits;
// softmax and reference + offset / (W, H) happen in P1a (the same routines as the other fused kernels), their
// backward in P3 (/root/reference/models/ops/modules/ms_deform_attn.py:100-108).
template <typename VT, typename GT, bool kFuse = false>
__global__ void __launch_bounds__(kTileThreads, 2)
bwd_d32_tile_kernel(const D32Params p, const int grid_mode) {
  using C = Chunk<VT>;
  using Gc = Chunk<GT>;
  using Vec = typename C::Vec;
  using GVec = typename Gc::Vec;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileShared& sh = *reinterpret_cast<TileShared*>(smem_raw);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int M = p.M;

  // ---- once per CTA: level tables, window shapes, tile enumeration, cleared counters ------------------
  if (t < kTileL) {
    sh.tab.H[t] = (int)p.shapes[2 * t];
    sh.tab.W[t] = (int)p.shapes[2 * t + 1];
    sh.tab.start[t] = (int)p.lsi[t];
  }
  for (int i = t; i < kRowCap + 8; i += kTileThreads) sh.off[i] = 0;
  for (int i = t; i < kTileWarps * (kBins + 1); i += kTileThreads) (&sh.whist[0][0])[i] = 0;
  if (t < kTileL * 3) (&sh.anchor_sum[0][0])[t] = 0;
  __syncthreads();
  if (t == 0) {
    const bool whole = p.S <= kWholeMaxRows;
    int base = 0, tiles = 0, all_fixed = 1;
    for (int l = 0; l < kTileL; ++l) {
      const int H = sh.tab.H[l], W = sh.tab.W[l];
      const int ww = whole ? W : min(W, kWinDim), wh = whole ? H : min(H, kWinDim);
      sh.ww[l] = ww; sh.wh[l] = wh; sh.wbase[l] = base;
      sh.fixed[l] = (ww == W && wh == H);
      all_fixed &= sh.fixed[l];
      base += ww * wh;
      sh.tile0[l] = tiles;
      sh.tiles_x[l] = (W + 7) >> 3;
      tiles += ((W + 7) >> 3) * ((H + 7) >> 3);
    }
    sh.wbase[kTileL] = base;
    sh.tile0[kTileL] = tiles;
    sh.n_tiles = grid_mode ? tiles : (p.Lq + kTileQ - 1) / kTileQ;
    sh.all_fixed = all_fixed;
    sh.n_left = 0;
    sh.next_pass = 0;
    const unsigned nu = (unsigned)p.N * (unsigned)sh.n_tiles * (unsigned)M;
    if (blockIdx.x < nu) decode_unit(sh, 0, grid_mode, blockIdx.x, (unsigned)sh.n_tiles, M);
  }
  __syncthreads();
  const unsigned n_tiles = (unsigned)sh.n_tiles;
  const unsigned n_units = (unsigned)p.N * n_tiles * (unsigned)M;   // the host checked that this fits
  const int R = sh.wbase[kTileL];             // window rows; leftover rows live at R + 1 + k
  const bool all_fixed = sh.all_fixed;

  // this thread's two samples: local query (t >> 4) + 32 i, level/point lp = t & 15
  const int lp = t & 15, l = lp >> 2;
  const int ql_dy = t >> 7, ql_dx = (t >> 4) & 7;          // grid mode: pixel of local query t >> 4 within the block

  if (blockIdx.x < n_units) {
    request_samples(sh, p, sh.unit[0], grid_mode, t);
    request_grad_rows(sh, p, sh.unit[0], grid_mode, t);
  }
  int slot_u = 0;
  for (unsigned u = blockIdx.x; u < n_units; u += gridDim.x, slot_u ^= 1) {
    const TileUnit un = sh.unit[slot_u];
    const bool more = u + gridDim.x < n_units;
    if (t == 0 && more) decode_unit(sh, slot_u ^ 1, grid_mode, u + gridDim.x, n_tiles, M);
    const int m = un.m, b = un.b;
    const unsigned qbase = (unsigned)b * (unsigned)p.Lq;
    const int Hl = sh.tab.H[l], Wl = sh.tab.W[l];
    const int wwl = sh.ww[l], whl = sh.wh[l];
    const bool fixedl = sh.fixed[l];
    const int Hq = sh.tab.H[un.lq], Wq = sh.tab.W[un.lq], startq = sh.tab.start[un.lq];

    // query index of local slot ql (0..63) of the tile, or -1
    auto tile_query = [&](int ql, int dy, int dx) -> int {
      if (!grid_mode) {
        const int q = un.x0 + ql;
        return q < p.Lq ? q : -1;
      }
      const int y = un.y0 + dy, x = un.x0 + dx;
      return (y < Hq && x < Wq) ? startq + y * Wq + x : -1;
    };

    // ---- P1a: geometry of this thread's two samples ------------------------------------------------
    cp_async_wait_all();         // this thread's own copies (its samples, its chunk of G) have landed
    float lh[2], lw[2], aw[2];
    int h_low[2], w_low[2];
    unsigned msk[2];
    int packed = 0, accepted = 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int q = tile_query((t >> 4) + 32 * i, ql_dy + 4 * i, ql_dx);
      msk[i] = 0; lh[i] = 0.f; lw[i] = 0.f; aw[i] = 0.f; h_low[i] = 0; w_low[i] = 0;
      float a_raw = -INFINITY;
      if (q >= 0) a_raw = sh.stage_aw[t + kTileThreads * i];
      // fused: the 16 logits of a (query, head) sit in 16 consecutive lanes (absent queries pass -inf)
      const float a = kFuse ? segment_softmax<kTileLP>(a_raw) : a_raw;
      if (q >= 0) {
        float2 xy = sh.stage_loc[t + kTileThreads * i];
        if (kFuse) {
          xy = fused_location(p, (long long)(qbase + q) * M + m, l, xy, Hl, Wl);
          aw[i] = a;                  // the softmax backward needs the weight of rejected samples too
        }
        SampleGeom<float> g;
        if (sample_geom<float>(xy.x, xy.y, Hl, Wl, g)) {
          msk[i] = g.mask | 16u;      // bit 4: accepted by the coarse test (a sample may have no valid corner)
          lh[i] = g.lh; lw[i] = g.lw; aw[i] = a;
          h_low[i] = g.h_low; w_low[i] = g.w_low;
          packed += (g.w_low + 1) | ((g.h_low + 1) << 16);
          accepted += 1;
        }
      }
      // (the query index rides along so that the epilogue needs nothing that lived through the row pass)
      sh.rec[t + kTileThreads * i] = make_float4(lh[i], lw[i], aw[i], __uint_as_float(((unsigned)(q + 1) << 5) | msk[i]));
    }
    if (!all_fixed) {
      // window anchor of non-whole levels: mean corner position of the unit's accepted samples
      // (lanes with the same level: lane bits 0, 1 and 4 vary; 16 samples of <= 4094 fit the packed 16-bit sums)
      packed += __shfl_xor_sync(kFull, packed, 1);
      accepted += __shfl_xor_sync(kFull, accepted, 1);
      packed += __shfl_xor_sync(kFull, packed, 2);
      accepted += __shfl_xor_sync(kFull, accepted, 2);
      packed += __shfl_xor_sync(kFull, packed, 16);
      accepted += __shfl_xor_sync(kFull, accepted, 16);
      if (!fixedl && (lane & 19) == 0 && accepted > 0) {
        atomicAdd(&sh.anchor_sum[l][0], packed & 0xffff);
        atomicAdd(&sh.anchor_sum[l][1], packed >> 16);
        atomicAdd(&sh.anchor_sum[l][2], accepted);
      }
    }
    __syncthreads();                                                                   // B1 (also: G, next unit's slot)

    // ---- P1b: count contributions per destination row ------------------------------------------------
    int x0 = 0, y0 = 0;
    if (!fixedl) {
      const int n = sh.anchor_sum[l][2];
      if (n > 0) {
        const float inv = 1.f / (float)n;
        x0 = (int)floorf((float)sh.anchor_sum[l][0] * inv - 0.5f * (float)wwl);
        y0 = (int)floorf((float)sh.anchor_sum[l][1] * inv - 0.5f * (float)whl);
        x0 = max(0, min(x0, Wl - wwl));
        y0 = max(0, min(y0, Hl - whl));
      }
    }
    // slot per corner: window row (12 bits) | rank << 12, or 0 when the corner has no entry in a window row's list
    unsigned slot[2][4];
    {
      const int wbasel = sh.wbase[l], startl = sh.tab.start[l];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int ux = w_low[i] - x0, uy = h_low[i] - y0;
        const bool inx[2] = {(unsigned)ux < (unsigned)wwl, (unsigned)(ux + 1) < (unsigned)wwl};
        const bool iny[2] = {(unsigned)uy < (unsigned)whl, (unsigned)(uy + 1) < (unsigned)whl};
        const int wr0 = wbasel + uy * wwl + ux;
        const int tok0 = startl + h_low[i] * Wl + w_low[i];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          slot[i][c] = 0;
          if ((msk[i] >> c) & 1u) {
            const int tok = tok0 + (c & 1) + (c & 2 ? Wl : 0);
            if (inx[c & 1] && iny[c >> 1]) {
              const int wr = wr0 + (c & 1) + (c & 2 ? wwl : 0);
              const int rank = atomicAdd(&sh.off[wr], 1);
              if (rank == 0) sh.rowtok[wr] = tok * M;
              slot[i][c] = 0x80000000u | ((unsigned)rank << 12) | (unsigned)wr;
            } else {
              const float wgt = (c & 2 ? lh[i] : 1.f - lh[i]) * (c & 1 ? lw[i] : 1.f - lw[i]);
              const unsigned sc = (unsigned)((t + kTileThreads * i) * 4 + c);
              if (!leftover_corner(sh, R, tok * M, wgt * aw[i], sc)) {
                // no room left in the unit's row table: this thread alone serves the row (slow; only when
                // the windows miss almost everything, which the host-side kernel choice avoids)
                const long long ro = ((long long)b * p.S * M + (long long)tok * M + m) * 8;
                sh.dbuf[sc] = lone_contribution<VT, GT>(reinterpret_cast<const Vec*>(p.value) + ro,
                                                        reinterpret_cast<GVec*>(p.grad_value) + ro,
                                                        sh.G + (sc >> 6) * 8, wgt * aw[i]);
              }
            }
          }
        }
      }
    }
    __syncthreads();                                                                   // B2
    if (more) request_samples(sh, p, sh.unit[slot_u ^ 1], grid_mode, t);

    // ---- scan of the window rows' counts (thread: rows 4t..4t+3), histogram of list lengths ------------
    int4 cnt4;
    int excl;
    {
      cnt4 = t < kRowCap / 4 ? reinterpret_cast<const int4*>(sh.off)[t] : make_int4(0, 0, 0, 0);
      const int r0 = t * kRowsPerThread;
      if (r0 + 0 >= R) cnt4.x = 0;
      if (r0 + 1 >= R) cnt4.y = 0;
      if (r0 + 2 >= R) cnt4.z = 0;
      if (r0 + 3 >= R) cnt4.w = 0;
      const int cs[4] = {cnt4.x, cnt4.y, cnt4.z, cnt4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = cs[j];
        if (c > 0) {
          if (c >> 5) atomicAdd(&sh.whist[warp][kSplit], c >> 5);
          if (c & 31) atomicAdd(&sh.whist[warp][c & 31], 1);
        }
      }
      const int run = cnt4.x + cnt4.y + cnt4.z + cnt4.w;
      int inc = run;
#pragma unroll
      for (int o = 1; o < kWarp; o <<= 1) {
        const int v = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += v;
      }
      excl = inc - run;
      if (lane == kWarp - 1) sh.wsum[warp] = inc;
    }
    __syncthreads();                                                                   // B3
    int row_off0;      // offset of this thread's first row
    {
      int before = lane < warp ? sh.wsum[lane] : 0;
      int total = lane < kTileWarps ? sh.wsum[lane] : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        before += __shfl_xor_sync(kFull, before, o);
        total += __shfl_xor_sync(kFull, total, o);
      }
      row_off0 = before + excl;
      // (rows >= R of the last touched int4 get offsets too: never read)
      if (t < kRowCap / 4)
        reinterpret_cast<int4*>(sh.off)[t] = make_int4(row_off0, row_off0 + cnt4.x, row_off0 + cnt4.x + cnt4.y,
                                                       row_off0 + cnt4.x + cnt4.y + cnt4.z);
      if (t == 0) sh.n_win = total;
      if (warp == kTileWarps - 1) {
        // per-warp scatter cursors: bins in descending length, warps in order within a bin
        const int bin = kSplit - lane;           // lanes 0..31 <-> bins 32..1
        int tot = 0;
#pragma unroll
        for (int w = 0; w < kTileWarps; ++w) tot += sh.whist[w][bin];
        int inc = tot;
#pragma unroll
        for (int o = 1; o < kWarp; o <<= 1) {
          const int v = __shfl_up_sync(kFull, inc, o);
          if (lane >= o) inc += v;
        }
        int cur = inc - tot;
#pragma unroll
        for (int w = 0; w < kTileWarps; ++w) {
          const int h = sh.whist[w][bin];
          sh.whist[w][bin] = cur;
          cur += h;
        }
        if (lane == kWarp - 1) sh.n_vrows = inc;
      }
    }
    __syncthreads();                                                                   // B4

    // ---- entries into their rows' lists; row list, longest first ----------------------------------------
    const int n_left = min(sh.n_left, kRowCap - R - 1);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float hh = 1.f - lh[i], hw = 1.f - lw[i];
      const float ah = hh * aw[i], al = lh[i] * aw[i];
      const float cf[4] = {ah * hw, ah * lw[i], al * hw, al * lw[i]};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (slot[i][c]) {
          const int pos = sh.off[slot[i][c] & 0xfff] + (int)((slot[i][c] >> 12) & 0x7ffff);
          sh.entries[pos] = make_uint2(__float_as_uint(cf[c]), (unsigned)((t + kTileThreads * i) * 4 + c));
        }
      }
    }
    {
      const int cs[4] = {cnt4.x, cnt4.y, cnt4.z, cnt4.w};
      int o = row_off0;
#pragma unroll
      for (int j = 0; j < kRowsPerThread; ++j) {
        const int c = cs[j];
        if (c > 0) {
          const unsigned r = (unsigned)(t * kRowsPerThread + j);
          int done = 0;
          for (int f = 0; f < (c >> 5); ++f, done += kSplit) {
            const int pos = atomicAdd(&sh.whist[warp][kSplit], 1);
            sh.rowlist[pos] = ((unsigned)(o + done) << 18) | (kSplit << 12) | r;
          }
          if (c & 31) {
            const int pos = atomicAdd(&sh.whist[warp][c & 31], 1);
            sh.rowlist[pos] = ((unsigned)(o + done) << 18) | ((unsigned)(c & 31) << 12) | r;
          }
          o += c;
        }
      }
    }
    const int n_vrows = sh.n_vrows;
    for (int k = t; k < n_left; k += kTileThreads)
      sh.rowlist[n_vrows + k] = ((unsigned)(kTileEntries - 1 - k) << 18) | (1u << 12) | (unsigned)(R + 1 + k);
    __syncthreads();                                                                   // B5

    // ---- P2: row-centric accumulation -------------------------------------------------------------------
    // 4 lanes own one destination row; lane h of group g owns chunks h and 4 + h (adjacent lanes cover adjacent 16 B
    // pieces: full sectors towards L2) and odd groups visit them in the opposite order so that the 8 lanes of a
    // quarter-warp read 8 different chunks = all 32 banks.
    {
      const int n_rows = n_vrows + n_left;
      const int g = lane >> 2, h = lane & 3;
      const int first = 4 * (g & 1) + h;
      const unsigned head = ((unsigned)b * (unsigned)p.S * (unsigned)M + (unsigned)m) * 8u;   // < 2^31 (host-checked)
      const Vec* vb = reinterpret_cast<const Vec*>(p.value);
      GVec* gvb = reinterpret_cast<GVec*>(p.grad_value);
      const float4* GA = sh.G + first;
      const float4* GB = sh.G + (first ^ 4);
      // Passes (8 rows each): the first three quarters go to the warps in serpentine order (lists are sorted by length,
      // so straight round robin would hand warp 0 the longest pass of every round), the rest is handed out by a
      // shared counter so that the warps reach the barrier together.  The next pass's row and value chunks are
      // requested before the current pass's entry loop (their L2 latency hides behind it).
      const int n_pass = (n_rows + kBatchRows - 1) / kBatchRows;
      const int static_rounds = (n_pass / kTileWarps) * 3 / 4;
      int round = 0;
      auto next_pass = [&]() -> int {
        int pass;
        if (round < static_rounds) {
          pass = round * kTileWarps + ((round & 1) ? kTileWarps - 1 - warp : warp);
        } else {
          unsigned v = 0;
          if (lane == 0)
            asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(&sh.next_pass)) : "memory");
          pass = static_rounds * kTileWarps + (int)__shfl_sync(kFull, v, 0);
        }
        ++round;
        return pass;
      };
      unsigned rl_n = 0u;
      float4 vA_n = zero4(), vB_n = zero4();
      auto request = [&](int pass) {
        const int j = pass * kBatchRows + g;
        rl_n = 0u; vA_n = zero4(); vB_n = zero4();
        if (pass < n_pass && j < n_rows) {
          rl_n = sh.rowlist[j] | 0x40000000u;        // bit 30 (above the 12-bit entry offset): the slot holds a row
          const unsigned ro = head + (unsigned)sh.rowtok[rl_n & 0xfffu] * 8u;
          vA_n = C::load(vb + (ro + first));
          vB_n = C::load(vb + (ro + (first ^ 4)));
        }
      };
      int cur = next_pass();
      request(cur);
      while (cur < n_pass) {
        const unsigned rl = rl_n;
        const float4 vA = vA_n, vB = vB_n;
        cur = next_pass();
        request(cur);
        const bool has = (rl >> 30) & 1u;
        const int c = (int)((rl >> 12) & 63u);
        const uint2* ep = sh.entries + ((rl >> 18) & 0xfffu);
        // lists are sorted by length: group 0 has the longest, group 7 the shortest of the pass
#if MSDA_TILE_EXPERIMENT == 3      // timing experiment: no entry loop at all
        const int cmax = 0, cmin = 0;
#else
        const int cmax = __shfl_sync(kFull, c, 0), cmin = __shfl_sync(kFull, c, 28);
#endif
        const unsigned ro = head + (unsigned)sh.rowtok[rl & 0xfffu] * 8u;
        float4 accA = zero4(), accB = zero4();
        int i = 0;
        // every group has at least cmin entries: two per trip, no predicates; the two dot products are reduced
        // together (lanes h = 0, 1 end up with the first, h = 2, 3 with the second)
        for (; i + 2 <= cmin; i += 2) {
          const uint2 e0 = ep[i], e1 = ep[i + 1];
          const unsigned q0 = (e0.y >> 6) << 3, q1 = (e1.y >> 6) << 3;
#if MSDA_TILE_EXPERIMENT == 1      // timing experiment: half of the G-row reads
          const float4 a0 = GA[q0], b0 = a0, a1 = GA[q1], b1 = a1;
#else
          const float4 a0 = GA[q0], b0 = GB[q0], a1 = GA[q1], b1 = GB[q1];
#endif
          fma4(accA, __uint_as_float(e0.x), a0);
          fma4(accB, __uint_as_float(e0.x), b0);
          fma4(accA, __uint_as_float(e1.x), a1);
          fma4(accB, __uint_as_float(e1.x), b1);
#if MSDA_TILE_EXPERIMENT != 2      // (2: timing experiment without the dot products)
          const float d0 = dot4(a0, vA) + dot4(b0, vB);
          const float d1 = dot4(a1, vA) + dot4(b1, vB);
          float r = ((h & 2) ? d1 : d0) + __shfl_xor_sync(kFull, (h & 2) ? d0 : d1, 2);
          r += __shfl_xor_sync(kFull, r, 1);
          if ((h & 1) == 0) sh.dbuf[(h & 2) ? e1.y : e0.y] = r;
#endif
        }
        for (; i < cmax; ++i) {
          const bool act = i < c;
          uint2 e = make_uint2(0u, 0u);
          if (act) e = ep[i];
          const unsigned qq = (e.y >> 6) << 3;
          const float4 a0 = GA[qq], b0 = GB[qq];
          fma4(accA, __uint_as_float(e.x), a0);
          fma4(accB, __uint_as_float(e.x), b0);
          float dp = dot4(a0, vA) + dot4(b0, vB);
          dp += __shfl_xor_sync(kFull, dp, 2);      // same association as the paired trips: results do not depend
          dp += __shfl_xor_sync(kFull, dp, 1);      // on where in its row's list an entry landed
          if (act && h == 0) sh.dbuf[e.y] = dp;
        }
        if (has) {
          Gc::red_add(gvb + (ro + first), accA);
          Gc::red_add(gvb + (ro + (first ^ 4)), accB);
        }
      }
    }
    {
      const int Rz = sh.wbase[kTileL];
      for (int i = t; i < Rz; i += kTileThreads) sh.off[i] = 0;
      for (int i = t; i < kTileWarps * (kBins + 1); i += kTileThreads) (&sh.whist[0][0])[i] = 0;
      if (t < kTileL * 3) (&sh.anchor_sum[0][0])[t] = 0;
      if (t == 0) sh.n_left = 0;
    }
    __syncthreads();                                                                   // B6

    // ---- P3: per-sample gradients (cuh:156-158); the next unit's grad_out rows are requested (G is free again).
    // Everything is re-read from shared memory: nothing but the loop state lives across the row pass. ------------
    {
      const TileUnit un3 = sh.unit[slot_u];
      if (more) request_grad_rows(sh, p, sh.unit[slot_u ^ 1], grid_mode, t);
      if (t == 0) sh.next_pass = 0;
      const int l3 = (t & 15) >> 2;
      const float Wf = (float)sh.tab.W[l3], Hf = (float)sh.tab.H[l3];
      const unsigned qb3 = (unsigned)un3.b * (unsigned)p.Lq;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float4 rc = sh.rec[t + kTileThreads * i];
        const unsigned mk = __float_as_uint(rc.w);
        if (!kFuse && (mk >> 5)) {
          const unsigned gs = ((qb3 + (mk >> 5) - 1u) * (unsigned)p.M + (unsigned)un3.m) * kTileLP + (t & 15);
          float pa = 0.f, px = 0.f, py = 0.f;
          if (mk & 15u) {
            const float4 d = reinterpret_cast<const float4*>(sh.dbuf)[t + kTileThreads * i];
            const float d1 = (mk & 1u) ? d.x : 0.f, d2 = (mk & 2u) ? d.y : 0.f;
            const float d3 = (mk & 4u) ? d.z : 0.f, d4 = (mk & 8u) ? d.w : 0.f;
            const float hh = 1.f - rc.x, hw = 1.f - rc.y;
            pa = hh * hw * d1 + hh * rc.y * d2 + rc.x * hw * d3 + rc.x * rc.y * d4;
            px = (hh * (d2 - d1) + rc.x * (d4 - d3)) * rc.z * Wf;
            py = (hw * (d3 - d1) + rc.y * (d4 - d2)) * rc.z * Hf;
          }
          if (!kFuse) {
            st_stream(p.grad_aw + gs, pa);
            st_stream(reinterpret_cast<float2*>(p.grad_loc) + gs, make_float2(px, py));
          }
        }
        if (kFuse) {
          // (every lane takes part in the shuffles; absent queries carry zeros)
          const bool live = (mk >> 5) != 0;
          const float a = live ? rc.z : 0.f;
          float pa = 0.f, px = 0.f, py = 0.f;
          if (live && (mk & 15u)) {
            const float4 d = reinterpret_cast<const float4*>(sh.dbuf)[t + kTileThreads * i];
            const float d1 = (mk & 1u) ? d.x : 0.f, d2 = (mk & 2u) ? d.y : 0.f;
            const float d3 = (mk & 4u) ? d.z : 0.f, d4 = (mk & 8u) ? d.w : 0.f;
            const float hh = 1.f - rc.x, hw = 1.f - rc.y;
            pa = hh * hw * d1 + hh * rc.y * d2 + rc.x * hw * d3 + rc.x * rc.y * d4;
            px = (hh * (d2 - d1) + rc.x * (d4 - d3)) * a * Wf;
            py = (hw * (d3 - d1) + rc.y * (d4 - d2)) * a * Hf;
          }
          // softmax backward over the 16 samples of the (query, head): d/d logit_s = a_s (ga_s - sum_j a_j ga_j)
          float dot = a * pa;
#pragma unroll
          for (int o = 1; o < kTileLP; o <<= 1) dot += __shfl_xor_sync(kFull, dot, o);
          // d/d reference point of (query, level) = sum over points (here) and heads (other units: atomics)
          float rx = px, ry = py;
          rx += __shfl_xor_sync(kFull, rx, 1); ry += __shfl_xor_sync(kFull, ry, 1);
          rx += __shfl_xor_sync(kFull, rx, 2); ry += __shfl_xor_sync(kFull, ry, 2);
          if (live) {
            const unsigned q3 = qb3 + (mk >> 5) - 1u;
            const unsigned gs = (q3 * (unsigned)p.M + (unsigned)un3.m) * kTileLP + (t & 15);
            st_stream(p.grad_aw + gs, a * (pa - dot));
            st_stream(reinterpret_cast<float2*>(p.grad_loc) + gs, make_float2(__fdiv_rn(px, Wf), __fdiv_rn(py, Hf)));
            if (p.grad_ref && (t & 3) == 0) {
              float* gr = p.grad_ref + ((size_t)q3 * kTileL + l3) * 2;
              atomicAdd(gr, rx);
              atomicAdd(gr + 1, ry);
            }
          }
        }
      }
    }
    __syncthreads();                                                                   // B7
  }
}

static_assert(2 * (sizeof(TileShared) + 1024) <= 228 * 1024, "two units per SM");
inline size_t tile_smem_bytes() { return sizeof(TileShared); }

}  // namespace msda
