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
// (rows x queries) matrix times the tile's grad_out rows.  One UNIT = (sample b, tile of 64 queries, head m);
// a KEY = (level, row y, left column x0) = the row pair (y, x0), (y, x0 + 1) that the two x-corners of a
// bilinear sample fall on (x0 = w_low in [-1, W - 1]): a sample contributes to two keys (top, bottom row).
//   geometry  one thread per bilinear sample: sample_geom (the same routine as every other kernel); the sample's
//             state is parked in shared memory.
//   sort      per sample and key ONE 4-byte shared-memory counter increment (native ATOMS.ADD) on the key's slot
//             in a per-level window; the returned rank places the 16-byte entry
//             {coef left, coef right, sample * 2 + top|bottom} in a counting sort by key; keys are then listed
//             longest list first (counting sort by list length, lists longer than 32 are cut).
//   keys      key-centric: 4 lanes own one key (8 channels each), 8 keys per warp pass.  The two value rows
//             are loaded ONCE, the two grad_value rows are accumulated in REGISTERS over the key's entries
//             (one broadcast LDS.128 + one 128 B G-row read per entry, no read-modify-write anywhere)
//             and leave the SM as ONE red.global.add.v4.f32 per row and unit.  The same pass takes the dot
//             products <G[q], value row> the location / weight gradients need and parks them per (sample, corner).
//   epilogue  one thread per sample: grad_attn_weight and grad_sampling_loc from its four dots (cuh:156-158).
// Every phase is its own (non-inlined) function that keeps its working set in registers and talks to the others
// through shared memory: with one flat kernel body the 64-register budget spilled inside the key loop, and every
// instruction of these latency-bound phases is paid at full latency (32 warps per SM).
//
// Windows: per level min(W + 1, 20) x min(H, 20) keys positioned at the mean sample position of the unit
// (data-driven: correctness never depends on where the queries are), or the whole pyramid when its
// H x (W + 1) keys fit the table (ARCTIC: 1098).  Row pairs outside the window are served on their own after the window
// keys (no combining, any number of them: their entries carry the pair's token).  Query tiles: 8x8 pixel blocks of each level when
// Lq == S (encoder self-attention: query i is token i), else 64 consecutive queries.
// The next unit's sampling locations / weights / grad_out rows are copied asynchronously (cp.async) while the
// current unit is processed.
#pragma once

#include "msda_d32.cuh"

namespace msda {

#ifndef MSDA_TILE_EXPERIMENT
#define MSDA_TILE_EXPERIMENT 0
#endif
#ifndef MSDA_TILE_WAVES
#define MSDA_TILE_WAVES 8     // grid = 2 resident CTAs per SM x this many; CTAs stride over the units (2 / 4 / 8 / 16:
                              // 424 / 420 / 404 / 406 us at C2: shorter CTAs let the hardware balance the tail)
#endif

constexpr int kTileQ = 64;                        // queries per unit
constexpr int kTileThreads = 512;
constexpr int kTileWarps = kTileThreads / kWarp;
constexpr int kTileLP = 16;                       // L = P = 4
constexpr int kTileL = 4;
constexpr int kTileSamples = kTileQ * kTileLP;    // 1024: two per thread
constexpr int kTileEntries = kTileSamples * 2;    // 2048 row-pair contributions at most (two per sample)
constexpr int kTileDots = kTileSamples * 4;       // one dot product per bilinear corner
#ifndef MSDA_TILE_WIN
#define MSDA_TILE_WIN 20
#endif
#ifndef MSDA_TILE_STATIC_4THS
#define MSDA_TILE_STATIC_4THS 2   // share of the key passes handed out statically, the rest through a shared counter
                                  // (0 / 1 / 2 / 3 / 4 quarters: 414 / 397* / 405 / 420 / 414 us at C2; * with 8 waves)
#endif
constexpr int kWinDim = MSDA_TILE_WIN;            // window edge per level when the pyramid is large
constexpr int kRowCap = 1664;                     // window keys (4 levels x 20 x 20, or a whole small pyramid)
constexpr int kWholeMaxRows = 1500;               // host side: pyramids up to this many tokens are expected to be windowed whole
constexpr int kWholeMaxKeys = kRowCap;            // device side: ... when their H x (W + 1) keys fit the table
constexpr int kSplit = 32;                        // longer entry lists are cut into virtual keys
constexpr int kBins = kSplit + 1;
constexpr int kRowListCap = kRowCap + kTileEntries / kSplit;
constexpr int kBatchRows = 8;                     // 4 lanes per key
static_assert(kRowCap % 4 == 0 && kRowCap / 4 <= kTileThreads, "the scan reads one int4 of counters per thread");

struct TileUnit {          // decoded once per unit by thread 0
  int m, b, lq, y0, x0, pad[3];
};

struct TileShared {
  float4 G[kTileQ * 8];                 // grad_out rows of the unit's queries (8 KB)
  uint4 entries[kTileEntries];          // window keys' lists from the front: {coef left, coef right, byte offset of the pair's two
                                        // dots in dbuf, byte offset of the query's row in G}; row pairs outside the windows from
                                        // the back: {coef left, coef right, G offset << 16 | dbuf offset, packed token} (32 KB)
  float dbuf[kTileDots];                // <G[q], value row> per (sample, corner) (16 KB)
  float4 rec[kTileSamples];             // {lh, lw, attention weight, (query + 1) << 5 | accepted << 4 | corner mask} (16 KB)
  int hw[kTileSamples];                 // h_low << 16 | (w_low & 0xffff) per sample (4 KB)
  float2 stage_loc[kTileSamples];       // the NEXT unit's sampling locations / weights, copied asynchronously while
  float stage_aw[kTileSamples];         // this unit is processed (12 KB): their HBM latency never heads a unit
  float2 stage_ref[kTileQ * kTileL];    // fused prologue: the next unit's reference points, one per (query, level) (2 KB)
  int off[kRowCap + 8];                 // counts per key; then exclusive offsets
  int rowtok[kRowCap + 8];              // (token * M of the key's LEFT row) << 2 | right row exists << 1 | left row exists,
                                        // of every touched window key
  unsigned rowlist[kRowListCap];        // first entry << 18 | count << 12 | key, longest first
  int whist[kTileWarps][kBins + 1];     // per-warp histogram of list lengths -> per-warp scatter cursors
  int wsum[kTileWarps];
  LevelTab tab;
  int ww[kTileL], wh[kTileL], wbase[kTileL + 1], fixed[kTileL];
  int tile0[kTileL + 2], tiles_x[kTileL];   // first tile of each level's pixel blocks, of the linear tail, and the count
  int qstart[kTileL + 1];                   // first query of each level's blocks (from the shapes alone), then of the tail
  int anchor_sum[kTileL][3];            // sum (w_low + 1), sum (h_low + 1), count of accepted samples per level
  TileUnit unit[2];                     // this unit and the next one (alternating)
  int n_left, n_win, n_vrows, n_tiles, all_fixed, next_pass;
};
static_assert(2 * (sizeof(TileShared) + 1024) <= 228 * 1024, "two units per SM");

__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
// shared-memory accesses by 32-bit shared address
__device__ __forceinline__ float4 lds128(unsigned addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds128u(unsigned addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(unsigned addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ TileShared& tile_smem() {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  return *reinterpret_cast<TileShared*>(smem_raw);
}

// unit -> (head, sample, tile origin).  grid mode: tile = an 8x8 pixel block of one level, then 64-query pieces of the
// queries beyond the pyramid (lq == kTileL); linear mode: 64 consecutive queries.
// The pixel blocks are enumerated from the level SHAPES alone -- query qstart[l] + y * W_l + x, qstart = running sum of
// H * W -- so every query in [0, Lq) belongs to exactly one unit whatever level_start_index says (gaps, other orders,
// overlapping levels: the operator never checks the layout, cuh:258-297).  When the queries ARE the pyramid's tokens in
// the usual layout (the encoder) the blocks are pixel blocks and their samples fall into the unit's windows; otherwise
// the windows just catch less.
__device__ __forceinline__ void decode_unit(TileShared& sh, int slot, int grid_mode, unsigned u, unsigned n_tiles,
                                            int M) {
  TileUnit d;
  d.m = (int)(u % (unsigned)M);
  const unsigned um = u / (unsigned)M;
  const int tile = (int)(um % n_tiles);
  d.b = (int)(um / n_tiles);
  if (!grid_mode) {
    d.lq = kTileL; d.y0 = 0; d.x0 = tile * kTileQ;
  } else {
    d.lq = (tile >= sh.tile0[1]) + (tile >= sh.tile0[2]) + (tile >= sh.tile0[3]) + (tile >= sh.tile0[4]);
    const int tt = tile - sh.tile0[d.lq];
    if (d.lq == kTileL) {
      d.y0 = 0;
      d.x0 = sh.qstart[kTileL] + tt * kTileQ;
    } else {
      const int ty = tt / sh.tiles_x[d.lq];
      d.y0 = ty * 8;
      d.x0 = (tt - ty * sh.tiles_x[d.lq]) * 8;
    }
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

// query index of local slot (ql; pixel dy, dx of the 8x8 block) of a unit's tile, or -1
__device__ __forceinline__ int unit_query(const TileShared& sh, const TileUnit& un, int grid_mode, int Lq, int ql, int dy,
                                          int dx) {
  int q;
  if (un.lq == kTileL) {
    q = un.x0 + ql;
  } else {
    const int y = un.y0 + dy, x = un.x0 + dx, W = sh.tab.W[un.lq];
    if (y >= sh.tab.H[un.lq] || x >= W) return -1;
    q = sh.qstart[un.lq] + y * W + x;
  }
  return q < Lq ? q : -1;
}

// Asynchronous copies of a unit's inputs: every thread copies exactly what it will read itself (its two samples),
// plus one 16 B chunk of the tile's grad_out rows (read by others: a barrier follows the wait).
__device__ __noinline__ void request_samples(const D32Params& p, int slot, int grid_mode) {
  TileShared& sh = tile_smem();
  const int t = threadIdx.x;
  const TileUnit un = sh.unit[slot];
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
  if (p.ref != nullptr && t < kTileQ * kTileL) {   // fused prologue (read by other threads: tile_geometry has the barrier)
    const int ql = t >> 2;
    const int q = unit_query(sh, un, grid_mode, p.Lq, ql, ql >> 3, ql & 7);
    if (q >= 0) cp_async_8(&sh.stage_ref[t], reinterpret_cast<const float2*>(p.ref) + ((qbase + q) * kTileL + (t & 3)));
  }
}
__device__ __noinline__ void request_grad_rows(const D32Params& p, int slot, int grid_mode) {
  TileShared& sh = tile_smem();
  const int t = threadIdx.x;
  const TileUnit un = sh.unit[slot];
  const int ql = t >> 3;
  const int qq = unit_query(sh, un, grid_mode, p.Lq, ql, ql >> 3, ql & 7);
  const unsigned qbase = (unsigned)un.b * (unsigned)p.Lq;
  if (qq >= 0) cp_async_16(&sh.G[t], reinterpret_cast<const float4*>(p.grad_out) + (((qbase + qq) * p.M + un.m) * 8u + (t & 7)));
  else sh.G[t] = zero4();
}

// A row pair outside the unit's window: no combining, it is served on its own after the window keys.  Its entry goes to
// the back of the entry array (window entries + leftovers <= 2048 = all pairs of the unit: they cannot meet) and carries
// the pair's token itself, so any number of pairs may miss the windows (locations unrelated to the query grid).
__device__ __forceinline__ void leftover_pair(TileShared& sh, int packed_tok, float coef_l, float coef_r, unsigned sc2) {
  const int k = atomicAdd(&sh.n_left, 1);
  sh.entries[kTileEntries - 1 - k] = make_uint4(__float_as_uint(coef_l), __float_as_uint(coef_r),
                                                (((sc2 >> 5) * 128u) << 16) | (sc2 * 8u), (unsigned)packed_tok);
}

// ---- geometry: this thread's two samples (local query (t >> 4) + 32 i, level / point t & 15) -----------------------
// kFuse: `loc` holds raw sampling offsets, `aw` raw attention logits; softmax and reference + offset / (W, H) happen here
// (the same routines as the other fused kernels; /root/reference/models/ops/modules/ms_deform_attn.py:100-108).
template <bool kFuse>
__device__ __noinline__ void tile_geometry(const D32Params& p, int slot, int grid_mode) {
  TileShared& sh = tile_smem();
  const int t = threadIdx.x, lane = t & 31;
  const int lp = t & 15, l = lp >> 2;
  const TileUnit un = sh.unit[slot];
  const int Hl = sh.tab.H[l], Wl = sh.tab.W[l];
  cp_async_wait_all();         // this thread's own copies (its samples, its chunk of G) have landed
  if (kFuse) __syncthreads();  // ... and the reference points, which other threads copied
  int packed = 0, accepted = 0;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int q = unit_query(sh, un, grid_mode, p.Lq, (t >> 4) + 32 * i, (t >> 7) + 4 * i, (t >> 4) & 7);
    float lh = 0.f, lw = 0.f, aw = 0.f;
    unsigned msk = 0;
    int hw = 0;
    float a_raw = -INFINITY;
    if (q >= 0) a_raw = sh.stage_aw[t + kTileThreads * i];
    // fused: the 16 logits of a (query, head) sit in 16 consecutive lanes (absent queries pass -inf)
    const float a = kFuse ? segment_softmax<kTileLP>(a_raw) : a_raw;
    if (q >= 0) {
      float2 xy = sh.stage_loc[t + kTileThreads * i];
      if (kFuse) {
        xy = fused_location(sh.stage_ref[((t >> 4) + 32 * i) * kTileL + l], xy, Hl, Wl);
        aw = a;                     // the softmax backward needs the weight of rejected samples too
      }
      SampleGeom<float> g;
      if (sample_geom<float>(xy.x, xy.y, Hl, Wl, g)) {
        msk = g.mask | 16u;         // bit 4: accepted by the coarse test (a sample may have no valid corner)
        lh = g.lh; lw = g.lw; aw = a;
        hw = (g.h_low << 16) | (g.w_low & 0xffff);
        packed += (g.w_low + 1) | ((g.h_low + 1) << 16);
        accepted += 1;
      }
    }
    sh.rec[t + kTileThreads * i] = make_float4(lh, lw, aw, __uint_as_float(((unsigned)(q + 1) << 5) | msk));
    sh.hw[t + kTileThreads * i] = hw;
  }
  if (!sh.all_fixed) {
    // window anchor of non-whole levels: mean corner position of the unit's accepted samples
    // (lanes with the same level: lane bits 0, 1 and 4 vary; 16 samples of <= 4094 fit the packed 16-bit sums)
    packed += __shfl_xor_sync(kFull, packed, 1);
    accepted += __shfl_xor_sync(kFull, accepted, 1);
    packed += __shfl_xor_sync(kFull, packed, 2);
    accepted += __shfl_xor_sync(kFull, accepted, 2);
    packed += __shfl_xor_sync(kFull, packed, 16);
    accepted += __shfl_xor_sync(kFull, accepted, 16);
    if (!sh.fixed[l] && (lane & 19) == 0 && accepted > 0) {
      atomicAdd(&sh.anchor_sum[l][0], packed & 0xffff);
      atomicAdd(&sh.anchor_sum[l][1], packed >> 16);
      atomicAdd(&sh.anchor_sum[l][2], accepted);
    }
  }
}

// ---- sort: count per key, scan, entries into their keys' lists, key list longest first ------------------------------
// (contains barriers: called by every thread of the CTA)
__device__ __noinline__ void tile_sort(const D32Params& p, int next_slot, int grid_mode) {
  TileShared& sh = tile_smem();
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int l = (t & 15) >> 2;
  const int M = p.M;
  const int R = sh.wbase[kTileL];             // window keys
  unsigned slot_of[2][2];                     // per row pair (0 top, 1 bottom): key (12 bits) | rank << 12 | bit 31, or 0
  {
    const int Hl = sh.tab.H[l], Wl = sh.tab.W[l];
    const int wwl = sh.ww[l], whl = sh.wh[l];
    // window of keys: x0 in [X0, X0 + ww), y in [Y0, Y0 + wh); anchored at the mean sample position of the unit
    int x0 = -1, y0 = 0;
    if (!sh.fixed[l]) {
      const int n = sh.anchor_sum[l][2];
      x0 = 0;
      if (n > 0) {
        const float inv = 1.f / (float)n;       // the sums are of (w_low + 1), (h_low + 1)
        x0 = (int)floorf((float)sh.anchor_sum[l][0] * inv - 0.5f - 0.5f * (float)wwl);
        y0 = (int)floorf((float)sh.anchor_sum[l][1] * inv - 0.5f * (float)whl);
      }
      x0 = max(-1, min(x0, Wl - wwl));
      y0 = max(0, min(y0, Hl - whl));
    }
    const int wbasel = sh.wbase[l], startl = sh.tab.start[l];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const unsigned msk = __float_as_uint(sh.rec[t + kTileThreads * i].w) & 15u;
      const int hw = sh.hw[t + kTileThreads * i];
      const int h_low = hw >> 16, w_low = (int)(short)(hw & 0xffff);
      const int ux = w_low - x0, uy = h_low - y0;
      const bool inx = (unsigned)ux < (unsigned)wwl;
      const int wr0 = wbasel + uy * wwl + ux;
      const int tok0 = startl + h_low * Wl + w_low;          // left token of the top pair (may not exist)
      const int side = (w_low >= 0 ? 1 : 0) | (w_low + 1 <= Wl - 1 ? 2 : 0);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        slot_of[i][r] = 0;
        const unsigned pair = (msk >> (2 * r)) & 3u;
        if (pair) {
          const int packed_tok = (((tok0 + r * Wl) * M) << 2) | side;
          if (inx && (unsigned)(uy + r) < (unsigned)whl) {
            const int wr = wr0 + r * wwl;
            const int rank = atomicAdd(&sh.off[wr], 1);
            if (rank == 0) sh.rowtok[wr] = packed_tok;
            slot_of[i][r] = 0x80000000u | ((unsigned)rank << 12) | (unsigned)wr;
          } else {
            const float4 rc = sh.rec[t + kTileThreads * i];
            const float wy = (r ? rc.x : 1.f - rc.x) * rc.z;
            const float cl = (pair & 1u) ? wy * (1.f - rc.y) : 0.f, cr = (pair & 2u) ? wy * rc.y : 0.f;
            leftover_pair(sh, packed_tok, cl, cr, (unsigned)((t + kTileThreads * i) * 2 + r));
          }
        }
      }
    }
  }
  __syncthreads();                                                                     // counts complete
  if (next_slot >= 0) request_samples(p, next_slot, grid_mode);    // this unit's staged inputs are consumed

  // ---- scan of the window keys' counts (thread: keys 4t..4t+3), histogram of list lengths
  int4 cnt4 = t < kRowCap / 4 ? reinterpret_cast<const int4*>(sh.off)[t] : make_int4(0, 0, 0, 0);
  int excl;
  {
    const int r0 = t * 4;
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
  __syncthreads();
  int row_off0;      // offset of this thread's first key
  {
    int before = lane < warp ? sh.wsum[lane] : 0;
    int total = lane < kTileWarps ? sh.wsum[lane] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      before += __shfl_xor_sync(kFull, before, o);
      total += __shfl_xor_sync(kFull, total, o);
    }
    row_off0 = before + excl;
    // (keys >= R of the last touched int4 get offsets too: never read)
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
  __syncthreads();

  // ---- entries into their keys' lists; key list, longest first
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    if (slot_of[i][0] | slot_of[i][1]) {
      const float4 rc = sh.rec[t + kTileThreads * i];
      const unsigned msk = __float_as_uint(rc.w);
      const float hw = 1.f - rc.y;
      const float ay[2] = {(1.f - rc.x) * rc.z, rc.x * rc.z};
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (slot_of[i][r]) {
          const unsigned pair = (msk >> (2 * r)) & 3u;
          const int pos = sh.off[slot_of[i][r] & 0xfff] + (int)((slot_of[i][r] >> 12) & 0x7ffff);
          const float cl = (pair & 1u) ? ay[r] * hw : 0.f, cr = (pair & 2u) ? ay[r] * rc.y : 0.f;
          const unsigned sc2 = (unsigned)((t + kTileThreads * i) * 2 + r);
          sh.entries[pos] = make_uint4(__float_as_uint(cl), __float_as_uint(cr), sc2 * 8u, (sc2 >> 5) * 128u);
        }
      }
    }
  }
  {
    const int cs[4] = {cnt4.x, cnt4.y, cnt4.z, cnt4.w};
    int o = row_off0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = cs[j];
      if (c > 0) {
        const unsigned r = (unsigned)(t * 4 + j);
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
}

// ---- keys: key-centric accumulation -------------------------------------------------------------------------------
// 4 lanes own one key = the row pair (y, x0), (y, x0 + 1); lane h of group g owns chunks h and 4 + h of both rows
// (adjacent lanes cover adjacent 16 B pieces: full sectors towards L2) and odd groups visit their two chunks in
// the opposite order so that the 8 lanes of a quarter-warp read 8 different chunks of G = all 32 banks.
// One entry = one sample's contribution to BOTH rows: one read of the query's grad_out row serves two
// accumulations and two dot products.
template <typename VT, typename GT>
__device__ __noinline__ void tile_keys(const void* value, void* grad_value, unsigned head, unsigned row_stride) {
  using C = Chunk<VT>;
  using Gc = Chunk<GT>;
  using Vec = typename C::Vec;
  using GVec = typename Gc::Vec;
  TileShared& sh = tile_smem();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_rows = sh.n_vrows;
  const int g = lane >> 2, h = lane & 3;
  const int first = 4 * (g & 1) + h;
  const Vec* vb = reinterpret_cast<const Vec*>(value);
  GVec* gvb = reinterpret_cast<GVec*>(grad_value);
  const unsigned ga_base = (unsigned)__cvta_generic_to_shared(sh.G + first);
  const unsigned gb_base = (unsigned)__cvta_generic_to_shared(sh.G + (first ^ 4));
  const unsigned ent_base = (unsigned)__cvta_generic_to_shared(sh.entries);
  const unsigned dot_base = (unsigned)__cvta_generic_to_shared(sh.dbuf) + (unsigned)(h >> 1) * 4u;
  // Passes (8 keys each): the first half goes to the warps in serpentine order (lists are sorted by length,
  // so straight round robin would hand warp 0 the longest pass of every round), the rest is handed out by a
  // shared counter so that the warps reach the barrier together.
  const int n_pass = (n_rows + kBatchRows - 1) / kBatchRows;
  const int static_rounds = (n_pass / kTileWarps) * MSDA_TILE_STATIC_4THS / 4;
  int round = 0;
  for (;;) {
    int cur;
    if (round < static_rounds) {
      cur = round * kTileWarps + ((round & 1) ? kTileWarps - 1 - warp : warp);
    } else {
      unsigned v = 0;
      if (lane == 0)
        asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(&sh.next_pass)) : "memory");
      cur = static_rounds * kTileWarps + (int)__shfl_sync(kFull, v, 0);
    }
    ++round;
    if (cur >= n_pass) break;
    const int j = cur * kBatchRows + g;
    const bool has = j < n_rows;
    const unsigned rl = has ? sh.rowlist[j] : 0u;
    const int c = (int)((rl >> 12) & 63u);
    // lists are sorted by length: group 0 has the longest, group 7 the shortest of the pass
#if MSDA_TILE_EXPERIMENT == 3      // timing experiment: no entry loop at all
    const int cmax = 0, cmin = 0;
#else
    const int cmax = __shfl_sync(kFull, c, 0), cmin = __shfl_sync(kFull, c, 28);
#endif
    const int packed_tok = has ? sh.rowtok[rl & 0xfffu] : 0;
    const bool has_l = packed_tok & 1, has_r = packed_tok & 2;
    const unsigned ro_l = head + (unsigned)((packed_tok >> 2) * 8), ro_r = ro_l + row_stride;
    float4 vLA = zero4(), vLB = zero4(), vRA = zero4(), vRB = zero4();
    if (has_l) { vLA = C::load(vb + (ro_l + first)); vLB = C::load(vb + (ro_l + (first ^ 4))); }
    if (has_r) { vRA = C::load(vb + (ro_r + first)); vRB = C::load(vb + (ro_r + (first ^ 4))); }
    float4 aLA = zero4(), aLB = zero4(), aRA = zero4(), aRB = zero4();
    // the two dot products of an entry are reduced together: lanes h = 0, 1 end up with the left row's, h = 2, 3
    // with the right row's; dbuf index = sample * 4 + (top | bottom) * 2 + (left | right)
    // shared-memory addresses are carried as 32-bit offsets: one add per access, no generic-address arithmetic
    auto one = [&](const uint4 e, bool store) {
      const float4 a0 = lds128(ga_base + e.w), b0 = lds128(gb_base + e.w);
      const float cl = __uint_as_float(e.x), cr = __uint_as_float(e.y);
      fma4(aLA, cl, a0);
      fma4(aLB, cl, b0);
      fma4(aRA, cr, a0);
      fma4(aRB, cr, b0);
#if MSDA_TILE_EXPERIMENT != 2      // (2: timing experiment without the dot products)
      const float dl = dot4(a0, vLA) + dot4(b0, vLB);
      const float dr = dot4(a0, vRA) + dot4(b0, vRB);
      float r = ((h & 2) ? dr : dl) + __shfl_xor_sync(kFull, (h & 2) ? dl : dr, 2);
      r += __shfl_xor_sync(kFull, r, 1);
      if (store && (h & 1) == 0) sts32(dot_base + e.z, r);
#endif
    };
    const unsigned ep = ent_base + (rl >> 18) * 16u;
    int i = 0;
#pragma unroll 2
    for (; i < cmin; ++i) one(lds128u(ep + i * 16), true);          // every group has at least cmin entries: no predicates
    for (; i < cmax; ++i) {
      const bool act = i < c;
      uint4 e = make_uint4(0u, 0u, 0u, 0u);
      if (act) e = lds128u(ep + i * 16);
      one(e, act);
    }
    if (has_l) {
      Gc::red_add(gvb + (ro_l + first), aLA);
      Gc::red_add(gvb + (ro_l + (first ^ 4)), aLB);
    }
    if (has_r) {
      Gc::red_add(gvb + (ro_r + first), aRA);
      Gc::red_add(gvb + (ro_r + (first ^ 4)), aRB);
    }
  }
  // ---- row pairs outside the windows: one pair per 4-lane group, nothing to combine (their entries carry the token)
  {
    const int n_left = sh.n_left;
    for (int k0 = warp * kBatchRows; k0 < n_left; k0 += kTileWarps * kBatchRows) {
      const int k = k0 + g;
      const bool has = k < n_left;
      uint4 e = make_uint4(0u, 0u, 0u, 0u);
      if (has) e = lds128u(ent_base + (unsigned)(kTileEntries - 1 - k) * 16u);
      const int packed_tok = (int)e.w;
      const bool has_l = has && (packed_tok & 1), has_r = has && (packed_tok & 2);
      const unsigned ro_l = head + (unsigned)((packed_tok >> 2) * 8), ro_r = ro_l + row_stride;
      float4 vLA = zero4(), vLB = zero4(), vRA = zero4(), vRB = zero4();
      if (has_l) { vLA = C::load(vb + (ro_l + first)); vLB = C::load(vb + (ro_l + (first ^ 4))); }
      if (has_r) { vRA = C::load(vb + (ro_r + first)); vRB = C::load(vb + (ro_r + (first ^ 4))); }
      const float4 a0 = lds128(ga_base + (e.z >> 16)), b0 = lds128(gb_base + (e.z >> 16));
      const float cl = __uint_as_float(e.x), cr = __uint_as_float(e.y);
      const float dl = dot4(a0, vLA) + dot4(b0, vLB);
      const float dr = dot4(a0, vRA) + dot4(b0, vRB);
      float r = ((h & 2) ? dr : dl) + __shfl_xor_sync(kFull, (h & 2) ? dl : dr, 2);
      r += __shfl_xor_sync(kFull, r, 1);
      if (has && (h & 1) == 0) sts32(dot_base + (e.z & 0xffffu), r);
      if (has_l) {
        Gc::red_add(gvb + (ro_l + first), scale4(cl, a0));
        Gc::red_add(gvb + (ro_l + (first ^ 4)), scale4(cl, b0));
      }
      if (has_r) {
        Gc::red_add(gvb + (ro_r + first), scale4(cr, a0));
        Gc::red_add(gvb + (ro_r + (first ^ 4)), scale4(cr, b0));
      }
    }
  }
  __syncwarp();
  // nothing in this phase reads the counters: clear them for the next unit before the barrier that follows
  {
    const int t = threadIdx.x;
    const int Rz = sh.wbase[kTileL];
    for (int i = t; i < Rz; i += kTileThreads) sh.off[i] = 0;
    for (int i = t; i < kTileWarps * (kBins + 1); i += kTileThreads) (&sh.whist[0][0])[i] = 0;
    if (t < kTileL * 3) (&sh.anchor_sum[0][0])[t] = 0;
  }
}

// ---- epilogue: per-sample gradients (cuh:156-158) from the four dots of a sample.  kFuse: also the backward of the
// softmax and of reference + offset / (W, H). -----------------------------------------------------------------------
template <bool kFuse>
__device__ __noinline__ void tile_epilogue(const D32Params& p, int slot) {
  TileShared& sh = tile_smem();
  const int t = threadIdx.x;
  const TileUnit un = sh.unit[slot];
  if (t == 0) { sh.next_pass = 0; sh.n_left = 0; }      // (read by the key phase: reset behind its barrier)
  const int l = (t & 15) >> 2;
  const float Wf = (float)sh.tab.W[l], Hf = (float)sh.tab.H[l];
  const unsigned qb = (unsigned)un.b * (unsigned)p.Lq;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float4 rc = sh.rec[t + kTileThreads * i];
    const unsigned mk = __float_as_uint(rc.w);
    const bool live = (mk >> 5) != 0;
    float pa = 0.f, px = 0.f, py = 0.f;
    if (live && (mk & 15u)) {
      const float4 d = reinterpret_cast<const float4*>(sh.dbuf)[t + kTileThreads * i];
      const float d1 = (mk & 1u) ? d.x : 0.f, d2 = (mk & 2u) ? d.y : 0.f;
      const float d3 = (mk & 4u) ? d.z : 0.f, d4 = (mk & 8u) ? d.w : 0.f;
      const float hh = 1.f - rc.x, hw = 1.f - rc.y;
      pa = hh * hw * d1 + hh * rc.y * d2 + rc.x * hw * d3 + rc.x * rc.y * d4;
      px = (hh * (d2 - d1) + rc.x * (d4 - d3)) * rc.z * Wf;
      py = (hw * (d3 - d1) + rc.y * (d4 - d2)) * rc.z * Hf;
    }
    const unsigned q = qb + (mk >> 5) - 1u;
    const unsigned gs = (q * (unsigned)p.M + (unsigned)un.m) * kTileLP + (t & 15);
    if (!kFuse) {
      if (live) {
        st_stream(p.grad_aw + gs, pa);
        st_stream(reinterpret_cast<float2*>(p.grad_loc) + gs, make_float2(px, py));
      }
    } else {
      // (every lane takes part in the shuffles; absent queries carry zeros)
      const float a = live ? rc.z : 0.f;
      // softmax backward over the 16 samples of the (query, head): d/d logit_s = a_s (ga_s - sum_j a_j ga_j)
      float dot = a * pa;
#pragma unroll
      for (int o = 1; o < kTileLP; o <<= 1) dot += __shfl_xor_sync(kFull, dot, o);
      // d/d reference point of (query, level) = sum over points (here) and heads (other units: atomics)
      float rx = px, ry = py;
      rx += __shfl_xor_sync(kFull, rx, 1); ry += __shfl_xor_sync(kFull, ry, 1);
      rx += __shfl_xor_sync(kFull, rx, 2); ry += __shfl_xor_sync(kFull, ry, 2);
      if (live) {
        st_stream(p.grad_aw + gs, a * (pa - dot));
        st_stream(reinterpret_cast<float2*>(p.grad_loc) + gs, make_float2(__fdiv_rn(px, Wf), __fdiv_rn(py, Hf)));
        if (p.grad_ref && (t & 3) == 0) {
          float* gr = p.grad_ref + ((size_t)q * kTileL + l) * 2;
          atomicAdd(gr, rx);
          atomicAdd(gr + 1, ry);
        }
      }
    }
  }
}

template <typename VT, typename GT, bool kFuse = false>
__global__ void __launch_bounds__(kTileThreads, 2)
bwd_d32_tile_kernel(const __grid_constant__ D32Params p, const int grid_mode) {
  TileShared& sh = tile_smem();
  const int t = threadIdx.x;
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
    // a KEY is (level, row y, left column x0) with x0 = w_low in [-1, W - 1]: the row pair (y, x0), (y, x0 + 1)
    int keys_whole = 0;
    for (int l = 0; l < kTileL; ++l) keys_whole += sh.tab.H[l] * (sh.tab.W[l] + 1);
    const bool whole = keys_whole <= kWholeMaxKeys;
    int base = 0, tiles = 0, all_fixed = 1, q0 = 0;
    for (int l = 0; l < kTileL; ++l) {
      const int H = sh.tab.H[l], W = sh.tab.W[l];
      const int ww = whole ? W + 1 : min(W + 1, kWinDim), wh = whole ? H : min(H, kWinDim);
      sh.ww[l] = ww; sh.wh[l] = wh; sh.wbase[l] = base;
      sh.fixed[l] = (ww == W + 1 && wh == H);
      all_fixed &= sh.fixed[l];
      base += ww * wh;
      sh.tile0[l] = tiles;
      sh.qstart[l] = q0;
      sh.tiles_x[l] = (W + 7) >> 3;
      // a level whose pixels do not all fit into the query range gets no blocks: its queries go to the linear tail
      if (grid_mode && H > 0 && W > 0 && (long long)q0 + (long long)H * W <= p.Lq) {
        q0 += H * W;
        tiles += ((W + 7) >> 3) * ((H + 7) >> 3);
      }
    }
    sh.wbase[kTileL] = base;
    sh.tile0[kTileL] = tiles;
    sh.qstart[kTileL] = q0;
    tiles += (p.Lq - q0 + kTileQ - 1) / kTileQ;      // linear mode: q0 == 0, every tile is one of these
    sh.tile0[kTileL + 1] = tiles;
    sh.n_tiles = tiles;
    sh.all_fixed = all_fixed;
    sh.n_left = 0;
    sh.next_pass = 0;
    const unsigned nu = (unsigned)p.N * (unsigned)sh.n_tiles * (unsigned)M;
    if (blockIdx.x < nu) decode_unit(sh, 0, grid_mode, blockIdx.x, (unsigned)sh.n_tiles, M);
  }
  __syncthreads();
  const unsigned n_tiles = (unsigned)sh.n_tiles;
  const unsigned n_units = (unsigned)p.N * n_tiles * (unsigned)M;   // the host checked that this fits

  if (blockIdx.x < n_units) {
    request_samples(p, 0, grid_mode);
    request_grad_rows(p, 0, grid_mode);
  }
#if MSDA_TILE_EXPERIMENT == 9      // timing experiment: cycles per phase (thread 0 of a few CTAs prints them)
  long long tk[6] = {0, 0, 0, 0, 0, 0};
  long long t_prev = clock64();
  int n_done = 0;
#define MSDA_TICK(k) do { const long long now__ = clock64(); tk[k] += now__ - t_prev; t_prev = now__; } while (0)
#else
#define MSDA_TICK(k) do { } while (0)
#endif
  int slot = 0;
  for (unsigned u = blockIdx.x; u < n_units; u += gridDim.x, slot ^= 1) {
    const bool more = u + gridDim.x < n_units;
    if (t == 0 && more) decode_unit(sh, slot ^ 1, grid_mode, u + gridDim.x, n_tiles, M);
    tile_geometry<kFuse>(p, slot, grid_mode);
    MSDA_TICK(0);
    __syncthreads();                 // sample records, anchor sums, G, the next unit's slot
    MSDA_TICK(5);
    tile_sort(p, more ? (slot ^ 1) : -1, grid_mode);
    MSDA_TICK(1);
    __syncthreads();                 // entries and key list complete
    MSDA_TICK(5);
    {
      const TileUnit un = sh.unit[slot];
      // (< 2^31 chunks: host-checked)
      tile_keys<VT, GT>(p.value, p.grad_value, ((unsigned)un.b * (unsigned)p.S * (unsigned)M + (unsigned)un.m) * 8u, (unsigned)M * 8u);
    }
    MSDA_TICK(2);
    __syncthreads();                 // dots complete; G is free again
    MSDA_TICK(4);
    if (more) request_grad_rows(p, slot ^ 1, grid_mode);
    tile_epilogue<kFuse>(p, slot);
    MSDA_TICK(3);
    __syncthreads();
    MSDA_TICK(5);
#if MSDA_TILE_EXPERIMENT == 9
    ++n_done;
#endif
  }
#if MSDA_TILE_EXPERIMENT == 9
  if ((t == 0 || t == 480) && (blockIdx.x == 0 || blockIdx.x == 700))
    printf("cta %d thread %d: %d units; cycles/unit geometry %lld sort %lld keys %lld epilogue %lld | wait after keys %lld other barriers %lld\n",
           blockIdx.x, t, n_done, tk[0] / max(n_done, 1), tk[1] / max(n_done, 1), tk[2] / max(n_done, 1), tk[3] / max(n_done, 1),
           tk[4] / max(n_done, 1), tk[5] / max(n_done, 1));
#endif
}

inline size_t tile_smem_bytes() { return sizeof(TileShared); }

}  // namespace msda
