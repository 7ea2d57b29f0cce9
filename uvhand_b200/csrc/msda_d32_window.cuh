// fp32 backward for ENCODER self-attention (query i is value token i, so queries can be visited as
// pixel tiles): the value rows a tile needs are staged in shared memory once, and the gathers of the
// backward read them from there.
//
// Why (profiles/r01_tuning_matrix.md, compiled-out experiments): the backward is bound by the SM's port
// into L2, which carries the grad_value reductions (433 us by themselves at the DETR-encoder shape) AND
// the read requests of the value gathers that miss L1 (another ~80 us; L1 hit rate 17 %).  With the
// gathers served from shared memory the kernel runs at the reduction floor.
//
// Unit = (sample b, head m, tile of 8 x 8 query pixels of one level); one CTA of 16 warps per unit, one
// 8-lane group per query.  Per unit:
//   1. thread 0 lays out the WINDOW: for every level the tile's footprint (its pixel-centre references
//      mapped to that level) plus a halo of kWinHalo pixels;
//   2. all threads copy those rows of value[b, :, m, :] with cp.async (each row once, 128 B);
//   3. phase A as in msda_d32.cuh, plus for every sample the window row of its (h_low, w_low) corner, or
//      -1 when any of its 4 corners falls outside the window -- such samples gather from global memory
//      exactly like the plain kernel, so results never depend on where the offsets point;
//   4. phase B as in bwd_d32_kernel.  Reductions go to L2 as before.
// The tile is only a visiting order and the window only a cache: every query is visited once and every
// sample computes with the same arithmetic as the plain kernel.
//
// RESULT (B200): parity-green in every test family, and SLOWER than bwd_d32_kernel --
//   DETR encoder 588 us vs 498 us (uniform locations, window useless: 606 vs 558), ARCTIC encoder 414 vs 309 us,
// although the upper-bound experiment (gathers redirected to shared memory inside the plain kernel) reaches
// 434 us.  The window itself pays (588 vs 606), what it rides on does not: the per-unit structure (two CTA
// barriers, a serial window layout, 16-warp CTAs, ragged tile edges, 4 queries of one head per warp) costs
// ~50 us even when the window is useless, and neighbouring tiles running at the same time send their
// coarse-level reductions to the same rows (striding the tile order across the image recovers 22 us: 564).
// Every tile-ordered backward of this round lost for these reasons (profiles/r01_tuning_matrix.md).
// NOT selected automatically (option "bwd_window" = 2 runs it).
#pragma once

#include "msda_d32.cuh"

namespace msda {

constexpr int kWinWarps = 16;
constexpr int kWinThreads = kWinWarps * kWarp;
constexpr int kWinTileLog2 = 3;            // 8 x 8 query pixels = 64 slots = 16 warps x 4 groups
constexpr int kWinHalo = 4;                // pixels around the tile's footprint, every level
constexpr int kWinRows = 600;              // window capacity in rows of 128 B (8x8 tile, halo 4: 581)
constexpr int kWinLevels = 4;

struct WinTab {
  int H[kWinLevels], W[kWinLevels], start[kWinLevels];
  int qstart[kWinLevels];                  // first query of level l in the pyramid order
  int tiles_x[kWinLevels], tfirst[kWinLevels + 1];
  int n_tiles;
  int spatial;                             // 1: Lq == sum_l H_l*W_l, tiles are pixel blocks; 0: chunks of 64 queries, no window
};

struct WinGeom {                            // the current unit's window, per level
  int y0[kWinLevels], x0[kWinLevels], h[kWinLevels], w[kWinLevels], base[kWinLevels];
  int rows;
};

__host__ __device__ constexpr size_t window_smem_bytes() {
  return (size_t)kWinWarps * kItemsPerWarp * rec_stride(16) * 16   // sample records
         + (size_t)kWinWarps * kItemsPerWarp * 16 * 4              // window row of every sample
         + (size_t)kWinRows * 128;                                 // the window
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}

template <int kM>
__global__ void __launch_bounds__(kWinThreads, 2)
bwd_d32_window_kernel(const D32Params p) {
  constexpr int kL = kWinLevels, kP = 4, LP = kL * kP;
  constexpr int stride = rec_stride(LP);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ WinTab tab;
  __shared__ WinGeom geo;
  const int M = kM > 0 ? kM : p.M;
  if (threadIdx.x == 0) {
    int t = 0, q = 0;
    for (int l = 0; l < kL; ++l) {
      const int h = (int)p.shapes[2 * l], w = (int)p.shapes[2 * l + 1];
      tab.H[l] = h;
      tab.W[l] = w;
      tab.start[l] = (int)p.lsi[l];
      tab.qstart[l] = q;
      q += (h > 0 && w > 0) ? h * w : 0;
      tab.tiles_x[l] = w > 0 ? (w + (1 << kWinTileLog2) - 1) >> kWinTileLog2 : 0;
      tab.tfirst[l] = t;
      t += tab.tiles_x[l] * (h > 0 ? (h + (1 << kWinTileLog2) - 1) >> kWinTileLog2 : 0);
    }
    tab.tfirst[kL] = t;
    tab.spatial = (q == p.Lq && t > 0) ? 1 : 0;
    tab.n_tiles = tab.spatial ? t : (p.Lq + 63) / 64;
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = lane >> 3, sub = lane & 7;
  int4* warp_recs = reinterpret_cast<int4*>(smem_raw) + warp * (kItemsPerWarp * stride);
  const int4* my_recs = warp_recs + group * stride;
  int* all_wrows = reinterpret_cast<int*>(smem_raw + (size_t)kWinWarps * kItemsPerWarp * stride * 16);
  int* warp_wrows = all_wrows + warp * (kItemsPerWarp * LP);
  const int* my_wrows = warp_wrows + group * LP;
  float4* window = reinterpret_cast<float4*>(smem_raw + (size_t)kWinWarps * kItemsPerWarp * stride * 16 +
                                             (size_t)kWinWarps * kItemsPerWarp * LP * 4);

  const long long n_units = (long long)p.N * M * tab.n_tiles;
  for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
    const int m = (int)(u % M);
    const long long rest = u / M;
#ifdef MSDA_EXPERIMENT_WINDOW_STRIDE     // timing experiment: concurrently running CTAs far apart in the image
    const int t = (int)(((rest % tab.n_tiles) * MSDA_EXPERIMENT_WINDOW_STRIDE) % tab.n_tiles);
#else
    const int t = (int)(rest % tab.n_tiles);
#endif
    const int b = (int)(rest / tab.n_tiles);
    int lq = 0, qy0 = 0, qx0 = 0, Hq = 1, Wq = 1;
    if (tab.spatial) {
      while (lq + 1 < kL && t >= tab.tfirst[lq + 1]) ++lq;
      const int tl = t - tab.tfirst[lq];
      const int ty = tl / tab.tiles_x[lq], tx = tl - ty * tab.tiles_x[lq];
      qy0 = ty << kWinTileLog2;
      qx0 = tx << kWinTileLog2;
      Hq = tab.H[lq];
      Wq = tab.W[lq];
    }
    // query served by a slot of this unit, or -1
    auto slot_query = [&](int slot) {
      if (tab.spatial) {
        const int r = slot >> kWinTileLog2, c = slot & ((1 << kWinTileLog2) - 1);
        return (qy0 + r < Hq && qx0 + c < Wq) ? tab.qstart[lq] + (qy0 + r) * Wq + qx0 + c : -1;
      }
      const long long q = (long long)t * 64 + slot;
      return q < p.Lq ? (int)q : -1;
    };

    __syncthreads();   // everybody is done with the previous unit's window and geometry
    if (threadIdx.x == 0) {
      // footprint of the tile's pixel-centre references in every level, plus the halo; the window rows
      // of a level are [y0, y0+h) x [x0, x0+w) (may stick out of the level: those rows are never used)
      const int qy1 = min(qy0 + (1 << kWinTileLog2), Hq), qx1 = min(qx0 + (1 << kWinTileLog2), Wq);
      int rows = 0;
      // the widest halo (<= kWinHalo) with which all levels fit; levels that still do not fit gather globally
      for (int halo = kWinHalo; halo >= 0; --halo) {
        rows = 0;
        bool all_fit = true;
        for (int l = 0; l < kL; ++l) {
          int ylo = 0, xlo = 0, h = 0, w = 0;
          if (tab.spatial && tab.H[l] > 0 && tab.W[l] > 0) {
            const float sy = (float)tab.H[l] / (float)Hq, sx = (float)tab.W[l] / (float)Wq;
            ylo = (int)floorf((qy0 + 0.5f) * sy - 0.5f) - halo;
            xlo = (int)floorf((qx0 + 0.5f) * sx - 0.5f) - halo;
            h = (int)floorf((qy1 - 0.5f) * sy - 0.5f) + 1 + halo - ylo + 1;
            w = (int)floorf((qx1 - 0.5f) * sx - 0.5f) + 1 + halo - xlo + 1;
            if ((long long)h * w > kWinRows - rows) {
              h = w = 0;
              all_fit = false;
            }
          }
          geo.y0[l] = ylo; geo.x0[l] = xlo; geo.h[l] = h; geo.w[l] = w; geo.base[l] = rows;
          rows += h * w;
        }
        if (all_fit) break;
      }
      geo.rows = rows;
    }
    __syncthreads();

    // stage the window: one 8-lane group per row, 16 B per lane
    {
      const char* vbase = reinterpret_cast<const char*>(p.value) + ((long long)b * p.S * M + m) * 128 + sub * 16;
#pragma unroll
      for (int l = 0; l < kL; ++l) {
        const int w = geo.w[l], n = geo.h[l] * w;
        for (int i = threadIdx.x >> 3; i < n; i += kWinThreads / 8) {
          const int ry = i / w, rx = i - ry * w;
          const int y = geo.y0[l] + ry, x = geo.x0[l] + rx;
          if (y >= 0 && y < tab.H[l] && x >= 0 && x < tab.W[l])
            cp_async16(reinterpret_cast<char*>(window + (geo.base[l] + i) * 8) + sub * 16,
                       vbase + (long long)(tab.start[l] + y * tab.W[l] + x) * M * 128);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }

    // phase A while the copies fly: this warp's 4 queries, one thread per sample
    const int slot0 = warp * kItemsPerWarp;
    for (int s = lane; s < kItemsPerWarp * LP; s += kWarp) {
      const int grp = s / LP, ls = s % LP, l = ls / kP;
      const int q = slot_query(slot0 + grp);
      if (q >= 0) {
        const long long gs = ((((long long)b * p.Lq + q) * M) + m) * LP + ls;
        const float2 xy = ld_stream(reinterpret_cast<const float2*>(p.loc) + gs);
        const float a = ld_stream(p.aw + gs);
        SampleGeom<float> g;
        int4 rec = make_int4(0, 0, 0, 0);
        int wrow = -1;
        if (sample_geom<float>(xy.x, xy.y, tab.H[l], tab.W[l], g)) {
          rec.x = (tab.start[l] + g.h_low * tab.W[l] + g.w_low) * M * 16 + (int)g.mask;
          rec.y = __float_as_int(g.lh);
          rec.z = __float_as_int(g.lw);
          rec.w = __float_as_int(a);
          const int ry = g.h_low - geo.y0[l], rx = g.w_low - geo.x0[l];
          if (ry >= 0 && ry + 1 < geo.h[l] && rx >= 0 && rx + 1 < geo.w[l]) wrow = geo.base[l] + ry * geo.w[l] + rx;
        }
        warp_recs[grp * stride + ls] = rec;
        warp_wrows[grp * LP + ls] = wrow;
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();   // window complete, records visible

    // phase B
    const int q_mine = slot_query(slot0 + group);
    const bool live = q_mine >= 0;
    const int q = live ? q_mine : 0;
    const long long item = (((long long)b * p.Lq + q) * M) + m;
    const long long head_off = ((long long)b * p.S * M + m) * 8 + sub;
    const float4* vb = reinterpret_cast<const float4*>(p.value) + head_off;
    float4* gvb = reinterpret_cast<float4*>(p.grad_value) + head_off;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) g = ld_stream(reinterpret_cast<const float4*>(p.grad_out) + item * 8 + sub);
    float2* gl_out = reinterpret_cast<float2*>(p.grad_loc) + item * LP;
    float* ga_out = p.grad_aw + item * LP;
    const int dx = M * 8;

#pragma unroll
    for (int l = 0; l < kL; ++l) {
      const int wm = tab.W[l] * M;
      const float Wf = (float)tab.W[l], Hf = (float)tab.H[l];
      const int ww = geo.w[l] * 8;              // window row stride of this level, in float4
      float pa[4], px[4], py[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int4 rec = live ? my_recs[l * kP + j] : make_int4(0, 0, 0, 0);
        const int wrow = live ? my_wrows[l * kP + j] : -1;
        const float lh = __int_as_float(rec.y), lw = __int_as_float(rec.z), a = __int_as_float(rec.w);
        const float hh = 1.f - lh, hw = 1.f - lw;
        const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
        const long long o0 = (long long)(rec.x >> 4) * 8, o1 = o0 + (long long)wm * 8;
        const bool ok1 = rec.x & 1, ok2 = rec.x & 2, ok3 = rec.x & 4, ok4 = rec.x & 8;
        float4 v1 = make_float4(0.f, 0.f, 0.f, 0.f), v2 = v1, v3 = v1, v4 = v1;
        if (wrow >= 0) {
          const float4* wp = window + wrow * 8 + sub;
          if (ok1) v1 = wp[0];
          if (ok2) v2 = wp[8];
          if (ok3) v3 = wp[ww];
          if (ok4) v4 = wp[ww + 8];
        } else {
          if (ok1) v1 = Chunk<float>::load(vb + o0);
          if (ok2) v2 = Chunk<float>::load(vb + o0 + dx);
          if (ok3) v3 = Chunk<float>::load(vb + o1);
          if (ok4) v4 = Chunk<float>::load(vb + o1 + dx);
        }
        const float4 t = scale4(a, g);
        if (ok1) Chunk<float>::red_add(gvb + o0, scale4(w1, t));
        if (ok2) Chunk<float>::red_add(gvb + o0 + dx, scale4(w2, t));
        if (ok3) Chunk<float>::red_add(gvb + o1, scale4(w3, t));
        if (ok4) Chunk<float>::red_add(gvb + o1 + dx, scale4(w4, t));
        const float d1 = dot4(g, v1), d2 = dot4(g, v2), d3 = dot4(g, v3), d4 = dot4(g, v4);
        pa[j] = w1 * d1 + w2 * d2 + w3 * d3 + w4 * d4;
        px[j] = (hh * (d2 - d1) + lh * (d4 - d3)) * a * Wf;
        py[j] = (hw * (d3 - d1) + lw * (d4 - d2)) * a * Hf;
      }
      const float ta = group8_sum4(pa, sub);
      const float tx = group8_sum4(px, sub);
      const float ty = group8_sum4(py, sub);
      const int s = l * kP + (sub >> 1);
      if (live) {
        if (sub & 1) st_stream(ga_out + s, ta);
        else st_stream(gl_out + s, make_float2(tx, ty));
      }
    }
  }
}

}  // namespace msda
