// Backward for SMALL pyramids (H2OTR / ARCTIC: 224^2 crops -> levels 28^2, 14^2, 7^2, 4^2): the
// coarse levels are accumulated ON CHIP.
//
// The plain backward sends one 128 B reduction per sample corner to L2 and is bound by the SM's
// port into L2.  When a level is small, all of it -- for one (sample, head) -- fits in
// a few KB: at the ARCTIC shape levels 2 and 3 together are 65 rows = 8.1 KB, and they receive HALF
// of all corner reductions.  Here every warp keeps a private fp32 copy ("map") of those levels for
// the (sample, head) it is working on, adds into it with plain LDS/FADD/STS, and sends each map row
// to L2 once when it moves on.  No locality assumption: the whole level is resident, so decoder
// queries (random references) profit like encoder queries.
//
// What makes the plain read-modify-write safe (shared-memory float atomics would cost as much as the
// L2 reductions they replace, profiles/r01_microbench_smem.md):
//   * a map belongs to ONE warp -- no other warp ever touches it;
//   * inside the warp the 4 groups work on 4 QUERIES OF THE SAME HEAD and may hit the same row in
//     the same step, so their map updates run one group after the other (4 predicated passes);
//     the 4 corners of one sample are distinct rows and go together.
// Work split: the flattened (sample, head, query) space is cut into equal contiguous ranges, one per
// warp, so a warp changes (sample, head) -- and flushes -- once or twice in its life.
//
// Levels that do not fit keep using red.global, so the kernel is correct for any shapes.  L = 4, P = 4, fp32.
//
// RESULT (B200, ARCTIC shape N=32, S=1045): correct, and SLOWER than the plain backward --
//   encoder 402 us vs 306 us, decoder (300 queries) 125 us vs 70 us
//   same kernel with the maps disabled (all reductions to L2): 326 / 84 us  (cost of the 4-queries-per-warp order)
//   same kernel updating only one group in four (racy, 1/4 of the map work): 315 / 96 us
// i.e. taking half of the reductions off the SM->L2 port saves ~11 us, and the serialised shared-memory
// updates that replace them cost ~90 us: each map row costs ~3.9 SM cycles (one group at a time, LDS -> FMA
// -> STS chains, 24 warps per SM because of the 75 KB of maps) against ~2.6 for the reduction it replaces.
// It therefore is NOT selected automatically (option "bwd_maps" = 2 runs it); it stays in the tree, tested,
// as the measured basis of DESIGN.md 3.3.
#pragma once

#include "msda_d32.cuh"

namespace msda {

constexpr int kMapRows = 65;              // map rows (128 B each) per warp: 8.1 KB = ARCTIC levels 2 + 3 (49 + 16);
                                          // with the records 75 KB per CTA -> 3 CTAs per SM
constexpr int kMapsCtasPerSm = 3;

struct MapTab {
  int base[kMaxLevelsD32];   // first map row of level l, or -1 when the level stays on the red.global path
  int rows[kMaxLevelsD32];
  int used;
};

__host__ __device__ constexpr size_t maps_smem_bytes(int LP) {
  return (size_t)kD32Warps * (kItemsPerWarp * rec_stride(LP) * 16 + kMapRows * 128);
}

// record for this kernel: token index (not yet times M) << 4 | corner mask, lh, lw, a
__device__ __forceinline__ int4 make_token_record(float x, float y, float a, int H, int W) {
  SampleGeom<float> g;
  if (!sample_geom<float>(x, y, H, W, g)) return make_int4(0, 0, 0, 0);
  int4 r;
  r.x = (g.h_low * W + g.w_low) * 16 + (int)g.mask;   // token inside its level; negative at the top/left border
  r.y = __float_as_int(g.lh);
  r.z = __float_as_int(g.lw);
  r.w = __float_as_int(a);
  return r;
}

template <int kL, int kP, int kM>
__global__ void __launch_bounds__(kD32Threads, kMapsCtasPerSm)
bwd_d32_maps_kernel(const D32Params p) {
  static_assert(kL > 0 && kP == 4, "specialised for 4 points per level");
  constexpr int LP = kL * kP;
  constexpr int stride = rec_stride(LP);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ LevelTab tab;
  __shared__ MapTab maps;
  load_level_tab(tab, p);
  if (threadIdx.x == 0) {
    int used = 0;
    for (int l = 0; l < kL; ++l) maps.base[l] = -1;
    for (int l = kL - 1; l >= 0; --l) {   // coarsest first: most reductions per row
      const long long rows = (long long)tab.H[l] * tab.W[l];
      maps.rows[l] = 0;
      if (tab.H[l] > 0 && tab.W[l] > 0 && rows <= kMapRows - used) {
        maps.base[l] = used;
        maps.rows[l] = (int)rows;
        used += (int)rows;
      }
    }
    maps.used = used;
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = lane >> 3, sub = lane & 7;
  const int M = kM > 0 ? kM : p.M;
  int4* warp_recs = reinterpret_cast<int4*>(smem_raw) + warp * (kItemsPerWarp * stride);
  const int4* my_recs = warp_recs + group * stride;
  float4* map = reinterpret_cast<float4*>(smem_raw + (size_t)kD32Warps * kItemsPerWarp * stride * 16) +
                (size_t)warp * kMapRows * 8;

  // this warp's contiguous share of the flattened (sample*head, query) space
  const long long total = (long long)p.N * M * p.Lq;
  const long long n_warps = (long long)gridDim.x * kD32Warps;
  const long long per = (total + n_warps - 1) / n_warps;
  const long long f0 = ((long long)blockIdx.x * kD32Warps + warp) * per;
  const long long f1 = f0 + per < total ? f0 + per : total;

  long long cur_bm = -1;   // the (sample, head) the map currently belongs to

  auto flush = [&](long long bm) {
    // every map row -> one vector reduction into grad_value; untouched chunks are skipped
    const int b = (int)(bm / M), m = (int)(bm % M);
    float4* gv = reinterpret_cast<float4*>(p.grad_value) + ((long long)b * p.S * M + m) * 8 + sub;
#pragma unroll
    for (int l = 0; l < kL; ++l) {
      const int base = maps.base[l];
      if (base < 0) continue;
      for (int r = group; r < maps.rows[l]; r += kItemsPerWarp) {
        const float4 v = map[(base + r) * 8 + sub];
        if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f)
          Chunk<float>::red_add(gv + (long long)(tab.start[l] + r) * M * 8, v);
      }
    }
  };
  auto clear = [&]() {
    for (int i = lane; i < maps.used * 8; i += kWarp) map[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  };

  for (long long f = f0; f < f1;) {
    const long long bm = f / p.Lq;
    const int q0 = (int)(f - bm * p.Lq);
    if (bm != cur_bm) {
      if (cur_bm >= 0) flush(cur_bm);
      __syncwarp();
      clear();
      cur_bm = bm;
    }
    // up to 4 queries of this (sample, head), not past its last query nor past this warp's range
    long long n_here = p.Lq - q0;
    if (n_here > f1 - f) n_here = f1 - f;
    if (n_here > kItemsPerWarp) n_here = kItemsPerWarp;
    const int b = (int)(bm / M), m = (int)(bm % M);

    // phase A: 4 * LP samples over the lanes, one thread per sample
    for (int s = lane; s < kItemsPerWarp * LP; s += kWarp) {
      const int grp = s / LP, ls = s % LP, l = ls / kP;
      if (grp < n_here) {
        const long long gs = ((((long long)b * p.Lq + q0 + grp) * M) + m) * LP + ls;
        const float2 xy = ld_stream(reinterpret_cast<const float2*>(p.loc) + gs);
        const float a = ld_stream(p.aw + gs);
        warp_recs[grp * stride + ls] = make_token_record(xy.x, xy.y, a, tab.H[l], tab.W[l]);
      }
    }
    __syncwarp();

    const bool live = group < n_here;
    const long long item = (((long long)b * p.Lq + q0 + (live ? group : 0)) * M) + m;
    const long long head_off = ((long long)b * p.S * M + m) * 8 + sub;
    const float4* vb = reinterpret_cast<const float4*>(p.value) + head_off;
    float4* gvb = reinterpret_cast<float4*>(p.grad_value) + head_off;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) g = ld_stream(reinterpret_cast<const float4*>(p.grad_out) + item * 8 + sub);
    float2* gl_out = reinterpret_cast<float2*>(p.grad_loc) + item * LP;
    float* ga_out = p.grad_aw + item * LP;

#pragma unroll
    for (int l = 0; l < kL; ++l) {
      const int W = tab.W[l];
      const float Wf = (float)W, Hf = (float)tab.H[l];
#ifdef MSDA_MAPS_EXPERIMENT_NO_MAPS        // timing experiment: everything through red.global
      const int mbase = -1;
#else
      const int mbase = maps.base[l];          // warp-uniform
#endif
      const long long lvl = (long long)tab.start[l] * M * 8;
      const int dx = M * 8;
      float pa[4], px[4], py[4];
      // gathers, dot products and -- for a level that is not on chip -- the reductions, as in bwd_d32_kernel
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int4 r = live ? my_recs[l * kP + j] : make_int4(0, 0, 0, 0);
        const float lh = __int_as_float(r.y), lw = __int_as_float(r.z), a = __int_as_float(r.w);
        const float hh = 1.f - lh, hw = 1.f - lw;
        const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
        const long long o0 = lvl + (long long)(r.x >> 4) * M * 8, o1 = o0 + (long long)W * M * 8;
        const bool ok1 = r.x & 1, ok2 = r.x & 2, ok3 = r.x & 4, ok4 = r.x & 8;
        float4 v1 = make_float4(0.f, 0.f, 0.f, 0.f), v2 = v1, v3 = v1, v4 = v1;
        if (ok1) v1 = Chunk<float>::load(vb + o0);
        if (ok2) v2 = Chunk<float>::load(vb + o0 + dx);
        if (ok3) v3 = Chunk<float>::load(vb + o1);
        if (ok4) v4 = Chunk<float>::load(vb + o1 + dx);
        if (mbase < 0) {
          const float4 t = scale4(a, g);
          if (ok1) Chunk<float>::red_add(gvb + o0, scale4(w1, t));
          if (ok2) Chunk<float>::red_add(gvb + o0 + dx, scale4(w2, t));
          if (ok3) Chunk<float>::red_add(gvb + o1, scale4(w3, t));
          if (ok4) Chunk<float>::red_add(gvb + o1 + dx, scale4(w4, t));
        }
        const float d1 = dot4(g, v1), d2 = dot4(g, v2), d3 = dot4(g, v3), d4 = dot4(g, v4);
        pa[j] = w1 * d1 + w2 * d2 + w3 * d3 + w4 * d4;
        px[j] = (hh * (d2 - d1) + lh * (d4 - d3)) * a * Wf;
        py[j] = (hw * (d3 - d1) + lw * (d4 - d2)) * a * Hf;
      }
      if (mbase >= 0) {
        // map updates: one group at a time (two queries may hit the same row), the 4 points of a
        // group in order (they may too), the 4 corners of a point together (distinct rows).
        // Weights are rebuilt from the record instead of being kept alive across the gathers.
#pragma unroll
        for (int gsel = 0; gsel < kItemsPerWarp; ++gsel) {
#ifdef MSDA_MAPS_EXPERIMENT_RACY           // timing experiment: a quarter of the map work, wrong results
          if (gsel == 0 && live) {
#else
          if (group == gsel && live) {
#endif
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int4 r = my_recs[l * kP + j];
              const float lh = __int_as_float(r.y), lw = __int_as_float(r.z);
              const float hh = 1.f - lh, hw = 1.f - lw;
              const float4 t = scale4(__int_as_float(r.w), g);
              float4* m00 = map + (mbase + (r.x >> 4)) * 8 + sub;
              // the 4 corners are distinct rows: read all, then update, then write (one latency, not four)
              const bool k1 = r.x & 1, k2 = r.x & 2, k3 = r.x & 4, k4 = r.x & 8;
              float4 a1, a2, a3, a4;
              if (k1) a1 = m00[0];
              if (k2) a2 = m00[8];
              if (k3) a3 = m00[W * 8];
              if (k4) a4 = m00[W * 8 + 8];
              if (k1) { fma4(a1, hh * hw, t); m00[0] = a1; }
              if (k2) { fma4(a2, hh * lw, t); m00[8] = a2; }
              if (k3) { fma4(a3, lh * hw, t); m00[W * 8] = a3; }
              if (k4) { fma4(a4, lh * lw, t); m00[W * 8 + 8] = a4; }
            }
          }
          __syncwarp();
        }
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
    __syncwarp();
    f += n_here;
  }
  if (cur_bm >= 0) {
    __syncwarp();
    flush(cur_bm);
  }
}

}  // namespace msda
