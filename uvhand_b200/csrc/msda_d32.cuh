// D = 32 kernels -- the shipped head geometry (util/settings.py:102-120: d_model 256 / 8 heads)
// and the performance path of this library.
//
// Work decomposition (forward and backward):
//   * an ITEM is one (sample b, query q, head m); it owns 32 output channels = one 128 B row.
//   * 8 lanes serve one item, each lane owns 4 consecutive channels, so one warp instruction
//     moves FOUR complete value rows (fp32: 4 x 128 B with LDG.128; bf16: 4 x 64 B with LDG.64)
//     and one red.global.add.v4.f32 per lane replaces four scalar atomics in the backward.
//   * a warp therefore works on 4 consecutive items (same query, 4 neighbouring heads when
//     M % 4 == 0), whose sampling locations / weights / outputs are contiguous in memory:
//     every streaming access of the warp is one fully coalesced segment.
//
// Two phases per warp pass, separated only by __syncwarp (no block barrier anywhere):
//   A  "geometry": the 4*L*P samples of the pass are spread over the 32 lanes (one thread per
//      sample, no redundancy); each computes sample_geom(), the four corner offsets and the
//      four bilinear weights, and parks them as a 32 B record in the warp's private slice of
//      shared memory.
//   B  "gather": every 8-lane group walks its item's L*P records (two broadcast LDS.128 per
//      sample), issues the 4 corner row loads, and accumulates.
//
// The reference instead recomputes all of phase A in each of the 32 channel threads and issues
// 64 predicated scalar loads per thread (cuh:272-296); its backward adds 2*L*P block barriers
// and a serial thread-0 reduction per block (cuh:376-394).
#pragma once

#include "msda_device.cuh"
#include "msda_generic.cuh"

namespace msda {

constexpr int kD32Warps = 8;                  // warps per CTA
constexpr int kD32Threads = kD32Warps * kWarp;
constexpr int kItemsPerWarp = 4;              // 8 lanes per item
constexpr int kMaxLevelsD32 = 8;
constexpr int kMaxLPD32 = 32;                 // L*P limit of the specialised kernels

struct __align__(16) CornerRec {
  float w[4];  // forward: bilinear weight * attention weight; backward: (lh, lw, a, unused)
  int off[4];  // offset of each corner row in Chunk units relative to the item's base; <0 = outside
};

struct LevelTab {
  int H[kMaxLevelsD32], W[kMaxLevelsD32], start[kMaxLevelsD32];
};

struct D32Params {
  const void* value;
  const int64_t* shapes;
  const int64_t* lsi;
  const float* loc;
  const float* aw;
  const float* grad_out;  // backward only
  float* out;             // forward only
  void* grad_value;       // backward only
  float* grad_loc;
  float* grad_aw;
  int N, S, M, L, Lq, P;
  long long n_items;
};

__device__ __forceinline__ void load_level_tab(LevelTab& tab, const D32Params& p) {
  if (threadIdx.x < p.L) {
    tab.H[threadIdx.x] = (int)p.shapes[2 * threadIdx.x];
    tab.W[threadIdx.x] = (int)p.shapes[2 * threadIdx.x + 1];
    tab.start[threadIdx.x] = (int)p.lsi[threadIdx.x];
  }
  __syncthreads();
}

// Phase A for one sample.  kBackward selects what goes into rec.w.
template <bool kBackward>
__device__ __forceinline__ void make_record(CornerRec& rec, float x, float y, float a, int H, int W,
                                            int start, int M) {
  SampleGeom<float> g;
  if (!sample_geom<float>(x, y, H, W, g)) {
    rec.w[0] = rec.w[1] = rec.w[2] = rec.w[3] = 0.f;
    rec.off[0] = rec.off[1] = rec.off[2] = rec.off[3] = -1;
    return;
  }
  const int chunks_per_token = M * 8;  // 8 chunks of 4 channels per head row
  const int r00 = (start + g.h_low * W + g.w_low) * chunks_per_token;
  rec.off[0] = (g.mask & 1u) ? r00 : -1;
  rec.off[1] = (g.mask & 2u) ? r00 + chunks_per_token : -1;
  rec.off[2] = (g.mask & 4u) ? r00 + W * chunks_per_token : -1;
  rec.off[3] = (g.mask & 8u) ? r00 + (W + 1) * chunks_per_token : -1;
  if (kBackward) {
    rec.w[0] = g.lh;
    rec.w[1] = g.lw;
    rec.w[2] = a;
    rec.w[3] = 0.f;
  } else {
    const float hh = 1.f - g.lh, hw = 1.f - g.lw;
    rec.w[0] = hh * hw * a;
    rec.w[1] = hh * g.lw * a;
    rec.w[2] = g.lh * hw * a;
    rec.w[3] = g.lh * g.lw * a;
  }
}

// Phase A for the whole warp pass: 4*LP samples over 32 lanes, coalesced loc / weight loads.
//
// Records are parked as two planes (weights, offsets) so that the 16 B stores of consecutive
// lanes and the broadcast loads of phase B are both free of bank conflicts.
template <bool kBackward>
__device__ __forceinline__ void geometry_phase(float4* rec_w, int4* rec_o, const D32Params& p,
                                               const LevelTab& tab, long long item0, int LP, int lane) {
  const int n_samples = kItemsPerWarp * LP;
  const long long s_base = item0 * LP;
  const long long s_end = p.n_items * LP;
  for (int s = lane; s < n_samples; s += kWarp) {
    const long long gs = s_base + s;
    if (gs >= s_end) break;
    const int ls = s % LP;
    const int l = ls / p.P;
    const float2 xy = ld_stream(reinterpret_cast<const float2*>(p.loc) + gs);
    const float a = ld_stream(p.aw + gs);
    CornerRec rec;
    make_record<kBackward>(rec, xy.x, xy.y, a, tab.H[l], tab.W[l], tab.start[l], p.M);
    rec_w[s] = make_float4(rec.w[0], rec.w[1], rec.w[2], rec.w[3]);
    rec_o[s] = make_int4(rec.off[0], rec.off[1], rec.off[2], rec.off[3]);
  }
}

__device__ __forceinline__ void fma4(float4& acc, float w, const float4& v) {
  acc.x = fmaf(w, v.x, acc.x);
  acc.y = fmaf(w, v.y, acc.y);
  acc.z = fmaf(w, v.z, acc.z);
  acc.w = fmaf(w, v.w, acc.w);
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
}

// ---------------------------------------------------------------------------------------------
// Forward
// ---------------------------------------------------------------------------------------------
template <typename VT, int kLP>  // kLP = compile-time L*P (0: runtime)
__global__ void __launch_bounds__(kD32Threads)
fwd_d32_kernel(const D32Params p) {
  using C = Chunk<VT>;
  using Vec = typename C::Vec;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ LevelTab tab;
  load_level_tab(tab, p);

  const int LP = kLP ? kLP : p.L * p.P;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = lane >> 3, sub = lane & 7;
  float4* rec_w = reinterpret_cast<float4*>(smem_raw) + warp * (2 * kItemsPerWarp * LP);
  int4* rec_o = reinterpret_cast<int4*>(rec_w + kItemsPerWarp * LP);
  const float4* my_w = rec_w + group * LP;
  const int4* my_o = rec_o + group * LP;

  const long long n_pass = (p.n_items + kItemsPerWarp - 1) / kItemsPerWarp;
  const long long pass_stride = (long long)gridDim.x * kD32Warps;
  for (long long pass = (long long)blockIdx.x * kD32Warps + warp; pass < n_pass; pass += pass_stride) {
    const long long item0 = pass * kItemsPerWarp;
    geometry_phase<false>(rec_w, rec_o, p, tab, item0, LP, lane);
    __syncwarp();

    const long long item = item0 + group;
    if (item < p.n_items) {
      const int m = (int)(item % p.M);
      const int b = (int)(item / ((long long)p.M * p.Lq));
      const Vec* vb = reinterpret_cast<const Vec*>(p.value) + ((long long)b * p.S * p.M + m) * 8 + sub;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int s = 0; s < LP; ++s) {
        const float4 w = my_w[s];
        const int4 o = my_o[s];
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0, v2 = v0, v3 = v0;
        if (o.x >= 0) v0 = C::load(vb + o.x);
        if (o.y >= 0) v1 = C::load(vb + o.y);
        if (o.z >= 0) v2 = C::load(vb + o.z);
        if (o.w >= 0) v3 = C::load(vb + o.w);
        fma4(acc, w.x, v0);
        fma4(acc, w.y, v1);
        fma4(acc, w.z, v2);
        fma4(acc, w.w, v3);
      }
      st_stream(reinterpret_cast<float4*>(p.out) + item * 8 + sub, acc);
    }
    __syncwarp();  // records are rewritten by the next pass
  }
}

// ---------------------------------------------------------------------------------------------
// Backward
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float group8_sum(float v) {
  v += __shfl_xor_sync(kFull, v, 4);
  v += __shfl_xor_sync(kFull, v, 2);
  v += __shfl_xor_sync(kFull, v, 1);
  return v;
}

template <typename VT, int kLP>
__global__ void __launch_bounds__(kD32Threads)
bwd_d32_kernel(const D32Params p) {
  using C = Chunk<VT>;
  using Vec = typename C::Vec;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ LevelTab tab;
  load_level_tab(tab, p);

  const int LP = kLP ? kLP : p.L * p.P;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = lane >> 3, sub = lane & 7;
  float4* rec_w = reinterpret_cast<float4*>(smem_raw) + warp * (2 * kItemsPerWarp * LP);
  int4* rec_o = reinterpret_cast<int4*>(rec_w + kItemsPerWarp * LP);
  const float4* my_w = rec_w + group * LP;
  const int4* my_o = rec_o + group * LP;

  const long long n_pass = (p.n_items + kItemsPerWarp - 1) / kItemsPerWarp;
  const long long pass_stride = (long long)gridDim.x * kD32Warps;
  for (long long pass = (long long)blockIdx.x * kD32Warps + warp; pass < n_pass; pass += pass_stride) {
    const long long item0 = pass * kItemsPerWarp;
    geometry_phase<true>(rec_w, rec_o, p, tab, item0, LP, lane);
    __syncwarp();

    // Groups past the end still take part in the shuffles (with zero data).
    const long long item = item0 + group;
    const bool live = item < p.n_items;
    const long long item_c = live ? item : (p.n_items - 1);
    const int m = (int)(item_c % p.M);
    const int b = (int)(item_c / ((long long)p.M * p.Lq));
    const long long head_off = ((long long)b * p.S * p.M + m) * 8 + sub;
    const Vec* vb = reinterpret_cast<const Vec*>(p.value) + head_off;
    Vec* gvb = reinterpret_cast<Vec*>(p.grad_value) + head_off;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) g = ld_stream(reinterpret_cast<const float4*>(p.grad_out) + item * 8 + sub);

    for (int s = 0; s < LP; ++s) {
      const float4 r = my_w[s];  // lh, lw, a
      const int4 o = my_o[s];
      const int l = s / p.P;
      const float lh = r.x, lw = r.y, a = r.z;
      const float hh = 1.f - lh, hw = 1.f - lw;
      const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
      float4 v1 = make_float4(0.f, 0.f, 0.f, 0.f), v2 = v1, v3 = v1, v4 = v1;
      const bool ok1 = live && o.x >= 0, ok2 = live && o.y >= 0, ok3 = live && o.z >= 0,
                 ok4 = live && o.w >= 0;
      if (ok1) v1 = C::load(vb + o.x);
      if (ok2) v2 = C::load(vb + o.y);
      if (ok3) v3 = C::load(vb + o.z);
      if (ok4) v4 = C::load(vb + o.w);
      const float4 t = make_float4(g.x * a, g.y * a, g.z * a, g.w * a);
      if (ok1) C::red_add(gvb + o.x, make_float4(w1 * t.x, w1 * t.y, w1 * t.z, w1 * t.w));
      if (ok2) C::red_add(gvb + o.y, make_float4(w2 * t.x, w2 * t.y, w2 * t.z, w2 * t.w));
      if (ok3) C::red_add(gvb + o.z, make_float4(w3 * t.x, w3 * t.y, w3 * t.z, w3 * t.w));
      if (ok4) C::red_add(gvb + o.w, make_float4(w4 * t.x, w4 * t.y, w4 * t.z, w4 * t.w));
      const float d1 = dot4(g, v1), d2 = dot4(g, v2), d3 = dot4(g, v3), d4 = dot4(g, v4);
      float ga = w1 * d1 + w2 * d2 + w3 * d3 + w4 * d4;
      float gx = (hh * (d2 - d1) + lh * (d4 - d3)) * a * (float)tab.W[l];
      float gy = (hw * (d3 - d1) + lw * (d4 - d2)) * a * (float)tab.H[l];
      ga = group8_sum(ga);
      gx = group8_sum(gx);
      gy = group8_sum(gy);
      if (live && sub == (s & 7)) {
        st_stream(reinterpret_cast<float2*>(p.grad_loc) + item * LP + s, make_float2(gx, gy));
        st_stream(p.grad_aw + item * LP + s, ga);
      }
    }
    __syncwarp();
  }
}

}  // namespace msda
