// Forward for D = 32, L = P = 4 with EIGHT channels per lane ("wide" forward).
// Replaces ms_deformable_im2col_gpu_kernel (/root/reference/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299).
//
// fwd_d32_kernel gives 8 lanes to one (query, head) item: 3.4 warp instructions per gathered 128 B row, and the
// round-1 decomposition showed half of its time is instruction issue (profiles/r01_tuning_matrix.md).  Here 4 lanes serve
// an item and a warp takes 8 consecutive items (= the 8 heads of ONE query when M = 8):
//   * fp32 value rows are read with 256-bit loads (sm_100 LDG.E.256: 4 lanes x 32 B = one row), bf16 rows with
//     128-bit loads (8 bf16 per lane), so one warp instruction moves EIGHT rows -- a corner of all 8 heads is one
//     contiguous 1 KB piece of `value`;
//   * record decode, weights, addresses and predicates are paid once per 8 rows instead of once per 4.
// Phase A (one thread per bilinear sample, 16 B records in the warp's shared-memory slice) is the same as in
// fwd_d32_kernel; records are padded to 17 per item so that the eight 4-lane groups read from distinct banks.
#pragma once

#include "msda_d32.cuh"

namespace msda {

#ifndef MSDA_FWD_WIDE_MINB
#define MSDA_FWD_WIDE_MINB 4          // resident CTAs per SM (64 registers)
#endif
// Grid cap in units of "CTAs the device holds at once" and the one-pass-ahead request of the raw samples.  Measured at
// C2 / C5 bf16 (us): 8 waves 111.5 / 1703 (round 2 until late), 1 wave (persistent warps, ~9 passes each) 106.5 / 1597,
// 1 wave + prefetch 110.7 / 1664 (16 B of spills at 64 registers; 105.5 / 1596 at 3 CTAs per SM), 2 waves + prefetch
// 114.8 / 1737; 1 wave at 3 / 5 / 6 CTAs per SM 106.5 / 114.9 / 114.7.
#ifndef MSDA_FWD_WIDE_PREFETCH
#define MSDA_FWD_WIDE_PREFETCH 0
#endif
#ifndef MSDA_FWD_WIDE_WAVES
#define MSDA_FWD_WIDE_WAVES 1
#endif

constexpr int kWideItems = 8;          // items per warp pass
constexpr int kWideLP = 16;
constexpr int kWideStride = kWideLP + 1;

struct Row8 {
  float4 lo, hi;
};

template <typename VT>
struct Wide;
#ifndef MSDA_FWD_WIDE_F32_V8
#define MSDA_FWD_WIDE_F32_V8 0         // 1: one 256-bit load per lane and row (measured slower: LDG.E.ENL2.256 does not use L1)
#endif
template <>
struct Wide<float> {
  using Vec = float4;                  // row = 8 of these
#if MSDA_FWD_WIDE_F32_V8
  // a lane owns Vec 2*sub and 2*sub+1
  static __device__ __forceinline__ Row8 load(const Vec* p) {
    MSDA_CHK(p, 32, 0);
    Row8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z),
                   "=f"(r.hi.w)
                 : "l"(p));
    return r;
  }
  static constexpr int kVecPerLane = 2, kSecond = 1;
#else
  // a lane owns Vec sub and 4 + sub: each of the two 128-bit loads covers 64 contiguous bytes of every row
  static __device__ __forceinline__ Row8 load(const Vec* p) {
    MSDA_CHK(p, 16, 0);
    MSDA_CHK(p + 4, 16, 0);
    Row8 r;
    r.lo = __ldg(p);
    r.hi = __ldg(p + 4);
    return r;
  }
  static constexpr int kVecPerLane = 1, kSecond = 4;
#endif
};
template <>
struct Wide<__nv_bfloat16> {
  using Vec = uint2;                   // 4 bf16; row = 8 of these
  static __device__ __forceinline__ Row8 load(const Vec* p) {
    MSDA_CHK(p, 16, 0);
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
    Row8 v;
    v.lo.x = __uint_as_float(r.x << 16); v.lo.y = __uint_as_float(r.x & 0xffff0000u);
    v.lo.z = __uint_as_float(r.y << 16); v.lo.w = __uint_as_float(r.y & 0xffff0000u);
    v.hi.x = __uint_as_float(r.z << 16); v.hi.y = __uint_as_float(r.z & 0xffff0000u);
    v.hi.z = __uint_as_float(r.w << 16); v.hi.w = __uint_as_float(r.w & 0xffff0000u);
    return v;
  }
  static constexpr int kVecPerLane = 2, kSecond = 1;
};

__device__ __forceinline__ void fma8(float4& lo, float4& hi, float w, const Row8& v) {
  fma4(lo, w, v.lo);
  fma4(hi, w, v.hi);
}

template <typename VT, int kM, bool kFuse = false>   // kM: compile-time head count (0: runtime)
__global__ void __launch_bounds__(kD32Threads, MSDA_FWD_WIDE_MINB)
fwd_d32_wide_kernel(const D32Params p) {
  using Wd = Wide<VT>;
  using Vec = typename Wd::Vec;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ LevelTab tab;
  load_level_tab(tab, p);
  constexpr int kL = 4, kP = 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = lane >> 2, sub = lane & 3;
  int4* warp_recs = reinterpret_cast<int4*>(smem_raw) + warp * (kWideItems * kWideStride);
  const int4* my_recs = warp_recs + group * kWideStride;
  const int M = kM > 0 ? kM : p.M;
  int wm[kL];
#pragma unroll
  for (int l = 0; l < kL; ++l) wm[l] = tab.W[l] * M;

  const long long n_pass = (p.n_items + kWideItems - 1) / kWideItems;
  const long long pass_stride = (long long)gridDim.x * kD32Warps;
  // raw locations / weights of a pass: 8 items x 16 samples over 32 lanes (4 each), coalesced.  With MSDA_FWD_WIDE_PREFETCH
  // the next pass's are requested before this pass starts gathering (their HBM latency then hides behind the gathers).
  float2 xy[4];
  float a[4];
  auto load_raw = [&](long long ps) {
    const long long s_base = ps * kWideItems * kWideLP, s_end = p.n_items * kWideLP;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long gs = s_base + lane + i * kWarp;
      xy[i] = make_float2(0.f, 0.f);
      a[i] = -INFINITY;
      if (gs < s_end) {
        xy[i] = ld_stream(reinterpret_cast<const float2*>(p.loc) + gs);
        a[i] = ld_stream(p.aw + gs);
      }
    }
  };
  const long long pass0 = (long long)blockIdx.x * kD32Warps + warp;
  if (MSDA_FWD_WIDE_PREFETCH && pass0 < n_pass) load_raw(pass0);
  for (long long pass = pass0; pass < n_pass; pass += pass_stride) {
    const long long item0 = pass * kWideItems;
    // ---- phase A ------------------------------------------------------------------------------------------------
    {
      const long long s_base = item0 * kWideLP, s_end = p.n_items * kWideLP;
      if (!MSDA_FWD_WIDE_PREFETCH) load_raw(pass);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int s = lane + i * kWarp;
        const bool in = s_base + s < s_end;
        const int grp = s >> 4, ls = s & 15, l = ls >> 2;
        float2 loc = xy[i];
        float w = a[i];
        if (kFuse) {
          w = segment_softmax<kWideLP>(w);
          if (in) loc = fused_location(p, item0 + grp, l, loc, tab.H[l], tab.W[l]);
        }
        if (in)
          warp_recs[grp * kWideStride + ls] = make_record<false>(loc.x, loc.y, w, tab.H[l], tab.W[l], tab.start[l], M);
      }
      if (MSDA_FWD_WIDE_PREFETCH && pass + pass_stride < n_pass) load_raw(pass + pass_stride);
    }
    __syncwarp();

    // ---- phase B: 4 lanes per item, 8 channels per lane -------------------------------------------------------
    const long long item = item0 + group;
    if (item < p.n_items) {
      const int m = (int)(item % M);
      const int b = (int)(item / ((long long)M * p.Lq));
      const Vec* vb = reinterpret_cast<const Vec*>(p.value) + ((long long)b * p.S * M + m) * 8 + sub * Wd::kVecPerLane;
      float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
#pragma unroll
      for (int l = 0; l < kL; ++l) {
#pragma unroll
        for (int pt = 0; pt < kP; ++pt) {
          const int4 r = my_recs[l * kP + pt];
          const int row = r.x >> 4;
          const float ah = __int_as_float(r.y), al = __int_as_float(r.z), lw = __int_as_float(r.w);
          const float w2 = ah * lw, w4 = al * lw;
          const float w1 = ah - w2, w3 = al - w4;
          const Vec* r0 = vb + (long long)row * 8;
          const Vec* r1 = r0 + (long long)wm[l] * 8;
          if (r.x & 1) fma8(lo, hi, w1, Wd::load(r0));
          if (r.x & 2) fma8(lo, hi, w2, Wd::load(r0 + M * 8));
          if (r.x & 4) fma8(lo, hi, w3, Wd::load(r1));
          if (r.x & 8) fma8(lo, hi, w4, Wd::load(r1 + M * 8));
        }
      }
      // the output row is fp32: a bf16 lane holds channels 8 sub .. 8 sub + 7, an fp32 lane the chunks it loaded
      constexpr bool kF32 = sizeof(VT) == 4;
      float4* o = reinterpret_cast<float4*>(p.out) + item * 8 + sub * (kF32 ? Wd::kVecPerLane : 2);
      st_stream(o, lo);
      st_stream(o + (kF32 ? Wd::kSecond : 1), hi);
    }
    __syncwarp();  // records are rewritten by the next pass
  }
}

inline size_t wide_smem_bytes() { return (size_t)kD32Warps * kWideItems * kWideStride * 16; }

}  // namespace msda
