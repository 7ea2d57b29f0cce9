// Shared device-side pieces of the MSDeformAttn kernels (sm_100a).
//
// The ONE routine that decides integer work -- pixel floor, corner validity, the coarse skip --
// is sample_geom() below.  Forward, backward and the index probe all inline it, so they cannot
// disagree with each other, and tests compare its decisions bit-for-bit with the CPU oracle.
//
// Semantics follow the reference kernels as COMPILED:
//   /root/reference/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:285-288  (coordinate + skip)
//   /root/reference/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:38-78    (floor + corner masks)
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace msda {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

template <typename T>
struct Num;
template <>
struct Num<float> {
  // `loc * size - 0.5` is emitted as one FFMA by nvcc for the reference; make that explicit so
  // the floor cannot move with compiler flags.
  static __device__ __forceinline__ float mad(float a, float b, float c) { return __fmaf_rn(a, b, c); }
  static __device__ __forceinline__ float flr(float x) { return floorf(x); }
};
template <>
struct Num<double> {
  static __device__ __forceinline__ double mad(double a, double b, double c) { return __fma_rn(a, b, c); }
  static __device__ __forceinline__ double flr(double x) { return floor(x); }
};

template <typename T>
struct SampleGeom {
  int h_low, w_low;  // floor of the pixel coordinate; each may be -1 at the top/left border
  unsigned mask;     // bit i <=> corner v(i+1) lies inside the level
  T lh, lw;          // fractional parts (exact)
};

// Returns false when the sample contributes nothing: outside (-1, size) on either axis, NaN or
// +-inf (every comparison is false for those), exactly as cuh:288.
template <typename T>
__device__ __forceinline__ bool sample_geom(T loc_x, T loc_y, int H, int W, SampleGeom<T>& g) {
  const T Hf = (T)H, Wf = (T)W;
  const T h_im = Num<T>::mad(loc_y, Hf, (T)-0.5);
  const T w_im = Num<T>::mad(loc_x, Wf, (T)-0.5);
  if (!(h_im > (T)-1 && w_im > (T)-1 && h_im < Hf && w_im < Wf)) return false;
  const T hf = Num<T>::flr(h_im), wf = Num<T>::flr(w_im);
  g.h_low = (int)hf;
  g.w_low = (int)wf;
  g.lh = h_im - hf;
  g.lw = w_im - wf;
  const bool top = g.h_low >= 0, left = g.w_low >= 0;
  const bool bottom = g.h_low + 1 <= H - 1, right = g.w_low + 1 <= W - 1;
  g.mask = (top && left ? 1u : 0u) | (top && right ? 2u : 0u) | (bottom && left ? 4u : 0u) |
           (bottom && right ? 8u : 0u);
  return true;
}

// ---- optional bounds checking (-DMSDA_BOUNDS_CHECK=1: the "checked" library the tests also run) ----
// compute-sanitizer is not available on the GPU pool this was developed on, so the kernels carry
// their own memcheck for the accesses whose addresses depend on DATA (sampling locations): every
// gathered value row and every grad_value reduction is compared with the extent of the tensor the
// host registered for this launch; a violation prints once and traps (the launch fails).
#ifndef MSDA_BOUNDS_CHECK
#define MSDA_BOUNDS_CHECK 0
#endif
#if MSDA_BOUNDS_CHECK
__device__ const char* g_chk_lo[2];   // [0] value, [1] grad_value (or its fp32 scratch)
__device__ const char* g_chk_hi[2];
__device__ __forceinline__ void chk_access(const void* p, unsigned bytes, int which) {
  const char* c = reinterpret_cast<const char*>(p);
  if (c < g_chk_lo[which] || c + bytes > g_chk_hi[which]) {
    printf("msda bounds check: %s access of %u B at %p outside [%p, %p)\n", which ? "grad_value" : "value", bytes, p,
           (const void*)g_chk_lo[which], (const void*)g_chk_hi[which]);
    __trap();
  }
}
#define MSDA_CHK(p, bytes, which) chk_access((p), (bytes), (which))
#else
#define MSDA_CHK(p, bytes, which) ((void)0)
#endif

// ---- memory helpers ---------------------------------------------------------------------

// Streaming (read-once) inputs: do not let them evict gathered value rows from L1.
__device__ __forceinline__ float ld_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float2 ld_stream(const float2* p) {
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream(float4* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_stream(float2* p, float2 v) {
  asm volatile("st.global.cs.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void st_stream(float* p, float v) {
  asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// One value row chunk = 4 consecutive channels of one (token, head).
// fp32 storage: 16 B, bf16 storage: 8 B.  Rows are gathered through the read-only path and DO
// allocate in L1 (neighbouring queries hit the same rows).
template <typename VT>
struct Chunk;
template <>
struct Chunk<float> {
  using Vec = float4;
  static __device__ __forceinline__ float4 load(const Vec* p) {
    MSDA_CHK(p, 16, 0);
#ifdef MSDA_EXPERIMENT_NO_GATHER   // timing experiment: no value gathers (rows made up from the address)
    const float f = __uint_as_float(0x3f800000u | ((unsigned)(size_t)p & 0xffffu));
    return make_float4(f, f, f, f);
#endif
#ifdef MSDA_EXPERIMENT_VALUE_NO_ALLOCATE   // timing experiment: gathered rows do not allocate in L1 either
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
#endif
    return __ldg(p);
  }
  // red.global.add.v4.f32: one 16 B reduction instead of four scalar atomics (sm_90+).
  static __device__ __forceinline__ void red_add(Vec* p, float4 v) {
    MSDA_CHK(p, 16, 1);
#ifdef MSDA_EXPERIMENT_NO_RED   // timing experiment (tools/build_variants.sh): the backward without its reductions
    if (v.x == 1.2345e-30f) *reinterpret_cast<volatile float*>(p) = v.y;
    return;
#endif
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
  }
};
template <>
struct Chunk<__nv_bfloat16> {
  using Vec = uint2;
  static __device__ __forceinline__ float4 load(const Vec* p) {
    MSDA_CHK(p, 8, 0);
    const uint2 r = __ldg(p);
    float4 v;
    v.x = __uint_as_float(r.x << 16);
    v.y = __uint_as_float(r.x & 0xffff0000u);
    v.z = __uint_as_float(r.y << 16);
    v.w = __uint_as_float(r.y & 0xffff0000u);
    return v;
  }
  static __device__ __forceinline__ void red_add(Vec* p, float4 v) {
    MSDA_CHK(p, 8, 1);
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
    const __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
    const unsigned a = *reinterpret_cast<const unsigned*>(&lo);
    const unsigned b = *reinterpret_cast<const unsigned*>(&hi);
    asm volatile("red.global.add.noftz.v2.bf16x2 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
  }
};

}  // namespace msda
