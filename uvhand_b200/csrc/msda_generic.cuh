// Any-shape kernels: every D, L, P; float and double; 64-bit addressing throughout.
//
// These exist so that the whole input space of the reference operator is served (its test
// sweeps D in {30,32,64,71,1025,2048,3096} in fp64, models/ops/test.py:85-86).  They are not
// the performance path -- the D=32 kernels in msda_d32.cuh are -- but they are still written
// for the machine: one warp per (sample, query, head) with lanes over channels so every value
// row access is a coalesced segment, location/weight scalars loaded once per warp and
// broadcast by the hardware, and the backward's grad_loc / grad_weight sums reduced with warp
// shuffles (the reference reduces them through shared memory with block barriers and a serial
// loop on thread 0, cuh:376-394).
#pragma once

#include "msda_device.cuh"

namespace msda {

struct Dims {
  int N, S, M, D, L, Lq, P;
};

constexpr int kGenericThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kGenericThreads)
fwd_generic_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes,
                   const int64_t* __restrict__ lsi, const T* __restrict__ loc,
                   const T* __restrict__ aw, Dims d, T* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long n_items = (long long)d.N * d.Lq * d.M;
  const long long n_warps = (long long)gridDim.x * (kGenericThreads / kWarp);
  const int LP = d.L * d.P;
  const long long row_stride = (long long)d.M * d.D;
  for (long long item = (long long)blockIdx.x * (kGenericThreads / kWarp) + (threadIdx.x >> 5);
       item < n_items; item += n_warps) {
    const int m = (int)(item % d.M);
    const int b = (int)(item / ((long long)d.M * d.Lq));
    const T* vb = value + (long long)b * d.S * row_stride + (long long)m * d.D;
    const T* locp = loc + item * LP * 2;
    const T* awp = aw + item * LP;
    // Channel chunks outermost: a lane owns channel c0+lane; geometry is recomputed per chunk
    // (one chunk when D <= 32).
    for (int c0 = 0; c0 < d.D; c0 += kWarp) {
      const int c = c0 + lane;
      const bool live = c < d.D;
      T col = 0;
      for (int l = 0; l < d.L; ++l) {
        const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
        const T* vl = vb + (long long)lsi[l] * row_stride;
        for (int p = 0; p < d.P; ++p) {
          const int s = l * d.P + p;
          const T x = locp[2 * s], y = locp[2 * s + 1], a = awp[s];
          SampleGeom<T> g;
          if (!sample_geom<T>(x, y, H, W, g)) continue;
          if (!live) continue;
          const T hh = 1 - g.lh, hw = 1 - g.lw;
          const long long r00 = ((long long)g.h_low * W + g.w_low) * row_stride + c;
          if (g.mask & 1u) MSDA_CHK(vl + r00, sizeof(T), 0);
          if (g.mask & 2u) MSDA_CHK(vl + r00 + row_stride, sizeof(T), 0);
          if (g.mask & 4u) MSDA_CHK(vl + r00 + (long long)W * row_stride, sizeof(T), 0);
          if (g.mask & 8u) MSDA_CHK(vl + r00 + (long long)(W + 1) * row_stride, sizeof(T), 0);
          const T v1 = (g.mask & 1u) ? vl[r00] : (T)0;
          const T v2 = (g.mask & 2u) ? vl[r00 + row_stride] : (T)0;
          const T v3 = (g.mask & 4u) ? vl[r00 + (long long)W * row_stride] : (T)0;
          const T v4 = (g.mask & 8u) ? vl[r00 + (long long)(W + 1) * row_stride] : (T)0;
          const T w1 = hh * hw, w2 = hh * g.lw, w3 = g.lh * hw, w4 = g.lh * g.lw;
          col += (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4) * a;
        }
      }
      if (live) out[item * d.D + c] = col;
    }
  }
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

template <typename T>
__global__ void __launch_bounds__(kGenericThreads)
bwd_generic_kernel(const T* __restrict__ grad_out, const T* __restrict__ value,
                   const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi,
                   const T* __restrict__ loc, const T* __restrict__ aw, Dims d,
                   T* __restrict__ grad_value, T* __restrict__ grad_loc, T* __restrict__ grad_aw) {
  const int lane = threadIdx.x & 31;
  const long long n_items = (long long)d.N * d.Lq * d.M;
  const long long n_warps = (long long)gridDim.x * (kGenericThreads / kWarp);
  const int LP = d.L * d.P;
  const long long row_stride = (long long)d.M * d.D;
  for (long long item = (long long)blockIdx.x * (kGenericThreads / kWarp) + (threadIdx.x >> 5);
       item < n_items; item += n_warps) {
    const int m = (int)(item % d.M);
    const int b = (int)(item / ((long long)d.M * d.Lq));
    const long long head_base = (long long)b * d.S * row_stride + (long long)m * d.D;
    const T* g_row = grad_out + item * d.D;
    const T* locp = loc + item * LP * 2;
    const T* awp = aw + item * LP;
    for (int l = 0; l < d.L; ++l) {
      const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
      const long long lvl_base = head_base + (long long)lsi[l] * row_stride;
      for (int p = 0; p < d.P; ++p) {
        const int s = l * d.P + p;
        const T x = locp[2 * s], y = locp[2 * s + 1], a = awp[s];
        SampleGeom<T> g;
        T gx = 0, gy = 0, ga = 0;
        if (sample_geom<T>(x, y, H, W, g)) {  // warp-uniform branch
          const T hh = 1 - g.lh, hw = 1 - g.lw;
          const T w1 = hh * hw, w2 = hh * g.lw, w3 = g.lh * hw, w4 = g.lh * g.lw;
          const long long r00 = lvl_base + ((long long)g.h_low * W + g.w_low) * row_stride;
          const long long r01 = r00 + row_stride;
          const long long r10 = r00 + (long long)W * row_stride;
          const long long r11 = r10 + row_stride;
          for (int c = lane; c < d.D; c += kWarp) {
            const T top = g_row[c];
            const T t = top * a;
            T v1 = 0, v2 = 0, v3 = 0, v4 = 0;
            if (g.mask & 1u) { MSDA_CHK(value + r00 + c, sizeof(T), 0); MSDA_CHK(grad_value + r00 + c, sizeof(T), 1); v1 = value[r00 + c]; atomicAdd(grad_value + r00 + c, w1 * t); }
            if (g.mask & 2u) { MSDA_CHK(value + r01 + c, sizeof(T), 0); MSDA_CHK(grad_value + r01 + c, sizeof(T), 1); v2 = value[r01 + c]; atomicAdd(grad_value + r01 + c, w2 * t); }
            if (g.mask & 4u) { MSDA_CHK(value + r10 + c, sizeof(T), 0); MSDA_CHK(grad_value + r10 + c, sizeof(T), 1); v3 = value[r10 + c]; atomicAdd(grad_value + r10 + c, w3 * t); }
            if (g.mask & 8u) { MSDA_CHK(value + r11 + c, sizeof(T), 0); MSDA_CHK(grad_value + r11 + c, sizeof(T), 1); v4 = value[r11 + c]; atomicAdd(grad_value + r11 + c, w4 * t); }
            const T gh = hw * (v3 - v1) + g.lw * (v4 - v2);
            const T gw = hh * (v2 - v1) + g.lh * (v4 - v3);
            ga += top * (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);
            gx += (T)W * gw * t;
            gy += (T)H * gh * t;
          }
          gx = warp_sum(gx);
          gy = warp_sum(gy);
          ga = warp_sum(ga);
        }
        if (lane == 0) {
          grad_loc[(item * LP + s) * 2] = gx;
          grad_loc[(item * LP + s) * 2 + 1] = gy;
          grad_aw[item * LP + s] = ga;
        }
      }
    }
  }
}

// Index probe: the decisions of sample_geom for every sample, one thread per sample.
__global__ void __launch_bounds__(256)
sample_index_kernel(const int64_t* __restrict__ shapes, const float* __restrict__ loc, long long n_samples,
                    int L, int P, int32_t* __restrict__ h_low, int32_t* __restrict__ w_low,
                    uint8_t* __restrict__ mask) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_samples;
       i += (long long)gridDim.x * blockDim.x) {
    const int l = (int)((i / P) % L);
    const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
    SampleGeom<float> g;
    if (sample_geom<float>(loc[2 * i], loc[2 * i + 1], H, W, g)) {
      h_low[i] = g.h_low;
      w_low[i] = g.w_low;
      mask[i] = (uint8_t)(g.mask | 16u);
    } else {
      h_low[i] = 0;
      w_low[i] = 0;
      mask[i] = 0;
    }
  }
}

}  // namespace msda
