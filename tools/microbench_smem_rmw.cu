// Microbenchmark: what does it cost to accumulate 128 B gradient rows in a warp-private shared-memory
// window (LDS.128 + FADD + STS.128, no atomics) instead of pushing every row to L2 with red.global?
// One persistent CTA of 8 warps per SM; every warp owns a 24 KB window (192 rows of 32 floats).
// A "step" = one bilinear sample = 4 corner rows (r, r+1, r+ww, r+ww+1); 8 lanes x float4 per row.
//
//   mode 0  RMW only
//   mode 1  RMW + 4 row gathers (LDG.128) from a 45 MB buffer (L2 resident), like the value reads
//   mode 2  mode 1, but every 4th step sends its rows to L2 with red.global.add.v4.f32 instead of the window
//   mode 3  red.global only + gathers (the round-1 backward's inner loop)
//   mode 4  LDS.128 broadcast only (all lanes one address): cost of handing a 16 B record to the warp
//   mode 5  SHFL only (4 per step)
//   mode 6  mode 1 + 1 LDS.128 broadcast + 4 SHFL per step (everything a window-mode sample needs)
//   mode 7  mode 1 with 2 samples interleaved per step when they do not alias (software ILP)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_smem_rmw tools/microbench_smem_rmw.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                 \
  do {                                                                        \
    cudaError_t e = (x);                                                      \
    if (e != cudaSuccess) {                                                   \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                \
    }                                                                         \
  } while (0)

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

constexpr int kWarps = 8;
constexpr int kWinRows = 192;
constexpr int kWW = 14;

__device__ __forceinline__ void red4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(kWarps * 32, 1)
rmw_kernel(const float4* __restrict__ gsrc, float* gdst, unsigned span_rows, int iters, unsigned seed, float* sink) {
  extern __shared__ __align__(16) float4 smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 3, sub = lane & 7;
  float4* win = smem + warp * (kWinRows * 8);
  for (int i = lane; i < kWinRows * 8; i += 32) win[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();
  const unsigned wid = blockIdx.x * kWarps + warp;
  const int corner_off = (grp & 1) + (grp >> 1) * kWW;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float sh = (float)lane;
  for (int i = 0; i < iters; ++i) {
    const unsigned h = hash32(seed + wid * iters + i);
    const int r = (int)(h % (unsigned)(kWinRows - kWW - 2)) + corner_off;
    const unsigned grow = (hash32(h) % (span_rows - 4096)) + (grp & 1) + (grp >> 1) * 1336;  // 4 scattered rows
    if (MODE == 4) {
      const float4 rec = smem[(h & 255)];
      acc.x += rec.x; acc.y += rec.y; acc.z += rec.z; acc.w += rec.w;
      continue;
    }
    if (MODE == 5) {
#pragma unroll
      for (int k = 0; k < 4; ++k) sh += __shfl_xor_sync(0xffffffffu, sh, 1 << k);
      continue;
    }
    float4 v = make_float4(1.f, 1.f, 1.f, 1.f);
    if (MODE == 1 || MODE == 2 || MODE == 3 || MODE == 6 || MODE == 7) v = __ldg(gsrc + (size_t)grow * 8 + sub);
    if (MODE == 6) {
      const float4 rec = smem[(h & 255)];
      v.x += rec.x;
      float d = v.x * v.y;
#pragma unroll
      for (int k = 0; k < 4; ++k) d += __shfl_xor_sync(0xffffffffu, d, 1 << k);
      acc.w += d;
    }
    const bool to_l2 = MODE == 3 || (MODE == 2 && (i & 3) == 0);
    if (to_l2) {
      red4(gdst + (size_t)grow * 32 + sub * 4, v);
    } else if (MODE == 7) {
      // second, independent sample in the same step (different half of the window => never aliases)
      const int r2 = (int)(hash32(h + 77u) % (unsigned)(kWinRows / 2 - kWW - 2)) + kWinRows / 2 + corner_off;
      const int r1 = (int)(h % (unsigned)(kWinRows / 2 - kWW - 2)) + corner_off;
      float4 a = win[r1 * 8 + sub], b = win[r2 * 8 + sub];
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      b.x += v.y; b.y += v.x; b.z += v.w; b.w += v.z;
      win[r1 * 8 + sub] = a;
      win[r2 * 8 + sub] = b;
    } else {
      float4 a = win[r * 8 + sub];
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      win[r * 8 + sub] = a;
    }
  }
  __syncwarp();
  for (int i = lane; i < kWinRows * 8; i += 32) {
    const float4 a = win[i];
    acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
  }
  if (acc.x + acc.y + acc.z + acc.w + sh == -12345.f) sink[0] = acc.x;
}

template <int MODE>
void run(const char* name, const float4* src, float* dst, unsigned span_rows, int iters, float* sink, int sms, double rows_per_step) {
  const size_t smem = (size_t)kWarps * kWinRows * 128;
  CK(cudaFuncSetAttribute(rmw_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0));
    rmw_kernel<MODE><<<sms, kWarps * 32, smem>>>(src, dst, span_rows, iters, 1234u + rep, sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  const double steps_per_sm = (double)kWarps * iters;
  const double clk = best * 1e-3 * 1.965e9;
  printf("%-58s %8.1f us  %6.2f clk/step/SM  %6.2f clk/row/SM\n", name, best * 1e3, clk / steps_per_sm,
         clk / steps_per_sm / rows_per_step);
}

int main() {
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const unsigned span_rows = 45u * 1024 * 1024 / 128;
  float4* src;
  float *dst, *sink;
  CK(cudaMalloc(&src, (size_t)span_rows * 128));
  CK(cudaMalloc(&dst, (size_t)span_rows * 128));
  CK(cudaMalloc(&sink, 16));
  CK(cudaMemset(src, 0, (size_t)span_rows * 128));
  CK(cudaMemset(dst, 0, (size_t)span_rows * 128));
  const int iters = 20000;
  printf("SMs %d, %d warps/SM, window %d rows/warp, %d steps/warp (1 step = 4 corner rows)\n", sms, kWarps, kWinRows, iters);
  run<0>("0 RMW (LDS.128+FADD+STS.128) only", src, dst, span_rows, iters, sink, sms, 4);
  run<1>("1 RMW + 4 LDG.128 row gathers (L2)", src, dst, span_rows, iters, sink, sms, 4);
  run<2>("2 as 1, every 4th step red.global.v4 instead of RMW", src, dst, span_rows, iters, sink, sms, 4);
  run<3>("3 red.global.v4 + gathers (round-1 inner loop)", src, dst, span_rows, iters, sink, sms, 4);
  run<4>("4 LDS.128 broadcast only (per instr)", src, dst, span_rows, iters, sink, sms, 1);
  run<5>("5 SHFL x4 only (per 4 instr)", src, dst, span_rows, iters, sink, sms, 4);
  run<6>("6 as 1 + record LDS.128 + 4 SHFL", src, dst, span_rows, iters, sink, sms, 4);
  run<7>("7 as 1 but 2 independent samples per step (8 rows)", src, dst, span_rows, iters, sink, sms, 8);
  return 0;
}
