// Microbenchmark: accumulate 128 B gradient rows into a CTA-SHARED shared-memory window with atomics,
// so that any warp may hit any row (what an on-chip pre-reduction of grad_value needs).
// Layout as in the kernels: 8 lanes per row, lane owns 4 consecutive channels; a warp instruction
// touches 4 rows (the 4 corners of one bilinear sample).
//
//   mode 0  atomicAdd(float) x4 per lane          (sm_100a: ATOMS.CAST.SPIN CAS loop)
//   mode 1  atomicAdd(unsigned) x4 per lane       (native RED/ATOMS.ADD; fixed-point accumulation)
//   mode 2  atomicAdd(unsigned long long) x4      (native 64-bit; window is twice as large)
//   mode 3  non-atomic LDS.128 + FADD + STS.128   (racy; lower bound for a read-modify-write)
//   +4      the same plus 4 LDG.128 row gathers from a 45 MB L2-resident buffer per step
// j-rotation: in step j group g touches channel (j + g) & 3 of its chunk, so that the 4 groups of an
// instruction fall into distinct banks.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/microbench_smem_atomic tools/microbench_smem_atomic.cu
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

constexpr int kWinRows = 576;
constexpr int kWW = 24;

template <int MODE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
atom_kernel(const float4* __restrict__ gsrc, unsigned span_rows, int iters, unsigned seed, float* sink) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int M = MODE & 3;
  constexpr bool kGather = MODE >= 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 3, sub = lane & 7;
  const int words = kWinRows * 32 * (M == 2 ? 2 : 1);
  for (int i = threadIdx.x; i < words; i += WARPS * 32) reinterpret_cast<unsigned*>(smem_raw)[i] = 0;
  __syncthreads();
  const unsigned wid = blockIdx.x * WARPS + warp;
  const int corner_off = (grp & 1) + (grp >> 1) * kWW;
  float acc = 0.f;
  for (int i = 0; i < iters; ++i) {
    const unsigned h = hash32(seed + wid * iters + i);
    const int r = (int)(h % (unsigned)(kWinRows - kWW - 2)) + corner_off;
    float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    if (kGather) {
      const unsigned grow = (hash32(h) % (span_rows - 4096)) + (grp & 1) + (grp >> 1) * 1336;
      v = __ldg(gsrc + (size_t)grow * 8 + sub);
      v.x += 1.f;
    }
    const float vv[4] = {v.x, v.y, v.z, v.w};
    if (M == 0) {
      float* w = reinterpret_cast<float*>(smem_raw) + r * 32 + sub * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(w + ((j + grp) & 3), vv[(j + grp) & 3]);
    } else if (M == 1) {
      unsigned* w = reinterpret_cast<unsigned*>(smem_raw) + r * 32 + sub * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(w + ((j + grp) & 3), (unsigned)__float2int_rn(vv[(j + grp) & 3] * 1024.f));
    } else if (M == 2) {
      unsigned long long* w = reinterpret_cast<unsigned long long*>(smem_raw) + r * 32 + sub * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        atomicAdd(w + ((j + grp) & 3), (unsigned long long)__float2ll_rn(vv[(j + grp) & 3] * 1048576.f));
    } else {
      float4* w = reinterpret_cast<float4*>(smem_raw) + r * 8 + sub;
      float4 a = *w;
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      *w = a;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < words; i += WARPS * 32) acc += (float)reinterpret_cast<unsigned*>(smem_raw)[i];
  if (acc == -12345.f) sink[0] = acc;
}

template <int MODE, int WARPS>
void run(const char* name, const float4* src, unsigned span_rows, int iters, float* sink, int sms, int ctas_per_sm) {
  const size_t smem = (size_t)kWinRows * 128 * ((MODE & 3) == 2 ? 2 : 1);
  CK(cudaFuncSetAttribute(atom_kernel<MODE, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, atom_kernel<MODE, WARPS>, WARPS * 32, smem));
  if (occ < ctas_per_sm) ctas_per_sm = occ;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0));
    atom_kernel<MODE, WARPS><<<sms * ctas_per_sm, WARPS * 32, smem>>>(src, span_rows, iters, 1234u + rep, sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  const double steps_per_sm = (double)WARPS * ctas_per_sm * iters;
  const double clk = best * 1e-3 * 1.965e9;
  printf("%-64s %2d warps x %d CTA/SM %8.1f us  %6.2f clk/row/SM\n", name, WARPS, ctas_per_sm, best * 1e3,
         clk / steps_per_sm / 4.0);
}

int main() {
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const unsigned span_rows = 45u * 1024 * 1024 / 128;
  float4* src;
  float* sink;
  CK(cudaMalloc(&src, (size_t)span_rows * 128));
  CK(cudaMalloc(&sink, 16));
  CK(cudaMemset(src, 0, (size_t)span_rows * 128));
  const int iters = 4000;
  printf("SMs %d, CTA-shared window of %d rows, %d steps/warp (1 step = 4 corner rows)\n", sms, kWinRows, iters);
  run<0, 16>("0 atomicAdd(float) x4", src, span_rows, iters, sink, sms, 3);
  run<1, 16>("1 atomicAdd(u32) x4 (fixed point)", src, span_rows, iters, sink, sms, 3);
  run<2, 16>("2 atomicAdd(u64) x4 (fixed point)", src, span_rows, iters, sink, sms, 1);
  run<3, 16>("3 LDS.128+FADD+STS.128 (racy reference)", src, span_rows, iters, sink, sms, 3);
  run<4, 16>("4 = 0 + gathers", src, span_rows, iters, sink, sms, 3);
  run<5, 16>("5 = 1 + gathers", src, span_rows, iters, sink, sms, 3);
  run<6, 16>("6 = 2 + gathers", src, span_rows, iters, sink, sms, 1);
  run<7, 16>("7 = 3 + gathers", src, span_rows, iters, sink, sms, 3);
  run<5, 8>("5 = 1 + gathers", src, span_rows, iters, sink, sms, 3);
  run<1, 8>("1 atomicAdd(u32) x4", src, span_rows, iters, sink, sms, 3);
  run<6, 32>("6 = 2 + gathers", src, span_rows, iters, sink, sms, 1);
  run<2, 32>("2 atomicAdd(u64) x4", src, span_rows, iters, sink, sms, 1);
  return 0;
}
