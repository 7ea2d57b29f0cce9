// Microbenchmark: 128 B row gathers through L1 -- cost per row for different lane mappings.
//   A: LDG.32, 32 lanes per row  (1 row  = 1 line per warp instruction)
//   B: LDG.64, 16 lanes per row  (2 rows = 2 lines per warp instruction)
//   C: LDG.128, 8 lanes per row  (4 rows = 4 lines per warp instruction)
//   D: LDG.128, 8 lanes per row, all 4 groups on the SAME row (1 line per instruction)
// Each warp gathers `rows_per_warp` pseudo-random rows from a window of `span_rows` rows starting at
// a per-CTA offset (window small => L1 resident; window = whole buffer => L2 resident).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_gather tools/microbench_gather.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                           \
  do {                                                                                  \
    cudaError_t e = (x);                                                                \
    if (e != cudaSuccess) {                                                             \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);    \
      exit(1);                                                                          \
    }                                                                                   \
  } while (0)

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

template <int MODE, int UNROLL>
__global__ void __launch_bounds__(256) gather_kernel(const float* __restrict__ src, float* __restrict__ out,
                                                     unsigned total_rows, unsigned span_rows, int iters) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  // per-CTA window start (so that co-resident CTAs do not share lines by accident)
  const unsigned win0 = (hash32(blockIdx.x * 2654435761u) % (total_rows - span_rows + 1));
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  const unsigned mask = span_rows - 1;  // span_rows is a power of two
  // one cheap LCG stream per row slot: mode A per warp, B per half warp, C per 8-lane group, D per warp
  unsigned st = hash32(warp * 4u + (MODE == 0 || MODE == 3 ? 0u : (MODE == 1 ? (lane >> 4) : (lane >> 3))) + 17u);
  for (int i = 0; i < iters; i += UNROLL) {
    if (MODE == 0) {
      float v[UNROLL * 4];
#pragma unroll
      for (int u = 0; u < UNROLL * 4; ++u) {
        st = st * 1664525u + 1013904223u;
        const unsigned row = win0 + ((st >> 9) & mask);
        v[u] = __ldg(src + (size_t)row * 32 + lane);
      }
#pragma unroll
      for (int u = 0; u < UNROLL * 4; ++u) acc0 += v[u];
    } else if (MODE == 1) {
      float2 v[UNROLL * 2];
#pragma unroll
      for (int u = 0; u < UNROLL * 2; ++u) {
        st = st * 1664525u + 1013904223u;
        const unsigned row = win0 + ((st >> 9) & mask);
        v[u] = __ldg(reinterpret_cast<const float2*>(src + (size_t)row * 32) + (lane & 15));
      }
#pragma unroll
      for (int u = 0; u < UNROLL * 2; ++u) { acc0 += v[u].x; acc1 += v[u].y; }
    } else {
      float4 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        st = st * 1664525u + 1013904223u;
        const unsigned row = win0 + ((st >> 9) & mask);
        v[u] = __ldg(reinterpret_cast<const float4*>(src + (size_t)row * 32) + (lane & 7));
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) { acc0 += v[u].x; acc1 += v[u].y; acc2 += v[u].z; acc3 += v[u].w; }
    }
  }
  const float r = acc0 + acc1 + acc2 + acc3;
  if (r == 123.456f) out[threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, const float* src, float* out, unsigned total_rows, unsigned span_rows, int ctas_per_sm) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int ctas = 148 * ctas_per_sm;
  const int iters = 2048;  // warp iterations of 4 rows
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0));
    gather_kernel<MODE, 4><<<ctas, 256>>>(src, out, total_rows, span_rows, iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double rows = (double)ctas * 8 * iters * 4;
  // cycles per row per SM at 1.9 GHz nominal (report rows/us too)
  printf("  %-44s %8.1f us  %7.2f TB/s  %6.3f ns/row/SM  (~%.2f clk/row/SM @1.9GHz)\n", name, best * 1e3,
         rows * 128 / (best * 1e-3) / 1e12, best * 1e6 / (rows / 148), best * 1e6 / (rows / 148) * 1.9);
}

int main(int argc, char** argv) {
  const unsigned total_rows = 524288;  // 64 MB
  float *src, *out;
  CK(cudaMalloc(&src, (size_t)total_rows * 128));
  CK(cudaMemset(src, 0, (size_t)total_rows * 128));
  CK(cudaMalloc(&out, 4096));
  for (int pass = 0; pass < 3; ++pass) {
    const unsigned span = pass == 0 ? 128u : (pass == 1 ? 1024u : 262144u);
    for (int occ = 4; occ <= 8; occ += 4) {
      printf("window %u rows (%.0f KB) per CTA, %d CTAs/SM (%s):\n", span, span * 128.0 / 1024, occ,
             pass == 0 ? "L1 resident" : (pass == 1 ? "128 KB/CTA: L1 thrash at 8 CTAs" : "32 MB window: L2 resident"));
      run<0>("A  LDG.32  x1 line  (32 lanes/row)", src, out, total_rows, span, occ);
      run<1>("B  LDG.64  x2 lines (16 lanes/row)", src, out, total_rows, span, occ);
      run<2>("C  LDG.128 x4 lines ( 8 lanes/row)", src, out, total_rows, span, occ);
      run<3>("D  LDG.128 x1 line  (4 groups same row)", src, out, total_rows, span, occ);
    }
  }
  CK(cudaDeviceSynchronize());
  printf("done\n");
  return 0;
}
