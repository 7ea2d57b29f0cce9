// Microbenchmark: what one SM's load path charges for a warp-wide GATHER instruction as a function of
// how many distinct 128 B lines it touches.  Not part of the product; evidence for DESIGN.md 3.4.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l1_gather_bench l1_gather_bench.cu && ./l1_gather_bench
//
// Every warp issues kIters x 8 independent loads; the lanes of a load are split into groups of
// `lanes_per_line` lanes that read consecutive pieces of one line, and every group picks its line
// pseudo-randomly (a register LCG: no index loads) from a table of `lines` lines:
//   lines * 128 B <= ~100 KB -> L1 hits (global) ; 64 MB -> L2 hits ; shared-memory table -> LDS.
// Reported: SM cycles per load instruction and per line touched, at 32 resident warps per SM.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int kBytes> struct Ld;
template <> struct Ld<4>  { using T = float;  static __device__ float g(const void* p) { return __ldg((const float*)p); }
                            static __device__ float s(const void* p) { return *(const volatile float*)p; } };
template <> struct Ld<8>  { using T = float2; static __device__ float g(const void* p) { float2 v = __ldg((const float2*)p); return v.x + v.y; }
                            static __device__ float s(const void* p) { float2 v = *(const float2*)p; return v.x + v.y; } };
template <> struct Ld<16> { using T = float4; static __device__ float g(const void* p) { float4 v = __ldg((const float4*)p); return v.x + v.y + v.z + v.w; }
                            static __device__ float s(const void* p) { float4 v = *(const float4*)p; return v.x + v.y + v.z + v.w; } };

constexpr int kIters = 256;

template <int kBytes, bool kShared, int mode = 0>
__global__ void __launch_bounds__(256) gather_kernel(const char* table, unsigned line_mask, int lanes_per_line, float* sink,
                                                     long long* cycles, unsigned big_mask = 0) {
  extern __shared__ __align__(128) char smem[];
  const int lane = threadIdx.x & 31;
  const int group = lane / lanes_per_line, piece = lane % lanes_per_line;
  if (kShared) {
    for (int i = threadIdx.x; i < (int)(line_mask + 1) * 32; i += blockDim.x) ((float*)smem)[i] = (float)i;
    __syncthreads();
  }
  const char* base = kShared ? smem : table;
  unsigned state = (blockIdx.x * 8 + (threadIdx.x >> 5)) * 977u + group * 7919u + 12345u;
  float acc = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < kIters; ++it) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      state = state * 1664525u + 1013904223u;
      unsigned line = (state >> 8) & line_mask;
      size_t off = (size_t)line * 128;
      if (mode == 1 && ((state >> 28) & 3) == 0) off = (size_t)((state >> 4) & big_mask) * 128;            // this group misses
      if (mode == 2 && ((it * 8 + j) & 3) == 0) off = (size_t)((state >> 4) & big_mask) * 128;             // this instruction misses
      if (mode == 3) off = (size_t)line * (256 * 1024 + 128);                                              // one line per 256 KB
      if (mode == 4) off = (size_t)line * 1024;                                                            // one line per KB: one head's rows
      if (mode == 5) off = (size_t)line * 1024 + (size_t)group * 128;                                      // ... four heads side by side
      if (mode == 6) off = (size_t)(__shfl_sync(0xffffffffu, state, 0) >> 8 & line_mask & ~3u) * 128 + (size_t)group * 128;  // 4 ADJACENT lines (512 B) per instruction
      const char* p = base + off + piece * kBytes;
      v[j] = kShared ? Ld<kBytes>::s(p) : Ld<kBytes>::g(p);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += v[j];
  }
  const long long t1 = clock64();
  if (acc == 12345.678f) sink[0] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int kBytes, bool kShared, int mode = 0>
void run(const char* name, const char* table, unsigned lines, int lanes_per_line, float* sink, long long* d_cycles, int sms,
         unsigned big_mask = 0) {
  const int ctas = sms * 4;   // 4 x 256 threads = 32 warps per SM
  const size_t smem = kShared ? (size_t)lines * 128 : 0;
  auto k = gather_kernel<kBytes, kShared, mode>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; ++rep) k<<<ctas, 256, smem, 0>>>(table, lines - 1, lanes_per_line, sink, d_cycles, big_mask);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<<<ctas, 256, smem, 0>>>(table, lines - 1, lanes_per_line, sink, d_cycles, big_mask);
  cudaEventRecord(e1);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s: launch failed: %s\n", name, cudaGetErrorString(cudaGetLastError())); return; }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  long long* h = (long long*)malloc(sizeof(long long) * ctas);
  cudaMemcpy(h, d_cycles, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < ctas; ++i) mean += (double)h[i];
  mean /= ctas;
  free(h);
  // per SM: 32 warps x kIters x 8 loads, all resident at once
  const double loads_per_sm = 32.0 * kIters * 8;
  const int n_lines = 32 / lanes_per_line;
  printf("%-34s %2d lines/instr x %3d B   %7.2f cycles/instr  %5.2f cycles/line  %6.1f B/clk/SM   (%.1f us)\n", name, n_lines,
         lanes_per_line * kBytes, mean / loads_per_sm, mean / loads_per_sm / n_lines, loads_per_sm * 32 * kBytes / mean, ms * 1e3);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t big = 96ull << 20;
  char* table;
  float* sink;
  long long* cyc;
  cudaMalloc(&table, big);
  cudaMemset(table, 0, big);
  cudaMalloc(&sink, 4);
  cudaMalloc(&cyc, sizeof(long long) * sms * 4);
  const unsigned l1_lines = 256;          // 32 KB: L1 hits
  const unsigned l2_lines = (64u << 20) / 128;    // 64 MB: L2 hits (126 MB L2)
  printf("== global loads, table in L1 (32 KB)\n");
  run<4, false>("LDG.32  32 lanes/line", table, l1_lines, 32, sink, cyc, sms);
  run<8, false>("LDG.64  16 lanes/line", table, l1_lines, 16, sink, cyc, sms);
  run<16, false>("LDG.128  8 lanes/line (fp32 fwd)", table, l1_lines, 8, sink, cyc, sms);
  run<8, false>("LDG.64   8 lanes/line (bf16 fwd)", table, l1_lines, 8, sink, cyc, sms);
  run<16, false>("LDG.128  4 lanes/line (bf16 wide)", table, l1_lines, 4, sink, cyc, sms);
  run<8, false>("LDG.64   4 lanes/line", table, l1_lines, 4, sink, cyc, sms);
  run<16, false>("LDG.128  2 lanes/line", table, l1_lines, 2, sink, cyc, sms);
  run<16, false>("LDG.128  1 lane /line", table, l1_lines, 1, sink, cyc, sms);
  printf("== global loads, table in L2 (64 MB)\n");
  run<4, false>("LDG.32  32 lanes/line", table, l2_lines, 32, sink, cyc, sms);
  run<16, false>("LDG.128  8 lanes/line (fp32 fwd)", table, l2_lines, 8, sink, cyc, sms);
  run<8, false>("LDG.64   8 lanes/line (bf16 fwd)", table, l2_lines, 8, sink, cyc, sms);
  run<16, false>("LDG.128  4 lanes/line (bf16 wide)", table, l2_lines, 4, sink, cyc, sms);
  printf("== LDG.128, 8 lanes/line: hits mixed with misses; a table of 256 lines spread over 64 MB (one per 256 KB)\n");
  run<16, false, 1>("25% of the GROUPS miss to L2", table, l1_lines, 8, sink, cyc, sms, l2_lines / 2 - 1);
  run<16, false, 2>("25% of the INSTRUCTIONS miss to L2", table, l1_lines, 8, sink, cyc, sms, l2_lines / 2 - 1);
  run<16, false, 3>("256 L1-resident lines on 256 pages", table, l1_lines, 8, sink, cyc, sms, 0);
  run<16, false, 4>("256 lines at a 1 KB stride (one head)", table, l1_lines, 8, sink, cyc, sms, 0);
  run<16, false, 4>("64 lines at a 1 KB stride (one head)", table, 64, 8, sink, cyc, sms, 0);
  run<16, false, 5>("256 x 4 lines, groups = 4 heads", table, l1_lines, 8, sink, cyc, sms, 0);
  run<16, false, 5>("64 x 4 lines, groups = 4 heads", table, 64, 8, sink, cyc, sms, 0);
  run<16, false, 6>("L2: 4 adjacent lines (512 B) per instr", table, l2_lines, 8, sink, cyc, sms, 0);
  run<16, false, 0>("L2: 4 random lines per instr", table, l2_lines, 8, sink, cyc, sms, 0);
  run<16, false, 6>("L1: 4 adjacent lines per instr", table, l1_lines, 8, sink, cyc, sms, 0);
  printf("== shared-memory loads (32 KB table per CTA)\n");
  run<4, true>("LDS.32  32 lanes/row", table, l1_lines, 32, sink, cyc, sms);
  run<16, true>("LDS.128  8 lanes/row", table, l1_lines, 8, sink, cyc, sms);
  run<8, true>("LDS.64   8 lanes/row", table, l1_lines, 8, sink, cyc, sms);
  run<16, true>("LDS.128  4 lanes/row", table, l1_lines, 4, sink, cyc, sms);
  return 0;
}
