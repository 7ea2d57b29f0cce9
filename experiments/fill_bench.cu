// Microbenchmark: zero-fill of a grad_value-sized buffer (45.5 MB at the DETR-encoder shape): cudaMemsetAsync vs a plain
// 16 B-store kernel at several grid sizes.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fill_bench fill_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) fill_kernel(float4* p, size_t n) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = z;
}
__global__ void __launch_bounds__(256) fill_kernel4(float4* p, size_t n) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) { p[i] = z; p[i + stride] = z; p[i + 2 * stride] = z; p[i + 3 * stride] = z; }
  for (; i < n; i += stride) p[i] = z;
}
int main() {
  for (size_t bytes : {(size_t)45512704, (size_t)34242560, (size_t)368640000}) {
    float4 *d, *other;
    cudaMalloc(&d, bytes); cudaMalloc(&other, 256u << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time = [&](const char* name, auto fn) {
      float tot = 0;
      for (int r = 0; r < 12; ++r) {
        cudaMemsetAsync(other, 1, 256u << 20, 0);   // push the buffer out of L2
        cudaEventRecord(e0); fn(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r >= 2) tot += ms;
      }
      printf("%9.1f MB  %-28s %7.2f us  %6.0f GB/s\n", bytes / 1e6, name, tot / 10 * 1e3, bytes / (tot / 10) / 1e6);
    };
    time("cudaMemsetAsync", [&] { cudaMemsetAsync(d, 0, bytes, 0); });
    for (int ctas : {148 * 2, 148 * 4, 148 * 8, 148 * 16, 148 * 64})
      { char nm[64]; snprintf(nm, 64, "store kernel, %d CTAs", ctas); time(nm, [&] { fill_kernel<<<ctas, 256>>>(d, bytes / 16); }); }
    time("store kernel x4, 1184 CTAs", [&] { fill_kernel4<<<1184, 256>>>(d, bytes / 16); });
    cudaFree(d); cudaFree(other);
  }
  return 0;
}
