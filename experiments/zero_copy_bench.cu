// Microbenchmark: PCIe throughput when the SMs read / write pinned host memory directly (zero-copy) instead of the
// copy engines.  Evidence for the host-buffer entry points' design; not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o zero_copy_bench zero_copy_bench.cu && ./zero_copy_bench
#include <cstdio>
#include <cuda_runtime.h>

__global__ void copy_kernel(const float4* __restrict__ src, float4* __restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float4 v;
    asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(src + i));
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  }
}
// 4 loads in flight per thread
__global__ void copy_kernel4(const float4* __restrict__ src, float4* __restrict__ dst, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += 4 * stride) {
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) if (i + j * stride < n) v[j] = __ldcs(src + i + j * stride);
#pragma unroll
    for (int j = 0; j < 4; ++j) if (i + j * stride < n) __stcs(dst + i + j * stride, v[j]);
  }
}

int main() {
  const size_t bytes = 160ull << 20, n = bytes / 16;
  float4 *h_in, *h_out, *d_a, *d_b;
  cudaHostAlloc(&h_in, bytes, cudaHostAllocMapped);
  cudaHostAlloc(&h_out, bytes, cudaHostAllocMapped);
  cudaMalloc(&d_a, bytes);
  cudaMalloc(&d_b, bytes);
  for (size_t i = 0; i < n; ++i) h_in[i] = make_float4(1.f, 2.f, 3.f, 4.f);
  cudaMemset(d_b, 0, bytes);
  cudaStream_t s1, s2;
  cudaStreamCreate(&s1); cudaStreamCreate(&s2);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto time = [&](const char* name, auto fn) {
    fn(); cudaDeviceSynchronize();
    cudaEventRecord(e0, 0);
    cudaStreamWaitEvent(s1, e0, 0); cudaStreamWaitEvent(s2, e0, 0);
    const int reps = 5;
    for (int r = 0; r < reps; ++r) fn();
    cudaEvent_t j1, j2; cudaEventCreate(&j1); cudaEventCreate(&j2);
    cudaEventRecord(j1, s1); cudaEventRecord(j2, s2);
    cudaStreamWaitEvent(0, j1, 0); cudaStreamWaitEvent(0, j2, 0);
    cudaEventRecord(e1, 0);
    cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    printf("%-64s %7.3f ms  %6.1f GB/s per direction   (%s)\n", name, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  };
  for (int ctas : {148, 592, 2368}) {
    printf("-- %d CTAs x 256 threads\n", ctas);
    time("SM reads host -> device", [&] { copy_kernel<<<ctas, 256, 0, s1>>>(h_in, d_a, n); });
    time("SM reads host -> device, 4 loads in flight", [&] { copy_kernel4<<<ctas, 256, 0, s1>>>(h_in, d_a, n); });
    time("SM writes device -> host", [&] { copy_kernel<<<ctas, 256, 0, s2>>>(d_b, h_out, n); });
    time("both at once (two kernels)", [&] { copy_kernel4<<<ctas / 2, 256, 0, s1>>>(h_in, d_a, n); copy_kernel<<<ctas / 2, 256, 0, s2>>>(d_b, h_out, n); });
    time("one kernel: host -> host", [&] { copy_kernel4<<<ctas, 256, 0, s1>>>(h_in, h_out, n); });
  }
  printf("-- copy engines\n");
  time("cudaMemcpyAsync H2D", [&] { cudaMemcpyAsync(d_a, h_in, bytes, cudaMemcpyHostToDevice, s1); });
  time("cudaMemcpyAsync D2H", [&] { cudaMemcpyAsync(h_out, d_b, bytes, cudaMemcpyDeviceToHost, s2); });
  time("both at once", [&] { cudaMemcpyAsync(d_a, h_in, bytes, cudaMemcpyHostToDevice, s1); cudaMemcpyAsync(h_out, d_b, bytes, cudaMemcpyDeviceToHost, s2); });
  time("copy engine H2D + SM writes to host", [&] { cudaMemcpyAsync(d_a, h_in, bytes, cudaMemcpyHostToDevice, s1); copy_kernel<<<592, 256, 0, s2>>>(d_b, h_out, n); });
  time("SM reads host + copy engine D2H", [&] { copy_kernel4<<<592, 256, 0, s1>>>(h_in, d_a, n); cudaMemcpyAsync(h_out, d_b, bytes, cudaMemcpyDeviceToHost, s2); });
  return 0;
}
