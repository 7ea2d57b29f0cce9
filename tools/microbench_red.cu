// Microbenchmark: how fast can the SMs push 128 B row reductions into L2?
// Variants: scalar red.f32 (32 lanes/row), red.v2.f32 (16 lanes/row), red.v4.f32 (8 lanes/row),
// plain st.v4 (no reduction, same bytes), TMA bulk reduce (cp.reduce.async.bulk, 128 B per op from smem).
// Rows are pseudo-random inside a `span_rows` window (45 MB by default = L2 resident, like grad_value at C2).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_red tools/microbench_red.cu && ./microbench_red
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

// mode 0: scalar red, 1: v2, 2: v4, 3: st.v4, 4: v4 but only 1 of 4 rows distinct per instr (same row for all 4 groups)
template <int MODE>
__global__ void __launch_bounds__(256) red_kernel(float* dst, unsigned span_rows, int rows_per_thread_group, unsigned seed) {
  const unsigned gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31;
  const unsigned warp = gtid >> 5;
  const float v = 1.0f;
  for (int i = 0; i < rows_per_thread_group; ++i) {
    if (MODE == 0) {
      // one row per warp instruction, 4 instructions for 4 rows
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned row = hash32(seed + (warp * rows_per_thread_group + i) * 4 + k) % span_rows;
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + (size_t)row * 32 + lane), "f"(v) : "memory");
      }
    } else if (MODE == 1) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const unsigned row = hash32(seed + (warp * rows_per_thread_group + i) * 4 + k * 2 + (lane >> 4)) % span_rows;
        asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(dst + (size_t)row * 32 + (lane & 15) * 2), "f"(v), "f"(v) : "memory");
      }
    } else if (MODE == 2 || MODE == 4) {
      const unsigned grp = MODE == 4 ? 0 : (lane >> 3);
      const unsigned row = hash32(seed + (warp * rows_per_thread_group + i) * 4 + grp) % span_rows;
      asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + (size_t)row * 32 + (lane & 7) * 4), "f"(v), "f"(v), "f"(v), "f"(v) : "memory");
    } else if (MODE == 3) {
      const unsigned row = hash32(seed + (warp * rows_per_thread_group + i) * 4 + (lane >> 3)) % span_rows;
      asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + (size_t)row * 32 + (lane & 7) * 4), "f"(v), "f"(v), "f"(v), "f"(v) : "memory");
    }
  }
}

// TMA bulk reduce: each warp stages 4 rows (4 x 128 B) in smem, one lane per row issues
// cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 of 128 B.
template <int ROW_BYTES>
__global__ void __launch_bounds__(256) tma_red_kernel(float* dst, unsigned span_rows, int iters, unsigned seed) {
  __shared__ __align__(128) float stage[8][2][4][32];  // warp, double buffer, row, channel
  const unsigned gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const unsigned warp = gtid >> 5;
  for (int i = 0; i < iters; ++i) {
    const int buf = i & 1;
    // make sure the bulk ops that read this buffer two iterations ago are done reading
    if (lane < 4) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
    float4* s = reinterpret_cast<float4*>(&stage[wib][buf][lane >> 3][(lane & 7) * 4]);
    *s = make_float4(1.f, 1.f, 1.f, 1.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane < 4) {
      const unsigned row = hash32(seed + (warp * iters + i) * 4 + lane) % span_rows;
      const unsigned saddr = (unsigned)__cvta_generic_to_shared(&stage[wib][buf][lane][0]);
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst + (size_t)row * 32),
                   "r"(saddr), "n"(ROW_BYTES)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (lane < 4) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// Mixed: warps [0, LSU_WARPS) of every CTA reduce with red.global.add.v4.f32, the others with TMA bulk
// reduce.  If the two paths had separate ports out of the SM their throughputs would add.
template <int LSU_WARPS>
__global__ void __launch_bounds__(256) mixed_red_kernel(float* dst, unsigned span_rows, int iters, unsigned seed) {
  __shared__ __align__(128) float stage[8][2][4][32];
  const unsigned gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const unsigned warp = gtid >> 5;
  if ((int)wib < LSU_WARPS) {
    for (int i = 0; i < iters; ++i) {
      const unsigned row = hash32(seed + (warp * iters + i) * 4 + (lane >> 3)) % span_rows;
      asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + (size_t)row * 32 + (lane & 7) * 4), "f"(1.f), "f"(1.f), "f"(1.f), "f"(1.f) : "memory");
    }
    return;
  }
  for (int i = 0; i < iters; ++i) {
    const int buf = i & 1;
    if (lane < 4) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
    float4* s = reinterpret_cast<float4*>(&stage[wib][buf][lane >> 3][(lane & 7) * 4]);
    *s = make_float4(1.f, 1.f, 1.f, 1.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane < 4) {
      const unsigned row = hash32(seed + (warp * iters + i) * 4 + lane) % span_rows;
      const unsigned saddr = (unsigned)__cvta_generic_to_shared(&stage[wib][buf][lane][0]);
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst + (size_t)row * 32),
                   "r"(saddr), "n"(128)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (lane < 4) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char** argv) {
  const unsigned span_rows = argc > 1 ? (unsigned)atoi(argv[1]) : 355568u;  // 45.5 MB of 128 B rows
  const int total_rows = 22757376;                                         // C2 backward: samples * 4 corners
  float* dst;
  CK(cudaMalloc(&dst, (size_t)span_rows * 128));
  CK(cudaMemset(dst, 0, (size_t)span_rows * 128));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int threads = 256, warps_per_cta = 8;
  const int ctas = 148 * 16;
  const int per_group = total_rows / 4 / (ctas * warps_per_cta);  // iterations (4 rows per warp-iteration)
  const double bytes = (double)per_group * 4 * ctas * warps_per_cta * 128.0;
  const char* names[] = {"red.f32 scalar (32 lanes/row)", "red.v2.f32 (16 lanes/row)", "red.v4.f32 (8 lanes/row)",
                         "st.v4.f32 (no reduction)", "red.v4.f32, 4 groups same row"};
  for (int mode = 0; mode < 5; ++mode) {
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      CK(cudaEventRecord(e0));
      switch (mode) {
        case 0: red_kernel<0><<<ctas, threads>>>(dst, span_rows, per_group, rep * 7919u); break;
        case 1: red_kernel<1><<<ctas, threads>>>(dst, span_rows, per_group, rep * 7919u); break;
        case 2: red_kernel<2><<<ctas, threads>>>(dst, span_rows, per_group, rep * 7919u); break;
        case 3: red_kernel<3><<<ctas, threads>>>(dst, span_rows, per_group, rep * 7919u); break;
        case 4: red_kernel<4><<<ctas, threads>>>(dst, span_rows, per_group, rep * 7919u); break;
      }
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (ms < best) best = ms;
    }
    printf("%-36s %8.1f us   %7.2f TB/s payload   (%d rows of 128 B over a %.1f MB window)\n", names[mode], best * 1e3,
           bytes / (best * 1e-3) / 1e12, (int)(bytes / 128), span_rows * 128.0 / 1e6);
  }
  {
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      CK(cudaEventRecord(e0));
      tma_red_kernel<128><<<ctas, threads>>>(dst, span_rows, per_group, rep * 7919u);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (ms < best) best = ms;
    }
    printf("%-36s %8.1f us   %7.2f TB/s payload\n", "cp.reduce.async.bulk 128 B/op", best * 1e3, bytes / (best * 1e-3) / 1e12);
  }
  for (int split = 0; split < 3; ++split) {
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      CK(cudaEventRecord(e0));
      if (split == 0) mixed_red_kernel<4><<<ctas, threads>>>(dst, span_rows, per_group, rep * 7919u);
      if (split == 1) mixed_red_kernel<5><<<ctas, threads>>>(dst, span_rows, per_group, rep * 7919u);
      if (split == 2) mixed_red_kernel<6><<<ctas, threads>>>(dst, span_rows, per_group, rep * 7919u);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (ms < best) best = ms;
    }
    printf("mixed: %d of 8 warps red.v4, rest TMA      %8.1f us   %7.2f TB/s payload\n", 4 + split, best * 1e3,
           bytes / (best * 1e-3) / 1e12);
  }
  CK(cudaDeviceSynchronize());
  printf("done\n");
  return 0;
}
