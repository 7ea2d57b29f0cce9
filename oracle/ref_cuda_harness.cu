// TEST INFRASTRUCTURE ONLY.
//
// C-ABI harness around the REFERENCE's own CUDA kernels so that they can be run on the GPU
// box as a second oracle and as an informational speed comparison.  The kernel header is
// #included from where it lies under /root/reference (path injected by oracle/Makefile as
// MSDA_REF_CUH); no reference source is copied into this repository.  The output,
// oracle/_ref/libmsda_refcuda.so, is git-ignored.
//
// What is wrapped: the two templated launchers
//   ms_deformable_im2col_cuda<T>   models/ops/src/cuda/ms_deform_im2col_cuda.cuh:923-954
//   ms_deformable_col2im_cuda<T>   models/ops/src/cuda/ms_deform_im2col_cuda.cuh:956-1327
// with the chunk loop of ms_deform_attn_cuda.cu:50-75 / :119-148 restated around them
// (that host file itself does not compile against torch 2.11).
#include MSDA_REF_CUH

#include <cstdint>

namespace {

template <typename T>
int ref_fwd(const T* value, const int64_t* shapes, const int64_t* lsi, const T* loc, const T* aw,
            int N, int S, int M, int D, int L, int Lq, int P, int im2col_step, T* out,
            cudaStream_t stream) {
  const int step = N < im2col_step ? N : im2col_step;
  if (step <= 0 || N % step != 0) return -1;
  const long long per_value = (long long)S * M * D;
  const long long per_loc = (long long)Lq * M * L * P * 2;
  const long long per_aw = (long long)Lq * M * L * P;
  const long long per_out = (long long)Lq * M * D;
  cudaMemsetAsync(out, 0, sizeof(T) * N * per_out, stream);  // at::zeros, .cu:54
  for (int n = 0; n < N / step; ++n) {
    ms_deformable_im2col_cuda<T>(stream, value + n * step * per_value, shapes, lsi,
                                 loc + n * step * per_loc, aw + n * step * per_aw, step, S, M, D,
                                 L, Lq, P, out + n * step * per_out);
  }
  return (int)cudaGetLastError();
}

template <typename T>
int ref_bwd(const T* grad_out, const T* value, const int64_t* shapes, const int64_t* lsi,
            const T* loc, const T* aw, int N, int S, int M, int D, int L, int Lq, int P,
            int im2col_step, T* grad_value, T* grad_loc, T* grad_aw, cudaStream_t stream) {
  const int step = N < im2col_step ? N : im2col_step;
  if (step <= 0 || N % step != 0) return -1;
  const long long per_value = (long long)S * M * D;
  const long long per_loc = (long long)Lq * M * L * P * 2;
  const long long per_aw = (long long)Lq * M * L * P;
  const long long per_out = (long long)Lq * M * D;
  // zeros_like x3, .cu:121-123
  cudaMemsetAsync(grad_value, 0, sizeof(T) * N * per_value, stream);
  cudaMemsetAsync(grad_loc, 0, sizeof(T) * N * per_loc, stream);
  cudaMemsetAsync(grad_aw, 0, sizeof(T) * N * per_aw, stream);
  for (int n = 0; n < N / step; ++n) {
    ms_deformable_col2im_cuda<T>(stream, grad_out + n * step * per_out,
                                 value + n * step * per_value, shapes, lsi,
                                 loc + n * step * per_loc, aw + n * step * per_aw, step, S, M, D,
                                 L, Lq, P, grad_value + n * step * per_value,
                                 grad_loc + n * step * per_loc, grad_aw + n * step * per_aw);
  }
  return (int)cudaGetLastError();
}

}  // namespace

extern "C" {

int msda_ref_fwd_f32(const float* value, const int64_t* shapes, const int64_t* lsi,
                     const float* loc, const float* aw, int N, int S, int M, int D, int L, int Lq,
                     int P, int im2col_step, float* out, void* stream) {
  return ref_fwd<float>(value, shapes, lsi, loc, aw, N, S, M, D, L, Lq, P, im2col_step, out,
                        (cudaStream_t)stream);
}
int msda_ref_fwd_f64(const double* value, const int64_t* shapes, const int64_t* lsi,
                     const double* loc, const double* aw, int N, int S, int M, int D, int L,
                     int Lq, int P, int im2col_step, double* out, void* stream) {
  return ref_fwd<double>(value, shapes, lsi, loc, aw, N, S, M, D, L, Lq, P, im2col_step, out,
                         (cudaStream_t)stream);
}
int msda_ref_bwd_f32(const float* grad_out, const float* value, const int64_t* shapes,
                     const int64_t* lsi, const float* loc, const float* aw, int N, int S, int M,
                     int D, int L, int Lq, int P, int im2col_step, float* grad_value,
                     float* grad_loc, float* grad_aw, void* stream) {
  return ref_bwd<float>(grad_out, value, shapes, lsi, loc, aw, N, S, M, D, L, Lq, P, im2col_step,
                        grad_value, grad_loc, grad_aw, (cudaStream_t)stream);
}
int msda_ref_bwd_f64(const double* grad_out, const double* value, const int64_t* shapes,
                     const int64_t* lsi, const double* loc, const double* aw, int N, int S, int M,
                     int D, int L, int Lq, int P, int im2col_step, double* grad_value,
                     double* grad_loc, double* grad_aw, void* stream) {
  return ref_bwd<double>(grad_out, value, shapes, lsi, loc, aw, N, S, M, D, L, Lq, P,
                         im2col_step, grad_value, grad_loc, grad_aw, (cudaStream_t)stream);
}

}  // extern "C"
