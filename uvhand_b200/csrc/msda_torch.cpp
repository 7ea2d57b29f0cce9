// PyTorch-extension host layer: the python module `MultiScaleDeformableAttention` with the two
// symbols the reference exports, same names, same positional argument order, same results.
//
//   reference binding ............ /root/reference/models/ops/src/vision.cpp:13-16
//   reference dispatch ........... /root/reference/models/ops/src/ms_deform_attn.h:20-61
//   reference host launchers ..... /root/reference/models/ops/src/cuda/ms_deform_attn_cuda.cu:20-80, 83-153
//
// Torch is plumbing here (device memory, current stream, device guard); the work is done by the
// torch-free C ABI in libmsda_b200.so (include/msda_b200.h).  Differences from the reference host
// code, all deliberate:
//   * a device guard is taken, and launch errors raise instead of being printf'ed (cuh:948-952);
//   * output / grad_loc / grad_weight are allocated uninitialised (they are fully written) --
//     only grad_value is zero-filled;
//   * the im2col_step precondition is kept (same error) but the batch is processed in one launch;
//   * value may additionally be bfloat16 (with float32 locations / weights): opt-in mode.
// CPU tensors are rejected exactly like the reference ("Not implemented on the CPU").
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <vector>

#include "../../include/msda_b200.h"

namespace {

void check_rc(int rc, const char* what) {
  TORCH_CHECK(rc == MSDA_OK, what, ": ", msda_b200_strerror(rc), " (code ", rc, ")");
}

struct Sizes {
  int N, S, M, D, L, Lq, P;
};

Sizes common_checks(const at::Tensor& value, const at::Tensor& spatial_shapes,
                    const at::Tensor& level_start_index, const at::Tensor& sampling_loc,
                    const at::Tensor& attn_weight, int im2col_step) {
  // ms_deform_attn.h:28-38: non-CUDA value -> "Not implemented on the CPU"
  TORCH_CHECK(value.is_cuda(), "Not implemented on the CPU");
  // ms_deform_attn_cuda.cu:28-38
  TORCH_CHECK(value.is_contiguous(), "value tensor has to be contiguous");
  TORCH_CHECK(spatial_shapes.is_contiguous(), "spatial_shapes tensor has to be contiguous");
  TORCH_CHECK(level_start_index.is_contiguous(), "level_start_index tensor has to be contiguous");
  TORCH_CHECK(sampling_loc.is_contiguous(), "sampling_loc tensor has to be contiguous");
  TORCH_CHECK(attn_weight.is_contiguous(), "attn_weight tensor has to be contiguous");
  TORCH_CHECK(spatial_shapes.is_cuda(), "spatial_shapes must be a CUDA tensor");
  TORCH_CHECK(level_start_index.is_cuda(), "level_start_index must be a CUDA tensor");
  TORCH_CHECK(sampling_loc.is_cuda(), "sampling_loc must be a CUDA tensor");
  TORCH_CHECK(attn_weight.is_cuda(), "attn_weight must be a CUDA tensor");
  // .data<int64_t>() in the reference throws for any other dtype (.cu:67-68)
  TORCH_CHECK(spatial_shapes.scalar_type() == at::kLong, "spatial_shapes must be int64");
  TORCH_CHECK(level_start_index.scalar_type() == at::kLong, "level_start_index must be int64");
  TORCH_CHECK(value.dim() == 4, "value must be [N, S, M, D]");
  TORCH_CHECK(sampling_loc.dim() == 6 && sampling_loc.size(5) == 2, "sampling_loc must be [N, Lq, M, L, P, 2]");
  TORCH_CHECK(attn_weight.dim() == 5, "attn_weight must be [N, Lq, M, L, P]");
  TORCH_CHECK(spatial_shapes.dim() == 2 && spatial_shapes.size(1) == 2, "spatial_shapes must be [L, 2]");
  const auto dev = value.device();
  TORCH_CHECK(spatial_shapes.device() == dev && level_start_index.device() == dev &&
                  sampling_loc.device() == dev && attn_weight.device() == dev,
              "all tensors must be on the same CUDA device");
  Sizes s;
  s.N = (int)value.size(0);
  s.S = (int)value.size(1);
  s.M = (int)value.size(2);
  s.D = (int)value.size(3);
  s.L = (int)spatial_shapes.size(0);
  s.Lq = (int)sampling_loc.size(1);
  s.P = (int)sampling_loc.size(4);
  TORCH_CHECK(sampling_loc.size(0) == s.N && sampling_loc.size(2) == s.M && sampling_loc.size(3) == s.L,
              "sampling_loc shape does not match value / spatial_shapes");
  TORCH_CHECK(attn_weight.size(0) == s.N && attn_weight.size(1) == s.Lq && attn_weight.size(2) == s.M &&
                  attn_weight.size(3) == s.L && attn_weight.size(4) == s.P,
              "attn_weight shape does not match sampling_loc");
  TORCH_CHECK(level_start_index.numel() == s.L, "level_start_index must have one entry per level");
  // ms_deform_attn_cuda.cu:50-52
  const int step = std::min(s.N, im2col_step);
  TORCH_CHECK(msda_b200_check_im2col_step(s.N, im2col_step) == MSDA_OK, "batch(", s.N,
              ") must divide im2col_step(", step, ")");
  return s;
}

enum class Mode { F32, F64, BF16 };

Mode pick_mode(const at::Tensor& value, const at::Tensor& sampling_loc, const at::Tensor& attn_weight) {
  const auto vt = value.scalar_type();
  if (vt == at::kFloat || vt == at::kDouble) {
    // AT_DISPATCH_FLOATING_TYPES on value + .data<scalar_t>() on the others (.cu:64-71)
    TORCH_CHECK(sampling_loc.scalar_type() == vt && attn_weight.scalar_type() == vt,
                "sampling_loc and attn_weight must have the dtype of value (", value.dtype(), ")");
    return vt == at::kFloat ? Mode::F32 : Mode::F64;
  }
  if (vt == at::kBFloat16) {
    TORCH_CHECK(sampling_loc.scalar_type() == at::kFloat && attn_weight.scalar_type() == at::kFloat,
                "bfloat16 value requires float32 sampling_loc and attn_weight");
    return Mode::BF16;
  }
  TORCH_CHECK(false, "\"ms_deform_attn\" not implemented for '", value.dtype(), "'");
  return Mode::F32;
}

}  // namespace

at::Tensor ms_deform_attn_forward(const at::Tensor& value, const at::Tensor& spatial_shapes,
                                  const at::Tensor& level_start_index, const at::Tensor& sampling_loc,
                                  const at::Tensor& attn_weight, const int im2col_step) {
  const Sizes s = common_checks(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step);
  const Mode mode = pick_mode(value, sampling_loc, attn_weight);
  const c10::cuda::CUDAGuard guard(value.device());
  void* stream = at::cuda::getCurrentCUDAStream().stream();
  const auto out_opts = mode == Mode::BF16 ? sampling_loc.options() : value.options();
  at::Tensor output = at::empty({s.N, s.Lq, s.M, s.D}, out_opts);
  const int64_t* shp = spatial_shapes.data_ptr<int64_t>();
  const int64_t* lsi = level_start_index.data_ptr<int64_t>();
  int rc = MSDA_OK;
  switch (mode) {
    case Mode::F32:
      rc = msda_fwd_f32(value.data_ptr<float>(), shp, lsi, sampling_loc.data_ptr<float>(),
                        attn_weight.data_ptr<float>(), s.N, s.S, s.M, s.D, s.L, s.Lq, s.P,
                        output.data_ptr<float>(), stream);
      break;
    case Mode::F64:
      rc = msda_fwd_f64(value.data_ptr<double>(), shp, lsi, sampling_loc.data_ptr<double>(),
                        attn_weight.data_ptr<double>(), s.N, s.S, s.M, s.D, s.L, s.Lq, s.P,
                        output.data_ptr<double>(), stream);
      break;
    case Mode::BF16:
      rc = msda_fwd_bf16(reinterpret_cast<const uint16_t*>(value.data_ptr()), shp, lsi,
                         sampling_loc.data_ptr<float>(), attn_weight.data_ptr<float>(), s.N, s.S, s.M,
                         s.D, s.L, s.Lq, s.P, output.data_ptr<float>(), stream);
      break;
  }
  check_rc(rc, "ms_deform_attn_forward");
  return output.view({s.N, s.Lq, s.M * s.D});  // .cu:77
}

std::vector<at::Tensor> ms_deform_attn_backward(const at::Tensor& value, const at::Tensor& spatial_shapes,
                                                const at::Tensor& level_start_index,
                                                const at::Tensor& sampling_loc, const at::Tensor& attn_weight,
                                                const at::Tensor& grad_output, const int im2col_step) {
  const Sizes s = common_checks(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step);
  const Mode mode = pick_mode(value, sampling_loc, attn_weight);
  TORCH_CHECK(grad_output.is_contiguous(), "grad_output tensor has to be contiguous");  // .cu:98
  TORCH_CHECK(grad_output.is_cuda(), "grad_output must be a CUDA tensor");              // .cu:105
  TORCH_CHECK(grad_output.device() == value.device(), "grad_output must be on value's device");
  TORCH_CHECK(grad_output.scalar_type() == sampling_loc.scalar_type(),
              "grad_output must have the dtype of sampling_loc");
  TORCH_CHECK(grad_output.numel() == (int64_t)s.N * s.Lq * s.M * s.D, "grad_output must be [N, Lq, M*D]");
  const c10::cuda::CUDAGuard guard(value.device());
  void* stream = at::cuda::getCurrentCUDAStream().stream();
  at::Tensor grad_value = at::zeros_like(value);  // accumulated into (.cu:121)
  at::Tensor grad_loc = at::empty_like(sampling_loc);
  at::Tensor grad_aw = at::empty_like(attn_weight);
  if (s.L == 0 || s.P == 0 || s.S == 0) {
    grad_loc.zero_();
    grad_aw.zero_();
  }
  const int64_t* shp = spatial_shapes.data_ptr<int64_t>();
  const int64_t* lsi = level_start_index.data_ptr<int64_t>();
  int rc = MSDA_OK;
  switch (mode) {
    case Mode::F32:
      rc = msda_bwd_f32(grad_output.data_ptr<float>(), value.data_ptr<float>(), shp, lsi,
                        sampling_loc.data_ptr<float>(), attn_weight.data_ptr<float>(), s.N, s.S, s.M, s.D,
                        s.L, s.Lq, s.P, grad_value.data_ptr<float>(), grad_loc.data_ptr<float>(),
                        grad_aw.data_ptr<float>(), stream);
      break;
    case Mode::F64:
      rc = msda_bwd_f64(grad_output.data_ptr<double>(), value.data_ptr<double>(), shp, lsi,
                        sampling_loc.data_ptr<double>(), attn_weight.data_ptr<double>(), s.N, s.S, s.M, s.D,
                        s.L, s.Lq, s.P, grad_value.data_ptr<double>(), grad_loc.data_ptr<double>(),
                        grad_aw.data_ptr<double>(), stream);
      break;
    case Mode::BF16: {
      // accurate mode by default: fp32 accumulation scratch, one rounding into the bf16 grad_value
      int direct = 0;
      msda_b200_get_option("bf16_direct_red", &direct);
      at::Tensor scratch;
      float* scratch_ptr = nullptr;
      if (!direct) {
        scratch = at::zeros(value.sizes(), value.options().dtype(at::kFloat));
        scratch_ptr = scratch.data_ptr<float>();
      }
      rc = msda_bwd_bf16(grad_output.data_ptr<float>(), reinterpret_cast<const uint16_t*>(value.data_ptr()),
                         shp, lsi, sampling_loc.data_ptr<float>(), attn_weight.data_ptr<float>(), s.N, s.S,
                         s.M, s.D, s.L, s.Lq, s.P, reinterpret_cast<uint16_t*>(grad_value.data_ptr()),
                         grad_loc.data_ptr<float>(), grad_aw.data_ptr<float>(), scratch_ptr, stream);
      break;
    }
  }
  check_rc(rc, "ms_deform_attn_backward");
  return {grad_value, grad_loc, grad_aw};
}

// ---- fused prologue: softmax + location arithmetic folded into the operator -------------------
// Not in the reference extension; replaces, inside MSDeformAttn.forward, the elementwise steps of
// /root/reference/models/ops/modules/ms_deform_attn.py:100-108.
namespace {

Sizes fused_checks(const at::Tensor& value, const at::Tensor& spatial_shapes, const at::Tensor& level_start_index,
                   const at::Tensor& ref, const at::Tensor& offsets, const at::Tensor& logits) {
  TORCH_CHECK(value.is_cuda(), "Not implemented on the CPU");
  for (const at::Tensor* t : {&value, &spatial_shapes, &level_start_index, &ref, &offsets, &logits}) {
    TORCH_CHECK(t->is_contiguous(), "fused MSDeformAttn: every tensor has to be contiguous");
    TORCH_CHECK(t->is_cuda() && t->device() == value.device(), "fused MSDeformAttn: all tensors must be on value's device");
  }
  TORCH_CHECK(value.scalar_type() == at::kFloat && ref.scalar_type() == at::kFloat &&
                  offsets.scalar_type() == at::kFloat && logits.scalar_type() == at::kFloat,
              "fused MSDeformAttn is float32 only");
  TORCH_CHECK(spatial_shapes.scalar_type() == at::kLong && level_start_index.scalar_type() == at::kLong,
              "spatial_shapes / level_start_index must be int64");
  TORCH_CHECK(value.dim() == 4 && offsets.dim() == 6 && offsets.size(5) == 2, "value [N,S,M,D], sampling_offsets [N,Lq,M,L,P,2]");
  Sizes s;
  s.N = (int)value.size(0); s.S = (int)value.size(1); s.M = (int)value.size(2); s.D = (int)value.size(3);
  s.L = (int)spatial_shapes.size(0); s.Lq = (int)offsets.size(1); s.P = (int)offsets.size(4);
  TORCH_CHECK(offsets.size(0) == s.N && offsets.size(2) == s.M && offsets.size(3) == s.L,
              "sampling_offsets shape does not match value / spatial_shapes");
  TORCH_CHECK(logits.numel() == (int64_t)s.N * s.Lq * s.M * s.L * s.P, "attn_logits must be [N, Lq, M, L*P]");
  TORCH_CHECK(ref.dim() == 4 && ref.size(0) == s.N && ref.size(1) == s.Lq && ref.size(2) == s.L && ref.size(3) == 2,
              "reference_points must be [N, Lq, L, 2]");
  TORCH_CHECK(level_start_index.numel() == s.L, "level_start_index must have one entry per level");
  return s;
}

}  // namespace

at::Tensor ms_deform_attn_fused_forward(const at::Tensor& value, const at::Tensor& spatial_shapes,
                                        const at::Tensor& level_start_index, const at::Tensor& reference_points,
                                        const at::Tensor& sampling_offsets, const at::Tensor& attn_logits) {
  const Sizes s = fused_checks(value, spatial_shapes, level_start_index, reference_points, sampling_offsets, attn_logits);
  const c10::cuda::CUDAGuard guard(value.device());
  at::Tensor output = at::empty({s.N, s.Lq, s.M, s.D}, value.options());
  check_rc(msda_fused_fwd_f32(value.data_ptr<float>(), spatial_shapes.data_ptr<int64_t>(),
                              level_start_index.data_ptr<int64_t>(), reference_points.data_ptr<float>(),
                              sampling_offsets.data_ptr<float>(), attn_logits.data_ptr<float>(), s.N, s.S, s.M,
                              s.D, s.L, s.Lq, s.P, output.data_ptr<float>(),
                              at::cuda::getCurrentCUDAStream().stream()),
           "ms_deform_attn_fused_forward");
  return output.view({s.N, s.Lq, s.M * s.D});
}

// -> [grad_value, grad_sampling_offsets, grad_attn_logits, grad_reference_points (empty tensor if not wanted)]
std::vector<at::Tensor> ms_deform_attn_fused_backward(const at::Tensor& value, const at::Tensor& spatial_shapes,
                                                      const at::Tensor& level_start_index,
                                                      const at::Tensor& reference_points,
                                                      const at::Tensor& sampling_offsets,
                                                      const at::Tensor& attn_logits, const at::Tensor& grad_output,
                                                      const bool want_grad_reference) {
  const Sizes s = fused_checks(value, spatial_shapes, level_start_index, reference_points, sampling_offsets, attn_logits);
  TORCH_CHECK(grad_output.is_contiguous() && grad_output.is_cuda() && grad_output.scalar_type() == at::kFloat &&
                  grad_output.numel() == (int64_t)s.N * s.Lq * s.M * s.D,
              "grad_output must be a contiguous float32 CUDA tensor [N, Lq, M*D]");
  const c10::cuda::CUDAGuard guard(value.device());
  at::Tensor grad_value = at::zeros_like(value);
  at::Tensor grad_off = at::empty_like(sampling_offsets);
  at::Tensor grad_logits = at::empty_like(attn_logits);
  at::Tensor grad_ref = want_grad_reference ? at::zeros_like(reference_points) : at::Tensor();
  if (s.L == 0 || s.P == 0 || s.S == 0) {
    grad_off.zero_();
    grad_logits.zero_();
  }
  check_rc(msda_fused_bwd_f32(grad_output.data_ptr<float>(), value.data_ptr<float>(),
                              spatial_shapes.data_ptr<int64_t>(), level_start_index.data_ptr<int64_t>(),
                              reference_points.data_ptr<float>(), sampling_offsets.data_ptr<float>(),
                              attn_logits.data_ptr<float>(), s.N, s.S, s.M, s.D, s.L, s.Lq, s.P,
                              grad_value.data_ptr<float>(), grad_off.data_ptr<float>(), grad_logits.data_ptr<float>(),
                              want_grad_reference ? grad_ref.data_ptr<float>() : nullptr,
                              at::cuda::getCurrentCUDAStream().stream()),
           "ms_deform_attn_fused_backward");
  return {grad_value, grad_off, grad_logits, grad_ref};
}

// value producer side: mask + cast of the value projection's output in one pass
namespace {
const uint8_t* mask_ptr(const at::Tensor& value, const c10::optional<at::Tensor>& mask, int64_t rows) {
  if (!mask.has_value() || !mask->defined()) return nullptr;
  TORCH_CHECK(mask->is_cuda() && mask->device() == value.device() && mask->is_contiguous() && mask->numel() == rows &&
                  (mask->scalar_type() == at::kBool || mask->scalar_type() == at::kByte),
              "padding mask must be a contiguous bool/uint8 CUDA tensor with one entry per token");
  return reinterpret_cast<const uint8_t*>(mask->data_ptr());
}
}  // namespace

at::Tensor ms_deform_attn_value_to_bf16(const at::Tensor& value, const c10::optional<at::Tensor>& padding_mask) {
  TORCH_CHECK(value.is_cuda(), "Not implemented on the CPU");
  TORCH_CHECK(value.is_contiguous() && value.scalar_type() == at::kFloat && value.dim() >= 2, "value must be contiguous float32 [..., C]");
  const int64_t cols = value.size(-1), rows = value.numel() / std::max<int64_t>(cols, 1);
  const c10::cuda::CUDAGuard guard(value.device());
  at::Tensor out = at::empty(value.sizes(), value.options().dtype(at::kBFloat16));
  check_rc(msda_value_pack_bf16(value.data_ptr<float>(), mask_ptr(value, padding_mask, rows), rows, (int)cols,
                                reinterpret_cast<uint16_t*>(out.data_ptr()), at::cuda::getCurrentCUDAStream().stream()),
           "ms_deform_attn_value_to_bf16");
  return out;
}

at::Tensor ms_deform_attn_value_grad_from_bf16(const at::Tensor& grad, const c10::optional<at::Tensor>& padding_mask) {
  TORCH_CHECK(grad.is_cuda(), "Not implemented on the CPU");
  TORCH_CHECK(grad.is_contiguous() && grad.scalar_type() == at::kBFloat16 && grad.dim() >= 2, "grad must be contiguous bfloat16 [..., C]");
  const int64_t cols = grad.size(-1), rows = grad.numel() / std::max<int64_t>(cols, 1);
  const c10::cuda::CUDAGuard guard(grad.device());
  at::Tensor out = at::empty(grad.sizes(), grad.options().dtype(at::kFloat));
  check_rc(msda_value_unpack_grad_bf16(reinterpret_cast<const uint16_t*>(grad.data_ptr()), mask_ptr(grad, padding_mask, rows),
                                       rows, (int)cols, out.data_ptr<float>(), at::cuda::getCurrentCUDAStream().stream()),
           "ms_deform_attn_value_grad_from_bf16");
  return out;
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("ms_deform_attn_value_to_bf16", &ms_deform_attn_value_to_bf16, "mask + cast of the value projection output");
  m.def("ms_deform_attn_value_grad_from_bf16", &ms_deform_attn_value_grad_from_bf16, "backward of value_to_bf16");
  m.def("ms_deform_attn_forward", &ms_deform_attn_forward, "ms_deform_attn_forward");
  m.def("ms_deform_attn_backward", &ms_deform_attn_backward, "ms_deform_attn_backward");
  // additions (not in the reference module)
  m.def("ms_deform_attn_fused_forward", &ms_deform_attn_fused_forward, "softmax + location prologue fused in");
  m.def("ms_deform_attn_fused_backward", &ms_deform_attn_fused_backward, "backward of the fused operator");
  m.def("set_option", [](const std::string& key, int value) {
    check_rc(msda_b200_set_option(key.c_str(), value), "set_option");
  });
  m.def("launch_count", []() { return (int64_t)msda_b200_launch_count(); });
  m.def("last_path", [](bool backward) { return std::string(msda_b200_last_path(backward ? 1 : 0)); });
  m.def("abi_version", []() { return msda_b200_abi_version(); });
}
