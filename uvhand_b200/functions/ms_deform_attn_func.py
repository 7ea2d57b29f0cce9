"""Autograd boundary of the operator -- same class name, same ``apply`` signature and the same
returned gradients as the reference's
/root/reference/models/ops/functions/ms_deform_attn_func.py:21-39, so every model of the
reference tree (H2OTR/ARCTIC, AssemblyHands, DINO, DN-DAB) picks the sm_100a kernels up unchanged.

    MSDeformAttnFunction.apply(value, value_spatial_shapes, value_level_start_index,
                               sampling_locations, attention_weights, im2col_step)

dtype policy at this boundary ("dtype policy"):
  * float32 value                     -> fp32 kernels (all three reference copies end up here);
  * float64 value with float64 loc/w  -> fp64 kernels (what models/dino/ops allows; the reference's
                                         own test.py fp64 checks need it);
  * bfloat16 value, float32 loc/w     -> native bf16-value kernels ONLY when the opt-in switch is
                                         on (``set_bf16_value_mode(True)`` or MSDA_B200_BF16_VALUE=1);
  * anything else                     -> ``value.to(float32)``, which is what models/ops does for
                                         every input (func.py:26,37).
There is no fallback implementation here: the extension must be present.
"""
from __future__ import annotations

import os

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import load_extension

MSDA = load_extension()

_bf16_value_mode = os.environ.get("MSDA_B200_BF16_VALUE", "0") not in ("", "0", "false", "False")


def set_bf16_value_mode(enabled: bool) -> None:
    """Opt in/out of handing bfloat16 ``value`` tensors to the kernels as they are (half the
    gather bytes; grad_value is then accumulated in bfloat16).  Off by default = reference numerics."""
    global _bf16_value_mode
    _bf16_value_mode = bool(enabled)


def get_bf16_value_mode() -> bool:
    return _bf16_value_mode


class bf16_value_mode:
    """``with bf16_value_mode(True): ...`` -- scoped form of :func:`set_bf16_value_mode` (the mode in force when the
    forward runs is remembered for that call's backward)."""

    def __init__(self, enabled: bool = True):
        self.enabled = bool(enabled)

    def __enter__(self):
        self.prev = get_bf16_value_mode()
        set_bf16_value_mode(self.enabled)
        return self

    def __exit__(self, *exc):
        set_bf16_value_mode(self.prev)
        return False


def _kernel_value(value: torch.Tensor, sampling_locations: torch.Tensor, native_bf16: bool) -> torch.Tensor:
    if value.dtype == torch.float32:
        return value
    if value.dtype == torch.float64 and sampling_locations.dtype == torch.float64:
        return value
    if value.dtype == torch.bfloat16 and native_bf16 and sampling_locations.dtype == torch.float32:
        return value
    return value.to(torch.float32)


class MSDeformAttnFunction(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        ctx.native_bf16 = _bf16_value_mode
        output = MSDA.ms_deform_attn_forward(
            _kernel_value(value, sampling_locations, ctx.native_bf16), value_spatial_shapes, value_level_start_index,
            sampling_locations, attention_weights, ctx.im2col_step)
        # like the reference, the tensor as given (not the cast copy) is what is kept alive
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index,
                              sampling_locations, attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights = \
            ctx.saved_tensors
        grad_value, grad_sampling_loc, grad_attn_weight = MSDA.ms_deform_attn_backward(
            _kernel_value(value, sampling_locations, ctx.native_bf16), value_spatial_shapes, value_level_start_index,
            sampling_locations, attention_weights, grad_output.contiguous(), ctx.im2col_step)
        return grad_value, None, None, grad_sampling_loc, grad_attn_weight, None


class ValueToBF16Function(Function):
    """value_proj output (float32 [N, S, C]) -> the bf16 ``value`` the kernels gather from, with the padding mask
    applied in the same pass (/root/reference/models/ops/modules/ms_deform_attn.py:96-98 does
    ``masked_fill`` as a pass of its own, and a bf16 mode would add a cast pass).  SURVEY.md 8f rank 2."""

    @staticmethod
    def forward(ctx, value, padding_mask):
        mask = None if padding_mask is None else padding_mask.contiguous()
        ctx.mask = mask
        return MSDA.ms_deform_attn_value_to_bf16(value.contiguous(), mask)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad):
        return MSDA.ms_deform_attn_value_grad_from_bf16(grad.contiguous(), ctx.mask), None


class MSDeformAttnFusedFunction(Function):
    """The operator with MSDeformAttn.forward's softmax and location arithmetic folded in
    (/root/reference/models/ops/modules/ms_deform_attn.py:100-108):

        MSDeformAttnFusedFunction.apply(value, value_spatial_shapes, value_level_start_index,
                                        reference_points,      # [N, Lq, L, 2]
                                        sampling_offsets,      # [N, Lq, M, L, P, 2]  raw Linear output
                                        attention_logits)      # [N, Lq, M, L*P]      raw Linear output

    == MSDeformAttnFunction.apply(value, shapes, lsi, reference_points[:, :, None, :, None, :]
                                  + sampling_offsets / (W_l, H_l), softmax(attention_logits), step)
    with identical sampling locations (same fp32 operations) and weights equal to rounding.  float32,
    D = 32, L <= 8, L*P <= 32 only (``fused_supported``); not part of the reference API -- the
    MSDeformAttn module uses it when constructed with ``fuse_prologue=True``."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, reference_points,
                sampling_offsets, attention_logits):
        out = MSDA.ms_deform_attn_fused_forward(value, value_spatial_shapes, value_level_start_index,
                                                reference_points, sampling_offsets, attention_logits)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, reference_points,
                              sampling_offsets, attention_logits)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, lsi, ref, offsets, logits = ctx.saved_tensors
        want_ref = ctx.needs_input_grad[3]
        gv, goff, glogit, gref = MSDA.ms_deform_attn_fused_backward(
            value, shapes, lsi, ref, offsets, logits, grad_output.contiguous(), want_ref)
        return gv, None, None, (gref if want_ref else None), goff, glogit


def fused_supported(value: torch.Tensor, sampling_offsets: torch.Tensor, reference_points: torch.Tensor) -> bool:
    """Whether the fused kernels serve these operands (else callers use MSDeformAttnFunction)."""
    return (value.is_cuda and value.dtype == torch.float32 and sampling_offsets.dtype == torch.float32
            and value.dim() == 4 and value.shape[3] == 32 and sampling_offsets.dim() == 6
            and sampling_offsets.shape[3] <= 8 and sampling_offsets.shape[3] * sampling_offsets.shape[4] <= 32
            and reference_points.shape[-1] == 2 and value.shape[1] * value.shape[2] < (1 << 27))
