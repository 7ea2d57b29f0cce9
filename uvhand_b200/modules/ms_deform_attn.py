"""``MSDeformAttn`` -- the nn.Module that produces the operator's inputs and consumes its output.

Drop-in for /root/reference/models/ops/modules/ms_deform_attn.py:30-140 (and its two sibling
copies): same constructor, same four ``nn.Linear`` children with the same
names (``sampling_offsets``, ``attention_weights``, ``value_proj``, ``output_proj``) so reference
checkpoints load unchanged, same initialisation, same forward contract.  The hot op goes through
``MSDeformAttnFunction`` (sm_100a kernels); everything else here is plain torch plumbing around it.

Reference-point forms accepted (last dim of ``reference_points``):
  2   (x, y) per level                           -- ms_deform_attn.py:105-108
  42  21 hand keypoints (x0,y0,...,x20,y20); their mean is the sampling centre -- :110-128 (UVHand)
  4   (cx, cy, w, h) boxes                       -- models/dn_dab_dino_deformable_detr/ops/modules/ms_deform_attn.py:106-108
"""
from __future__ import annotations

import math
import warnings

import torch
import torch.nn.functional as F
from torch import nn

import os

from ..functions.ms_deform_attn_func import (MSDeformAttnFunction, MSDeformAttnFusedFunction, ValueToBF16Function,
                                             bf16_value_mode, fused_supported)


def _current_bf16_mode() -> bool:
    """the process-wide bf16 switch as it stands (modules without value_dtype leave it alone)"""
    from ..functions.ms_deform_attn_func import get_bf16_value_mode
    return get_bf16_value_mode()


def _power_of_two(n) -> bool:
    if not isinstance(n, int) or n < 0:
        raise ValueError(f"invalid input for _is_power_of_2: {n} (type: {type(n)})")
    return n != 0 and (n & (n - 1)) == 0


class MSDeformAttn(nn.Module):
    def __init__(self, d_model: int = 256, n_levels: int = 4, n_heads: int = 8, n_points: int = 4,
                 fuse_prologue: bool | None = None, value_dtype: torch.dtype | None = None):
        """Two switches that are not in the reference (same parameters, same checkpoints either way):

        ``fuse_prologue`` (default ON; ``False`` or env MSDA_B200_FUSE_PROLOGUE=0 turns it off): the kernels compute
        softmax(attention_weights) and reference + offsets / (W, H) themselves instead of materialising
        sampling_locations / attention_weights (locations bit-identical to torch's, results
        equal to rounding).  Applies to fp32, 32 channels per head, 2- or 42-dim reference points; everything
        else (box references, other head sizes, half / double modules) takes the unfused kernels as before.

        ``value_dtype=torch.bfloat16`` (default None = the reference's fp32 value): the value projection's output is
        written ONCE as bf16 with the padding mask applied in the same pass and the bf16-value kernels gather from
        it (half the gather bytes); grad_value is accumulated in fp32 and rounded once.
        Output and the other gradients stay fp32; tolerance of this mode: 1e-2 (BASELINE.json north_star)."""
        super().__init__()
        if fuse_prologue is None:
            fuse_prologue = os.environ.get("MSDA_B200_FUSE_PROLOGUE", "1") not in ("", "0", "false", "False")
        self.fuse_prologue = bool(fuse_prologue)
        if value_dtype not in (None, torch.float32, torch.bfloat16):
            raise ValueError("value_dtype must be None, torch.float32 or torch.bfloat16")
        self.value_dtype = None if value_dtype == torch.float32 else value_dtype
        if d_model % n_heads != 0:
            raise ValueError(f"d_model must be divisible by n_heads, but got {d_model} and {n_heads}")
        if not _power_of_two(d_model // n_heads):
            warnings.warn("MSDeformAttn: channels per head is not a power of two; the sm_100a fast "
                          "path is specialised for 32 channels per head, other sizes use the "
                          "any-shape kernels.")
        self.im2col_step = 64          # kept for interface parity (ms_deform_attn.py:48)
        self.d_model = d_model
        self.n_levels = n_levels
        self.n_heads = n_heads
        self.n_points = n_points

        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self._reset_parameters()

    def _reset_parameters(self):
        # Offsets start as a ring: head h looks along angle 2*pi*h/M, point i sits i+1 pixels out
        # (ms_deform_attn.py:62-70); attention starts uniform (:73-74).
        nn.init.constant_(self.sampling_offsets.weight, 0.0)
        angle = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        ring = torch.stack([angle.cos(), angle.sin()], dim=-1)
        ring = ring / ring.abs().max(dim=-1, keepdim=True)[0]
        ring = ring.view(self.n_heads, 1, 1, 2).repeat(1, self.n_levels, self.n_points, 1)
        ring = ring * torch.arange(1, self.n_points + 1, dtype=torch.float32).view(1, 1, -1, 1)
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(ring.reshape(-1))
        nn.init.constant_(self.attention_weights.weight, 0.0)
        nn.init.constant_(self.attention_weights.bias, 0.0)
        nn.init.xavier_uniform_(self.value_proj.weight)
        nn.init.constant_(self.value_proj.bias, 0.0)
        nn.init.xavier_uniform_(self.output_proj.weight)
        nn.init.constant_(self.output_proj.bias, 0.0)

    @staticmethod
    def _sampling_locations(reference_points, offsets, spatial_shapes):
        """reference_points [N, Lq, L, C], offsets [N, Lq, M, L, P, 2] -> locations, same shape as offsets."""
        # (W_l, H_l): offsets are in pixels of their own level
        wh = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], dim=-1)
        wh = wh[None, None, None, :, None, :]
        c = reference_points.shape[-1]
        if c == 2:
            return reference_points[:, :, None, :, None, :] + offsets / wh
        if c == 42:
            # centre of the 21 keypoints (the "center refine" variant the reference ships)
            kp = reference_points[:, :, None, :, None, :]
            centre = torch.stack([kp[..., 0::2].mean(-1), kp[..., 1::2].mean(-1)], dim=-1)
            return centre + offsets / wh
        if c == 4:
            n_points = offsets.shape[4]
            box = reference_points[:, :, None, :, None, :]
            return box[..., :2] + offsets / n_points * box[..., 2:] * 0.5
        raise ValueError(f"Last dim of reference_points must be 2, 4 or 42, but get {c} instead.")

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes,
                input_level_start_index, input_padding_mask=None):
        """query [N, Lq, C]; reference_points [N, Lq, L, 2|4|42] in [0,1] image coordinates;
        input_flatten [N, S, C] with S = sum_l H_l*W_l; input_spatial_shapes [L, 2] = (H_l, W_l);
        input_level_start_index [L]; input_padding_mask [N, S] (True = padding).  -> [N, Lq, C]"""
        n, len_q, _ = query.shape
        n, len_in, _ = input_flatten.shape
        assert (input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum() == len_in

        value = self.value_proj(input_flatten)
        bf16_value = (self.value_dtype == torch.bfloat16 and value.is_cuda and value.dtype == torch.float32
                      and self.d_model // self.n_heads == 32 and self.n_levels <= 8
                      and self.n_levels * self.n_points <= 32)
        if bf16_value:
            value = ValueToBF16Function.apply(value, input_padding_mask)      # mask + cast in one pass
        elif input_padding_mask is not None:
            value = value.masked_fill(input_padding_mask[..., None], float(0))
        value = value.view(n, len_in, self.n_heads, self.d_model // self.n_heads)
        offsets = self.sampling_offsets(query).view(
            n, len_q, self.n_heads, self.n_levels, self.n_points, 2)
        weights = self.attention_weights(query).view(
            n, len_q, self.n_heads, self.n_levels * self.n_points)
        if self.fuse_prologue and not bf16_value and reference_points.shape[-1] in (2, 42):
            centre = reference_points
            if reference_points.shape[-1] == 42:     # keypoint centre (ms_deform_attn.py:123-124), [N, Lq, L, 2]
                centre = torch.stack([reference_points[..., 0::2].mean(-1), reference_points[..., 1::2].mean(-1)], -1)
            if fused_supported(value, offsets, centre):
                out = MSDeformAttnFusedFunction.apply(value, input_spatial_shapes, input_level_start_index,
                                                      centre.to(value.dtype).contiguous(), offsets, weights)
                return self.output_proj(out)
        weights = F.softmax(weights, -1).view(n, len_q, self.n_heads, self.n_levels, self.n_points)
        locations = self._sampling_locations(reference_points, offsets, input_spatial_shapes)

        if value.dtype == torch.float16:
            # models/dino/ops/modules/ms_deform_attn.py:124-131: half goes through fp32 and back
            out = MSDeformAttnFunction.apply(value.float(), input_spatial_shapes, input_level_start_index,
                                             locations.float(), weights.float(), self.im2col_step)
            out = out.to(torch.float16)
        else:
            if locations.dtype != weights.dtype:
                locations = locations.to(weights.dtype)
            with bf16_value_mode(bf16_value or _current_bf16_mode()):
                out = MSDeformAttnFunction.apply(value, input_spatial_shapes, input_level_start_index,
                                                 locations.contiguous(), weights, self.im2col_step)
        return self.output_proj(out)
