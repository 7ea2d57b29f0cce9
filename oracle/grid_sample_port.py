"""TEST INFRASTRUCTURE ONLY -- port of the reference's CPU path for MSDeformAttn.

The reference's CPU implementation of the hot path is the pure-PyTorch function
``ms_deform_attn_core_pytorch`` (/root/reference/models/ops/functions/ms_deform_attn_func.py:42-62):
per level it reshapes that level's slice of ``value`` to an image batch ``[N*M, D, H, W]``,
maps the normalised sampling locations to grid_sample's [-1, 1] convention (``2*loc - 1``,
func.py:48) and calls ``F.grid_sample(mode='bilinear', padding_mode='zeros',
align_corners=False)`` (func.py:56-57); the L*P sampled columns are then weighted by the
attention weights and summed (func.py:60-61).

The arithmetic itself lives in a third-party dependency that is not under /root/reference:
``torch.nn.functional.grid_sample`` (the reference pins pytorch==1.13.0, install.sh:4; this
image has torch 2.11.0 -- bilinear/zeros/align_corners=False semantics are unchanged).  This
port calls the same library routine.  It is what bench.py times as the CPU baseline
(``cpu_baseline.kind = "port"``) because the Python reference cannot travel to the GPU box.
tests/test_oracle.py pins it to golden vectors written by the unmodified reference.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def msda_grid_sample(value: torch.Tensor, spatial_shapes, sampling_locations: torch.Tensor,
                     attention_weights: torch.Tensor) -> torch.Tensor:
    """value [N,S,M,D], spatial_shapes L x (H,W), sampling_locations [N,Lq,M,L,P,2] (x,y in
    [0,1] image-normalised), attention_weights [N,Lq,M,L,P]  ->  [N, Lq, M*D].

    Differentiable through autograd (that is the CPU backward the baseline times)."""
    n, s, m, d = value.shape
    _, lq, _, lv, p, _ = sampling_locations.shape
    sizes = [(int(h), int(w)) for h, w in spatial_shapes]
    assert sum(h * w for h, w in sizes) == s
    # one image batch per (sample, head): [N*M, D, S]
    planes = value.permute(0, 2, 3, 1).reshape(n * m, d, s)
    grid = (sampling_locations * 2 - 1).permute(0, 2, 1, 3, 4, 5).reshape(n * m, lq, lv, p, 2)
    start = 0
    cols = []
    for lvl, (h, w) in enumerate(sizes):
        img = planes[:, :, start:start + h * w].reshape(n * m, d, h, w)
        start += h * w
        cols.append(F.grid_sample(img, grid[:, :, lvl], mode="bilinear",
                                  padding_mode="zeros", align_corners=False))   # [N*M, D, Lq, P]
    sampled = torch.cat(cols, dim=-1)                                             # [N*M, D, Lq, L*P]
    wts = attention_weights.permute(0, 2, 1, 3, 4).reshape(n * m, 1, lq, lv * p)
    out = (sampled * wts).sum(-1)                                                 # [N*M, D, Lq]
    return out.reshape(n, m * d, lq).permute(0, 2, 1).contiguous()
