"""Synthetic workloads of the MSDeformAttn path: the shapes BASELINE.json names and the input
distributions SURVEY.md section 8(d) defines.  Host-side helpers for tests and bench.py only --
nothing here is on the operator's execution path.

Shapes (M=8 heads, D=32 channels/head, L=4 levels, P=4 points unless stated):
  C1  models/ops/test.py:21-28        N=1|2, M=2, D=2, Lq=2, L=2, P=2, levels (6,4),(3,2)
  C2  Deformable-DETR encoder         N=2, levels of an 800x1333 image at strides 8..64, Lq=S=22223
  C3  H2OTR/ARCTIC (224^2, window 32) N=32, levels 28^2,14^2,7^2,4^2 (S=1045); encoder Lq=S, decoder Lq=300
  C4  DINO 4-scale                    encoder = C2; decoder Lq=900 (eval) / 1098 (train with DN queries)
  C5  high-res stress                 N=8, levels (256,264),(128,132),(64,66),(32,33), Lq=S=89760, bf16 value
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Tuple

import torch


@dataclass(frozen=True)
class Workload:
    name: str
    N: int
    levels: Tuple[Tuple[int, int], ...]
    Lq: int
    M: int = 8
    D: int = 32
    P: int = 4
    loc_dist: str = "encoder"
    note: str = ""

    @property
    def L(self) -> int:
        return len(self.levels)

    @property
    def S(self) -> int:
        return sum(h * w for h, w in self.levels)

    @property
    def items(self) -> int:
        """queries x heads: the unit of the throughput metric."""
        return self.N * self.Lq * self.M

    def dims(self):
        return (self.N, self.S, self.M, self.D, self.L, self.Lq, self.P)

    def bytes_fwd(self, value_bytes: int = 4) -> int:
        """Algorithmic bytes, forward: |value| + |loc| + |w| + |out|."""
        n_value = self.N * self.S * self.M * self.D
        n_w = self.N * self.Lq * self.M * self.L * self.P
        n_out = self.N * self.Lq * self.M * self.D
        return n_value * value_bytes + n_w * 2 * 4 + n_w * 4 + n_out * 4

    def bytes_bwd(self, value_bytes: int = 4) -> int:
        """|grad_out| + |value| + |loc| + |w| + |grad_value| + |grad_loc| + |grad_w|."""
        n_value = self.N * self.S * self.M * self.D
        n_w = self.N * self.Lq * self.M * self.L * self.P
        n_out = self.N * self.Lq * self.M * self.D
        return n_out * 4 + 2 * n_value * value_bytes + 2 * (n_w * 2 * 4 + n_w * 4)


def _pyramid(h: int, w: int, strides) -> Tuple[Tuple[int, int], ...]:
    return tuple((math.ceil(h / s), math.ceil(w / s)) for s in strides)


DETR_LEVELS = _pyramid(800, 1333, (8, 16, 32, 64))        # (100,167),(50,84),(25,42),(13,21)
ARCTIC_LEVELS = ((28, 28), (14, 14), (7, 7), (4, 4))
HIRES_LEVELS = ((256, 264), (128, 132), (64, 66), (32, 33))


def _sum_hw(levels) -> int:
    return sum(h * w for h, w in levels)


WORKLOADS = {
    "c1_testpy": Workload("c1_testpy", 2, ((6, 4), (3, 2)), 2, M=2, D=2, P=2, loc_dist="uniform",
                          note="BASELINE configs[0]"),
    "c2_detr_encoder": Workload("c2_detr_encoder", 2, DETR_LEVELS, _sum_hw(DETR_LEVELS),
                                note="BASELINE configs[1]: N=2, M=8, D=32, L=4, P=4, 800x1333"),
    "c3_arctic_encoder": Workload("c3_arctic_encoder", 32, ARCTIC_LEVELS, _sum_hw(ARCTIC_LEVELS),
                                  note="BASELINE configs[2], per-GPU window of 32 frames"),
    "c3_arctic_decoder": Workload("c3_arctic_decoder", 32, ARCTIC_LEVELS, 300, loc_dist="twostage"),
    "c4_dino_decoder": Workload("c4_dino_decoder", 2, DETR_LEVELS, 900, loc_dist="twostage",
                                note="BASELINE configs[3], eval queries"),
    "c4_dino_decoder_train": Workload("c4_dino_decoder_train", 2, DETR_LEVELS, 1098, loc_dist="twostage"),
    "c5_hires": Workload("c5_hires", 8, HIRES_LEVELS, _sum_hw(HIRES_LEVELS),
                         note="BASELINE configs[4]: bf16 value"),
    "c5_hires_contended": Workload("c5_hires_contended", 8, HIRES_LEVELS, _sum_hw(HIRES_LEVELS),
                                   loc_dist="contended"),
}


def level_tensors(levels, device="cpu"):
    shapes = torch.as_tensor(list(levels), dtype=torch.long, device=device).reshape(-1, 2)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    return shapes, lsi


def encoder_reference_points(levels, device="cpu") -> torch.Tensor:
    """Pixel centres of every level, (x, y) normalised -- what the encoder uses as references
    (arctic_transformer.py:310-323 with valid_ratios = 1).  [S, 2]"""
    pts = []
    for h, w in levels:
        ys = (torch.arange(h, device=device, dtype=torch.float32) + 0.5) / h
        xs = (torch.arange(w, device=device, dtype=torch.float32) + 0.5) / w
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        pts.append(torch.stack([gx.reshape(-1), gy.reshape(-1)], -1))
    return torch.cat(pts, 0)


def make_inputs(wl: Workload, seed: int = 0, device="cpu", value_dtype=torch.float32,
                loc_dist: str | None = None, pad_frac: float = 0.1, offset_px: float = 2.0):
    """Seeded synthetic inputs on `device` (the generator lives on that device).

    value ~ N(0,1) with the trailing `pad_frac` of every level zeroed (padding mask,
    ms_deform_attn.py:97-98); w = softmax(N(0,1)) over L*P; grad_out ~ N(0,1); locations:
      uniform    U[0,1)                                                    (test.py:34)
      encoder    pixel-centre references + N(0, offset_px^2) pixel offsets  (needs Lq == S)
      twostage   references U[-1,1] (arctic_transformer.py:230) + the same offsets
      contended  references N(0.5, 0.02^2) + offsets: atomics pile onto few rows
    """
    dist = loc_dist or wl.loc_dist
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    N, S, M, D, L, Lq, P = wl.dims()
    shapes, lsi = level_tensors(wl.levels, device)
    value = torch.randn(N, S, M, D, generator=g, device=device, dtype=torch.float32)
    if pad_frac > 0:
        start = 0
        for h, w in wl.levels:
            n_pad = int(h * w * pad_frac)
            if n_pad:
                value[:, start + h * w - n_pad:start + h * w] = 0
            start += h * w
    wh = torch.stack([shapes[:, 1], shapes[:, 0]], -1).to(torch.float32)      # (W, H)
    offs = torch.randn(N, Lq, M, L, P, 2, generator=g, device=device) * offset_px
    if dist == "uniform":
        loc = torch.rand(N, Lq, M, L, P, 2, generator=g, device=device)
    elif dist == "encoder":
        if Lq != S:
            raise ValueError("encoder locations need Lq == S")
        ref = encoder_reference_points(wl.levels, device)
        loc = ref[None, :, None, None, None, :] + offs / wh[None, None, None, :, None, :]
    elif dist == "twostage":
        ref = torch.rand(N, Lq, 1, 1, 1, 2, generator=g, device=device) * 2 - 1
        loc = ref + offs / wh[None, None, None, :, None, :]
    elif dist == "contended":
        ref = 0.5 + 0.02 * torch.randn(N, Lq, 1, 1, 1, 2, generator=g, device=device)
        loc = ref + offs / wh[None, None, None, :, None, :]
    else:
        raise ValueError(dist)
    aw = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g, device=device), -1).view(N, Lq, M, L, P)
    grad_out = torch.randn(N, Lq, M * D, generator=g, device=device)
    return {
        "value": value.to(value_dtype).contiguous(), "shapes": shapes, "lsi": lsi,
        "loc": loc.contiguous(), "aw": aw.contiguous(), "grad_out": grad_out.contiguous(),
    }
