"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference; the GPU box has none):

    python oracle/make_golden.py

It imports the reference's own ``ms_deform_attn_core_pytorch``
(/root/reference/models/ops/functions/ms_deform_attn_func.py:42-62) -- the function the
reference's test (models/ops/test.py:31-60) scores its CUDA kernels against -- runs it in
float64 on seeded inputs, takes its autograd gradients, and stores inputs + outputs.  The
reference module imports its CUDA extension at top level (func.py:18), which does not exist
here, so an empty stub module of that name is registered first; the function used never
touches it.

Input recipe for the ``testpy_*`` cases is the reference test's own
(test.py:21-28,33-36: ``torch.manual_seed(3)``, CPU ``torch.rand``), so the vectors are the
ones that test would have produced.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF_OPS = "/root/reference/models/ops"
OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_reference():
    if not os.path.isdir(REF_OPS):
        raise SystemExit("make_golden.py needs /root/reference (build container only)")
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    sys.path.insert(0, REF_OPS)
    from functions.ms_deform_attn_func import ms_deform_attn_core_pytorch  # noqa: E402
    return ms_deform_attn_core_pytorch


def level_start_index(shapes: torch.Tensor) -> torch.Tensor:
    # test.py:24
    return torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))


def run_case(ref_fn, name, value, shapes, loc, aw, grad_out, note):
    """value/loc/aw/grad_out are float32 tensors; the reference runs on their float64 images."""
    v = value.double().requires_grad_(True)
    lo = loc.double().requires_grad_(True)
    a = aw.double().requires_grad_(True)
    out = ref_fn(v, shapes, lo, a)
    out.backward(grad_out.double())
    with torch.no_grad():
        out32 = ref_fn(value, shapes, loc, aw)
    np.savez_compressed(
        os.path.join(OUT_DIR, name + ".npz"),
        value=value.numpy(), shapes=shapes.numpy(), lsi=level_start_index(shapes).numpy(),
        loc=loc.numpy(), aw=aw.numpy(), grad_out=grad_out.numpy(),
        out=out.detach().numpy(), grad_value=v.grad.numpy(), grad_loc=lo.grad.numpy(),
        grad_aw=a.grad.numpy(), out_ref_fp32=out32.numpy(), note=np.array(note))
    print(f"{name}: out {tuple(out.shape)} |out|max {out.abs().max():.4e}")


def testpy_inputs(N, M, D, Lq, L, P, shapes):
    S = int(shapes.prod(1).sum())
    value = torch.rand(N, S, M, D) * 0.01
    loc = torch.rand(N, Lq, M, L, P, 2)
    aw = torch.rand(N, Lq, M, L, P) + 1e-5
    aw /= aw.sum(-1, keepdim=True).sum(-2, keepdim=True)
    return value, loc, aw


def main():
    os.makedirs(OUT_DIR, exist_ok=True)
    ref = load_reference()
    shapes_t = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long)

    # C1 as in the file: N=1, M=2, D=2, Lq=2, L=2, P=2, seed 3, first draw after the seed.
    torch.manual_seed(3)
    value, loc, aw = testpy_inputs(1, 2, 2, 2, 2, 2, shapes_t)
    g = torch.randn(1, 2, 4)
    run_case(ref, "testpy_n1", value, shapes_t, loc, aw, g,
             "models/ops/test.py:21-36 shape and recipe, seed 3, N=1")

    # BASELINE.json configs[0]: same shape with N=2.
    torch.manual_seed(3)
    value, loc, aw = testpy_inputs(2, 2, 2, 2, 2, 2, shapes_t)
    g = torch.randn(2, 2, 4)
    run_case(ref, "testpy_n2", value, shapes_t, loc, aw, g,
             "BASELINE.json configs[0]: test.py recipe with N=2, seed 3")

    # test.py:85-86 channel sweep, the two small odd ones (the big ones are covered by
    # properties on the GPU, not by fixtures).
    for D in (30, 71):
        torch.manual_seed(3)
        value, loc, aw = testpy_inputs(1, 2, D, 2, 2, 2, shapes_t)
        g = torch.randn(1, 2, 2 * D)
        run_case(ref, f"testpy_d{D}", value, shapes_t, loc, aw, g,
                 f"test.py:63-78 check_gradient_numerical shape, channels={D}, seed 3")

    # Out-of-range locations (two-stage references are in [-1,1], arctic_transformer.py:230),
    # non-square levels, a 1x2 level.
    torch.manual_seed(11)
    sh = torch.as_tensor([(6, 4), (3, 5), (1, 2)], dtype=torch.long)
    N, M, D, Lq, L, P = 2, 2, 4, 5, 3, 2
    S = int(sh.prod(1).sum())
    value = torch.randn(N, S, M, D)
    loc = torch.rand(N, Lq, M, L, P, 2) * 2.2 - 0.6
    aw = torch.softmax(torch.randn(N, Lq, M, L * P), -1).view(N, Lq, M, L, P)
    g = torch.randn(N, Lq, M * D)
    run_case(ref, "outside_small", value, sh, loc, aw, g,
             "loc ~ U[-0.6,1.6): zero padding and the coarse skip; levels 6x4,3x5,1x2")

    # The shipped head geometry: M=8, D=32, L=4, P=4 (util/settings.py:102-120), small levels.
    torch.manual_seed(12)
    sh = torch.as_tensor([(8, 11), (4, 6), (2, 3), (1, 2)], dtype=torch.long)
    N, M, D, Lq, L, P = 1, 8, 32, 23, 4, 4
    S = int(sh.prod(1).sum())
    value = torch.randn(N, S, M, D)
    loc = torch.rand(N, Lq, M, L, P, 2) * 1.2 - 0.1
    aw = torch.softmax(torch.randn(N, Lq, M, L * P), -1).view(N, Lq, M, L, P)
    g = torch.randn(N, Lq, M * D)
    run_case(ref, "d32_m8_l4_p4", value, sh, loc, aw, g,
             "M=8,D=32,L=4,P=4 (shipped head geometry), loc ~ U[-0.1,1.1)")

    # Encoder-like: queries are the pixel centres of every level, offsets of a few pixels
    # (modules/ms_deform_attn.py:62-70,105-108; arctic_transformer.py:310-323).
    torch.manual_seed(13)
    sh = torch.as_tensor([(5, 7), (3, 4)], dtype=torch.long)
    N, M, D, L, P = 2, 4, 8, 2, 3
    S = int(sh.prod(1).sum())
    refs = []
    for (h, w) in sh.tolist():
        ys, xs = torch.meshgrid((torch.arange(h) + 0.5) / h, (torch.arange(w) + 0.5) / w, indexing="ij")
        refs.append(torch.stack([xs.reshape(-1), ys.reshape(-1)], -1))
    refp = torch.cat(refs, 0)                                         # [S,2] (x,y)
    Lq = S
    norm = torch.stack([sh[:, 1], sh[:, 0]], -1).float()               # (W,H)
    off = torch.randn(N, Lq, M, L, P, 2) * 1.5 + 0.37                  # keep off exact half-pixels
    loc = refp[None, :, None, None, None, :] + off / norm[None, None, None, :, None, :]
    value = torch.randn(N, S, M, D)
    aw = torch.softmax(torch.randn(N, Lq, M, L * P), -1).view(N, Lq, M, L, P)
    g = torch.randn(N, Lq, M * D)
    run_case(ref, "encoder_like", value, sh, loc.contiguous(), aw, g,
             "Lq=S, refs = pixel centres, offsets ~ N(0.37, 1.5 px)")


if __name__ == "__main__":
    main()
