#!/usr/bin/env python
"""The reference's own operator test, /root/reference/models/ops/test.py, re-stated against this package:
same shapes (N=1, M=2, D=2, Lq=2, L=2, P=2, levels 6x4 and 3x2, test.py:21-28), same seed and CPU
`torch.rand` inputs (:28-36), same three checks with the same print lines (:31-78) and the same channel sweep
(:85-86).  The reference compares its CUDA extension with ms_deform_attn_core_pytorch; here the comparison
partner is the port of that function under oracle/ (checker only).  Needs a B200.

    python tools/reference_test.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.autograd import gradcheck  # noqa: E402

from oracle.grid_sample_port import msda_grid_sample  # noqa: E402
from uvhand_b200.functions import MSDeformAttnFunction  # noqa: E402

N, M, D = 1, 2, 2
Lq, L, P = 2, 2, 2
shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long).cuda()
level_start_index = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
S = sum([(H * W).item() for H, W in shapes])
levels = [tuple(int(v) for v in hw) for hw in shapes.cpu()]
torch.manual_seed(3)


def inputs(channels=D):
    value = torch.rand(N, S, M, channels).cuda() * 0.01
    sampling_locations = torch.rand(N, Lq, M, L, P, 2).cuda()
    attention_weights = torch.rand(N, Lq, M, L, P).cuda() + 1e-5
    attention_weights /= attention_weights.sum(-1, keepdim=True).sum(-2, keepdim=True)
    return value, sampling_locations, attention_weights


@torch.no_grad()
def check_forward_equal_with_pytorch(dtype, name):
    value, loc, w = (t.to(dtype) for t in inputs())
    out_ref = msda_grid_sample(value.cpu(), levels, loc.cpu(), w.cpu()).to(dtype)
    out = MSDeformAttnFunction.apply(value, shapes, level_start_index, loc, w, 2).cpu()
    if dtype == torch.float64:
        fwdok = torch.allclose(out, out_ref)
    else:
        fwdok = torch.allclose(out, out_ref, rtol=1e-2, atol=1e-3)
    max_abs_err = (out - out_ref).abs().max()
    max_rel_err = ((out - out_ref).abs() / out_ref.abs()).max()
    print(f'* {fwdok} check_forward_equal_with_pytorch_{name}: max_abs_err {max_abs_err:.2e} max_rel_err {max_rel_err:.2e}')
    return fwdok


def check_gradient_numerical(channels=4, grad_value=True, grad_sampling_loc=True, grad_attn_weight=True):
    value, loc, w = (t.double() for t in inputs(channels))
    value.requires_grad = grad_value
    loc.requires_grad = grad_sampling_loc
    w.requires_grad = grad_attn_weight
    gradok = gradcheck(MSDeformAttnFunction.apply, (value, shapes, level_start_index, loc, w, 2))
    print(f'* {gradok} check_gradient_numerical(D={channels})')
    return gradok


if __name__ == '__main__':
    ok = check_forward_equal_with_pytorch(torch.float64, "double")
    ok &= check_forward_equal_with_pytorch(torch.float32, "float")
    for channels in [30, 32, 64, 71, 1025, 2048, 3096]:
        ok &= check_gradient_numerical(channels, True, True, True)
    sys.exit(0 if ok else 1)
