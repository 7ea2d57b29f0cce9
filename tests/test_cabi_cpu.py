"""CPU tests: the C-ABI library loads, exports every symbol include/msda_b200.h declares, and its
host-side logic (argument validation, the im2col_step precondition, options) behaves -- no
compute call is made without a GPU."""
import os
import re
import subprocess

import pytest
import torch

from uvhand_b200 import build as msda_build
from uvhand_b200 import cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _built():
    msda_build.build_cabi()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "msda_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(msda_[a-z0-9_]+)\s*\(", text))


def test_header_and_binding_agree():
    assert _declared_symbols() == set(cabi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    out = subprocess.run(["nm", "-D", "--defined-only", cabi.LIB_PATH], check=True, capture_output=True,
                         text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if ln.strip()}
    missing = _declared_symbols() - exported
    assert not missing, missing
    L = cabi.lib()
    for name in cabi.SYMBOLS:
        assert getattr(L, name) is not None


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", cabi.LIB_PATH], check=True, capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_abi_version_and_strerror():
    assert cabi.lib().msda_b200_abi_version() == 1
    assert cabi.strerror(0) == "ok"
    assert "im2col_step" in cabi.strerror(-3)
    assert "NULL" in cabi.strerror(-1)


@pytest.mark.parametrize("batch,step,ok", [(2, 64, True), (64, 64, True), (128, 64, True), (2048, 64, True),
                                           (6, 4, False), (65, 64, False), (3, 2, False), (2, 2, True),
                                           (1, 2, True), (0, 64, True), (4, 0, False), (4, -1, False)])
def test_im2col_step_precondition(batch, step, ok):
    # ms_deform_attn_cuda.cu:50-52: im2col_step_ = min(batch, im2col_step); batch % im2col_step_ == 0
    rc = cabi.lib().msda_b200_check_im2col_step(batch, step)
    assert (rc == 0) == ok
    if not ok:
        assert rc == -3


def test_argument_validation_without_gpu():
    L = cabi.lib()
    # negative dimension
    assert L.msda_fwd_f32(None, None, None, None, None, -1, 1, 1, 1, 1, 1, 1, None, None) == -2
    # empty problem: nothing to do, success, nothing launched
    n0 = cabi.launch_count()
    assert L.msda_fwd_f32(None, None, None, None, None, 0, 10, 8, 32, 4, 5, 4, None, None) == 0
    assert L.msda_bwd_f32(None, None, None, None, None, None, 2, 10, 8, 32, 4, 0, 4, None, None, None, None) == 0
    assert cabi.launch_count() == n0
    # null pointers with a non-empty problem
    assert L.msda_fwd_f32(None, None, None, None, None, 1, 10, 8, 32, 4, 5, 4, None, None) == -1
    assert L.msda_bwd_f64(None, None, None, None, None, None, 1, 10, 8, 32, 4, 5, 4, None, None, None, None) == -1
    assert L.msda_sample_index_f32(None, None, 1, 1, 1, 1, 1, None, None, None, None) == -1
    # size overflow
    big = 2 ** 31 - 1
    assert L.msda_fwd_f32(None, None, None, None, None, big, big, 8, 32, 4, 5, 4, None, None) == -2
    # fused prologue entry points: same argument discipline; D != 32 or L*P > 32 is "unsupported", not a fallback
    assert L.msda_fused_fwd_f32(None, None, None, None, None, None, 0, 10, 8, 32, 4, 5, 4, None, None) == 0
    assert L.msda_fused_fwd_f32(None, None, None, None, None, None, 1, 10, 8, 32, 4, 5, 4, None, None) == -1
    assert L.msda_fused_bwd_f32(None, None, None, None, None, None, None, 1, 10, 8, 32, 4, 5, 4,
                                None, None, None, None, None) == -1
    import ctypes
    buf = (ctypes.c_float * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert L.msda_fused_fwd_f32(p, p, p, p, p, p, 1, 10, 8, 30, 4, 5, 4, p, None) == -4      # D = 30
    assert L.msda_fused_fwd_f32(p, p, p, p, p, p, 1, 10, 8, 32, 5, 5, 7, p, None) == -4      # L*P = 35


def test_options_roundtrip():
    for key in ("fwd_path", "bwd_path", "split_mode", "bwd_tiles", "fwd_wide", "fwd_slab"):
        assert cabi.get_option(key) == 0
        cabi.set_option(key, 1)
        assert cabi.get_option(key) == 1
        cabi.set_option(key, 0)
    with pytest.raises(cabi.MsdaError):
        cabi.set_option("fwd_path", 99)
    with pytest.raises(cabi.MsdaError):
        cabi.set_option("no_such_option", 0)


def test_extension_rejects_cpu_tensors_like_the_reference():
    # ms_deform_attn.h:37,60: AT_ERROR("Not implemented on the CPU")
    msda_build.build_torch_ext()
    import uvhand_b200
    M = uvhand_b200.load_extension()
    shapes = torch.tensor([[2, 3]], dtype=torch.long)
    lsi = torch.zeros(1, dtype=torch.long)
    value = torch.zeros(1, 6, 1, 4)
    loc = torch.zeros(1, 2, 1, 1, 1, 2)
    aw = torch.zeros(1, 2, 1, 1, 1)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        M.ms_deform_attn_forward(value, shapes, lsi, loc, aw, 64)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        M.ms_deform_attn_backward(value, shapes, lsi, loc, aw, torch.zeros(1, 2, 4), 64)


def test_reference_import_names_resolve():
    msda_build.build_torch_ext()
    import importlib
    import uvhand_b200
    uvhand_b200.load_extension()
    mod = importlib.import_module("MultiScaleDeformableAttention")   # functions/ms_deform_attn_func.py:18
    assert callable(mod.ms_deform_attn_forward) and callable(mod.ms_deform_attn_backward)
    from uvhand_b200.functions import MSDeformAttnFunction
    from uvhand_b200.modules import MSDeformAttn
    m = MSDeformAttn(d_model=256, n_levels=4, n_heads=8, n_points=4)
    assert set(dict(m.named_parameters())) == {
        "sampling_offsets.weight", "sampling_offsets.bias", "attention_weights.weight",
        "attention_weights.bias", "value_proj.weight", "value_proj.bias", "output_proj.weight",
        "output_proj.bias"}
    assert m.im2col_step == 64
    assert MSDeformAttnFunction.apply is not None


def test_module_sampling_locations_match_reference_formulas():
    from uvhand_b200.modules import MSDeformAttn
    torch.manual_seed(0)
    N, Lq, M, L, P = 2, 3, 4, 2, 3
    shapes = torch.tensor([[5, 7], [3, 4]])
    off = torch.randn(N, Lq, M, L, P, 2)
    norm = torch.stack([shapes[..., 1], shapes[..., 0]], -1)
    ref2 = torch.rand(N, Lq, L, 2)
    want = ref2[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    assert torch.equal(MSDeformAttn._sampling_locations(ref2, off, shapes), want)
    ref42 = torch.rand(N, Lq, L, 42)
    cx = ref42[:, :, None, :, None, 0::2].mean(-1).unsqueeze(-1)
    cy = ref42[:, :, None, :, None, 1::2].mean(-1).unsqueeze(-1)
    want = torch.cat([cx, cy], -1) + off / norm[None, None, None, :, None, :]
    assert torch.allclose(MSDeformAttn._sampling_locations(ref42, off, shapes), want, atol=0, rtol=0)
    ref4 = torch.rand(N, Lq, L, 4)
    want = ref4[:, :, None, :, None, :2] + off / P * ref4[:, :, None, :, None, 2:] * 0.5
    assert torch.equal(MSDeformAttn._sampling_locations(ref4, off, shapes), want)
    with pytest.raises(ValueError):
        MSDeformAttn._sampling_locations(torch.rand(N, Lq, L, 3), off, shapes)


def test_module_init_matches_reference_ring():
    from uvhand_b200.modules import MSDeformAttn
    m = MSDeformAttn(256, 4, 8, 4)
    b = m.sampling_offsets.bias.detach().view(8, 4, 4, 2)
    # head 0 looks along +x, point i sits i+1 pixels out, identical on every level
    assert torch.allclose(b[0, :, :, 0], torch.tensor([1., 2., 3., 4.]).expand(4, 4))
    assert torch.allclose(b[0, :, :, 1], torch.zeros(4, 4), atol=1e-6)
    assert torch.allclose(b[2, 0, :, 1], torch.tensor([1., 2., 3., 4.]))
    assert float(m.sampling_offsets.weight.abs().max()) == 0.0
    assert float(m.attention_weights.weight.abs().max()) == 0.0 and float(m.attention_weights.bias.abs().max()) == 0.0
    assert float(m.value_proj.bias.abs().max()) == 0.0


def test_host_plan_argument_validation():
    """HostPlan hands bare pointers to the C ABI: dtype, contiguity and size are checked on the Python side
    (no device needed: the plan object is built without its native handle)."""
    import numpy as np
    hp = cabi.HostPlan.__new__(cabi.HostPlan)
    hp.dims = (2, 30, 2, 4, 2, 3, 2)       # N, S, M, D, L, Lq, P
    hp._h = None
    sizes = hp._sizes()
    good = {"value": np.zeros(sizes["value"], np.float32), "shapes": np.zeros((2, 2), np.int64),
            "lsi": np.zeros(2, np.int64), "loc": np.zeros(sizes["loc"], np.float32),
            "aw": np.zeros(sizes["aw"], np.float32), "out": np.zeros(sizes["out"], np.float32)}
    hp._validate(**good)
    for name, bad in (("value", good["value"].astype(np.float64)), ("shapes", good["shapes"].astype(np.int32)),
                      ("loc", np.zeros(2 * sizes["loc"], np.float32)[::2]), ("out", np.zeros(sizes["out"] - 1, np.float32))):
        with pytest.raises(ValueError, match=name):
            hp._validate(**{**good, name: bad})
    ro = np.zeros(sizes["out"], np.float32)
    ro.flags.writeable = False
    with pytest.raises(ValueError, match="writeable"):
        hp._validate(**{**good, "out": ro})
