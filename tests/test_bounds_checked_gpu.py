"""The parity suite once more through the CHECKED build of the library (-DMSDA_BOUNDS_CHECK=1,
uvhand_b200/lib/libmsda_b200_checked.so): inside the kernels every gathered value row and every
grad_value reduction -- the accesses whose addresses depend on the sampling locations -- is compared with
the extent of its tensor, and a violation traps.  This is the project's own memcheck (compute-sanitizer is
not available on the GPU pool).  Subprocesses, because the library is chosen at load time and a trap
poisons the CUDA context."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECKED = os.path.join(ROOT, "uvhand_b200", "lib", "libmsda_b200_checked.so")


def _env():
    if not os.path.exists(CHECKED):
        pytest.fail(f"{CHECKED} is not built (python -m uvhand_b200.build)")
    return dict(os.environ, MSDA_B200_LIB=CHECKED)


def test_parity_suite_is_clean_under_the_bounds_checked_kernels():
    k = ("golden_f32 or golden_f64 or random_geometry_fuzz or shape_sweep or special_locations or invalid_corners "
         "or touch_pattern or mid_size or fused_prologue_matches or fused_prologue_fuzz or fused_locations")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_msda_gpu.py"), "-m", "gpu",
                        "-x", "-q", "-k", k], env=_env(), capture_output=True, text=True, timeout=1500, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "passed" in r.stdout and "bounds check" not in r.stdout + r.stderr


PROBE = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r)
from uvhand_b200 import cabi
assert cabi.LIB_PATH.endswith("_checked.so")
N, S, M, D, L, Lq, P = 1, 30, 8, 32, 2, 16, 4
shapes = torch.tensor([[6, 4], [3, 2]], dtype=torch.long, device="cuda")
lsi = torch.tensor([0, %d], dtype=torch.long, device="cuda")       # level 1 starts at / beyond the end of value
value = torch.randn(N, S, M, D, device="cuda")
loc = torch.rand(N, Lq, M, L, P, 2, device="cuda")
aw = torch.rand(N, Lq, M, L, P, device="cuda")
out = torch.empty(N, Lq, M * D, device="cuda")
cabi.set_option("fwd_path", %d)
cabi.forward("f32", value.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), loc.data_ptr(), aw.data_ptr(),
             (N, S, M, D, L, Lq, P), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("finished")
"""


@pytest.mark.parametrize("path", [0, 1], ids=["d32", "anyshape"])
def test_the_checker_catches_a_gather_past_the_end_of_value(path):
    """Negative control: a level_start_index that points past the tensor makes the kernels gather out of
    range -- silently in the normal build, loudly in the checked one."""
    good = subprocess.run([sys.executable, "-c", PROBE % (ROOT, 24, path)], env=_env(), capture_output=True, text=True,
                          timeout=300)
    assert good.returncode == 0 and "finished" in good.stdout, good.stderr[-2000:]
    bad = subprocess.run([sys.executable, "-c", PROBE % (ROOT, 4000, path)], env=_env(), capture_output=True, text=True,
                         timeout=300)
    assert bad.returncode != 0 and "finished" not in bad.stdout
    assert "msda bounds check: value access" in bad.stdout + bad.stderr
