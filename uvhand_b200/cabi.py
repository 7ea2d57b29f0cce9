"""ctypes binding of the C ABI in include/msda_b200.h (uvhand_b200/lib/libmsda_b200.so).

This is the binding a non-torch host would write (INTEGRATION.md shows it next to the reference
call sites).  Device-pointer entry points take raw addresses (e.g. ``tensor.data_ptr()``) and a
``cudaStream_t`` as an integer; host-buffer entry points take numpy arrays.  Nothing in here
computes anything: a missing library is an error, never a fallback.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# MSDA_B200_LIB lets tools/ time experimental builds of the same ABI; the default is the in-tree build.
LIB_PATH = os.environ.get("MSDA_B200_LIB") or os.path.join(_HERE, "lib", "libmsda_b200.so")

# every symbol include/msda_b200.h declares (tests check that the library exports exactly these)
SYMBOLS = (
    "msda_b200_abi_version", "msda_b200_strerror", "msda_b200_check_im2col_step",
    "msda_fwd_f32", "msda_fwd_f64", "msda_fwd_bf16",
    "msda_bwd_f32", "msda_bwd_f64", "msda_bwd_bf16",
    "msda_fused_fwd_f32", "msda_fused_bwd_f32",
    "msda_value_pack_bf16", "msda_value_unpack_grad_bf16",
    "msda_sample_index_f32",
    "msda_host_plan_create", "msda_host_plan_create_bf16", "msda_host_plan_destroy", "msda_host_plan_traffic",
    "msda_host_register", "msda_host_unregister", "msda_host_fwd_f32", "msda_host_fwd_bwd_f32", "msda_host_fwd_bwd_bf16",
    "msda_host_pipe_create", "msda_host_pipe_create_bf16", "msda_host_pipe_destroy", "msda_host_pipe_submit_f32",
    "msda_host_pipe_submit_bf16", "msda_host_pipe_wait",
    "msda_host_pipe_traffic",
    "msda_b200_set_option", "msda_b200_get_option", "msda_b200_launch_count", "msda_b200_last_path",
)

_lib = None


class MsdaError(RuntimeError):
    def __init__(self, rc: int, what: str):
        self.rc = rc
        super().__init__(f"{what}: {strerror(rc)} (code {rc})")


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is not built; run `python -m uvhand_b200.build`")
        L = ctypes.CDLL(LIB_PATH)
        L.msda_b200_strerror.restype = c_char_p
        L.msda_b200_strerror.argtypes = [c_int]
        L.msda_b200_last_path.restype = c_char_p
        L.msda_b200_last_path.argtypes = [c_int]
        L.msda_b200_launch_count.restype = c_int64
        L.msda_b200_set_option.argtypes = [c_char_p, c_int]
        L.msda_b200_get_option.argtypes = [c_char_p, ctypes.POINTER(c_int)]
        L.msda_b200_check_im2col_step.argtypes = [c_int, c_int]
        seven = [c_int] * 7
        for name in ("msda_fwd_f32", "msda_fwd_f64", "msda_fwd_bf16"):
            getattr(L, name).argtypes = [c_void_p] * 5 + seven + [c_void_p, c_void_p]
        for name in ("msda_bwd_f32", "msda_bwd_f64"):
            getattr(L, name).argtypes = [c_void_p] * 6 + seven + [c_void_p] * 4
        L.msda_bwd_bf16.argtypes = [c_void_p] * 6 + seven + [c_void_p] * 5
        L.msda_fused_fwd_f32.argtypes = [c_void_p] * 6 + seven + [c_void_p, c_void_p]
        L.msda_fused_bwd_f32.argtypes = [c_void_p] * 7 + seven + [c_void_p] * 5
        L.msda_sample_index_f32.argtypes = [c_void_p, c_void_p] + [c_int] * 5 + [c_void_p] * 4
        L.msda_value_pack_bf16.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]
        L.msda_value_unpack_grad_bf16.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]
        L.msda_host_plan_create.argtypes = [ctypes.POINTER(c_void_p), c_int] + seven
        L.msda_host_plan_create_bf16.argtypes = [ctypes.POINTER(c_void_p), c_int] + seven
        L.msda_host_fwd_bwd_bf16.argtypes = [c_void_p] * 11
        L.msda_host_plan_destroy.argtypes = [c_void_p]
        L.msda_host_plan_destroy.restype = None
        L.msda_host_plan_traffic.argtypes = [c_void_p, ctypes.POINTER(c_int64), ctypes.POINTER(c_int64)]
        L.msda_host_pipe_create.argtypes = [ctypes.POINTER(c_void_p), c_int, c_int] + [c_int] * 7
        L.msda_host_pipe_destroy.argtypes = [c_void_p]
        L.msda_host_pipe_destroy.restype = None
        L.msda_host_pipe_create_bf16.argtypes = [ctypes.POINTER(c_void_p), c_int, c_int] + [c_int] * 7
        L.msda_host_pipe_submit_f32.argtypes = [c_void_p, c_int] + [c_void_p] * 10
        L.msda_host_pipe_submit_bf16.argtypes = [c_void_p, c_int] + [c_void_p] * 10
        L.msda_host_pipe_wait.argtypes = [c_void_p, c_int]
        L.msda_host_pipe_traffic.argtypes = [c_void_p, ctypes.POINTER(c_int64), ctypes.POINTER(c_int64)]
        L.msda_host_register.argtypes = [c_void_p, c_int64]
        L.msda_host_unregister.argtypes = [c_void_p]
        L.msda_host_fwd_f32.argtypes = [c_void_p] * 7
        L.msda_host_fwd_bwd_f32.argtypes = [c_void_p] * 11
        _lib = L
    return _lib


def strerror(rc: int) -> str:
    return lib().msda_b200_strerror(int(rc)).decode()


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise MsdaError(rc, what)


def set_option(key: str, value: int) -> None:
    check(lib().msda_b200_set_option(key.encode(), int(value)), f"set_option({key})")


def get_option(key: str) -> int:
    v = c_int(0)
    check(lib().msda_b200_get_option(key.encode(), ctypes.byref(v)), f"get_option({key})")
    return v.value


def launch_count() -> int:
    return int(lib().msda_b200_launch_count())


def last_path(backward: bool = False) -> str:
    return lib().msda_b200_last_path(1 if backward else 0).decode()


_SUFFIX = {"f32": "f32", "f64": "f64", "bf16": "bf16"}


def forward(kind: str, value_ptr, shapes_ptr, lsi_ptr, loc_ptr, aw_ptr, dims, out_ptr, stream=0) -> None:
    """dims = (N, S, M, D, L, Lq, P); pointers are device addresses (ints)."""
    fn = getattr(lib(), "msda_fwd_" + _SUFFIX[kind])
    check(fn(value_ptr, shapes_ptr, lsi_ptr, loc_ptr, aw_ptr, *dims, out_ptr, stream), "msda_fwd_" + kind)


def backward(kind: str, go_ptr, value_ptr, shapes_ptr, lsi_ptr, loc_ptr, aw_ptr, dims,
             gv_ptr, gl_ptr, ga_ptr, stream=0, bf16_scratch_ptr=None) -> None:
    """bf16: `bf16_scratch_ptr` (zeroed fp32 buffer, value-sized) selects the accurate accumulation
    mode; None = direct bf16x2 reductions into a zeroed grad_value."""
    fn = getattr(lib(), "msda_bwd_" + _SUFFIX[kind])
    if kind == "bf16":
        rc = fn(go_ptr, value_ptr, shapes_ptr, lsi_ptr, loc_ptr, aw_ptr, *dims, gv_ptr, gl_ptr, ga_ptr,
                bf16_scratch_ptr, stream)
    else:
        rc = fn(go_ptr, value_ptr, shapes_ptr, lsi_ptr, loc_ptr, aw_ptr, *dims, gv_ptr, gl_ptr, ga_ptr, stream)
    check(rc, "msda_bwd_" + kind)


def fused_forward(value_ptr, shapes_ptr, lsi_ptr, ref_ptr, offsets_ptr, logits_ptr, dims, out_ptr, stream=0) -> None:
    """msda_fused_fwd_f32: softmax + location prologue folded into the operator (fp32, D = 32)."""
    check(lib().msda_fused_fwd_f32(value_ptr, shapes_ptr, lsi_ptr, ref_ptr, offsets_ptr, logits_ptr, *dims, out_ptr,
                                   stream), "msda_fused_fwd_f32")


def fused_backward(go_ptr, value_ptr, shapes_ptr, lsi_ptr, ref_ptr, offsets_ptr, logits_ptr, dims,
                   gv_ptr, goff_ptr, glogit_ptr, gref_ptr=None, stream=0) -> None:
    """msda_fused_bwd_f32; gv (and gref when given) must be zero on entry."""
    check(lib().msda_fused_bwd_f32(go_ptr, value_ptr, shapes_ptr, lsi_ptr, ref_ptr, offsets_ptr, logits_ptr, *dims,
                                   gv_ptr, goff_ptr, glogit_ptr, gref_ptr, stream), "msda_fused_bwd_f32")


def sample_index(shapes_ptr, loc_ptr, N, Lq, M, L, P, h_ptr, w_ptr, mask_ptr, stream=0) -> None:
    check(lib().msda_sample_index_f32(shapes_ptr, loc_ptr, N, Lq, M, L, P, h_ptr, w_ptr, mask_ptr, stream),
          "msda_sample_index_f32")


class HostPlan:
    """Owns device buffers + stream for one problem size; host (numpy) arrays in and out."""

    def __init__(self, N, S, M, D, L, Lq, P, device: int = 0, value_bf16: bool = False):
        """value_bf16: `value` goes up and `grad_value` comes down as bf16 bit patterns (numpy uint16 arrays)."""
        self.dims = (N, S, M, D, L, Lq, P)
        self.value_bf16 = bool(value_bf16)
        self._h = c_void_p()
        create = lib().msda_host_plan_create_bf16 if self.value_bf16 else lib().msda_host_plan_create
        check(create(ctypes.byref(self._h), device, *self.dims), "msda_host_plan_create")

    def close(self):
        if self._h:
            lib().msda_host_plan_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def traffic(self):
        a, b = c_int64(0), c_int64(0)
        check(lib().msda_host_plan_traffic(self._h, ctypes.byref(a), ctypes.byref(b)), "msda_host_plan_traffic")
        return a.value, b.value

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(c_void_p)

    def _sizes(self):
        N, S, M, D, L, Lq, P = self.dims
        n_aw = N * Lq * M * L * P
        return {"value": N * S * M * D, "shapes": 2 * L, "lsi": L, "loc": 2 * n_aw, "aw": n_aw,
                "grad_out": N * Lq * M * D, "out": N * Lq * M * D, "gv": N * S * M * D, "gl": 2 * n_aw, "ga": n_aw}

    def _validate(self, **arrays):
        """The C ABI takes bare pointers: a wrong dtype, a strided view or a short buffer would be read / overrun
        silently, so every array is checked against the plan here."""
        import numpy as np
        sizes = self._sizes()
        for name, a in arrays.items():
            want = np.int64 if name in ("shapes", "lsi") else np.float32
            if getattr(self, "value_bf16", False) and name in ("value", "gv"):
                want = np.uint16
            if not isinstance(a, np.ndarray) or a.dtype != want:
                raise ValueError(f"HostPlan: `{name}` must be a numpy array of {np.dtype(want).name}")
            if not a.flags["C_CONTIGUOUS"]:
                raise ValueError(f"HostPlan: `{name}` must be C-contiguous")
            if a.size != sizes[name]:
                raise ValueError(f"HostPlan: `{name}` has {a.size} elements, the plan needs {sizes[name]}")
            if name in ("out", "gv", "gl", "ga") and not a.flags["WRITEABLE"]:
                raise ValueError(f"HostPlan: `{name}` must be writeable")

    def forward(self, value, shapes, lsi, loc, aw, out):
        self._validate(value=value, shapes=shapes, lsi=lsi, loc=loc, aw=aw, out=out)
        check(lib().msda_host_fwd_f32(self._h, self._p(value), self._p(shapes), self._p(lsi), self._p(loc),
                                      self._p(aw), self._p(out)), "msda_host_fwd_f32")

    def forward_backward(self, value, shapes, lsi, loc, aw, grad_out, out, gv, gl, ga):
        self._validate(value=value, shapes=shapes, lsi=lsi, loc=loc, aw=aw, grad_out=grad_out, out=out, gv=gv, gl=gl,
                       ga=ga)
        fn = lib().msda_host_fwd_bwd_bf16 if self.value_bf16 else lib().msda_host_fwd_bwd_f32
        check(fn(self._h, self._p(value), self._p(shapes), self._p(lsi), self._p(loc),
                                          self._p(aw), self._p(grad_out), self._p(out), self._p(gv),
                                          self._p(gl), self._p(ga)), "msda_host_fwd_bwd_f32")


class HostPipe(HostPlan):
    """`depth` host-buffer steps in flight on one device (msda_host_pipe_*): submit() returns once a step is enqueued,
    wait() once its results are in the caller's arrays.  The arrays of a slot must not be touched between the two."""

    def __init__(self, N, S, M, D, L, Lq, P, device: int = 0, depth: int = 2, value_bf16: bool = False):
        self.dims = (N, S, M, D, L, Lq, P)
        self.value_bf16 = bool(value_bf16)
        self.depth = int(depth)
        self._h = c_void_p()
        self._keep = [None] * self.depth          # the arrays of a step in flight must outlive it
        create = lib().msda_host_pipe_create_bf16 if self.value_bf16 else lib().msda_host_pipe_create
        check(create(ctypes.byref(self._h), device, self.depth, *self.dims), "msda_host_pipe_create")

    def close(self):
        if self._h:
            lib().msda_host_pipe_destroy(self._h)
            self._h = c_void_p()
            self._keep = [None] * self.depth

    def traffic(self):
        a, b = c_int64(0), c_int64(0)
        check(lib().msda_host_pipe_traffic(self._h, ctypes.byref(a), ctypes.byref(b)), "msda_host_pipe_traffic")
        return a.value, b.value

    def submit(self, slot, value, shapes, lsi, loc, aw, grad_out, out, gv, gl, ga):
        if not 0 <= slot < self.depth:
            raise ValueError(f"HostPipe: slot {slot} outside [0, {self.depth})")
        self._validate(value=value, shapes=shapes, lsi=lsi, loc=loc, aw=aw, grad_out=grad_out, out=out, gv=gv, gl=gl,
                       ga=ga)
        fn = lib().msda_host_pipe_submit_bf16 if self.value_bf16 else lib().msda_host_pipe_submit_f32
        check(fn(self._h, slot, self._p(value), self._p(shapes), self._p(lsi), self._p(loc), self._p(aw), self._p(grad_out),
                 self._p(out), self._p(gv), self._p(gl), self._p(ga)), "msda_host_pipe_submit")
        self._keep[slot] = (value, shapes, lsi, loc, aw, grad_out, out, gv, gl, ga)

    def wait(self, slot: int = -1):
        check(lib().msda_host_pipe_wait(self._h, slot), "msda_host_pipe_wait")
        for i in (range(self.depth) if slot < 0 else (slot,)):
            self._keep[i] = None

    def forward(self, *a, **k):
        raise NotImplementedError("HostPipe runs forward + backward steps; use HostPlan.forward")

    def forward_backward(self, *a):
        self.submit(0, *a)
        self.wait(0)


def host_register(arr) -> None:
    if not arr.flags["C_CONTIGUOUS"]:
        raise ValueError("host_register: array must be C-contiguous")
    check(lib().msda_host_register(arr.ctypes.data_as(c_void_p), arr.nbytes), "msda_host_register")


def host_unregister(arr) -> None:
    check(lib().msda_host_unregister(arr.ctypes.data_as(c_void_p)), "msda_host_unregister")
