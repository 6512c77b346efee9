"""ctypes binding of libcmbl_b200.so (C ABI: include/cmbl_b200.h).

The product library is `libcmbl_b200.so` next to this file, built from csrc/*.cu for sm_100a.  There is NO CPU fallback:
if the library is missing or CUDA is unavailable, loading raises.  (The unit tests of the kernel index logic load a host
emulator build of the same sources explicitly by path — `load(path=...)` — that build is never picked up implicitly.)
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_PATH = os.environ.get("CMBL_B200_LIB", os.path.join(_HERE, "libcmbl_b200.so"))      # override: kernel-variant experiments

MAP, FOURIER = 0, 1
OP_L, OP_LH, OP_LINV, OP_LHINV = 0, 1, 2, 3


class CmblError(RuntimeError):
    """Raised for any non-zero status of the C ABI (mirrors the reference's `error(...)` exceptions)."""


class DatasetDesc(ctypes.Structure):
    _fields_ = [("Npol", c_int), ("Nb", c_int), ("Cf", c_void_p), ("Cn", c_void_p), ("Cnhat", c_void_p),
                ("B", c_void_p), ("Bhat", c_void_p), ("Mf", c_void_p), ("mask_pix", c_void_p), ("d", c_void_p)]


# name -> (restype, argtypes); every symbol include/cmbl_b200.h declares
SIGNATURES = {
    "cmbl_last_error": (c_char_p, []),
    "cmbl_version": (c_char_p, []),
    "cmbl_launch_count": (c_longlong, []),
    "cmbl_profile_begin": (c_int, []),
    "cmbl_profile_end": (c_char_p, []),
    "cmbl_plan_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_double, c_int]),
    "cmbl_plan_destroy": (c_int, [c_void_p]),
    "cmbl_plan_grids": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_double)]),
    "cmbl_cl_to_cov": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_double, c_void_p, c_void_p]),
    "cmbl_rfft2": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "cmbl_irfft2": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "cmbl_diag_mul": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "cmbl_field_axpby": (c_int, [c_void_p, c_int, POINTER(c_double), c_int, c_void_p, POINTER(c_double), c_int, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "cmbl_qu_eb": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "cmbl_blockdiag_ieb": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "cmbl_dot": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, POINTER(c_double), c_void_p]),
    "cmbl_lenseflow_create": (c_int, [POINTER(c_void_p), c_void_p, c_int, c_int, c_int, c_int]),
    "cmbl_lenseflow_destroy": (c_int, [c_void_p]),
    "cmbl_lenseflow_precompute": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "cmbl_lenseflow_apply": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "cmbl_lenseflow_apply_host": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "cmbl_lenseflow_apply_host_async": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "cmbl_lenseflow_host_sync": (c_int, []),
    "cmbl_lenseflow_grad": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "cmbl_max_lensing_step": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, POINTER(c_double), c_void_p]),
    "cmbl_lenseflow_kernel_path": (c_int, [c_void_p]),
    "cmbl_lenseflow_get_p": (c_int, [c_void_p, c_int, c_void_p]),
    "cmbl_cg_create": (c_int, [POINTER(c_void_p), c_void_p, POINTER(DatasetDesc), c_void_p]),
    "cmbl_cg_destroy": (c_int, [c_void_p]),
    "cmbl_cg_begin": (c_int, [c_void_p, c_void_p, c_int, POINTER(c_double), c_void_p]),
    "cmbl_cg_step": (c_int, [c_void_p, POINTER(c_double), c_void_p]),
    "cmbl_cg_mark_best": (c_int, [c_void_p, c_void_p]),
    "cmbl_cg_result": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "cmbl_wiener_cg": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_double, c_int, POINTER(c_int), POINTER(c_double), c_void_p]),
    "cmbl_gradientf_logpdf": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "cmbl_comm_unique_id": (c_int, [c_void_p]),
    "cmbl_comm_init": (c_int, [POINTER(c_void_p), c_int, c_int, c_void_p]),
    "cmbl_comm_destroy": (c_int, [c_void_p]),
    "cmbl_comm_allreduce": (c_int, [c_void_p, POINTER(c_double), c_int, c_int, c_void_p]),
    "cmbl_wiener_cg_sharded": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_double, c_int, POINTER(c_int), POINTER(c_double), c_void_p]),
}


class Library:
    """Loaded C ABI with checked calls: `lib.call("cmbl_rfft2", ...)` raises CmblError on failure."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise CmblError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). cmblensing.jl_b200 has no CPU fallback.")
        self.path = path
        self.cdll = ctypes.CDLL(path)
        missing = []
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(self.cdll, name)
            except AttributeError:
                missing.append(name)
                continue
            fn.restype, fn.argtypes = res, args
        if missing and not os.environ.get("CMBL_B200_ALLOW_MISSING"):      # (the override exists for A/B timing against older builds)
            raise CmblError(f"{path} does not export: {', '.join(missing)}")
        self.is_emulator = b"emulator" in self.cdll.cmbl_version()

    def call(self, name: str, *args):
        rc = getattr(self.cdll, name)(*args)
        if rc != 0:
            raise CmblError(f"{name}: {self.cdll.cmbl_last_error().decode()} (status {rc})")

    def launch_count(self) -> int:
        return int(self.cdll.cmbl_launch_count())


_default = None


def load(path: str | None = None) -> Library:
    """Load (once) the product library; an explicit `path` loads that file instead (tests only)."""
    global _default
    if path is not None:
        return Library(path)
    if _default is None:
        _default = Library(DEFAULT_PATH)
    return _default
