"""ctypes binding of libmsr3d_b200.so (the C ABI in include/msr3d_b200.h).

There is NO fallback: if the shared library is missing or a call returns a
non-zero status this module raises.  Nothing here touches ``oracle/``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MSR3D_LIB") or os.path.join(_HERE, "libmsr3d_b200.so")   # MSR3D_LIB: development builds (make trace)

c_int = ctypes.c_int
c_float = ctypes.c_float
c_void_p = ctypes.c_void_p
c_ll = ctypes.c_longlong


class Msr3dError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a with the in-tree Makefile."""
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "all"],
                       capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise Msr3dError("building libmsr3d_b200.so failed")
    return LIB_PATH


_lib = None

# name -> argtypes (restype is always int status)
_P = c_void_p
_PROTOS = {
    "msr3d_furthest_point_sampling_f32": [_P, _P, c_int, c_int, c_int, _P],
    "msr3d_gather_points_f32": [_P, _P, _P, c_int, c_int, c_int, c_int, _P],
    "msr3d_gather_points_grad_f32": [_P, _P, _P, c_int, c_int, c_int, c_int, _P],
    "msr3d_ball_query_f32": [_P, _P, _P, c_int, c_int, c_int, c_float, c_int, _P],
    "msr3d_group_points_f32": [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P],
    "msr3d_group_points_grad_f32": [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P],
    "msr3d_three_nn_f32": [_P, _P, _P, _P, c_int, c_int, c_int, _P],
    "msr3d_three_interpolate_f32": [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P],
    "msr3d_three_interpolate_grad_f32": [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P],
    # group B
    "msr3d_pack_weight_f32": [_P, c_int, c_int, c_ll, c_ll, _P, _P, c_int, c_int, _P, _P],
    "msr3d_linear_bf16x3_f32": [_P, _P, _P, c_ll, c_int, c_int, _P, _P, _P, _P, _P, c_ll, c_int, c_int, c_int, _P],
    "msr3d_linear_tc_f32": [_P, c_ll, c_ll, c_int, c_int, _P, _P, _P, c_ll, c_int, c_int, c_int,
                            c_int, c_int, c_int, _P],
    "msr3d_sa_group_f32": [_P, c_int, c_int, c_int, c_int, c_float, c_int, _P, _P, _P, _P],
    "msr3d_scatter_xyz_cols_f32": [_P, _P, c_ll, c_int, c_int, _P],
    "msr3d_sa_mlp_fused_f32": [_P, c_int, c_int, _P, c_int, _P, _P, c_int, c_int, c_int, c_int,
                               _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P, c_ll, c_int, _P],
}


def register_protos(protos: dict) -> None:
    """Other modules of this package add the prototypes of their entry points."""
    _PROTOS.update(protos)
    if _lib is not None:
        _bind(_lib, protos)


def _bind(lib, protos) -> None:
    for name, argtypes in protos.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale: fail loudly
        fn.argtypes = argtypes
        fn.restype = c_int


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Msr3dError(
                f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C msr3d_b200/csrc`). msr3d_b200 has no CPU or PyTorch fallback.")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.msr3d_last_error.restype = ctypes.c_char_p
        _lib.msr3d_abi_version.restype = c_int
        _lib.msr3d_launch_count.restype = ctypes.c_ulonglong
        _lib.msr3d_packed_weight_bytes.restype = ctypes.c_longlong
        _lib.msr3d_packed_weight_bytes.argtypes = [c_int, c_int, c_int, c_int]
        _bind(_lib, _PROTOS)
    return _lib


def call(name: str, *args) -> None:
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise Msr3dError(f"{name} failed (status {rc}): {lib().msr3d_last_error().decode()}")


def launch_count() -> int:
    return int(lib().msr3d_launch_count())


def exported_symbols_in_header() -> list:
    """Every `msr3d_*` function declared in include/msr3d_b200.h (for the load test)."""
    import re
    hdr = os.path.join(os.path.dirname(_HERE), "include", "msr3d_b200.h")
    text = open(hdr).read()
    return sorted(set(re.findall(r"\b(msr3d_[a-z0-9_]+)\s*\(", text)))
