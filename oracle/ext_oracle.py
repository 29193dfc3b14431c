"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by msr3d_b200/).

ctypes front-end of ``oracle/pointnet2_oracle.c`` exposing the same nine
functions, argument order and allocation rules as the reference's pybind module
``pointnet2._ext`` (modules/third_party/pointnet2/_ext_src/src/bindings.cpp:6-19),
but on CPU float32/int32 torch tensors.  Used as

* the checker in ``tests/`` (CUDA path vs this, bit-exact for int32 outputs),
* the ``pointnet2._ext`` stand-in when the *unmodified* reference Python modules
  are imported in the authoring container to produce ``tests/golden/*``
  (oracle/refshim.py), and
* the native half of bench.py's ``cpu_baseline`` / ``--impl reference`` legs.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")

_i32p = ctypes.POINTER(ctypes.c_int32)
_f32p = ctypes.POINTER(ctypes.c_float)


def _cpu_has_fma() -> bool:
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("flags"):
                    return " fma " in line + " "
    except OSError:
        pass
    return False


def build(force: bool = False) -> str:
    """Compile the C oracle with the committed Makefile (gcc only)."""
    name = "libpointnet2_oracle.so" if _cpu_has_fma() else "libpointnet2_oracle_nofma.so"
    path = os.path.join(_BUILD, name)
    src = os.path.join(_HERE, "pointnet2_oracle.c")
    stale = (not os.path.exists(path)) or os.path.getmtime(path) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)
    return path


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_set_threads.argtypes = [ctypes.c_int]
        _lib.oracle_get_threads.restype = ctypes.c_int
        _lib.oracle_opt_n_threads.argtypes = [ctypes.c_int]
        _lib.oracle_opt_n_threads.restype = ctypes.c_int
    return _lib


def set_threads(n: int) -> None:
    lib().oracle_set_threads(int(n))


def opt_n_threads(n: int) -> int:
    return lib().oracle_opt_n_threads(int(n))


#: 1 = the FMA contraction the reference binary (nvcc default -fmad=true) uses.
FMAD = 1


def _f(t: torch.Tensor):
    return ctypes.cast(t.data_ptr(), _f32p)


def _i(t: torch.Tensor):
    return ctypes.cast(t.data_ptr(), _i32p)


def _chk(t: torch.Tensor, dtype, name: str) -> None:
    # utils.h:5-25 — contiguity / dtype assertions become RuntimeError.
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be a {'float' if dtype == torch.float32 else 'int'} tensor")
    if t.device.type != "cpu":
        raise RuntimeError("oracle works on CPU tensors only")


_ref_mod = None


def ref_cuda():
    """The reference's own CUDA extension compiled for sm_100 into oracle/_ref (oracle/build_ref.py): serves
    CUDA tensors when the oracle port runs on the GPU (bench.py's `gpu_baseline`: reference kernels + stock
    torch ops on the same B200)."""
    global _ref_mod
    if _ref_mod is None:
        import importlib.util
        path = os.path.join(_HERE, "_ref", "pointnet2_ref_ext.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} not found (python oracle/build_ref.py builds it where /root/reference exists)")
        spec = importlib.util.spec_from_file_location("pointnet2_ref_ext", path)
        _ref_mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_ref_mod)
    return _ref_mod


def furthest_point_sampling(points: torch.Tensor, nsamples: int) -> torch.Tensor:
    if points.is_cuda:
        return ref_cuda().furthest_point_sampling(points, int(nsamples))
    _chk(points, torch.float32, "points")
    b, n, _ = points.shape
    out = torch.zeros(b, nsamples, dtype=torch.int32)
    rc = lib().oracle_furthest_point_sampling(_f(points), _i(out), b, n, int(nsamples), FMAD)
    assert rc == 0
    return out


def gather_points(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    _chk(points, torch.float32, "points"); _chk(idx, torch.int32, "idx")
    b, c, n = points.shape
    m = idx.shape[1]
    out = torch.zeros(b, c, m, dtype=torch.float32)
    lib().oracle_gather_points(_f(points), _i(idx), _f(out), b, c, n, m)
    return out


def gather_points_grad(grad_out: torch.Tensor, idx: torch.Tensor, n: int) -> torch.Tensor:
    _chk(grad_out, torch.float32, "grad_out"); _chk(idx, torch.int32, "idx")
    b, c, m = grad_out.shape
    out = torch.zeros(b, c, n, dtype=torch.float32)
    lib().oracle_gather_points_grad(_f(grad_out), _i(idx), _f(out), b, c, int(n), m)
    return out


def ball_query(new_xyz: torch.Tensor, xyz: torch.Tensor, radius: float, nsample: int) -> torch.Tensor:
    if new_xyz.is_cuda:
        return ref_cuda().ball_query(new_xyz, xyz, float(radius), int(nsample))
    _chk(new_xyz, torch.float32, "new_xyz"); _chk(xyz, torch.float32, "xyz")
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = torch.zeros(b, m, nsample, dtype=torch.int32)
    lib().oracle_ball_query(_f(new_xyz), _f(xyz), _i(idx), b, n, m,
                            ctypes.c_float(radius), int(nsample), FMAD)
    return idx


def group_points(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    _chk(points, torch.float32, "points"); _chk(idx, torch.int32, "idx")
    b, c, n = points.shape
    _, npoint, ns = idx.shape
    out = torch.zeros(b, c, npoint, ns, dtype=torch.float32)
    lib().oracle_group_points(_f(points), _i(idx), _f(out), b, c, n, npoint, ns)
    return out


def group_points_grad(grad_out: torch.Tensor, idx: torch.Tensor, n: int) -> torch.Tensor:
    _chk(grad_out, torch.float32, "grad_out"); _chk(idx, torch.int32, "idx")
    b, c, npoint, ns = grad_out.shape
    out = torch.zeros(b, c, n, dtype=torch.float32)
    lib().oracle_group_points_grad(_f(grad_out), _i(idx), _f(out), b, c, int(n), npoint, ns)
    return out


def three_nn(unknown: torch.Tensor, known: torch.Tensor) -> List[torch.Tensor]:
    _chk(unknown, torch.float32, "unknowns"); _chk(known, torch.float32, "knows")
    b, n, _ = unknown.shape
    m = known.shape[1]
    idx = torch.zeros(b, n, 3, dtype=torch.int32)
    dist2 = torch.zeros(b, n, 3, dtype=torch.float32)
    lib().oracle_three_nn(_f(unknown), _f(known), _f(dist2), _i(idx), b, n, m, FMAD)
    return [dist2, idx]


def three_interpolate(points: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    _chk(points, torch.float32, "points"); _chk(idx, torch.int32, "idx"); _chk(weight, torch.float32, "weight")
    b, c, m = points.shape
    n = idx.shape[1]
    out = torch.zeros(b, c, n, dtype=torch.float32)
    lib().oracle_three_interpolate(_f(points), _i(idx), _f(weight), _f(out), b, c, m, n, FMAD)
    return out


def three_interpolate_grad(grad_out: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor, m: int) -> torch.Tensor:
    _chk(grad_out, torch.float32, "grad_out"); _chk(idx, torch.int32, "idx"); _chk(weight, torch.float32, "weight")
    b, c, n = grad_out.shape
    out = torch.zeros(b, c, m, dtype=torch.float32)
    lib().oracle_three_interpolate_grad(_f(grad_out), _i(idx), _f(weight), _f(out), b, c, n, int(m))
    return out
