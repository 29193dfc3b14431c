"""`pointnet2._ext`-compatible operator module backed by libmsr3d_b200.so.

Same nine names, argument order, dtypes, output shapes and error behaviour as the
reference's pybind module (modules/third_party/pointnet2/_ext_src/src/bindings.cpp:6-19):
inputs must be contiguous float32 / int32 CUDA tensors (utils.h:5-25 -> RuntimeError),
CPU tensors are refused ("CPU not supported", ball_query.cpp:27-29), outputs are
allocated here on the input's device, and every launch goes to torch's current
stream (ball_query_gpu.cu:49).  `msr3d_b200.compat.install()` publishes this
module as ``pointnet2._ext`` so the reference's pointnet2_utils.py binds to it.
"""
from __future__ import annotations

import torch

from . import _C


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _check(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if t.dtype != dtype:
        kind = "float" if dtype == torch.float32 else "int"
        raise RuntimeError(f"{name} must be a {kind} tensor")
    if not t.is_cuda:
        raise RuntimeError("CPU not supported")


def furthest_point_sampling(points: torch.Tensor, nsamples: int) -> torch.Tensor:
    _check(points, torch.float32, "points")
    B, N, _ = points.shape
    out = torch.empty(B, nsamples, dtype=torch.int32, device=points.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(points.device):
        _C.call("msr3d_furthest_point_sampling_f32", points.data_ptr(), out.data_ptr(),
                B, N, int(nsamples), _stream())
    return out


def gather_points(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    _check(points, torch.float32, "points"); _check(idx, torch.int32, "idx")
    B, C, N = points.shape
    M = idx.shape[1]
    out = torch.empty(B, C, M, dtype=torch.float32, device=points.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(points.device):
        _C.call("msr3d_gather_points_f32", points.data_ptr(), idx.data_ptr(), out.data_ptr(),
                B, C, N, M, _stream())
    return out


def gather_points_grad(grad_out: torch.Tensor, idx: torch.Tensor, n: int) -> torch.Tensor:
    _check(grad_out, torch.float32, "grad_out"); _check(idx, torch.int32, "idx")
    B, C, M = grad_out.shape
    out = torch.zeros(B, C, int(n), dtype=torch.float32, device=grad_out.device)
    if grad_out.numel() == 0 or out.numel() == 0:
        return out
    with torch.cuda.device(grad_out.device):
        _C.call("msr3d_gather_points_grad_f32", grad_out.data_ptr(), idx.data_ptr(),
                out.data_ptr(), B, C, int(n), M, _stream())
    return out


def ball_query(new_xyz: torch.Tensor, xyz: torch.Tensor, radius: float, nsample: int) -> torch.Tensor:
    _check(new_xyz, torch.float32, "new_xyz"); _check(xyz, torch.float32, "xyz")
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = torch.empty(B, M, int(nsample), dtype=torch.int32, device=xyz.device)
    if idx.numel() == 0:
        return idx
    with torch.cuda.device(xyz.device):
        _C.call("msr3d_ball_query_f32", new_xyz.data_ptr(), xyz.data_ptr(), idx.data_ptr(),
                B, N, M, float(radius), int(nsample), _stream())
    return idx


def group_points(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    _check(points, torch.float32, "points"); _check(idx, torch.int32, "idx")
    B, C, N = points.shape
    _, P, S = idx.shape
    out = torch.empty(B, C, P, S, dtype=torch.float32, device=points.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(points.device):
        _C.call("msr3d_group_points_f32", points.data_ptr(), idx.data_ptr(), out.data_ptr(),
                B, C, N, P, S, _stream())
    return out


def group_points_grad(grad_out: torch.Tensor, idx: torch.Tensor, n: int) -> torch.Tensor:
    _check(grad_out, torch.float32, "grad_out"); _check(idx, torch.int32, "idx")
    B, C, P, S = grad_out.shape
    out = torch.zeros(B, C, int(n), dtype=torch.float32, device=grad_out.device)
    if grad_out.numel() == 0 or out.numel() == 0:
        return out
    with torch.cuda.device(grad_out.device):
        _C.call("msr3d_group_points_grad_f32", grad_out.data_ptr(), idx.data_ptr(),
                out.data_ptr(), B, C, int(n), P, S, _stream())
    return out


def three_nn(unknowns: torch.Tensor, knows: torch.Tensor):
    _check(unknowns, torch.float32, "unknowns"); _check(knows, torch.float32, "knows")
    B, n, _ = unknowns.shape
    m = knows.shape[1]
    dist2 = torch.empty(B, n, 3, dtype=torch.float32, device=unknowns.device)
    idx = torch.empty(B, n, 3, dtype=torch.int32, device=unknowns.device)
    with torch.cuda.device(unknowns.device):
        _C.call("msr3d_three_nn_f32", unknowns.data_ptr(), knows.data_ptr(), dist2.data_ptr(),
                idx.data_ptr(), B, n, m, _stream())
    return [dist2, idx]


def three_interpolate(points: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    _check(points, torch.float32, "points"); _check(idx, torch.int32, "idx")
    _check(weight, torch.float32, "weight")
    B, c, m = points.shape
    n = idx.shape[1]
    out = torch.empty(B, c, n, dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        _C.call("msr3d_three_interpolate_f32", points.data_ptr(), idx.data_ptr(),
                weight.data_ptr(), out.data_ptr(), B, c, m, n, _stream())
    return out


def three_interpolate_grad(grad_out: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor, m: int) -> torch.Tensor:
    _check(grad_out, torch.float32, "grad_out"); _check(idx, torch.int32, "idx")
    _check(weight, torch.float32, "weight")
    B, c, n = grad_out.shape
    out = torch.empty(B, c, int(m), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        _C.call("msr3d_three_interpolate_grad_f32", grad_out.data_ptr(), idx.data_ptr(),
                weight.data_ptr(), out.data_ptr(), B, c, n, int(m), _stream())
    return out
