"""Device-memory plumbing: torch CUDA tensors are used only as typed HBM buffers + stream handles."""
from __future__ import annotations

import ctypes

import torch


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("boosted_detr_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def ptr(t):
    """Device pointer of a contiguous tensor (None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.to(torch.float32).contiguous()
    return t


def i32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.int32 or not t.is_contiguous():
        t = t.to(torch.int32).contiguous()
    return t


def empty(*shape, dtype=torch.float32):
    return torch.empty(*shape, dtype=dtype, device=require_cuda())


def zeros(*shape, dtype=torch.float32):
    return torch.zeros(*shape, dtype=dtype, device=require_cuda())
