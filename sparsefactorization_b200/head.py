"""Head of the FLATTEN pooling: ``final(V.view(B, -1))`` with ``final = nn.Linear(n_vec * n_channels_V, n_class)``
(SyntheticExperiments/psf.py:60,80; LRA/psf.py:13-22,109) as two streaming CUDA kernels (csrc/psf_head.cu).

At the headline shape the contraction length is 131072 and ``n_class`` is 1: a pass over V that is bound by HBM (13 us
forward, 20 us backward inside the step; the library's gemv route takes 31 + 85 us in isolation).  CUDA only; shapes the
kernels do not cover (more than 16 outputs, K not a multiple of 4, non-fp32) use ``nn.Linear`` as before.
"""
from __future__ import annotations

import ctypes as C  # noqa: F401
import os

import torch
from torch import nn

from . import _capi

MAX_CLASSES = 16
launches = 0          # kernels of libpsf_b200.so launched through this module (bench.py's gpu_launches)


class _FlatHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, V2d, weight, bias):
        lib = _capi.load()
        V2d, weight = V2d.contiguous(), weight.contiguous()
        B, K = V2d.shape
        nc = weight.shape[0]
        out = torch.empty(B, nc, dtype=torch.float32, device=V2d.device)
        with torch.cuda.device(V2d.device):
            ws = torch.empty(max(int(lib.psf_head_workspace_bytes(B, K, nc, 0)), 256), dtype=torch.uint8, device=V2d.device)
            rc = lib.psf_head_fwd(V2d.data_ptr(), weight.data_ptr(), bias.data_ptr() if bias is not None else None,
                                  out.data_ptr(), B, K, nc, ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
        _capi.check(rc, "psf_head_fwd")
        global launches
        launches += 2                              # chunk sums + their reduction
        ctx.save_for_backward(V2d, weight)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        lib = _capi.load()
        V2d, weight = ctx.saved_tensors
        g = g.contiguous()
        B, K = V2d.shape
        nc = weight.shape[0]
        dV = torch.empty_like(V2d)
        dW = torch.empty_like(weight)
        with torch.cuda.device(V2d.device):
            ws = torch.empty(max(int(lib.psf_head_workspace_bytes(B, K, nc, 1)), 256), dtype=torch.uint8, device=V2d.device)
            rc = lib.psf_head_bwd(V2d.data_ptr(), weight.data_ptr(), g.data_ptr(), dV.data_ptr(), dW.data_ptr(), B, K, nc,
                                  ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
        _capi.check(rc, "psf_head_bwd")
        global launches
        launches += 2 if B >= 8 else 1             # fused dV / dW pass (+ the reduction over the batch slices)
        return dV, dW, (g.sum(0) if ctx.has_bias else None)


def covers(lin: nn.Module, V2d: torch.Tensor) -> bool:
    return (os.environ.get("PSF_HEAD_KERNELS", "1") != "0" and isinstance(lin, nn.Linear) and V2d.is_cuda and V2d.dim() == 2 and V2d.dtype == torch.float32
            and lin.weight.dtype == torch.float32 and lin.out_features <= MAX_CLASSES and V2d.shape[1] % 4 == 0
            and V2d.shape[1] >= 4096 and 0 < V2d.shape[0] <= 65535)


def flat_linear(lin: nn.Linear, V2d: torch.Tensor) -> torch.Tensor:
    """``lin(V2d)``; through the streaming kernels when they cover the shape."""
    if covers(lin, V2d):
        return _FlatHead.apply(V2d, lin.weight, lin.bias)
    return lin(V2d)


def apply_head(final: nn.Module, V2d: torch.Tensor) -> torch.Tensor:
    """``final(V2d)`` for the reference's two head forms: a Linear, or Linear -> GELU -> Linear (LRA/psf.py:18-21)."""
    if isinstance(final, nn.Linear):
        return flat_linear(final, V2d)
    if isinstance(final, nn.Sequential) and len(final) > 0 and isinstance(final[0], nn.Linear):
        x = flat_linear(final[0], V2d)
        for layer in list(final)[1:]:
            x = layer(x)
        return x
    return final(V2d)
