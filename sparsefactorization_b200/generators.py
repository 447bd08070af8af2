"""All W generators of a PSFNet in one fused CUDA kernel each way.

Replaces the per-factor ``fs[m](data)`` calls of the reference (SyntheticExperiments/psf.py:110-126, :175):
``Linear(E,H) -> GELU -> Linear(H,L)`` for every m, writing W directly in the chain's streaming
(tap-major) layout.  fp32 FMA arithmetic (no TF32); the backward pass recomputes the hidden activations.
Shapes the kernel does not cover (E or H != 32, L > 16, other layer stacks) stay on the batched-GEMM path
of ``_psfnet.generate_factors``.  CUDA only.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import torch
from torch import nn

from . import _capi

E_FUSED, H_FUSED, L_MAX = 32, 32, 16


class _Stats:
    def __init__(self):
        self.launches = 0


stats = _Stats()


def fusable(fs: Sequence[nn.Module], data: torch.Tensor) -> bool:
    """True when every generator is Linear(32,32) -> exact GELU -> Linear(32, L <= 16) on fp32 CUDA data."""
    if not (data.is_cuda and data.dtype == torch.float32 and data.shape[-1] == E_FUSED and len(fs) >= 1):
        return False
    L = None
    for f in fs:
        net = getattr(f, "network", None)
        if net is None or len(net) != 3:
            return False
        a, act, b = net
        if not (isinstance(a, nn.Linear) and isinstance(act, nn.GELU) and isinstance(b, nn.Linear)):
            return False
        if getattr(act, "approximate", "none") != "none" or a.bias is None or b.bias is None:
            return False
        if (a.in_features, a.out_features, b.in_features) != (E_FUSED, H_FUSED, H_FUSED) or b.out_features > L_MAX:
            return False
        if L is not None and b.out_features != L:
            return False
        L = b.out_features
        if a.weight.dtype != torch.float32:
            return False
    return True


class _FusedGenerators(torch.autograd.Function):
    """x [rows,32], w1 [n_W,32,32], b1 [n_W,32], w2 [n_W,L,32], b2 [n_W,L]  ->  W [n_W, L, rows]."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        lib = _capi.load()
        x, w1, b1, w2, b2 = (t.contiguous() for t in (x, w1, b1, w2, b2))
        rows = x.shape[0]
        n_W, L = w2.shape[0], w2.shape[1]
        out = torch.empty(n_W, L, rows, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = lib.psf_generators_fwd(x.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                        out.data_ptr(), rows, n_W, E_FUSED, H_FUSED, L, L * rows, rows,
                                        torch.cuda.current_stream().cuda_stream)
        _capi.check(rc, "psf_generators_fwd")
        stats.launches += 1
        ctx.save_for_backward(x, w1, b1, w2, b2)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        lib = _capi.load()
        x, w1, b1, w2, b2 = ctx.saved_tensors
        dout = dout.contiguous()
        rows = x.shape[0]
        n_W, L = w2.shape[0], w2.shape[1]
        dx = torch.empty_like(x)
        dw1, db1, dw2, db2 = (torch.empty_like(t) for t in (w1, b1, w2, b2))
        with torch.cuda.device(x.device):
            nbytes = int(lib.psf_generators_workspace_bytes(n_W))
            ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=x.device)
            rc = lib.psf_generators_bwd(x.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                        dout.data_ptr(), dx.data_ptr(), dw1.data_ptr(), db1.data_ptr(), dw2.data_ptr(),
                                        db2.data_ptr(), ws.data_ptr(), ws.numel(), rows, n_W, E_FUSED, H_FUSED, L,
                                        L * rows, rows, torch.cuda.current_stream().cuda_stream)
        _capi.check(rc, "psf_generators_bwd")
        stats.launches += 2
        return dx, dw1, db1, dw2, db2


def fused_generators(fs: Sequence[nn.Module], x2d: torch.Tensor) -> torch.Tensor:
    """``W[m, l, row] = fs[m](x2d)[row, l]`` for every generator; x2d ``[rows, 32]``; result ``[n_W, L, rows]``."""
    w1 = torch.stack([f.network[0].weight for f in fs], 0)
    b1 = torch.stack([f.network[0].bias for f in fs], 0)
    w2 = torch.stack([f.network[2].weight for f in fs], 0)
    b2 = torch.stack([f.network[2].bias for f in fs], 0)
    return _FusedGenerators.apply(x2d, w1, b1, w2, b2)
