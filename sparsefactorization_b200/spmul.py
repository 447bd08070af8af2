"""Drop-in for the reference's ``spmul/spmul.py`` + the ``spmul_cuda`` extension module.

``SparseMultiply.apply(F, V, offsets=None, n_block=16, n_thread_vec=64, n_thread_dim=16,
n_thread_link=16)`` keeps the reference signature (spmul.py:15); the four launch-tuning
integers are accepted and ignored -- the library sizes its own grid.  ``forward_host`` /
``backward_host`` mirror the pybind functions of spmul_cuda.cu:163-166.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _capi
from .chain import _DTYPES, _offsets_array, _require_cuda


def get_offsets(n_link_all):
    """LongTensor ``[0, 1, 2, 4, ...]`` on the GPU (spmul.py:8-9)."""
    return torch.tensor([0] + [2 ** k for k in range(n_link_all - 1)], dtype=torch.int64,
                        requires_grad=False).cuda()


def _check(F, V):
    _require_cuda(F, V)
    if F.dim() != 3 or V.dim() != 3 or F.shape[:2] != V.shape[:2]:
        raise ValueError(f"expected F [B,n,L] and V [B,n,D], got {tuple(F.shape)} and {tuple(V.shape)}")
    if F.dtype not in _DTYPES or V.dtype != F.dtype:
        raise TypeError("F and V must both be float32 or both float64")


def forward_host(F, V, offsets, n_block=16, n_thread_vec=64, n_thread_dim=16):
    """Z[b,p,d] = sum_k F[b,p,k] * V[b,(p+offsets[k]) % n,d]   (spmul_cuda.cu:24,31-59)."""
    lib = _capi.load()
    _check(F, V)
    B, n, L = F.shape
    D = V.shape[2]
    F, V = F.contiguous(), V.contiguous()
    offs_arr, _ = _offsets_array(offsets, L)
    Z = torch.empty_like(V)
    with torch.cuda.device(F.device):
        rc = lib.psf_factor_fwd(_DTYPES[F.dtype], F.data_ptr(), V.data_ptr(), Z.data_ptr(),
                                C.cast(offs_arr, C.POINTER(C.c_int64)), B, n, L, D,
                                torch.cuda.current_stream().cuda_stream)
    _capi.check(rc, "psf_factor_fwd")
    return Z


def backward_host(dJdZ, F, V, offsets, n_block=16, n_thread_vec=64, n_thread_dim=16, n_thread_link=16):
    """[dJdF, dJdV]   (spmul_cuda.cu:79-80,105-108,114-159)."""
    lib = _capi.load()
    _check(F, V)
    B, n, L = F.shape
    D = V.shape[2]
    F, V, dJdZ = F.contiguous(), V.contiguous(), dJdZ.contiguous()
    offs_arr, _ = _offsets_array(offsets, L)
    dF, dV = torch.empty_like(F), torch.empty_like(V)
    with torch.cuda.device(F.device):
        rc = lib.psf_factor_bwd(_DTYPES[F.dtype], dJdZ.data_ptr(), F.data_ptr(), V.data_ptr(), dF.data_ptr(),
                                dV.data_ptr(), C.cast(offs_arr, C.POINTER(C.c_int64)), B, n, L, D,
                                torch.cuda.current_stream().cuda_stream)
    _capi.check(rc, "psf_factor_bwd")
    return [dF, dV]


class SparseMultiply(torch.autograd.Function):

    @staticmethod
    def forward(ctx, F, V, offsets=None, n_block=16, n_thread_vec=64, n_thread_dim=16, n_thread_link=16):
        if offsets is None:
            offsets = get_offsets(F.shape[-1])
        ctx.offsets = offsets
        ctx.save_for_backward(F, V)
        return forward_host(F, V, offsets, n_block, n_thread_vec, n_thread_dim)

    @staticmethod
    def backward(ctx, dJdZ):
        F, V = ctx.saved_tensors
        dJdF, dJdV = backward_host(dJdZ, F, V, ctx.offsets)
        return dJdF, dJdV, None, None, None, None, None
