"""Shared pieces of the PSFNet variants: chord indices, MLP builders, the batched W generators and
the fused chain call.  The public, reference-named modules are ``psf.py`` (Synthetic ctor),
``lra/psf.py`` and ``genome/psf.py``.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Union

import torch
import torch.nn.functional as F
from torch import nn

from . import _capi
from .chain import chord_chain


# ---------------------------------------------------------------------------------------------
# chord pattern (integer, bit-exact with the reference builder, psf.py:7-32)
# ---------------------------------------------------------------------------------------------
def chord_offsets(n_vec: int, n_link: int) -> List[int]:
    lib = _capi.load()
    arr = (C.c_int64 * n_link)()
    _capi.check(lib.psf_chord_offsets(n_vec, n_link, arr), "psf_chord_offsets")
    return list(arr)


def chord_index_tensor(n_vec: int, n_link: int, device="cpu") -> torch.Tensor:
    """int64 ``[2, n_vec*n_link]`` (rows, cols): what ``torch.tensor(get_chord_indices_assym(...))`` holds."""
    lib = _capi.load()
    device = torch.device(device)
    out = torch.empty(2, n_vec * n_link, dtype=torch.int64, device=device)
    if device.type == "cuda":
        with torch.cuda.device(device):
            rc = lib.psf_chord_indices_device(n_vec, n_link, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    else:
        p = C.cast(out.data_ptr(), C.POINTER(C.c_int64))
        p2 = C.cast(out.data_ptr() + 8 * n_vec * n_link, C.POINTER(C.c_int64))
        rc = lib.psf_chord_indices_host(n_vec, n_link, p, p2)
    _capi.check(rc, "psf_chord_indices")
    return out


def get_chord_indices_assym(n_vec: int, n_link: int):
    """Drop-in for psf.py:7-32: ``(rows, cols)`` as two Python lists of length n_vec*n_link;
    entry i*n_link+k is (i, (i + o_k) % n_vec) with o_0 = 0, o_k = 2**(k-1)."""
    idx = chord_index_tensor(n_vec, n_link, "cpu")
    return idx[0].tolist(), idx[1].tolist()


def chord_mask(N: int, b: int = 2, symmetric: bool = False, self_loop: bool = True) -> torch.Tensor:
    """Dense ``N x N`` 0/1 chord mask (float64), non-parametric/chord_mask_mat.m (random_order=false)."""
    lib = _capi.load()
    out = torch.empty(N, N, dtype=torch.float64)
    rc = lib.psf_chord_mask_host(N, b, int(symmetric), int(self_loop), C.cast(out.data_ptr(), C.POINTER(C.c_double)))
    _capi.check(rc, "psf_chord_mask_host")
    return out


# ---------------------------------------------------------------------------------------------
# MLP builders (psf.py:35-60): ints are Linear widths, any str is a GELU, closing Linear
# ---------------------------------------------------------------------------------------------
def MakeMLP(cfg: List[Union[str, int]], in_channels: int, out_channels: int) -> nn.Sequential:
    layers: List[nn.Module] = []
    width = in_channels
    for item in cfg:
        if isinstance(item, int):
            layers.append(nn.Linear(width, item))
            width = item
        else:
            layers.append(nn.GELU())
    layers.append(nn.Linear(width, out_channels))
    return nn.Sequential(*layers)


class MLPBlock(nn.Module):
    def __init__(self, cfg, in_dim, out_dim):
        super().__init__()
        self.network = MakeMLP(cfg, in_dim, out_dim)   # the attribute name fixes the state-dict keys

    def forward(self, data):
        return self.network(data)


# ---------------------------------------------------------------------------------------------
# all n_W generators at once, emitting W in the streaming (tap-major) order
# ---------------------------------------------------------------------------------------------
def _same_structure(fs: Sequence[MLPBlock]) -> bool:
    ref = [(type(l), getattr(l, "in_features", None), getattr(l, "out_features", None)) for l in fs[0].network]
    for f in fs[1:]:
        cur = [(type(l), getattr(l, "in_features", None), getattr(l, "out_features", None)) for l in f.network]
        if cur != ref:
            return False
    return all(t in (nn.Linear, nn.GELU) for t, _, _ in ref) and all(
        getattr(l, "approximate", "none") == "none" for l in fs[0].network if isinstance(l, nn.GELU))


def generate_factors(fs: Sequence[MLPBlock], data: torch.Tensor) -> torch.Tensor:
    """``W[m] = fs[m](data)`` for every m (psf.py:175), computed as batched GEMMs.

    data ``[B, n, E]`` -> logical ``[n_W, B, n, L]`` stored tap-major (physical ``[n_W, L, B*n]``):
    the last Linear of every generator is evaluated as ``weight @ hidden^T`` so its output lands
    with the row index innermost, which is the order the chain kernel streams.  Parameters stay
    in their per-module tensors (state-dict unchanged); the stacks below are tiny.
    """
    n_W = len(fs)
    B, n, E = data.shape
    x_shared = data.reshape(B * n, E)
    if not _same_structure(fs):
        # heterogeneous generators: evaluate one by one (reference order), then re-lay
        return torch.stack([f(data) for f in fs], 0)
    x = None          # [n_W, B*n, width] once the generators have diverged
    layers = list(zip(*[list(f.network) for f in fs]))
    last_linear = max(i for i, ls in enumerate(layers) if isinstance(ls[0], nn.Linear))
    for i, ls in enumerate(layers):
        if isinstance(ls[0], nn.GELU):
            if x is None:      # GELU straight on the shared input
                x_shared = F.gelu(x_shared)
            else:
                x = F.gelu(x)
            continue
        w = torch.stack([l.weight for l in ls], 0)                 # [n_W, out, in]
        bvec = torch.stack([l.bias for l in ls], 0)                # [n_W, out]
        out_f = w.shape[1]
        if i == last_linear:
            if x is None:
                xt = x_shared.t().unsqueeze(0).expand(n_W, -1, -1)   # [n_W, in, B*n]
            else:
                xt = x.transpose(1, 2)
            wt = torch.baddbmm(bvec.unsqueeze(-1), w, xt)          # [n_W, L, B*n]
            return wt.view(n_W, out_f, B, n).permute(0, 2, 3, 1)   # logical [n_W, B, n, L]
        if x is None:
            # one GEMM for all generators: [B*n, in] x [in, n_W*out]
            h = torch.addmm(bvec.reshape(-1), x_shared, w.reshape(n_W * out_f, -1).t())
            x = h.view(B * n, n_W, out_f).permute(1, 0, 2)         # [n_W, B*n, out], no copy
        else:
            x = torch.baddbmm(bvec.unsqueeze(1), x, w.transpose(1, 2))
    raise RuntimeError("generator has no Linear layer")


class PSFChainMixin:
    """forward() tail shared by the variants: W generation + fused chain (psf.py:165-188)."""

    def _apply_chain(self, data: torch.Tensor, V: torch.Tensor) -> torch.Tensor:
        W = generate_factors(self.fs, data)
        return chord_chain(W, V, use_residuals=self.use_residuals)

    def _make_indices(self):
        self.chord_indicies = chord_index_tensor(self.n_vec, self.n_links, "cuda" if self.use_cuda else "cpu")
