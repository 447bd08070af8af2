"""Shared pieces of the PSFNet variants: chord indices, MLP builders, the batched W generators and
the fused chain call.  The public, reference-named modules are ``psf.py`` (Synthetic ctor),
``lra/psf.py`` and ``genome/psf.py``.
"""
from __future__ import annotations

import os

import ctypes as C
from typing import List, Sequence, Union

import torch
import torch.nn.functional as F
from torch import nn

from . import _capi
from . import generators as _gen
from .chain import chord_chain, preferred_interleave, row_permutation


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


class _TapMajorLastLinear(torch.autograd.Function):
    """Last generator layer for all n_W factors, output tap-major.

    h2d  ``[B*n, n_W*H]``  hidden activations of every generator side by side (one GEMM upstream)
    w    ``[n_W, L, H]``, b ``[n_W, L]``
    out  ``[n_W, L, B*n]`` = w[m] @ h_m^T + b[m]   (row index innermost: the chain's streaming order)

    The backward pass is written out because autograd's version is pathological here: the weight
    gradient is a GEMM with a 15 x 32 output and K = B*n (millions), for which cuBLAS falls back to a
    gemv-style kernel (88 ms per step at B = 132).  It is evaluated split-K as a strided batched GEMM.
    """

    KC = 2048      # split-K chunk (rows of B*n per partial product)

    @staticmethod
    def forward(ctx, h2d, w, b):
        n_W, L, H = w.shape
        BN = h2d.shape[0]
        hid3 = h2d.view(BN, n_W, H).permute(1, 0, 2)                     # [n_W, BN, H] strided view
        out = torch.baddbmm(b.unsqueeze(-1), w, hid3.transpose(1, 2))     # [n_W, L, BN]
        ctx.save_for_backward(h2d, w)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        h2d, w = ctx.saved_tensors
        n_W, L, H = w.shape
        BN = h2d.shape[0]
        dout = dout.contiguous()
        dh = dw = db = None
        if ctx.needs_input_grad[0]:
            dh = torch.empty_like(h2d)
            dh3 = dh.view(BN, n_W, H)
            for m in range(n_W):                                          # [BN, L] @ [L, H] -> strided slice
                torch.mm(dout[m].t(), w[m], out=dh3[:, m, :])
        if ctx.needs_input_grad[1]:
            kc = min(_TapMajorLastLinear.KC, BN)
            S, main = BN // kc, (BN // kc) * kc                           # S full chunks; the remainder separately
            h4 = h2d[:main].view(S, kc, n_W, H)
            dw = torch.empty_like(w)
            for m in range(n_W):
                a = dout[m, :, :main].reshape(L, S, kc).permute(1, 0, 2)  # [S, L, kc]
                part = torch.bmm(a, h4[:, :, m, :])                       # [S, L, H]
                torch.sum(part, dim=0, out=dw[m])
                if main < BN:
                    dw[m].addmm_(dout[m, :, main:], h2d[main:].view(BN - main, n_W, H)[:, m, :])
        if ctx.needs_input_grad[2]:
            db = dout.sum(-1)
        return dh, dw, db


USE_FUSED_GENERATORS = True      # tests flip this to compare against the batched-GEMM route


def generate_factors(fs: Sequence[MLPBlock], data: torch.Tensor, interleave: int = 1) -> torch.Tensor:
    """``W[m] = fs[m](data)`` for every m (psf.py:175), computed as batched GEMMs.

    ``interleave = Q > 1``: the generators are row-wise, so permuting the rows of ``data`` once makes
    every W (and dW on the way back) come out in the quad-interleaved row order the second-generation
    chain kernel streams (include/psf_b200.h); the permutation's backward is autograd's.

    data ``[B, n, E]`` -> logical ``[n_W, B, n, L]`` stored tap-major (physical ``[n_W, L, B*n]``):
    the last Linear of every generator is evaluated as ``weight @ hidden^T`` so its output lands
    with the row index innermost, which is the order the chain kernel streams.  Parameters stay
    in their per-module tensors (state-dict unchanged); the stacks below are tiny.
    """
    n_W = len(fs)
    B, n, E = data.shape
    if interleave > 1:
        data = data.view(B, n // (4 * interleave), interleave, 4, E).permute(0, 2, 1, 3, 4).reshape(B, n, E)
    x_shared = data.reshape(B * n, E)
    if USE_FUSED_GENERATORS and _gen.fusable(fs, data):
        wt, _ = _gen.fused_generators(fs, x_shared, n, interleave)  # [n_W, L, B*n], one kernel
        return wt.view(n_W, wt.shape[1], B, n).permute(0, 2, 3, 1)
    if not _same_structure(fs):
        # heterogeneous generators: evaluate one by one (reference order), then re-lay
        return torch.stack([f(data) for f in fs], 0)
    # activations of all generators side by side: [B*n, n_W*width] contiguous, so that elementwise
    # layers run on a dense buffer; None while the generators still share their input
    h2d = None
    width = E
    layers = list(zip(*[list(f.network) for f in fs]))
    last_linear = max(i for i, ls in enumerate(layers) if isinstance(ls[0], nn.Linear))
    for i, ls in enumerate(layers):
        if isinstance(ls[0], nn.GELU):
            if h2d is None:
                x_shared = F.gelu(x_shared)
            else:
                h2d = F.gelu(h2d)
            continue
        w = torch.stack([l.weight for l in ls], 0)                 # [n_W, out, in]
        bvec = torch.stack([l.bias for l in ls], 0)                # [n_W, out]
        out_f = w.shape[1]
        if i == last_linear:
            if h2d is None:                                        # no hidden layer at all
                h2d = x_shared.repeat(1, n_W)
            wt = _TapMajorLastLinear.apply(h2d, w, bvec)           # [n_W, L, B*n]
            return wt.view(n_W, out_f, B, n).permute(0, 2, 3, 1)   # logical [n_W, B, n, L]
        if h2d is None:
            # one GEMM for all generators: [B*n, in] x [in, n_W*out]
            h2d = torch.addmm(bvec.reshape(-1), x_shared, w.reshape(n_W * out_f, -1).t())
        else:
            x3 = h2d.view(B * n, n_W, width).permute(1, 0, 2)      # [n_W, B*n, width]
            y3 = torch.baddbmm(bvec.unsqueeze(1), x3, w.transpose(1, 2))
            h2d = y3.permute(1, 0, 2).reshape(B * n, n_W * out_f)
        width = out_f
    raise RuntimeError("generator has no Linear layer")


def generate_all(fs: Sequence[MLPBlock], g: MLPBlock, data: torch.Tensor, interleave: int = 1,
                 pre_interleaved: bool = False):
    """``(W, V0)``: every ``fs[m](data)`` and ``g(data)`` (psf.py:165, :175).  One fused kernel when the shapes
    allow (W tap-major in the requested row order, V0 natural ``[B, n, C]``), else the GEMM routes.
    ``pre_interleaved``: the rows of ``data`` already are in the quad-interleaved order (fused route only)."""
    B, n, E = data.shape
    if USE_FUSED_GENERATORS and _gen.fusable(fs, data, g):
        xp = data
        if interleave > 1 and not pre_interleaved:
            xp = data.view(B, n // (4 * interleave), interleave, 4, E).permute(0, 2, 1, 3, 4)
        wt, v0 = _gen.fused_generators(fs, xp.reshape(B * n, E), n, interleave, g)
        return wt.view(len(fs), wt.shape[1], B, n).permute(0, 2, 3, 1), v0.view(B, n, -1)
    if pre_interleaved and interleave > 1:
        raise RuntimeError("pre-interleaved rows need the fused generator kernels")
    return generate_factors(fs, data, interleave), g(data)


class _SkinnyLinear(torch.autograd.Function):
    """``x @ W^T + b`` for a very narrow input (init_linear = Linear(2, E), psf.py:136-141) over millions of rows: one
    streaming kernel each way (csrc/psf_embed.cu: a 277 MB write forward, one pass over dy for dW and db backward).  The
    library route - addmm forward, split-K batched GEMM and two column reductions backward - took 0.46 ms of the 8.4 ms
    step at cfg2.  Shapes outside the kernels (out_features not a multiple of 4, > 128) keep that route."""

    KC = 4096
    launches = 0      # kernels of libpsf_b200.so launched (bench.py's gpu_launches)

    @staticmethod
    def _native(x2d, w):
        E, IN = w.shape
        return (os.environ.get("PSF_INIT_LINEAR_KERNELS", "1") != "0" and x2d.dtype == torch.float32 and w.dtype == torch.float32 and IN <= 4 and E % 4 == 0 and E <= 128
                and 256 % (E // 4) == 0 and x2d.shape[0] > 0)

    @staticmethod
    def forward(ctx, x2d, w, b):
        x2d = x2d.contiguous()
        ctx.save_for_backward(x2d, w)
        ctx.native = _SkinnyLinear._native(x2d, w)
        if not ctx.native:
            return torch.addmm(b, x2d, w.t())
        lib = _capi.load()
        wc, bc = w.contiguous(), b.contiguous()
        y = torch.empty(x2d.shape[0], w.shape[0], dtype=torch.float32, device=x2d.device)
        with torch.cuda.device(x2d.device):
            rc = lib.psf_init_linear_fwd(x2d.data_ptr(), wc.data_ptr(), bc.data_ptr(), y.data_ptr(), x2d.shape[0], w.shape[1],
                                         w.shape[0], torch.cuda.current_stream().cuda_stream)
        _capi.check(rc, "psf_init_linear_fwd")
        _SkinnyLinear.launches += 1
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        x2d, w = ctx.saved_tensors
        dy = dy.contiguous()
        rows = x2d.shape[0]
        dx = dy @ w if ctx.needs_input_grad[0] else None
        if ctx.native:
            lib = _capi.load()
            dw, db = torch.empty_like(w, memory_format=torch.contiguous_format), torch.empty(w.shape[0], dtype=w.dtype, device=w.device)
            with torch.cuda.device(x2d.device):
                nbytes = int(lib.psf_init_linear_workspace_bytes(rows, w.shape[1], w.shape[0]))
                ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=x2d.device)
                rc = lib.psf_init_linear_bwd(x2d.data_ptr(), dy.data_ptr(), dw.data_ptr(), db.data_ptr(), rows, w.shape[1],
                                             w.shape[0], ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
            _capi.check(rc, "psf_init_linear_bwd")
            _SkinnyLinear.launches += 2            # the pass over dy + the reduction of the CTA partials
            return dx, dw, db
        kc = min(_SkinnyLinear.KC, rows)
        S, main = rows // kc, (rows // kc) * kc                                            # full chunks + remainder
        part = torch.bmm(dy[:main].view(S, kc, -1).transpose(1, 2), x2d[:main].reshape(S, kc, -1))   # [S, E, in]
        dw = part.sum(0)
        if main < rows:
            dw = dw + dy[main:].t() @ x2d[main:]
        return dx, dw, dy.sum(0)


def skinny_linear(lin: nn.Linear, x: torch.Tensor) -> torch.Tensor:
    """``lin(x)`` through :class:`_SkinnyLinear` when the input is narrow, on CUDA and batched; else ``lin(x)``."""
    if lin.in_features <= 4 and x.is_cuda and lin.bias is not None and x.dim() >= 2:
        y = _SkinnyLinear.apply(x.reshape(-1, lin.in_features), lin.weight, lin.bias)
        return y.view(*x.shape[:-1], lin.out_features)
    return lin(x)


class PSFChainMixin:
    """forward() tail shared by the variants: W generation + fused chain (psf.py:165-188)."""

    def _early_interleave(self, raw: torch.Tensor) -> int:
        """Row interleave Q (> 1) if this batch will run the fused generators + gen-2 chain: the caller may then
        permute the rows of the RAW input (tokens / 2 floats per row) instead of the embedded rows.  1 otherwise."""
        if not (USE_FUSED_GENERATORS and raw.is_cuda and self.embedding_size == _gen.E_FUSED
                and _gen.fusable_modules(self.fs, self.g)):
            return 1
        L = self.fs[0].network[-1].out_features
        return preferred_interleave(len(self.fs), raw.shape[0], self.n_vec, L, self.g.network[-1].out_features)

    def _row_perm(self, Q: int, device) -> torch.Tensor:
        key = (Q, str(device))
        cache = self.__dict__.setdefault("_perm_cache", {})
        if key not in cache:
            cache[key] = row_permutation(self.n_vec, Q, device)
        return cache[key]

    def _chain_from_data(self, data: torch.Tensor, v_dropout=None, pre_interleaved: int = 1) -> torch.Tensor:
        """``V = g(data)`` (+ the LRA variants' dropout2), ``W_m = fs[m](data)``, then the chain.
        ``pre_interleaved = Q > 1``: the rows of ``data`` are already quad-interleaved (see _early_interleave)."""
        n_W, (B, n, _) = len(self.fs), data.shape
        L = self.fs[0].network[-1].out_features
        Cv = self.g.network[-1].out_features
        homog = _same_structure(self.fs)
        Q = preferred_interleave(n_W, B, n, L, Cv) if data.dtype == torch.float32 and homog else 1
        if pre_interleaved > 1:
            if pre_interleaved != Q or data.dtype != torch.float32:
                raise RuntimeError("rows were pre-interleaved for a kernel that is not going to run")
            W, V = generate_all(self.fs, self.g, data, Q, pre_interleaved=True)
        elif homog:
            W, V = generate_all(self.fs, self.g, data, Q)
        else:
            W, V = generate_factors(self.fs, data, 1), self.g(data)
        if v_dropout is not None:
            V = v_dropout(V)
        return chord_chain(W, V, use_residuals=self.use_residuals, interleave=Q)

    @torch.no_grad()
    def attention_map(self, data: torch.Tensor) -> torch.Tensor:
        """Dense ``[B, n, n]`` product of the factors generated from the embedded ``data`` ``[B, n, E]`` (what
        ChangedPSF.forward returns as ``W_final``, LRA/attention_maps/imdb_inference.py:24-69)."""
        from .chain import attention_map as _amap
        n_W, (B, n, _) = len(self.fs), data.shape
        L = self.fs[0].network[-1].out_features
        homog = _same_structure(self.fs)
        Q = preferred_interleave(n_W, B, n, L, n) if data.dtype == torch.float32 and homog and n % 8 == 0 else 1
        return _amap(generate_factors(self.fs, data, Q), interleave=Q)

    def _apply_chain(self, data: torch.Tensor, V: torch.Tensor) -> torch.Tensor:
        n_W, (B, n, _) = len(self.fs), data.shape
        L = self.fs[0].network[-1].out_features
        # (heterogeneous generators are evaluated one by one in the reference layout: natural row order)
        Q = preferred_interleave(n_W, B, n, L, V.shape[2]) if V.dtype == torch.float32 and _same_structure(self.fs) else 1
        W = generate_factors(self.fs, data, Q)
        return chord_chain(W, V, use_residuals=self.use_residuals, interleave=Q)

    def _make_indices(self):
        self.chord_indicies = chord_index_tensor(self.n_vec, self.n_links, "cuda" if self.use_cuda else "cpu")
