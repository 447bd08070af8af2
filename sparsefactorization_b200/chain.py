"""The fused chord-factor chain as one autograd operator.

Replaces the Python loop over ``torch_sparse.spmm`` in the reference's ``PSFNet.forward``
(SyntheticExperiments/psf.py:167-188; LRA/psf.py:219-240): all ``n_W`` factors, forward and
backward, are two calls into libpsf_b200.so.  CUDA only -- there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _capi

_PATHS = {"auto": _capi.PATH_AUTO, "generic": _capi.PATH_GENERIC, "fused": _capi.PATH_FUSED,
          "fused_gen1": _capi.PATH_FUSED_GEN1}
_PATH_NAMES = {_capi.PATH_AUTO: "auto", _capi.PATH_GENERIC: "generic", _capi.PATH_FUSED: "fused"}
_DTYPES = {torch.float32: _capi.PSF_F32, torch.float64: _capi.PSF_F64}

last_path_used: Optional[str] = None   # "fused" / "generic": which implementation the last forward took


class _Stats:
    """Launch counter and optional CUDA-event timing of the chain kernels (used by bench.py)."""

    def __init__(self):
        self.launches = 0          # calls into psf_chain_fwd / psf_chain_bwd that launched kernels
        self.timing = False
        self.events = {"fwd": [], "bwd": []}

    def reset(self, timing: bool = False):
        self.launches = 0
        self.timing = timing
        self.events = {"fwd": [], "bwd": []}

    def mean_ms(self, kind: str) -> float:
        """Mean device time of the recorded launches (call after torch.cuda.synchronize())."""
        ev = self.events[kind]
        return sum(a.elapsed_time(b) for a, b in ev) / len(ev) if ev else float("nan")


stats = _Stats()


class _Timed:
    def __init__(self, kind):
        self.kind = kind

    def __enter__(self):
        stats.launches += 1
        if stats.timing:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()          # current stream == the stream the kernel is launched on
        return self

    def __exit__(self, *exc):
        if stats.timing:
            self.b.record()
            stats.events[self.kind].append((self.a, self.b))
        return False


def _require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("sparsefactorization_b200: the PSF chain runs on CUDA only (no CPU fallback); "
                               "got a tensor on %s" % t.device)


def _offsets_array(offsets: Optional[Sequence[int]], L: int):
    if offsets is None:
        return None, None
    if isinstance(offsets, torch.Tensor):
        offsets = offsets.detach().cpu().tolist()
    offsets = [int(o) for o in offsets]
    if len(offsets) != L:
        raise ValueError(f"need {L} offsets, got {len(offsets)}")
    arr = (C.c_int64 * L)(*offsets)
    return arr, offsets


def _dense_nonoverlapping(t: torch.Tensor) -> bool:
    # every stride order of a 4-d dense tensor is a permutation of a contiguous one
    sizes_strides = sorted(((st, sz) for sz, st in zip(t.shape, t.stride()) if sz > 1))
    expect = 1
    for st, sz in sizes_strides:
        if st != expect:
            return False
        expect *= sz
    return True


def _desc(W: torch.Tensor, V0: torch.Tensor, residual: bool, offs_arr, path: int,
          interleave: int = 1) -> _capi.ChainDesc:
    n_W, B, n, L = W.shape
    d = _capi.ChainDesc()
    d.B, d.n, d.C = B, n, V0.shape[2]
    d.n_W, d.L = n_W, L
    d.residual = 1 if residual else 0
    d.dtype = _DTYPES[W.dtype]
    sf, sb, sr, st = W.stride()
    d.ws = _capi.WStrides(sf, sb, sr, st)
    d.offsets = C.cast(offs_arr, C.POINTER(C.c_int64)) if offs_arr is not None else None
    d.path = path
    d.interleave = int(interleave)
    return d


class _ChordChainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, W, V0, residual, offsets, path, interleave=1):
        global last_path_used
        lib = _capi.load()
        _require_cuda(W, V0)
        if W.dim() != 4 or V0.dim() != 3:
            raise ValueError("W must be [n_W, B, n, L] and V0 [B, n, C]")
        n_W, B, n, L = W.shape
        if V0.shape[0] != B or V0.shape[1] != n:
            raise ValueError(f"shape mismatch: W {tuple(W.shape)} vs V0 {tuple(V0.shape)}")
        if W.dtype not in _DTYPES or V0.dtype != W.dtype:
            raise TypeError("W and V0 must both be float32 or both float64")
        if not _dense_nonoverlapping(W):
            W = W.contiguous()
        V0 = V0.contiguous()
        offs_arr, offs_list = _offsets_array(offsets, L)
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        out = torch.empty_like(V0)
        with torch.cuda.device(W.device):
            d = _desc(W, V0, residual, offs_arr, _PATHS[path], interleave)
            nbytes = lib.psf_chain_workspace_bytes(C.byref(d), 1 if need_grad else 0)
            ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=W.device)
            used = C.c_int(0)
            stream = torch.cuda.current_stream().cuda_stream
            with _Timed("fwd"):
                rc = lib.psf_chain_fwd(C.byref(d), W.data_ptr(), V0.data_ptr(), out.data_ptr(), ws.data_ptr(),
                                       ws.numel(), 1 if need_grad else 0, C.byref(used), stream)
            _capi.check(rc, "psf_chain_fwd")
        last_path_used = _PATH_NAMES.get(used.value, str(used.value))
        if need_grad:
            ctx.save_for_backward(W)
            ctx.ws = ws
            ctx.residual = residual
            ctx.offs_list = offs_list
            ctx.path_used = _PATHS[path] if path == "fused_gen1" else used.value
            ctx.interleave = interleave
            ctx.v_shape = V0.shape
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dOut):
        lib = _capi.load()
        (W,) = ctx.saved_tensors
        dOut = dOut.contiguous()
        L = W.shape[3]
        offs_arr, _ = _offsets_array(ctx.offs_list, L)
        dW = torch.empty_strided(W.shape, W.stride(), dtype=W.dtype, device=W.device)
        dV0 = torch.empty(ctx.v_shape, dtype=W.dtype, device=W.device)
        with torch.cuda.device(W.device):
            V_like = dV0
            # same path as the forward (the saved states are laid out per path)
            d = _desc(W, V_like, ctx.residual, offs_arr, ctx.path_used, ctx.interleave)
            used = C.c_int(0)
            stream = torch.cuda.current_stream().cuda_stream
            with _Timed("bwd"):
                rc = lib.psf_chain_bwd(C.byref(d), W.data_ptr(), dOut.data_ptr(), dW.data_ptr(), dV0.data_ptr(),
                                       ctx.ws.data_ptr(), ctx.ws.numel(), C.byref(used), stream)
            _capi.check(rc, "psf_chain_bwd")
        return dW, dV0, None, None, None, None


def chord_chain(W: torch.Tensor, V0: torch.Tensor, use_residuals: bool = False,
                offsets: Optional[Sequence[int]] = None, path: str = "auto", interleave: int = 1) -> torch.Tensor:
    """Apply ``W[n_W-1] ... W[0]`` to ``V0``.

    W  : logical shape ``[n_W, B, n, L]``, any dense stride order.  The streaming ("tap-major")
         order -- physical ``[n_W, L, B, n]``, see :func:`empty_tap_major` -- is what the fused
         kernel consumes directly; the reference order (physical ``[n_W, B, n, L]``) runs through
         the per-factor path.
    V0 : ``[B, n, C]``.   offsets: tap offsets (default chord: 0, 1, 2, 4, ... mod n).
    interleave : ``Q > 1`` says the row axis of W (and of the returned dW) is stored quad-interleaved
         (include/psf_b200.h; :func:`preferred_interleave` / :func:`row_permutation`), which is what the
         second-generation fused kernel streams.  V0 and the result stay in natural row order.
    Returns ``V_{n_W}`` with ``V_m = W_m V_{m-1} (+ V_0 if use_residuals)``.
    """
    return _ChordChainFn.apply(W, V0, bool(use_residuals), offsets, path, int(interleave))


def attention_map(W: torch.Tensor, offsets: Optional[Sequence[int]] = None, path: str = "auto",
                  interleave: int = 1) -> torch.Tensor:
    """Dense ``[B, n, n]`` matrix ``W[n_W-1] ... W[0]`` of every sample: the chain applied to the identity, as the
    reference's attention-map scripts build it (``W_final = spmm(..., W_final)`` starting from ``torch.eye(n)``,
    LRA/attention_maps/imdb_inference.py:41-59, pathfinder_inference.py:57-81).  No residual, no gradient.
    ``out[b] @ V0[b]`` equals ``chord_chain(W, V0)[b]``.  The fused kernels treat it as ``C = n`` channels."""
    n_W, B, n, L = W.shape
    eye = torch.eye(n, dtype=W.dtype, device=W.device).expand(B, n, n).contiguous()
    with torch.no_grad():
        return chord_chain(W.detach(), eye, False, offsets, path, interleave)


def empty_tap_major(n_W: int, B: int, n: int, L: int, dtype=torch.float32, device="cuda") -> torch.Tensor:
    """Logical ``[n_W, B, n, L]`` view of a physical ``[n_W, L, B, n]`` buffer (row stride 1)."""
    return torch.empty(n_W, L, B, n, dtype=dtype, device=device).permute(0, 2, 3, 1)


def to_tap_major(W: torch.Tensor, interleave: int = 1, src_interleave: int = 1) -> torch.Tensor:
    """Re-lay a logical ``[n_W, B, n, L]`` stack into the streaming order with the library's
    tiled transpose (values unchanged), optionally changing the row order (``interleave`` of the
    result, ``src_interleave`` of the input).  Not differentiable; for benches and tests."""
    lib = _capi.load()
    _require_cuda(W)
    n_W, B, n, L = W.shape
    dst = empty_tap_major(n_W, B, n, L, W.dtype, W.device)
    with torch.cuda.device(W.device):
        rc = lib.psf_w_relayout_ex(_DTYPES[W.dtype], W.data_ptr(), _capi.WStrides(*W.stride()), int(src_interleave),
                                   dst.data_ptr(), _capi.WStrides(*dst.stride()), int(interleave), n_W, B, n, L,
                                   torch.cuda.current_stream().cuda_stream)
        _capi.check(rc, "psf_w_relayout_ex")
    return dst


def row_permutation(n: int, interleave: int, device="cpu") -> torch.Tensor:
    """``perm`` with ``x_interleaved = x_natural[perm]``: the natural row stored at every position of a
    quad-interleaved row axis (include/psf_b200.h: position ``(u % Q)*(n/Q) + 4*(u//Q) + (i & 3)``, ``u = i >> 2``)."""
    Q = int(interleave)
    idx = torch.arange(n, device=device)
    if Q <= 1:
        return idx
    if n % (4 * Q):
        raise ValueError("quad interleave needs n % (4*Q) == 0")
    return idx.view(n // (4 * Q), Q, 4).permute(1, 0, 2).reshape(n)


def _shape_desc(n_W: int, B: int, n: int, L: int, C_: int, residual: bool, interleave: int, path: str):
    d = _capi.ChainDesc()
    d.B, d.n, d.C, d.n_W, d.L, d.residual, d.dtype = B, n, C_, n_W, L, int(residual), _capi.PSF_F32
    d.ws = _capi.WStrides(L * B * n, n, 1, B * n)
    d.offsets = None
    d.path = _PATHS[path]
    d.interleave = int(interleave)
    return d


def preferred_interleave(n_W: int, B: int, n: int, L: int, C_: int = 8) -> int:
    """Row interleave Q the fastest fused kernel wants for this problem (1 = natural order)."""
    lib = _capi.load()
    return int(lib.psf_chain_preferred_interleave(C.byref(_shape_desc(n_W, B, n, L, C_, True, 1, "auto"))))


def fused_generation(n_W: int, B: int, n: int, L: int, C_: int = 8, interleave: int = 1, path: str = "auto") -> int:
    """2 / 1 / 0: which fused kernel generation a tap-major problem with this row order would run."""
    lib = _capi.load()
    return int(lib.psf_chain_fused_generation(C.byref(_shape_desc(n_W, B, n, L, C_, True, interleave, path))))


def chord_chain_list(Ws: Sequence[torch.Tensor], V0: torch.Tensor, use_residuals: bool = False,
                     offsets: Optional[Sequence[int]] = None, path: str = "auto") -> torch.Tensor:
    """Convenience for reference-shaped inputs: ``Ws[m]`` is ``[B, n, L]`` as ``fs[m](data)`` returns."""
    return chord_chain(torch.stack(list(Ws), 0), V0, use_residuals, offsets, path)


def fused_plan(n_W: int, B: int, n: int, L: int, C_: int = 8, residual: bool = True, interleave: int = 1,
               path: str = "auto"):
    """Launch geometry of the fused path for a tap-major problem, or None if unsupported."""
    lib = _capi.load()
    d = _shape_desc(n_W, B, n, L, C_, residual, interleave, path)
    vals = [C.c_int(0) for _ in range(5)]
    rc = lib.psf_chain_fused_plan(C.byref(d), *[C.byref(v) for v in vals])
    if rc != 0:
        return None
    keys = ("cluster", "rows_per_cta", "threads", "smem_bytes", "resident_clusters")
    return dict(zip(keys, (v.value for v in vals)))
