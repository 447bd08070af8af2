"""All generators of a PSFNet in one fused CUDA kernel each way.

Replaces the per-factor ``fs[m](data)`` calls and ``g(data)`` of the reference
(SyntheticExperiments/psf.py:110-126, :165, :175): ``Linear(E,H) -> GELU -> Linear(H,out)`` for every
W generator and for the V generator, writing W directly in the chain's streaming (tap-major) layout and
V0 in its natural ``[B, n, C]`` layout.  The GEMM-shaped pieces run on the tcgen05 tensor cores as error-compensated
3xTF32 with the A operands in tensor memory (csrc/psf_gen_tc_sm100.cu, psf_gen_tc_bwd_sm100.cu; ``PSF_GEN_TC=0`` selects
the FP32-FMA / mma.sync kernels of csrc/psf_gen_sm100.cu); the backward pass recomputes the hidden activations.  Shapes
the kernels do not cover (E or H != 32, more than 16 outputs, other layer stacks) stay on the batched-GEMM path of
``_psfnet.generate_factors``.  CUDA only.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch
from torch import nn

from . import _capi

E_FUSED, H_FUSED, OUT_MAX = 32, 32, 16
SAVE_PREACTIVATIONS = False   # True: the forward pass saves them ([n_gen, 32, rows], 4.2 GB at the headline shape) and
                              # the backward pass reads them back instead of recomputing; measured no faster


class _Stats:
    """Launch counter and optional CUDA-event timing of the generator kernels (used by bench.py)."""

    def __init__(self):
        self.reset()

    def reset(self, timing: bool = False):
        self.launches = 0
        self.timing = timing
        self.events = {"fwd": [], "bwd": []}

    def mean_ms(self, kind: str) -> float:
        ev = self.events[kind]
        return sum(a.elapsed_time(b) for a, b in ev) / len(ev) if ev else float("nan")


stats = _Stats()


class _Timed:
    def __init__(self, kind, launches):
        self.kind, self.n = kind, launches

    def __enter__(self):
        stats.launches += self.n
        if stats.timing:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if stats.timing:
            self.b.record()
            stats.events[self.kind].append((self.a, self.b))
        return False


def _mlp_ok(f: nn.Module) -> Optional[int]:
    """out width if ``f.network`` is Linear(32,32) -> exact GELU -> Linear(32, out <= 16) in fp32, else None."""
    net = getattr(f, "network", None)
    if net is None or len(net) != 3:
        return None
    a, act, b = net
    if not (isinstance(a, nn.Linear) and isinstance(act, nn.GELU) and isinstance(b, nn.Linear)):
        return None
    if getattr(act, "approximate", "none") != "none" or a.bias is None or b.bias is None:
        return None
    if (a.in_features, a.out_features, b.in_features) != (E_FUSED, H_FUSED, H_FUSED) or b.out_features > OUT_MAX:
        return None
    if a.weight.dtype != torch.float32 or not a.weight.is_cuda:
        return None
    return b.out_features


def fusable_modules(fs: Sequence[nn.Module], g: Optional[nn.Module] = None) -> bool:
    """True when every generator (and ``g`` if given) has the fused kernel's shape (fp32 CUDA parameters)."""
    if len(fs) < 1:
        return False
    outs = {_mlp_ok(f) for f in fs}
    if len(outs) != 1 or None in outs:
        return False
    return g is None or _mlp_ok(g) is not None


def fusable(fs: Sequence[nn.Module], data: torch.Tensor, g: Optional[nn.Module] = None) -> bool:
    """``fusable_modules`` and fp32 CUDA data of width 32."""
    if not (data.is_cuda and data.dtype == torch.float32 and data.shape[-1] == E_FUSED):
        return False
    return fusable_modules(fs, g)


def _ptrs(ts):
    return _capi.MlpPtrs(*[t.data_ptr() if t is not None else None for t in ts])


class _FusedGenerators(torch.autograd.Function):
    """x [rows,32]; fs params stacked (w1 [n_W,32,32], b1 [n_W,32], w2 [n_W,L,32], b2 [n_W,L]); optional g params
    ->  W [n_W, L, rows] (row order of x) and V0 [rows, C] (natural row order; empty without g)."""

    @staticmethod
    def forward(ctx, x, n, interleave, w1, b1, w2, b2, gw1, gb1, gw2, gb2):
        lib = _capi.load()
        x, w1, b1, w2, b2 = (t.contiguous() for t in (x, w1, b1, w2, b2))
        has_g = gw1 is not None
        gs = tuple(t.contiguous() for t in (gw1, gb1, gw2, gb2)) if has_g else (None,) * 4
        if has_g and any(t.data_ptr() % 16 for t in gs):      # the kernels stage weights in 16-byte pieces
            gs = tuple(t.clone() for t in gs)
        rows = x.shape[0]
        n_W, L = w2.shape[0], w2.shape[1]
        Cg = gs[2].shape[0] if has_g else 0
        d = _capi.GenDesc(rows, n, n_W, E_FUSED, H_FUSED, L, L * rows, rows, int(interleave), Cg)
        out = torch.empty(n_W, L, rows, dtype=torch.float32, device=x.device)
        v0 = torch.empty(rows, Cg, dtype=torch.float32, device=x.device)
        # pre-activations for the backward pass (saves it the recompute): only when a gradient will be asked for
        pre = None
        if SAVE_PREACTIVATIONS and any(ctx.needs_input_grad):
            pre = torch.empty(n_W + (1 if has_g else 0), H_FUSED, rows, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device), _Timed("fwd", 1):
            rc = lib.psf_generators_fwd(C.byref(d), x.data_ptr(), _ptrs((w1, b1, w2, b2)), _ptrs(gs), out.data_ptr(),
                                        v0.data_ptr() if has_g else None, pre.data_ptr() if pre is not None else None,
                                        torch.cuda.current_stream().cuda_stream)
        _capi.check(rc, "psf_generators_fwd")
        ctx.save_for_backward(x, w1, b1, w2, b2, *(gs if has_g else ()))
        ctx.pre = pre
        ctx.desc = d
        ctx.has_g = has_g
        return out, v0

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout, dv0):
        lib = _capi.load()
        saved = ctx.saved_tensors
        x, w1, b1, w2, b2 = saved[:5]
        gs = saved[5:] if ctx.has_g else (None,) * 4
        dout = dout.contiguous()
        dv0 = dv0.contiguous() if ctx.has_g else None
        dx = torch.empty_like(x)
        dfs = tuple(torch.empty_like(t) for t in (w1, b1, w2, b2))
        dgs = tuple(torch.empty_like(t) for t in gs) if ctx.has_g else (None,) * 4
        d = ctx.desc
        with torch.cuda.device(x.device):
            nbytes = int(lib.psf_generators_workspace_bytes(d.n_W))
            ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=x.device)
            with _Timed("bwd", 3):       # weight-image prep + backward kernel + the fixed-order reduction of the per-CTA slices
                rc = lib.psf_generators_bwd(C.byref(d), x.data_ptr(), _ptrs((w1, b1, w2, b2)), _ptrs(gs),
                                            dout.data_ptr(), dv0.data_ptr() if ctx.has_g else None,
                                            ctx.pre.data_ptr() if ctx.pre is not None else None, dx.data_ptr(),
                                            _ptrs(dfs), _ptrs(dgs), ws.data_ptr(), ws.numel(),
                                            torch.cuda.current_stream().cuda_stream)
        _capi.check(rc, "psf_generators_bwd")
        ctx.pre = None
        return (dx, None, None, *dfs, *dgs)


def fused_generators(fs: Sequence[nn.Module], x2d: torch.Tensor, n: int, interleave: int = 1,
                     g: Optional[nn.Module] = None):
    """``W[m, l, row] = fs[m](x2d)[row, l]`` for every generator (x2d ``[rows, 32]``, result ``[n_W, L, rows]``) and,
    with ``g``, ``V0 = g(x2d)`` as ``[rows, C]`` in natural row order (x2d rows are quad-interleaved per sample of
    ``n`` rows when ``interleave > 1``)."""
    w1 = torch.stack([f.network[0].weight for f in fs], 0)
    b1 = torch.stack([f.network[0].bias for f in fs], 0)
    w2 = torch.stack([f.network[2].weight for f in fs], 0)
    b2 = torch.stack([f.network[2].bias for f in fs], 0)
    gp = (g.network[0].weight, g.network[0].bias, g.network[2].weight, g.network[2].bias) if g is not None else (None,) * 4
    W, V0 = _FusedGenerators.apply(x2d, n, interleave, w1, b1, w2, b2, *gp)
    return W, (V0 if g is not None else None)
