#!/usr/bin/env python
"""Generator-only micro-benchmark: the fused fs[m] kernels (fwd / bwd) vs the batched-GEMM route, CUDA events."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from sparsefactorization_b200 import _psfnet
from sparsefactorization_b200._psfnet import MLPBlock, generate_factors


def time_ms(fn, iters, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16384)
    ap.add_argument("--B", type=int, default=132)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--gemm", type=int, default=0, help="also time the batched-GEMM route")
    ap.add_argument("--recompute", type=int, default=1, help="1: backward recomputes the pre-activations")
    ap.add_argument("--g", type=int, default=0, help="1: with the V generator (C = 8 outputs), as inside the model")
    ap.add_argument("--interleave", type=int, default=1, help="quad-interleave factor of the rows (the model uses 4 at n = 16384)")
    a = ap.parse_args()
    from sparsefactorization_b200 import generators
    generators.SAVE_PREACTIVATIONS = not a.recompute
    n_W = (a.n - 1).bit_length()
    L = n_W + 1
    torch.manual_seed(0)
    fs = torch.nn.ModuleList([MLPBlock([32, "GELU"], 32, L) for _ in range(n_W)]).cuda()
    data = torch.randn(a.B, a.n, 32, device="cuda", requires_grad=True)
    gW = torch.randn(n_W, L, a.B, a.n, device="cuda").permute(0, 2, 3, 1)
    res = {"n": a.n, "B": a.B, "n_W": n_W}
    if a.g:
        # the model's call: all W generators and the V generator in one launch each way
        g = MLPBlock([32, "GELU"], 32, 8).cuda()
        x2d = data.detach().reshape(a.B * a.n, 32).requires_grad_(True)
        gV = torch.randn(a.B * a.n, 8, device="cuda")
        gWt = torch.randn(n_W, L, a.B * a.n, device="cuda")
        holder = {}

        def fwd_g():
            holder["o"] = generators.fused_generators(fs, x2d, a.n, a.interleave, g)

        def bwd_g():
            for p_ in list(fs.parameters()) + list(g.parameters()):
                p_.grad = None
            x2d.grad = None
            torch.autograd.backward(holder["o"], (gWt, gV), retain_graph=True)

        res["with_g"] = {"interleave": a.interleave, "fwd_ms": round(time_ms(fwd_g, a.iters), 3), "bwd_ms": round(time_ms(bwd_g, a.iters), 3)}
        print(json.dumps(res))
        return
    for fused in ([True, False] if a.gemm else [True]):
        _psfnet.USE_FUSED_GENERATORS = fused
        holder = {}

        def fwd():
            holder["W"] = generate_factors(fs, data)

        def bwd():
            for p in fs.parameters():
                p.grad = None
            data.grad = None
            holder["W"].backward(gW, retain_graph=True)

        tf = time_ms(fwd, a.iters)
        tb = time_ms(bwd, a.iters)
        gf = 2 * a.B * a.n * n_W * (32 * 32 + 32 * L) / 1e9
        key = "fused" if fused else "gemm"
        res[key] = {"fwd_ms": round(tf, 3), "bwd_ms": round(tb, 3), "fwd_TFLOPs": round(gf / tf, 2),
                    "bwd_TFLOPs": round((2 * gf + 2 * a.B * a.n * n_W * 32 * 32 / 1e9) / tb, 2)}
        del holder
    print(json.dumps(res))


if __name__ == "__main__":
    main()
