#!/usr/bin/env python
"""Chain-only micro-benchmark (BASELINE cfg5): fused fwd / bwd, CUDA-event timing, achieved
algorithmic GB/s (fwd = 4(S_W + 2 S_V), bwd = 4(2 S_W + 3 S_V) bytes per sequence)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from sparsefactorization_b200 import chain


def alg_bytes(n, n_W, L, C):
    S_W, S_V = n_W * n * L, n * C
    return 4 * (S_W + 2 * S_V), 4 * (2 * S_W + 3 * S_V)


def time_ms(fn, iters, warmup=3):
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
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", "--seq-len", dest="n", type=int, default=16384, help="--seq-len: the spelling that survives torchrun's own option parser")
    ap.add_argument("--B", type=int, default=148)
    ap.add_argument("--C", type=int, default=8)
    ap.add_argument("--n_W", type=int, default=0)
    ap.add_argument("--residual", type=int, default=1)
    ap.add_argument("--path", default="auto")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--peak", type=float, default=6450.3)
    ap.add_argument("--gen", type=int, default=2, help="fused kernel generation: 2 = quad-interleaved W rows, 1 = natural")
    a = ap.parse_args()
    n_W = a.n_W or (a.n - 1).bit_length()
    L = n_W + 1
    rank, world = int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))   # under torchrun: replicas
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    g = torch.Generator(device=dev).manual_seed(0)
    Q = chain.preferred_interleave(n_W, a.B, a.n, L, a.C) if a.gen == 2 and a.path == "auto" else 1
    if a.gen == 1 and a.path == "auto":
        a.path = "fused_gen1"
    W = chain.empty_tap_major(n_W, a.B, a.n, L, device=dev)
    W.copy_(0.25 * torch.randn(n_W, a.B, a.n, L, generator=g, device=dev) / L ** 0.5)
    W.requires_grad_(True)
    V0 = torch.randn(a.B, a.n, a.C, generator=g, device=dev).requires_grad_(True)
    dOut = torch.randn(a.B, a.n, a.C, generator=g, device=dev)
    plan = chain.fused_plan(n_W, a.B, a.n, L, a.C, bool(a.residual), interleave=Q, path=a.path)
    gen = chain.fused_generation(n_W, a.B, a.n, L, a.C, interleave=Q, path=a.path)
    fb, bb = alg_bytes(a.n, n_W, L, a.C)

    def fwd_nograd():
        with torch.no_grad():
            chain.chord_chain(W, V0, bool(a.residual), path=a.path, interleave=Q)

    holder = {}

    def fwd():
        holder["out"] = chain.chord_chain(W, V0, bool(a.residual), path=a.path, interleave=Q)

    def bwd():
        W.grad = None
        V0.grad = None
        holder["out"].backward(dOut, retain_graph=True)

    t_inf, _ = time_ms(fwd_nograd, a.iters)
    t_fwd, _ = time_ms(fwd, a.iters)
    used = chain.last_path_used
    t_bwd, _ = time_ms(bwd, a.iters)
    res = {
        "n": a.n, "B": a.B, "C": a.C, "n_W": n_W, "residual": a.residual, "rank": rank, "world": world, "path": used, "generation": gen, "interleave": Q, "plan": plan,
        "fwd_nograd_ms": round(t_inf, 4), "fwd_ms": round(t_fwd, 4), "bwd_ms": round(t_bwd, 4),
        "fwd_GBps": round(a.B * fb / t_fwd / 1e6, 1), "bwd_GBps": round(a.B * bb / t_bwd / 1e6, 1),
        "fwdbwd_GBps": round(a.B * (fb + bb) / (t_fwd + t_bwd) / 1e6, 1),
        "frac_of_measured_peak": round(a.B * (fb + bb) / (t_fwd + t_bwd) / 1e6 / a.peak, 4),
        "seq_per_s_chain": round(a.B / (t_fwd + t_bwd) * 1e3, 1),
    }
    print(json.dumps(res))


if __name__ == "__main__":
    main()
