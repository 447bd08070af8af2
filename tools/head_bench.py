#!/usr/bin/env python
"""Micro-benchmark of the two small streaming operators around the chain: init_linear (csrc/psf_embed.cu) and the FLATTEN
head (csrc/psf_head.cu), each against the torch / cuBLAS route on the same tensors (CUDA events, inputs larger than L2 are
flushed between iterations by a 256 MB fill whose time is subtracted)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch import nn

from sparsefactorization_b200 import _psfnet, head


def t_us(fn, it=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / it * 1e3


def main():
    B, n = 132, 16384
    big = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
    flush = t_us(lambda: big.zero_())
    lin = nn.Linear(2, 32).cuda()
    x = torch.randn(B, n, 2, device="cuda")
    g = torch.randn(B, n, 32, device="cuda")
    for name, f in (("init_linear kernels", lambda: _psfnet.skinny_linear(lin, x)), ("init_linear torch", lambda: lin(x))):
        def fb():
            lin.zero_grad()
            big.zero_()
            f().backward(g)
        print(name, "fwd us", round(t_us(f), 1), "fwd+bwd us", round(t_us(fb) - flush, 1))
    K = n * 8
    fin = nn.Linear(K, 1).cuda()
    V = torch.randn(B, K, device="cuda", requires_grad=True)
    gl = torch.randn(B, 1, device="cuda")
    for name, f in (("head kernels", lambda: head.flat_linear(fin, V)), ("head torch", lambda: fin(V))):
        def fb():
            V.grad = None
            fin.zero_grad()
            big.zero_()
            f().backward(gl)
        print(name, "fwd us", round(t_us(f), 1), "fwd+bwd us", round(t_us(fb) - flush, 1))


if __name__ == "__main__":
    main()
