#!/usr/bin/env python
"""Small fused forward+backward cases for compute-sanitizer (memcheck / racecheck), checked against the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import oracle
from sparsefactorization_b200 import chain

torch.manual_seed(0)
for (B, n, n_W, res) in [(2, 128, 7, True), (3, 516, 9, False), (2, 4096, 4, True)]:
    L = n_W + 1
    W = 0.2 * torch.randn(n_W, B, n, L)
    V0 = torch.randn(B, n, 8)
    dOut = torch.randn(B, n, 8)
    Wc, Vc = W.clone().requires_grad_(True), V0.clone().requires_grad_(True)
    idx = torch.tensor(oracle.get_chord_indices_assym(n, L))
    want = oracle.chain_ref(idx, [Wc[m] for m in range(n_W)], Vc, n, res)
    want.backward(dOut)
    Wd = chain.to_tap_major(W.cuda()).requires_grad_(True)
    Vd = V0.cuda().requires_grad_(True)
    out = chain.chord_chain(Wd, Vd, res)
    out.backward(dOut.cuda())
    torch.cuda.synchronize()
    assert chain.last_path_used == "fused"
    np.testing.assert_allclose(out.detach().cpu().numpy(), want.detach().numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(Wd.grad.cpu().numpy(), Wc.grad.numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(Vd.grad.cpu().numpy(), Vc.grad.numpy(), rtol=1e-4, atol=1e-4)
    print("ok", B, n, n_W, res, flush=True)
