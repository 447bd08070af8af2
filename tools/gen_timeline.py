#!/usr/bin/env python
"""Timeline of block 0 of the backward generator kernel (PSF_TC_DEBUG build): per event the SM clock relative to the start.
    python -m sparsefactorization_b200.build --out sparsefactorization_b200/_lib/dbg.so -DPSF_TC_DEBUG
    PSF_B200_LIB=.../dbg.so python tools/gen_timeline.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sparsefactorization_b200 import _capi, generators
from sparsefactorization_b200._psfnet import MLPBlock

NAMES = {1: "epi wait PD", 2: "epi got PD", 3: "epi math done", 4: "epi got BFREE", 5: "epi published",
         10: "bld wait DS", 11: "bld got DS", 12: "bld built", 13: "drn wait GRAD", 14: "drn got GRAD", 15: "drn done",
         20: "mma wait DPRE", 21: "mma got DPRE", 22: "mma DX issued", 23: "mma PD issued", 24: "mma got B/AT/DFREE", 25: "mma GRAD issued"}


def main():
    n, B, n_W = 16384, 8, 14
    L = n_W + 1
    torch.manual_seed(0)
    fs = torch.nn.ModuleList([MLPBlock([32, "GELU"], 32, L) for _ in range(n_W)]).cuda()
    x = torch.randn(B * n, 32, device="cuda", requires_grad=True)
    W, _ = generators.fused_generators(fs, x, n, 1, None)
    W.backward(torch.randn_like(W))
    torch.cuda.synchronize()
    lib = _capi.load()
    per = 2048
    buf = (C.c_longlong * (24 * per * 2))()
    lib.psf_debug_timeline.restype = C.c_int
    lib.psf_debug_timeline(buf)
    ev = []
    for w in range(24):
        for i in range(per):
            code, clk = buf[(w * per + i) * 2], buf[(w * per + i) * 2 + 1]
            if code:
                ev.append((clk, w, code // 100000, code % 100000 - 1))
    ev.sort()
    t0 = ev[0][0]
    lo, hi = int(sys.argv[1]) if len(sys.argv) > 1 else 20, int(sys.argv[2]) if len(sys.argv) > 2 else 24
    for t, warp, e, step in ev:
        if lo <= step < hi and warp in (0, 16, 17, 18, 20):
            print(f"{t - t0:9d}  warp {warp:2d}  step {step:3d}  {NAMES.get(e, e)}")


if __name__ == "__main__":
    main()
