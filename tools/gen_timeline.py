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

NAMES = {1: "epi wait PD", 2: "epi got PD", 3: "epi math done", 4: "epi got BFREE", 5: "epi published", 6: "epi STS issued", 7: "epi proxy fence done", 8: "epi tmem st issued", 9: "epi tmem st done",
         10: "bld wait DS", 11: "bld got DS", 12: "bld built", 13: "drn wait GRAD", 14: "drn got GRAD", 15: "drn done",
         20: "mma wait DPRE", 21: "mma got DPRE", 22: "mma DX issued", 23: "mma PD issued", 24: "mma got B/AT/DFREE", 25: "mma GRAD issued", 26: "mma got ATD", 30: "pd got X", 31: "pd got W", 32: "pd got DOUTT", 40: "epi got PD(s+1)"}


WARPS = tuple(int(w) for w in os.environ.get('TL_WARPS', '0,16,17,18,20').split(','))


def main():
    n, B, n_W = 16384, int(os.environ.get('TL_B', '16')), 14
    L = n_W + 1
    torch.manual_seed(0)
    fs = torch.nn.ModuleList([MLPBlock([32, "GELU"], 32, L) for _ in range(n_W)]).cuda()
    x = torch.randn(B * n, 32, device="cuda", requires_grad=True)
    W, _ = generators.fused_generators(fs, x, n, 1, None)
    W.backward(torch.randn_like(W))
    torch.cuda.synchronize()
    lib = _capi.load()
    per = 4096
    buf = (C.c_uint32 * (24 * per * 2))()
    lib.psf_debug_timeline.restype = C.c_int
    lib.psf_debug_timeline(buf)
    ev = []
    for w in range(24):
        for i in range(per):
            code, clk = buf[(w * per + i) * 2], buf[(w * per + i) * 2 + 1]
            if code:
                ev.append((clk, w, code // 100000, code % 100000 - 1))
    ev.sort()
    if os.environ.get("TL_SUMMARY", "1") == "1":
        # mean time between consecutive events of each warp, over all steps: where every role spends its step
        from collections import defaultdict
        for w in WARPS:
            seq = [(t, e, st) for t, ww, e, st in ev if ww == w]
            acc, cnt = defaultdict(float), defaultdict(int)
            for (t1, e1, s1), (t2, e2, s2) in zip(seq, seq[1:]):
                if s1 < 4:
                    continue
                acc[(e1, e2)] += (t2 - t1) & 0xffffffff
                cnt[(e1, e2)] += 1
            steps = max(cnt.values()) if cnt else 1
            tot = sum(acc.values()) / steps
            print(f"warp {w}: {tot:.0f} cycles per step over {steps} steps")
            for k in sorted(acc, key=lambda k: -acc[k]):
                if cnt[k] >= steps // 4:
                    print(f"    {NAMES.get(k[0], k[0]):>22s} -> {NAMES.get(k[1], k[1]):<22s} {acc[k] / cnt[k]:8.0f} x {cnt[k]}")
        return
    t0 = ev[0][0]
    lo, hi = int(sys.argv[1]) if len(sys.argv) > 1 else 20, int(sys.argv[2]) if len(sys.argv) > 2 else 24
    for t, warp, e, step in ev:
        if lo <= step < hi and warp in WARPS:
            print(f"{t - t0:9d}  warp {warp:2d}  step {step:3d}  {NAMES.get(e, e)}")


if __name__ == "__main__":
    main()
