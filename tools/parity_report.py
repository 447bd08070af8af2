#!/usr/bin/env python
"""Module-level parity report at the BASELINE shapes: the drop-in PSFNet on the GPU against the oracle module on the CPU
(fp32 and fp64), logits and every parameter gradient, rtol 1e-5 / atol 1e-6 with the float64-anchored fallback
(tests/parity.py).  One JSON line per parameter kind; `python tools/parity_report.py > profiles/r02_parity.jsonl`."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch

import parity
from sparsefactorization_b200 import data as D


def cases(which):
    syn = dict(embedding_size=32, Ws=[32, "GELU"], V=[32, "GELU"], n_channels_V=8, pooling_type="FLATTEN", head=["linear"],
               use_residuals=True)
    out = {}
    out["cfg1"] = ("synthetic", dict(syn, vocab_size=1, add_init_linear_layer=True, n_vec=128, n_W=7, n_class=1,
                                     use_pos_embedding=False, problem="adding"), 40, "adding")
    out["cfg2"] = ("synthetic", dict(syn, vocab_size=1, add_init_linear_layer=True, n_vec=16384, n_W=14, n_class=1,
                                     use_pos_embedding=False, problem="adding"), 40, "adding")
    out["cfg3"] = ("synthetic", dict(syn, vocab_size=6, add_init_linear_layer=False, n_vec=8192, n_W=13, n_class=4,
                                     use_pos_embedding=True, problem="order"), 40, "order")
    out["genome"] = ("genome", dict(vocab_size=5, embedding_size=32, n_vec=2048, n_W=11, Ws=[32, "GELU"], V=[32, "GELU"],
                                    n_channels_V=32, n_class=2, pooling_type="FLATTEN", head=["linear"], use_residuals=True,
                                    dropout1_p=0.0, dropout2_p=0.0, dropout3_p=0.0, init_embedding_weights=False,
                                    use_pos_embedding=True), 8, "tokens")
    return {k: out[k] for k in which}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="cfg1,cfg3,cfg2,genome")
    ap.add_argument("--batch", type=int, default=0)
    a = ap.parse_args()
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    for name, (variant, kw, B, kind) in cases(a.cases.split(",")).items():
        B = a.batch or B
        n = kw["n_vec"]
        if kind == "adding":
            X, Y = D.adding(B, n, seed=1234)
            loss = "mse"
        elif kind == "order":
            X, Y = D.temporal_order(B, n, seed=1234)
            loss = "ce"
        else:
            g = torch.Generator().manual_seed(7)
            X = torch.randint(0, kw["vocab_size"], (B, n), generator=g)
            Y = torch.randint(0, kw["n_class"], (B,), generator=g)
            loss = "ce"
        t0 = time.time()
        recs, info = parity.module_parity(variant, kw, X, Y, loss)
        for r in parity.summarize(recs):
            print(json.dumps(dict(case=name, **info, **r)), flush=True)
        print(json.dumps({"case": name, "all_ok": all(r["ok"] for r in recs), "seconds": round(time.time() - t0, 1)}), flush=True)


if __name__ == "__main__":
    main()
