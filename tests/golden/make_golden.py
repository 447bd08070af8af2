#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the reference at /root/reference.

Run ONCE in the build container (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py

What it pins:
  * chord_indices.npz ....... output of the reference's own ``get_chord_indices_assym``
    (SyntheticExperiments/psf.py:7-32), executed from its source; small cases stored in
    full, every case also as a sha256 of the int64 little-endian bytes.
  * psfnet_*.npz ............ the reference's own ``PSFNet`` modules
    (SyntheticExperiments/psf.py, LRA/psf.py), imported unmodified, run forward + backward
    on seeded CPU inputs.  Their only un-vendored dependency on this path,
    ``torch_sparse.spmm`` (torch-sparse==0.6.11, requirements.txt:146), is supplied by the
    shim below that restates the package's published algorithm.  The fixtures hold the
    state dict, inputs, logits, loss, all parameter grads, and the chain-level tensors
    (W_m, V0, chain output and their grads) captured with hooks.
  * checkpoint_manifest.json  parameter names/shapes of the PSF checkpoints shipped in
    LRA/attention_maps/ (state-dict compatibility target).
"""
import ast
import hashlib
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


# ---- torch_sparse shim: published algorithm of torch_sparse.spmm 0.6.11 --------------------
def _shim_spmm(index, value, m, n, matrix):
    assert n == matrix.size(-2)
    row, col = index[0], index[1]
    matrix = matrix if matrix.dim() > 1 else matrix.unsqueeze(-1)
    out = matrix.index_select(-2, col)
    out = out * value.unsqueeze(-1)
    shape = list(out.shape)
    shape[-2] = m
    idx = row.view([1] * (out.dim() - 2) + [-1, 1]).expand_as(out)
    return torch.zeros(shape, dtype=out.dtype).scatter_add_(-2, idx, out)


def _install_shim():
    mod = types.ModuleType("torch_sparse")
    mod.spmm = _shim_spmm
    sys.modules["torch_sparse"] = mod


def _load_module(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _ref_chord_fn():
    """Exec only ``get_chord_indices_assym`` out of the reference source."""
    src = open(os.path.join(REF, "SyntheticExperiments/psf.py")).read()
    tree = ast.parse(src)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "get_chord_indices_assym"][0]
    ns = {}
    exec("import itertools\n" + ast.get_source_segment(src, fn), ns)
    return ns["get_chord_indices_assym"]


def make_chord_indices():
    fn = _ref_chord_fn()
    cases = [(8, 4), (16, 5), (100, 8), (128, 8), (1000, 11), (1024, 11), (1024, 12), (2000, 12),
             (4096, 13), (4097, 13), (8192, 14), (16384, 15), (32768, 16), (65536, 17)]
    store, digests = {}, {}
    for n, L in cases:
        rows, cols = fn(n, L)
        arr = np.asarray([rows, cols], dtype="<i8")
        digests[f"{n}_{L}"] = hashlib.sha256(arr.tobytes()).hexdigest()
        if n <= 1024:
            store[f"idx_{n}_{L}"] = arr
    np.savez_compressed(os.path.join(OUT, "chord_indices.npz"), **store)
    json.dump(digests, open(os.path.join(OUT, "chord_indices_sha256.json"), "w"), indent=1, sort_keys=True)
    print("chord indices:", len(cases), "cases")


def _run_case(name, net, X, Y, loss_fn, squeeze):
    """Forward + backward through the reference module, capturing chain-level tensors."""
    cap = {"W": [None] * len(net.fs)}

    def hook_f(m_idx):
        def h(_mod, _inp, out):
            out.retain_grad()
            cap["W"][m_idx] = out
        return h

    def hook_g(_mod, _inp, out):
        out.retain_grad()
        cap["V0"] = out

    def hook_final(_mod, inp):
        inp[0].retain_grad()
        cap["final_in"] = inp[0]

    hs = [f.register_forward_hook(hook_f(i)) for i, f in enumerate(net.fs)]
    hs.append(net.g.register_forward_hook(hook_g))
    hs.append(net.final.register_forward_pre_hook(hook_final))
    net.zero_grad()
    pred = net(X)
    loss = loss_fn(pred.squeeze() if squeeze else pred, Y)
    loss.backward()
    for h in hs:
        h.remove()
    out = {"X": X.numpy(), "Y": Y.numpy(), "pred": pred.detach().numpy(), "loss": loss.detach().numpy()}
    for k, v in net.state_dict().items():
        out["param/" + k] = v.numpy()
    for k, p in net.named_parameters():
        if p.grad is not None:
            out["grad/" + k] = p.grad.numpy()
    out["chain/W"] = np.stack([w.detach().numpy() for w in cap["W"]])          # [n_W, B, n, L]
    out["chain/dW"] = np.stack([w.grad.numpy() for w in cap["W"]])
    out["chain/V0"] = cap["V0"].detach().numpy()                                # [B, n, C] (pre-dropout2; p=0 here)
    out["chain/dV0"] = cap["V0"].grad.numpy()
    out["chain/final_in"] = cap["final_in"].detach().numpy()
    out["chain/dfinal_in"] = cap["final_in"].grad.numpy()
    np.savez_compressed(os.path.join(OUT, f"psfnet_{name}.npz"), **out)
    print(f"psfnet_{name}: loss={float(loss):.6f} pred[0]={pred.detach().flatten()[:3].tolist()}")


def make_psfnet_goldens():
    _install_shim()
    syn = _load_module(os.path.join(REF, "SyntheticExperiments/psf.py"), "ref_syn_psf")
    lra = _load_module(os.path.join(REF, "LRA/psf.py"), "ref_lra_psf")

    # --- cfg1: Adding, n=128, literal config of synthetic_training_config.py:4-18 --------
    torch.manual_seed(42)
    n = 128
    net = syn.PSFNet(vocab_size=1, add_init_linear_layer=True, embedding_size=32, n_vec=n, n_W=7,
                     Ws=[32, "GELU"], V=[32, "GELU"], n_channels_V=8, n_class=1, pooling_type="FLATTEN",
                     head=["linear"], use_cuda=False, use_residuals=True, use_pos_embedding=False,
                     problem="adding")
    g = torch.Generator().manual_seed(1234)
    B = 4
    X = torch.zeros(B, n, 2)
    X[:, :, 0] = 1 - 2 * torch.rand(B, n, generator=g)
    Y = torch.zeros(B)
    for b in range(B):
        p = torch.randperm(n, generator=g)[:2]
        X[b, p, 1] = 1
        Y[b] = 0.5 + X[b, p, 0].sum() / 4
    _run_case("adding_n128", net, X, Y, torch.nn.MSELoss(), squeeze=True)

    # --- Temporal order, n=64, pos embedding (synthetic_training_config.py:72-86) ---------
    torch.manual_seed(43)
    n = 64
    net = syn.PSFNet(vocab_size=6, add_init_linear_layer=False, embedding_size=32, n_vec=n, n_W=6,
                     Ws=[32, "GELU"], V=[32, "GELU"], n_channels_V=8, n_class=4, pooling_type="FLATTEN",
                     head=["linear"], use_cuda=False, use_residuals=True, use_pos_embedding=True,
                     problem="order")
    X = torch.randint(0, 4, (3, n, 1), generator=g)
    X[:, 5, 0] = 4
    X[:, 40, 0] = 5
    Y = torch.tensor([1, 1, 1])
    _run_case("order_n64", net, X, Y, torch.nn.CrossEntropyLoss(), squeeze=True)

    # --- LRA ListOps-shaped: non-power-of-two n, CLS pooling, no residual ------------------
    torch.manual_seed(44)
    n = 100
    net = lra.PSFNet(vocab_size=17, embedding_size=24, n_vec=n, n_W=7, Ws=[16, "GELU"], V=[16, "GELU"],
                     n_channels_V=16, n_class=10, pooling_type="CLS", head=["linear"], use_cuda=False,
                     use_residuals=False, dropout1_p=0, dropout2_p=0, dropout3_p=0,
                     init_embedding_weights=False, use_pos_embedding=True, problem="listops")
    X = torch.randint(0, 15, (3, n), generator=g)
    X[:, 0] = 16
    X[:, 70:] = 15
    Y = torch.tensor([3, 0, 9])
    _run_case("listops_n100", net, X, Y, torch.nn.CrossEntropyLoss(), squeeze=False)

    # --- LRA Pathfinder-shaped: duplicate tap (2^(L-2) == n), FLATTEN, non-linear head -----
    torch.manual_seed(45)
    n = 64
    net = lra.PSFNet(vocab_size=225, embedding_size=16, n_vec=n, n_W=7, Ws=[16, "GELU"], V=[16, "GELU"],
                     n_channels_V=8, n_class=2, pooling_type="FLATTEN", head=["non-linear", 12],
                     use_cuda=False, use_residuals=False, dropout1_p=0, dropout2_p=0, dropout3_p=0,
                     init_embedding_weights=True, use_pos_embedding=True, problem="pathfinder")
    X = torch.randint(0, 225, (2, n), generator=g)
    Y = torch.tensor([1, 0])
    _run_case("pathfinder_n64_duptap", net, X, Y, torch.nn.CrossEntropyLoss(), squeeze=False)


def make_checkpoint_manifest():
    base = os.path.join(REF, "LRA/attention_maps")
    man = {}
    for fn in ["PSF_5.pt", "cifar10_epoch35.pt", "imdb_epoch138.pt", "pathfinder_epoch27.pt"]:
        sd = torch.load(os.path.join(base, fn), map_location="cpu", weights_only=True)
        man[fn] = {k: list(v.shape) for k, v in sd.items()}
    json.dump(man, open(os.path.join(OUT, "checkpoint_manifest.json"), "w"), indent=1, sort_keys=True)
    print("checkpoint manifest:", {k: len(v) for k, v in man.items()})


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; fixtures can only be regenerated in the build container")
    make_chord_indices()
    make_psfnet_goldens()
    make_checkpoint_manifest()
