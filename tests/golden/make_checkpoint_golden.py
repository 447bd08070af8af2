#!/usr/bin/env python
"""Golden fixtures from the PSF checkpoints the reference ships (LRA/attention_maps/*.pt, BSD-3).

Run ONCE in the build container (the reference tree does not exist on the GPU box):

    python tests/golden/make_checkpoint_golden.py

For each checkpoint: build the reference's own ``PSFNet`` (LRA/psf.py, imported unmodified; ``torch_sparse.spmm`` supplied
by the shim of make_golden.py), ``load_state_dict(strict=True)``, run ``eval()`` forward on two seeded sequences on the CPU,
and store the state dict, the input and the logits in ``tests/golden/ckpt_<name>.npz``.  The tests then load the same
tensors into the oracle (CPU) and into the drop-in module (GPU) with ``strict=True`` and compare the logits.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

CKPT_DIR = os.path.join(mg.REF, "LRA/attention_maps")
# LRA/psf_training_config.py:31-117; PSF_5.pt is the 32-wide pathfinder model of LRA/attention_maps/pathfinder_inference.py
BASE = dict(use_cuda=False, dropout1_p=0, dropout2_p=0, dropout3_p=0)
CONFIGS = {
    "pathfinder_epoch27": dict(vocab_size=225, embedding_size=32, n_vec=1024, n_W=11, Ws=[128, 'GELU'], V=[128, 'GELU'],
                               n_channels_V=32, n_class=2, pooling_type='FLATTEN', head=['linear'], use_residuals=False,
                               init_embedding_weights=False, use_pos_embedding=True, problem='pathfinder'),
    "PSF_5": dict(vocab_size=225, embedding_size=32, n_vec=1024, n_W=11, Ws=[32, 'GELU'], V=[32, 'GELU'],
                  n_channels_V=32, n_class=2, pooling_type='FLATTEN', head=['linear'], use_residuals=False,
                  init_embedding_weights=False, use_pos_embedding=True, problem='pathfinder'),
    "imdb_epoch138": dict(vocab_size=97, embedding_size=32, n_vec=4097, n_W=12, Ws=[128, 'GELU'], V=[128, 'GELU'],
                          n_channels_V=32, n_class=2, pooling_type='CLS', head=['linear'], use_residuals=True,
                          init_embedding_weights=True, use_pos_embedding=False, problem='imdb'),
    "cifar10_epoch35": dict(vocab_size=256, embedding_size=16, n_vec=1024, n_W=10, Ws=[16, 'GELU'], V=[16, 'GELU'],
                            n_channels_V=16, n_class=10, pooling_type='FLATTEN', head=['non-linear', 16], use_residuals=False,
                            init_embedding_weights=False, use_pos_embedding=True, problem='cifar10'),
}


def main():
    mg._install_shim()
    ref = mg._load_module(os.path.join(mg.REF, "LRA/psf.py"), "ref_lra_psf_ckpt")
    meta = {}
    for name, cfg in CONFIGS.items():
        sd = torch.load(os.path.join(CKPT_DIR, name + ".pt"), map_location="cpu")
        kw = dict(BASE, **cfg)
        try:
            net = ref.PSFNet(**kw)
            net.load_state_dict(sd, strict=True)
        except Exception as e:      # a config that does not match the checkpoint is a bug in this script
            raise SystemExit("%s: %s" % (name, e))
        net.eval()
        g = torch.Generator().manual_seed(1234)
        X = torch.randint(0, cfg["vocab_size"] - (2 if cfg["pooling_type"] == "CLS" else 0), (2, cfg["n_vec"]), generator=g)
        if cfg["pooling_type"] == "CLS":
            X[:, 0] = cfg["vocab_size"] - 1
        with torch.no_grad():
            logits = net(X)
        out = {"sd::" + k: v.detach().numpy() for k, v in sd.items()}
        out["X"] = X.numpy()
        out["logits"] = logits.numpy()
        np.savez_compressed(os.path.join(HERE, "ckpt_%s.npz" % name), **out)
        meta[name] = {"config": cfg, "logits": logits.tolist(), "n_params": int(sum(v.numel() for v in sd.values()))}
        print(name, "logits", logits.tolist())
    json.dump(meta, open(os.path.join(HERE, "ckpt_meta.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
