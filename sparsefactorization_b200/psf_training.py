"""Training driver: the reference's per-task scripts (SyntheticExperiments/psf_training.py:1-129, LRA/*_training.py,
Genome_Clf/genome_clf_training.py) as ONE data-parallel program, one process per GPU.

    python -m sparsefactorization_b200.psf_training --task adding --n_vec 16384 --data-dir /data/adding
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
        -m sparsefactorization_b200.psf_training --task order --n_vec 8192 --synthetic 4000

Model and loop settings come from the same config dicts the reference reads (``synthetic_training_config.config``,
``lra.psf_training_config.config``, ``genome.genome_training_config.config``); ``batch_size`` is per GPU.  With
``--data-dir`` the tensors are the files the reference's preprocessing scripts write (same names); with ``--synthetic N``
they are N seeded sequences with the reference generators' distributions (``data.py``) - there is no dataset download here.
Every rank holds the full tensors on its GPU and trains on its contiguous shard (``dp.TrainModelDP``); rank 0 prints the
reference's per-epoch lines and writes ``{problem}_epoch{e}_acc{a}.pt`` under the reference's saving rule.
"""
from __future__ import annotations

import argparse
import os

import torch
from torch import nn, optim

SYNTHETIC_TASKS = ("adding", "order")
LRA_TASKS = ("listops", "cifar10", "pathfinder", "imdb")
GENOME_TASKS = ("genome",)


def model_settings(task: str, n_vec: int | None = None):
    """(family, PSFNet keyword arguments, training settings) of a task, from the reference's config dicts."""
    if task in SYNTHETIC_TASKS:
        from .synthetic_training_config import config
        m, tr = config[task]["models"]["PSF"], config[task]["training"]
        if not n_vec or n_vec & (n_vec - 1):
            raise ValueError("the synthetic tasks take --n_vec, a power of two (2**7 .. 2**14 in the reference)")
        # psf_training.py:28-45: dim -> embedding_size, n_W = log2(n_vec)
        kw = dict(vocab_size=m["vocab_size"], add_init_linear_layer=m["add_init_linear_layer"], embedding_size=m["dim"],
                  n_vec=n_vec, n_W=n_vec.bit_length() - 1, Ws=m["Ws"], V=m["V"], n_channels_V=m["n_channels_V"],
                  n_class=m["n_class"], pooling_type=m["pooling_type"], head=m["head"], use_cuda=m["use_cuda"],
                  use_residuals=m["use_residuals"], use_pos_embedding=m["use_pos_embedding"], problem=m["problem"])
        return "synthetic", kw, tr
    if task in LRA_TASKS:
        from .lra.psf_training_config import config
        return "lra", dict(config[task]["model"]), config[task]["training"]
    if task in GENOME_TASKS:
        from .genome.genome_training_config import config
        m = dict(config["DDcDNA"]["PSF"])
        m.pop("name")
        return "genome", m, config["DDcDNA"]["training"]
    raise ValueError("unknown task %r" % task)


def data_files(task: str, n_vec: int | None):
    """File names of the (train, val, test) x (data, target) tensors as the reference scripts load them."""
    if task in SYNTHETIC_TASKS:                       # psf_training.py:63-70
        return [("%s_%d_%s.pt" % (task, n_vec, s), "%s_%d_%s_target.pt" % (task, n_vec, s)) for s in ("train", "val", "test")]
    if task in GENOME_TASKS:                          # genome_clf_training.py:121-128
        return [("DDcDNA16384_%s.pt" % s, "DDcDNA16384_%s_targets.pt" % s) for s in ("train", "val", "test")]
    return [("%s_clean.pt" % s, "target_%s_clean.pt" % s) for s in ("train", "val", "test")]       # listops_training.py:56-63


def make_net(family: str, kw: dict):
    if family == "synthetic":
        from .psf import PSFNet
    elif family == "lra":
        from .lra.psf import PSFNet
    else:
        from .genome.psf import PSFNet
    return PSFNet(**kw)


def load_split(task, family, kw, data_dir, synthetic, device, seed):
    """[(X, Y)] x 3 on ``device``."""
    if data_dir:
        out = []
        for fx, fy in data_files(task, kw["n_vec"]):
            X, Y = torch.load(os.path.join(data_dir, fx)), torch.load(os.path.join(data_dir, fy))
            if family != "synthetic" or task == "order":
                X, Y = X.to(torch.int64), Y.to(torch.int64)
            out.append((X.to(device), Y.to(device)))
        return out
    from . import data as synth
    sizes = (synthetic, max(synthetic // 5, 1), max(synthetic // 5, 1))
    out = []
    for i, cnt in enumerate(sizes):
        if task == "adding":
            out.append(synth.adding(cnt, kw["n_vec"], seed + i, device))
        elif task == "order":
            out.append(synth.temporal_order(cnt, kw["n_vec"], seed + i, device))
        else:
            out.append(synth.listops_shaped(cnt, kw["n_vec"], kw["vocab_size"], kw["n_class"], seed + i, device))
    return out


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--task", required=True, choices=SYNTHETIC_TASKS + LRA_TASKS + GENOME_TASKS)
    ap.add_argument("--n_vec", type=int, default=None, help="sequence length of the synthetic tasks")
    ap.add_argument("--data-dir", default=None, help="directory with the reference's preprocessed .pt tensors")
    ap.add_argument("--synthetic", type=int, default=0, help="train on this many generated sequences instead")
    ap.add_argument("--epochs", type=int, default=None, help="override num_train_steps")
    ap.add_argument("--batch-size", type=int, default=None, help="override the per-GPU batch size")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--saving-criteria", type=float, default=99.5)
    a = ap.parse_args(argv)
    if not a.data_dir and a.synthetic <= 0:
        ap.error("give --data-dir or --synthetic N")

    from . import dp
    from .psf_utils import count_params, seed_everything
    family, kw, tr = model_settings(a.task, a.n_vec)
    local = int(os.environ.get("LOCAL_RANK", tr["device_id"]))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not dp.is_dist():
        torch.distributed.init_process_group("nccl", device_id=device)
    seed_everything(a.seed)
    net = make_net(family, kw).cuda()
    if dp.rank() == 0:
        print(torch.cuda.get_device_name(local), "x", dp.world_size())
        print('Number of trainable parameters', count_params(net))
    optimizer = optim.Adam(net.parameters(), lr=tr["learning_rate"])
    loss = nn.MSELoss() if a.task == "adding" else nn.CrossEntropyLoss()
    train, val, test = load_split(a.task, family, kw, a.data_dir, a.synthetic, device, a.seed)
    problem = kw.get("problem") or a.task
    history = dp.TrainModelDP(net, train, val, test, n_epochs=a.epochs or tr["num_train_steps"], test_freq=tr["eval_frequency"],
                              optimizer=optimizer, loss=loss, problem=problem if problem == "adding" else (problem or "order"),
                              saving_criteria=a.saving_criteria, batch_size=a.batch_size or tr["batch_size"],
                              clip_grad_norm=1.0 if family == "genome" else None)     # Genome_Clf/psf_utils.py:73
    if dp.is_dist():
        torch.distributed.destroy_process_group()
    return history


if __name__ == "__main__":
    main()
