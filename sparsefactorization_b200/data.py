"""Synthetic inputs with the distributions of the reference's generators, seeded and produced
directly on the target device (``SyntheticExperiments/synth_data_generation.py:8-70``; the
reference generator is an unseeded Python loop writing .pt files).  ListOps-shaped data follows
the PAD / CLS conventions of LRA/listops_training.py:65-72 and LRA/psf_training_config.py:4.
"""
from __future__ import annotations

import torch


def _two_positions(n_seq: int, n: int, gen: torch.Generator, device) -> torch.Tensor:
    # two distinct positions per sequence, uniformly chosen
    p1 = torch.randint(0, n, (n_seq,), generator=gen, device=device)
    p2 = torch.randint(0, n - 1, (n_seq,), generator=gen, device=device)
    p2 = p2 + (p2 >= p1).long()
    return torch.stack([torch.minimum(p1, p2), torch.maximum(p1, p2)], 1)


def adding(n_seq: int, n: int, seed: int = 1234, device="cpu"):
    """X [n_seq, n, 2] fp32: column 0 ~ U(-1, 1), column 1 marks two positions; Y = 0.5 + (x1+x2)/4."""
    gen = torch.Generator(device=device).manual_seed(seed)
    x = 1 - 2 * torch.rand(n_seq, n, generator=gen, device=device)
    pos = _two_positions(n_seq, n, gen, device)
    mark = torch.zeros(n_seq, n, device=device)
    mark.scatter_(1, pos, 1.0)
    X = torch.stack([x, mark], -1)
    Y = 0.5 + x.gather(1, pos).sum(1) / 4
    return X, Y


def temporal_order(n_seq: int, n: int, seed: int = 1234, device="cpu"):
    """X [n_seq, n, 1] int64 tokens {0..3} with two markers from {4, 5}; label = 2*[v1==5] + [v2==5]."""
    gen = torch.Generator(device=device).manual_seed(seed)
    X = torch.randint(0, 4, (n_seq, n), generator=gen, device=device)
    pos = _two_positions(n_seq, n, gen, device)
    vals = torch.randint(4, 6, (n_seq, 2), generator=gen, device=device)
    X.scatter_(1, pos, vals)
    Y = 2 * (vals[:, 0] == 5).long() + (vals[:, 1] == 5).long()
    return X.unsqueeze(-1), Y


def listops_shaped(n_seq: int, n: int, vocab_size: int = 17, n_class: int = 10, seed: int = 1234, device="cpu"):
    """X [n_seq, n] int64: CLS (= vocab-1) first, tokens {0..vocab-3}, PAD (= vocab-2) tail; Y uniform."""
    gen = torch.Generator(device=device).manual_seed(seed)
    X = torch.randint(0, vocab_size - 2, (n_seq, n), generator=gen, device=device)
    length = torch.randint(max(2, n // 4), n, (n_seq, 1), generator=gen, device=device)
    ar = torch.arange(n, device=device).unsqueeze(0)
    X = torch.where(ar >= length, torch.full_like(X, vocab_size - 2), X)
    X[:, 0] = vocab_size - 1
    Y = torch.randint(0, n_class, (n_seq,), generator=gen, device=device)
    return X, Y
