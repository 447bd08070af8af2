"""Drop-in for the reference's ``SyntheticExperiments/psf.py`` (Adding / Temporal Order).

Same public names and signatures -- ``get_chord_indices_assym``, ``MakeMLP``, ``MLPBlock``,
``PSFNet`` (15-argument constructor, psf.py:63-78) -- same submodule names, so reference
state dicts load with ``strict=True``.  The Python loop over ``torch_sparse.spmm``
(psf.py:172-188) is replaced by one fused CUDA operator (``chain.chord_chain``).
"""
from __future__ import annotations

import torch
from torch import nn

from .head import apply_head
from ._psfnet import (MakeMLP, MLPBlock, PSFChainMixin, chord_mask, get_chord_indices_assym,  # noqa: F401
                      skinny_linear)

__all__ = ["get_chord_indices_assym", "chord_mask", "MakeMLP", "MLPBlock", "PSFNet"]


class PSFNet(PSFChainMixin, nn.Module):
    def __init__(self,
                 vocab_size,
                 add_init_linear_layer,
                 embedding_size,
                 n_vec,
                 n_W,
                 Ws,
                 V,
                 n_channels_V,
                 n_class,
                 pooling_type,
                 head,
                 use_cuda,
                 use_residuals,
                 use_pos_embedding,
                 problem):
        super().__init__()
        self.vocab_size = vocab_size
        self.add_init_linear_layer = add_init_linear_layer
        self.embedding_size = embedding_size
        self.n_vec = n_vec
        self.n_W = n_W
        self.n_links = n_W + 1
        self.Ws = Ws
        self.V = V
        self.n_channels_V = n_channels_V
        self.n_class = n_class
        self.pooling_type = pooling_type
        self.head = head
        self.use_cuda = use_cuda
        self.use_residuals = use_residuals
        self.use_pos_embedding = use_pos_embedding
        self.problem = problem

        # creation order follows psf.py:97-141 so that a seeded init matches the reference's
        self.embedding = nn.Embedding(vocab_size, embedding_size)
        self.pos_embedding = nn.Embedding(n_vec, embedding_size)
        self.fs = nn.ModuleList([MLPBlock(Ws, embedding_size, self.n_links) for _ in range(n_W)])
        self.g = MLPBlock(V, embedding_size, n_channels_V)
        if head[0] == 'linear':
            self.final = nn.Linear(n_vec * n_channels_V, n_class, bias=True)
        if add_init_linear_layer:
            self.init_linear = nn.Linear(2, embedding_size, bias=True)
        self._make_indices()     # attribute ``chord_indicies`` (sic), int64 [2, n_vec*n_links]; not a buffer

    def forward(self, data):
        # the gen-2 chain kernel wants W with quad-interleaved rows: permute the raw input rows (a few bytes
        # each) rather than the embedded ones; everything up to the generators is row-wise
        Q = self._early_interleave(data)
        perm = self._row_perm(Q, data.device) if Q > 1 else None
        if perm is not None:
            data = data.index_select(1, perm)
        if self.problem == 'order':
            data = self.embedding(data).squeeze(-2)
        if self.add_init_linear_layer:
            data = skinny_linear(self.init_linear, data)
        if self.use_pos_embedding:
            positions = perm if perm is not None else torch.arange(0, self.n_vec, device=data.device)
            data = data + self.pos_embedding(positions.expand(data.size(0), self.n_vec))
        V = self._chain_from_data(data, pre_interleaved=Q)   # V = g(data), all n_W factors (+ residual V_0 after each)
        return apply_head(self.final, V.reshape(V.size(0), -1))
