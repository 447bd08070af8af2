"""Drop-in for the reference's ``LRA/psf.py`` (18-argument ``PSFNet``, LRA/psf.py:64-82):
dropouts, ``padding_idx`` for imdb/listops, CLS / FLATTEN pooling, linear / non-linear head."""
from __future__ import annotations

import torch
from torch import nn

from ..head import apply_head
from .._psfnet import MakeMLP, MLPBlock, PSFChainMixin, get_chord_indices_assym  # noqa: F401

__all__ = ["get_chord_indices_assym", "MakeMLP", "MLPBlock", "PSFNet"]


def _make_head(head, pooling_type, n_vec, n_channels_V, n_class):
    width = n_vec * n_channels_V if pooling_type == 'FLATTEN' else n_channels_V
    if pooling_type not in ('FLATTEN', 'CLS'):
        return None
    if head[0] == 'linear':
        return nn.Linear(width, n_class)
    if head[0] == 'non-linear':
        return nn.Sequential(nn.Linear(width, head[1]), nn.GELU(), nn.Linear(head[1], n_class))
    return None


class PSFNet(PSFChainMixin, nn.Module):
    _HAS_PROBLEM = True

    def __init__(self,
                 vocab_size,
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
                 dropout1_p,
                 dropout2_p,
                 dropout3_p,
                 init_embedding_weights,
                 use_pos_embedding,
                 problem):
        super().__init__()
        self.vocab_size = vocab_size
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
        self.dropout1_p = dropout1_p
        self.dropout2_p = dropout2_p
        self.dropout3_p = dropout3_p
        self.init_embedding_weights = init_embedding_weights
        self.use_pos_embedding = use_pos_embedding
        self.problem = problem

        self._make_embedding()
        self.pos_embedding = nn.Embedding(n_vec, embedding_size)
        if init_embedding_weights:
            self.init_embed_weights()
        self.fs = nn.ModuleList([MLPBlock(Ws, embedding_size, self.n_links) for _ in range(n_W)])
        self.g = MLPBlock(V, embedding_size, n_channels_V)
        final = _make_head(head, pooling_type, n_vec, n_channels_V, n_class)
        if final is not None:
            self.final = final
        self.dropout1 = nn.Dropout(dropout1_p)
        self.dropout2 = nn.Dropout(dropout2_p)
        self.dropout3 = nn.Dropout(dropout3_p)
        self._make_indices()

    def _make_embedding(self):
        # LRA/psf.py:106-116: PAD = vocab_size-2 for the text tasks; other problems get no embedding
        # in the reference (AttributeError later) -- here that is a constructor error.
        if self.problem in ('imdb', 'listops'):
            self.embedding = nn.Embedding(self.vocab_size, self.embedding_size, padding_idx=self.vocab_size - 2)
        elif self.problem in ('cifar10', 'pathfinder'):
            self.embedding = nn.Embedding(self.vocab_size, self.embedding_size)
        else:
            raise ValueError(f"unknown LRA problem {self.problem!r}")

    def init_embed_weights(self):
        initrange = 0.1
        self.embedding.weight.data.uniform_(-initrange, initrange)
        self.embedding.weight.requires_grad = True

    def _embed(self, data):
        data = self.embedding(data)
        if self.use_pos_embedding:
            positions = torch.arange(0, self.n_vec, device=data.device).expand(data.size(0), self.n_vec)
            data = data + self.pos_embedding(positions)
        return self.dropout1(data)

    def forward(self, data):
        data = self._embed(data)
        V = self._chain_from_data(data, self.dropout2)   # V = dropout2(g(data)), then the factor chain
        V = self.dropout3(V)
        if self.pooling_type == 'CLS':
            V = V[:, 0, :]
        return apply_head(self.final, V.reshape(V.size(0), -1))

    def forward_with_attention_map(self, data):
        """``(logits, W_final)`` like ``ChangedPSF.forward`` of the reference's attention-map scripts
        (LRA/attention_maps/imdb_inference.py:24-69): W_final is the dense ``[B, n, n]`` product of the factors.
        Inference helper: call it in eval mode (the factors of the map and of the logits then see the same input)."""
        emb = self._embed(data)
        V = self._chain_from_data(emb, self.dropout2)
        V = self.dropout3(V)
        if self.pooling_type == 'CLS':
            V = V[:, 0, :]
        return apply_head(self.final, V.reshape(V.size(0), -1)), self.attention_map(emb)
