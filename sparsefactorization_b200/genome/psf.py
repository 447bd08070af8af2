"""Drop-in for the reference's ``Genome_Clf/psf.py``: the LRA ``PSFNet`` without the ``problem``
argument and with a plain embedding (Genome_Clf/psf.py:64-81,103-106)."""
from __future__ import annotations

from torch import nn

from .._psfnet import MakeMLP, MLPBlock, get_chord_indices_assym  # noqa: F401
from ..lra.psf import PSFNet as _LRAPSFNet

__all__ = ["get_chord_indices_assym", "MakeMLP", "MLPBlock", "PSFNet"]


class PSFNet(_LRAPSFNet):
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
                 use_pos_embedding):
        super().__init__(vocab_size, embedding_size, n_vec, n_W, Ws, V, n_channels_V, n_class, pooling_type, head,
                         use_cuda, use_residuals, dropout1_p, dropout2_p, dropout3_p, init_embedding_weights,
                         use_pos_embedding, problem=None)

    def _make_embedding(self):
        self.embedding = nn.Embedding(self.vocab_size, self.embedding_size)
