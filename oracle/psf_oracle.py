"""CPU restatement of the reference's PSF chord-factor chain (torch, CPU only).

TEST INFRASTRUCTURE.  Never imported by the product package.

What is restated, and from where (all paths relative to /root/reference):

* chord COO indices ........ SyntheticExperiments/psf.py:7-32 (same function is
  repeated in LRA/psf.py:7-32, Genome_Clf/psf.py:7-32, attention_block.py:14-39)
* dense chord mask ......... non-parametric/chord_mask_mat.m:1-38
* sparse-dense product ..... ``torch_sparse.spmm`` of the pinned dependency
  torch-sparse==0.6.11 on torch-scatter==2.0.8 (requirements.txt:145-146).  That
  package is NOT vendored in /root/reference and is not installable here, so its
  published algorithm is restated:  ``scatter_add(matrix.index_select(-2, col) *
  value.unsqueeze(-1), row, dim=-2, dim_size=m)``; call sites psf.py:178-184,
  LRA/psf.py:230-236.
* one factor, explicit loops  spmul/spmul_cuda.cu:24 (fwd), :79-80 (dJdV),
  :105-108 (dJdF) -- the reference's own independent statement of the same math.
* chain loop + residual ..... SyntheticExperiments/psf.py:167-188
* model plumbing ............ SyntheticExperiments/psf.py:63-191, LRA/psf.py:64-249,
  Genome_Clf/psf.py:64-239

Parity pinning.  The reference has no tests, golden vectors or fixtures for this path
(SURVEY.md section 4, 8c).  What pins this oracle:
  1. chord indices: bit-exact against the reference's own function, executed from
     /root/reference when the fixtures were generated (tests/golden/make_golden.py);
  2. model plumbing + chain loop: golden vectors produced by importing the reference's
     own ``psf.py`` modules (Synthetic and LRA) in this container with ``torch_sparse``
     replaced by a shim implementing the published ``spmm`` algorithm above
     (tests/golden/make_golden.py, fixtures under tests/golden/);
  3. ``spmm`` itself: cross-checked against a dense fp64 chord-matrix product and against
     the explicit-loop formulas of spmul_cuda.cu.
Item 3 is a restatement check, not a run of the third-party package, so for the
floating-point sparse product itself this oracle is "parity unpinned" in the strict sense
(the real torch_sparse binary could not be executed here).
"""
from __future__ import annotations

import math
from typing import List, Sequence, Union

import numpy as np
import torch
from torch import nn

__all__ = [
    "chord_offsets",
    "get_chord_indices_assym",
    "chord_mask",
    "spmm_ref",
    "factor_apply_shift",
    "factor_backward_shift",
    "chain_ref",
    "chain_shift_ref",
    "chain_dense_fp64",
    "chord_dense_matrix",
    "OraclePSFNet",
    "oracle_make_mlp",
]


# --------------------------------------------------------------------------------------
# integer part: chord pattern
# --------------------------------------------------------------------------------------
def chord_offsets(n_vec: int, n_link: int) -> List[int]:
    """Column offset of every link of a row: 0 (self) then 2^0, 2^1, ... taken mod n_vec.

    psf.py:27 builds ``[i] + [(i + 2**k) % n_vec for k in range(n_link-1)]``; the offset
    list is that expression at i = 0.
    """
    return [0] + [(1 << k) % n_vec for k in range(n_link - 1)]


def get_chord_indices_assym(n_vec: int, n_link: int):
    """COO (rows, cols) of the one-sided chord pattern incl. the diagonal (psf.py:7-32).

    Entry e = i*n_link + k holds rows[e] = i and cols[e] = (i + offset_k) % n_vec.
    Returned as two Python lists of ints, like the reference.
    """
    i = np.arange(n_vec, dtype=np.int64)[:, None]
    off = np.asarray(chord_offsets(n_vec, n_link), dtype=np.int64)[None, :]
    rows = np.broadcast_to(i, (n_vec, n_link)).reshape(-1)
    cols = ((i + off) % n_vec).reshape(-1)
    return rows.tolist(), cols.tolist()


def chord_mask(N: int, b: int = 2, symmetric: bool = False, self_loop: bool = True) -> np.ndarray:
    """Dense 0/1 chord mask, chord_mask_mat.m:17-38 (0-based here).

    L = floor(log2 N / log2 b) forward links b^0..b^(L-1); ``symmetric`` adds the
    backward links, ``self_loop`` the diagonal.  ``random_order`` (MATLAB ``rng(0)``
    permutation, chord_mask_mat.m:12-15) is not reproducible outside MATLAB: out of scope.
    """
    M = np.zeros((N, N), dtype=np.float64)
    L = int(math.floor(math.log2(N) / math.log2(b) + 1e-12))
    steps = np.array([b ** k for k in range(L)], dtype=np.int64)
    for i in range(N):
        M[i, (i + steps) % N] = 1
        if symmetric:
            M[i, (i - steps) % N] = 1
        if self_loop:
            M[i, i] = 1
    return M


# --------------------------------------------------------------------------------------
# floating-point part: one factor and the chain
# --------------------------------------------------------------------------------------
def spmm_ref(index: torch.Tensor, value: torch.Tensor, m: int, n: int, matrix: torch.Tensor) -> torch.Tensor:
    """torch_sparse.spmm (0.6.11) restated; leading dims of value/matrix broadcast."""
    assert n == matrix.size(-2)
    row, col = index[0], index[1]
    if matrix.dim() == 1:
        matrix = matrix.unsqueeze(-1)
    gathered = matrix.index_select(-2, col) * value.unsqueeze(-1)
    out_shape = list(gathered.shape)
    out_shape[-2] = m
    out = torch.zeros(out_shape, dtype=gathered.dtype, device=gathered.device)
    idx = row.view([1] * (gathered.dim() - 2) + [-1, 1]).expand_as(gathered)
    return out.scatter_add_(-2, idx, gathered)


def factor_apply_shift(F: torch.Tensor, V: torch.Tensor, offsets: Sequence[int]) -> torch.Tensor:
    """Z[b,p,d] = sum_k F[b,p,k] * V[b,(p+offsets[k]) % n,d]   (spmul_cuda.cu:24)."""
    n = V.size(1)
    Z = torch.zeros_like(V)
    for k, off in enumerate(offsets):
        Z = Z + F[:, :, k : k + 1] * torch.roll(V, shifts=-(int(off) % n), dims=1)
    return Z


def factor_backward_shift(dZ: torch.Tensor, F: torch.Tensor, V: torch.Tensor, offsets: Sequence[int]):
    """(dJdF, dJdV) of one factor, spmul_cuda.cu:79-80 and :105-108."""
    n = V.size(1)
    dF = torch.zeros_like(F)
    dV = torch.zeros_like(V)
    for k, off in enumerate(offsets):
        off = int(off) % n
        dF[:, :, k] = (dZ * torch.roll(V, shifts=-off, dims=1)).sum(-1)
        dV = dV + torch.roll(F[:, :, k : k + 1] * dZ, shifts=off, dims=1)
    return dF, dV


def chain_ref(index: torch.Tensor, Ws: Sequence[torch.Tensor], V0: torch.Tensor, n_vec: int, use_residuals: bool) -> torch.Tensor:
    """The reference's inner loop, psf.py:167-188, over ``spmm_ref``.

    Ws[m] is [B, n, L]; V0 is [B, n, C].  ``res_conn`` is V0, captured once (psf.py:168-169)
    and re-added after every factor (psf.py:187-188).
    """
    V = V0
    res = V0
    for W in Ws:
        V = spmm_ref(index, W.reshape(W.size(0), W.size(1) * W.size(2)), n_vec, n_vec, V)
        if use_residuals:
            V = V + res
    return V


def chain_shift_ref(Ws: Sequence[torch.Tensor], V0: torch.Tensor, offsets: Sequence[int], use_residuals: bool) -> torch.Tensor:
    """Same chain through the explicit-shift formula (independent of the COO path)."""
    V = V0
    for W in Ws:
        V = factor_apply_shift(W, V, offsets)
        if use_residuals:
            V = V + V0
    return V


def chord_dense_matrix(W: torch.Tensor, offsets: Sequence[int]) -> torch.Tensor:
    """Dense [B, n, n] matrix of one factor (duplicate taps sum), for small-n checks."""
    B, n, L = W.shape
    A = torch.zeros(B, n, n, dtype=W.dtype)
    rows = torch.arange(n)
    for k, off in enumerate(offsets):
        A[:, rows, (rows + int(off)) % n] += W[:, :, k]
    return A


def chain_dense_fp64(Ws: Sequence[torch.Tensor], V0: torch.Tensor, offsets: Sequence[int], use_residuals: bool) -> torch.Tensor:
    V0 = V0.double()
    V = V0
    for W in Ws:
        V = chord_dense_matrix(W.double(), offsets) @ V
        if use_residuals:
            V = V + V0
    return V


# --------------------------------------------------------------------------------------
# model plumbing (literal loop of the reference forward, on top of spmm_ref)
# --------------------------------------------------------------------------------------
def oracle_make_mlp(cfg: List[Union[str, int]], in_dim: int, out_dim: int) -> nn.Sequential:
    """psf.py:35-47: ints are Linear widths, any str is a GELU, then a closing Linear."""
    mods: List[nn.Module] = []
    width = in_dim
    for item in cfg:
        if isinstance(item, int):
            mods.append(nn.Linear(width, item))
            width = item
        else:
            mods.append(nn.GELU())
    mods.append(nn.Linear(width, out_dim))
    return nn.Sequential(*mods)


class _OracleMLP(nn.Module):
    def __init__(self, cfg, in_dim, out_dim):
        super().__init__()
        self.network = oracle_make_mlp(cfg, in_dim, out_dim)  # attribute name fixes state-dict keys

    def forward(self, x):
        return self.network(x)


class OraclePSFNet(nn.Module):
    """PSFNet whose forward is the reference loop over ``spmm_ref``; CPU only.

    variant = "synthetic" (SyntheticExperiments/psf.py:63-191), "lra" (LRA/psf.py:64-249)
    or "genome" (Genome_Clf/psf.py:64-239: LRA minus ``problem``, plain embedding).
    Submodule names follow the reference so state dicts interchange.
    """

    def __init__(self, variant: str, *, vocab_size, embedding_size, n_vec, n_W, Ws, V, n_channels_V,
                 n_class, pooling_type, head, use_residuals, use_pos_embedding, problem=None,
                 add_init_linear_layer=False, dropout1_p=0.0, dropout2_p=0.0, dropout3_p=0.0,
                 init_embedding_weights=False, use_cuda=False):
        super().__init__()
        assert variant in ("synthetic", "lra", "genome")
        self.variant = variant
        self.n_vec, self.n_W, self.n_links = n_vec, n_W, n_W + 1
        self.use_residuals, self.use_pos_embedding = use_residuals, use_pos_embedding
        self.problem, self.pooling_type, self.head = problem, pooling_type, head
        self.add_init_linear_layer = add_init_linear_layer

        pad = vocab_size - 2 if (variant == "lra" and problem in ("imdb", "listops")) else None
        self.embedding = nn.Embedding(vocab_size, embedding_size, padding_idx=pad)
        self.pos_embedding = nn.Embedding(n_vec, embedding_size)
        if variant != "synthetic" and init_embedding_weights:
            self.embedding.weight.data.uniform_(-0.1, 0.1)
        self.fs = nn.ModuleList([_OracleMLP(Ws, embedding_size, self.n_links) for _ in range(n_W)])
        self.g = _OracleMLP(V, embedding_size, n_channels_V)

        flat = n_vec * n_channels_V
        if variant == "synthetic":
            if head[0] == "linear":
                self.final = nn.Linear(flat, n_class, bias=True)
            if add_init_linear_layer:
                self.init_linear = nn.Linear(2, embedding_size, bias=True)
        else:
            width = flat if pooling_type == "FLATTEN" else n_channels_V
            if head[0] == "linear":
                self.final = nn.Linear(width, n_class)
            elif head[0] == "non-linear":
                self.final = nn.Sequential(nn.Linear(width, head[1]), nn.GELU(), nn.Linear(head[1], n_class))
            self.dropout1 = nn.Dropout(dropout1_p)
            self.dropout2 = nn.Dropout(dropout2_p)
            self.dropout3 = nn.Dropout(dropout3_p)
        rows, cols = get_chord_indices_assym(n_vec, self.n_links)
        self.chord_indicies = torch.tensor([rows, cols])

    def forward(self, data):
        if self.variant == "synthetic":
            if self.problem == "order":
                data = self.embedding(data).squeeze(-2)
            if self.add_init_linear_layer:
                data = self.init_linear(data)
        else:
            data = self.embedding(data)
        if self.use_pos_embedding:
            pos = torch.arange(0, self.n_vec).expand(data.size(0), self.n_vec)
            data = data + self.pos_embedding(pos)
        if self.variant != "synthetic":
            data = self.dropout1(data)
        V = self.g(data)
        if self.variant != "synthetic":
            V = self.dropout2(V)
        Ws = [f(data) for f in self.fs]
        V = chain_ref(self.chord_indicies, Ws, V, self.n_vec, self.use_residuals)
        if self.variant != "synthetic":
            V = self.dropout3(V)
            if self.pooling_type == "CLS":
                V = V[:, 0, :]
        return self.final(V.reshape(V.size(0), -1))
