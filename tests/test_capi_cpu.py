"""CPU-side checks of the C-ABI library and the host mirror (no GPU compute)."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest
import torch

import oracle
from conftest import GOLDEN, ROOT
from sparsefactorization_b200 import _capi


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "psf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(psf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _capi.load()
    names = _declared_symbols()
    assert len(names) >= 12
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/psf_b200.h but not exported"
    assert set(names) == set(_capi.EXPORTED_SYMBOLS)
    assert lib.psf_version() >= 100


def test_error_strings_and_argument_checks():
    lib = _capi.load()
    assert lib.psf_error_string(0) == b"ok"
    assert b"argument" in lib.psf_error_string(-1)
    arr = (C.c_int64 * 4)()
    assert lib.psf_chord_offsets(0, 4, arr) == -1
    assert lib.psf_chord_offsets(16, 4, None) == -1


@pytest.mark.parametrize("n,L", [(8, 4), (100, 8), (128, 8), (1000, 11), (1024, 12), (2000, 12), (4097, 13), (16384, 15)])
def test_host_chord_indices_bit_exact_vs_oracle(n, L):
    from sparsefactorization_b200.psf import get_chord_indices_assym
    assert get_chord_indices_assym(n, L) == oracle.get_chord_indices_assym(n, L)


def test_host_chord_indices_vs_reference_golden():
    from sparsefactorization_b200._psfnet import chord_index_tensor, chord_offsets
    z = np.load(os.path.join(GOLDEN, "chord_indices.npz"))
    for key in z.files:
        _, n, L = key.split("_")
        got = chord_index_tensor(int(n), int(L)).numpy()
        assert got.dtype == np.int64 and np.array_equal(got, z[key]), key
    assert chord_offsets(1024, 12) == oracle.chord_offsets(1024, 12)
    assert chord_offsets(100, 9) == oracle.chord_offsets(100, 9)


@pytest.mark.parametrize("N,b,sym,self_", [(16, 2, False, True), (64, 2, True, True), (27, 3, False, False), (100, 2, True, False)])
def test_chord_mask_vs_oracle(N, b, sym, self_):
    from sparsefactorization_b200.psf import chord_mask
    assert np.array_equal(chord_mask(N, b, sym, self_).numpy(), oracle.chord_mask(N, b, sym, self_))


def test_workspace_query_monotone():
    lib = _capi.load()
    d = _capi.ChainDesc()
    d.B, d.n, d.C, d.n_W, d.L, d.residual, d.dtype = 4, 1024, 8, 10, 11, 1, _capi.PSF_F32
    d.ws = _capi.WStrides(11 * 4 * 1024, 1024, 1, 4 * 1024)
    a = lib.psf_chain_workspace_bytes(C.byref(d), 0)
    b = lib.psf_chain_workspace_bytes(C.byref(d), 1)
    assert a == 3 * 4 * 1024 * 8 * 4 and b == 13 * 4 * 1024 * 8 * 4


def test_state_dict_layout_matches_shipped_checkpoints():
    """Parameter names/shapes of the reference's shipped PSF checkpoints (LRA/attention_maps/*.pt)."""
    from sparsefactorization_b200.lra.psf import PSFNet
    man = json.load(open(os.path.join(GOLDEN, "checkpoint_manifest.json")))
    cfgs = {
        # LRA/psf_training_config.py:60-88 (pathfinder), :89-117 (imdb), :31-59 (cifar10)
        "pathfinder_epoch27.pt": dict(vocab_size=225, embedding_size=32, n_vec=1024, n_W=11, Ws=[128, 'GELU'],
                                      V=[128, 'GELU'], n_channels_V=32, n_class=2, pooling_type='FLATTEN',
                                      head=['linear'], problem='pathfinder'),
        "imdb_epoch138.pt": dict(vocab_size=97, embedding_size=32, n_vec=4097, n_W=12, Ws=[128, 'GELU'],
                                 V=[128, 'GELU'], n_channels_V=32, n_class=2, pooling_type='CLS', head=['linear'],
                                 problem='imdb'),
        "cifar10_epoch35.pt": dict(vocab_size=256, embedding_size=16, n_vec=1024, n_W=10, Ws=[16, 'GELU'],
                                   V=[16, 'GELU'], n_channels_V=16, n_class=10, pooling_type='FLATTEN',
                                   head=['non-linear', 16], problem='cifar10'),
    }
    for fn, cfg in cfgs.items():
        net = PSFNet(use_cuda=False, use_residuals=False, dropout1_p=0, dropout2_p=0, dropout3_p=0,
                     init_embedding_weights=False, use_pos_embedding=True, **cfg)
        mine = {k: list(v.shape) for k, v in net.state_dict().items()}
        assert mine == man[fn], fn
        assert "chord_indicies" not in mine and net.chord_indicies.shape == (2, cfg["n_vec"] * (cfg["n_W"] + 1))


def test_cpu_tensors_fail_loudly():
    from sparsefactorization_b200.chain import chord_chain
    from sparsefactorization_b200.spmul import forward_host
    with pytest.raises(RuntimeError, match="CUDA only"):
        chord_chain(torch.zeros(2, 1, 8, 4), torch.zeros(1, 8, 8))
    with pytest.raises(RuntimeError, match="CUDA only"):
        forward_host(torch.zeros(1, 8, 4), torch.zeros(1, 8, 2), [0, 1, 2, 4])


def test_synthetic_data_shapes_and_labels():
    from sparsefactorization_b200 import data
    X, Y = data.adding(16, 128, seed=1)
    assert X.shape == (16, 128, 2) and X.dtype == torch.float32 and Y.shape == (16,)
    assert torch.all(X[:, :, 1].sum(1) == 2) and X[:, :, 0].abs().max() <= 1
    assert torch.allclose(Y, 0.5 + (X[:, :, 0] * X[:, :, 1]).sum(1) / 4)
    X, Y = data.temporal_order(32, 64, seed=2)
    assert X.shape == (32, 64, 1) and X.dtype == torch.int64 and int(Y.max()) <= 3
    marks = (X.squeeze(-1) >= 4)
    assert torch.all(marks.sum(1) == 2)
    vals = X.squeeze(-1)[marks].view(32, 2)
    assert torch.equal(Y, 2 * (vals[:, 0] == 5).long() + (vals[:, 1] == 5).long())
    X, Y = data.listops_shaped(8, 100, vocab_size=17)
    assert X.shape == (8, 100) and torch.all(X[:, 0] == 16) and int(X.max()) == 16 and int(Y.max()) <= 9


def test_small_operator_wrappers_keep_linear_off_the_gpu():
    """head.flat_linear / _psfnet.skinny_linear only take the CUDA kernels for CUDA fp32 tensors."""
    from torch import nn
    from sparsefactorization_b200 import head
    from sparsefactorization_b200._psfnet import skinny_linear
    lin = nn.Linear(8192, 2)
    V = torch.randn(3, 8192)
    assert not head.covers(lin, V)
    torch.testing.assert_close(head.apply_head(lin, V), lin(V))
    seq = nn.Sequential(nn.Linear(8192, 16), nn.GELU(), nn.Linear(16, 4))
    torch.testing.assert_close(head.apply_head(seq, V), seq(V))
    init = nn.Linear(2, 32)
    x = torch.randn(2, 64, 2)
    torch.testing.assert_close(skinny_linear(init, x), init(x))
