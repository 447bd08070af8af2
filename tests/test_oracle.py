"""The oracle against the reference-generated golden fixtures (CPU only)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

import oracle
from conftest import GOLDEN, load_golden

RTOL, ATOL = 1e-5, 1e-6   # BASELINE.json north_star: fp32 tolerance (reassociation only)


def test_chord_indices_bit_exact_small():
    z = np.load(os.path.join(GOLDEN, "chord_indices.npz"))
    assert len(z.files) >= 6
    for key in z.files:
        _, n, L = key.split("_")
        rows, cols = oracle.get_chord_indices_assym(int(n), int(L))
        got = np.asarray([rows, cols], dtype="<i8")
        assert np.array_equal(got, z[key]), key


def test_chord_indices_sha256_all_sizes():
    dig = json.load(open(os.path.join(GOLDEN, "chord_indices_sha256.json")))
    assert "16384_15" in dig and "4097_13" in dig and "65536_17" in dig
    for key, want in dig.items():
        n, L = map(int, key.split("_"))
        rows, cols = oracle.get_chord_indices_assym(n, L)
        got = hashlib.sha256(np.asarray([rows, cols], dtype="<i8").tobytes()).hexdigest()
        assert got == want, key


def test_offsets_match_indices_and_duplicate_tap():
    assert oracle.chord_offsets(1024, 12) == [0, 1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 0]
    rows, cols = oracle.get_chord_indices_assym(1024, 12)
    assert cols[0] == 0 and cols[11] == 0           # taps 0 and 11 both hit the diagonal


@pytest.mark.parametrize("N", [8, 16, 64, 128])
def test_chord_mask_equals_densified_indices(N):
    # chord_mask_mat.m with b=2, self=true: L_mask = log2 N forward links + diagonal == COO pattern with L = log2 N + 1
    L = int(np.log2(N)) + 1
    rows, cols = oracle.get_chord_indices_assym(N, L)
    dense = np.zeros((N, N))
    dense[rows, cols] = 1
    assert np.array_equal(dense, oracle.chord_mask(N, 2, False, True))
    sym = oracle.chord_mask(N, 2, True, True)
    assert np.array_equal(sym, np.maximum(dense, dense.T))
    assert oracle.chord_mask(N, 2, False, False).trace() == 0


@pytest.mark.parametrize("n,L,C", [(16, 5, 3), (128, 8, 8), (100, 8, 4), (64, 8, 2)])
def test_spmm_restatement_vs_dense_and_shift_fp64(n, L, C):
    g = torch.Generator().manual_seed(n + L)
    B = 3
    W = torch.randn(B, n, L, generator=g, dtype=torch.float64, requires_grad=True)
    V = torch.randn(B, n, C, generator=g, dtype=torch.float64, requires_grad=True)
    offs = oracle.chord_offsets(n, L)
    idx = torch.tensor(oracle.get_chord_indices_assym(n, L))
    out = oracle.spmm_ref(idx, W.reshape(B, n * L), n, n, V)
    dense = oracle.chord_dense_matrix(W.detach(), offs) @ V.detach()
    assert torch.allclose(out, dense, rtol=1e-12, atol=1e-12)
    assert torch.allclose(out, oracle.factor_apply_shift(W.detach(), V.detach(), offs), rtol=1e-12, atol=1e-12)
    dZ = torch.randn(B, n, C, generator=g, dtype=torch.float64)
    out.backward(dZ)
    dF, dV = oracle.factor_backward_shift(dZ, W.detach(), V.detach(), offs)
    assert torch.allclose(W.grad, dF, rtol=1e-12, atol=1e-12)
    assert torch.allclose(V.grad, dV, rtol=1e-12, atol=1e-12)


def _build(case, z):
    common = dict(use_cuda=False)
    if case == "adding_n128":
        return oracle.OraclePSFNet("synthetic", vocab_size=1, add_init_linear_layer=True, embedding_size=32, n_vec=128,
                                   n_W=7, Ws=[32, "GELU"], V=[32, "GELU"], n_channels_V=8, n_class=1,
                                   pooling_type="FLATTEN", head=["linear"], use_residuals=True,
                                   use_pos_embedding=False, problem="adding", **common), torch.nn.MSELoss(), True
    if case == "order_n64":
        return oracle.OraclePSFNet("synthetic", vocab_size=6, add_init_linear_layer=False, embedding_size=32, n_vec=64,
                                   n_W=6, Ws=[32, "GELU"], V=[32, "GELU"], n_channels_V=8, n_class=4,
                                   pooling_type="FLATTEN", head=["linear"], use_residuals=True,
                                   use_pos_embedding=True, problem="order", **common), torch.nn.CrossEntropyLoss(), True
    if case == "listops_n100":
        return oracle.OraclePSFNet("lra", vocab_size=17, embedding_size=24, n_vec=100, n_W=7, Ws=[16, "GELU"],
                                   V=[16, "GELU"], n_channels_V=16, n_class=10, pooling_type="CLS", head=["linear"],
                                   use_residuals=False, use_pos_embedding=True, problem="listops",
                                   **common), torch.nn.CrossEntropyLoss(), False
    if case == "pathfinder_n64_duptap":
        return oracle.OraclePSFNet("lra", vocab_size=225, embedding_size=16, n_vec=64, n_W=7, Ws=[16, "GELU"],
                                   V=[16, "GELU"], n_channels_V=8, n_class=2, pooling_type="FLATTEN",
                                   head=["non-linear", 12], use_residuals=False, use_pos_embedding=True,
                                   init_embedding_weights=True, problem="pathfinder",
                                   **common), torch.nn.CrossEntropyLoss(), False
    raise KeyError(case)


CASES = ["adding_n128", "order_n64", "listops_n100", "pathfinder_n64_duptap"]


@pytest.mark.parametrize("case", CASES)
def test_oracle_psfnet_matches_reference_golden(case):
    z = load_golden(f"psfnet_{case}.npz")
    net, loss_fn, squeeze = _build(case, z)
    sd = {k[len("param/"):]: torch.from_numpy(v) for k, v in z.items() if k.startswith("param/")}
    net.load_state_dict(sd, strict=True)          # names/shapes identical to the reference module
    X, Y = torch.from_numpy(z["X"]), torch.from_numpy(z["Y"])
    pred = net(X)
    loss = loss_fn(pred.squeeze() if squeeze else pred, Y)
    loss.backward()
    np.testing.assert_allclose(pred.detach().numpy(), z["pred"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(loss.detach().numpy(), z["loss"], rtol=RTOL, atol=ATOL)
    grads = {k[len("grad/"):]: v for k, v in z.items() if k.startswith("grad/")}
    assert grads
    for k, p in net.named_parameters():
        if k in grads:
            np.testing.assert_allclose(p.grad.numpy(), grads[k], rtol=1e-4, atol=ATOL, err_msg=k)
        else:
            assert p.grad is None or float(p.grad.abs().sum()) == 0.0, k


@pytest.mark.parametrize("case", CASES)
def test_oracle_chain_matches_reference_golden(case):
    """Chain-level tensors captured from the reference run: W_m, V0 -> final_in and their grads."""
    z = load_golden(f"psfnet_{case}.npz")
    W = torch.from_numpy(z["chain/W"]).requires_grad_(True)       # [n_W, B, n, L]
    V0 = torch.from_numpy(z["chain/V0"]).requires_grad_(True)
    n_W, B, n, L = W.shape
    residual = case in ("adding_n128", "order_n64")
    idx = torch.tensor(oracle.get_chord_indices_assym(n, L))
    out = oracle.chain_ref(idx, [W[m] for m in range(n_W)], V0, n, residual)
    fin = torch.from_numpy(z["chain/final_in"])
    dfin = torch.from_numpy(z["chain/dfinal_in"])
    if case == "listops_n100":                                      # CLS pooling: final sees row 0 only
        got = out[:, 0, :]
        dout = torch.zeros_like(out)
        dout[:, 0, :] = dfin
    else:
        got = out.reshape(B, -1)
        dout = dfin.reshape(out.shape)
    np.testing.assert_allclose(got.detach().numpy(), fin.numpy(), rtol=RTOL, atol=ATOL)
    out.backward(dout)
    np.testing.assert_allclose(W.grad.numpy(), z["chain/dW"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(V0.grad.numpy(), z["chain/dV0"], rtol=RTOL, atol=ATOL)
    # shift formulation (spmul_cuda.cu:24) agrees with the COO formulation
    offs = oracle.chord_offsets(n, L)
    out2 = oracle.chain_shift_ref([W[m].detach() for m in range(n_W)], V0.detach(), offs, residual)
    np.testing.assert_allclose(out2.numpy(), out.detach().numpy(), rtol=RTOL, atol=ATOL)
