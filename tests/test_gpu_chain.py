"""GPU parity tests: the CUDA paths (through the C-ABI) against the CPU oracle and the golden
fixtures.  Tolerance for fp32: rtol 1e-5 / atol 1e-6 (BASELINE.json north_star), inputs drawn
from default-initialised generators as the reference would produce them."""
import numpy as np
import pytest
import torch

import oracle
from conftest import load_golden

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-6


def dev():
    return torch.device("cuda:0")


def make_inputs(B, n, n_W, C, seed, E=32, H=32, L=None):
    """Generator-scaled W [n_W,B,n,L] and V0 [B,n,C] on the CPU (fp32)."""
    L = L if L is not None else n_W + 1
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    data = torch.randn(B, n, E, generator=g)
    fs = [oracle.oracle_make_mlp([H, "GELU"], E, L) for _ in range(n_W)]
    gnet = oracle.oracle_make_mlp([H, "GELU"], E, C)
    with torch.no_grad():
        W = torch.stack([f(data) for f in fs], 0) if n_W else torch.zeros(0, B, n, L)
        V0 = gnet(data)
    return W.contiguous(), V0.contiguous()


def oracle_chain(W, V0, residual, dOut, offsets=None):
    n_W, B, n, L = W.shape
    Wc = W.clone().requires_grad_(True)
    Vc = V0.clone().requires_grad_(True)
    if offsets is None:
        idx = torch.tensor(oracle.get_chord_indices_assym(n, L))
        out = oracle.chain_ref(idx, [Wc[m] for m in range(n_W)], Vc, n, residual)
    else:
        out = oracle.chain_shift_ref([Wc[m] for m in range(n_W)], Vc, offsets, residual)
    out.backward(dOut)
    return out.detach(), Wc.grad, Vc.grad


def run_ours(W, V0, residual, dOut, path, tap_major, offsets=None):
    from sparsefactorization_b200 import chain
    Wd = W.to(dev())
    if tap_major:
        Wd = chain.to_tap_major(Wd)
    Wd.requires_grad_(True)
    Vd = V0.to(dev()).requires_grad_(True)
    out = chain.chord_chain(Wd, Vd, residual, offsets=offsets, path=path)
    used = chain.last_path_used
    out.backward(dOut.to(dev()))
    torch.cuda.synchronize()
    return out.detach().cpu(), Wd.grad.cpu(), Vd.grad.cpu(), used


def assert_close(got, want, what):
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=RTOL, atol=ATOL, err_msg=what)


# ---------------------------------------------------------------------------------------------
def test_chord_indices_device_bit_exact():
    from sparsefactorization_b200._psfnet import chord_index_tensor
    for n, L in [(8, 4), (100, 8), (1024, 12), (4097, 13), (16384, 15)]:
        got = chord_index_tensor(n, L, dev()).cpu()
        want = torch.tensor(oracle.get_chord_indices_assym(n, L))
        assert got.dtype == torch.int64 and torch.equal(got, want), (n, L)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("B,n,L,D,offsets", [
    (3, 128, 8, 8, None),
    (2, 100, 5, 3, [0, 1, 7, 50, 99]),
    (2, 64, 8, 16, None),            # duplicate tap: 2^6 == n
    (1, 4097, 13, 4, None),
])
def test_sparse_multiply_matches_oracle(dtype, B, n, L, D, offsets):
    from sparsefactorization_b200.spmul import SparseMultiply, get_offsets
    g = torch.Generator().manual_seed(B * n + L)
    F = (0.3 * torch.randn(B, n, L, generator=g)).to(dtype)
    V = torch.randn(B, n, D, generator=g).to(dtype)
    dZ = torch.randn(B, n, D, generator=g).to(dtype)
    offs_host = offsets if offsets is not None else [0] + [2 ** k for k in range(L - 1)]
    want = oracle.factor_apply_shift(F, V, offs_host)
    wdF, wdV = oracle.factor_backward_shift(dZ, F, V, offs_host)
    Fd, Vd = F.to(dev()).requires_grad_(True), V.to(dev()).requires_grad_(True)
    off_arg = None if offsets is None else torch.tensor(offsets, device=dev())
    if offsets is None:
        assert torch.equal(get_offsets(L).cpu(), torch.tensor([0] + [2 ** k for k in range(L - 1)]))
    Z = SparseMultiply.apply(Fd, Vd, off_arg, 16, 64, 16, 16)
    Z.backward(dZ.to(dev()))
    tol = dict(rtol=1e-5, atol=1e-6) if dtype == torch.float32 else dict(rtol=1e-12, atol=1e-12)   # the stated fp32 bar
    np.testing.assert_allclose(Z.detach().cpu().numpy(), want.numpy(), **tol)
    np.testing.assert_allclose(Fd.grad.cpu().numpy(), wdF.numpy(), **tol)
    np.testing.assert_allclose(Vd.grad.cpu().numpy(), wdV.numpy(), **tol)


CHAIN_CASES = [
    # B, n, n_W, C, residual
    (4, 128, 7, 8, True),        # cfg1 shape
    (2, 64, 7, 8, False),        # duplicate tap
    (3, 1024, 10, 8, True),
    (2, 2000, 11, 8, False),     # non power of two
    (2, 4096, 12, 8, True),      # largest single-CTA block
    (2, 8192, 13, 8, True),      # cfg3 shape: 2-CTA cluster
    (1, 516, 9, 8, True),        # ragged last warp tile
]


@pytest.mark.parametrize("B,n,n_W,C,residual", CHAIN_CASES)
@pytest.mark.parametrize("path,tap_major", [("generic", False), ("generic", True), ("fused", True), ("fused_gen1", True)])
def test_chain_matches_oracle(B, n, n_W, C, residual, path, tap_major):
    W, V0 = make_inputs(B, n, n_W, C, seed=n + n_W)
    dOut = torch.randn(B, n, C, generator=torch.Generator().manual_seed(7)) / n
    want, wdW, wdV = oracle_chain(W, V0, residual, dOut)
    got, gdW, gdV, used = run_ours(W, V0, residual, dOut, path, tap_major)
    assert used == ("fused" if path == "fused_gen1" else path)     # "fused" = the newest generation that fits
    assert_close(got, want, "out")
    assert_close(gdV, wdV, "dV0")
    assert_close(gdW, wdW, "dW")


def test_chain_cfg2_shape_fused_vs_oracle():
    """n = 16384, 14 factors, C = 8: the BASELINE headline shape (4-CTA cluster), small batch."""
    B, n, n_W, C = 2, 16384, 14, 8
    W, V0 = make_inputs(B, n, n_W, C, seed=99)
    dOut = torch.randn(B, n, C, generator=torch.Generator().manual_seed(8)) / n
    want, wdW, wdV = oracle_chain(W, V0, True, dOut)
    got, gdW, gdV, used = run_ours(W, V0, True, dOut, "auto", True)
    assert used == "fused"
    assert_close(got, want, "out")
    assert_close(gdV, wdV, "dV0")
    assert_close(gdW, wdW, "dW")


@pytest.mark.parametrize("B,n,n_W,C,residual", [(2, 4097, 12, 4, True), (2, 100, 7, 16, False), (1, 2000, 11, 128, False),
                                                (2, 256, 8, 32, True)])
def test_chain_auto_falls_back_to_generic_and_matches(B, n, n_W, C, residual):
    W, V0 = make_inputs(B, n, n_W, C, seed=n)
    dOut = torch.randn(B, n, C, generator=torch.Generator().manual_seed(9)) / n
    want, wdW, wdV = oracle_chain(W, V0, residual, dOut)
    got, gdW, gdV, used = run_ours(W, V0, residual, dOut, "auto", False)
    assert used == "generic"
    assert_close(got, want, "out")
    assert_close(gdV, wdV, "dV0")
    assert_close(gdW, wdW, "dW")


def test_chain_arbitrary_offsets_fused():
    B, n, n_W, C, L = 2, 512, 3, 8, 6
    W, V0 = make_inputs(B, n, n_W, C, seed=5, L=L)
    offs = [0, 3, 8, 100, 509, 256]
    dOut = torch.randn(B, n, C, generator=torch.Generator().manual_seed(10)) / n
    want, wdW, wdV = oracle_chain(W, V0, True, dOut, offsets=offs)
    for path in ("generic", "fused"):
        got, gdW, gdV, used = run_ours(W, V0, True, dOut, path, True, offsets=offs)
        assert used == path
        assert_close(got, want, "out " + path)
        assert_close(gdV, wdV, "dV0 " + path)
        assert_close(gdW, wdW, "dW " + path)


def test_chain_fp64_generic():
    B, n, n_W, C = 2, 128, 7, 8
    W, V0 = make_inputs(B, n, n_W, C, seed=3)
    W, V0 = W.double(), V0.double()
    dOut = torch.randn(B, n, C, dtype=torch.float64, generator=torch.Generator().manual_seed(11))
    want, wdW, wdV = oracle_chain(W, V0, True, dOut)
    got, gdW, gdV, used = run_ours(W, V0, True, dOut, "auto", False)
    assert used == "generic"
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(gdW.numpy(), wdW.numpy(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(gdV.numpy(), wdV.numpy(), rtol=1e-12, atol=1e-12)


def test_chain_no_grad_and_empty_cases():
    from sparsefactorization_b200 import chain
    W, V0 = make_inputs(3, 256, 8, 8, seed=12)
    idx = torch.tensor(oracle.get_chord_indices_assym(256, 9))
    want = oracle.chain_ref(idx, [W[m] for m in range(8)], V0, 256, True)
    with torch.no_grad():
        got = chain.chord_chain(chain.to_tap_major(W.to(dev())), V0.to(dev()), True)
    assert chain.last_path_used == "fused"
    assert_close(got.cpu(), want, "no-grad forward (3 rotating state slots)")
    # zero factors: identity
    out = chain.chord_chain(torch.zeros(0, 3, 256, 9, device=dev()), V0.to(dev()), True)
    assert torch.equal(out.cpu(), V0)
    # empty batch
    out = chain.chord_chain(torch.zeros(4, 0, 256, 5, device=dev()), torch.zeros(0, 256, 8, device=dev()), False)
    assert out.shape == (0, 256, 8)


@pytest.mark.parametrize("case,residual", [("adding_n128", True), ("order_n64", True), ("listops_n100", False),
                                           ("pathfinder_n64_duptap", False)])
def test_chain_against_reference_golden(case, residual):
    """Chain-level tensors captured from the reference's own module run (tests/golden/make_golden.py)."""
    z = load_golden(f"psfnet_{case}.npz")
    W, V0 = torch.from_numpy(z["chain/W"]), torch.from_numpy(z["chain/V0"])
    n_W, B, n, L = W.shape
    fin, dfin = torch.from_numpy(z["chain/final_in"]), torch.from_numpy(z["chain/dfinal_in"])
    if case == "listops_n100":
        dOut = torch.zeros(B, n, V0.shape[2])
        dOut[:, 0, :] = dfin
    else:
        dOut = dfin.reshape(B, n, -1)
    paths = [("generic", False)] + ([("fused", True)] if V0.shape[2] == 8 and n % 4 == 0 else [])
    for path, tm in paths:
        got, gdW, gdV, used = run_ours(W, V0, residual, dOut, path, tm)
        got_fin = got[:, 0, :] if case == "listops_n100" else got.reshape(B, -1)
        assert_close(got_fin, fin, f"final_in {path}")
        assert_close(gdW, torch.from_numpy(z["chain/dW"]), f"dW {path}")
        assert_close(gdV, torch.from_numpy(z["chain/dV0"]), f"dV0 {path}")


# ---- full-size, size-independent properties (BASELINE cfg2 shape, batch 40) -------------------
def _full_size_W(B, n, n_W, L, seed):
    from sparsefactorization_b200 import chain
    g = torch.Generator(device=dev()).manual_seed(seed)
    W = chain.empty_tap_major(n_W, B, n, L, device=dev())
    W.copy_(0.25 * torch.randn(n_W, B, n, L, generator=g, device=dev()) / np.sqrt(L))
    return W


def test_full_size_fused_equals_generic_and_is_linear():
    from sparsefactorization_b200 import chain
    B, n, n_W, C = 40, 16384, 14, 8
    L = n_W + 1
    W = _full_size_W(B, n, n_W, L, 1).requires_grad_(True)
    g = torch.Generator(device=dev()).manual_seed(2)
    Va = torch.randn(B, n, C, generator=g, device=dev()).requires_grad_(True)
    Vb = torch.randn(B, n, C, generator=g, device=dev())
    dOut = torch.randn(B, n, C, generator=g, device=dev()) / n
    outs = {}
    for path in ("fused", "generic"):
        W.grad = None
        Va.grad = None
        out = chain.chord_chain(W, Va, True, path=path)
        assert chain.last_path_used == path
        out.backward(dOut)
        outs[path] = (out.detach(), W.grad.clone(), Va.grad.clone())
    for a, b, what in zip(outs["fused"], outs["generic"], ("out", "dW", "dV0")):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5, msg=lambda m: f"{what}: {m}")
    with torch.no_grad():
        lhs = chain.chord_chain(W, 0.5 * Va + 2.0 * Vb, True)
        rhs = 0.5 * outs["fused"][0] + 2.0 * chain.chord_chain(W, Vb, True)
    torch.testing.assert_close(lhs, rhs, rtol=1e-4, atol=1e-4)


def test_full_size_shift_factors_are_a_rotation():
    """Factors that select a single tap: the chain is a cyclic rotation by the sum of the offsets."""
    from sparsefactorization_b200 import chain
    B, n, n_W, C = 8, 16384, 14, 8
    L = n_W + 1
    offs = oracle.chord_offsets(n, L)
    W = chain.empty_tap_major(n_W, B, n, L, device=dev())
    W.zero_()
    taps = [(3 * m + 1) % L for m in range(n_W)]
    for m, k in enumerate(taps):
        W[m, :, :, k] = 1.0
    V0 = torch.randn(B, n, C, device=dev())
    with torch.no_grad():
        out = chain.chord_chain(W, V0, False)
    assert chain.last_path_used == "fused"
    shift = sum(offs[k] for k in taps) % n
    assert torch.equal(out, torch.roll(V0, -shift, dims=1))
