"""GPU parity tests of the second-generation fused chain (quad-interleaved clusters, psf_chain2_sm100.cu):
against the CPU oracle at sizes it finishes in seconds, against the per-factor CUDA path and size-independent
properties at full size.  Tolerance for fp32: rtol 1e-5 / atol 1e-6 (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

import oracle
from test_gpu_chain import assert_close, dev, make_inputs, oracle_chain

pytestmark = pytest.mark.gpu


def run_gen2(W, V0, residual, dOut, expect_gen=2):
    """W [n_W,B,n,L] natural (CPU) -> quad-interleaved tap-major on the GPU -> chain -> natural-order results."""
    from sparsefactorization_b200 import chain
    n_W, B, n, L = W.shape
    Q = chain.preferred_interleave(n_W, B, n, L, V0.shape[2])
    assert chain.fused_generation(n_W, B, n, L, V0.shape[2], interleave=Q) == expect_gen
    Wd = chain.to_tap_major(W.to(dev()), interleave=Q).requires_grad_(True)
    Vd = V0.to(dev()).requires_grad_(True)
    out = chain.chord_chain(Wd, Vd, residual, interleave=Q)
    assert chain.last_path_used == "fused"
    out.backward(dOut.to(dev()))
    torch.cuda.synchronize()
    inv = torch.argsort(chain.row_permutation(n, Q))
    return out.detach().cpu(), Wd.grad.cpu()[:, :, inv, :], Vd.grad.cpu(), Q


def test_relayout_interleave_matches_row_permutation():
    from sparsefactorization_b200 import chain
    n_W, B, n, L = 2, 3, 256, 9
    W = torch.randn(n_W, B, n, L, device=dev())
    for Q in (1, 2, 4, 8):
        Wi = chain.to_tap_major(W, interleave=Q)
        assert Wi.stride(2) == 1
        perm = chain.row_permutation(n, Q, dev())
        assert torch.equal(Wi, W[:, :, perm, :]), Q
        back = chain.to_tap_major(Wi, interleave=1, src_interleave=Q)
        assert torch.equal(back, W), Q


@pytest.mark.parametrize("B,n,residual", [
    (5, 128, True),       # single-CTA shapes below 2048 rows: one warp per 128-row tile, one pass (cfg1 shape)
    (3, 256, False),
    (40, 512, True),
    (3, 1024, True),
    (3, 2048, True),      # Q = 1: no cluster, tap 3 is the first local tap
    (2, 2048, False),
    (3, 4096, True),      # Q = 2
    (2, 8192, True),      # Q = 4 (cfg3 shape)
    (2, 8192, False),
    (3, 16384, True),     # Q = 8 (cfg2 shape)
    (1, 16384, False),
    (2, 32768, True),     # 4096-row blocks, one CTA per SM, Q = 8
    (2, 65536, False),    # Q = 16: non-portable cluster size
])
def test_gen2_chain_matches_oracle(B, n, residual):
    n_W = n.bit_length() - 1
    W, V0 = make_inputs(B, n, n_W, 8, seed=n + B)
    dOut = torch.randn(B, n, 8, generator=torch.Generator().manual_seed(17)) / n
    want, wdW, wdV = oracle_chain(W, V0, residual, dOut)
    got, gdW, gdV, Q = run_gen2(W, V0, residual, dOut)
    assert Q == (max(1, n // 2048) if n <= 16384 else n // 4096)
    assert_close(got, want, "out")
    assert_close(gdV, wdV, "dV0")
    assert_close(gdW, wdW, "dW")


@pytest.mark.parametrize("B,n,C,residual", [
    (2, 2048, 128, False),    # BASELINE cfg4 chain shape (ListOps-shaped: n = 2048, 128 channels, no residual)
    (2, 4096, 16, True),      # Q = 2 clusters x 2 slabs
    (3, 128, 32, True),       # genome-style channel count on the smallest block
    (1, 16384, 16, True),
    (1, 65536, 16, False),    # 16-CTA clusters x 2 slabs
])
def test_gen2_chain_wide_channels_match_oracle(B, n, C, residual):
    """C = 8 G channels: every 8-channel slab of a sample runs as its own pseudo-sample on the same W; the slabs'
    dW contributions meet in a pre-zeroed buffer through red.add (summation order across slabs is not fixed)."""
    n_W = n.bit_length() - 1
    W, V0 = make_inputs(B, n, n_W, C, seed=n + C)
    dOut = torch.randn(B, n, C, generator=torch.Generator().manual_seed(19)) / n
    want, wdW, wdV = oracle_chain(W, V0, residual, dOut)
    got, gdW, gdV, _ = run_gen2(W, V0, residual, dOut)
    assert_close(got, want, "out")
    assert_close(gdV, wdV, "dV0")
    assert_close(gdW, wdW, "dW")


@pytest.mark.parametrize("n_W", [1, 2, 3])
@pytest.mark.parametrize("n", [512, 4096])
def test_gen2_short_chains(n, n_W):
    """Fewer factors than log2(n) (the taps still are the full chord set): first/last-factor special cases and the
    every-other-factor residual accumulation with odd and even chain lengths."""
    L = n.bit_length()
    W, V0 = make_inputs(2, n, n_W, 8, seed=n + n_W, L=L)
    dOut = torch.randn(2, n, 8, generator=torch.Generator().manual_seed(23)) / n
    want, wdW, wdV = oracle_chain(W, V0, True, dOut)
    got, gdW, gdV, _ = run_gen2(W, V0, True, dOut)
    assert_close(got, want, "out")
    assert_close(gdV, wdV, "dV0")
    assert_close(gdW, wdW, "dW")


@pytest.mark.parametrize("B,n", [(2, 256), (1, 1024), (1, 4096)])
def test_attention_map_is_the_dense_chain(B, n):
    """chain applied to the identity (C = n channels through the fused kernels) against the oracle's loop, and
    against its defining property  map @ V0 == chain(V0)."""
    from sparsefactorization_b200 import chain
    n_W = n.bit_length() - 1
    W, V0 = make_inputs(B, n, n_W, 8, seed=3 * n)
    Q = chain.preferred_interleave(n_W, B, n, n_W + 1, n)
    Wd = chain.to_tap_major(W.to(dev()), interleave=Q)
    amap = chain.attention_map(Wd, interleave=Q)
    assert amap.shape == (B, n, n) and chain.last_path_used == "fused"
    if n <= 1024:
        idx = torch.tensor(oracle.get_chord_indices_assym(n, n_W + 1))
        want = oracle.chain_ref(idx, [W[m] for m in range(n_W)], torch.eye(n).expand(B, n, n).contiguous(), n, False)
        assert_close(amap.cpu(), want, "attention map")
    with torch.no_grad():
        out = chain.chord_chain(Wd, V0.to(dev()), False, interleave=Q)
    torch.testing.assert_close(torch.bmm(amap, V0.to(dev())), out, rtol=1e-4, atol=1e-5)


def test_gen2_no_grad_forward_rotating_slots():
    from sparsefactorization_b200 import chain
    B, n, n_W = 2, 4096, 12
    W, V0 = make_inputs(B, n, n_W, 8, seed=21)
    idx = torch.tensor(oracle.get_chord_indices_assym(n, n_W + 1))
    want = oracle.chain_ref(idx, [W[m] for m in range(n_W)], V0, n, True)
    with torch.no_grad():
        got = chain.chord_chain(chain.to_tap_major(W.to(dev()), interleave=2), V0.to(dev()), True, interleave=2)
    assert_close(got.cpu(), want, "no-grad forward")


def test_gen2_rejects_wrong_interleave_and_generic_request():
    from sparsefactorization_b200 import chain
    from sparsefactorization_b200._capi import PsfError
    W = chain.empty_tap_major(12, 1, 4096, 13, device=dev()).zero_()
    V0 = torch.zeros(1, 4096, 8, device=dev())
    with pytest.raises(PsfError):
        chain.chord_chain(W, V0, True, interleave=4)          # the kernel for n = 4096 wants Q = 2
    with pytest.raises(PsfError):
        chain.chord_chain(W, V0, True, interleave=2, path="generic")


@pytest.mark.parametrize("n,B", [(2048, 700), (16384, 40)])
def test_gen2_full_size_equals_generic_and_is_linear(n, B):
    """More samples than resident clusters (persistent loop, exchange-buffer reuse); the per-factor CUDA
    path on the same values is the comparator, plus linearity in V0."""
    from sparsefactorization_b200 import chain
    n_W, C = n.bit_length() - 1, 8
    L = n_W + 1
    Q = chain.preferred_interleave(n_W, B, n, L, C)
    g = torch.Generator(device=dev()).manual_seed(3)
    Wn = 0.25 * torch.randn(n_W, B, n, L, generator=g, device=dev()) / np.sqrt(L)
    Va = torch.randn(B, n, C, generator=g, device=dev()).requires_grad_(True)
    Vb = torch.randn(B, n, C, generator=g, device=dev())
    dOut = torch.randn(B, n, C, generator=g, device=dev()) / n
    Wi = chain.to_tap_major(Wn, interleave=Q).requires_grad_(True)
    Wg = chain.to_tap_major(Wn).requires_grad_(True)
    del Wn
    out2 = chain.chord_chain(Wi, Va, True, interleave=Q)
    out2.backward(dOut)
    dV2, Va.grad = Va.grad.clone(), None
    out1 = chain.chord_chain(Wg, Va, True, path="generic")
    assert chain.last_path_used == "generic"
    out1.backward(dOut)
    inv = torch.argsort(chain.row_permutation(n, Q, dev()))
    torch.testing.assert_close(out2, out1, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dV2, Va.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(Wi.grad[:, :, inv, :], Wg.grad, rtol=1e-4, atol=1e-5)
    with torch.no_grad():
        lhs = chain.chord_chain(Wi, 0.5 * Va + 2.0 * Vb, True, interleave=Q)
        rhs = 0.5 * out2 + 2.0 * chain.chord_chain(Wi, Vb, True, interleave=Q)
    torch.testing.assert_close(lhs, rhs, rtol=1e-4, atol=1e-4)


def test_gen2_shift_factors_are_a_rotation():
    """Factors that select a single tap: the chain is a cyclic rotation by the sum of the offsets (bit-exact)."""
    from sparsefactorization_b200 import chain
    B, n, n_W, C = 5, 16384, 14, 8
    L = n_W + 1
    Q = chain.preferred_interleave(n_W, B, n, L, C)
    offs = oracle.chord_offsets(n, L)
    for start in (0, 1, 2):
        W = chain.empty_tap_major(n_W, B, n, L, device=dev())
        W.zero_()
        taps = [(3 * m + start) % L for m in range(n_W)]
        for m, k in enumerate(taps):
            W[m, :, :, k] = 1.0          # constant along the row axis: the row order does not matter
        V0 = torch.randn(B, n, C, device=dev())
        with torch.no_grad():
            out = chain.chord_chain(W, V0, False, interleave=Q)
        shift = sum(offs[k] for k in taps) % n
        assert torch.equal(out, torch.roll(V0, -shift, dims=1)), start


@pytest.mark.parametrize("problem", ["adding", "order"])
def test_psfnet_large_n_uses_gen2_and_matches_oracle_module(problem):
    """Module level at n = 4096 (Q = 2): the raw input rows are permuted, the fused generators emit W
    quad-interleaved and V0 in natural order; logits and gradients must match the oracle's restatement of the
    reference module (SyntheticExperiments/psf.py) on the same weights, at rtol 1e-5 / atol 1e-6 with the float64-anchored
    fallback (tests/parity.py).  `order` adds the token and position embeddings (psf.py:151-162)."""
    import parity
    from sparsefactorization_b200 import generators
    n, n_W, B = 4096, 12, 2
    adding = problem == "adding"
    kw = dict(vocab_size=1 if adding else 6, add_init_linear_layer=adding, embedding_size=32, n_vec=n, n_W=n_W,
              Ws=[32, "GELU"], V=[32, "GELU"], n_channels_V=8, n_class=1 if adding else 4, pooling_type="FLATTEN",
              head=["linear"], use_residuals=True, use_pos_embedding=not adding, problem=problem)
    g = torch.Generator().manual_seed(6)
    if adding:
        X = torch.stack([1 - 2 * torch.rand(B, n, generator=g), (torch.rand(B, n, generator=g) < 0.001).float()], -1)
        Y = torch.rand(B, generator=g)
    else:
        X = torch.randint(0, 6, (B, n, 1), generator=g)
        Y = torch.randint(0, 4, (B,), generator=g)
    before = generators.stats.launches
    recs, info = parity.module_parity("synthetic", kw, X, Y, "mse" if adding else "ce")
    assert info["path"] == "fused" and generators.stats.launches > before            # the fused generator kernels ran
    bad = [r for r in recs if not r["ok"]]
    assert not bad, bad