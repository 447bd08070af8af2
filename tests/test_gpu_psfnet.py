"""Module-level parity on the GPU: the drop-in PSFNet variants against the golden vectors the
reference's own modules produced (logits, loss, every parameter gradient)."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _build(case):
    from sparsefactorization_b200.psf import PSFNet as SynPSFNet
    from sparsefactorization_b200.lra.psf import PSFNet as LraPSFNet
    if case == "adding_n128":      # synthetic_training_config.py:4-18 at n=128 (BASELINE cfg1)
        return SynPSFNet(vocab_size=1, add_init_linear_layer=True, embedding_size=32, n_vec=128, n_W=7,
                         Ws=[32, "GELU"], V=[32, "GELU"], n_channels_V=8, n_class=1, pooling_type="FLATTEN",
                         head=["linear"], use_cuda=True, use_residuals=True, use_pos_embedding=False,
                         problem="adding"), torch.nn.MSELoss(), True, "fused"
    if case == "order_n64":
        return SynPSFNet(vocab_size=6, add_init_linear_layer=False, embedding_size=32, n_vec=64, n_W=6,
                         Ws=[32, "GELU"], V=[32, "GELU"], n_channels_V=8, n_class=4, pooling_type="FLATTEN",
                         head=["linear"], use_cuda=True, use_residuals=True, use_pos_embedding=True,
                         problem="order"), torch.nn.CrossEntropyLoss(), True, "fused"
    if case == "listops_n100":
        return LraPSFNet(vocab_size=17, embedding_size=24, n_vec=100, n_W=7, Ws=[16, "GELU"], V=[16, "GELU"],
                         n_channels_V=16, n_class=10, pooling_type="CLS", head=["linear"], use_cuda=True,
                         use_residuals=False, dropout1_p=0, dropout2_p=0, dropout3_p=0,
                         init_embedding_weights=False, use_pos_embedding=True,
                         problem="listops"), torch.nn.CrossEntropyLoss(), False, "generic"
    if case == "pathfinder_n64_duptap":
        return LraPSFNet(vocab_size=225, embedding_size=16, n_vec=64, n_W=7, Ws=[16, "GELU"], V=[16, "GELU"],
                         n_channels_V=8, n_class=2, pooling_type="FLATTEN", head=["non-linear", 12], use_cuda=True,
                         use_residuals=False, dropout1_p=0, dropout2_p=0, dropout3_p=0,
                         init_embedding_weights=True, use_pos_embedding=True,
                         problem="pathfinder"), torch.nn.CrossEntropyLoss(), False, "fused"
    raise KeyError(case)


@pytest.mark.parametrize("case", ["adding_n128", "order_n64", "listops_n100", "pathfinder_n64_duptap"])
def test_psfnet_matches_reference_golden(case):
    from sparsefactorization_b200 import chain
    z = load_golden(f"psfnet_{case}.npz")
    net, loss_fn, squeeze, want_path = _build(case)
    sd = {k[len("param/"):]: torch.from_numpy(v) for k, v in z.items() if k.startswith("param/")}
    net.load_state_dict(sd, strict=True)
    net = net.cuda()
    assert net.chord_indicies.is_cuda and net.chord_indicies.dtype == torch.int64
    X, Y = torch.from_numpy(z["X"]).cuda(), torch.from_numpy(z["Y"]).cuda()
    pred = net(X)
    assert chain.last_path_used == want_path
    loss = loss_fn(pred.squeeze() if squeeze else pred, Y)
    loss.backward()
    # stated tolerance (rtol 1e-5 / atol 1e-6) against the reference's own fp32 results
    import parity
    rec = parity.check("logits", pred.detach(), torch.from_numpy(z["pred"]))
    assert rec["ok"], rec
    np.testing.assert_allclose(loss.detach().cpu().numpy(), z["loss"], rtol=1e-5, atol=1e-6)
    grads = {k[len("grad/"):]: v for k, v in z.items() if k.startswith("grad/")}
    for k, p in net.named_parameters():
        if k in grads:
            rec = parity.check(k, p.grad, torch.from_numpy(grads[k]))
            assert rec["ok"], rec
        else:
            assert p.grad is None or float(p.grad.abs().sum()) == 0.0, k


def test_generate_factors_matches_per_module_call():
    from sparsefactorization_b200._psfnet import MLPBlock, generate_factors
    torch.manual_seed(0)
    fs = torch.nn.ModuleList([MLPBlock([32, "GELU"], 32, 15) for _ in range(5)]).cuda()
    data = torch.randn(3, 256, 32, device="cuda")
    W = generate_factors(fs, data)
    assert W.shape == (5, 3, 256, 15) and W.stride(2) == 1          # tap-major: row index innermost
    for m, f in enumerate(fs):
        torch.testing.assert_close(W[m], f(data), rtol=1e-5, atol=1e-6)
    # deeper / GELU-free generator configs still batch
    fs2 = torch.nn.ModuleList([MLPBlock([16, "GELU", 24, "GELU"], 32, 9) for _ in range(3)]).cuda()
    W2 = generate_factors(fs2, data)
    for m, f in enumerate(fs2):
        torch.testing.assert_close(W2[m], f(data), rtol=1e-5, atol=1e-6)
    fs3 = torch.nn.ModuleList([MLPBlock([], 32, 9) for _ in range(2)]).cuda()
    W3 = generate_factors(fs3, data)
    for m, f in enumerate(fs3):
        torch.testing.assert_close(W3[m], f(data), rtol=1e-5, atol=1e-6)


def test_train_model_loop_learns_adding_small():
    """TrainModel end to end on a tiny Adding problem: loss must drop (smoke-level learning check)."""
    from torch.utils.data import DataLoader
    from sparsefactorization_b200 import data as D
    from sparsefactorization_b200.psf import PSFNet
    from sparsefactorization_b200.psf_utils import DatasetCreator, TrainModel, count_params, seed_everything
    seed_everything(42)
    n = 128
    net = PSFNet(1, True, 32, n, 7, [32, "GELU"], [32, "GELU"], 8, 1, "FLATTEN", ["linear"], True, True, False,
                 "adding").cuda()
    assert count_params(net) == 15809
    X, Y = D.adding(640, n, seed=1)
    Xv, Yv = D.adding(80, n, seed=2)
    mk = lambda a, b, sh: DataLoader(DatasetCreator(a, b), batch_size=40, shuffle=sh, drop_last=True)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    hist = TrainModel(net, mk(X, Y, True), mk(Xv, Yv, False), mk(Xv, Yv, False), 3, 1, opt, torch.nn.MSELoss(),
                      "adding", saving_criteria=1000.0)
    assert len(hist) == 3 and hist[-1]["train_loss"] < hist[0]["train_loss"]
    assert np.isfinite(hist[-1]["test_loss"])


@pytest.mark.parametrize("B,n,n_W,L", [(3, 256, 5, 15), (1, 100, 3, 9), (2, 1000, 2, 16), (2, 4096, 12, 13)])
def test_fused_generators_match_modules_forward_and_backward(B, n, n_W, L):
    """psf_generators_fwd/bwd (one kernel each way, hidden activations recomputed) against the per-module
    torch evaluation fs[m](data) of the reference (psf.py:175) in fp64, and against the fp32 torch route."""
    from sparsefactorization_b200 import _psfnet, generators
    from sparsefactorization_b200._psfnet import MLPBlock, generate_factors
    torch.manual_seed(n + L)
    fs = torch.nn.ModuleList([MLPBlock([32, "GELU"], 32, L) for _ in range(n_W)]).cuda()
    data = torch.randn(B, n, 32, device="cuda", requires_grad=True)
    gW = torch.randn(n_W, B, n, L, device="cuda") / n
    assert generators.fusable(fs, data)

    def run(fused):
        _psfnet.USE_FUSED_GENERATORS = fused
        try:
            for p in fs.parameters():
                p.grad = None
            data.grad = None
            before = generators.stats.launches
            W = generate_factors(fs, data)
            assert (generators.stats.launches > before) == fused
            assert W.shape == (n_W, B, n, L) and W.stride(2) == 1
            W.backward(gW)
            return W.detach().clone(), data.grad.clone(), [p.grad.clone() for p in fs.parameters()]
        finally:
            _psfnet.USE_FUSED_GENERATORS = True

    W1, dx1, g1 = run(True)
    W0, dx0, g0 = run(False)
    # fp64 ground truth from the modules themselves
    fs64 = torch.nn.ModuleList([MLPBlock([32, "GELU"], 32, L) for _ in range(n_W)]).cuda().double()
    fs64.load_state_dict({k: v.double() for k, v in fs.state_dict().items()})
    d64 = data.detach().double().requires_grad_(True)
    W64 = torch.stack([f(d64) for f in fs64], 0)
    W64.backward(gW.double())
    g64 = [p.grad for p in fs64.parameters()]

    def check(got, ref32, ref64, what):
        err = (got.double() - ref64).abs().max().item()
        err32 = (ref32.double() - ref64).abs().max().item()
        scale = ref64.abs().max().item()
        assert err <= 4 * err32 + 2e-6 * scale + 1e-9, f"{what}: ours {err:.3e} vs torch-fp32 {err32:.3e} (scale {scale:.3e})"

    check(W1, W0, W64.detach(), "W")
    torch.testing.assert_close(W1, W64.float(), rtol=1e-5, atol=1e-6)
    check(dx1, dx0, d64.grad, "d data")
    for a, b, c, (name, _) in zip(g1, g0, g64, fs.named_parameters()):
        check(a, b, c, name)


def test_fused_generators_full_size_equal_gemm_route():
    """BASELINE cfg2 shape (n = 16384, 14 generators, 15 taps), batch 8: fused kernels vs the batched-GEMM route."""
    from sparsefactorization_b200 import _psfnet
    from sparsefactorization_b200._psfnet import MLPBlock, generate_factors
    torch.manual_seed(1)
    B, n, n_W, L = 8, 16384, 14, 15
    fs = torch.nn.ModuleList([MLPBlock([32, "GELU"], 32, L) for _ in range(n_W)]).cuda()
    data = torch.randn(B, n, 32, device="cuda", requires_grad=True)
    gW = torch.randn(n_W, L, B, n, device="cuda").permute(0, 2, 3, 1) / n
    res = {}
    for fused in (True, False):
        _psfnet.USE_FUSED_GENERATORS = fused
        try:
            for p in fs.parameters():
                p.grad = None
            data.grad = None
            W = generate_factors(fs, data, interleave=8)
            W.backward(gW)
            res[fused] = (W.detach().clone(), data.grad.clone(), [p.grad.clone() for p in fs.parameters()])
        finally:
            _psfnet.USE_FUSED_GENERATORS = True
    torch.testing.assert_close(res[True][0], res[False][0], rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(res[True][1], res[False][1], rtol=1e-4, atol=1e-7)
    for a, b in zip(res[True][2], res[False][2]):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5 * float(b.abs().max()))


@pytest.mark.parametrize("save_pre", [True, False])
@pytest.mark.parametrize("B,n,n_W,L,C,Q", [(2, 256, 3, 9, 8, 1), (2, 256, 4, 9, 8, 4), (1, 1024, 2, 11, 16, 8),
                                           (3, 128, 2, 8, 3, 2)])
def test_fused_generators_with_v_generator(B, n, n_W, L, C, Q, save_pre, monkeypatch):
    """generate_all: every fs[m](data) and g(data) from one kernel; W comes out in the requested (quad-interleaved)
    row order, V0 in natural order; all gradients against the per-module fp64 evaluation."""
    from sparsefactorization_b200 import chain, generators
    from sparsefactorization_b200._psfnet import MLPBlock, generate_all
    monkeypatch.setattr(generators, "SAVE_PREACTIVATIONS", save_pre)   # backward from saved / recomputed pre-activations
    torch.manual_seed(B * n + L + Q)
    fs = torch.nn.ModuleList([MLPBlock([32, "GELU"], 32, L) for _ in range(n_W)]).cuda()
    g = MLPBlock([32, "GELU"], 32, C).cuda()
    data = torch.randn(B, n, 32, device="cuda", requires_grad=True)
    gW = torch.randn(n_W, B, n, L, device="cuda") / n       # gradient w.r.t. W in NATURAL row order
    gV = torch.randn(B, n, C, device="cuda") / n
    perm = chain.row_permutation(n, Q, "cuda")
    assert generators.fusable(fs, data, g)
    before = generators.stats.launches
    W, V0 = generate_all(fs, g, data, Q)
    assert generators.stats.launches == before + 1
    assert W.shape == (n_W, B, n, L) and W.stride(2) == 1 and V0.shape == (B, n, C)
    (W * gW[:, :, perm, :]).sum().add((V0 * gV).sum()).backward()

    mods64 = torch.nn.ModuleList([MLPBlock([32, "GELU"], 32, L) for _ in range(n_W)] + [MLPBlock([32, "GELU"], 32, C)])
    mods64 = mods64.cuda().double()
    sd = {f"{i}.{k}": v.double() for i, f in enumerate(list(fs) + [g]) for k, v in f.state_dict().items()}
    mods64.load_state_dict(sd)
    d64 = data.detach().double().requires_grad_(True)
    W64 = torch.stack([f(d64) for f in mods64[:n_W]], 0)
    V64 = mods64[n_W](d64)
    ((W64 * gW.double()).sum() + (V64 * gV.double()).sum()).backward()

    torch.testing.assert_close(W.detach(), W64[:, :, perm, :].float(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(V0.detach(), V64.float(), rtol=1e-5, atol=1e-6)

    def close(a, b, what):
        scale = float(b.abs().max())
        torch.testing.assert_close(a.double(), b, rtol=1e-4, atol=2e-6 * scale + 1e-9, msg=lambda m: f"{what}: {m}")

    close(data.grad, d64.grad, "d data")
    for (name, p32), p64 in zip(list(fs.named_parameters()) + [("g." + k, v) for k, v in g.named_parameters()],
                                mods64.parameters()):
        close(p32.grad, p64.grad, name)


def test_lra_forward_with_attention_map():
    """(logits, W_final) of the inference helper: logits equal forward(); the map applied to V0 = g(embedded data)
    reproduces the chain output the head sees (no residual in the LRA configs)."""
    from sparsefactorization_b200.lra.psf import PSFNet
    torch.manual_seed(3)
    n, n_W = 256, 8
    net = PSFNet(vocab_size=17, embedding_size=32, n_vec=n, n_W=n_W, Ws=[32, "GELU"], V=[32, "GELU"], n_channels_V=8,
                 n_class=10, pooling_type="FLATTEN", head=["linear"], use_cuda=True, use_residuals=False,
                 dropout1_p=0.1, dropout2_p=0.1, dropout3_p=0.1, init_embedding_weights=False, use_pos_embedding=True,
                 problem="listops").cuda().eval()
    X = torch.randint(0, 15, (3, n), device="cuda")
    with torch.no_grad():
        logits, amap = net.forward_with_attention_map(X)
        assert amap.shape == (3, n, n)
        torch.testing.assert_close(logits, net(X), rtol=1e-6, atol=1e-6)
        emb = net._embed(X)
        V = torch.bmm(amap, net.g(emb))
        torch.testing.assert_close(net.final(V.reshape(3, -1)), logits, rtol=1e-4, atol=1e-5)
