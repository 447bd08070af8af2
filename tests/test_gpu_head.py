"""FLATTEN head kernels (csrc/psf_head.cu) against nn.Linear: fp32 result within rtol 1e-5 / atol 1e-6 of the fp32 reference or
no further from the float64 result than it (tests/parity.py)."""
import os
import sys

import pytest
import torch
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,K,nc,bias", [(40, 131072, 1, True), (132, 16384 * 8, 4, True), (3, 8192, 10, True), (7, 4096, 16, False),
                                         (9, 8 * 1000, 2, True), (33, 65536, 1, True)])
def test_flat_head_matches_linear(B, K, nc, bias):
    import parity
    from sparsefactorization_b200 import head
    torch.manual_seed(B + nc)
    lin = nn.Linear(K, nc, bias=bias).cuda()
    V = torch.randn(B, K, device="cuda", requires_grad=True)
    assert head.covers(lin, V)
    out = head.flat_linear(lin, V)
    g = torch.randn_like(out)
    out.backward(g)
    got = [out.detach(), V.grad.clone(), lin.weight.grad.clone()] + ([lin.bias.grad.clone()] if bias else [])
    V.grad = None
    lin.zero_grad()
    ref = lin(V)
    ref.backward(g)
    want = [ref.detach(), V.grad.clone(), lin.weight.grad.clone()] + ([lin.bias.grad.clone()] if bias else [])
    lin64 = nn.Linear(K, nc, bias=bias).cuda().double()
    lin64.load_state_dict({k: v.double() for k, v in lin.state_dict().items()})
    V64 = V.detach().double().requires_grad_(True)
    r64 = lin64(V64)
    r64.backward(g.double())
    want64 = [r64.detach(), V64.grad, lin64.weight.grad] + ([lin64.bias.grad] if bias else [])
    for name, a, b, c in zip(("logits", "dV", "dW", "dbias"), got, want, want64):
        rec = parity.check(name, a, b, c)
        assert rec["ok"], rec


def test_shapes_outside_the_kernels_use_linear():
    from sparsefactorization_b200 import head
    lin = nn.Linear(8192, 32).cuda()
    V = torch.randn(4, 8192, device="cuda")
    assert not head.covers(lin, V)                       # more than 16 outputs
    assert not head.covers(nn.Linear(8, 2).cuda(), torch.randn(4, 8, device="cuda"))      # CLS pooling: tiny K
    torch.testing.assert_close(head.apply_head(lin, V), lin(V))
    seq = nn.Sequential(nn.Linear(8192, 16), nn.GELU(), nn.Linear(16, 10)).cuda()
    torch.testing.assert_close(head.apply_head(seq, V), seq(V), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("rows,IN,E", [(16384 * 5, 2, 32), (1000, 2, 32), (4099, 4, 128), (77, 1, 4), (5000, 3, 16)])
def test_init_linear_kernels_match_linear(rows, IN, E):
    """init_linear (psf.py:136-141) through csrc/psf_embed.cu against nn.Linear."""
    import parity
    from sparsefactorization_b200._psfnet import skinny_linear
    torch.manual_seed(rows)
    lin = nn.Linear(IN, E).cuda()
    x = torch.randn(rows, IN, device="cuda")
    x = x.view(5, -1, IN) if rows % 5 == 0 else x.view(1, -1, IN)
    y = skinny_linear(lin, x)
    g = torch.randn_like(y)
    y.backward(g)
    got = [y.detach(), lin.weight.grad.clone(), lin.bias.grad.clone()]
    lin.zero_grad()
    ref = lin(x)
    ref.backward(g)
    want = [ref.detach(), lin.weight.grad.clone(), lin.bias.grad.clone()]
    lin64 = nn.Linear(IN, E).cuda().double()
    lin64.load_state_dict({k: v.double() for k, v in lin.state_dict().items()})
    r64 = lin64(x.double())
    r64.backward(g.double())
    want64 = [r64.detach(), lin64.weight.grad, lin64.bias.grad]
    for name, a, b, c in zip(("y", "dW", "db"), got, want, want64):
        rec = parity.check(name, a, b, c)
        assert rec["ok"], rec
