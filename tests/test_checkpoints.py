"""The PSF checkpoints the reference ships (LRA/attention_maps/*.pt): ``load_state_dict(strict=True)`` and the logits
of the reference's own module on the same input (fixtures: tests/golden/make_checkpoint_golden.py)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
META = json.load(open(os.path.join(GOLDEN, "ckpt_meta.json")))
NAMES = sorted(META)


def _load(name):
    z = np.load(os.path.join(GOLDEN, "ckpt_%s.npz" % name))
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")}
    return sd, torch.from_numpy(z["X"]), torch.from_numpy(z["logits"])


def _kw(name):
    return dict(META[name]["config"], dropout1_p=0, dropout2_p=0, dropout3_p=0)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_the_reference_on_shipped_weights(name):
    import oracle
    sd, X, want = _load(name)
    net = oracle.OraclePSFNet("lra", use_cuda=False, **_kw(name))
    net.load_state_dict(sd, strict=True)
    net.eval()
    with torch.no_grad():
        got = net(X)
    assert sum(v.numel() for v in sd.values()) == META[name]["n_params"]
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_dropin_loads_and_reproduces_shipped_checkpoints(name):
    import oracle
    import parity
    from sparsefactorization_b200.lra.psf import PSFNet
    sd, X, want = _load(name)
    net = PSFNet(use_cuda=True, **_kw(name))
    net.load_state_dict(sd, strict=True)          # same parameter names and shapes, no extra buffers
    net = net.cuda().eval()
    with torch.no_grad():
        got = net(X.cuda())
    ref64 = oracle.OraclePSFNet("lra", use_cuda=False, **_kw(name)).double()
    ref64.load_state_dict({k: v.double() for k, v in sd.items()}, strict=True)
    ref64.eval()
    with torch.no_grad():
        want64 = ref64(X)
    rec = parity.check("logits " + name, got, want, want64)
    assert rec["ok"], rec
