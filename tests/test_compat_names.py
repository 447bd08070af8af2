"""The reference's own module names resolve to the drop-in modules when a compat directory is first on sys.path
(boundary nit: `import psf`, `import spmul_cuda`)."""
import importlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from sparsefactorization_b200 import compat

CASES = {
    # what the reference scripts import (psf_training.py:1-3, LRA/listops_training.py:1-3, genome_clf_training.py:1-3,
    # spmul/sf_parametric_cuda.py:7, spmul/spmul.py:2)
    "synthetic": {"psf": ["PSFNet"], "psf_utils": ["DatasetCreator", "count_params", "seed_everything", "TrainModel"],
                  "synthetic_training_config": ["config"]},
    "lra": {"psf": ["PSFNet"], "psf_utils": ["DatasetCreator", "count_params", "seed_everything", "TrainPSF"],
            "psf_training_config": ["config"]},
    "genome": {"psf": ["PSFNet"], "psf_utils": ["DatasetCreator", "count_params", "seed_everything", "TrainPSF"],
               "genome_training_config": ["config"]},
    "spmul": {"spmul": ["SparseMultiply", "get_offsets"], "spmul_cuda": ["forward_host", "backward_host"]},
}


@pytest.mark.parametrize("flavour", sorted(CASES))
def test_reference_module_names_import(flavour):
    code = ["import sys", "sys.path.insert(0, %r)" % compat.path(flavour), "sys.path.insert(1, %r)" % ROOT]
    for mod, names in CASES[flavour].items():
        code.append("import %s as m" % mod)
        code.append("assert 'sparsefactorization_b200' in (getattr(m.%s, '__module__', '') or m.__doc__), m" % names[0]
                    if names[0] != "config" else "assert isinstance(m.config, dict)")
        for n in names:
            code.append("assert hasattr(m, %r), (%r, %r)" % (n, mod, n))
    r = subprocess.run([sys.executable, "-c", "\n".join(code)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]


def test_path_cli():
    r = subprocess.run([sys.executable, "-m", "sparsefactorization_b200.compat", "lra"], capture_output=True, text=True, cwd=ROOT,
                       timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == compat.path("lra")
    with pytest.raises(ValueError):
        compat.path("nope")
    assert importlib.util.find_spec("sparsefactorization_b200.compat") is not None
