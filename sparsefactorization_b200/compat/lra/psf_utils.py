"""LRA/psf_utils.py by name."""
from sparsefactorization_b200.lra.psf_utils import *  # noqa: F401,F403
