"""LRA/psf_training_config.py by name."""
from sparsefactorization_b200.lra.psf_training_config import config  # noqa: F401
