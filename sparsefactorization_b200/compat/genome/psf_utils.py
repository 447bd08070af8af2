"""Genome_Clf/psf_utils.py by name."""
from sparsefactorization_b200.genome.psf_utils import *  # noqa: F401,F403
