"""Genome_Clf/genome_training_config.py by name (PSF entry)."""
from sparsefactorization_b200.genome.genome_training_config import config  # noqa: F401
