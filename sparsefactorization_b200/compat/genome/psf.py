"""Genome_Clf/psf.py by name."""
from sparsefactorization_b200.genome.psf import PSFNet, MakeMLP, MLPBlock, get_chord_indices_assym  # noqa: F401
