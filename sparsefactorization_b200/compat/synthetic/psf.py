"""SyntheticExperiments/psf.py by name."""
from sparsefactorization_b200.psf import *  # noqa: F401,F403
from sparsefactorization_b200.psf import PSFNet, MakeMLP, MLPBlock, get_chord_indices_assym  # noqa: F401
