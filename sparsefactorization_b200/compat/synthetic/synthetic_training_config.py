"""SyntheticExperiments/synthetic_training_config.py by name (PSF entries)."""
from sparsefactorization_b200.synthetic_training_config import config  # noqa: F401
