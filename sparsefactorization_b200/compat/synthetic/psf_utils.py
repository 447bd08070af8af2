"""SyntheticExperiments/psf_utils.py by name."""
from sparsefactorization_b200.psf_utils import DatasetCreator, TrainModel, TrainPSF, count_params, seed_everything  # noqa: F401
