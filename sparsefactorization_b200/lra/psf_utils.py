"""Drop-in for the reference's ``LRA/psf_utils.py`` (:10-128): same names, ``DatasetCreator`` always casts the target to
int64 (:37), ``TrainPSF`` scores by arg-max."""
from __future__ import annotations

from .. import psf_utils as _base
from ..psf_utils import TrainPSF, count_params, seed_everything  # noqa: F401

__all__ = ["seed_everything", "DatasetCreator", "count_params", "TrainPSF"]


class DatasetCreator(_base.DatasetCreator):
    def __init__(self, data, labels):
        super().__init__(data, labels, long_labels=True)
