"""Directories that carry the reference's own module names (``psf``, ``psf_utils``, ``spmul``, ``spmul_cuda``,
``synthetic_training_config`` ...), one per reference folder, so that its scripts run unchanged with the directory first on
``sys.path``::

    PYTHONPATH=$(python -m sparsefactorization_b200.compat synthetic):$REPO python -P psf_training.py      # SyntheticExperiments/
    PYTHONPATH=$(python -m sparsefactorization_b200.compat lra):$REPO python -P listops_training.py        # LRA/
    PYTHONPATH=$(python -m sparsefactorization_b200.compat genome):$REPO python -P genome_clf_training.py  # Genome_Clf/
    PYTHONPATH=$(python -m sparsefactorization_b200.compat spmul):$REPO python -P sf_parametric_cuda.py    # spmul/

(``$REPO``: the root of this repository, so that the package itself is importable.  ``-P``, Python >= 3.11: do not put the script's own directory - which holds the reference's ``psf.py`` - first.)
Each file is a one-line re-export of the module documented in INTEGRATION.md section 1.
"""
import os

FLAVOURS = ("synthetic", "lra", "genome", "spmul")


def path(flavour: str) -> str:
    """Directory to put first on sys.path for the reference folder ``flavour``."""
    if flavour not in FLAVOURS:
        raise ValueError("flavour must be one of %s" % (FLAVOURS,))
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), flavour)
