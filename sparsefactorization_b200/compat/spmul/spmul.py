"""spmul/spmul.py by name."""
from sparsefactorization_b200.spmul import SparseMultiply, get_offsets  # noqa: F401
