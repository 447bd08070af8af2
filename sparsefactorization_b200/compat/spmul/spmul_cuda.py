"""The pybind module of spmul/spmul_cuda.cu:163-166 by name: forward_host / backward_host with its argument lists."""
from sparsefactorization_b200.spmul import backward_host, forward_host  # noqa: F401
