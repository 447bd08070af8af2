"""CPU oracle for the PSF chord-factor chain.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  Nothing under
``sparsefactorization_b200/`` imports it; the product path has no CPU fallback.

Parity status: see the header of ``oracle/psf_oracle.py``.
"""
from .psf_oracle import *  # noqa: F401,F403
