"""sparsefactorization_b200 -- B200-native chord-factor chain behind the reference's PSF API.

    from sparsefactorization_b200.psf import PSFNet, get_chord_indices_assym     # Synthetic ctor
    from sparsefactorization_b200.lra.psf import PSFNet                          # LRA ctor
    from sparsefactorization_b200.spmul import SparseMultiply, get_offsets
    from sparsefactorization_b200.psf_utils import TrainModel, TrainPSF, seed_everything
    from sparsefactorization_b200.chain import chord_chain                       # the fused operator
"""
__version__ = "0.1.0"
