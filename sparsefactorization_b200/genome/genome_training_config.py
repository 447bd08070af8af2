"""PSF entry of the reference's genome configuration (Genome_Clf/genome_training_config.py:2-24 and its ``training`` block),
built with the same helper as the LRA tasks: ``config["DDcDNA"]["PSF"]`` are the keyword arguments of ``genome.psf.PSFNet`` plus
``name`` (the reference's model selector), ``config["DDcDNA"]["training"]`` the loop settings.  The xformer baselines of that
file are outside this path and are not listed."""
from ..lra.psf_training_config import _task

_t = _task(None, 6, 32, 16384, 14, 32, 32, 2, "FLATTEN", ['linear'], False, (0.2, 0, 0.8), False, False, 16, 0.0001, 100)
_model = {"name": "psf", **{k: v for k, v in _t["model"].items() if k != "problem"}}

config = {"DDcDNA": {"PSF": _model, "training": _t["training"]}}
