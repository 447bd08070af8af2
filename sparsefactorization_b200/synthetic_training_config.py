"""PSF entries of the reference's ``synthetic_training_config.config`` (SyntheticExperiments/synthetic_training_config.py:4-18,
62-68 for Adding, :72-86, :130-136 for Temporal Order): same keys, same values, so that a script written against the
reference's dict reads this one unchanged.  The xformer baselines of that file (Transformer, Linformer, Performer) are
outside this path and are not listed.
"""

_TRAINING = {
    "device_id": 0,          # under torchrun the launcher's LOCAL_RANK wins (psf_training.py)
    "batch_size": 40,        # per GPU
    "learning_rate": 0.001,
    "eval_frequency": 1,
    "num_train_steps": 20,   # epochs
}


def _psf(problem, vocab_size, n_class, init_linear, pos_embedding):
    return {
        "add_init_linear_layer": init_linear,
        "vocab_size": vocab_size,
        "dim": 32,
        "Ws": [32, 'GELU'],
        "V": [32, 'GELU'],
        "pooling_type": "FLATTEN",
        "head": ['linear'],
        "n_class": n_class,
        "n_channels_V": 8,
        "use_cuda": True,
        "use_residuals": True,
        "use_pos_embedding": pos_embedding,
        "problem": problem,
    }


config = {
    "adding": {"models": {"PSF": _psf("adding", 1, 1, True, False)}, "training": dict(_TRAINING)},
    "order": {"models": {"PSF": _psf("order", 6, 4, False, True)}, "training": dict(_TRAINING)},
}
