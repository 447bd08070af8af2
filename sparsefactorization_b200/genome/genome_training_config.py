"""PSF entry of the reference's genome configuration (Genome_Clf/genome_training_config.py:2-24 and its ``training``
block): keyword arguments of ``genome.psf.PSFNet`` (``name`` is the reference's model selector and is not one of them)."""

config = {
    "DDcDNA": {
        "PSF": {
            "name": "psf",
            "vocab_size": 6,
            "embedding_size": 32,
            "n_vec": 16384,
            "n_W": 14,
            "Ws": [32, 'GELU'],
            "V": [32, 'GELU'],
            "n_channels_V": 32,
            "n_class": 2,
            "pooling_type": "FLATTEN",
            "head": ['linear'],
            "use_cuda": True,
            "use_residuals": False,
            "dropout1_p": 0.2,
            "dropout2_p": 0,
            "dropout3_p": 0.8,
            "init_embedding_weights": False,
            "use_pos_embedding": False,
        },
        "training": {
            "device_id": 0,
            "batch_size": 16,
            "learning_rate": 0.0001,
            "eval_frequency": 1,
            "num_train_steps": 100,
        },
    },
}
