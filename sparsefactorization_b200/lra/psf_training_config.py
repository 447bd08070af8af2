"""The reference's LRA hyper-parameters (LRA/psf_training_config.py:2-118): ``config[task]["model"]`` are the keyword
arguments of ``lra.psf.PSFNet``, ``config[task]["training"]`` the loop settings.  Same keys and values as the reference.
"""


def _task(problem, vocab_size, embedding_size, n_vec, n_W, width, n_channels_V, n_class, pooling_type, head, use_residuals,
          dropouts, init_embedding_weights, use_pos_embedding, batch_size, learning_rate, num_train_steps):
    return {
        "model": {
            "vocab_size": vocab_size,
            "embedding_size": embedding_size,
            "n_vec": n_vec,
            "n_W": n_W,
            "Ws": [width, 'GELU'],
            "V": [width, 'GELU'],
            "n_channels_V": n_channels_V,
            "n_class": n_class,
            "pooling_type": pooling_type,          # "FLATTEN" or "CLS"
            "head": head,                          # ['linear'] or ['non-linear', hidden]
            "use_cuda": True,
            "use_residuals": use_residuals,
            "dropout1_p": dropouts[0],             # after the embeddings
            "dropout2_p": dropouts[1],             # after V
            "dropout3_p": dropouts[2],             # before the final layer
            "init_embedding_weights": init_embedding_weights,
            "use_pos_embedding": use_pos_embedding,
            "problem": problem,
        },
        "training": {
            "device_id": 0,
            "batch_size": batch_size,
            "learning_rate": learning_rate,
            "eval_frequency": 1,
            "num_train_steps": num_train_steps,
        },
    }


config = {
    # 15 tokens + PAD + CLS; 1999 positions + CLS
    "listops": _task("listops", 15 + 1 + 1, 512, 1999 + 1, 11, 128, 128, 10, "CLS", ['linear'], False, (0, 0, 0), False, True,
                     32, 0.001, 7),
    "cifar10": _task("cifar10", 256, 16, 1024, 10, 16, 16, 10, "FLATTEN", ['non-linear', 16], False, (0, 0.2, 0.8), False, True,
                     32, 0.001, 35),
    "pathfinder": _task("pathfinder", 225, 32, 1024, 11, 128, 32, 2, "FLATTEN", ['linear'], False, (0, 0, 0), False, True,
                        64, 0.001, 45),
    # 95 symbols + PAD + CLS; 4096 positions + CLS
    "imdb": _task("imdb", 95 + 1 + 1, 32, 4096 + 1, 12, 128, 32, 2, "CLS", ['linear'], True, (0.4, 0, 0), True, False,
                  32, 0.0001, 145),
}
