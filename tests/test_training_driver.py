"""The training driver (psf_training.py) and the config dicts it reads: host logic on the CPU, one tiny run on the GPU."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from sparsefactorization_b200 import psf_training
from sparsefactorization_b200.lra.psf_training_config import config as lra_config
from sparsefactorization_b200.synthetic_training_config import config as syn_config


def test_synthetic_settings_follow_the_reference_script():
    # SyntheticExperiments/psf_training.py:28-45: dim -> embedding_size, n_W = log2(n_vec); synthetic_training_config.py:62-68
    family, kw, tr = psf_training.model_settings("adding", 16384)
    assert family == "synthetic"
    assert (kw["n_vec"], kw["n_W"], kw["embedding_size"], kw["n_channels_V"]) == (16384, 14, 32, 8)
    assert kw["add_init_linear_layer"] and not kw["use_pos_embedding"] and kw["use_residuals"]
    assert (tr["batch_size"], tr["learning_rate"], tr["num_train_steps"]) == (40, 0.001, 20)
    _, kw, _ = psf_training.model_settings("order", 8192)
    assert (kw["vocab_size"], kw["n_class"], kw["n_W"], kw["use_pos_embedding"]) == (6, 4, 13, True)
    with pytest.raises(ValueError):
        psf_training.model_settings("adding", 2000)
    assert set(syn_config) == {"adding", "order"}


def test_lra_and_genome_settings():
    # LRA/psf_training_config.py:5-10, :64-65, :95-97
    family, kw, tr = psf_training.model_settings("listops")
    assert family == "lra" and (kw["n_vec"], kw["n_W"], kw["embedding_size"], kw["n_channels_V"]) == (2000, 11, 512, 128)
    assert kw["Ws"] == [128, 'GELU'] and kw["pooling_type"] == "CLS" and tr["batch_size"] == 32
    assert lra_config["pathfinder"]["model"]["n_W"] == 11 and lra_config["pathfinder"]["model"]["n_vec"] == 1024   # duplicate tap
    assert lra_config["imdb"]["model"]["n_vec"] == 4097 and lra_config["imdb"]["training"]["learning_rate"] == 0.0001
    family, kw, tr = psf_training.model_settings("genome")
    assert family == "genome" and "name" not in kw and "problem" not in kw and tr["batch_size"] == 16
    assert (kw["n_vec"], kw["n_W"], kw["n_channels_V"], kw["dropout3_p"]) == (16384, 14, 32, 0.8)


def test_data_file_names():
    assert psf_training.data_files("adding", 128)[0] == ("adding_128_train.pt", "adding_128_train_target.pt")
    assert psf_training.data_files("listops", None)[2] == ("test_clean.pt", "target_test_clean.pt")
    assert psf_training.data_files("genome", None)[1] == ("DDcDNA16384_val.pt", "DDcDNA16384_val_targets.pt")


def test_needs_a_data_source():
    with pytest.raises(SystemExit):
        psf_training.main(["--task", "adding", "--n_vec", "128"])


@pytest.mark.gpu
def test_driver_runs_adding_on_the_gpu(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    hist = psf_training.main(["--task", "adding", "--n_vec", "128", "--synthetic", "160", "--epochs", "2", "--batch-size", "40"])
    assert len(hist) == 2 and all(torch.isfinite(torch.tensor(h["train_loss"])) for h in hist)
    assert "test_acc" in hist[0]


@pytest.mark.gpu
def test_driver_reads_reference_files(tmp_path, monkeypatch):
    from sparsefactorization_b200 import data
    monkeypatch.chdir(tmp_path)
    for split, (fx, fy) in zip((0, 1, 2), psf_training.data_files("order", 128)):
        X, Y = data.temporal_order(80, 128, seed=split)
        torch.save(X, tmp_path / fx)
        torch.save(Y, tmp_path / fy)
    hist = psf_training.main(["--task", "order", "--n_vec", "128", "--data-dir", str(tmp_path), "--epochs", "1", "--batch-size", "40"])
    assert len(hist) == 1 and 0.0 <= hist[0]["test_acc"] <= 100.0
