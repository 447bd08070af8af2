"""bench.py contract on the CPU: the reference arm prints one JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-batch", "1"], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "psf_fwd_bwd_seqs_per_sec_n16384"
    assert line["unit"] == "seq/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "seq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["n"] == 16384 and "workload" in line["config"]


def test_non_rank0_reference_prints_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_algorithmic_bytes_match_baseline_md():
    sys.path.insert(0, ROOT)
    import bench
    fb, bb = bench.Workload("cfg2").alg_bytes_per_seq()
    assert round(fb / 1e6, 3) == 14.811 and round(bb / 1e6, 3) == 29.098
    assert round((fb + bb) / 1e6, 3) == 43.909
    fb, bb = bench.Workload("cfg3").alg_bytes_per_seq()
    assert round((fb + bb) / 1e6, 3) == 19.202
    fb, bb = bench.Workload("cfg4").alg_bytes_per_seq()
    assert round((fb + bb) / 1e6, 3) == 8.487                                  # cfg4


def test_reference_arm_other_config():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "cfg3", "--steps", "1",
                          "--warmup", "0", "--ref-batch", "1"], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["config"]["n"] == 8192 and line["metric"].endswith("order_n8192") and line["value"] > 0
