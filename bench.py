#!/usr/bin/env python
"""Headline benchmark: PSF fwd+bwd sequences/s at n = 16384 (BASELINE.json metric, config cfg2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one training step of the drop-in PSFNet on one batch of synthetic Adding data at
n = 16384 (the literal PSF config of SyntheticExperiments/synthetic_training_config.py:4-18):
generators -> fused chord chain -> head -> MSE loss -> backward -> gradient all-reduce (N > 1) ->
fused Adam step.  One JSON line is printed by rank 0:

  value      whole-job seq/s with the batch resident in HBM (CUDA events, max over ranks)
  e2e        same step fed from pinned HOST buffers: H2D of every step's X, Y and D2H of every step's loss inside
             the timed region (the copy of step i+1 and the read of loss i overlap the device work of their neighbours)
  roofline   the dominant kernel of the chord-chain path BASELINE.json's metric names (psf_chain2_bwd_kernel):
             algorithmic bytes / measured launch time vs the measured HBM copy peak (MEASURED_PEAKS.json);
             `chain` holds fwd, bwd and fwd+bwd figures
  generators the fused W/V generator kernels (FP32-FMA bound; they are the larger share of the step)
  cpu_baseline  the CPU oracle (oracle/, torch on the host cores) on a bounded sample of the same workload
  clocks     SM clock / throttle reasons sampled during the timed region (NVML)

`--impl reference` times the reference's own CPU path for this workload: the oracle port
(the reference's operator, torch_sparse.spmm, is not installable here), all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_VEC, N_W, C_V, EMB = 16384, 14, 8, 32
METRIC = "psf_fwd_bwd_seqs_per_sec_n16384"
UNIT = "seq/s"


def alg_bytes_per_seq(n=N_VEC, n_W=N_W, L=N_W + 1, C=C_V):
    """SURVEY.md 8(d) / BASELINE.md 5: compulsory fp32 traffic of the chain per sequence."""
    S_W, S_V = n_W * n * L, n * C
    return 4 * (S_W + 2 * S_V), 4 * (2 * S_W + 3 * S_V)


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def psf_kwargs(use_cuda):
    return dict(vocab_size=1, add_init_linear_layer=True, embedding_size=EMB, n_vec=N_VEC, n_W=N_W,
                Ws=[32, "GELU"], V=[32, "GELU"], n_channels_V=C_V, n_class=1, pooling_type="FLATTEN",
                head=["linear"], use_cuda=use_cuda, use_residuals=True, use_pos_embedding=False, problem="adding")


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_oracle_seqs_per_sec(batch, steps, warmup):
    import torch
    import oracle
    from sparsefactorization_b200 import data
    try:        # all the host threads this process may use (torchrun pins OMP_NUM_THREADS to 1)
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except (AttributeError, RuntimeError):
        pass
    torch.manual_seed(42)
    kw = psf_kwargs(False)
    net = oracle.OraclePSFNet("synthetic", **kw)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    X, Y = data.adding(batch, N_VEC, seed=1234)
    loss_fn = torch.nn.MSELoss()

    def step():
        opt.zero_grad(set_to_none=True)
        out = loss_fn(net(X).squeeze(), Y)
        out.backward()
        opt.step()
        return float(out.detach())

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.ref_batch
    steps = min(args.steps, 5)
    warmup = min(args.warmup, 1)
    val, ms, threads = cpu_oracle_seqs_per_sec(batch, steps, warmup)
    sample = (f"oracle port of the reference CPU path (PSFNet loop over restated torch_sparse.spmm), "
              f"{steps} steps x {batch} sequences of n={N_VEC} after {warmup} warm-up, fwd+bwd+Adam")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": round(ms, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "PSF Adding n=16384 fwd+bwd+Adam (BASELINE cfg2) on host cores", "n": N_VEC,
                   "n_W": N_W, "channels": C_V, "batch_per_step": batch},
        "cpu_baseline": {"value": round(val, 4), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": round(val, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from sparsefactorization_b200 import chain, data, dp as dpmod, generators
    from sparsefactorization_b200.psf import PSFNet
    from sparsefactorization_b200.psf_utils import seed_everything

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the PSF chain has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    seed_everything(42)                      # identical init on every rank (psf_training.py:16)
    net = PSFNet(**psf_kwargs(True)).to(dev)
    dp = dpmod.DataParallel(net) if world > 1 else None
    if dp is not None:
        dp.broadcast_parameters()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, fused=True)
    loss_fn = torch.nn.MSELoss()
    # rank-sharded synthetic data, generated on the device (weak scaling: B sequences per GPU)
    X, Y = data.adding(B, N_VEC, seed=1234 + rank, device=dev)
    Xh, Yh = X.cpu().pin_memory(), Y.cpu().pin_memory()

    def step(xb, yb):
        return dpmod.train_step(net, dp, opt, loss_fn, xb, yb)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(max(args.warmup, 3)):
        step(X, Y)
    assert chain.last_path_used == "fused", f"fused chain kernel not used ({chain.last_path_used})"

    sampler = ClockSampler(local_rank)
    sampler.start()
    chain.stats.reset(timing=True)
    generators.stats.reset(timing=True)
    total_ms = timed(lambda: step(X, Y), args.steps)
    launches = chain.stats.launches + generators.stats.launches     # kernels of libpsf_b200.so in the timed region
    t_fwd, t_bwd = chain.stats.mean_ms("fwd"), chain.stats.mean_ms("bwd")
    g_fwd, g_bwd = generators.stats.mean_ms("fwd"), generators.stats.mean_ms("bwd")
    chain.stats.reset(timing=False)
    generators.stats.reset(timing=False)
    plan_gen = chain.fused_generation(N_W, B, N_VEC, N_W + 1, C_V,
                                      interleave=chain.preferred_interleave(N_W, B, N_VEC, N_W + 1, C_V))
    clocks = sampler.stop()

    # end to end: host buffers in, loss out, every step.  Every step's X and Y are copied from pinned host memory
    # inside the timed region; the copy of step i+1 is issued on a side stream before step i's loss is read back,
    # so it overlaps the device work (what a prefetching loader does).
    copy_stream = torch.cuda.Stream(device=dev)

    def issue_copy():
        with torch.cuda.stream(copy_stream):
            xb = Xh.to(dev, non_blocking=True)
            yb = Yh.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return xb, yb, ev

    e2e_state = {"next": None, "left": 0, "pending": None, "losses": []}
    loss_host = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]

    def read_pending():
        if e2e_state["pending"] is not None:
            buf, ev = e2e_state["pending"]
            ev.synchronize()
            e2e_state["losses"].append(float(buf[0]))     # the device -> host read of that step's loss
            e2e_state["pending"] = None

    def e2e_step():
        main = torch.cuda.current_stream()
        xb, yb, ev = e2e_state["next"] if e2e_state["next"] is not None else issue_copy()
        main.wait_event(ev)
        xb.record_stream(main)
        yb.record_stream(main)
        e2e_state["left"] -= 1
        e2e_state["next"] = issue_copy() if e2e_state["left"] > 0 else None
        loss = step(xb, yb)
        # the loss is copied to pinned host memory asynchronously and read one step later (the last one right away),
        # so that launching the next step does not wait for this one to finish
        buf = loss_host[e2e_state["left"] & 1]
        buf.copy_(loss.detach().reshape(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(main)
        read_pending()
        e2e_state["pending"] = (buf, ev)
        if e2e_state["left"] == 0:
            read_pending()

    e2e_state["left"] = 2
    for _ in range(2):
        e2e_step()
    e2e_state["left"] = args.steps           # exactly `steps` host->device copies inside the timed region
    e2e_state["losses"] = []
    e2e_ms = timed(e2e_step, args.steps)
    assert len(e2e_state["losses"]) == args.steps and all(v == v for v in e2e_state["losses"])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    fb, bb = alg_bytes_per_seq()
    peak, peak_src = measured_peak_gbs()
    gbs = lambda nbytes, ms: B * nbytes / ms / 1e6
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "r01_chain_traffic.json")))
        if prof.get("batch") == B and prof.get("generation") == plan_gen:
            traffic = prof.get("bwd_dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {
        "kernel": "psf_chain2_bwd_kernel" if plan_gen == 2 else "psf_chain_bwd_kernel", "bound": "hbm", "achieved": round(gbs(bb, t_bwd), 1), "peak": peak,
        "unit": "GB/s", "frac": round(gbs(bb, t_bwd) / peak, 4), "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": B * bb, "launch_ms": round(t_bwd, 4),
        "chain": {"fwd_ms": round(t_fwd, 4), "bwd_ms": round(t_bwd, 4), "fwd_GBps": round(gbs(fb, t_fwd), 1),
                  "bwd_GBps": round(gbs(bb, t_bwd), 1), "fwd_bwd_GBps": round(gbs(fb + bb, t_fwd + t_bwd), 1),
                  "fwd_bwd_frac": round(gbs(fb + bb, t_fwd + t_bwd) / peak, 4),
                  "share_of_step": round((t_fwd + t_bwd) * args.steps / total_ms, 4)},
    }
    # the W/V generators (15 small MLPs per row) are FP32-FMA work, not HBM work: reported next to the chain
    gen_flop = 2.0 * B * N_VEC * (N_W * (EMB * 32 + 32 * (N_W + 1)) + EMB * 32 + 32 * C_V)        # forward
    gen_flop_bwd = 2.0 * gen_flop + 2.0 * B * N_VEC * (N_W + 1) * EMB * 32                         # + recompute
    fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12
    gen_stats = {"kernels": "psf_gen_fwd_kernel / psf_gen_bwd_kernel", "bound": "fp32 fma (forward exact fp32; backward parameter-gradient GEMMs on mma.sync with 3xTF32 compensation)",
                 "fwd_ms": round(g_fwd, 4), "bwd_ms": round(g_bwd, 4),
                 "fwd_TFLOPs": round(gen_flop / g_fwd / 1e9, 2), "bwd_TFLOPs": round(gen_flop_bwd / g_bwd / 1e9, 2),
                 "fp32_peak_TFLOPs": round(fp32_peak, 1), "share_of_step": round((g_fwd + g_bwd) * args.steps / total_ms, 4)}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        val, ms, threads = cpu_oracle_seqs_per_sec(args.ref_batch, 2, 1)
        cpu = {"value": round(val, 4), "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"oracle PSFNet (reference loop over restated spmm) fwd+bwd+Adam, 2 steps x "
                         f"{args.ref_batch} sequences of n={N_VEC} after 1 warm-up, {ms:.0f} ms/step"}
    line = {
        "metric": METRIC, "value": round(B * world * args.steps / total_ms * 1e3, 2), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(total_ms / args.steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "PSF Adding n=16384 training step: generators + fused chord chain + head + MSE + "
                               "backward + Adam (BASELINE configs[1])",
                   "n": N_VEC, "n_W": N_W, "links": N_W + 1, "channels": C_V, "embedding": EMB,
                   "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                   "optimizer_in_step": True, "chain_kernel_generation": plan_gen,
                   "l2": "inputs larger than L2: the W stream alone is 13.8 MB/sequence x batch >> 126 MB"},
        "e2e": {"value": round(B * world * args.steps / e2e_ms * 1e3, 2), "unit": UNIT,
                "h2d_bytes_per_step": int(Xh.numel() * 4 + Yh.numel() * 4) * world, "d2h_bytes_per_step": 4 * world},
        "gpu_launches": launches,
        "roofline": roofline,
        "generators": gen_stats,
        "cpu_baseline": cpu,
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=132, help="sequences per GPU per step (weak scaling)")
    ap.add_argument("--ref-batch", type=int, default=4, help="sequences per step of the CPU reference sample")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
