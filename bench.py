#!/usr/bin/env python
"""Headline benchmark: PSF fwd+bwd sequences/s at n = 16384 (BASELINE.json metric, config cfg2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference] [--config cfg2|cfg3|cfg4]
                    [--min-seconds S]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

`--config` selects another BASELINE.json workload (cfg3: Temporal Order n = 8192, cfg4: ListOps-shaped n = 2048, C = 128);
the default, cfg2, is the configuration the metric is quoted on.

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
  generators the fused W/V generator kernels (tcgen05 + tensor memory; they are the larger share of the step)
  reference_batch  the same step at the reference's own batch size (40 per step at cfg2 / cfg3)
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

METRIC = "psf_fwd_bwd_seqs_per_sec_n16384"
UNIT = "seq/s"


class Workload:
    """One BASELINE.json config: model constructor arguments, synthetic data, loss."""

    def __init__(self, name):
        self.name = name
        if name == "cfg2":      # Adding n = 2^14 (synthetic_training_config.py:4-18,62-68) - the config the metric is quoted on
            self.n, self.n_W, self.C, self.E = 16384, 14, 8, 32
            self.variant, self.loss, self.ref_batch, self.batch = "synthetic", "mse", 40, 132
            self.metric = METRIC
            self.label = "PSF Adding n=16384 (BASELINE configs[1])"
            self.kw = dict(vocab_size=1, add_init_linear_layer=True, embedding_size=32, n_vec=self.n, n_W=self.n_W,
                           Ws=[32, "GELU"], V=[32, "GELU"], n_channels_V=8, n_class=1, pooling_type="FLATTEN",
                           head=["linear"], use_residuals=True, use_pos_embedding=False, problem="adding")
        elif name == "cfg3":    # Temporal Order n = 2^13 (synthetic_training_config.py:72-86,130-136)
            self.n, self.n_W, self.C, self.E = 8192, 13, 8, 32
            self.variant, self.loss, self.ref_batch, self.batch = "synthetic", "ce", 40, 264
            self.metric = "psf_fwd_bwd_seqs_per_sec_order_n8192"
            self.label = "PSF Temporal Order n=8192 (BASELINE configs[2])"
            self.kw = dict(vocab_size=6, add_init_linear_layer=False, embedding_size=32, n_vec=self.n, n_W=self.n_W,
                           Ws=[32, "GELU"], V=[32, "GELU"], n_channels_V=8, n_class=4, pooling_type="FLATTEN",
                           head=["linear"], use_residuals=True, use_pos_embedding=True, problem="order")
        elif name == "cfg4":    # LRA ListOps-shaped (LRA/psf_training_config.py:2-30), n padded to 2048 as BASELINE.json says
            self.n, self.n_W, self.C, self.E = 2048, 11, 128, 512
            self.variant, self.loss, self.ref_batch, self.batch = "lra", "ce", 32, 32
            self.metric = "psf_fwd_bwd_seqs_per_sec_listops_n2048"
            self.label = "PSF ListOps-shaped n=2048, vocab 17 = 15 tokens + PAD + CLS (BASELINE configs[3])"
            self.kw = dict(vocab_size=17, embedding_size=512, n_vec=self.n, n_W=self.n_W, Ws=[128, "GELU"], V=[128, "GELU"],
                           n_channels_V=128, n_class=10, pooling_type="CLS", head=["linear"], use_residuals=False,
                           dropout1_p=0, dropout2_p=0, dropout3_p=0, init_embedding_weights=False,
                           use_pos_embedding=True, problem="listops")
        else:
            raise KeyError(name)
        self.L = self.n_W + 1

    def data(self, batch, seed, device="cpu"):
        from sparsefactorization_b200 import data
        if self.name == "cfg2":
            return data.adding(batch, self.n, seed=seed, device=device)
        if self.name == "cfg3":
            return data.temporal_order(batch, self.n, seed=seed, device=device)
        return data.listops_shaped(batch, self.n, vocab_size=17, n_class=10, seed=seed, device=device)

    def loss_fn(self):
        import torch
        return torch.nn.MSELoss() if self.loss == "mse" else torch.nn.CrossEntropyLoss()

    def net(self, use_cuda):
        if self.variant == "synthetic":
            from sparsefactorization_b200.psf import PSFNet
        else:
            from sparsefactorization_b200.lra.psf import PSFNet
        return PSFNet(use_cuda=use_cuda, **self.kw)

    def oracle_net(self):
        import oracle
        return oracle.OraclePSFNet(self.variant, use_cuda=False, **self.kw)

    def alg_bytes_per_seq(self):
        """SURVEY.md 8(d) / BASELINE.md 5: compulsory fp32 traffic of the chain per sequence (fwd, bwd)."""
        S_W, S_V = self.n_W * self.n * self.L, self.n * self.C
        return 4 * (S_W + 2 * S_V), 4 * (2 * S_W + 3 * S_V)


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_oracle_seqs_per_sec(wl, batch, steps, warmup):
    import torch
    try:        # all the host threads this process may use (torchrun pins OMP_NUM_THREADS to 1)
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except (AttributeError, RuntimeError):
        pass
    torch.manual_seed(42)
    net = wl.oracle_net()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    X, Y = wl.data(batch, 1234)
    loss_fn = wl.loss_fn()

    def step():
        opt.zero_grad(set_to_none=True)
        pred = net(X)
        out = loss_fn(pred.squeeze() if wl.loss == "mse" else pred, Y)
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
    wl = Workload(args.config)
    batch = args.ref_batch
    steps = min(args.steps, 5)
    warmup = min(args.warmup, 1)
    val, ms, threads = cpu_oracle_seqs_per_sec(wl, batch, steps, warmup)
    sample = (f"oracle port of the reference CPU path (PSFNet loop over restated torch_sparse.spmm), "
              f"{steps} steps x {batch} sequences of n={wl.n} after {warmup} warm-up, fwd+bwd+Adam")
    line = {
        "impl": "reference", "metric": wl.metric, "value": round(val, 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": round(ms, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.label + " fwd+bwd+Adam on host cores", "n": wl.n,
                   "n_W": wl.n_W, "channels": wl.C, "batch_per_step": batch},
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
    from sparsefactorization_b200 import _psfnet, chain, dp as dpmod, generators, head as headmod
    from sparsefactorization_b200.psf_utils import seed_everything
    wl = Workload(args.config)

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

    B = args.batch or wl.batch
    seed_everything(42)                      # identical init on every rank (psf_training.py:16)
    net = wl.net(True).to(dev)
    dp = dpmod.DataParallel(net) if world > 1 else None
    if dp is not None:
        dp.broadcast_parameters()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, fused=True)
    loss_fn = wl.loss_fn()
    # rank-sharded synthetic data, generated on the device (weak scaling: B sequences per GPU)
    X, Y = wl.data(B, 1234 + rank, device=dev)
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
    steps = args.steps
    small0 = headmod.launches + _psfnet._SkinnyLinear.launches          # init_linear / head kernels launched before the timed region
    total_ms = timed(lambda: step(X, Y), steps)
    if args.min_seconds > 0 and total_ms < 1e3 * args.min_seconds:      # a sustained-clock run: re-time with enough steps
        steps = int(steps * 1e3 * args.min_seconds / total_ms) + 1
        chain.stats.reset(timing=True)
        generators.stats.reset(timing=True)
        small0 = headmod.launches + _psfnet._SkinnyLinear.launches
        total_ms = timed(lambda: step(X, Y), steps)
    launches = (chain.stats.launches + generators.stats.launches + headmod.launches + _psfnet._SkinnyLinear.launches
                - small0)     # kernels of libpsf_b200.so in the timed region
    t_fwd, t_bwd = chain.stats.mean_ms("fwd"), chain.stats.mean_ms("bwd")
    g_fwd, g_bwd = generators.stats.mean_ms("fwd"), generators.stats.mean_ms("bwd")
    chain.stats.reset(timing=False)
    generators.stats.reset(timing=False)
    plan_gen = chain.fused_generation(wl.n_W, B, wl.n, wl.L, wl.C,
                                      interleave=chain.preferred_interleave(wl.n_W, B, wl.n, wl.L, wl.C))
    clocks = sampler.stop()

    # the reference's own batch size (synthetic_training_config.py:64 / LRA config: 40 / 32 per step), same model: at
    # cfg2 40 samples are 33 + 7 clusters of the persistent chain kernels, i.e. a 21 %-full second wave
    ref_b = None
    if wl.ref_batch != B and not args.no_ref_batch:
        Xr, Yr = wl.data(wl.ref_batch, 4321 + rank, device=dev)
        for _ in range(3):
            step(Xr, Yr)
        chain.stats.reset(timing=True)
        generators.stats.reset(timing=True)
        rsteps = max(steps, 10)
        rms = timed(lambda: step(Xr, Yr), rsteps)
        ref_b = {"batch_per_gpu": wl.ref_batch, "ms_per_step": round(rms / rsteps, 4),
                 "value": round(wl.ref_batch * world * rsteps / rms * 1e3, 2), "unit": UNIT,
                 "chain_fwd_ms": round(chain.stats.mean_ms("fwd"), 4), "chain_bwd_ms": round(chain.stats.mean_ms("bwd"), 4)}
        chain.stats.reset(timing=False)
        generators.stats.reset(timing=False)
        del Xr, Yr

    # end to end: host buffers in, loss out, every step.  Every step's X and Y are copied from pinned host memory
    # inside the timed region; the copy of step i+1 is issued on a side stream before step i's loss is read back,
    # so it overlaps the device work (what a prefetching loader does).
    copy_stream = torch.cuda.Stream(device=dev)
    # two device-side input slots, allocated once: the loop itself allocates nothing (a fresh cross-stream allocation per
    # step costs a cudaMalloc or an allocator synchronisation every now and then, which a short run then reports as the
    # step time)
    x_slots = [torch.empty_like(Xh, device=dev) for _ in range(2)]
    y_slots = [torch.empty_like(Yh, device=dev) for _ in range(2)]
    slot_free = [None, None]                 # event: the step that last read the slot has been enqueued and finished

    def issue_copy(slot):
        with torch.cuda.stream(copy_stream):
            if slot_free[slot] is not None:
                copy_stream.wait_event(slot_free[slot])
            x_slots[slot].copy_(Xh, non_blocking=True)
            y_slots[slot].copy_(Yh, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return slot, ev

    e2e_state = {"next": None, "left": 0, "pending": None, "losses": [], "slot": 0}
    loss_host = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]

    def read_pending():
        if e2e_state["pending"] is not None:
            buf, ev = e2e_state["pending"]
            ev.synchronize()
            e2e_state["losses"].append(float(buf[0]))     # the device -> host read of that step's loss
            e2e_state["pending"] = None

    def e2e_step():
        main = torch.cuda.current_stream()
        if e2e_state["next"] is None:
            e2e_state["next"] = issue_copy(e2e_state["slot"])
        slot, ev = e2e_state["next"]
        main.wait_event(ev)
        e2e_state["left"] -= 1
        e2e_state["slot"] = slot ^ 1
        e2e_state["next"] = issue_copy(slot ^ 1) if e2e_state["left"] > 0 else None
        loss = step(x_slots[slot], y_slots[slot])
        done = torch.cuda.Event()
        done.record(main)
        slot_free[slot] = done
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

    e2e_state["left"] = 3
    for _ in range(3):
        e2e_step()
    e2e_state["left"] = args.steps           # exactly `steps` host->device copies inside the timed region
    e2e_state["losses"] = []
    e2e_ms = timed(e2e_step, args.steps)
    assert len(e2e_state["losses"]) == args.steps and all(v == v for v in e2e_state["losses"])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    fb, bb = wl.alg_bytes_per_seq()
    peak, peak_src = measured_peak_gbs()
    gbs = lambda nbytes, ms: B * nbytes / ms / 1e6
    # DRAM traffic of the dominant kernel comes from an ncu capture (profiles/), which exists for one batch size only
    traffic = "not measured for this configuration (needs an ncu capture; profiles/ holds cfg2 at batch 132)"
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "r01_chain_traffic.json")))
        if wl.name == "cfg2" and prof.get("batch") == B and prof.get("generation") == plan_gen:
            traffic = prof.get("bwd_dram_bytes_per_launch")
    except Exception:
        pass
    share = lambda ms: round(ms * steps / total_ms, 4) if ms == ms else None
    roofline = {
        "kernel": "psf_chain2_bwd_kernel" if plan_gen == 2 else ("psf_chain_bwd_kernel" if plan_gen == 1 else "psf_generic"),
        "bound": "hbm", "achieved": round(gbs(bb, t_bwd), 1), "peak": peak,
        "unit": "GB/s", "frac": round(gbs(bb, t_bwd) / peak, 4), "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": B * bb, "launch_ms": round(t_bwd, 4),
        "chain": {"fwd_ms": round(t_fwd, 4), "bwd_ms": round(t_bwd, 4), "fwd_GBps": round(gbs(fb, t_fwd), 1),
                  "bwd_GBps": round(gbs(bb, t_bwd), 1), "fwd_bwd_GBps": round(gbs(fb + bb, t_fwd + t_bwd), 1),
                  "fwd_bwd_frac": round(gbs(fb + bb, t_fwd + t_bwd) / peak, 4),
                  "frac_of_8TBps_nominal": round(gbs(fb + bb, t_fwd + t_bwd) / 8000.0, 4),
                  "share_of_step": share(t_fwd + t_bwd)},
    }
    if wl.C > 16:
        roofline["note"] = ("C = %d: arithmetic intensity ~%d FLOP/B, the chain is bound by the shared-memory / FP32-FMA pipes rather "
                            "than HBM (SURVEY.md 8d); the HBM fraction is reported for completeness" % (wl.C, 6 * wl.C // 32 * 4))
    # the W/V generators (15 small MLPs per row): GEMM-shaped work on tcgen05 (3xTF32), GELU on the SIMT pipes
    rows = B * wl.n
    gen_flop = 2.0 * rows * (wl.n_W * (wl.E * 32 + 32 * wl.L) + wl.E * 32 + 32 * wl.C)                 # forward
    gen_flop_bwd = 2.0 * gen_flop + 2.0 * rows * (wl.n_W + 1) * wl.E * 32                              # + recompute
    fused_gen = g_fwd == g_fwd
    if fused_gen:
        gen_stats = {"kernels": "psf_gen_fwd_tc_kernel / psf_gen_bwd_tc2_kernel (tcgen05 + tensor memory, 3xTF32; backward: "
                                "row-local products and parameter gradients all on the tensor pipe)",
                     "bound": "forward: MUFU / issue (erf GELU per hidden unit); backward: issue slots and barrier latency of the "
                              "epilogue warps (tensor pipe 36 % busy, about 1.5 k of 4 k cycles per tile-generator) - "
                              "profiles/r02_gen_notes.md",
                     "fwd_ms": round(g_fwd, 4), "bwd_ms": round(g_bwd, 4),
                     "fwd_TFLOPs": round(gen_flop / g_fwd / 1e9, 2), "bwd_TFLOPs": round(gen_flop_bwd / g_bwd / 1e9, 2),
                     "share_of_step": share(g_fwd + g_bwd)}
    else:
        gen_stats = {"kernels": "batched cuBLAS GEMMs (generator shape E=%d, H=128 is outside the fused kernels)" % wl.E}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        val, ms, threads = cpu_oracle_seqs_per_sec(wl, args.ref_batch, 2, 1)
        cpu = {"value": round(val, 4), "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"oracle PSFNet (reference loop over restated spmm) fwd+bwd+Adam, 2 steps x "
                         f"{args.ref_batch} sequences of n={wl.n} after 1 warm-up, {ms:.0f} ms/step"}
    w_mb = 4.0 * wl.n_W * wl.n * wl.L / 1e6
    line = {
        "metric": wl.metric, "value": round(B * world * steps / total_ms * 1e3, 2), "unit": UNIT, "n_gpus": world,
        "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": round(total_ms / steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.label + " training step: generators + fused chord chain + head + loss + backward + Adam",
                   "n": wl.n, "n_W": wl.n_W, "links": wl.L, "channels": wl.C, "embedding": wl.E,
                   "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                   "optimizer_in_step": True, "chain_kernel_generation": plan_gen,
                   "l2": "inputs larger than L2: the W stream alone is %.1f MB/sequence x batch %d = %.0f MB vs 126 MB" % (w_mb, B, w_mb * B)},
        "e2e": {"value": round(B * world * args.steps / e2e_ms * 1e3, 2), "unit": UNIT,
                "h2d_bytes_per_step": int(Xh.numel() * Xh.element_size() + Yh.numel() * Yh.element_size()) * world,
                "d2h_bytes_per_step": 4 * world},
        "gpu_launches": launches,
        "roofline": roofline,
        "generators": gen_stats,
        "reference_batch": ref_b,
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
    ap.add_argument("--config", default="cfg2", choices=["cfg2", "cfg3", "cfg4"],
                    help="BASELINE.json workload: cfg2 Adding n=16384 (the metric's config, default), cfg3 Temporal Order n=8192, "
                         "cfg4 ListOps-shaped n=2048")
    ap.add_argument("--batch", type=int, default=0, help="sequences per GPU per step (weak scaling); 0 = the config's default "
                                                          "(cfg2: 132 = four waves of the 33 resident clusters)")
    ap.add_argument("--min-seconds", type=float, default=0.0, help="re-time with enough steps to fill this many seconds (sustained clocks)")
    ap.add_argument("--no-ref-batch", action="store_true", help="skip the extra measurement at the reference's own batch size")
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
