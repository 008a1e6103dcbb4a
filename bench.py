#!/usr/bin/env python
"""bench.py — filter + smoother + log-likelihood-gradient throughput of the temporally-parallel
state-space GP path (BASELINE.json metric), on synthetic data.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]; SURVEY.md §8d "Config 2"): Matern52(variance 1, lengthscale 1),
noise 0.1, FP64, N = 1e6 irregular time steps per GPU: dt_k = 0.004 * U(0.5, 1.5) (RandomState
31415926), y = obs_noise(sinu(t), 0.1, seed 0), 1 % of the observations set to NaN.

* ``value``: time-steps/s of ONE device-resident step = pkf (+log-likelihood) + pks + pkf_backward
  (gradient w.r.t. the LGSSM fields) through the C ABI, LGSSM (Fs, Qs, y, P0, H, R) already in HBM.
* ``e2e``: the same metric through the reference-facing model API with HOST buffers, per step:
  H2D of (t, y) from pinned memory -> StateSpaceGP.maximum_log_likelihood_objective() + gradient
  w.r.t. the unconstrained hyper-parameters (discretisation + filter + adjoint scans) ->
  predict_f at N query times (merge + discretise + filter + smoother) -> D2H of ll, gradient,
  posterior mean and variance.
* ``roofline``: dominant kernel of the device-resident step, timed with CUDA events on its stream
  inside the timed region (library option "timing"); algorithmic bytes per DESIGN.md.
* ``cpu_baseline``: the CPU oracle (restated reference: pkf + autograd gradient + pks in TFP's
  scan order, torch-CPU on all host cores) on the same workload (rank 0, N = 1 only).
* ``configs``: the other BASELINE.json configurations measured in the same run (few steps each, same timing
  rules): configs[2] RBF order 6 (d = 6), N = 1e7 TOTAL, time-sharded over --gpus (strong scaling); configs[3]
  quasi-periodic Periodic(order 5) x Matern32 (d = 24), N = 1e6, single GPU; configs[4] Matern52 + RBF6 (d = 9),
  1.25e7 steps per GPU (N = 1e8 on 8 GPUs) and the log-likelihood over a 32 x 32 grid of hyper-parameter settings
  (batch-sharded over --gpus).  Each entry carries its own roofline and a bounded oracle-port cpu_baseline.
* ``sharded_check`` (--gpus > 1): max relative error of the time-sharded step against the unsharded step of the
  same series computed on every rank (ll, smoothed moments, gradients), d = 3 and d = 6.
* ``--impl reference``: only the CPU arm (the reference's TF stack is not installable here; the
  oracle port is the stand-in), same metric/config.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PER_GPU = 1_000_000
NOISE = 0.1
print_line = lambda obj: print(json.dumps(obj))
METRIC = "filter+smoother+grad timesteps/s"
UNIT = "timesteps/s"


def sinu(t):
    return np.sin(np.pi * t) + np.sin(2 * np.pi * t) + np.cos(3 * np.pi * t)


def make_series(n, seed_offset=0):
    """fixed-rate irregular sampling, 1 % missing observations."""
    rng = np.random.RandomState(31415926 + seed_offset)
    t = np.cumsum(0.004 * rng.uniform(0.5, 1.5, size=n))
    x = sinu(t)
    rng_y = np.random.RandomState(0 + seed_offset)
    y = x + np.sqrt(NOISE) * rng_y.normal(x, math.sqrt(NOISE), (n,))  # obs_noise quirk: noise drawn with mean x
    miss = np.random.RandomState(7 + seed_offset).choice(n, size=n // 100, replace=False)
    y[miss] = np.nan
    return t, y


def make_series_sunspot(n, seed=666):
    """monthly sampling (28..31 days cycling like calendar months, in years), an 11-year
    quasi-cycle + noise of variance 10 (sunspot/common.py:24), 1 % missing."""
    months = np.array([31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31], dtype=np.float64) / 365.25
    t = np.cumsum(np.resize(months, n))
    rng = np.random.RandomState(seed)
    y = 50.0 + 40.0 * np.sin(2 * np.pi * t / 11.0) * (1 + 0.3 * np.sin(2 * np.pi * t / 80.0)) + np.sqrt(10.0) * rng.randn(n)
    y[rng.choice(n, size=n // 100, replace=False)] = np.nan
    return t, y


def make_series_weekly(n, seed=42):
    """weekly sampling with +-1 day jitter (in years), seasonal cycle + trend, noise 0.05."""
    rng = np.random.RandomState(seed)
    t = np.cumsum(1.0 / 52.0 + rng.uniform(-1.0, 1.0, size=n) / 365.25)
    y = 0.02 * t + np.sin(2 * np.pi * t) + 0.3 * np.sin(4 * np.pi * t) + np.sqrt(0.05) * rng.randn(n)
    y[rng.choice(n, size=n // 100, replace=False)] = np.nan
    return t, y


# the other BASELINE.json configurations (name -> kernel factory on a kernels module, noise, series, sizes)
def extra_configs(world):
    return {
        "rbf6_n1e7_time_sharded": dict(
            baseline_config=2, kern=lambda K: K.RBF(1.0, 1.0, order=6, balancing_iter=5), noise=10.0,
            series=make_series_sunspot, n_total=10_000_000, scaling="strong",
            what="RBF order 6 balancing_iter 5 (sunspots-shaped), N = 1e7 total, time-sharded over n_gpus"),
        "qp5_n1e6": dict(
            baseline_config=3, kern=lambda K: K.Periodic(K.SquaredExponential(5.0, 1.0), period=1.0, order=5) * K.Matern32(0.1, 50.0),
            noise=0.05, series=make_series_weekly, n_total=1_000_000, scaling="single-gpu",
            what="Periodic(SE(5,1), period 1, order 5) x Matern32(0.1, 50) (CO2-shaped, d = 24), N = 1e6, FP64, 1 GPU"),
        "qp5_n1e6_fp32": dict(
            baseline_config=3, kern=lambda K: K.Periodic(K.SquaredExponential(5.0, 1.0), period=1.0, order=5) * K.Matern32(0.1, 50.0),
            noise=0.05, series=make_series_weekly, n_total=1_000_000, scaling="single-gpu", dtype="f32",
            what="the same in the FP32 opt-in mode (FP32 storage at the C ABI, FP64 arithmetic in the DMMA kernels)"),
        "m52rbf6_n1p25e7_per_gpu": dict(
            baseline_config=4, kern=lambda K: K.Matern52(1.0, 1.0) + K.RBF(1.0, 1.0, order=6, balancing_iter=5), noise=NOISE,
            series=make_series, n_total=12_500_000 * world, scaling="weak",
            what="Matern52 + RBF order 6 (d = 9), 1.25e7 steps per GPU time-sharded (N = 1e8 on 8 GPUs)"),
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_step(O, torch, ssm, y, n):
    """One CPU step of the restated reference: pkf + gradient (autograd through the parallel scan) + pks."""
    P0, Fs, Qs, H, R = [x.clone().requires_grad_(True) for x in ssm]
    fm, fP, ll = O.pkf((P0, Fs, Qs, H, R), y[:, None], True, max_parallel=max(n, 2))
    grads = torch.autograd.grad(ll, (P0, Fs, Qs, H, R))
    with torch.no_grad():
        sm, sP = O.pks(ssm, fm.detach(), fP.detach(), max_parallel=max(n, 2))
    return float(ll), grads, sm, sP


def cpu_reference_arm(steps, warmup, sample_n):
    """Times the CPU restatement of the reference path (oracle port) with all host threads."""
    import torch
    import __graft_entry__ as entry
    try:
        torch.set_num_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1
    except Exception:
        pass
    O = entry.import_oracle()
    t, y = make_series(sample_n)
    cov = O.Matern52(1.0, 1.0)
    with torch.no_grad():
        ssm = cov.get_ssm(t[:, None], torch.tensor([[NOISE]], dtype=torch.float64))
    for _ in range(warmup):
        oracle_step(O, torch, ssm, y, sample_n)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle_step(O, torch, ssm, y, sample_n)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return sample_n / dt, dt, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_n = 250_000
    val, dt, cores = cpu_reference_arm(args.steps, max(args.warmup, 1), sample_n)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "Matern52 single series N=1e6 per GPU, FP64, log-lik + gradient + RTS smoother "
                               "(configs[1]); each step a bounded sample of 250,000 steps of it",
                   "sample_steps": sample_n},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "first 250,000 time steps of the workload per step; oracle port (torch-CPU "
                                   "restatement of pkf + autograd gradient + pks); the TF reference is not installable"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print_line(line)


def time_steps_on_device(step, W, K, barrier, dist, dev, torch):
    for _ in range(W):
        out = step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K):
        out = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / K
    if dist is not None:
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt)
    del out
    return ms


def shard_lgssm(cfg, world, rank, dev, torch, kernels, ops):
    """Synthetic series of the config, this rank's contiguous shard discretised on the device."""
    n_total = cfg["n_total"]
    n = n_total // world
    t_host, y_host = cfg["series"](n_total)
    lo, hi = rank * n, (rank + 1) * n
    with torch.no_grad():
        sde = cfg["kern"](kernels).get_sde()
    F, Pinf = sde.F.to(dev).contiguous(), sde.P0.to(dev).contiguous()
    H = sde.H.to(dev).reshape(-1).contiguous()
    R = torch.tensor([cfg["noise"]], dtype=torch.float64, device=dev)
    t_dev = torch.as_tensor(t_host[lo:hi]).to(dev)
    t_prev = 0.0 if lo == 0 else float(t_host[lo - 1])
    dts = t_dev - torch.cat([torch.tensor([t_prev], dtype=torch.float64, device=dev), t_dev[:-1]])
    y_dev = torch.as_tensor(y_host[lo:hi]).to(dev)
    Fs, Qs = ops.discretise(F, Pinf, dts)
    if cfg.get("dtype") == "f32":
        Pinf, Fs, Qs, H, R, y_dev = (x.float().contiguous() for x in (Pinf, Fs, Qs, H, R, y_dev))
    return n, Pinf, Fs, Qs, H, R, y_dev, (t_host, y_host)


def oracle_sample_baseline(cfg, sample_n):
    """Oracle-port CPU baseline of one config on its first sample_n time steps (all host threads)."""
    import torch
    import __graft_entry__ as entry
    O = entry.import_oracle()
    t, y = cfg["series"](sample_n)
    cov = cfg["kern"](O)
    with torch.no_grad():
        ssm = cov.get_ssm(t[:, None], torch.tensor([[cfg["noise"]]], dtype=torch.float64))
    t0 = time.perf_counter()
    oracle_step(O, torch, ssm, y, sample_n)
    dt = time.perf_counter() - t0
    return {"value": sample_n / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"first {sample_n} time steps of the workload once ({dt:.1f} s): oracle port (pkf + autograd gradient "
                      f"+ pks in TFP scan order, torch-CPU)"}


def run_extra_config(name, cfg, world, rank, dev, dist, barrier, torch, kernels, ops, pdist, h, hbm_peak, fp64_peak, W, K,
                     with_cpu, xchg=None):
    if cfg["scaling"] == "single-gpu" and world > 1:
        return {"skipped": "single-GPU configuration: measured at n_gpus = 1"}
    # set-up (allocations, discretisation) first; the ranks then agree that everybody is ready before any collective
    # or peer exchange of the timed steps is issued (a rank that failed here must not leave the others waiting)
    setup_err = None
    try:
        n, Pinf, Fs, Qs, H, R, y_dev, _ = shard_lgssm(cfg, world, rank, dev, torch, kernels, ops)
        torch.cuda.synchronize()
    except Exception as e:
        setup_err = repr(e)
    if dist is not None:
        ok = torch.tensor([0.0 if setup_err else 1.0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok) == 0.0:
            return {"error": setup_err or "set-up failed on another rank"}
    elif setup_err:
        return {"error": setup_err}
    d = Fs.shape[1]
    g_ll = torch.ones(1, dtype=Fs.dtype, device=dev)
    esz = Fs.element_size()
    shard = pdist.TimeShard(rank, world, dist, exchange=xchg) if world > 1 else None

    def step():
        if shard is None:
            return ops.pkfs_grad(Pinf, Fs, Qs, H, R, y_dev, g_ll)
        return shard.filter_smoother_grad(Pinf, Fs, Qs, H, R, y_dev, g_ll)

    ms = time_steps_on_device(step, W, K, barrier, dist, dev, torch)
    h.set_option("timing", 1)
    h.timing_report()
    for _ in range(2):
        out = step()
    barrier()
    kt = h.timing_report()
    h.set_option("timing", 0)
    del out
    n_total = n * world
    alg = esz * (12 * d * d + 4 * d + 2)
    gbs = alg * n_total / (ms * 1e-3) / 1e9
    flops = 30.0 * d ** 3  # useful flops of this implementation per time step: 15 d x d products (K1 3, K2 3, K3 9)
    tfl = flops * n_total / (ms * 1e-3) / 1e12
    entry = {
        "baseline_config": cfg["baseline_config"], "workload": cfg["what"], "state_dim": d, "dtype": "f64" if esz == 8 else "f32 storage / f64 arithmetic",
        "n_total": n_total, "n_per_gpu": n, "n_gpus": world, "scaling": cfg["scaling"], "steps": K, "warmup": W,
        "ms_per_step": ms, "value": n_total / (ms * 1e-3), "unit": UNIT,
        "roofline": {"bound": "hbm" if d <= 9 else "fp64", "alg_bytes_per_timestep": alg, "achieved_gbs": gbs,
                     "hbm_peak_gbs": hbm_peak * world, "frac_of_hbm_peak": gbs / (hbm_peak * world),
                     "useful_flops_per_timestep": flops, "achieved_tflops": tfl, "fp64_peak_tflops": fp64_peak * world,
                     "frac_of_fp64_peak": tfl / (fp64_peak * world),
                     "fp64_peak_source": "DMMA m8n8k4 full-chip rate measured with scripts/sm_probe.cu (profiles/r02_sm_probe.txt)"},
        "kernels": {k: {"launches": c // 2, "avg_us": 1e3 * t / max(c, 1)} for k, (c, t) in kt.items()},
    }
    del Fs, Qs, y_dev
    torch.cuda.empty_cache()
    if with_cpu and rank == 0 and world == 1:
        try:
            entry["cpu_baseline"] = oracle_sample_baseline(cfg, 20_000 if d > 9 else 50_000)
        except Exception as e:  # pragma: no cover
            entry["cpu_baseline"] = {"error": repr(e)}
    return entry


def run_grid_config(world, rank, dev, dist, barrier, torch, kernels, ops, W):
    """configs[4]b: log-likelihood over a 32 x 32 grid of (Matern52 lengthscale, RBF lengthscale) in [0.1, 10],
    Matern52 + RBF6 (d = 9), N = 1e5, batch-sharded over the ranks (no data-path collective)."""
    from pssgp_b200 import batch
    n = 100_000
    t_host, y_host = make_series(n)
    ls = np.logspace(-1, 1, 32)
    settings = [(a, b) for a in ls for b in ls]
    mk = lambda a, b: kernels.Matern52(1.0, float(a)) + kernels.RBF(1.0, float(b), order=6, balancing_iter=5)
    data = (torch.as_tensor(t_host[:, None]).to(dev), torch.as_tensor(y_host[:, None]).to(dev))
    batch.grid_log_likelihood(mk, settings[:8 * world], data, NOISE, rank=rank, world=world, dist=dist, device=dev)  # warm-up
    barrier()
    t0 = time.perf_counter()
    ll = batch.grid_log_likelihood(mk, settings, data, NOISE, rank=rank, world=world, dist=dist, device=dev)
    barrier()
    dt = time.perf_counter() - t0
    if dist is not None:
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt)
    # the same with the hyper-parameter gradient of every setting (pssgp_grid_loglik_grad + native Jacobians), on the first
    # 256 settings
    grad_info = None
    try:
        sub = settings[:256]
        batch.grid_log_likelihood(mk, sub[:8 * world], data, NOISE, rank=rank, world=world, dist=dist, device=dev, with_grad=True)
        barrier()
        t1 = time.perf_counter()
        llg, dpar, dnz = batch.grid_log_likelihood(mk, sub, data, NOISE, rank=rank, world=world, dist=dist, device=dev,
                                                   with_grad=True)
        barrier()
        dtg = time.perf_counter() - t1
        if dist is not None:
            tg = torch.tensor([dtg], dtype=torch.float64, device=dev)
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
            dtg = float(tg)
        grad_info = {"settings": len(sub), "seconds": dtg, "value": len(sub) / dtg, "unit": "settings/s",
                     "finite": bool(torch.isfinite(dpar).all() and torch.isfinite(dnz).all()),
                     "ll_matches": bool(torch.allclose(llg, ll[:len(sub)], rtol=1e-10, atol=0))}
    except Exception as e:
        grad_info = {"error": repr(e)}
    return {"with_gradient": grad_info,
            "baseline_config": 4, "workload": "grid log-likelihood, 32 x 32 settings of Matern52 + RBF6 (d = 9), N = 1e5, "
                                              "batch-sharded over n_gpus", "settings": len(settings), "n": n,
            "seconds": dt, "value": len(settings) / dt, "unit": "settings/s", "timesteps_per_s": len(settings) * n / dt,
            "n_gpus": world, "path": "native batched SDE (pssgp_sde_batch, host C++) + one pssgp_grid_loglik call per rank "
            "(4 concurrent settings)", "finite": bool(torch.isfinite(ll).all()), "ll_max": float(ll.max()),
            "argmax_setting": [float(x) for x in settings[int(ll.argmax())]]}


def run_model_e2e(dev, torch, kernels, K):
    """End to end through StateSpaceGP with HOST buffers for the d > 4 configurations (N = 1e6): one step = pinned (t, y)
    to the device, log-likelihood + gradient w.r.t. the hyper-parameters, predict_f at N pinned queries back into pinned
    memory.  Exercises get_sde (host), pssgp_discretise / _backward, the fused filter + adjoint, merge, filter + smoother."""
    from pssgp_b200.model import StateSpaceGP
    out = {}
    n = 1_000_000
    t_host, y_host = make_series(n)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    t_pin, y_pin, q_pin = pin(t_host[:, None]), pin(y_host[:, None]), pin((t_host + 0.002)[:, None])
    mean_pin = torch.empty((n, 1), dtype=torch.float64).pin_memory()
    var_pin = torch.empty((n, 1), dtype=torch.float64).pin_memory()
    for name, mk in (("rbf6_d6", lambda: kernels.RBF(1.0, 1.0, order=6, balancing_iter=5)),
                     ("m52rbf6_d9", lambda: kernels.Matern52(1.0, 1.0) + kernels.RBF(1.0, 1.0, order=6, balancing_iter=5))):
        model = StateSpaceGP((t_pin, y_pin), mk(), noise_variance=NOISE, parallel=True, max_parallel=2 * n)

        def step():
            model.data = (t_pin, y_pin)
            ll = model.maximum_log_likelihood_objective()
            grads = torch.autograd.grad(ll, model.trainable_variables)
            model.predict_f(q_pin, out=(mean_pin, var_pin))
            return float(ll.detach()), [float(g) for g in grads]

        for _ in range(3):
            r = step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(K):
            r = step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / K
        out[name] = {"n": n, "ms_per_step": dt * 1e3, "value": n / dt, "unit": UNIT, "finite": bool(np.isfinite(r[0])),
                     "h2d_bytes_per_step": int(8 * 3 * n), "d2h_bytes_per_step": int(8 * 2 * n)}
        del model
        torch.cuda.empty_cache()
    return out


def sharded_check(world, rank, dev, dist, torch, kernels, ops, pdist, xchg=None):
    """Time-sharded step against the unsharded step of the same series (computed on every rank), d = 3 and d = 6."""
    out = {}
    for name, mk, n_rank in (("matern52_d3", lambda: kernels.Matern52(1.0, 1.0), 60_000),
                             ("rbf6_d6", lambda: kernels.RBF(1.0, 1.0, order=6, balancing_iter=5), 20_000)):
        n_total = n_rank * world
        t_host, y_host = make_series(n_total, seed_offset=5)
        with torch.no_grad():
            sde = mk().get_sde()
        F, Pinf = sde.F.to(dev).contiguous(), sde.P0.to(dev).contiguous()
        H = sde.H.to(dev).reshape(-1).contiguous()
        R = torch.tensor([NOISE], dtype=torch.float64, device=dev)
        t_dev = torch.as_tensor(t_host).to(dev)
        dts = t_dev - torch.cat([torch.zeros(1, dtype=torch.float64, device=dev), t_dev[:-1]])
        y_dev = torch.as_tensor(y_host).to(dev)
        Fs, Qs = ops.discretise(F, Pinf, dts)
        g = torch.full((1,), 1.1, dtype=torch.float64, device=dev)
        (fms, fPs, ll), (sms, sPs), (dP0, dFs, dQs, dH, dR) = ops.pkfs_grad(Pinf, Fs, Qs, H, R, y_dev, g)
        lo, hi = rank * n_rank, (rank + 1) * n_rank
        sh = pdist.TimeShard(rank, world, dist, exchange=xchg)
        ll2, sms2, sPs2, (dP02, dFs2, dQs2, dH2, dR2) = sh.filter_smoother_grad(
            Pinf, Fs[lo:hi].contiguous(), Qs[lo:hi].contiguous(), H, R, y_dev[lo:hi].contiguous(), g)
        rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))
        errs = torch.tensor([rel(ll2, ll), rel(sms2, sms[lo:hi]), rel(sPs2, sPs[lo:hi]), rel(dFs2, dFs[lo:hi]),
                             rel(dQs2, dQs[lo:hi]), rel(dP02, dP0), rel(dH2, dH), rel(dR2, dR)], dtype=torch.float64, device=dev)
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
        names = ["ll", "sms", "sPs", "dFs", "dQs", "dP0", "dH", "dR"]
        out[name] = {"n_total": n_total, "max_rel_err": float(errs.max()), "per_output": dict(zip(names, [float(e) for e in errs]))}
    out["max_rel_err"] = max(v["max_rel_err"] for v in out.values())
    return out


def sequential_numba_baseline(sample_n=200_000):
    """numba-JIT sequential kf + ks (pssgp/kalman/sequential.py restated), ONE host core."""
    import torch
    import __graft_entry__ as entry
    O = entry.import_oracle()
    import seq_numba
    t, y = make_series(sample_n)
    with torch.no_grad():
        ssm = O.Matern52(1.0, 1.0).get_ssm(t[:, None], torch.tensor([[NOISE]], dtype=torch.float64))
    P0, Fs, Qs, H, R = [np.ascontiguousarray(x.numpy()) for x in ssm]
    seq_numba.kfs(P0, Fs[:1000], Qs[:1000], H, R, y[:1000])  # JIT compile
    t0 = time.perf_counter()
    seq_numba.kfs(P0, Fs, Qs, H, R, y)
    dt = time.perf_counter() - t0
    return {"value": sample_n / dt, "unit": "filter+smoother timesteps/s", "cores": 1, "kind": "port",
            "sample": f"first {sample_n} time steps of the workload, sequential kf + ks (numba-JIT restatement of "
                      f"pssgp/kalman/sequential.py:11-73), no gradient (the reference obtains it by TF autodiff)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=N_PER_GPU, help="time steps per GPU (default: the BASELINE workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE configs (configs[2..4])")
    ap.add_argument("--only", default=None, help="comma-separated names of the extra configs to run")
    args = ap.parse_args()
    # exactly ONE line on stdout (the JSON): anything a library prints on fd 1 meanwhile (NCCL's version banner) goes to
    # stderr instead
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    global print_line
    print_line = lambda obj: (real_stdout.write(json.dumps(obj) + "\n"), real_stdout.flush())
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import __graft_entry__ as entry
    entry.import_package()
    from pssgp_b200 import _lib, kernels, ops
    from pssgp_b200.model import StateSpaceGP

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n = args.n
    d = 3
    W = max(args.warmup, 3)
    K = args.steps

    # ---- synthetic workload -------------------------------------------------------------------------
    from pssgp_b200 import dist as pdist
    t_host, y_host = make_series(n * world)  # one long series, time-sharded across ranks
    lo, hi = rank * n, (rank + 1) * n
    cov = kernels.Matern52(1.0, 1.0)
    with torch.no_grad():
        sde = cov.get_sde()
    F = sde.F.to(dev).contiguous()
    Pinf = sde.P0.to(dev).contiguous()
    H = sde.H.to(dev).reshape(-1).contiguous()
    R = torch.tensor([NOISE], dtype=torch.float64, device=dev)
    t_dev = torch.as_tensor(t_host[lo:hi]).to(dev)
    t_prev = 0.0 if lo == 0 else float(t_host[lo - 1])
    dts = t_dev - torch.cat([torch.tensor([t_prev], dtype=torch.float64, device=dev), t_dev[:-1]])
    y_dev = torch.as_tensor(y_host[lo:hi]).to(dev)
    Fs, Qs = ops.discretise(F, Pinf, dts)
    g_ll = torch.ones(1, dtype=torch.float64, device=dev)
    xchg, xchg_kind = None, "none"
    if world > 1:
        xchg_kind = "nccl collectives"
        if os.environ.get("PSSGP_EXCHANGE", "peer") == "peer":
            try:
                xchg = pdist.PeerExchange(rank, world, dist, dev)
                xchg_kind = "NVLink peer stores (pssgp_peer_exchange)"
            except Exception as e:  # symmetric memory unavailable: NCCL collectives
                print(f"[bench] peer exchange unavailable ({e!r}); using NCCL collectives", file=sys.stderr)
        ok = torch.tensor([1.0 if xchg is not None else 0.0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok) == 0.0:
            xchg, xchg_kind = None, "nccl collectives"
    shard = pdist.TimeShard(rank, world, dist, exchange=xchg) if world > 1 else None
    h = _lib.handle(local_rank)

    def device_step():
        if shard is None:
            # one C-ABI call: pkf (+ll), pks and pkf_backward sharing their passes over the LGSSM
            (fms, fPs, ll), (sms, sPs), grads = ops.pkfs_grad(Pinf, Fs, Qs, H, R, y_dev, g_ll)
            return ll, sms, sPs, grads
        return shard.filter_smoother_grad(Pinf, Fs, Qs, H, R, y_dev, g_ll)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ---------------------------------------------------------------------
    for _ in range(W):
        out = device_step()
    barrier()
    # time sharding with the peer exchange is kernels only (no NCCL call, no host synchronisation): the whole step is
    # captured once and replayed as ONE CUDA graph launch, which takes the per-launch host work of the ~20 small
    # operations around the three exchanges off the critical path of every rank
    graphed = False
    eager_step = device_step
    if shard is not None and xchg is not None and os.environ.get("PSSGP_GRAPH", "1") == "1":
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(2):
                    out = eager_step()
            torch.cuda.current_stream(dev).wait_stream(side)
            barrier()
            graph = torch.cuda.CUDAGraph()
            lc0 = h.launch_count()
            with torch.cuda.graph(graph, stream=side):
                graph_out = eager_step()
            graph_launches = h.launch_count() - lc0   # kernels of this library recorded into the graph
            ok = torch.tensor([1.0], device=dev)
        except Exception as e:
            print(f"[bench] CUDA-graph capture of the sharded step failed ({e!r}); eager launches", file=sys.stderr)
            ok = torch.tensor([0.0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok) == 1.0:
            graphed = True

            def device_step():
                graph.replay()
                return graph_out
            for _ in range(W):
                out = device_step()
            barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = h.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K):
        out = device_step()
    e1.record()
    barrier()
    ms_dev = e0.elapsed_time(e1) / K
    launches = graph_launches if graphed else (h.launch_count() - launches0) // max(K, 1)
    device_step = eager_step   # the per-kernel event pass below needs eager launches
    # the same K steps once more with a CUDA-event pair around every kernel launch (library option "timing"):
    # per-kernel durations for the roofline; kept out of the loop above because every event record costs
    # a few microseconds of stream time
    h.set_option("timing", 1)
    h.timing_report()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e2.record()
    for _ in range(K):
        out = device_step()
    e3.record()
    barrier()
    ms_instrumented = e2.elapsed_time(e3) / K
    ktimes = h.timing_report()
    h.set_option("timing", 0)
    clock_info = clocks.stop() if rank == 0 else None
    xchg_fail = xchg.failures() if xchg is not None else 0
    if dist is not None:
        tt = torch.tensor([ms_dev], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_dev = float(tt)
    value = n * world / (ms_dev * 1e-3)

    # ---- end-to-end through the model API with host buffers --------------------------------------------
    e2e = None
    if world == 1:
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        t_pin, y_pin = pin(t_host[:, None]), pin(y_host[:, None])
        q_host = t_host + 0.002  # N query times interleaved with the training grid
        q_pin = pin(q_host[:, None])
        model = StateSpaceGP((t_pin, y_pin), kernels.Matern52(1.0, 1.0), noise_variance=NOISE, parallel=True,
                             max_parallel=2 * n)

        mean_pin = torch.empty((n, 1), dtype=torch.float64).pin_memory()
        var_pin = torch.empty((n, 1), dtype=torch.float64).pin_memory()

        def e2e_step():
            model.data = (t_pin, y_pin)  # H2D of this step's inputs from pinned memory
            # queries from pinned memory, posterior mean / variance read back into pinned memory (D2H of the result);
            # non_blocking: the read-back runs on a copy stream while the training step below computes
            mean, var = model.predict_f(q_pin, out=(mean_pin, var_pin), non_blocking=True)
            ll = model.maximum_log_likelihood_objective()
            grads = torch.autograd.grad(ll, model.trainable_variables)
            model.synchronize()          # posterior buffers valid
            return float(ll.detach()), [float(g) for g in grads], mean, var

        for _ in range(W):
            r = e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(K):
            r = e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / K
        e2e = {"value": n / dt, "unit": UNIT, "h2d_bytes_per_step": int(8 * 2 * n + 8 * n),
               "d2h_bytes_per_step": int(8 * 2 * n + 8 * 4), "ms_per_step": dt * 1e3,
               "api": "StateSpaceGP.data= (pinned t, y) ; predict_f(N pinned queries, out=pinned mean/var, non_blocking=True) ; "
                      "maximum_log_likelihood_objective + autograd.grad ; model.synchronize()"}
    else:
        # time-sharded: every rank holds ITS shard of the series in pinned host memory; one step = H2D of the shard,
        # discretise, sharded filter + smoother + gradient (exchanges over NVLink), gradient pulled back to the SDE,
        # D2H of (ll, gradient) and of the posterior mean / variance at the shard's times (dist.TimeShard.series_step)
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        t_pin, y_pin = pin(t_host[lo:hi]), pin(y_host[lo:hi])
        mean_pin = torch.empty(n, dtype=torch.float64).pin_memory()
        var_pin = torch.empty(n, dtype=torch.float64).pin_memory()

        def e2e_step():
            return shard.series_step(F, Pinf, H, R, t_pin, y_pin, t_prev, out=(mean_pin, var_pin))

        for _ in range(W):
            r = e2e_step()
        barrier()
        # the same step as one CUDA graph per rank (H2D of the shard ... D2H of the results): kernels, copies and peer
        # exchanges replay from one launch; falls back to the eager step if any rank cannot capture
        e2e_graphed = False
        if xchg is not None and os.environ.get("PSSGP_GRAPH", "1") == "1":
            try:
                replay = shard.capture_series_step(F, Pinf, H, R, t_pin, y_pin, t_prev, (mean_pin, var_pin))
                okc = torch.tensor([1.0], device=dev)
            except Exception as e:
                print(f"[bench] CUDA-graph capture of the series step failed ({e!r}); eager step", file=sys.stderr)
                okc = torch.tensor([0.0], device=dev)
            dist.all_reduce(okc, op=dist.ReduceOp.MIN)
            if float(okc) == 1.0:
                r_eager, eager_fn = r, e2e_step
                for _ in range(W):
                    r = replay()
                barrier()
                same = torch.tensor([1.0 if abs(float(r[0][0]) - float(r_eager[0][0])) <= 1e-12 * abs(float(r_eager[0][0]))
                                     else 0.0], device=dev)
                dist.all_reduce(same, op=dist.ReduceOp.MIN)
                if float(same) == 1.0:
                    e2e_step, e2e_graphed = replay, True
                else:   # never seen; the eager step is the reference behaviour
                    print("[bench] graph replay of the series step differs from the eager step; eager step", file=sys.stderr)
                    e2e_step = eager_fn
        t0 = time.perf_counter()
        for _ in range(K):
            r = e2e_step()
        barrier()
        dt_t = torch.tensor([(time.perf_counter() - t0) / K], dtype=torch.float64, device=dev)
        dist.all_reduce(dt_t, op=dist.ReduceOp.MAX)
        dt = float(dt_t)
        e2e = {"value": n * world / dt, "unit": UNIT, "h2d_bytes_per_step": int(8 * 2 * n * world),
               "d2h_bytes_per_step": int((8 * 2 * n + 8 * (2 + 2 * d * d + d)) * world), "ms_per_step": dt * 1e3,
               "api": "dist.TimeShard.series_step: per rank pinned (t, y) shard -> device, discretise, sharded filter + "
                      "smoother + gradient, -> host (ll, dF, dPinf, dH, dR) + pinned posterior mean/var of the shard; "
                      "bytes are totals over the ranks, time is the max over ranks",
               "finite": bool(torch.isfinite(r[0][0])), "cuda_graph": e2e_graphed}
    # ---- the other BASELINE configurations + the sharded-vs-unsharded check (all ranks take part) ----------
    peaks_all = {}
    try:
        peaks_all = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_all = float(peaks_all.get("hbm_gbs", 6650.0))
    FP64_PEAK_TFLOPS = 37.1  # DMMA full-chip rate, scripts/sm_probe.cu on this pool's B200 (profiles/r02_sm_probe.txt)
    del Fs, Qs, out
    torch.cuda.empty_cache()
    extras, check = {}, None
    if not args.no_extra:
        Kx, Wx = max(2, min(K, 3)), 3
        for name, cfg in extra_configs(world).items():
            if args.only and name not in args.only.split(","):
                continue
            try:
                extras[name] = run_extra_config(name, cfg, world, rank, dev, dist, barrier, torch, kernels, ops, pdist, h,
                                                hbm_all, FP64_PEAK_TFLOPS, Wx, Kx, not args.no_cpu_baseline, xchg)
            except Exception as e:  # keep the headline line alive
                extras[name] = {"error": repr(e)}
                torch.cuda.empty_cache()
        if not args.only or "grid_1024" in args.only.split(","):
            try:
                extras["grid_1024"] = run_grid_config(world, rank, dev, dist, barrier, torch, kernels, ops, Wx)
            except Exception as e:
                extras["grid_1024"] = {"error": repr(e)}
        if world == 1 and (not args.only or "model_e2e" in args.only.split(",")):
            try:
                extras["model_e2e"] = run_model_e2e(dev, torch, kernels, Kx)
            except Exception as e:
                extras["model_e2e"] = {"error": repr(e)}
                torch.cuda.empty_cache()
        if world > 1:
            try:
                check = sharded_check(world, rank, dev, dist, torch, kernels, ops, pdist, xchg)
                check["exchange"] = xchg_kind
            except Exception as e:
                check = {"error": repr(e)}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ------------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    s = 8
    alg_bytes = {  # algorithmic bytes per time step of each kernel
        "pkf_reduce": s * (2 * d * d + 1), "pkf_apply": s * (3 * d * d + d + 1), "pkf_mid": 0,
        "pkf_apply_fused": s * (3 * d * d + d + 1), "pks_bwd_apply_fused": s * (8 * d * d + 2 * d + 1),
        "pks_reduce": s * (3 * d * d + d), "pks_apply": s * (4 * d * d + 2 * d), "pks_mid": 0,
        "pkf_bwd_reduce": s * (3 * d * d + d + 1), "pkf_bwd_apply": s * (5 * d * d + d + 1), "pkf_bwd_mid": 0,
    }
    per_kernel = {k: {"launches": c, "avg_us": 1e3 * ms / max(c, 1)} for k, (c, ms) in ktimes.items()}
    dom = max(ktimes.items(), key=lambda kv: kv[1][1])[0] if ktimes else None
    roofline = None
    if dom is not None:
        avg_s = ktimes[dom][1] / max(ktimes[dom][0], 1) * 1e-3
        achieved = alg_bytes.get(dom, 0) * n / avg_s / 1e9
        # measured DRAM traffic of that kernel per launch (ncu --set full capture of the same workload, committed
        # under profiles/); null when the capture does not cover this kernel or this n
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            if int(tr.get("n", 0)) == n and dom in tr["kernels"]:
                traffic = tr["kernels"][dom]["dram_bytes_read"] + tr["kernels"][dom]["dram_bytes_write"]
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                    "frac": achieved / hbm_peak, "traffic": traffic, "alg_bytes_per_launch": alg_bytes.get(dom, 0) * n,
                    "peak_source": peak_src,
                    "share_of_step": ktimes[dom][1] / sum(v[1] for v in ktimes.values())}
    if e2e is not None and e2e.get("value"):
        # the north star's compulsory-traffic floor beside the API-boundary roofline: a fully fused path would only read
        # (t, y) and write the smoothed moments, s (2 + d^2 + d) bytes per step
        floor_bytes = s * (2 + d * d + d)
        e2e["compulsory_floor"] = {"bytes_per_timestep": floor_bytes,
                                   "hbm_floor_ms_per_step": floor_bytes * n / (hbm_peak * 1e9) * 1e3,
                                   "achieved_gbs_per_gpu": floor_bytes * e2e["value"] / world / 1e9,
                                   "pcie_bytes_per_step_per_gpu": (e2e["h2d_bytes_per_step"] + e2e["d2h_bytes_per_step"]) // world}
    stage_bytes = s * (12 * d * d + 4 * d + 2)
    step_roof = {"alg_bytes_per_timestep": stage_bytes, "achieved_gbs": stage_bytes * n * world / (ms_dev * 1e-3) / 1e9,
                 "frac_of_hbm_peak": stage_bytes * n * world / (ms_dev * 1e-3) / 1e9 / (hbm_peak * world)}

    # ---- CPU baseline (oracle port) on the host cores ------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sample_n = min(n, 1_000_000)
        val, dt, cores = cpu_reference_arm(1, 1, sample_n)
        cpu = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"the full workload once ({sample_n} time steps, {dt:.1f} s): oracle port = torch-CPU "
                         f"restatement of the reference's pkf + autograd gradient + pks; TF reference not installable"}
        try:
            cpu["sequential_kf_ks"] = sequential_numba_baseline()
        except Exception as e:  # pragma: no cover
            cpu["sequential_kf_ks"] = {"error": repr(e)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "Matern52 single series N=1e6 per GPU (one series of n_gpus*1e6 steps, time-sharded), "
                               "FP64, log-lik + gradient + RTS smoother (configs[1])",
                   "n_per_gpu": n, "state_dim": d, "l2_policy": "inputs larger than L2 (Fs+Qs+y = 152 MB, "
                   "plus 96+96+144 MB of outputs per step; L2 = 126 MB)",
                   "parallelism": "time-sharded x%d" % world if world > 1 else "single GPU", "exchange": xchg_kind,
                   "cuda_graph": graphed, "exchange_failures": xchg_fail},
        "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "step_roofline": step_roof,
        "kernels": per_kernel, "ms_per_step_with_kernel_events": ms_instrumented, "cpu_baseline": cpu, "clocks": clock_info,
        "configs": extras, "sharded_check": check,
    }
    print_line(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
