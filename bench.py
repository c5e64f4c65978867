#!/usr/bin/env python
"""Benchmark of the hot path: fitMPS two-site sweep throughput in sample-bonds/s (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A *step* is one full sweep (backward + forward half-sweep = 2*(T-1) bond updates, each = loss +
gradient + optimiser step + truncated SVD + environment update) over the whole synthetic training
set.  Workload at N=1: BASELINE.json configs[1] (synthetic trendy-sine, N=100k, T=100, d=12,
chi_max=40, Legendre, KLD, TSGO, eta=0.01, log_level=0 -- SURVEY 8d (B)).  With --gpus N each rank
holds that many samples (weak scaling); the only collective is the per-bond gradient all-reduce.

`value`  : device-resident throughput (inputs in HBM before the timed region), CUDA events.
`e2e`    : same metric through the host-facing API with HOST buffers: every step copies the scaled
           series and the cores host->device, runs one sweep and reads the cores back.
`roofline`: the dominant kernel (bond_grad_kernel, FP64 DMMA GEMM): algorithmic flops
           2*N*(d*chi_l)*(d*chi_r) per launch / CUDA-event time of that kernel inside the timed region.
`cpu_baseline`: the restated reference algorithm (oracle/, numpy + C hot loop) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(name="trendy_sine_N100k_T100_d12_chi40", N=100_000, T=100, d=12, chi_max=40, eta=0.01, seed=1,
                periods=((12.0, 15.0), (16.0, 19.0)), slopes=(-3.0, 0.0, 3.0), sigma=0.1, chi_init=4)
FP64_PEAK_TFLOPS = 35.4      # cuBLAS DGEMM measured on this pool (profiles/r01_dgemm_calib.txt); MEASURED_PEAKS.json has no FP64 entry
FP64_PEAK_NOTE = "fallback: same-pool cuBLAS DGEMM 8192^3 = 35.4 TFLOP/s (DMMA issue peak 37.1); MEASURED_PEAKS.json carries no FP64 figure"
# dram__bytes_read.sum + dram__bytes_write.sum of one steady-state launch of the gradient kernel at config B shapes
# (ncu --set full, profiles/r01_bond_grad_kr_ncu_full.txt); algorithmic bytes are 84 MB, the kernel is tensor bound
NCU_TRAFFIC_BYTES = 184.6e6
NCU_TRAFFIC_NOTE = "profiles/r01_bond_grad_kr_ncu_full.txt (ncu --set full, launch 30 of a config-B sweep): 175.5 MB read + 9.1 MB written"


def trendy_sine(T, n, period, slopes, sigma, rng):
    """x_t = sin(2 pi t / tau + psi) + m t / T + sigma n_t  (reference src/Simulation/toy_data.jl:53-85)."""
    tau = rng.uniform(period[0], period[1], size=n)
    m = rng.choice(np.asarray(slopes, dtype=np.float64), size=n)
    psi = rng.uniform(0.0, 2 * np.pi, size=n)
    t = np.arange(1, T + 1, dtype=np.float64)
    X = np.sin(2 * np.pi / tau[:, None] * t[None, :] + psi[:, None]) + m[:, None] * t[None, :] / T
    X += sigma * rng.standard_normal((n, T))
    return X


def make_data(N, T, seed, w=WORKLOAD):
    rng = np.random.default_rng(seed)
    n0 = N // 2
    X = np.concatenate([trendy_sine(T, n0, w["periods"][0], w["slopes"], w["sigma"], rng),
                        trendy_sine(T, N - n0, w["periods"][1], w["slopes"], w["sigma"], rng)], axis=0)
    y = np.concatenate([np.zeros(n0, dtype=np.int64), np.ones(N - n0, dtype=np.int64)])
    return X, y


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    td = None
    if world > 1:
        import torch
        import torch.distributed as td_
        torch.cuda.set_device(local)
        td_.init_process_group("nccl", device_id=torch.device("cuda", local))
        td = td_
    return rank, world, local, td


def barrier(td, local):
    if td is not None:
        import torch
        td.barrier(device_ids=[local])
        torch.cuda.synchronize()


def max_over_ranks(td, local, v):
    if td is None:
        return v
    import torch
    t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{local}")
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t.item())


# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(cores, Xs_sorted, counts, w, n_sample, n_bonds, nthreads, d):
    """Restated reference algorithm on the host: `n_bonds` bond updates of a backward half-sweep on
    `n_sample` samples (drawn evenly from both classes) starting from the given (trained,
    chi-saturated) cores.  Returns (sample_bonds_per_s, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mpstime_oracle as o            # noqa: E402  (bench-only use of the oracle: the CPU baseline)
    import bond_ref                       # noqa: E402
    ns = [n_sample // 2, n_sample - n_sample // 2]
    off = np.concatenate([[0], np.cumsum(counts)])
    idx = np.concatenate([np.arange(off[c], off[c] + min(ns[c], counts[c])) for c in range(len(counts))])
    csub = np.array([min(ns[c], counts[c]) for c in range(len(counts))])
    phi = o.encode(Xs_sorted[:, idx].T, d)
    T = phi.shape[1]

    def lg(B, L, R, xl, xr, cnts, train_sep=False):
        return bond_ref.loss_grad_kld(B, L, R, xl, xr, cnts, train_sep, nthreads)

    orig = o.loss_grad_KLD
    o.loss_grad_KLD = lg                  # C restatement of the literal per-sample loop (oracle/bond_ref.c)
    try:
        t0 = time.time()
        o.fit_sweeps(cores, phi, csub, nsweeps=1, chi_max=w["chi_max"], eta=w["eta"], max_bonds=n_bonds)
        dt = time.time() - t0
    finally:
        o.loss_grad_KLD = orig
    return len(idx) * min(n_bonds, 2 * (T - 1)) / dt, dt


def run_reference(args):
    """--impl reference: the reference's own algorithm (CPU) on this box's host cores.  The reference is
    Julia (absent here), so this runs the oracle port with every host thread on the hot loop."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import mpstime_jl_b200 as m           # only for the host-side preprocessing + start MPS
    w = WORKLOAD
    ncores = os.cpu_count() or 1
    n_sample = 1024
    X, y = make_data(n_sample, w["T"], w["seed"])
    opts = m.MPSOptions(d=w["d"], chi_max=w["chi_max"], eta=w["eta"])
    Xs, _ = m.transform_train_data(X.T, opts)
    Xs_sorted, _, ys, _, classes, counts = m.sort_by_class(Xs, X, y)
    # chi-saturated starting state: random MPS already at chi_max (a trained MPS sits there after one sweep)
    cores = m.generate_starting_mps(w["chi_max"], w["T"], w["d"], 2, seed=1234)
    n_bonds = 24
    rates, times = [], []
    for it in range(args.warmup + args.steps):
        r, dt = cpu_reference_rate(cores, Xs_sorted, counts, w, n_sample, n_bonds, ncores, w["d"])
        if it >= args.warmup:
            rates.append(r); times.append(dt)
    v = float(np.mean(rates))
    sample = f"{n_bonds} bond updates (backward half-sweep from the right edge, chi<= {w['chi_max']}) on {n_sample} samples per step"
    print(json.dumps({
        "impl": "reference", "metric": "fitMPS sample-bonds/sec per sweep", "value": v, "unit": "sample-bonds/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["name"], "T": w["T"], "d": w["d"], "chi_max": w["chi_max"], "sample": sample},
        "cpu_baseline": {"value": v, "unit": "sample-bonds/s", "cores": ncores, "kind": "port", "sample": sample,
                         "note": "restated reference algorithm (numpy + C hot loop, LAPACK gesdd), not Julia; the reference's own loop is single-threaded"},
        "e2e": {"value": v, "unit": "sample-bonds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def impute_bench(m, ctx, args):
    """Second half of BASELINE.json's metric: MPS_impute instances/s (configs[3] shape: T=256, d=16, chi=64,
    128 contiguous missing sites with a random start, median on the dx=1e-4 grid of 20 001 points) on a bounded
    batch of instances.  The class MPS is a random chi=64 MPS (imputation cost does not depend on the values)."""
    T, d, chi, K = 256, 16, 64, 128
    n = args.impute_instances
    rng = np.random.default_rng(3)
    cores = m.generate_starting_mps(chi, T, d, 1, seed=7)
    ctx.model_init(T, 1, d, chi)
    ctx.set_cores(cores)
    t = np.arange(1, T + 1)
    X = np.sin(2 * np.pi * t[:, None] / 24.0 + rng.uniform(0, 2 * np.pi, n)[None, :]) * 0.45 + 0.04 * rng.standard_normal((T, n))
    X = np.clip(X, -1, 1)
    mask = np.zeros((T, n), dtype=np.uint8)
    for i, s0 in enumerate(rng.integers(0, T - K + 1, n)):
        mask[s0:s0 + K, i] = 1
    grid = m.make_grid((-1.0, 1.0), 1e-4)
    ctx.impute_batch(0, X[:, :256], mask[:, :256], grid)                   # warm-up (>= one instance per SM: sizes the work buffers)
    t0 = time.time()
    out = ctx.impute_batch(0, X, mask, grid, method="median")
    dt = time.time() - t0
    assert np.isfinite(out).all()
    return {"metric": "MPS_impute instances/sec", "value": n / dt, "unit": "instances/s", "instances": n, "T": T, "d": d,
            "chi": chi, "missing": K, "grid_points": len(grid), "method": "median", "seconds": dt,
            "note": "host buffers in/out through mpst_impute_batch (e2e); bounded batch of configs[3]"}


# ------------------------------------------------------------------------------------------------
def pinned_copy(a):
    """numpy view of a page-locked copy of `a` (torch is only the pinned allocator here)."""
    import torch
    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
    out = t.numpy()
    out[...] = a
    _PINNED.append(t)
    return out


_PINNED = []


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--samples-per-gpu", dest="n", type=int, default=0, help="override samples per GPU (debug)")
    ap.add_argument("--series-length", dest="t", type=int, default=0, help="override series length (debug)")
    ap.add_argument("--local-dim", dest="d", type=int, default=0, help="override d (debug: config C shapes)")
    ap.add_argument("--chi-max", dest="chi", type=int, default=0, help="override chi_max (debug: config C shapes)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-impute", action="store_true")
    ap.add_argument("--impute-instances", type=int, default=2048)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank, world, local, td = dist_setup(args.gpus)
    import mpstime_jl_b200 as m
    w = dict(WORKLOAD)
    if args.n:
        w["N"] = args.n
    if args.t:
        w["T"] = args.t
    if args.d:
        w["d"] = args.d
    if args.chi:
        w["chi_max"] = args.chi
    if args.n or args.t or args.d or args.chi:
        w["name"] = "trendy_sine_N%d_T%d_d%d_chi%d(debug override)" % (w["N"], w["T"], w["d"], w["chi_max"])
    N_local, T, d, chi_max = w["N"], w["T"], w["d"], w["chi_max"]
    N_global = N_local * world
    # every rank generates only its own shard: per class, rank r holds samples [r*N_local/2, (r+1)*N_local/2)
    X, y = make_data(N_local, T, w["seed"] + 1000 * rank)
    opts = m.MPSOptions(d=d, chi_max=chi_max, eta=w["eta"], log_level=0, verbosity=-1, nsweeps=1)
    # each rank normalises its own shard (same distribution on every rank; a real multi-GPU fitMPS call
    # computes the statistics once on the host before sharding, mpstime.jl_b200/api.py)
    Xs, _ = m.transform_train_data(X.T, opts)
    Xs_sorted, _, ys, _, classes, counts = m.sort_by_class(Xs, X, y)
    counts_global = counts * world

    ctx = m.Context(local)
    if world > 1:
        m.dist.init_comm(ctx)
    cores0 = m.generate_starting_mps(w["chi_init"], T, d, 2, seed=1234)
    topts = m.make_opts(chi_max=chi_max, eta=w["eta"])

    # ---- device-resident arm --------------------------------------------------------------------
    ctx.train_load_x(Xs_sorted, counts, d, chi_max, n_global=N_global, counts_global=counts_global)
    ctx.set_cores(cores0)
    if args.warmup:
        ctx.sweep(topts, args.warmup, record=False)
    ctx.profile_enable(True)
    ctx.profile_reset()
    clocks = ClockSampler(local)
    barrier(td, local)
    if rank == 0:
        clocks.start()
    l0 = ctx.launch_count()
    ctx.timer_start()
    t_wall = time.time()
    lo, gn, chi = ctx.sweep(topts, args.steps, record=True)      # K sweeps: build LE, K x (backward + forward), normalize!
    chis = [chi]
    ms = ctx.timer_stop()
    wall = time.time() - t_wall
    barrier(td, local)
    clk = clocks.stop() if rank == 0 else None
    launches = ctx.launch_count() - l0
    prof = ctx.profile_get()
    ctx.profile_enable(False)
    ms = max_over_ranks(td, local, ms)
    sample_bonds = args.steps * 2 * (T - 1) * N_global
    value = sample_bonds / (ms * 1e-3)

    # ---- end-to-end arm: host buffers in, cores out, every step -----------------------------------
    cores_host = ctx.get_cores()
    h2d = Xs_sorted.nbytes + sum(c.nbytes for c in cores_host)
    d2h = sum(c.nbytes for c in cores_host) + 2 * (T - 1) * 20
    # the step's input lives in pinned host memory, already in the wire layout (Julia column-major T x N == C-order
    # (N, T)), so train_load_x hands the pinned pointer straight to the C ABI (no host-side transpose)
    X_pinned = pinned_copy(np.ascontiguousarray(Xs_sorted.T)).T

    def e2e_step(cores):
        ctx.train_load_x(X_pinned, counts, d, chi_max, n_global=N_global, counts_global=counts_global)
        ctx.set_cores(cores)
        ctx.sweep(topts, 1, record=True)
        return ctx.get_cores()

    if args.warmup:
        cores_host = e2e_step(cores_host)              # untimed: first-touch of the pinned buffer, allocator state
    barrier(td, local)
    t0 = time.time()
    for _ in range(args.steps):
        cores_host = e2e_step(cores_host)
    barrier(td, local)
    e2e_s = max_over_ranks(td, local, time.time() - t0)
    e2e = sample_bonds / e2e_s

    if td is not None:
        td.barrier(device_ids=[local])
        td.destroy_process_group()
    if rank != 0:
        return
    gk_ms, gk_n, gk_fl = prof["grad_kernel"]
    achieved = gk_fl / (gk_ms * 1e-3) / 1e12 if gk_ms > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "bond_grad_kr_kernel", "achieved": achieved, "peak": FP64_PEAK_TFLOPS,
                "unit": "TFLOP/s", "frac": achieved / FP64_PEAK_TFLOPS, "traffic": NCU_TRAFFIC_BYTES,
                "traffic_source": NCU_TRAFFIC_NOTE, "peak_source": FP64_PEAK_NOTE,
                "launches": gk_n, "avg_launch_ms": gk_ms / max(gk_n, 1),
                "algorithmic_flops_per_launch": gk_fl / max(gk_n, 1)}
    breakdown = {k: {"ms": round(v[0], 3), "n": v[1]} for k, v in prof.items() if v[1]}
    out = {
        "metric": "fitMPS sample-bonds/sec per sweep", "value": value, "unit": "sample-bonds/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["name"], "N_per_gpu": N_local, "N_global": N_global, "T": T, "d": d, "chi_max": chi_max,
                   "encoding": "Legendre_No_Norm", "loss": "KLD", "bbopt": "TSGO", "eta": w["eta"], "update_iters": 1,
                   "parallelism": f"sample-sharded x{world}, one NCCL all-reduce per bond" if world > 1 else "single GPU",
                   "l2_policy": "inputs larger than L2: env cache %.1f GB + series %.2f GB per GPU" % (
                       (T * N_local * chi_max * 8) / 1e9, Xs_sorted.nbytes / 1e9),
                   "mean_chi": float(np.mean(np.concatenate(chis)))},
        "clocks": clk, "gpu_launches": int(launches),
        "e2e": {"value": e2e, "unit": "sample-bonds/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "roofline": roofline, "device_time_breakdown_ms": breakdown, "wall_s_timed": wall,
    }
    if not args.no_impute:
        try:
            out["impute"] = impute_bench(m, ctx, args)
        except Exception as e:
            out["impute"] = {"value": None, "error": repr(e)}
    if not args.no_cpu_baseline:
        try:
            n_sample, n_bonds = 1024, 24
            r, dt = cpu_reference_rate(cores_host, Xs_sorted, counts, w, n_sample, n_bonds, 1, d)
            out["cpu_baseline"] = {"value": r, "unit": "sample-bonds/s", "cores": 1, "kind": "port", "seconds": dt,
                                   "sample": f"{n_bonds} bond updates of a backward half-sweep on {n_sample} samples, same T/d/chi, "
                                             "trained cores; 1 thread like the reference's own hot loop (numpy + C port, not Julia)"}
        except Exception as e:          # the baseline is reporting only
            out["cpu_baseline"] = {"value": None, "error": repr(e)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
