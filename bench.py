#!/usr/bin/env python
"""Benchmark of the hot path: fitMPS two-site sweep throughput in sample-bonds/s and MPS_impute instances/s
(BASELINE.json).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload C|B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (default, every N): BASELINE.json configs[2], the north star -- synthetic 2-class trendy-sine series,
N_global = 1 000 000, T = 256, d = 16, chi_max = 64, Legendre, KLD, TSGO, eta = 0.01 -- as STRONG scaling: the
1 M samples are sharded over the N ranks (per class, contiguous ranges), one NCCL all-reduce of [gradient | loss] per
bond, SVD replicated.  It fits one B200 (131 GB of environments), so N = 1 runs it too and the driver's 1/2/4/8 curve
is one workload.  If it does not fit the GPU found at N = 1, the run falls back to configs[1] and says so.

A *step* is a fixed block of 32 consecutive bond updates of the running sweep (each = flatten + loss + gradient +
all-reduce + optimiser step + truncated SVD + environment update): a whole sweep is 510 bonds (46 s on one GPU), too
long for W + K steps.  The blocks continue one another in sweep order (backward, forward, backward ...) after ONE
full untimed sweep that brings every link to chi_max, so the timed bonds are real, chi-saturated sweep bonds.

`value`  : device-resident throughput (series and environments in HBM before the timed region), CUDA events, max over
           ranks: K * 32 * N_global / time.
`e2e`    : the same metric through the host-facing call sequence of one fitMPS sweep with HOST buffers: copy the scaled
           series (pinned) and the cores host->device, build the environments, run ONE WHOLE sweep (510 bonds) with
           per-bond loss read-back, normalise, read the cores back.  One such call is timed (it is 16 steps' worth of
           bonds); bytes are per call.
`roofline`: the dominant kernel (the bond-gradient DMMA GEMM; its name comes from the library, not a constant):
           algorithmic flops 2*N_local*(d*chi_l)*(d*chi_r) per launch / CUDA-event time of that kernel inside the timed
           region; peak = FP64 GEMM rate measured live on this GPU (torch.matmul float64 = cuBLAS DGEMM).
`cpu_baseline`: the restated reference algorithm (oracle/: numpy + the C loop of oracle/bond_ref.c) on a bounded sample.
`config_B`: (N = 1 only) a short run of configs[1] (N = 100k, T = 100, d = 12, chi = 40; step = whole sweep).
`impute` : configs[3] (MPS_impute, T = 256, d = 16, chi = 64, 128 contiguous missing, dx = 1e-4 grid, median and ITS),
           instances sharded over the ranks with no communication, with its own roofline and CPU figure.
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

WORKLOADS = {
    "C": dict(key="C", name="north_star_trendy_sine_N1M_T256_d16_chi64", N=1_000_000, T=256, d=16, chi_max=64, eta=0.01,
              seed=2, periods=((20.0, 30.0), (32.0, 48.0)), slopes=(-3.0, 0.0, 3.0), sigma=0.1, chi_init=4,
              bonds_per_step=32, scaling="strong"),
    "B": dict(key="B", name="trendy_sine_N100k_T100_d12_chi40", N=100_000, T=100, d=12, chi_max=40, eta=0.01, seed=1,
              periods=((12.0, 15.0), (16.0, 19.0)), slopes=(-3.0, 0.0, 3.0), sigma=0.1, chi_init=4,
              bonds_per_step=0, scaling="weak"),          # 0 = whole sweep per step
}
METRIC = "fitMPS sample-bonds/sec per sweep"
FP64_FALLBACK_TFLOPS = 35.4   # same-pool cuBLAS DGEMM 8192^3 measured in round 1 (profiles/r01_dgemm_calib.txt)
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "kernel_traffic.json")   # ncu dram bytes per launch, by kernel + shape


def trendy_sine(T, n, period, slopes, sigma, rng):
    """x_t = sin(2 pi t / tau + psi) + m t / T + sigma n_t  (reference src/Simulation/toy_data.jl:53-85)."""
    tau = rng.uniform(period[0], period[1], size=n)
    m = rng.choice(np.asarray(slopes, dtype=np.float64), size=n)
    psi = rng.uniform(0.0, 2 * np.pi, size=n)
    t = np.arange(1, T + 1, dtype=np.float64)
    X = np.sin(2 * np.pi / tau[:, None] * t[None, :] + psi[:, None]) + m[:, None] * t[None, :] / T
    X += sigma * rng.standard_normal((n, T))
    return X


def make_data(N, T, seed, w):
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


def reduce_over_ranks(td, local, v, op="max"):
    if td is None:
        return v
    import torch
    t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{local}")
    td.all_reduce(t, op=td.ReduceOp.MAX if op == "max" else td.ReduceOp.SUM)
    return float(t.item())


def measure_fp64_gemm_peak(local):
    """FP64 GEMM rate of this GPU, live: torch.matmul float64 8192^3 (cuBLAS DGEMM), best of 6, CUDA events --
    the same way MEASURED_PEAKS.json measures bf16; it carries no FP64 figure."""
    try:
        import torch
        dev = torch.device("cuda", local)
        n = 8192
        a = torch.randn(n, n, dtype=torch.float64, device=dev)
        b = torch.randn(n, n, dtype=torch.float64, device=dev)
        c = torch.empty(n, n, dtype=torch.float64, device=dev)
        torch.matmul(a, b, out=c)
        torch.cuda.synchronize(dev)
        best = 0.0
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b, out=c)
            e1.record()
            torch.cuda.synchronize(dev)
            best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        del a, b, c
        torch.cuda.empty_cache()
        return best, "measured live: torch.matmul float64 8192^3 (cuBLAS DGEMM), best of 6, CUDA events; MEASURED_PEAKS.json has no FP64 entry"
    except Exception as e:            # the peak is a denominator only; never fail the run for it
        return FP64_FALLBACK_TFLOPS, f"fallback {FP64_FALLBACK_TFLOPS} TFLOP/s (round-1 cuBLAS DGEMM calibration); live probe failed: {e!r}"


def measured_hbm_peak():
    try:
        j = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(j["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6550.0, "fallback 6550 GB/s (B200_PROFILING.md)"


GRAD_KERNEL_NAMES = {1: "bond_grad_kr_kernel", 2: "bond_grad_kernel"}


def kernel_traffic(name, d, chi, n_local):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this
    kernel at this bond shape (profiles/kernel_traffic.json keeps bytes per sample of the captured launch; the traffic
    of these kernels is proportional to the sample count), or None when no capture exists for it."""
    try:
        tab = json.load(open(TRAFFIC_FILE))
        e = tab.get(f"{name}:d{d}:chi{chi}")
        return (float(e["bytes_per_sample"]) * n_local, e["source"]) if e else (None, "no ncu capture for this kernel/shape")
    except Exception:
        return None, "profiles/kernel_traffic.json missing"


# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(cores, Xs_sorted, counts, w, n_sample, n_bonds, nthreads):
    """Restated reference algorithm on the host (oracle/: numpy + the C loop oracle/bond_ref.c): `n_bonds` bond
    updates of a backward half-sweep on `n_sample` samples (drawn evenly from both classes) from the given cores.
    The loss/gradient loop and the environment update are O(N) per bond; the LAPACK SVD is N-independent (amortised
    over 1 M samples in the named config), so the two are timed apart.  Only chi-saturated bonds count towards the rate
    (the two cheap bonds at the chain edge only move the label inwards): all but 4 of the 510 bonds of a sweep are
    saturated.  Returns a dict."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mpstime_oracle as o            # noqa: E402  (bench-only use of the oracle: the CPU baseline)
    import bond_ref                       # noqa: E402
    d = w["d"]
    ns = [n_sample // 2, n_sample - n_sample // 2]
    off = np.concatenate([[0], np.cumsum(counts)])
    idx = np.concatenate([np.arange(off[c], off[c] + min(ns[c], counts[c])) for c in range(len(counts))])
    csub = np.array([min(ns[c], counts[c]) for c in range(len(counts))])
    phi = o.encode(Xs_sorted[:, idx].T, d)
    N, T = phi.shape[0], phi.shape[1]
    cs = [c.copy() for c in cores]
    ones = np.ones((N, 1))
    t_env0 = time.time()
    LE = o.construct_caches(cs, phi, going_left=True)
    t_build = time.time() - t_env0
    RE = {}
    t_loop = t_svd = t_env = 0.0
    work = 0.0
    done = 0
    for j in range(T - 2, max(T - 2 - n_bonds, -1), -1):
        l, r = j, j + 1
        L = LE[l - 1] if l > 0 else ones
        R = RE[r + 1] if r < T - 1 else ones
        B, dims = o.flatten_bt(cs[l], cs[r])
        sat = dims[0] == min(w["chi_max"], d ** min(l, 30)) and dims[2] == min(w["chi_max"], d ** min(T - 1 - r, 30)) and \
            dims[0] * dims[2] == max(c.shape[0] for c in cs) * max(c.shape[2] for c in cs)
        t0 = time.time()
        lo, G = bond_ref.loss_grad_kld(B, L, R, phi[:, l], phi[:, r], csub, False, nthreads)      # Loss_Grad_KLD
        if sat:
            t_loop += time.time() - t0
        Bn = B - w["eta"] * G / np.linalg.norm(G)
        Bn /= np.linalg.norm(Bn)
        t0 = time.time()
        cl, cr, S = o.decompose_bt(Bn, dims, True, w["chi_max"], 1e-10)                           # LAPACK gesdd
        if sat:
            t_svd += time.time() - t0
        cs[l], cs[r] = cl, cr
        t0 = time.time()
        Wflat = np.ascontiguousarray(cr.transpose(1, 2, 0)).reshape(-1, order="F")                # [s + d*(b + chi_r*k)]
        RE[r] = bond_ref.env_update(phi[:, r], R, Wflat, cr.shape[0], nthreads)                   # update_caches!
        if sat:
            t_env += time.time() - t0
            work += N * 4.0 * dims[0] * d * d * dims[2]
            done += 1
    rate_on = N * done / (t_loop + t_env)
    return {"rate_loop_only": rate_on, "rate_with_svd": N * done / (t_loop + t_env + t_svd), "n": N, "bonds": done,
            "t_loop": t_loop, "t_env": t_env, "t_svd": t_svd, "t_build": t_build, "gflops_loop": work / t_loop / 1e9}


def chi_saturated_cores(m, w, C=2):
    """Starting state for CPU-only timing: a random MPS already at chi_max (a trained MPS sits there after one sweep)."""
    return m.generate_starting_mps(w["chi_max"], w["T"], w["d"], C, seed=1234)


def run_reference(args):
    """--impl reference: the reference's own algorithm (CPU) on this box's host cores.  The reference is Julia (absent
    here, 250 pinned packages, no network), so this runs the oracle port: the literal per-sample loop of
    Loss_Grad_KLD / update_caches! in C (oracle/bond_ref.c) and LAPACK for the SVD.  The reference's loop is
    single-threaded by construction (loss_functions.jl:322-432, @turbo SIMD only); here it may use every host thread
    (per-thread accumulators), which only flatters the CPU arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import mpstime_jl_b200 as m           # host-side preprocessing + start MPS only
    w = WORKLOADS[args.workload if args.workload in WORKLOADS else "C"]
    ncores = os.cpu_count() or 1
    n_sample, n_bonds = (1024, 4) if w["key"] == "C" else (1024, 8)
    X, y = make_data(n_sample, w["T"], w["seed"], w)
    opts = m.MPSOptions(d=w["d"], chi_max=w["chi_max"], eta=w["eta"])
    Xs, _ = m.transform_train_data(X.T, opts)
    Xs_sorted, _, ys, _, classes, counts = m.sort_by_class(Xs, X, y)
    cores = chi_saturated_cores(m, w)
    rates, times, last = [], [], None
    for it in range(args.warmup + args.steps):
        r = cpu_reference_rate(cores, Xs_sorted, counts, w, n_sample, n_bonds, ncores)
        if it >= args.warmup:
            rates.append(r["rate_loop_only"]); times.append(r["t_loop"] + r["t_env"]); last = r
    v = float(np.mean(rates))
    sample = (f"{last['bonds']} chi-saturated bond updates of a backward half-sweep (after the 2 edge bonds) on {n_sample} samples "
              f"per step; O(N) part only (loss/gradient loop {last['t_loop']:.2f} s + env update {last['t_env']:.2f} s); the "
              f"N-independent LAPACK SVD ({last['t_svd']:.2f} s for these bonds) is amortised over N = {w['N']} in the named config "
              f"and left out, which favours the CPU arm; with it the sample rate is {last['rate_with_svd']:.0f}/s")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "sample-bonds/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)),
        "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["name"], "T": w["T"], "d": w["d"], "chi_max": w["chi_max"], "N_global": w["N"], "sample": sample},
        "cpu_baseline": {"value": v, "unit": "sample-bonds/s", "cores": ncores, "kind": "port", "sample": sample,
                         "loop_gflops": last["gflops_loop"],
                         "note": "restated reference algorithm (numpy + C hot loop, LAPACK gesdd), not Julia; the reference's own loop is single-threaded"},
        "e2e": {"value": v, "unit": "sample-bonds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------
def impute_flops_per_instance(T, K, d, chi, G):
    """per imputed site 2*G*d^2 + 3*G (pdf on the grid) + 2*d*chi + 2*d*chi^2 (projection into the next
    core) + 2*d*chi^3 (right-Gram / orthogonalisation); per known site 2*d*chi^2."""
    return K * (2.0 * G * d * d + 3.0 * G + 2.0 * d * chi + 2.0 * d * chi * chi + 2.0 * d * chi ** 3) + (T - K) * 2.0 * d * chi * chi


def impute_bench(m, ctx, args, rank, world, td, local, fp64_peak):
    """BASELINE.json configs[3]: MPS_impute on synthetic instances, T = 256, d = 16, chi = 64, one contiguous block of
    128 missing sites with a uniform random start (missing_data_mechanisms.jl:146-153), grid dx = 1e-4 (G = 20 001),
    methods median and ITS (1 trajectory, pre-drawn uniforms); `n` instances PER RANK, no communication.  The class MPS
    is a random chi = 64 MPS kept at unit running scale (the cost of imputation does not depend on the values).
    Host buffers in, host buffers out (this is the reference-facing mpst_impute_batch call: an e2e figure)."""
    T, d, chi, K = 256, 16, 64, 128
    n = args.impute_instances
    rng = np.random.default_rng(3 + rank)
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
    U = rng.uniform(0.0, 1.0, size=(n, 1, K))
    ctx.impute_batch(0, X[:, :296], mask[:, :296], grid)                   # warm-up: sizes the work buffers
    res = {}
    flops = impute_flops_per_instance(T, K, d, chi, len(grid))
    for method in ("median", "ITS"):
        ctx.profile_enable(True)
        ctx.profile_reset()
        barrier(td, local)
        t0 = time.time()
        out = ctx.impute_batch(0, X, mask, grid, method=method, uniforms=U if method == "ITS" else None)
        dt = time.time() - t0
        kms = ctx.profile_get()["impute"][0]
        ctx.profile_enable(False)
        assert np.isfinite(out).all()
        dt = reduce_over_ranks(td, local, dt)
        kms = reduce_over_ranks(td, local, kms)
        ach = flops * n / (kms * 1e-3) / 1e12 if kms > 0 else 0.0
        res[method] = {"value": n * world / dt, "unit": "instances/s", "seconds": dt, "kernel_ms": kms,
                       "roofline": {"bound": "tensor", "kernel": "impute_kernel", "achieved": ach, "peak": fp64_peak[0],
                                    "unit": "TFLOP/s", "frac": ach / fp64_peak[0], "traffic": None,
                                    "algorithmic_flops_per_instance": flops, "peak_source": fp64_peak[1],
                                    "note": "one persistent launch per batch; flops per SURVEY 8(d): pdf 2Gd^2+3G, projection, 2d chi^3 Gram per missing site"}}
    out = {"metric": "MPS_impute instances/sec", "value": res["median"]["value"], "unit": "instances/s",
           "instances_per_gpu": n, "n_gpus": world, "T": T, "d": d, "chi": chi, "missing": K, "grid_points": len(grid),
           "median": res["median"], "ITS": res["ITS"],
           "note": "host buffers in/out through mpst_impute_batch; instances sharded over ranks, no communication; "
                   "bounded batch of BASELINE.json configs[3] (100k instances)"}
    if rank == 0 and not args.no_cpu_baseline:
        try:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import mpstime_oracle as o
            genc = o.encode(grid, d)
            cls = o.expand_label_index(cores)[0]
            nb = 4
            t0 = time.time()
            for i in range(nb):
                o.impute_series(cls, X[:, i], list(np.nonzero(mask[:, i])[0]), grid, genc, d, method="median")
            dtc = time.time() - t0
            out["cpu_baseline"] = {"value": nb / dtc, "unit": "instances/s", "cores": 1, "kind": "port", "seconds": dtc,
                                   "sample": f"{nb} instances of the same shape, median, numpy restatement (oracle.impute_series), BLAS threads as numpy chooses"}
        except Exception as e:
            out["cpu_baseline"] = {"value": None, "error": repr(e)}
    return out


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


_T0 = time.time()


def log(rank, msg):
    if rank == 0:
        print("[bench %.1fs] %s" % (time.time() - _T0, msg), file=sys.stderr, flush=True)


def fits_on_gpu(local, w, n_local):
    """Environment ring + series + work buffers of the workload against free device memory."""
    try:
        import torch
        free, total = torch.cuda.mem_get_info(local)
    except Exception:
        return True, 0, 0
    npad = (n_local + 127) // 128 * 128 + 128
    need = 8.0 * npad * (w["T"] * w["chi_max"] + w["T"] + 2 * w["d"] + 3 * w["chi_max"] + 8) + 3e9
    return need < free, need, free


def run_training(m, ctx, w, args, rank, world, local, td, steps, warmup, fp64_peak, do_e2e=True):
    """One workload on the current ranks.  Returns the result dict (rank 0) plus the trained host cores."""
    T, d, chi_max = w["T"], w["d"], w["chi_max"]
    if w["scaling"] == "strong":
        N_global = w["N"]
        N_local = N_global // world
        N_global = N_local * world
    else:
        N_local = w["N"]
        N_global = N_local * world
    # every rank generates only its own shard (same distribution, rank-dependent seed) and normalises it with its own
    # statistics; a real multi-GPU fitMPS call computes them once on the host before sharding (mpstime.jl_b200/api.py)
    X, y = make_data(N_local, T, w["seed"] + 1000 * rank, w)
    opts = m.MPSOptions(d=d, chi_max=chi_max, eta=w["eta"], log_level=0, verbosity=-1, nsweeps=1)
    Xs, _ = m.transform_train_data(X.T, opts)
    Xs_sorted, _, ys, _, classes, counts = m.sort_by_class(Xs, X, y)
    del X, Xs
    counts_global = counts * world
    cores0 = m.generate_starting_mps(w["chi_init"], T, d, 2, seed=1234)
    topts = m.make_opts(chi_max=chi_max, eta=w["eta"])
    nb_sweep = 2 * (T - 1)
    bps = w["bonds_per_step"] or nb_sweep

    # ---- device-resident arm --------------------------------------------------------------------
    ctx.train_load_x(Xs_sorted, counts, d, chi_max, n_global=N_global, counts_global=counts_global)
    ctx.set_cores(cores0)
    ctx.sweep_bonds(topts, nb_sweep, restart=True, record=False)           # one whole untimed sweep: links reach chi_max
    cores_host = ctx.get_cores()                                            # trained, chi-saturated, label on the last site
    log(rank, "warm sweep done")
    for _ in range(warmup):
        ctx.sweep_bonds(topts, bps, record=False)
    ctx.profile_enable(True)
    ctx.profile_reset()
    ctx.debug_set("grad_kr_launches", 0)
    ctx.debug_set("grad_tile_launches", 0)
    SVD_KEYS = ("svd_calls", "svd_iters_sum", "svd_round2", "svd_jacobi", "svd_fast", "svd_serial")
    for k in SVD_KEYS:
        ctx.debug_set(k, 0)
    clocks = ClockSampler(local)
    barrier(td, local)
    if rank == 0:
        clocks.start()
    l0 = ctx.launch_count()
    ctx.timer_start()
    t_wall = time.time()
    chis = []
    for _ in range(steps):
        lo, gn, chi = ctx.sweep_bonds(topts, bps, record=True)
        chis.append(chi)
    ms = ctx.timer_stop()
    wall = time.time() - t_wall
    barrier(td, local)
    clk = clocks.stop() if rank == 0 else None
    launches = ctx.launch_count() - l0
    prof = ctx.profile_get()
    ctx.profile_enable(False)
    n_kr, n_tile = ctx.debug_get("grad_kr_launches"), ctx.debug_get("grad_tile_launches")
    grad_kernel = GRAD_KERNEL_NAMES[1 if n_kr >= n_tile else 2]          # the kernel that ran most of the timed launches
    grad_mix = {"bond_grad_kr_kernel": n_kr, "bond_grad_kernel": n_tile}
    svd_stats = {k: ctx.debug_get(k) for k in SVD_KEYS}
    ms = reduce_over_ranks(td, local, ms)
    sample_bonds = steps * bps * N_global
    value = sample_bonds / (ms * 1e-3)
    log(rank, "timed region done: %.1f ms per bond" % (ms / (steps * bps)))

    # ---- end-to-end arm: one fitMPS-style call from host buffers ---------------------------------------
    e2e = None
    if do_e2e:
        h2d = Xs_sorted.nbytes + sum(c.nbytes for c in cores_host)
        d2h = sum(c.nbytes for c in cores_host) + nb_sweep * 20
        # the input lives in pinned host memory, already in the wire layout (Julia column-major T x N == C-order (N, T))
        X_pinned = pinned_copy(np.ascontiguousarray(Xs_sorted.T)).T
        n_e2e = 1 if w["key"] == "C" else max(1, min(steps, 3))
        if w["key"] != "C":                # cheap workload: one untimed call first (first touch of the pinned buffer)
            ctx.train_load_x(X_pinned, counts, d, chi_max, n_global=N_global, counts_global=counts_global)
            ctx.set_cores(cores_host)
            ctx.sweep(topts, 1, record=True)
        barrier(td, local)
        t0 = time.time()
        ch = cores_host
        for _ in range(n_e2e):
            ctx.train_load_x(X_pinned, counts, d, chi_max, n_global=N_global, counts_global=counts_global)
            ctx.set_cores(ch)
            ctx.sweep(topts, 1, record=True)          # build LE, backward + forward half-sweep, normalize!
            ch = ctx.get_cores()
        barrier(td, local)
        e2e_s = reduce_over_ranks(td, local, time.time() - t0)
        log(rank, "e2e done")
        e2e = {"value": n_e2e * nb_sweep * N_global / e2e_s, "unit": "sample-bonds/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "seconds": e2e_s, "calls_timed": n_e2e,
               "step": f"one whole sweep ({nb_sweep} bonds) per call: mpst_train_load_x (pinned host series) + mpst_set_core x T + "
                       "mpst_sweep (environments, all bonds, per-bond loss read-back, normalize!) + mpst_get_core x T; "
                       "bytes are per call and per rank"}
        _PINNED.clear()
    if rank != 0:
        return None, cores_host, (Xs_sorted, counts)
    gk_ms, gk_n, gk_fl = prof["grad_kernel"]
    achieved = gk_fl / (gk_ms * 1e-3) / 1e12 if gk_ms > 0 else 0.0
    traffic, traffic_src = kernel_traffic(grad_kernel, d, chi_max, N_local)
    roofline = {"bound": "tensor", "kernel": grad_kernel, "kernel_launch_mix": grad_mix, "achieved": achieved,
                "peak": fp64_peak[0], "unit": "TFLOP/s", "frac": achieved / fp64_peak[0], "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": fp64_peak[1], "launches": gk_n,
                "avg_launch_ms": gk_ms / max(gk_n, 1), "algorithmic_flops_per_launch": gk_fl / max(gk_n, 1),
                "algorithmic_bytes_per_launch": 8.0 * N_local * (2 * chi_max + 2 * d + 1)}
    breakdown = {k: {"ms": round(v[0], 3), "n": v[1]} for k, v in prof.items() if v[1]}
    step_desc = (f"{bps} consecutive bond updates of the running sweep (after one full untimed sweep; blocks continue in sweep order)"
                 if w["bonds_per_step"] else f"one whole sweep = {nb_sweep} bond updates")
    out = {
        "metric": METRIC, "value": value, "unit": "sample-bonds/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True,
        "scaling": w["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["name"], "N_per_gpu": N_local, "N_global": N_global, "T": T, "d": d, "chi_max": chi_max,
                   "encoding": "Legendre_No_Norm", "loss": "KLD", "bbopt": "TSGO", "eta": w["eta"], "update_iters": 1,
                   "step": step_desc, "bonds_per_step": bps,
                   "parallelism": f"sample-sharded x{world}, one NCCL all-reduce per bond, SVD replicated" if world > 1 else "single GPU",
                   "l2_policy": "inputs larger than L2: env cache %.1f GB + series %.2f GB per GPU" % (
                       (T * N_local * chi_max * 8) / 1e9, Xs_sorted.nbytes / 1e9),
                   "mean_chi": float(np.mean(np.concatenate(chis)))},
        "clocks": clk, "gpu_launches": int(launches),
        "roofline": roofline, "device_time_breakdown_ms": breakdown, "wall_s_timed": wall,
        "ms_per_bond": ms / (steps * bps),
        "svd_stats": dict(svd_stats, note="timed region: splits, subspace iterations summed over the fast-path splits, splits that "
                                          "needed a second round of iterations, exact-Jacobi fallbacks, fast-path splits, of which in the serial (fully orthonormalising) loop"),
    }
    if e2e is not None:
        out["e2e"] = e2e
    return out, cores_host, (Xs_sorted, counts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "C", "B"],
                    help="C = BASELINE configs[2] north star (default, strong scaling); B = configs[1] (weak scaling)")
    ap.add_argument("--samples", dest="n", type=int, default=0, help="override N (global for C, per GPU for B; debug)")
    ap.add_argument("--series-length", dest="t", type=int, default=0, help="override series length (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-impute", action="store_true")
    ap.add_argument("--no-config-b", action="store_true")
    ap.add_argument("--impute-instances", type=int, default=16384, help="imputation instances per GPU")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank, world, local, td = dist_setup(args.gpus)
    import mpstime_jl_b200 as m
    fp64_peak = measure_fp64_gemm_peak(local)
    key = "C" if args.workload == "auto" else args.workload
    w = dict(WORKLOADS[key])
    note = None
    if args.n:
        w["N"] = args.n
    if args.t:
        w["T"] = args.t
    if args.n or args.t:
        w["name"] += "(debug override N=%d T=%d)" % (w["N"], w["T"])
    if key == "C":
        ok, need, free = fits_on_gpu(local, w, w["N"] // world)
        ok = reduce_over_ranks(td, local, 0.0 if ok else 1.0) == 0.0
        if not ok and args.workload == "auto":
            note = ("configs[2] needs %.0f GB per GPU at %d GPU(s) but %.0f GB are free: fell back to configs[1]" % (need / 1e9, world, free / 1e9))
            w = dict(WORKLOADS["B"])
    ctx = m.Context(local)
    if world > 1:
        m.dist.init_comm(ctx)

    out, cores_host, data = run_training(m, ctx, w, args, rank, world, local, td, args.steps, args.warmup, fp64_peak)
    if rank == 0 and note:
        out["config"]["fallback"] = note

    # CPU baseline on the same workload (rank 0, bounded sample, 1 thread like the reference's own hot loop)
    if rank == 0 and not args.no_cpu_baseline:
        try:
            n_sample, n_bonds = (1024, 8) if w["key"] == "C" else (2048, 12)
            r = cpu_reference_rate(cores_host, data[0], data[1], w, n_sample, n_bonds, 1)
            out["cpu_baseline"] = {
                "value": r["rate_loop_only"], "unit": "sample-bonds/s", "cores": 1, "kind": "port",
                "seconds": r["t_loop"] + r["t_env"] + r["t_svd"], "loop_gflops": r["gflops_loop"],
                "sample": f"{r['bonds']} chi-saturated bond updates of a backward half-sweep on {r['n']} samples, same T/d/chi, the trained cores of this "
                          f"run; 1 thread like the reference's own hot loop (numpy + C port, not Julia); O(N) part only: loop "
                          f"{r['t_loop']:.2f} s + env {r['t_env']:.2f} s; the N-independent LAPACK SVD took {r['t_svd']:.2f} s "
                          f"(with it: {r['rate_with_svd']:.0f} sample-bonds/s on this sample)"}
        except Exception as e:          # the baseline is reporting only
            out["cpu_baseline"] = {"value": None, "error": repr(e)}
    del data

    # configs[1] beside the north star (single GPU only: the headline of round 1, kept for continuity)
    if world == 1 and w["key"] == "C" and not args.no_config_b:
        try:
            ob, _, _ = run_training(m, ctx, dict(WORKLOADS["B"]), args, rank, world, local, td, steps=3, warmup=2,
                                    fp64_peak=fp64_peak, do_e2e=True)
            out["config_B"] = {k: ob[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "config", "roofline", "e2e",
                                                  "device_time_breakdown_ms", "ms_per_bond", "gpu_launches", "svd_stats")}
        except Exception as e:
            out["config_B"] = {"value": None, "error": repr(e)}

    # the reference-facing call itself: fitMPS(X_train, y_train, X_test, y_test, opts) with the reference's default
    # log_level = 3 (per-sweep train / test loss, KL divergence, accuracy, confusion on the device), raw host arrays in,
    # trained MPS out -- three sweeps of configs[1], host preprocessing and every read-back inside the timed region
    if world == 1 and w["key"] == "C" and not args.no_config_b and out.get("config_B", {}).get("value"):
        try:
            wb = WORKLOADS["B"]
            Xb, yb = make_data(wb["N"], wb["T"], 77, wb)
            Xt, yt = make_data(10_000, wb["T"], 78, wb)
            nsw = 3
            ob = m.MPSOptions(d=wb["d"], chi_max=wb["chi_max"], eta=wb["eta"], nsweeps=nsw, log_level=3, verbosity=-1)
            t0 = time.time()
            mps, info, _ = m.fitMPS(Xb, yb, Xt, yt, ob, device=local)
            dt = time.time() - t0
            nb = 2 * (wb["T"] - 1)
            out["config_B"]["api_fitMPS"] = {
                "value": nsw * nb * wb["N"] / dt, "unit": "sample-bonds/s", "seconds": dt, "sweeps": nsw,
                "train_acc": float(info["train_acc"][-1]), "test_acc": float(info["test_acc"][-1]),
                "note": "wall clock of one fitMPS call (N = 100k train + 10k test, T = 100, d = 12, chi_max = 40, 3 sweeps from a random "
                        "chi_init = 4 start, log_level = 3): normalisation + class sort on the host, series upload, start MPS, "
                        "3 x 198 bond updates driven one C-ABI call at a time, 5 evaluations of both sets on the device, "
                        "normalize!, cores read back"}
        except Exception as e:
            out["config_B"]["api_fitMPS"] = {"value": None, "error": repr(e)}

    if not args.no_impute:
        try:
            imp = impute_bench(m, ctx, args, rank, world, td, local, fp64_peak)
            if rank == 0:
                out["impute"] = imp
        except Exception as e:
            if rank == 0:
                out["impute"] = {"value": None, "error": repr(e)}
    if td is not None:
        td.barrier(device_ids=[local])
        td.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
