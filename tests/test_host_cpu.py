"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, the host
preprocessing mirrors the oracle, option validation mirrors the reference's errors, and the
sample-sharding plan reproduces the full gradient (world_size 2, gloo)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(pkg):
    header = open(os.path.join(ROOT, "include", "mpstime_b200.h")).read()
    declared = set(re.findall(r"\b(mpst_[a-z0-9_]+)\s*\(", header))
    declared -= {"mpst_ctx", "mpst_train_opts"}
    assert len(declared) >= 20
    assert declared == set(pkg.SIGNATURES), declared ^ set(pkg.SIGNATURES)
    from mpstime_jl_b200 import _lib
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.mpst_version() >= 100


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.MPSTError):
        pkg.Context(0)


def test_product_never_imports_oracle():
    pkgdir = os.path.join(ROOT, "mpstime.jl_b200")
    for dirpath, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "mpstime_oracle" not in src and "oracle/" not in src, f


def test_preprocess_matches_oracle(pkg, oracle):
    rng = np.random.default_rng(0)
    Xtr = rng.standard_normal((50, 20)).cumsum(axis=1)
    Xte = 1.3 * rng.standard_normal((9, 20)).cumsum(axis=1)
    opts = pkg.MPSOptions()
    Xs, norms = pkg.transform_train_data(Xtr.T, opts)
    Xo, no = oracle.transform_train_data(Xtr.T)
    assert np.array_equal(Xs, Xo)
    Xt, oob = pkg.transform_test_data(Xte.T, norms, opts)
    Xto, oobo = oracle.transform_test_data(Xte.T, no)
    assert np.array_equal(Xt, Xto) and oob == oobo
    back = pkg.invert_test_transform(Xt, oob, norms, opts)
    backo = oracle.invert_test_transform(Xto, oobo, no)
    assert np.array_equal(np.isnan(back), np.isnan(backo)) and np.allclose(back, backo, equal_nan=True)
    opts2 = pkg.MPSOptions(sigmoid_transform=False, data_bounds=(0.1, 0.9))
    Xs2, n2 = pkg.transform_train_data(Xtr.T, opts2)
    Xo2, _ = oracle.transform_train_data(Xtr.T, sigmoid_transform=False, data_bounds=(0.1, 0.9))
    assert np.array_equal(Xs2, Xo2) and abs(Xs2.min() + 0.8) < 1e-12 and abs(Xs2.max() - 0.8) < 1e-12
    # golden: the reference's own saved data (see test_oracle.py)
    g = np.load(os.path.join(ROOT, "tests", "golden", "ecg200_legendre.npz"))
    Xg, _ = pkg.transform_train_data(g["X_orig"].T, opts)
    assert np.abs(np.sqrt(1.5) * Xg.T - g["phi_ref"][:, :, 1]).max() < 1e-14


def test_sort_and_start_mps(pkg, oracle):
    y = np.array([1, 0, 1, 0, 2, 0])
    X = np.arange(12.0).reshape(2, 6)
    Xs, _, ys, order, classes, counts = pkg.sort_by_class(X, None, y)
    assert list(order) == [1, 3, 5, 0, 2, 4] and list(counts) == [3, 2, 1] and list(classes) == [0, 1, 2]
    cores = pkg.generate_starting_mps(4, 7, 3, 2, seed=1)
    assert cores[-1].shape == (3, 3, 1, 2) and cores[0].shape == (1, 3, 3)
    assert abs(oracle._norm2_general(cores) - 1.0) < 1e-12
    for A in cores[:-1]:                                   # left-orthonormal
        a, s, b = A.shape
        M = A.reshape(a * s, b)
        assert np.abs(M.T @ M - np.eye(b)).max() < 1e-12


def test_option_validation_mirrors_reference(pkg):
    with pytest.raises(ValueError, match="Optim"):
        pkg.MPSOptions(bbopt="Optim")._check()
    with pytest.raises(ValueError):
        pkg.MPSOptions(encoding="Fourier")._check()
    assert pkg.MPSOptions()._check() == "legendre_no_norm"
    o = pkg.MPSOptions()
    assert (o.nsweeps, o.chi_max, o.eta, o.d, o.cutoff, o.chi_init, o.init_rng) == (10, 25, 0.01, 5, 1e-10, 4, 1234)
    clf = pkg.MPSClassifier()
    assert (clf.nsweeps, clf.chi_max, clf.d, clf.exit_early) == (5, 15, 2, True)


def test_shard_ranges(pkg):
    counts = [7, 5, 11]
    for world in (1, 2, 3, 8):
        seen = []
        for r in range(world):
            rg = pkg.dist.shard_ranges(counts, r, world)
            assert len(rg) == 3
            seen.append(rg)
        for c in range(3):
            off = sum(counts[:c])
            assert seen[0][c][0] == off and seen[-1][c][1] == off + counts[c]
            for r in range(world - 1):
                assert seen[r][c][1] == seen[r + 1][c][0]


_GLOO_SCRIPT = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "oracle"))
import numpy as np, torch, torch.distributed as td
import mpstime_oracle as o
import mpstime_jl_b200 as m
td.init_process_group("gloo")
rank, world = td.get_rank(), td.get_world_size()
assert m.dist.rank_world() == (rank, world)
rng = np.random.default_rng(0)
N, d, cl, cr, C = 41, 3, 4, 5, 2
counts = np.array([17, 24])
xl = o.legendre_encode(rng.uniform(-1, 1, N), d); xr = o.legendre_encode(rng.uniform(-1, 1, N), d)
L = rng.standard_normal((N, cl)); R = rng.standard_normal((N, cr))
B = rng.standard_normal((d * cl * d * cr, C))
for loss in ("KLD", "MSE"):
    fn = o.loss_grad_KLD if loss == "KLD" else o.loss_grad_MSE
    lo_full, G_full = fn(B, L, R, xl, xr, counts)
    X = np.arange(N, dtype=np.float64)[None, :]
    _, cloc, idx = m.dist.shard_samples(X, counts, rank, world)
    lo_loc, G_loc = fn(B, L[idx], R[idx], xl[idx], xr[idx], cloc)
    # local normaliser is the LOCAL N; the device kernel uses the global one: rescale, then sum
    buf = torch.from_numpy(np.concatenate([G_loc.reshape(-1) * len(idx) / N, [lo_loc * len(idx) / N]]))
    td.all_reduce(buf)
    G = buf[:-1].numpy().reshape(G_full.shape); lo = float(buf[-1])
    assert abs(lo - lo_full) < 1e-12 * abs(lo_full), (lo, lo_full)
    assert np.abs(G - G_full).max() < 1e-12 * np.abs(G_full).max()
# imputation / classification: contiguous instance blocks per rank, results concatenated in rank order on every rank
n_inst = 7
b, e = m.dist.shard_instances(n_inst, rank, world)
full = np.arange(n_inst * 3, dtype=np.float64).reshape(n_inst, 1, 3)
got = m.dist.gather_instances((full[b:e], 2.0 * full[b:e]))
assert np.array_equal(got[0], full) and np.array_equal(got[1], 2.0 * full)
assert np.array_equal(m.dist.gather_instances(full[b:e, 0]), full[:, 0])
td.barrier()
if rank == 0: print("GLOO_OK")
'''


def test_sharded_gradient_allreduce_gloo(tmp_path):
    """N>1 host logic on CPU: per-class sharding + one all-reduce of [grad, loss] == full result."""
    script = tmp_path / "gloo_case.py"
    script.write_text(_GLOO_SCRIPT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731", str(script), ROOT],
                         capture_output=True, text=True, env=env, timeout=240)
    assert "GLOO_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU port of the reference loop on the host cores) runs without a GPU and prints
    one JSON line with the keys the driver reads."""
    import json, subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "fitMPS sample-bonds/sec per sweep"
    assert line["unit"] == "sample-bonds/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"].startswith("north_star_trendy_sine_N1M_T256_d16_chi64")
    assert line["scaling"] == "strong"


def _julia_ccalls(src):
    """[(symbol, return type, [argument types])] of every `ccall(sym(:name), Ret, (T1, T2, ...), ...)` in the shim."""
    out = []
    for m in re.finditer(r"ccall\(sym\(:(\w+)\),\s*(\w+),\s*\(", src):
        i = m.end()
        depth, j = 1, i
        while depth:
            depth += {"(": 1, ")": -1}.get(src[j], 0)
            j += 1
        body = src[i:j - 1]
        args, cur, br = [], "", 0
        for ch in body:
            if ch == "{":
                br += 1
            elif ch == "}":
                br -= 1
            if ch == "," and br == 0:
                args.append(cur.strip()); cur = ""
            else:
                cur += ch
        if cur.strip():
            args.append(cur.strip())
        out.append((m.group(1), m.group(2), args))
    return out


def test_julia_shim_signatures_match_the_abi(pkg):
    """julia/B200Backend.jl cannot run here (no Julia in the image); its ccall signatures and struct mirrors are checked
    against the ctypes binding the GPU parity tests drive, and it must bind every symbol the header declares."""
    import ctypes as C
    from mpstime_jl_b200 import _lib
    src = open(os.path.join(ROOT, "julia", "B200Backend.jl")).read()
    jl2c = {
        "Cint": C.c_int, "Int64": C.c_int64, "Float64": C.c_double, "Cstring": C.c_char_p, "Ptr{Cvoid}": C.c_void_p,
        "Ref{Ptr{Cvoid}}": C.POINTER(C.c_void_p), "Ptr{Float64}": _lib.c_double_p, "Ref{Float64}": _lib.c_double_p,
        "Ptr{Int64}": _lib.c_i64_p, "Ptr{Int32}": _lib.c_i32_p, "Ref{Int32}": _lib.c_i32_p, "Ptr{UInt8}": _lib.c_u8_p,
        "Ref{TrainOpts}": C.POINTER(_lib.TrainOpts), "Ref{ImputeOpts}": C.POINTER(_lib.ImputeOpts),
    }
    calls = _julia_ccalls(src)
    assert len(calls) >= len(pkg.SIGNATURES)
    seen = set()
    for name, ret, args in calls:
        assert name in pkg.SIGNATURES, f"shim calls unknown symbol {name}"
        res, argtypes = pkg.SIGNATURES[name]
        assert jl2c[ret] is res, (name, ret, res)
        assert len(args) == len(argtypes), (name, args, argtypes)
        for k, (a, t) in enumerate(zip(args, argtypes)):
            assert jl2c[a] is t, f"{name}: argument {k} is {a} in the shim, {t} in the binding"
        seen.add(name)
    assert seen == set(pkg.SIGNATURES), set(pkg.SIGNATURES) - seen
    # struct mirrors: same field names, order and widths as the ctypes structures (= the C structs)
    width = {"Int32": C.c_int32, "Float64": C.c_double}
    for jl_name, cstruct in (("TrainOpts", _lib.TrainOpts), ("ImputeOpts", _lib.ImputeOpts)):
        body = re.search(r"struct %s\b[^\n]*\n(.*?)\nend" % jl_name, src, re.S).group(1)
        fields = re.findall(r"(\w+)::(\w+)", body)
        assert [(n, width[t]) for n, t in fields] == [(n, t) for n, t in cstruct._fields_], jl_name
    # the seams of SURVEY 8(b) are all defined
    for fn in ("function fitMPS(W::MPS, training_states_meta::EncodedTimeSeriesSet", "function classify(mps::TrainedMPS",
               "function get_predictions(imp::ImputationProblem", "function get_predictions_batch(", "function eval_loss(::ImputationLoss",
               "function mlj_fit(m::MPSClassifier"):
        assert fn in src, fn


def test_options_refuse_what_the_device_path_does_not_implement(pkg):
    """options that change the reference's results must raise, never be silently ignored; the encoding name
    is normalised once (':Legendre', 'Legendre_No_Norm', ...)."""
    for kw in (dict(projected_basis=True, encoding="Fourier"), dict(projected_basis=True, encoding="Uniform"),
               dict(encode_classes_separately=True), dict(dtype=np.complex128), dict(encoding="hist_split_fourier", d=4),
               dict(encoding="hist_split_uniform", d=5, aux_basis_dim=2),
               dict(dtype=np.float32), dict(svd_alg="nonsense"), dict(use_legacy_ITensor=True), dict(loss_grad="Mixed")):
        with pytest.raises(ValueError):
            pkg.MPSOptions(**kw)._check()
    for enc in (":Legendre", "Legendre_No_Norm", "legendre_norm", ":Uniform"):
        o = pkg.MPSOptions(encoding=enc)
        name = o._check()
        assert pkg.preprocess.encoding_range(enc) == pkg.preprocess.encoding_range(name)
        Xs, _ = pkg.transform_train_data(np.random.default_rng(0).standard_normal((7, 9)), o)
        a, b = pkg.preprocess.encoding_range(name)
        assert Xs.min() >= a - 1e-12 and Xs.max() <= b + 1e-12


def test_streamk_schedule_builder(tmp_path):
    """The gradient kernels' host-side stream-K schedule (csrc/streamk.h): every (unit, chunk) covered exactly once,
    CTAs balanced to one chunk, slots contiguous per unit in ascending chunk order, for every walk order."""
    exe = str(tmp_path / "streamk_check")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I/usr/local/cuda/include", "-x", "c++",
                           os.path.join(ROOT, "tests", "cpp", "streamk_check.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "STREAMK_OK" in out.stdout, out.stdout + out.stderr


def test_normalize_matches_direct_contraction(pkg):
    """api._normalize (normalize!(W), RealRealHighDimension.jl:852) through BLAS contractions with a running rescale
    equals the direct three-operand contraction, label core anywhere in the chain, and yields <W|W> = 1."""
    api = sys.modules[pkg.MPSOptions.__module__]
    rng = np.random.default_rng(0)
    T, chi, d, C = 9, 6, 3, 2
    for pos in (0, 4, T - 1):
        cores = []
        for j in range(T):
            cl, cr = (1 if j == 0 else chi), (1 if j == T - 1 else chi)
            cores.append(rng.standard_normal((cl, d, cr, C)) if j == pos else rng.standard_normal((cl, d, cr)))
        out = api._normalize(cores)
        E = np.ones((1, 1, 1))
        for A in out:
            E = np.einsum("ab,asmc,bsnc->mnc", E[:, :, 0], A, A) if A.ndim == 4 else np.einsum("abc,asm,bsn->mnc", E, A, A)
        assert abs(float(E[0, 0, :].sum()) - 1.0) < 1e-12
        ratio = out[0].ravel()[0] / cores[0].ravel()[0]
        assert all(np.allclose(o, c * ratio, rtol=1e-13, atol=0) for o, c in zip(out, cores))


@pytest.mark.parametrize("knee,r1,r2,rank,expect", [(0, 0.95, 0.95, None, 128), (0, 0.9, 0.9, None, None), (30, 0.6, 0.99, None, None),
                                                    (0, 0.9, 0.9, 64, 64), (0, 0.9, 0.9, 70, 70)])
def test_two_pass_split_model_equals_lapack(knee, r1, r2, rank, expect):
    """The algorithm behind chi_max > 80 on the device (numpy model, tests/svd_two_pass_model.py): kept dimension as
    LAPACK + truncate!, sigma / product to rounding, V orthonormal; pass 2's residuals need sigma_1^2 of the whole
    matrix as their yardstick.  The CUDA implementation is compared with the oracle in tests/test_gpu_zz_wide_links.py."""
    import svd_two_pass_model as tp
    rng = np.random.default_rng(7)
    m, n = 768, 384
    r = rank or n
    U, _ = np.linalg.qr(rng.standard_normal((m, r)))
    V, _ = np.linalg.qr(rng.standard_normal((n, r)))
    k = np.arange(r)
    s = np.where(k < knee, r1 ** np.minimum(k, knee), r1 ** knee * r2 ** np.maximum(k - knee, 0))
    M = (U * s) @ V.T
    M /= np.linalg.norm(M)
    kl, sl, prod = tp.lapack_truncated(M, 128, 1e-10)
    out = tp.two_pass_split(M, 128, 1e-10, rng)
    assert out is not None
    c, US, Vk, P = out
    assert c == kl and (expect is None or c == expect)
    assert np.abs(np.sqrt(P) - sl).max() < 1e-12 and np.abs(US @ Vk.T - prod).max() < 1e-12
    assert np.abs(Vk.T @ Vk - np.eye(c)).max() < 1e-9


def test_out_of_bounds_rescale_equals_the_per_series_loop(pkg):
    """transform_test_data (utils.jl:202-278): the vectorised out-of-bounds rescale is the reference's per-series loop
    (shift up when the minimum is below 0, then divide when the maximum is above 1), bit for bit, and
    invert_test_transform undoes it."""
    rng = np.random.default_rng(0)
    opts = pkg.MPSOptions(d=4, chi_max=8)
    Xtr = rng.standard_normal((200, 30)).cumsum(1)
    _, norms = pkg.transform_train_data(Xtr.T, opts)
    Xte = rng.standard_normal((500, 30)).cumsum(1) * 1.4 + 0.3
    got, oob = pkg.transform_test_data(Xte.T, norms, opts)
    ref = norms.apply(Xte.T)
    want_oob = []
    for i in range(ref.shape[1]):                     # the literal loop
        col = ref[:, i]
        lb_s, ub_s = 0.0, 1.0
        lo, hi = col.min(), col.max()
        if lo < 0:
            col -= lo
            hi = col.max()
            lb_s = float(lo)
        if hi > 1:
            col /= hi
            ub_s = float(hi)
        if (lb_s, ub_s) != (0.0, 1.0):
            want_oob.append((i, lb_s, ub_s))
    a, b = pkg.api.encoding_range(opts.encoding)
    assert np.array_equal(got, (b - a) * ref + a) and oob == want_oob and len(oob) > 10
    assert got.min() >= a and got.max() <= b
    back = pkg.invert_test_transform(got, oob, norms, opts)
    assert np.abs(back - Xte.T).max() < 1e-6 * np.abs(Xte).max()
    one, oob1 = pkg.transform_test_data(Xte[7], norms, opts)
    assert np.array_equal(one, got[:, 7])
