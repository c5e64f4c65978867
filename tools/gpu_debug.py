"""Development aid: run every device component once against the oracle and print the deviations
(each block in its own try/except so one gpurun call reports as much as possible)."""
import sys, os, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import mpstime_oracle as o
import mpstime_jl_b200 as m

ctx = m.Context(0)
rng = np.random.default_rng(0)

def block(name):
    def deco(fn):
        t = time.time()
        try:
            fn()
            print(f"[ok ] {name}  ({time.time()-t:.2f}s)", flush=True)
        except Exception:
            print(f"[ERR] {name}", flush=True)
            traceback.print_exc()
        return fn
    return deco

@block("encode")
def _():
    x = rng.uniform(-1, 1, 1000)
    for b, d in (("legendre_no_norm", 12), ("legendre_norm", 7), ("fourier", 6), ("sahand", 6), ("uniform", 4)):
        xx = x if b not in ("sahand",) else (x + 1) / 2
        dev = ctx.encode(xx, d, b); ref = o.encode(xx, d, b)
        print("   ", b, d, np.abs(dev - ref).max())
    xx = (x + 1) / 2
    print("    stoudenmire", np.abs(ctx.encode(xx, 2, "stoudenmire") - o.encode(xx, 2, "stoudenmire")).max())

def rand_bond(N, d, chi_l, chi_r, C, counts=None):
    if counts is None:
        counts = [N // C] * C; counts[-1] += N - sum(counts)
    xl = o.legendre_encode(rng.uniform(-1, 1, N), d); xr = o.legendre_encode(rng.uniform(-1, 1, N), d)
    L = rng.standard_normal((N, chi_l)) / np.sqrt(chi_l) if chi_l > 1 else np.ones((N, 1))
    R = rng.standard_normal((N, chi_r)) / np.sqrt(chi_r) if chi_r > 1 else np.ones((N, 1))
    B = rng.standard_normal((d * chi_l * d * chi_r, C)); B /= np.linalg.norm(B)
    return B, L, R, xl, xr, np.array(counts)

@block("bond_loss_grad")
def _():
    for (N, d, cl, cr, C) in ((300, 4, 3, 5, 2), (1000, 12, 40, 40, 2), (257, 5, 1, 7, 3), (100, 6, 9, 1, 1), (2000, 16, 64, 64, 2)):
        B, L, R, xl, xr, counts = rand_bond(N, d, cl, cr, C)
        for loss in ("KLD", "MSE"):
            t = time.time()
            lo, G, yh = ctx.bond_loss_grad(B, L, R, xl, xr, counts, loss=loss, want_yhat=True)
            dt = time.time() - t
            fn = o.loss_grad_KLD if loss == "KLD" else o.loss_grad_MSE
            lo_r, G_r = fn(B, L, R, xl, xr, counts)
            yr = o.bond_yhat(B, L, R, xl, xr)
            if loss == "KLD":
                own = np.repeat(np.arange(C), counts)
                ey = np.abs(yh[np.arange(N), own] - yr[np.arange(N), own]).max()
            else:
                ey = np.abs(yh - yr).max()
            print(f"    N={N} d={d} chi=({cl},{cr}) C={C} {loss}: loss rel {abs(lo-lo_r)/abs(lo_r):.2e} "
                  f"gnorm rel {abs(np.linalg.norm(G)-np.linalg.norm(G_r))/np.linalg.norm(G_r):.2e} "
                  f"G maxabs {np.abs(G-G_r).max()/np.abs(G_r).max():.2e} yhat {ey:.2e}  ({dt:.2f}s)")

@block("bond_split")
def _():
    for (d, cl, cr, C, gl, chimax) in ((4, 3, 5, 2, True, 8), (4, 3, 5, 2, False, 8), (12, 40, 40, 2, True, 40), (12, 40, 40, 2, False, 40),
                                      (5, 1, 6, 2, True, 10), (5, 6, 1, 2, False, 10), (16, 64, 64, 2, True, 64)):
        B = rng.standard_normal((d * cl * d * cr, C))
        # low-rank + noise like a trained bond
        B /= np.linalg.norm(B)
        t = time.time()
        c_l, c_r, sig = ctx.bond_split(B, d, cl, cr, gl, chimax)
        dt = time.time() - t
        r_l, r_r, rs = o.decompose_bt(B, (cl, d, cr), gl, chimax, 1e-10)
        if gl:
            prod = np.einsum("asmc,mtb->btasc", c_l, c_r); prod_r = np.einsum("asmc,mtb->btasc", r_l, r_r)
        else:
            prod = np.einsum("asm,mtbc->btasc", c_l, c_r); prod_r = np.einsum("asm,mtbc->btasc", r_l, r_r)
        orth = c_r.reshape(c_r.shape[0], -1) if gl else c_l.reshape(-1, c_l.shape[2]).T
        print(f"    d={d} chi=({cl},{cr}) C={C} left={gl}: chi {len(sig)} vs {len(rs)}  sigma rel {np.abs(sig-rs).max()/rs.max():.2e} "
              f"product {np.abs(prod-prod_r).max():.2e} orth {np.abs(orth@orth.T-np.eye(orth.shape[0])).max():.2e} ({dt:.2f}s)")

@block("sweep vs oracle")
def _():
    N, T, d, C = 200, 10, 4, 2
    X, y = o.synthetic_two_class(N, T, seed=1)
    Xs, _ = o.transform_train_data(X.T)
    phi, ys, order, counts, classes = o.encode_dataset(Xs, y, d)
    cores = o.random_start_mps(T, d, 4, C, seed=3)
    rec = []
    new = o.fit_sweeps(cores, phi, counts, nsweeps=2, chi_max=12, eta=0.05, record=rec)
    ctx.train_load_x(Xs[:, order], counts, d, 12)
    ctx.set_cores(cores)
    back = ctx.get_cores()
    print("    core roundtrip", max(np.abs(a - b).max() for a, b in zip(cores, back)))
    opts = m.make_opts(chi_max=12, eta=0.05)
    lo, gn, chi = ctx.sweep(opts, 2)
    rl = np.array([r["loss"] for r in rec]); rg = np.array([r["gradnorm"] for r in rec]); rc = np.array([r["chi"] for r in rec])
    print("    loss rel", np.abs(lo - rl).max() / np.abs(rl).max(), " gnorm rel", (np.abs(gn - rg) / rg).max(), " chi equal", np.array_equal(chi, rc))
    print("    first bonds dev", list(zip(np.round(lo[:4], 6), chi[:4])), "oracle", list(zip(np.round(rl[:4], 6), rc[:4])))
    dev_cores = ctx.get_cores()
    ov_d = o.overlaps(dev_cores, phi); ov_r = o.overlaps(new, phi)
    print("    final overlaps maxabs", np.abs(np.abs(ov_d) - np.abs(ov_r)).max(), "norm", o._norm2_general(dev_cores))
    yh, am = ctx.overlaps(X_TxN=Xs[:, order])
    print("    device overlaps vs oracle-on-device-cores", np.abs(yh - ov_d).max(), "pred equal", np.array_equal(am, np.argmax(ov_r**2, 1)))

@block("sweep MSE / GD / iters=2")
def _():
    N, T, d, C = 150, 6, 3, 3
    rs = np.random.default_rng(5)
    X = rs.standard_normal((N, T)).cumsum(1); y = rs.integers(0, C, N)
    Xs, _ = o.transform_train_data(X.T)
    phi, ys, order, counts, classes = o.encode_dataset(Xs, y, d)
    cores = o.random_start_mps(T, d, 3, C, seed=4)
    for kw in (dict(loss="MSE"), dict(bbopt="GD", eta=0.02), dict(update_iters=2), dict(train_sep=True), dict(rescale=(True, False))):
        rec = []
        okw = dict(nsweeps=1, chi_max=8, eta=0.05); okw.update(kw)
        new = o.fit_sweeps(cores, phi, counts, record=rec, **okw)
        ctx.train_load_x(Xs[:, order], counts, d, 8); ctx.set_cores(cores)
        mk = dict(chi_max=8, eta=okw["eta"]); mk.update({k: v for k, v in kw.items() if k != "eta"})
        lo, gn, chi = ctx.sweep(m.make_opts(**mk), 1)
        rl = np.array([r["loss"] for r in rec]); rc = np.array([r["chi"] for r in rec])
        print("   ", kw, "loss rel", np.abs(lo - rl).max() / np.abs(rl).max(), "chi equal", np.array_equal(chi, rc))

@block("timing config B bond")
def _():
    N, d, chi = 100000, 12, 40
    B, L, R, xl, xr, counts = rand_bond(N, d, chi, chi, 2)
    ctx.profile_enable(True)
    for rep in range(2):
        ctx.profile_reset()
        lo, G = ctx.bond_loss_grad(B, L, R, xl, xr, counts)
        print("   ", {k: v for k, v in ctx.profile_get().items() if v[1]})
    D = (d * chi) ** 2
    print("    grad GEMM flops", 2 * N * D / 1e9, "GF")
    t = time.time(); c_l, c_r, sig = ctx.bond_split(B, d, chi, chi, True, chi); print("    split 960x480 wall", time.time() - t)
    ctx.profile_enable(False)
