import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import mpstime_oracle as o
import mpstime_jl_b200 as m
ctx = m.Context(0)
N, T, d, C = 200, 10, 4, 2
X, y = o.synthetic_two_class(N, T, seed=1)
Xs, _ = o.transform_train_data(X.T)
phi, ys, order, counts, classes = o.encode_dataset(Xs, y, d)
cores = o.random_start_mps(T, d, 4, C, seed=3)
rec = []
new = o.fit_sweeps(cores, phi, counts, nsweeps=2, chi_max=12, eta=0.05, record=rec)
ctx.train_load_x(Xs[:, order], counts, d, 12); ctx.set_cores(cores)
lo, gn, chi = ctx.sweep(m.make_opts(chi_max=12, eta=0.05), 2)
rl = np.array([r["loss"] for r in rec]); rg = np.array([r["gradnorm"] for r in rec])
print("free-running per-bond rel dev of loss:")
print(np.array2string(np.abs(lo - rl) / np.abs(rl), precision=1, max_line_width=200))
print("gnorm:"); print(np.array2string(np.abs(gn - rg) / np.abs(rg), precision=1, max_line_width=200))
# teacher-forced: replay oracle states bond by bond
print("teacher-forced:")
states = []
def run_tf():
    cs = [c.copy() for c in cores]
    worst = 0
    nb = 0
    for it in range(2):
        for going_left in (True, False):
            js = range(T - 2, -1, -1) if going_left else range(T - 1)
            for j in js:
                # oracle state before this bond = cs; device: load cores + rebuild envs
                ctx.train_load_x(Xs[:, order], counts, d, 12); ctx.set_cores(cs)
                # build the environments the bond needs with the oracle (then compare via device bond_step)
                # device: build LE via build_env only valid when label on last site; instead use oracle envs through bond_loss_grad
                LE = {}; prev = np.ones((N, 1))
                for k in range(j):
                    prev = o.env_step_left(phi[:, k], prev, cs[k]); LE[k] = prev
                nxt = np.ones((N, 1)); RE = {}
                for k in range(T - 1, j + 1, -1):
                    nxt = o.env_step_right(phi[:, k], nxt, cs[k]); RE[k] = nxt
                L = LE[j - 1] if j > 0 else np.ones((N, 1)); R = RE[j + 2] if j + 1 < T - 1 else np.ones((N, 1))
                B, dims = o.flatten_bt(cs[j], cs[j + 1])
                lo_r, G_r = o.loss_grad_KLD(B, L, R, phi[:, j], phi[:, j + 1], counts)
                lo_d, G_d = ctx.bond_loss_grad(B, L, R, phi[:, j], phi[:, j + 1], counts)
                Bn, _, _ = o.apply_update(B, L, R, phi[:, j], phi[:, j + 1], counts, eta=0.05)
                cl, cr, S = o.decompose_bt(Bn, dims, going_left, 12, 1e-10)
                dl, dr, Sd = ctx.bond_split(Bn, d, dims[0], dims[2], going_left, 12)
                e1 = abs(lo_d - lo_r) / abs(lo_r); e2 = np.abs(G_d - G_r).max() / np.abs(G_r).max()
                e3 = (np.abs(Sd - S).max() / S.max()) if len(Sd) == len(S) else 9.9
                if going_left:
                    pd_ = np.einsum("asmc,mtb->btasc", dl, dr); pr_ = np.einsum("asmc,mtb->btasc", cl, cr)
                else:
                    pd_ = np.einsum("asm,mtbc->btasc", dl, dr); pr_ = np.einsum("asm,mtbc->btasc", cl, cr)
                e4 = np.abs(pd_ - pr_).max()
                gap = (S[-1] - (np.linalg.svd(Bn.T.reshape(C, dims[2], d, dims[0], d).transpose(3,0,4,2,1).reshape(dims[0]*C*d, d*dims[2]) if going_left else Bn.T.reshape(C, dims[2], d, dims[0], d).transpose(1,0,2,4,3).reshape(dims[2]*C*d, d*dims[0]), compute_uv=False)[len(S):len(S)+1].sum())) / S[0]
                print(f"  it{it} {'L' if going_left else 'R'} j={j} dims={dims} loss {e1:.1e} G {e2:.1e} sigma {e3:.1e} prod {e4:.1e} chi {len(Sd)}/{len(S)} gap(s_k-s_k+1)/s1 {gap:.2e}")
                cs[j], cs[j + 1] = cl, cr
run_tf()
