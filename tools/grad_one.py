"""Development aid: one config-B sized bond gradient (N samples, d=12, chi=40, 2 classes) through the C ABI."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mpstime_jl_b200 as m
ctx = m.Context(0)
rng = np.random.default_rng(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
d, chi, C = 12, 40, 2
B = rng.standard_normal((d * chi * d * chi, C)); B /= np.linalg.norm(B)
L = rng.standard_normal((N, chi)); R = rng.standard_normal((N, chi))
xl = rng.standard_normal((N, d)); xr = rng.standard_normal((N, d))
counts = np.array([N // 2, N - N // 2])
ctx.profile_enable(True)
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    t = time.time(); lo, G = ctx.bond_loss_grad(B, L, R, xl, xr, counts); print("wall ms", 1e3 * (time.time() - t), lo)
ms, n, fl = ctx.profile_get()["grad_kernel"]
print("grad_kernel avg ms %.4f  TFLOP/s %.2f" % (ms / n, fl / (ms * 1e-3) / 1e12))
