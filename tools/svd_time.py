import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mpstime_jl_b200 as m
ctx = m.Context(0)
rng = np.random.default_rng(0)
d, chi = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (12, 40)
B = rng.standard_normal((d * chi * d * chi, 2)); B /= np.linalg.norm(B)
ts = []
for rep in range(4):
    t = time.time()
    try:
        ctx.bond_split(B, d, chi, chi, True, chi)
    except Exception as e:
        pass
    ts.append(time.time() - t)
print(os.environ.get("MPST_SVD_SKIP", "0"), os.environ.get("MPST_SVD_FIXED"), "min wall ms", 1e3 * min(ts))
