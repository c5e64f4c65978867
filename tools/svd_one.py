import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mpstime_jl_b200 as m
ctx = m.Context(0)
rng = np.random.default_rng(0)
d, chi, C = 12, 40, 2
a = rng.standard_normal((chi, d, chi, C)); b = rng.standard_normal((chi, d, chi))
B = np.einsum("asmc,mtb->btasc", a, b).reshape(-1, C); B /= np.linalg.norm(B)
G = rng.standard_normal(B.shape) * np.exp(-rng.uniform(0, 25, size=(B.shape[0], 1)))
B = B - 0.01 * G / np.linalg.norm(G); B /= np.linalg.norm(B)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    t = time.time(); ctx.bond_split(B, d, chi, chi, True, chi); print("wall ms", 1e3 * (time.time() - t))
