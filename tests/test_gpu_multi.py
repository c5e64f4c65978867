"""Two-rank run of the sharded sweep (NCCL all-reduce of the per-bond gradient inside the library): every rank
must end with bit-identical cores, equal to the single-GPU result on the concatenated data to rounding.
Needs 2 GPUs; skipped otherwise."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_SCRIPT = r'''
import os, sys, hashlib
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "oracle"))
import numpy as np, torch, torch.distributed as td
import mpstime_oracle as o
import mpstime_jl_b200 as m
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
td.init_process_group("nccl", device_id=torch.device("cuda", local))
N, T, d, C = 400, 8, 4, 2
X, y = o.synthetic_two_class(N, T, seed=1)
Xs, _ = o.transform_train_data(X.T)
phi, ys, order, counts, classes = o.encode_dataset(Xs, y, d)
Xs = Xs[:, order]
cores = o.random_start_mps(T, d, 4, C, seed=3)
ctx = m.Context(local)
m.dist.init_comm(ctx)
Xl, cl, idx = m.dist.shard_samples(Xs, counts, rank, world)
ctx.train_load_x(Xl, cl, d, 10, n_global=N, counts_global=counts)
ctx.set_cores(cores)
opts = m.make_opts(chi_max=10, eta=0.3, loss="MSE", bbopt="GD")
lo, gn, chi = ctx.sweep(opts, 1)
dev = ctx.get_cores()
h = hashlib.sha256(b"".join(np.ascontiguousarray(c).tobytes() for c in dev)).hexdigest()
hs = [None] * world
td.all_gather_object(hs, h)
assert len(set(hs)) == 1, hs                       # bit-identical cores on every rank, no broadcast
rec = []
new = o.fit_sweeps(cores, phi, counts, nsweeps=1, chi_max=10, eta=0.3, loss="MSE", bbopt="GD", record=rec)
rl = np.array([r["loss"] for r in rec])
assert np.array_equal(chi, [r["chi"] for r in rec])
assert np.abs(lo - rl).max() < 1e-8 * np.abs(rl).max(), np.abs(lo - rl).max()
# KLD: first bonds strictly
ctx.train_load_x(Xl, cl, d, 10, n_global=N, counts_global=counts); ctx.set_cores(cores)
lo2, gn2, chi2 = ctx.sweep(m.make_opts(chi_max=10, eta=0.05), 1)
rec2 = []
o.fit_sweeps(cores, phi, counts, nsweeps=1, chi_max=10, eta=0.05, record=rec2, max_bonds=4)
for k in range(4):
    assert abs(lo2[k] - rec2[k]["loss"]) < 1e-9 * abs(rec2[k]["loss"]) and abs(gn2[k] - rec2[k]["gradnorm"]) < 1e-8 * rec2[k]["gradnorm"]
# public API under torch.distributed: fitMPS shards the samples, classify / imputation shard by instance (no data-path
# collective); every rank ends with the same labels / imputed series as its own unsharded call
Xr, yr = o.synthetic_two_class(160, 10, seed=5)
mo = m.MPSOptions(d=4, chi_max=8, nsweeps=2, eta=0.05, verbosity=-1, log_level=0)
mps, _, _ = m.fitMPS(Xr, yr, opts=mo)
hs = [None] * world
td.all_gather_object(hs, hashlib.sha256(b"".join(np.ascontiguousarray(c).tobytes() for c in mps.mps)).hexdigest())
assert len(set(hs)) == 1, hs
assert np.array_equal(m.classify(mps, Xr, distributed=True), m.classify(mps, Xr))
imp = m.init_imputation_problem(mps, Xr, yr, dx=1e-3, verbosity=-1)
inst = list(range(7)); miss = [[2, 3, 4]] * 3 + [[5, 6]] * 4
for method in ("median", "ITS"):
    a = m.get_predictions_batch(imp, int(yr.min()), inst, miss, method=method, distributed=True, return_err=True)
    b = m.get_predictions_batch(imp, int(yr.min()), inst, miss, method=method, return_err=True)
    assert all(np.array_equal(x, y2, equal_nan=True) for x, y2 in zip(a, b)), method
td.barrier()
if rank == 0: print("MULTI_OK")
'''


def test_two_rank_sweep_matches_oracle(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "multi_case.py"
    script.write_text(_SCRIPT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29741", str(script), ROOT],
                         capture_output=True, text=True, timeout=600)
    assert "MULTI_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
