"""Thin object wrapper over the C ABI: one `Context` per GPU.  numpy arrays in / out, Fortran
(Julia) memory order at the boundary.  This is what the Julia shim's `ccall`s do, in Python."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import ImputeOpts, TrainOpts, c_double_p, c_i32_p, c_i64_p, c_u8_p

BASIS_IDS = {
    "legendre_no_norm": 0, "legendre": 0, "legendre_norm": 1, "fourier": 2, "stoudenmire": 3,
    "sahand": 4, "uniform": 5, "precomputed": 100,
    # data-driven / time-dependent real bases: evaluated from per-site coefficient tables (set_encoding_table)
    "table_legendre_proj": 101, "table_sahand_legendre": 102, "table_split": 103,
}
BASIS_RANGE = {0: (-1.0, 1.0), 1: (-1.0, 1.0), 2: (-1.0, 1.0), 3: (0.0, 1.0), 4: (0.0, 1.0), 5: (0.0, 1.0),
               101: (-1.0, 1.0), 102: (-1.0, 1.0)}
LOSS_IDS = {"KLD": 0, "MSE": 1}
OPT_IDS = {"TSGO": 0, "GD": 1}
METHOD_IDS = {"median": 0, "mean": 1, "mode": 2, "ITS": 3}
TIMER_NAMES = ["encode", "flatten", "fwd", "grad", "update", "svd", "env", "allreduce", "impute", "grad_kernel"]


class MPSTError(RuntimeError):
    pass


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def make_opts(loss="KLD", bbopt="TSGO", train_sep=False, update_iters=1, rescale=(False, True), chi_max=25,
              eta=0.01, cutoff=1e-10):
    o = TrainOpts()
    o.loss_kind = LOSS_IDS[str(loss).upper()]
    o.opt_kind = OPT_IDS[str(bbopt).upper()]
    o.train_sep = int(bool(train_sep))
    o.update_iters = int(update_iters)
    o.rescale_before = int(bool(rescale[0]))
    o.rescale_after = int(bool(rescale[1]))
    o.chi_max = int(chi_max)
    o.eta = float(eta)
    o.cutoff = float(cutoff)
    return o


class Context:
    """Opaque mpst_ctx handle bound to one CUDA device."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.mpst_create(C.byref(h), int(device))
        if rc != 0:
            raise MPSTError(f"mpst_create failed ({rc}): no usable sm_100 GPU at index {device}; "
                            "this library has no CPU fallback")
        self.h = h
        self.device = device
        self.T = self.d = self.C = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.mpst_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise MPSTError(f"libmpstime_b200 error {rc}: {self.lib.mpst_last_error(self.h).decode()}")

    # ---- multi-GPU --------------------------------------------------------------------------
    def comm_unique_id(self):
        buf = (C.c_ubyte * 128)()
        rc = self.lib.mpst_comm_unique_id(buf)
        if rc != 0:
            raise MPSTError(f"mpst_comm_unique_id failed ({rc})")
        return bytes(buf)

    def comm_init(self, uid, rank, world):
        buf = (C.c_ubyte * 128).from_buffer_copy(uid)
        self._chk(self.lib.mpst_comm_init(self.h, buf, int(rank), int(world)))

    # ---- K1 ---------------------------------------------------------------------------------
    def encode(self, x, d, basis="legendre_no_norm"):
        """x: any shape, values in the basis range -> (..., d) float64 (complex128 for complex bases)."""
        bid = BASIS_IDS[basis.lower()]
        x = _f64(x)
        n = x.size
        cplx = bid in (2, 3, 4)
        out = np.empty((n, (2 if cplx else 1) * d), dtype=np.float64)
        self._chk(self.lib.mpst_encode(self.h, bid, int(d), _dp(x.reshape(-1)), n, _dp(out)))
        if cplx:
            out = out.view(np.complex128)
        return out.reshape(x.shape + (d,))

    def set_encoding_table(self, kind, n_sites, d, ip, dp):
        """Per-site coefficient tables of a data-driven / time-dependent encoding (encodings_host.py builds them)."""
        ip = np.ascontiguousarray(ip, dtype=np.int32).reshape(n_sites, -1)
        dp = np.ascontiguousarray(dp, dtype=np.float64).reshape(n_sites, -1)
        self._chk(self.lib.mpst_set_encoding_table(self.h, int(kind), int(n_sites), int(d), ip.ctypes.data_as(c_i32_p),
                                                   ip.shape[1], _dp(dp), dp.shape[1]))

    def encode_site(self, site, x):
        """K1 at one site of the context's current encoding -> (n, d)."""
        x = _f64(x).reshape(-1)
        out = np.empty((x.size, self.d), dtype=np.float64)
        self._chk(self.lib.mpst_encode_site(self.h, int(site), _dp(x), x.size, _dp(out)))
        return out

    # ---- training set -----------------------------------------------------------------------
    def model_init(self, T, C_, d, chi_max, basis="legendre_no_norm"):
        self._chk(self.lib.mpst_model_init(self.h, int(T), int(C_), int(d), int(chi_max),
                                           basis if isinstance(basis, int) else BASIS_IDS[basis.lower()]))
        self.T, self.C, self.d = int(T), int(C_), int(d)

    def train_load_x(self, X_TxN, class_counts, d, chi_max, basis="legendre_no_norm", n_global=0,
                     counts_global=None):
        """X_TxN: (T, N) array, series are columns, class-sorted, already in the encoding range."""
        X = np.asarray(X_TxN, dtype=np.float64)
        T, N = X.shape
        Xh = np.ascontiguousarray(X.T)                       # == Julia column-major T x N
        cc = np.ascontiguousarray(class_counts, dtype=np.int64)
        cg = np.ascontiguousarray(counts_global if counts_global is not None else class_counts, dtype=np.int64)
        self._chk(self.lib.mpst_train_load_x(self.h, _dp(Xh), N, T, cc.ctypes.data_as(c_i64_p), len(cc),
                                             basis if isinstance(basis, int) else BASIS_IDS[basis.lower()], int(d), int(chi_max), int(n_global or N),
                                             cg.ctypes.data_as(c_i64_p)))
        self.T, self.C, self.d = T, len(cc), int(d)

    def train_load_phi(self, phi_NTd, class_counts, chi_max, n_global=0, counts_global=None):
        """phi_NTd: (N, T, d) encoded samples (C order == Julia d x T x N column-major)."""
        phi = _f64(phi_NTd)
        N, T, d = phi.shape
        cc = np.ascontiguousarray(class_counts, dtype=np.int64)
        cg = np.ascontiguousarray(counts_global if counts_global is not None else class_counts, dtype=np.int64)
        self._chk(self.lib.mpst_train_load_phi(self.h, _dp(phi), N, T, cc.ctypes.data_as(c_i64_p), len(cc), d,
                                               int(chi_max), int(n_global or N), cg.ctypes.data_as(c_i64_p)))
        self.T, self.C, self.d = T, len(cc), d

    # ---- cores: python-side shape (chi_l, d, chi_r[, C]) ---------------------------------------
    def set_core(self, site, A):
        A = np.asarray(A, dtype=np.float64)
        has_label = A.ndim == 4
        wire = np.asfortranarray(A)                          # a fastest ... c slowest
        self._chk(self.lib.mpst_set_core(self.h, int(site), _dp(wire.reshape(-1, order="F")), A.shape[0], A.shape[2],
                                         int(has_label)))

    def set_cores(self, cores):
        for j, A in enumerate(cores):
            self.set_core(j, A)

    def get_core(self, site):
        cl, cr, hl = C.c_int32(), C.c_int32(), C.c_int32()
        self._chk(self.lib.mpst_get_core_dims(self.h, int(site), C.byref(cl), C.byref(cr), C.byref(hl)))
        shape = (cl.value, self.d, cr.value) + ((self.C,) if hl.value else ())
        flat = np.empty(int(np.prod(shape)), dtype=np.float64)
        self._chk(self.lib.mpst_get_core(self.h, int(site), _dp(flat)))
        return np.ascontiguousarray(flat.reshape(shape, order="F"))

    def get_cores(self):
        return [self.get_core(j) for j in range(self.T)]

    # ---- training ---------------------------------------------------------------------------
    def build_env(self, going_left=True):
        self._chk(self.lib.mpst_build_env(self.h, int(bool(going_left))))

    def bond_step(self, lid, going_left, opts):
        lo, gn, chi = C.c_double(), C.c_double(), C.c_int32()
        self._chk(self.lib.mpst_bond_step(self.h, int(lid), int(bool(going_left)), C.byref(opts), C.byref(lo),
                                          C.byref(gn), C.byref(chi)))
        return lo.value, gn.value, chi.value

    def bond_step_quiet(self, lid, going_left, opts):
        """bond step without reading loss / gradient norm back (no extra host sync)."""
        self._chk(self.lib.mpst_bond_step(self.h, int(lid), int(bool(going_left)), C.byref(opts), None, None, None))

    def sweep(self, opts, nsweeps, record=True):
        nb = nsweeps * 2 * (self.T - 1)
        if record:
            lo = np.zeros(nb)
            gn = np.zeros(nb)
            chi = np.zeros(nb, dtype=np.int32)
            self._chk(self.lib.mpst_sweep(self.h, C.byref(opts), int(nsweeps), _dp(lo), _dp(gn),
                                          chi.ctypes.data_as(c_i32_p)))
            return lo, gn, chi
        self._chk(self.lib.mpst_sweep(self.h, C.byref(opts), int(nsweeps), None, None, None))
        return None

    def sweep_bonds(self, opts, n_bonds, restart=False, record=True):
        """`n_bonds` consecutive bond updates of the sweep cycle, continuing where the last call stopped."""
        if record:
            lo = np.zeros(n_bonds)
            gn = np.zeros(n_bonds)
            chi = np.zeros(n_bonds, dtype=np.int32)
            self._chk(self.lib.mpst_sweep_bonds(self.h, C.byref(opts), int(n_bonds), int(bool(restart)), _dp(lo), _dp(gn),
                                                chi.ctypes.data_as(c_i32_p)))
            return lo, gn, chi
        self._chk(self.lib.mpst_sweep_bonds(self.h, C.byref(opts), int(n_bonds), int(bool(restart)), None, None, None))
        return None

    # ---- K7 ---------------------------------------------------------------------------------
    def overlaps(self, X_TxN=None, phi_NTd=None):
        """returns (yhat (n, C), argmax (n,) 0-based class index)."""
        if phi_NTd is not None:
            data = _f64(phi_NTd)
            n = data.shape[0]
        else:
            X = np.asarray(X_TxN, dtype=np.float64)
            n = X.shape[1]
            data = np.ascontiguousarray(X.T)
        yh = np.empty((n, self.C), dtype=np.float64)
        am = np.empty(n, dtype=np.int64)
        self._chk(self.lib.mpst_overlaps(self.h, _dp(data), n, _dp(yh), am.ctypes.data_as(c_i64_p)))
        return yh, am

    def eval_metrics(self, X_TxN=None, phi_NTd=None, labels=None):
        """MSE_loss_acc_conf on the device (summary.jl:60-114).  No data argument: the resident training set (this
        rank's shard).  Returns (sums [mse_sum, kld_sum, n_correct], conf (C, C) int64, n)."""
        sums = np.zeros(3)
        conf = np.zeros((self.C, self.C), dtype=np.int64)
        if X_TxN is None and phi_NTd is None:
            self._chk(self.lib.mpst_eval_metrics(self.h, None, 0, None, _dp(sums), conf.ctypes.data_as(c_i64_p)))
            return sums, conf, int(conf.sum())
        if phi_NTd is not None:
            data = _f64(phi_NTd)
            n = data.shape[0]
        else:
            X = np.asarray(X_TxN, dtype=np.float64)
            n = X.shape[1]
            data = np.ascontiguousarray(X.T)
        lab = np.ascontiguousarray(labels, dtype=np.int64)
        assert lab.shape == (n,)
        self._chk(self.lib.mpst_eval_metrics(self.h, _dp(data), n, lab.ctypes.data_as(c_i64_p), _dp(sums),
                                             conf.ctypes.data_as(c_i64_p)))
        return sums, conf, n

    # ---- K8 ---------------------------------------------------------------------------------
    def impute_batch(self, class_idx, X_TxN, missing_TxN, grid, method="median", uniforms=None, n_traj=1,
                     max_jump=-1.0, impute_order="forwards", get_err=False, rejection_threshold=None, max_trials=10,
                     return_err=False):
        """K8.  Returns out (n, n_traj, T) [and err (n, n_traj, T) with return_err].  `uniforms`: (n, n_traj, K_max)
        for plain ITS; with `rejection_threshold` a flat stream (n, U), U >= n_traj * K_max * max_trials."""
        if impute_order not in ("forwards", "backwards"):
            raise ValueError('impute_order must be either "forwards" or "backwards"')      # MPS_methods.jl:119
        if get_err or return_err or rejection_threshold is not None or impute_order != "forwards":
            return self._impute_batch_ex(class_idx, X_TxN, missing_TxN, grid, method, uniforms, n_traj, max_jump,
                                         impute_order, get_err, rejection_threshold, max_trials, return_err)
        X = np.asarray(X_TxN, dtype=np.float64)
        T, n = X.shape
        Xh = np.ascontiguousarray(X.T)
        mh = np.ascontiguousarray(np.asarray(missing_TxN, dtype=np.uint8).T)
        grid = _f64(grid)
        out = np.empty((n, n_traj, T), dtype=np.float64)
        u = None
        if uniforms is not None:
            u = _f64(uniforms)
        self._chk(self.lib.mpst_impute_batch(self.h, int(class_idx), _dp(Xh), mh.ctypes.data_as(c_u8_p), n,
                                             METHOD_IDS[method], _dp(grid), len(grid),
                                             _dp(u) if u is not None else None, int(n_traj), float(max_jump), _dp(out)))
        return out

    def _impute_batch_ex(self, class_idx, X_TxN, missing_TxN, grid, method, uniforms, n_traj, max_jump, impute_order,
                         get_err, rejection_threshold, max_trials, return_err):
        X = np.asarray(X_TxN, dtype=np.float64)
        T, n = X.shape
        Xh = np.ascontiguousarray(X.T)
        mh = np.ascontiguousarray(np.asarray(missing_TxN, dtype=np.uint8).T)
        grid = _f64(grid)
        nt = n_traj if method == "ITS" else 1
        out = np.empty((n, nt, T), dtype=np.float64)
        err = np.zeros((n, nt, T), dtype=np.float64)
        io = ImputeOpts()
        io.backwards = int(impute_order == "backwards")
        io.get_err = int(bool(get_err))
        io.max_trials = int(max_trials)
        io.rejection_threshold = -1.0 if rejection_threshold is None else float(rejection_threshold)
        io.max_jump = -1.0 if max_jump is None else float(max_jump)
        u, per = None, 0
        if uniforms is not None:
            u = _f64(uniforms)
            per = u.size // max(n, 1)
        self._chk(self.lib.mpst_impute_batch_ex(self.h, int(class_idx), _dp(Xh), mh.ctypes.data_as(c_u8_p), n,
                                                METHOD_IDS[method], _dp(grid), len(grid),
                                                _dp(u) if u is not None else None, int(per), int(nt), C.byref(io),
                                                _dp(out), _dp(err)))
        return (out, err) if return_err else out

    # ---- test entries -------------------------------------------------------------------------
    def bond_loss_grad(self, B, L, R, xl, xr, class_counts, loss="KLD", train_sep=False, want_yhat=False):
        """B (D, C); L (N, chi_l); R (N, chi_r); xl, xr (N, d).  Returns (loss, grad (D, C)[, yhat (N, C)])."""
        B = np.asarray(B, dtype=np.float64)
        D, Cn = B.shape
        L, R, xl, xr = _f64(L), _f64(R), _f64(xl), _f64(xr)
        N, d = xl.shape
        cc = np.ascontiguousarray(class_counts, dtype=np.int64)
        Bw = np.ascontiguousarray(B.T)                        # column-major D x C
        G = np.empty((Cn, D), dtype=np.float64)
        lo = C.c_double()
        yh = np.empty((N, Cn), dtype=np.float64) if want_yhat else None
        self._chk(self.lib.mpst_bond_loss_grad(self.h, _dp(Bw), _dp(L), _dp(R), _dp(xl), _dp(xr), N, d, L.shape[1],
                                               R.shape[1], cc.ctypes.data_as(c_i64_p), Cn, LOSS_IDS[loss.upper()],
                                               int(bool(train_sep)), C.byref(lo), _dp(G),
                                               _dp(yh) if want_yhat else None))
        self.T = 0
        if want_yhat:
            return lo.value, np.ascontiguousarray(G.T), yh
        return lo.value, np.ascontiguousarray(G.T)

    def bond_split(self, B, d, chi_l, chi_r, going_left, chi_max, cutoff=1e-10):
        """B (D, C) -> (core_l, core_r, sigma) in python core shapes."""
        B = np.asarray(B, dtype=np.float64)
        D, Cn = B.shape
        Bw = np.ascontiguousarray(B.T)
        kmax = max(1, min(chi_max, d * (chi_r if going_left else chi_l)))
        cl_shape = (chi_l, d, kmax) + ((Cn,) if going_left else ())
        cr_shape = (kmax, d, chi_r) + (() if going_left else (Cn,))
        cl = np.zeros(int(np.prod(cl_shape)))
        cr = np.zeros(int(np.prod(cr_shape)))
        sig = np.zeros(kmax)
        chi = C.c_int32()
        self._chk(self.lib.mpst_bond_split(self.h, _dp(Bw), d, chi_l, chi_r, Cn, int(bool(going_left)), int(chi_max),
                                           float(cutoff), C.byref(chi), _dp(cl), _dp(cr), _dp(sig)))
        k = chi.value
        cl_shape = (chi_l, d, k) + ((Cn,) if going_left else ())
        cr_shape = (k, d, chi_r) + (() if going_left else (Cn,))
        core_l = cl[: int(np.prod(cl_shape))].reshape(cl_shape, order="F")
        core_r = cr[: int(np.prod(cr_shape))].reshape(cr_shape, order="F")
        return np.ascontiguousarray(core_l), np.ascontiguousarray(core_r), sig[:k].copy()

    # ---- profiling ----------------------------------------------------------------------------
    def profile_enable(self, on=True):
        self._chk(self.lib.mpst_profile_enable(self.h, int(on)))

    def profile_reset(self):
        self._chk(self.lib.mpst_profile_reset(self.h))

    def profile_get(self):
        """{family: (device ms, launches, algorithmic work)} since the last reset."""
        ms = np.zeros(len(TIMER_NAMES))
        n = np.zeros(len(TIMER_NAMES), dtype=np.int64)
        wk = np.zeros(len(TIMER_NAMES))
        self._chk(self.lib.mpst_profile_get(self.h, _dp(ms), n.ctypes.data_as(c_i64_p), _dp(wk)))
        return {k: (float(ms[i]), int(n[i]), float(wk[i])) for i, k in enumerate(TIMER_NAMES)}

    def debug_set(self, name, value):
        self._chk(self.lib.mpst_debug_set(self.h, name.encode(), int(value)))

    def debug_get(self, name):
        return int(self.lib.mpst_debug_get(self.h, name.encode()))

    def timer_start(self):
        self._chk(self.lib.mpst_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        self._chk(self.lib.mpst_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(self.lib.mpst_launch_count(self.h))
