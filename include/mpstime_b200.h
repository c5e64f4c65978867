/*
 * mpstime_b200.h -- C ABI of libmpstime_b200.so: the B200-native (sm_100a) implementation of the
 * MPSTime.jl hot path (fitMPS two-site sweep, classify, MPS_impute).
 *
 * The reference (hugopstackhouse/MPSTime.jl) is pure Julia and has no FFI today; these entry
 * points are what a Julia shim would `ccall` at the seams named next to each function
 * (paths relative to the reference's src/).  INTEGRATION.md shows the Julia side.
 *
 * Conventions
 *  - every function returns 0 on success, a negative MPST_E_* code otherwise; the message is
 *    available from mpst_last_error(ctx).  No exception crosses the boundary.
 *  - all pointer arguments are caller-owned HOST memory, Float64 / Int64, column-major exactly as
 *    Julia lays arrays out.  Device memory is owned by the opaque context.
 *  - a context is bound to one CUDA device and is not re-entrant (one Julia task at a time).
 *  - sites are 0-based in this ABI (Julia site j  <->  j-1 here).
 *  - core layout on the wire: dims (chi_left, d, chi_right[, C]) column-major, i.e.
 *    idx = a + chi_left*(s + d*(b + chi_right*c)); the class axis C is present only on the one
 *    core that carries the label index "f(x)" (utils.jl:342-354 find_label).
 *  - bond tensor layout: the reference's `BondTensor = Matrix` (D x C), column c flattened with
 *    s_l fastest: idx = s_l + d*(a + chi_l*(s_r + d*b))  (Training/loss_functions.jl:193-262).
 */
#ifndef MPSTIME_B200_H
#define MPSTIME_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mpst_ctx mpst_ctx;

enum {
    MPST_OK = 0,
    MPST_E_INVALID = -1,   /* bad argument / wrong call order */
    MPST_E_CUDA = -2,      /* CUDA runtime error */
    MPST_E_NCCL = -3,      /* NCCL error or libnccl not loadable */
    MPST_E_NUMERIC = -4,   /* non-finite loss, SVD did not converge */
    MPST_E_UNSUPPORTED = -5
};

/* Encodings/bases.jl: data-independent real/complex bases (basis_structs.jl:101-283) */
enum {
    MPST_BASIS_LEGENDRE_NO_NORM = 0, /* bases.jl:77-92, norm=false (default encoding, options.jl:111) */
    MPST_BASIS_LEGENDRE_NORM = 1,    /* bases.jl:86-89 */
    MPST_BASIS_FOURIER = 2,          /* bases.jl:23-42   (complex: out is re,im interleaved) */
    MPST_BASIS_STOUDENMIRE = 3,      /* bases.jl:13-20   (complex, d = 2) */
    MPST_BASIS_SAHAND = 4,           /* bases.jl:53-74   (complex, d even) */
    MPST_BASIS_UNIFORM = 5,          /* bases.jl:2-5 */
    MPST_BASIS_PRECOMPUTED = 100,    /* caller passes phi (custom bases) */
    /* data-driven / time-dependent real bases: the host computes per-site coefficient tables once
     * (bases.jl:294-397, splitbases.jl:13-110), the device evaluates them (mpst_set_encoding_table) */
    MPST_BASIS_TABLE_LEGENDRE_PROJ = 101,   /* projected Legendre, bases.jl:95-108, 381-397           */
    MPST_BASIS_TABLE_SAHAND_LEGENDRE = 102, /* Sahand-Legendre (time dependent or not), :111-129      */
    MPST_BASIS_TABLE_SPLIT = 103            /* histogram / uniform split bases, splitbases.jl:113-163 */
};

enum { MPST_LOSS_KLD = 0, MPST_LOSS_MSE = 1 };     /* loss_functions.jl:322-432 / 561-619 */
enum { MPST_OPT_TSGO = 0, MPST_OPT_GD = 1 };       /* loss_functions.jl:59-86 / 27-57 */
enum { MPST_IMPUTE_MEDIAN = 0, MPST_IMPUTE_MEAN = 1, MPST_IMPUTE_MODE = 2, MPST_IMPUTE_ITS = 3 };

/* Mirrors the MPSOptions fields the sweep reads (Structs/options.jl:106-134). */
typedef struct mpst_train_opts {
    int32_t loss_kind;        /* loss_grad  :KLD | :MSE                         */
    int32_t opt_kind;         /* bbopt      :TSGO | :GD                         */
    int32_t train_sep;        /* train_classes_separately (KLD only)            */
    int32_t update_iters;     /* update_iters                                   */
    int32_t rescale_before;   /* rescale[1]                                     */
    int32_t rescale_after;    /* rescale[2]                                     */
    int32_t chi_max;          /* chi_max                                        */
    int32_t reserved;
    double eta;               /* eta                                            */
    double cutoff;            /* cutoff (relative, sum of discarded sigma^2)    */
} mpst_train_opts;

/* ---- lifetime -------------------------------------------------------------------------- */
int mpst_version(void);
int mpst_create(mpst_ctx** out, int device_id);
/* model without a training set (classify / impute with cores passed in): T sites, C classes. */
int mpst_model_init(mpst_ctx* ctx, int T, int C, int d, int chi_max, int basis_id);
int mpst_destroy(mpst_ctx* ctx);
const char* mpst_last_error(mpst_ctx* ctx);

/* ---- multi-GPU plumbing (one process per GPU; the only collective is the per-bond gradient
 *      all-reduce).  The 128-byte id comes from rank 0 and is broadcast by the host
 *      language (Distributed.jl / torch.distributed). ------------------------------------- */
int mpst_comm_unique_id(void* id128);
int mpst_comm_init(mpst_ctx* ctx, const void* id128, int rank, int world);

/* ---- K1: basis encoding.  Replaces encode_TS -> Basis.encode (Encodings/encodings.jl:18-27,
 *      bases.jl:13-92).  x: n values already in the encoding range; out: d x n column-major
 *      (2 x d x n for complex bases). ---------------------------------------------------- */
int mpst_encode(mpst_ctx* ctx, int basis_id, int d, const double* x, int64_t n, double* out);

/* ---- coefficient tables of a data-driven / time-dependent encoding (replaces the `encoding_args` that
 *      opts.encoding.init returns and encode_TS threads through, Encodings/encodings.jl:1-31).  n_sites = 1 for a
 *      time-independent basis, else T.  Per site `ni` Int32 and `nd` Float64 values:
 *        LEGENDRE_PROJ  : ip = d orders (0-based) followed by the inverse map order -> slot (-1 = unused) for orders
 *                         0..L;  dp = { scale, L }
 *        SAHAND_LEGENDRE: ip = { npts, has_data };  dp = { x_1, h, minx, scale, cVecs[d*d] (row n = basis function,
 *                         column i = power of x), c[npts + 2] quadratic-B-spline coefficients of the KDE density }
 *        SPLIT          : ip = { nbins, aux_dim, aux_basis_id };  dp = nbins + 1 bin edges
 *      Call before mpst_train_load_x / mpst_model_init with basis_id = kind; the table stays until replaced. ---- */
int mpst_set_encoding_table(mpst_ctx* ctx, int kind, int n_sites, int d, const int32_t* ip, int64_t ni,
                            const double* dp, int64_t nd);
/* K1 at one site of the context's current encoding (table or built-in): x: n values, out: d x n column-major. */
int mpst_encode_site(mpst_ctx* ctx, int site, const double* x, int64_t n, double* out);

/* ---- training set.  Replaces the PState / EncodedTimeSeriesSet / PCache containers
 *      (Structs/structs.jl:2-33) for the 4-arg fitMPS seam (RealRealHighDimension.jl:587).
 *      Samples MUST be sorted by class (asserted at :624); class_counts = class_distribution
 *      (encodings.jl:151-152).  X: T x N column-major (X_train_scaled, series are columns).
 *      n_global / counts_global: totals over all ranks (== local values on one GPU). ------- */
int mpst_train_load_x(mpst_ctx* ctx, const double* X, int64_t N, int T, const int64_t* class_counts,
                      int C, int basis_id, int d, int chi_max, int64_t n_global,
                      const int64_t* counts_global);
/* phi: d x T x N column-major (pstate[j][s] of sample i at s + d*(j + T*i)). */
int mpst_train_load_phi(mpst_ctx* ctx, const double* phi, int64_t N, int T,
                        const int64_t* class_counts, int C, int d, int chi_max, int64_t n_global,
                        const int64_t* counts_global);

/* ---- MPS cores in / out (dense-ified ITensors; generate_startingMPS stays on the host,
 *      RealRealHighDimension.jl:1-41). --------------------------------------------------- */
int mpst_set_core(mpst_ctx* ctx, int site, const double* data, int chi_l, int chi_r, int has_label);
int mpst_get_core_dims(mpst_ctx* ctx, int site, int* chi_l, int* chi_r, int* has_label);
int mpst_get_core(mpst_ctx* ctx, int site, double* out);

/* ---- K6 x (T-1): construct_caches (RealRealHighDimension.jl:45-103). -------------------- */
int mpst_build_env(mpst_ctx* ctx, int going_left);

/* ---- one bond: flatten_bt -> apply_update -> decomposeBT -> update_caches!
 *      (RealRealHighDimension.jl:733-762 / 777-801).  lid = left site of the bond (0-based).
 *      loss_out / gradnorm_out: loss and ||grad||_F of the first optimiser iteration. -------- */
int mpst_bond_step(mpst_ctx* ctx, int lid, int going_left, const mpst_train_opts* opts,
                   double* loss_out, double* gradnorm_out, int* chi_new_out);

/* ---- the sweep loop of fitMPS(W, train, test, opts) (RealRealHighDimension.jl:726-852):
 *      builds LE, runs nsweeps x (backward + forward) half-sweeps, then normalize!(W).
 *      Optional per-bond outputs have nsweeps*2*(T-1) entries (NULL to skip). ---------------- */
int mpst_sweep(mpst_ctx* ctx, const mpst_train_opts* opts, int nsweeps, double* per_bond_loss,
               double* per_bond_gradnorm, int32_t* per_bond_chi);

/* ---- a block of `n_bonds` consecutive bond updates of the same (backward, forward) cycle that mpst_sweep runs,
 *      continuing where the previous call stopped (restart != 0, or the first call after mpst_set_core /
 *      mpst_train_load_*, rebuilds LE and starts at the right edge; the label index must then be on the last
 *      site).  No normalize!(W) at the end: this is the sweep loop cut into pieces, for callers that interleave
 *      logging / early exit with training (RealRealHighDimension.jl:726-850) and for bench.py's fixed-size steps
 *      at the north-star shape.  Optional per-bond outputs have n_bonds entries. ------------------------------ */
int mpst_sweep_bonds(mpst_ctx* ctx, const mpst_train_opts* opts, int n_bonds, int restart, double* per_bond_loss,
                     double* per_bond_gradnorm, int32_t* per_bond_chi);

/* ---- K7: contract_mps / classify / MSE_loss_acc (summary.jl:4-136).  X: T x n column-major in
 *      the encoding range (basis from the loaded training set) or phi (d x T x n) when the
 *      context was loaded with precomputed phi.  yhat: C x n column-major; argmax: n (0-based
 *      class index, first maximum of |yhat|^2). ------------------------------------------- */
int mpst_overlaps(mpst_ctx* ctx, const double* X_or_phi, int64_t n, double* yhat, int64_t* argmax);

/* ---- per-sweep evaluation on the device: MSE_loss_acc / MSE_loss_acc_conf (summary.jl:33-114), which fitMPS runs on
 *      the train and test sets after every sweep when log_level > 0 (RealRealHighDimension.jl:813-845).  Runs the K7
 *      chain and reduces on the device; only 3 + C*C numbers come back.  X_or_phi as in mpst_overlaps with
 *      label_idx[i] the 0-based class index of sample i; X_or_phi == NULL evaluates the training set already resident
 *      on the device (labels = its sorted class ranges; n and label_idx ignored; on a sharded run: this rank's shard).
 *      sums[3] = { sum_i 0.5*|yhat_i - onehot_i|^2, sum_i -log|yhat_i,label|^2, number correct } (the caller divides by
 *      the global sample count); conf: C x C row-major counts conf[true][predicted] (NULL to skip). ------------------ */
int mpst_eval_metrics(mpst_ctx* ctx, const double* X_or_phi, int64_t n, const int64_t* label_idx, double* sums,
                      int64_t* conf);

/* ---- K8: MPS_impute batch (Imputation/imputation.jl:264-410 get_predictions ->
 *      MPS_methods.jl:42-180 precondition + impute_at! -> sampling_utils.jl:64-316).
 *      Uses the class_idx slice of the context's current MPS (expand_label_index,
 *      utils.jl:356-370).  X: T x n scaled series with missing entries already filled
 *      (imputation.jl:290-291); missing: T x n bytes (1 = impute); xgrid: G grid points
 *      (imputation.jl:90); uniforms: K_max x n_traj x n draws for ITS (NULL otherwise);
 *      out: T x n_traj x n; max_jump < 0 disables the mode jump filter. ------------------- */
int mpst_impute_batch(mpst_ctx* ctx, int class_idx, const double* X, const uint8_t* missing,
                      int64_t n, int method, const double* xgrid, int G, const double* uniforms,
                      int n_traj, double max_jump, double* out);

/* ---- the keyword arguments of impute_median / impute_mean / impute_mode / impute_ITS (MPS_methods.jl:201-347) -- */
typedef struct mpst_impute_opts {
    int32_t backwards;           /* impute_order = :backwards (MPS_methods.jl:113-118); 0 = :forwards             */
    int32_t get_err;             /* get_wmad (median, sampling_utils.jl:192-196) / get_std (mean, :89-97)         */
    int32_t max_trials;          /* ITS: draws per site when rejecting (reference default 10)                     */
    int32_t reserved;
    double rejection_threshold;  /* ITS: accept when |x - median| < threshold * WMAD (:295-311); < 0 = :none      */
    double max_jump;             /* mode: jump filter (:104-158); < 0 = nothing                                   */
} mpst_impute_opts;
/* As mpst_impute_batch with every option.  err_out (T x n_traj x n, or NULL): the method's error bar at each imputed
 * site (WMAD for median with get_err and for ITS with rejection, standard deviation for mean with get_err, else 0).
 * With rejection the uniforms are a flat stream per instance -- uniforms_per_instance >= n_traj * K_max * max_trials
 * values, consumed in order over sites and trajectories exactly as the reference's shared MersenneTwister is
 * (MPS_methods.jl:324); without rejection the layout is that of mpst_impute_batch and uniforms_per_instance is
 * ignored. */
int mpst_impute_batch_ex(mpst_ctx* ctx, int class_idx, const double* X, const uint8_t* missing, int64_t n, int method,
                         const double* xgrid, int G, const double* uniforms, int64_t uniforms_per_instance,
                         int n_traj, const mpst_impute_opts* opts, double* out, double* err_out);

/* ---- test / benchmark entry: teacher-forced K2 in isolation on caller-provided operands
 *      (Loss_Grad_KLD / Loss_Grad_MSE, loss_functions.jl:322-432, 561-619).
 *      B: D x C; L: chi_l x N; R: chi_r x N; xl, xr: d x N (all column-major, i.e. one sample per
 *      column).  grad_out: D x C.  yhat_out: C x N (NULL to skip; KLD fills own-class only). - */
int mpst_bond_loss_grad(mpst_ctx* ctx, const double* B, const double* L, const double* R,
                        const double* xl, const double* xr, int64_t N, int d, int chi_l, int chi_r,
                        const int64_t* class_counts, int C, int loss_kind, int train_sep,
                        double* loss_out, double* grad_out, double* yhat_out);

/* ---- test entry: K5 in isolation (decomposeBT, RealRealHighDimension.jl:146-203).
 *      B: D x C bond tensor; returns the two new cores in the wire layout
 *      (core_l: chi_l x d x chi_new [x C if !going_left... see DESIGN.md]), sigma: chi_new. --- */
int mpst_bond_split(mpst_ctx* ctx, const double* B, int d, int chi_l, int chi_r, int C,
                    int going_left, int chi_max, double cutoff, int* chi_new, double* core_l,
                    double* core_r, double* sigma);

/* ---- timing hooks for bench.py: device time (ms, CUDA events on the context's stream) spent in
 *      each kernel family since the last reset, and launch counts. ------------------------ */
enum { MPST_T_ENCODE = 0, MPST_T_FLATTEN, MPST_T_FWD, MPST_T_GRAD, MPST_T_UPDATE, MPST_T_SVD,
       MPST_T_ENV, MPST_T_ALLREDUCE, MPST_T_IMPUTE, MPST_T_GRADK /* bond_grad_kernel alone */, MPST_T_COUNT };
int mpst_profile_enable(mpst_ctx* ctx, int on);
int mpst_profile_get(mpst_ctx* ctx, double* ms /*MPST_T_COUNT*/, int64_t* launches /*MPST_T_COUNT*/,
                     double* work /*MPST_T_COUNT: algorithmic flops (GEMM families) or bytes (encode)*/);
/* device-side stopwatch on the context's stream (CUDA events): start, ..., stop -> elapsed ms */
int mpst_timer_start(mpst_ctx* ctx);
int mpst_timer_stop(mpst_ctx* ctx, double* ms);
int mpst_profile_reset(mpst_ctx* ctx);
int64_t mpst_launch_count(mpst_ctx* ctx);

/* ---- test / experiment switches.  The library reads MPST_<NAME> environment variables exactly once, in
 *      mpst_create; afterwards a switch changes only through mpst_debug_set (name without the MPST_ prefix, case
 *      insensitive: "NO_ENV_REUSE", "SVD_NOSUB", "SVD_NOHALF", "GRAD_NOKR", "KRAO_NOREG", "DENSE_FWD", ...).
 *      mpst_debug_get returns a switch, or which code path the last call took: "svd_path" (1 tall-Gram, 2 wide-Gram,
 *      3 subspace iteration, 4 fused Jacobi, 5 three-kernel Jacobi), "svd_iters", "svd_restarts", "grad_kernel"
 *      (1 register-operand, 2 shared-memory tiles), "grad_variant", "krao_kernel" (1 register-operand, 2 tiles),
 *      "krao_variant", "fwd_path" (1 factorised + cached environment, 2 factorised, 3 dense); -1 = unknown name. */
int mpst_debug_set(mpst_ctx* ctx, const char* name, int value);
int64_t mpst_debug_get(mpst_ctx* ctx, const char* name);

#ifdef __cplusplus
}
#endif
#endif /* MPSTIME_B200_H */
