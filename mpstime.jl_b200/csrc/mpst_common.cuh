// Shared declarations for libmpstime_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/mpstime_b200.h"

#define MPST_MAX_D 32
#define MPST_MAX_CHI 128
#define MPST_TILE 128          // sample padding granularity

// ---------------------------------------------------------------------------------------------
// device storage of one MPS core.  Two orientations: the link that points towards
// the bond being optimised is the slowest ("k") index so that the core is directly the weight
// matrix [p = s + d*env_link][k] of the Khatri-Rao GEMM that updates the environment.
//   LEFT : idx = s + d*(a + chi_l*b) + d*chi_l*chi_r*c     (site, left link, right link[, class])
//   RIGHT: idx = s + d*(b + chi_r*a) + d*chi_l*chi_r*c     (site, right link, left link[, class])
// ---------------------------------------------------------------------------------------------
enum { ORIENT_LEFT = 0, ORIENT_RIGHT = 1 };

struct Core {
    double* dev = nullptr;
    size_t cap = 0;   // doubles allocated
    int chi_l = 0, chi_r = 0, has_label = 0, orient = ORIENT_LEFT;
};

struct CoreView {       // strides (in doubles) of (s, a = left link, b = right link, c = class)
    const double* p;
    long ss, sa, sb, sc;
};

struct GradSeg {        // one stream-K segment of the gradient GEMM (host-built, see bond_grad)
    int cls, tp, tq, slot;
    int64_t chunk_begin, chunk_end;
};

struct Timer {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
};

// Run-time switches for tests / experiments.  They are read from the environment (MPST_<NAME>) exactly ONCE, in
// mpst_create, and can be changed afterwards only through mpst_debug_set: a stray variable on one rank cannot
// change the numerics half-way through a sharded sweep, and nothing on the per-bond path calls getenv().
enum {
    F_DENSE_FWD = 0, F_NO_ENV_REUSE, F_GRAD_T128, F_GRAD_NOKR, F_IMPUTE_NODBUF, F_IMPUTE_DEBUG, F_KRAO_NOREG,
    F_SVD_INNER, F_SVD_DEBUG, F_SVD_SKIP, F_SVD_FIXED, F_SVD_FULL, F_SVD_PB64, F_SVD_LEGACY, F_SVD_NOSUB, F_SVD_OVS,
    F_SVD_NOHALF, F_SVD_HALF_FROM, F_SVD_IT, F_SVD_NOGRAPH, F_GRAD_KC, F_IMPUTE_NOSERIES, F_IMPUTE_FULLSYM, F_GRAD_PHASES, F_SVD_EIGSMEM, F_SVD_SERIAL, F_SVD_CHOLSEQ, F_SVD_PROBE, F_SVD_SYNCFIRST, F_SVD_NOPREP, F_KRAO_NOSLAB, F_KRAO_SLAB_MI, F_KRAO_SLAB_MIN, F_SVD_FIRST, F_SVD_NO2PASS, F_SVD_GRAMREG, F_COUNT
};
// which code path the last call took (mpst_debug_get): lets the parity tests assert that they exercised the
// kernels the benchmark runs, and lets bench.py name the kernel it reports a roofline for
enum {
    L_SVD_PATH = 0,    // 1 tall-Gram, 2 wide-Gram, 3 subspace iteration, 4 fused Jacobi, 5 three-kernel Jacobi
    L_SVD_ITERS,       // subspace iterations of the last split
    L_SVD_RESTARTS,    // restarts without the column-scaling shortcut
    L_GRAD_KERNEL,     // 1 bond_grad_kr_kernel (register operands), 2 bond_grad_kernel (shared-memory tiles)
    L_GRAD_VARIANT,    // kr: MA*10000 + S*1000 + KC*10 + SR;  tiles: TP*1000 + TQ
    L_KRAO_KERNEL,     // 1 krao_reg_kernel, 2 krao_gemm_kernel
    L_KRAO_VARIANT,    // reg: NI;  tiles: TN
    L_FWD_PATH,        // 1 factorised + cached environment, 2 factorised, 3 dense
    L_KRAO_REG_MASK,   // bit NI set for every krao_reg_kernel<NI> launched since the mask was last cleared (debug_set)
    L_GRAD_KR_LAUNCHES,   // launches of bond_grad_kr_kernel / bond_grad_kernel since last cleared (debug_set): bench.py
    L_GRAD_TILE_LAUNCHES, // names the kernel that did most of the timed work instead of assuming one
    L_SVD_CALLS,          // cumulative SVD statistics (cleared with debug_set): splits, subspace iterations summed over the
    L_SVD_ITERS_SUM,      // fast-path splits, splits that needed a second round of iterations, splits that fell back to the
    L_SVD_ROUND2,         // exact Jacobi, fast-path splits (Gram or subspace)
    L_SVD_JACOBI,
    L_SVD_FAST,
    L_SVD_SERIAL,         // fast-path splits that ran the serial (fully orthonormalising) loop
    L_KRAO_SLAB_LAUNCHES, // launches of krao_slab_kernel since last cleared
    L_SVD_TWOPASS,        // fast-path splits that needed the deflated second pass (chi_max > 80)
    L_COUNT
};

struct EncTable {        // per-site coefficient tables of a data-driven / time-dependent encoding (encode_table.cu)
    int kind = 0, nsites = 0, d = 0;
    int64_t istride = 0, dstride = 0;
    int* ip = nullptr;      // device [nsites][istride]
    double* dp = nullptr;   // device [nsites][dstride]
};

struct SegTable {       // cached stream-K schedule of one (kernel variant, shape, class ranges) combination
    std::vector<int64_t> key;
    GradSeg* segs = nullptr;     // device
    int* cta_ptr = nullptr;
    int* tile_slot = nullptr;
    int nseg = 0;
    uint64_t last_use = 0;
};

struct SvdGraph {       // one captured round of the subspace SVD (svd_subspace.cu)
    std::vector<int64_t> key;
    void* exec = nullptr;        // cudaGraphExec_t
    int64_t launches = 0;
    uint64_t last_use = 0;
};

struct mpst_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    // training set
    int64_t N = 0, Npad = 0, Nglobal = 0;
    int T = 0, d = 0, C = 0, chi_max = 0, basis = 0;
    bool have_phi = false;
    std::vector<int64_t> counts, counts_global, class_off;
    double* X = nullptr;        // [T][Npad]   (x mode)
    double* PHI = nullptr;      // [T][Npad][d] (phi mode)
    double* phi_l = nullptr;    // [Npad][d] two-site window
    double* phi_r = nullptr;
    double* env = nullptr;      // [T][Npad][chi_max] shared LE/RE slots (SURVEY 7 "cache memory")
    std::vector<int> env_chi;   // link dim held by each slot
    double* ones = nullptr;     // [Npad] == 1.0 (dummy boundary environments)
    std::vector<Core> cores;
    // bond scratch
    double* B = nullptr;        // [C][D]
    double* G = nullptr;        // [C][D] (+1 slot for the loss, contiguous for the all-reduce)
    size_t Dcap = 0;
    double* yhat = nullptr;     // [C][Npad]
    double* w = nullptr;        // [C][Npad]
    double* Z = nullptr;        // scratch rows x ldz for the dense forward
    size_t Zcap = 0;
    double* part = nullptr;     // stream-K partial tiles
    size_t partcap = 0;
    double* red = nullptr;      // reduction scratch (doubles)
    size_t redcap = 0;
    double* scal = nullptr;     // device scalars [16]
    double* hscal = nullptr;    // pinned host mirror
    std::vector<SegTable> segtabs;   // stream-K schedules, built once per shape (no per-bond host work / sync)
    uint64_t seg_clock = 0;
    EncTable enc;
    int kr_deep = 1;            // 1: prefer the 64-sample x 4-stage ring of the register-operand gradient kernel where it fits
    int flag[F_COUNT] = {0};
    int last[L_COUNT] = {0};
    // cursor of mpst_sweep_bonds: next bond of the (backward, forward) cycle, -1 = not started
    int sw_cursor = -1;
    int* nonfinite = nullptr;   // device flag: a loss / gradient weight / norm was NaN or Inf (checked once per bond)
    // jacobi workspace
    double* S = nullptr;        // (m+n) x npad column-major
    size_t Scap = 0;
    double* gpart = nullptr;
    size_t gpartcap = 0;
    double* wbuf = nullptr;
    size_t wbufcap = 0;
    double* colnorm = nullptr;  // [npad]
    int* perm = nullptr;        // [npad]
    double* sub = nullptr;      // subspace-SVD workspace
    size_t subcap = 0;
    // deflated two-pass split (chi_max > 80, svd_subspace.cu): while st_on, a finished subspace pass hands its triplets to
    // the staging blocks stU (m x k, U S) / stV (n x k) / stP (sigma^2) at column st_col0 instead of writing cores;
    // st_kept = device scalar holding the weight the first pass kept (nullptr during the first pass)
    bool st_on = false;
    int st_col0 = 0;
    double* st_kept = nullptr;
    double *stU = nullptr, *stV = nullptr, *stP = nullptr;
    double* stbuf = nullptr;
    size_t stbufcap = 0;
    double* gws = nullptr;      // split-K partial products of the small GEMMs
    double* kslab = nullptr;    // K6: slab-major copy of the weight matrix (krao_slab.cu)
    size_t kslabcap = 0;
    double* gws2 = nullptr;     // ... of the GEMMs on the side stream
    size_t gws2cap = 0;
    cudaStream_t stream2 = nullptr;            // side stream of the subspace SVD (Gram + Cholesky next to the big product)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_upload = nullptr;
    bool svd_prepare_only = false;             // svd_subspace_device: build / upload the round's CUDA graph, launch nothing
    void* svd_uploaded = nullptr;              // graph exec whose upload is in flight on the side stream
    void* imp_ptr[16] = {nullptr};   // imputation work buffers (grow-only, reused across mpst_impute_batch calls)
    size_t imp_cap[16] = {0};
    // per-bond subspace-iteration count learned during training (svd_subspace.cu): svd_slot = bond being split
    // (-1: none), svd_its[b] = iterations to start with, svd_floor[b] = smallest count that has not failed yet
    // provenance of the environment slots: slot j is reusable as the unlabelled forward factor of a bond when it
    // was computed (dir 1 = LE, 2 = RE) from the current core j and the current neighbouring slot
    uint64_t gen = 1;
    std::vector<uint64_t> core_ver, env_ver, env_core_ver, env_src_ver;
    std::vector<int> env_dir;
    int svd_slot = -1;
    std::vector<int> svd_its, svd_floor, svd_calm;
    // iteration count proven on the most recently split bond of the same shape: a bond without history of its own starts
    // from its neighbour's count instead of the conservative default (spectra change slowly along the chain)
    int svd_hint_m = 0, svd_hint_n = 0, svd_hint_its = 0, svd_hint_floor = 0;
    std::vector<SvdGraph> svd_graphs;
    std::vector<char> svd_nohalf, svd_pen;   // bonds on which the column-scaling shortcut of the subspace iteration broke down once
    // capacities (doubles) of the training buffers: a re-load with the same or a smaller shape reuses them
    size_t cap_X = 0, cap_PHI = 0, cap_phi = 0, cap_env = 0, cap_ones = 0, cap_yw = 0;
    size_t gwscap = 0;
    int* flags = nullptr;       // dataflow counters of the fused Jacobi sweep
    size_t flagcap = 0;
    int* iscal = nullptr;       // device ints [16]
    int* hiscal = nullptr;      // pinned
    double* meta = nullptr;     // class offsets (int64) + loss normalisers, 256 doubles
    double* hmeta = nullptr;    // pinned
    int meta_key = 0;
    double* tmp = nullptr;      // generic staging buffer
    size_t tmpcap = 0;
    int sm_count = 148;
    // nccl
    void* nccl_comm = nullptr;
    int rank = 0, world = 1;
    // profiling
    bool prof = false;
    double prof_ms[MPST_T_COUNT] = {0};
    int64_t prof_n[MPST_T_COUNT] = {0};
    double prof_work[MPST_T_COUNT] = {0};
    cudaEvent_t tm0 = nullptr, tm1 = nullptr;
    int64_t launches = 0;
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> pending;
    std::vector<cudaEvent_t> evpool;
};

#define CUDA_TRY(ctx, expr)                                                                       \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                      \
            return MPST_E_CUDA;                                                                   \
        }                                                                                         \
    } while (0)

#define TRY(expr)                                                                                 \
    do {                                                                                          \
        int _r = (expr);                                                                          \
        if (_r != MPST_OK) return _r;                                                             \
    } while (0)

static inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// ---- kernels (defined in the .cu files) -------------------------------------------------------
int launch_encode(mpst_ctx* c, int basis, int d, const double* x, int64_t n, double* out, int64_t ldo);
int launch_encode_site(mpst_ctx* c, int site, const double* x, int64_t n, double* out, int64_t ldo);
int launch_permute_core(mpst_ctx* c, const double* src, double* dst, int d, int chi_l, int chi_r,
                        int C, long ss, long sa, long sb, long sc, long ds, long da, long db, long dc);
int launch_krao_gemm_rows(mpst_ctx* c, const double* x, const double* E, const double* W, double* out,
                          int64_t row_begin, int64_t row_end, int d, int chi, int n_out, int64_t ldw,
                          int64_t ldo);
int launch_krao_gemm(mpst_ctx* c, const double* x, const double* E, const double* W, double* out,
                     int64_t N, int d, int chi, int n_out, int64_t ldw, int64_t ldo);
int launch_flatten(mpst_ctx* c, CoreView l, CoreView r, int d, int chi_l, int chi_m, int chi_r, int C,
                   double* B);
int launch_sumsq(mpst_ctx* c, const double* v, int64_t n, double* out_dev);
int launch_axpy(mpst_ctx* c, double* B, const double* G, int64_t n, const double* gnorm2_dev,
                double eta, int tsgo);
int launch_scale_dev(mpst_ctx* c, double* v, int64_t n, const double* norm2_dev);
int impute_batch(mpst_ctx* c, int class_idx, const double* X, const uint8_t* missing, int64_t n,
                 int method, const double* xgrid, int G, const double* uniforms, int64_t uniforms_per_instance,
                 int n_traj, const mpst_impute_opts* io, double* out, double* err_out);

// stream-K schedule cache (api.cu)
SegTable* segtable_find(mpst_ctx* c, const std::vector<int64_t>& key);
int segtable_add(mpst_ctx* c, const std::vector<int64_t>& key, const std::vector<GradSeg>& segs,
                 const std::vector<int>& cta_ptr, const std::vector<int>& tile_slot, SegTable** out);

// profiling helpers
void prof_begin(mpst_ctx* c, int kind);
void prof_end(mpst_ctx* c, int kind);
int ensure_buf(mpst_ctx* c, double** p, size_t* cap, size_t need);
