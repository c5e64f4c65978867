# B200Backend.jl -- the Julia side of libmpstime_b200.so (include/mpstime_b200.h).
#
# Drop into MPSTime.jl as src/B200Backend/B200Backend.jl and `include` it at the end of src/MPSTime.jl (after
# Training/, Imputation/ and MLJIntegration/ are loaded).  Nothing in MPSOptions / Options / TrainedMPS changes
# (their field layout is the JLD2 on-disk format, Structs/options.jl:11-39): the backend is selected with
#     ENV["MPSTIME_BACKEND"] = "b200"          (or B200Backend.enable!())
# next to the existing use_legacy_ITensor switch (Training/RealRealHighDimension.jl:556-560).
#
# Seams swapped (same signatures and return values as the reference):
#   fitMPS(W::MPS, train::EncodedTimeSeriesSet, test::EncodedTimeSeriesSet, opts::Options)   RealRealHighDimension.jl:587
#   classify(mps::TrainedMPS, test::EncodedTimeSeriesSet)                                      summary.jl:116
#   get_predictions(imp, class, instance, missing_sites, method; ...)                          Imputation/imputation.jl:264
#   get_predictions_batch(imp, class, instances, missing_sites_list, method; ...)              new, used by eval_loss
#   eval_loss(::ImputationLoss, ...)                                                           hyperopt_utils.jl:174-231
#   MMI.fit / MMI.predict for MPSClassifier                                                    MLJ_integration.jl:32-62
#
# Julia is not installed in the build image of the library, so this file has not been executed there; its `ccall`
# signatures are checked mechanically against the ctypes binding that the GPU parity tests drive
# (tests/test_host_cpu.py::test_julia_shim_signatures_match_the_abi).
module B200Backend

using ITensors, ITensorMPS, Libdl, Random, Statistics, StatsBase
import ..MPSTime: Options, MPSOptions, TrainedMPS, EncodedTimeSeriesSet, PState, ImputationProblem, KLDLoss, MSELoss,
                  find_label, get_siteinds, transform_train_data, transform_test_data, invert_test_transform,
                  init_imputation_problem, ImputationLoss, MPSClassifier, MMI

const LIB = Ref{Ptr{Cvoid}}(C_NULL)
lib() = (LIB[] == C_NULL && (LIB[] = dlopen(get(ENV, "MPSTIME_B200_LIB", "libmpstime_b200.so"))); LIB[])
sym(s::Symbol) = dlsym(lib(), s)

const ENABLED = Ref(get(ENV, "MPSTIME_BACKEND", "") == "b200")
enable!(on::Bool=true) = (ENABLED[] = on)
enabled() = ENABLED[]

# ---- mirrors of the C structs --------------------------------------------------------------------------------------
struct TrainOpts            # mpst_train_opts
    loss_kind::Int32; opt_kind::Int32; train_sep::Int32; update_iters::Int32
    rescale_before::Int32; rescale_after::Int32; chi_max::Int32; reserved::Int32
    eta::Float64; cutoff::Float64
end

struct ImputeOpts           # mpst_impute_opts
    backwards::Int32; get_err::Int32; max_trials::Int32; reserved::Int32
    rejection_threshold::Float64; max_jump::Float64
end

const BASIS_IDS = Dict("Legendre" => 0, "Legendre_Norm" => 1, "Fourier" => 2, "Stoudenmire" => 3, "Sahand" => 4, "Uniform" => 5)
const METHOD_IDS = Dict(:median => 0, :mean => 1, :mode => 2, :ITS => 3)
const PRECOMPUTED = 100

# ---- context -------------------------------------------------------------------------------------------------------
mutable struct Ctx
    h::Ptr{Cvoid}
    function Ctx(dev::Integer=0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall(sym(:mpst_create), Cint, (Ref{Ptr{Cvoid}}, Cint), r, dev)
        rc == 0 || error("mpst_create failed ($rc): no usable sm_100 GPU at index $dev; this backend has no CPU fallback")
        c = new(r[])
        finalizer(x -> ccall(sym(:mpst_destroy), Cint, (Ptr{Cvoid},), x.h), c)
        return c
    end
end
const CTX = Dict{Int,Ctx}()
context(dev::Integer=parse(Int, get(ENV, "LOCAL_RANK", "0"))) = get!(() -> Ctx(dev), CTX, dev)

last_error(c::Ctx) = unsafe_string(ccall(sym(:mpst_last_error), Cstring, (Ptr{Cvoid},), c.h))
chk(c::Ctx, rc) = rc == 0 || error("libmpstime_b200 error $rc: " * last_error(c))
version() = ccall(sym(:mpst_version), Cint, ())

# ---- multi-GPU: one Julia worker per GPU; rank 0 creates the id and sends the 128 bytes to the others ---------------
function comm_unique_id()
    id = zeros(UInt8, 128)
    rc = ccall(sym(:mpst_comm_unique_id), Cint, (Ptr{Cvoid},), id)
    rc == 0 || error("mpst_comm_unique_id failed ($rc)")
    return id
end
comm_init(c::Ctx, id::Vector{UInt8}, rank::Integer, world::Integer) =
    chk(c, ccall(sym(:mpst_comm_init), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint), c.h, id, rank, world))

# ---- K1 ------------------------------------------------------------------------------------------------------------
function encode(c::Ctx, basis_id::Integer, d::Integer, x::Vector{Float64})
    cplx = basis_id in (2, 3, 4)
    out = Matrix{Float64}(undef, (cplx ? 2 : 1) * d, length(x))
    chk(c, ccall(sym(:mpst_encode), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Int64, Ptr{Float64}),
                 c.h, basis_id, d, x, length(x), out))
    return cplx ? reinterpret(ComplexF64, out) : out
end

# ---- per-site coefficient tables of the data-driven / time-dependent real encodings (encode_table.cu) ----------------
const TABLE_LEGENDRE_PROJ = 101; const TABLE_SAHAND_LEGENDRE = 102; const TABLE_SPLIT = 103

set_encoding_table(c::Ctx, kind::Integer, n_sites::Integer, d::Integer, ip::Matrix{Int32}, dp::Matrix{Float64}) =   # columns = sites
    chk(c, ccall(sym(:mpst_set_encoding_table), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Int32}, Int64, Ptr{Float64}, Int64),
                 c.h, kind, n_sites, d, ip, size(ip, 1), dp, size(dp, 1)))

function encode_site(c::Ctx, site0::Integer, d::Integer, x::Vector{Float64})
    out = Matrix{Float64}(undef, d, length(x))
    chk(c, ccall(sym(:mpst_encode_site), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Int64, Ptr{Float64}), c.h, site0, x, length(x), out))
    return out
end

# `encoding_args` (what opts.encoding.init returned, encodings.jl:112-120) -> device table.  fitMPS below does not need
# this (it ships the reference's own PStates as phi); it lets classify / repeated fits skip the per-sample encode on the host.
function encoding_table(opts::Options, enc_args::AbstractVector, T::Integer)
    d = opts.d; name = opts.encoding.name
    if startswith(name, "Projected Legendre")
        ds = enc_args[1]                                                   # Vector (per site) of d orders
        L = maximum(maximum.(ds))
        ip = fill(Int32(-1), d + L + 1, T); dp = zeros(2, T)
        for t in 1:T
            ip[1:d, t] .= ds[t]
            for (k, l) in enumerate(ds[t]); ip[d+l+1, t] = k - 1; end
            dmax = maximum(ds[t])
            dp[1, t] = endswith(name, "Norm") && dmax > 0 ? 1 / sqrt(sqrt((2dmax + 1) / 2) * dmax) : 1.0
            dp[2, t] = maximum(ds[t])
        end
        return TABLE_LEGENDRE_PROJ, T, ip, dp
    elseif startswith(name, "Sahand-Legendre")
        td = opts.encoding.istimedependent
        kdes, minxs, scales, cVecs = td ? enc_args : ([enc_args[1]], [enc_args[2]], [enc_args[3]], [enc_args[4]])
        ns = length(cVecs); npts = 2048
        ip = zeros(Int32, 2, ns); dp = zeros(4 + d * d + npts + 2, ns)
        for t in 1:ns
            ip[1, t] = npts; dp[2, t] = 1.0; dp[4, t] = 1.0
            (isassigned(kdes, t) && !iszero(cVecs[t])) || continue
            ik = Main.MPSTime.KernelDensity.InterpKDE(kdes[t])
            ip[2, t] = 1
            dp[1, t] = first(kdes[t].x); dp[2, t] = step(kdes[t].x); dp[3, t] = minxs[t]; dp[4, t] = scales[t]
            dp[5:4+d*d, t] .= vec(permutedims(cVecs[t]))                   # row n = basis function, column i = power
            dp[5+d*d:end, t] .= collect(ik.itp.itp.itp.coefs)              # padded quadratic-B-spline coefficients c[0..npts+1]
        end
        return TABLE_SAHAND_LEGENDRE, ns, ip, dp
    else                                                                   # SplitBasis over a data-independent auxiliary basis
        aux_enc_args, (bins, aux_dim, aux_enc) = enc_args
        rows = eltype(bins) <: Number ? [bins] : bins
        nb = length(rows[1]) - 1
        ip = repeat(Int32[nb, aux_dim, BASIS_IDS[replace(aux_enc.name, "_No_Norm" => "")]], 1, length(rows))
        return TABLE_SPLIT, length(rows), ip, reduce(hcat, rows)
    end
end

# ---- model / training set ------------------------------------------------------------------------------------------
model_init(c::Ctx, T, C, d, chi_max, basis_id) =
    chk(c, ccall(sym(:mpst_model_init), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Cint), c.h, T, C, d, chi_max, basis_id))

# X: T x N, series are columns, class-sorted, already in the encoding range
train_load_x(c::Ctx, X::Matrix{Float64}, counts::Vector{Int64}, basis_id, d, chi_max; n_global=size(X, 2), counts_global=counts) =
    chk(c, ccall(sym(:mpst_train_load_x), Cint,
                 (Ptr{Cvoid}, Ptr{Float64}, Int64, Cint, Ptr{Int64}, Cint, Cint, Cint, Cint, Int64, Ptr{Int64}),
                 c.h, X, size(X, 2), size(X, 1), counts, length(counts), basis_id, d, chi_max, n_global, counts_global))

# phi: d x T x N (pstate[j][s] of sample i)
train_load_phi(c::Ctx, phi::Array{Float64,3}, counts::Vector{Int64}, chi_max; n_global=size(phi, 3), counts_global=counts) =
    chk(c, ccall(sym(:mpst_train_load_phi), Cint,
                 (Ptr{Cvoid}, Ptr{Float64}, Int64, Cint, Ptr{Int64}, Cint, Cint, Cint, Int64, Ptr{Int64}),
                 c.h, phi, size(phi, 3), size(phi, 2), counts, length(counts), size(phi, 1), chi_max, n_global, counts_global))

set_core(c::Ctx, site0::Integer, data::Vector{Float64}, chi_l, chi_r, has_label) =
    chk(c, ccall(sym(:mpst_set_core), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Cint, Cint), c.h, site0, data, chi_l, chi_r, has_label))

function get_core(c::Ctx, site0::Integer, d::Integer, C::Integer)
    cl = Ref{Int32}(0); cr = Ref{Int32}(0); lab = Ref{Int32}(0)
    chk(c, ccall(sym(:mpst_get_core_dims), Cint, (Ptr{Cvoid}, Cint, Ref{Int32}, Ref{Int32}, Ref{Int32}), c.h, site0, cl, cr, lab))
    buf = Vector{Float64}(undef, cl[] * d * cr[] * (lab[] == 1 ? C : 1))
    chk(c, ccall(sym(:mpst_get_core), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), c.h, site0, buf))
    return buf, Int(cl[]), Int(cr[]), lab[] == 1
end

build_env(c::Ctx, going_left::Bool) = chk(c, ccall(sym(:mpst_build_env), Cint, (Ptr{Cvoid}, Cint), c.h, going_left))

function bond_step(c::Ctx, lid0::Integer, going_left::Bool, o::TrainOpts)
    lo = Ref{Float64}(0); gn = Ref{Float64}(0); chi = Ref{Int32}(0)
    chk(c, ccall(sym(:mpst_bond_step), Cint, (Ptr{Cvoid}, Cint, Cint, Ref{TrainOpts}, Ref{Float64}, Ref{Float64}, Ref{Int32}),
                 c.h, lid0, going_left, Ref(o), lo, gn, chi))
    return lo[], gn[], Int(chi[])
end

function sweep(c::Ctx, o::TrainOpts, nsweeps::Integer, T::Integer)
    nb = 2 * (T - 1) * nsweeps
    loss = zeros(nb); gn = zeros(nb); chi = zeros(Int32, nb)
    chk(c, ccall(sym(:mpst_sweep), Cint, (Ptr{Cvoid}, Ref{TrainOpts}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
                 c.h, Ref(o), nsweeps, loss, gn, chi))
    return loss, gn, chi
end

function sweep_bonds(c::Ctx, o::TrainOpts, n_bonds::Integer; restart::Bool=false)
    loss = zeros(n_bonds); gn = zeros(n_bonds); chi = zeros(Int32, n_bonds)
    chk(c, ccall(sym(:mpst_sweep_bonds), Cint, (Ptr{Cvoid}, Ref{TrainOpts}, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
                 c.h, Ref(o), n_bonds, restart, loss, gn, chi))
    return loss, gn, chi
end

# ---- K7 ------------------------------------------------------------------------------------------------------------
function overlaps(c::Ctx, X_or_phi::Array{Float64}, n::Integer, C::Integer)
    yhat = Matrix{Float64}(undef, C, n); am = Vector{Int64}(undef, n)
    chk(c, ccall(sym(:mpst_overlaps), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int64}), c.h, X_or_phi, n, yhat, am))
    return yhat, am .+ 1
end

# MSE_loss_acc_conf (summary.jl:95-114) reduced on the device; X_or_phi === nothing: the resident training set
function eval_metrics(c::Ctx, X_or_phi::Union{Nothing,Array{Float64}}, n::Integer, label_idx0::Union{Nothing,Vector{Int64}}, C::Integer)
    sums = zeros(3); conf = zeros(Int64, C, C)
    chk(c, ccall(sym(:mpst_eval_metrics), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}),
                 c.h, isnothing(X_or_phi) ? C_NULL : X_or_phi, n, isnothing(label_idx0) ? C_NULL : label_idx0, sums, conf))
    ntot = sum(conf)
    return sums[1] / ntot, sums[2] / ntot, sums[3] / ntot, permutedims(conf)      # C row-major conf[true][pred]
end

# ---- K8 ------------------------------------------------------------------------------------------------------------
function impute_batch(c::Ctx, class_idx0::Integer, X::Matrix{Float64}, missing::Matrix{UInt8}, method::Symbol, xvals::Vector{Float64};
                      uniforms::Union{Nothing,Array{Float64}}=nothing, n_traj::Integer=1, max_jump::Float64=-1.0)
    T, n = size(X)
    out = Array{Float64,3}(undef, T, method == :ITS ? n_traj : 1, n)
    chk(c, ccall(sym(:mpst_impute_batch), Cint,
                 (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{UInt8}, Int64, Cint, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Float64, Ptr{Float64}),
                 c.h, class_idx0, X, missing, n, METHOD_IDS[method], xvals, length(xvals),
                 isnothing(uniforms) ? C_NULL : uniforms, n_traj, max_jump, out))
    return out
end

function impute_batch_ex(c::Ctx, class_idx0::Integer, X::Matrix{Float64}, missing::Matrix{UInt8}, method::Symbol, xvals::Vector{Float64},
                         io::ImputeOpts; uniforms::Union{Nothing,Array{Float64}}=nothing, n_traj::Integer=1)
    T, n = size(X)
    nt = method == :ITS ? n_traj : 1
    out = Array{Float64,3}(undef, T, nt, n); err = zeros(T, nt, n)
    per = isnothing(uniforms) ? 0 : div(length(uniforms), max(n, 1))
    chk(c, ccall(sym(:mpst_impute_batch_ex), Cint,
                 (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{UInt8}, Int64, Cint, Ptr{Float64}, Cint, Ptr{Float64}, Int64, Cint,
                  Ref{ImputeOpts}, Ptr{Float64}, Ptr{Float64}),
                 c.h, class_idx0, X, missing, n, METHOD_IDS[method], xvals, length(xvals),
                 isnothing(uniforms) ? C_NULL : uniforms, per, nt, Ref(io), out, err))
    return out, err
end

# ---- test / benchmark entries (kept bound so that the Julia test-suite can teacher-force single kernels) ------------
function bond_loss_grad(c::Ctx, B::Matrix{Float64}, L::Matrix{Float64}, R::Matrix{Float64}, xl::Matrix{Float64}, xr::Matrix{Float64},
                        counts::Vector{Int64}, loss_kind::Integer, train_sep::Bool)
    N = size(xl, 2); d = size(xl, 1)
    G = similar(B); lo = Ref{Float64}(0)
    chk(c, ccall(sym(:mpst_bond_loss_grad), Cint,
                 (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Cint, Cint, Cint,
                  Ptr{Int64}, Cint, Cint, Cint, Ref{Float64}, Ptr{Float64}, Ptr{Float64}),
                 c.h, B, L, R, xl, xr, N, d, size(L, 1), size(R, 1), counts, length(counts), loss_kind, train_sep, lo, G, C_NULL))
    return lo[], G
end

function bond_split(c::Ctx, B::Matrix{Float64}, d, chi_l, chi_r, going_left::Bool, chi_max, cutoff)
    C = size(B, 2); kmax = max(1, min(chi_max, d * (going_left ? chi_r : chi_l)))
    cl = zeros(chi_l * d * kmax * (going_left ? C : 1)); cr = zeros(kmax * d * chi_r * (going_left ? 1 : C)); sig = zeros(kmax)
    chi = Ref{Int32}(0)
    chk(c, ccall(sym(:mpst_bond_split), Cint,
                 (Ptr{Cvoid}, Ptr{Float64}, Cint, Cint, Cint, Cint, Cint, Cint, Float64, Ref{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                 c.h, B, d, chi_l, chi_r, C, going_left, chi_max, cutoff, chi, cl, cr, sig))
    return cl, cr, sig[1:chi[]], Int(chi[])
end

profile_enable(c::Ctx, on::Bool) = chk(c, ccall(sym(:mpst_profile_enable), Cint, (Ptr{Cvoid}, Cint), c.h, on))
profile_reset(c::Ctx) = chk(c, ccall(sym(:mpst_profile_reset), Cint, (Ptr{Cvoid},), c.h))
function profile_get(c::Ctx)
    ms = zeros(10); n = zeros(Int64, 10); work = zeros(10)
    chk(c, ccall(sym(:mpst_profile_get), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}), c.h, ms, n, work))
    return ms, n, work
end
timer_start(c::Ctx) = chk(c, ccall(sym(:mpst_timer_start), Cint, (Ptr{Cvoid},), c.h))
function timer_stop(c::Ctx)
    ms = Ref{Float64}(0)
    chk(c, ccall(sym(:mpst_timer_stop), Cint, (Ptr{Cvoid}, Ref{Float64}), c.h, ms))
    return ms[]
end
launch_count(c::Ctx) = ccall(sym(:mpst_launch_count), Int64, (Ptr{Cvoid},), c.h)
debug_set(c::Ctx, name::String, value::Integer) = chk(c, ccall(sym(:mpst_debug_set), Cint, (Ptr{Cvoid}, Cstring, Cint), c.h, name, value))
debug_get(c::Ctx, name::String) = ccall(sym(:mpst_debug_get), Int64, (Ptr{Cvoid}, Cstring), c.h, name)

# =====================================================================================================================
# ITensor <-> wire layout.  Wire: (left link, site, right link[, label]) column-major, missing boundary links = size 1.
# =====================================================================================================================
function dense_core(W::MPS, j::Integer, sites, label_idx)
    l = j > 1 ? commonind(W[j-1], W[j]) : nothing
    r = j < length(W) ? commonind(W[j], W[j+1]) : nothing
    lab = !isnothing(label_idx) && hasind(W[j], label_idx)
    is = Index[]
    isnothing(l) || push!(is, l)
    push!(is, sites[j])
    isnothing(r) || push!(is, r)
    lab && push!(is, label_idx)
    A = Array{Float64}(array(W[j], is...))
    return vec(A), isnothing(l) ? 1 : dim(l), isnothing(r) ? 1 : dim(r), lab
end

function upload_mps!(c::Ctx, W::MPS)
    sites = get_siteinds(W)
    _, label_idx = find_label(W)
    for j in eachindex(W)
        data, cl, cr, lab = dense_core(W, j, sites, label_idx)
        set_core(c, j - 1, data, cl, cr, lab)
    end
    return sites, label_idx
end

function download_mps!(c::Ctx, W::MPS, sites, label_idx, d::Integer, C::Integer)
    T = length(W)
    links = Vector{Index}(undef, T + 1)
    links[1] = Index(1, "Link,l=0")
    for j in 1:T
        buf, cl, cr, lab = get_core(c, j - 1, d, C)
        links[j+1] = Index(cr, "Link,l=$j")
        is = lab ? (links[j], sites[j], links[j+1], label_idx) : (links[j], sites[j], links[j+1])
        A = itensor(reshape(buf, dim.(is)...), is...)
        j == 1 && (A *= onehot(links[1] => 1))                 # drop the size-1 boundary links
        j == T && (A *= onehot(links[T+1] => 1))
        W[j] = A
    end
    return W
end

train_opts(opts::Options) = TrainOpts(opts.loss_grad isa KLDLoss ? 0 : 1, uppercase(opts.bbopt.fl) == "TSGO" ? 0 : 1,
                                      opts.train_classes_separately, opts.update_iters, opts.rescale[1], opts.rescale[2],
                                      opts.chi_max, 0, opts.eta, opts.cutoff)

function check_supported(opts::Options)
    opts.loss_grad isa KLDLoss || opts.loss_grad isa MSELoss || throw(ArgumentError("B200 backend: loss_grad must be KLD or MSE"))
    uppercase(opts.bbopt.fl) in ("TSGO", "GD") || throw(ArgumentError("B200 backend: bbopt must be TSGO or GD (loss_functions.jl:166-170)"))
    opts.dtype == Float64 || throw(ArgumentError("B200 backend: the array training path is Float64-only (loss_functions.jl:343)"))
    opts.encoding.iscomplex && throw(ArgumentError("B200 backend: complex encodings cannot be trained on the array path"))
end

# phi (d x T x N) from the reference's own PStates: works for every real encoding, data-driven or not
function pack_phi(ts::Vector{PState}, T::Integer, d::Integer)
    phi = Array{Float64}(undef, d, T, length(ts))
    for (i, ps) in enumerate(ts), j in 1:T
        phi[:, j, i] .= ps.pstate[j]
    end
    return phi
end

label_indices0(ts::Vector{PState}) = Int64[Int64(ps.label_index) - 1 for ps in ts]

# ---- fitMPS(W, train, test, opts): the body of RealRealHighDimension.jl:587-890 ------------------------------------------
function fitMPS(W::MPS, training_states_meta::EncodedTimeSeriesSet, testing_states_meta::EncodedTimeSeriesSet, opts::Options; test_run=false)
    check_supported(opts)
    c = context()
    ts = training_states_meta.timeseries
    T = length(W); d = opts.d
    counts = Int64.(training_states_meta.class_distribution); C = length(counts)
    train_load_phi(c, pack_phi(ts, T, d), counts, opts.chi_max)
    sites, label_idx = upload_mps!(c, W)
    has_test = !isempty(testing_states_meta)
    phi_test = has_test ? pack_phi(testing_states_meta.timeseries, T, d) : nothing
    lab_test = has_test ? label_indices0(testing_states_meta.timeseries) : nothing

    training_information = Dict("train_loss" => Float64[], "train_acc" => Float64[], "test_loss" => Float64[],
                                "time_taken" => Float64[], "train_KL_div" => Float64[])
    if has_test
        training_information["test_acc"] = Float64[]; training_information["test_KL_div"] = Float64[]
        training_information["test_conf"] = Matrix{Int}[]
    end
    function log!(elapsed)                                                # :657-689, :813-845, :854-885
        opts.log_level > 0 || return nothing
        mse, kld, acc, _ = eval_metrics(c, nothing, 0, nothing, C)
        push!(training_information["train_loss"], mse); push!(training_information["train_acc"], acc)
        push!(training_information["train_KL_div"], kld); push!(training_information["time_taken"], elapsed)
        if has_test
            mse_t, kld_t, acc_t, conf = eval_metrics(c, phi_test, length(lab_test), lab_test, C)
            push!(training_information["test_loss"], mse_t); push!(training_information["test_acc"], acc_t)
            push!(training_information["test_KL_div"], kld_t); push!(training_information["test_conf"], conf)
        end
        opts.verbosity > -1 && println("Training KL Div. $kld | Training acc. $acc.")
        return acc
    end
    log!(0.0)
    o = train_opts(opts)
    nb = 2 * (T - 1)
    for itS in 1:opts.nsweeps                                              # :726
        t0 = time()
        sweep_bonds(c, o, nb; restart=(itS == 1))                          # backward :731 + forward :776 half-sweeps
        acc = log!(time() - t0)
        opts.exit_early && acc == 1.0 && break                             # :847-849
    end
    download_mps!(c, W, sites, label_idx, d, C)
    normalize!(W)                                                          # :852
    upload_mps!(c, W)
    log!(NaN)
    return TrainedMPS(W, MPSOptions(opts), training_states_meta), training_information, testing_states_meta
end

# ---- classify(mps, test_states) (summary.jl:116-136) -----------------------------------------------------------------------
function classify(mps::TrainedMPS, test_states::EncodedTimeSeriesSet)
    c = context()
    W = mps.mps; T = length(W)
    pss = test_states.timeseries
    d = length(pss[1].pstate[1])
    _, label_idx = find_label(W)
    C = dim(label_idx)
    model_init(c, T, C, d, maxlinkdim(W), PRECOMPUTED)
    upload_mps!(c, W)
    _, am = overlaps(c, pack_phi(pss, T, d), length(pss), C)
    labels = sort(unique([ps.label for ps in mps.train_data.timeseries]))
    return Int64[labels[a] for a in am]
end

# ---- get_predictions (imputation.jl:264-410), batched over instances ----------------------------------------------------------
# instances: indices into the test series of `class`; missing_sites_list[k]: 1-based sites to impute in instance k.
function get_predictions_batch(imp::ImputationProblem, class, instances::AbstractVector{<:Integer}, missing_sites_list, method::Symbol=:median;
                               impute_order::Symbol=:forwards, invert_transform::Bool=true, rseed::Integer=1, num_trajectories::Integer=1,
                               max_jump=nothing, get_wmad::Bool=false, get_std::Bool=false, rejection_threshold=:none, max_trials::Integer=10)
    impute_order in (:forwards, :backwards) || throw(ArgumentError("impute_order must be either \":forwards\" or \":backwards\""))
    haskey(METHOD_IDS, method) || error("Invalid method. Choose :mean, :mode, :median or :ITS")
    c = context()
    mps = imp.mpss[imp.class_map[class]]                                   # label-free class MPS (utils.jl:356-370)
    T = length(mps); d = imp.opts.d
    if imp.opts.encoding.isdatadriven || imp.opts.encoding.istimedependent
        # K8 table mode: per-site grid states (`xvals_enc[site]`, imputation.jl:92-100) and per-site encoding of the known /
        # imputed values are evaluated on the device from the tables built out of the reference's own `enc_args`
        kind, ns, ip, dp = encoding_table(imp.opts, imp.enc_args, T)
        set_encoding_table(c, kind, ns, d, ip, dp)
        model_init(c, T, 1, d, maxlinkdim(mps), kind)
    else
        model_init(c, T, 1, d, maxlinkdim(mps), BASIS_IDS[replace(imp.opts.encoding.name, "_No_Norm" => "")])
    end
    sites = get_siteinds(mps)
    for j in 1:T
        data, cl, cr, _ = dense_core(mps, j, sites, nothing)
        set_core(c, j - 1, data, cl, cr, false)
    end
    cl_inds = (1:length(imp.y_test))[imp.y_test .== class]
    n = length(instances)
    X_train_scaled, norms = transform_train_data(imp.X_train; opts=imp.opts)          # hoisted out of the per-instance loop (:287)
    raw = imp.X_test[cl_inds[instances], :]
    X = Matrix{Float64}(undef, T, n); mask = zeros(UInt8, T, n)
    oobs = Vector{Any}(undef, n)
    fill_value = mean(imp.X_train[:])
    for k in 1:n
        ts = copy(raw[k, :]); ts[missing_sites_list[k]] .= fill_value                 # :290
        X[:, k], oobs[k] = transform_test_data(ts, norms; opts=imp.opts)              # :291
        mask[missing_sites_list[k], k] .= 1
    end
    Kmax = maximum(length.(missing_sites_list))
    rejecting = method == :ITS && rejection_threshold != :none
    uniforms = nothing
    if method == :ITS                                                      # the reference's stream: one MersenneTwister per call (:324)
        per = rejecting ? num_trajectories * Kmax * max_trials : num_trajectories * Kmax
        uniforms = Matrix{Float64}(undef, per, n)
        for k in 1:n
            rng = MersenneTwister(rseed); uniforms[:, k] .= rand(rng, per)
        end
    end
    io = ImputeOpts(impute_order == :backwards, (method == :median && get_wmad) || (method == :mean && get_std), max_trials, 0,
                    rejecting ? Float64(rejection_threshold) : -1.0, isnothing(max_jump) ? -1.0 : Float64(max_jump))
    out, err = impute_batch_ex(c, 0, X, mask, method, imp.x_guess_range.xvals, io; uniforms=uniforms, n_traj=num_trajectories)
    tss = [[out[:, tr, k] for tr in 1:size(out, 2)] for k in 1:n]
    errs = [method in (:median, :mean) ? [err[:, tr, k] for tr in 1:size(out, 2)] : [nothing for _ in 1:size(out, 2)] for k in 1:n]
    targets = Vector{Vector{Float64}}(undef, n)
    for k in 1:n
        if invert_transform                                                # :337-394
            for tr in eachindex(tss[k])
                if !isnothing(errs[k][tr])
                    errs[k][tr] .+= tss[k][tr]
                end
                tss[k][tr] = invert_test_transform(tss[k][tr], oobs[k], norms; opts=imp.opts)
                if !isnothing(errs[k][tr])
                    e = try invert_test_transform(errs[k][tr], oobs[k], norms; opts=imp.opts) catch; fill(NaN, T) end
                    errs[k][tr] = e .- tss[k][tr]
                end
            end
            targets[k] = raw[k, :]
        else
            targets[k], _ = transform_test_data(raw[k, :], norms; opts=imp.opts)
        end
    end
    return tss, errs, targets
end

function get_predictions(imp::ImputationProblem, class, instance::Integer, missing_sites::Vector{<:Integer}, method::Symbol=:median; kwargs...)
    tss, errs, targets = get_predictions_batch(imp, class, [instance], [missing_sites], method; kwargs...)
    return tss[1], errs[1], targets[1]
end

# ---- eval_loss(::ImputationLoss) (hyperopt_utils.jl:174-231): one batched call per window instead of numval x windows calls
function eval_loss(::ImputationLoss, mps::TrainedMPS, X_val::AbstractMatrix, y_val::AbstractVector, windows=nothing; p_fold=nothing,
                   distribute::Bool=false, method::Symbol=:median)
    imp = init_imputation_problem(mps, X_val, y_val, verbosity=-5)
    cmap = countmap(y_val)
    loss_by_window = zeros(length(windows))
    numval = size(X_val, 1)
    for (iw, impute_sites) in enumerate(windows)
        for (cls, cnt) in pairs(cmap)
            tss, _, targets = get_predictions_batch(imp, cls, collect(1:cnt), [impute_sites for _ in 1:cnt], method)
            for k in 1:cnt
                loss_by_window[iw] += mean(abs.(tss[k][1][impute_sites] .- targets[k][impute_sites]))     # MAE, metrics.jl:2-20
            end
        end
    end
    return loss_by_window ./ numval
end

# ---- MLJ: MPSClassifier keeps its fields and traits (MLJ_integration.jl:2-30, 65-70); fit / predict route to the working
#      fitMPS(X, y, opts) / classify(mps, X) pair, which reach the seams above when the backend is enabled --------------------------
function mlj_fit(m::MPSClassifier, verbosity::Int, X, y, decode)
    opts = MPSOptions(m; verbosity=verbosity)
    mps, info, _ = Main.MPSTime.fitMPS(permutedims(X), y, opts)            # X arrives transposed from MMI.reformat
    return ((decode, mps), nothing, (info = info,))
end
mlj_predict(::MPSClassifier, fitresult, Xnew) = fitresult[1].(Main.MPSTime.classify(fitresult[2], permutedims(Xnew)))

end # module
