# gab1pde_dropin.jl — drop-in Julia front end for libgab1pde.so (B200 batched solver).
#
# HOW TO USE.  This is a plain top-level file, NOT a module, on purpose: the reference's API is a set of functions
# that its scripts bring into `Main` with `include(...)`, and in Julia a name that already has a binding in `Main` is
# not replaced by `using SomeModule` (the existing binding wins, with a warning).  What does replace a method is a
# second definition with the same positional signature in the same module.  So:
#
#     include("basepdesolver.jl"); include("basepdesolver_rect.jl"); include("pulsechase_solver.jl")
#     include("get_param_posteriors.jl")           # run_ensemble, Diffs, kvals, ...
#     include("sapdesolver.jl")                    # or sapdesolver_memb-SFK.jl — defines Co, R, dr, tf, fbatch_*_mt
#     include("/path/to/julia/gab1pde_dropin.jl")  # LAST: overwrites the methods below with the GPU ones
#
# Every function below repeats the reference's positional signature and keyword list verbatim (tests/test_julia_surface.py
# checks this against the reference source), so each definition OVERWRITES the reference's CPU method; callers such as
# run_base_model.jl:83,91 or `gsa(fbatch_concs_mt, eFAST(), pbounds; samples=1000, batch=true)` (GSA_concs.jl:81) then reach
# the GPU without any edit.  Keyword defaults that the reference takes from file-level globals (`Co=Co, D=Diffs,
# kvals=kvals, R=R, dr=dr, tf=tf`, sapdesolver.jl:288-460) are written the same way here and are evaluated in `Main` at
# call time, exactly as in the reference.  The file may also be included on its own (none of the reference's files
# loaded): then the functions are simply defined, and the globals must be supplied as keywords.
#
# The GSA wrappers (pmap_fun_*, fbatch_*) exist twice in the reference with identical signatures: sapdesolver.jl binds
# them to `sapdesolver`, sapdesolver_memb-SFK.jl:288-474 re-binds them to `sapdesolver_membSFK`.  Which one is wanted
# is the switch `GAB1PDE.MEMBSFK_WRAPPERS[]`; it defaults to "whichever of the two files was included", detected by
# `isdefined(Main, :sapdesolver_membSFK) && !isdefined(Main, :sapdesolver)`, and can be set by hand.
#
# All arithmetic happens in the shared library (hand-written sm_100a CUDA, include/gab1pde.h).  This file marshals
# arrays and evaluates the two inputs whose last bits depend on Julia itself: the default `dt` (Base `sum` is a SIMD
# reduction) and the grid `collect(0.0:dr:R)`.
#
# NOTE: no Julia runtime exists in the image this library is developed in (nor on its GPU boxes), so this file has
# never been executed; tests/test_julia_surface.py checks its signatures, keyword names and default literals against
# the reference source mechanically, and the Python twin (host.py) is what the test-suite executes.

module GAB1PDE      # helpers only — nothing here has a name the reference uses, and nothing is exported

const LIB = get(ENV, "GAB1PDE_LIB", joinpath(@__DIR__, "..", "myers-furcht-et-al_gab1-shp2-pde-model_b200", "libgab1pde.so"))
"true: pmap_fun_* / fbatch_* solve with sapdesolver_membSFK (sapdesolver_memb-SFK.jl:288-474); false: with sapdesolver"
const MEMBSFK_WRAPPERS = Ref(isdefined(Main, :sapdesolver_membSFK) && !isdefined(Main, :sapdesolver))
"safety cap of the `while error > tol` loop (sapdesolver_memb-SFK.jl:177 has none and hangs); flagged by ST_ITER_CAP"
const ITER_CAP = Ref(2000)
"number of GPUs a batch is sharded over inside one call; 0 = all visible"
const N_DEVICES = Ref(0)
"true: the GSA entry points (fbatch_*_mt, pmap_fun_*), whose reference signatures leave no room for a keyword, certify their
batches (gab1_solve_batch_certified: ill-conditioned sets are re-solved with the strict kernels; about 2.3x the cost)"
const CERTIFY = Ref(false)

# struct gab1_opts (include/gab1pde.h) — field order and types must match exactly
struct Opts
    abi_version::Int32; geometry::Int32; sfk_mode::Int32; bc_loop::Int32; save_rule::Int32; pg1tot_form::Int32
    out_mode::Int32; matrix_mask::UInt32; maxiters::Int32; Nr::Int32; Nts::Int32; arith::Int32
    tol::Float64; R::Float64; dr::Float64; tf::Float64; dt_save::Float64; t_prechase::Float64
    pct_mul::Float64; pct_div::Float64; n_devices::Int32; reserved::Int32; device_ids::Ptr{Int32}
end

const OUT_FINAL4, OUT_FULL, OUT_SIX, OUT_PCT_BOUND, OUT_FINAL_STATE = Int32(0), Int32(1), Int32(2), Int32(3), Int32(4)
const ST_NAN, ST_ITER_CAP, ST_SHORT, ST_OVERFLOW, ST_THROW = 1, 2, 4, 8, 16
const MATRICES = (:iSFK, :aSFK, :GRB2, :GAB1, :SHP2, :G2G1, :G2PG1, :G2PG1S, :PG1, :PG1S, :PG1tot, :PG1Stot)
const VECTORS = (:pE, :mE, :mES, :mESmES, :E, :EG2, :EG2G1, :EG2PG1, :EG2PG1S, :EGFR_SHP2, :t_out)
const MASK_FITTING = UInt32((1 << 1) | (1 << 9) | (1 << 7))

make_opts(; R, dr, tf, Nts, dt_save=tf / Nts, maxiters, tol, geometry=0, sfk_mode=0, bc_loop=0, save_rule=0,
          pg1tot_form=0, out_mode=OUT_FULL, matrix_mask=0x0fff, t_prechase=-1.0, pct_mul=1.0, pct_div=1.0,
          n_devices=N_DEVICES[]) =
    Opts(1, geometry, sfk_mode, bc_loop, save_rule, pg1tot_form, out_mode, matrix_mask, maxiters,
         Int32(ceil(R / dr)), Nts, 0, tol, R, dr, tf, dt_save, t_prechase, pct_mul, pct_div, n_devices, 0, C_NULL)

out_doubles(o::Opts) = ccall((:gab1_out_doubles_per_set, LIB), Int64, (Ref{Opts},), o)
last_error() = unsafe_string(ccall((:gab1_last_error, LIB), Cstring, ()))

# dt exactly as the reference's keyword default (basepdesolver.jl:30), evaluated by Julia
default_dt(D, k, dr) = 1.0 / (2.0 * (maximum(D) / (dr .^ 2) + sum(k) / 4)) * 0.99

"Pinned, device-mapped host matrix (gab1_host_alloc): the kernels write it while the time loop runs, so a 2.5 GB ensemble
result needs no staged copy.  Freed by a finalizer.  Small results use ordinary arrays."
function result_matrix(n::Integer, S::Integer)
    bytes = 8 * n * S
    bytes < (32 << 20) && return zeros(Float64, n, S)
    p = ccall((:gab1_host_alloc_near, LIB), Ptr{Float64}, (Csize_t, Int32), bytes, 0)
    p == C_NULL && return zeros(Float64, n, S)
    a = unsafe_wrap(Array, p, (Int(n), Int(S)); own=false)
    finalizer(x -> ccall((:gab1_host_free, LIB), Cvoid, (Ptr{Cvoid},), pointer(x)), a)
    a
end

struct Batch
    o::Opts; out::Matrix{Float64}; status::Vector{Int32}; n_saved::Vector{Int32}
    n_steps::Vector{Int64}; n_bc_iters::Vector{Int64}; r::Vector{Float64}; dt::Vector{Float64}
    resolved_strict::Vector{Int}       # certify=true: the sets (1-based) that were re-solved with the strict kernels
end

"Run S parameter sets (rows of Dmat S×7 and kmat S×17). Co is a 5-vector shared by all sets or S×5.
certify=true goes through gab1_solve_batch_certified: every set whose final state responds to a one-ulp change of Co (two
extra fast solves) is re-solved with the strict kernels, so that the 1e-9 contract holds for ill-conditioned sets too."
function solve(o::Opts, Co, Dmat, kmat, dt::Vector{Float64}, r::Vector{Float64}; certify::Bool=false)
    S = size(Dmat, 1)
    length(r) == o.Nr + 1 || throw(BoundsError(r, o.Nr + 1))           # the reference indexes r[Nr+1]
    Dt = permutedims(Float64.(Dmat)); kt = permutedims(Float64.(kmat))   # row-major for C
    Cot, stride = Co isa AbstractVector ? (Float64.(Co), 0) : (permutedims(Float64.(Co)), 5)
    n = out_doubles(o)
    out = result_matrix(n, S); status = zeros(Int32, S); n_saved = zeros(Int32, S)
    n_steps = zeros(Int64, S); n_bc = zeros(Int64, S)
    if certify
        resolved = zeros(Int32, S); n_res = Ref{Int64}(0)
        rc = ccall((:gab1_solve_batch_certified, LIB), Cint,
                   (Ref{Opts}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                    Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}, Float64, Ptr{Int32}, Ref{Int64}),
                   o, S, Cot, stride, Dt, kt, dt, r, out, status, n_saved, n_steps, n_bc, 0.0, resolved, n_res)
        rc == 0 || error("gab1_solve_batch_certified: " * last_error())
        return Batch(o, out, status, n_saved, n_steps, n_bc, r, dt, findall(!=(0), resolved))
    end
    rc = ccall((:gab1_solve_batch, LIB), Cint,
               (Ref{Opts}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}),
               o, S, Cot, stride, Dt, kt, dt, r, out, status, n_saved, n_steps, n_bc)
    rc == 0 || error("gab1_solve_batch: " * last_error())
    Batch(o, out, status, n_saved, n_steps, n_bc, r, dt, Int[])
end

function matrix(b::Batch, name::Symbol, j::Int)
    m = findfirst(==(name), MATRICES) - 1
    off = ccall((:gab1_full_matrix_offset, LIB), Int64, (Ref{Opts}, Int32), b.o, m)
    P, C = b.o.Nr + 1, b.o.Nts + 1
    reshape(b.out[off+1:off+P*C, j], P, C)           # column-major (Nr+1)×(Nts+1), as the reference builds it
end
function vector(b::Batch, name::Symbol, j::Int)
    v = findfirst(==(name), VECTORS) - 1
    off = ccall((:gab1_full_vector_offset, LIB), Int64, (Ref{Opts}, Int32), b.o, v)
    b.out[off+1:off+b.o.Nts+1, j]
end

# the reference's solvers are generic in the element type only so that ForwardDiff duals flow through; the GPU path
# takes Float64 (and, for pdesolver_fitting, duals through gab1_solve_tangent) and says so instead of guessing
f64(x, what) = eltype(x) <: AbstractFloat || eltype(x) <: Integer ? Float64.(x) :
    throw(ArgumentError("gab1pde_dropin: $what has element type $(eltype(x)); the GPU path takes real numbers " *
                        "(ForwardDiff duals are supported through pdesolver_fitting)"))

"Batched sibling of `pdesolver` (basepdesolver.jl:25-312): one row of Dmat / kmat per parameter set."
function pdesolver_batch(Co, Dmat, kmat; R=10.0, dr=0.1, tf=5.0, Nts=100,
                         dt=[default_dt(Dmat[j, :], kmat[j, :], dr) for j in axes(Dmat, 1)],
                         dt_save=tf / Nts, maxiters=100, tol=1.0e-6, r=collect(0.0:dr:R), certify=false, kw...)
    solve(make_opts(; R, dr, tf, Nts, dt_save, maxiters, tol, kw...), Co, Dmat, kmat, Float64.(dt), r; certify)
end

"Batched sibling of `sapdesolver` / `sapdesolver_membSFK` (sapdesolver.jl:55-280, sapdesolver_memb-SFK.jl:55-281)."
function sapdesolver_batch(Co, Dmat, kmat; R=10.0, dr=0.2, tf=5.0,
                           dt=[default_dt(Dmat[j, :], kmat[j, :], dr) for j in axes(Dmat, 1)],
                           maxiters=20, tol=1.0e-3, membSFK=false, out_mode=OUT_FINAL4, r=collect(0.0:dr:R),
                           iter_cap=ITER_CAP[], certify=false)
    # membSFK: the reference's `while error > tol` has no cap (sapdesolver_memb-SFK.jl:177) and spins forever on a fixed
    # point that never meets tol; the library stops such a step after `iter_cap` passes and flags the set (ST_ITER_CAP)
    o = make_opts(; R, dr, tf, Nts=1, maxiters=membSFK ? iter_cap : maxiters, tol, out_mode,
                  sfk_mode=membSFK ? 1 : 0, bc_loop=membSFK ? 1 : 0, pg1tot_form=membSFK ? 1 : 0)
    solve(o, Co, Dmat, kmat, Float64.(dt), r; certify)
end

function sol_tuple(b::Batch, j; extra=false, ncol=b.o.Nts + 1)
    mats = NamedTuple{MATRICES}(Tuple(matrix(b, n, j)[:, 1:ncol] for n in MATRICES))
    vecs = NamedTuple{VECTORS[1:9]}(Tuple(vector(b, n, j)[1:ncol] for n in VECTORS[1:9]))
    extra ? merge(mats, (EGFR_SHP2=vector(b, :EGFR_SHP2, j)[1:ncol],), vecs) : merge(mats, vecs)
end
check_throw(b) = (b.status[1] & ST_THROW != 0) && throw(InexactError(:Int64, Int64, NaN))      # Int64(ceil(tf/dt))
check_overflow(b) = (b.status[1] & ST_OVERFLOW != 0) && throw(BoundsError())                  # column Nts+2 of a fixed-size output

# which model a `model_fun` argument of run_ensemble names — by NAME, so that Main.pdesolver_rect (the reference's
# function object, overwritten or not) selects the rectangular kernels
function variant(model_fun)
    n = nameof(model_fun)
    n === :pdesolver ? NamedTuple() :
    n === :pdesolver_membSFK ? (sfk_mode=1,) :
    n === :pdesolver_rect ? (geometry=1, pg1tot_form=1) :
    n === :pulsechase_solver ? NamedTuple() :
    throw(ArgumentError("run_ensemble on the GPU knows pdesolver, pdesolver_membSFK, pdesolver_rect and pulsechase_solver, not $n"))
end

six(Co, Dm, km; R, dr, tf, tol, maxiters, membSFK=MEMBSFK_WRAPPERS[]) =
    sapdesolver_batch(Co, Dm, km; R, dr, tf, tol, maxiters, membSFK, out_mode=OUT_SIX, certify=CERTIFY[])
one(b) = (b.status[1] & ST_THROW != 0) ? throw(ArgumentError("reducing over an empty collection is not allowed")) : b.out[:, 1]
row(v) = reshape(Float64.(v), 1, :)

end # module GAB1PDE


## ================================================================ single solves — the reference's names and signatures
# Positional types and keyword lists are copied from the reference so that these definitions replace its methods.

"pdesolver (basepdesolver.jl:25-312) → (sol, r, t_out, dt)"
function pdesolver(Co::AbstractVector, D::AbstractVector, k::AbstractVector{T};
    R=10.0,
    dr=0.1,
    tf=5.0,
    Nts=100,
    dt=1.0/(2.0*(maximum(D)/(dr.^2) + sum(k)/4))*0.99,
    dt_save=tf/Nts,
    maxiters=100,
    tol=1.0e-6) where T
    G = GAB1PDE
    b = G.pdesolver_batch(G.f64(Co, "Co"), G.row(G.f64(D, "D")), G.row(G.f64(k, "k")); R, dr, tf, Nts, dt=[Float64(dt)], dt_save, maxiters, tol)
    G.check_throw(b); G.check_overflow(b)
    return G.sol_tuple(b, 1), b.r, G.vector(b, :t_out, 1), dt
end

"pdesolver_membSFK (basepdesolver.jl:350-636): aSFK diffusivity 1e-32"
function pdesolver_membSFK(Co::AbstractVector, D::AbstractVector, k::AbstractVector{T};
    R::Float64=10.0, dr::Float64=0.1, tf::Float64=5.0,
    Nts::Int64=100, dt::Float64=1.0/(2.0*(maximum(D)/dr.^2 + sum(k)/4))*0.99,
    dt_save = tf/Nts,
    maxiters=20,
    tol::Float64=1.0e-6) where T
    G = GAB1PDE
    b = G.pdesolver_batch(G.f64(Co, "Co"), G.row(G.f64(D, "D")), G.row(G.f64(k, "k")); R, dr, tf, Nts, dt=[dt], dt_save, maxiters, tol, sfk_mode=1)
    G.check_throw(b); G.check_overflow(b)
    return G.sol_tuple(b, 1), b.r, G.vector(b, :t_out, 1), dt
end

"pdesolver_rect (basepdesolver_rect.jl:23-294): planar Laplacian; outputs hold 1 + #snapshots columns"
function pdesolver_rect(Co::Vector{Float64}, D::Vector{Float64}, k::Vector{Float64};
    R::Float64=10.0, dr::Float64=0.1, tf::Float64=5.0,
    Nts::Int64=100, dt::Float64=1.0/(2.0*(maximum(D)/dr.^2 + sum(k)/4))*0.99,
    dt_save = tf/Nts,
    maxiters=20,
    tol::Float64=1.0e-6)
    G = GAB1PDE
    # the reference grows its outputs with hcat/push! (:250-279): one column per snapshot actually taken, no upper
    # bound; the library holds Nts+1 columns, so a schedule that would take more is an error here, not a truncation
    b = G.pdesolver_batch(Co, G.row(D), G.row(k); R, dr, tf, Nts, dt=[dt], dt_save, maxiters, tol, geometry=1, pg1tot_form=1)
    G.check_throw(b)
    (b.status[1] & G.ST_OVERFLOW != 0) && error("pdesolver_rect: more than Nts+1 snapshots are due (dt_save too small for the library's fixed-size output)")
    nc = Int(b.n_saved[1])
    return G.sol_tuple(b, 1; extra=true, ncol=nc), b.r, G.vector(b, :t_out, 1)[1:nc], dt
end

"pdesolver_membSFK_rect (basepdesolver_rect.jl:298-569): 8-element D (D[3:8] used), both SFK diffusivities 1e-32, modulus snapshots"
function pdesolver_membSFK_rect(Co::Vector{Float64}, D::Vector{Float64}, k::Vector{Float64};
    R::Float64=10.0, dr::Float64=0.1, tf::Float64=5.0,
    Nts::Int64=100, dt::Float64=1.0/(2.0*(maximum(D)/dr.^2 + sum(k)/4))*0.99,
    maxiters=20,
    tol::Float64=1.0e-6)
    G = GAB1PDE
    D7 = D[[1; 3:8]]                                    # D[2] is never read (:305-312); D[8] must exist, as in the reference
    b = G.pdesolver_batch(Co, G.row(D7), G.row(k); R, dr, tf, Nts, dt=[dt], maxiters, tol, geometry=1, sfk_mode=2, save_rule=1, pg1tot_form=1)
    G.check_throw(b)
    (b.status[1] & G.ST_OVERFLOW != 0) && error("pdesolver_membSFK_rect: the modulus rule (i-1) % round(Nt/Nts) == 0 takes more than Nts+1 snapshots")
    nc = Int(b.n_saved[1])
    return G.sol_tuple(b, 1; extra=true, ncol=nc), b.r, G.vector(b, :t_out, 1)[1:nc], dt
end

"pulsechase_solver (pulsechase_solver.jl:29-318): pdesolver with kp := 0 once t >= t_prechase (:156-158)"
function pulsechase_solver(Co::AbstractVector, D::AbstractVector, k::AbstractVector{T};
    R::Float64=10.0, dr::Float64=0.1,
    t_prechase::Float64=5.0,
    t_chase::Float64=2.0,
    tf::Float64=t_prechase + t_chase,
    Nts::Int=100,
    dt::Float64=1.0/(2.0*(maximum(D)/dr.^2 + sum(k)/4))*0.99,
    dt_save=tf/Nts,
    maxiters=20,
    tol::Float64 = 1.0e-6) where T
    G = GAB1PDE
    b = G.pdesolver_batch(G.f64(Co, "Co"), G.row(G.f64(D, "D")), G.row(G.f64(k, "k")); R, dr, tf, Nts, dt=[dt], dt_save, maxiters, tol, t_prechase)
    G.check_throw(b); G.check_overflow(b)
    return G.sol_tuple(b, 1; extra=true), b.r, G.vector(b, :t_out, 1), t_prechase, t_chase, dt_save
end

"""pdesolver_fitting (basepdesolver.jl:674-932): p = [D; k; Co]; `T = Float64` solves on the GPU, `T <: ForwardDiff.Dual`
propagates the partials through the whole time loop on the GPU (gab1_solve_tangent) and re-wraps the outputs as duals of
the caller's tag, so `ForwardDiff.gradient(testf, x)`, LBFGS under `AutoForwardDiff()` and NUTS
(param_fitting+inference_finitediff.jl:128-151,188-240,308-370) run unchanged."""
function pdesolver_fitting(p::AbstractVector{T};
    Diff_inds = 1:7,
    k_inds = Diff_inds[end] .+ (1:17),
    Co_inds = k_inds[end] .+ (1:5),
    R=10.0, dr=0.1, tf=5.0,
    Nts=100,
    dt_save=tf/Nts,
    maxiters=20,
    tol=1.0e-6) where T
    G = GAB1PDE
    D = p[Diff_inds]; k = p[k_inds]; Co = p[Co_inds]
    dt = 1.0/(2.0*(maximum(D)/(dr.^2) + sum(k)/4))*0.99                       # :696, a Dual when p is
    if T <: AbstractFloat
        b = G.pdesolver_batch(Co, G.row(D), G.row(k); R, dr, tf, Nts, dt=[Float64(dt)], dt_save, maxiters, tol, matrix_mask=G.MASK_FITTING)
        if b.status[1] & G.ST_THROW != 0                                       # :730-735
            return (PG1S=zeros(10,10), G2PG1S=zeros(10,10), EG2PG1S=zeros(10,10)), ones(10), ones(10), dt
        end
        sol = (aSFK=G.matrix(b, :aSFK, 1), PG1S=G.matrix(b, :PG1S, 1), G2PG1S=G.matrix(b, :G2PG1S, 1), EG2PG1S=G.vector(b, :EG2PG1S, 1))
        return sol, b.r, G.vector(b, :t_out, 1), dt
    end
    return gab1_fitting_dual(p, D, k, Co, dt; R, dr, tf, Nts, dt_save, maxiters, tol)
end

# forward mode: loaded only when ForwardDiff is (the fitting scripts load it; the others never reach this method)
function gab1_fitting_dual(p, D, k, Co, dt; R, dr, tf, Nts, dt_save, maxiters, tol)
    G = GAB1PDE
    FD = Base.require(Base.PkgId(Base.UUID("f6369f11-7733-5829-9624-2563aa707210"), "ForwardDiff"))
    T = eltype(p)
    T <: FD.Dual || throw(ArgumentError("pdesolver_fitting on the GPU takes Float64 or ForwardDiff.Dual parameters, not $T"))
    N = FD.npartials(T)
    seeds = zeros(Float64, 30, N, 1)
    for d in 1:N
        seeds[1:7, d, 1] = [FD.partials(x, d) for x in D]
        seeds[8:24, d, 1] = [FD.partials(x, d) for x in k]
        seeds[25:29, d, 1] = [FD.partials(x, d) for x in Co]
        seeds[30, d, 1] = FD.partials(dt, d)
    end
    o = G.make_opts(; R, dr, tf, Nts, dt_save, maxiters, tol, matrix_mask=G.MASK_FITTING, n_devices=1)
    r = collect(0.0:dr:R)
    n = G.out_doubles(o)
    out = zeros(Float64, n, 1 + N, 1); status = zeros(Int32, 1); n_saved = zeros(Int32, 1); n_steps = zeros(Int64, 1); n_bc = zeros(Int64, 1)
    rc = ccall((:gab1_solve_tangent, G.LIB), Cint,
               (Ref{G.Opts}, Int64, Int32, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}),
               o, 1, N, FD.value.(Co), 0, FD.value.(D), FD.value.(k), [FD.value(dt)], seeds, r, out, status, n_saved, n_steps, n_bc)
    rc == 0 || error("gab1_solve_tangent: " * G.last_error())
    if status[1] & G.ST_THROW != 0
        z() = zeros(T, 10, 10)
        return (PG1S=z(), G2PG1S=z(), EG2PG1S=z()), ones(10), ones(10), dt
    end
    P, C = o.Nr + 1, o.Nts + 1
    dual(i) = T(out[i, 1, 1], FD.Partials(ntuple(d -> out[i, 1 + d, 1], N)))
    mat(name) = (off = ccall((:gab1_full_matrix_offset, G.LIB), Int64, (Ref{G.Opts}, Int32), o, findfirst(==(name), G.MATRICES) - 1);
                 reshape([dual(off + i) for i in 1:P*C], P, C))
    vec(name) = (off = ccall((:gab1_full_vector_offset, G.LIB), Int64, (Ref{G.Opts}, Int32), o, findfirst(==(name), G.VECTORS) - 1);
                 [dual(off + i) for i in 1:C])
    return (aSFK=mat(:aSFK), PG1S=mat(:PG1S), G2PG1S=mat(:G2PG1S), EG2PG1S=vec(:EG2PG1S)), r, vec(:t_out), dt
end

function gab1_sa(Co, D, k, membSFK; kw...)
    G = GAB1PDE
    b = G.sapdesolver_batch(Co, G.row(D), G.row(k); membSFK, kw...)
    G.check_throw(b); P = b.o.Nr + 1; o = b.out
    return (iSFK=o[1:P, 1], aSFK=o[P+1:2P, 1], PG1tot=o[2P+1:3P, 1], PG1Stot=o[3P+1:4P, 1]), b.r
end

"sapdesolver (sapdesolver.jl:55-280); R, dr, tf default to the file-level globals (:11-14)"
function sapdesolver(Co::Vector{Float64}, D::Vector{Float64}, k::Vector{Float64};
    R::Float64=R, dr::Float64=dr, tf::Float64=tf,
    dt::Float64=1.0/(2.0*(maximum(D)/dr.^2 + sum(k)/4))*0.99,
    maxiters=20,
    tol::Float64 = 1.0e-3)
    return gab1_sa(Co, D, k, false; R, dr, tf, dt=[dt], maxiters, tol)
end

"sapdesolver_membSFK (sapdesolver_memb-SFK.jl:55-281); `maxiters` is accepted and unused, as in the reference"
function sapdesolver_membSFK(Co::Vector{Float64}, D::Vector{Float64}, k::Vector{Float64};
    R::Float64=R, dr::Float64=dr, tf::Float64=tf,
    dt::Float64=1.0/(2.0*(maximum(D)/dr.^2 + sum(k)/4))*0.99,
    maxiters=20,
    tol::Float64 = 1.0e-3)
    return gab1_sa(Co, D, k, true; R, dr, tf, dt=[dt], maxiters, tol)
end


## ================================================================ ensemble drivers (get_param_posteriors.jl)

"run_ensemble (get_param_posteriors.jl:135-168): the Threads.@threads loop over sets (:147) is one batched GPU call"
function run_ensemble(model_fun, ensemble, Co;
    dr=0.2, R=10.0, tf=5.0, Nts=100, tol=1e-4, maxit=20,
    D_inds=1:7, k_inds=8:24,
    show_prog=true,
    )
    G = GAB1PDE
    b = G.pdesolver_batch(Co, ensemble[:, D_inds], ensemble[:, k_inds]; R, dr, tf, Nts, tol, maxiters=maxit, G.variant(model_fun)...)
    rect = nameof(model_fun) === :pdesolver_rect
    retro_full_df = DataFrame()
    for j in axes(ensemble, 1)
        # the reference lets an exception out of the threaded loop: InexactError from Int64(ceil(tf/dt)), BoundsError
        # from a snapshot beyond column Nts+1
        (b.status[j] & G.ST_THROW != 0) && throw(InexactError(:Int64, Int64, NaN))
        (b.status[j] & G.ST_OVERFLOW != 0) && throw(BoundsError())
        b.status[j] & G.ST_NAN != 0 && continue                       # any(isnan.(sol.PG1S)) → skipped (:155)
        nc = rect ? Int(b.n_saved[j]) : Nts + 1
        append!(retro_full_df, DataFrame(r=[b.r], t_sol=[G.vector(b, :t_out, j)[1:nc]], sol=G.sol_tuple(b, j; extra=rect, ncol=nc), index=j))
    end
    return retro_full_df
end

"run_ensemble_pc (get_param_posteriors.jl:204-236) over pulsechase_solver"
function run_ensemble_pc(model_fun, ensemble, Co;
    dr=0.2, R=10.0,
    t_prechase=5.0,
    t_chase=2.0,
    Nts=100, tol=1e-4, maxit=20,
    D_inds=1:7, k_inds=8:24
    )
    G = GAB1PDE
    nameof(model_fun) === :pulsechase_solver || throw(ArgumentError("run_ensemble_pc on the GPU runs pulsechase_solver, not $(nameof(model_fun))"))
    b = G.pdesolver_batch(Co, ensemble[:, D_inds], ensemble[:, k_inds]; R, dr, tf=t_prechase + t_chase, Nts, tol, maxiters=maxit, t_prechase)
    retro_full_df = DataFrame()
    for j in axes(ensemble, 1)
        (b.status[j] & G.ST_THROW != 0) && throw(InexactError(:Int64, Int64, NaN))
        (b.status[j] & G.ST_OVERFLOW != 0) && throw(BoundsError())
        b.status[j] & G.ST_NAN != 0 && continue
        append!(retro_full_df, DataFrame(r=[b.r], t_sol=[G.vector(b, :t_out, j)], sol=G.sol_tuple(b, j; extra=true), index=j))
    end
    return retro_full_df
end


## ================================================================ GSA batch functions (sapdesolver.jl:288-476, sapdesolver_memb-SFK.jl:288-474)
# Keyword defaults `Co=Co, D=Diffs, kvals=kvals, R=R, dr=dr, tf=tf` are the reference's own: Main's globals at call time.

function pmap_fun_allpars(p; Co=Co, D=Diffs, kvals=kvals, R=R, dr=dr, tf=tf)
    lenCo = length(Co); lenDiffs = length(D); lenkvals = length(kvals)
    Co_inds = 1:lenCo
    Diffs_inds = (Co_inds[end]+1):(Co_inds[end] + lenDiffs)
    kvals_inds = (Diffs_inds[end]+1):(Diffs_inds[end] + lenkvals)
    G = GAB1PDE
    return G.one(G.six(p[Co_inds], G.row(p[Diffs_inds]), G.row(p[kvals_inds]); R, dr, tf, tol=1.0e-3, maxiters=20))
end

function pmap_fun_dk(p; Co=Co, D=Diffs, kvals=kvals, R=R, dr=dr, tf=tf, maxiters=100)
    lenDiffs = length(D); lenkvals = length(kvals)
    Diffs_inds = 1:lenDiffs
    kvals_inds = (Diffs_inds[end]+1):(Diffs_inds[end] + lenkvals)
    G = GAB1PDE
    return G.one(G.six(Co, G.row(p[Diffs_inds]), G.row(p[kvals_inds]); R, dr, tf, tol=1e-3, maxiters))
end

function pmap_fun_dk_combD(p; Co=Co, D=Diffs, kvals=kvals, R=R, dr=dr, tf=tf)
    lenkvals = length(kvals)
    kvals_inds = 2:(1 + lenkvals)
    G = GAB1PDE
    return G.one(G.six(Co, G.row(D .* (p[1]/D[1])), G.row(p[kvals_inds]); R, dr, tf, tol=1.0e-3, maxiters=20))
end

function pmap_fun_concs(p; Co=Co, D=Diffs, kvals=kvals, R=R, dr=dr, tf=tf, tol=1e-3, maxiters=20)
    G = GAB1PDE
    return G.one(G.six(p, G.row(D), G.row(kvals); R, dr, tf, tol, maxiters))
end

# batched forms: one GPU call per matrix of columns.  A column whose solve or reduction throws in the reference
# comes back as zeros(numout) — fbatch_*_mt's `catch` (sapdesolver.jl:378-382), pmap's `on_error` (:363-366);
# the plain pmap forms (fbatch, fbatch_dk_combD, fbatch_concs) let the exception out, and so do these.
function gab1_fbatch(Co, Dm, km; R, dr, tf, tol, maxiters, rethrow)
    G = GAB1PDE
    b = G.six(Co, Dm, km; R, dr, tf, tol, maxiters)
    rethrow && any(b.status .& G.ST_THROW .!= 0) && throw(ArgumentError("reducing over an empty collection is not allowed"))
    return Matrix(b.out)
end

function fbatch(p_batch; numout=6, Co=Co, D=Diffs, kvals=kvals, R=R, dr=dr, tf=tf)
    nC = length(Co); nD = length(D); nk = length(kvals)
    P = p_batch
    return gab1_fbatch(permutedims(P[1:nC, :]), permutedims(P[nC+1:nC+nD, :]), permutedims(P[nC+nD+1:nC+nD+nk, :]); R, dr, tf, tol=1.0e-3, maxiters=20, rethrow=true)
end

function fbatch_dk(p_batch; numout=6, Co=Co, D=Diffs, kvals=kvals, R=R, dr=dr, tf=tf)
    P = exp.(p_batch); nD = length(D); nk = length(kvals)
    return gab1_fbatch(Co, permutedims(P[1:nD, :]), permutedims(P[nD+1:nD+nk, :]); R, dr, tf, tol=1e-3, maxiters=100, rethrow=false)
end

function fbatch_dk_mt(p_batch; numout=6, Co=Co, D=Diffs, kvals=kvals, R=R, dr=dr, tf=tf, maxiters=20)
    P = exp.(p_batch); nD = length(D); nk = length(kvals)
    return gab1_fbatch(Co, permutedims(P[1:nD, :]), permutedims(P[nD+1:nD+nk, :]); R, dr, tf, tol=1e-3, maxiters, rethrow=false)
end

function fbatch_dk_combD(p_batch; numout=6, Co=Co, D=Diffs, kvals=kvals, R=R, dr=dr, tf=tf)
    P = p_batch; nk = length(kvals); S = size(P, 2)
    Dm = permutedims(hcat([D .* (P[1, i]/D[1]) for i in 1:S]...))
    return gab1_fbatch(Co, Dm, permutedims(P[2:1+nk, :]); R, dr, tf, tol=1.0e-3, maxiters=20, rethrow=true)
end

function fbatch_concs(p_batch; numout=6, Co=Co, D=Diffs, kvals=kvals, R=R, dr=dr, tf=tf)
    P = p_batch; S = size(P, 2)
    return gab1_fbatch(permutedims(P), repeat(GAB1PDE.row(D), S), repeat(GAB1PDE.row(kvals), S); R, dr, tf, tol=1e-3, maxiters=20, rethrow=true)
end

function fbatch_concs_mt(p_batch; numout=6, Co=Co, D=Diffs, kvals=kvals, R=R, dr=dr, tf=tf)
    P = exp.(p_batch); S = size(P, 2)
    return gab1_fbatch(permutedims(P), repeat(GAB1PDE.row(D), S), repeat(GAB1PDE.row(kvals), S); R, dr, tf, tol=1e-3, maxiters=20, rethrow=false)
end


## ================================================================ new, batched entry points (north_star: "a batched sibling takes a matrix of parameter sets")
"pdesolver over the rows of Dmat (S×7) and kmat (S×17); returns a GAB1PDE.Batch (raw block + per-set diagnostics)"
pdesolver_batch(Co, Dmat, kmat; kw...) = GAB1PDE.pdesolver_batch(Co, Dmat, kmat; kw...)
"sapdesolver / sapdesolver_membSFK (membSFK=true) over the rows of Dmat and kmat"
sapdesolver_batch(Co, Dmat, kmat; kw...) = GAB1PDE.sapdesolver_batch(Co, Dmat, kmat; kw...)

"""
    ensemble_quantiles(ensemble, Co; probs=(:median, 0.5-0.341, 0.5+0.341), matrices=(:aSFK, :PG1tot, :PG1Stot), kw...)

The summary surfaces of run_base_model.jl:103-174 (`median(stack, dims=3)`, `quantile(stack[node, col, :], p)` over the sets
run_ensemble keeps) computed on the GPU: the full solutions never leave the device.  Returns
`(Dict(name => Array (Nr+1) × ncols × length(probs)), n_valid, r, status)`; run_ensemble's solver defaults.
"""
function ensemble_quantiles(ensemble, Co; probs=(:median, 0.5 - 0.341, 0.5 + 0.341), matrices=(:aSFK, :PG1tot, :PG1Stot),
                            columns=nothing, dr=0.2, R=10.0, tf=5.0, Nts=100, tol=1e-4, maxit=20, D_inds=1:7, k_inds=8:24, kw...)
    G = GAB1PDE
    mask = UInt32(sum(1 << (findfirst(==(m), G.MATRICES) - 1) for m in matrices))
    o = G.make_opts(; R, dr, tf, Nts, maxiters=maxit, tol, out_mode=G.OUT_FULL, matrix_mask=mask, kw...)
    c0, c1 = columns === nothing ? (0, Nts + 1) : columns
    S = size(ensemble, 1)
    Dt = permutedims(Float64.(ensemble[:, D_inds])); kt = permutedims(Float64.(ensemble[:, k_inds]))
    dt = [G.default_dt(ensemble[j, D_inds], ensemble[j, k_inds], dr) for j in 1:S]
    r = collect(0.0:dr:R)
    p = Float64[x === :median ? -1.0 : x for x in probs]
    q = zeros(Float64, o.Nr + 1, c1 - c0, length(p), length(matrices))      # C order [matrix][p][column][node]
    status = zeros(Int32, S); n_saved = zeros(Int32, S); n_steps = zeros(Int64, S); n_bc = zeros(Int64, S); nv = zeros(Int64, 1)
    rc = ccall((:gab1_solve_ensemble_quantiles, G.LIB), Cint,
               (Ref{G.Opts}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, UInt32, Int32,
                Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
               o, S, Float64.(Co), 0, Dt, kt, dt, r, mask, c0, c1, length(p), p, q, status, n_saved, n_steps, n_bc, nv)
    rc == 0 || error("gab1_solve_ensemble_quantiles: " * G.last_error())
    order = sort(collect(matrices); by=m -> findfirst(==(m), G.MATRICES))
    return Dict(name => q[:, :, :, i] for (i, name) in enumerate(order)), nv[1], r, status
end

"`linear_interpolation(r, y)(0:dr_new:R)` along the node axis (first dimension): the re-gridding of run_base_model.jl:108-119"
function regrid(values::AbstractArray, r::AbstractVector; dr_new=0.1, R=10.0)
    x = collect(0.0:dr_new:R)
    out = similar(values, Float64, (length(x), size(values)[2:end]...))
    for (j, xj) in enumerate(x)
        i = clamp(searchsortedlast(r, xj), 1, length(r) - 1)
        w = (xj - r[i]) / (r[i+1] - r[i])
        selectdim(out, 1, j) .= (1 - w) .* selectdim(values, 1, i) .+ w .* selectdim(values, 1, i + 1)
    end
    return out, x
end

"""Synthetic prior ensemble drawn on the device (gab1_sample_prior; the prior half of generate_ensemble,
get_param_posteriors.jl:53-76, from the library's own Philox stream): 22 log-normal (mu, sigma) in the order of include/gab1pde.h.
Returns (Dmat S×7, kmat S×17), ready for `sapdesolver_batch` / `fbatch_*`."""
function sample_prior(S::Integer, seed::Integer, mu::Vector{Float64}, sigma::Vector{Float64}; EGF=1.67e-3, Kdd=0.38)
    length(mu) == 22 && length(sigma) == 22 || throw(ArgumentError("mu and sigma must hold 22 entries"))
    Dt = zeros(Float64, 7, S); kt = zeros(Float64, 17, S)
    rc = ccall((:gab1_sample_prior, GAB1PDE.LIB), Cint,
               (Int64, UInt64, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Ptr{Float64}, Ptr{Float64}),
               S, seed, mu, sigma, EGF, Kdd, Dt, kt)
    rc == 0 || error("gab1_sample_prior: " * GAB1PDE.last_error())
    return permutedims(Dt), permutedims(kt)
end
