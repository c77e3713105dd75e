# GAB1PDE.jl — drop-in Julia front end for libgab1pde.so (B200 batched solver).
#
# `include("GAB1PDE.jl")` *after* the reference's own includes replaces the CPU solvers and ensemble drivers of
# basepdesolver.jl, basepdesolver_rect.jl, sapdesolver.jl, sapdesolver_memb-SFK.jl, pulsechase_solver.jl and
# get_param_posteriors.jl with methods of the same names, argument order, keyword defaults and return shapes, and
# adds the batched siblings `pdesolver_batch` / `sapdesolver_batch`.  All arithmetic happens in the shared library
# (hand-written sm_100a CUDA); this file only marshals arrays and evaluates the two inputs whose last bits depend
# on Julia itself: the default `dt` (Base `sum` is a SIMD reduction) and the grid `collect(0.0:dr:R)`.
#
# NOTE: no Julia runtime exists in the image this library is developed in, so this file has not been executed
# there; it is kept in lock-step with the Python twin (host.py), which the test-suite does execute.
module GAB1PDE

export pdesolver, pdesolver_membSFK, pdesolver_rect, pdesolver_membSFK_rect, pdesolver_fitting, pulsechase_solver,
       sapdesolver, sapdesolver_membSFK, pdesolver_batch, sapdesolver_batch, run_ensemble, run_ensemble_pc,
       pmap_fun_dk, pmap_fun_allpars, pmap_fun_dk_combD, pmap_fun_concs, fbatch_dk_mt, fbatch_concs_mt

using DataFrames

const LIB = get(ENV, "GAB1PDE_LIB", joinpath(@__DIR__, "..", "myers-furcht-et-al_gab1-shp2-pde-model_b200", "libgab1pde.so"))

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

make_opts(; R, dr, tf, Nts, dt_save=tf / Nts, maxiters, tol, geometry=0, sfk_mode=0, bc_loop=0, save_rule=0,
          pg1tot_form=0, out_mode=OUT_FULL, matrix_mask=0x0fff, t_prechase=-1.0, pct_mul=1.0, pct_div=1.0,
          n_devices=0) =
    Opts(1, geometry, sfk_mode, bc_loop, save_rule, pg1tot_form, out_mode, matrix_mask, maxiters,
         Int32(ceil(R / dr)), Nts, 0, tol, R, dr, tf, dt_save, t_prechase, pct_mul, pct_div, n_devices, 0, C_NULL)

out_doubles(o::Opts) = ccall((:gab1_out_doubles_per_set, LIB), Int64, (Ref{Opts},), o)

# dt exactly as the reference's keyword default (basepdesolver.jl:30), evaluated by Julia
default_dt(D, k, dr) = 1.0 / (2.0 * (maximum(D) / (dr .^ 2) + sum(k) / 4)) * 0.99

struct Batch
    o::Opts; out::Matrix{Float64}; status::Vector{Int32}; n_saved::Vector{Int32}
    n_steps::Vector{Int64}; n_bc_iters::Vector{Int64}; r::Vector{Float64}; dt::Vector{Float64}
end

"Run S parameter sets (rows of Dmat S×7 and kmat S×17). Co is a 5-vector shared by all sets or S×5."
function solve(o::Opts, Co, Dmat, kmat, dt::Vector{Float64}, r::Vector{Float64})
    S = size(Dmat, 1)
    length(r) == o.Nr + 1 || throw(BoundsError(r, o.Nr + 1))           # the reference indexes r[Nr+1]
    Dt = permutedims(Float64.(Dmat)); kt = permutedims(Float64.(kmat))   # row-major for C
    Cot, stride = Co isa AbstractVector ? (Float64.(Co), 0) : (permutedims(Float64.(Co)), 5)
    n = out_doubles(o)
    out = zeros(Float64, n, S); status = zeros(Int32, S); n_saved = zeros(Int32, S)
    n_steps = zeros(Int64, S); n_bc = zeros(Int64, S)
    rc = ccall((:gab1_solve_batch, LIB), Cint,
               (Ref{Opts}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}),
               o, S, Cot, stride, Dt, kt, dt, r, out, status, n_saved, n_steps, n_bc)
    rc == 0 || error("gab1_solve_batch: " * unsafe_string(ccall((:gab1_last_error, LIB), Cstring, ())))
    Batch(o, out, status, n_saved, n_steps, n_bc, r, dt)
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

## ---------------------------------------------------------------- batched siblings (new)
"Batched sibling of `pdesolver` (basepdesolver.jl:25-312)."
function pdesolver_batch(Co, Dmat, kmat; R=10.0, dr=0.1, tf=5.0, Nts=100,
                         dt=[default_dt(Dmat[j, :], kmat[j, :], dr) for j in axes(Dmat, 1)],
                         dt_save=tf / Nts, maxiters=100, tol=1.0e-6, r=collect(0.0:dr:R), kw...)
    solve(make_opts(; R, dr, tf, Nts, dt_save, maxiters, tol, kw...), Co, Dmat, kmat, Float64.(dt), r)
end

"Batched sibling of `sapdesolver` / `sapdesolver_membSFK` (sapdesolver.jl:55-280, sapdesolver_memb-SFK.jl:55-281)."
function sapdesolver_batch(Co, Dmat, kmat; R=10.0, dr=0.2, tf=5.0,
                           dt=[default_dt(Dmat[j, :], kmat[j, :], dr) for j in axes(Dmat, 1)],
                           maxiters=20, tol=1.0e-3, membSFK=false, out_mode=OUT_FINAL4, r=collect(0.0:dr:R),
                           iter_cap=100_000)
    # membSFK: the reference's `while error > tol` has no cap (sapdesolver_memb-SFK.jl:177) and spins forever on a fixed
    # point that never meets tol; the library stops such a step after `iter_cap` passes and flags the set (ST_ITER_CAP)
    o = make_opts(; R, dr, tf, Nts=1, maxiters=membSFK ? iter_cap : maxiters, tol, out_mode,
                  sfk_mode=membSFK ? 1 : 0, bc_loop=membSFK ? 1 : 0, pg1tot_form=membSFK ? 1 : 0)
    solve(o, Co, Dmat, kmat, Float64.(dt), r)
end

"""
    ensemble_quantiles(ensemble, Co; probs=(:median, 0.5-0.341, 0.5+0.341), matrices=(:aSFK, :PG1tot, :PG1Stot), kw...)

The summary surfaces of run_base_model.jl:103-174 (`median(stack, dims=3)`, `quantile(stack[node, col, :], p)` over the sets
run_ensemble keeps) computed on the GPU: the full solutions never leave the device.  Returns
`(Dict(name => Array (Nr+1) × ncols × length(probs)), n_valid, r, status)`; run_ensemble's solver defaults.
"""
function ensemble_quantiles(ensemble, Co; probs=(:median, 0.5 - 0.341, 0.5 + 0.341), matrices=(:aSFK, :PG1tot, :PG1Stot),
                            columns=nothing, dr=0.2, R=10.0, tf=5.0, Nts=100, tol=1e-4, maxit=20, D_inds=1:7, k_inds=8:24, kw...)
    mask = UInt32(sum(1 << (findfirst(==(m), MATRICES) - 1) for m in matrices))
    o = make_opts(; R, dr, tf, Nts, maxiters=maxit, tol, out_mode=OUT_FULL, matrix_mask=mask, kw...)
    c0, c1 = columns === nothing ? (0, Nts + 1) : columns
    S = size(ensemble, 1)
    Dt = permutedims(Float64.(ensemble[:, D_inds])); kt = permutedims(Float64.(ensemble[:, k_inds]))
    dt = [default_dt(ensemble[j, D_inds], ensemble[j, k_inds], dr) for j in 1:S]
    r = collect(0.0:dr:R)
    p = Float64[x === :median ? -1.0 : x for x in probs]
    q = zeros(Float64, o.Nr + 1, c1 - c0, length(p), length(matrices))      # C order [matrix][p][column][node]
    status = zeros(Int32, S); n_saved = zeros(Int32, S); n_steps = zeros(Int64, S); n_bc = zeros(Int64, S); nv = zeros(Int64, 1)
    rc = ccall((:gab1_solve_ensemble_quantiles, LIB), Cint,
               (Ref{Opts}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, UInt32, Int32,
                Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
               o, S, Float64.(Co), 0, Dt, kt, dt, r, mask, c0, c1, length(p), p, q, status, n_saved, n_steps, n_bc, nv)
    rc == 0 || error("gab1_solve_ensemble_quantiles: " * unsafe_string(ccall((:gab1_last_error, LIB), Cstring, ())))
    order = sort(collect(matrices); by=m -> findfirst(==(m), MATRICES))
    Dict(name => q[:, :, :, i] for (i, name) in enumerate(order)), nv[1], r, status
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
    out, x
end

## ---------------------------------------------------------------- single solves, reference names
function _sol(b::Batch, j; extra=false, ncol=b.o.Nts + 1)
    mats = NamedTuple{MATRICES}(Tuple(matrix(b, n, j)[:, 1:ncol] for n in MATRICES))
    vecs = NamedTuple{VECTORS[1:9]}(Tuple(vector(b, n, j)[1:ncol] for n in VECTORS[1:9]))
    extra ? merge(mats, (EGFR_SHP2=vector(b, :EGFR_SHP2, j)[1:ncol],), vecs) : merge(mats, vecs)
end
_check(b) = (b.status[1] & ST_THROW != 0) && throw(InexactError(:Int64, Int64, NaN))

"pdesolver (basepdesolver.jl:25-312) → (sol, r, t_out, dt)"
function pdesolver(Co::AbstractVector, D::AbstractVector, k::AbstractVector; R=10.0, dr=0.1, tf=5.0, Nts=100,
                   dt=default_dt(D, k, dr), dt_save=tf / Nts, maxiters=100, tol=1.0e-6)
    b = pdesolver_batch(Co, reshape(D, 1, :), reshape(k, 1, :); R, dr, tf, Nts, dt=[dt], dt_save, maxiters, tol)
    _check(b); (b.status[1] & ST_OVERFLOW != 0) && throw(BoundsError())
    _sol(b, 1), b.r, vector(b, :t_out, 1), dt
end
"pdesolver_membSFK (basepdesolver.jl:350-636)"
function pdesolver_membSFK(Co::AbstractVector, D::AbstractVector, k::AbstractVector; R=10.0, dr=0.1, tf=5.0, Nts=100,
                           dt=default_dt(D, k, dr), dt_save=tf / Nts, maxiters=20, tol=1.0e-6)
    b = pdesolver_batch(Co, reshape(D, 1, :), reshape(k, 1, :); R, dr, tf, Nts, dt=[dt], dt_save, maxiters, tol, sfk_mode=1)
    _check(b); _sol(b, 1), b.r, vector(b, :t_out, 1), dt
end
"pdesolver_rect (basepdesolver_rect.jl:23-294): outputs hold 1 + #snapshots columns"
function pdesolver_rect(Co::Vector{Float64}, D::Vector{Float64}, k::Vector{Float64}; R=10.0, dr=0.1, tf=5.0, Nts=100,
                        dt=default_dt(D, k, dr), dt_save=tf / Nts, maxiters=20, tol=1.0e-6)
    b = pdesolver_batch(Co, reshape(D, 1, :), reshape(k, 1, :); R, dr, tf, Nts, dt=[dt], dt_save, maxiters, tol,
                        geometry=1, pg1tot_form=1)
    _check(b); nc = Int(b.n_saved[1])
    _sol(b, 1; extra=true, ncol=nc), b.r, vector(b, :t_out, 1)[1:nc], dt
end
"pdesolver_membSFK_rect (basepdesolver_rect.jl:298-569): 8-element D, both SFK diffusivities 1e-32, modulus snapshots"
function pdesolver_membSFK_rect(Co::Vector{Float64}, D::Vector{Float64}, k::Vector{Float64}; R=10.0, dr=0.1, tf=5.0,
                                Nts=100, dt=default_dt(D, k, dr), maxiters=20, tol=1.0e-6)
    D7 = D[[1; 3:8]]
    b = pdesolver_batch(Co, reshape(D7, 1, :), reshape(k, 1, :); R, dr, tf, Nts, dt=[dt], maxiters, tol,
                        geometry=1, sfk_mode=2, save_rule=1, pg1tot_form=1)
    _check(b); nc = Int(b.n_saved[1])
    _sol(b, 1; extra=true, ncol=nc), b.r, vector(b, :t_out, 1)[1:nc], dt
end
"Float64 path of pdesolver_fitting (basepdesolver.jl:674-932); the ForwardDiff.Dual method is below (solve_tangent)"
function pdesolver_fitting(p::AbstractVector{Float64}; Diff_inds=1:7, k_inds=Diff_inds[end] .+ (1:17),
                           Co_inds=k_inds[end] .+ (1:5), R=10.0, dr=0.1, tf=5.0, Nts=100, dt_save=tf / Nts,
                           maxiters=20, tol=1.0e-6)
    D, k, Co = p[Diff_inds], p[k_inds], p[Co_inds]
    dt = default_dt(D, k, dr)
    b = pdesolver_batch(Co, reshape(D, 1, :), reshape(k, 1, :); R, dr, tf, Nts, dt=[dt], dt_save, maxiters, tol,
                        matrix_mask=UInt32((1 << 1) | (1 << 9) | (1 << 7)))
    if b.status[1] & ST_THROW != 0
        return (PG1S=zeros(10, 10), G2PG1S=zeros(10, 10), EG2PG1S=zeros(10, 10)), ones(10), ones(10), dt
    end
    (aSFK=matrix(b, :aSFK, 1), PG1S=matrix(b, :PG1S, 1), G2PG1S=matrix(b, :G2PG1S, 1), EG2PG1S=vector(b, :EG2PG1S, 1)),
    b.r, vector(b, :t_out, 1), dt
end
## ---------------------------------------------------------------- forward mode: pdesolver_fitting on ForwardDiff duals
# ForwardDiff.gradient(testf, x) / AutoForwardDiff() / Turing NUTS call pdesolver_fitting with eltype(p) <: Dual
# (param_fitting+inference_finitediff.jl:128-151,188-240,308-370).  This method peels values and partials off the duals,
# lets the library propagate the partials through the whole time loop (gab1_solve_tangent), and re-wraps the outputs as
# duals with the caller's tag, so the calling code (loss, testf, turing_model) runs unchanged.
using ForwardDiff: Dual, Partials, value, partials

"values + partials of S sets: seeds is 30 × n_dir × S ([D;k;Co;dt] partials); returns out (n, 1+n_dir, S) and diagnostics"
function solve_tangent(o::Opts, Co, Dmat, kmat, dt::Vector{Float64}, seeds::Array{Float64,3}, r::Vector{Float64})
    S = size(Dmat, 1); n_dir = size(seeds, 2)
    Dt = permutedims(Float64.(Dmat)); kt = permutedims(Float64.(kmat))
    Cot, stride = Co isa AbstractVector ? (Float64.(Co), 0) : (permutedims(Float64.(Co)), 5)
    n = out_doubles(o)
    out = zeros(Float64, n, 1 + n_dir, S); status = zeros(Int32, S); n_saved = zeros(Int32, S)
    n_steps = zeros(Int64, S); n_bc = zeros(Int64, S)
    rc = ccall((:gab1_solve_tangent, LIB), Cint,
               (Ref{Opts}, Int64, Int32, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}),
               o, S, n_dir, Cot, stride, Dt, kt, dt, seeds, r, out, status, n_saved, n_steps, n_bc)
    rc == 0 || error("gab1_solve_tangent: " * unsafe_string(ccall((:gab1_last_error, LIB), Cstring, ())))
    out, status, n_saved, n_steps, n_bc
end

function pdesolver_fitting(p::AbstractVector{Dual{Tg,Float64,N}}; Diff_inds=1:7, k_inds=Diff_inds[end] .+ (1:17),
                           Co_inds=k_inds[end] .+ (1:5), R=10.0, dr=0.1, tf=5.0, Nts=100, dt_save=tf / Nts,
                           maxiters=20, tol=1.0e-6) where {Tg,N}
    D, k, Co = p[Diff_inds], p[k_inds], p[Co_inds]
    dt = 1.0 / (2.0 * (maximum(D) / (dr .^ 2) + sum(k) / 4)) * 0.99          # a Dual, as at basepdesolver.jl:696
    seeds = zeros(Float64, 30, N, 1)
    for d in 1:N
        seeds[1:7, d, 1] = [partials(x, d) for x in D]
        seeds[8:24, d, 1] = [partials(x, d) for x in k]
        seeds[25:29, d, 1] = [partials(x, d) for x in Co]
        seeds[30, d, 1] = partials(dt, d)
    end
    o = make_opts(; R, dr, tf, Nts, dt_save, maxiters, tol, matrix_mask=UInt32((1 << 1) | (1 << 9) | (1 << 7)))
    r = collect(0.0:dr:R)
    out, status, = solve_tangent(o, value.(Co), reshape(value.(D), 1, :), reshape(value.(k), 1, :), [value(dt)], seeds, r)
    if status[1] & ST_THROW != 0
        z() = zeros(eltype(p), 10, 10)
        return (PG1S=z(), G2PG1S=z(), EG2PG1S=z()), ones(10), ones(10), dt
    end
    P, C = o.Nr + 1, o.Nts + 1
    dual(i) = Dual{Tg}(out[i, 1, 1], Partials(ntuple(d -> out[i, 1 + d, 1], N)))
    mat(name) = (off = ccall((:gab1_full_matrix_offset, LIB), Int64, (Ref{Opts}, Int32), o, findfirst(==(name), MATRICES) - 1);
                 reshape([dual(off + i) for i in 1:P*C], P, C))
    vec(name) = (off = ccall((:gab1_full_vector_offset, LIB), Int64, (Ref{Opts}, Int32), o, findfirst(==(name), VECTORS) - 1);
                 [dual(off + i) for i in 1:C])
    (aSFK=mat(:aSFK), PG1S=mat(:PG1S), G2PG1S=mat(:G2PG1S), EG2PG1S=vec(:EG2PG1S)), r, vec(:t_out), dt
end

"""Synthetic prior ensemble drawn on the device (gab1_sample_prior; the prior half of generate_ensemble,
get_param_posteriors.jl:53-76, from the library's own Philox stream): 22 log-normal (mu, sigma) in the order of include/gab1pde.h.
Returns (Dmat S×7, kmat S×17), ready for `sapdesolver_batch` / `fbatch_*`."""
function sample_prior(S::Integer, seed::Integer, mu::Vector{Float64}, sigma::Vector{Float64}; EGF=1.67e-3, Kdd=0.38)
    length(mu) == 22 && length(sigma) == 22 || throw(ArgumentError("mu and sigma must hold 22 entries"))
    Dt = zeros(Float64, 7, S); kt = zeros(Float64, 17, S)
    rc = ccall((:gab1_sample_prior, LIB), Cint,
               (Int64, UInt64, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Ptr{Float64}, Ptr{Float64}),
               S, seed, mu, sigma, EGF, Kdd, Dt, kt)
    rc == 0 || error("gab1_sample_prior: " * unsafe_string(ccall((:gab1_last_error, LIB), Cstring, ())))
    permutedims(Dt), permutedims(kt)
end

"pulsechase_solver (pulsechase_solver.jl:29-318)"
function pulsechase_solver(Co::AbstractVector, D::AbstractVector, k::AbstractVector; R=10.0, dr=0.1, t_prechase=5.0,
                           t_chase=2.0, tf=t_prechase + t_chase, Nts=100, dt=default_dt(D, k, dr), dt_save=tf / Nts,
                           maxiters=20, tol=1.0e-6)
    b = pdesolver_batch(Co, reshape(D, 1, :), reshape(k, 1, :); R, dr, tf, Nts, dt=[dt], dt_save, maxiters, tol, t_prechase)
    _check(b); _sol(b, 1; extra=true), b.r, vector(b, :t_out, 1), t_prechase, t_chase, dt_save
end
function _sa(Co, D, k, membSFK; kw...)
    b = sapdesolver_batch(Co, reshape(D, 1, :), reshape(k, 1, :); membSFK, kw...)
    _check(b); P = b.o.Nr + 1; o = b.out
    (iSFK=o[1:P, 1], aSFK=o[P+1:2P, 1], PG1tot=o[2P+1:3P, 1], PG1Stot=o[3P+1:4P, 1]), b.r
end
"sapdesolver (sapdesolver.jl:55-280)"
sapdesolver(Co::Vector{Float64}, D::Vector{Float64}, k::Vector{Float64}; R=10.0, dr=0.2, tf=5.0,
            dt=default_dt(D, k, dr), maxiters=20, tol=1.0e-3) = _sa(Co, D, k, false; R, dr, tf, dt=[dt], maxiters, tol)
"sapdesolver_membSFK (sapdesolver_memb-SFK.jl:55-281)"
sapdesolver_membSFK(Co::Vector{Float64}, D::Vector{Float64}, k::Vector{Float64}; R=10.0, dr=0.2, tf=5.0,
                    dt=default_dt(D, k, dr), maxiters=20, tol=1.0e-3) = _sa(Co, D, k, true; R, dr, tf, dt=[dt], maxiters, tol)

## ---------------------------------------------------------------- ensemble drivers
_variant(f) = f === pdesolver_membSFK ? (sfk_mode=1,) : f === pdesolver_rect ? (geometry=1, pg1tot_form=1) : NamedTuple()

"run_ensemble (get_param_posteriors.jl:135-168): the Threads.@threads loop becomes one batched GPU call"
function run_ensemble(model_fun, ensemble, Co; dr=0.2, R=10.0, tf=5.0, Nts=100, tol=1e-4, maxit=20,
                      D_inds=1:7, k_inds=8:24, show_prog=true)
    b = pdesolver_batch(Co, ensemble[:, D_inds], ensemble[:, k_inds]; R, dr, tf, Nts, tol, maxiters=maxit, _variant(model_fun)...)
    rect = model_fun === pdesolver_rect
    df = DataFrame()
    for j in axes(ensemble, 1)
        b.status[j] & ST_NAN != 0 && continue                         # any(isnan.(sol.PG1S)) → skipped (:155)
        nc = rect ? Int(b.n_saved[j]) : Nts + 1
        append!(df, DataFrame(r=[b.r], t_sol=[vector(b, :t_out, j)[1:nc]], sol=_sol(b, j; extra=rect, ncol=nc), index=j))
    end
    df
end
"run_ensemble_pc (get_param_posteriors.jl:204-236)"
function run_ensemble_pc(model_fun, ensemble, Co; dr=0.2, R=10.0, t_prechase=5.0, t_chase=2.0, Nts=100, tol=1e-4,
                         maxit=20, D_inds=1:7, k_inds=8:24)
    b = pdesolver_batch(Co, ensemble[:, D_inds], ensemble[:, k_inds]; R, dr, tf=t_prechase + t_chase, Nts, tol,
                        maxiters=maxit, t_prechase)
    df = DataFrame()
    for j in axes(ensemble, 1)
        b.status[j] & ST_NAN != 0 && continue
        append!(df, DataFrame(r=[b.r], t_sol=[vector(b, :t_out, j)], sol=_sol(b, j; extra=true), index=j))
    end
    df
end

## ---------------------------------------------------------------- GSA batch functions (sapdesolver.jl:288-476)
_six(Co, Dm, km; R, dr, tf, tol, maxiters, membSFK=false) =
    sapdesolver_batch(Co, Dm, km; R, dr, tf, tol, maxiters, membSFK, out_mode=OUT_SIX)
_one(b) = (b.status[1] & ST_THROW != 0) ? throw(ArgumentError("reducing over an empty collection is not allowed")) : b.out[:, 1]

pmap_fun_dk(p; Co, D=nothing, kvals=nothing, R=10.0, dr=0.2, tf=5.0, maxiters=100) =
    _one(_six(Co, reshape(p[1:7], 1, :), reshape(p[8:24], 1, :); R, dr, tf, tol=1e-3, maxiters))
pmap_fun_allpars(p; R=10.0, dr=0.2, tf=5.0, kw...) =
    _one(_six(p[1:5], reshape(p[6:12], 1, :), reshape(p[13:29], 1, :); R, dr, tf, tol=1e-3, maxiters=20))
pmap_fun_dk_combD(p; Co, D, R=10.0, dr=0.2, tf=5.0, kw...) =
    _one(_six(Co, reshape(D .* (p[1] / D[1]), 1, :), reshape(p[2:18], 1, :); R, dr, tf, tol=1e-3, maxiters=20))
pmap_fun_concs(p; D, kvals, R=10.0, dr=0.2, tf=5.0, tol=1e-3, maxiters=20, kw...) =
    _one(_six(p, reshape(D, 1, :), reshape(kvals, 1, :); R, dr, tf, tol, maxiters))

"fbatch_dk_mt (sapdesolver.jl:371-387): 24×S log-space in, 6×S out; a column that throws in the reference is zeros(6)"
function fbatch_dk_mt(p_batch; numout=6, Co, D=nothing, kvals=nothing, R=10.0, dr=0.2, tf=5.0, maxiters=20)
    P = exp.(p_batch)
    _six(Co, permutedims(P[1:7, :]), permutedims(P[8:24, :]); R, dr, tf, tol=1e-3, maxiters).out
end
"fbatch_concs_mt (sapdesolver.jl:460-476): 5×S log-space initial concentrations"
function fbatch_concs_mt(p_batch; numout=6, Co=nothing, D, kvals, R=10.0, dr=0.2, tf=5.0)
    P = exp.(p_batch); S = size(P, 2)
    _six(permutedims(P), repeat(reshape(D, 1, :), S), repeat(reshape(kvals, 1, :), S); R, dr, tf, tol=1e-3, maxiters=20).out
end

end # module
