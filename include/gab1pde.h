/*
 * gab1pde.h — C ABI of libgab1pde.so, the B200 (sm_100a) batched solver for the
 * EGFR/GRB2/GAB1/SHP2/SFK reaction–diffusion model.
 *
 * The reference (pauljmyers/Myers-Furcht-et-al_GAB1-SHP2-PDE-model) has no FFI layer:
 * its boundary is the Julia call surface.  Each entry point below names the Julia
 * function(s) it stands in for (file:line under the reference's Julia/ directory).
 * A Julia maintainer binds these with `ccall` (see INTEGRATION.md and julia/gab1pde_dropin.jl).
 *
 * All buffers are owned by the caller.  Plain pointers and sizes only.
 * Every function returns 0 on success and a negative code on failure; the message
 * for the calling thread's last failure is available from gab1_last_error().
 * A parameter set that diverges numerically is NOT a failure: it is reported in
 * `status` and its values propagate NaN/Inf exactly as the reference does.
 */
#ifndef GAB1PDE_H
#define GAB1PDE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GAB1_ABI_VERSION 1

/* sizes fixed by the model (basepdesolver.jl:43-68,79) */
#define GAB1_N_CO 5      /* Co = [SFK, GRB2, GAB1, SHP2, EGFR]                        */
#define GAB1_N_D 7       /* D  = [D_S, D_G2, D_G2G1, D_G2G1S2, D_G1, D_G1S2, D_S2]    */
#define GAB1_N_K 17      /* k  = [kS2f,kS2r,kG1f,kG1r,kG2f,kG2r,kG1p,kG1dp,kSa,kSi,
                                   kp,kdp,kEGFf,kEGFr,EGF,kdf,kdr]                     */
#define GAB1_N_CYTO 10   /* cytosolic species carried on the radial grid              */
#define GAB1_N_MEMB 8    /* membrane species (ODEs at r = R)                           */

/* geometry: which Laplacian the interior update uses */
enum {
  GAB1_GEOM_SPHERICAL = 0,   /* basepdesolver.jl:151-179                               */
  GAB1_GEOM_RECT = 1         /* basepdesolver_rect.jl:132-160                          */
};

/* sfk_mode: which diffusivity the SFK species use */
enum {
  GAB1_SFK_DIFFUSIBLE = 0,   /* iSFK and aSFK both use D[0]        (pdesolver)          */
  GAB1_SFK_MEMBRANE = 1,     /* aSFK uses 1e-32                    (basepdesolver.jl:366,477,530;
                                                                    sapdesolver_memb-SFK.jl:62,134,187) */
  GAB1_SFK_BOTH_FROZEN = 2   /* iSFK and aSFK both use 1e-32       (basepdesolver_rect.jl:305-306)      */
};

/* bc_loop: form of the membrane fixed-point loop */
enum {
  GAB1_BC_FOR_BREAK = 0,     /* `for _ in 1:maxiters … error <= tol && break`  (basepdesolver.jl:197-242) */
  GAB1_BC_WHILE = 1          /* `error = 2tol; while error > tol`              (sapdesolver_memb-SFK.jl:175-222);
                                maxiters acts only as a safety cap that raises GAB1_ST_ITER_CAP */
};

/* save_rule: when a snapshot column is written (FULL / PCT_BOUND output) */
enum {
  GAB1_SAVE_T_GE_TSAVE = 0,  /* `t += dt; if t >= t_save` (basepdesolver.jl:265-295)   */
  GAB1_SAVE_MODULUS = 1      /* `(i-1) % round(Nt/Nts) == 0` (basepdesolver_rect.jl:336,526) */
};

/* pg1tot_form: association order of the derived PG1tot output */
enum {
  GAB1_PG1TOT_VIA_STOT = 0,  /* (G2PG1 + pGAB1) + (PG1S + G2PG1S)   basepdesolver.jl:299-300, sapdesolver.jl:255-256 */
  GAB1_PG1TOT_CHAIN = 1      /* ((G2PG1 + pGAB1) + PG1S) + G2PG1S   basepdesolver_rect.jl:261, sapdesolver_memb-SFK.jl:256 */
};

/* out_mode: what is written per parameter set */
enum {
  GAB1_OUT_FINAL4 = 0,       /* sapdesolver result: iSFK, aSFK, PG1tot, PG1Stot at t_end, 4*(Nr+1) doubles
                                (sapdesolver.jl:245-279)                                */
  GAB1_OUT_FULL = 1,         /* pdesolver result: selected (Nr+1)x(Nts+1) column-major matrices, then the
                                11 time-series vectors (basepdesolver.jl:83-111,268-311) */
  GAB1_OUT_SIX = 2,          /* pmap_fun_dk reduction: r1/2, r1/10 of aSFK and PG1Stot, centre/surface ratio,
                                volume average (sapdesolver.jl:343-356); zeros(6) if it would throw */
  GAB1_OUT_PCT_BOUND = 3,    /* % SHP2-bound GAB1 from the last snapshot column (run_base_model.jl:269-276) */
  GAB1_OUT_FINAL_STATE = 4   /* all 10 cytosolic profiles then the 8 membrane values at t_end:
                                10*(Nr+1)+8 doubles (diagnostic; same state sapdesolver holds at :245) */
};

/* matrices of GAB1_OUT_FULL, in the reference NamedTuple order (basepdesolver.jl:303-310) */
enum {
  GAB1_M_iSFK = 0, GAB1_M_aSFK, GAB1_M_GRB2, GAB1_M_GAB1, GAB1_M_SHP2, GAB1_M_G2G1,
  GAB1_M_G2PG1, GAB1_M_G2PG1S, GAB1_M_PG1, GAB1_M_PG1S, GAB1_M_PG1tot, GAB1_M_PG1Stot,
  GAB1_N_MATRICES
};
/* vectors of GAB1_OUT_FULL (length Nts+1 each), after the matrices */
enum {
  GAB1_V_pE = 0, GAB1_V_mE, GAB1_V_mES, GAB1_V_mESmES, GAB1_V_E, GAB1_V_EG2, GAB1_V_EG2G1,
  GAB1_V_EG2PG1, GAB1_V_EG2PG1S,
  GAB1_V_EGFR_SHP2,          /* basepdesolver_rect.jl:264, pulsechase_solver.jl:289 */
  GAB1_V_t_out,
  GAB1_N_VECTORS
};
#define GAB1_MASK_ALL_MATRICES 0xFFFu
/* the three matrices pdesolver_fitting returns (basepdesolver.jl:753-758,929) */
#define GAB1_MASK_FITTING ((1u << GAB1_M_aSFK) | (1u << GAB1_M_PG1S) | (1u << GAB1_M_G2PG1S))

/* per-set status bits */
#define GAB1_ST_NAN 1u        /* a NaN reached the outputs (reference: any(isnan.(sol.PG1S)) filter,
                                 get_param_posteriors.jl:155)                            */
#define GAB1_ST_ITER_CAP 2u   /* GAB1_BC_WHILE hit the safety cap                        */
#define GAB1_ST_SHORT 4u      /* fewer than Nts+1 snapshot columns were written          */
#define GAB1_ST_OVERFLOW 8u   /* more snapshots were due than columns exist (the reference would raise
                                 BoundsError / keep growing); extra snapshots dropped    */
#define GAB1_ST_THROW 16u     /* GAB1_OUT_SIX: the reference reduction would throw (empty `minimum`),
                                 column is zeros(6) as fbatch_*_mt's catch produces (sapdesolver.jl:378-382) */

typedef struct gab1_opts {
  int32_t abi_version;    /* must be GAB1_ABI_VERSION */
  int32_t geometry;       /* GAB1_GEOM_*        */
  int32_t sfk_mode;       /* GAB1_SFK_*         */
  int32_t bc_loop;        /* GAB1_BC_*          */
  int32_t save_rule;      /* GAB1_SAVE_*        */
  int32_t pg1tot_form;    /* GAB1_PG1TOT_*      */
  int32_t out_mode;       /* GAB1_OUT_*         */
  uint32_t matrix_mask;   /* GAB1_OUT_FULL: bit m set => matrix m is materialised */
  int32_t maxiters;       /* solver kwarg `maxiters` */
  int32_t Nr;             /* ceil(R/dr); the grid has Nr+1 nodes */
  int32_t Nts;            /* solver kwarg `Nts`; Nts+1 snapshot columns */
  int32_t arith;          /* 0 = fast (FMA-contracted, hoisted reciprocals; |rel err| << 1e-9),
                             1 = strict (operation-for-operation IEEE restatement, bit-identical to the oracle) */
  double tol;             /* solver kwarg `tol` */
  double R, dr, tf;       /* solver kwargs */
  double dt_save;         /* solver kwarg `dt_save` (= tf/Nts by default) */
  double t_prechase;      /* < 0: off.  Otherwise kp := 0 from the first step whose start time t satisfies
                             t_prechase + dt > t >= t_prechase (pulsechase_solver.jl:156-158) */
  double pct_mul, pct_div;/* GAB1_OUT_PCT_BOUND: membrane term is EG2PG1S*pct_mul/pct_div
                             (volCF, surfCF in run_base_model.jl:274) */
  int32_t n_devices;      /* host entry point only: number of GPUs to shard over; 0 => all visible */
  int32_t reserved;
  const int32_t* device_ids; /* n_devices CUDA ordinals, or NULL => 0..n_devices-1 */
} gab1_opts;

/* Fill `o` with the defaults of pdesolver (basepdesolver.jl:25-33) for a given grid. */
void gab1_opts_init(gab1_opts* o, double R, double dr, double tf, int32_t Nts);

/* Doubles written per parameter set for o->out_mode (0 on invalid options). */
int64_t gab1_out_doubles_per_set(const gab1_opts* o);
/* Offset (in doubles, from the start of a set's FULL block) of matrix m / vector v; -1 if absent. */
int64_t gab1_full_matrix_offset(const gab1_opts* o, int32_t m);
int64_t gab1_full_vector_offset(const gab1_opts* o, int32_t v);

/* dt exactly as the solvers' keyword default, `1.0/(2.0*(maximum(D)/(dr.^2) + sum(k)/4))*0.99`
 * (basepdesolver.jl:30), with sum(k) taken left to right.  Writes S values. */
int gab1_default_dt(int64_t S, const double* D /* S x 7 */, const double* k /* S x 17 */,
                    double dr, double* dt /* S */);

/*
 * Solve S independent parameter sets.  HOST buffers; shards the sets over o->n_devices GPUs
 * (one host thread + one stream per device, no collective) and gathers into `out`.
 *
 * Stands in for the per-set loops around the reference solvers:
 *   run_ensemble(pdesolver|pdesolver_membSFK|pdesolver_rect, …)   get_param_posteriors.jl:135-168
 *   fbatch_dk_mt / fbatch_concs_mt → pmap_fun_* → sapdesolver[_membSFK]   sapdesolver.jl:330-387,432-476
 * and, with S = 1, for a single pdesolver / sapdesolver / pulsechase_solver call
 *   basepdesolver.jl:25-312,350-636; basepdesolver_rect.jl:23-294,298-569;
 *   sapdesolver.jl:55-280; sapdesolver_memb-SFK.jl:55-281; pulsechase_solver.jl:29-318.
 *
 *   Co        S x 5 row-major when Co_stride == 5; one shared Co[5] when Co_stride == 0
 *   D, k      S x 7, S x 17 row-major
 *   dt        S time steps (the caller evaluates the keyword default so that the Julia value is used)
 *   r         Nr+1 grid nodes, the caller's collect(0.0:dr:R)
 *   out       S * gab1_out_doubles_per_set(o) doubles, set-major
 *   status    S words of GAB1_ST_* bits              (may be NULL)
 *   n_saved   S: snapshot columns written incl. the initial one (FULL/PCT)   (may be NULL)
 *   n_steps   S: time steps taken (= ceil(tf/dt))     (may be NULL)
 *   n_bc_iters S: total membrane fixed-point iterations (may be NULL)
 */
int gab1_solve_batch(const gab1_opts* o, int64_t S,
                     const double* Co, int64_t Co_stride,
                     const double* D, const double* k, const double* dt, const double* r,
                     double* out, int32_t* status, int32_t* n_saved,
                     int64_t* n_steps, int64_t* n_bc_iters);

/*
 * gab1_solve_batch with the 1e-9 contract certified for every non-diverging set, ill-conditioned ones included (opt-in,
 * about 2.3x the cost).  The fast kernels sit ~100 ulps of accumulated rounding from the reference's arithmetic; a set whose
 * final state responds to a ONE-ulp change of the initial concentrations by `response` or more (relative; <= 0 selects the
 * default 1e-12), or whose control flow responds at all, is re-solved with the strict kernels (arith = 1, every operation
 * of basepdesolver.jl:150-242 in source order) and its rows replaced.  Same buffers as gab1_solve_batch, plus
 *   resolved    S flags, 1 = the set was re-solved strictly   (may be NULL)
 *   n_resolved  their number                                  (may be NULL)
 * With o->arith == 1 this is gab1_solve_batch.
 */
int gab1_solve_batch_certified(const gab1_opts* o, int64_t S,
                               const double* Co, int64_t Co_stride,
                               const double* D, const double* k, const double* dt, const double* r,
                               double* out, int32_t* status, int32_t* n_saved,
                               int64_t* n_steps, int64_t* n_bc_iters,
                               double response, int32_t* resolved, int64_t* n_resolved);

/*
 * Same computation with every buffer already resident on CUDA device `device`
 * (device pointers), enqueued on `stream` (a cudaStream_t passed as void*; NULL = default
 * stream) without synchronising.  `workspace` must hold gab1_workspace_bytes(S) bytes of
 * device memory.  This is the entry point the resident-throughput benchmark times.
 */
size_t gab1_workspace_bytes(int64_t S);
int gab1_solve_batch_device(const gab1_opts* o, int32_t device, void* stream, int64_t S,
                            const double* Co, int64_t Co_stride,
                            const double* D, const double* k, const double* dt, const double* r,
                            double* out, int32_t* status, int32_t* n_saved,
                            int64_t* n_steps, int64_t* n_bc_iters,
                            void* workspace);

/* The sharding gab1_solve_batch applies (no collective: the parameter sets are independent, get_param_posteriors.jl:147,
 * sapdesolver.jl:377).  Pure host code; bounds has n_shards+1 entries.
 *   gab1_plan_shards  full snapshot output: bounds[g]..bounds[g+1] is the contiguous range of set indices device g solves,
 *                     balanced by total step count ceil(tf/dt), so that a device's result is one block of `out`;
 *   gab1_deal_shards  small per-set outputs (FINAL4 / SIX / PCT_BOUND / FINAL_STATE): the sets in descending step-count
 *                     order go one by one to the least-loaded shard (LPT) — equal total work AND the same mix of long, short and
 *                     diverging solves per device; perm[bounds[g] .. bounds[g+1]) are shard g's set indices (ascending). */
int gab1_plan_shards(int64_t S, const double* dt, double tf, int32_t n_shards, int64_t* bounds);
int gab1_deal_shards(int64_t S, const double* dt, double tf, int32_t n_shards, int64_t* perm, int64_t* bounds);

/*
 * Order statistics across the parameter sets of an ensemble — what the reference's figure scripts compute from the stacked
 * full solutions: `median(stack, dims=3)` for the whole surface and `quantile(stack[node, end, :], 0.5 -+ 0.341)` for the
 * credible band at t = tf (run_base_model.jl:103-174).  For every selected matrix, every snapshot column c in [c0, c1) and
 * every node, the values of the sets WITHOUT GAB1_ST_NAN (run_ensemble drops the others, get_param_posteriors.jl:155) are
 * sorted on the device and the np statistics p[] are taken with the definitions of Julia's Statistics stdlib:
 * p[i] in [0, 1] -> quantile(v, p[i]) (default alpha = beta = 1); p[i] < 0 -> median(v) (middle(a, b) = a/2 + b/2 for an even
 * count, which differs from quantile(v, 0.5) in the last bit).
 *   matrices   bit m set => matrix m (GAB1_M_*), a subset of o->matrix_mask; o->out_mode must be GAB1_OUT_FULL
 *   q          [matrix (ascending m)][i][c - c0][node], doubles
 *   n_valid    number of sets that entered the statistics
 * At most 16384 sets per call (one row of all sets, padded to a power of two, is sorted in 200 KB of shared memory).
 *
 * gab1_solve_ensemble_quantiles: HOST buffers; solves the ensemble on one GPU (o->device_ids[0], default 0), keeps the full
 * result in HBM and returns only the statistics and the per-set diagnostics — 0.5 MB per set never crosses PCIe.
 * gab1_ensemble_quantiles_device: the reduction alone, on a FULL block already resident on `device` (device pointers except
 * `p`; `workspace` holds gab1_quantiles_workspace_bytes(S) bytes; `n_valid` is a device pointer), enqueued on `stream`.
 */
size_t gab1_quantiles_workspace_bytes(int64_t S);
int gab1_ensemble_quantiles_device(const gab1_opts* o, int32_t device, void* stream, int64_t S,
                                   const double* out, const int32_t* status, uint32_t matrices,
                                   int32_t c0, int32_t c1, int32_t np, const double* p,
                                   double* q, int64_t* n_valid, void* workspace);
int gab1_solve_ensemble_quantiles(const gab1_opts* o, int64_t S,
                                  const double* Co, int64_t Co_stride,
                                  const double* D, const double* k, const double* dt, const double* r,
                                  uint32_t matrices, int32_t c0, int32_t c1, int32_t np, const double* p,
                                  double* q, int32_t* status, int32_t* n_saved,
                                  int64_t* n_steps, int64_t* n_bc_iters, int64_t* n_valid);

/*
 * Forward-mode derivatives ("tangents") of the solve — what the reference obtains by running pdesolver_fitting on
 * ForwardDiff dual numbers (basepdesolver.jl:674-932 is generic in T): ForwardDiff.gradient(testf, …)
 * param_fitting+inference_finitediff.jl:128-151, the LBFGS loss under AutoForwardDiff :188-240, NUTS through
 * turing_model :308-370.  Along each of n_dir input directions the partials of EVERY input are given as one seed row
 *     seeds[set][d][0..29] = d/d(dir d) of [ D(7) ; k(17) ; Co(5) ; dt ]      (pdesolver_fitting's packed order, then dt)
 * — dt is computed inside pdesolver_fitting from p (:696) and so carries partials; gab1_default_dt_tangent evaluates
 * dt and fills slot 29 of every seed row by the same rules.  Decisions (loop exits, snapshot tests, ceil) look at the
 * values alone, exactly as comparisons on dual numbers do, so status / n_saved / n_steps / n_bc_iters are the primal's.
 *   out      S * (1 + n_dir) * gab1_out_doubles_per_set(o) doubles; per set: block 0 = values (same layout and
 *            numbers as gab1_solve_batch), block 1 + d = partials of every output along direction d
 * Supported: bc_loop = GAB1_BC_FOR_BREAK with maxiters >= 1, save_rule = GAB1_SAVE_T_GE_TSAVE, t_prechase < 0,
 * out_mode FULL / FINAL4 / PCT_BOUND / FINAL_STATE (GAB1_OUT_SIX is piecewise constant in the parameters),
 * Nr <= 128, 1 <= n_dir <= 64.  Anything else fails loudly (negative return).  `arith` is ignored (fast forms).
 */
#define GAB1_N_SEED 30
int gab1_default_dt_tangent(int64_t S, int32_t n_dir, const double* D, const double* k, double dr,
                            double* dt /* S */, double* seeds /* S x n_dir x 30, slot 29 written */);
int gab1_solve_tangent(const gab1_opts* o, int64_t S, int32_t n_dir,
                       const double* Co, int64_t Co_stride,
                       const double* D, const double* k, const double* dt,
                       const double* seeds, const double* r,
                       double* out, int32_t* status, int32_t* n_saved,
                       int64_t* n_steps, int64_t* n_bc_iters);
/* device-resident twin (device pointers, enqueued on `stream`, workspace of gab1_workspace_bytes(S) bytes) */
int gab1_solve_tangent_device(const gab1_opts* o, int32_t device, void* stream, int64_t S, int32_t n_dir,
                              const double* Co, int64_t Co_stride,
                              const double* D, const double* k, const double* dt,
                              const double* seeds, const double* r,
                              double* out, int32_t* status, int32_t* n_saved,
                              int64_t* n_steps, int64_t* n_bc_iters, void* workspace);

/*
 * Synthetic prior ensembles: the prior half of generate_ensemble (get_param_posteriors.jl:38-86) — per set,
 * 22 independent log-normal draws exp(mu[i] + sigma[i] * z) in the order
 *     D(7) ; Kd_S2, kS2r, Kd_G2, kG2r, kG1f, kG1r, kEGFf, kEGFr, kdf ; kG1p, kG1dp, kSa, kSi, kp, kdp
 * assembled into D[7] and k[17] with k_f = k_r / Kd for the SHP2 and GRB2 pairs (:75-76), k[14] = EGF, kdr = kdf * Kdd.
 * The stream is the library's own (the reference draws from Julia's default RNG): Philox4x32-10 with counter
 * (set index, draw pair) and key = seed, two 53-bit uniforms per call, Box-Muller — any host can restate it
 * (oracle/sampler_oracle.py).  _device: D and k are device pointers, the kernel is enqueued on `stream`; the host
 * variant fills host buffers.
 */
#define GAB1_N_PRIOR_NORMALS 22
int gab1_sample_prior_device(int32_t device, void* stream, int64_t S, uint64_t seed, const double* mu /* 22, host */,
                             const double* sigma /* 22, host */, double EGF, double Kdd,
                             double* D /* S x 7 */, double* k /* S x 17 */);
int gab1_sample_prior(int64_t S, uint64_t seed, const double* mu, const double* sigma, double EGF, double Kdd,
                      double* D, double* k);

/* gab1_solve_batch keeps one slab of device memory and one stream per GPU between calls (no cudaMalloc/cudaFree per
 * call; calls that target the same GPU take turns).  This frees them; the next call re-creates what it needs. */
void gab1_release_device_memory(void);

/* Pinned, device-mapped host allocations.  When `out` of gab1_solve_batch lives in such memory the kernels write the
 * snapshots straight into it while the time loop runs (no device copy of the output, no D2H phase). */
void* gab1_host_alloc(size_t bytes);
void gab1_host_free(void* p);
/* Same, with the pages placed on the NUMA node the GPU `device` is attached to (the calling thread is moved onto that
 * node's CPUs while the pages are touched); falls back to gab1_host_alloc where the platform exposes no topology.
 * gab1_device_numa_node: that node, or -1.  GAB1_NUMA=0 in the environment disables the placement. */
void* gab1_host_alloc_near(size_t bytes, int32_t device);
int gab1_device_numa_node(int32_t device);

/* Sustained FP64 FMA rate of `device` in TFLOP/s (2 flop per DFMA), measured by a register-resident
 * DFMA kernel; the roofline denominator bench.py reports. Negative on failure. */
double gab1_measure_fp64_tflops(int32_t device, double seconds);

/* Diagnostic: worst relative error of the hardware reciprocal seed and of the kernel's reciprocal (seed + one cubic
 * step) over 4M log-spaced operands in [lo, hi]. */
int gab1_debug_recip_error(int32_t device, double lo, double hi, double* seed_err, double* recip_err);

/* Number of launches of this library's kernels since load (for the benchmark's gpu_launches). */
int64_t gab1_kernel_launches(void);

int gab1_device_count(void);
int gab1_version(void);
const char* gab1_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* GAB1PDE_H */
