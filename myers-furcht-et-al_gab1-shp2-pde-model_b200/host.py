"""Host-side mirror of the reference's Julia call surface, on top of the C ABI.

Every function keeps the name, argument meaning, keyword defaults and return shape of the Julia function
it mirrors (file:line cited per function; all under the reference's Julia/ directory), so that the parity
tests read like the reference's own call sites.  The Julia twin of this file is julia/gab1pde_dropin.jl.

`Frontend(backend)` binds the surface to a backend object with a `.solve(opts, Co, D, k, dt, r)` method.
The module-level functions are bound to the CUDA library (abi.CudaBackend); tests bind a second Frontend
to the CPU oracle to check it.  The product never does.
"""
from __future__ import annotations

from collections import namedtuple
from dataclasses import dataclass

import numpy as np

from . import abi, params

Sol21 = namedtuple("Sol21", abi.MATRIX_NAMES + abi.VECTOR_NAMES[:9])                       # basepdesolver.jl:303-310
Sol22 = namedtuple("Sol22", abi.MATRIX_NAMES + ("EGFR_SHP2",) + abi.VECTOR_NAMES[:9])       # basepdesolver_rect.jl:282-290
SolFit = namedtuple("SolFit", ("aSFK", "PG1S", "G2PG1S", "EG2PG1S"))                        # basepdesolver.jl:929
SolSA = namedtuple("SolSA", ("iSFK", "aSFK", "PG1tot", "PG1Stot"))                          # sapdesolver.jl:270-278
EnsembleRow = namedtuple("EnsembleRow", ("r", "t_sol", "sol", "index"))                     # get_param_posteriors.jl:158


@dataclass
class BatchResult:
    """What a *_batch sibling returns: the raw set-major output block plus per-set diagnostics."""
    opts: abi.Opts
    out: np.ndarray          # (S, doubles_per_set)
    status: np.ndarray       # (S,) GAB1_ST_* bits
    n_saved: np.ndarray      # (S,) snapshot columns written
    n_steps: np.ndarray      # (S,) Nt
    n_bc_iters: np.ndarray   # (S,) membrane fixed-point iterations
    r: np.ndarray
    dt: np.ndarray
    resolved_strict: np.ndarray = None    # certify=True: the sets that were re-solved with the strict kernels

    def matrix(self, name: str) -> np.ndarray:
        """(S, Nr+1, Nts+1) view of one FULL matrix (Julia layout: node index fastest)."""
        o = self.opts
        off = abi.full_matrix_offset(o, abi.MATRIX_NAMES.index(name))
        if off < 0:
            raise KeyError(f"{name} was masked out")
        P, Cn = o.Nr + 1, o.Nts + 1
        return self.out[:, off:off + P * Cn].reshape(-1, Cn, P).transpose(0, 2, 1)

    def vector(self, name: str) -> np.ndarray:
        o = self.opts
        off = abi.full_vector_offset(o, abi.VECTOR_NAMES.index(name))
        return self.out[:, off:off + o.Nts + 1]


@dataclass
class TangentResult:
    """Values and forward-mode partials of a batch: out[:, 0] is what the primal call returns, out[:, 1 + d] the
    partials of every output along seed direction d (same layout)."""
    opts: abi.Opts
    out: np.ndarray          # (S, 1 + n_dir, doubles_per_set)
    status: np.ndarray
    n_saved: np.ndarray
    n_steps: np.ndarray
    n_bc_iters: np.ndarray
    r: np.ndarray
    dt: np.ndarray           # (S,)
    seeds: np.ndarray        # (S, n_dir, 30) as used (slot 29 = partial of dt)

    def matrix(self, name: str) -> np.ndarray:
        """(S, 1 + n_dir, Nr+1, Nts+1)"""
        o = self.opts
        off = abi.full_matrix_offset(o, abi.MATRIX_NAMES.index(name))
        if off < 0:
            raise KeyError(f"{name} was masked out")
        P, Cn = o.Nr + 1, o.Nts + 1
        return self.out[:, :, off:off + P * Cn].reshape(self.out.shape[0], -1, Cn, P).transpose(0, 1, 3, 2)

    def vector(self, name: str) -> np.ndarray:
        """(S, 1 + n_dir, Nts+1)"""
        o = self.opts
        off = abi.full_vector_offset(o, abi.VECTOR_NAMES.index(name))
        return self.out[:, :, off:off + o.Nts + 1]


def _grid(R, dr, r):
    return params.julia_range(dr, R) if r is None else np.asarray(r, dtype=np.float64)


class Frontend:
    def __init__(self, backend):
        self.backend = backend

    # ------------------------------------------------------------------ batched siblings (new)
    def pdesolver_batch(self, Co, Dmat, kmat, *, R=10.0, dr=0.1, tf=5.0, Nts=100, dt=None, dt_save=None,
                        maxiters=100, tol=1e-6, geometry=abi.GEOM_SPHERICAL, sfk_mode=abi.SFK_DIFFUSIBLE,
                        save_rule=abi.SAVE_T_GE_TSAVE, pg1tot_form=abi.PG1TOT_VIA_STOT, matrices=None,
                        t_prechase=-1.0, out_mode=abi.OUT_FULL, pct_mul=1.0, pct_div=1.0, r=None, certify=False) -> BatchResult:
        """Batched sibling of pdesolver (basepdesolver.jl:25-312): one row of Dmat/kmat per parameter set.
        certify=True: sets whose answer responds to a one-ulp input change are re-solved with the strict kernels (_certify)."""
        mask = abi.MASK_ALL if matrices is None else sum(1 << abi.MATRIX_NAMES.index(m) for m in matrices)
        o = abi.make_opts(R=R, dr=dr, tf=tf, Nts=Nts, dt_save=dt_save, maxiters=maxiters, tol=tol, geometry=geometry,
                          sfk_mode=sfk_mode, bc_loop=abi.BC_FOR_BREAK, save_rule=save_rule, pg1tot_form=pg1tot_form,
                          out_mode=out_mode, matrix_mask=mask, t_prechase=t_prechase, pct_mul=pct_mul, pct_div=pct_div)
        return self._run(o, Co, Dmat, kmat, dt, dr, _grid(R, dr, r), certify)

    def sapdesolver_batch(self, Co, Dmat, kmat, *, R=10.0, dr=0.2, tf=5.0, dt=None, maxiters=20, tol=1e-3,
                          membSFK=False, out_mode=abi.OUT_FINAL4, r=None, iter_cap=2000, certify=False) -> BatchResult:
        """Batched sibling of sapdesolver / sapdesolver_membSFK (sapdesolver.jl:55-280, sapdesolver_memb-SFK.jl:55-281)."""
        o = abi.make_opts(R=R, dr=dr, tf=tf, Nts=1, maxiters=maxiters, tol=tol, out_mode=out_mode,
                          sfk_mode=abi.SFK_MEMBRANE if membSFK else abi.SFK_DIFFUSIBLE,
                          bc_loop=abi.BC_WHILE if membSFK else abi.BC_FOR_BREAK,
                          pg1tot_form=abi.PG1TOT_CHAIN if membSFK else abi.PG1TOT_VIA_STOT)
        if membSFK:
            # `maxiters` is accepted but unused by the while-loop form (sapdesolver_memb-SFK.jl:58,177): a fixed point
            # that never meets `tol` spins forever in the reference.  The library needs a finite cap to stay bounded
            # and reports GAB1_ST_ITER_CAP for a set that reaches it.  The default comes from a sweep over caps
            # (tools/cap_sweep.py, profiles/r2_cap_sweep_*.jsonl): on 20 000 HeLa sets (posterior-like and wide-prior) the
            # SAME sets are flagged at every cap from 20 to 100 000 — a step that converges at all does so in fewer than
            # 20 passes — and the pass time is flat up to ~5000; 2000 leaves a 100x margin (2.5-2.7 % of the sets reach it)
            o.maxiters = int(iter_cap)
        return self._run(o, Co, Dmat, kmat, dt, dr, _grid(R, dr, r), certify)

    def ensemble_quantiles(self, ensemble, Co, *, probs=("median", 0.5 - 0.341, 0.5 + 0.341),
                           matrices=("aSFK", "PG1tot", "PG1Stot"), columns=None, dr=0.2, R=10.0, tf=5.0, Nts=100, tol=1e-4,
                           maxit=20, D_inds=slice(0, 7), k_inds=slice(7, 24), geometry=abi.GEOM_SPHERICAL,
                           sfk_mode=abi.SFK_DIFFUSIBLE, pg1tot_form=abi.PG1TOT_VIA_STOT):
        """The summary surfaces of run_base_model.jl:103-174 without the full solutions leaving the device: for each
        matrix, `median(stack, dims=3)` / `quantile(stack[node, column, :], p)` over the sets run_ensemble would keep
        (NaN sets dropped, get_param_posteriors.jl:155), run_ensemble's solver defaults.  Returns
        ({name: array (len(probs), n_columns, Nr+1)}, n_valid, r, status); `columns` = (c0, c1), default all Nts+1."""
        ensemble = np.asarray(ensemble, dtype=np.float64)
        mask = sum(1 << abi.MATRIX_NAMES.index(m) for m in matrices)
        o = abi.make_opts(R=R, dr=dr, tf=tf, Nts=Nts, maxiters=maxit, tol=tol, geometry=geometry, sfk_mode=sfk_mode,
                          pg1tot_form=pg1tot_form, out_mode=abi.OUT_FULL, matrix_mask=mask)
        c0, c1 = (0, Nts + 1) if columns is None else columns
        Dmat = np.ascontiguousarray(ensemble[:, D_inds])
        kmat = np.ascontiguousarray(ensemble[:, k_inds])
        dtv = params.default_dt(Dmat, kmat, dr)
        r = params.julia_range(dr, R)
        q, n_valid, status, *_ = self.backend.solve_quantiles(o, Co, Dmat, kmat, dtv, r, mask, c0, c1, probs)
        order = sorted(matrices, key=abi.MATRIX_NAMES.index)
        return {name: q[i] for i, name in enumerate(order)}, n_valid, r, status

    @staticmethod
    def regrid(values, r, dr_new=0.1, R=10.0):
        """`linear_interpolation(r, y)(0:dr_new:R)` along the node axis (the last one) — the re-gridding run_base_model.jl
        applies to the summary surfaces (:108-119, 134-145, 161-172).  A host-side operation on the ~0.5 MB of statistics
        ensemble_quantiles returns.  Interpolations.jl is not in the reference tree: its linear rule is restated as
        (1 - w) * y[i] + w * y[i+1] with w = (x - r[i]) / (r[i+1] - r[i]) (unpinned in the last bit).
        Returns (values on the new grid, the new grid)."""
        r = np.asarray(r, dtype=np.float64)
        x = params.julia_range(dr_new, R)
        i = np.clip(np.searchsorted(r, x, side="right") - 1, 0, len(r) - 2)
        w = (x - r[i]) / (r[i + 1] - r[i])
        v = np.asarray(values, dtype=np.float64)
        return (1.0 - w) * v[..., i] + w * v[..., i + 1], x

    def _run(self, o, Co, Dmat, kmat, dt, dr, r, certify=False) -> BatchResult:
        Dmat = np.ascontiguousarray(Dmat, dtype=np.float64).reshape(-1, abi.N_D)
        kmat = np.ascontiguousarray(kmat, dtype=np.float64).reshape(-1, abi.N_K)
        dtv = params.default_dt(Dmat, kmat, dr) if dt is None else np.broadcast_to(np.asarray(dt, float), (Dmat.shape[0],)).copy()
        if certify and hasattr(self.backend, "solve_certified"):       # the C ABI's own entry point (gab1_solve_batch_certified)
            out, status, n_saved, n_steps, n_bc, idx = self.backend.solve_certified(o, Co, Dmat, kmat, dtv, r, self.CERTIFY_RESPONSE)
            res = BatchResult(o, out, status, n_saved, n_steps, n_bc, r, dtv)
            res.resolved_strict = idx
            return res
        out, status, n_saved, n_steps, n_bc = self.backend.solve(o, Co, Dmat, kmat, dtv, r)
        res = BatchResult(o, out, status, n_saved, n_steps, n_bc, r, dtv)
        if certify and hasattr(self.backend, "strict_twin"):
            self._certify(res, Co, Dmat, kmat)
        return res

    # one-ulp response above which a set is re-solved strictly: the fast kernels differ from the reference's arithmetic by
    # ~100 ulps of accumulated rounding, so a response of 1e-12 to ONE ulp keeps their distance below 1e-9 with a margin of ten
    CERTIFY_RESPONSE = 1e-12

    def _certify(self, res: BatchResult, Co, Dmat, kmat) -> None:
        """Python restatement of gab1_solve_batch_certified (the C ABI entry point `certify=True` uses; this one is kept as
        its cross-check, tests/test_gpu_census.py).
        Makes the 1e-9 contract hold for EVERY non-diverging set, ill-conditioned ones included (opt-in, ~2.3x the cost).

        The explicit scheme's time step ignores the second-order rate x concentration terms (basepdesolver.jl:30), so a few
        per mille of wide prior draws sit at the edge of stability: an alternating mode amplifies last-bit differences by
        1e5..1e13 over the 4e4 steps without blowing up, and no arithmetic but the reference's own reproduces the reference
        there (tests/test_gpu_census.py).  Those sets are found by their response to a one-ulp change of the initial
        concentrations — two fast solves of the final state — and re-solved with the strict kernels (arith = 1: every
        operation of the reference in source order, bit-identical to the oracle), whose results replace the fast ones."""
        o = res.opts
        fs = abi.Opts.from_buffer_copy(bytes(o))
        fs.device_ids = None
        fs.out_mode = abi.OUT_FINAL_STATE
        Co = np.ascontiguousarray(Co, dtype=np.float64)

        def final_state(Cox):
            if o.out_mode == abi.OUT_FINAL_STATE and Cox is Co:
                return res.out, res.status, res.n_bc_iters
            out, status, _, _, n_bc = self.backend.solve(abi.Opts.from_buffer_copy(bytes(fs)), Cox, Dmat, kmat, res.dt, res.r)
            return out, status, n_bc

        a, sa, ba = final_state(Co)
        b, sb, bb = final_state(np.nextafter(Co, np.inf))
        fin = np.isfinite(a)
        scale = np.where(fin, np.abs(a), 0).max(axis=1, keepdims=True)
        den = np.maximum(np.abs(a), 1e-6 * scale)
        with np.errstate(invalid="ignore", divide="ignore"):
            e = np.where(fin & (den > 0), np.abs(b - a) / den, 0.0)
        e = np.where(np.isnan(a) != np.isnan(b), np.inf, e).max(axis=1)
        both_diverge = ((sa & abi.ST_NAN) != 0) & ((sb & abi.ST_NAN) != 0)          # NaN either way: dropped by every caller
        flagged = ((e >= self.CERTIFY_RESPONSE) | (ba != bb) | (sa != sb)) & ~both_diverge
        idx = np.flatnonzero(flagged)
        res.resolved_strict = idx
        if len(idx) == 0:
            return
        so = abi.Opts.from_buffer_copy(bytes(o))
        so.device_ids = None
        Cos = Co if Co.ndim == 1 else np.ascontiguousarray(Co[idx])
        out, status, n_saved, n_steps, n_bc = self.backend.strict_twin().solve(so, Cos, Dmat[idx], kmat[idx], res.dt[idx], res.r)
        res.out[idx] = out
        res.status[idx], res.n_saved[idx], res.n_steps[idx], res.n_bc_iters[idx] = status, n_saved, n_steps, n_bc

    # ------------------------------------------------------------------ single solves (reference names)
    def _single_full(self, Co, D, k, kw, *, trim=False, extra=False, **fixed):
        res = self.pdesolver_batch(Co, np.asarray(D, float)[None, :], np.asarray(k, float)[None, :], **kw, **fixed)
        if res.status[0] & abi.ST_THROW:
            raise ArithmeticError("InexactError: Int64(ceil(tf/dt))")      # basepdesolver.jl:72
        if res.status[0] & abi.ST_OVERFLOW:
            # basepdesolver.jl:271 raises BoundsError; the rect solvers grow their outputs instead
            # (basepdesolver_rect.jl:250-279) — the library's block holds Nts+1 columns, so it refuses rather than truncates
            raise IndexError("more than Nts+1 snapshots are due (BoundsError in pdesolver; beyond the library's fixed-size "
                             "output for the rect solvers)")
        ncol = int(res.n_saved[0]) if trim else res.opts.Nts + 1
        mats = [res.matrix(n)[0][:, :ncol] for n in abi.MATRIX_NAMES]
        vecs = [res.vector(n)[0][:ncol] for n in abi.VECTOR_NAMES]
        t_out = vecs[10]
        if extra:
            sol = Sol22(*mats, vecs[9], *vecs[:9])
        else:
            sol = Sol21(*mats, *vecs[:9])
        return sol, res.r, t_out, float(res.dt[0]), res

    def pdesolver(self, Co, D, k, *, R=10.0, dr=0.1, tf=5.0, Nts=100, dt=None, dt_save=None, maxiters=100, tol=1.0e-6):
        """pdesolver (basepdesolver.jl:25-312) -> (sol, r, t_out, dt)."""
        kw = dict(R=R, dr=dr, tf=tf, Nts=Nts, dt=dt, dt_save=dt_save, maxiters=maxiters, tol=tol)
        return self._single_full(Co, D, k, kw)[:4]

    def pdesolver_membSFK(self, Co, D, k, *, R=10.0, dr=0.1, tf=5.0, Nts=100, dt=None, dt_save=None, maxiters=20, tol=1.0e-6):
        """pdesolver_membSFK (basepdesolver.jl:350-636): aSFK diffusivity 1e-32."""
        kw = dict(R=R, dr=dr, tf=tf, Nts=Nts, dt=dt, dt_save=dt_save, maxiters=maxiters, tol=tol)
        return self._single_full(Co, D, k, kw, sfk_mode=abi.SFK_MEMBRANE)[:4]

    def pdesolver_rect(self, Co, D, k, *, R=10.0, dr=0.1, tf=5.0, Nts=100, dt=None, dt_save=None, maxiters=20, tol=1.0e-6):
        """pdesolver_rect (basepdesolver_rect.jl:23-294): planar Laplacian; outputs hold 1 + #snapshots columns."""
        kw = dict(R=R, dr=dr, tf=tf, Nts=Nts, dt=dt, dt_save=dt_save, maxiters=maxiters, tol=tol)
        return self._single_full(Co, D, k, kw, trim=True, extra=True, geometry=abi.GEOM_RECT,
                                 pg1tot_form=abi.PG1TOT_CHAIN)[:4]

    def pdesolver_membSFK_rect(self, Co, D, k, *, R=10.0, dr=0.1, tf=5.0, Nts=100, dt=None, maxiters=20, tol=1.0e-6):
        """pdesolver_membSFK_rect (basepdesolver_rect.jl:298-569): 8-element D (D[3:8] used, :307-312), both SFK
        diffusivities 1e-32 (:305-306), modulus snapshot rule (:336,526).  The default dt uses maximum(D) over all 8."""
        D = np.asarray(D, float)
        if D.shape[0] != 8:
            raise IndexError("BoundsError: pdesolver_membSFK_rect indexes D[8]")
        k = np.asarray(k, float)
        if dt is None:
            sk = 0.0
            for v in k:
                sk = sk + v
            dt = 1.0 / (2.0 * (D.max() / (dr * dr) + sk / 4)) * 0.99
        D7 = np.concatenate([D[:1], D[2:]])
        kw = dict(R=R, dr=dr, tf=tf, Nts=Nts, dt=dt, dt_save=None, maxiters=maxiters, tol=tol)
        return self._single_full(Co, D7, k, kw, trim=True, extra=True, geometry=abi.GEOM_RECT,
                                 sfk_mode=abi.SFK_BOTH_FROZEN, save_rule=abi.SAVE_MODULUS,
                                 pg1tot_form=abi.PG1TOT_CHAIN)[:4]

    def pdesolver_fitting(self, p, *, Diff_inds=range(0, 7), k_inds=None, Co_inds=None, R=10.0, dr=0.1, tf=5.0,
                          Nts=100, dt_save=None, maxiters=20, tol=1.0e-6):
        """Float64 path of pdesolver_fitting (basepdesolver.jl:674-932): p = [D; k; Co], dt computed inside (:696),
        outputs aSFK, PG1S, G2PG1S matrices and the EG2PG1S vector; dummy zeros when Nt overflows (:730-735)."""
        p = np.asarray(p, float)
        Diff_inds = list(Diff_inds)
        k_inds = [Diff_inds[-1] + 1 + i for i in range(17)] if k_inds is None else list(k_inds)
        Co_inds = [k_inds[-1] + 1 + i for i in range(5)] if Co_inds is None else list(Co_inds)
        D, k, Co = p[Diff_inds], p[k_inds], p[Co_inds]
        res = self.pdesolver_batch(Co, D[None, :], k[None, :], R=R, dr=dr, tf=tf, Nts=Nts, dt_save=dt_save,
                                   maxiters=maxiters, tol=tol, matrices=("aSFK", "PG1S", "G2PG1S"))
        if res.status[0] & abi.ST_THROW:
            z = np.zeros((10, 10))
            return SolFit(None, z, z.copy(), z.copy()), np.ones(10), np.ones(10), float(res.dt[0])
        sol = SolFit(res.matrix("aSFK")[0], res.matrix("PG1S")[0], res.matrix("G2PG1S")[0], res.vector("EG2PG1S")[0])
        return sol, res.r, res.vector("t_out")[0], float(res.dt[0])

    # ------------------------------------------------------------------ forward-mode (ForwardDiff.Dual) path
    def pdesolver_tangent_batch(self, Co, Dmat, kmat, seeds, *, R=10.0, dr=0.1, tf=5.0, Nts=100, dt=None, dt_save=None,
                                maxiters=20, tol=1e-6, geometry=abi.GEOM_SPHERICAL, sfk_mode=abi.SFK_DIFFUSIBLE,
                                pg1tot_form=abi.PG1TOT_VIA_STOT, matrices=None, out_mode=abi.OUT_FULL, pct_mul=1.0,
                                pct_div=1.0, r=None) -> TangentResult:
        """The solver on dual numbers, batched: what `pdesolver_fitting(p::Vector{<:ForwardDiff.Dual})` computes
        (basepdesolver.jl:674-932), for S parameter sets and n_dir partials each.  seeds: (S, n_dir, 30), partials of
        [D; k; Co; dt].  dt=None: dt and its partials are computed from D, k inside, as pdesolver_fitting does (:696);
        a given dt is a constant unless the caller filled seed slot 29."""
        mask = abi.MASK_ALL if matrices is None else sum(1 << abi.MATRIX_NAMES.index(m) for m in matrices)
        o = abi.make_opts(R=R, dr=dr, tf=tf, Nts=Nts, dt_save=dt_save, maxiters=maxiters, tol=tol, geometry=geometry,
                          sfk_mode=sfk_mode, pg1tot_form=pg1tot_form, out_mode=out_mode, matrix_mask=mask, pct_mul=pct_mul,
                          pct_div=pct_div)
        Dmat = np.ascontiguousarray(Dmat, dtype=np.float64).reshape(-1, abi.N_D)
        kmat = np.ascontiguousarray(kmat, dtype=np.float64).reshape(-1, abi.N_K)
        seeds = np.ascontiguousarray(seeds, dtype=np.float64)
        if dt is None:
            dtv, seeds = self.backend.default_dt_tangent(Dmat, kmat, dr, seeds)
        else:
            dtv = np.broadcast_to(np.asarray(dt, float), (Dmat.shape[0],)).copy()
        rr = _grid(R, dr, r)
        out, status, n_saved, n_steps, n_bc = self.backend.solve_tangent(o, Co, Dmat, kmat, dtv, seeds, rr)
        return TangentResult(o, out, status, n_saved, n_steps, n_bc, rr, dtv, seeds)

    def pdesolver_fitting_dual(self, p, dp, *, R=10.0, dr=0.1, tf=5.0, Nts=100, dt_save=None, maxiters=20, tol=1.0e-6):
        """pdesolver_fitting on Dual inputs (basepdesolver.jl:674-932): p = [D; k; Co] (29,), dp (n_dir, 29) its partials.
        Returns ((sol, dsol), r, (t_out, dt_out), (dt, ddt)); dsol fields carry a leading n_dir axis."""
        p = np.asarray(p, float)
        dp = np.atleast_2d(np.asarray(dp, float))
        seeds = np.zeros((1, dp.shape[0], abi.N_SEED))
        seeds[0, :, :29] = dp
        res = self.pdesolver_tangent_batch(p[24:29], p[None, :7], p[None, 7:24], seeds, R=R, dr=dr, tf=tf, Nts=Nts,
                                           dt_save=dt_save, maxiters=maxiters, tol=tol,
                                           matrices=("aSFK", "PG1S", "G2PG1S"))
        if res.status[0] & abi.ST_THROW:
            z = np.zeros((10, 10))
            return (SolFit(None, z, z.copy(), z.copy()), None), np.ones(10), (np.ones(10), None), (float(res.dt[0]), res.seeds[0, :, 29])
        m = {n: res.matrix(n)[0] for n in ("aSFK", "PG1S", "G2PG1S")}
        v = res.vector("EG2PG1S")[0]
        t = res.vector("t_out")[0]
        sol = SolFit(m["aSFK"][0], m["PG1S"][0], m["G2PG1S"][0], v[0])
        dsol = SolFit(m["aSFK"][1:], m["PG1S"][1:], m["G2PG1S"][1:], v[1:])
        return (sol, dsol), res.r, (t[0], t[1:]), (float(res.dt[0]), res.seeds[0, :, 29].copy())

    def fitting_loss_and_gradient(self, pvals_in, mu, sigma, *, param_inds, pvals0, Co, R=10.0, dr=0.2, tf=5.0, Nts=100,
                                  tol=1e-3, maxiters=20):
        """`loss` of param_fitting+inference_finitediff.jl:188-226 and its ForwardDiff gradient (:238-240), for a batch of
        points: pvals_in (S, n) are log-parameters, x2[param_inds] = exp.(pvals_in) (:201-202, 0-based indices into
        [D; k] here), ŷ = % SHP2-bound GAB1 (:211-216), loss = (μ - ŷ)^2 / σ^2 (:219).
        Returns (loss (S,), grad (S, n), yhat (S,)); loss = inf where the reference returns Inf (NaN)."""
        x = np.atleast_2d(np.asarray(pvals_in, float))
        S, n = x.shape
        P = np.tile(np.asarray(pvals0, float), (S, 1))
        ex = np.exp(x)
        P[:, list(param_inds)] = ex
        seeds = np.zeros((S, n, abi.N_SEED))
        for i, j in enumerate(param_inds):
            seeds[:, i, j] = ex[:, i]
        volCF, surfCF = params.conversion_factors(R)
        res = self.pdesolver_tangent_batch(Co, P[:, :7], P[:, 7:24], seeds, R=R, dr=dr, tf=tf, Nts=Nts, tol=tol,
                                           maxiters=maxiters, out_mode=abi.OUT_PCT_BOUND, pct_mul=volCF, pct_div=surfCF)
        yhat = res.out[:, 0, 0]
        dy = res.out[:, 1:, 0]
        loss = (mu - yhat) ** 2 / sigma ** 2
        grad = (-2.0 * (mu - yhat) / sigma ** 2)[:, None] * dy
        bad = np.isnan(loss)
        loss = np.where(bad, np.inf, loss)
        return loss, grad, yhat

    def pulsechase_solver(self, Co, D, k, *, R=10.0, dr=0.1, t_prechase=5.0, t_chase=2.0, tf=None, Nts=100, dt=None,
                          dt_save=None, maxiters=20, tol=1.0e-6):
        """pulsechase_solver (pulsechase_solver.jl:29-318): pdesolver with kp := 0 once t >= t_prechase (:156-158);
        returns (sol, r, t_out, t_prechase, t_chase, dt_save) (:318)."""
        tf = t_prechase + t_chase if tf is None else tf
        kw = dict(R=R, dr=dr, tf=tf, Nts=Nts, dt=dt, dt_save=dt_save, maxiters=maxiters, tol=tol)
        sol, r, t_out, _, res = self._single_full(Co, D, k, kw, extra=True, t_prechase=t_prechase)
        return sol, r, t_out, t_prechase, t_chase, res.opts.dt_save

    def _single_sa(self, Co, D, k, membSFK, **kw):
        res = self.sapdesolver_batch(Co, np.asarray(D, float)[None, :], np.asarray(k, float)[None, :], membSFK=membSFK, **kw)
        if res.status[0] & abi.ST_THROW:
            raise ArithmeticError("InexactError: Int64(ceil(tf/dt))")      # sapdesolver.jl:90
        P = res.opts.Nr + 1
        o = res.out[0]
        return SolSA(o[:P], o[P:2 * P], o[2 * P:3 * P], o[3 * P:4 * P]), res.r

    def sapdesolver(self, Co, D, k, *, R=10.0, dr=0.2, tf=5.0, dt=None, maxiters=20, tol=1.0e-3):
        """sapdesolver (sapdesolver.jl:55-280); R, dr, tf default to the file-level globals (:11-14)."""
        return self._single_sa(Co, D, k, False, R=R, dr=dr, tf=tf, dt=dt, maxiters=maxiters, tol=tol)

    def sapdesolver_membSFK(self, Co, D, k, *, R=10.0, dr=0.2, tf=5.0, dt=None, maxiters=20, tol=1.0e-3):
        """sapdesolver_membSFK (sapdesolver_memb-SFK.jl:55-281)."""
        return self._single_sa(Co, D, k, True, R=R, dr=dr, tf=tf, dt=dt, maxiters=maxiters, tol=tol)

    # ------------------------------------------------------------------ ensemble drivers
    def run_ensemble(self, model_fun, ensemble, Co, *, dr=0.2, R=10.0, tf=5.0, Nts=100, tol=1e-4, maxit=20,
                     D_inds=range(0, 7), k_inds=range(7, 24), show_prog=True):
        """run_ensemble (get_param_posteriors.jl:135-168): full solutions for every row of `ensemble`, sets whose
        PG1S contains NaN dropped (:155); rows carry the set's `index`.  The Threads.@threads loop (:147) is one
        batched GPU call here.  model_fun is one of this frontend's pdesolver / pdesolver_membSFK / pdesolver_rect
        (or its name)."""
        name = model_fun if isinstance(model_fun, str) else model_fun.__name__
        variant = {"pdesolver": {}, "pdesolver_membSFK": dict(sfk_mode=abi.SFK_MEMBRANE),
                   "pdesolver_rect": dict(geometry=abi.GEOM_RECT, pg1tot_form=abi.PG1TOT_CHAIN)}[name]
        ensemble = np.asarray(ensemble, float)
        res = self.pdesolver_batch(Co, ensemble[:, list(D_inds)], ensemble[:, list(k_inds)], R=R, dr=dr, tf=tf, Nts=Nts,
                                   tol=tol, maxiters=maxit, **variant)
        rect = name == "pdesolver_rect"
        self._raise_like_the_reference(res)
        return self._rows(res, trim=rect, extra=rect)

    @staticmethod
    def _rows(res, *, trim, extra):
        """One EnsembleRow of views per set whose PG1S holds no NaN (get_param_posteriors.jl:155-161; 1-based `index`).  The 23
        per-set views come out of one zip over the (S, …) arrays — iterating an ndarray hands out its rows at C speed — which
        halves the host time of a 5000-set ensemble against indexing each array per set (94 -> 43 ms)."""
        arrs = [res.matrix(n) for n in abi.MATRIX_NAMES] + [res.vector(n) for n in abi.VECTOR_NAMES]
        nan = ((res.status & abi.ST_NAN) != 0).tolist()
        ncs = res.n_saved.tolist()
        r = res.r
        rows = []
        for j, f in enumerate(zip(*arrs)):
            if nan[j]:
                continue
            if trim:                       # the rect solvers' outputs hold 1 + #snapshots columns (basepdesolver_rect.jl:250-279)
                nc = ncs[j]
                f = tuple(a[:, :nc] for a in f[:12]) + tuple(a[:nc] for a in f[12:])
            sol = Sol22(*f[:12], f[21], *f[12:21]) if extra else Sol21(*f[:12], *f[12:21])
            rows.append(EnsembleRow(r, f[22], sol, j + 1))
        return rows

    @staticmethod
    def _raise_like_the_reference(res):
        """The reference's threaded loops let a solver exception out (get_param_posteriors.jl:147-163): an InexactError from
        Int64(ceil(tf/dt)), a BoundsError from a snapshot beyond column Nts+1.  (pdesolver_rect grows its outputs without
        bound instead, basepdesolver_rect.jl:250-279; the library's output block holds Nts+1 columns, so more snapshots
        than that is an error here rather than a silent truncation.)"""
        bad = np.flatnonzero(res.status & abi.ST_THROW)
        if len(bad):
            raise ArithmeticError(f"InexactError: Int64(ceil(tf/dt)) for set {int(bad[0]) + 1} (and {len(bad) - 1} more)")
        bad = np.flatnonzero(res.status & abi.ST_OVERFLOW)
        if len(bad):
            raise IndexError(f"more than Nts+1 snapshots are due for set {int(bad[0]) + 1} (and {len(bad) - 1} more)")

    def run_ensemble_pc(self, model_fun, ensemble, Co, *, dr=0.2, R=10.0, t_prechase=5.0, t_chase=2.0, Nts=100,
                        tol=1e-4, maxit=20, D_inds=range(0, 7), k_inds=range(7, 24)):
        """run_ensemble_pc (get_param_posteriors.jl:204-236) over pulsechase_solver."""
        ensemble = np.asarray(ensemble, float)
        res = self.pdesolver_batch(Co, ensemble[:, list(D_inds)], ensemble[:, list(k_inds)], R=R, dr=dr,
                                   tf=t_prechase + t_chase, Nts=Nts, tol=tol, maxiters=maxit, t_prechase=t_prechase)
        self._raise_like_the_reference(res)
        return self._rows(res, trim=False, extra=True)

    # ------------------------------------------------------------------ GSA batch functions
    def _six(self, Co, Dmat, kmat, *, R, dr, tf, tol, maxiters, membSFK, certify=False):
        res = self.sapdesolver_batch(Co, Dmat, kmat, R=R, dr=dr, tf=tf, tol=tol, maxiters=maxiters, membSFK=membSFK,
                                     out_mode=abi.OUT_SIX, certify=certify)
        return res

    def pmap_fun_dk(self, p, *, Co=None, D=None, kvals=None, R=10.0, dr=0.2, tf=5.0, maxiters=100, membSFK=False):
        """pmap_fun_dk (sapdesolver.jl:330-357): p = [D; k] -> 6 scalars.  Raises where the reference would."""
        p = np.asarray(p, float)
        Co = params.base_Co(R) if Co is None else Co
        res = self._six(Co, p[None, :7], p[None, 7:24], R=R, dr=dr, tf=tf, tol=1e-3, maxiters=maxiters, membSFK=membSFK)
        if res.status[0] & abi.ST_THROW:
            raise ValueError("ArgumentError: reducing over an empty collection is not allowed")
        return res.out[0].copy()

    def pmap_fun_allpars(self, p, *, R=10.0, dr=0.2, tf=5.0, membSFK=False):
        """pmap_fun_allpars (sapdesolver.jl:288-319): p = [Co; D; k]; solver defaults tol=1e-3, maxiters=20."""
        p = np.asarray(p, float)
        res = self._six(p[:5], p[None, 5:12], p[None, 12:29], R=R, dr=dr, tf=tf, tol=1e-3, maxiters=20, membSFK=membSFK)
        if res.status[0] & abi.ST_THROW:
            raise ValueError("ArgumentError: reducing over an empty collection is not allowed")
        return res.out[0].copy()

    def pmap_fun_dk_combD(self, p, *, Co=None, D=None, R=10.0, dr=0.2, tf=5.0, membSFK=False):
        """pmap_fun_dk_combD (sapdesolver.jl:391-421): p[1]/D[1] scales every diffusivity; p[2:18] are the k."""
        p = np.asarray(p, float)
        D = params.DIFFS_BASE if D is None else np.asarray(D, float)
        Co = params.base_Co(R) if Co is None else Co
        mult = p[0] / D[0]
        res = self._six(Co, (D * mult)[None, :], p[None, 1:18], R=R, dr=dr, tf=tf, tol=1e-3, maxiters=20, membSFK=membSFK)
        if res.status[0] & abi.ST_THROW:
            raise ValueError("ArgumentError: reducing over an empty collection is not allowed")
        return res.out[0].copy()

    def pmap_fun_concs(self, p, *, D=None, kvals=None, R=10.0, dr=0.2, tf=5.0, tol=1e-3, maxiters=20, membSFK=False):
        """pmap_fun_concs (sapdesolver.jl:432-451): p = Co."""
        D = params.DIFFS_BASE if D is None else D
        kvals = params.KVALS_BASE if kvals is None else kvals
        res = self._six(np.asarray(p, float), np.asarray(D, float)[None, :], np.asarray(kvals, float)[None, :],
                        R=R, dr=dr, tf=tf, tol=tol, maxiters=maxiters, membSFK=membSFK)
        if res.status[0] & abi.ST_THROW:
            raise ValueError("ArgumentError: reducing over an empty collection is not allowed")
        return res.out[0].copy()

    def fbatch_dk_mt(self, p_batch, *, numout=6, Co=None, D=None, kvals=None, R=10.0, dr=0.2, tf=5.0, maxiters=20,
                     membSFK=False, certify=False):
        """fbatch_dk_mt (sapdesolver.jl:371-387): p_batch is 24 x S in natural-log space; returns 6 x S; a column whose
        solve or reduction throws in the reference is zeros(6) (:378-382)."""
        P = np.exp(np.asarray(p_batch, float))
        Co = params.base_Co(R) if Co is None else Co
        res = self._six(Co, P[:7].T, P[7:24].T, R=R, dr=dr, tf=tf, tol=1e-3, maxiters=maxiters, membSFK=membSFK, certify=certify)
        return res.out.T.copy()

    def fbatch_concs_mt(self, p_batch, *, numout=6, Co=None, D=None, kvals=None, R=10.0, dr=0.2, tf=5.0, membSFK=False):
        """fbatch_concs_mt (sapdesolver.jl:460-476): p_batch is 5 x S log-space initial concentrations."""
        P = np.exp(np.asarray(p_batch, float))
        D = params.DIFFS_BASE if D is None else np.asarray(D, float)
        kvals = params.KVALS_BASE if kvals is None else np.asarray(kvals, float)
        S = P.shape[1]
        res = self._six(P.T.copy(), np.tile(D, (S, 1)), np.tile(kvals, (S, 1)), R=R, dr=dr, tf=tf, tol=1e-3, maxiters=20,
                        membSFK=membSFK)
        return res.out.T.copy()

    def pct_shp2_bound_gab1(self, Co, Dmat, kmat, *, R=10.0, dr=0.2, tf=5.0, Nts=100, tol=1e-4, maxiters=20):
        """% SHP2-bound GAB1 per set (run_base_model.jl:269-276) fused into the solve."""
        volCF, surfCF = params.conversion_factors(R)
        res = self.pdesolver_batch(Co, Dmat, kmat, R=R, dr=dr, tf=tf, Nts=Nts, tol=tol, maxiters=maxiters,
                                   out_mode=abi.OUT_PCT_BOUND, pct_mul=volCF, pct_div=surfCF)
        return res.out[:, 0].copy(), res


_default = Frontend(abi.CudaBackend())
for _n in ("pdesolver_batch", "sapdesolver_batch", "pdesolver", "pdesolver_membSFK", "pdesolver_rect",
           "pdesolver_membSFK_rect", "pdesolver_fitting", "pulsechase_solver", "sapdesolver", "sapdesolver_membSFK",
           "run_ensemble", "run_ensemble_pc", "pmap_fun_dk", "pmap_fun_allpars", "pmap_fun_dk_combD", "pmap_fun_concs",
           "fbatch_dk_mt", "fbatch_concs_mt", "pct_shp2_bound_gab1", "pdesolver_tangent_batch", "pdesolver_fitting_dual",
           "fitting_loss_and_gradient"):
    globals()[_n] = getattr(_default, _n)
