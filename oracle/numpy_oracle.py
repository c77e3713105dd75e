"""Second, independent CPU restatement of the reference solvers in NumPy — TEST INFRASTRUCTURE, NOT PRODUCT.

Written separately from oracle/gab1_oracle.c (different data layout, own argument handling) and required by
tests/test_oracle_cross.py to agree with it BIT FOR BIT: every elementwise NumPy operation is one IEEE-754 binary64
operation per element, applied in the reference's source order, so two faithful restatements cannot differ.
The time loop is a Python loop: small cases only.  PARITY UNPINNED against the reference itself (no Julia here).

Follows basepdesolver.jl:25-312 (pdesolver), :350-636 (membSFK), basepdesolver_rect.jl:23-294, :298-569,
sapdesolver.jl:55-280, sapdesolver_memb-SFK.jl:55-281, pulsechase_solver.jl:156-158.
"""
from __future__ import annotations

import math

import numpy as np

CYTO = ("iSFK", "aSFK", "GAB1", "pGAB1", "GRB2", "G2G1", "G2PG1", "SHP2", "PG1S", "G2PG1S")
MEMB = ("mE", "mES", "mESmES", "E", "EG2", "EG2G1", "EG2PG1", "EG2PG1S")


def _julia_maximum_abs_rel(new, old):
    with np.errstate(divide="ignore", invalid="ignore"):
        e = np.abs(1.0 - new / old)
    return float("nan") if np.isnan(e).any() else float(e.max())


def solve(Co, D, k, r, *, dr, tf, dt, Nts=100, dt_save=None, maxiters=100, tol=1e-6, rect=False, sfk_mode=0,
          while_loop=False, modulus_rule=False, chain_pg1tot=False, t_prechase=None, snapshots=True, cap=10**9):
    """Returns dict(sol=..., t_out=..., n_saved=..., n_bc=..., final=..., memb=...)."""
    f8 = np.float64
    Co = [f8(x) for x in Co]
    D = [f8(x) for x in D]
    k = [f8(x) for x in k]
    r = np.asarray(r, dtype=f8)
    dr, tf, dt, tol = f8(dr), f8(tf), f8(dt), f8(tol)
    dt_save = tf / f8(Nts) if dt_save is None else f8(dt_save)
    P = r.shape[0]
    Nr = P - 1
    Nt = int(math.ceil(tf / dt))
    D_S, D_G2, D_G2G1, D_G2G1S2, D_G1, D_G1S2, D_S2 = D
    D_Si, D_Sa = D_S, D_S
    if sfk_mode == 1:
        D_Sa = f8(1e-32)
    elif sfk_mode == 2:
        D_Si = D_Sa = f8(1e-32)
    Dq = dict(iSFK=D_Si, aSFK=D_Sa, GAB1=D_G1, pGAB1=D_G1, GRB2=D_G2, G2G1=D_G2G1, G2PG1=D_G2G1, SHP2=D_S2,
              PG1S=D_G1S2, G2PG1S=D_G2G1S2)
    (kS2f, kS2r, kG1f, kG1r, kG2f, kG2r, kG1p, kG1dp, kSa, kSi, kp, kdp, kEGFf, kEGFr, EGF, kdf, kdr) = k
    CoSFK, CoG2, CoG1, CoS2, CoEGFR = Co

    old = {n: np.zeros(P) for n in CYTO}
    new = {n: np.zeros(P) for n in CYTO}
    old["iSFK"][:] = CoSFK
    old["GAB1"][:] = CoG1
    old["GRB2"][:] = CoG2
    old["SHP2"][:] = CoS2
    m_old = {n: f8(0.0) for n in MEMB}
    m_new = {n: f8(0.0) for n in MEMB}
    m_old["mE"] = CoEGFR

    C = Nts + 1
    names_out = ("iSFK", "aSFK", "GRB2", "GAB1", "SHP2", "G2G1", "G2PG1", "G2PG1S", "PG1", "PG1S", "PG1tot", "PG1Stot")
    mats = {n: np.zeros((P, C)) for n in names_out}
    vecs = {n: np.zeros(C) for n in ("pE",) + MEMB + ("EGFR_SHP2", "t_out")}
    mats["iSFK"][:, 0] = CoSFK
    mats["GRB2"][:, 0] = CoG2
    mats["SHP2"][:, 0] = CoS2
    mats["GAB1"][:, 0] = CoG1
    vecs["mE"][0] = CoEGFR

    t = f8(0.0)
    t_save = dt_save
    nts = 1
    n_bc = 0
    overflow = False
    modulus_step = f8(np.rint(f8(Nt) / f8(Nts))) if modulus_rule else None
    dr2 = dr * dr
    inner = slice(1, Nr)
    up_s, dn_s = slice(2, Nr + 1), slice(0, Nr - 1)
    a = 1 / (r[inner] * dr) if not rect else None

    for step in range(1, Nt + 1):
        if t_prechase is not None:
            if f8(t_prechase) + dt > t and t >= f8(t_prechase):
                kp = f8(0.0)
        c = {n: old[n][inner] for n in CYTO}
        L = {}
        for n in CYTO:
            up, um, uc = old[n][up_s], old[n][dn_s], old[n][inner]
            if rect:
                L[n] = Dq[n] * (up - 2.0 * uc + um) / dr2
            else:
                L[n] = Dq[n] * (a * (up - um) + (up - 2.0 * uc + um) / dr2)
        new["iSFK"][inner] = (L["iSFK"] + kSi * c["aSFK"]) * dt + c["iSFK"]
        new["aSFK"][inner] = (L["aSFK"] - kSi * c["aSFK"]) * dt + c["aSFK"]
        new["GAB1"][inner] = (L["GAB1"] - kG1f * c["GAB1"] * c["GRB2"] + kG1r * c["G2G1"] - kG1p * c["aSFK"] * c["GAB1"]
                              + kG1dp * c["pGAB1"]) * dt + c["GAB1"]
        new["pGAB1"][inner] = (L["pGAB1"] - kG1f * c["pGAB1"] * c["GRB2"] + kG1r * c["G2PG1"] + kG1p * c["aSFK"] * c["GAB1"]
                               - kG1dp * c["pGAB1"] - kS2f * c["SHP2"] * c["pGAB1"] + kS2r * c["PG1S"]) * dt + c["pGAB1"]
        new["GRB2"][inner] = (L["GRB2"] - kG1f * c["GAB1"] * c["GRB2"] + kG1r * c["G2G1"] - kG1f * c["pGAB1"] * c["GRB2"]
                              + kG1r * c["G2PG1"] - kG1f * c["GRB2"] * c["PG1S"] + kG1r * c["G2PG1S"]) * dt + c["GRB2"]
        new["G2G1"][inner] = (L["G2G1"] + kG1f * c["GAB1"] * c["GRB2"] - kG1r * c["G2G1"] - kG1p * c["aSFK"] * c["G2G1"]
                              + kG1dp * c["G2PG1"]) * dt + c["G2G1"]
        new["G2PG1"][inner] = (L["G2PG1"] + kG1f * c["pGAB1"] * c["GRB2"] - kG1r * c["G2PG1"] + kG1p * c["aSFK"] * c["G2G1"]
                               - kG1dp * c["G2PG1"] - kS2f * c["SHP2"] * c["G2PG1"] + kS2r * c["G2PG1S"]) * dt + c["G2PG1"]
        new["SHP2"][inner] = (L["SHP2"] - kS2f * c["SHP2"] * c["pGAB1"] + kS2r * c["PG1S"] - kS2f * c["SHP2"] * c["G2PG1"]
                              + kS2r * c["G2PG1S"]) * dt + c["SHP2"]
        new["PG1S"][inner] = (L["PG1S"] + kS2f * c["SHP2"] * c["pGAB1"] - kS2r * c["PG1S"] - kG1f * c["GRB2"] * c["PG1S"]
                              + kG1r * c["G2PG1S"]) * dt + c["PG1S"]
        new["G2PG1S"][inner] = (L["G2PG1S"] + kG1f * c["GRB2"] * c["PG1S"] - kG1r * c["G2PG1S"] + kS2f * c["SHP2"] * c["G2PG1"]
                                - kS2r * c["G2PG1S"]) * dt + c["G2PG1S"]
        for n in CYTO:
            new[n][0] = new[n][1]

        err = tol * 2.0
        it = 0
        with np.errstate(all="ignore"):
            while True:
                if not while_loop:
                    if it >= maxiters:
                        break
                else:
                    if not (err > tol) or it >= cap:
                        break
                it += 1
                cyto_old = np.array([new[n][Nr] for n in CYTO])
                memb_old = np.array([m_new[n] for n in MEMB])
                M = m_new
                I = {n: new[n][Nr - 1] for n in CYTO}
                Etot = 2.0 * (M["E"] + M["EG2"] + M["EG2G1"] + M["EG2PG1"] + M["EG2PG1S"])
                b = {}
                b["iSFK"] = I["iSFK"] / (1 + kSa * Etot * dr / D_Si)
                b["aSFK"] = I["aSFK"] + kSa * b["iSFK"] * Etot * dr / D_Sa
                b["GAB1"] = (kG1r * M["EG2G1"] * dr / D_G1 + I["GAB1"]) / (1 + kG1f * M["EG2"] * dr / D_G1)
                b["pGAB1"] = (kG1r * M["EG2PG1"] * dr / D_G1 + I["pGAB1"]) / (1 + kG1f * M["EG2"] * dr / D_G1)
                b["GRB2"] = (kG2r * M["EG2"] * dr / D_G2 + I["GRB2"]) / (1 + kG2f * M["E"] * dr / D_G2)
                b["G2G1"] = (kG2r * M["EG2G1"] * dr / D_G2G1 + I["G2G1"]) / (1 + kG2f * M["E"] * dr / D_G2G1)
                b["G2PG1"] = (kG2r * M["EG2PG1"] * dr / D_G2G1 + I["G2PG1"]) / (1 + kG2f * M["E"] * dr / D_G2G1)
                b["SHP2"] = (kS2r * M["EG2PG1S"] * dr / D_S2 + I["SHP2"]) / (1 + kS2f * M["EG2PG1"] * dr / D_S2)
                b["PG1S"] = (kG1r * M["EG2PG1S"] * dr / D_G1S2 + I["PG1S"]) / (1 + kG1f * M["EG2"] * dr / D_G1S2)
                b["G2PG1S"] = (kG2r * M["EG2PG1S"] * dr / D_G2G1S2 + I["G2PG1S"]) / (1 + kG2f * M["E"] * dr / D_G2G1S2)
                for n in CYTO:
                    new[n][Nr] = b[n]
                o = m_old
                n_ = {}
                n_["mE"] = (-kEGFf * EGF * o["mE"] + kEGFr * o["mES"]) * dt + o["mE"]
                n_["mES"] = (kEGFf * EGF * o["mE"] - kEGFr * o["mES"] - 2 * kdf * o["mES"] * o["mES"] + 2 * kdr * o["mESmES"]) * dt + o["mES"]
                n_["mESmES"] = (kdf * o["mES"] * o["mES"] - kdr * o["mESmES"] - kp * o["mESmES"] + kdp * o["E"]) * dt + o["mESmES"]
                n_["E"] = (kp * o["mESmES"] - kdp * o["E"] - kG2f * o["E"] * b["GRB2"] + kG2r * o["EG2"] - kG2f * o["E"] * b["G2G1"]
                           + kG2r * o["EG2G1"] - kG2f * o["E"] * b["G2PG1"] + kG2r * o["EG2PG1"] - kG2f * o["E"] * b["G2PG1S"]
                           + kG2r * o["EG2PG1S"]) * dt + o["E"]
                n_["EG2"] = (kG2f * b["GRB2"] * o["E"] - kG2r * o["EG2"] - kG1f * b["GAB1"] * o["EG2"] + kG1r * o["EG2G1"]
                             - kG1f * b["pGAB1"] * o["EG2"] + kG1r * o["EG2PG1"] - kG1f * b["PG1S"] * o["EG2"]
                             + kG1r * o["EG2PG1S"]) * dt + o["EG2"]
                n_["EG2G1"] = (kG2f * b["G2G1"] * o["E"] - kG2r * o["EG2G1"] + kG1f * b["GAB1"] * o["EG2"] - kG1r * o["EG2G1"]) * dt + o["EG2G1"]
                n_["EG2PG1"] = (kG2f * b["G2PG1"] * o["E"] - kG2r * o["EG2PG1"] + kG1f * b["pGAB1"] * o["EG2"] - kG1r * o["EG2PG1"]
                                - kS2f * b["SHP2"] * o["EG2PG1"] + kS2r * o["EG2PG1S"]) * dt + o["EG2PG1"]
                n_["EG2PG1S"] = (kS2f * b["SHP2"] * o["EG2PG1"] - kS2r * o["EG2PG1S"] + kG1f * b["PG1S"] * o["EG2"] - kG1r * o["EG2PG1S"]
                                 + kG2f * b["G2PG1S"] * o["E"] - kG2r * o["EG2PG1S"]) * dt + o["EG2PG1S"]
                m_new = n_
                cyto_new = np.array([b[n] for n in CYTO])
                memb_new = np.array([n_[n] for n in MEMB])
                err = _julia_maximum_abs_rel(np.concatenate([cyto_new, memb_new]), np.concatenate([cyto_old, memb_old]))
                if not while_loop and err <= tol:
                    break
        n_bc += it
        for n in CYTO:
            old[n][:] = new[n]
        m_old = dict(m_new)

        if snapshots:
            Etot = 2.0 * (m_new["E"] + m_new["EG2"] + m_new["EG2G1"] + m_new["EG2PG1"] + m_new["EG2PG1S"])
            t = t + dt
            if modulus_rule:
                with np.errstate(all="ignore"):
                    save = bool(np.fmod(f8(step), modulus_step) == 0.0)
            else:
                save = bool(t >= t_save)
            if save:
                if nts >= C:
                    overflow = True
                else:
                    cidx = nts
                    nts += 1
                    for on, sn in (("iSFK", "iSFK"), ("aSFK", "aSFK"), ("GRB2", "GRB2"), ("GAB1", "GAB1"), ("SHP2", "SHP2"),
                                   ("G2G1", "G2G1"), ("G2PG1", "G2PG1"), ("G2PG1S", "G2PG1S"), ("PG1", "pGAB1"), ("PG1S", "PG1S")):
                        mats[on][:, cidx] = new[sn]
                    stot = new["PG1S"] + new["G2PG1S"]
                    mats["PG1Stot"][:, cidx] = stot
                    if chain_pg1tot:
                        mats["PG1tot"][:, cidx] = new["G2PG1"] + new["pGAB1"] + new["PG1S"] + new["G2PG1S"]
                    else:
                        mats["PG1tot"][:, cidx] = new["G2PG1"] + new["pGAB1"] + stot
                    with np.errstate(all="ignore"):
                        vecs["pE"][cidx] = Etot * 100.0 / CoEGFR
                        vecs["EGFR_SHP2"][cidx] = m_new["EG2PG1S"] * 100.0 / CoEGFR
                    for n in MEMB:
                        vecs[n][cidx] = m_new[n]
                    vecs["t_out"][cidx] = t
                if not modulus_rule:
                    t_save = t_save + dt_save

    stot = new["PG1S"] + new["G2PG1S"]
    ptot = (new["G2PG1"] + new["pGAB1"] + new["PG1S"] + new["G2PG1S"]) if chain_pg1tot else (new["G2PG1"] + new["pGAB1"] + stot)
    return dict(mats=mats, vecs=vecs, n_saved=nts, n_bc=n_bc, Nt=Nt, overflow=overflow,
                final={n: new[n].copy() for n in CYTO}, memb=dict(m_new), PG1Stot=stot, PG1tot=ptot)


def six_scalars(r, aSFK, PG1Stot, R):
    """pmap_fun_dk reductions (sapdesolver.jl:343-356); raises ValueError where Julia's `minimum` of an empty
    selection would throw."""
    r = np.asarray(r, dtype=np.float64)

    def ls(y, f):
        mx = float("nan") if np.isnan(y).any() else y.max()
        sel = r[y >= f * mx]
        if sel.size == 0:
            raise ValueError("reducing over an empty collection")
        return R - sel.min()

    out = [ls(aSFK, 0.5), ls(aSFK, 0.1), ls(PG1Stot, 0.5), ls(PG1Stot, 0.1)]
    with np.errstate(all="ignore"):
        out.append(PG1Stot[0] / PG1Stot[-1])
    y = PG1Stot * (r * r)
    acc = np.float64(0.0)
    for i in range(len(r) - 1):
        acc = acc + (r[i + 1] - r[i]) * (y[i] + y[i + 1])
    out.append(0.5 * acc * 3.0 / math.pow(R, 3.0))
    return np.array(out)
