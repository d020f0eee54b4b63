#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Golden fixtures for the NMTF path from the REAL
reference (oracle/_ref): python oracle/make_ref.py && python oracle/gen_golden_nmtf.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refimport  # noqa: E402
from gen_golden import make_problem, quiet  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
HYP = {"alphatau": 3.0, "betatau": 1.0, "alpha0": 6.0, "beta0": 2.0}


def gen_params(ref, out):
    rng = np.random.default_rng(707)
    for name, (I, J, K, L, dens) in {"a": (12, 10, 3, 4, 0.7), "b": (35, 48, 5, 3, 0.35)}.items():
        R, M = make_problem(rng, I, J, max(K, L), dens)
        F, S, G = rng.exponential(1.0, (I, K)), rng.exponential(0.8, (K, L)), rng.exponential(0.7, (J, L))
        tau = 1.7
        lamFk, lamGl = rng.gamma(2.0, 1.0, K), rng.gamma(2.0, 1.0, L)
        lamF, lamG, lamS = rng.gamma(2.0, 1.0, (I, K)), rng.gamma(2.0, 1.0, (J, L)), rng.gamma(2.0, 0.3, (K, L))
        for ard in (True, False):
            h = dict(HYP, lambdaS=lamS)
            if not ard:
                h.update(lambdaF=lamF, lambdaG=lamG)
            m = ref.bnmtf_gibbs(R, M, K, L, ard, h)
            m.F, m.S, m.G, m.tau = F.copy(), S.copy(), G.copy(), tau
            m.lambdaFk, m.lambdaGl = lamFk.copy(), lamGl.copy()
            tF = np.stack([m.tauF(k) for k in range(K)], 1)
            mF = np.stack([m.muF(tF[:, k], k) for k in range(K)], 1)
            tG = np.stack([m.tauG(l) for l in range(L)], 1)
            mG = np.stack([m.muG(tG[:, l], l) for l in range(L)], 1)
            tS = np.array([[m.tauS(k, l) for l in range(L)] for k in range(K)])
            mS = np.array([[m.muS(tS[k, l], k, l) for l in range(L)] for k in range(K)])
            p = "tg_%s_%s_" % (name, "ard" if ard else "noard")
            out.update({p + "R": R, p + "M": M, p + "F": F, p + "S": S, p + "G": G, p + "tau": tau,
                        p + "lamFk": lamFk, p + "lamGl": lamGl, p + "lamF": lamF, p + "lamG": lamG, p + "lamS": lamS,
                        p + "tauF": tF, p + "muF": mF, p + "tauS": tS, p + "muS": mS, p + "tauG": tG, p + "muG": mG,
                        p + "beta_s": m.beta_s(), p + "alpha_s": m.alpha_s(),
                        p + "betaFk_s": np.array([m.betaFk_s(k) for k in range(K)]) if ard else np.zeros(K),
                        p + "betaGl_s": np.array([m.betaGl_s(l) for l in range(L)]) if ard else np.zeros(L),
                        p + "hyper": np.array([HYP[k] for k in ("alphatau", "betatau", "alpha0", "beta0")])})


def gen_icm_run(ref, out):
    rng = np.random.default_rng(808)
    hyper = {"alphatau": 1.0, "betatau": 1.0, "alpha0": 1.0, "beta0": 1.0}
    I, J, K, L, its = 26, 21, 4, 3, 6
    R, M = make_problem(rng, I, J, max(K, L), 0.6)
    lamS = rng.gamma(2.0, 0.3, (K, L))
    for ard in (True, False):
        lamF, lamG = rng.gamma(2.0, 0.5, (I, K)), rng.gamma(2.0, 0.5, (J, L))
        h = dict(hyper, lambdaS=lamS)
        if not ard:
            h.update(lambdaF=lamF, lambdaG=lamG)
        m = ref.nmtf_icm(R, M, K, L, ard, h)
        np.random.seed(9)
        m.initialise("random", "random")
        F0, S0, G0, tau0 = m.F.copy(), m.S.copy(), m.G.copy(), m.tau
        lF0, lG0 = m.lambdaFk.copy(), m.lambdaGl.copy()
        quiet(m.run, its)
        p = "ti_%s_" % ("ard" if ard else "noard")
        out.update({p + "R": R, p + "M": M, p + "F0": F0, p + "S0": S0, p + "G0": G0, p + "tau0": tau0,
                    p + "lamFk0": lF0, p + "lamGl0": lG0, p + "lamF": lamF, p + "lamG": lamG, p + "lamS": lamS,
                    p + "its": its, p + "all_F": m.all_F, p + "all_S": m.all_S, p + "all_G": m.all_G,
                    p + "all_tau": m.all_tau, p + "all_lambdaFk": m.all_lambdaFk, p + "all_lambdaGl": m.all_lambdaGl,
                    p + "MSE": np.array(m.all_performances["MSE"]), p + "R2": np.array(m.all_performances["R^2"]),
                    p + "Rp": np.array(m.all_performances["Rp"]),
                    p + "hyper": np.array([hyper[k] for k in ("alphatau", "betatau", "alpha0", "beta0")])})


VB_FIELDS = ["mu_F", "tau_F", "exp_F", "var_F", "mu_S", "tau_S", "exp_S", "var_S", "mu_G", "tau_G", "exp_G", "var_G"]
VB_SCAL = ["exp_tau", "exp_logtau", "alpha_s", "beta_s"]
VB_ARD = ["alphaFk_s", "betaFk_s", "exp_lambdaFk", "exp_loglambdaFk", "alphaGl_s", "betaGl_s", "exp_lambdaGl", "exp_loglambdaGl"]


def gen_vb_run(ref, out):
    rng = np.random.default_rng(909)
    hyper = {"alphatau": 1.0, "betatau": 1.0, "alpha0": 1.0, "beta0": 1.0}
    I, J, K, L, its = 24, 19, 4, 3, 8
    R, M = make_problem(rng, I, J, max(K, L), 0.65)
    lamS = rng.gamma(2.0, 0.3, (K, L))
    for ard in (True, False):
        lamF, lamG = rng.gamma(2.0, 0.5, (I, K)), rng.gamma(2.0, 0.5, (J, L))
        h = dict(hyper, lambdaS=lamS)
        if not ard:
            h.update(lambdaF=lamF, lambdaG=lamG)
        m = ref.bnmtf_vb(R, M, K, L, ard, h)
        np.random.seed(13)
        m.initialise("random", "random")
        p = "tv_%s_" % ("ard" if ard else "noard")
        fields = VB_FIELDS + VB_SCAL + (VB_ARD if ard else [])
        for f in fields:
            out[p + f + "0"] = np.array(getattr(m, f), dtype=np.float64).copy()
        out[p + "elbo0"] = m.elbo()
        out[p + "esd0"] = m.exp_square_diff()
        # single-step updates from the initial state (what tests/code/test_bnmtf_vb.py pokes)
        m2 = ref.bnmtf_vb(R, M, K, L, ard, h)
        for f in fields:
            v = getattr(m, f)
            setattr(m2, f, np.array(v, dtype=np.float64).copy() if np.ndim(v) else v)
        m2.update_F(1); m2.update_exp_F(1); m2.update_S(2, 1); m2.update_exp_S(2, 1); m2.update_G(2); m2.update_exp_G(2)
        for f in VB_FIELDS:
            out[p + "step_" + f] = np.array(getattr(m2, f), dtype=np.float64).copy()
        elbos, taus, mses = [], [], []
        for _ in range(its):
            quiet(m.run, 1)
            elbos.append(m.elbo()); taus.append(m.exp_tau); mses.append(m.all_performances["MSE"][0])
        for f in fields:
            out[p + f] = np.array(getattr(m, f), dtype=np.float64).copy()
        out.update({p + "R": R, p + "M": M, p + "lamF": lamF, p + "lamG": lamG, p + "lamS": lamS, p + "its": its,
                    p + "elbo": np.array(elbos), p + "all_exp_tau": np.array(taus), p + "MSE": np.array(mses),
                    p + "hyper": np.array([hyper[k] for k in ("alphatau", "betatau", "alpha0", "beta0")])})


def gen_np_run(ref, out):
    rng = np.random.default_rng(1010)
    I, J, K, L, its = 23, 27, 3, 4, 7
    R, M = make_problem(rng, I, J, max(K, L), 0.6, nonneg=True)
    m = ref.nmtf_np(R, M, K, L)
    np.random.seed(5)
    m.initialise("random", "random")
    F0, S0, G0 = m.F.copy(), m.S.copy(), m.G.copy()
    m1 = ref.nmtf_np(R, M, K, L)
    m1.F, m1.S, m1.G = F0.copy(), S0.copy(), G0.copy()
    m1.update_F(1); m1.update_S(2, 1); m1.update_G(3)
    quiet(m.run, its)
    out.update({"tn_R": R, "tn_M": M, "tn_F0": F0, "tn_S0": S0, "tn_G0": G0, "tn_its": its,
                "tn_F": m.F, "tn_S": m.S, "tn_G": m.G, "tn_step_F": m1.F, "tn_step_S": m1.S, "tn_step_G": m1.G,
                "tn_idiv": m.compute_I_div(), "tn_MSE": np.array(m.all_performances["MSE"]),
                "tn_Rp": np.array(m.all_performances["Rp"])})


def main():
    ref = refimport.load()
    if ref is None:
        raise SystemExit("oracle/_ref missing: run `python oracle/make_ref.py` first")
    out = {}
    gen_params(ref, out)
    gen_icm_run(ref, out)
    gen_vb_run(ref, out)
    gen_np_run(ref, out)
    path = os.path.join(OUT, "nmtf_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%d arrays, %.1f KB)" % (path, len(out), os.path.getsize(path) / 1024.0))


if __name__ == "__main__":
    main()
