#!/usr/bin/env python
"""
TEST INFRASTRUCTURE ONLY.

Generate the golden fixtures under ``tests/golden/`` by running the REAL
reference (shimmed copy in ``oracle/_ref``, see ``oracle/make_ref.py``) on small
seeded inputs.  Run in the build container (needs /root/reference):

    python oracle/make_ref.py && python oracle/gen_golden.py

The fixtures pin (a) the numpy oracle (tests/test_oracle_golden.py, CPU) and
(b) the CUDA path (tests/test_gpu_*.py) to the reference's own outputs.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refimport  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def make_problem(rng, I, J, K, density, nonneg=False):
    Ut, Vt = rng.exponential(1.0, (I, K)), rng.exponential(1.0, (J, K))
    R = Ut @ Vt.T + rng.normal(0, 1.0, (I, J))
    if nonneg:
        R = np.abs(R) + 0.05
    M = (rng.random((I, J)) < density).astype(np.float64)
    for i in range(I):
        if M[i].sum() == 0:
            M[i, rng.integers(J)] = 1
    for j in range(J):
        if M[:, j].sum() == 0:
            M[rng.integers(I), j] = 1
    return R, M


def quiet(fn, *a, **kw):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **kw)


def gen_gibbs_params(ref, out):
    rng = np.random.default_rng(101)
    hyper = {"alphatau": 3.0, "betatau": 1.0, "alpha0": 6.0, "beta0": 2.0}
    for name, (I, J, K, dens) in {"a": (13, 9, 4, 0.7), "b": (40, 57, 6, 0.3)}.items():
        R, M = make_problem(rng, I, J, K, dens)
        U, V = rng.exponential(1.0, (I, K)), rng.exponential(0.7, (J, K))
        tau = 2.5
        lambdak = rng.gamma(2.0, 1.0, K)
        lamU, lamV = rng.gamma(2.0, 1.0, (I, K)), rng.gamma(2.0, 1.0, (J, K))
        for ard in (True, False):
            h = dict(hyper)
            if not ard:
                h.update(lambdaU=lamU, lambdaV=lamV)
            m = ref.bnmf_gibbs(R, M, K, ard, h)
            m.U, m.V, m.tau, m.lambdak = U.copy(), V.copy(), tau, lambdak.copy()
            tU = np.stack([m.tauU(k) for k in range(K)], 1)
            mU = np.stack([m.muU(tU[:, k], k) for k in range(K)], 1)
            tV = np.stack([m.tauV(k) for k in range(K)], 1)
            mV = np.stack([m.muV(tV[:, k], k) for k in range(K)], 1)
            p = "gibbs_%s_%s_" % (name, "ard" if ard else "noard")
            out.update({p + "R": R, p + "M": M, p + "U": U, p + "V": V, p + "tau": tau,
                        p + "lambdak": lambdak, p + "lamU": lamU, p + "lamV": lamV,
                        p + "tauU": tU, p + "muU": mU, p + "tauV": tV, p + "muV": mV,
                        p + "beta_s": m.beta_s(), p + "alpha_s": m.alpha_s(),
                        p + "betak_s": np.array([m.betak_s(k) for k in range(K)]) if ard else np.zeros(K),
                        p + "hyper": np.array([hyper[k] for k in ("alphatau", "betatau", "alpha0", "beta0")])})


def gen_icm_run(ref, out):
    rng = np.random.default_rng(202)
    hyper = {"alphatau": 1.0, "betatau": 1.0, "alpha0": 1.0, "beta0": 1.0}
    I, J, K, its = 30, 22, 5, 8
    R, M = make_problem(rng, I, J, K, 0.6)
    for ard in (True, False):
        h = dict(hyper)
        lamU, lamV = rng.gamma(2.0, 0.5, (I, K)), rng.gamma(2.0, 0.5, (J, K))
        if not ard:
            h.update(lambdaU=lamU, lambdaV=lamV)
        m = ref.nmf_icm(R, M, K, ard, h)
        np.random.seed(7)
        m.initialise("random")
        U0, V0, tau0, lam0 = m.U.copy(), m.V.copy(), m.tau, m.lambdak.copy()
        quiet(m.run, its)
        p = "icm_%s_" % ("ard" if ard else "noard")
        out.update({p + "R": R, p + "M": M, p + "U0": U0, p + "V0": V0, p + "tau0": tau0, p + "lambdak0": lam0,
                    p + "lamU": lamU, p + "lamV": lamV, p + "its": its,
                    p + "all_U": m.all_U, p + "all_V": m.all_V, p + "all_tau": m.all_tau,
                    p + "all_lambdak": m.all_lambdak,
                    p + "MSE": np.array(m.all_performances["MSE"]), p + "R2": np.array(m.all_performances["R^2"]),
                    p + "Rp": np.array(m.all_performances["Rp"]),
                    p + "hyper": np.array([hyper[k] for k in ("alphatau", "betatau", "alpha0", "beta0")])})


VB_FIELDS = ["mu_U", "tau_U", "exp_U", "var_U", "mu_V", "tau_V", "exp_V", "var_V"]
VB_SCAL = ["exp_tau", "exp_logtau", "alpha_s", "beta_s"]
VB_ARD = ["alphak_s", "betak_s", "exp_lambdak", "exp_loglambdak"]


def gen_vb_run(ref, out):
    rng = np.random.default_rng(303)
    hyper = {"alphatau": 1.0, "betatau": 1.0, "alpha0": 1.0, "beta0": 1.0}
    I, J, K, its = 28, 19, 4, 10
    R, M = make_problem(rng, I, J, K, 0.65)
    for ard in (True, False):
        h = dict(hyper)
        lamU, lamV = rng.gamma(2.0, 0.5, (I, K)), rng.gamma(2.0, 0.5, (J, K))
        if not ard:
            h.update(lambdaU=lamU, lambdaV=lamV)
        m = ref.bnmf_vb(R, M, K, ard, h)
        np.random.seed(11)
        m.initialise("random")
        p = "vb_%s_" % ("ard" if ard else "noard")
        for f in VB_FIELDS + VB_SCAL + (VB_ARD if ard else []):
            out[p + f + "0"] = np.array(getattr(m, f), dtype=np.float64).copy()
        out[p + "elbo0"] = m.elbo()
        out[p + "esd0"] = m.exp_square_diff()
        # single-step pieces from the initial state (what the reference tests poke)
        m2 = ref.bnmf_vb(R, M, K, ard, h)
        for f in VB_FIELDS + VB_SCAL + (VB_ARD if ard else []):
            setattr(m2, f, np.array(getattr(m, f), dtype=np.float64).copy() if np.ndim(getattr(m, f)) else getattr(m, f))
        m2.update_U(1); m2.update_exp_U(1); m2.update_V(2); m2.update_exp_V(2)
        out.update({p + "step_mu_U": m2.mu_U, p + "step_tau_U": m2.tau_U, p + "step_exp_U": m2.exp_U,
                    p + "step_var_U": m2.var_U, p + "step_mu_V": m2.mu_V, p + "step_tau_V": m2.tau_V,
                    p + "step_exp_V": m2.exp_V, p + "step_var_V": m2.var_V})
        elbos = []
        # run() prints ELBO but does not store it: iterate one at a time
        perfs = {"MSE": [], "R^2": [], "Rp": []}
        all_exp_tau = []
        for _ in range(its):
            quiet(m.run, 1)
            elbos.append(m.elbo())
            all_exp_tau.append(m.exp_tau)
            for k_ in perfs:
                perfs[k_].append(m.all_performances[k_][0])
        for f in VB_FIELDS + VB_SCAL + (VB_ARD if ard else []):
            out[p + f] = np.array(getattr(m, f), dtype=np.float64).copy()
        out.update({p + "R": R, p + "M": M, p + "lamU": lamU, p + "lamV": lamV, p + "its": its,
                    p + "elbo": np.array(elbos), p + "all_exp_tau": np.array(all_exp_tau),
                    p + "MSE": np.array(perfs["MSE"]), p + "R2": np.array(perfs["R^2"]), p + "Rp": np.array(perfs["Rp"]),
                    p + "hyper": np.array([hyper[k] for k in ("alphatau", "betatau", "alpha0", "beta0")])})


def gen_np_run(ref, out):
    rng = np.random.default_rng(404)
    I, J, K, its = 25, 31, 4, 12
    R, M = make_problem(rng, I, J, K, 0.55, nonneg=True)
    m = ref.nmf_np(R, M, K)
    np.random.seed(3)
    m.initialise("random")
    U0, V0 = m.U.copy(), m.V.copy()
    m1 = ref.nmf_np(R, M, K)
    m1.U, m1.V = U0.copy(), V0.copy()
    m1.update_U(2); m1.update_V(1)
    quiet(m.run, its)
    out.update({"np_R": R, "np_M": M, "np_U0": U0, "np_V0": V0, "np_its": its, "np_U": m.U, "np_V": m.V,
                "np_step_U": m1.U, "np_step_V": m1.V, "np_idiv": m.compute_I_div(),
                "np_MSE": np.array(m.all_performances["MSE"]), "np_R2": np.array(m.all_performances["R^2"]),
                "np_Rp": np.array(m.all_performances["Rp"])})


def gen_distributions(ref, out):
    mus = np.array([-400.0, -45.0, -31.0, -29.0, -10.0, -3.0, -1.0, -0.1, 0.0, 0.3, 1.0, 2.5, 8.0, 40.0, 1e3])
    taus = np.array([1e-3, 0.25, 1.0, 4.0, 2000.0])
    mm, tt = np.meshgrid(mus, taus, indexing="ij")
    mm, tt = mm.ravel(), tt.ravel()
    e_vec = np.array(ref.TN_vector_expectation(mm, tt), dtype=np.float64)
    v_vec = np.array(ref.TN_vector_variance(mm, tt), dtype=np.float64)
    e_sc = np.array([ref.TN_expectation(a, b) for a, b in zip(mm, tt)])
    v_sc = np.array([ref.TN_variance(a, b) for a, b in zip(mm, tt)])
    ab = np.array([[1.0, 1.0], [3.0, 2.0], [40001.0, 1234.5], [4.0e7, 3.9e7], [0.5, 7.0]])
    out.update({"dist_mu": mm, "dist_tau": tt, "dist_tn_exp_vec": e_vec, "dist_tn_var_vec": v_vec,
                "dist_tn_exp": e_sc, "dist_tn_var": v_sc,
                "dist_gamma_ab": ab,
                "dist_gamma_exp": np.array([ref.gamma_expectation(a, b) for a, b in ab]),
                "dist_gamma_explog": np.array([ref.gamma_expectation_log(a, b) for a, b in ab]),
                "dist_gamma_mode": np.array([ref.gamma_mode(a, b) for a, b in ab])})


def gen_metrics(ref, out):
    rng = np.random.default_rng(505)
    I, J, K = 17, 12, 3
    R, M = make_problem(rng, I, J, K, 0.5)
    P = R + rng.normal(0, 0.5, (I, J))
    m = ref.bnmf_gibbs(R, M, K, True, {"alphatau": 1, "betatau": 1, "alpha0": 1, "beta0": 1})
    out.update({"met_R": R, "met_M": M, "met_P": P,
                "met_vals": np.array([m.compute_MSE(M, R, P), m.compute_R2(M, R, P), m.compute_Rp(M, R, P)])})
    # predict / quality from supplied traces (bnmf_gibbs.py:222-302)
    its = 9
    m.all_U = rng.exponential(1.0, (its, I, K)); m.all_V = rng.exponential(1.0, (its, J, K))
    m.all_tau = rng.gamma(2.0, 1.0, its); m.all_lambdak = rng.gamma(2.0, 1.0, (its, K))
    Mt = (rng.random((I, J)) < 0.3).astype(float)
    perf = m.predict(Mt, 2, 3)
    out.update({"met_all_U": m.all_U, "met_all_V": m.all_V, "met_all_tau": m.all_tau, "met_all_lambdak": m.all_lambdak,
                "met_Mtest": Mt, "met_predict": np.array([perf["MSE"], perf["R^2"], perf["Rp"]]),
                "met_quality": np.array([m.quality(q, 2, 3) for q in ("loglikelihood", "BIC", "AIC", "MSE", "ELBO")])})


def main():
    ref = refimport.load()
    if ref is None:
        raise SystemExit("oracle/_ref missing: run `python oracle/make_ref.py` first")
    os.makedirs(OUT, exist_ok=True)
    out = {}
    gen_gibbs_params(ref, out)
    gen_icm_run(ref, out)
    gen_vb_run(ref, out)
    gen_np_run(ref, out)
    gen_distributions(ref, out)
    gen_metrics(ref, out)
    path = os.path.join(OUT, "nmf_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%d arrays, %.1f KB)" % (path, len(out), os.path.getsize(path) / 1024.0))


if __name__ == "__main__":
    main()
