#!/usr/bin/env python
"""
TEST INFRASTRUCTURE ONLY.

Golden fixtures on the configurations BASELINE.json names, produced by the REAL
reference (shimmed copy in ``oracle/_ref``, see ``oracle/make_ref.py``) on the
reference's own datasets:

  C1  data/toy/bnmf/R.txt (100 x 80, fully observed), K=10, ARD
      (experiments/experiments_toy/convergence/nmf_{vb,icm,np}.py): VB, ICM and NP
      trajectories from a seeded initialisation.
  C2  data/drug_sensitivity/GDSC/processed_all/ic50.txt through load_gdsc_ic50
      (705 x 139, 80.7 % observed), K=20, VB + ARD, 200 iterations
      (experiments/experiments_gdsc/convergence/nmf_vb.py): ELBO / MSE per iteration.
  Gibbs chains (north_star correctness level 3): 10 seeded reference chains on the
      toy matrix (K=10, ARD, 500 iterations, 10 % held out) and on GDSC (K=20, ARD,
      200 iterations, burn_in 180, thinning 2 -- the settings of
      experiments/experiments_gdsc/cross_validation/nmf_gibbs_ard/nmf_gibbs_ard_xval.py,
      10 % held out): per chain the held-out MSE / R^2 / Rp and the posterior-mean
      prediction of the held-out entries.

Run in the build container (needs /root/reference):

    python oracle/make_ref.py && python oracle/gen_golden_configs.py [c1] [c2] [chains]

Writes tests/golden/config_{c1,c2,chains}.npz.  The input matrices travel inside the
fixtures (the GPU box has no /root/reference).
"""
import contextlib
import io
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refimport  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
REFERENCE = "/root/reference"
HYPER = {"alphatau": 1.0, "betatau": 1.0, "alpha0": 1.0, "beta0": 1.0}

VB_FIELDS = ["mu_U", "tau_U", "exp_U", "var_U", "mu_V", "tau_V", "exp_V", "var_V"]
VB_SCAL = ["exp_tau", "exp_logtau", "alpha_s", "beta_s"]
VB_ARD = ["alphak_s", "betak_s", "exp_lambdak", "exp_loglambdak"]


def quiet(fn, *a, **kw):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **kw)


def load_toy():
    R = np.loadtxt(os.path.join(REFERENCE, "data", "toy", "bnmf", "R.txt"))
    return R, np.ones(R.shape)


def load_gdsc():
    """load_gdsc_ic50 of the reference (data/drug_sensitivity/load_data.py:37-52), imported from the read-only checkout; it has no Python-2 syntax."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "ref_load_data", os.path.join(REFERENCE, "data", "drug_sensitivity", "load_data.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.load_gdsc_ic50(os.path.join(REFERENCE, "data", "drug_sensitivity", "GDSC", "processed_all", "ic50.txt"))


def vb_state(m, ard):
    return {f: np.array(getattr(m, f), dtype=np.float64).copy() for f in VB_FIELDS + VB_SCAL + (VB_ARD if ard else [])}


def run_vb(ref, R, M, K, its, seed, out, p):
    m = ref.bnmf_vb(R, M, K, True, dict(HYPER))
    np.random.seed(seed)
    m.initialise("random")
    for f, v in vb_state(m, True).items():
        out[p + f + "0"] = v
    elbo, mse, r2, rp, etau = [], [], [], [], []
    for _ in range(its):                 # run() prints the ELBO but does not keep it: one iteration at a time
        quiet(m.run, 1)
        elbo.append(m.elbo()); etau.append(m.exp_tau)
        mse.append(m.all_performances["MSE"][0]); r2.append(m.all_performances["R^2"][0]); rp.append(m.all_performances["Rp"][0])
    for f, v in vb_state(m, True).items():
        out[p + f] = v
    out.update({p + "its": its, p + "elbo": np.array(elbo), p + "all_exp_tau": np.array(etau), p + "MSE": np.array(mse),
                p + "R2": np.array(r2), p + "Rp": np.array(rp)})


def gen_c1(ref):
    out = {}
    R, M = load_toy()
    K = 10
    out.update({"R": R, "M": M, "K": K})
    t0 = time.time()
    run_vb(ref, R, M, K, 500, 0, out, "vb_")
    print("C1 VB   500 it: %.1f s, final MSE %.6f ELBO %.4f" % (time.time() - t0, out["vb_MSE"][-1], out["vb_elbo"][-1]))
    # ICM (experiments_toy/convergence/nmf_icm.py; 0.1 floor of nmf_icm.py:159-167)
    t0 = time.time()
    m = ref.nmf_icm(R, M, K, True, dict(HYPER))
    np.random.seed(0)
    m.initialise("random")
    out.update({"icm_U0": m.U.copy(), "icm_V0": m.V.copy(), "icm_tau0": m.tau, "icm_lambdak0": m.lambdak.copy()})
    its = 500
    quiet(m.run, its)
    out.update({"icm_its": its, "icm_U": m.U.copy(), "icm_V": m.V.copy(), "icm_all_tau": np.array(m.all_tau),
                "icm_all_lambdak": np.array(m.all_lambdak), "icm_MSE": np.array(m.all_performances["MSE"]),
                "icm_R2": np.array(m.all_performances["R^2"]), "icm_Rp": np.array(m.all_performances["Rp"])})
    print("C1 ICM  %d it: %.1f s, final MSE %.6f" % (its, time.time() - t0, out["icm_MSE"][-1]))
    # NP (experiments_toy/convergence/nmf_np.py): multiplicative updates need R > 0 where observed
    t0 = time.time()
    Rn = np.abs(R) + 1e-3 * (R == 0)
    m = ref.nmf_np(Rn, M, K)
    np.random.seed(0)
    m.initialise("random")
    out.update({"np_R": Rn, "np_U0": m.U.copy(), "np_V0": m.V.copy()})
    its = 500
    quiet(m.run, its)
    out.update({"np_its": its, "np_U": m.U.copy(), "np_V": m.V.copy(), "np_MSE": np.array(m.all_performances["MSE"]),
                "np_Rp": np.array(m.all_performances["Rp"]), "np_idiv": m.compute_I_div()})
    print("C1 NP   %d it: %.1f s, final MSE %.6f" % (its, time.time() - t0, out["np_MSE"][-1]))
    np.savez_compressed(os.path.join(OUT, "config_c1.npz"), **out)


def gen_c2(ref):
    out = {}
    R, M = load_gdsc()
    K = 20
    out.update({"R": R, "M": M, "K": K})
    t0 = time.time()
    run_vb(ref, R, M, K, 200, 0, out, "vb_")
    print("C2 VB   200 it on %s: %.1f s, final MSE %.4f ELBO %.4f" % (R.shape, time.time() - t0, out["vb_MSE"][-1], out["vb_elbo"][-1]))
    np.savez_compressed(os.path.join(OUT, "config_c2.npz"), **out)


def held_out_mask(M, fraction, seed):
    """A held-out tenth of the observed entries that leaves no empty row or column in the training mask."""
    rng = np.random.default_rng(seed)
    obs = np.flatnonzero(M.ravel() != 0)
    for _ in range(1000):
        test = np.zeros(M.size)
        test[rng.choice(obs, size=int(round(fraction * obs.size)), replace=False)] = 1
        test = test.reshape(M.shape)
        train = M - test
        if train.sum(1).min() > 0 and train.sum(0).min() > 0:
            return train, test
    raise RuntimeError("no valid split")


def gibbs_chains(ref, R, M, K, its, burn_in, thinning, n_chains, out, p):
    train, test = held_out_mask(M, 0.1, 12345)
    ti, tj = np.nonzero(test)
    out.update({p + "R": R, p + "M_train": train, p + "M_test": test, p + "K": K, p + "its": its, p + "burn_in": burn_in,
                p + "thinning": thinning})
    perf, preds, train_mse, taus = [], [], [], []
    for c in range(n_chains):
        t0 = time.time()
        m = ref.bnmf_gibbs(R, train, K, True, dict(HYPER))
        np.random.seed(1000 + c)
        m.initialise("random")
        quiet(m.run, its)
        pf = m.predict(test, burn_in, thinning)
        eU, eV, _, _ = m.approx_expectation(burn_in, thinning)
        preds.append((eU @ eV.T)[ti, tj])
        perf.append([pf["MSE"], pf["R^2"], pf["Rp"]])
        train_mse.append(m.all_performances["MSE"][-1])
        taus.append(np.mean(m.all_tau[burn_in::thinning]))
        print("  %schain %d: %.1f s, held-out MSE %.4f, train MSE %.4f" % (p, c, time.time() - t0, pf["MSE"], train_mse[-1]))
    out.update({p + "perf": np.array(perf), p + "pred": np.array(preds), p + "train_mse": np.array(train_mse),
                p + "tau_mean": np.array(taus)})


def gen_chains(ref):
    out = {}
    R, M = load_toy()
    gibbs_chains(ref, R, M, 10, 500, 250, 2, 10, out, "toy_")
    R, M = load_gdsc()
    gibbs_chains(ref, R, M, 20, 200, 180, 2, 10, out, "gdsc_")
    np.savez_compressed(os.path.join(OUT, "config_chains.npz"), **out)


def main():
    ref = refimport.load()
    assert ref is not None, "run oracle/make_ref.py first (needs /root/reference)"
    which = sys.argv[1:] or ["c1", "c2", "chains"]
    os.makedirs(OUT, exist_ok=True)
    if "c1" in which:
        gen_c1(ref)
    if "c2" in which:
        gen_c2(ref)
    if "chains" in which:
        gen_chains(ref)


if __name__ == "__main__":
    main()
