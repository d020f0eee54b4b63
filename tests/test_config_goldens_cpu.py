"""CPU: the numpy oracle against the goldens the REAL reference produced on the configurations BASELINE.json names
(oracle/gen_golden_configs.py): toy matrix K = 10 and GDSC K = 20.  The GPU tests of tests/test_gpu_configs.py compare
the CUDA path with the same fixtures over the full iteration counts; here the oracle is pinned on their first iterations."""
import os

import numpy as np

import nmf_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
HYPER = {"alphatau": 1.0, "betatau": 1.0, "alpha0": 1.0, "beta0": 1.0}
VB_FIELDS = ["mu_U", "tau_U", "exp_U", "var_U", "mu_V", "tau_V", "exp_V", "var_V"]
VB_SCAL = ["exp_tau", "exp_logtau", "alpha_s", "beta_s"]
VB_ARD = ["alphak_s", "betak_s", "exp_lambdak", "exp_loglambdak"]


def vb_state(g):
    st = {f: g["vb_" + f + "0"].copy() for f in VB_FIELDS + VB_ARD}
    st.update({f: float(g["vb_" + f + "0"]) for f in VB_SCAL})
    return st


def check_vb(name, n_it):
    g = np.load(os.path.join(GOLD, name))
    R, M = g["R"], g["M"]
    st = vb_state(g)
    for it in range(n_it):
        orc.vb_iteration(R, M, st, HYPER, True)
        mse = orc.metrics(M, R, st["exp_U"] @ st["exp_V"].T)["MSE"]
        np.testing.assert_allclose(mse, g["vb_MSE"][it], rtol=1e-8)
        with np.errstate(all="ignore"):
            elbo = orc.vb_elbo(R, M, st, HYPER, True)
        if np.isfinite(g["vb_elbo"][it]):
            np.testing.assert_allclose(elbo, g["vb_elbo"][it], rtol=1e-8)
        else:
            assert elbo == g["vb_elbo"][it]
        np.testing.assert_allclose(st["exp_tau"], g["vb_all_exp_tau"][it], rtol=1e-8)
    return g


def test_oracle_vb_on_the_toy_matrix():
    check_vb("config_c1.npz", 25)


def test_oracle_vb_on_gdsc_including_the_elbo_underflow():
    g = check_vb("config_c2.npz", 24)      # the reference's ELBO is -inf from iteration 21 on
    assert np.isfinite(g["vb_elbo"][:20]).all() and not np.isfinite(g["vb_elbo"][20:24]).any()


def test_oracle_icm_and_np_on_the_toy_matrix():
    g = np.load(os.path.join(GOLD, "config_c1.npz"))
    R, M = g["R"], g["M"]
    U, V, tau, lam = g["icm_U0"].copy(), g["icm_V0"].copy(), float(g["icm_tau0"]), g["icm_lambdak0"].copy()
    for it in range(25):
        tau, lam = orc.icm_iteration(R, M, U, V, tau, lam, HYPER, True)
        np.testing.assert_allclose(orc.metrics(M, R, U @ V.T)["MSE"], g["icm_MSE"][it], rtol=1e-8)
        np.testing.assert_allclose(tau, g["icm_all_tau"][it], rtol=1e-8)
        np.testing.assert_allclose(lam, g["icm_all_lambdak"][it], rtol=1e-8)
    U, V = g["np_U0"].copy(), g["np_V0"].copy()
    for it in range(25):
        orc.np_iteration(g["np_R"], M, U, V)
        np.testing.assert_allclose(orc.metrics(M, g["np_R"], U @ V.T)["MSE"], g["np_MSE"][it], rtol=1e-8)


def test_chain_fixture_agrees_with_the_published_numbers():
    """Toy training MSE ~1.026 after 500 iterations on the fully observed matrix
    (experiments/experiments_toy/convergence/results/nmf_gibbs_all_performances.txt; ours hold out a tenth); GDSC
    10-fold held-out MSE 708.55 with folds between 694.9 and 725.4
    (experiments/experiments_gdsc/cross_validation/nmf_gibbs_ard/results.txt)."""
    g = np.load(os.path.join(GOLD, "config_chains.npz"))
    assert g["toy_perf"].shape == (10, 3) and g["gdsc_perf"].shape == (10, 3)
    assert 0.9 < g["toy_train_mse"].mean() < 1.05
    assert 670 < g["gdsc_perf"][:, 0].mean() < 730
