"""CPU: pin the numpy oracle (oracle/nmf_oracle.py) to golden vectors produced
by the real reference (oracle/gen_golden.py)."""
import numpy as np
import pytest

import nmf_oracle as orc

RTOL = 1e-11


def close(a, b, rtol=RTOL, atol=0.0):
    np.testing.assert_allclose(np.asarray(a, dtype=float), np.asarray(b, dtype=float), rtol=rtol, atol=atol)


def hyper_of(g, p):
    h = g[p + "hyper"]
    return {"alphatau": h[0], "betatau": h[1], "alpha0": h[2], "beta0": h[3]}


@pytest.mark.parametrize("name", ["a", "b"])
@pytest.mark.parametrize("ard", [True, False])
def test_gibbs_conditionals(golden, name, ard):
    g, p = golden, "gibbs_%s_%s_" % (name, "ard" if ard else "noard")
    R, M, U, V, tau = g[p + "R"], g[p + "M"], g[p + "U"], g[p + "V"], float(g[p + "tau"])
    I, J = R.shape
    K = U.shape[1]
    lamU = np.broadcast_to(g[p + "lambdak"], (I, K)) if ard else g[p + "lamU"]
    lamV = np.broadcast_to(g[p + "lambdak"], (J, K)) if ard else g[p + "lamV"]
    tU, mU, tV, mV = orc.gibbs_params_all(R, M, U, V, tau, lamU, lamV)
    close(tU, g[p + "tauU"]); close(tV, g[p + "tauV"])
    # mu = (-lam + tau*b)/a can cancel: compare on the posterior-sd scale too
    close(mU, g[p + "muU"], rtol=1e-9, atol=1e-12); close(mV, g[p + "muV"], rtol=1e-9, atol=1e-12)
    h = hyper_of(g, p)
    close(orc.beta_s(R, M, U, V, h["betatau"]), g[p + "beta_s"])
    if ard:
        close([orc.betak_s(U, V, h["beta0"], k) for k in range(K)], g[p + "betak_s"])


@pytest.mark.parametrize("ard", [True, False])
def test_icm_trajectory(golden, ard):
    g, p = golden, "icm_%s_" % ("ard" if ard else "noard")
    R, M = g[p + "R"], g[p + "M"]
    U, V, tau, lam = g[p + "U0"].copy(), g[p + "V0"].copy(), float(g[p + "tau0"]), g[p + "lambdak0"].copy()
    h = hyper_of(g, p)
    for it in range(int(g[p + "its"])):
        tau, lam = orc.icm_iteration(R, M, U, V, tau, lam, h, ard, g[p + "lamU"], g[p + "lamV"])
        close(U, g[p + "all_U"][it], rtol=1e-9); close(V, g[p + "all_V"][it], rtol=1e-9)
        close(tau, g[p + "all_tau"][it], rtol=1e-9)
        if ard:
            close(lam, g[p + "all_lambdak"][it], rtol=1e-9)
        perf = orc.metrics(M, R, U @ V.T)
        close([perf["MSE"], perf["R^2"], perf["Rp"]], [g[p + "MSE"][it], g[p + "R2"][it], g[p + "Rp"][it]], rtol=1e-9)


@pytest.mark.parametrize("ard", [True, False])
def test_vb_trajectory(golden, ard):
    g, p = golden, "vb_%s_" % ("ard" if ard else "noard")
    R, M = g[p + "R"], g[p + "M"]
    h = hyper_of(g, p)
    fields = ["mu_U", "tau_U", "exp_U", "var_U", "mu_V", "tau_V", "exp_V", "var_V",
              "exp_tau", "exp_logtau", "alpha_s", "beta_s"]
    if ard:
        fields += ["alphak_s", "betak_s", "exp_lambdak", "exp_loglambdak"]
    st = {f: (g[p + f + "0"].copy() if g[p + f + "0"].ndim else float(g[p + f + "0"])) for f in fields}
    close(orc.vb_exp_square_diff(R, M, st["exp_U"], st["var_U"], st["exp_V"], st["var_V"]), g[p + "esd0"])
    close(orc.vb_elbo(R, M, st, h, ard, g[p + "lamU"], g[p + "lamV"]), g[p + "elbo0"], rtol=1e-10)
    for it in range(int(g[p + "its"])):
        orc.vb_iteration(R, M, st, h, ard, g[p + "lamU"], g[p + "lamV"])
        close(orc.vb_elbo(R, M, st, h, ard, g[p + "lamU"], g[p + "lamV"]), g[p + "elbo"][it], rtol=1e-8)
        close(st["exp_tau"], g[p + "all_exp_tau"][it], rtol=1e-8)
        perf = orc.metrics(M, R, st["exp_U"] @ st["exp_V"].T)
        close(perf["MSE"], g[p + "MSE"][it], rtol=1e-8)
    for f in fields:
        close(st[f], g[p + f], rtol=1e-7, atol=1e-12)


def test_vb_single_step(golden):
    for ard in (True, False):
        g, p = golden, "vb_%s_" % ("ard" if ard else "noard")
        R, M = g[p + "R"], g[p + "M"]
        I, J = R.shape
        K = g[p + "exp_U0"].shape[1]
        st = {f: g[p + f + "0"].copy() for f in ["mu_U", "tau_U", "exp_U", "var_U", "mu_V", "tau_V", "exp_V", "var_V"]}
        exp_tau = float(g[p + "exp_tau0"])
        lamU = np.broadcast_to(g[p + "exp_lambdak0"], (I, K)) if ard else g[p + "lamU"]
        lamV = np.broadcast_to(g[p + "exp_lambdak0"], (J, K)) if ard else g[p + "lamV"]
        # update_U(1) from the initial state, then update_V(2) (gen_golden.gen_vb_run)
        E = orc.residual(R, M, st["exp_U"], st["exp_V"])
        k = 1
        y = st["exp_V"][:, k]; y2 = st["var_V"][:, k] + y * y
        t = exp_tau * (M @ y2); b = E @ y + st["exp_U"][:, k] * (M @ (y * y))
        m = (-lamU[:, k] + exp_tau * b) / t
        close(t, g[p + "step_tau_U"][:, k]); close(m, g[p + "step_mu_U"][:, k], rtol=1e-9, atol=1e-12)
        close(orc.tn_expectation(m, t), g[p + "step_exp_U"][:, k], rtol=1e-9, atol=1e-300)
        close(orc.tn_variance(m, t), g[p + "step_var_U"][:, k], rtol=1e-9, atol=1e-300)


def test_np_trajectory(golden):
    g = golden
    R, M = g["np_R"], g["np_M"]
    U, V = g["np_U0"].copy(), g["np_V0"].copy()
    for _ in range(int(g["np_its"])):
        orc.np_iteration(R, M, U, V)
    close(U, g["np_U"], rtol=1e-8); close(V, g["np_V"], rtol=1e-8)
    close(orc.np_i_div(R, M, U, V), g["np_idiv"], rtol=1e-8)
    perf = orc.metrics(M, R, U @ V.T)
    close([perf["MSE"], perf["R^2"], perf["Rp"]], [g["np_MSE"][-1], g["np_R2"][-1], g["np_Rp"][-1]], rtol=1e-8)


def test_distributions(golden):
    g = golden
    mu, tau = g["dist_mu"], g["dist_tau"]
    close(orc.tn_expectation(mu, tau), g["dist_tn_exp_vec"], rtol=1e-12, atol=0)
    close(orc.tn_variance(mu, tau), g["dist_tn_var_vec"], rtol=1e-12, atol=0)
    close(orc.tn_expectation(mu, tau), g["dist_tn_exp"], rtol=1e-12, atol=0)
    close(orc.tn_variance(mu, tau), g["dist_tn_var"], rtol=1e-12, atol=0)
    ab = g["dist_gamma_ab"]
    close([orc.gamma_expectation(a, b) for a, b in ab], g["dist_gamma_exp"], rtol=0)
    close([orc.gamma_expectation_log(a, b) for a, b in ab], g["dist_gamma_explog"], rtol=1e-15)
    close([orc.gamma_mode(a, b) for a, b in ab], g["dist_gamma_mode"], rtol=0)
    # reference literals (tests/code/distributions/test_gamma.py:11-37, test_truncated_normal.py:12-38)
    assert orc.gamma_expectation_log(2.0, 3.0) == -0.67582795356964265
    assert orc.gamma_expectation(2.0, 3.0) == 2.0 / 3.0 and orc.gamma_mode(2.0, 3.0) == 1.0 / 3.0
    assert orc.tn_variance(-1.0, 2000.0) == pytest.approx((1.0 / 2000.0) ** 2, rel=1e-15)
    assert orc.tn_expectation(-1.0, 2000.0) == pytest.approx(1.0 / 2000.0, rel=1e-15)


def test_metrics(golden):
    g = golden
    perf = orc.metrics(g["met_M"], g["met_R"], g["met_P"])
    close([perf["MSE"], perf["R^2"], perf["Rp"]], g["met_vals"], rtol=1e-13)
