"""CPU: pin the NMTF numpy oracle to golden vectors produced by the real reference."""
import numpy as np
import pytest

import nmtf_oracle as orc


@pytest.fixture(scope="module")
def gold():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nmtf_golden.npz"))


def close(a, b, rtol=1e-10, atol=0.0):
    np.testing.assert_allclose(np.asarray(a, dtype=float), np.asarray(b, dtype=float), rtol=rtol, atol=atol)


@pytest.mark.parametrize("name", ["a", "b"])
@pytest.mark.parametrize("ard", [True, False])
def test_gibbs_conditionals(gold, name, ard):
    g, p = gold, "tg_%s_%s_" % (name, "ard" if ard else "noard")
    R, M, F, S, G, tau = g[p + "R"], g[p + "M"], g[p + "F"], g[p + "S"], g[p + "G"], float(g[p + "tau"])
    I, J = R.shape
    K, L = S.shape
    lamF = np.broadcast_to(g[p + "lamFk"], (I, K)) if ard else g[p + "lamF"]
    lamG = np.broadcast_to(g[p + "lamGl"], (J, L)) if ard else g[p + "lamG"]
    tF, mF, tS, mS, tG, mG = orc.params_all(R, M, F, S, G, tau, lamF, g[p + "lamS"], lamG)
    close(tF, g[p + "tauF"]); close(tS, g[p + "tauS"]); close(tG, g[p + "tauG"])
    close(mF, g[p + "muF"], 1e-9, 1e-11); close(mS, g[p + "muS"], 1e-9, 1e-11); close(mG, g[p + "muG"], 1e-9, 1e-11)
    close(orc.beta_s(R, M, F, S, G, g[p + "hyper"][1]), g[p + "beta_s"])


@pytest.mark.parametrize("ard", [True, False])
def test_icm_trajectory(gold, ard):
    g, p = gold, "ti_%s_" % ("ard" if ard else "noard")
    R, M = g[p + "R"], g[p + "M"]
    F, S, G, tau = g[p + "F0"].copy(), g[p + "S0"].copy(), g[p + "G0"].copy(), float(g[p + "tau0"])
    lF, lG = g[p + "lamFk0"].copy(), g[p + "lamGl0"].copy()
    h = g[p + "hyper"]
    hyper = {"alphatau": h[0], "betatau": h[1], "alpha0": h[2], "beta0": h[3]}
    for it in range(int(g[p + "its"])):
        tau = orc.icm_iteration(R, M, F, S, G, tau, lF, lG, hyper, ard, g[p + "lamS"], g[p + "lamF"], g[p + "lamG"])
        close(F, g[p + "all_F"][it], 1e-8); close(S, g[p + "all_S"][it], 1e-8); close(G, g[p + "all_G"][it], 1e-8)
        close(tau, g[p + "all_tau"][it], 1e-8)
        if ard:
            close(lF, g[p + "all_lambdaFk"][it], 1e-8); close(lG, g[p + "all_lambdaGl"][it], 1e-8)


@pytest.mark.parametrize("ard", [True, False])
def test_vb_trajectory_and_steps(gold, ard):
    g, p = gold, "tv_%s_" % ("ard" if ard else "noard")
    R, M, lamS = g[p + "R"], g[p + "M"], g[p + "lamS"]
    h = g[p + "hyper"]
    hyper = {"alphatau": h[0], "betatau": h[1], "alpha0": h[2], "beta0": h[3]}
    fields = orc.VB_FIELDS + orc.VB_SCAL + (orc.VB_ARD if ard else [])
    st = {f: (g[p + f + "0"].copy() if g[p + f + "0"].ndim else float(g[p + f + "0"])) for f in fields}
    close(orc.vb_exp_square_diff(R, M, st), g[p + "esd0"])
    close(orc.vb_elbo(R, M, st, hyper, ard, lamS, g[p + "lamF"], g[p + "lamG"]), g[p + "elbo0"], 1e-10)
    # single-step F(1), S(2,1), G(2) as generated
    s2 = {f: (np.array(v).copy() if np.ndim(v) else v) for f, v in st.items()}
    lam = s2["exp_lambdaFk"][1] if ard else g[p + "lamF"][:, 1]
    s2["tau_F"][:, 1], s2["mu_F"][:, 1] = orc.vb_update_F(R, M, s2, lam, 1)
    s2["exp_F"][:, 1] = orc.tn_expectation(s2["mu_F"][:, 1], s2["tau_F"][:, 1])
    s2["var_F"][:, 1] = orc.tn_variance(s2["mu_F"][:, 1], s2["tau_F"][:, 1])
    s2["tau_S"][2, 1], s2["mu_S"][2, 1] = orc.vb_update_S(R, M, s2, lamS, 2, 1)
    s2["exp_S"][2, 1] = float(orc.tn_expectation(s2["mu_S"][2, 1], s2["tau_S"][2, 1]))
    s2["var_S"][2, 1] = float(orc.tn_variance(s2["mu_S"][2, 1], s2["tau_S"][2, 1]))
    lam = s2["exp_lambdaGl"][2] if ard else g[p + "lamG"][:, 2]
    s2["tau_G"][:, 2], s2["mu_G"][:, 2] = orc.vb_update_G(R, M, s2, lam, 2)
    for f in ("tau_F", "mu_F", "exp_F", "tau_S", "mu_S", "exp_S", "tau_G", "mu_G"):
        close(s2[f], g[p + "step_" + f], 1e-9, 1e-12)
    for it in range(int(g[p + "its"])):
        orc.vb_iteration(R, M, st, hyper, ard, lamS, g[p + "lamF"], g[p + "lamG"])
        close(orc.vb_elbo(R, M, st, hyper, ard, lamS, g[p + "lamF"], g[p + "lamG"]), g[p + "elbo"][it], 1e-8)
        close(st["exp_tau"], g[p + "all_exp_tau"][it], 1e-8)
    for f in fields:
        close(st[f], g[p + f], 1e-7, 1e-12)


def test_np_trajectory(gold):
    g = gold
    R, M = g["tn_R"], g["tn_M"]
    F, S, G = g["tn_F0"].copy(), g["tn_S0"].copy(), g["tn_G0"].copy()
    for _ in range(int(g["tn_its"])):
        orc.np_iteration(R, M, F, S, G)
    close(F, g["tn_F"], 1e-8); close(S, g["tn_S"], 1e-8); close(G, g["tn_G"], 1e-8)
    close(orc.np_i_div(R, M, F, S, G), g["tn_idiv"], 1e-8)
