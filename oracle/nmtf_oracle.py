"""
TEST INFRASTRUCTURE ONLY -- CPU oracle for the BNMTF / NMTF hot path
(R ~ F S G^T).  numpy float64 restatement in residual form of
code/models/bnmtf_gibbs.py and nmtf_icm.py; pinned against golden vectors from
the real reference (oracle/gen_golden_nmtf.py -> tests/golden/nmtf_golden.npz,
tests/test_oracle_golden_nmtf.py).  Never imported by bnmtf_ard_b200/.
"""
import numpy as np

from nmf_oracle import MINIMUM_TN, col_params, gamma_mode, tn_mode  # noqa: F401


def residual(R, M, F, S, G):
    return M * (R - F @ S @ G.T)


def beta_s(R, M, F, S, G, betatau):
    """bnmtf_gibbs.py:248-250."""
    return betatau + 0.5 * (residual(R, M, F, S, G) ** 2).sum()


def f_params(E, M, F, S, G, tau, lamF, k):
    """tauF(k), muF(.,k)  (bnmtf_gibbs.py:268-275): the NMF column conditional
    with the other factor replaced by W = G S^T (column k: w = S[k] G^T)."""
    W = G @ S.T
    return col_params(E, M, F, W, tau, lamF, k)


def g_params(E, M, F, S, G, tau, lamG, l):
    """tauG(l), muG(.,l)  (bnmtf_gibbs.py:285-292) with Z = F S."""
    Z = F @ S
    return col_params(E.T, M.T, G, Z, tau, lamG, l)


def s_params(E, M, F, S, G, tau, lamS, k, l):
    """tauS(k,l), muS(.,k,l)  (bnmtf_gibbs.py:277-283)."""
    q = np.outer(F[:, k], G[:, l])
    a = tau * (M * q * q).sum()
    b = (M * (E + S[k, l] * q) * q).sum()
    return a, (-lamS[k, l] + tau * b) / a


def params_all(R, M, F, S, G, tau, lamF, lamS, lamG):
    """Every conditional from ONE fixed state (what the reference tests call)."""
    E = residual(R, M, F, S, G)
    K, L = S.shape
    tF, mF = np.empty_like(F), np.empty_like(F)
    tG, mG = np.empty_like(G), np.empty_like(G)
    tS, mS = np.empty_like(S), np.empty_like(S)
    for k in range(K):
        tF[:, k], mF[:, k] = f_params(E, M, F, S, G, tau, lamF[:, k], k)
    for l in range(L):
        tG[:, l], mG[:, l] = g_params(E, M, F, S, G, tau, lamG[:, l], l)
    for k in range(K):
        for l in range(L):
            tS[k, l], mS[k, l] = s_params(E, M, F, S, G, tau, lamS, k, l)
    return tF, mF, tS, mS, tG, mG


def _lam(lam, n, K):
    lam = np.asarray(lam, dtype=np.float64)
    return np.broadcast_to(lam[None, :], (n, K)) if lam.ndim == 1 else lam


def iteration(R, M, F, S, G, tau, lamF, lamS, lamG, new_F, new_S, new_G):
    """The F, S, G sweeps of one iteration (bnmtf_gibbs.py:193-211 / nmtf_icm.py:194-214).
    new_X(indices..., mu, a) returns the new value(s); state is updated in place.
    Returns the conditional parameters used and the final residual sum of squares."""
    I, J = R.shape
    K, L = S.shape
    lamF, lamG = _lam(lamF, I, K), _lam(lamG, J, L)
    E = residual(R, M, F, S, G)
    tF, mF = np.empty_like(F), np.empty_like(F)
    tS, mS = np.empty_like(S), np.empty_like(S)
    tG, mG = np.empty_like(G), np.empty_like(G)
    W = G @ S.T
    for k in range(K):
        a, mu = col_params(E, M, F, W, tau, lamF[:, k], k)
        tF[:, k], mF[:, k] = a, mu
        x = np.asarray(new_F(k, mu, a), dtype=np.float64)
        E -= M * np.outer(x - F[:, k], W[:, k])
        F[:, k] = x
    for k in range(K):
        for l in range(L):
            a, mu = s_params(E, M, F, S, G, tau, lamS, k, l)
            tS[k, l], mS[k, l] = a, mu
            x = float(new_S(k, l, mu, a))
            E -= M * ((x - S[k, l]) * np.outer(F[:, k], G[:, l]))
            S[k, l] = x
    Z = F @ S
    Et, Mt = np.ascontiguousarray(E.T), np.ascontiguousarray(M.T)
    for l in range(L):
        a, mu = col_params(Et, Mt, G, Z, tau, lamG[:, l], l)
        tG[:, l], mG[:, l] = a, mu
        x = np.asarray(new_G(l, mu, a), dtype=np.float64)
        Et -= Mt * np.outer(x - G[:, l], Z[:, l])
        G[:, l] = x
    return tF, mF, tS, mS, tG, mG, float((Et ** 2).sum())


def icm_iteration(R, M, F, S, G, tau, lamFk, lamGl, hyper, ARD, lamS, lamF=None, lamG=None):
    """One iteration of nmtf_icm.run (nmtf_icm.py:186-217)."""
    I, J = R.shape
    K, L = S.shape
    if ARD:
        for k in range(K):
            lamFk[k] = gamma_mode(hyper["alpha0"] + I, hyper["beta0"] + F[:, k].sum())
        for l in range(L):
            lamGl[l] = gamma_mode(hyper["alpha0"] + J, hyper["beta0"] + G[:, l].sum())
        lF, lG = lamFk, lamGl
    else:
        lF, lG = lamF, lamG
    mode = lambda mu: np.maximum(tn_mode(mu), MINIMUM_TN)  # noqa: E731
    out = iteration(R, M, F, S, G, tau, lF, lamS, lG,
                    lambda k, mu, a: mode(mu), lambda k, l, mu, a: max(max(0.0, mu), MINIMUM_TN), lambda l, mu, a: mode(mu))
    tau = gamma_mode(hyper["alphatau"] + M.sum() / 2.0, hyper["betatau"] + 0.5 * out[-1])
    return tau


def gibbs_replay_iteration(R, M, F, S, G, tau, lamF, lamS, lamG, F_draws, S_draws, G_draws):
    """One Gibbs iteration with the draws supplied (bnmtf_gibbs.py:193-211)."""
    return iteration(R, M, F, S, G, tau, lamF, lamS, lamG,
                     lambda k, mu, a: F_draws[:, k], lambda k, l, mu, a: S_draws[k, l], lambda l, mu, a: G_draws[:, l])


# --------------------------------------------------------------------------- #
# BNMTF VB (code/models/bnmtf_vb.py)                                          #
# --------------------------------------------------------------------------- #
import math  # noqa: E402

from scipy.special import erfc, gammaln  # noqa: E402

from nmf_oracle import gamma_expectation, gamma_expectation_log, tn_expectation, tn_variance  # noqa: E402

VB_FIELDS = ["mu_F", "tau_F", "exp_F", "var_F", "mu_S", "tau_S", "exp_S", "var_S", "mu_G", "tau_G", "exp_G", "var_G"]
VB_SCAL = ["exp_tau", "exp_logtau", "alpha_s", "beta_s"]
VB_ARD = ["alphaFk_s", "betaFk_s", "exp_lambdaFk", "exp_loglambdaFk", "alphaGl_s", "betaGl_s", "exp_lambdaGl", "exp_loglambdaGl"]


def vb_exp_square_diff(R, M, st):
    """bnmtf_vb.py:319-324."""
    eF, vF, eS, vS, eG, vG = st["exp_F"], st["var_F"], st["exp_S"], st["var_S"], st["exp_G"], st["var_G"]
    t1 = (M * (R - eF @ eS @ eG.T) ** 2).sum()
    t2 = (M * ((vF + eF ** 2) @ (vS + eS ** 2) @ (vG + eG ** 2).T - (eF ** 2) @ (eS ** 2) @ (eG ** 2).T)).sum()
    t3 = (M * (vF @ ((eS @ eG.T) ** 2 - (eS ** 2) @ (eG.T ** 2)))).sum()
    t4 = (M * (((eF @ eS) ** 2 - (eF ** 2) @ (eS ** 2)) @ vG.T)).sum()
    return float(t1 + t2 + t3 + t4)


def vb_update_F(R, M, st, lamF, k):
    """bnmtf_vb.py:336-348 (tau_F[:,k], mu_F[:,k])."""
    eF, vF, eS, vS, eG, vG, et = st["exp_F"], st["var_F"], st["exp_S"], st["var_S"], st["exp_G"], st["var_G"], st["exp_tau"]
    var_SkG = (vS[k] + eS[k] ** 2) @ (vG + eG ** 2).T - (eS[k] ** 2) @ (eG ** 2).T
    w = eS[k] @ eG.T
    tau = et * ((var_SkG + w ** 2) @ M.T)
    diff = (M * ((R - eF @ eS @ eG.T + np.outer(eF[:, k], w)) * w)).sum(axis=1)
    cov = (M * ((eS[k] * (eF @ eS)) @ vG.T - np.outer(eF[:, k], (eS[k] ** 2) @ vG.T))).sum(axis=1)
    mu = (1.0 / tau) * (-lamF + et * diff - et * cov)
    return tau, mu


def vb_update_S(R, M, st, lamS, k, l):
    """bnmtf_vb.py:350-362."""
    eF, vF, eS, vS, eG, vG, et = st["exp_F"], st["var_F"], st["exp_S"], st["var_S"], st["exp_G"], st["var_G"], st["exp_tau"]
    tau = et * (M * np.outer(vF[:, k] + eF[:, k] ** 2, vG[:, l] + eG[:, l] ** 2)).sum()
    q = np.outer(eF[:, k], eG[:, l])
    diff = (M * ((R - eF @ eS @ eG.T + eS[k, l] * q) * q)).sum()
    covG = (M * np.outer(eF[:, k] * (eF @ eS[:, l] - eF[:, k] * eS[k, l]), vG[:, l])).sum()
    covF = (M * np.outer(vF[:, k], eG[:, l] * (eS[k] @ eG.T - eS[k, l] * eG[:, l]))).sum()
    mu = (1.0 / tau) * (-lamS[k, l] + et * diff - et * covG - et * covF)
    return tau, mu


def vb_update_G(R, M, st, lamG, l):
    """bnmtf_vb.py:364-375."""
    eF, vF, eS, vS, eG, vG, et = st["exp_F"], st["var_F"], st["exp_S"], st["var_S"], st["exp_G"], st["var_G"], st["exp_tau"]
    var_FSl = (vF + eF ** 2) @ (vS[:, l] + eS[:, l] ** 2) - (eF ** 2) @ (eS[:, l] ** 2)
    z = eF @ eS[:, l]
    tau = et * ((var_FSl + z ** 2).T @ M)
    diff = (M * ((R - eF @ eS @ eG.T + np.outer(z, eG[:, l])).T * z).T).sum(axis=0)
    cov = (M * (vF @ (eS[:, l] * (eS @ eG.T).T).T - np.outer(vF @ (eS[:, l] ** 2), eG[:, l]))).sum(axis=0)
    mu = (1.0 / tau) * (-lamG + et * diff - et * cov)
    return tau, mu


def vb_iteration(R, M, st, hyper, ARD, lamS, lamF=None, lamG=None):
    """One iteration of bnmtf_vb.run (bnmtf_vb.py:204-232); st updated in place."""
    I, J = R.shape
    K, L = st["exp_S"].shape
    if ARD:
        for k in range(K):
            st["alphaFk_s"][k] = hyper["alpha0"] + I
            st["betaFk_s"][k] = hyper["beta0"] + st["exp_F"][:, k].sum()
            st["exp_lambdaFk"][k] = gamma_expectation(st["alphaFk_s"][k], st["betaFk_s"][k])
            st["exp_loglambdaFk"][k] = gamma_expectation_log(st["alphaFk_s"][k], st["betaFk_s"][k])
        for l in range(L):
            st["alphaGl_s"][l] = hyper["alpha0"] + J
            st["betaGl_s"][l] = hyper["beta0"] + st["exp_G"][:, l].sum()
            st["exp_lambdaGl"][l] = gamma_expectation(st["alphaGl_s"][l], st["betaGl_s"][l])
            st["exp_loglambdaGl"][l] = gamma_expectation_log(st["alphaGl_s"][l], st["betaGl_s"][l])
    for k in range(K):
        lam = st["exp_lambdaFk"][k] if ARD else lamF[:, k]
        st["tau_F"][:, k], st["mu_F"][:, k] = vb_update_F(R, M, st, lam, k)
        st["exp_F"][:, k] = tn_expectation(st["mu_F"][:, k], st["tau_F"][:, k])
        st["var_F"][:, k] = tn_variance(st["mu_F"][:, k], st["tau_F"][:, k])
    for k in range(K):
        for l in range(L):
            st["tau_S"][k, l], st["mu_S"][k, l] = vb_update_S(R, M, st, lamS, k, l)
            st["exp_S"][k, l] = float(tn_expectation(st["mu_S"][k, l], st["tau_S"][k, l]))
            st["var_S"][k, l] = float(tn_variance(st["mu_S"][k, l], st["tau_S"][k, l]))
    for l in range(L):
        lam = st["exp_lambdaGl"][l] if ARD else lamG[:, l]
        st["tau_G"][:, l], st["mu_G"][:, l] = vb_update_G(R, M, st, lam, l)
        st["exp_G"][:, l] = tn_expectation(st["mu_G"][:, l], st["tau_G"][:, l])
        st["var_G"][:, l] = tn_variance(st["mu_G"][:, l], st["tau_G"][:, l])
    st["alpha_s"] = hyper["alphatau"] + M.sum() / 2.0
    st["beta_s"] = hyper["betatau"] + 0.5 * vb_exp_square_diff(R, M, st)
    st["exp_tau"] = gamma_expectation(st["alpha_s"], st["beta_s"])
    st["exp_logtau"] = gamma_expectation_log(st["alpha_s"], st["beta_s"])


def vb_elbo(R, M, st, hyper, ARD, lamS, lamF=None, lamG=None):
    """bnmtf_vb.py:248-299."""
    I, J = R.shape
    K, L = st["exp_S"].shape
    total = M.sum() / 2.0 * (st["exp_logtau"] - math.log(2 * math.pi)) - st["exp_tau"] / 2.0 * vb_exp_square_diff(R, M, st)
    a0, b0 = hyper["alpha0"], hyper["beta0"]
    if ARD:
        for X, n, e in (("Fk", I, "exp_F"), ("Gl", J, "exp_G")):
            el, ell = st["exp_lambda" + X], st["exp_loglambda" + X]
            total += a0 * math.log(b0) - gammaln(a0) + (a0 - 1.0) * ell.sum() - b0 * el.sum()
            total += n * np.log(el).sum() - (el * st[e]).sum()
    else:
        total += np.log(lamF).sum() - (lamF * st["exp_F"]).sum()
        total += np.log(lamG).sum() - (lamG * st["exp_G"]).sum()
    total += np.log(lamS).sum() - (lamS * st["exp_S"]).sum()
    at, bt = hyper["alphatau"], hyper["betatau"]
    total += at * math.log(bt) - gammaln(at) + (at - 1.0) * st["exp_logtau"] - bt * st["exp_tau"]
    if ARD:
        for X in ("Fk", "Gl"):
            al, be = st["alpha" + X + "_s"], st["beta" + X + "_s"]
            total += (-(al * np.log(be)).sum() + gammaln(al).sum() - ((al - 1.0) * st["exp_loglambda" + X]).sum()
                      + (be * st["exp_lambda" + X]).sum())
    for X, n in (("F", I * K), ("G", J * L), ("S", K * L)):
        mu, tau_, e, v = st["mu_" + X], st["tau_" + X], st["exp_" + X], st["var_" + X]
        total += (-0.5 * np.log(tau_).sum() + n / 2.0 * math.log(2 * math.pi)
                  + np.log(0.5 * erfc(-mu * np.sqrt(tau_) / math.sqrt(2))).sum()
                  + (tau_ / 2.0 * (v + (e - mu) ** 2)).sum())
    total += (-st["alpha_s"] * math.log(st["beta_s"]) + gammaln(st["alpha_s"])
              - (st["alpha_s"] - 1.0) * st["exp_logtau"] + st["beta_s"] * st["exp_tau"])
    return float(total)


# --------------------------------------------------------------------------- #
# NMTF NP (code/models/nmtf_np.py:139-195)                                    #
# --------------------------------------------------------------------------- #
def np_iteration(R, M, F, S, G):
    """One iteration of nmtf_np.run: update_F(k) for all k, update_S(k,l) for all (k,l),
    update_G(l) for all l, each with R_pred recomputed (nmtf_np.py:148-155,173-195)."""
    K, L = S.shape
    with np.errstate(all="ignore"):
        for k in range(K):
            P = F @ S @ G.T
            SG = S[k] @ G.T
            F[:, k] = F[:, k] * (M * R / P * SG).sum(axis=1) / (M * SG).sum(axis=1)
        for k in range(K):
            for l in range(L):
                P = F @ S @ G.T
                FG = M * np.outer(F[:, k], G[:, l])
                S[k, l] = S[k, l] * (R * FG / P).sum() / FG.sum()
        for l in range(L):
            P = F @ S @ G.T
            FS = F @ S[:, l]
            G[:, l] = G[:, l] * ((M * R / P).T * FS).T.sum(axis=0) / (M.T * FS).T.sum(axis=0)


def np_i_div(R, M, F, S, G):
    """nmtf_np.py:222-225."""
    Rx = np.where(M != 0, R, 1.0)
    P = F @ S @ G.T
    with np.errstate(all="ignore"):
        return float((M * (Rx * np.log(Rx / P) - Rx + P)).sum())
