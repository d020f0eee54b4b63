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
