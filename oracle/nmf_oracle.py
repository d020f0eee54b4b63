"""
TEST INFRASTRUCTURE ONLY -- CPU oracle for the BNMF / NMF hot path.

A plain-numpy float64 restatement, in residual form, of the per-iteration
inference loops of the reference's ``code/models`` NMF classes.  Nothing under
``bnmtf_ard_b200/`` may import this module: it is the checker used by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg.

Pinning: ``tests/test_oracle_golden.py`` checks every function here against
golden vectors produced by the real reference (``oracle/gen_golden.py`` ->
``tests/golden/*.npz``) and, when ``oracle/_ref`` is present, against the live
reference.  The random *streams* of the reference (numpy MT19937 + Chopin
tables) are parity-unpinned by design; samplers are validated distributionally.

Each function cites the reference file:line it follows (paths relative to the
reference checkout).
"""
import math

import numpy as np
from scipy.special import erfc, gammaln, psi

SQRT2 = math.sqrt(2.0)
MINIMUM_TN = 0.1          # code/models/nmf_icm.py:59


# --------------------------------------------------------------------------- #
# distributions                                                               #
# --------------------------------------------------------------------------- #
def _norm_pdf(x):
    return np.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)


def _guard(v):
    """`v if (v >= 0 and finite) else 0`  (truncated_normal_vector.py:61,73)."""
    v = np.asarray(v, dtype=np.float64)
    ok = np.isfinite(v) & (v >= 0.0)
    return np.where(ok, v, 0.0)


def tn_expectation(mu, tau):
    """code/models/distributions/truncated_normal_vector.py:53-61 (and scalar
    truncated_normal.py:47-55)."""
    mu = np.asarray(mu, dtype=np.float64)
    tau = np.asarray(tau, dtype=np.float64)
    with np.errstate(all="ignore"):
        sigma = 1.0 / np.sqrt(tau)
        x = -mu / sigma
        lambdax = _norm_pdf(x) / (0.5 * erfc(x / SQRT2))
        e = mu + sigma * lambdax
        e = np.where(mu < -30.0 * sigma, 1.0 / (np.abs(mu) * tau), e)
    return _guard(e)


def tn_variance(mu, tau):
    """code/models/distributions/truncated_normal_vector.py:64-73."""
    mu = np.asarray(mu, dtype=np.float64)
    tau = np.asarray(tau, dtype=np.float64)
    with np.errstate(all="ignore"):
        sigma = 1.0 / np.sqrt(tau)
        x = -mu / sigma
        lambdax = _norm_pdf(x) / (0.5 * erfc(x / SQRT2))
        deltax = lambdax * (lambdax - x)
        v = sigma ** 2 * (1.0 - deltax)
        v = np.where(mu < -30.0 * sigma, (1.0 / (np.abs(mu) * tau)) ** 2, v)
    return _guard(v)


def tn_mode(mu):
    """truncated_normal_vector.py:76-78."""
    return np.maximum(0.0, np.asarray(mu, dtype=np.float64))


def tn_cdf(x, mu, tau):
    """Analytic CDF of N(mu, 1/tau) truncated to [0, inf): the distributional
    contract of TN_draw (truncated_normal.py:37-44, rtnorm.py:16-87)."""
    x = np.asarray(x, dtype=np.float64)
    sigma = 1.0 / math.sqrt(tau)
    a = -mu / sigma
    z = (x - mu) / sigma
    # P(a <= Z <= z) / P(Z >= a), in the numerically safe upper-tail form
    den = 0.5 * erfc(a / SQRT2)
    num = den - 0.5 * erfc(z / SQRT2)
    if den == 0.0 or a > 6.0:
        # far tail: use the ratio of erfc's through log1p-safe difference
        from scipy.special import log_ndtr
        lr = log_ndtr(-z) - log_ndtr(-a)
        return np.where(x < 0, 0.0, -np.expm1(lr))
    return np.where(x < 0, 0.0, num / den)


def tn_cdf_vec(x, mu, tau):
    """tn_cdf for arrays of (x, mu, tau): 1 - P(Z >= z) / P(Z >= a) through log_ndtr, stable in
    both tails."""
    from scipy.special import log_ndtr
    x, mu, tau = (np.asarray(v, dtype=np.float64) for v in (x, mu, tau))
    s = np.sqrt(tau)
    lr = log_ndtr(-(x - mu) * s) - log_ndtr(mu * s)
    return np.where(x < 0, 0.0, -np.expm1(lr))


def gamma_expectation(alpha, beta):
    """code/models/distributions/gamma.py:17-19."""
    return float(alpha) / float(beta)


def gamma_expectation_log(alpha, beta):
    """gamma.py:22-24."""
    return float(psi(float(alpha))) - math.log(float(beta))


def gamma_mode(alpha, beta):
    """gamma.py:27-29."""
    return (float(alpha) - 1.0) / float(beta)


# --------------------------------------------------------------------------- #
# metrics                                                                     #
# --------------------------------------------------------------------------- #
def metrics(M, R, R_pred):
    """compute_MSE / compute_R2 / compute_Rp  (bnmf_gibbs.py:252-270)."""
    M = np.asarray(M, dtype=np.float64)
    n = float(M.sum())
    mse = float((M * (R - R_pred) ** 2).sum() / n)
    mean = (M * R).sum() / n
    ss_tot = float((M * (R - mean) ** 2).sum())
    ss_res = float((M * (R - R_pred) ** 2).sum())
    r2 = 1.0 - ss_res / ss_tot if ss_tot != 0.0 else np.inf
    mean_pred = (M * R_pred).sum() / n
    cov = (M * (R - mean) * (R_pred - mean_pred)).sum()
    var_r = (M * (R - mean) ** 2).sum()
    var_p = (M * (R_pred - mean_pred) ** 2).sum()
    with np.errstate(all="ignore"):
        rp = float(cov / (math.sqrt(var_r) * math.sqrt(var_p))) if var_r * var_p != 0 else float("nan")
    return {"MSE": mse, "R^2": r2, "Rp": rp}


# --------------------------------------------------------------------------- #
# BNMF Gibbs / ICM conditionals (single column, from a given state)           #
# --------------------------------------------------------------------------- #
def residual(R, M, U, V):
    return M * (R - U @ V.T)


def beta_s(R, M, U, V, betatau):
    """bnmf_gibbs.py:191-193."""
    return betatau + 0.5 * (residual(R, M, U, V) ** 2).sum()


def betak_s(U, V, beta0, k):
    """bnmf_gibbs.py:199-201."""
    return beta0 + U[:, k].sum() + V[:, k].sum()


def col_params(E, M, X, Y, tau, lam, k):
    """Conditional (precision a, mean mu) of column k of X given the current
    masked residual E = M*(R - X Y^T).  U direction: bnmf_gibbs.py:203-210;
    V direction is the same call on transposed (E, M) (bnmf_gibbs.py:212-219).

        a_i  = tau * sum_j M_ij Y_jk^2
        mu_i = (-lam_i + tau * sum_j M_ij (E_ij + X_ik Y_jk) Y_jk) / a_i
    """
    y = Y[:, k]
    s = M @ (y * y)
    a = tau * s
    b = E @ y + X[:, k] * s
    with np.errstate(all="ignore"):
        mu = (-lam + tau * b) / a
    return a, mu


def gibbs_params_all(R, M, U, V, tau, lamU, lamV):
    """tauU(k), muU(.,k), tauV(k), muV(.,k) for every k from ONE fixed state
    (what the reference tests call column by column)."""
    E = residual(R, M, U, V)
    K = U.shape[1]
    tU, mU = np.empty_like(U), np.empty_like(U)
    tV, mV = np.empty_like(V), np.empty_like(V)
    for k in range(K):
        tU[:, k], mU[:, k] = col_params(E, M, U, V, tau, lamU[:, k], k)
        tV[:, k], mV[:, k] = col_params(E.T, M.T, V, U, tau, lamV[:, k], k)
    return tU, mU, tV, mV


def _lam_matrix(lam, n, K):
    lam = np.asarray(lam, dtype=np.float64)
    if lam.ndim == 1:               # ARD: lambdak[k] shared by every row
        return np.broadcast_to(lam[None, :], (n, K))
    return lam


def sweep(E, M, X, Y, tau, lam, new_value):
    """One K-column sweep over X (rows of E), updating E in place.

    ``new_value(k, mu, a, x_old)`` returns the new column.  Follows the loop
    bodies of bnmf_gibbs.py:155-164 / nmf_icm.py:155-167.  Returns (A, MU) --
    the conditional parameters that were used for each column.
    """
    n, K = X.shape
    lam = _lam_matrix(lam, n, K)
    A, MU = np.empty_like(X), np.empty_like(X)
    for k in range(K):
        a, mu = col_params(E, M, X, Y, tau, lam[:, k], k)
        A[:, k], MU[:, k] = a, mu
        x_new = np.asarray(new_value(k, mu, a, X[:, k]), dtype=np.float64)
        delta = x_new - X[:, k]
        E -= M * np.outer(delta, Y[:, k])
        X[:, k] = x_new
    return A, MU


def icm_value(k, mu, a, x_old):
    """nmf_icm.py:159-160: mode then floor at MINIMUM_TN."""
    return np.maximum(tn_mode(mu), MINIMUM_TN)


def icm_iteration(R, M, U, V, tau, lambdak, hyper, ARD, lamU=None, lamV=None):
    """One iteration of nmf_icm.run (nmf_icm.py:148-172).  State is updated in
    place; returns the new tau and lambdak."""
    I, J = R.shape
    K = U.shape[1]
    if ARD:
        for k in range(K):
            lambdak[k] = gamma_mode(hyper["alpha0"] + I + J, betak_s(U, V, hyper["beta0"], k))
        lU = lV = lambdak
    else:
        lU, lV = lamU, lamV
    E = residual(R, M, U, V)
    sweep(E, M, U, V, tau, lU, icm_value)
    Et = np.ascontiguousarray(E.T)
    sweep(Et, np.ascontiguousarray(M.T), V, U, tau, lV, icm_value)
    tau = gamma_mode(hyper["alphatau"] + M.sum() / 2.0, hyper["betatau"] + 0.5 * (Et ** 2).sum())
    return tau, lambdak


def gibbs_replay_iteration(R, M, U, V, tau, lamU, lamV, U_draws, V_draws):
    """Replay of one Gibbs iteration's U and V sweeps (bnmf_gibbs.py:155-164)
    with the *draws supplied* (the sampler stream is unpinned): returns the
    conditional parameters (tauU, muU, tauV, muV) used at every column, and
    the residual sum of squares afterwards.  U, V are updated in place."""
    E = residual(R, M, U, V)
    tU, mU = sweep(E, M, U, V, tau, lamU, lambda k, mu, a, x: U_draws[:, k])
    Et = np.ascontiguousarray(E.T)
    tV, mV = sweep(Et, np.ascontiguousarray(M.T), V, U, tau, lamV, lambda k, mu, a, x: V_draws[:, k])
    return tU, mU, tV, mV, float((Et ** 2).sum())


# --------------------------------------------------------------------------- #
# BNMF VB                                                                     #
# --------------------------------------------------------------------------- #
def vb_exp_square_diff(R, M, eU, vU, eV, vV):
    """bnmf_vb.py:241-244."""
    return float((M * ((R - eU @ eV.T) ** 2
                       + ((vU + eU ** 2) @ (vV + eV ** 2).T - (eU ** 2) @ (eV ** 2).T))).sum())


def vb_sweep(E, M, muX, tauX, eX, vX, eY, vY, exp_tau, lam):
    """update_U(k)+update_exp_U(k) for k=0..K-1 (bnmf_vb.py:251-255,275-278);
    E is the masked residual of the *expectations* and is updated in place."""
    n, K = eX.shape
    lam = _lam_matrix(lam, n, K)
    for k in range(K):
        y = eY[:, k]
        y2 = vY[:, k] + y * y
        t = exp_tau * (M @ y2)
        b = E @ y + eX[:, k] * (M @ (y * y))
        with np.errstate(all="ignore"):
            m = (-lam[:, k] + exp_tau * b) / t
        tauX[:, k], muX[:, k] = t, m
        e_new, v_new = tn_expectation(m, t), tn_variance(m, t)
        E -= M * np.outer(e_new - eX[:, k], y)
        eX[:, k], vX[:, k] = e_new, v_new


def vb_iteration(R, M, st, hyper, ARD, lamU=None, lamV=None):
    """One iteration of bnmf_vb.run (bnmf_vb.py:155-173).  ``st`` is a dict with
    mu_U,tau_U,exp_U,var_U,mu_V,tau_V,exp_V,var_V,exp_tau,exp_logtau,alpha_s,
    beta_s and (ARD) alphak_s,betak_s,exp_lambdak,exp_loglambdak; updated in
    place."""
    I, J = R.shape
    K = st["exp_U"].shape[1]
    if ARD:
        for k in range(K):
            st["alphak_s"][k] = hyper["alpha0"] + I + J
            st["betak_s"][k] = hyper["beta0"] + st["exp_U"][:, k].sum() + st["exp_V"][:, k].sum()
            st["exp_lambdak"][k] = gamma_expectation(st["alphak_s"][k], st["betak_s"][k])
            st["exp_loglambdak"][k] = gamma_expectation_log(st["alphak_s"][k], st["betak_s"][k])
        lU = lV = st["exp_lambdak"]
    else:
        lU, lV = lamU, lamV
    E = residual(R, M, st["exp_U"], st["exp_V"])
    vb_sweep(E, M, st["mu_U"], st["tau_U"], st["exp_U"], st["var_U"], st["exp_V"], st["var_V"], st["exp_tau"], lU)
    Et, Mt = np.ascontiguousarray(E.T), np.ascontiguousarray(M.T)
    vb_sweep(Et, Mt, st["mu_V"], st["tau_V"], st["exp_V"], st["var_V"], st["exp_U"], st["var_U"], st["exp_tau"], lV)
    st["alpha_s"] = hyper["alphatau"] + M.sum() / 2.0
    st["beta_s"] = hyper["betatau"] + 0.5 * vb_exp_square_diff(
        R, M, st["exp_U"], st["var_U"], st["exp_V"], st["var_V"])
    st["exp_tau"] = gamma_expectation(st["alpha_s"], st["beta_s"])
    st["exp_logtau"] = gamma_expectation_log(st["alpha_s"], st["beta_s"])


def vb_elbo(R, M, st, hyper, ARD, lamU=None, lamV=None):
    """bnmf_vb.py:191-232."""
    I, J = R.shape
    K = st["exp_U"].shape[1]
    size_Omega = M.sum()
    esd = vb_exp_square_diff(R, M, st["exp_U"], st["var_U"], st["exp_V"], st["var_V"])
    total = size_Omega / 2.0 * (st["exp_logtau"] - math.log(2 * math.pi)) - st["exp_tau"] / 2.0 * esd
    if ARD:
        a0, b0 = hyper["alpha0"], hyper["beta0"]
        total += a0 * math.log(b0) - gammaln(a0) + (a0 - 1.0) * st["exp_loglambdak"].sum() - b0 * st["exp_lambdak"].sum()
        total += I * np.log(st["exp_lambdak"]).sum() - (st["exp_lambdak"] * st["exp_U"]).sum()
        total += J * np.log(st["exp_lambdak"]).sum() - (st["exp_lambdak"] * st["exp_V"]).sum()
    else:
        total += np.log(lamU).sum() - (lamU * st["exp_U"]).sum()
        total += np.log(lamV).sum() - (lamV * st["exp_V"]).sum()
    at, bt = hyper["alphatau"], hyper["betatau"]
    total += at * math.log(bt) - gammaln(at) + (at - 1.0) * st["exp_logtau"] - bt * st["exp_tau"]
    if ARD:
        total += (- (st["alphak_s"] * np.log(st["betak_s"])).sum() + gammaln(st["alphak_s"]).sum()
                  - ((st["alphak_s"] - 1.0) * st["exp_loglambdak"]).sum() + (st["betak_s"] * st["exp_lambdak"]).sum())
    for X, n in (("U", I), ("V", J)):
        mu, tau_, e, v = st["mu_" + X], st["tau_" + X], st["exp_" + X], st["var_" + X]
        total += (-0.5 * np.log(tau_).sum() + n * K / 2.0 * math.log(2 * math.pi)
                  + np.log(0.5 * erfc(-mu * np.sqrt(tau_) / SQRT2)).sum()
                  + (tau_ / 2.0 * (v + (e - mu) ** 2)).sum())
    total += (-st["alpha_s"] * math.log(st["beta_s"]) + gammaln(st["alpha_s"])
              - (st["alpha_s"] - 1.0) * st["exp_logtau"] + st["beta_s"] * st["exp_tau"])
    return float(total)


# --------------------------------------------------------------------------- #
# NMF NP (multiplicative updates)                                             #
# --------------------------------------------------------------------------- #
def np_sweep(Rm, M, P, X, Y):
    """update_U(k) for k=0..K-1 (nmf_np.py:120-122) with P = X Y^T maintained
    by rank-1 corrections instead of recomputed; V direction: transposed call
    (nmf_np.py:124-126).  Rm = M*R."""
    K = X.shape[1]
    for k in range(K):
        y = Y[:, k]
        with np.errstate(all="ignore"):
            num = (M * (Rm / P)) @ y
            den = M @ y
            x_new = X[:, k] * num / den
        P += np.outer(x_new - X[:, k], y)
        X[:, k] = x_new


def np_iteration(R, M, U, V):
    """One iteration of nmf_np.run (nmf_np.py:107-111); U, V in place."""
    Rm = M * R
    P = U @ V.T
    np_sweep(Rm, M, P, U, V)
    Pt = np.ascontiguousarray(P.T)
    np_sweep(np.ascontiguousarray(Rm.T), np.ascontiguousarray(M.T), Pt, V, U)


def np_i_div(R, M, U, V):
    """nmf_np.py:160-163."""
    Rx = np.where(M != 0, R, 1.0)
    P = U @ V.T
    with np.errstate(all="ignore"):
        return float((M * (Rx * np.log(Rx / P) - Rx + P)).sum())
