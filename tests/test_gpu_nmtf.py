"""GPU parity tests for the tri-factorisation path (bnmtf_gibbs, nmtf_icm): golden
vectors from the real reference and the numpy oracle on seeded inputs."""
import os

import numpy as np
import pytest

import nmtf_oracle as orc

pytestmark = pytest.mark.gpu
RTOL_STEP, RTOL_TRAJ = 1e-9, 1e-6


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nmtf_golden.npz"))


def close(a, b, rtol, atol=0.0):
    np.testing.assert_allclose(np.asarray(a, dtype=float), np.asarray(b, dtype=float), rtol=rtol, atol=atol)


def mu_close(mu, mu_ref, tau_ref, rtol=RTOL_STEP):
    mu, mu_ref = np.asarray(mu, dtype=float), np.asarray(mu_ref, dtype=float)
    scale = np.maximum(np.abs(mu_ref), 1.0 / np.sqrt(np.asarray(tau_ref, dtype=float)))
    assert np.max(np.abs(mu - mu_ref) / scale) < rtol


def hyper_of(g, p, ard):
    h = g[p + "hyper"]
    d = {"alphatau": h[0], "betatau": h[1], "alpha0": h[2], "beta0": h[3], "lambdaS": g[p + "lamS"]}
    if not ard:
        d.update(lambdaF=g[p + "lamF"], lambdaG=g[p + "lamG"])
    return d


def make_problem(seed, I, J, K, L, density):
    rng = np.random.default_rng(seed)
    R = rng.exponential(1.0, (I, K)) @ rng.exponential(1.0, (K, L)) @ rng.exponential(1.0, (J, L)).T + rng.normal(0, 1, (I, J))
    M = (rng.random((I, J)) < density).astype(float)
    M[np.arange(I), rng.integers(0, J, I)] = 1
    M[rng.integers(0, I, J), np.arange(J)] = 1
    return R, M, rng


@pytest.mark.parametrize("name", ["a", "b"])
@pytest.mark.parametrize("ard", [True, False])
def test_conditionals_golden(gold, name, ard):
    from bnmtf_ard_b200 import bnmtf_gibbs, nmtf_icm
    g, p = gold, "tg_%s_%s_" % (name, "ard" if ard else "noard")
    K, L = g[p + "S"].shape
    for cls in (bnmtf_gibbs, nmtf_icm):
        m = cls(g[p + "R"], g[p + "M"], K, L, ard, hyper_of(g, p, ard))
        m.F, m.S, m.G, m.tau = g[p + "F"].copy(), g[p + "S"].copy(), g[p + "G"].copy(), float(g[p + "tau"])
        m.lambdaFk, m.lambdaGl = g[p + "lamFk"].copy(), g[p + "lamGl"].copy()
        for k in range(K):
            t = m.tauF(k)
            close(t, g[p + "tauF"][:, k], RTOL_STEP); mu_close(m.muF(t, k), g[p + "muF"][:, k], g[p + "tauF"][:, k])
        for l in range(L):
            t = m.tauG(l)
            close(t, g[p + "tauG"][:, l], RTOL_STEP); mu_close(m.muG(t, l), g[p + "muG"][:, l], g[p + "tauG"][:, l])
        for k in range(K):
            for l in range(L):
                t = m.tauS(k, l)
                close(t, g[p + "tauS"][k, l], RTOL_STEP); mu_close(m.muS(t, k, l), g[p + "muS"][k, l], g[p + "tauS"][k, l])
        close(m.beta_s(), g[p + "beta_s"], RTOL_STEP)
        assert m.alpha_s() == float(g[p + "alpha_s"])
        if ard:
            close([m.betaFk_s(k) for k in range(K)], g[p + "betaFk_s"], 1e-13)
            close([m.betaGl_s(l) for l in range(L)], g[p + "betaGl_s"], 1e-13)


@pytest.mark.parametrize("ard", [True, False])
def test_icm_trajectory_golden(gold, ard):
    from bnmtf_ard_b200 import nmtf_icm
    g, p = gold, "ti_%s_" % ("ard" if ard else "noard")
    K, L = g[p + "S0"].shape
    m = nmtf_icm(g[p + "R"], g[p + "M"], K, L, ard, hyper_of(g, p, ard))
    m.verbose = False
    m.F, m.S, m.G, m.tau = g[p + "F0"].copy(), g[p + "S0"].copy(), g[p + "G0"].copy(), float(g[p + "tau0"])
    m.lambdaFk, m.lambdaGl = g[p + "lamFk0"].copy(), g[p + "lamGl0"].copy()
    its = int(g[p + "its"])
    m.run(its)
    close(m.all_F, g[p + "all_F"], RTOL_TRAJ); close(m.all_S, g[p + "all_S"], RTOL_TRAJ); close(m.all_G, g[p + "all_G"], RTOL_TRAJ)
    close(m.all_tau, g[p + "all_tau"], RTOL_TRAJ)
    if ard:
        close(m.all_lambdaFk, g[p + "all_lambdaFk"], RTOL_TRAJ); close(m.all_lambdaGl, g[p + "all_lambdaGl"], RTOL_TRAJ)
    close(m.all_performances["MSE"], g[p + "MSE"], RTOL_TRAJ)
    close(m.all_performances["R^2"], g[p + "R2"], RTOL_TRAJ)
    # Rp is 0/0 when the prediction is constant (all factors at the 0.1 floor): the reference
    # returns rounding noise there (1e-17), so compare only where it is defined
    ok = np.abs(g[p + "Rp"]) > 1e-9
    close(np.asarray(m.all_performances["Rp"])[ok], g[p + "Rp"][ok], RTOL_TRAJ)


# the last two: odd K and L on the per-row-chain cluster sweep (F side with the projection epilogue; G side)
SHAPES = [(40, 30, 3, 4, 0.7), (150, 220, 5, 3, 0.5), (64, 3000, 4, 6, 0.6), (700, 90, 7, 2, 0.8), (48, 2600, 5, 7, 0.5),
          (3100, 70, 5, 3, 0.6), (70, 60, 3, 40, 0.8),      # L = 40, the reference's largest model-selection setting
          (6, 17500, 3, 4, 0.97)]                           # rows of more than 16384 observed entries: streamed rows (kernels_long.cuh)


@pytest.mark.parametrize("shape", SHAPES)
def test_gibbs_iteration_vs_oracle(shape, monkeypatch):
    """One Gibbs iteration on the GPU; the oracle replays the same draws and must
    reproduce every conditional (tau, mu) of F, S and G that the kernels used."""
    from bnmtf_ard_b200 import bnmtf_gibbs
    from bnmtf_ard_b200._lib import F_MU, F_TAU, SIDE_U, SIDE_V
    monkeypatch.setenv("BNMTF_TRACE_PARAMS", "1")
    I, J, K, L, dens = shape
    R, M, rng = make_problem(I * 31 + J, I, J, K, L, dens)
    hyper = {"alphatau": 1.0, "betatau": 1.0, "alpha0": 1.0, "beta0": 1.0, "lambdaS": 0.1}
    m = bnmtf_gibbs(R, M, K, L, True, hyper)
    m.verbose = False
    np.random.seed(4)
    m.initialise("random", "random")
    F0, S0, G0, tau0 = m.F.copy(), m.S.copy(), m.G.copy(), m.tau
    assert (F0 >= 0).all() and (S0 >= 0).all() and (G0 >= 0).all() and tau0 > 0
    m.run(1)
    eng = m._engine()
    if 2048 <= max(I, J) <= 5000:      # the side with >= 2048 columns runs on the per-row-chain cluster sweep (kernels_v5.cuh)
        assert eng.layout_info()["v2_U" if J >= I else "v2_V"] == 5, eng.layout_info()
    lamF, lamG = m.all_lambdaFk[0], m.all_lambdaGl[0]
    F, S, G = F0.copy(), S0.copy(), G0.copy()
    tF, mF, tS, mS, tG, mG, rss = orc.gibbs_replay_iteration(R, M, F, S, G, tau0, lamF, m.lambdaS, lamG,
                                                            m.all_F[0], m.all_S[0], m.all_G[0])
    close(eng.get_factor(SIDE_U, F_TAU), tF, RTOL_STEP); mu_close(eng.get_factor(SIDE_U, F_MU), mF, tF)
    close(eng.get_S(F_TAU), tS, RTOL_STEP); mu_close(eng.get_S(F_MU), mS, tS, 1e-8)
    close(eng.get_factor(SIDE_V, F_TAU), tG, RTOL_STEP); mu_close(eng.get_factor(SIDE_V, F_MU), mG, tG)
    assert (m.all_F[0] >= 0).all() and (m.all_S[0] >= 0).all() and (m.all_G[0] >= 0).all()
    close(m.all_performances["MSE"][0], rss / M.sum(), 1e-9)
    assert (lamF > 0).all() and (lamG > 0).all() and m.all_tau[0] > 0


@pytest.mark.parametrize("shape", SHAPES[:3] + SHAPES[4:])
def test_icm_iterations_vs_oracle(shape):
    from bnmtf_ard_b200 import nmtf_icm
    I, J, K, L, dens = shape
    R, M, rng = make_problem(I * 17 + J, I, J, K, L, dens)
    hyper = {"alphatau": 1.0, "betatau": 1.0, "alpha0": 1.0, "beta0": 1.0, "lambdaS": 0.1}
    F0, S0, G0 = rng.exponential(1.0, (I, K)), rng.exponential(1.0, (K, L)), rng.exponential(1.0, (J, L))
    m = nmtf_icm(R, M, K, L, True, hyper); m.verbose = False
    m.F, m.S, m.G, m.tau, m.lambdaFk, m.lambdaGl = F0.copy(), S0.copy(), G0.copy(), 1.1, np.ones(K), np.ones(L)
    m.run(3)
    F, S, G, tau, lF, lG = F0.copy(), S0.copy(), G0.copy(), 1.1, np.ones(K), np.ones(L)
    for _ in range(3):
        tau = orc.icm_iteration(R, M, F, S, G, tau, lF, lG, hyper, True, 0.1 * np.ones((K, L)))
    close(m.F, F, RTOL_TRAJ); close(m.S, S, RTOL_TRAJ); close(m.G, G, RTOL_TRAJ); close(m.tau, tau, RTOL_TRAJ)
    close(m.lambdaFk, lF, RTOL_TRAJ); close(m.lambdaGl, lG, RTOL_TRAJ)


def test_gibbs_reproducible_and_converges():
    from bnmtf_ard_b200 import bnmtf_gibbs
    I, J, K, L = 100, 80, 4, 3
    R, M, _ = make_problem(5, I, J, K, L, 0.7)
    hyper = {"alphatau": 1.0, "betatau": 1.0, "alpha0": 1.0, "beta0": 1.0, "lambdaS": 0.1}
    runs = []
    for _ in range(2):
        np.random.seed(0)
        m = bnmtf_gibbs(R, M, K, L, True, hyper); m.verbose = False
        m.train("random", "random", 300)
        runs.append(m)
    assert np.array_equal(runs[0].all_F, runs[1].all_F) and np.array_equal(runs[0].all_S, runs[1].all_S)
    mse = runs[0].all_performances["MSE"]
    assert 0.6 < np.mean(mse[-50:]) < 1.4 and mse[-1] < mse[0]
    perf = runs[0].predict(M, 150, 2)
    assert perf["MSE"] < 1.4 and perf["Rp"] > 0.9
    assert runs[0].all_S.shape == (300, K, L) and runs[0].all_lambdaGl.shape == (300, L)
    assert runs[0].quality("loglikelihood", 150, 2) < 0


# ------------------------------------------------------------------ BNMTF VB
VB_FIELDS = orc.VB_FIELDS
VB_SCAL = orc.VB_SCAL
VB_ARD = orc.VB_ARD


def load_vb(g, p, ard):
    from bnmtf_ard_b200 import bnmtf_vb
    K, L = g[p + "exp_S0"].shape
    m = bnmtf_vb(g[p + "R"], g[p + "M"], K, L, ard, hyper_of(g, p, ard))
    m.verbose = False
    for f in VB_FIELDS + (VB_ARD if ard else []):
        setattr(m, f, g[p + f + "0"].copy())
    for f in VB_SCAL:
        setattr(m, f, float(g[p + f + "0"]))
    return m


@pytest.mark.parametrize("ard", [True, False])
def test_vb_single_step_golden(gold, ard):
    g, p = gold, "tv_%s_" % ("ard" if ard else "noard")
    m = load_vb(g, p, ard)
    close(m.exp_square_diff(), g[p + "esd0"], RTOL_STEP)
    close(m.elbo(), g[p + "elbo0"], RTOL_STEP)
    m.update_F(1); m.update_exp_F(1); m.update_S(2, 1); m.update_exp_S(2, 1); m.update_G(2); m.update_exp_G(2)
    for X in ("F", "S", "G"):
        close(getattr(m, "tau_" + X), g[p + "step_tau_" + X], RTOL_STEP)
        mu_close(getattr(m, "mu_" + X), g[p + "step_mu_" + X], g[p + "step_tau_" + X], 1e-8)
        close(getattr(m, "exp_" + X), g[p + "step_exp_" + X], 1e-7, 1e-300)
        close(getattr(m, "var_" + X), g[p + "step_var_" + X], 1e-6, 1e-300)


@pytest.mark.parametrize("ard", [True, False])
def test_vb_trajectory_golden(gold, ard):
    g, p = gold, "tv_%s_" % ("ard" if ard else "noard")
    m = load_vb(g, p, ard)
    its = int(g[p + "its"])
    m.run(its)
    ok = np.isfinite(g[p + "elbo"])
    close(np.asarray(m.all_elbo)[ok], g[p + "elbo"][ok], RTOL_TRAJ)
    assert np.array_equal(np.isfinite(m.all_elbo), ok)
    close(m.all_exp_tau, g[p + "all_exp_tau"], RTOL_TRAJ)
    close(m.all_performances["MSE"], g[p + "MSE"], RTOL_TRAJ)
    for f in VB_FIELDS + VB_SCAL + (VB_ARD if ard else []):
        close(getattr(m, f), g[p + f], 1e-5, 1e-9)
    close(m.quality("ELBO"), g[p + "elbo"][-1], RTOL_TRAJ)


# (the shapes with >= 2048 columns on one side: F / G sweep on the per-row-chain cluster kernel with the covariance term)
@pytest.mark.parametrize("shape", [(150, 220, 5, 3, 0.5), (64, 3000, 4, 6, 0.6)] + SHAPES[4:])
def test_vb_iterations_vs_oracle(shape):
    from bnmtf_ard_b200 import bnmtf_vb
    I, J, K, L, dens = shape
    R, M, rng = make_problem(I * 13 + J, I, J, K, L, dens)
    hyper = {"alphatau": 1.0, "betatau": 1.0, "alpha0": 1.0, "beta0": 1.0, "lambdaS": 0.1}
    v = bnmtf_vb(R, M, K, L, True, hyper); v.verbose = False
    np.random.seed(2)
    v.initialise("random", "random")
    st = {f: getattr(v, f).copy() for f in VB_FIELDS + VB_ARD}
    st.update({f: float(getattr(v, f)) for f in VB_SCAL})
    lamS = 0.1 * np.ones((K, L))
    close(v.elbo(), orc.vb_elbo(R, M, st, hyper, True, lamS), 1e-8)
    v.run(3)
    for _ in range(3):
        orc.vb_iteration(R, M, st, hyper, True, lamS)
    close(v.exp_F, st["exp_F"], 1e-6, 1e-10); close(v.exp_S, st["exp_S"], 1e-6, 1e-10); close(v.exp_G, st["exp_G"], 1e-6, 1e-10)
    close(v.exp_tau, st["exp_tau"], RTOL_TRAJ)
    close(v.all_elbo[-1], orc.vb_elbo(R, M, st, hyper, True, lamS), RTOL_TRAJ)


# ------------------------------------------------------------------ NMTF NP
def test_np_golden(gold):
    from bnmtf_ard_b200 import nmtf_np
    g = gold
    K, L = g["tn_S0"].shape
    m = nmtf_np(g["tn_R"], g["tn_M"], K, L)
    m.verbose = False
    m.F, m.S, m.G = g["tn_F0"].copy(), g["tn_S0"].copy(), g["tn_G0"].copy()
    m.update_F(1); m.update_S(2, 1); m.update_G(3)
    close(m.F, g["tn_step_F"], RTOL_STEP); close(m.S, g["tn_step_S"], RTOL_STEP); close(m.G, g["tn_step_G"], RTOL_STEP)
    m.F, m.S, m.G = g["tn_F0"].copy(), g["tn_S0"].copy(), g["tn_G0"].copy()
    m.run(int(g["tn_its"]))
    close(m.F, g["tn_F"], RTOL_TRAJ); close(m.S, g["tn_S"], RTOL_TRAJ); close(m.G, g["tn_G"], RTOL_TRAJ)
    close(m.compute_I_div(), g["tn_idiv"], RTOL_TRAJ); close(m.all_i_div[-1], g["tn_idiv"], RTOL_TRAJ)
    close(m.all_performances["MSE"], g["tn_MSE"], RTOL_TRAJ); close(m.all_performances["Rp"], g["tn_Rp"], RTOL_TRAJ)


@pytest.mark.parametrize("shape", [(120, 2600, 3, 4, 0.5), (6, 17500, 3, 4, 0.97)])
def test_np_iterations_vs_oracle(shape):
    from bnmtf_ard_b200 import nmtf_np
    I, J, K, L, dens = shape
    R, M, rng = make_problem(77, I, J, K, L, dens)
    R = np.abs(R) + 0.05
    F0, S0, G0 = rng.random((I, K)) + 0.1, rng.random((K, L)) + 0.1, rng.random((J, L)) + 0.1
    m = nmtf_np(R, M, K, L); m.verbose = False
    m.F, m.S, m.G = F0.copy(), S0.copy(), G0.copy()
    m.run(2)
    F, S, G = F0.copy(), S0.copy(), G0.copy()
    for _ in range(2):
        orc.np_iteration(R, M, F, S, G)
    close(m.F, F, RTOL_TRAJ); close(m.S, S, RTOL_TRAJ); close(m.G, G, RTOL_TRAJ)
    close(m.all_i_div[-1], orc.np_i_div(R, M, F, S, G), RTOL_TRAJ)


@pytest.mark.parametrize("cls_name", ["bnmtf_gibbs", "nmtf_icm", "bnmtf_vb", "nmtf_np"])
def test_graph_replay_equals_eager_loop(cls_name, monkeypatch):
    """Graph replays of the tri-factorisation run loops give bitwise the chain of the eager loop."""
    import bnmtf_ard_b200 as pkg
    rng = np.random.default_rng(12)
    I, J, K, L = 70, 2100, 3, 4        # the F side runs the cluster sweep (with the projection epilogue)
    R = np.abs(rng.exponential(1, (I, K)) @ rng.exponential(1, (K, L)) @ rng.exponential(1, (J, L)).T + rng.normal(0, 1, (I, J))) + 0.05
    M = (rng.random((I, J)) < 0.6).astype(float)
    M[:, 0] = 1; M[0, :] = 1
    hyper = {"alphatau": 1.0, "betatau": 1.0, "alpha0": 1.0, "beta0": 1.0, "lambdaS": np.ones((K, L))}
    out = []
    for no_graph in ("0", "1"):
        if no_graph == "1":
            monkeypatch.setenv("BNMTF_NO_GRAPH", "1")
        np.random.seed(9)
        cls = getattr(pkg, cls_name)
        m = cls(R, M, K, L) if cls_name == "nmtf_np" else cls(R, M, K, L, True, hyper)
        m.verbose = False
        m.initialise("random", "random")
        m.run(12)
        if cls_name in ("bnmtf_gibbs", "nmtf_icm"):
            out.append((m.all_F.copy(), m.all_S.copy(), m.all_G.copy(), m.all_tau.copy(), m.all_lambdaFk.copy(), m.all_lambdaGl.copy()))
        elif cls_name == "bnmtf_vb":
            out.append((m.exp_F.copy(), m.exp_S.copy(), m.exp_G.copy(), np.array(m.all_performances["MSE"])))
        else:
            out.append((m.F.copy(), m.S.copy(), m.G.copy(), np.array(m.all_performances["MSE"])))
    for a, b in zip(*out):
        assert np.array_equal(a, b)
