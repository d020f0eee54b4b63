"""GPU parity tests: the CUDA path (through the C ABI / the Python classes)
against (a) golden vectors produced by the real reference and (b) the numpy
oracle on seeded inputs.  Tolerances follow BASELINE.json north_star:
single-step conditionals 1e-9 relative, deterministic trajectories 1e-6."""
import os

import numpy as np
import pytest

import nmf_oracle as orc

pytestmark = pytest.mark.gpu

RTOL_STEP = 1e-9
RTOL_TRAJ = 1e-6


def hyper_of(g, p, ard, I=None, K=None):
    h = g[p + "hyper"]
    d = {"alphatau": h[0], "betatau": h[1], "alpha0": h[2], "beta0": h[3]}
    if not ard:
        d.update(lambdaU=g[p + "lamU"], lambdaV=g[p + "lamV"])
    return d


def close(a, b, rtol, atol=0.0):
    np.testing.assert_allclose(np.asarray(a, dtype=float), np.asarray(b, dtype=float), rtol=rtol, atol=atol)


def mu_close(mu, mu_ref, tau_ref, rtol=RTOL_STEP):
    """mu = (-lam + tau*b)/a can cancel: tolerance is relative to max(|mu|, posterior sd)."""
    mu, mu_ref = np.asarray(mu), np.asarray(mu_ref)
    scale = np.maximum(np.abs(mu_ref), 1.0 / np.sqrt(np.asarray(tau_ref)))
    assert np.max(np.abs(mu - mu_ref) / scale) < rtol


def make_problem(seed, I, J, K, density, nonneg=False):
    rng = np.random.default_rng(seed)
    R = rng.exponential(1.0, (I, K)) @ rng.exponential(1.0, (J, K)).T + rng.normal(0, 1, (I, J))
    if nonneg:
        R = np.abs(R) + 0.05
    M = (rng.random((I, J)) < density).astype(float)
    M[np.arange(I), rng.integers(0, J, I)] = 1
    M[rng.integers(0, I, J), np.arange(J)] = 1
    return R, M, rng


@pytest.mark.parametrize("name", ["a", "b"])
@pytest.mark.parametrize("ard", [True, False])
def test_gibbs_conditionals_golden(golden, name, ard):
    from bnmtf_ard_b200 import bnmf_gibbs, nmf_icm
    g, p = golden, "gibbs_%s_%s_" % (name, "ard" if ard else "noard")
    for cls in (bnmf_gibbs, nmf_icm):
        m = cls(g[p + "R"], g[p + "M"], g[p + "U"].shape[1], ard, hyper_of(g, p, ard))
        m.U, m.V, m.tau, m.lambdak = g[p + "U"].copy(), g[p + "V"].copy(), float(g[p + "tau"]), g[p + "lambdak"].copy()
        K = m.K
        for k in range(K):
            tU = m.tauU(k)
            close(tU, g[p + "tauU"][:, k], RTOL_STEP)
            mu_close(m.muU(tU, k), g[p + "muU"][:, k], g[p + "tauU"][:, k])
            tV = m.tauV(k)
            close(tV, g[p + "tauV"][:, k], RTOL_STEP)
            mu_close(m.muV(tV, k), g[p + "muV"][:, k], g[p + "tauV"][:, k])
        close(m.beta_s(), g[p + "beta_s"], RTOL_STEP)
        assert m.alpha_s() == float(g[p + "alpha_s"])
        if ard:
            close([m.betak_s(k) for k in range(K)], g[p + "betak_s"], 1e-13)


@pytest.mark.parametrize("ard", [True, False])
def test_icm_trajectory_golden(golden, ard):
    from bnmtf_ard_b200 import nmf_icm
    g, p = golden, "icm_%s_" % ("ard" if ard else "noard")
    m = nmf_icm(g[p + "R"], g[p + "M"], g[p + "U0"].shape[1], ard, hyper_of(g, p, ard))
    m.verbose = False
    m.U, m.V, m.tau, m.lambdak = g[p + "U0"].copy(), g[p + "V0"].copy(), float(g[p + "tau0"]), g[p + "lambdak0"].copy()
    its = int(g[p + "its"])
    m.run(its)
    close(m.all_U, g[p + "all_U"], RTOL_TRAJ); close(m.all_V, g[p + "all_V"], RTOL_TRAJ)
    close(m.all_tau, g[p + "all_tau"], RTOL_TRAJ)
    if ard:
        close(m.all_lambdak, g[p + "all_lambdak"], RTOL_TRAJ)
    close(m.all_performances["MSE"], g[p + "MSE"], RTOL_TRAJ)
    close(m.all_performances["R^2"], g[p + "R2"], RTOL_TRAJ)
    close(m.all_performances["Rp"], g[p + "Rp"], RTOL_TRAJ)
    assert len(m.all_times) == its and all(b >= a for a, b in zip(m.all_times, m.all_times[1:]))


VB_FIELDS = ["mu_U", "tau_U", "exp_U", "var_U", "mu_V", "tau_V", "exp_V", "var_V"]
VB_SCAL = ["exp_tau", "exp_logtau", "alpha_s", "beta_s"]
VB_ARD = ["alphak_s", "betak_s", "exp_lambdak", "exp_loglambdak"]


def load_vb(g, p, ard):
    from bnmtf_ard_b200 import bnmf_vb
    m = bnmf_vb(g[p + "R"], g[p + "M"], g[p + "exp_U0"].shape[1], ard, hyper_of(g, p, ard))
    m.verbose = False
    for f in VB_FIELDS + (VB_ARD if ard else []):
        setattr(m, f, g[p + f + "0"].copy())
    for f in VB_SCAL:
        setattr(m, f, float(g[p + f + "0"]))
    return m


@pytest.mark.parametrize("ard", [True, False])
def test_vb_single_step_golden(golden, ard):
    g, p = golden, "vb_%s_" % ("ard" if ard else "noard")
    m = load_vb(g, p, ard)
    close(m.exp_square_diff(), g[p + "esd0"], RTOL_STEP)
    close(m.elbo(), g[p + "elbo0"], RTOL_STEP)
    m.update_U(1); m.update_exp_U(1); m.update_V(2); m.update_exp_V(2)
    close(m.tau_U, g[p + "step_tau_U"], RTOL_STEP); close(m.tau_V, g[p + "step_tau_V"], RTOL_STEP)
    mu_close(m.mu_U, g[p + "step_mu_U"], g[p + "step_tau_U"]); mu_close(m.mu_V, g[p + "step_mu_V"], g[p + "step_tau_V"])
    close(m.exp_U, g[p + "step_exp_U"], 1e-8, 1e-300); close(m.var_U, g[p + "step_var_U"], 1e-8, 1e-300)
    close(m.exp_V, g[p + "step_exp_V"], 1e-8, 1e-300); close(m.var_V, g[p + "step_var_V"], 1e-8, 1e-300)


@pytest.mark.parametrize("ard", [True, False])
def test_vb_trajectory_golden(golden, ard):
    g, p = golden, "vb_%s_" % ("ard" if ard else "noard")
    m = load_vb(g, p, ard)
    its = int(g[p + "its"])
    m.run(its)
    close(m.all_elbo, g[p + "elbo"], RTOL_TRAJ)
    close(m.all_exp_tau, g[p + "all_exp_tau"], RTOL_TRAJ)
    close(m.all_performances["MSE"], g[p + "MSE"], RTOL_TRAJ)
    close(m.all_performances["R^2"], g[p + "R2"], RTOL_TRAJ)
    close(m.all_performances["Rp"], g[p + "Rp"], RTOL_TRAJ)
    for f in VB_FIELDS + VB_SCAL + (VB_ARD if ard else []):
        close(getattr(m, f), g[p + f], 1e-5, 1e-9)
    close(m.elbo(), g[p + "elbo"][-1], RTOL_TRAJ)
    close(m.quality("ELBO"), g[p + "elbo"][-1], RTOL_TRAJ)


def test_np_golden(golden):
    from bnmtf_ard_b200 import nmf_np
    g = golden
    m = nmf_np(g["np_R"], g["np_M"], g["np_U0"].shape[1])
    m.verbose = False
    m.U, m.V = g["np_U0"].copy(), g["np_V0"].copy()
    m.update_U(2); m.update_V(1)
    close(m.U, g["np_step_U"], RTOL_STEP); close(m.V, g["np_step_V"], RTOL_STEP)
    m.U, m.V = g["np_U0"].copy(), g["np_V0"].copy()
    m.run(int(g["np_its"]))
    close(m.U, g["np_U"], RTOL_TRAJ); close(m.V, g["np_V"], RTOL_TRAJ)
    close(m.compute_I_div(), g["np_idiv"], RTOL_TRAJ)
    close(m.all_i_div[-1], g["np_idiv"], RTOL_TRAJ)
    close(m.all_performances["MSE"], g["np_MSE"], RTOL_TRAJ)
    close(m.all_performances["Rp"], g["np_Rp"], RTOL_TRAJ)


# shapes chosen to exercise every row-team configuration: one warp per row with
# 2/4/8/16 entries per thread, multi-warp teams (named barriers), and ragged rows
# The last four also exercise the cluster-tile kernel (>= 2048 columns): tail tiles,
# tail row sets, both sides tiled, and spill placement of crowded bank classes.
SHAPES = [(37, 23, 3, 0.8), (120, 150, 5, 0.9), (90, 400, 4, 0.55), (64, 3000, 6, 0.7), (700, 140, 7, 0.8),
          (33, 9000, 3, 0.9), (100, 4096, 5, 0.5), (2300, 2500, 4, 0.3), (41, 5000, 3, 0.45), (70, 2049, 8, 0.95),
          (9, 17500, 3, 0.97),     # rows of ~17000 observed entries, streamed from global memory (kernels_long.cuh)
          (36, 5120, 4, 0.92)]     # row segments of ~590 entries (20 per lane): cluster sweep with the tensor-memory park


@pytest.mark.parametrize("shape", SHAPES)
def test_gibbs_iteration_vs_oracle(shape, monkeypatch):
    """One full Gibbs iteration on the GPU; the oracle replays the same draws and
    must reproduce every conditional (tau, mu) the kernel used, and the residual
    statistics afterwards."""
    from bnmtf_ard_b200 import bnmf_gibbs
    from bnmtf_ard_b200._lib import F_MU, F_TAU, SIDE_U, SIDE_V
    monkeypatch.setenv("BNMTF_TRACE_PARAMS", "1")
    I, J, K, dens = shape
    R, M, rng = make_problem(I * 1000 + J, I, J, K, dens)
    hyper = {"alphatau": 1.0, "betatau": 1.0, "alpha0": 1.0, "beta0": 1.0}
    m = bnmf_gibbs(R, M, K, True, hyper)
    m.verbose = False
    np.random.seed(5)
    m.initialise("random")
    U0, V0, tau0 = m.U.copy(), m.V.copy(), m.tau
    assert (U0 >= 0).all() and (V0 >= 0).all() and tau0 > 0
    m.run(1)
    eng = m._engine()
    muU, tauU = eng.get_factor(SIDE_U, F_MU), eng.get_factor(SIDE_U, F_TAU)
    muV, tauV = eng.get_factor(SIDE_V, F_MU), eng.get_factor(SIDE_V, F_TAU)
    lam = m.all_lambdak[0]
    assert (m.all_U[0] >= 0).all() and (m.all_V[0] >= 0).all()
    U, V = U0.copy(), V0.copy()
    tU, mU, tV, mV, rss = orc.gibbs_replay_iteration(R, M, U, V, tau0, lam, lam, m.all_U[0], m.all_V[0])
    close(tauU, tU, RTOL_STEP); close(tauV, tV, RTOL_STEP)
    mu_close(muU, mU, tU); mu_close(muV, mV, tV)
    perf = orc.metrics(M, R, m.all_U[0] @ m.all_V[0].T)
    close(m.all_performances["MSE"][0], perf["MSE"], 1e-9)
    close(m.all_performances["R^2"][0], perf["R^2"], 1e-9)
    close(m.all_performances["Rp"][0], perf["Rp"], 1e-9)
    # lambda_k and tau are draws: check support and that they moved
    assert (lam > 0).all() and m.all_tau[0] > 0 and m.all_tau[0] != tau0


def test_cluster_kernel_is_selected():
    from bnmtf_ard_b200._engine import NMFEngine
    R, M, _ = make_problem(3, 100, 4096, 3, 0.5)
    eng = NMFEngine(R, M, 5)
    info = eng.layout_info()
    assert info["v2_U"] in (1, 4, 5) and info["v2_V"] == 0 and info["v2_npt_U"] in (8, 12)
    eng.close()


def test_streamed_rows_are_selected_and_column_params_match():
    """Rows beyond 16384 observed entries: the row sweep streams the residual from global memory; the single-column
    conditionals (tauU / muU of the class API) go through the same kernel and must match the oracle."""
    from bnmtf_ard_b200 import bnmf_gibbs, bnmtf_gibbs
    from bnmtf_ard_b200._engine import NMFEngine
    I, J, K, dens = SHAPES[10]
    R, M, rng = make_problem(11, I, J, K, dens)
    eng = NMFEngine(R, M, K)
    info = eng.layout_info()
    assert info["tpr_U"] == 512 and info["npt_U"] == 0 and info["v2_U"] == 0 and info["npt_V"] > 0
    eng.close()
    hyper = {"alphatau": 1.0, "betatau": 1.0, "alpha0": 1.0, "beta0": 1.0}
    m = bnmf_gibbs(R, M, K, True, hyper); m.verbose = False
    m.U, m.V, m.tau, m.lambdak = rng.exponential(1.0, (I, K)), rng.exponential(1.0, (J, K)), 0.7, np.array([0.5, 1.0, 2.0])
    E = orc.residual(R, M, m.U, m.V)
    for k in range(K):
        t_ref, mu_ref = orc.col_params(E, M, m.U, m.V, m.tau, m.lambdak[k], k)
        t = m.tauU(k)
        close(t, t_ref, RTOL_STEP); mu_close(m.muU(t, k), mu_ref, t_ref)
    close(m.beta_s(), orc.beta_s(R, M, m.U, m.V, 1.0), RTOL_STEP)
    # the tri-factorisation classes stream such rows too (round 2; parity: tests/test_gpu_nmtf.py, shape 6 x 17500)
    t3 = bnmtf_gibbs(R, M, 2, 2, True, dict(hyper, lambdaS=0.1)); t3.verbose = False
    np.random.seed(3)
    t3.initialise('random', 'random')
    t3.run(2)
    assert np.isfinite(t3.all_performances["MSE"]).all() and (t3.F >= 0).all() and (t3.G >= 0).all()


@pytest.mark.parametrize("shape", SHAPES[:4] + SHAPES[6:8] + SHAPES[10:])
def test_icm_vb_np_iterations_vs_oracle(shape):
    from bnmtf_ard_b200 import bnmf_vb, nmf_icm, nmf_np
    I, J, K, dens = shape
    R, M, rng = make_problem(I * 77 + J, I, J, K, dens, nonneg=True)
    hyper = {"alphatau": 1.0, "betatau": 1.0, "alpha0": 1.0, "beta0": 1.0}
    U0, V0 = rng.exponential(1.0, (I, K)), rng.exponential(1.0, (J, K))
    # ICM
    m = nmf_icm(R, M, K, True, hyper); m.verbose = False
    m.U, m.V, m.tau, m.lambdak = U0.copy(), V0.copy(), 1.3, np.ones(K)
    m.run(3)
    U, V, tau, lam = U0.copy(), V0.copy(), 1.3, np.ones(K)
    for _ in range(3):
        tau, lam = orc.icm_iteration(R, M, U, V, tau, lam, hyper, True)
    close(m.U, U, RTOL_TRAJ); close(m.V, V, RTOL_TRAJ); close(m.tau, tau, RTOL_TRAJ); close(m.lambdak, lam, RTOL_TRAJ)
    # NP
    n = nmf_np(R, M, K); n.verbose = False
    n.U, n.V = U0.copy(), V0.copy()
    n.run(3)
    U, V = U0.copy(), V0.copy()
    for _ in range(3):
        orc.np_iteration(R, M, U, V)
    close(n.U, U, RTOL_TRAJ); close(n.V, V, RTOL_TRAJ)
    close(n.all_i_div[-1], orc.np_i_div(R, M, U, V), RTOL_TRAJ)
    # VB
    v = bnmf_vb(R, M, K, True, hyper); v.verbose = False
    np.random.seed(1)
    v.initialise("random")
    st = {f: getattr(v, f).copy() for f in VB_FIELDS + VB_ARD}
    st.update({f: float(getattr(v, f)) for f in VB_SCAL})
    close(v.elbo(), orc.vb_elbo(R, M, st, hyper, True), 1e-8)
    v.run(3)
    for _ in range(3):
        orc.vb_iteration(R, M, st, hyper, True)
    close(v.exp_U, st["exp_U"], 1e-6, 1e-10); close(v.exp_V, st["exp_V"], 1e-6, 1e-10)
    close(v.exp_tau, st["exp_tau"], RTOL_TRAJ)
    close(v.all_elbo[-1], orc.vb_elbo(R, M, st, hyper, True), RTOL_TRAJ)


def test_gibbs_reproducible_and_converges():
    """Same numpy seed => identical chain (experiments_toy/convergence/nmf_gibbs.py:52,65-66);
    the chain converges to the noise level on a synthetic low-rank matrix."""
    from bnmtf_ard_b200 import bnmf_gibbs
    I, J, K = 100, 80, 5
    R, M, _ = make_problem(9, I, J, K, 0.7)
    hyper = {"alphatau": 1.0, "betatau": 1.0, "alpha0": 1.0, "beta0": 1.0}
    runs = []
    for _ in range(2):
        np.random.seed(0)
        m = bnmf_gibbs(R, M, K, True, hyper); m.verbose = False
        m.train("random", 200)
        runs.append(m)
    assert np.array_equal(runs[0].all_U, runs[1].all_U) and np.array_equal(runs[0].all_tau, runs[1].all_tau)
    mse = runs[0].all_performances["MSE"]
    assert 0.7 < mse[-1] < 1.3 and mse[-1] < mse[0]
    Mt = 1.0 - M
    Mt[0, 0] = 1.0
    perf = runs[0].predict(M, 100, 2)
    assert perf["MSE"] < 1.3 and perf["Rp"] > 0.9
    assert runs[0].all_U.shape == (200, I, K) and runs[0].all_lambdak.shape == (200, K)


def test_reference_literal_cases():
    """The known-answer cases of the reference's own tests
    (tests/code/test_bnmf_gibbs.py:146-220)."""
    from bnmtf_ard_b200 import bnmf_gibbs
    I, J, K = 5, 3, 2
    R, M = np.ones((I, J)), np.ones((I, J))
    M[0, 0], M[2, 2], M[3, 1] = 0, 0, 0
    lambdaU, lambdaV = 2 * np.ones((I, K)), 3 * np.ones((J, K))
    hp = {'alphatau': 3, 'betatau': 1, 'alpha0': 6, 'beta0': 2, 'lambdaU': lambdaU, 'lambdaV': lambdaV}
    m = bnmf_gibbs(R, M, K, False, hp)
    m.initialise('exp')
    assert np.all(m.U == 0.5) and np.all(m.V == 1. / 3.)
    assert m.alpha_s() == 3 + 6.
    assert abs(m.beta_s() - (1 + .5 * (12 * (2. / 3.) ** 2))) < 1e-14
    m.tau = 3.
    tauU = 3. * np.array([[2. / 9., 2. / 9.], [1. / 3., 1. / 3.], [2. / 9., 2. / 9.], [2. / 9., 2. / 9.], [1. / 3., 1. / 3.]])
    muU = 1. / tauU * (3. * np.array([[2. * (5. / 6.) * (1. / 3.), 10. / 18.], [15. / 18., 15. / 18.], [10. / 18., 10. / 18.],
                                       [10. / 18., 10. / 18.], [15. / 18., 15. / 18.]]) - lambdaU)
    tauV = 3. * np.ones((J, K))
    muV = 1. / tauV * (3. * 4. * (5. / 6.) * (1. / 2.) * np.ones((J, K)) - lambdaV)
    for k in range(K):
        close(m.tauU(k), tauU[:, k], 1e-14); close(m.muU(tauU[:, k], k), muU[:, k], 1e-13)
        close(m.tauV(k), tauV[:, k], 1e-14); close(m.muV(tauV[:, k], k), muV[:, k], 1e-13)
    m2 = bnmf_gibbs(R, M, K, True, hp)
    m2.initialise('exp')
    assert all(m2.lambdak == 3.) and all(m2.alphak_s(k) == 6 + I + J for k in range(K))
    close([m2.betak_s(k) for k in range(K)], [2 + I / 3. + J / 3.] * K, 1e-14)


@pytest.mark.parametrize("shape", [(2300, 2500, 6, 0.3), (150, 120, 6, 0.8)])
def test_gibbs_kernel_draws_follow_their_conditionals(shape, monkeypatch):
    """The draws made inside the sweep kernels (cluster kernel for the first shape, row-team
    kernel for the second) are TN(mu, tau) variates of the conditionals the kernel reports:
    F(draw; mu, tau) must be uniform (KS), on a problem whose surplus components are shrunk
    (mu < 0: exponential-proposal branch) as well as on active ones (mu > 0)."""
    from scipy import stats
    from bnmtf_ard_b200 import bnmf_gibbs
    from bnmtf_ard_b200._lib import F_MU, F_TAU, SIDE_U, SIDE_V
    monkeypatch.setenv("BNMTF_TRACE_PARAMS", "1")
    I, J, K, dens = shape
    R, M, rng = make_problem(I + J, I, J, 2, dens)          # rank-2 data, K = 6 model
    hyper = {"alphatau": 1.0, "betatau": 1.0, "alpha0": 1.0, "beta0": 1.0}
    m = bnmf_gibbs(R, M, K, True, hyper)
    m.verbose = False
    np.random.seed(17)
    m.initialise("random")
    m.run(6)
    eng = m._engine()
    for side, X in ((SIDE_U, m.all_U[-1]), (SIDE_V, m.all_V[-1])):
        mu, tau = eng.get_factor(side, F_MU).ravel(), eng.get_factor(side, F_TAU).ravel()
        x = X.ravel()
        assert (x >= 0).all() and np.isfinite(x).all()
        a = -mu * np.sqrt(tau)
        for sel, min_frac in ((a <= 0, 0.05), (a > 0, 0.05), ((a > -1) & (a < 1), 0.0)):
            if x.size >= 5000:
                assert sel.mean() >= min_frac, "branch not exercised: %.3f" % sel.mean()
            if sel.sum() < 150:
                continue
            u = orc.tn_cdf_vec(x[sel], mu[sel], tau[sel])
            d, p = stats.kstest(u, "uniform")
            assert p > 1e-4, (side, sel.mean(), d, p)


@pytest.mark.parametrize("cls_name,shape", [("bnmf_gibbs", (120, 90, 4)), ("bnmf_vb", (120, 90, 4)), ("nmf_np", (60, 50, 3)),
                                            ("bnmf_gibbs", (64, 2300, 5))])
def test_graph_replay_equals_eager_loop(cls_name, shape, monkeypatch):
    """Runs of >= 8 iterations replay the iteration as a CUDA graph with the iteration index on the
    device; BNMTF_NO_GRAPH=1 forces the eager loop.  Same seed => bitwise the same chain and traces."""
    import bnmtf_ard_b200 as pkg
    I, J, K = shape
    R, M, _ = make_problem(I + 7 * J, I, J, K, 0.7, nonneg=True)
    hyper = {"alphatau": 1.0, "betatau": 1.0, "alpha0": 1.0, "beta0": 1.0}
    out = []
    for no_graph in ("0", "1"):
        if no_graph == "1":
            monkeypatch.setenv("BNMTF_NO_GRAPH", "1")
        np.random.seed(4)
        cls = getattr(pkg, cls_name)
        m = cls(R, M, K) if cls_name == "nmf_np" else cls(R, M, K, True, hyper)
        m.verbose = False
        m.initialise("random")
        m.run(25)
        if cls_name == "bnmf_gibbs":
            out.append((m.all_U.copy(), m.all_V.copy(), m.all_tau.copy(), m.all_lambdak.copy(), np.array(m.all_performances["MSE"])))
        elif cls_name == "bnmf_vb":
            out.append((m.exp_U.copy(), m.exp_V.copy(), np.array(m.all_elbo), np.array(m.all_exp_tau), np.array(m.all_performances["MSE"])))
        else:
            out.append((m.U.copy(), m.V.copy(), np.array(m.all_i_div), np.array(m.all_performances["MSE"])))
    for a, b in zip(*out):
        assert np.array_equal(a, b)
