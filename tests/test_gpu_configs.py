"""GPU parity on the configurations BASELINE.json names (round-1 verdict: every earlier GPU parity test ran K <= 8).

  C3  strips of the 20000 x 20000, rho = 0.2, K = 50 benchmark matrix in the benchmark's packed layout (20 entries per
      lane, 11 rows per set, several persistent waves of row sets): Gibbs replay against the oracle at 1e-9, ICM / VB
      trajectories at 1e-6  (bnmf_gibbs.py:148-172,203-219; nmf_icm.py:148-172; bnmf_vb.py:155-188).
  C1  the reference's toy matrix data/toy/bnmf/R.txt, K = 10, ARD: VB / ICM / NP trajectories of 500 iterations against
      goldens made by the REAL reference (oracle/gen_golden_configs.py -> tests/golden/config_c1.npz).
  C2  the GDSC IC50 matrix (705 x 139), K = 20, VB + ARD, 200 iterations: ELBO / MSE per iteration against the real
      reference (tests/golden/config_c2.npz); the reference's ELBO underflows to -inf from iteration 21 on and the
      engine must say so too.
  C4  strips of 10000 x 8000, rho = 0.3, K = L = 20: BNMTF Gibbs replay and VB iterations (bnmtf_gibbs.py:185-218,
      bnmtf_vb.py:204-245,319-375).
  C5  887 x 545, rho = 0.8, K = 30: VB / ICM / NP iterations, alone and through the batched engine.
"""
import os

import numpy as np
import pytest

import nmf_oracle as orc
import nmtf_oracle as torc

pytestmark = pytest.mark.gpu
RTOL_STEP, RTOL_TRAJ = 1e-9, 1e-6
HYPER = {"alphatau": 1.0, "betatau": 1.0, "alpha0": 1.0, "beta0": 1.0}
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

VB_FIELDS = ["mu_U", "tau_U", "exp_U", "var_U", "mu_V", "tau_V", "exp_V", "var_V"]
VB_SCAL = ["exp_tau", "exp_logtau", "alpha_s", "beta_s"]
VB_ARD = ["alphak_s", "betak_s", "exp_lambdak", "exp_loglambdak"]


def close(a, b, rtol, atol=0.0):
    np.testing.assert_allclose(np.asarray(a, dtype=float), np.asarray(b, dtype=float), rtol=rtol, atol=atol)


def mu_close(mu, mu_ref, tau_ref, rtol=RTOL_STEP):
    """mu = (-lam + tau*b)/a can cancel: tolerance relative to max(|mu|, posterior sd)."""
    mu, mu_ref = np.asarray(mu, dtype=float), np.asarray(mu_ref, dtype=float)
    scale = np.maximum(np.abs(mu_ref), 1.0 / np.sqrt(np.asarray(tau_ref, dtype=float)))
    err = float(np.max(np.abs(mu - mu_ref) / scale))
    assert err < rtol, err


def synth(I, J, K, density, seed=20250, nonneg=False):
    """bench.py's recipe: R = U* V*^T + N(0,1), U*, V* ~ Exp(1), M ~ Bernoulli(rho), no empty row / column."""
    rng = np.random.default_rng(seed)
    R = rng.exponential(1.0, (I, K)) @ rng.exponential(1.0, (J, K)).T + rng.standard_normal((I, J))
    if nonneg:
        R = np.abs(R) + 0.05
    M = (rng.random((I, J)) < density)
    M[np.arange(I), rng.integers(0, J, I)] = True
    M[rng.integers(0, I, J), np.arange(J)] = True
    return R, M.astype(np.float64), rng


# ------------------------------------------------------------------------------------------------ C3
C3_STRIPS = [(600, 20000), (20000, 600)]
# the 20000 x 20000 matrix has a longest (row, tile) segment of ~590 entries -> 20 entries per lane; a strip's own
# maximum is a little shorter, so the packing floor of the full matrix is handed over (what a sharded run does too)
C3_FLOOR = [0, 600, 0, 600]


def c3_model(cls, R, M, K=50):
    m = cls(R, M, K, True, dict(HYPER))
    m.verbose = False
    m._layout_floor = C3_FLOOR
    return m


def assert_bench_layout(m, shape):
    info = m._engine().layout_info()
    side = "U" if shape[1] >= shape[0] else "V"
    assert info["v2_" + side] >= 1, info
    assert info["v2_npt_" + side] == 20 and info["v2_rows_per_set_" + side] == 11, info
    return info


@pytest.mark.parametrize("shape", C3_STRIPS)
def test_c3_strip_gibbs_replay(shape, monkeypatch):
    from bnmtf_ard_b200 import bnmf_gibbs
    from bnmtf_ard_b200._lib import F_MU, F_TAU, SIDE_U, SIDE_V
    monkeypatch.setenv("BNMTF_TRACE_PARAMS", "1")
    I, J = shape
    K = 50
    R, M, _ = synth(I, J, K, 0.2)
    m = c3_model(bnmf_gibbs, R, M)
    np.random.seed(5)
    m.initialise("random")
    info = assert_bench_layout(m, shape)
    nsets = -(-max(I, J) // 11) if min(I, J) == 600 else 0
    assert -(-600 // 11) > 33 or nsets > 33      # more row sets than resident clusters: the persistent loop wraps
    U0, V0, tau0 = m.U.copy(), m.V.copy(), m.tau
    m.run(2)                                      # two iterations: the second starts from factors the kernels wrote
    eng = m._engine()
    lam = m.all_lambdak[1]
    U, V = m.all_U[0].copy(), m.all_V[0].copy()
    tU, mU, tV, mV, rss = orc.gibbs_replay_iteration(R, M, U, V, m.all_tau[0], lam, lam, m.all_U[1], m.all_V[1])
    close(eng.get_factor(SIDE_U, F_TAU), tU, RTOL_STEP); close(eng.get_factor(SIDE_V, F_TAU), tV, RTOL_STEP)
    mu_close(eng.get_factor(SIDE_U, F_MU), mU, tU); mu_close(eng.get_factor(SIDE_V, F_MU), mV, tV)
    close(m.all_performances["MSE"][1], rss / M.sum(), 1e-9)
    perf = orc.metrics(M, R, m.all_U[1] @ m.all_V[1].T)
    close(m.all_performances["R^2"][1], perf["R^2"], 1e-9); close(m.all_performances["Rp"][1], perf["Rp"], 1e-9)
    assert (m.all_U >= 0).all() and (m.all_V >= 0).all() and (lam > 0).all() and m.all_tau[1] > 0
    assert info["nnz_U"] == int(M.sum())


@pytest.mark.parametrize("shape", C3_STRIPS[:1])
def test_c3_strip_icm_vb_trajectories(shape):
    from bnmtf_ard_b200 import bnmf_vb, nmf_icm
    I, J = shape
    K = 50
    R, M, rng = synth(I, J, K, 0.2, nonneg=True)
    U0, V0 = rng.exponential(1.0, (I, K)), rng.exponential(1.0, (J, K))
    m = c3_model(nmf_icm, R, M)
    m.U, m.V, m.tau, m.lambdak = U0.copy(), V0.copy(), 1.3, np.ones(K)
    m.run(3)
    assert_bench_layout(m, shape)
    U, V, tau, lam = U0.copy(), V0.copy(), 1.3, np.ones(K)
    for _ in range(3):
        tau, lam = orc.icm_iteration(R, M, U, V, tau, lam, HYPER, True)
    close(m.U, U, RTOL_TRAJ); close(m.V, V, RTOL_TRAJ); close(m.tau, tau, RTOL_TRAJ); close(m.lambdak, lam, RTOL_TRAJ)
    close(m.all_performances["MSE"][-1], orc.metrics(M, R, U @ V.T)["MSE"], RTOL_TRAJ)
    v = c3_model(bnmf_vb, R, M)
    np.random.seed(1)
    v.initialise("random")
    st = {f: getattr(v, f).copy() for f in VB_FIELDS + VB_ARD}
    st.update({f: float(getattr(v, f)) for f in VB_SCAL})
    v.run(3)
    for _ in range(3):
        orc.vb_iteration(R, M, st, HYPER, True)
    close(v.exp_U, st["exp_U"], RTOL_TRAJ, 1e-10); close(v.exp_V, st["exp_V"], RTOL_TRAJ, 1e-10)
    close(v.exp_tau, st["exp_tau"], RTOL_TRAJ)
    ref_elbo = orc.vb_elbo(R, M, st, HYPER, True)
    if np.isfinite(ref_elbo):
        close(v.all_elbo[-1], ref_elbo, RTOL_TRAJ)
    else:
        assert v.all_elbo[-1] == ref_elbo
    close(v.all_performances["MSE"][-1], orc.metrics(M, R, st["exp_U"] @ st["exp_V"].T)["MSE"], RTOL_TRAJ)


# ------------------------------------------------------------------------------------------------ C1
@pytest.fixture(scope="module")
def c1():
    return np.load(os.path.join(GOLD, "config_c1.npz"))


@pytest.fixture(scope="module")
def c2():
    return np.load(os.path.join(GOLD, "config_c2.npz"))


def run_vb_golden(g):
    """bnmf_vb from the golden's initial state for its iteration count; checks the ELBO / MSE trajectory and the end state."""
    from bnmtf_ard_b200 import bnmf_vb
    K, p = int(g["K"]), "vb_"
    m = bnmf_vb(g["R"], g["M"], K, True, dict(HYPER))
    m.verbose = False
    for f in VB_FIELDS + VB_ARD:
        setattr(m, f, g[p + f + "0"].copy())
    for f in VB_SCAL:
        setattr(m, f, float(g[p + f + "0"]))
    its = int(g[p + "its"])
    m.run(its)
    ref_elbo, elbo = g[p + "elbo"], np.asarray(m.all_elbo)
    ok = np.isfinite(ref_elbo)
    close(elbo[ok], ref_elbo[ok], RTOL_TRAJ)
    assert np.array_equal(elbo[~ok], ref_elbo[~ok]), "the reference's ELBO is -inf where log(0.5 erfc(.)) underflows (bnmf_vb.py:222,228)"
    close(m.all_performances["MSE"], g[p + "MSE"], RTOL_TRAJ)
    close(m.all_performances["R^2"], g[p + "R2"], RTOL_TRAJ)
    close(m.all_performances["Rp"], g[p + "Rp"], RTOL_TRAJ)
    close(m.all_exp_tau, g[p + "all_exp_tau"], RTOL_TRAJ)
    for f in ["exp_U", "exp_V", "exp_lambdak", "exp_tau", "beta_s"]:
        close(getattr(m, f), g[p + f], 1e-5, 1e-8)
    return m, int(ok.sum())


def test_c1_toy_vb_500_iterations_golden(c1):
    m, finite = run_vb_golden(c1)
    assert finite == 500 and abs(m.all_performances["MSE"][-1] - 0.8025) < 1e-3   # the reference's published toy trace ends at 0.804


def test_c1_toy_icm_500_iterations_golden(c1):
    from bnmtf_ard_b200 import nmf_icm
    g, K = c1, int(c1["K"])
    m = nmf_icm(g["R"], g["M"], K, True, dict(HYPER))
    m.verbose = False
    m.U, m.V, m.tau, m.lambdak = g["icm_U0"].copy(), g["icm_V0"].copy(), float(g["icm_tau0"]), g["icm_lambdak0"].copy()
    m.run(int(g["icm_its"]))
    close(m.all_performances["MSE"], g["icm_MSE"], RTOL_TRAJ)
    close(m.all_performances["R^2"], g["icm_R2"], RTOL_TRAJ)
    close(m.all_performances["Rp"], g["icm_Rp"], RTOL_TRAJ)
    close(m.all_tau, g["icm_all_tau"], RTOL_TRAJ)
    close(m.all_lambdak, g["icm_all_lambdak"], 1e-5)
    close(m.U, g["icm_U"], 1e-5, 1e-8); close(m.V, g["icm_V"], 1e-5, 1e-8)


def test_c1_toy_np_500_iterations_golden(c1):
    from bnmtf_ard_b200 import nmf_np
    g, K = c1, int(c1["K"])
    m = nmf_np(g["np_R"], g["M"], K)
    m.verbose = False
    m.U, m.V = g["np_U0"].copy(), g["np_V0"].copy()
    m.run(int(g["np_its"]))
    close(m.all_performances["MSE"], g["np_MSE"], RTOL_TRAJ)
    close(m.all_performances["Rp"], g["np_Rp"], RTOL_TRAJ)
    close(m.compute_I_div(), g["np_idiv"], RTOL_TRAJ)
    close(m.U, g["np_U"], 1e-5, 1e-8); close(m.V, g["np_V"], 1e-5, 1e-8)


# ------------------------------------------------------------------------------------------------ C2
def test_c2_gdsc_vb_200_iterations_golden(c2):
    assert c2["R"].shape == (705, 139) and abs(c2["M"].mean() - 0.8088) < 1e-3      # load_data.py:1-13,47-52 (705 of 707 cell lines kept)
    m, finite = run_vb_golden(c2)
    assert 0 < finite < 200          # the reference itself reports ELBO = -inf on this matrix after ~20 iterations


# ------------------------------------------------------------------------------------------------ C4
def synth3(I, J, K, L, density, seed=20250):
    rng = np.random.default_rng(seed)
    R = rng.exponential(1.0, (I, K)) @ rng.exponential(1.0, (K, L)) @ rng.exponential(1.0, (J, L)).T + rng.standard_normal((I, J))
    M = (rng.random((I, J)) < density)
    M[np.arange(I), rng.integers(0, J, I)] = True
    M[rng.integers(0, I, J), np.arange(J)] = True
    return R, M.astype(np.float64), rng


C4_STRIPS = [(400, 8000), (10000, 320)]
HYPER3 = dict(HYPER, lambdaS=0.1)


@pytest.mark.parametrize("shape", C4_STRIPS)
def test_c4_strip_bnmtf_gibbs_replay(shape, monkeypatch):
    from bnmtf_ard_b200 import bnmtf_gibbs
    from bnmtf_ard_b200._lib import F_MU, F_TAU, SIDE_U, SIDE_V
    monkeypatch.setenv("BNMTF_TRACE_PARAMS", "1")
    I, J = shape
    K = L = 20
    R, M, _ = synth3(I, J, K, L, 0.3)
    m = bnmtf_gibbs(R, M, K, L, True, dict(HYPER3))
    m.verbose = False
    np.random.seed(4)
    m.initialise("random", "random")
    m.run(2)
    eng = m._engine()
    assert eng.layout_info()["v2_U" if J >= I else "v2_V"] == 5, eng.layout_info()   # the per-row-chain cluster sweep
    lamF, lamG = m.all_lambdaFk[1], m.all_lambdaGl[1]
    F, S, G = m.all_F[0].copy(), m.all_S[0].copy(), m.all_G[0].copy()
    tF, mF, tS, mS, tG, mG, rss = torc.gibbs_replay_iteration(R, M, F, S, G, m.all_tau[0], lamF, m.lambdaS, lamG,
                                                             m.all_F[1], m.all_S[1], m.all_G[1])
    close(eng.get_factor(SIDE_U, F_TAU), tF, RTOL_STEP); mu_close(eng.get_factor(SIDE_U, F_MU), mF, tF)
    close(eng.get_S(F_TAU), tS, RTOL_STEP); mu_close(eng.get_S(F_MU), mS, tS, 1e-8)
    close(eng.get_factor(SIDE_V, F_TAU), tG, RTOL_STEP); mu_close(eng.get_factor(SIDE_V, F_MU), mG, tG)
    close(m.all_performances["MSE"][1], rss / M.sum(), 1e-9)


@pytest.mark.parametrize("shape", C4_STRIPS)
def test_c4_strip_bnmtf_vb_iterations(shape):
    from bnmtf_ard_b200 import bnmtf_vb
    I, J = shape
    K = L = 20
    R, M, _ = synth3(I, J, K, L, 0.3)
    v = bnmtf_vb(R, M, K, L, True, dict(HYPER3))
    v.verbose = False
    np.random.seed(2)
    v.initialise("random", "random")
    st = {f: getattr(v, f).copy() for f in torc.VB_FIELDS + torc.VB_ARD}
    st.update({f: float(getattr(v, f)) for f in torc.VB_SCAL})
    lamS = 0.1 * np.ones((K, L))
    e0, e0_ref = v.elbo(), torc.vb_elbo(R, M, st, HYPER3, True, lamS)
    if np.isfinite(e0_ref):
        close(e0, e0_ref, 1e-8)
    else:
        assert e0 == e0_ref
    v.run(2)
    for _ in range(2):
        torc.vb_iteration(R, M, st, HYPER3, True, lamS)
    close(v.exp_F, st["exp_F"], RTOL_TRAJ, 1e-10); close(v.exp_S, st["exp_S"], RTOL_TRAJ, 1e-10); close(v.exp_G, st["exp_G"], RTOL_TRAJ, 1e-10)
    close(v.exp_tau, st["exp_tau"], RTOL_TRAJ)
    close(v.all_performances["MSE"][-1], orc.metrics(M, R, st["exp_F"] @ st["exp_S"] @ st["exp_G"].T)["MSE"], RTOL_TRAJ)
    # ELBO at benchmark scale: the reference formula underflows (log of erfc == 0) -- the oracle and the engine must agree on that
    e_ref = torc.vb_elbo(R, M, st, HYPER3, True, lamS)
    if np.isfinite(e_ref):
        close(v.all_elbo[-1], e_ref, RTOL_TRAJ)
    else:
        assert v.all_elbo[-1] == e_ref, (v.all_elbo[-1], e_ref)


# ------------------------------------------------------------------------------------------------ C5
def c5_problem():
    return synth(887, 545, 30, 0.8, seed=887545, nonneg=True)


def test_c5_k30_vb_icm_np_iterations():
    from bnmtf_ard_b200 import bnmf_vb, nmf_icm, nmf_np
    R, M, rng = c5_problem()
    I, J, K = 887, 545, 30
    U0, V0 = rng.exponential(1.0, (I, K)), rng.exponential(1.0, (J, K))
    m = nmf_icm(R, M, K, True, dict(HYPER)); m.verbose = False
    m.U, m.V, m.tau, m.lambdak = U0.copy(), V0.copy(), 1.3, np.ones(K)
    m.run(5)
    U, V, tau, lam = U0.copy(), V0.copy(), 1.3, np.ones(K)
    for _ in range(5):
        tau, lam = orc.icm_iteration(R, M, U, V, tau, lam, HYPER, True)
    close(m.U, U, RTOL_TRAJ); close(m.V, V, RTOL_TRAJ); close(m.tau, tau, RTOL_TRAJ); close(m.lambdak, lam, RTOL_TRAJ)
    n = nmf_np(R, M, K); n.verbose = False
    n.U, n.V = U0.copy(), V0.copy()
    n.run(5)
    U, V = U0.copy(), V0.copy()
    for _ in range(5):
        orc.np_iteration(R, M, U, V)
    close(n.U, U, RTOL_TRAJ); close(n.V, V, RTOL_TRAJ)
    close(n.all_i_div[-1], orc.np_i_div(R, M, U, V), RTOL_TRAJ)
    v = bnmf_vb(R, M, K, True, dict(HYPER)); v.verbose = False
    np.random.seed(1)
    v.initialise("random")
    st = {f: getattr(v, f).copy() for f in VB_FIELDS + VB_ARD}
    st.update({f: float(getattr(v, f)) for f in VB_SCAL})
    v.run(5)
    for _ in range(5):
        orc.vb_iteration(R, M, st, HYPER, True)
    close(v.exp_U, st["exp_U"], RTOL_TRAJ, 1e-10); close(v.exp_V, st["exp_V"], RTOL_TRAJ, 1e-10)
    close(v.exp_tau, st["exp_tau"], RTOL_TRAJ)
    e_ref = orc.vb_elbo(R, M, st, HYPER, True)
    if np.isfinite(e_ref):
        close(v.all_elbo[-1], e_ref, RTOL_TRAJ)
    else:
        assert v.all_elbo[-1] == e_ref


def test_c5_k30_batched_equals_single_and_oracle():
    """A K = 30 model of the CTRP-shaped sweep driven through the batched engine (one launch per update for all models)
    ends bit for bit where the same model ends alone, and both match the oracle."""
    from bnmtf_ard_b200 import _batch, bnmf_vb
    R, M, _ = c5_problem()
    Ks = [30, 17, 30]
    floor = _batch.layout_floor([M])
    singles, models = [], []
    for batched in (False, True):
        for K in Ks:
            v = bnmf_vb(R, M, K, True, dict(HYPER)); v.verbose = False
            v._layout_floor = floor
            np.random.seed(K)
            if batched:
                with _batch.deferred(v):
                    v.train(iterations=4, init_UV="random")
                models.append(v)
            else:
                v.train(iterations=4, init_UV="random")
                singles.append(v)
    _batch.run_deferred(models)
    for a, b in zip(models, singles):
        assert np.array_equal(a.exp_U, b.exp_U) and np.array_equal(a.exp_V, b.exp_V) and a.exp_tau == b.exp_tau
        assert np.array_equal(np.asarray(a.all_elbo), np.asarray(b.all_elbo))
    v = singles[0]                       # and the K = 30 model against the oracle
    ref = bnmf_vb(R, M, 30, True, dict(HYPER)); ref.verbose = False
    ref._layout_floor = floor
    np.random.seed(30)
    ref.initialise("random")
    st = {f: getattr(ref, f).copy() for f in VB_FIELDS + VB_ARD}
    st.update({f: float(getattr(ref, f)) for f in VB_SCAL})
    for _ in range(4):
        orc.vb_iteration(R, M, st, HYPER, True)
    close(v.exp_U, st["exp_U"], RTOL_TRAJ, 1e-10); close(v.exp_V, st["exp_V"], RTOL_TRAJ, 1e-10)
