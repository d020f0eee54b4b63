"""GPU tests of the on-device posterior summaries: a chain that averages its
draws on the device (summarise_on_device, no stored traces) must give the approx_expectation /
predict / quality of the same chain with stored traces -- evaluated here with the oracle's numpy
restatement of the reference formulas (bnmf_gibbs.py:222-302, bnmtf_gibbs.py:302-380)."""
import math

import numpy as np
import pytest

import nmf_oracle as orc

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def problem(seed, I, J, K, density=0.7):
    rng = np.random.default_rng(seed)
    R = rng.exponential(1.0, (I, K)) @ rng.exponential(1.0, (J, K)).T + rng.normal(0, 0.5, (I, J))
    M = (rng.random((I, J)) < density).astype(float)
    M[np.arange(I), rng.integers(0, J, I)] = 1
    M[rng.integers(0, I, J), np.arange(J)] = 1
    M_test = ((rng.random((I, J)) < 0.3) * (1 - M)).astype(float)
    return R, M, M_test


HYP = {"alphatau": 1.0, "betatau": 1.0, "alpha0": 1.0, "beta0": 1.0}


def host_mean(traces, burn_in, thinning):
    idx = range(burn_in, len(traces), thinning)
    return np.array([traces[i] for i in idx]).sum(axis=0) / float(len(idx))


def loglik(n, exp_tau, ss_res):
    return n / 2.0 * (math.log(exp_tau) - math.log(2 * math.pi)) - exp_tau / 2.0 * ss_res


@pytest.mark.parametrize("cls_name", ["bnmf_gibbs", "nmf_icm"])
@pytest.mark.parametrize("burn_in,thinning", [(10, 3), (0, 1), (29, 5)])
def test_nmf_summary_equals_traces(cls_name, burn_in, thinning):
    import bnmtf_ard_b200 as pkg
    cls = getattr(pkg, cls_name)
    R, M, M_test = problem(5, 70, 50, 4)
    its = 30
    runs = []
    for on_device in (False, True):
        np.random.seed(11)
        m = cls(R, M, 4, True, HYP)
        m.verbose = False
        m.initialise("random")
        if on_device:
            m.summarise_on_device(burn_in, thinning)
        m.run(its)
        runs.append(m)
    a, b = runs
    assert b.all_U is None and b.all_V is None
    np.testing.assert_array_equal(a.U, b.U)
    np.testing.assert_array_equal(a.all_tau, b.all_tau)
    eU, eV, etau, elam = b.approx_expectation(burn_in, thinning)
    # same summation order as numpy's axis-0 sum: bit-identical
    np.testing.assert_array_equal(eU, host_mean(a.all_U, burn_in, thinning))
    np.testing.assert_array_equal(eV, host_mean(a.all_V, burn_in, thinning))
    np.testing.assert_allclose(elam, host_mean(a.all_lambdak, burn_in, thinning), rtol=1e-14)
    np.testing.assert_allclose(etau, host_mean(a.all_tau, burn_in, thinning), rtol=1e-14)
    ref = orc.metrics(M_test, R, eU @ eV.T)
    got = b.predict(M_test, burn_in, thinning)
    for key in ("MSE", "R^2", "Rp"):
        np.testing.assert_allclose(got[key], ref[key], rtol=RTOL)
        np.testing.assert_allclose(got[key], a.predict(M_test, burn_in, thinning)[key], rtol=RTOL)
    ss_res = float((M * (R - eU @ eV.T) ** 2).sum())
    np.testing.assert_allclose(b.quality("loglikelihood", burn_in, thinning), loglik(M.sum(), etau, ss_res), rtol=RTOL)
    np.testing.assert_allclose(b.quality("MSE", burn_in, thinning), ss_res / M.sum(), rtol=RTOL)
    for metric in ("BIC", "AIC", "loglikelihood", "MSE"):
        np.testing.assert_allclose(b.quality(metric, burn_in, thinning), a.quality(metric, burn_in, thinning), rtol=RTOL)
    with pytest.raises(ValueError):
        b.approx_expectation(burn_in + 1, thinning)


def test_nmf_summary_keep_traces_and_rerun():
    from bnmtf_ard_b200 import bnmf_gibbs
    R, M, M_test = problem(6, 40, 30, 3)
    np.random.seed(3)
    m = bnmf_gibbs(R, M, 3, False, dict(HYP, lambdaU=np.ones((40, 3)), lambdaV=np.ones((30, 3))))
    m.verbose = False
    m.initialise("exp")
    m.summarise_on_device(4, 2, keep_traces=True)
    m.run(12)
    np.testing.assert_array_equal(m.approx_expectation(4, 2)[0], host_mean(m.all_U, 4, 2))
    assert m.approx_expectation(4, 2)[3] is None
    # other (burn_in, thinning) fall back to the stored traces
    np.testing.assert_array_equal(m.approx_expectation(2, 3)[1], host_mean(m.all_V, 2, 3))
    # a second run restarts the sums (all_U is rebuilt by every run() in the reference)
    m.run(8)
    np.testing.assert_array_equal(m.approx_expectation(4, 2)[0], host_mean(m.all_U, 4, 2))
    m.summarise_on_device(0, None)
    m.run(5)
    assert m.all_U.shape[0] == 5


@pytest.mark.parametrize("cls_name", ["bnmtf_gibbs", "nmtf_icm"])
def test_nmtf_summary_equals_traces(cls_name):
    import bnmtf_ard_b200 as pkg
    cls = getattr(pkg, cls_name)
    R, M, M_test = problem(8, 60, 45, 3)
    K, L, its, burn_in, thinning = 3, 4, 24, 6, 2
    runs = []
    for on_device in (False, True):
        np.random.seed(21)
        m = cls(R, M, K, L, True, dict(HYP, lambdaS=np.ones((K, L))))
        m.verbose = False
        m.initialise("random", "random")
        if on_device:
            m.summarise_on_device(burn_in, thinning)
        m.run(its)
        runs.append(m)
    a, b = runs
    assert b.all_F is None and b.all_S is None and b.all_G is None
    np.testing.assert_array_equal(a.F, b.F)
    np.testing.assert_array_equal(a.S, b.S)
    np.testing.assert_array_equal(a.G, b.G)
    eF, eS, eG, etau, elF, elG = b.approx_expectation(burn_in, thinning)
    np.testing.assert_array_equal(eF, host_mean(a.all_F, burn_in, thinning))
    np.testing.assert_array_equal(eS, host_mean(a.all_S, burn_in, thinning))
    np.testing.assert_array_equal(eG, host_mean(a.all_G, burn_in, thinning))
    np.testing.assert_allclose(elF, host_mean(a.all_lambdaFk, burn_in, thinning), rtol=1e-14)
    np.testing.assert_allclose(elG, host_mean(a.all_lambdaGl, burn_in, thinning), rtol=1e-14)
    np.testing.assert_allclose(etau, host_mean(a.all_tau, burn_in, thinning), rtol=1e-14)
    ref = orc.metrics(M_test, R, eF @ eS @ eG.T)
    got = b.predict(M_test, burn_in, thinning)
    for key in ("MSE", "R^2", "Rp"):
        np.testing.assert_allclose(got[key], ref[key], rtol=RTOL)
    for metric in ("BIC", "AIC", "loglikelihood", "MSE"):
        np.testing.assert_allclose(b.quality(metric, burn_in, thinning), a.quality(metric, burn_in, thinning), rtol=RTOL)


def test_point_estimate_predict_on_device():
    """predict(M_pred) of the VB / NP classes runs on the device from the current factors."""
    from bnmtf_ard_b200 import bnmf_vb, nmf_np, bnmtf_vb, nmtf_np
    R, M, M_test = problem(9, 50, 40, 3)
    R = np.abs(R) + 0.05
    np.random.seed(2)
    m = bnmf_vb(R, M, 3, True, HYP); m.verbose = False
    m.initialise("random"); m.run(15)
    ref = orc.metrics(M_test, R, m.exp_U @ m.exp_V.T)
    got = m.predict(M_test)
    for key in ref:
        np.testing.assert_allclose(got[key], ref[key], rtol=RTOL)
    m = nmf_np(R, M, 3); m.verbose = False
    m.initialise("random"); m.run(15)
    ref, got = orc.metrics(M_test, R, m.U @ m.V.T), m.predict(M_test)
    for key in ref:
        np.testing.assert_allclose(got[key], ref[key], rtol=RTOL)
    m = bnmtf_vb(R, M, 3, 2, True, dict(HYP, lambdaS=np.ones((3, 2)))); m.verbose = False
    m.initialise("random", "random"); m.run(10)
    ref, got = orc.metrics(M_test, R, m.exp_F @ m.exp_S @ m.exp_G.T), m.predict(M_test)
    for key in ref:
        np.testing.assert_allclose(got[key], ref[key], rtol=RTOL)
    m = nmtf_np(R, M, 3, 2); m.verbose = False
    m.initialise("random", "random"); m.run(10)
    ref, got = orc.metrics(M_test, R, m.F @ m.S @ m.G.T), m.predict(M_test)
    for key in ref:
        np.testing.assert_allclose(got[key], ref[key], rtol=RTOL)


def test_cv_driver_uses_device_summary(tmp_path):
    from bnmtf_ard_b200 import bnmf_gibbs
    from bnmtf_ard_b200.cross_validation.matrix_cross_validation import MatrixCrossValidation
    import random
    R, M, _ = problem(10, 40, 30, 3, density=0.9)
    seen = []

    class spy(bnmf_gibbs):
        def train(self, **kw):
            self.verbose = False
            bnmf_gibbs.train(self, **kw)
            seen.append((self._summary, self.all_U is None))

    random.seed(1); np.random.seed(1)
    cv = MatrixCrossValidation(method=spy, R=R, M=M, K=3,
                               parameter_search=[{"K": 2, "ARD": True, "hyperparameters": HYP}],
                               train_config={"iterations": 20, "init_UV": "random"},
                               predict_config={"burn_in": 5, "thinning": 2},
                               file_performance=str(tmp_path / "log.txt"))
    cv.verbose = False
    cv.run()
    assert len(seen) == 3 and all(s == ((5, 2), True) for s in seen)
    (perf,) = cv.average_performances.values()
    assert 0 < perf["MSE"] < 10
