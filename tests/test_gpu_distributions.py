"""GPU: the device samplers / moments that replace code/models/distributions.
The reference's streams (numpy MT19937 + Chopin tables) cannot be matched, so
draws are validated distributionally (KS against the analytic CDF) across the
regimes of the reference sampler (rtnorm.py:96-217: a < -2.004, table, a > 3.487)
and far beyond; moments and guards are checked exactly."""
import numpy as np
import pytest
from scipy import stats

import nmf_oracle as orc

pytestmark = pytest.mark.gpu

N = 40000
P_MIN = 1e-4   # 30 KS tests in this file: a correct sampler fails with probability ~3e-3


@pytest.mark.parametrize("a", [-10.0, -3.0, -2.0, -0.5, 0.0, 0.7, 2.0, 3.4, 3.6, 4.0, 4.1, 6.0, 12.0, 40.0])
@pytest.mark.parametrize("tau", [0.25, 9.0])
def test_tn_draw_ks(a, tau):
    from bnmtf_ard_b200.distributions._device import tn_draw
    sigma = 1.0 / np.sqrt(tau)
    mu = -a * sigma                      # standardised lower bound a = -mu sqrt(tau)
    x = tn_draw(np.full(N, mu), np.full(N, tau), seed=int(1000 * abs(a) + tau * 7 + 3))
    assert (x >= 0).all() and np.isfinite(x).all()
    d, p = stats.kstest(x, lambda v: orc.tn_cdf(v, mu, tau))
    assert p > P_MIN, "TN KS failed for a=%s tau=%s: D=%.4f p=%.2e" % (a, tau, d, p)


def test_tn_draw_guards_and_api():
    from bnmtf_ard_b200.distributions import TN_draw, TN_vector_draw, TN_mode, TN_vector_mode
    # tau == 0 -> 0.0  (tests/code/distributions/test_truncated_normal.py:48-52)
    assert TN_draw(0.32, 0.0) == 0.0
    out = TN_vector_draw([0.32, 1.0, -5.0], [0.0, 3.0, 2.0])
    assert out[0] == 0.0 and out[1] >= 0.0 and out[2] >= 0.0 and len(out) == 3
    for _ in range(20):
        assert TN_draw(1.0, 3.0) >= 0.0
    assert TN_mode(-2.0) == 0.0 and TN_mode(3.0) == 3.0
    assert list(TN_vector_mode(np.array([-1.0, 2.0]))) == [0.0, 2.0]
    # numpy.random.seed makes the device stream reproducible
    np.random.seed(3); a = TN_vector_draw([1.0] * 5, [2.0] * 5)
    np.random.seed(3); b = TN_vector_draw([1.0] * 5, [2.0] * 5)
    assert a == b


def test_tn_moments_golden(golden):
    from bnmtf_ard_b200.distributions import TN_expectation, TN_variance, TN_vector_expectation, TN_vector_variance
    g = golden
    mu, tau = g["dist_mu"], g["dist_tau"]
    e, v = np.array(TN_vector_expectation(mu, tau)), np.array(TN_vector_variance(mu, tau))
    np.testing.assert_allclose(e, g["dist_tn_exp_vec"], rtol=1e-9, atol=1e-300)
    # var = sigma^2 (1 - lambda(lambda - x)) cancels for large x (erfc/exp differ from scipy's
    # by an ulp, amplified by x^2): absolute slack on the scale sigma^2 next to the 1e-9 relative bar
    scale = 1.0 / tau
    assert np.all(np.abs(v - g["dist_tn_var_vec"]) <= 1e-9 * np.abs(g["dist_tn_var_vec"]) + 2e-11 * scale)
    # reference literals (test_truncated_normal.py:12-38)
    sigma = 0.5773502691896258
    lam = stats.norm.pdf(-1.0 / sigma) / (1 - stats.norm.cdf(-1.0 / sigma))
    assert TN_expectation(1.0, 3.0) == pytest.approx(1.0 + sigma * lam, rel=1e-12)
    assert TN_variance(1.0, 3.0) == pytest.approx(sigma ** 2 * (1 - lam * (lam + 1.0 / sigma)), rel=1e-11)
    assert TN_expectation(-1.0, 2000.0) == pytest.approx(1.0 / 2000.0, rel=1e-15)
    assert TN_variance(-1.0, 2000.0) == pytest.approx((1.0 / 2000.0) ** 2, rel=1e-15)
    # non-finite inputs are guarded to 0 like the reference
    assert TN_expectation(float("nan"), 1.0) == 0.0 and TN_variance(1.0, 0.0) == 0.0


@pytest.mark.parametrize("alpha,beta", [(0.3, 2.0), (1.0, 1.0), (2.0, 3.0), (51.0, 7.5), (40001.0, 1234.5), (4.0e7, 3.9e7)])
def test_gamma_draw_ks(alpha, beta):
    from bnmtf_ard_b200.distributions._device import gamma_draws
    x = gamma_draws(np.full(N, alpha), np.full(N, beta), seed=int(alpha * 13 + 1))
    assert (x > 0).all()
    d, p = stats.kstest(x, stats.gamma(a=alpha, scale=1.0 / beta).cdf)
    assert p > P_MIN, "Gamma KS failed for alpha=%s beta=%s: D=%.4f p=%.2e" % (alpha, beta, d, p)


def test_exponential_draw_ks_and_gamma_api():
    from bnmtf_ard_b200.distributions import exponential_draw, gamma_draw, gamma_expectation, gamma_expectation_log, gamma_mode
    from bnmtf_ard_b200.distributions._device import exp_draws
    for lam in (0.1, 1.0, 25.0):
        x = exp_draws(np.full(N, lam), seed=int(lam * 10) + 5)
        d, p = stats.kstest(x, stats.expon(scale=1.0 / lam).cdf)
        assert p > P_MIN and (x > 0).all()
    assert exponential_draw(2.0) > 0.0 and gamma_draw(2.0, 3.0) > 0.0
    # tests/code/distributions/test_gamma.py:11-37
    assert gamma_expectation(2.0, 3.0) == 2.0 / 3.0
    assert gamma_expectation_log(2.0, 3.0) == -0.67582795356964265
    assert gamma_mode(2.0, 3.0) == 1.0 / 3.0


def test_streams_are_independent():
    """Different (row, k) counters give uncorrelated draws; same counters identical draws."""
    from bnmtf_ard_b200.distributions._device import tn_draw
    mu, tau = np.full(N, 0.5), np.full(N, 2.0)
    a, b, c = tn_draw(mu, tau, seed=11), tn_draw(mu, tau, seed=11), tn_draw(mu, tau, seed=12)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert abs(np.corrcoef(a, c)[0, 1]) < 0.03 and abs(np.corrcoef(a[:-1], a[1:])[0, 1]) < 0.03
