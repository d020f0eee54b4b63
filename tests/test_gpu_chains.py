"""Chain-level validation of the Gibbs sampler (BASELINE.json north_star, correctness level 3: "posterior-mean U V^T
and test MSE within the spread of 10 reference chains").

The reference's random stream (MT19937 + Chopin tables) cannot be matched, so the engine's chains are compared with the
reference's as samples of the same posterior: tests/golden/config_chains.npz holds 10 seeded chains of the REAL reference
(oracle/gen_golden_configs.py chains) on
  * the toy matrix data/toy/bnmf/R.txt, K = 10, ARD, 500 iterations, burn-in 250, thinning 2, and
  * the GDSC IC50 matrix, K = 20, ARD, 200 iterations, burn-in 180, thinning 2 -- the settings of
    experiments/experiments_gdsc/cross_validation/nmf_gibbs_ard/nmf_gibbs_ard_xval.py (published 10-fold MSE 708.55),
each with the same tenth of the observed entries held out.  Ten engine chains on the same split must give
  * held-out MSEs from the same distribution (Welch t and two-sample KS, p > 1e-3),
  * a mean posterior prediction of the held-out entries that differs from the reference chains' mean by less than
    the reference's own chain-to-chain spread, and
  * training MSE and posterior-mean tau inside the reference chains' range (+- 4 sd).
"""
import os

import numpy as np
import pytest
from scipy import stats

pytestmark = pytest.mark.gpu
HYPER = {"alphatau": 1.0, "betatau": 1.0, "alpha0": 1.0, "beta0": 1.0}
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config_chains.npz")


def engine_chains(g, p, n_chains, on_device):
    from bnmtf_ard_b200 import bnmf_gibbs
    R, train, test = g[p + "R"], g[p + "M_train"], g[p + "M_test"]
    K, its, burn, thin = int(g[p + "K"]), int(g[p + "its"]), int(g[p + "burn_in"]), int(g[p + "thinning"])
    ti, tj = np.nonzero(test)
    perf, preds, train_mse, taus = [], [], [], []
    for c in range(n_chains):
        m = bnmf_gibbs(R, train, K, True, dict(HYPER))
        m.verbose = False
        if on_device:
            m.summarise_on_device(burn, thin)
        np.random.seed(2000 + c)
        m.initialise("random")
        m.run(its)
        pf = m.predict(test, burn, thin)
        eU, eV, etau, _ = m.approx_expectation(burn, thin)
        preds.append((eU @ eV.T)[ti, tj])
        perf.append([pf["MSE"], pf["R^2"], pf["Rp"]])
        train_mse.append(m.all_performances["MSE"][-1])
        taus.append(etau)
    return np.array(perf), np.array(preds), np.array(train_mse), np.array(taus)


@pytest.mark.parametrize("p,on_device", [("toy_", False), ("gdsc_", False), ("gdsc_", True)])
def test_engine_chains_match_the_spread_of_ten_reference_chains(p, on_device):
    g = np.load(GOLD)
    ref_perf, ref_pred = g[p + "perf"], g[p + "pred"]
    n = ref_perf.shape[0]
    assert n == 10
    perf, pred, train_mse, taus = engine_chains(g, p, n, on_device)
    # held-out MSE: same distribution
    t = stats.ttest_ind(perf[:, 0], ref_perf[:, 0], equal_var=False)
    ks = stats.ks_2samp(perf[:, 0], ref_perf[:, 0])
    msg = "held-out MSE engine %.4f +- %.4f, reference %.4f +- %.4f (t p=%.3g, KS p=%.3g)" % (
        perf[:, 0].mean(), perf[:, 0].std(ddof=1), ref_perf[:, 0].mean(), ref_perf[:, 0].std(ddof=1), t.pvalue, ks.pvalue)
    print(msg)
    assert t.pvalue > 1e-3 and ks.pvalue > 1e-3, msg
    for col, name in ((1, "R^2"), (2, "Rp")):
        assert stats.ttest_ind(perf[:, col], ref_perf[:, col], equal_var=False).pvalue > 1e-3, name
    # posterior-mean prediction of the held-out entries
    ref_mean, eng_mean = ref_pred.mean(axis=0), pred.mean(axis=0)
    spread = np.sqrt(((ref_pred - ref_mean) ** 2).mean(axis=1))         # each reference chain against the reference mean
    diff = float(np.sqrt(((eng_mean - ref_mean) ** 2).mean()))
    eng_spread = np.sqrt(((pred - eng_mean) ** 2).mean(axis=1))
    print("mean-prediction rms difference %.4f; reference chain spread %.4f .. %.4f; engine chain spread %.4f .. %.4f" % (
        diff, spread.min(), spread.max(), eng_spread.min(), eng_spread.max()))
    assert diff < spread.min(), (diff, spread)
    assert 0.6 * spread.mean() < eng_spread.mean() < 1.6 * spread.mean()    # the chains mix like the reference's
    # training error and noise precision
    for got, want, name in ((train_mse, g[p + "train_mse"], "train MSE"), (taus, g[p + "tau_mean"], "tau")):
        sd = max(want.std(ddof=1), 1e-12)
        assert abs(got.mean() - want.mean()) < 4 * sd * np.sqrt(2.0 / n) + 0.002 * abs(want.mean()), (name, got.mean(), want.mean(), sd)
