"""CPU: the C-ABI library loads and exports every symbol include/bnmtf_b200.h
declares; host-side constructor validation reproduces the reference's messages;
compute entry points fail loudly without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "bnmtf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bnmtf_[A-Za-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from bnmtf_ard_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
        assert n in _lib.SIGNATURES, "no ctypes signature for %s" % n
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.bnmtf_abi_version() == 1


def test_no_cpu_fallback():
    from bnmtf_ard_b200 import _lib, bnmf_gibbs
    from bnmtf_ard_b200.distributions import TN_vector_expectation
    if _lib.load().bnmtf_device_count() > 0:
        pytest.skip("GPU present")
    m = bnmf_gibbs(np.ones((3, 2)), np.ones((3, 2)), 2, True, {'alphatau': 1, 'betatau': 1, 'alpha0': 1, 'beta0': 1})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.initialise('exp')
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        TN_vector_expectation([1.0], [1.0])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "bnmtf_ard_b200")
    for root, _d, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, fn)).read()
                assert "oracle" not in text.lower().replace("# oracle", ""), "%s mentions the oracle" % fn
                assert "/root/reference" not in text


@pytest.mark.parametrize("clsname", ["bnmf_gibbs", "nmf_icm", "bnmf_vb"])
def test_constructor_messages(clsname):
    """tests/code/test_bnmf_gibbs.py:14-105 (identical blocks in test_nmf_icm.py / test_bnmf_vb.py)."""
    import bnmtf_ard_b200 as pkg
    cls = getattr(pkg, clsname)
    M = np.ones((2, 3))
    I, J, K = 5, 3, 1
    lambdaU, lambdaV = np.ones((I, K)), np.ones((J, K))
    hp = {'alphatau': 3, 'betatau': 1, 'alpha0': 6, 'beta0': 2}
    with pytest.raises(AssertionError) as error:
        cls(np.ones(3), M, K, True, hp)
    assert str(error.value) == "Input matrix R is not a two-dimensional array, but instead 1-dimensional."
    with pytest.raises(AssertionError) as error:
        cls(np.ones((4, 3, 2)), M, K, True, hp)
    assert str(error.value) == "Input matrix R is not a two-dimensional array, but instead 3-dimensional."
    with pytest.raises(AssertionError) as error:
        cls(np.ones((3, 2)), M, K, True, hp)
    assert str(error.value) == "Input matrix R is not of the same size as the indicator matrix M: (3, 2) and (2, 3) respectively."
    R4 = np.ones((2, 3))
    with pytest.raises(AssertionError) as error:
        cls(R4, M, K, False, {'alphatau': 3, 'betatau': 1, 'lambdaU': np.ones((3, 1)), 'lambdaV': lambdaV})
    assert str(error.value) == "Prior matrix lambdaU has the wrong shape: (3, 1) instead of (2, 1)."
    with pytest.raises(AssertionError) as error:
        cls(R4, M, K, False, {'alphatau': 3, 'betatau': 1, 'lambdaU': np.ones((2, 1)), 'lambdaV': np.ones((4, 1))})
    assert str(error.value) == "Prior matrix lambdaV has the wrong shape: (4, 1) instead of (3, 1)."
    with pytest.raises(AssertionError) as error:
        cls(R4, [[1, 1, 1], [0, 0, 0]], K, True, hp)
    assert str(error.value) == "Fully unobserved row in R, row 1."
    with pytest.raises(AssertionError) as error:
        cls(R4, [[1, 1, 0], [1, 0, 0]], K, True, hp)
    assert str(error.value) == "Fully unobserved column in R, column 2."
    # a valid construction: scalars broadcast, attributes as in the reference
    I, J, K = 3, 2, 2
    m = cls(np.ones((I, J)), np.ones((I, J)), K, False, {'alphatau': 3, 'betatau': 1, 'lambdaU': 2., 'lambdaV': 3.})
    assert (m.I, m.J, m.K, m.size_Omega, m.alphatau, m.betatau) == (I, J, K, I * J, 3, 1)
    assert np.array_equal(m.lambdaU, 2. * np.ones((I, K))) and np.array_equal(m.lambdaV, 3. * np.ones((J, K)))
    with pytest.raises(AssertionError) as error:
        m.initialise('FAIL')
    assert str(error.value) == "Unknown initialisation option: FAIL. Should be in ['random', 'exp']."


def test_np_constructor_and_run_guard():
    from bnmtf_ard_b200 import nmf_np
    with pytest.raises(AssertionError) as error:
        nmf_np(np.ones(3), np.ones((2, 3)), 1)
    assert str(error.value) == "Input matrix R is not a two-dimensional array, but instead 1-dimensional."
    with pytest.raises(AssertionError) as error:
        nmf_np(np.ones((2, 3)), [[1, 1, 1], [0, 0, 0]], 1)
    assert str(error.value) == "Fully unobserved row in R, row 1."
    R = np.array([[1., 2.], [3., 4.]]); M = np.array([[1, 0], [1, 1]])
    m = nmf_np(R, M, 2)
    assert np.array_equal(m.R_excl_unknown, [[1., 1.], [3., 4.]])     # nmf_np.py:54-57
    with pytest.raises(AssertionError) as error:
        m.initialise('FAIL')
    assert str(error.value) == "Unrecognised init option for U,V: FAIL. Should be one in ['ones', 'random', 'exponential']."
    with pytest.raises(AssertionError) as error:
        m.run(1)
    assert str(error.value) == "U and V have not been initialised - please run NMF.initialise() first."
    m.initialise('ones')
    assert np.array_equal(m.U, np.ones((2, 2)))


def test_host_metrics_and_predict(golden):
    """compute_MSE/R2/Rp, approx_expectation, predict, quality on injected traces
    (tests/code/test_bnmf_gibbs.py:256-377) against the reference's own outputs."""
    from bnmtf_ard_b200 import bnmf_gibbs
    g = golden
    R, M, P = g["met_R"], g["met_M"], g["met_P"]
    m = bnmf_gibbs(R, M, 3, True, {'alphatau': 1, 'betatau': 1, 'alpha0': 1, 'beta0': 1})
    np.testing.assert_allclose([m.compute_MSE(M, R, P), m.compute_R2(M, R, P), m.compute_Rp(M, R, P)], g["met_vals"], rtol=1e-14)
    m.all_U, m.all_V, m.all_tau, m.all_lambdak = list(g["met_all_U"]), list(g["met_all_V"]), g["met_all_tau"], g["met_all_lambdak"]
    perf = m.predict(g["met_Mtest"], 2, 3)
    np.testing.assert_allclose([perf["MSE"], perf["R^2"], perf["Rp"]], g["met_predict"], rtol=1e-13)
    np.testing.assert_allclose([m.quality(q, 2, 3) for q in ("loglikelihood", "BIC", "AIC", "MSE", "ELBO")],
                               g["met_quality"], rtol=1e-13)
    with pytest.raises(AssertionError) as error:
        m.quality('FAIL', 2, 3)
    assert str(error.value) == "Unrecognised metric for model quality: FAIL."


@pytest.mark.parametrize("clsname", ["bnmtf_gibbs", "nmtf_icm", "bnmtf_vb"])
def test_nmtf_constructor_messages(clsname):
    """tests/code/test_bnmtf_gibbs.py:14-120 (same blocks in test_nmtf_icm.py / test_bnmtf_vb.py)."""
    import bnmtf_ard_b200 as pkg
    cls = getattr(pkg, clsname)
    M = np.ones((2, 3))
    I, J, K, L = 5, 3, 1, 2
    lambdaF, lambdaS, lambdaG = np.ones((I, K)), np.ones((K, L)), np.ones((J, L))
    hp = {'alphatau': 3, 'betatau': 1, 'alpha0': 6, 'beta0': 2, 'lambdaS': lambdaS}
    with pytest.raises(AssertionError) as error:
        cls(np.ones(3), M, K, L, True, hp)
    assert str(error.value) == "Input matrix R is not a two-dimensional array, but instead 1-dimensional."
    with pytest.raises(AssertionError) as error:
        cls(np.ones((3, 2)), M, K, L, True, hp)
    assert str(error.value) == "Input matrix R is not of the same size as the indicator matrix M: (3, 2) and (2, 3) respectively."
    R4 = np.ones((2, 3))
    base = {'alphatau': 3, 'betatau': 1}
    with pytest.raises(AssertionError) as error:
        cls(R4, M, K, L, False, dict(base, lambdaF=np.ones((3, 1)), lambdaS=lambdaS, lambdaG=lambdaG))
    assert str(error.value) == "Prior matrix lambdaF has the wrong shape: (3, 1) instead of (2, 1)."
    with pytest.raises(AssertionError) as error:
        cls(R4, M, K, L, False, dict(base, lambdaF=np.ones((2, 1)), lambdaS=np.ones((2, 2)), lambdaG=lambdaG))
    assert str(error.value) == "Prior matrix lambdaS has the wrong shape: (2, 2) instead of (1, 2)."
    with pytest.raises(AssertionError) as error:
        cls(R4, M, K, L, False, dict(base, lambdaF=np.ones((2, 1)), lambdaS=lambdaS, lambdaG=np.ones((4, 2))))
    assert str(error.value) == "Prior matrix lambdaG has the wrong shape: (4, 2) instead of (3, 2)."
    with pytest.raises(AssertionError) as error:
        cls(R4, [[1, 1, 1], [0, 0, 0]], K, L, True, hp)
    assert str(error.value) == "Fully unobserved row in R, row 1."
    with pytest.raises(AssertionError) as error:
        cls(R4, [[1, 1, 0], [1, 0, 0]], K, L, True, hp)
    assert str(error.value) == "Fully unobserved column in R, column 2."
    m = cls(np.ones((3, 2)), np.ones((3, 2)), 2, 3, False, dict(base, lambdaF=2., lambdaS=4., lambdaG=3.))
    assert (m.I, m.J, m.K, m.L, m.size_Omega) == (3, 2, 2, 3, 6)
    assert np.array_equal(m.lambdaF, 2. * np.ones((3, 2))) and np.array_equal(m.lambdaS, 4. * np.ones((2, 3)))
    assert np.array_equal(m.lambdaG, 3. * np.ones((2, 3)))
    with pytest.raises(AssertionError) as error:
        m.initialise('FAIL', 'exp')
    assert str(error.value) == "Unknown initialisation option for F and G: FAIL. Should be in ['kmeans', 'random', 'exp']."
    with pytest.raises(AssertionError) as error:
        m.initialise('exp', 'FAIL')
    assert str(error.value) == "Unknown initialisation option for S: FAIL. Should be in ['random', 'exp']."
    assert m.number_parameters() == 3 * 2 + 2 * 3 + 2 * 3 + 1


@pytest.mark.parametrize("clsname", ["bnmtf_gibbs", "nmtf_icm", "bnmtf_vb", "nmtf_np"])
def test_nmtf_size_limits_are_raised_in_the_constructor(clsname):
    """The engine's limits ("Limits": L <= 48, K*L <= 3631) surface at construction with a clear
    message, never in the middle of run(); the reference's largest model-selection setting K = L = 40 is inside them."""
    import bnmtf_ard_b200 as pkg
    from bnmtf_ard_b200._nmf_common import NMTF_MAX_KL, NMTF_MAX_L
    assert (NMTF_MAX_L, NMTF_MAX_KL) == (48, 3631)
    cls = getattr(pkg, clsname)
    R, M = np.ones((4, 5)), np.ones((4, 5))
    hp = {'alphatau': 1., 'betatau': 1., 'alpha0': 1., 'beta0': 1., 'lambdaS': 1.}
    make = (lambda K, L: cls(R, M, K, L)) if clsname == "nmtf_np" else (lambda K, L: cls(R, M, K, L, True, hp))
    make(40, 40)
    with pytest.raises(ValueError, match="L = 49 is not supported"):
        make(2, 49)
    with pytest.raises(ValueError, match=r"K\*L = 3640 is not supported"):
        make(91, 40)


def test_nmtf_np_constructor_and_guards():
    from bnmtf_ard_b200 import nmtf_np
    with pytest.raises(AssertionError) as error:
        nmtf_np(np.ones((2, 3)), [[1, 1, 1], [0, 0, 0]], 1, 2)
    assert str(error.value) == "Fully unobserved row in R, row 1."
    m = nmtf_np(np.array([[1., 2.], [3., 4.]]), np.array([[1, 0], [1, 1]]), 2, 1)
    assert np.array_equal(m.R_excl_unknown, [[1., 1.], [3., 4.]])
    with pytest.raises(AssertionError) as error:
        m.initialise('FAIL', 'ones')
    assert str(error.value) == "Unrecognised init option for F,G: FAIL. Should be one in ['kmeans', 'ones', 'random', 'exponential']."
    with pytest.raises(AssertionError) as error:
        m.initialise('ones', 'FAIL')
    assert str(error.value) == "Unrecognised init option for S: FAIL. Should be one in ['ones', 'random', 'exponential']."
    with pytest.raises(AssertionError) as error:
        m.run(1)
    assert str(error.value) == "F, S and G have not been initialised - please run NMTF.initialise() first."
    m.initialise('ones', 'ones')
    assert m.F.shape == (2, 2) and m.S.shape == (2, 1) and m.G.shape == (2, 1)
    np.testing.assert_array_equal(m.triple_dot(m.F, m.S, m.G.T), 2. * np.ones((2, 2)))


def test_performances_all_equals_rowwise():
    """The vectorised per-iteration metrics are element for element the row-wise ones."""
    import numpy as np
    from bnmtf_ard_b200 import _lib
    from bnmtf_ard_b200._engine import NMFEngine
    eng = NMFEngine.__new__(NMFEngine)
    eng.size_Omega, eng.ss_total = 1234.0, 987.654321
    rng = np.random.default_rng(0)
    stats = np.zeros((50, _lib.NSTATS))
    stats[:, _lib.ST_RSS] = rng.uniform(1, 2000, 50)
    stats[:, _lib.ST_SUM_E] = rng.normal(0, 30, 50)
    stats[:, _lib.ST_SUM_RCE] = rng.normal(0, 300, 50)
    allp = eng.performances_all(stats)
    for i, row in enumerate(stats):
        one = eng.performances(row)
        for key in ("MSE", "R^2", "Rp"):
            assert one[key] == allp[key][i] or (np.isnan(one[key]) and np.isnan(allp[key][i]))
    eng.h = None
