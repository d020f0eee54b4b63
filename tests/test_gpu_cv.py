"""GPU: the cross-validation drivers running the device models (serial, thread-pool parallel, nested)."""
import contextlib
import io

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def problem(I=60, J=40, K=3, seed=0):
    rng = np.random.default_rng(seed)
    R = rng.exponential(1, (I, K)) @ rng.exponential(1, (J, K)).T + rng.normal(0, 0.5, (I, J))
    M = (rng.random((I, J)) < 0.8).astype(float)
    M[np.arange(I), rng.integers(0, J, I)] = 1
    M[rng.integers(0, I, J), np.arange(J)] = 1
    return R, M


class QuietVB(object):
    """bnmf_vb behind the method protocol, without the per-iteration prints."""
    def __new__(cls, R, M, **parameters):
        from bnmtf_ard_b200 import bnmf_vb
        m = bnmf_vb(R, M, **parameters)
        m.verbose = False
        return m


def test_parallel_cv_equals_serial_cv(tmp_path):
    from bnmtf_ard_b200.cross_validation.matrix_cross_validation import MatrixCrossValidation
    from bnmtf_ard_b200.cross_validation.parallel_matrix_cross_validation import ParallelMatrixCrossValidation
    R, M = problem()
    hp = {'alphatau': 1., 'betatau': 1., 'alpha0': 1., 'beta0': 1.}
    search = [{'K': k, 'ARD': True, 'hyperparameters': hp} for k in (2, 3, 5)]
    out = []
    for cls, extra in ((MatrixCrossValidation, {}), (ParallelMatrixCrossValidation, {'P': 4})):
        np.random.seed(4)
        f = str(tmp_path / (cls.__name__ + ".txt"))
        with contextlib.redirect_stdout(io.StringIO()):
            cv = cls(method=QuietVB, R=R, M=M, K=4, parameter_search=search, train_config={'iterations': 30, 'init_UV': 'exp'},
                     predict_config={}, file_performance=f, **extra)
            cv.run()
            best, perf = cv.find_best_parameters('MSE', True)
        out.append((cv.performances['MSE'], best['K'], open(f).read()))
    # VB with init 'exp' is deterministic: the concurrent folds must give exactly the serial numbers
    assert out[0][0] == out[1][0] and out[0][1] == out[1][1] and out[0][2] == out[1][2]
    assert out[0][1] in (3, 5) and "Best performances:" in out[0][2]


def test_nested_cv_gibbs(tmp_path):
    from bnmtf_ard_b200 import bnmf_gibbs
    from bnmtf_ard_b200.cross_validation.nested_matrix_cross_validation import MatrixNestedCrossValidation
    R, M = problem(50, 30, 2, 1)
    hp = {'alphatau': 1., 'betatau': 1., 'alpha0': 1., 'beta0': 1.}

    class QuietGibbs(object):
        def __new__(cls, R, M, **parameters):
            m = bnmf_gibbs(R, M, **parameters); m.verbose = False
            return m
    np.random.seed(2)
    with contextlib.redirect_stdout(io.StringIO()):
        cv = MatrixNestedCrossValidation(method=QuietGibbs, R=R, M=M, K=3, P=2,
                                         parameter_search=[{'K': k, 'ARD': True, 'hyperparameters': hp} for k in (1, 2)],
                                         train_config={'iterations': 40, 'init_UV': 'random'},
                                         predict_config={'burn_in': 20, 'thinning': 2}, file_performance=str(tmp_path / "o.txt"),
                                         files_nested_performances=[str(tmp_path / ("n%d.txt" % i)) for i in range(3)])
        cv.run(parallel=False)
    assert len(cv.all_performances['MSE']) == 3 and np.isfinite(cv.average_performances['MSE'])
    assert cv.average_performances['Rp'] > 0.5
    assert open(str(tmp_path / "o.txt")).read().startswith("Average performances:")
