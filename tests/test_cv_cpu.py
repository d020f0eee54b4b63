"""CPU: the cross-validation callers (bnmtf_ard_b200/cross_validation) against golden outputs of the REAL
reference (oracle/gen_golden_cv.py): seeded mask / fold generators are identical, and the four drivers
write byte-identical log files for a deterministic stand-in model."""
import contextlib
import io
import json
import os
import random
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from gen_golden_cv import StandIn  # noqa: E402


@pytest.fixture(scope="module")
def gold():
    g = np.load(os.path.join(ROOT, "tests", "golden", "cv_golden.npz"))
    logs = json.load(open(os.path.join(ROOT, "tests", "golden", "cv_logs.json")))
    return g, logs


def test_mask_generators_match_reference(gold):
    from bnmtf_ard_b200.cross_validation import mask
    g, _ = gold
    M = g["cv_M"]
    I, J = M.shape
    random.seed(5)
    tr, te = mask.compute_folds(I, J, 4, M)
    assert np.array_equal(np.array(tr), g["folds_train"]) and np.array_equal(np.array(te), g["folds_test"])
    random.seed(6)
    a, b = mask.generate_M(I, J, 0.4, M)
    assert np.array_equal(a, g["gen_train"]) and np.array_equal(b, g["gen_test"])
    random.seed(7)
    a, b = mask.try_generate_M(I, J, 0.35, 50, M)
    assert np.array_equal(a, g["try_train"]) and np.array_equal(b, g["try_test"])
    random.seed(8)
    tr, te = mask.compute_folds_attempts(I, J, 3, 100, None)
    assert np.array_equal(np.array(tr), g["folds_full_train"]) and np.array_equal(np.array(te), g["folds_full_test"])
    assert np.array_equal(np.array(mask.nonzero_indices(M)), g["nonzero"])
    assert np.array_equal(np.array(mask.compute_Ms(list(g["folds_test"]))), g["compute_Ms"])
    assert np.array_equal(mask.calc_inverse_M(g["folds_test"][0], M), g["inverse"])
    assert mask.check_empty_rows_columns(M) and not mask.check_empty_rows_columns(np.zeros((2, 2)))
    with pytest.raises(AssertionError) as err:
        mask.generate_M(2, 2, 0.1, np.array([[1., 0.], [0., 1.]]))
    assert "Specified 0.1 fraction missing" in str(err.value)


def test_stratified_folds_properties():
    from bnmtf_ard_b200.cross_validation import mask
    rng = np.random.default_rng(3)
    I, J = 30, 50
    M = (rng.random((I, J)) < 0.7).astype(float)
    M[:, 0] = 1; M[0, :] = 1
    np.random.seed(0)
    for fn, axis in ((mask.compute_folds_stratify_rows_attempts, 1), (mask.compute_folds_stratify_columns_attempts, 0)):
        tr, te = fn(I, J, 5, 100, M)
        assert np.array_equal(sum(te), M)                           # the test folds partition Omega
        for a, b in zip(tr, te):
            assert np.array_equal(a + b, M) and mask.check_empty_rows_columns(a)
        counts = np.array([t.sum(axis=axis) for t in te])           # stratified: per row/column the folds differ by <= 1
        assert (counts.max(axis=0) - counts.min(axis=0)).max() <= 1


def test_driver_logs_match_reference(gold, tmp_path):
    from bnmtf_ard_b200.cross_validation.matrix_cross_validation import MatrixCrossValidation
    from bnmtf_ard_b200.cross_validation.matrix_single_cross_validation import MatrixSingleCrossValidation
    from bnmtf_ard_b200.cross_validation.nested_matrix_cross_validation import MatrixNestedCrossValidation
    from bnmtf_ard_b200.cross_validation.parallel_matrix_cross_validation import ParallelMatrixCrossValidation
    g, logs = gold
    R, M = g["cv_R"], g["cv_M"]
    search = [{'K': 2, 'scale': 1.0}, {'K': 3, 'scale': 0.9}, {'K': 4, 'scale': 1.2}]
    with contextlib.redirect_stdout(io.StringIO()):
        np.random.seed(1)
        f = str(tmp_path / "a.txt")
        c = MatrixCrossValidation(StandIn, R, M, 3, search, {'iterations': 5}, {'burn_in': 1, 'thinning': 2}, f)
        c.run(); best = c.find_best_parameters('MSE', True); c.fout.close()
        assert open(f).read() == logs["matrix"] and json.dumps([best[0], best[1]]) == logs["matrix_best"]
        np.random.seed(2)
        f = str(tmp_path / "b.txt")
        c = MatrixSingleCrossValidation(StandIn, R, M, 4, search[1], {'iterations': 7}, {'burn_in': 2, 'thinning': 1}, f)
        c.run(); c.fout.close()
        assert open(f).read() == logs["single"]
        np.random.seed(3)
        f = str(tmp_path / "c.txt"); nested = [str(tmp_path / ("n%d.txt" % i)) for i in range(3)]
        c = MatrixNestedCrossValidation(StandIn, R, M, 3, 2, search, {'iterations': 4}, {'burn_in': 0, 'thinning': 1}, f, nested)
        c.run(parallel=False); c.fout.close()
        assert open(f).read() == logs["nested"]
        assert [open(n).read() for n in nested] == logs["nested_inner"]
        # the thread-pool driver: same folds, same bookkeeping; predict() is called without predict_config
        # like the reference's run_model (parallel_matrix_cross_validation.py:28-31)
        np.random.seed(1)
        f = str(tmp_path / "p.txt")
        p = ParallelMatrixCrossValidation(StandIn, R, M, 3, search, {'iterations': 5}, {'burn_in': 1, 'thinning': 2}, f, P=3)
        p.run(); p.fout.close()
        np.random.seed(1)
        f2 = str(tmp_path / "s.txt")
        s = MatrixCrossValidation(StandIn, R, M, 3, search, {'iterations': 5}, {}, f2)
        s.run(); s.fout.close()
        assert open(f).read() == open(f2).read()
        with pytest.raises(Exception):
            p.run_model(None, None, None)
        # the batched driver with a model that cannot defer its run: fitted in place, the reference's log
        from bnmtf_ard_b200.cross_validation.batched_matrix_cross_validation import BatchedMatrixCrossValidation
        np.random.seed(1)
        f = str(tmp_path / "bt.txt")
        c = BatchedMatrixCrossValidation(StandIn, R, M, 3, search, {'iterations': 5}, {'burn_in': 1, 'thinning': 2}, f)
        c.run(); best = c.find_best_parameters('MSE', True); c.fout.close()
        assert open(f).read() == logs["matrix"] and json.dumps([best[0], best[1]]) == logs["matrix_best"]


def test_stratified_folds_equal_sklearn():
    """The vectorised stratified assignment reproduces scikit-learn's StratifiedKFold(shuffle=True)
    draw for draw (same numpy global state in, same folds out)."""
    from sklearn.model_selection import StratifiedKFold
    from bnmtf_ard_b200.cross_validation import mask
    rng = np.random.default_rng(8)
    for n_classes, n_splits in ((17, 5), (60, 10), (5, 3)):
        sizes = rng.integers(n_splits, 40, n_classes)
        labels = np.repeat(np.arange(n_classes) * 3 + 1, sizes)
        np.random.seed(123)
        ref = np.empty(labels.size, dtype=int)
        for f, (_, test) in enumerate(StratifiedKFold(n_splits=n_splits, shuffle=True).split(np.zeros(labels.size), labels)):
            ref[test] = f
        state_ref = np.random.get_state()[1].copy()
        np.random.seed(123)
        got = mask._stratified_test_folds(labels, n_splits)
        assert np.array_equal(got, ref)
        assert np.array_equal(np.random.get_state()[1], state_ref)      # same consumption of the global stream
