#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Golden fixtures for the mask / fold generators and the log files of the
cross-validation drivers, produced by the REAL reference (oracle/_ref) with a deterministic stand-in
model: python oracle/make_ref.py && python oracle/gen_golden_cv.py"""
import contextlib
import io
import json
import os
import random
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refimport  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


class StandIn(object):
    """Deterministic method(R, M, **parameters) / train / predict used to exercise the drivers on the CPU."""

    def __init__(self, R, M, K, scale=1.0):
        self.R, self.M, self.K, self.scale = np.array(R, dtype=float), np.array(M, dtype=float), K, scale

    def train(self, iterations):
        self.level = (self.M * self.R).sum() / self.M.sum() * self.scale + 0.01 * self.K + 0.001 * iterations

    def predict(self, M_test, burn_in=0, thinning=1):
        M_test = np.array(M_test, dtype=float)
        mse = (M_test * (self.R - self.level) ** 2).sum() / M_test.sum()
        return {'MSE': float(mse), 'R^2': float(1.0 - mse / (1.0 + burn_in)), 'Rp': float(0.5 / thinning)}


def main():
    cv = refimport.load_cv()
    if cv is None:
        raise SystemExit("oracle/_ref missing")
    rng = np.random.default_rng(77)
    I, J = 14, 9
    M = (rng.random((I, J)) < 0.8).astype(float)
    M[np.arange(I), rng.integers(0, J, I)] = 1
    M[rng.integers(0, I, J), np.arange(J)] = 1
    R = rng.normal(2.0, 1.0, (I, J))
    out = {"cv_M": M, "cv_R": R}
    random.seed(5)
    tr, te = cv.mask.compute_folds(I, J, 4, M)
    out["folds_train"], out["folds_test"] = np.array(tr), np.array(te)
    random.seed(6)
    a, b = cv.mask.generate_M(I, J, 0.4, M)
    out["gen_train"], out["gen_test"] = a, b
    random.seed(7)
    a, b = cv.mask.try_generate_M(I, J, 0.35, 50, M)
    out["try_train"], out["try_test"] = a, b
    random.seed(8)
    tr, te = cv.mask.compute_folds_attempts(I, J, 3, 100, None)
    out["folds_full_train"], out["folds_full_test"] = np.array(tr), np.array(te)
    out["nonzero"] = np.array(cv.mask.nonzero_indices(M))
    out["compute_Ms"] = np.array(cv.mask.compute_Ms(list(out["folds_test"])))
    out["inverse"] = cv.mask.calc_inverse_M(out["folds_test"][0], M)
    # driver logs (stratified folds come from sklearn + numpy's global stream)
    logs = {}
    search = [{'K': 2, 'scale': 1.0}, {'K': 3, 'scale': 0.9}, {'K': 4, 'scale': 1.2}]
    with tempfile.TemporaryDirectory() as d, contextlib.redirect_stdout(io.StringIO()):
        np.random.seed(1)
        f = os.path.join(d, "a.txt")
        c = cv.MatrixCrossValidation(StandIn, R, M, 3, search, {'iterations': 5}, {'burn_in': 1, 'thinning': 2}, f)
        c.run(); best = c.find_best_parameters('MSE', True); c.fout.close()
        logs["matrix"] = open(f).read(); logs["matrix_best"] = json.dumps([best[0], best[1]])
        np.random.seed(2)
        f = os.path.join(d, "b.txt")
        c = cv.MatrixSingleCrossValidation(StandIn, R, M, 4, search[1], {'iterations': 7}, {'burn_in': 2, 'thinning': 1}, f)
        c.run(); c.fout.close()
        logs["single"] = open(f).read()
        np.random.seed(3)
        f = os.path.join(d, "c.txt"); nested = [os.path.join(d, "n%d.txt" % i) for i in range(3)]
        c = cv.MatrixNestedCrossValidation(StandIn, R, M, 3, 2, search, {'iterations': 4}, {'burn_in': 0, 'thinning': 1}, f, nested)
        c.run(parallel=False); c.fout.close()
        logs["nested"] = open(f).read()
        logs["nested_inner"] = [open(n).read() for n in nested]
    with open(os.path.join(OUT, "cv_logs.json"), "w") as fjson:
        json.dump(logs, fjson, indent=1)
    np.savez_compressed(os.path.join(OUT, "cv_golden.npz"), **out)
    print("wrote cv_golden.npz (%d arrays) and cv_logs.json" % len(out))


if __name__ == "__main__":
    main()
