"""cProfile of the batched cross-validation driver on the C5-shaped sweep (887 x 545, 10 folds x K = 1..Kmax,
BNMF VB 200 iterations): where does the host time go?"""
import argparse
import contextlib
import cProfile
import io
import os
import pstats
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kmax", type=int, default=30)
    ap.add_argument("--method", default="vb")
    ap.add_argument("--top", type=int, default=45)
    ap.add_argument("--workers", type=int, default=8)
    ap.add_argument("--noprofile", action="store_true")
    args = ap.parse_args()
    import bnmtf_ard_b200 as pkg
    from bnmtf_ard_b200.cross_validation.batched_matrix_cross_validation import BatchedMatrixCrossValidation
    rng = np.random.default_rng(5)
    I, J = 887, 545
    R = rng.exponential(1, (I, 10)) @ rng.exponential(1, (J, 10)).T + rng.normal(0, 1, (I, J))
    M = (rng.random((I, J)) < 0.8).astype(float)
    M[np.arange(I), rng.integers(0, J, I)] = 1; M[rng.integers(0, I, J), np.arange(J)] = 1
    hp = {'alphatau': 1., 'betatau': 1., 'alpha0': 1., 'beta0': 1.}
    cls0 = {"vb": pkg.bnmf_vb, "icm": pkg.nmf_icm, "gibbs": pkg.bnmf_gibbs}[args.method]

    class Quiet(object):
        def __new__(cls, R, M, **p):
            m = cls0(R, M, **p); m.verbose = False
            return m
    search = [{'K': k, 'ARD': True, 'hyperparameters': hp} for k in range(1, args.kmax + 1)]
    pc = {} if args.method == "vb" else {'burn_in': 100, 'thinning': 2}

    def go():
        with contextlib.redirect_stdout(io.StringIO()):
            cv = BatchedMatrixCrossValidation(method=Quiet, R=R, M=M, K=10, parameter_search=box['search'],
                                              train_config={'iterations': 200, 'init_UV': 'random'}, predict_config=pc,
                                              file_performance="/tmp/cv_prof.txt", workers=args.workers)
            cv.run()
        return cv
    box = {'search': search[:2]}
    go()                      # warm-up: context, module load, pool growth
    box['search'] = search
    if args.noprofile:
        import time
        for rep in range(2):
            t0 = time.perf_counter()
            cv = go()
            print("workers", args.workers, "total_s", time.perf_counter() - t0, "phases", cv.timing, "device_s", cv.device_seconds)
        return
    pr = cProfile.Profile()
    pr.enable()
    cv = go()
    pr.disable()
    print("phases", cv.timing, "launches", cv.launches)
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(args.top)
    print(s.getvalue())


if __name__ == "__main__":
    main()
