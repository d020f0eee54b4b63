"""Where does a small fit spend its time?  (C5-shaped 887 x 545 BNMF VB, 200 iterations.)
Prints wall-clock per phase for one fit and the throughput of P concurrent fits."""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--its", type=int, default=200)
    ap.add_argument("--fits", type=int, default=40)
    args = ap.parse_args()
    from bnmtf_ard_b200 import bnmf_vb, bnmf_gibbs
    rng = np.random.default_rng(5)
    I, J = 887, 545
    R = rng.exponential(1, (I, 10)) @ rng.exponential(1, (J, 10)).T + rng.normal(0, 1, (I, J))
    M = (rng.random((I, J)) < 0.8).astype(float)
    M[np.arange(I), rng.integers(0, J, I)] = 1; M[rng.integers(0, I, J), np.arange(J)] = 1
    Mt = ((rng.random((I, J)) < 0.1) * M).astype(float)
    Mtr = M - Mt
    hp = {'alphatau': 1., 'betatau': 1., 'alpha0': 1., 'beta0': 1.}

    def fit(K, cls=bnmf_vb, phases=None):
        t = [time.perf_counter()]
        m = cls(R, Mtr, K, True, hp); m.verbose = False
        m._engine(); t.append(time.perf_counter())
        m.initialise('random'); t.append(time.perf_counter())
        if cls is bnmf_gibbs:
            m.summarise_on_device(args.its // 2, 2)
        m.run(args.its); t.append(time.perf_counter())
        p = m.predict(Mt) if cls is bnmf_vb else m.predict(Mt, args.its // 2, 2); t.append(time.perf_counter())
        if phases is not None:
            phases.append({"K": K, "create_ms": 1e3 * (t[1] - t[0]), "init_ms": 1e3 * (t[2] - t[1]), "run_ms": 1e3 * (t[3] - t[2]),
                           "device_run_ms": 1e3 * m.all_times[-1], "predict_ms": 1e3 * (t[4] - t[3])})
        return p['MSE']

    fit(5)   # warm-up (context, module load)
    out = {}
    for cls in (bnmf_vb, bnmf_gibbs):
        ph = []
        for K in (5, 10, 30):
            fit(K, cls, ph)
        out[cls.__name__ + " phases"] = ph
    for P in (1, 2, 4, 10, 20):
        t0 = time.perf_counter()
        with ThreadPoolExecutor(P) as ex:
            list(ex.map(lambda i: fit(1 + i % 30), range(args.fits)))
        dt = time.perf_counter() - t0
        out["P=%d" % P] = {"fits": args.fits, "seconds": dt, "fits_per_s": args.fits / dt}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
