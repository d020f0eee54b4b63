"""Experiment helper: the cluster sweep on clusters of 16 CTAs (BNMTF_V2_CS=16) against the default clusters of 8.
The tile partition differs, so the row sums associate differently: ICM trajectories must agree to rounding."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(out):
    from bnmtf_ard_b200 import nmf_icm
    rng = np.random.default_rng(5)
    I, J, K = 40, 6144, 4
    R = np.abs(rng.exponential(1.0, (I, K)) @ rng.exponential(1.0, (J, K)).T + rng.normal(0, 1, (I, J))) + 0.05
    M = (rng.random((I, J)) < 0.9).astype(float)
    m = nmf_icm(R, M, K, True, {"alphatau": 1.0, "betatau": 1.0, "alpha0": 1.0, "beta0": 1.0})
    m.verbose = False
    m.U, m.V, m.tau, m.lambdak = rng.exponential(1.0, (I, K)), rng.exponential(1.0, (J, K)), 1.3, np.ones(K)
    m.run(3)
    info = m._engine().layout_info()
    print(os.environ.get("BNMTF_V2_CS"), {k: info[k] for k in ("v2_U", "v2_npt_U", "v2_rows_per_set_U")})
    np.savez(out, U=m.U, V=m.V, tau=m.tau)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1])
    else:
        outs = []
        for cs in ("8", "16"):
            out = "/tmp/cs%s.npz" % cs
            subprocess.check_call([sys.executable, __file__, out], env=dict(os.environ, BNMTF_V2_CS=cs))
            outs.append(np.load(out))
        for f in ("U", "V", "tau"):
            a, b = outs[0][f], outs[1][f]
            print(f, "max rel diff", float(np.max(np.abs(a - b) / np.maximum(np.abs(a), 1e-300))))
            assert np.allclose(a, b, rtol=1e-9, atol=1e-12)
        print("CS16_CHECK PASS")
