#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Golden fixtures for the masked K-means initialisation from the
REAL reference (oracle/_ref/BNMTF_ARD/code/models/kmeans/kmeans.py), seeded through Python's
`random` exactly as the reference does: python oracle/make_ref.py && python oracle/gen_golden_kmeans.py"""
import contextlib
import importlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refimport  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    if refimport.load() is None:
        raise SystemExit("oracle/_ref missing")
    KMeans = importlib.import_module("BNMTF_ARD.code.models.kmeans.kmeans").KMeans
    rng = np.random.default_rng(4242)
    out = {}
    cases = {"a": (40, 25, 4, 0.6, 3), "b": (120, 60, 6, 0.35, 11), "c": (30, 18, 9, 0.5, 5)}   # c: many clusters -> empty-cluster path
    for name, (n, m, K, dens, seed) in cases.items():
        centers = rng.normal(0, 3, (K, m))
        X = centers[rng.integers(0, K, n)] + rng.normal(0, 0.7, (n, m))
        M = (rng.random((n, m)) < dens).astype(float)
        M[np.arange(n), rng.integers(0, m, n)] = 1
        M[rng.integers(0, n, m), np.arange(m)] = 1
        with contextlib.redirect_stdout(io.StringIO()):
            km = KMeans(X, M, K)
            km.initialise(seed)
            c0 = np.array(km.centroids, dtype=float)
            km.cluster()
        out.update({"km_%s_X" % name: X, "km_%s_M" % name: M, "km_%s_K" % name: K, "km_%s_seed" % name: seed,
                    "km_%s_centroids0" % name: c0, "km_%s_assign" % name: np.array(km.cluster_assignments),
                    "km_%s_result" % name: km.clustering_results,
                    "km_%s_centroids" % name: np.array(km.centroids, dtype=float) * np.array(km.mask_centroids),
                    "km_%s_mask" % name: np.array(km.mask_centroids, dtype=float), "km_%s_dist" % name: km.distances})
    path = os.path.join(OUT, "kmeans_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%d arrays)" % (path, len(out)))


if __name__ == "__main__":
    main()
