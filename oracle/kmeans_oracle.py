"""
TEST INFRASTRUCTURE ONLY -- numpy restatement of the masked K-means of
code/models/kmeans/kmeans.py (assignment :88-125, update :131-170, cluster :70-84),
pinned against the real reference by tests/test_oracle_golden_kmeans.py.
"""
import random

import numpy as np


def assign(X, M, centroids, mask_centroids):
    """closest_cluster for every point: lowest index among the smallest defined MSEs (kmeans.py:105-125)."""
    n, K = X.shape[0], centroids.shape[0]
    best = np.full(n, -1)
    best_mse = np.zeros(n)
    best_def = np.zeros(n, dtype=bool)
    for c in range(K):
        mm = M * mask_centroids[c][None, :]
        cnt = mm.sum(axis=1)
        with np.errstate(all="ignore"):
            mse = (mm * (X - centroids[c][None, :]) ** 2).sum(axis=1) / cnt
        defined = cnt != 0
        take = (~best_def) | (defined & (mse < best_mse))
        best = np.where(take, c, best)
        best_mse = np.where(take, np.where(defined, mse, 0.0), best_mse)
        best_def = np.where(take, defined, best_def)
    return best, np.where(best_def, best_mse, np.nan)


def cluster(X, M, K, seed, max_iterations=200):
    """initialise(seed) + cluster() of the reference, singleton resolution of empty clusters."""
    n, m = X.shape
    random.seed(seed)
    mins = [X[M[:, j] != 0, j].min() for j in range(m)]
    maxs = [X[M[:, j] != 0, j].max() for j in range(m)]
    centroids = np.array([[random.uniform(mins[j], maxs[j]) for j in range(m)] for _ in range(K)])
    masks = np.ones((K, m))
    assignments = np.full(n, -1)
    n_unique = len(set(row.tobytes() for row in np.ascontiguousarray(X)))
    iteration, change = 1, True
    while change:
        iteration += 1
        new, dist = assign(X, M, centroids, masks)
        change = bool(np.any(new != assignments))
        assignments = new.copy()
        members = [list(np.flatnonzero(assignments == c)) for c in range(K)]
        flags = np.zeros(K, dtype=bool)

        def update_cluster(c):
            if len(members[c]) == 0:
                flags[c] = False
                if n_unique >= K:
                    idx = int(dist.argmax())
                    old = assignments[idx]
                    dist[idx] = 0.0
                    assignments[idx] = c
                    members[c] = [idx]
                    flags[c] = True
                    members[old].remove(idx)
                    update_cluster(old)
            else:
                flags[c] = True
        for c in range(K):
            update_cluster(c)
        for c in range(K):
            if flags[c]:
                sel = assignments == c
                cnt = M[sel].sum(axis=0)
                with np.errstate(all="ignore"):
                    centroids[c] = np.where(cnt > 0, (M[sel] * X[sel]).sum(axis=0) / cnt, 0.0)
                masks[c] = (cnt > 0).astype(float)
        if iteration >= max_iterations:
            break
    return assignments, centroids * masks, masks
