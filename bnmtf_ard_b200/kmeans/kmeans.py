"""
Masked K-means over the rows of a partially observed matrix -- B200 engine.
Drop-in for the reference class code/models/kmeans/kmeans.py (used by
init_FG='kmeans' of the tri-factorisation models): KMeans(X, M, K, resolve_empty),
initialise(seed), cluster(); attributes centroids, mask_centroids,
cluster_assignments, data_point_assignments, distances, clustering_results.

The distance / assignment pass (O(nnz * K)) and the centroid means run on the GPU
over the packed observed entries; the bookkeeping of the reference (change
detection, resolution of empty clusters by moving the point furthest from its
centroid, kmeans.py:131-158) is replayed on the host on length-I arrays.
"""
import ctypes as C
import random

import numpy

from .. import _lib
from .._lib import check, dptr

max_iterations = 200  # safeguard - if it takes more than this many iterations, stop


class KMeans:
    verbose = False     # the reference prints the iteration number and cluster sizes

    def __init__(self, X, M, K, resolve_empty='singleton', _engine=None, _side=0):
        self.X = numpy.copy(X)
        self.M = numpy.copy(M)
        self.K = K
        self.resolve_empty = resolve_empty

        assert len(self.X.shape) == 2, "Input matrix X is not a two-dimensional array, but instead %s-dimensional." % len(self.X.shape)
        assert self.X.shape == self.M.shape, "Input matrix X is not of the same size as the indicator matrix M: %s and %s respectively." % (self.X.shape, self.M.shape)
        assert self.K > 0, "K should be greater than 0."

        (self.no_points, self.no_coordinates) = self.X.shape
        self.no_unique_points = len(set(numpy.ascontiguousarray(row).tobytes() for row in self.X))

        if self.verbose and self.no_points < self.K:
            print("Want %s clusters but only have %s datapoints!" % (self.K, self.no_points))
        if self.verbose and self.no_unique_points < self.K:
            print("Want %s clusters but only have %s unique datapoints!" % (self.K, self.no_unique_points))

        # Assert none of the rows are entirely unknown values
        row_counts = (self.M != 0).sum(axis=1)
        empty = numpy.flatnonzero(row_counts == 0)
        assert empty.size == 0, "Fully unobserved row in X, row %s." % (empty[0] if empty.size else -1)

        # Columns can be entirely unknown - they just don't influence the clustering - but we need to remove them
        columns_to_remove = [int(j) for j in numpy.flatnonzero((self.M != 0).sum(axis=0) == 0)]
        if len(columns_to_remove) > 0:
            if self.verbose:
                print("WARNING: removed columns %s for K-means clustering as they have no observed datapoints." % columns_to_remove)
            self.X = numpy.delete(self.X, columns_to_remove, axis=1)
            self.M = numpy.delete(self.M, columns_to_remove, axis=1)
            self.no_coordinates -= len(columns_to_remove)
            _engine = None

        # Initialise the distances from data points to the assigned cluster centroids to zeros
        self.distances = numpy.zeros(self.no_points)
        self._eng, self._side = _engine, _side

    # -- device plumbing ---------------------------------------------------------
    def _engine(self):
        if self._eng is None:
            from .._engine import NMFEngine
            self._eng, self._side = NMFEngine(self.X, self.M, 1, distributed=False), 0
        return self._eng

    def _device_assign(self):
        eng = self._engine()
        cen = _lib.f64(numpy.array(self.centroids, dtype=float), (self.K, self.no_coordinates))
        msk = numpy.ascontiguousarray(numpy.array(self.mask_centroids) != 0, dtype=numpy.uint8)
        assign = numpy.empty(self.no_points, dtype=numpy.int32)
        dist = numpy.empty(self.no_points)
        check(eng.lib.bnmtf_kmeans_assign(eng.h, self._side, self.K, dptr(cen), msk.ctypes.data_as(C.POINTER(C.c_uint8)),
                                          assign.ctypes.data_as(C.POINTER(C.c_int32)), dptr(dist)))
        return assign, dist

    def _device_update(self, flags):
        eng = self._engine()
        cen = _lib.f64(numpy.array(self.centroids, dtype=float), (self.K, self.no_coordinates)).copy()
        msk = numpy.ascontiguousarray(numpy.array(self.mask_centroids) != 0, dtype=numpy.uint8).copy()
        assign = numpy.ascontiguousarray(self.cluster_assignments, dtype=numpy.int32)
        flg = numpy.ascontiguousarray(flags, dtype=numpy.int32)
        check(eng.lib.bnmtf_kmeans_update(eng.h, self._side, self.K, assign.ctypes.data_as(C.POINTER(C.c_int32)),
                                          flg.ctypes.data_as(C.POINTER(C.c_int32)), dptr(cen),
                                          msk.ctypes.data_as(C.POINTER(C.c_uint8))))
        self.centroids = [list(row) for row in cen]
        self.mask_centroids = msk.astype(float)

    """ Initialise the cluster centroids randomly """
    def initialise(self, seed=None):
        if seed is not None:
            random.seed(seed)

        # Compute the mins and maxes of the columns - i.e. the min and max of each dimension
        eng = self._engine()
        mins, maxs = numpy.empty(self.no_coordinates), numpy.empty(self.no_coordinates)
        check(eng.lib.bnmtf_kmeans_minmax(eng.h, self._side, dptr(mins), dptr(maxs)))
        self.mins, self.maxs = list(mins), list(maxs)

        # Randomly initialise the cluster centroids
        self.centroids = [self.random_cluster_centroid() for k in range(0, self.K)]
        self.cluster_assignments = numpy.array([-1 for d in range(0, self.no_points)])
        self.mask_centroids = numpy.ones((self.K, self.no_coordinates))

    # Randomly place a new cluster centroids, picking uniformly between the min and max of each coordinate
    def random_cluster_centroid(self):
        centroid = []
        for coordinate in range(0, self.no_coordinates):
            value = random.uniform(self.mins[coordinate], self.maxs[coordinate])
            centroid.append(value)
        return centroid

    """ Perform the clustering, until there is no change """
    def cluster(self):
        iteration = 1
        change = True
        while change:
            if self.verbose:
                print("Iteration: %s." % iteration)
            iteration += 1
            change = self.assignment()
            self.update()

            if iteration >= max_iterations:
                if self.verbose:
                    print("WARNING: did not converge, stopped after %s iterations." % max_iterations)
                break

        # At the end, we create a binary matrix indicating which points were assigned to which cluster
        self.create_matrix()

    """ Assign each data point to the closest cluster, and return whether any reassignments were made """
    def assignment(self):
        assign, dist = self._device_assign()
        change = bool(numpy.any(assign != self.cluster_assignments))
        self.cluster_assignments = assign.astype(int)
        self.distances = dist
        self.data_point_assignments = [list(numpy.flatnonzero(assign == c)) for c in range(0, self.K)]
        return change

    """ Update the centroids to the mean of the points assigned to it.
        If for a coordinate there are no known values, we set this cluster's mask to 0 there.
        If a cluster has no points assigned to it at all, we randomly re-initialise it. """
    def update(self):
        self._mean_flags = numpy.zeros(self.K, dtype=numpy.int32)
        for c in range(0, self.K):
            self.update_cluster(c)
        if self._mean_flags.any():
            self._device_update(self._mean_flags)

    # Update for one specific cluster (membership bookkeeping on the host, means on the device afterwards)
    def update_cluster(self, c):
        if len(self.data_point_assignments[c]) == 0:
            self._mean_flags[c] = 0
            # Reassign a datapoint to this cluster, as long as there are enough
            # unique datapoints. Either furthest away (singleton) or random.
            if self.no_unique_points >= self.K:
                if self.resolve_empty == 'singleton':
                    # Find the point currently furthest away from its centroid
                    index_furthest_away = self.find_point_furthest_away()
                    old_cluster = self.cluster_assignments[index_furthest_away]

                    # Add point to new cluster: its centroid is the point itself (= the mean of a single point)
                    self.distances[index_furthest_away] = 0.0
                    self.cluster_assignments[index_furthest_away] = c
                    self.data_point_assignments[c] = [index_furthest_away]
                    self._mean_flags[c] = 1

                    # Remove from old cluster and update
                    self.data_point_assignments[old_cluster].remove(index_furthest_away)
                    self.update_cluster(old_cluster)
                else:
                    # Randomly re-initialise this point
                    self.centroids[c] = self.random_cluster_centroid()
                    self.mask_centroids[c] = numpy.ones(self.no_coordinates)
        else:
            self._mean_flags[c] = 1

    # Find data point furthest away from its current cluster centroid
    def find_point_furthest_away(self):
        data_point_index = self.distances.argmax()
        return data_point_index

    # Create a binary matrix indicating the clustering (so size [no_points x K])
    def create_matrix(self):
        self.clustering_results = numpy.zeros((self.no_points, self.K))
        self.clustering_results[numpy.arange(self.no_points), self.cluster_assignments] = 1
        if self.verbose:
            print([sum(column) for column in self.clustering_results.T])
