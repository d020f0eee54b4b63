// Masked K-means over the rows of R (code/models/kmeans/kmeans.py), used by
// init_FG='kmeans' of the tri-factorisation classes (bnmtf_gibbs.py:140-151).
// The points are the rows of one side's row layout; the coordinates are its columns.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "kernels.cuh"

namespace bnmtf {

// assignment() / closest_cluster() / compute_MSE() (kmeans.py:88-125): one warp per point.
// cm = [Kc][ld] pairs (centroid value, centroid mask as 0/1).  The closest cluster is the
// lowest index among those with the smallest defined MSE; a cluster without overlap has
// no MSE (None) and is only kept while nothing better has been seen (kmeans.py:112-116).
static __global__ void kmeans_assign_kernel(const double* __restrict__ vals, const uint32_t* __restrict__ cols,
                                     const int64_t* __restrict__ rowstart, int n_loc, int64_t row0, uint32_t padcol,
                                     const double* __restrict__ cm, int64_t ld, int Kc, int* __restrict__ assign,
                                     double* __restrict__ dist) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n_loc) return;
  const int64_t beg = rowstart[row], end = rowstart[row + 1];
  int best = -1;
  double best_mse = 0.0;
  bool best_defined = false;
  for (int c = 0; c < Kc; ++c) {
    const double2* __restrict__ cc = reinterpret_cast<const double2*>(cm) + (size_t)c * ld;
    double num = 0.0, cnt = 0.0;
    for (int64_t q = beg + lane; q < end; q += 32) {
      const uint32_t j = cols[q];
      if (j == padcol) continue;
      const double2 v = __ldg(cc + j);
      const double d = vals[q] - v.x;
      num = fma(v.y * d, d, num);
      cnt += v.y;
    }
    num = warp_sum(num); cnt = warp_sum(cnt);
    const bool defined = cnt != 0.0;
    const double mse = defined ? num / cnt : 0.0;
    if (!best_defined || (defined && mse < best_mse)) { best = c; best_mse = mse; best_defined = defined; }
  }
  if (lane == 0) {
    assign[row0 + row] = best;
    dist[row0 + row] = best_defined ? best_mse : __longlong_as_double(0x7ff8000000000000ULL);
  }
}

// update_cluster() (kmeans.py:131-170): one warp per coordinate of the transposed layout (entries = points that
// observe the coordinate).  centroid[c][j] = mean of the observed values of the points assigned to c, mask 1;
// 0 and mask 0 if none.  Only clusters with flag[c] != 0 are written.
static __global__ void kmeans_update_kernel(const double* __restrict__ vals, const uint32_t* __restrict__ cols,
                                     const int64_t* __restrict__ rowstart, int n_loc, int64_t row0, uint32_t padcol,
                                     const int* __restrict__ assign, const int* __restrict__ flag, int Kc,
                                     double* __restrict__ cm, int64_t ld) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // coordinate
  const int lane = threadIdx.x & 31;
  if (row >= n_loc) return;
  const int64_t beg = rowstart[row], end = rowstart[row + 1];
  for (int c = 0; c < Kc; ++c) {
    if (!flag[c]) continue;
    double sum = 0.0, cnt = 0.0;
    for (int64_t q = beg + lane; q < end; q += 32) {
      const uint32_t i = cols[q];
      if (i == padcol) continue;
      if (assign[i] == c) { sum += vals[q]; cnt += 1.0; }
    }
    sum = warp_sum(sum); cnt = warp_sum(cnt);
    if (lane == 0)
      reinterpret_cast<double2*>(cm)[(size_t)c * ld + row0 + row] = cnt > 0.0 ? make_double2(sum / cnt, 1.0) : make_double2(0.0, 0.0);
  }
}

// per-coordinate minimum and maximum of the observed values (kmeans.py:50-52)
static __global__ void kmeans_minmax_kernel(const double* __restrict__ vals, const uint32_t* __restrict__ cols,
                                     const int64_t* __restrict__ rowstart, int n_loc, int64_t row0, uint32_t padcol,
                                     double* __restrict__ mins, double* __restrict__ maxs) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n_loc) return;
  double lo = __longlong_as_double(0x7ff0000000000000ULL), hi = -lo;
  for (int64_t q = rowstart[row] + lane; q < rowstart[row + 1]; q += 32) {
    if (cols[q] == padcol) continue;
    lo = fmin(lo, vals[q]); hi = fmax(hi, vals[q]);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) { mins[row0 + row] = lo; maxs[row0 + row] = hi; }
}

}  // namespace bnmtf
