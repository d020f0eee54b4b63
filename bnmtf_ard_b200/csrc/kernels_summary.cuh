// On-device posterior summaries: running sums of the Gibbs/ICM draws over
// burn_in::thinning (approx_expectation, bnmf_gibbs.py:222-230, bnmtf_gibbs.py:302-315) and the
// held-out metrics of predict() (compute_MSE / compute_R2 / compute_Rp, bnmf_gibbs.py:252-270)
// evaluated on a list of (row, column, value) test entries.
#pragma once
#include <cstdint>

struct SumSeg { double* dst; const double* src; int64_t n; };
struct SumArgs { SumSeg seg[6]; int nseg; };

// dst += src for every segment; summation order over iterations is the launch order, i.e. the
// order of numpy's axis-0 sum in the reference.
static __global__ void summary_accum_kernel(SumArgs a) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int s = 0; s < a.nseg; ++s) {
    double* __restrict__ d = a.seg[s].dst;
    const double* __restrict__ x = a.seg[s].src;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.seg[s].n; i += stride) d[i] += x[i];
  }
}

static __global__ void summary_mean_kernel(double* __restrict__ dst, const double* __restrict__ sum, int64_t n, double count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = sum[i] / count;
}

constexpr int PRED_THREADS = 256;

template <int NQ>
__device__ __forceinline__ void pred_block_reduce(double (&v)[NQ], double* __restrict__ out) {
  __shared__ double sh[NQ][PRED_THREADS / 32];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) sh[q][w] = v[q];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      double s = 0;
      for (int i = 0; i < PRED_THREADS / 32; ++i) s += sh[q][i];
      out[(size_t)blockIdx.x * NQ + q] = s;
    }
  }
}

// pass 1: prediction of every test entry (A_i . B_j, or A_i S B_j^T when S != nullptr) and the
// per-block sums of r and p.  A: [nA][ka], B: [nB][kb] row-major.
static __global__ void __launch_bounds__(PRED_THREADS) predict_pass1_kernel(
    const double* __restrict__ A, int ka, const double* __restrict__ B, int kb, const double* __restrict__ S,
    const int32_t* __restrict__ rows, const int32_t* __restrict__ cols, const double* __restrict__ vals, int64_t n,
    double* __restrict__ pred, double* __restrict__ partial) {
  double acc[2] = {0.0, 0.0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
    const double* a = A + (size_t)rows[e] * ka;
    const double* b = B + (size_t)cols[e] * kb;
    double p = 0.0;
    if (S == nullptr) {
      for (int k = 0; k < ka; ++k) p += a[k] * b[k];
    } else {
      for (int k = 0; k < ka; ++k) {
        double t = 0.0;
        for (int l = 0; l < kb; ++l) t += S[k * kb + l] * b[l];
        p += a[k] * t;
      }
    }
    pred[e] = p;
    acc[0] += vals[e];
    acc[1] += p;
  }
  pred_block_reduce<2>(acc, partial);
}

// pass 2: centred sums; means[0] = mean of r, means[1] = mean of p
static __global__ void __launch_bounds__(PRED_THREADS) predict_pass2_kernel(
    const double* __restrict__ vals, const double* __restrict__ pred, int64_t n, const double* __restrict__ means,
    double* __restrict__ partial) {
  const double mr = means[0], mp = means[1];
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
    const double r = vals[e], p = pred[e];
    acc[0] += (r - p) * (r - p);
    acc[1] += (r - mr) * (r - mr);
    acc[2] += (p - mp) * (p - mp);
    acc[3] += (r - mr) * (p - mp);
  }
  pred_block_reduce<4>(acc, partial);
}

// fixed-order fold of the per-block partials.  stage 0: out[0..1] = means; stage 1: the metrics
// out[8] = {n, MSE, R^2, Rp, SS_res, SS_tot, mean_real, mean_pred}
static __global__ void predict_fold_kernel(const double* __restrict__ partial, int nblocks, int nq, int stage, double n,
                                    double* __restrict__ means, double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s[4] = {0, 0, 0, 0};
  for (int b = 0; b < nblocks; ++b)
    for (int q = 0; q < nq; ++q) s[q] += partial[(size_t)b * nq + q];
  if (stage == 0) {
    means[0] = s[0] / n;
    means[1] = s[1] / n;
  } else {
    out[0] = n;
    out[1] = s[0] / n;
    out[2] = s[1] != 0.0 ? 1.0 - s[0] / s[1] : INFINITY;
    out[3] = s[3] / (sqrt(s[1]) * sqrt(s[2]));
    out[4] = s[0];
    out[5] = s[1];
    out[6] = means[0];
    out[7] = means[1];
  }
}
