// Kernels specific to the tri-factorisation R ~ F S G^T (code/models/bnmtf_gibbs.py,
// nmtf_icm.py).
//
// F and G sweeps are the NMF row sweeps with the "other factor" replaced by
// W = G S^T (tauF/muF, bnmtf_gibbs.py:268-275) and Z = F S (tauG/muG, :285-292).
//
// The S sweep (K*L sequential scalar updates, :200-203,277-283) is restated in
// contraction form.  With q_ij = vec(F_i (x) G_j) and the residual e_ij,
//     tauS(k,l) = tau * T[d,d],           d = (k,l)
//     muS(k,l)  = (-lambdaS_d + tau (c_d + S_d T[d,d])) / tauS
//     c_d = sum_Omega e_ij q_ij,d          T[d,d'] = sum_Omega q_ij,d q_ij,d'
// and an update S_d += delta changes e by -delta q_d, i.e. c -= delta T[:,d].
// T factorises over rows: T = sum_i (F_i F_i^T) (x) H_i with H_i = sum_{j in Omega_i} G_j G_j^T,
// so one pass over the observed entries (H_i, and P_il = sum_j e_ij G_jl) replaces the
// K*L passes of the reference; the K*L sequential updates then run on the small
// tensors in one CTA.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "kernels.cuh"

namespace bnmtf {

// out_km[c][i] = sum_q X_rm[i][q] * Smat[q*s_q + c*s_c]  (i < n), zero padded to ld
__global__ void combine_km_kernel(const double* __restrict__ X_rm, int64_t n, int kq, const double* __restrict__ Smat, int s_q,
                                  int s_c, double* __restrict__ out_km, int64_t ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (i >= ld) return;
  double acc = 0.0;
  if (i < n)
    for (int q = 0; q < kq; ++q) acc = fma(X_rm[i * kq + q], Smat[q * s_q + c * s_c], acc);
  out_km[(size_t)c * ld + i] = acc;
}

__device__ __forceinline__ void pair_decode(int p, int L, int& a, int& b) {
  a = 0;
  int rem = p;
  while (rem >= L - a) { rem -= L - a; ++a; }
  b = a + rem;
}
__host__ __device__ __forceinline__ int pair_index(int a, int b, int L) {   // a <= b
  return a * L - (a * (a - 1)) / 2 + (b - a);
}

// H_i[pair(l,l')] = sum_{j in Omega_i} G_jl G_jl'   -- one warp per row of the row layout
template <int PMAX>
__global__ void nmtf_gram_kernel(const uint32_t* __restrict__ cols, const int64_t* __restrict__ rowstart, int n_loc, int64_t row0,
                                 uint32_t padcol, const double* __restrict__ G_rm, int L, double* __restrict__ Hrow, int NP) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n_loc) return;
  int pa[PMAX], pb[PMAX];
  double acc[PMAX];
#pragma unroll
  for (int t = 0; t < PMAX; ++t) {
    const int p = lane + 32 * t;
    pa[t] = 0; pb[t] = 0; acc[t] = 0.0;
    if (p < NP) pair_decode(p, L, pa[t], pb[t]);
  }
  const int64_t beg = rowstart[row], end = rowstart[row + 1];
  for (int64_t q0 = beg; q0 < end; q0 += 32) {
    const uint32_t mycol = (q0 + lane < end) ? cols[q0 + lane] : padcol;
    for (int s = 0; s < 32; ++s) {
      const uint32_t j = __shfl_sync(0xffffffffu, mycol, s);
      if (j == padcol) continue;
      const double* __restrict__ g = G_rm + (size_t)j * L;
#pragma unroll
      for (int t = 0; t < PMAX; ++t) acc[t] = fma(__ldg(g + pa[t]), __ldg(g + pb[t]), acc[t]);
    }
  }
#pragma unroll
  for (int t = 0; t < PMAX; ++t) {
    const int p = lane + 32 * t;
    if (p < NP) Hrow[(size_t)(row0 + row) * NP + p] = acc[t];
  }
}

// c[k][l] = sum_i F[i][k] P[i][l]    (one block per (k,l), fixed-order tree)
__global__ void nmtf_reduce_c_kernel(const double* __restrict__ F, const double* __restrict__ P, int64_t I, int K, int L,
                                     double* __restrict__ c) {
  __shared__ double sh[32];
  const int k = blockIdx.x / L, l = blockIdx.x % L;
  double a = 0.0;
  for (int64_t i = threadIdx.x; i < I; i += blockDim.x) a = fma(F[i * K + k], P[i * L + l], a);
  a = block_sum(a, sh);
  if (threadIdx.x == 0) c[blockIdx.x] = a;
}

// TT[pk][pl] = sum_i F[i][ka] F[i][kb] H[i][pl]; block = 32 (pl) x 8 (row slices), fixed order
__global__ void nmtf_reduce_T_kernel(const double* __restrict__ F, const double* __restrict__ H, int64_t I, int K, int NPK, int NPL,
                                     double* __restrict__ TT) {
  __shared__ double sh[8][33];
  const int pk = blockIdx.y;
  const int pl = blockIdx.x * 32 + threadIdx.x;
  int ka, kb;
  pair_decode(pk, K, ka, kb);
  double a = 0.0;
  if (pl < NPL)
    for (int64_t i = threadIdx.y; i < I; i += 8) a = fma(F[i * K + ka] * F[i * K + kb], H[i * NPL + pl], a);
  sh[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && pl < NPL) {
    double t = 0.0;
    for (int z = 0; z < 8; ++z) t += sh[z][threadIdx.x];
    TT[(size_t)pk * NPL + pl] = t;
  }
}

// The K*L sequential S updates on the small tensors (single CTA).
//   Gibbs: S_kl ~ TN(muS, tauS)  bnmtf_gibbs.py:200-203      ICM: max(max(0,mu),0.1)  nmtf_icm.py:203-207
// single_d >= 0: only report (tau, mu) of entry d from the current state.
template <int MODE>
__global__ void nmtf_s_sweep_kernel(const double* __restrict__ c_in, const double* __restrict__ TT, int K, int L, double* S,
                                    const double* __restrict__ lamS, const double* __restrict__ scal, uint64_t seed,
                                    uint32_t iteration, double* mu_out, double* tau_out, int single_d) {
  extern __shared__ double shs[];
  const int KL = K * L, NPL = L * (L + 1) / 2;
  double* c = shs;            // [KL]
  double* sv = c + KL;        // [KL]
  double* bc = sv + KL;       // [1]
  for (int d = threadIdx.x; d < KL; d += blockDim.x) { c[d] = c_in[d]; sv[d] = S[d]; }
  __syncthreads();
  const double tau = scal[0];
  const int d_begin = single_d >= 0 ? single_d : 0, d_end = single_d >= 0 ? single_d + 1 : KL;
  for (int d = d_begin; d < d_end; ++d) {
    const int k = d / L, l = d - k * L;
    if (threadIdx.x == 0) {
      const double Tdd = TT[(size_t)pair_index(k, k, K) * NPL + pair_index(l, l, L)];
      const double ct = tau * Tdd;                                        // bnmtf_gibbs.py:279
      const double cm = (-lamS[d] + tau * (c[d] + sv[d] * Tdd)) / ct;     // bnmtf_gibbs.py:283
      double x_new;
      if (MODE == M_GIBBS) {
        Rng rng(seed, RNG_S, (uint64_t)d, 0u, iteration);
        x_new = tn_draw(cm, ct, rng);
      } else {
        x_new = fmax(fmax(0.0, cm), MINIMUM_TN);
      }
      if (mu_out) { mu_out[d] = cm; tau_out[d] = ct; }
      if (single_d >= 0) x_new = sv[d];
      bc[0] = x_new - sv[d];
      sv[d] = x_new;
    }
    __syncthreads();
    const double delta = bc[0];
    if (delta != 0.0) {
      for (int d2 = threadIdx.x; d2 < KL; d2 += blockDim.x) {
        const int k2 = d2 / L, l2 = d2 - k2 * L;
        const int pk = pair_index(min(k, k2), max(k, k2), K), pl = pair_index(min(l, l2), max(l, l2), L);
        c[d2] = fma(-delta, TT[(size_t)pk * NPL + pl], c[d2]);
      }
    }
    __syncthreads();
  }
  if (single_d < 0)
    for (int d = threadIdx.x; d < KL; d += blockDim.x) S[d] = sv[d];
}

// lambdaF_k ~ Gamma(alpha0 + I, beta0 + sum_i F_ik), lambdaG_l likewise with J
// (bnmtf_gibbs.py:187-191,252-266; ICM: modes, nmtf_icm.py:188-192).  One block per column.
template <int MODE>
__global__ void nmtf_lambda_kernel(const double* __restrict__ F, int64_t I, int K, const double* __restrict__ G, int64_t J, int L,
                                   Hyper h, double* colsum, double* lamF, double* lamG, uint64_t seed, uint32_t iteration,
                                   int update) {
  __shared__ double sh[32];
  const int c = blockIdx.x;
  const bool isF = c < K;
  const double* X = isF ? F : G;
  const int64_t n = isF ? I : J;
  const int kf = isF ? K : L, col = isF ? c : c - K;
  double a = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) a += X[i * kf + col];
  a = block_sum(a, sh);
  if (threadIdx.x == 0) {
    colsum[c] = a;
    if (h.ARD && update) {
      const double alpha = h.alpha0 + (double)n, beta = h.beta0 + a;
      double lam;
      if (MODE == M_GIBBS) {
        Rng rng(seed, RNG_LAMBDA, (uint64_t)c, 0u, iteration);
        lam = gamma_draw(alpha, beta, rng);
      } else {
        lam = (alpha - 1.0) / beta;
      }
      (isF ? lamF : lamG)[col] = lam;
    }
  }
}

__global__ void init_vector_kernel(double* __restrict__ X, int64_t n, const double* __restrict__ lam, int random, uint64_t seed,
                                   uint32_t purpose) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  if (random) {
    Rng rng(seed, purpose, (uint64_t)idx, 0u, 0u);
    X[idx] = exp_draw(lam[idx], rng);
  } else {
    X[idx] = 1.0 / lam[idx];
  }
}

}  // namespace bnmtf
