// Row sweep for rows too long for registers (more than 16384 observed entries and not eligible for the cluster
// kernel): the residual of the row lives in a global scratch array, one CTA of 512 threads walks the row once per
// column -- the correction of column k-1 is folded into the pass of column k, so every column costs one read and
// one write of the residual plus two gathers per entry.  Same conditionals, same update code (conditional_update,
// tn_draw_fast) and the same per-row statistics as sweep_body; only the fixed summation order differs (512 threads
// striding the row).  Serves single_k = -1 (sweep, with the optional projection epilogue and the BNMTF VB covariance term
// of the tri-factorisation), -2 (statistics only), -3 / -4 (projections / masked sums) and >= 0 (parameters of one column).
#pragma once
#include "kernels.cuh"

namespace bnmtf {

constexpr int LONG_T = 512;

// fixed-order block sum of NR values; result valid on thread 0
template <int NR>
__device__ __forceinline__ void long_reduce(double (&v)[NR], double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < NR; ++r) v[r] = warp_sum(v[r]);
  __syncthreads();               // red[] free again
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < NR; ++r) red[w * NR + r] = v[r];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < LONG_T / 32; ++q) {
#pragma unroll
      for (int r = 0; r < NR; ++r) v[r] += red[q * NR + r];
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(LONG_T) sweep_long_kernel(const SweepParams p, double* __restrict__ escr) {
  constexpr int NV = (MODE == M_VB) ? 2 : 1;
  extern __shared__ double smem[];
  const int K = p.K, t = threadIdx.x, lane = t & 31;
  double* xs = smem;            // [K] current values of the row
  double* xvar = xs + K;        // [K]
  double* xmu = xvar + K;       // [K]
  double* xtau = xmu + K;       // [K]
  double* red = xtau + K;       // [3 * 16]
  double* bc = red + 3 * (LONG_T / 32);   // [2]
  double* fsv = bc + 2;         // [covL] (X S)_il of the row   (BNMTF VB)
  double* arow = fsv + p.covL;  // [covL] A_il of the row
  double* mat = arow + p.covL;  // [K][TN_PRE_DOUBLES] (Gibbs)
  const bool fast_draw = (MODE == M_GIBBS) && p.single_k == -1;
  const bool has_cov = (MODE == M_VB) && (p.covA != nullptr);
  const bool want_proj = p.single_k == -1 && p.Ykm2 != nullptr && MODE != M_NP;   // epilogue: project the final residual
  const uint32_t iteration = p.iter_ptr ? *p.iter_ptr : p.iteration;
  const bool dry = (p.single_k == -2);
  const double sgn = (MODE == M_NP) ? 1.0 : -1.0;   // NP keeps the prediction, the others the residual

  for (int row = blockIdx.x; row < p.n_loc; row += gridDim.x) {
    const int64_t grow = p.row0 + row;
    const int64_t base = p.rowstart[row];
    const int64_t len = p.rowstart[row + 1] - base;   // multiple of LONG_T
    __syncthreads();
    if (fast_draw)
      for (int kk = t; kk < K; kk += LONG_T)
        tn_pre_material_store(p.seed, p.purpose, (uint64_t)grow, (uint32_t)kk, iteration, mat + TN_PRE_DOUBLES * kk);
    for (int kk = t; kk < K; kk += LONG_T) {
      xs[kk] = p.Xval[grow * K + kk];
      if (MODE == M_VB) { xvar[kk] = p.Xvar[grow * K + kk]; xmu[kk] = p.Xmu[grow * K + kk]; xtau[kk] = p.Xtau[grow * K + kk]; }
    }
    __syncthreads();
    if (has_cov) {
      for (int l = t; l < p.covL; l += LONG_T) {
        double a = 0.0;
        for (int kk = 0; kk < K; ++kk) a = fma(xs[kk], p.covS[kk * p.cov_sk + l * p.cov_sl], a);
        fsv[l] = a;
        arow[l] = p.covA[grow * p.covL + l];
      }
      __syncthreads();
    }

    // ---- residual of the row from the current factors
    if (p.single_k != -4) {
      for (int64_t q = t; q < len; q += LONG_T) {
        const uint32_t c = p.cols[base + q] * NV;
        double e = (MODE == M_NP) ? 0.0 : p.vals[base + q];
        for (int k = 0; k < K; ++k) e = fma(sgn * xs[k], __ldg(p.Ykm + (size_t)k * p.ldy * NV + c), e);
        escr[base + q] = e;
      }
    }

    if (p.single_k == -3 || p.single_k == -4) {
      // projections of the residual onto the columns of a second factor and, in VB mode, the masked sums of its second
      // moments and variances; NP: the row parts of nmtf_np.update_S (as sweep_body, kernels.cuh)
      for (int l = 0; l < p.K2; ++l) {
        const double* __restrict__ y2 = p.Ykm2 + (size_t)l * p.ldy2 * NV;
        double s3[3] = {0.0, 0.0, 0.0};
        for (int64_t q = t; q < len; q += LONG_T) {
          const uint32_t c = p.cols[base + q] * NV;
          if (MODE == M_VB) {
            const double2 yy = __ldg(reinterpret_cast<const double2*>(y2 + c));
            s3[0] = fma(yy.x, yy.x, s3[0]); s3[1] += yy.y;
            if (p.single_k == -3) s3[2] = fma(escr[base + q], yy.x, s3[2]);
          } else if (MODE == M_NP) {
            const double g = __ldg(y2 + c);
            const double ratio = (c != p.padcol) ? p.vals[base + q] / escr[base + q] : 0.0;
            s3[2] = fma(ratio, g, s3[2]); s3[1] += g;
          } else {
            s3[2] = fma(escr[base + q], __ldg(y2 + c), s3[2]);
          }
        }
        long_reduce<3>(s3, red);
        if (t == 0) {
          if (p.out_mu) p.out_mu[grow * p.K2 + l] = s3[2];
          if (p.out_tau) p.out_tau[grow * p.K2 + l] = s3[1];
          if (p.out3) p.out3[grow * p.K2 + l] = s3[1] - s3[0];
        }
      }
      continue;
    }

    double acc_esd = 0.0, acc_q = 0.0, acc_prior = 0.0;   // thread 0 only
    const int k_begin = (p.single_k >= 0) ? p.single_k : 0;
    const int k_end = (p.single_k >= 0) ? p.single_k + 1 : K;
    const bool skip_loop = dry && MODE != M_VB;
    double delta_prev = 0.0;
    int k_prev = -1;
    for (int k = k_begin; k < k_end && !skip_loop; ++k) {
      const double* __restrict__ y = p.Ykm + (size_t)k * p.ldy * NV;
      const double* __restrict__ yp = p.Ykm + (size_t)(k_prev < 0 ? 0 : k_prev) * p.ldy * NV;
      const bool fix = (k_prev >= 0) && (delta_prev != 0.0);
      double s[3] = {0.0, 0.0, 0.0};
      for (int64_t q = t; q < len; q += LONG_T) {
        const uint32_t c = p.cols[base + q] * NV;
        double e = escr[base + q];
        if (fix) { e = fma(sgn * delta_prev, __ldg(yp + c), e); escr[base + q] = e; }
        double vn;
        if (MODE == M_VB) {
          const double2 yy = __ldg(reinterpret_cast<const double2*>(y + c));
          vn = yy.x; s[1] += yy.y;
        } else {
          vn = __ldg(y + c);
        }
        if (MODE == M_NP) {
          const double ratio = (c != p.padcol) ? p.vals[base + q] / e : 0.0;
          s[0] = fma(vn, ratio, s[0]); s[1] += vn;
        } else {
          s[0] = fma(vn, vn, s[0]); s[2] = fma(e, vn, s[2]);
        }
      }
      long_reduce<3>(s, red);
      if (t == 0) {
        const double x_old = xs[k];
        const double lam = p.lamk ? p.lamk[k] : (p.lam_rm ? p.lam_rm[grow * K + k] : 0.0);
        const double tau = p.scal[0];
        const UpdateCtx uc{p.seed, p.purpose, iteration, dry};
        double cm, ct, vn = 0.0, x_new;
        if (has_cov) {   // diff_term - cov_term (bnmtf_vb.py:342-348)
          double cov = 0.0;
          for (int l = 0; l < p.covL; ++l) {
            const double skl = p.covS[k * p.cov_sk + l * p.cov_sl];
            cov = fma(skl * arow[l], fsv[l] - x_old * skl, cov);
          }
          s[2] -= cov;
        }
        if (fast_draw) {
          Rng rng(uc.seed, uc.purpose, (uint64_t)grow, (uint32_t)k, uc.iteration);
          rng.draw = 2;
          const TnPre m = tn_pre_load(mat + TN_PRE_DOUBLES * k);
          ct = tau * s[0];
          const double num = -lam + tau * (s[2] + x_old * s[0]);
          x_new = tn_draw_fast(num, ct, m, rng, cm);
        } else {
          x_new = conditional_update<MODE, false>(s, x_old, lam, tau, uc, (uint64_t)grow, k, xmu[k], xtau[k], xvar[k], ct, cm, vn,
                                                  acc_esd, acc_q, acc_prior);
        }
        if (MODE == M_VB) xvar[k] = vn;
        if (p.single_k >= 0) { p.out_tau[grow] = ct; p.out_mu[grow] = cm; x_new = x_old; }
        xmu[k] = cm; xtau[k] = ct;
        bc[0] = x_new - x_old;
        xs[k] = x_new;
        if (has_cov && x_new != x_old)
          for (int l = 0; l < p.covL; ++l) fsv[l] = fma(x_new - x_old, p.covS[k * p.cov_sk + l * p.cov_sl], fsv[l]);
      }
      __syncthreads();
      delta_prev = bc[0];
      k_prev = k;
    }

    if (want_proj && k_prev >= 0 && delta_prev != 0.0) {   // the projection reads the final residual: last correction in place
      const double* __restrict__ yp = p.Ykm + (size_t)k_prev * p.ldy * NV;
      for (int64_t q = t; q < len; q += LONG_T) {
        const uint32_t c = p.cols[base + q] * NV;
        escr[base + q] = fma(sgn * delta_prev, __ldg(yp + c), escr[base + q]);
      }
      delta_prev = 0.0;
    }

    // ---- last correction and the per-row statistics from the final residual
    if (p.rowstats != nullptr && p.single_k < 0) {
      if (MODE == M_VB && t < 32) acc_q = vb_row_q(xs, xvar, xmu, xtau, K, lane);
      const bool fix = (k_prev >= 0) && (delta_prev != 0.0);
      const double* __restrict__ yp = p.Ykm + (size_t)(k_prev < 0 ? 0 : k_prev) * p.ldy * NV;
      double st[3] = {0.0, 0.0, 0.0};
      double idiv[1] = {0.0};
      for (int64_t q = t; q < len; q += LONG_T) {
        const uint32_t c = p.cols[base + q] * NV;
        double e = escr[base + q];
        if (fix) e = fma(sgn * delta_prev, __ldg(yp + c), e);
        if (c != p.padcol * NV) {
          const double r = p.vals[base + q];
          const double ee = (MODE == M_NP) ? (r - e) : e;
          st[0] = fma(ee, ee, st[0]); st[1] += ee; st[2] = fma(r - p.rmean, ee, st[2]);
          if (MODE == M_NP) idiv[0] += r * log(r / e) - r + e;
        }
      }
      long_reduce<3>(st, red);
      if (MODE == M_NP) long_reduce<1>(idiv, red);
      if (t == 0) {
        double* rs = p.rowstats + (size_t)row * NRS;
        rs[RS_RSS] = st[0]; rs[RS_SUM_E] = st[1]; rs[RS_SUM_RCE] = st[2];
        rs[RS_ESDVAR] = acc_esd; rs[RS_ELBOQ] = acc_q; rs[RS_PRIOR] = acc_prior;
        rs[RS_IDIV] = idiv[0]; rs[RS_SPARE] = 0.0;
      }
    }
    if (want_proj) {
      // ---- projections of the residual the row ends with: out_mu[row][l] = sum_j e_j Y2[l][j] (P_il of the S statistics)
      for (int l = 0; l < p.K2; ++l) {
        const double* __restrict__ y2 = p.Ykm2 + (size_t)l * p.ldy2 * NV;
        double s1[1] = {0.0};
        for (int64_t q = t; q < len; q += LONG_T) s1[0] = fma(escr[base + q], __ldg(y2 + p.cols[base + q] * NV), s1[0]);
        long_reduce<1>(s1, red);
        if (t == 0) p.out_mu[grow * p.K2 + l] = s1[0];
      }
    }
    __syncthreads();
    if (p.single_k == -1) {
      for (int kk = t; kk < K; kk += LONG_T) {
        p.Xval[grow * K + kk] = xs[kk];
        if (MODE == M_VB) p.Xvar[grow * K + kk] = xvar[kk];
        if (p.Xmu != nullptr) { p.Xmu[grow * K + kk] = xmu[kk]; p.Xtau[grow * K + kk] = xtau[kk]; }
      }
    }
  }
}

}  // namespace bnmtf
