// Device kernels of the BNMF / NMF hot path (sm_100a).
//
// Formulation: the K column updates of one factor are
// conditionally independent across rows (bnmf_gibbs.py:155-158: U[:,k] for all
// rows given V; rows never interact), so a *row team* keeps the residual of one
// row of R -- observed entries only -- in registers and runs all K column
// updates before touching memory again.  The residual is recomputed from R at
// the start of each sweep, so nothing but the factors is carried between
// sweeps.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "rng.cuh"

namespace bnmtf {

enum { M_GIBBS = 0, M_ICM = 1, M_VB = 2, M_NP = 3 };

// per-row statistics written by a sweep (deterministic: one writer per row)
enum { RS_RSS = 0, RS_SUM_E = 1, RS_SUM_RCE = 2, RS_ESDVAR = 3, RS_ELBOQ = 4, RS_PRIOR = 5, RS_IDIV = 6, RS_SPARE = 7, NRS = 8 };

constexpr double MINIMUM_TN = 0.1;  // nmf_icm.py:59

struct SweepParams {
  // packed observed entries of this side: rows padded to multiples of tpr
  const double* __restrict__ vals;
  const uint32_t* __restrict__ cols;
  const int64_t* __restrict__ rowstart;  // [n_loc+1], slot offsets
  int n_loc;                             // local rows
  int64_t row0;                          // global index of local row 0
  int K;
  int tpr;                               // threads per row team (multiple of 32)
  uint32_t padcol;                       // column index that reads 0 from Ykm
  const double* __restrict__ Ykm;        // other factor, k-major [K][ldy][NV]
  int64_t ldy;
  double* Xval;                          // own factor, row-major [n_glob_pad][K]
  double* Xvar;                          // VB
  double* Xmu;                           // VB state / optional parity output
  double* Xtau;                          // VB state / optional parity output
  const double* __restrict__ lam_rm;     // [n_glob][K] or nullptr
  const double* __restrict__ lamk;       // [K] (ARD) or nullptr
  const double* __restrict__ scal;       // scal[0] = tau (exp_tau)
  double* out_tau;                       // single_k >= 0: [n_glob] outputs
  double* out_mu;
  double* rowstats;                      // [n_loc][NRS] or nullptr
  double rmean;
  uint64_t seed;
  uint32_t iteration;
  const uint32_t* iter_ptr;              // when set (graph replays) the iteration index is read from the device
  uint32_t purpose;
  // Small models are launch-bound: a sweep (single_k == -1) can also leave the k-major gather copy of the factor it
  // updates (what rm_to_km_kernel would write behind it: [K][km_ld][NV], NV == 2 for VB) and its slot of the draw traces
  // (what trace_copy_kernel would copy), so the iteration needs neither of the two launches.
  double* km_out;                        // or nullptr
  int64_t km_ld;
  double* trace_out;                     // [iterations][n_glob][K] staging buffer of this factor, or nullptr
  int64_t trace_n;                       // n_glob (rows of the factor)
  int single_k;                          // >=0 one column, no state change; -1 sweep; -2 dry (stats only);
                                         // -3 project: out_mu[row][l] = sum_j e_j Y2[l][j], l < K2 (NMTF S statistics)
                                         // -4 masked sums only (no residual): out_tau = sum y2, out3 = sum y2 - sum y^2
  const double* __restrict__ Ykm2;       // second gather source [K2][ldy2][NV] (project modes)
  int64_t ldy2;
  int K2;
  double* out3;
  // BNMTF VB covariance term of update_F / update_G (bnmtf_vb.py:343,371):
  //   cov_i = sum_l S(k,l) A_il ((X S)_il - x_ik S(k,l)),  A = masked sums of the other factor's variance
  const double* __restrict__ covA;       // [n_glob][covL] or nullptr
  const double* __restrict__ covS;       // S(k,l) = covS[k*cov_sk + l*cov_sl]
  int covL, cov_sk, cov_sl;
};

__device__ __forceinline__ void bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Two warp sums for the price of one: afterwards lanes 0..15 hold the total of a, lanes 16..31 the total of b.
// The first exchange hands each half the value it keeps; the remaining rounds are those of warp_sum on a single value,
// pairing the same partial sums in the same order, so the totals are bitwise warp_sum(a) and warp_sum(b).
// (Shuffles share the shared-memory pipe with the gathers of the sweeps: 10 SHFL instead of 20.)
__device__ __forceinline__ double warp_sum_pair(double a, double b, int lane) {
  const bool up = (lane & 16) != 0;
  double v = (up ? b : a) + __shfl_xor_sync(0xffffffffu, up ? a : b, 16);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Team-wide sum of NR values, result valid on the team leader warp (all of its
// lanes).  Fixed order => bitwise reproducible.  Non-leader warps only arrive.
template <int NR>
__device__ __forceinline__ void team_reduce_to_leader(double (&v)[NR], double* red, int nw, int wteam, int lane,
                                                      int barA, int tpr) {
#pragma unroll
  for (int r = 0; r < NR; ++r) v[r] = warp_sum(v[r]);
  if (nw == 1) return;
  if (wteam != 0) {
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < NR; ++r) red[wteam * NR + r] = v[r];
    }
    __syncwarp();
    bar_arrive(barA, tpr);
  } else {
    bar_sync(barA, tpr);
    for (int w = 1; w < nw; ++w) {
#pragma unroll
      for (int r = 0; r < NR; ++r) v[r] += red[w * NR + r];
    }
  }
}

// The update of one factor entry from the reduced sums of its row:
//   s0 = sum y^2 (NP: numerator), s1 = sum (var_y + y^2) (NP: denominator), s2 = sum e*y.
// Returns the conditional precision / mean actually used (ct, cm), the new value
// and (VB) the new variance, and accumulates the VB ELBO/ESD row terms.
//   Gibbs: bnmf_gibbs.py:203-210 + TN_vector_draw     ICM: nmf_icm.py:159-160
//   VB   : bnmf_vb.py:251-255,275-278                  NP : nmf_np.py:120-122
struct UpdateCtx {
  uint64_t seed;
  uint32_t purpose, iteration;
  bool dry;
};

// Gibbs update with the Philox stream opened (and its first uniform drawn) before the
// row sums are known; see tn_draw_pre.
__device__ __forceinline__ double gibbs_update_pre(const double (&s)[3], double x_old, double lam, double tau, Rng& rng,
                                                   double u1, double& ct, double& cm) {
  ct = tau * s[0];                                        // bnmtf_gibbs tauU, bnmf_gibbs.py:205
  const double num = -lam + tau * (s[2] + x_old * s[0]);  // numerator of muU, bnmf_gibbs.py:210
  return tn_draw_pre(num, ct, u1, rng, cm);
}
// WITH_Q = false: the caller evaluates the q(U) terms of the ELBO itself (vb_row_q)
template <int MODE, bool WITH_Q = true>
__device__ __forceinline__ double conditional_update(const double (&s)[3], double x_old, double lam, double tau,
                                                     const UpdateCtx& u, uint64_t grow, int k, double old_mu,
                                                     double old_tau, double old_var, double& ct, double& cm,
                                                     double& v_new, double& acc_esd, double& acc_q, double& acc_prior) {
  double x_new;
  v_new = old_var;
  if (MODE == M_NP) {
    x_new = x_old * s[0] / s[1];  // nmf_np.py:122
    ct = s[1]; cm = x_new;
  } else if (MODE == M_VB) {
    ct = tau * s[1];                                   // bnmf_vb.py:254
    cm = (-lam + tau * (s[2] + x_old * s[0])) / ct;    // bnmf_vb.py:255
    double en, vn;
    if (u.dry) { cm = old_mu; ct = old_tau; en = x_old; vn = old_var; }
    else tn_moments(cm, ct, en, vn);
    x_new = en;
    v_new = vn;
    acc_esd += (vn + en * en) * s[1] - (en * en) * s[0];
    if (WITH_Q)
      acc_q += -0.5 * log(ct) + log_half_erfc_ref(-cm * sqrt(ct) / 1.4142135623730951)
               + 0.5 * ct * (vn + (en - cm) * (en - cm));
    acc_prior += lam * en;
  } else {
    ct = tau * s[0];                                   // bnmf_gibbs.py:205
    cm = (-lam + tau * (s[2] + x_old * s[0])) / ct;    // bnmf_gibbs.py:210
    if (MODE == M_GIBBS) {
      Rng rng(u.seed, u.purpose, grow, (uint32_t)k, u.iteration);
      const double u1 = rng.u01();
      x_new = gibbs_update_pre(s, x_old, lam, tau, rng, u1, ct, cm);   // bnmf_gibbs.py:158
    } else {
      x_new = fmax(fmax(0.0, cm), MINIMUM_TN);         // nmf_icm.py:159-160
    }
  }
  return x_new;
}

// q(U_i) part of the ELBO for one row (bnmf_vb.py:219-229): sum_k -0.5 log tau_ik + log(0.5 erfc(-mu_ik sqrt(tau_ik)/sqrt2))
// + 0.5 tau_ik (var_ik + (exp_ik - mu_ik)^2).  Each column is updated once per sweep, so the final parameters of
// the row are the ones every update ended with; one lane per column keeps the two logarithms and the erfc off the
// leader's serial path.  Called by the leader warp of the team; fixed summation order.
__device__ __forceinline__ double vb_row_q(const double* xs, const double* xvar, const double* xmu, const double* xtau,
                                           int K, int lane) {
  double qq = 0.0;
  for (int kk = lane; kk < K; kk += 32) {
    const double ct = xtau[kk], cm = xmu[kk], en = xs[kk], vn = xvar[kk];
    qq += -0.5 * log(ct) + log_half_erfc_ref(-cm * sqrt(ct) / 1.4142135623730951) + 0.5 * ct * (vn + (en - cm) * (en - cm));
  }
  return warp_sum(qq);
}

// -----------------------------------------------------------------------------
// Row-fused K-column sweep.  One team of `tpr` threads per row; thread t owns the
// entries t, t+tpr, ... of the row (coalesced loads), NPT of them at most.
//   Gibbs : bnmf_gibbs.py:155-164 (tauU/muU :203-210, TN_vector_draw)
//   ICM   : nmf_icm.py:155-167
//   VB    : bnmf_vb.py:167-174 (update_U :251-255, update_exp_U :275-278)
//   NP    : nmf_np.py:107-126
// -----------------------------------------------------------------------------
// HOLDV: keep the gathered column of the other factor in registers between the
// accumulation and the rank-1 correction (2 gathers per entry and column instead
// of 3); switched off for very long rows where the registers are needed for e[].
// second moment Var + E^2 of a VB factor entry as every gather copy stores it (bnmf_vb.py:254): one expression for all
// the kernels that write such copies, so that they agree bit for bit
__device__ __forceinline__ double second_moment(double e, double v) { return fma(e, e, v); }

template <int NPT>
struct SweepCfg {
  static constexpr int MAXT = NPT < 16 ? 128 : 512;
  static constexpr bool HOLDV = NPT <= 16;
};

// bid / nblk: index of this thread block among the nblk blocks that share the rows of p
// (blockIdx.x / gridDim.x for a single model; the batched launch of kernels_batch.cuh maps
// blockIdx.y to the model).
template <int NPT, int MODE>
__device__ __forceinline__ void sweep_body(const SweepParams& p, const int bid, const int nblk) {
  constexpr int NV = (MODE == M_VB) ? 2 : 1;
  constexpr bool HOLDV = SweepCfg<NPT>::HOLDV;
  extern __shared__ double smem[];
  const int tpr = p.tpr, K = p.K;
  const int teams = blockDim.x / tpr;
  const int team = threadIdx.x / tpr;
  const int t = threadIdx.x - team * tpr;
  const int lane = threadIdx.x & 31;
  const int wteam = t >> 5;
  const int nw = tpr >> 5;
  const int barA = 1 + 2 * team, barB = 2 + 2 * team;
  const int team_stride = 4 * K + 3 * nw + 2 + 2 * p.covL + (MODE == M_GIBBS ? TN_PRE_DOUBLES * K : 0);
  double* xs = smem + (size_t)team * team_stride;  // [K] current values of the row
  double* xvar = xs + K;                            // [K] VB variance
  double* xmu = xvar + K;                           // [K] conditional mean
  double* xtau = xmu + K;                           // [K] conditional precision
  double* red = xtau + K;                           // [3*nw]
  double* bc = red + 3 * nw;                        // [2]
  double* fsv = bc + 2;                             // [covL] (X S)_il of the row   (BNMTF VB)
  double* arow = fsv + p.covL;                      // [covL] A_il of the row
  double* mat = arow + p.covL;                      // [K][6] Gibbs: random material prepared per column (tn_pre_material)
  const bool fast_draw = (MODE == M_GIBBS) && p.single_k == -1;
  const bool has_cov = (MODE == M_VB) && (p.covA != nullptr);
  const bool leader = (t == 0);
  const uint32_t iteration = p.iter_ptr ? *p.iter_ptr : p.iteration;
  const bool dry = (p.single_k == -2);

  for (int row = bid * teams + team; row < p.n_loc; row += nblk * teams) {
    const int64_t grow = p.row0 + row;
    const int64_t base = p.rowstart[row];
    const int npt = (int)((p.rowstart[row + 1] - base) / tpr);

    double e[NPT];  // residual (NP: prediction p)
    double q[(MODE == M_NP) ? NPT : 1];  // NP: r
    uint32_t c[NPT];
#pragma unroll
    for (int n = 0; n < NPT; ++n) {
      if (n < npt) {
        e[n] = p.vals[base + (int64_t)n * tpr + t];
        c[n] = p.cols[base + (int64_t)n * tpr + t] * NV;
      } else {
        e[n] = 0.0;
        c[n] = p.padcol * NV;
      }
      if (MODE == M_NP) { q[n] = e[n]; e[n] = 0.0; }
    }
    if (fast_draw) {   // every thread of the team prepares the draw material of some columns, once per row
      for (int kk = t; kk < K; kk += tpr)
        tn_pre_material_store(p.seed, p.purpose, (uint64_t)grow, (uint32_t)kk, iteration, mat + TN_PRE_DOUBLES * kk);
    }
    for (int kk = t; kk < K; kk += tpr) {
      xs[kk] = p.Xval[grow * K + kk];
      if (MODE == M_VB) {
        xvar[kk] = p.Xvar[grow * K + kk];
        xmu[kk] = p.Xmu[grow * K + kk];
        xtau[kk] = p.Xtau[grow * K + kk];
      }
    }
    if (nw == 1) __syncwarp(); else bar_sync(barB, tpr);
    if (has_cov) {
      for (int l = t; l < p.covL; l += tpr) {
        double a = 0.0;
        for (int kk = 0; kk < K; ++kk) a = fma(xs[kk], p.covS[kk * p.cov_sk + l * p.cov_sl], a);
        fsv[l] = a;
        arow[l] = p.covA[grow * p.covL + l];
      }
      if (nw == 1) __syncwarp(); else bar_sync(barB, tpr);
    }

    // ---- residual of the row from the current factors: e = r - sum_k x_k y_k
    for (int k = 0; k < K && p.single_k != -4; ++k) {
      const double xk = xs[k];
      const double* __restrict__ y = p.Ykm + (size_t)k * p.ldy * NV;
#pragma unroll
      for (int n = 0; n < NPT; ++n) {
        if (MODE == M_NP) e[n] = fma(xk, __ldg(y + c[n]), e[n]);
        else e[n] = fma(-xk, __ldg(y + c[n]), e[n]);
      }
    }

    if (p.single_k == -3 || p.single_k == -4) {
      // projections of the residual onto the columns of a second factor (bnmtf_gibbs.py:277-283:
      // sum_j M_ij E_ij G_jl, the row part of the S_kl numerators) and, in VB mode, the masked
      // sums of its second moments and variances (bnmtf_vb.py:343,352-357)
      for (int l = 0; l < p.K2; ++l) {
        const double* __restrict__ y2 = p.Ykm2 + (size_t)l * p.ldy2 * NV;
        double s[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int n = 0; n < NPT; ++n) {
          if (MODE == M_VB) {
            const double2 yy = __ldg(reinterpret_cast<const double2*>(y2 + c[n]));
            s[0] = fma(yy.x, yy.x, s[0]); s[1] += yy.y; s[2] = fma(e[n], yy.x, s[2]);
          } else if (MODE == M_NP) {   // numerator / denominator row parts of nmtf_np.update_S (nmtf_np.py:181-188)
            const double g = __ldg(y2 + c[n]);
            const double ratio = (c[n] != p.padcol) ? q[n] / e[n] : 0.0;
            s[2] = fma(ratio, g, s[2]); s[1] += g;
          } else {
            s[2] = fma(e[n], __ldg(y2 + c[n]), s[2]);
          }
        }
        team_reduce_to_leader<3>(s, red, nw, wteam, lane, barA, tpr);
        if (leader) {
          if (p.out_mu) p.out_mu[grow * p.K2 + l] = s[2];
          if (p.out_tau) p.out_tau[grow * p.K2 + l] = s[1];
          if (p.out3) p.out3[grow * p.K2 + l] = s[1] - s[0];
        }
        if (nw > 1) bar_sync(barB, tpr);
      }
      continue;
    }

    double acc_esd = 0.0, acc_q = 0.0, acc_prior = 0.0;  // leader only

    const int k_begin = (p.single_k >= 0) ? p.single_k : 0;
    const int k_end = (p.single_k >= 0) ? p.single_k + 1 : K;
    // Gibbs/ICM/NP statistics need no column loop in dry mode
    const bool skip_loop = dry && MODE != M_VB;
    for (int k = k_begin; k < k_end && !skip_loop; ++k) {
      const double* __restrict__ y = p.Ykm + (size_t)k * p.ldy * NV;
      double v[HOLDV ? NPT : 1];
      double s[3] = {0.0, 0.0, 0.0};  // [0]=sum y^2 (NP: num), [1]=sum (var+y^2) (NP: den), [2]=sum e*y
#pragma unroll
      for (int n = 0; n < NPT; ++n) {
        double vn;
        if (MODE == M_VB) {
          const double2 yy = __ldg(reinterpret_cast<const double2*>(y + c[n]));
          vn = yy.x;
          s[1] += yy.y;
        } else {
          vn = __ldg(y + c[n]);
        }
        if (HOLDV) v[n] = vn;
        if (MODE == M_NP) {
          // padded slots (and only those) read y == 0 exactly; skip the 0/0
          const double ratio = (c[n] != p.padcol) ? q[n] / e[n] : 0.0;
          s[0] = fma(vn, ratio, s[0]);
          s[1] += vn;
        } else {
          s[0] = fma(vn, vn, s[0]);
          s[2] = fma(e[n], vn, s[2]);
        }
      }
      if (MODE == M_NP) {
        double r2[2] = {s[0], s[1]};
        team_reduce_to_leader<2>(r2, red, nw, wteam, lane, barA, tpr);
        s[0] = r2[0]; s[1] = r2[1];
      } else {
        team_reduce_to_leader<3>(s, red, nw, wteam, lane, barA, tpr);
      }
      if (leader) {
        const double x_old = xs[k];
        const double lam = p.lamk ? p.lamk[k] : (p.lam_rm ? p.lam_rm[grow * K + k] : 0.0);
        const double tau = p.scal[0];
        const UpdateCtx uc{p.seed, p.purpose, iteration, dry};
        double cm, ct, vn;
        if (has_cov) {   // diff_term - cov_term (bnmtf_vb.py:342-348)
          double cov = 0.0;
          for (int l = 0; l < p.covL; ++l) {
            const double skl = p.covS[k * p.cov_sk + l * p.cov_sl];
            cov = fma(skl * arow[l], fsv[l] - x_old * skl, cov);
          }
          s[2] -= cov;
        }
        double x_new;
        if (fast_draw) {
          Rng rng(uc.seed, uc.purpose, (uint64_t)grow, (uint32_t)k, uc.iteration);
          rng.draw = 2;                                         // Philox blocks 0 and 1 made the material
          const TnPre m = tn_pre_load(mat + TN_PRE_DOUBLES * k);
          ct = tau * s[0];                                      // tauU, bnmf_gibbs.py:205
          const double num = -lam + tau * (s[2] + x_old * s[0]);   // numerator of muU, bnmf_gibbs.py:210
          x_new = tn_draw_fast(num, ct, m, rng, cm);
        } else {
          // the q(U) terms of the ELBO are evaluated lane-parallel after the column loop (vb_row_q)
          x_new = conditional_update<MODE, false>(s, x_old, lam, tau, uc, (uint64_t)grow, k, xmu[k], xtau[k], xvar[k], ct, cm,
                                                  vn, acc_esd, acc_q, acc_prior);
        }
        if (MODE == M_VB) xvar[k] = vn;
        if (p.single_k >= 0) {
          p.out_tau[grow] = ct;
          p.out_mu[grow] = cm;
          x_new = x_old;
        }
        xmu[k] = cm; xtau[k] = ct;
        bc[0] = x_new - x_old;
        xs[k] = x_new;
        if (has_cov && x_new != x_old)
          for (int l = 0; l < p.covL; ++l) fsv[l] = fma(x_new - x_old, p.covS[k * p.cov_sk + l * p.cov_sl], fsv[l]);
      }
      if (nw == 1) __syncwarp(); else bar_sync(barB, tpr);
      const double delta = bc[0];
      if (delta != 0.0) {
#pragma unroll
        for (int n = 0; n < NPT; ++n) {
          const double vn = HOLDV ? v[n] : __ldg(y + c[n]);
          if (MODE == M_NP) e[n] = fma(delta, vn, e[n]);
          else e[n] = fma(-delta, vn, e[n]);
        }
      }
    }

    if (MODE == M_VB && wteam == 0 && p.rowstats != nullptr && p.single_k < 0) acc_q = vb_row_q(xs, xvar, xmu, xtau, K, lane);

    // ---- optional epilogue of a sweep: projections of the residual the row ends with onto the columns of a second
    // factor, out_mu[row][l] = sum_j e_j Y2[l][j] (BNMTF VB: P_il for the S statistics, bnmtf_vb.py:352-357 -- the
    // residual after update_F is the one update_S starts from, so no separate residual pass is needed)
    if (p.single_k == -1 && p.Ykm2 != nullptr && MODE != M_NP) {
      for (int l = 0; l < p.K2; ++l) {
        const double* __restrict__ y2 = p.Ykm2 + (size_t)l * p.ldy2 * NV;
        double s1[1] = {0.0};
#pragma unroll
        for (int n = 0; n < NPT; ++n) s1[0] = fma(e[n], __ldg(y2 + c[n]), s1[0]);
        team_reduce_to_leader<1>(s1, red, nw, wteam, lane, barA, tpr);
        if (leader) p.out_mu[grow * p.K2 + l] = s1[0];
        if (nw > 1) bar_sync(barB, tpr);
      }
    }

    // ---- per-row statistics from the final residual
    if (p.rowstats != nullptr && p.single_k < 0) {
      double st[3] = {0.0, 0.0, 0.0};
      double idiv = 0.0;
#pragma unroll
      for (int n = 0; n < NPT; ++n) {
        if (n < npt) {
          const double r = (MODE == M_NP) ? q[n] : p.vals[base + (int64_t)n * tpr + t];
          const double ee = (MODE == M_NP) ? (r - e[n]) : e[n];
          const bool valid = (c[n] != p.padcol * NV);
          if (valid) {
            st[0] = fma(ee, ee, st[0]);
            st[1] += ee;
            st[2] = fma(r - p.rmean, ee, st[2]);
            if (MODE == M_NP) idiv += r * log(r / e[n]) - r + e[n];  // nmf_np.py:160-163
          }
        }
      }
      team_reduce_to_leader<3>(st, red, nw, wteam, lane, barA, tpr);
      if (nw > 1) bar_sync(barB, tpr);  // red[] reuse
      if (MODE == M_NP) {
        double r1[1] = {idiv};
        team_reduce_to_leader<1>(r1, red, nw, wteam, lane, barA, tpr);
        idiv = r1[0];
      }
      if (leader) {
        double* rs = p.rowstats + (size_t)row * NRS;
        rs[RS_RSS] = st[0]; rs[RS_SUM_E] = st[1]; rs[RS_SUM_RCE] = st[2];
        rs[RS_ESDVAR] = acc_esd; rs[RS_ELBOQ] = acc_q; rs[RS_PRIOR] = acc_prior;
        rs[RS_IDIV] = idiv; rs[RS_SPARE] = 0.0;
      }
    }
    if (nw == 1) __syncwarp(); else bar_sync(barB, tpr);

    // ---- write the row back (coalesced)
    if (p.single_k == -1) {
      for (int kk = t; kk < K; kk += tpr) {
        p.Xval[grow * K + kk] = xs[kk];
        if (MODE == M_VB) p.Xvar[grow * K + kk] = xvar[kk];
        if (p.Xmu != nullptr) { p.Xmu[grow * K + kk] = xmu[kk]; p.Xtau[grow * K + kk] = xtau[kk]; }
        if (p.km_out != nullptr) {
          if (MODE == M_VB) reinterpret_cast<double2*>(p.km_out)[(size_t)kk * p.km_ld + grow] = make_double2(xs[kk], second_moment(xs[kk], xvar[kk]));
          else p.km_out[(size_t)kk * p.km_ld + grow] = xs[kk];
        }
        if (p.trace_out != nullptr) p.trace_out[((size_t)iteration * p.trace_n + grow) * K + kk] = xs[kk];
      }
    }
    if (nw == 1) __syncwarp(); else bar_sync(barB, tpr);
  }
}

template <int NPT, int MODE>
__global__ void __launch_bounds__(SweepCfg<NPT>::MAXT, 1) sweep_kernel(const SweepParams p) {
  sweep_body<NPT, MODE>(p, (int)blockIdx.x, (int)gridDim.x);
}

// -----------------------------------------------------------------------------
// Row-major factor [n_pad][K]  ->  k-major gather layout [K][ld][NV], zero padded
// for i >= n.  NV == 2 (VB): (exp, var + exp^2) pairs (bnmf_vb.py:254).
// -----------------------------------------------------------------------------
// Block-major copies for the blocked cluster sweep (kernels_v4.cuh): the two columns of a block are interleaved per
// row, YE[(kb * ld + i) * 2 + j] = X[i][2 kb + j] and (VB) YQ[...] = Var + X^2 of the same entry (bnmf_vb.py:254);
// zero for columns >= K and rows >= n.
template <bool VB>
__global__ void rm_to_kmB_kernel(const double* __restrict__ Xval, const double* __restrict__ Xvar, double* __restrict__ YE,
                                 double* __restrict__ YQ, int64_t n, int K, int64_t ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int kb = blockIdx.y;
  if (i >= ld) return;
  double2 e = make_double2(0.0, 0.0), q = make_double2(0.0, 0.0);
  if (i < n) {
    const int k = 2 * kb;
    e.x = Xval[i * K + k];
    if (VB) q.x = second_moment(e.x, Xvar[i * K + k]);
    if (k + 1 < K) {
      e.y = Xval[i * K + k + 1];
      if (VB) q.y = second_moment(e.y, Xvar[i * K + k + 1]);
    }
  }
  reinterpret_cast<double2*>(YE)[(size_t)kb * ld + i] = e;
  if (VB) reinterpret_cast<double2*>(YQ)[(size_t)kb * ld + i] = q;
}

// Both gather copies in one launch (sides whose other side runs the per-row-chain cluster sweep need the two of them after
// every sweep): the block-major copies as above plus the k-major copy of rm_to_km_kernel, [K][ld][NV] (NV == 2: the
// (E, Var + E^2) pairs) -- the same values, bit for bit.
template <bool VB>
__global__ void rm_to_km_both_kernel(const double* __restrict__ Xval, const double* __restrict__ Xvar, double* __restrict__ Ykm,
                                     double* __restrict__ YE, double* __restrict__ YQ, int64_t n, int K, int64_t ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int kb = blockIdx.y;
  if (i >= ld) return;
  double2 e = make_double2(0.0, 0.0), q = make_double2(0.0, 0.0);
  const int k = 2 * kb;
  if (i < n) {
    e.x = Xval[i * K + k];
    if (VB) q.x = second_moment(e.x, Xvar[i * K + k]);
    if (k + 1 < K) {
      e.y = Xval[i * K + k + 1];
      if (VB) q.y = second_moment(e.y, Xvar[i * K + k + 1]);
    }
  }
  reinterpret_cast<double2*>(YE)[(size_t)kb * ld + i] = e;
  if (VB) {
    reinterpret_cast<double2*>(YQ)[(size_t)kb * ld + i] = q;
    reinterpret_cast<double2*>(Ykm)[(size_t)k * ld + i] = make_double2(e.x, q.x);
    if (k + 1 < K) reinterpret_cast<double2*>(Ykm)[(size_t)(k + 1) * ld + i] = make_double2(e.y, q.y);
  } else {
    Ykm[(size_t)k * ld + i] = e.x;
    if (k + 1 < K) Ykm[(size_t)(k + 1) * ld + i] = e.y;
  }
}

template <int NV>
__device__ __forceinline__ void rm_to_km_body(const double* __restrict__ Xval, const double* __restrict__ Xvar,
                                              double* __restrict__ Ykm, int64_t n, int K, int64_t ld, int bx, int by) {
  __shared__ double tile[32][33];
  __shared__ double tile2[(NV == 2) ? 32 : 1][33];
  const int64_t i0 = (int64_t)bx * 32;
  const int k0 = by * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t i = i0 + r;
    const int k = k0 + threadIdx.x;
    double a = 0.0, b = 0.0;
    if (i < n && k < K) {
      a = Xval[i * K + k];
      if (NV == 2) b = second_moment(a, Xvar[i * K + k]);
    }
    tile[r][threadIdx.x] = a;
    if (NV == 2) tile2[r][threadIdx.x] = b;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int k = k0 + r;
    const int64_t i = i0 + threadIdx.x;
    if (k < K && i < ld) {
      if (NV == 2) {
        double2 o = make_double2(tile[threadIdx.x][r], tile2[threadIdx.x][r]);
        reinterpret_cast<double2*>(Ykm)[(size_t)k * ld + i] = o;
      } else {
        Ykm[(size_t)k * ld + i] = tile[threadIdx.x][r];
      }
    }
  }
}

template <int NV>
__global__ void rm_to_km_kernel(const double* __restrict__ Xval, const double* __restrict__ Xvar, double* __restrict__ Ykm,
                                int64_t n, int K, int64_t ld) {
  rm_to_km_body<NV>(Xval, Xvar, Ykm, n, K, ld, (int)blockIdx.x, (int)blockIdx.y);
}

// deterministic block sum (fixed tree), blockDim.x <= 1024, result on all threads
__device__ __forceinline__ double block_sum(double v, double* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  for (int i = 0; i < nw; ++i) r += sh[i];
  return r;
}

struct Hyper {
  int ARD;
  double alphatau, betatau, alpha0, beta0;
  double size_Omega;        // global
  double sumlog_lamU, sumlog_lamV;  // non-ARD VB ELBO constants
};

// device scalar block layout
enum { SC_TAU = 0, SC_LOGTAU = 1, SC_ALPHA_S = 2, SC_BETA_S = 3, SC_COUNT = 8 };

// -----------------------------------------------------------------------------
// Start of an iteration: column sums of both factors and the lambda_k update.
//   Gibbs: lambdak[k] ~ Gamma(alpha0+I+J, beta0+sum U_k+sum V_k) bnmf_gibbs.py:150-152,195-201
//   ICM  : mode (alpha-1)/beta                                    nmf_icm.py:150-152
//   VB   : alphak_s, betak_s, exp_lambdak, exp_loglambdak         bnmf_vb.py:157-160,246-249,270-273
// One block per k.  colsum[2][K]; lamstate[4][K] = alphak_s, betak_s, exp_lambdak(lambdak), exp_loglambdak.
// -----------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ void lambda_body(const double* __restrict__ Ukm, int64_t ldu, int64_t I,
                                            const double* __restrict__ Vkm, int64_t ldv, int64_t J, int K, const Hyper& h,
                                            double* colsum, double* lamstate, double* lamk, uint64_t seed,
                                            uint32_t iteration, int update, const uint32_t* iter_ptr, const int k,
                                            double* lam_trace = nullptr) {
  constexpr int NV = (MODE == M_VB) ? 2 : 1;
  __shared__ double sh[32];
  if (iter_ptr) iteration = *iter_ptr;
  double a = 0.0, b = 0.0;
  // (unrolled: the loads of eight steps are in flight together; the order of the additions is the loop's)
#pragma unroll 8
  for (int64_t i = threadIdx.x; i < I; i += blockDim.x) a += Ukm[((size_t)k * ldu + i) * NV];
#pragma unroll 8
  for (int64_t j = threadIdx.x; j < J; j += blockDim.x) b += Vkm[((size_t)k * ldv + j) * NV];
  a = block_sum(a, sh);
  b = block_sum(b, sh);
  if (threadIdx.x == 0) {
    colsum[k] = a;
    colsum[K + k] = b;
    if (h.ARD && update) {
      const double alpha = h.alpha0 + (double)I + (double)J;
      const double beta = h.beta0 + a + b;
      double lam;
      if (MODE == M_GIBBS) {
        Rng rng(seed, RNG_LAMBDA, (uint64_t)k, 0u, iteration);
        lam = gamma_draw(alpha, beta, rng);
      } else if (MODE == M_ICM) {
        lam = (alpha - 1.0) / beta;
      } else {
        lam = alpha / beta;
        lamstate[0 * K + k] = alpha;
        lamstate[1 * K + k] = beta;
        lamstate[3 * K + k] = digamma(alpha) - log(beta);
      }
      lamstate[2 * K + k] = lam;
      lamk[k] = lam;
    }
    // graph replays: this iteration's slot of all_lambdak (what trace_copy_kernel would copy)
    if (lam_trace != nullptr && update) lam_trace[(size_t)iteration * K + k] = lamk[k];
  }
}

template <int MODE>
__global__ void lambda_kernel(const double* __restrict__ Ukm, int64_t ldu, int64_t I, const double* __restrict__ Vkm,
                              int64_t ldv, int64_t J, int K, Hyper h, double* colsum, double* lamstate, double* lamk,
                              uint64_t seed, uint32_t iteration, int update, const uint32_t* iter_ptr, double* lam_trace) {
  lambda_body<MODE>(Ukm, ldu, I, Vkm, ldv, J, K, h, colsum, lamstate, lamk, seed, iteration, update, iter_ptr, (int)blockIdx.x, lam_trace);
}

// -----------------------------------------------------------------------------
// End of an iteration: deterministic reduction of the per-row statistics, the
// tau update and (VB) the ELBO.  Single block.
//   Gibbs: tau ~ Gamma(alphatau+|Omega|/2, betatau+RSS/2)  bnmf_gibbs.py:167,187-193
//   ICM  : mode                                              nmf_icm.py:170
//   VB   : update_tau/update_exp_tau/elbo                    bnmf_vb.py:176-180,191-244,265-268
// stats_out[BNMTF_NSTATS] layout: see include/bnmtf_b200.h.
// -----------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ void finalize_body(const double* __restrict__ rsU, int64_t nU, const double* __restrict__ rsV,
                                              int64_t nV, int K, const Hyper& h, const double* __restrict__ colsum,
                                              const double* __restrict__ lamstate, double* scal, double* stats_out,
                                              uint64_t seed, uint32_t iteration, int update, int64_t I, int64_t J,
                                              uint32_t* iter_ctr) {
  __shared__ double sh[32];
  __shared__ double tot[2 * NRS];
  // graph replays: the iteration index lives on the device; this kernel ends the iteration and advances it
  if (iter_ctr) { iteration = *iter_ctr; stats_out += (size_t)iteration * 16; }
  for (int sidx = 0; sidx < 2; ++sidx) {
    const double* rs = sidx == 0 ? rsU : rsV;
    const int64_t n = sidx == 0 ? nU : nV;
    // only VB consumes the U-side statistics and the ELBO/ESD columns
    auto needed = [&](int f) -> bool {
      return (MODE == M_VB) ? (f <= RS_PRIOR) : (sidx == 1 && (f <= RS_SUM_RCE || (MODE == M_NP && f == RS_IDIV)));
    };
    // one walk over the rows for all fields (the loads of a row are independent; per field the order of the sum is
    // the one of a walk per field)
    double acc[NRS];
#pragma unroll
    for (int f = 0; f < NRS; ++f) acc[f] = 0.0;
    if (rs != nullptr) {
#pragma unroll 4
      for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
#pragma unroll
        for (int f = 0; f < NRS; ++f)
          if (needed(f)) acc[f] += rs[i * NRS + f];
      }
    }
#pragma unroll
    for (int f = 0; f < NRS; ++f) {
      double a = 0.0;
      if (rs != nullptr && needed(f)) a = block_sum(acc[f], sh);
      if (threadIdx.x == 0) tot[sidx * NRS + f] = a;
    }
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  // residual statistics always come from the V-side sweep (last of the iteration)
  const double rss = tot[NRS + RS_RSS], sum_e = tot[NRS + RS_SUM_E], sum_rce = tot[NRS + RS_SUM_RCE];
  for (int i = 0; i < 16; ++i) stats_out[i] = 0.0;
  stats_out[0] = rss; stats_out[1] = sum_e; stats_out[2] = sum_rce;
  const double alpha_s = h.alphatau + h.size_Omega / 2.0;
  double beta_s = h.betatau + 0.5 * rss;
  double tau = scal[SC_TAU];
  if (MODE == M_GIBBS) {
    if (update) { Rng rng(seed, RNG_TAU, 0ull, 0u, iteration); tau = gamma_draw(alpha_s, beta_s, rng); }
  } else if (MODE == M_ICM) {
    if (update) tau = (alpha_s - 1.0) / beta_s;
  } else if (MODE == M_VB) {
    const double esd = rss + tot[NRS + RS_ESDVAR];   // bnmf_vb.py:241-244
    beta_s = h.betatau + 0.5 * esd;
    double exp_logtau = scal[SC_LOGTAU];
    double a_s = scal[SC_ALPHA_S], b_s = scal[SC_BETA_S];
    if (update) {
      a_s = alpha_s; b_s = beta_s;
      tau = a_s / b_s;
      exp_logtau = digamma(a_s) - log(b_s);
      scal[SC_LOGTAU] = exp_logtau; scal[SC_ALPHA_S] = a_s; scal[SC_BETA_S] = b_s;
    }
    const double log2pi = 1.8378770664093454836;
    double elbo = h.size_Omega / 2.0 * (exp_logtau - log2pi) - tau / 2.0 * esd;
    if (h.ARD) {
      double s_loglam = 0.0, s_lam = 0.0, s_logexplam = 0.0, s_lamcol = 0.0, q_lam = 0.0;
      for (int k = 0; k < K; ++k) {
        const double al = lamstate[0 * K + k], be = lamstate[1 * K + k], el = lamstate[2 * K + k], ell = lamstate[3 * K + k];
        s_loglam += ell; s_lam += el; s_logexplam += log(el);
        s_lamcol += el * (colsum[k] + colsum[K + k]);
        q_lam += -al * log(be) + lgamma(al) - (al - 1.0) * ell + be * el;
      }
      elbo += h.alpha0 * log(h.beta0) - lgamma(h.alpha0) + (h.alpha0 - 1.0) * s_loglam - h.beta0 * s_lam;
      elbo += ((double)I + (double)J) * s_logexplam - s_lamcol;
      elbo += q_lam;
    } else {
      elbo += h.sumlog_lamU - tot[RS_PRIOR] + h.sumlog_lamV - tot[NRS + RS_PRIOR];
    }
    elbo += h.alphatau * log(h.betatau) - lgamma(h.alphatau) + (h.alphatau - 1.0) * exp_logtau - h.betatau * tau;
    elbo += tot[RS_ELBOQ] + (double)I * K / 2.0 * log2pi + tot[NRS + RS_ELBOQ] + (double)J * K / 2.0 * log2pi;
    elbo += -a_s * log(b_s) + lgamma(a_s) - (a_s - 1.0) * exp_logtau + b_s * tau;
    stats_out[5] = esd; stats_out[6] = elbo; stats_out[8] = exp_logtau;
  } else {
    stats_out[7] = tot[NRS + RS_IDIV];
  }
  if (update) scal[SC_TAU] = tau;
  stats_out[3] = tau; stats_out[4] = beta_s; stats_out[9] = alpha_s;
  if (iter_ctr && update) *iter_ctr = iteration + 1;
}

template <int MODE>
__global__ void __launch_bounds__(1024) finalize_kernel(const double* __restrict__ rsU, int64_t nU, const double* __restrict__ rsV, int64_t nV,
                                int K, Hyper h, const double* __restrict__ colsum, const double* __restrict__ lamstate,
                                double* scal, double* stats_out, uint64_t seed, uint32_t iteration, int update,
                                int64_t I, int64_t J, uint32_t* iter_ctr) {
  finalize_body<MODE>(rsU, nU, rsV, nV, K, h, colsum, lamstate, scal, stats_out, seed, iteration, update, I, J, iter_ctr);
}

// Graph replays: copy the draws of the iteration *iter_ptr into the trace staging buffers
// (all_U[it], all_V[it], all_lambdak[it]); runs before finalize_kernel advances the index.
static __global__ void trace_copy_kernel(double* __restrict__ dstU, const double* __restrict__ srcU, int64_t nU, double* __restrict__ dstV,
                                  const double* __restrict__ srcV, int64_t nV, double* __restrict__ dstL,
                                  const double* __restrict__ srcL, int64_t nL, const uint32_t* __restrict__ iter_ptr) {
  const size_t it = *iter_ptr;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (dstU) for (int64_t i = t0; i < nU; i += stride) dstU[it * nU + i] = srcU[i];
  if (dstV) for (int64_t i = t0; i < nV; i += stride) dstV[it * nV + i] = srcV[i];
  if (dstL) for (int64_t i = t0; i < nL; i += stride) dstL[it * nL + i] = srcL[i];
}

// ---- packing -----------------------------------------------------------------
// observed entries per row of a dense mask block (n x m, row-major, ld)
static __global__ void count_rows_kernel(const uint8_t* __restrict__ M, int64_t n, int64_t m, int64_t ld, int* __restrict__ rowlen) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  int cnt = 0;
  for (int64_t j = lane; j < m; j += 32) cnt += (M[row * ld + j] != 0);
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) rowlen[row] = cnt;
}

// compact one row per warp, in column order, then pad the tail of the row
static __global__ void pack_rows_kernel(const double* __restrict__ R, const uint8_t* __restrict__ M, int64_t n, int64_t m, int64_t ld,
                                 const int64_t* __restrict__ rowstart, uint32_t padcol, double* __restrict__ vals,
                                 uint32_t* __restrict__ cols) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  int64_t pos = rowstart[row];
  const int64_t end = rowstart[row + 1];
  for (int64_t j0 = 0; j0 < m; j0 += 32) {
    const int64_t j = j0 + lane;
    const bool obs = (j < m) && (M[row * ld + j] != 0);
    const unsigned bal = __ballot_sync(0xffffffffu, obs);
    if (obs) {
      const int off = __popc(bal & ((1u << lane) - 1u));
      vals[pos + off] = R[row * ld + j];
      cols[pos + off] = (uint32_t)j;
    }
    pos += __popc(bal);
  }
  for (int64_t q = pos + lane; q < end; q += 32) { vals[q] = 0.0; cols[q] = padcol; }
}

// dense transpose of an n x m block (doubles and mask bytes)
template <typename T>
__global__ void transpose_kernel(const T* __restrict__ A, int64_t n, int64_t m, int64_t lda, T* __restrict__ B, int64_t ldb) {
  __shared__ T tile[32][33];
  const int64_t i0 = (int64_t)blockIdx.y * 32, j0 = (int64_t)blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t i = i0 + r, j = j0 + threadIdx.x;
    if (i < n && j < m) tile[r][threadIdx.x] = A[i * lda + j];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t j = j0 + r, i = i0 + threadIdx.x;
    if (i < n && j < m) B[j * ldb + i] = tile[threadIdx.x][r];
  }
}

// sum of r and of (r-c)^2 over the packed values; padded slots are flagged by cols == padcol
static __global__ void data_stats_kernel(const double* __restrict__ vals, const uint32_t* __restrict__ cols, int64_t nslots,
                                  uint32_t padcol, double center, double* __restrict__ partial) {
  __shared__ double sh[32];
  double a = 0.0, b = 0.0, cnt = 0.0;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nslots; q += (int64_t)gridDim.x * blockDim.x) {
    if (cols[q] != padcol) { const double r = vals[q]; a += r; b += (r - center) * (r - center); cnt += 1.0; }
  }
  a = block_sum(a, sh); b = block_sum(b, sh); cnt = block_sum(cnt, sh);
  if (threadIdx.x == 0) { partial[blockIdx.x * 3 + 0] = cnt; partial[blockIdx.x * 3 + 1] = a; partial[blockIdx.x * 3 + 2] = b; }
}

// ---- initialise(): Exp(rate lambda) draws or 1/lambda -----------------------
static __global__ void init_factor_kernel(double* __restrict__ X, int64_t n, int K, const double* __restrict__ lam_rm,
                                   const double* __restrict__ lamk, int random, uint64_t seed, uint32_t purpose) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * K) return;
  const int64_t i = idx / K;
  const int k = (int)(idx - i * K);
  const double lam = lamk ? lamk[k] : lam_rm[idx];
  if (random) {
    Rng rng(seed, purpose, (uint64_t)i, (uint32_t)k, 0u);
    X[idx] = exp_draw(lam, rng);
  } else {
    X[idx] = 1.0 / lam;
  }
}

// VB initialise: tau_X = 1, moments of TN(mu, 1)  (bnmf_vb.py:118-135)
static __global__ void vb_init_moments_kernel(const double* __restrict__ mu, double* __restrict__ tau, double* __restrict__ ex,
                                       double* __restrict__ var, int64_t n) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  tau[idx] = 1.0;
  double e, v;
  tn_moments(mu[idx], 1.0, e, v);
  ex[idx] = e; var[idx] = v;
}

// ---- distributions API kernels ------------------------------------------------
static __global__ void api_tn_draw_kernel(const double* mu, const double* tau, double* out, int64_t n, uint64_t seed) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Rng rng(seed, RNG_API, (uint64_t)i, 1u, 0u);
  out[i] = tn_draw(mu[i], tau[i], rng);
}
static __global__ void api_tn_moments_kernel(const double* mu, const double* tau, double* e, double* v, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  tn_moments(mu[i], tau[i], e[i], v[i]);
}
static __global__ void api_gamma_draw_kernel(const double* alpha, const double* beta, double* out, int64_t n, uint64_t seed) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Rng rng(seed, RNG_API, (uint64_t)i, 2u, 0u);
  out[i] = gamma_draw(alpha[i], beta[i], rng);
}
static __global__ void api_exp_draw_kernel(const double* lambda, double* out, int64_t n, uint64_t seed) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Rng rng(seed, RNG_API, (uint64_t)i, 3u, 0u);
  out[i] = exp_draw(lambda[i], rng);
}

}  // namespace bnmtf
