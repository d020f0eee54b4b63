// Blocked cluster sweep with row statistics taken off the dependency chain ("v4").
//
// The cluster sweep of kernels_v2.cuh is bound by a latency chain per factor column: pass over the
// entries -> warp reduction -> exchange over distributed shared memory -> conditional -> delta
// broadcast (profiles/r01_final_sweep_ncu.md: 2.1 k + 0.6 k + 2.4 k cycles per column).  Two facts
// shorten it:
//
//  1. Of the three sums a conditional needs (bnmf_gibbs.py:203-210), only d_k = sum_j e_j y_jk
//     depends on the residual.  a_k = sum_j M_ij y_jk^2 (VB: q_k = sum_j M_ij (Var y + E[y]^2)_jk,
//     bnmf_vb.py:254) depends on the mask and the OTHER factor alone, and so does the cross term
//     g = sum_j y_j,k y_j,k+1 of two neighbouring columns.  The residual pass that opens a sweep
//     gathers every y_jk anyway, so it accumulates a_k, q_k, g on the side (phase A), the cluster
//     sums them once per row set, and the updater prepares sigma_k = (tau a_k)^-1/2 etc. BEFORE the
//     row sums of a column arrive: the conditional on the chain shrinks to a handful of FMAs.
//  2. With g known, two columns share one pass and one exchange (B = 2): the pass yields
//     h_0 = sum e y_0 and h_1 = sum e y_1 from the residual before the block; the updater draws
//     column 0, gets delta_0, corrects h_1 <- h_1 - delta_0 g (the dependence of column k+1 on
//     column k is linear in the residual) and draws column 1.  Same Gauss-Seidel order and the same
//     conditionals as bnmf_gibbs.py:155-164 / nmf_icm.py:155-167 / bnmf_vb.py:167-174; only the
//     association of the float64 sums differs.  Both columns of a block sit interleaved in shared
//     memory (16 bytes per entry), so a pass costs one LDS.128 per entry and block where v2 issued
//     two LDS.64 per entry and column.
//
// Everything else follows v2: a cluster of CS CTAs owns RS rows, CTA c owns column tile c, one
// compute warp per row segment with the residual in registers, warp 0 = TMA producer + updater
// (lane = row, every CTA updates its rows redundantly from the same sums, bitwise identical),
// records travel by st.async with mbarrier completion, packed layout of pack_tiles_kernel.
// Two tile slots of (tw + V2_PAD) x 16 bytes: the pass of block n first applies the deltas of
// block n-1 (reading slot n-1), signals that slot free -- the control warp refills it with block
// n+1 while the dot products of block n (slot n) and the exchange run.
#pragma once
#include "kernels_v2.cuh"

namespace bnmtf {

constexpr int V4_B = 2;          // columns per block

struct Sweep4Layout {
  size_t tiles, mbar, delta, xs, xvar, xmu, xtau, stat_own, stat_tot, part, rowst, proj, zmat, prof, total;   // offsets in doubles
  int nstat;                     // statistics per row: a[KP], g[NBK], (VB) q[KP]
};

__host__ __device__ inline Sweep4Layout sweep4_layout(int tw, int rs, int K, int cs, bool vb, bool want_mu, bool gibbs, int K2 = 0) {
  Sweep4Layout L;
  const int NBK = (K + V4_B - 1) / V4_B, KP = NBK * V4_B;
  L.nstat = KP + NBK + (vb ? KP : 0);
  size_t o = 0;
  L.tiles = o; o += (size_t)2 * (tw + V2_PAD) * V4_B;
  L.mbar = o; o += 4;                                  // 2 tile slots + 2 record buffers
  L.delta = o; o += (size_t)V4_B * rs + (((size_t)V4_B * rs) & 1);
  L.xs = o; o += (size_t)rs * K;
  L.xvar = o; o += vb ? (size_t)rs * K : 0;
  L.xmu = o; o += (want_mu || vb) ? (size_t)rs * K : 0;
  L.xtau = o; o += (want_mu || vb) ? (size_t)rs * K : 0;
  L.stat_own = o; o += (size_t)rs * L.nstat;           // this CTA's tile part of the row statistics (read by the whole cluster)
  L.stat_tot = o; o += (size_t)rs * L.nstat;           // cluster totals
  L.part = o; o += (size_t)2 * cs * rs * V4_B;         // [2][CS][RS][B] records of h, double buffered by block parity
  L.rowst = o; o += (size_t)cs * rs * 4;               // [CS][RS][4] row statistics of the final residual (meet in CTA 0)
  L.proj = o; o += (size_t)cs * rs * K2;               // [CS][RS][K2] projection partials (meet in CTA 0)
  L.zmat = o; o += gibbs ? (size_t)rs * (2 * NBK) * 2 : 0;   // [RS][KP][2]: the two prepared normals of every draw of the set
  L.prof = o; o += 32;                                 // phase clocks (timing experiments)
  L.total = o;
  return L;
}

// Sum four values over the warp with 6 double shuffles (12 SHFL) instead of 20: afterwards the lanes
// 8 i .. 8 i + 7 hold the warp total of v[i] in the return value.  Fixed pairing -> bitwise reproducible.
__device__ __forceinline__ double warp_sum_scatter4(double v0, double v1, double v2, double v3, int lane) {
  const bool up16 = (lane & 16) != 0;
  const double w0 = (up16 ? v2 : v0) + __shfl_xor_sync(0xffffffffu, up16 ? v0 : v2, 16);
  const double w1 = (up16 ? v3 : v1) + __shfl_xor_sync(0xffffffffu, up16 ? v1 : v3, 16);
  const bool up8 = (lane & 8) != 0;
  double u = (up8 ? w1 : w0) + __shfl_xor_sync(0xffffffffu, up8 ? w0 : w1, 8);
  u += __shfl_xor_sync(0xffffffffu, u, 4);
  u += __shfl_xor_sync(0xffffffffu, u, 2);
  u += __shfl_xor_sync(0xffffffffu, u, 1);
  return u;
}

// 16-byte shared-memory load by 32-bit shared address.  volatile: stays behind the mbarrier wait that
// guards the tile (both are volatile asm) but the FMAs consuming it are free to move.
__device__ __forceinline__ void lds_pair(uint32_t addr, double& a, double& b) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
__device__ __forceinline__ double ld_dsmem_f64(uint32_t cluster_addr) {
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(cluster_addr));
  return v;
}

__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f64(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
__device__ __forceinline__ void st_dsmem_f64(uint32_t cluster_addr, double v) {
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(cluster_addr), "d"(v) : "memory");
}

// The random material of a truncated-normal draw (rng.cuh: tn_pre_material) is prepared for a whole row set while the
// set's statistics pass runs: compute warp r makes the material of row r, one column per lane, and writes the six
// doubles to a scratch table in global memory ([CTA][RS][KP][6], L2 resident).  The two normals every draw starts from
// are mirrored in shared memory; the exponential-proposal variates, needed only when the standardised bound exceeds
// TN_FAST_A0, are fetched from the table inside that branch.
static __device__ __noinline__ void tn_pre_material_row(uint64_t seed, uint32_t purpose, uint64_t index, uint32_t k, uint32_t iteration,
                                                        double* __restrict__ gdst, uint32_t zdst) {
  Rng rng(seed, purpose, index, k, iteration);
  const TnPre m = tn_pre_material(rng);
  gdst[0] = m.z1; gdst[1] = m.z2; gdst[2] = m.e1; gdst[3] = m.l1; gdst[4] = m.e2; gdst[5] = m.l2;
  sts_f64(zdst, m.z1); sts_f64(zdst + 8, m.z2);
}

// Truncated-normal draw from the standardised lower bound a = -mu / sigma and the prepared random material: the
// branches and proposals of tn_draw_lazy (rng.cuh), the same draws bit for bit.  Returns sigma * (z - a) >= 0;
// ok = false when both proposals were rejected (the caller then draws by the inverse CDF, tn_draw_fallback).
__device__ __forceinline__ double tn_draw_bound(double a, double sigma, double z1, double z2, const double* __restrict__ gmat, bool& ok) {
  double x;
  if (!(a > TN_FAST_A0)) {
    ok = (z1 >= a);
    x = z1 - a;
    if (!ok) { ok = (z2 >= a); x = z2 - a; }
  } else {
    const double e1 = gmat[2], l1 = gmat[3], e2 = gmat[4], l2 = gmat[5];
    const double sq = sqrt(a * a + 4.0);
    const double alpha = 0.5 * (a + sq);
    const double ia = a < 256.0 ? 0.5 * (sq - a) : 1.0 / alpha, am = a - alpha;
    x = e1 * ia;
    double t = am + x;
    ok = (l1 >= t * t);
    if (!ok) { x = e2 * ia; t = am + x; ok = (l2 >= t * t); }
  }
  const double d = sigma * x;
  return (d >= 0.0 && isfinite(d)) ? d : 0.0;
}

// phase clocks of cluster 0 / CTA 0 (timing experiments: build with -DBNMTF_V2_PROF, run with BNMTF_V2_DBG=2); summed in
// shared memory (a global read-modify-write per phase would stall the warp for an L2 round trip) and written out at the end
#ifdef BNMTF_V2_PROF
#define P4_T0() long long t_prof = prof_on ? clock64() : 0
#define P4_ADD(slot)                                                       \
  do {                                                                     \
    if (prof_on) {                                                         \
      const long long t_now = clock64();                                   \
      if (lane == 0) prof_sm[slot] += (unsigned long long)(t_now - t_prof); \
      t_prof = t_now;                                                      \
    }                                                                      \
  } while (0)
#else
#define P4_T0() do { } while (0)
#define P4_ADD(slot) do { } while (0)
#endif

// p.Ykm: block-major copy of the other factor's values, [NBK][ldy][2] (rm_to_kmB_kernel); p.Ykm2 (VB only): the same
// layout for Var + E^2; p.gmat (Gibbs only): scratch of gridDim.x * RS * KP * 6 doubles for the random material.
// WPC: warps per SM (two CTAs of WPC / 2 warps): 24 -> 11 rows per CTA at 80 registers, 22 -> 10 rows at 88,
// 20 -> 9 rows at 96 (more gathers in flight per warp, fewer rows in flight).
template <int NPT, int MODE, int CS, int WPC>
__global__ void __launch_bounds__(WPC / 2 * 32, 2) sweep4_kernel(const Sweep2Params p) {
  constexpr int B = V4_B;
  constexpr int RS = WPC / 2 - 1;
  constexpr int NTHR = (RS + 1) * 32;
  constexpr bool VB = (MODE == M_VB);
  constexpr int CH = WPC >= 24 ? 2 : 4;                   // entries whose gathers are issued together (phase B)
  constexpr int NACC = WPC >= 24 ? 1 : 2;                 // accumulator pairs of the dot products
  static_assert(MODE != M_NP, "the NP updates run on the row-team kernel");
  static_assert(CS == 8, "the record exchange assumes 8 CTAs per cluster");
  static_assert(NPT % 4 == 0, "entries per lane: multiple of 4");
  const int c = (int)cg::this_cluster().block_rank();
  const int cid = blockIdx.x / CS, nclusters = gridDim.x / CS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = p.K, tw = p.tw;
  const int NBK = (K + B - 1) / B, KP = NBK * B;
  const int TA = VB ? 2 * NBK : NBK;                      // tiles of phase A (VB: values and second moments alternate)
  const int total = TA + NBK;                             // tiles per set
  const uint32_t slot_bytes = (uint32_t)(tw + V2_PAD) * B * 8u;
  const uint32_t tile_bytes = (uint32_t)tw * B * 8u;
  const bool want_mu = p.Xmu != nullptr;

  extern __shared__ __align__(128) double smem4[];
  const Sweep4Layout L = sweep4_layout(tw, RS, K, CS, VB, want_mu, MODE == M_GIBBS, 0);
  const int NSTAT = L.nstat;
  // shared-memory addresses (32-bit, warp uniform): everything below is addressed through them
  const uint32_t sm = smem_u32(smem4);
  const uint32_t tiles_a = sm + (uint32_t)L.tiles * 8u, mbar_a = sm + (uint32_t)L.mbar * 8u, delta_a = sm + (uint32_t)L.delta * 8u;
  const uint32_t xs_a = sm + (uint32_t)L.xs * 8u, stat_own_a = sm + (uint32_t)L.stat_own * 8u, stat_tot_a = sm + (uint32_t)L.stat_tot * 8u;
  const uint32_t part_a = sm + (uint32_t)L.part * 8u, rowst_a = sm + (uint32_t)L.rowst * 8u, zmat_a = sm + (uint32_t)L.zmat * 8u;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem4 + L.mbar);
  double* xs = smem4 + L.xs;                              // [RS][K]
  double* xvar = smem4 + L.xvar;
  double* xmu = smem4 + L.xmu;
  double* xtau = smem4 + L.xtau;
#ifdef BNMTF_V2_PROF
  unsigned long long* prof_sm = reinterpret_cast<unsigned long long*>(smem4 + L.prof);
  if (threadIdx.x < 32) prof_sm[threadIdx.x] = 0ull;
#endif

  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; ++s) mbar_init(&mbar[s], 1);   // 0-1: tile slots, 2-3: records of even / odd blocks
    fence_mbar_init();
  }
  for (int q = threadIdx.x; q < 2 * V2_PAD * B; q += blockDim.x) {   // the zero entries read by padded slots
    const int s = q / (V2_PAD * B), z = q - s * (V2_PAD * B);
    sts_f64(tiles_a + (uint32_t)s * slot_bytes + tile_bytes + (uint32_t)z * 8u, 0.0);
  }
  __syncthreads();

  // statistics of a row: a_k at [k], g of block m at [KP + m], (VB) q_k at [KP + NBK + k].
  // Sum the CS tile parts in a fixed order (every CTA, bitwise identical).
  auto gather_stats = [&]() {
    for (int q = threadIdx.x; q < RS * NSTAT; q += blockDim.x) {
      double t = 0.0;
#pragma unroll
      for (int cc = 0; cc < CS; ++cc) t += ld_dsmem_f64(map_to_cta(stat_own_a + (uint32_t)q * 8u, (uint32_t)cc));
      sts_f64(stat_tot_a + (uint32_t)q * 8u, t);
    }
  };
  // this CTA's part of the random-material table
  double* gmat_cta = (MODE == M_GIBBS) ? p.gmat + (size_t)blockIdx.x * RS * KP * TN_PRE_DOUBLES : nullptr;

  if (warp == 0) {
    // =========================== control warp ===========================
    const double* ybase = p.Ykm + (size_t)c * tw * B;     // this CTA's tile of block 0
    const double* qbase = VB ? p.Ykm2 + (size_t)c * tw * B : nullptr;
    const size_t ystride = (size_t)p.ldy * B;
    auto issue_tile = [&](uint32_t gseq, int t) {         // one elected lane; tile t of the set's sequence
      const double* src = t >= TA ? ybase + (size_t)(t - TA) * ystride
                                  : (!VB ? ybase + (size_t)t * ystride : ((t & 1) ? qbase : ybase) + (size_t)(t >> 1) * ystride);
      const uint32_t slot = (gseq + (uint32_t)t) & 1u;
      fence_proxy_async();
      mbar_expect_tx(&mbar[slot], tile_bytes);
      tma_load_1d(smem4 + L.tiles + (size_t)slot * (slot_bytes / 8u), src, tile_bytes, &mbar[slot]);
    };
    uint32_t gseq = 0;
    uint32_t rec_phase[2] = {0u, 0u};
    const uint32_t iteration = p.iter_ptr ? *p.iter_ptr : p.iteration;
    const double tau = p.scal[0];
#ifdef BNMTF_V2_PROF
    const bool prof_on = (p.dbg & 2) && cid == 0 && c == 0 && p.prof != nullptr;
#endif
    for (int set = cid; set < p.nsets; set += nclusters) {
      const int rows_here = min(RS, p.n_loc - set * RS);
      const int64_t row_base = p.row0 + (int64_t)set * RS;
      P4_T0();
      for (int q = threadIdx.x; q < rows_here * K; q += blockDim.x) xs[q] = p.Xval[(size_t)row_base * K + q];
      if (lane == 0) {                       // prime both slots
        issue_tile(gseq, 0);
        if (total > 1) issue_tile(gseq, 1);
      }
      cluster_sync_all();                    // S1: xs visible; every CTA has left the previous set
      for (int t = 0; t < TA; ++t) {         // phase A: refill the slot each tile leaves
        __syncthreads();
        if (lane == 0 && t + 2 < total) issue_tile(gseq, t + 2);
      }
      cluster_sync_all();                    // S2: the tile parts of the row statistics are complete in every CTA
      gather_stats();
      __syncthreads();
      P4_ADD(0);                             // set start + phase A (control side)
      double acc_esd = 0.0, acc_prior = 0.0;
      const int64_t grow = row_base + lane;
      const bool sampler = (lane < rows_here);
      const uint32_t st_a = stat_tot_a + (uint32_t)((sampler ? lane : 0) * NSTAT) * 8u;   // this lane's row of the statistics
      const double* gmat_row = (MODE == M_GIBBS) ? gmat_cta + (size_t)(sampler ? lane : 0) * KP * TN_PRE_DOUBLES : nullptr;
      for (int n = 0; n < NBK; ++n) {
        const int k0 = n * B, nb = min(B, K - k0);
        if (n >= 1) {
          named_sync(2, NTHR);               // every compute warp has applied the deltas of block n-1: its slot is free
          if (lane == 0 && TA + n + 1 < total) issue_tile(gseq, TA + n + 1);
        }
        P4_ADD(2);                           // waiting for the slot
        // ---- everything that does not need the residual sums happens before they arrive.  With
        //        num = -lam + tau (h + x_old a),   prec = tau a (VB: tau q),   sigma = prec^-1/2
        // Gibbs needs the standardised bound  -num sigma = c1 - c2 h,  c1 = sigma (lam - tau x_old a), c2 = sigma tau;
        // ICM / VB need the mean  num / prec = c1 + c2 h,  c1 = (-lam + tau x_old a) / prec, c2 = tau / prec.
        double x_old[B], sg[B], c1[B], c2[B], z1[B], z2[B], ctv[B];
        const double g01 = lds_f64(st_a + (uint32_t)(KP + n) * 8u);
#pragma unroll
        for (int j = 0; j < B; ++j) {
          x_old[j] = 0.0; sg[j] = 0.0; c1[j] = 0.0; c2[j] = 0.0; z1[j] = 0.0; z2[j] = 0.0; ctv[j] = 0.0;
          if (sampler && j < nb) {
            const int k = k0 + j;
            const double lam = p.lamk ? p.lamk[k] : (p.lam_rm ? p.lam_rm[grow * K + k] : 0.0);
            const double a = lds_f64(st_a + (uint32_t)k * 8u);                                   // sum_j M y^2
            const double ct = tau * (VB ? lds_f64(st_a + (uint32_t)(KP + NBK + k) * 8u) : a);    // tauU, bnmf_gibbs.py:205 / bnmf_vb.py:254
            x_old[j] = xs[lane * K + k];
            ctv[j] = ct;
            if (MODE == M_GIBBS) {
              sg[j] = rsqrt(ct);
              c2[j] = sg[j] * tau;
              c1[j] = sg[j] * (lam - tau * (x_old[j] * a));
              z1[j] = lds_f64(zmat_a + (uint32_t)((lane * KP + k) * 2) * 8u);
              z2[j] = lds_f64(zmat_a + (uint32_t)((lane * KP + k) * 2 + 1) * 8u);
            } else {
              sg[j] = 1.0 / ct;                                      // ICM / VB: reciprocal of the precision
              c2[j] = tau * sg[j];
              c1[j] = (tau * (x_old[j] * a) - lam) * sg[j];
            }
          }
        }
        P4_ADD(1);                           // preparation
        if (lane == 0) mbar_expect_tx(&mbar[2 + (n & 1)], (uint32_t)(CS * RS * B * 8));
        mbar_wait(&mbar[2 + (n & 1)], rec_phase[n & 1] & 1u);
        rec_phase[n & 1]++;
        P4_ADD(3);                           // waiting for the records of all CTAs
        double h[B], x_new[B], vn[B];
#pragma unroll
        for (int j = 0; j < B; ++j) { h[j] = 0.0; x_new[j] = 0.0; vn[j] = 0.0; }
        if (sampler) {
          {
            double rec[CS][B];
            const uint32_t src = part_a + (uint32_t)(((n & 1) * CS * RS + lane) * B) * 8u;
#pragma unroll
            for (int cc = 0; cc < CS; ++cc) {
#pragma unroll
              for (int j = 0; j < B; ++j) rec[cc][j] = lds_f64(src + (uint32_t)((cc * RS * B + j) * 8));
            }
#pragma unroll
            for (int j = 0; j < B; ++j)        // fixed pairing: ((0+1)+(2+3)) + ((4+5)+(6+7))
              h[j] = ((rec[0][j] + rec[1][j]) + (rec[2][j] + rec[3][j])) + ((rec[4][j] + rec[5][j]) + (rec[6][j] + rec[7][j]));
          }
          double dprev = 0.0;
#pragma unroll
          for (int j = 0; j < B; ++j) {
            double d = 0.0;
            if (j < nb) {
              if (j > 0) h[j] = fma(-dprev, g01, h[j]);      // column k0+1 sees the residual after column k0 moved
              if (p.dbg & 1) x_new[j] = x_old[j] + 1e-3 * h[j];
              else if (MODE == M_GIBBS) {
                const double ab = fma(-c2[j], h[j], c1[j]);   // standardised bound -mu / sigma
                bool ok;
                x_new[j] = tn_draw_bound(ab, sg[j], z1[j], z2[j], gmat_row + (size_t)(k0 + j) * TN_PRE_DOUBLES, ok);
                if (!(sg[j] < CUDART_INF)) { x_new[j] = 0.0; ok = true; }   // precision 0: TN_vector_draw returns 0 (truncated_normal_vector.py:41)
                if (!ok) {                                     // both proposals rejected (rare): inverse CDF, off the fast path
                  const int k = k0 + j;
                  const double lam = p.lamk ? p.lamk[k] : (p.lam_rm ? p.lam_rm[grow * K + k] : 0.0);
                  const double a = lds_f64(st_a + (uint32_t)k * 8u);
                  x_new[j] = tn_draw_fallback(-lam + tau * (h[j] + x_old[j] * a), tau * a, p.seed, p.purpose, (uint64_t)grow, (uint32_t)k, iteration).x;
                }
              } else if (MODE == M_ICM) {
                x_new[j] = fmax(fmax(0.0, fma(c2[j], h[j], c1[j])), MINIMUM_TN);   // nmf_icm.py:159-160
              } else {
                tn_moments(fma(c2[j], h[j], c1[j]), ctv[j], x_new[j], vn[j]);   // bnmf_vb.py:255,275-278
              }
              d = x_new[j] - x_old[j];
            }
            sts_f64(delta_a + (uint32_t)(j * RS + lane) * 8u, d);
            dprev = d;
          }
        }
        P4_ADD(4);                           // record sums + conditionals
        named_arrive(3, NTHR);               // deltas of block n are in this CTA's shared memory
        // ---- bookkeeping, off the chain: the new values, the conditional parameters, the VB sums
        if (sampler) {
#pragma unroll
          for (int j = 0; j < B; ++j) {
            if (j < nb) {
              const int k = k0 + j;
              xs[lane * K + k] = x_new[j];
              if (want_mu || VB) {
                const double lam = p.lamk ? p.lamk[k] : (p.lam_rm ? p.lam_rm[grow * K + k] : 0.0);
                const double a = lds_f64(st_a + (uint32_t)k * 8u);
                const double q = VB ? lds_f64(st_a + (uint32_t)(KP + NBK + k) * 8u) : a;
                const double num = -lam + tau * (h[j] + x_old[j] * a);   // numerator of muU, bnmf_gibbs.py:210
                xmu[lane * K + k] = (MODE == M_GIBBS) ? num * sg[j] * sg[j] : num * sg[j];
                xtau[lane * K + k] = tau * q;
                if (VB) {
                  xvar[lane * K + k] = vn[j];
                  acc_esd += (vn[j] + x_new[j] * x_new[j]) * q - (x_new[j] * x_new[j]) * a;
                  acc_prior += lam * x_new[j];
                }
              }
            }
          }
        }
        P4_ADD(5);                           // bookkeeping
      }
      // VB: q(U_i) part of the ELBO, one lane per column (bnmf_vb.py:219-229), off the chain
      double acc_q = 0.0;
      __syncthreads();                       // every warp has applied the last deltas; xs is final
      if (VB && p.rowstats != nullptr && c == 0) {
        for (int rr = 0; rr < rows_here; ++rr) {
          const double qq = vb_row_q(xs + (size_t)rr * K, xvar + (size_t)rr * K, xmu + (size_t)rr * K, xtau + (size_t)rr * K, K, lane);
          if (lane == rr) acc_q = qq;
        }
      }
      cluster_sync_all();                    // S3: the row statistics of the final residual are in CTA 0
      if (c == 0) {
        if (lane < rows_here && p.rowstats != nullptr) {
          double t3[3] = {0.0, 0.0, 0.0};
          for (int cc = 0; cc < CS; ++cc) {
            const uint32_t src = rowst_a + (uint32_t)((cc * RS + lane) * 4) * 8u;
            t3[0] += lds_f64(src); t3[1] += lds_f64(src + 8); t3[2] += lds_f64(src + 16);
          }
          double* rs = p.rowstats + (size_t)(set * RS + lane) * NRS;
          rs[RS_RSS] = t3[0]; rs[RS_SUM_E] = t3[1]; rs[RS_SUM_RCE] = t3[2];
          rs[RS_ESDVAR] = acc_esd; rs[RS_ELBOQ] = acc_q; rs[RS_PRIOR] = acc_prior;
          rs[RS_IDIV] = 0.0; rs[RS_SPARE] = 0.0;
        }
        for (int q = threadIdx.x; q < rows_here * K; q += blockDim.x) {
          const size_t g = (size_t)row_base * K + q;
          p.Xval[g] = xs[q];
          if (VB) p.Xvar[g] = xvar[q];
          if (want_mu) { p.Xmu[g] = xmu[q]; p.Xtau[g] = xtau[q]; }
        }
      }
      gseq += (uint32_t)total;
      __syncthreads();                       // xs / tiles are reused by the next set
      P4_ADD(6);
    }
  } else {
    // =========================== compute warps ===========================
    // (register diet: at 80 registers the residual segment takes 40 and its packed offsets 10; everything else in the
    // loops is a 32-bit shared address or a value recomputed where it is needed)
    const int r = warp - 1;                                 // row of the set owned by this warp
    uint32_t gseq = 0;
    // The gathers of a loop are issued CH at a time: the address of a chunk is made to depend on the last value of the
    // chunk before (through a mask that is zero at run time but unknown to the compiler).  Left alone the compiler either
    // serialises every gather behind its own FMAs (one load in flight) or issues all of them at once and spills the
    // packed offsets to local memory to make room -- 25 GB of local-memory traffic per sweep in the first version.
    const uint32_t zmask = (uint32_t)p.dbg & 0x40000000u;
#define V4_CHAIN(T_, y_) T_ += ((uint32_t)__double2hiint(y_) & zmask)
#ifdef BNMTF_V2_PROF
    const bool prof_on = (p.dbg & 2) && cid == 0 && c == 0 && warp == 1 && p.prof != nullptr;
#endif
    for (int set = cid; set < p.nsets; set += nclusters) {
      const bool live = r < min(RS, p.n_loc - set * RS);
      P4_T0();
#ifdef BNMTF_V2_PROF
      if (prof_on && lane == 0) prof_sm[16] += 1;
#endif
      double e[NPT];
      uint32_t cp[NPT / 2];
      {
        const int64_t seg = ((int64_t)(set * CS + c) * NPT) * RS * 32 + (int64_t)r * 32 + lane;
        const double* __restrict__ tv = p.tvals + seg;
        const uint16_t* __restrict__ tc = p.tcols + seg;
#pragma unroll
        for (int n = 0; n < NPT; ++n) e[n] = tv[(size_t)n * RS * 32];
#pragma unroll
        for (int n = 0; n < NPT; n += 2) {
          const uint32_t a = tc[(size_t)n * RS * 32], b = tc[(size_t)(n + 1) * RS * 32];
          cp[n / 2] = (a * (B * 8)) | ((b * (B * 8)) << 16);   // byte offsets inside a tile slot
        }
        const int rows_here = min(RS, p.n_loc - set * RS);
        const double* __restrict__ xsrc = p.Xval + (size_t)(p.row0 + (int64_t)set * RS) * K;
        for (int q = threadIdx.x; q < rows_here * K; q += blockDim.x) sts_f64(xs_a + (uint32_t)q * 8u, xsrc[q]);
      }
      cluster_sync_all();   // S1
      if (MODE == M_GIBBS && live) {   // the random material of this warp's row, one column per lane (see tn_pre_material_row)
        const uint32_t iteration = p.iter_ptr ? *p.iter_ptr : p.iteration;
        for (int k = lane; k < K; k += 32)
          tn_pre_material_row(p.seed, p.purpose, (uint64_t)(p.row0 + (int64_t)set * RS + r), (uint32_t)k, iteration,
                              gmat_cta + ((size_t)r * KP + k) * TN_PRE_DOUBLES, zmat_a + (uint32_t)((r * KP + k) * 2) * 8u);
      }
      P4_ADD(8);            // segment loads + set start barrier + random material

      // ---- phase A: residual of the segment, e = r - sum_k x_k y_k, and the residual-independent row sums
      for (int m = 0; m < NBK; ++m) {
        const int k0 = m * B;
        const uint32_t g = gseq + (uint32_t)(VB ? 2 * m : m);
        mbar_wait(&mbar[g & 1u], (g >> 1) & 1u);
        uint32_t T = tiles_a + (g & 1u) * slot_bytes;
        const double x0 = live ? lds_f64(xs_a + (uint32_t)(r * K + k0) * 8u) : 0.0;
        const double x1 = (live && k0 + 1 < K) ? lds_f64(xs_a + (uint32_t)(r * K + k0 + 1) * 8u) : 0.0;
        double a0 = 0.0, a1 = 0.0, gg = 0.0;
#pragma unroll
        for (int n0 = 0; n0 < NPT; n0 += 2) {
          double y0[2], y1[2];
          lds_pair(T + (cp[n0 / 2] & 0xFFFFu), y0[0], y1[0]);
          lds_pair(T + (cp[n0 / 2] >> 16), y0[1], y1[1]);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            e[n0 + i] = fma(-x0, y0[i], e[n0 + i]);
            e[n0 + i] = fma(-x1, y1[i], e[n0 + i]);
            a0 = fma(y0[i], y0[i], a0);
            a1 = fma(y1[i], y1[i], a1);
            gg = fma(y0[i], y1[i], gg);
          }
          V4_CHAIN(T, y1[1]);
        }
        __syncthreads();                     // the slot may be refilled
        {
          const double tot4 = warp_sum_scatter4(a0, a1, gg, 0.0, lane);
          if ((lane & 7) == 0 && lane < 24)
            sts_f64(stat_own_a + (uint32_t)(r * NSTAT + (lane == 0 ? k0 : (lane == 8 ? k0 + 1 : KP + m))) * 8u, tot4);
        }
        if (VB) {                            // second moments of the same block: plain masked sums
          const uint32_t g2 = g + 1u;
          mbar_wait(&mbar[g2 & 1u], (g2 >> 1) & 1u);
          uint32_t T2 = tiles_a + (g2 & 1u) * slot_bytes;
          double q0 = 0.0, q1 = 0.0;
#pragma unroll
          for (int n0 = 0; n0 < NPT; n0 += 2) {
            double y0[2], y1[2];
            lds_pair(T2 + (cp[n0 / 2] & 0xFFFFu), y0[0], y1[0]);
            lds_pair(T2 + (cp[n0 / 2] >> 16), y0[1], y1[1]);
            q0 += y0[0]; q1 += y1[0]; q0 += y0[1]; q1 += y1[1];
            V4_CHAIN(T2, y1[1]);
          }
          __syncthreads();
          const double tq = warp_sum_pair(q0, q1, lane);
          if ((lane & 15) == 0) sts_f64(stat_own_a + (uint32_t)(r * NSTAT + KP + NBK + k0 + (lane >> 4)) * 8u, tq);
        }
      }
      P4_ADD(9);            // phase A
      cluster_sync_all();   // S2
      gather_stats();
      __syncthreads();
      P4_ADD(10);           // statistics exchange

      // ---- phase B: the K column updates, B columns per exchange
      double d0 = 0.0, d1 = 0.0;
      for (int n = 0; n < NBK; ++n) {
        const uint32_t g = gseq + (uint32_t)(TA + n);
        uint32_t T = tiles_a + (g & 1u) * slot_bytes;
        if (n >= 1) {
          uint32_t Tp = tiles_a + ((g & 1u) ^ 1u) * slot_bytes;   // block n-1
#pragma unroll
          for (int n0 = 0; n0 < NPT; n0 += CH) {
            double y0[CH], y1[CH];
#pragma unroll
            for (int i = 0; i < CH; ++i) {
              const int nn = n0 + i;
              lds_pair(Tp + ((nn & 1) ? (cp[nn / 2] >> 16) : (cp[nn / 2] & 0xFFFFu)), y0[i], y1[i]);
            }
#pragma unroll
            for (int i = 0; i < CH; ++i) {
              const int nn = n0 + i;
              e[nn] = fma(-d1, y1[i], fma(-d0, y0[i], e[nn]));
            }
            V4_CHAIN(Tp, y1[CH - 1]);
          }
          named_arrive(2, NTHR);             // this warp no longer reads the slot of block n-1
        }
        P4_ADD(11);                          // corrections of block n-1
        mbar_wait(&mbar[g & 1u], (g >> 1) & 1u);
        P4_ADD(12);                          // waiting for the tile
        double hh;
        {
          double h0[NACC], h1[NACC];
#pragma unroll
          for (int q = 0; q < NACC; ++q) { h0[q] = 0.0; h1[q] = 0.0; }
#pragma unroll
          for (int n0 = 0; n0 < NPT; n0 += CH) {
            double y0[CH], y1[CH];
#pragma unroll
            for (int i = 0; i < CH; ++i) {
              const int nn = n0 + i;
              lds_pair(T + ((nn & 1) ? (cp[nn / 2] >> 16) : (cp[nn / 2] & 0xFFFFu)), y0[i], y1[i]);
            }
#pragma unroll
            for (int i = 0; i < CH; ++i) {
              h0[i % NACC] = fma(e[n0 + i], y0[i], h0[i % NACC]); h1[i % NACC] = fma(e[n0 + i], y1[i], h1[i % NACC]);
            }
            V4_CHAIN(T, y1[CH - 1]);
          }
          P4_ADD(13);                        // dot products
          if (NACC == 2) { h0[0] += h0[NACC - 1]; h1[0] += h1[NACC - 1]; }
          hh = warp_sum_pair(h0[0], h1[0], lane);   // lanes 0..15: h_0, lanes 16..31: h_1
        }
        if ((lane & 15) < CS) {              // lanes cc and 16 + cc send this warp's two sums to CTA cc
          const uint32_t dst = map_to_cta(part_a + (uint32_t)((((n & 1) * CS + c) * RS + r) * B + (lane >> 4)) * 8u, lane & 15);
          const uint32_t bar = map_to_cta(mbar_a + 16u + (uint32_t)(n & 1) * 8u, lane & 15);
          st_async_f64(dst, hh, bar);
        }
        P4_ADD(14);                          // reduction + record stores
        named_sync(3, NTHR);                 // the control warp of this CTA has written the deltas of block n
        d0 = live ? lds_f64(delta_a + (uint32_t)r * 8u) : 0.0;
        d1 = live ? lds_f64(delta_a + (uint32_t)(RS + r) * 8u) : 0.0;
        P4_ADD(15);                          // waiting for the deltas
      }

      // ---- last corrections, then the row statistics of the final residual
      {
        uint32_t Tp = tiles_a + ((gseq + (uint32_t)(total - 1)) & 1u) * slot_bytes;
        const double* __restrict__ tv = p.tvals + (((int64_t)(set * CS + c) * NPT) * RS * 32 + (int64_t)r * 32 + lane);
        double st0 = 0.0, st1 = 0.0, st2 = 0.0;
#pragma unroll
        for (int n0 = 0; n0 < NPT; n0 += 2) {
          double y0[2], y1[2];
          lds_pair(Tp + (cp[n0 / 2] & 0xFFFFu), y0[0], y1[0]);
          lds_pair(Tp + (cp[n0 / 2] >> 16), y0[1], y1[1]);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int nn = n0 + i;
            e[nn] = fma(-d1, y1[i], fma(-d0, y0[i], e[nn]));
            const double rr = tv[(size_t)nn * RS * 32];
            st0 = fma(e[nn], e[nn], st0);
            st1 += e[nn];
            st2 = fma(rr - p.rmean, e[nn], st2);   // padded slots hold e == 0
          }
          V4_CHAIN(Tp, y1[1]);
        }
        st0 = warp_sum(st0); st1 = warp_sum(st1); st2 = warp_sum(st2);
        if (lane == 0) {                     // CTA 0's copy, through distributed shared memory
          const uint32_t dst = map_to_cta(rowst_a + (uint32_t)((c * RS + r) * 4) * 8u, 0u);
          st_dsmem_f64(dst, st0); st_dsmem_f64(dst + 8, st1); st_dsmem_f64(dst + 16, st2);
        }
      }
      __syncthreads();      // (control warp: xs is final)
      cluster_sync_all();   // S3
      if (c == 0) {
        const int rows_here = min(RS, p.n_loc - set * RS);
        const size_t g0 = (size_t)(p.row0 + (int64_t)set * RS) * K;
        for (int q = threadIdx.x; q < rows_here * K; q += blockDim.x) {
          p.Xval[g0 + q] = xs[q];
          if (VB) p.Xvar[g0 + q] = xvar[q];
          if (want_mu) { p.Xmu[g0 + q] = xmu[q]; p.Xtau[g0 + q] = xtau[q]; }
        }
      }
      gseq += (uint32_t)total;
      __syncthreads();      // xs / tiles are reused by the next set
      P4_ADD(17);           // statistics, store, end of set
    }
  }
#ifdef BNMTF_V2_PROF
  __syncthreads();
  if ((p.dbg & 2) && cid == 0 && c == 0 && p.prof != nullptr && threadIdx.x < 32) p.prof[threadIdx.x] += prof_sm[threadIdx.x];
#endif
  cluster_sync_all();       // no CTA may exit while its shared memory can still be read or written remotely
}

}  // namespace bnmtf
