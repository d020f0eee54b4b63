// Blocked cluster sweep ("v3"): the v2 cluster sweep (kernels_v2.cuh) with B columns of the
// factor updated per cluster exchange instead of one.
//
// v2 is bound by a latency chain per column: partial sums -> DSMEM -> cluster barrier -> update
// -> delta broadcast.  The chain of column k+1 only depends on column k through the rank-1
// correction of the residual, and that dependence is linear:
//     b_j = sum_n e_n y_nj                                   (e: residual before the block)
//     after column i of the block moved by d_i:   b_j <- b_j - d_i * G_ij,  G_ij = sum_n y_ni y_nj
// so one pass over the entries yields a_j, b_j (and the VB sums) for the B columns of a block
// plus the B(B-1)/2 cross terms G_ij, ONE exchange brings them to the updater, and the updater
// runs the B conditionals back to back (lane = row), correcting b_j with the deltas it has just
// produced.  This is the same Gauss-Seidel order and the same conditionals as
// bnmf_gibbs.py:155-164 / nmf_icm.py:155-167 / bnmf_vb.py:167-174 -- only the association of
// the floating-point sums changes.  The residual corrections of a block are folded into the
// pass of the next block (as in v2), so 2B tile stages are live: the current block's and the
// previous block's.
//
// Everything else follows v2: cluster of CS CTAs, CTA c owns column tile c, one compute warp per
// row segment with the residual in registers, warp 0 = TMA producer + updater, every CTA
// updates its rows redundantly from the same partial sums (bitwise identical), packed layout of
// pack_tiles_kernel.
#pragma once
#include "kernels_v2.cuh"

namespace bnmtf {

// values per (row, CTA) record: a[B], b[B], G[B(B-1)/2], (VB) s1[B]; padded to 8 or 16 so that the
// halving reduction below ends with one value per lane group
template <int B, int MODE>
struct Sweep3Rec {
  static constexpr int RAW = (MODE == M_VB ? 3 : 2) * B + B * (B - 1) / 2;
  static constexpr int NVAL = RAW <= 8 ? 8 : 16;
  static_assert(RAW <= 16, "block too wide for the 16-value record");
};

constexpr int SWEEP3_RNDW = 8;   // columns of prepared random material kept in shared memory

struct Sweep3Layout {
  size_t tiles, mbar, delta, xs, xvar, xmu, xtau, part, stat, rnd, total;   // offsets in doubles
};

__host__ __device__ inline Sweep3Layout sweep3_layout(int tw, int nv, int rs, int K, int cs, int B, int nval, bool vb,
                                                      bool want_mu, bool gibbs) {
  Sweep3Layout L;
  size_t o = 0;
  L.tiles = o; o += (size_t)2 * (tw + V2_PAD) * B * nv;   // 2 slots x (tw + V2_PAD zero entries) x B*NV interleaved doubles
  L.mbar = o; o += 4;   // 2 tile mbarriers + 2 record mbarriers
  L.delta = o; o += (size_t)B * rs + (((size_t)B * rs) & 1);
  L.xs = o; o += (size_t)rs * K;
  L.xvar = o; o += vb ? (size_t)rs * K : 0;
  L.xmu = o; o += want_mu ? (size_t)rs * K : 0;
  L.xtau = o; o += want_mu ? (size_t)rs * K : 0;
  L.part = o; o += (size_t)2 * cs * nval * rs;       // [2][CS][NVAL][RS]
  L.stat = o; o += (size_t)cs * 3 * rs;              // [CS][3][RS]
  L.rnd = o; o += gibbs ? (size_t)SWEEP3_RNDW * rs * TN_PRE_DOUBLES : 0;   // [W][RS][6]: prepared random material
  L.total = o;
  return L;
}

// Sum NVAL values over the warp with NVAL-1+... shuffles instead of 5*NVAL: each halving step
// exchanges half of the live values with the partner lane.  On return every lane holds the warp
// total of value index (lane >> (5 - log2 NVAL)) in v[0].
template <int NVAL>
__device__ __forceinline__ void warp_sum_scatter(double (&v)[NVAL], int lane) {
  static_assert(NVAL == 8 || NVAL == 16, "");
  constexpr int STEPS = NVAL == 8 ? 3 : 4;
  int live = NVAL;
#pragma unroll
  for (int s = 0; s < STEPS; ++s) {
    const int bit = 16 >> s;
    const bool up = (lane & bit) != 0;
    live >>= 1;
#pragma unroll
    for (int i = 0; i < NVAL / 2; ++i) {
      if (i < live) {
        const double send = up ? v[i] : v[i + live];
        const double keep = up ? v[i + live] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
      }
    }
  }
#pragma unroll
  for (int bit = (16 >> STEPS); bit > 0; bit >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], bit);
}

// shared-memory loads by 32-bit shared address (no generic-address arithmetic in the hot loop)
__device__ __forceinline__ void lds_f64x2(uint32_t addr, double& a, double& b) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}

template <int EW>
__device__ __forceinline__ void load_entry(uint32_t addr, double (&y)[EW]) {
  static_assert(EW % 2 == 0, "entries are loaded as 16-byte pairs");
#pragma unroll
  for (int q = 0; q < EW; q += 2) lds_f64x2(addr + q * 8, y[q], y[q + 1]);
}

#define PROF_T0() long long t_prof = prof_on ? clock64() : 0
#define PROF_ADD(slot)                                             \
  do {                                                             \
    if (prof_on) {                                                 \
      const long long t_now = clock64();                           \
      if (lane == 0) p.prof[slot] += (unsigned long long)(t_now - t_prof); \
      t_prof = t_now;                                              \
    }                                                              \
  } while (0)

template <int NPT, int MODE, int CS, int DIV, int B>
__global__ void __launch_bounds__(Sweep2Cfg<NPT>::WPC / DIV * 32, DIV) sweep3_kernel(const Sweep2Params p) {
  constexpr int NV = (MODE == M_VB) ? 2 : 1;
  constexpr int RS = Sweep2Cfg<NPT>::WPC / DIV - 1;
  constexpr int NVAL = Sweep3Rec<B, MODE>::NVAL;
  constexpr int NG = B * (B - 1) / 2;
  constexpr int NTHREADS = (RS + 1) * 32;
  static_assert(MODE != M_NP, "the NP updates run on the v1 kernel");
  static_assert(CS == 8, "the record exchange below assumes 8 CTAs per cluster");
  cg::cluster_group cluster = cg::this_cluster();
  const int c = (int)cluster.block_rank();
  const int cid = blockIdx.x / CS, nclusters = gridDim.x / CS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = p.K, tw = p.tw;
  const int NBK = (K + B - 1) / B;                        // blocks per pass
  constexpr int EW = B * NV;                              // doubles per tile entry: the block's columns interleaved
  const int slot_doubles = (tw + V2_PAD) * EW;            // one stage: tw entries + the zero entries padded slots read
  const uint32_t block_bytes = (uint32_t)tw * EW * 8u;
  const bool want_mu = p.Xmu != nullptr;
  const bool prof_on = (p.dbg & 2) && blockIdx.x == 0 && (warp == 0 || warp == 1);

  extern __shared__ __align__(128) double smem3[];
  const Sweep3Layout L = sweep3_layout(tw, NV, RS, K, CS, B, NVAL, MODE == M_VB, want_mu, MODE == M_GIBBS);
  double* tiles = smem3 + L.tiles;                        // [2 slots][tw + V2_PAD][EW]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem3 + L.mbar);
  double* delta = smem3 + L.delta;                        // [B][RS]
  double* xs = smem3 + L.xs;                              // [RS][K]
  double* xvar = smem3 + L.xvar;
  double* xmu = smem3 + L.xmu;
  double* xtau = smem3 + L.xtau;
  double* part = smem3 + L.part;                          // [2][CS][NVAL][RS]
  double* stat = smem3 + L.stat;                          // [CS][3][RS]
  double* rnd = smem3 + L.rnd;                            // [W][RS][6]

  if (threadIdx.x == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_init(&mbar[2], 1);              // records of even blocks
    mbar_init(&mbar[3], 1);              // records of odd blocks
    fence_mbar_init();
  }
  // stages start as zeros: tail blocks leave stages unloaded, padded entries read the two
  // slots behind each tile
  for (int q = threadIdx.x; q < 2 * slot_doubles; q += blockDim.x) tiles[q] = 0.0;
  __syncthreads();

  // p.Ykm: block-major copy of the other factor, [ceil(K/B)][ldy][EW] (rm_to_kmB_kernel)
  const double* ybase = p.Ykm + (size_t)c * tw * EW;      // this CTA's tile of block 0
  const size_t ystride = (size_t)p.ldy * EW;

  // issue the TMA loads of block `blk` of the 2*NBK-block sequence of a set into its slot
  auto issue_block = [&](uint32_t gblk, int blk) {
    const int slot = (int)(gblk & 1u);
    fence_proxy_async();
    mbar_expect_tx(&mbar[slot], block_bytes);
    tma_load_1d(tiles + (size_t)slot * slot_doubles, ybase + (size_t)(blk % NBK) * ystride, block_bytes, &mbar[slot]);
  };

  if (warp == 0) {
    // =========================== control warp ===========================
    uint32_t gblk = 0;
    uint32_t rec_phase[2] = {0u, 0u};
    const UpdateCtx uc{p.seed, p.purpose, p.iter_ptr ? *p.iter_ptr : p.iteration, false};
    const double tau = p.scal[0];
    // Gibbs: the random material of CPR columns x RS rows is produced per round, lane = (column, row)
    constexpr int CPR = (32 / RS) < 1 ? 1 : (32 / RS);
    static_assert(B + 2 * CPR <= SWEEP3_RNDW || MODE != M_GIBBS, "random-material window too small");
    int gen = 0;
    auto gen_round = [&](int rows_here, int64_t row_base) {
      const int cj = lane / RS, rr = lane - cj * RS;
      const int k = gen + cj;
      if (cj < CPR && k < K && rr < rows_here) {
        Rng rng(uc.seed, uc.purpose, (uint64_t)(row_base + rr), (uint32_t)k, uc.iteration);
        const TnPre m = tn_pre_material(rng);
        double* dst = rnd + ((size_t)(k % SWEEP3_RNDW) * RS + rr) * TN_PRE_DOUBLES;
        dst[0] = m.z1; dst[1] = m.z2; dst[2] = m.e1; dst[3] = m.l1; dst[4] = m.e2; dst[5] = m.l2;
      }
      gen += CPR;
      __syncwarp();
    };
    for (int set = cid; set < p.nsets; set += nclusters) {
      const int rows_here = min(RS, p.n_loc - set * RS);
      if (lane == 0) {                       // prime both slots
        issue_block(gblk, 0);
        if (2 * NBK > 1) issue_block(gblk + 1, 1);
      }
      if (MODE == M_GIBBS) {                 // material of the first block(s), while the compute warps load their segments
        gen = 0;
        while (gen < min(K, B + CPR)) gen_round(rows_here, p.row0 + (int64_t)set * RS);
      }
      cluster_sync_all();
      for (int m = 0; m < NBK; ++m) {        // residual pass: refill the slot each block leaves
        __syncthreads();
        if (lane == 0 && m + 2 < 2 * NBK) issue_block(gblk + m + 2, m + 2);
      }
      double acc_esd = 0.0, acc_q = 0.0, acc_prior = 0.0;
      const int64_t grow = p.row0 + (int64_t)set * RS + lane;
      const bool sampler = (lane < rows_here);   // every CTA updates the RS rows redundantly (bitwise identical)
      PROF_T0();
      for (int n = 0; n < NBK; ++n) {
        const int k0 = n * B, nb = min(B, K - k0);
        // everything that does not need the row sums happens before they arrive
        double lam[B], x_old[B];
        TnPre mat[B];
#pragma unroll
        for (int j = 0; j < B; ++j) {
          lam[j] = 0.0; x_old[j] = 0.0;
          mat[j].z1 = 0.0; mat[j].z2 = 0.0; mat[j].e1 = 1.0; mat[j].l1 = 1.0; mat[j].e2 = 1.0; mat[j].l2 = 1.0;
          if (sampler && j < nb) {
            const int k = k0 + j;
            lam[j] = p.lamk ? p.lamk[k] : (p.lam_rm ? p.lam_rm[grow * K + k] : 0.0);
            x_old[j] = xs[lane * K + k];
            if (MODE == M_GIBBS) {
              mat[j] = tn_pre_load(rnd + ((size_t)(k % SWEEP3_RNDW) * RS + lane) * TN_PRE_DOUBLES);
            }
          }
        }
        if (MODE == M_GIBBS) {               // stay one round ahead; this runs while the compute warps do their pass
          __syncwarp();
          while (gen < min(K, k0 + B + CPR)) gen_round(rows_here, p.row0 + (int64_t)set * RS);
        }
        PROF_ADD(0);                         // prologue + material generation
        named_sync(1, NTHREADS);             // the compute warps of this CTA have finished their pass of block n
        PROF_ADD(1);                         // waiting for the local pass
        if (lane == 0 && n >= 1 && n + 1 < NBK) issue_block(gblk + NBK + n + 1, NBK + n + 1);   // slot of block n-1 is dead
        // the records of block n: CS x RS warps each send NVAL doubles; their st.async complete this CTA's mbarrier
        if (lane == 0) mbar_expect_tx(&mbar[2 + (n & 1)], (uint32_t)(CS * RS * NVAL * 8));
        mbar_wait(&mbar[2 + (n & 1)], rec_phase[n & 1] & 1u);
        rec_phase[n & 1]++;
        PROF_ADD(2);                         // waiting for the records
        // sum the CS records of every row in a fixed order; with RS <= 16 two lanes share a row (lane and lane + 16
        // take four CTAs each), halving the loads on this warp's critical path
        constexpr int RAW = Sweep3Rec<B, MODE>::RAW;
        constexpr bool SPLIT = (RS <= 16);
        double rec[RAW];
#pragma unroll
        for (int q = 0; q < RAW; ++q) rec[q] = 0.0;
        {
          const int rl = SPLIT ? (lane & 15) : lane, cc0 = SPLIT ? (lane >> 4) * (CS / 2) : 0;
          if (rl < rows_here) {
            const double* src = part + (size_t)(n & 1) * CS * NVAL * RS + rl;
#pragma unroll
            for (int cq = 0; cq < (SPLIT ? CS / 2 : CS); ++cq) {
#pragma unroll
              for (int q = 0; q < RAW; ++q) rec[q] += src[(size_t)((cc0 + cq) * NVAL + q) * RS];
            }
          }
          if (SPLIT) {
#pragma unroll
            for (int q = 0; q < RAW; ++q) rec[q] += __shfl_down_sync(0xffffffffu, rec[q], 16);
          }
        }
        if (sampler) {
          double d[B];
#pragma unroll
          for (int j = 0; j < B; ++j) {
            d[j] = 0.0;
            if (j < nb) {
              const int k = k0 + j;
              double bj = rec[B + j];
              // cross terms with the columns of this block that have already moved: pair (i, j), i < j
#pragma unroll
              for (int i = 0; i < j; ++i) bj = fma(-d[i], rec[2 * B + (i * (2 * B - i - 1)) / 2 + (j - i - 1)], bj);
              const double t3[3] = {rec[j], MODE == M_VB ? rec[2 * B + NG + j] : 0.0, bj};
              double ct, cm, vn = 0.0, x_new;
              if (MODE == M_GIBBS) {
                Rng rng(uc.seed, uc.purpose, (uint64_t)grow, (uint32_t)k, uc.iteration);
                rng.draw = 2;                // Philox blocks 0 and 1 of this stream made the material
                ct = tau * t3[0];                                            // tauU, bnmf_gibbs.py:205
                const double num = -lam[j] + tau * (t3[2] + x_old[j] * t3[0]);   // numerator of muU, :210
                bool fb = false;
                x_new = tn_draw_fast(num, ct, mat[j], rng, cm, &fb);
                if (prof_on) {
                  const unsigned fbm = __ballot_sync(__activemask(), fb);
                  if (lane == 0) { p.prof[20] += __popc(fbm); p.prof[21] += fbm != 0; p.prof[22] += 1; }
                }
              } else {
                x_new = conditional_update<MODE>(t3, x_old[j], lam[j], tau, uc, (uint64_t)grow, k, 0.0, 0.0, 0.0, ct, cm, vn,
                                                 acc_esd, acc_q, acc_prior);
              }
              xs[lane * K + k] = x_new;
              if (want_mu) { xmu[lane * K + k] = cm; xtau[lane * K + k] = ct; }
              if (MODE == M_VB) xvar[lane * K + k] = vn;
              d[j] = x_new - x_old[j];
            }
            delta[j * RS + lane] = d[j];
          }
        }
        PROF_ADD(3);                         // record sums + updates
        __syncthreads();                     // deltas of block n are in this CTA's shared memory
        PROF_ADD(4);
      }
      cluster_sync_all();                    // row-statistics partials are in CTA 0
      if (c == 0) {
        if (lane < rows_here && p.rowstats != nullptr) {
          double t3[3] = {0.0, 0.0, 0.0};
          for (int cc = 0; cc < CS; ++cc) {
            t3[0] += stat[(cc * 3 + 0) * RS + lane]; t3[1] += stat[(cc * 3 + 1) * RS + lane]; t3[2] += stat[(cc * 3 + 2) * RS + lane];
          }
          double* rs = p.rowstats + (size_t)(set * RS + lane) * NRS;
          rs[RS_RSS] = t3[0]; rs[RS_SUM_E] = t3[1]; rs[RS_SUM_RCE] = t3[2];
          rs[RS_ESDVAR] = acc_esd; rs[RS_ELBOQ] = acc_q; rs[RS_PRIOR] = acc_prior;
          rs[RS_IDIV] = 0.0; rs[RS_SPARE] = 0.0;
        }
        __syncthreads();
        for (int q = threadIdx.x; q < rows_here * K; q += blockDim.x) {
          const size_t g = (size_t)(p.row0 + set * RS) * K + q;
          p.Xval[g] = xs[q];
          if (MODE == M_VB) p.Xvar[g] = xvar[q];
          if (want_mu) { p.Xmu[g] = xmu[q]; p.Xtau[g] = xtau[q]; }
        }
      }
      gblk += 2 * NBK;
      __syncthreads();
    }
  } else {
    // =========================== compute warps ===========================
    const int r = warp - 1;                                 // row of the set owned by this warp
    const uint32_t tiles_u32 = smem_u32(tiles);
    const uint32_t slot_bytes = (uint32_t)slot_doubles * 8u;
    const uint32_t part_u32 = smem_u32(part), rbar_u32 = smem_u32(&mbar[2]);
    uint32_t gblk = 0;
    // after the scatter reduction lane `lane` holds value index vq and sends it to CTAs t, t+SPAN, ...
    constexpr int GROUP = 32 / NVAL;                        // lanes holding the same value
    const int vq = lane / GROUP, vt = lane % GROUP;
    for (int set = cid; set < p.nsets; set += nclusters) {
      const int rows_here = min(RS, p.n_loc - set * RS);
      const int64_t seg = ((int64_t)(set * CS + c) * NPT) * RS * 32 + (int64_t)r * 32 + lane;
      double e[NPT];
      uint32_t cp[NPT / 2];
#pragma unroll
      for (int n = 0; n < NPT; ++n) e[n] = p.tvals[seg + (int64_t)n * RS * 32];
#pragma unroll
      for (int n = 0; n < NPT; n += 2) {
        const uint32_t a = p.tcols[seg + (int64_t)n * RS * 32], b = p.tcols[seg + (int64_t)(n + 1) * RS * 32];
        cp[n / 2] = a | (b << 16);                           // entry indices inside a tile stage
      }
      for (int q = threadIdx.x - 32; q < rows_here * K; q += blockDim.x - 32) {
        const size_t g = (size_t)(p.row0 + set * RS) * K + q;
        xs[q] = p.Xval[g];
      }
      PROF_T0();
      cluster_sync_all();   // xs visible; every CTA has finished with the previous set
      PROF_ADD(8);          // set start barrier

      // ---- residual of the segment: e = r - sum_k x_k y_k   (blocks 0 .. NBK-1)
      for (int m = 0; m < NBK; ++m) {
        const uint32_t g = gblk + m;
        const int slot = (int)(g & 1u);
        mbar_wait(&mbar[slot], (g >> 1) & 1u);
        const int k0 = m * B, nb = min(B, K - k0);
        double xk[B];
#pragma unroll
        for (int j = 0; j < B; ++j) xk[j] = (r < rows_here && j < nb) ? xs[r * K + k0 + j] : 0.0;
        const uint32_t T = tiles_u32 + (uint32_t)slot * slot_bytes;
#pragma unroll
        for (int n = 0; n < NPT; ++n) {
          const uint32_t col = (n & 1) ? (cp[n / 2] >> 16) : (cp[n / 2] & 0xFFFFu);
          double y[EW];
          load_entry<EW>(T + col * (EW * 8), y);
#pragma unroll
          for (int j = 0; j < B; ++j) e[n] = fma(-xk[j], y[j * NV], e[n]);
        }
        __syncthreads();
      }

      PROF_ADD(9);          // residual pass
      // ---- the K column updates, B at a time   (blocks NBK .. 2 NBK - 1)
      double dprev[B];
#pragma unroll
      for (int j = 0; j < B; ++j) dprev[j] = 0.0;
      for (int n = 0; n < NBK; ++n) {
        const uint32_t g = gblk + NBK + n;
        const int slot = (int)(g & 1u);
        mbar_wait(&mbar[slot], (g >> 1) & 1u);
        PROF_ADD(10);                        // waiting for the tiles
        const uint32_t T = tiles_u32 + (uint32_t)slot * slot_bytes;
        // block n-1 for the folded corrections; at n == 0 there is nothing to fold (dprev == 0) and the
        // other slot is still being filled, so the current one stands in
        const uint32_t Tp = (n > 0) ? tiles_u32 + (uint32_t)(slot ^ 1) * slot_bytes : T;
        double v[NVAL];
#pragma unroll
        for (int q = 0; q < NVAL; ++q) v[q] = 0.0;
#pragma unroll
        for (int nn = 0; nn < NPT; ++nn) {
          const uint32_t eoff = ((nn & 1) ? (cp[nn / 2] >> 16) : (cp[nn / 2] & 0xFFFFu)) * (EW * 8);
          double yp[EW], yc[EW];
          load_entry<EW>(Tp + eoff, yp);
          load_entry<EW>(T + eoff, yc);
#pragma unroll
          for (int j = 0; j < B; ++j) e[nn] = fma(-dprev[j], yp[j * NV], e[nn]);
          double y[B];
#pragma unroll
          for (int j = 0; j < B; ++j) {
            y[j] = yc[j * NV];
            if (MODE == M_VB) v[2 * B + NG + j] += yc[j * NV + 1];
            v[j] = fma(y[j], y[j], v[j]);
            v[B + j] = fma(e[nn], y[j], v[B + j]);
          }
#pragma unroll
          for (int i = 0; i < B; ++i) {
#pragma unroll
            for (int j = i + 1; j < B; ++j) {
              const int q = 2 * B + (i * (2 * B - i - 1)) / 2 + (j - i - 1);
              v[q] = fma(y[i], y[j], v[q]);
            }
          }
        }
        PROF_ADD(11);                        // pass
        named_arrive(1, NTHREADS);           // this warp no longer reads the slot of block n-1
        warp_sum_scatter<NVAL>(v, lane);
        {
          // value vq goes to every CTA: the GROUP lanes that hold it share the CS destinations
          const size_t rec = ((size_t)(n & 1) * CS * NVAL + (size_t)c * NVAL + vq) * RS + r;
          const uint32_t dst_local = part_u32 + (uint32_t)rec * 8u, bar_local = rbar_u32 + (uint32_t)(n & 1) * 8u;
#pragma unroll
          for (int cc = vt; cc < CS; cc += GROUP) st_async_f64(map_to_cta(dst_local, cc), v[0], map_to_cta(bar_local, cc));
        }
        PROF_ADD(12);                        // reduction + DSMEM stores
        PROF_ADD(13);
        __syncthreads();                     // the control warp of this CTA has written the deltas
        PROF_ADD(14);                        // waiting for the deltas
#pragma unroll
        for (int j = 0; j < B; ++j) dprev[j] = (r < rows_here) ? delta[j * RS + r] : 0.0;
      }

      // ---- last corrections, then the row statistics of the final residual
      {
        const uint32_t g = gblk + 2 * NBK - 1;
        const uint32_t Tp = tiles_u32 + (uint32_t)(g & 1u) * slot_bytes;
        double st0 = 0.0, st1 = 0.0, st2 = 0.0;
#pragma unroll
        for (int n = 0; n < NPT; ++n) {
          const uint32_t col = (n & 1) ? (cp[n / 2] >> 16) : (cp[n / 2] & 0xFFFFu);
          double yp[EW];
          load_entry<EW>(Tp + col * (EW * 8), yp);
#pragma unroll
          for (int j = 0; j < B; ++j) e[n] = fma(-dprev[j], yp[j * NV], e[n]);
          const double rr = p.tvals[seg + (int64_t)n * RS * 32];
          st0 = fma(e[n], e[n], st0);
          st1 += e[n];
          st2 = fma(rr - p.rmean, e[n], st2);   // padded slots hold e == 0
        }
        st0 = warp_sum(st0); st1 = warp_sum(st1); st2 = warp_sum(st2);
        if (lane == 0) {
          double* dst = cluster.map_shared_rank(stat, 0);
          dst[(c * 3 + 0) * RS + r] = st0; dst[(c * 3 + 1) * RS + r] = st1; dst[(c * 3 + 2) * RS + r] = st2;
        }
      }
      cluster_sync_all();
      if (c == 0) {
        __syncthreads();
        for (int q = threadIdx.x; q < rows_here * K; q += blockDim.x) {
          const size_t g = (size_t)(p.row0 + set * RS) * K + q;
          p.Xval[g] = xs[q];
          if (MODE == M_VB) p.Xvar[g] = xvar[q];
          if (want_mu) { p.Xmu[g] = xmu[q]; p.Xtau[g] = xtau[q]; }
        }
      }
      gblk += 2 * NBK;
      __syncthreads();   // xs / tiles are reused by the next set
      PROF_ADD(15);      // statistics + store
      if (prof_on && lane == 0) p.prof[16] += 1;
    }
  }
  cluster_sync_all();    // no CTA may exit while its shared memory can still be written remotely
}

}  // namespace bnmtf
