// Cluster sweep with the pass running one column ahead of the conditionals ("look-ahead", opt-in: BNMTF_V2_LA=1).
//
// In the default cluster sweep (kernels_v2.cuh) the pass of column k needs delta_{k-1}: pass -> reduction -> exchange ->
// conditional -> delta form one dependency chain per column.  Here the pass of column k only needs delta_{k-2}.  With
// E = the residual corrected up to delta_{k-2} it produces
//     a_k = sum y_k^2,   h_k = sum E y_k,   g_k = sum y_{k-1} y_k          (three values per warp and column)
// and the control warp, which holds delta_{k-1} of its rows in registers, forms the numerator of the conditional from
//     sum e y_k = h_k - delta_{k-1} g_k                                     (e = E - delta_{k-1} y_{k-1}).
// So the compute warps work on column k while the conditional of column k-1 is evaluated: same Gauss-Seidel order and
// conditionals (bnmf_gibbs.py:203-210), a different association of one sum.  Costs: three gathers per entry and column
// instead of two, a 4-stage tile ring (tiles k-2, k-1, k and the prefetch), four record buffers (a CTA may send the
// records of column k+2 while a slower CTA still reads those of column k; k+4 is safe), deltas double-buffered and
// signalled by named barriers.  Gibbs and ICM only; no projection epilogue (NMF sweeps).
#pragma once
#include "kernels_v2.cuh"

namespace bnmtf {

__host__ __device__ inline size_t sweep2la_smem_doubles(int tw, int rs, int K, int cs, bool gibbs) {
  return (size_t)4 * (tw + V2_PAD) + 8 /*mbarriers: 0-3 tile stages, 4-7 records*/ + 2 * rs + 2 /*delta, double buffered*/ +
         (size_t)3 * rs * K /*xs, xmu, xtau*/ + (size_t)4 * cs * rs * 4 /*records, four buffers*/ + (size_t)cs * rs * 4 /*row statistics*/ +
         (gibbs ? (size_t)sweep2_rnd_window(rs) * rs * 6 : 0);
}

template <int NPT, int MODE, int CS>
__global__ void __launch_bounds__(Sweep2Cfg<NPT>::WPC / 2 * 32, 2) sweep2la_kernel(const Sweep2Params p) {
  static_assert(MODE == M_GIBBS || MODE == M_ICM, "look-ahead sweep: Gibbs and ICM");
  constexpr int RS = Sweep2Cfg<NPT>::WPC / 2 - 1;
  constexpr int NTHR = (RS + 1) * 32;
  constexpr int NST = 4;
  cg::cluster_group cluster = cg::this_cluster();
  const int c = (int)cluster.block_rank();
  const int cid = blockIdx.x / CS, nclusters = gridDim.x / CS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = p.K, tw = p.tw;
  const int tile_doubles = tw + V2_PAD;
  const uint32_t tile_bytes = (uint32_t)tw * 8u;

  extern __shared__ __align__(128) double smem2[];
  double* tiles = smem2;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(tiles + NST * tile_doubles);
  double* delta = reinterpret_cast<double*>(mbar + 8);    // [2][RS]
  double* xs = delta + 2 * RS + 2;                        // [RS][K]
  double* xmu = xs + (size_t)RS * K;
  double* xtau = xmu + (size_t)RS * K;
  double* part = xtau + (size_t)RS * K;                   // [4][CS][RS][4]: records of column k in buffer k & 3
  double* statp = part + (size_t)4 * CS * RS * 4;         // [CS][RS][4]
  double* rnd = statp + (size_t)CS * RS * 4;              // [RW][RS][6]
  const int total = 2 * K;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) mbar_init(&mbar[s], 1);
    fence_mbar_init();
  }
  for (int s = threadIdx.x; s < NST; s += blockDim.x)
    for (int z = 0; z < V2_PAD; ++z) tiles[s * tile_doubles + tw + z] = 0.0;
  __syncthreads();

  const double* ybase = p.Ykm + (size_t)c * tw;
  const size_t ystride = (size_t)p.ldy;
  auto tile_src = [&](int t) -> const double* { return ybase + (size_t)(t % K) * ystride; };
  auto stg = [&](uint32_t g) -> uint32_t { return g & 3u; };
  auto phs = [&](uint32_t g) -> uint32_t { return (g >> 2) & 1u; };
  // named barriers: 0 __syncthreads; 1, 4: "pass of an even / odd column done" (compute arrives, control syncs);
  // 2, 3: "deltas of an even / odd column written" (control arrives, compute syncs)
  auto bar_pass = [](int k) { return 1 + 3 * (k & 1); };
  auto bar_delta = [](int k) { return 2 + (k & 1); };

  if (warp == 0) {
    // =========================== control warp ===========================
    uint32_t gseq = 0;
    uint32_t rec_uses[4] = {0u, 0u, 0u, 0u};
    const UpdateCtx uc{p.seed, p.purpose, p.iter_ptr ? *p.iter_ptr : p.iteration, false};
    const double tau = p.scal[0];
    constexpr int RW = sweep2_rnd_window(RS);
    constexpr int CPR = (32 / RS) < 1 ? 1 : (32 / RS);
    int gen = 0;
    auto gen_round = [&](int rows_here, int64_t row_base) {
      const int cj = lane / RS, rr = lane - cj * RS;
      const int k = gen + cj;
      if (cj < CPR && k < K && rr < rows_here)
        tn_pre_material_store(uc.seed, uc.purpose, (uint64_t)(row_base + rr), (uint32_t)k, uc.iteration,
                              rnd + ((size_t)(k % RW) * RS + rr) * TN_PRE_DOUBLES);
      gen += CPR;
      __syncwarp();
    };
    for (int set = cid; set < p.nsets; set += nclusters) {
      const int rows_here = min(RS, p.n_loc - set * RS);
      for (int q = threadIdx.x; q < rows_here * K; q += blockDim.x) xs[q] = p.Xval[(size_t)(p.row0 + set * RS) * K + q];
      if (MODE == M_GIBBS) {
        gen = 0;
        while (gen < min(K, 1 + CPR)) gen_round(rows_here, p.row0 + (int64_t)set * RS);
      }
      if (lane == 0) {   // prime the ring: first NST tiles of this set's sequence (K residual + K update)
        fence_proxy_async();
        for (int q = 0; q < NST && q < total; ++q) {
          const uint32_t g = gseq + q;
          mbar_expect_tx(&mbar[stg(g)], tile_bytes);
          tma_load_1d(tiles + stg(g) * tile_doubles, tile_src(q), tile_bytes, &mbar[stg(g)]);
        }
      }
      cluster_sync_all();
      for (int q = 0; q < K; ++q) {          // residual pass: refill the stage each tile leaves
        __syncthreads();
        if (lane == 0 && q + NST < total) {
          const uint32_t g = gseq + q;
          fence_proxy_async();
          mbar_expect_tx(&mbar[stg(g)], tile_bytes);
          tma_load_1d(tiles + stg(g) * tile_doubles, tile_src(q + NST), tile_bytes, &mbar[stg(g)]);
        }
      }
      const int64_t grow = p.row0 + (int64_t)set * RS + lane;
      const bool sampler = (lane < rows_here);   // every CTA updates the RS rows redundantly (bitwise identical)
      double dlast = 0.0;                        // delta_{k-1} of this lane's row
      for (int k = 0; k < K; ++k) {
        double lam = 0.0, x_old = 0.0;
        TnPre mat;
        mat.z1 = mat.z2 = 0.0; mat.e1 = mat.l1 = mat.e2 = mat.l2 = 1.0;
        if (sampler) {
          lam = p.lamk ? p.lamk[k] : (p.lam_rm ? p.lam_rm[grow * K + k] : 0.0);
          x_old = xs[lane * K + k];
          if (MODE == M_GIBBS) mat = tn_pre_load(rnd + ((size_t)(k % RW) * RS + lane) * TN_PRE_DOUBLES);
        }
        if (MODE == M_GIBBS) {               // stay one round ahead of the updates
          __syncwarp();
          while (gen < min(K, k + 1 + CPR)) gen_round(rows_here, p.row0 + (int64_t)set * RS);
        }
        named_sync(bar_pass(k), NTHR);       // the compute warps of this CTA have finished their pass of column k
        if (lane == 0 && k >= 2 && K + k + 2 < total) {   // tile of column k-2 is dead: its stage takes column k+2
          const uint32_t gn = gseq + K + k + 2;
          fence_proxy_async();
          mbar_expect_tx(&mbar[stg(gn)], tile_bytes);
          tma_load_1d(tiles + stg(gn) * tile_doubles, tile_src(K + k + 2), tile_bytes, &mbar[stg(gn)]);
        }
        const int rb = k & 3;
        if (lane == 0) mbar_expect_tx(&mbar[4 + rb], (uint32_t)(CS * RS * 3 * 8));
        mbar_wait(&mbar[4 + rb], rec_uses[rb] & 1u);
        rec_uses[rb]++;
        if (sampler) {
          double a = 0.0, hh = 0.0, gg = 0.0;
#pragma unroll
          for (int cc = 0; cc < CS; ++cc) {
            const double* src = part + ((size_t)(rb * CS + cc) * RS + lane) * 4;
            a += src[0]; gg += src[1]; hh += src[2];
          }
          const double d = fma(-dlast, gg, hh);                        // sum_j e_j y_jk with e corrected by delta_{k-1}
          const double ct = tau * a;                                   // tauU, bnmf_gibbs.py:205
          const double num = -lam + tau * (d + x_old * a);             // numerator of muU, bnmf_gibbs.py:210
          double cm, x_new;
          if (MODE == M_GIBBS) x_new = tn_draw_lazy(num, ct, mat, uc.seed, uc.purpose, (uint64_t)grow, (uint32_t)k, uc.iteration, cm);
          else { cm = num / ct; x_new = fmax(fmax(0.0, cm), MINIMUM_TN); }   // nmf_icm.py:159-160
          xs[lane * K + k] = x_new;
          xmu[lane * K + k] = cm; xtau[lane * K + k] = ct;
          dlast = x_new - x_old;
          delta[(k & 1) * RS + lane] = dlast;
        }
        __syncwarp();
        named_arrive(bar_delta(k), NTHR);    // deltas of column k are in this CTA's shared memory
      }
      cluster_sync_all();                    // row-statistics partials are in CTA 0
      if (c == 0) {
        if (lane < rows_here && p.rowstats != nullptr) {
          double t3[3] = {0.0, 0.0, 0.0};
          for (int cc = 0; cc < CS; ++cc) {
            const double* src = statp + ((size_t)cc * RS + lane) * 4;
            t3[0] += src[0]; t3[1] += src[1]; t3[2] += src[2];
          }
          double* rs = p.rowstats + (size_t)(set * RS + lane) * NRS;
          rs[RS_RSS] = t3[0]; rs[RS_SUM_E] = t3[1]; rs[RS_SUM_RCE] = t3[2];
          rs[RS_ESDVAR] = 0.0; rs[RS_ELBOQ] = 0.0; rs[RS_PRIOR] = 0.0;
          rs[RS_IDIV] = 0.0; rs[RS_SPARE] = 0.0;
        }
        __syncthreads();
        for (int q = threadIdx.x; q < rows_here * K; q += blockDim.x) {
          const size_t g = (size_t)(p.row0 + set * RS) * K + q;
          p.Xval[g] = xs[q];
          if (p.Xmu != nullptr) { p.Xmu[g] = xmu[q]; p.Xtau[g] = xtau[q]; }
        }
      }
      gseq += total;
      __syncthreads();
    }
  } else {
    // =========================== compute warps ===========================
    const int r = warp - 1;                                 // row of the set owned by this warp
    double* stat0 = cluster.map_shared_rank(statp, 0);      // CTA 0's copy, through DSMEM
    const uint32_t part_u32 = smem_u32(part), rbar_u32 = smem_u32(&mbar[4]);
    uint32_t gseq = 0;
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
        cp[n / 2] = (a * 8u) | ((b * 8u) << 16);             // byte offsets inside a tile stage
      }
      for (int q = threadIdx.x; q < rows_here * K; q += blockDim.x) xs[q] = p.Xval[(size_t)(p.row0 + set * RS) * K + q];
      cluster_sync_all();   // xs visible; the previous set is finished everywhere

      // ---- residual of the segment: e = r - sum_k x_k y_k   (tile sequence 0 .. K-1)
      for (int q = 0; q < K; ++q) {
        const uint32_t g = gseq + q;
        mbar_wait(&mbar[stg(g)], phs(g));
        const char* __restrict__ T = reinterpret_cast<const char*>(tiles + stg(g) * tile_doubles);
        const double xk = (r < rows_here) ? xs[r * K + q] : 0.0;
#pragma unroll
        for (int n = 0; n < NPT; ++n) {
          const uint32_t off = (n & 1) ? (cp[n / 2] >> 16) : (cp[n / 2] & 0xFFFFu);
          e[n] = fma(-xk, *reinterpret_cast<const double*>(T + off), e[n]);
        }
        __syncthreads();
      }

      // ---- passes of the K columns, one column ahead of the conditionals   (tile sequence K .. 2K-1)
      for (int k = 0; k < K; ++k) {
        const uint32_t g = gseq + K + k;
        double d2 = 0.0;
        if (k >= 2) {
          named_sync(bar_delta(k - 2), NTHR);              // deltas of column k-2 are written
          d2 = (r < rows_here) ? delta[((k - 2) & 1) * RS + r] : 0.0;
        }
        mbar_wait(&mbar[stg(g)], phs(g));
        const char* __restrict__ T = reinterpret_cast<const char*>(tiles + stg(g) * tile_doubles);
        const char* __restrict__ T1 = reinterpret_cast<const char*>(tiles + stg(g + 3) * tile_doubles);   // column k-1
        const char* __restrict__ T2 = reinterpret_cast<const char*>(tiles + stg(g + 2) * tile_doubles);   // column k-2
        double sa = 0.0, sg = 0.0, sh = 0.0;
        if (k >= 2) {
#pragma unroll
          for (int n = 0; n < NPT; ++n) {
            const uint32_t off = (n & 1) ? (cp[n / 2] >> 16) : (cp[n / 2] & 0xFFFFu);
            e[n] = fma(-d2, *reinterpret_cast<const double*>(T2 + off), e[n]);
            const double v = *reinterpret_cast<const double*>(T + off);
            sa = fma(v, v, sa);
            sh = fma(e[n], v, sh);
            sg = fma(*reinterpret_cast<const double*>(T1 + off), v, sg);
          }
        } else {
#pragma unroll
          for (int n = 0; n < NPT; ++n) {
            const uint32_t off = (n & 1) ? (cp[n / 2] >> 16) : (cp[n / 2] & 0xFFFFu);
            const double v = *reinterpret_cast<const double*>(T + off);
            sa = fma(v, v, sa);
            sh = fma(e[n], v, sh);
            if (k == 1) sg = fma(*reinterpret_cast<const double*>(T1 + off), v, sg);
          }
        }
        named_arrive(bar_pass(k), NTHR);     // this warp no longer reads the tile of column k-2
        const double sah = warp_sum_pair(sa, sh, lane);   // lanes 0..15: sum of sa, lanes 16..31: sum of sh
        sg = warp_sum(sg);
        if ((lane & 15) < CS) {
          const uint32_t dst = map_to_cta(part_u32 + (uint32_t)(((k & 3) * CS + c) * RS + r) * 32u, lane & 15);
          const uint32_t bar = map_to_cta(rbar_u32 + (uint32_t)(k & 3) * 8u, lane & 15);
          st_async_f64(dst + ((lane & 16) ? 16u : 0u), sah, bar);
          if (lane < 16) st_async_f64(dst + 8, sg, bar);
        }
      }

      // ---- the last two corrections, then the row statistics of the final residual
      {
        double st0 = 0.0, st1 = 0.0, st2 = 0.0;
        double dA = 0.0, dB = 0.0;                         // deltas of columns K-2 and K-1
        const uint32_t gl = gseq + 2 * K - 1;              // tile of column K-1
        if (K >= 2) {
          named_sync(bar_delta(K - 2), NTHR);
          dA = (r < rows_here) ? delta[((K - 2) & 1) * RS + r] : 0.0;
        }
        named_sync(bar_delta(K - 1), NTHR);
        dB = (r < rows_here) ? delta[((K - 1) & 1) * RS + r] : 0.0;
        const char* __restrict__ TB = reinterpret_cast<const char*>(tiles + stg(gl) * tile_doubles);
        const char* __restrict__ TA = reinterpret_cast<const char*>(tiles + stg(gl + 3) * tile_doubles);
#pragma unroll
        for (int n = 0; n < NPT; ++n) {
          const uint32_t off = (n & 1) ? (cp[n / 2] >> 16) : (cp[n / 2] & 0xFFFFu);
          if (K >= 2) e[n] = fma(-dA, *reinterpret_cast<const double*>(TA + off), e[n]);
          e[n] = fma(-dB, *reinterpret_cast<const double*>(TB + off), e[n]);
          const double rr = p.tvals[seg + (int64_t)n * RS * 32];
          st0 = fma(e[n], e[n], st0);
          st1 += e[n];
          st2 = fma(rr - p.rmean, e[n], st2);   // padded slots hold e == 0
        }
        st0 = warp_sum(st0); st1 = warp_sum(st1); st2 = warp_sum(st2);
        if (lane == 0) {
          double* dst = stat0 + ((size_t)c * RS + r) * 4;
          dst[0] = st0; dst[1] = st1; dst[2] = st2;
        }
      }
      cluster_sync_all();
      if (c == 0) {
        __syncthreads();
        for (int q = threadIdx.x; q < rows_here * K; q += blockDim.x) {
          const size_t g = (size_t)(p.row0 + set * RS) * K + q;
          p.Xval[g] = xs[q];
          if (p.Xmu != nullptr) { p.Xmu[g] = xmu[q]; p.Xtau[g] = xtau[q]; }
        }
      }
      gseq += total;
      __syncthreads();   // xs / tiles are reused by the next set
    }
  }
  cluster_sync_all();    // no CTA may exit while its shared memory can still be written remotely
}

}  // namespace bnmtf
