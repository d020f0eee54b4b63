// Cluster / distributed-shared-memory sweep ("v2") for long rows.
//
// The v1 sweep (kernels.cuh) gathers the other factor straight from L2: one 32-byte
// sector per 8-byte value, 192 GB per sweep at the north-star size.  Here a
// thread-block cluster of CS CTAs owns a set of RS rows; CTA c owns column tile c
// (TW columns) of those rows:
//   * the residual of the RS x TW block (observed entries only) lives in registers,
//     one warp per row segment;
//   * column k of the other factor, restricted to the tile, is staged in shared
//     memory by 1-D TMA bulk copies (cp.async.bulk + mbarrier) through a 3- or 4-stage
//     ring and gathered from there (entries are placed so that a half-warp hits 16
//     distinct 8-byte banks; padded slots read a zero of their own bank class);
//   * the two per-row partial sums of a warp are reduced together (warp_sum_pair) and
//     sent to EVERY CTA's record buffer by st.async with mbarrier completion; the
//     control warp of each CTA sums the CS records in a fixed order and updates the RS
//     rows redundantly (lane = row, bitwise identical in all CTAs) -- no cluster
//     barrier and no cluster-scope fence per column;
//   * the rank-1 correction of column k is folded into the accumulation loop of
//     column k+1 ("lazy"), so nothing but e[] has to stay in registers.
// L2 -> SM traffic drops from 192 GB to ~31 GB per sweep.  What bounds the kernel is
// the per-column dependency chain (profiles/r01_final_sweep_ncu.md).  The variants measured
// slower in round 1 (tensor memory as a scratch area, two rows per warp, clusters of 16,
// the look-ahead pass) are recorded in profiles/r01_final_sweep_ncu.md and were removed.
#pragma once
#include <cooperative_groups.h>
#include <cstdint>
#include <cuda_runtime.h>
#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace bnmtf {

// ---- mbarrier / TMA bulk helpers (PTX) ---------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// ---- layout statistics: observed entries per row and the longest (row, tile) segment
static __global__ void count_rows_tiles_kernel(const uint8_t* __restrict__ M, int64_t n, int64_t m, int64_t ld, int tw,
                                        int* __restrict__ rowlen, int* __restrict__ maxseg) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  int total = 0, seg = 0, best = 0;
  for (int64_t j0 = 0; j0 < m; j0 += 32) {
    if (j0 % tw == 0) { best = max(best, seg); seg = 0; }
    const int64_t j = j0 + lane;
    const bool obs = (j < m) && (M[row * ld + j] != 0);
    const int c = __popc(__ballot_sync(0xffffffffu, obs));
    seg += c; total += c;
  }
  best = max(best, seg);
  if (lane == 0) { rowlen[row] = total; atomicMax(maxseg, best); }
}

// Padded slots read a zero behind the tile.  There are V2_PAD zero entries and the pad of lane l points at
// entry padcol + l % 16, i.e. at the bank class the lane's real entries have: pads never collide with the
// gathers of the other lanes (a single shared zero sat in one bank class and cost every mixed half-warp a
// replay: 25 % extra shared-memory wavefronts at the 20000 x 20000 size, 4 % with per-class zeros).
constexpr int V2_PAD = 16;
static __global__ void fill_pads_kernel(double* __restrict__ tvals, uint16_t* __restrict__ tcols, int64_t nslots, uint16_t padcol) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nslots; q += (int64_t)gridDim.x * blockDim.x) {
    tvals[q] = 0.0;
    tcols[q] = (uint16_t)(padcol + (q & 15));   // slot index = ... * 32 + lane
  }
}

// Slab layout: slot(set, c, n, r, lane) = ((((set*CS + c)*NPT + n)*RS + r)*32 + lane.
// One warp packs one (row, tile) segment.  Placement: the t-th observed entry (in
// column order) of bank class q = col_local % 16 goes to n = t/2, lane = q + 16*(t%2),
// so each half-warp of a gather instruction reads 16 distinct 8-byte banks; classes
// with more than 2*NPT entries spill, in canonical order, into the slots that
// shorter classes leave empty.  The placement depends only on the segment and NPT.
static __global__ void pack_tiles_kernel(const double* __restrict__ R, const uint8_t* __restrict__ M, int64_t n, int64_t m, int64_t ld,
                                  int cs, int tw, int rs, int npt, double* __restrict__ tvals, uint16_t* __restrict__ tcols) {
  const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (wid >= n * cs) return;
  const int64_t row = wid / cs;
  const int c = (int)(wid - row * cs);
  const int64_t set = row / rs;
  const int r = (int)(row - set * rs);
  const int64_t seg_base = ((set * cs + c) * npt) * (int64_t)rs * 32;
  const int64_t j_begin = (int64_t)c * tw, j_end = min(m, j_begin + tw);
  const int q = lane & 15;
  const int cap = 2 * npt;
  // pass 1: class totals
  int pc = 0;
  for (int64_t j0 = j_begin; j0 < j_end; j0 += 32) {
    const int64_t j = j0 + lane;
    pc += (j < j_end) && (M[row * ld + j] != 0);
  }
  const int cnt = pc + __shfl_xor_sync(0xffffffffu, pc, 16);   // entries of class q in the segment
  const int over = max(0, cnt - cap), holes = max(0, cap - cnt);
  // exclusive prefix sums over the 16 classes (both half-warps compute the same)
  int ovbase = over, holebase = holes;
  for (int o = 1; o < 16; o <<= 1) {
    const int a = __shfl_up_sync(0xffffffffu, ovbase, o, 16), b = __shfl_up_sync(0xffffffffu, holebase, o, 16);
    if (q >= o) { ovbase += a; holebase += b; }
  }
  ovbase -= over; holebase -= holes;
  // pass 2: placement
  int seen = 0;  // entries of this lane's column residue (mod 32) seen so far
  for (int64_t j0 = j_begin; j0 < j_end; j0 += 32) {
    const int64_t j = j0 + lane;
    const bool obs = (j < j_end) && (M[row * ld + j] != 0);
    const unsigned bal = __ballot_sync(0xffffffffu, obs);
    const int partner_seen = __shfl_xor_sync(0xffffffffu, seen, 16);
    int t = seen + partner_seen + ((lane >= 16) ? (int)((bal >> (lane - 16)) & 1u) : 0);
    int qq = q;
    const bool overflow = obs && (t >= cap);
    if (__any_sync(0xffffffffu, overflow)) {
      // canonical overflow index -> the o-th empty slot, in class order
      const int o = ovbase + (t - cap);
      int found_q = 0, found_t = 0;
      for (int z = 0; z < 16; ++z) {
        const int hb = __shfl_sync(0xffffffffu, holebase, z), hh = __shfl_sync(0xffffffffu, holes, z);
        const int cz = __shfl_sync(0xffffffffu, cnt, z);
        if (overflow && o >= hb && o < hb + hh) { found_q = z; found_t = cz + (o - hb); }
      }
      if (overflow) { qq = found_q; t = found_t; }
    }
    if (obs) {
      const int nn = t >> 1, ll = qq + 16 * (t & 1);
      const int64_t slot = seg_base + ((int64_t)nn * rs + r) * 32 + ll;
      tvals[slot] = R[row * ld + j];
      tcols[slot] = (uint16_t)(j - j_begin);
    }
    seen += obs ? 1 : 0;
  }
}

struct Sweep2Params {
  const double* __restrict__ tvals;
  const uint16_t* __restrict__ tcols;
  int nsets;
  int n_loc;
  int64_t row0;
  int K;
  int tw;                                // tile width in columns (multiple of 32)
  int rs;                                // rows per set == warps per CTA
  const double* __restrict__ Ykm;        // other factor, k-major [K][ldy][NV], zero padded to >= CS*tw
  int64_t ldy;
  double* Xval;
  double* Xvar;
  double* Xmu;
  double* Xtau;
  const double* __restrict__ lam_rm;
  const double* __restrict__ lamk;
  const double* __restrict__ scal;
  double* rowstats;                      // [n_loc][NRS]
  double rmean;
  uint64_t seed;
  uint32_t iteration;
  const uint32_t* iter_ptr;              // when set (graph replays) the iteration index is read from the device
  uint32_t purpose;
  // optional epilogue (NMTF F sweep): project the final residual onto the K2 columns of a second k-major factor,
  // proj_out[global row][l] = sum_j e_j Y2[l][j]   (the row part of the S_kl numerators, bnmtf_gibbs.py:277-283)
  const double* __restrict__ Ykm2;
  int64_t ldy2;
  int K2;
  double* proj_out;
  // kernels_v5.cuh: the second factor of the projection epilogue in the BLOCK-major layout [ceil(K2/2)][ldp][2] (the k-major
  // Ykm2 above belongs to kernels_v2.cuh; in VB mode Ykm2 holds the second moments of the other factor)
  const double* __restrict__ Yproj;
  int64_t ldp;
  // kernels_v5.cuh, BNMTF VB covariance term of update_F / update_G (bnmtf_vb.py:343,371), as SweepParams (kernels.cuh):
  //   cov_i = sum_l S(k,l) A_il ((X S)_il - x_ik S(k,l))
  const double* __restrict__ covA;       // [global row][covL] or nullptr
  const double* __restrict__ covS;       // S(k,l) = covS[k*cov_sk + l*cov_sl]
  int covL, cov_sk, cov_sl;
  double* gmat;                          // blocked cluster sweep (kernels_v4.cuh), Gibbs: scratch for the random material of the resident row sets
  int* set_counter;                      // kernels_v5.cuh: row sets handed out so far beyond the first round (zeroed before every launch)
  // multi-GPU, kernels_v5.cuh: the rows this rank updates are also written straight into the factor / statistics arrays of
  // the other ranks (peer memory over NVLink) -- the all-gather of the sweep's output rides on the sweep itself
  double* const* peer_val;               // [npeers] the other ranks' Xval
  double* const* peer_var;               // [npeers] Xvar (VB)
  double* const* peer_rowstats;          // [npeers] row statistics, indexed by GLOBAL row
  int npeers;
  int nst;                               // stages of the tile ring (3 or 4)
  int dbg;                               // timing experiments only: 1 = no update math, 2 = phase clocks into prof
  unsigned long long* prof;              // [32] clock64 sums of cluster 0 / CTA 0 (dbg & 2)
};

// dynamic shared memory of sweep2_kernel, in doubles
// columns of prepared random material kept in shared memory (Gibbs): one round of 32/rs columns ahead of the updates
__host__ __device__ constexpr int sweep2_rnd_window(int rs) { return (1 + 2 * ((32 / rs) < 1 ? 1 : (32 / rs))) <= 8 ? 8 : 16; }
__host__ __device__ inline size_t sweep2_smem_doubles(int tw, int nv, int rs, int K, int cs, int K2 = 0, bool gibbs = false, int nst = 3) {
  return (size_t)nst * (tw + V2_PAD) * nv + 8 /*mbarrier slots: 0-3 tile stages, 4-5 record buffers*/ + rs + 1 /*delta*/ + (size_t)4 * rs * K /*xs,xvar,xmu,xtau*/ +
         (size_t)2 * cs * rs * 4 /*part, double buffered*/ + (size_t)cs * rs * K2 /*projection partials*/ +
         (size_t)cs * rs * 4 /*row-statistics partials*/ + (gibbs ? (size_t)sweep2_rnd_window(rs) * rs * 6 : 0) /*random material*/;
}

// Warps per CTA that the register budget of NPT entries per lane allows.  Warp 0 of
// every CTA is the control warp (TMA producer; in CTA 0 also the updater of the RS
// rows of the set, lane = row); warps 1..WPC-1 each own one row segment, so a set
// has RS = WPC - 1 rows.  The two roles run separate loop nests with matching barrier
// sequences, so the compute warps never carry a function call in their loop.
// WPC is the largest block the register file allows (1 CTA per SM); smaller blocks
// (WPC/2, WPC/3: 2 or 3 CTAs of different clusters per SM) let one cluster's barrier
// and update latency overlap with another cluster's gathers.
#ifndef BNMTF_V2_WPC20
#define BNMTF_V2_WPC20 24   // warps per SM at 20 entries per lane
#endif
template <int NPT>
struct Sweep2Cfg {
  static constexpr int WPC = NPT <= 16 ? 32 : (NPT <= 18 ? 28 : BNMTF_V2_WPC20);
};

// Remote (DSMEM) store that signals the destination CTA's mbarrier when it lands: the receiver waits on its own
// mbarrier for the expected byte count -- no cluster barrier, no cluster/GPU-scope fence on either side.
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void st_async_f64(uint32_t remote_addr, double v, uint32_t remote_mbar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr),
               "l"(__double_as_longlong(v)), "r"(remote_mbar)
               : "memory");
}

__device__ __forceinline__ void named_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}


// phase clocks of cluster 0 / CTA 0 (timing experiments: build with -DBNMTF_V2_PROF, run with BNMTF_V2_DBG=2)
#ifdef BNMTF_V2_PROF
#define P2_T0() long long t_prof = prof_on ? clock64() : 0
#define P2_ADD(slot)                                                       \
  do {                                                                     \
    if (prof_on) {                                                         \
      const long long t_now = clock64();                                   \
      if (lane == 0) p.prof[slot] += (unsigned long long)(t_now - t_prof); \
      t_prof = t_now;                                                      \
    }                                                                      \
  } while (0)
#else
#define P2_T0() do { } while (0)
#define P2_ADD(slot) do { } while (0)
#endif

template <int NPT, int MODE, int CS, int DIV>
__global__ void __launch_bounds__(Sweep2Cfg<NPT>::WPC / DIV * 32, DIV) sweep2_kernel(const Sweep2Params p) {
  constexpr int NV = (MODE == M_VB) ? 2 : 1;
  constexpr int RS = Sweep2Cfg<NPT>::WPC / DIV - 1;
  constexpr int NTHR = (RS + 1) * 32;                     // threads of the CTA
  static_assert(MODE != M_NP, "the NP updates run on the v1 kernel");
  cg::cluster_group cluster = cg::this_cluster();
  const int c = (int)cluster.block_rank();
  const int cid = blockIdx.x / CS, nclusters = gridDim.x / CS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = p.K, tw = p.tw;
  const int tile_doubles = (tw + V2_PAD) * NV;
  const uint32_t tile_bytes = (uint32_t)tw * NV * 8u;

  extern __shared__ __align__(128) double smem2[];
  double* tiles = smem2;                                  // NST stages
  const int NST = p.nst;                                  // stages of the tile ring
  auto stg = [&](uint32_t g) -> uint32_t { return NST == 4 ? (g & 3u) : (g % 3u); };
  auto phs = [&](uint32_t g) -> uint32_t { return (NST == 4 ? (g >> 2) : (g / 3u)) & 1u; };
  uint64_t* mbar = reinterpret_cast<uint64_t*>(tiles + NST * tile_doubles);
  double* delta = reinterpret_cast<double*>(mbar + 8);    // [RS]
  double* xs = delta + RS + 1;                            // [RS][K]
  double* xvar = xs + (size_t)RS * K;
  double* xmu = xvar + (size_t)RS * K;
  double* xtau = xmu + (size_t)RS * K;
  double* part = xtau + (size_t)RS * K;                   // [2][CS][RS][4]: partial sums, double buffered by column parity
  double* proj = part + (size_t)2 * CS * RS * 4;          // [CS][RS][K2]: projection partials (meet in CTA 0)
  const int K2 = p.K2;
  double* rnd = proj + (size_t)CS * RS * K2 + (size_t)CS * RS * 4;   // [RW][RS][6]: prepared random material (Gibbs)
  double* statp = proj + (size_t)CS * RS * K2;            // [CS][RS][4]: row-statistics partials (own region: the last
                                                          // column's records in `part` may still be read when they arrive)
  const int total = 2 * K + K2;                           // tiles per set: residual pass, update pass, projection epilogue

  if (threadIdx.x == 0) {
    for (int s = 0; s < 6; ++s) mbar_init(&mbar[s], 1);   // 0-3: tile stages, 4-5: records of even / odd columns
    fence_mbar_init();
  }
  for (int s = threadIdx.x; s < NST; s += blockDim.x) {   // the zero slots read by padded entries
    for (int z = 0; z < V2_PAD * NV; ++z) tiles[s * tile_doubles + tw * NV + z] = 0.0;
  }
  __syncthreads();

  const double* ybase = p.Ykm + (size_t)c * tw * NV;      // this CTA's tile of column 0
  const size_t ystride = (size_t)p.ldy * NV;
  const double* ybase2 = K2 > 0 ? p.Ykm2 + (size_t)c * tw * NV : nullptr;
  const size_t ystride2 = (size_t)p.ldy2 * NV;
  auto tile_src = [&](int t) -> const double* {           // source of tile t of a set's sequence
    return t < 2 * K ? ybase + (size_t)(t % K) * ystride : ybase2 + (size_t)(t - 2 * K) * ystride2;
  };

  if (warp == 0) {
    // =========================== control warp ===========================
    uint32_t gseq = 0;
    uint32_t rec_phase[2] = {0u, 0u};
    const UpdateCtx uc{p.seed, p.purpose, p.iter_ptr ? *p.iter_ptr : p.iteration, false};
    const double tau = p.scal[0];
    // Gibbs: the random material of upcoming columns is prepared lane-parallel (lane = (column, row)) while the
    // compute warps run their pass; the conditional itself then costs an rsqrt and a compare (tn_draw_fast)
    constexpr int RW = sweep2_rnd_window(RS);
    constexpr int CPR = (32 / RS) < 1 ? 1 : (32 / RS);
    static_assert(1 + 2 * CPR <= RW || MODE != M_GIBBS, "random-material window too small");
    int gen = 0;
#ifdef BNMTF_V2_PROF
    const bool prof_on = (p.dbg & 2) && cid == 0 && c == 0 && p.prof != nullptr;
#endif
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
      P2_T0();
      for (int q = threadIdx.x; q < rows_here * K; q += blockDim.x) {
        const size_t g = (size_t)(p.row0 + set * RS) * K + q;
        xs[q] = p.Xval[g];
      }
      if (MODE == M_GIBBS) {
        gen = 0;
        while (gen < min(K, 1 + CPR)) gen_round(rows_here, p.row0 + (int64_t)set * RS);
      }
      if (lane == 0) {   // prime the ring: first NST tiles of this set's sequence (K init + K sweep)
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
      double acc_esd = 0.0, acc_q = 0.0, acc_prior = 0.0;
      P2_ADD(0);                             // set start + residual pass (control side)
      const int64_t grow = p.row0 + (int64_t)set * RS + lane;
      const bool sampler = (lane < rows_here);   // every CTA updates the RS rows redundantly (bitwise identical)
      for (int k = 0; k < K; ++k) {
        // everything that does not need the row sums happens before they arrive
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
        P2_ADD(0);                           // preparation + random material
        named_sync(1, NTHR);                 // the compute warps of this CTA have finished their pass of column k
        P2_ADD(1);
        if (lane == 0 && k >= 1 && K + k - 1 + NST < total) {   // tile of column k-1 is dead: its stage takes tile k-1+NST
          const uint32_t gn = gseq + K + k - 1 + NST;
          fence_proxy_async();
          mbar_expect_tx(&mbar[stg(gn)], tile_bytes);
          tma_load_1d(tiles + stg(gn) * tile_doubles, tile_src(K + k - 1 + NST), tile_bytes, &mbar[stg(gn)]);
        }
        // the partial sums of column k: CS x RS warps each send 2 (VB: 3) doubles by st.async; they complete this CTA's mbarrier
        if (lane == 0) mbar_expect_tx(&mbar[4 + (k & 1)], (uint32_t)(CS * RS * (MODE == M_VB ? 3 : 2) * 8));
        mbar_wait(&mbar[4 + (k & 1)], rec_phase[k & 1] & 1u);
        rec_phase[k & 1]++;
        P2_ADD(2);                           // waiting for the records of all CTAs
        if (sampler) {
          double t3[3] = {0.0, 0.0, 0.0};
#pragma unroll
          for (int cc = 0; cc < CS; ++cc) {
            const double* src = part + ((size_t)((k & 1) * CS + cc) * RS + lane) * 4;
            t3[0] += src[0]; t3[2] += src[2];
            if (MODE == M_VB) t3[1] += src[1];   // only the VB sweeps send the second sum
          }
          double ct, cm, vn = 0.0, x_new;
          if (p.dbg & 1) { x_new = x_old + 1e-3 * t3[0]; ct = cm = 0.0; }
          else if (MODE == M_GIBBS) {
            ct = tau * t3[0];                                            // tauU, bnmf_gibbs.py:205
            const double num = -lam + tau * (t3[2] + x_old * t3[0]);     // numerator of muU, bnmf_gibbs.py:210
            x_new = tn_draw_lazy(num, ct, mat, uc.seed, uc.purpose, (uint64_t)grow, (uint32_t)k, uc.iteration, cm);
          }
          else x_new = conditional_update<MODE>(t3, x_old, lam, tau, uc, (uint64_t)grow, k, 0.0, 0.0, 0.0, ct, cm, vn, acc_esd,
                                                acc_q, acc_prior);
          xs[lane * K + k] = x_new;
          xmu[lane * K + k] = cm; xtau[lane * K + k] = ct;
          if (MODE == M_VB) xvar[lane * K + k] = vn;
          delta[lane] = x_new - x_old;
        }
        P2_ADD(3);                           // record sums + conditional
        __syncthreads();                     // deltas of column k are in this CTA's shared memory
        P2_ADD(4);
      }
      if (K2 > 0) {                          // projection epilogue: keep the ring fed (tiles 2K .. 2K+K2-1)
        __syncthreads();                     // every warp has applied the last correction: the stage of tile 2K-1 is free
        if (lane == 0 && 2 * K - 1 + NST < total && K >= 2) {
          const uint32_t gn = gseq + 2 * K - 1 + NST;
          fence_proxy_async();
          mbar_expect_tx(&mbar[stg(gn)], tile_bytes);
          tma_load_1d(tiles + stg(gn) * tile_doubles, tile_src(2 * K - 1 + NST), tile_bytes, &mbar[stg(gn)]);
        }
        for (int q = 0; q < K2; ++q) {
          __syncthreads();
          if (lane == 0 && 2 * K + q + NST < total) {
            const uint32_t gn = gseq + 2 * K + q + NST;
            fence_proxy_async();
            mbar_expect_tx(&mbar[stg(gn)], tile_bytes);
            tma_load_1d(tiles + stg(gn) * tile_doubles, tile_src(2 * K + q + NST), tile_bytes, &mbar[stg(gn)]);
          }
        }
      }
      cluster_sync_all();                    // row-statistics (and projection) partials are in CTA 0
      if (c == 0) {
        if (lane < rows_here && p.rowstats != nullptr) {
          double t3[3] = {0.0, 0.0, 0.0};
          for (int cc = 0; cc < CS; ++cc) {
            const double* src = statp + ((size_t)cc * RS + lane) * 4;
            t3[0] += src[0]; t3[1] += src[1]; t3[2] += src[2];
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
          if (p.Xmu != nullptr) { p.Xmu[g] = xmu[q]; p.Xtau[g] = xtau[q]; }
        }
        for (int q = threadIdx.x; q < rows_here * K2; q += blockDim.x) {   // fixed-order sum of the CS projection partials
          const int rr = q / K2, l = q - rr * K2;
          double t = 0.0;
          for (int cc = 0; cc < CS; ++cc) t += proj[((size_t)cc * RS + rr) * K2 + l];
          p.proj_out[(size_t)(p.row0 + set * RS + rr) * K2 + l] = t;
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
#ifdef BNMTF_V2_PROF
    const bool prof_on = (p.dbg & 2) && cid == 0 && c == 0 && warp == 1 && p.prof != nullptr;
#endif
    for (int set = cid; set < p.nsets; set += nclusters) {
      const int rows_here = min(RS, p.n_loc - set * RS);
      const int64_t seg = ((int64_t)(set * CS + c) * NPT) * RS * 32 + (int64_t)r * 32 + lane;
      P2_T0();
#ifdef BNMTF_V2_PROF
      if (prof_on && lane == 0) p.prof[16] += 1;
#endif
      double e[NPT];
      uint32_t cp[NPT / 2];
#pragma unroll
      for (int n = 0; n < NPT; ++n) e[n] = p.tvals[seg + (int64_t)n * RS * 32];
#pragma unroll
      for (int n = 0; n < NPT; n += 2) {
        const uint32_t a = p.tcols[seg + (int64_t)n * RS * 32], b = p.tcols[seg + (int64_t)(n + 1) * RS * 32];
        cp[n / 2] = (a * (NV * 8)) | ((b * (NV * 8)) << 16);   // byte offsets inside a tile stage
      }
      for (int q = threadIdx.x; q < rows_here * K; q += blockDim.x) {
        const size_t g = (size_t)(p.row0 + set * RS) * K + q;
        xs[q] = p.Xval[g];
      }
      cluster_sync_all();   // xs visible; CTA 0 has finished with `part` of the previous set
      P2_ADD(8);            // segment loads + set start barrier

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

      P2_ADD(9);            // residual pass
      // ---- the K column updates   (tile sequence K .. 2K-1)
      double dprev = 0.0;
      for (int k = 0; k < K; ++k) {
        const uint32_t g = gseq + K + k;
        mbar_wait(&mbar[stg(g)], phs(g));
        P2_ADD(10);                          // waiting for the tile
        const char* __restrict__ T = reinterpret_cast<const char*>(tiles + stg(g) * tile_doubles);
        const char* __restrict__ Tp = reinterpret_cast<const char*>(tiles + stg(g + NST - 1) * tile_doubles);   // tile g-1
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        if (k > 0) {
#pragma unroll
          for (int n = 0; n < NPT; ++n) {
            const uint32_t off = (n & 1) ? (cp[n / 2] >> 16) : (cp[n / 2] & 0xFFFFu);
            e[n] = fma(-dprev, *reinterpret_cast<const double*>(Tp + off), e[n]);
            double v;
            if (MODE == M_VB) { const double2 vv = *reinterpret_cast<const double2*>(T + off); v = vv.x; s1 += vv.y; }
            else v = *reinterpret_cast<const double*>(T + off);
            s0 = fma(v, v, s0);
            s2 = fma(e[n], v, s2);
          }
        } else {
#pragma unroll
          for (int n = 0; n < NPT; ++n) {
            const uint32_t off = (n & 1) ? (cp[n / 2] >> 16) : (cp[n / 2] & 0xFFFFu);
            double v;
            if (MODE == M_VB) { const double2 vv = *reinterpret_cast<const double2*>(T + off); v = vv.x; s1 += vv.y; }
            else v = *reinterpret_cast<const double*>(T + off);
            s0 = fma(v, v, s0);
            s2 = fma(e[n], v, s2);
          }
        }
        P2_ADD(11);                          // pass
        named_arrive(1, (RS + 1) * 32);      // this warp no longer reads the tile of column k-1
        const double s02 = warp_sum_pair(s0, s2, lane);   // lanes 0..15: sum of s0, lanes 16..31: sum of s2
        if (MODE == M_VB) s1 = warp_sum(s1);
        if ((lane & 15) < CS) {   // lanes cc / 16 + cc send this warp's partial sums to CTA cc (every CTA updates redundantly)
          const uint32_t dst = map_to_cta(part_u32 + (uint32_t)(((k & 1) * CS + c) * RS + r) * 32u, lane & 15);
          const uint32_t bar = map_to_cta(rbar_u32 + (uint32_t)(k & 1) * 8u, lane & 15);
          st_async_f64(dst + ((lane & 16) ? 16u : 0u), s02, bar);
          if (MODE == M_VB && lane < 16) st_async_f64(dst + 8, s1, bar);
        }
        P2_ADD(12);                          // reduction + record stores
        __syncthreads();                     // the control warp of this CTA has written the deltas
        dprev = (r < rows_here) ? delta[r] : 0.0;
        P2_ADD(14);                          // waiting for the deltas
      }

      // ---- last correction, then the row statistics of the final residual
      {
        const uint32_t g = gseq + 2 * K - 1;
        const char* __restrict__ Tp = reinterpret_cast<const char*>(tiles + stg(g) * tile_doubles);
        double st0 = 0.0, st1 = 0.0, st2 = 0.0;
#pragma unroll
        for (int n = 0; n < NPT; ++n) {
          const uint32_t off = (n & 1) ? (cp[n / 2] >> 16) : (cp[n / 2] & 0xFFFFu);
          e[n] = fma(-dprev, *reinterpret_cast<const double*>(Tp + off), e[n]);
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
      if (K2 > 0) {
        // ---- projection epilogue: sum_j e_j Y2[l][j] over this segment, partials to CTA 0   (tiles 2K .. 2K+K2-1)
        double* proj0 = cluster.map_shared_rank(proj, 0);
        __syncthreads();
        for (int q = 0; q < K2; ++q) {
          const uint32_t g = gseq + 2 * K + q;
          mbar_wait(&mbar[stg(g)], phs(g));
          const char* __restrict__ T = reinterpret_cast<const char*>(tiles + stg(g) * tile_doubles);
          double s = 0.0;
#pragma unroll
          for (int n = 0; n < NPT; ++n) {
            const uint32_t off = (n & 1) ? (cp[n / 2] >> 16) : (cp[n / 2] & 0xFFFFu);
            s = fma(e[n], *reinterpret_cast<const double*>(T + off), s);
          }
          s = warp_sum(s);
          if (lane == 0) proj0[((size_t)c * RS + r) * K2 + q] = s;
          __syncthreads();
        }
      }
      cluster_sync_all();
      if (c == 0) {
        __syncthreads();
        for (int q = threadIdx.x; q < rows_here * K; q += blockDim.x) {
          const size_t g = (size_t)(p.row0 + set * RS) * K + q;
          p.Xval[g] = xs[q];
          if (MODE == M_VB) p.Xvar[g] = xvar[q];
          if (p.Xmu != nullptr) { p.Xmu[g] = xmu[q]; p.Xtau[g] = xtau[q]; }
        }
        for (int q = threadIdx.x; q < rows_here * K2; q += blockDim.x) {   // fixed-order sum of the CS projection partials
          const int rr = q / K2, l = q - rr * K2;
          double t = 0.0;
          for (int cc = 0; cc < CS; ++cc) t += proj[((size_t)cc * RS + rr) * K2 + l];
          p.proj_out[(size_t)(p.row0 + set * RS + rr) * K2 + l] = t;
        }
      }
      gseq += total;
      __syncthreads();   // xs / tiles are reused by the next set
      P2_ADD(15);        // statistics, store, end of set
    }
  }
  cluster_sync_all();    // no CTA may exit while its shared memory can still be written remotely
}

}  // namespace bnmtf
