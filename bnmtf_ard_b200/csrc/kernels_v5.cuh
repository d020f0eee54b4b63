// Cluster sweep with one dependency chain PER ROW ("v5").
//
// kernels_v2.cuh / kernels_v4.cuh run the rows of a set in lockstep: every compute warp of a CTA does its pass, then all
// of them wait while the records cross the cluster, the control warp draws for the whole set and the deltas come back.
// The shared-memory pipe -- the resource the sweep is bound by (profiles/r01_final_sweep_ncu.md) -- idles during every
// exchange, and only the second CTA of the SM fills the gaps (41-47 % wavefront utilisation, 8.1-8.5 ms per sweep at
// 20000 x 20000, K = 50).  Here nothing is shared between the rows of a CTA except the tile ring:
//
//   * one warp owns one row segment for the whole row (20 entries per lane in registers, as before) AND runs the row's
//     conditionals itself, redundantly in its 32 lanes and -- bitwise identically -- in the 8 CTAs of the cluster:
//     pass -> warp reduction -> st.async of the two sums to the row's mailbox in every CTA -> mbarrier wait on the own
//     mailbox (8 records) -> conditional -> rank-1 correction, with the deltas already in registers.  No control warp,
//     no named barrier, no delta hand-off through shared memory on the chain;
//   * the rows of a CTA drift apart by up to a tile-ring slot, so while one warp waits for its records the others keep
//     the shared-memory pipe busy.  The tile ring is a plain producer / consumer pipeline: warp 0 refills a slot by
//     TMA (cp.async.bulk) when all row warps have released it (`empty` mbarrier, one arrival per row warp);
//   * as in kernels_v4.cuh two columns share a pass and an exchange (B = 2, cross term g = sum y_k y_k+1), and the sums
//     that do not depend on the residual (a_k = sum M y^2, g, VB: q_k = sum M (Var y + E[y]^2); bnmf_gibbs.py:203-210,
//     bnmf_vb.py:251-255) come from the residual pass that opens a row.  The 8 warps of a row meet once per row on a
//     per-row mbarrier (remote arrivals, release / acquire at cluster scope), read each other's partial statistics through
//     distributed shared memory and each builds the row's table (c1, sigma | c2, precision) lane-parallel: the
//     conditional on the chain is a handful of FMAs;
//   * Gibbs: the random material of a row (rng.cuh: tn_pre_material) is made by the row's warp, one column per lane,
//     while the tiles of the residual pass arrive; the two normals every draw starts from sit in the table, the
//     exponential-proposal variates in an L2-resident scratch table.
//
// Same conditionals and the same Gauss-Seidel order as bnmf_gibbs.py:155-164 / nmf_icm.py:155-167 / bnmf_vb.py:167-174.
#pragma once
#include "kernels_v2.cuh"

namespace bnmtf {

constexpr int V5_B = 2;          // columns per block
#ifndef BNMTF_V5_CH
#define BNMTF_V5_CH 2
#endif

struct Sweep5Layout {
  size_t tiles, mbar, misc, stat_own, tab, gsum, part, cov, prof, total;   // offsets in doubles
  int covstride;                 // doubles per row of the covariance area: (fsv_l, A_l) pairs
  int nstat;                     // statistics per row: a[KP], g[NBK], (VB) q[KP]
  int sstride;                   // doubles per row of the statistics area
};

// doubles per CTA of the record a row's warps send to CTA 0 at the row's end: 3 statistics of the final residual (+ 1 pad) and,
// with the projection epilogue, the partial sums sum_j e_j Y2[l][j] of the K2 columns of the second factor (in blocks of two)
__host__ __device__ constexpr int sweep5_rec_width(int K2) { return 4 + 2 * ((K2 + V5_B - 1) / V5_B); }

__host__ __device__ inline Sweep5Layout sweep5_layout(int tw, int rs, int K, int cs, int nslot, int tabw, bool vb, int K2 = 0, int covL = 0) {
  Sweep5Layout L;
  const int NBK = (K + V5_B - 1) / V5_B, KP = NBK * V5_B;
  L.nstat = KP + NBK + (vb ? KP : 0);
  // the row's area also receives the records of the final residual ([CS][sweep5_rec_width], CTA 0 only) once the row's table is built
  const int recw = cs * sweep5_rec_width(K2);
  L.sstride = (L.nstat > recw ? L.nstat : recw) + ((L.nstat > recw ? L.nstat : recw) & 1);
  size_t o = 0;
  L.tiles = o; o += (size_t)nslot * (tw + V2_PAD) * V5_B;
  L.mbar = o; o += (size_t)2 * nslot + 4 * rs;          // full[nslot], empty[nslot], rec[rs][2], statbar[rs], rowbar[rs]
  L.misc = o; o += 8;                                   // tau, tensor-memory base, set queue: 4 entries (2 doubles) + 4 mbarriers
  L.stat_own = o; o += (size_t)rs * L.sstride;          // this CTA's tile part of the row statistics (read by the whole cluster)
  L.tab = o; o += (size_t)rs * KP * tabw;               // per (row, column): what the conditional needs besides the row sums
  L.gsum = o; o += (size_t)rs * NBK + ((rs * NBK) & 1); // cross terms of the blocks (cluster totals)
  L.part = o; o += (size_t)rs * 2 * cs * V5_B;          // [RS][2][CS][B] records of h, double buffered by exchange parity
  L.covstride = 2 * covL;
  L.cov = o; o += (size_t)rs * L.covstride;             // BNMTF VB: per row (fsv_l, A_il), l < covL
  L.prof = o;
#ifdef BNMTF_V2_PROF
  o += 32;                                              // phase clocks (timing experiments)
#endif
  L.total = o;
  return L;
}

__device__ __forceinline__ void lds_pair5(uint32_t addr, double& a, double& b) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
__device__ __forceinline__ double lds5(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts5(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_pair5(uint32_t addr, double a, double b) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ double ld_dsmem5(uint32_t cluster_addr) {
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(cluster_addr) : "memory");
  return v;
}
__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_bar) {   // release at cluster scope
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster_a(uint32_t bar, uint32_t parity) {
  // arrivals come from the whole cluster (release at cluster scope).  Polling with .acquire.cluster costs an L1
  // invalidation (CCTL.IVALL) per attempt; the loop polls relaxed and one cluster-scope acquire fence follows.
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.relaxed.cluster.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "fence.acquire.cluster;\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}

// Low half of a packed offset pair.  volatile: the compiler otherwise hoists the ten masked offsets of a row out of the
// pass loops and, out of registers, spills them -- a local-memory load per gather (seen in the SASS of the first build).
__device__ __forceinline__ uint32_t lo16_5(uint32_t x) {
  uint32_t r;
  asm volatile("and.b32 %0, %1, 0xFFFF;" : "=r"(r) : "r"(x));
  return r;
}

// a scalar of the model (tau) read where it is used instead of being kept in registers across the row
__device__ __forceinline__ double ld_scal5(const double* ptr) {
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(ptr));
  return v;
}

// ---- warp sums on the FP64 tensor core -----------------------------------------------------------------------------
// A butterfly of double shuffles is 10 SHFL per value, every one of them a trip through the shared-memory / shuffle
// pipe the gathers saturate -- and five of them in a row sit on each row's dependency chain.  mma.sync m8n8k4 (DMMA) sums
// across lanes on the tensor pipe instead: with A = 1 and B[k][n] = v(lane 4 n + k) the product holds the sums G_n of
// the eight lane quads, two per lane (D[i][2 q], D[i][2 q + 1], q = lane mod 4); their pair sums as A' against B' = 1 give
// the total in every lane.  Two DMMAs and one DADD per value, a fixed summation order, no shuffle.
__device__ __forceinline__ void dmma_884(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%4, %5};"
      : "=d"(d0), "=d"(d1)
      : "d"(a), "d"(b), "d"(0.0), "d"(0.0));
}
__device__ __forceinline__ double warp_sum_mma(double v) {
  double g0, g1, t0, t1;
  dmma_884(g0, g1, 1.0, v);
  dmma_884(t0, t1, g0 + g1, 1.0);
  return t0;
}
// sum of eight values held by the lanes 0..7 (and again by 8..15, 16..23, 24..31): one DMMA, quads 0-3 and 4-7, one DADD
__device__ __forceinline__ double oct_sum_mma(double v) {
  double g0, g1;
  dmma_884(g0, g1, 1.0, v);
  return g0 + g1;
}

// the random material of one (row, column): all six values go to the scratch table (L2), the first normal -- the one
// nearly every draw is made of -- is returned for the row's table
static __device__ __noinline__ double tn_pre_material_row5(uint64_t seed, uint32_t purpose, uint64_t index, uint32_t k, uint32_t iteration,
                                                           double* __restrict__ gdst) {
  Rng rng(seed, purpose, index, k, iteration);
  const TnPre m = tn_pre_material(rng);
  gdst[0] = m.z1; gdst[1] = m.z2; gdst[2] = m.e1; gdst[3] = m.l1; gdst[4] = m.e2; gdst[5] = m.l2;
  return m.z1;
}

// Truncated-normal draw from the standardised lower bound a = -mu / sigma and the prepared random material: the
// branches and proposals of tn_draw_lazy (rng.cuh).  Returns sigma * (z - a) >= 0; ok = false when both proposals
// were rejected (the caller then draws by the inverse CDF, tn_draw_fallback).
__device__ __forceinline__ double tn_draw_bound5(double a, double sigma, double z1, const double* gmat, bool& ok) {
  double x;
  if (!(a > TN_FAST_A0)) {
    ok = (z1 >= a);
    x = z1 - a;
    if (!ok) { const double z2 = __ldcg(gmat + 1); ok = (z2 >= a); x = z2 - a; }
  } else {
    // (written by another lane of this warp earlier in the row: read through L2, never through a stale L1 line)
    const double e1 = __ldcg(gmat + 2), l1 = __ldcg(gmat + 3), e2 = __ldcg(gmat + 4), l2 = __ldcg(gmat + 5);
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

// ---- tensor memory as the home of the residual segments -----------------------------------------------------------
// 20 entries per lane in registers are 40 registers; with the offsets (10) and the state of a row warp that does its own
// conditionals the 80 registers a thread may have at 24 warps per SM do not suffice, and what spills goes through the
// same L1 data pipe the gathers are bound by -- and, with nearly all of L1 carved out as shared memory, on to L2
// (first build: 1.1e9 local-load sectors, 45 GB of L2 traffic per sweep).  Tensor memory (256 KB per SM, reached
// only by tcgen05.ld / tcgen05.st on a data path of its own) is idle in a kernel without MMAs: each row warp keeps its
// segment in the 32 TMEM lanes of its quarter (warp % 4), NPT doubles = 2 NPT columns, and walks it in chunks of 4
// entries (tcgen05.ld.32x32b.x8) with the next chunk's load in flight while the current one is used.
struct TmChunk5 {         // 8 columns = 4 doubles per lane
  uint32_t r[8];
  __device__ __forceinline__ double get(int i) const { return __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]); }
  __device__ __forceinline__ void set(int i, double v) { r[2 * i] = (uint32_t)__double2loint(v); r[2 * i + 1] = (uint32_t)__double2hiint(v); }
};
__device__ __forceinline__ void tm5_ld8(uint32_t taddr, TmChunk5& ch) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(ch.r[0]), "=r"(ch.r[1]), "=r"(ch.r[2]), "=r"(ch.r[3]), "=r"(ch.r[4]), "=r"(ch.r[5]), "=r"(ch.r[6]), "=r"(ch.r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tm5_st8(uint32_t taddr, const TmChunk5& ch) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(ch.r[0]), "r"(ch.r[1]),
               "r"(ch.r[2]), "r"(ch.r[3]), "r"(ch.r[4]), "r"(ch.r[5]), "r"(ch.r[6]), "r"(ch.r[7])
               : "memory");
}
// the chunk is an in/out operand so that no use of its registers is scheduled above the wait
__device__ __forceinline__ void tm5_wait_ld(TmChunk5& ch) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(ch.r[0]), "+r"(ch.r[1]), "+r"(ch.r[2]), "+r"(ch.r[3]), "+r"(ch.r[4]), "+r"(ch.r[5]), "+r"(ch.r[6]), "+r"(ch.r[7])
               :
               : "memory");
}
__device__ __forceinline__ void tm5_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// TMEM columns a CTA of `warps` warps allocates: the warps of a lane quarter sit side by side (power of two >= 32)
__host__ __device__ constexpr int sweep5_tm_cols(int npt, int warps) {
  const int need = ((warps + 3) / 4) * 2 * npt;
  return need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
}

// p.Ykm: block-major copy of the other factor's values, [NBK][ldy][2] (rm_to_kmB_kernel); p.Ykm2 (VB only): the same
// layout for Var + E^2; p.gmat (Gibbs only): scratch of gridDim.x * RS * KP * 6 doubles for the random material.
// RS row warps + the producer warp per CTA, MINB CTAs (of different clusters) per SM, NSLOT tile slots.
// registers per thread: the whole register file over the resident threads, in allocation units of 8
__host__ __device__ constexpr int sweep5_regs(int rs, int minb) {
  return (65536 / ((rs + 1) * 32 * minb)) / 8 * 8 > 255 ? 255 : (65536 / ((rs + 1) * 32 * minb)) / 8 * 8;
}
// phase clocks of one row warp (cluster 0, CTA 0, warp 1): build with -DBNMTF_V2_PROF, run with BNMTF_V2_DBG=2
#ifdef BNMTF_V2_PROF
#define P5_T0() long long t_prof = prof_on ? clock64() : 0
#define P5_ADD(slot)                                                                           \
  do {                                                                                         \
    if (prof_on) {                                                                             \
      const long long t_now = clock64();                                                       \
      if (lane == 0) prof_sm[slot] += (unsigned long long)(t_now - t_prof);                    \
      t_prof = clock64();                                                                      \
    }                                                                                          \
  } while (0)
#else
#define P5_T0() do { } while (0)
#define P5_ADD(slot) do { } while (0)
#endif

// PROJ (tri-factorisation F sweep, Gibbs / ICM): an epilogue projects the residual the row ends with onto the K2 columns of
// a second block-major factor, proj_out[global row][l] = sum_j e_j Y2[l][j] (the row part of the S_kl numerators,
// bnmtf_gibbs.py:277-283): ceil(K2 / 2) more tiles per row set, the partial sums of the 8 segments meet in CTA 0.
// COV (tri-factorisation VB sweeps of F and G): the covariance term of update_F / update_G (bnmtf_vb.py:343,371),
//   cov_ik = sum_l S(k,l) A_il ((X S)_il - x_ik S(k,l)),   A = masked sums of the other factor's variances,
// is subtracted from the row's dot product.  The row's warp keeps (X S)_il and A_il in shared memory, one l per lane, and
// forms cov of both columns of a block -- plus the cross term that carries the first column's move into the second's
// cov -- by three tensor-core warp sums while the records of the block travel; (X S)_i is rank-1 updated after each block.
template <int NPT, int MODE, int CS, int RS, int NSLOT, int MINB, bool PROJ = false, bool COV = false>
__global__ void __maxnreg__(sweep5_regs(RS, MINB)) sweep5_kernel(const Sweep2Params p) {
  constexpr int B = V5_B;
  constexpr bool VB = (MODE == M_VB), GIBBS = (MODE == M_GIBBS);
  static_assert(!COV || VB, "the covariance term belongs to the VB updates");
  constexpr int TABW = 4;                                 // Gibbs: c1, sigma, z1, x_old; ICM: c1, c2, -, x_old; VB: c1, c2, precision, x_old
  constexpr int NCH = NPT / 4;                            // chunks of 4 entries
  constexpr int TMCOLS = sweep5_tm_cols(NPT, RS + 1);
  static_assert(MODE != M_NP, "the NP updates run on the row-team kernel");
  static_assert(CS == 8, "the record exchange assumes 8 CTAs per cluster");
  static_assert(NPT % 4 == 0, "entries per lane: multiple of 4");
  static_assert(NSLOT >= 2 && NSLOT <= 4, "tile slots");
  static_assert(TMCOLS * MINB <= 512, "tensor memory columns per SM");
  const int c = (int)cg::this_cluster().block_rank();
  const int cid = blockIdx.x / CS, nclusters = gridDim.x / CS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = p.K, tw = p.tw;
  const int NBK = (K + B - 1) / B, KP = NBK * B;
  const int TA = VB ? 2 * NBK : NBK;                      // tiles of the residual pass (VB: values and second moments alternate)
  const int K2 = PROJ ? p.K2 : 0, NBL = (K2 + B - 1) / B; // projection epilogue: blocks of the second factor
  const int RECW = sweep5_rec_width(K2);                  // doubles per CTA of a row's final record
  const int total = TA + NBK + NBL;                       // tiles per row set
  const uint32_t slot_bytes = (uint32_t)(tw + V2_PAD) * B * 8u;
  const uint32_t tile_bytes = (uint32_t)tw * B * 8u;
  const bool want_mu = p.Xmu != nullptr;

#ifdef BNMTF_V2_PROF
  const long long t_entry = clock64();
#endif
  extern __shared__ __align__(128) double smem5[];
  const int covL = COV ? p.covL : 0;
  const Sweep5Layout L = sweep5_layout(tw, RS, K, CS, NSLOT, TABW, VB, K2, covL);
  const uint32_t sm = smem_u32(smem5);
  const uint32_t tiles_a = sm + (uint32_t)L.tiles * 8u, mbar_a = sm + (uint32_t)L.mbar * 8u;
  const uint32_t full_a = mbar_a, empty_a = mbar_a + NSLOT * 8u, rec_a = mbar_a + 2 * NSLOT * 8u;
  const uint32_t statbar_a = rec_a + 2 * RS * 8u, rowbar_a = statbar_a + RS * 8u;
  const uint32_t stat_own_a = sm + (uint32_t)L.stat_own * 8u, tab_a = sm + (uint32_t)L.tab * 8u, gsum_a = sm + (uint32_t)L.gsum * 8u;
  const uint32_t part_a = sm + (uint32_t)L.part * 8u, misc_a = sm + (uint32_t)L.misc * 8u;
  const uint32_t setq_a = misc_a + 16u, setbar_a = misc_a + 32u;
  // Row sets are handed out dynamically: the first round statically (set = cluster index), then from a global counter.
  // CTAs that have an SM to themselves (the GPCs do not divide evenly into clusters) run a set 20 % faster than CTAs that
  // share one, and a cluster is as fast as its slowest CTA: with a static round-robin the kernel waited 1.5 of 7.5 ms for
  // the slowest clusters.  The producer lane of CTA 0 fetches the cluster's next set one set ahead and posts it to every
  // CTA's queue (4 entries; plain remote store + remote mbarrier arrive, release / acquire at cluster scope).
  auto next_set = [&](uint32_t it) -> int {               // the set of iteration `it` >= 1: queue entry it - 1
    const uint32_t q = it - 1u;
    mbar_wait_cluster_a(setbar_a + (q & 3u) * 8u, (q >> 2) & 1u);
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(setq_a + (q & 3u) * 4u) : "memory");
    return v;
  };
#ifdef BNMTF_V2_PROF
  unsigned long long* prof_sm = reinterpret_cast<unsigned long long*>(smem5 + L.prof);
  if (threadIdx.x < 32) prof_sm[threadIdx.x] = 0ull;
#endif
  auto slot_of = [](uint32_t g) -> uint32_t { return NSLOT == 2 ? (g & 1u) : (NSLOT == 4 ? (g & 3u) : g % 3u); };
  auto phase_of = [](uint32_t g) -> uint32_t { return (NSLOT == 2 ? (g >> 1) : (NSLOT == 4 ? (g >> 2) : g / 3u)) & 1u; };

  if (threadIdx.x == 0) {
    uint64_t* mb = reinterpret_cast<uint64_t*>(smem5 + L.mbar);
    for (int s = 0; s < NSLOT; ++s) { mbar_init(&mb[s], 1); mbar_init(&mb[NSLOT + s], RS); }
    for (int q = 0; q < 2 * RS; ++q) mbar_init(&mb[2 * NSLOT + q], 1);            // records
    for (int q = 0; q < RS; ++q) { mbar_init(&mb[2 * NSLOT + 2 * RS + q], CS); mbar_init(&mb[2 * NSLOT + 3 * RS + q], 1); }
    for (int q = 0; q < 4; ++q) mbar_init(reinterpret_cast<uint64_t*>(smem5 + L.misc) + 4 + q, 1);                    // set queue
    fence_mbar_init();
  }
  for (int q = threadIdx.x; q < NSLOT * V2_PAD * B; q += blockDim.x) {   // the zero entries read by padded slots
    const int s = q / (V2_PAD * B), z = q - s * (V2_PAD * B);
    sts5(tiles_a + (uint32_t)s * slot_bytes + tile_bytes + (uint32_t)z * 8u, 0.0);
  }
  // tensor memory for the residual segments of this CTA
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(misc_a + 8u), "n"(TMCOLS) : "memory");
    if (lane == 0) sts5(misc_a, p.scal[0]);               // tau, read by every conditional
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tm_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tm_base) : "r"(misc_a + 8u) : "memory");
  cluster_sync_all();        // every CTA's mbarriers exist (and tm_base is read) before anything arrives from a peer

  if (warp == 0) {
    // =========================== producer warp: the tile ring ===========================
    if (lane == 0) {
      const double* ybase = p.Ykm + (size_t)c * tw * B;   // this CTA's tile of block 0
      const double* qbase = VB ? p.Ykm2 + (size_t)c * tw * B : nullptr;
      const double* pbase = PROJ ? p.Yproj + (size_t)c * tw * B : nullptr;   // this CTA's tile of the second factor's block 0
      const size_t ystride = (size_t)p.ldy * B, pstride = (size_t)p.ldp * B;
      uint32_t g = 0, it = 0;
      for (int set = cid; set < p.nsets; set = next_set(++it)) {
        for (int t = 0; t < total; ++t, ++g) {
          const double* src = (PROJ && t >= TA + NBK) ? pbase + (size_t)(t - TA - NBK) * pstride
                              : t >= TA ? ybase + (size_t)(t - TA) * ystride
                                        : (!VB ? ybase + (size_t)t * ystride : ((t & 1) ? qbase : ybase) + (size_t)(t >> 1) * ystride);
          const uint32_t slot = slot_of(g);
          mbar_wait_a(empty_a + slot * 8u, phase_of(g) ^ 1u);   // all row warps have released the slot's previous tile
          fence_proxy_async();
          mbar_expect_tx_a(full_a + slot * 8u, tile_bytes);
          tma_load_1d(smem5 + L.tiles + (size_t)slot * (slot_bytes / 8u), src, tile_bytes,
                      reinterpret_cast<uint64_t*>(smem5 + L.mbar) + slot);
          if (t == 0 && c == 0) {                               // the cluster's set after this one, to all 8 CTAs
            const int nxt = nclusters + atomicAdd(p.set_counter, 1);
            for (uint32_t cc = 0; cc < (uint32_t)CS; ++cc) {
              asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(map_to_cta(setq_a + (it & 3u) * 4u, cc)), "r"(nxt) : "memory");
              mbar_arrive_remote(map_to_cta(setbar_a + (it & 3u) * 8u, cc));
            }
          }
        }
      }
    }
    __syncwarp();
  } else {
    // =========================== row warps ===========================
    const int r = warp - 1;
    // one counter per warp: sets done.  A warp skips a row only in the matrix's last -- partial -- set, which is also the
    // last set of its cluster, so rows done == sets done, the tile sequence stands at iset * total + t and the exchange
    // sequence at iset * NBK + n.
    uint32_t iset = 0;
    const uint32_t tab_r = tab_a + (uint32_t)(r * KP * TABW) * 8u;
    const uint32_t stat_r = stat_own_a + (uint32_t)(r * L.sstride) * 8u;
    const uint32_t gsum_r = gsum_a + (uint32_t)(r * NBK) * 8u;
    const uint32_t cov_r = sm + (uint32_t)(L.cov + (size_t)r * L.covstride) * 8u;
    const uint32_t tm = tm_base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * (uint32_t)(NPT * 2);
    double* gmat_row = GIBBS ? p.gmat + ((size_t)blockIdx.x * RS + r) * KP * TN_PRE_DOUBLES : nullptr;
    // The gathers of a chunk are issued two entries at a time: the address of the next pair is made to depend on the last
    // value of the pair before (through a mask that is zero at run time but unknown to the compiler).  Left alone the
    // compiler either serialises every gather behind its own FMAs or issues all of them at once and spills.
    const uint32_t zmask = (uint32_t)p.dbg & 0x40000000u;
#ifdef BNMTF_V2_PROF
    const bool prof_on = (p.dbg & 2) && cid == 0 && c == 0 && warp == 1 && p.prof != nullptr;
    const long long t_loop = clock64();
    if (prof_on && lane == 0) prof_sm[21] = (unsigned long long)(t_loop - t_entry);
#endif
#define V5_CHAIN(T_, y_) T_ += ((uint32_t)__double2hiint(y_) & zmask)
#define V5_OFF(nn_) (((nn_) & 1) ? (cp[(nn_) / 2] >> 16) : lo16_5(cp[(nn_) / 2]))

    for (int set = cid; set < p.nsets; set = next_set(iset)) {
      const bool live = r < min(RS, p.n_loc - set * RS);
      uint32_t g = iset * (uint32_t)total;        // position in the tile sequence
      if (!live) {             // a set with fewer rows: keep the tile ring's arrival counts whole
        for (int t = 0; t < total; ++t, ++g) {
          mbar_wait_a(full_a + slot_of(g) * 8u, phase_of(g));
          if (lane == 0) mbar_arrive_local(empty_a + slot_of(g) * 8u);
        }
        ++iset;
        continue;
      }
      const int64_t grow = p.row0 + (int64_t)set * RS + r;
      P5_T0();
#ifdef BNMTF_V2_PROF
      if (prof_on && lane == 0) prof_sm[16] += 1;
#endif
      const double* __restrict__ xrow = p.Xval + (size_t)grow * K;
      {                        // the row's current values -- and, Gibbs, the random material -- one column per lane
        const uint32_t iteration = p.iter_ptr ? *p.iter_ptr : p.iteration;
        for (int k = lane; k < K; k += 32) {
          const double z1 = GIBBS ? tn_pre_material_row5(p.seed, p.purpose, (uint64_t)grow, (uint32_t)k, iteration, gmat_row + (size_t)k * TN_PRE_DOUBLES) : 0.0;
          sts_pair5(tab_r + (uint32_t)(k * TABW + 2) * 8u, z1, xrow[k]);
        }
        if (lane == 0 && (K & 1)) sts_pair5(tab_r + (uint32_t)(K * TABW + 2) * 8u, 0.0, 0.0);   // the odd block's second column
        __syncwarp();
      }
      uint32_t cp[NPT / 2];
      {
        const double* __restrict__ tv = p.tvals + (((int64_t)(set * CS + c) * NPT) * RS * 32 + (int64_t)r * 32 + lane);
        const uint16_t* __restrict__ tc = p.tcols + (((int64_t)(set * CS + c) * NPT) * RS * 32 + (int64_t)r * 32 + lane);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {       // the segment's values go to tensor memory
          TmChunk5 t;
#pragma unroll
          for (int i = 0; i < 4; ++i) t.set(i, tv[(size_t)(4 * ch + i) * RS * 32]);
          tm5_st8(tm + 8u * ch, t);
        }
#pragma unroll
        for (int n = 0; n < NPT; n += 2) {
          const uint32_t a = tc[(size_t)n * RS * 32], b = tc[(size_t)(n + 1) * RS * 32];
          cp[n / 2] = (a * (B * 8)) | ((b * (B * 8)) << 16);   // byte offsets inside a tile slot
        }
      }
      P5_ADD(0);            // row start: random material, segment into tensor memory

      // ---- residual pass: e = r - sum_k x_k y_k, and the residual-independent row sums
      for (int m = 0; m < NBK; ++m) {
        const int k0 = m * B;
        const uint32_t slot = slot_of(g);
        const double x0 = lds5(tab_r + (uint32_t)(k0 * TABW + 3) * 8u), x1 = lds5(tab_r + (uint32_t)((k0 + 1) * TABW + 3) * 8u);
        mbar_wait_a(full_a + slot * 8u, phase_of(g));
        P5_ADD(1);          // residual pass: waiting for the tile
        uint32_t T = tiles_a + slot * slot_bytes;
        double a0 = 0.0, a1 = 0.0, gg = 0.0;
        TmChunk5 b[2];
        tm5_wait_st();
        tm5_ld8(tm, b[0]);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          TmChunk5& cur = b[ch & 1];
          tm5_wait_ld(cur);
          if (ch + 1 < NCH) tm5_ld8(tm + 8u * (ch + 1), b[(ch + 1) & 1]);
#pragma unroll
          for (int h2 = 0; h2 < 4; h2 += 2) {
            double y0[2], y1[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) lds_pair5(T + V5_OFF(4 * ch + h2 + i), y0[i], y1[i]);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              cur.set(h2 + i, fma(-x1, y1[i], fma(-x0, y0[i], cur.get(h2 + i))));
              a0 = fma(y0[i], y0[i], a0);
              a1 = fma(y1[i], y1[i], a1);
              gg = fma(y0[i], y1[i], gg);
            }
            V5_CHAIN(T, y1[1]);
          }
          tm5_st8(tm + 8u * ch, cur);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_local(empty_a + slot * 8u);
        ++g;
        {
          const double ta0 = warp_sum_mma(a0), ta1 = warp_sum_mma(a1), tg = warp_sum_mma(gg);
          if ((lane & 7) == 0 && lane < 24)
            sts5(stat_r + (uint32_t)(lane == 0 ? k0 : (lane == 8 ? k0 + 1 : KP + m)) * 8u, lane == 0 ? ta0 : (lane == 8 ? ta1 : tg));
        }
        if (VB) {              // second moments of the same block: plain masked sums
          const uint32_t slot2 = slot_of(g);
          mbar_wait_a(full_a + slot2 * 8u, phase_of(g));
          uint32_t T2 = tiles_a + slot2 * slot_bytes;
          double q0 = 0.0, q1 = 0.0;
#pragma unroll
          for (int n0 = 0; n0 < NPT; n0 += 2) {
            double y0[2], y1[2];
            lds_pair5(T2 + V5_OFF(n0), y0[0], y1[0]);
            lds_pair5(T2 + V5_OFF(n0 + 1), y0[1], y1[1]);
            q0 += y0[0]; q1 += y1[0]; q0 += y0[1]; q1 += y1[1];
            V5_CHAIN(T2, y1[1]);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive_local(empty_a + slot2 * 8u);
          ++g;
          const double tq0 = warp_sum_mma(q0), tq1 = warp_sum_mma(q1);
          if ((lane & 15) == 0) sts5(stat_r + (uint32_t)(KP + NBK + k0 + (lane >> 4)) * 8u, lane ? tq1 : tq0);
        }
      }

      P5_ADD(2);            // residual pass
      // ---- the 8 warps of the row meet: statistics of all tiles, then the row's table
      __syncwarp();
      if (lane < CS) mbar_arrive_remote(map_to_cta(statbar_a + (uint32_t)r * 8u, (uint32_t)lane));
      mbar_wait_cluster_a(statbar_a + (uint32_t)r * 8u, iset & 1u);
      P5_ADD(3);            // waiting for the row's other warps
      double va[2] = {0.0, 0.0};       // VB: a_k of this lane's columns, for the variance part of exp_square_diff at the row's end
      {
        const double tau = lds5(misc_a);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int k = lane + 32 * q;
          if (k < K) {
            double a = 0.0, qq = 0.0;
#pragma unroll
            for (int cc = 0; cc < CS; ++cc) a += ld_dsmem5(map_to_cta(stat_r + (uint32_t)k * 8u, (uint32_t)cc));
            if (VB) {
#pragma unroll
              for (int cc = 0; cc < CS; ++cc) qq += ld_dsmem5(map_to_cta(stat_r + (uint32_t)(KP + NBK + k) * 8u, (uint32_t)cc));
              va[q] = a;
            } else {
              qq = a;
            }
            const double lam = p.lamk ? p.lamk[k] : (p.lam_rm ? p.lam_rm[grow * K + k] : 0.0);
            const double x_old = lds5(tab_r + (uint32_t)(k * TABW + 3) * 8u);
            const double ct = tau * qq;                                  // tauU, bnmf_gibbs.py:205 / bnmf_vb.py:254
            // with num = -lam + tau (h + x_old a), prec = ct, sigma = prec^-1/2:
            //   Gibbs needs the standardised bound -num sigma = c1 - c2 h,  c1 = sigma (lam - tau x_old a), c2 = sigma tau;
            //   ICM / VB need the mean num / prec = c1 + c2 h,  c1 = (-lam + tau x_old a) / prec, c2 = tau / prec.
            if (GIBBS) {
              const double sg = rsqrt(ct);
              sts_pair5(tab_r + (uint32_t)(k * TABW) * 8u, sg * (lam - tau * (x_old * a)), sg);
            } else {
              const double ic = 1.0 / ct;
              sts_pair5(tab_r + (uint32_t)(k * TABW) * 8u, (tau * (x_old * a) - lam) * ic, tau * ic);
              if (VB) sts5(tab_r + (uint32_t)(k * TABW + 2) * 8u, ct);
            }
            if (want_mu && c == 0) p.Xtau[(size_t)grow * K + k] = ct;
          }
        }
      }
      if (lane < NBK) {
        double gt = 0.0;
#pragma unroll
        for (int cc = 0; cc < CS; ++cc) gt += ld_dsmem5(map_to_cta(stat_r + (uint32_t)(KP + lane) * 8u, (uint32_t)cc));
        sts5(gsum_r + (uint32_t)lane * 8u, gt);
      }
      __syncwarp();
      if (COV) {               // (X S)_il from the row's current values, and A_il
        for (int l = lane; l < covL; l += 32) {
          double a = 0.0;
          for (int kk = 0; kk < K; ++kk) a = fma(lds5(tab_r + (uint32_t)(kk * TABW + 3) * 8u), __ldg(p.covS + kk * p.cov_sk + l * p.cov_sl), a);
          sts_pair5(cov_r + (uint32_t)l * 16u, a, __ldg(p.covA + (size_t)grow * covL + l));
        }
        __syncwarp();
      }

      // ---- the K column updates, B columns per exchange.  The pass of block n applies the rank-1 corrections of block
      //      n - 1 (tile n - 1) and takes the dot products with block n (tile n) in one walk over the segment.
      double d0 = 0.0, d1 = 0.0;
      P5_ADD(4);            // table
      for (int n = 0; n < NBK; ++n) {
        const int k0 = n * B, nb = min(B, K - k0);
        const uint32_t slot = slot_of(g);
        mbar_wait_a(full_a + slot * 8u, phase_of(g));
        P5_ADD(5);          // update passes: waiting for the tile
        double hh;
        {
          uint32_t T = tiles_a + slot * slot_bytes;
          uint32_t Tp = tiles_a + slot_of(g - 1u) * slot_bytes;    // (n == 0: not read)
          double h0 = 0.0, h1 = 0.0;
          TmChunk5 b[2];
          tm5_wait_st();
          tm5_ld8(tm, b[0]);
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch) {
            TmChunk5& cur = b[ch & 1];
            tm5_wait_ld(cur);
            if (ch + 1 < NCH) tm5_ld8(tm + 8u * (ch + 1), b[(ch + 1) & 1]);
            if (n > 0) {
#pragma unroll
              for (int h2 = 0; h2 < 4; h2 += 2) {
                double p0[2], p1[2], y0[2], y1[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                  lds_pair5(Tp + V5_OFF(4 * ch + h2 + i), p0[i], p1[i]);
                  lds_pair5(T + V5_OFF(4 * ch + h2 + i), y0[i], y1[i]);
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                  const double ee = fma(-d1, p1[i], fma(-d0, p0[i], cur.get(h2 + i)));
                  cur.set(h2 + i, ee);
                  h0 = fma(ee, y0[i], h0); h1 = fma(ee, y1[i], h1);
                }
                V5_CHAIN(T, y1[1]); V5_CHAIN(Tp, p1[1]);
              }
              tm5_st8(tm + 8u * ch, cur);
            } else {
#pragma unroll
              for (int h2 = 0; h2 < 4; h2 += 2) {
                double y0[2], y1[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) lds_pair5(T + V5_OFF(4 * ch + h2 + i), y0[i], y1[i]);
#pragma unroll
                for (int i = 0; i < 2; ++i) { h0 = fma(cur.get(h2 + i), y0[i], h0); h1 = fma(cur.get(h2 + i), y1[i], h1); }
                V5_CHAIN(T, y1[1]);
              }
            }
          }
          P5_ADD(6);        // pass
          h0 = warp_sum_mma(h0); h1 = warp_sum_mma(h1);
          hh = (lane & 16) ? h1 : h0;
        }
        if (n > 0) {           // the tile of block n - 1 is no longer read by this warp
          __syncwarp();
          if (lane == 0) mbar_arrive_local(empty_a + slot_of(g - 1u) * 8u);
        }
        const uint32_t xseq = iset * (uint32_t)NBK + (uint32_t)n;   // exchanges of this warp so far (record buffer = parity)
        const uint32_t bpar = xseq & 1u;
        const uint32_t recbar = rec_a + (uint32_t)(r * 2 + bpar) * 8u;
        const uint32_t recbuf = part_a + (uint32_t)((r * 2 + bpar) * CS * B) * 8u;
        if (lane == 0) mbar_expect_tx_a(recbar, (uint32_t)(CS * B * 8));
        if ((lane & 15) < CS)    // lanes cc and 16 + cc send this warp's two sums to CTA cc
          st_async_f64(map_to_cta(recbuf + (uint32_t)(c * B + (lane >> 4)) * 8u, (uint32_t)(lane & 15)), hh,
                       map_to_cta(recbar, (uint32_t)(lane & 15)));
        // what the conditionals need besides the sums is fetched while the records travel
        const double g01 = lds5(gsum_r + (uint32_t)n * 8u);
        const double tau = lds5(misc_a);
        double tt[B][TABW];
#pragma unroll
        for (int j = 0; j < B; ++j) {
#pragma unroll
          for (int w = 0; w < TABW; w += 2) lds_pair5(tab_r + (uint32_t)((k0 + j) * TABW + w) * 8u, tt[j][w], tt[j][w + 1]);
        }
        double cov0 = 0.0, cov1 = 0.0, ccx = 0.0;
        if (COV) {             // cov of the block's two columns from (X S)_i as it stands, and the cross term
          const double xo0 = tt[0][3], xo1 = tt[1][3];
          for (int l = lane; l < covL; l += 32) {
            double f, a;
            lds_pair5(cov_r + (uint32_t)l * 16u, f, a);
            const double s0 = __ldg(p.covS + k0 * p.cov_sk + l * p.cov_sl);
            const double s1 = nb > 1 ? __ldg(p.covS + (k0 + 1) * p.cov_sk + l * p.cov_sl) : 0.0;
            cov0 = fma(s0 * a, f - xo0 * s0, cov0);
            cov1 = fma(s1 * a, f - xo1 * s1, cov1);
            ccx = fma(s1 * a, s0, ccx);
          }
          cov0 = warp_sum_mma(cov0); cov1 = warp_sum_mma(cov1); ccx = warp_sum_mma(ccx);
        }
        P5_ADD(7);          // reduction, records sent
        mbar_wait_a(recbar, (xseq >> 1) & 1u);
        P5_ADD(8);          // waiting for the records
        // the records of the 8 tiles: lane l takes record l mod 8, one DMMA adds records 0-3 and 4-7 -- the same bits in
        // every lane and in every CTA
        double hs0, hs1;
        lds_pair5(recbuf + (uint32_t)((lane & 7) * B) * 8u, hs0, hs1);
        hs0 = oct_sum_mma(hs0); hs1 = oct_sum_mma(hs1);
        d0 = 0.0; d1 = 0.0;
#pragma unroll
        for (int j = 0; j < B; ++j) {
          if (j < nb) {
            const int k = k0 + j;
            const double* tj = tt[j];
            const double x_old = tj[3];
            double h = j ? fma(-d0, g01, hs1) : hs0;           // column k0 + 1 sees the residual after column k0 moved
            if (COV) h -= j ? fma(d0, ccx, cov1) : cov0;       // ... and (X S)_i after column k0 moved
            double x_new, vn = 0.0, mu;
            if (GIBBS) {
              const double sg = tj[1];
              const double ab = fma(-(sg * tau), h, tj[0]);     // standardised bound -mu / sigma
              bool ok;
              x_new = tn_draw_bound5(ab, sg, tj[2], gmat_row + (size_t)k * TN_PRE_DOUBLES, ok);
              mu = -ab * sg;
              if (!(sg < CUDART_INF)) { x_new = 0.0; ok = true; }   // precision 0: TN_vector_draw returns 0 (truncated_normal_vector.py:41)
              if (!ok)                                          // both proposals rejected (rare): inverse CDF, off the fast path
                x_new = tn_draw_fallback(-ab / sg, 1.0 / (sg * sg), p.seed, p.purpose, (uint64_t)grow, (uint32_t)k, p.iter_ptr ? *p.iter_ptr : p.iteration).x;
            } else if (MODE == M_ICM) {
              mu = fma(tj[1], h, tj[0]);
              x_new = fmax(fmax(0.0, mu), MINIMUM_TN);          // nmf_icm.py:159-160
            } else {
              mu = fma(tj[1], h, tj[0]);
              tn_moments(mu, tj[2], x_new, vn);                 // bnmf_vb.py:255,275-278
            }
            if (j) d1 = x_new - x_old; else d0 = x_new - x_old;
            // The column's table entry is dead from here on: CTA 0 parks the new value there (VB: E, Var, precision, mu;
            // else: value, mu) and writes the row out once all 8 CTAs are through with it -- a CTA that lags behind still
            // reads the OLD value of its current block from global memory.
            if (c == 0) {
              __syncwarp();
              if (lane == 0) {
                sts_pair5(tab_r + (uint32_t)(k * TABW) * 8u, x_new, VB ? vn : mu);
                if (VB) sts5(tab_r + (uint32_t)(k * TABW + 3) * 8u, mu);
              }
            }
          }
        }
        if (COV) {             // (X S)_il += delta_k S(k,l)
          for (int l = lane; l < covL; l += 32) {
            double f = lds5(cov_r + (uint32_t)l * 16u);
            f = fma(d0, __ldg(p.covS + k0 * p.cov_sk + l * p.cov_sl), f);
            if (nb > 1) f = fma(d1, __ldg(p.covS + (k0 + 1) * p.cov_sk + l * p.cov_sl), f);
            sts5(cov_r + (uint32_t)l * 16u, f);
          }
          __syncwarp();
        }
        ++g;
        P5_ADD(9);          // conditionals
      }

      // ---- last corrections and the statistics of the final residual; they meet in CTA 0, which also writes the row out
      {
        const double* __restrict__ tv = p.tvals + (((int64_t)(set * CS + c) * NPT) * RS * 32 + (int64_t)r * 32 + lane);
        uint32_t Tp = tiles_a + slot_of(g - 1u) * slot_bytes;
        double st0 = 0.0, st1 = 0.0, st2 = 0.0;
        TmChunk5 b[2];
        tm5_wait_st();
        tm5_ld8(tm, b[0]);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          TmChunk5& cur = b[ch & 1];
          tm5_wait_ld(cur);
          if (ch + 1 < NCH) tm5_ld8(tm + 8u * (ch + 1), b[(ch + 1) & 1]);
#pragma unroll
          for (int h2 = 0; h2 < 4; h2 += 2) {
            double p0[2], p1[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) lds_pair5(Tp + V5_OFF(4 * ch + h2 + i), p0[i], p1[i]);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const double ee = fma(-d1, p1[i], fma(-d0, p0[i], cur.get(h2 + i)));
              const double rr = tv[(size_t)(4 * ch + h2 + i) * RS * 32];
              if (PROJ) cur.set(h2 + i, ee);
              st0 = fma(ee, ee, st0);
              st1 += ee;
              st2 = fma(rr - p.rmean, ee, st2);   // padded slots hold e == 0
            }
            V5_CHAIN(Tp, p1[1]);
          }
          if (PROJ) tm5_st8(tm + 8u * ch, cur);   // the projection reads the final residual
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_local(empty_a + slot_of(g - 1u) * 8u);
        st0 = warp_sum_mma(st0); st1 = warp_sum_mma(st1); st2 = warp_sum_mma(st2);
        const uint32_t rowbar = rowbar_a + (uint32_t)r * 8u;
        if (c == 0 && lane == 0) mbar_expect_tx_a(rowbar, (uint32_t)(CS * (3 + B * NBL) * 8));
        if (lane < 3)
          st_async_f64(map_to_cta(stat_r + (uint32_t)(c * RECW + lane) * 8u, 0u), lane == 0 ? st0 : (lane == 1 ? st1 : st2),
                       map_to_cta(rowbar, 0u));
        if (PROJ) {
          // ---- projection epilogue: sum_j e_j Y2[l][j] over this segment, two columns per tile, partial sums to CTA 0
          for (int m = 0; m < NBL; ++m) {
            const uint32_t slot = slot_of(g);
            mbar_wait_a(full_a + slot * 8u, phase_of(g));
            uint32_t T = tiles_a + slot * slot_bytes;
            double h0 = 0.0, h1 = 0.0;
            tm5_wait_st();
            tm5_ld8(tm, b[0]);
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
              TmChunk5& cur = b[ch & 1];
              tm5_wait_ld(cur);
              if (ch + 1 < NCH) tm5_ld8(tm + 8u * (ch + 1), b[(ch + 1) & 1]);
#pragma unroll
              for (int h2 = 0; h2 < 4; h2 += 2) {
                double y0[2], y1[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) lds_pair5(T + V5_OFF(4 * ch + h2 + i), y0[i], y1[i]);
#pragma unroll
                for (int i = 0; i < 2; ++i) { h0 = fma(cur.get(h2 + i), y0[i], h0); h1 = fma(cur.get(h2 + i), y1[i], h1); }
                V5_CHAIN(T, y1[1]);
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_local(empty_a + slot * 8u);
            ++g;
            h0 = warp_sum_mma(h0); h1 = warp_sum_mma(h1);
            if (lane < 2)
              st_async_f64(map_to_cta(stat_r + (uint32_t)(c * RECW + 4 + B * m + lane) * 8u, 0u), lane ? h1 : h0, map_to_cta(rowbar, 0u));
          }
        }
        if (c == 0) {
          double acc_q = 0.0, acc_prior = 0.0, acc_esd1 = 0.0, acc_esd2 = 0.0;
          __syncwarp();
          if (VB) {            // q(U_i) part of the ELBO, one lane per column (bnmf_vb.py:219-229); the table holds (E, Var, precision, mu)
            for (int kk = lane; kk < K; kk += 32) {
              double en, vn, ct, cm;
              lds_pair5(tab_r + (uint32_t)(kk * TABW) * 8u, en, vn);
              lds_pair5(tab_r + (uint32_t)(kk * TABW + 2) * 8u, ct, cm);
              acc_q += -0.5 * log(ct) + log_half_erfc_ref(-cm * sqrt(ct) / 1.4142135623730951) + 0.5 * ct * (vn + (en - cm) * (en - cm));
              acc_prior = fma(p.lamk ? p.lamk[kk] : (p.lam_rm ? p.lam_rm[grow * K + kk] : 0.0), en, acc_prior);
              acc_esd1 = fma(vn + en * en, ct, acc_esd1);              // (Var + E^2) tau q
              acc_esd2 = fma(en * en, kk < 32 ? va[0] : va[1], acc_esd2);   // E^2 a
            }
            acc_q = warp_sum(acc_q);
            acc_prior = warp_sum(acc_prior);
            acc_esd1 = warp_sum(acc_esd1);
            acc_esd2 = warp_sum(acc_esd2);
          }
          mbar_wait_a(rowbar, iset & 1u);      // every CTA has finished the row: nobody reads its old values any more
          for (int kk = lane; kk < K; kk += 32) {
            double v0, v1;
            lds_pair5(tab_r + (uint32_t)(kk * TABW) * 8u, v0, v1);
            p.Xval[(size_t)grow * K + kk] = v0;
            for (int pr = 0; pr < p.npeers; ++pr) p.peer_val[pr][(size_t)grow * K + kk] = v0;   // the other GPUs' copies
            if (VB) {
              p.Xvar[(size_t)grow * K + kk] = v1;
              for (int pr = 0; pr < p.npeers; ++pr) p.peer_var[pr][(size_t)grow * K + kk] = v1;
              if (want_mu) p.Xmu[(size_t)grow * K + kk] = lds5(tab_r + (uint32_t)(kk * TABW + 3) * 8u);
            } else if (want_mu) {
              p.Xmu[(size_t)grow * K + kk] = v1;
            }
          }
          if (PROJ) {          // fixed-order sum of the 8 segments' partial sums
            for (int l = lane; l < K2; l += 32) {
              double t = 0.0;
              for (int cc = 0; cc < CS; ++cc) t += lds5(stat_r + (uint32_t)(cc * RECW + 4 + l) * 8u);
              p.proj_out[(size_t)grow * K2 + l] = t;
            }
          }
          if (lane == 0 && p.rowstats != nullptr) {
            double t3[3] = {0.0, 0.0, 0.0};
            for (int cc = 0; cc < CS; ++cc) {
              const uint32_t src = stat_r + (uint32_t)(cc * RECW) * 8u;
              t3[0] += lds5(src); t3[1] += lds5(src + 8); t3[2] += lds5(src + 16);
            }
            double* rs = p.rowstats + (size_t)(set * RS + r) * NRS;
            rs[RS_RSS] = t3[0]; rs[RS_SUM_E] = t3[1]; rs[RS_SUM_RCE] = t3[2];
            rs[RS_ESDVAR] = VB ? acc_esd1 / lds5(misc_a) - acc_esd2 : 0.0; rs[RS_ELBOQ] = acc_q; rs[RS_PRIOR] = acc_prior;
            rs[RS_IDIV] = 0.0; rs[RS_SPARE] = 0.0;
            for (int pr = 0; pr < p.npeers; ++pr) {
              double* prs = p.peer_rowstats[pr] + (size_t)grow * NRS;
#pragma unroll
              for (int f = 0; f < NRS; ++f) prs[f] = rs[f];
            }
          }
        }
      }
      P5_ADD(10);           // last pass, statistics, row out
      ++iset;
    }
#ifdef BNMTF_V2_PROF
    if (prof_on && lane == 0) prof_sm[20] = (unsigned long long)(clock64() - t_loop);
#endif
#undef V5_CHAIN
#undef V5_OFF
  }
#ifdef BNMTF_V2_PROF
  __syncthreads();
  if ((p.dbg & 2) && cid == 0 && c == 0 && p.prof != nullptr && threadIdx.x == 0) prof_sm[22] = (unsigned long long)(clock64() - t_entry);
  __syncthreads();
  if ((p.dbg & 2) && cid == 0 && c == 0 && p.prof != nullptr && threadIdx.x < 32) p.prof[threadIdx.x] += prof_sm[threadIdx.x];
#endif
  // tensor memory back; no CTA may exit while its shared memory can still be read or written remotely
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_base), "n"(TMCOLS) : "memory");
  cluster_sync_all();
}

// Device-side barrier of the ranks after a sweep whose output went to the peers directly: every rank tells every other
// rank "my sweep number `epoch` and all its writes are done" (a system-scope release store into the peer's flag array)
// and waits until it has heard the same from all of them.  One warp; launched on the sweep's stream right behind it, so
// the sweep's peer writes are complete when the flags go out, and whatever follows on the stream sees the peers' rows.
static __global__ void peer_barrier_kernel(int* const* peer_flags, volatile int* my_flags, int rank, int world, int epoch) {
  const int r = threadIdx.x;
  if (r < world && r != rank) {
    __threadfence_system();
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(peer_flags[r] + rank), "r"(epoch) : "memory");
    int seen;
    do {
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(seen) : "l"(my_flags + r) : "memory");
    } while (seen - epoch < 0);
  }
}

}  // namespace bnmtf
