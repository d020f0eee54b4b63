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

// out_km[c][i] = sum_q X_rm[i][q] * Smat[q*s_q + c*s_c]  (i < n), zero padded to ld; out_kmB (optional): the same values in
// the block-major layout of the per-row-chain cluster sweep, out_kmB[((c / 2) * ld + i) * 2 + (c & 1)]
static __global__ void combine_km_kernel(const double* __restrict__ X_rm, int64_t n, int kq, const double* __restrict__ Smat, int s_q,
                                  int s_c, double* __restrict__ out_km, int64_t ld, double* __restrict__ out_kmB) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (i >= ld) return;
  double acc = 0.0;
  if (i < n)
    for (int q = 0; q < kq; ++q) acc = fma(X_rm[i * kq + q], Smat[q * s_q + c * s_c], acc);
  out_km[(size_t)c * ld + i] = acc;
  if (out_kmB) out_kmB[((size_t)(c >> 1) * ld + i) * 2 + (c & 1)] = acc;
}

// Masked row sums of the other factor's variational moments (bnmtf_vb.py:343,352-357,371):
//   outA[row][w] = sum_{j in Omega_row} Var[j][w],   outB[row][w] = sum_{j in Omega_row} (Var[j][w] + E[j][w]^2)   (WITH_B)
// One warp per row of the row layout, lane = column w (+ 32 for W > 32).  The rows of the ROW-MAJOR factor are gathered whole
// (W contiguous doubles per observed entry: every 32-byte sector fetched is used), where the projection mode of the row-team
// sweep gathers 8 bytes per sector from the k-major copy, once per column.  The 32 column indices of a step are one
// coalesced load, handed round by shuffles; eight entries are in flight per lane; fixed summation order per row.
template <int NC, bool WITH_B>
__global__ void __launch_bounds__(256) nmtf_masked_sums_kernel(const uint32_t* __restrict__ cols, const int64_t* __restrict__ rowstart,
                                                               int n_loc, int64_t row0, uint32_t padcol, const double* __restrict__ E_rm,
                                                               const double* __restrict__ V_rm, int W, double* __restrict__ outA,
                                                               double* __restrict__ outB) {
  const int row = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (row >= n_loc) return;
  constexpr int U = 8;
  double a[NC][U], b[NC][U];
#pragma unroll
  for (int q = 0; q < NC; ++q)
#pragma unroll
    for (int u = 0; u < U; ++u) { a[q][u] = 0.0; b[q][u] = 0.0; }
  const int64_t beg = rowstart[row], end = rowstart[row + 1];
  for (int64_t q0 = beg; q0 < end; q0 += 32) {
    const uint32_t mycol = (q0 + lane < end) ? cols[q0 + lane] : padcol;
#pragma unroll
    for (int r0 = 0; r0 < 32; r0 += U) {
      double v[NC][U], x[NC][U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t j = __shfl_sync(0xffffffffu, mycol, r0 + u);
#pragma unroll
        for (int q = 0; q < NC; ++q) {
          const int w = lane + 32 * q;
          const bool on = (j != padcol) && (w < W);
          v[q][u] = on ? __ldg(V_rm + (size_t)j * W + w) : 0.0;
          x[q][u] = (WITH_B && on) ? __ldg(E_rm + (size_t)j * W + w) : 0.0;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int q = 0; q < NC; ++q) {
          a[q][u] += v[q][u];
          if (WITH_B) b[q][u] += fma(x[q][u], x[q][u], v[q][u]);
        }
    }
  }
#pragma unroll
  for (int q = 0; q < NC; ++q) {
    const int w = lane + 32 * q;
    if (w < W) {
      const size_t o = (size_t)(row0 + row) * W + w;
      outA[o] = ((a[q][0] + a[q][1]) + (a[q][2] + a[q][3])) + ((a[q][4] + a[q][5]) + (a[q][6] + a[q][7]));
      if (WITH_B) outB[o] = ((b[q][0] + b[q][1]) + (b[q][2] + b[q][3])) + ((b[q][4] + b[q][5]) + (b[q][6] + b[q][7]));
    }
  }
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

// H_i[pair(l,l')] = sum_{j in Omega_i} G_jl G_jl'   -- one warp per row of the row layout.
// The G rows of 32 consecutive observed entries are staged in shared memory with cp.async (16-byte
// pieces, double buffered: the copies of the next 32 entries fly while the current ones are used).
// Lane b owns a 4 x 2 block of H_i (rows 4*bi.., columns 2*bj..) of the upper triangle: one 32-byte
// and one 16-byte shared load and 8 FMAs per entry; L <= 20 needs 30 blocks (NBLK = 1), L <= 28 two
// per lane, L <= 32 three.
__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NBLK, int LP, int WARPS, bool EVEN>   // LP: row length in shared memory (multiple of 4, >= L)
__global__ void __launch_bounds__(WARPS * 32) nmtf_gram_kernel(const uint32_t* __restrict__ cols, const int64_t* __restrict__ rowstart,
                                                                int n_loc, int64_t row0, uint32_t padcol, const double* __restrict__ G_rm,
                                                                int L, double* __restrict__ Hrow, int NP) {
  __shared__ __align__(32) double tile[WARPS][2][32][LP];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * WARPS + w;
  if (row >= n_loc) return;
  const int NBA = (L + 3) / 4, NBC = (L + 1) / 2;
  int bi[NBLK], bj[NBLK];
  double acc[NBLK][4][2];
#pragma unroll
  for (int b = 0; b < NBLK; ++b) {
    // enumerate the needed blocks (bj >= 2 bi) row by row: block id -> (bi, bj)
    int id = lane + 32 * b, a = 0;
    while (a < NBA && id >= NBC - 2 * a) { id -= NBC - 2 * a; ++a; }
    bi[b] = a < NBA ? a : -1;
    bj[b] = a < NBA ? 2 * a + id : 0;
#pragma unroll
    for (int x = 0; x < 4; ++x) { acc[b][x][0] = 0.0; acc[b][x][1] = 0.0; }
  }
  // columns L..LP-1 of both buffers stay zero, so partial blocks add nothing
  for (int q = lane; q < 2 * 32 * LP; q += 32) (&tile[w][0][0][0])[q] = 0.0;
  __syncwarp();
  const int64_t beg = rowstart[row], end = rowstart[row + 1];
  constexpr int PD = EVEN ? 2 : 1;             // doubles per copied piece (16-byte pieces need an even L)
  const int pieces = L / PD;
  auto stage = [&](int64_t q0, int buf) {
    const uint32_t mycol = (q0 + lane < end) ? cols[q0 + lane] : padcol;
    const int total = 32 * pieces;
    for (int idx0 = 0; idx0 < total; idx0 += 32) {
      const int idx = idx0 + lane;
      const int r = min(idx, total - 1) / pieces, pc = min(idx, total - 1) - r * pieces;
      const uint32_t j = __shfl_sync(0xffffffffu, mycol, r);
      if (idx < total) {
        double* dst = &tile[w][buf][r][PD * pc];
        if (j == padcol) { dst[0] = 0.0; if (EVEN) dst[1] = 0.0; }
        else if (EVEN) cp_async_16(dst, G_rm + (size_t)j * L + PD * pc);
        else cp_async_8(dst, G_rm + (size_t)j * L + pc);
      }
    }
    cp_async_commit();
  };
  if (beg < end) stage(beg, 0);
  int buf = 0;
  for (int64_t q0 = beg; q0 < end; q0 += 32, buf ^= 1) {
    if (q0 + 32 < end) { stage(q0 + 32, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncwarp();
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
#pragma unroll
      for (int b = 0; b < NBLK; ++b) {
        if (bi[b] < 0) continue;
        const double4 a = *reinterpret_cast<const double4*>(&tile[w][buf][r][4 * bi[b]]);
        const double2 c = *reinterpret_cast<const double2*>(&tile[w][buf][r][2 * bj[b]]);
        acc[b][0][0] = fma(a.x, c.x, acc[b][0][0]); acc[b][0][1] = fma(a.x, c.y, acc[b][0][1]);
        acc[b][1][0] = fma(a.y, c.x, acc[b][1][0]); acc[b][1][1] = fma(a.y, c.y, acc[b][1][1]);
        acc[b][2][0] = fma(a.z, c.x, acc[b][2][0]); acc[b][2][1] = fma(a.z, c.y, acc[b][2][1]);
        acc[b][3][0] = fma(a.w, c.x, acc[b][3][0]); acc[b][3][1] = fma(a.w, c.y, acc[b][3][1]);
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int b = 0; b < NBLK; ++b) {
    if (bi[b] < 0) continue;
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int y = 0; y < 2; ++y) {
        const int a = 4 * bi[b] + x, c = 2 * bj[b] + y;
        if (a <= c && c < L) Hrow[(size_t)(row0 + row) * NP + pair_index(a, c, L)] = acc[b][x][y];
      }
  }
}

// H_i on the FP64 tensor core.  H_i = G_Omega^T G_Omega is a small GEMM per row (L x n_i times n_i x L): with
// mma.sync.m8n8k4 (DMMA) a step takes FOUR observed entries j_0..j_3 and one 8 x 8 block (bi, bj) of H_i,
//   A[r][k] = G[j_k][8 bi + r],  B[k][c] = G[j_k][8 bj + c],  D += A B,
// and the fragment layouts (A: lane holds [lane / 4][lane % 4]; B: [lane % 4][lane / 4]) make both operands of lane
// (g, t) = (lane / 4, lane % 4) the SAME kind of element, x_b = G[j_t][8 b + g]: a lane loads ceil(L / 8) doubles per step
// -- straight from the row-major factor, eight lanes reading 64 contiguous bytes of one row, no shared memory -- and the
// warp issues one DMMA per block of the upper triangle (6 for L <= 24, 10 for L <= 32).  The 4 x 4-block FMA kernel
// below spends two 32-byte shared loads per 16 FMAs and is bound by them (1.2 ms at 10000 x 8000, 30 % observed, L = 20).
// One warp per row; the column indices of 8 steps are one coalesced load handed round by shuffles, the 8 steps' loads are
// issued before their DMMAs.  Fixed summation order per row (independent of the sharding).
__device__ __forceinline__ void dmma_acc_884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int NB>     // NB = ceil(L / 8) column blocks
__global__ void __launch_bounds__(128) nmtf_gram_mma_kernel(const uint32_t* __restrict__ cols, const int64_t* __restrict__ rowstart,
                                                            int n_loc, int64_t row0, uint32_t padcol, const double* __restrict__ G_rm,
                                                            int L, double* __restrict__ Hrow, int NP) {
  const int row = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (row >= n_loc) return;
  const int g = lane >> 2, t = lane & 3;
  constexpr int NBLK = NB * (NB + 1) / 2;
  double acc[NBLK][2];
#pragma unroll
  for (int q = 0; q < NBLK; ++q) { acc[q][0] = 0.0; acc[q][1] = 0.0; }
  bool colok[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) colok[b] = (8 * b + g) < L;
  const int64_t beg = rowstart[row], end = rowstart[row + 1];
  for (int64_t q0 = beg; q0 < end; q0 += 32) {
    const uint32_t mycol = (q0 + lane < end) ? cols[q0 + lane] : padcol;
    double x[8][NB];
#pragma unroll
    for (int s8 = 0; s8 < 8; ++s8) {
      const uint32_t j = __shfl_sync(0xffffffffu, mycol, 4 * s8 + t);
#pragma unroll
      for (int b = 0; b < NB; ++b) x[s8][b] = (j != padcol && colok[b]) ? __ldg(G_rm + (size_t)j * L + 8 * b + g) : 0.0;
    }
#pragma unroll
    for (int s8 = 0; s8 < 8; ++s8) {
      int q = 0;
#pragma unroll
      for (int bi = 0; bi < NB; ++bi)
#pragma unroll
        for (int bj = bi; bj < NB; ++bj, ++q) dmma_acc_884(acc[q][0], acc[q][1], x[s8][bi], x[s8][bj]);
    }
  }
  // lane (g, t) holds D[g][2 t], D[g][2 t + 1] of every block
  int q = 0;
#pragma unroll
  for (int bi = 0; bi < NB; ++bi)
#pragma unroll
    for (int bj = bi; bj < NB; ++bj, ++q) {
#pragma unroll
      for (int y = 0; y < 2; ++y) {
        const int a = 8 * bi + g, c = 8 * bj + 2 * t + y;
        if (a <= c && c < L) Hrow[(size_t)(row0 + row) * NP + pair_index(a, c, L)] = acc[q][y];
      }
    }
}

// L <= 20: the two half-warps of the row's warp take alternate entries; lane h of a half-warp owns a 4 x 4 block
// (bi <= bj) of the 5 x 5 block grid of H_i's upper triangle (15 blocks): two 32-byte shared loads and 16 FMAs per
// entry and lane, i.e. half the shared-memory traffic per FMA of the 4 x 2 mapping above (which was bound by it).
// The halves are added at the end (even entries + odd entries: a fixed order).
template <int WARPS, bool EVEN>
__global__ void __launch_bounds__(WARPS * 32) nmtf_gram44_kernel(const uint32_t* __restrict__ cols, const int64_t* __restrict__ rowstart,
                                                                  int n_loc, int64_t row0, uint32_t padcol, const double* __restrict__ G_rm,
                                                                  int L, double* __restrict__ Hrow, int NP) {
  constexpr int LP = 20;
  __shared__ __align__(32) double tile[WARPS][2][32][LP];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * WARPS + w;
  if (row >= n_loc) return;
  const int half = lane >> 4, hl = lane & 15;
  int bi = -1, bj = 0;
  {
    int id = hl, a = 0;
    while (a < 5 && id >= 5 - a) { id -= 5 - a; ++a; }
    if (a < 5) { bi = a; bj = a + id; }
  }
  double acc[4][4];
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) acc[x][y] = 0.0;
  for (int q = lane; q < 2 * 32 * LP; q += 32) (&tile[w][0][0][0])[q] = 0.0;   // columns L..LP-1 stay zero
  __syncwarp();
  const int64_t beg = rowstart[row], end = rowstart[row + 1];
  constexpr int PD = EVEN ? 2 : 1;
  const int pieces = L / PD;
  auto stage = [&](int64_t q0, int buf) {
    const uint32_t mycol = (q0 + lane < end) ? cols[q0 + lane] : padcol;
    const int total = 32 * pieces;
    for (int idx0 = 0; idx0 < total; idx0 += 32) {
      const int idx = idx0 + lane;
      const int r = min(idx, total - 1) / pieces, pc = min(idx, total - 1) - r * pieces;
      const uint32_t j = __shfl_sync(0xffffffffu, mycol, r);
      if (idx < total) {
        double* dst = &tile[w][buf][r][PD * pc];
        if (j == padcol) { dst[0] = 0.0; if (EVEN) dst[1] = 0.0; }
        else if (EVEN) cp_async_16(dst, G_rm + (size_t)j * L + PD * pc);
        else cp_async_8(dst, G_rm + (size_t)j * L + pc);
      }
    }
    cp_async_commit();
  };
  if (beg < end) stage(beg, 0);
  int buf = 0;
  for (int64_t q0 = beg; q0 < end; q0 += 32, buf ^= 1) {
    if (q0 + 32 < end) { stage(q0 + 32, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncwarp();
    if (bi >= 0) {
#pragma unroll 4
      for (int r = 0; r < 32; r += 2) {
        const double4 a = *reinterpret_cast<const double4*>(&tile[w][buf][r + half][4 * bi]);
        const double4 c = *reinterpret_cast<const double4*>(&tile[w][buf][r + half][4 * bj]);
        acc[0][0] = fma(a.x, c.x, acc[0][0]); acc[0][1] = fma(a.x, c.y, acc[0][1]); acc[0][2] = fma(a.x, c.z, acc[0][2]); acc[0][3] = fma(a.x, c.w, acc[0][3]);
        acc[1][0] = fma(a.y, c.x, acc[1][0]); acc[1][1] = fma(a.y, c.y, acc[1][1]); acc[1][2] = fma(a.y, c.z, acc[1][2]); acc[1][3] = fma(a.y, c.w, acc[1][3]);
        acc[2][0] = fma(a.z, c.x, acc[2][0]); acc[2][1] = fma(a.z, c.y, acc[2][1]); acc[2][2] = fma(a.z, c.z, acc[2][2]); acc[2][3] = fma(a.z, c.w, acc[2][3]);
        acc[3][0] = fma(a.w, c.x, acc[3][0]); acc[3][1] = fma(a.w, c.y, acc[3][1]); acc[3][2] = fma(a.w, c.z, acc[3][2]); acc[3][3] = fma(a.w, c.w, acc[3][3]);
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const double v = acc[x][y] + __shfl_xor_sync(0xffffffffu, acc[x][y], 16);
      const int a = 4 * bi + x, c = 4 * bj + y;
      if (half == 0 && bi >= 0 && a <= c && c < L) Hrow[(size_t)(row0 + row) * NP + pair_index(a, c, L)] = v;
    }
}

// c[k][l] = sum_i F[i][k] P[i][l]    (one block per (k,l), fixed-order tree)
static __global__ void nmtf_reduce_c_kernel(const double* __restrict__ F, const double* __restrict__ P, int64_t I, int K, int L,
                                     double* __restrict__ c) {
  __shared__ double sh[32];
  const int k = blockIdx.x / L, l = blockIdx.x % L;
  double a = 0.0;
  for (int64_t i = threadIdx.x; i < I; i += blockDim.x) a = fma(F[i * K + k], P[i * L + l], a);
  a = block_sum(a, sh);
  if (threadIdx.x == 0) c[blockIdx.x] = a;
}

// TT[pk][pl] = sum_i F[i][ka] F[i][kb] H[i][pl]   (pk = pair (ka <= kb) of K, pl = column of H, W columns).
// Stage 1: block (32 pl) x (8 row slices) accumulates a tile of RT_PKT pairs x 32 columns over one of RT_RC row
// chunks: H is read once per (pair tile, row) instead of once per pair, the F products of the chunk are staged in
// shared memory.  Stage 2 folds the RT_RC partial tiles in a fixed order.
constexpr int RT_PKT = 14;   // pairs per block
constexpr int RT_RC = 32;    // row chunks
constexpr int RT_ROWS = 64;  // rows of F products staged at a time
static __global__ void __launch_bounds__(256) nmtf_reduce_T_kernel(const double* __restrict__ F, const double* __restrict__ H, int64_t I, int K,
                                                            int NPK, int W, double* __restrict__ part) {
  __shared__ double ff[RT_ROWS][RT_PKT];
  __shared__ double red[8][RT_PKT][33];
  __shared__ int kab[RT_PKT][2];
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int pl = blockIdx.x * 32 + threadIdx.x;
  const int pk0 = blockIdx.y * RT_PKT;
  const int rc = blockIdx.z;
  const int64_t chunk = (I + RT_RC - 1) / RT_RC;
  const int64_t i_begin = rc * chunk, i_end = min(I, i_begin + chunk);
  if (tid < RT_PKT) {
    int ka = 0, kb = 0;
    if (pk0 + tid < NPK) pair_decode(pk0 + tid, K, ka, kb);
    kab[tid][0] = ka; kab[tid][1] = kb;
  }
  double acc[RT_PKT];
#pragma unroll
  for (int p = 0; p < RT_PKT; ++p) acc[p] = 0.0;
  for (int64_t i0 = i_begin; i0 < i_end; i0 += RT_ROWS) {
    __syncthreads();
    for (int q = tid; q < RT_ROWS * RT_PKT; q += 256) {
      const int r = q / RT_PKT, p = q - r * RT_PKT;
      const int64_t i = i0 + r;
      ff[r][p] = (i < i_end && pk0 + p < NPK) ? F[i * K + kab[p][0]] * F[i * K + kab[p][1]] : 0.0;
    }
    __syncthreads();
    if (pl < W) {
      for (int r = threadIdx.y; r < RT_ROWS; r += 8) {
        const int64_t i = i0 + r;
        if (i >= i_end) break;
        const double hv = H[i * W + pl];
#pragma unroll
        for (int p = 0; p < RT_PKT; ++p) acc[p] = fma(ff[r][p], hv, acc[p]);
      }
    }
  }
#pragma unroll
  for (int p = 0; p < RT_PKT; ++p) red[threadIdx.y][p][threadIdx.x] = acc[p];
  __syncthreads();
  if (pl < W) {
    for (int p = threadIdx.y; p < RT_PKT; p += 8) {
      if (pk0 + p >= NPK) continue;
      double t = 0.0;
      for (int z = 0; z < 8; ++z) t += red[z][p][threadIdx.x];
      part[((size_t)rc * NPK + pk0 + p) * W + pl] = t;
    }
  }
}
static __global__ void nmtf_reduce_T_fold_kernel(const double* __restrict__ part, int64_t n, double* __restrict__ TT) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  double t = 0.0;
  for (int rc = 0; rc < RT_RC; ++rc) t += part[(size_t)rc * n + q];
  TT[q] = t;
}

// The K*L sequential S updates on the small tensors (single CTA).
//   Gibbs: S_kl ~ TN(muS, tauS)  bnmtf_gibbs.py:200-203      ICM: max(max(0,mu),0.1)  nmtf_icm.py:203-207
// single_d >= 0: only report (tau, mu) of entry d from the current state.
// One barrier per step: c_d is read only at step d, so the move of step d-1 needs to reach the entries > d only.  While
// thread 0 computes the conditional of step d -- it applies delta_{d-1} to ITS entry c_d itself, the threads having applied
// delta_0 .. delta_{d-2} to it in the steps before -- the other entries take delta_{d-1}; the barrier at the end of the step
// publishes delta_d (double buffered) and closes the updates.  Every c_d sees the same FMAs in the same order as in the
// two-barrier form (conditional -> barrier -> update -> barrier), at about half its step time.
template <int MODE>
__global__ void __launch_bounds__(512) nmtf_s_sweep_kernel(const double* __restrict__ c_in, const double* __restrict__ TT, int K, int L, double* S,
                                    const double* __restrict__ lamS, const double* __restrict__ scal, uint64_t seed,
                                    uint32_t iteration, double* mu_out, double* tau_out, int single_d, const uint32_t* iter_ptr) {
  extern __shared__ double shs[];
  const int KL = K * L, NPL = L * (L + 1) / 2;
  double* c = shs;            // [KL]
  double* sv = c + KL;        // [KL]
  double* bc = sv + KL;       // [2] the moves of the last two steps
  double* mat = bc + 2;       // [KL][6] Gibbs: random material of every entry, prepared by all threads up front
  if (iter_ptr) iteration = *iter_ptr;
  for (int d = threadIdx.x; d < KL; d += blockDim.x) {
    c[d] = c_in[d]; sv[d] = S[d];
    if (MODE == M_GIBBS && single_d < 0) tn_pre_material_store(seed, RNG_S, (uint64_t)d, 0u, iteration, mat + TN_PRE_DOUBLES * d);
  }
  if (threadIdx.x < 2) bc[threadIdx.x] = 0.0;
  __syncthreads();
  const double tau = scal[0];
  const int d_begin = single_d >= 0 ? single_d : 0, d_end = single_d >= 0 ? single_d + 1 : KL;
  for (int d = d_begin; d < d_end; ++d) {
    const int k = d / L, l = d - k * L;
    // the move of the step before (0 at the first step) and the column of T it goes with
    const double dprev = bc[(d + 1) & 1];
    const int kp = d > d_begin ? (d - 1) / L : 0, lp = d > d_begin ? (d - 1) - kp * L : 0;
    if (threadIdx.x == 0) {
      double cd = c[d];
      if (dprev != 0.0)
        cd = fma(-dprev, TT[(size_t)pair_index(min(kp, k), max(kp, k), K) * NPL + pair_index(min(lp, l), max(lp, l), L)], cd);
      const double Tdd = TT[(size_t)pair_index(k, k, K) * NPL + pair_index(l, l, L)];
      const double lam_d = lamS[d];
      const double ct = tau * Tdd;                                        // bnmtf_gibbs.py:279
      const double cm = (-lam_d + tau * (cd + sv[d] * Tdd)) / ct;         // bnmtf_gibbs.py:283
      double x_new;
      if (MODE == M_GIBBS && single_d < 0) {
        Rng rng(seed, RNG_S, (uint64_t)d, 0u, iteration);
        rng.draw = 2;                                   // Philox blocks 0 and 1 made the material
        const TnPre m = tn_pre_load(mat + TN_PRE_DOUBLES * d);
        double cm2;
        x_new = tn_draw_fast(-lam_d + tau * (cd + sv[d] * Tdd), ct, m, rng, cm2);
      } else if (MODE == M_GIBBS) {
        x_new = sv[d];
      } else {
        x_new = fmax(fmax(0.0, cm), MINIMUM_TN);
      }
      if (mu_out) { mu_out[d] = cm; tau_out[d] = ct; }
      if (single_d >= 0) x_new = sv[d];
      bc[d & 1] = x_new - sv[d];
      sv[d] = x_new;
    }
    if (dprev != 0.0) {          // entries beyond d take the move of step d - 1 (entry d took it above, the ones before are done)
      for (int d2 = threadIdx.x; d2 < KL; d2 += blockDim.x) {
        if (d2 <= d) continue;
        const int k2 = d2 / L, l2 = d2 - k2 * L;
        const int pk = pair_index(min(kp, k2), max(kp, k2), K), pl = pair_index(min(lp, l2), max(lp, l2), L);
        c[d2] = fma(-dprev, TT[(size_t)pk * NPL + pl], c[d2]);
      }
    }
    __syncthreads();
  }
  if (single_d < 0)
    for (int d = threadIdx.x; d < KL; d += blockDim.x) S[d] = sv[d];
}

template <int MODE>
__global__ void nmtf_lambda_kernel(const double* __restrict__ F, int64_t I, int K, const double* __restrict__ G, int64_t J, int L,
                                   Hyper h, double* colsum, double* lamF, double* lamG, uint64_t seed, uint32_t iteration,
                                   int update, const uint32_t* iter_ptr) {
  __shared__ double sh[32];
  if (iter_ptr) iteration = *iter_ptr;
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

// ---------------------------------------------------------------------------------
// BNMTF VB (code/models/bnmtf_vb.py)
// ---------------------------------------------------------------------------------
// Pair-layout gather source of update_F / update_G (bnmtf_vb.py:337-339,365-366):
//   out[c][i] = ( w, var + w^2 ),  w = sum_q EX[i][q] ES(q,c),
//   var = sum_q (VX+EX^2)[i][q] (VS+ES^2)(q,c) - sum_q EX[i][q]^2 ES(q,c)^2      (var_SkG / var_FSl)
static __global__ void vb_combine_km_kernel(const double* __restrict__ EX, const double* __restrict__ VX, int64_t n, int kq,
                                     const double* __restrict__ ES, const double* __restrict__ VS, int s_q, int s_c,
                                     double* __restrict__ out_km, int64_t ld, double* __restrict__ outB_e, double* __restrict__ outB_q) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (i >= ld) return;
  double w = 0.0, a2 = 0.0, b2 = 0.0;
  if (i < n) {
    for (int q = 0; q < kq; ++q) {
      const double ex = EX[i * kq + q], vx = VX[i * kq + q];
      const double es = ES[q * s_q + c * s_c], vs = VS[q * s_q + c * s_c];
      w = fma(ex, es, w);
      a2 = fma(vx + ex * ex, vs + es * es, a2);
      b2 = fma(ex * ex, es * es, b2);
    }
  }
  reinterpret_cast<double2*>(out_km)[(size_t)c * ld + i] = make_double2(w, (a2 - b2) + w * w);
  if (outB_e) {              // block-major copies for the per-row-chain cluster sweep (kernels_v5.cuh)
    const size_t o = ((size_t)(c >> 1) * ld + i) * 2 + (c & 1);
    outB_e[o] = w; outB_q[o] = (a2 - b2) + w * w;
  }
}

// out[k][l] = sum_i f(i,k) P[i][l]  with f = F (Fvar == nullptr) or Fvar + F^2
static __global__ void nmtf_reduce_c2_kernel(const double* __restrict__ F, const double* __restrict__ Fvar, const double* __restrict__ P,
                                      int64_t I, int K, int L, double* __restrict__ out) {
  __shared__ double sh[32];
  const int k = blockIdx.x / L, l = blockIdx.x % L;
  double a = 0.0;
  for (int64_t i = threadIdx.x; i < I; i += blockDim.x) {
    const double f = F[i * K + k];
    a = fma(Fvar ? Fvar[i * K + k] + f * f : f, P[i * L + l], a);
  }
  a = block_sum(a, sh);
  if (threadIdx.x == 0) out[blockIdx.x] = a;
}

__device__ __forceinline__ double tmat(const double* TT, int k, int l, int k2, int l2, int K, int L, int NPL) {
  return TT[(size_t)pair_index(min(k, k2), max(k, k2), K) * NPL + pair_index(min(l, l2), max(l, l2), L)];
}

// The K*L sequential update_S(k,l) + update_exp_S(k,l) (bnmtf_vb.py:220-223,350-362,393-396) on the small tensors:
//   tau_S = Etau BS[k][l],  diff = c_d + ES_d T[d,d],
//   covG = sum_{k'!=k} ES(k',l) PA[pair(k,k')][l],  covF = sum_{l'!=l} ES(k,l') QC[pair(l,l')][k]
// mode: -1 full sweep, -2 dry (ELBO terms of the stored state), d >= 0 report (tau, mu) of entry d.
// misc[0] = sum_d q-terms of S for the ELBO (bnmtf_vb.py:290-292), misc[1] = sum lambdaS*ES (:268).
static __global__ void __launch_bounds__(512) nmtf_s_sweep_vb_kernel(const double* __restrict__ c_in, const double* __restrict__ TT, const double* __restrict__ BS,
                                       const double* __restrict__ PA, const double* __restrict__ QC, int K, int L, double* Sexp,
                                       double* Svar, double* Smu, double* Stau, const double* __restrict__ lamS,
                                       const double* __restrict__ scal, int mode, double* out_tau, double* out_mu, double* misc) {
  extern __shared__ double shs[];
  const int KL = K * L, NPL = L * (L + 1) / 2;
  double* c = shs;            // [KL]
  double* es = c + KL;        // [KL]
  double* bc = es + KL;       // [1]
  for (int d = threadIdx.x; d < KL; d += blockDim.x) { c[d] = c_in[d]; es[d] = Sexp[d]; }
  __syncthreads();
  const double etau = scal[0];
  double acc_q = 0.0, acc_prior = 0.0;
  const int d_begin = mode >= 0 ? mode : 0, d_end = mode >= 0 ? mode + 1 : KL;
  for (int d = d_begin; d < d_end; ++d) {
    const int k = d / L, l = d - k * L;
    if (threadIdx.x == 0) {
      double ct, cm, en, vn;
      if (mode == -2) {
        ct = Stau[d]; cm = Smu[d]; en = es[d]; vn = Svar[d];
      } else {
        const double Tdd = tmat(TT, k, l, k, l, K, L, NPL);
        ct = etau * BS[d];
        double covG = 0.0, covF = 0.0;
        for (int k2 = 0; k2 < K; ++k2)
          if (k2 != k) covG = fma(es[k2 * L + l], PA[(size_t)pair_index(min(k, k2), max(k, k2), K) * L + l], covG);
        for (int l2 = 0; l2 < L; ++l2)
          if (l2 != l) covF = fma(es[k * L + l2], QC[(size_t)pair_index(min(l, l2), max(l, l2), L) * K + k], covF);
        cm = (-lamS[d] + etau * (c[d] + es[d] * Tdd) - etau * covG - etau * covF) / ct;
        tn_moments(cm, ct, en, vn);
      }
      if (mode >= 0) {
        out_tau[0] = ct; out_mu[0] = cm;
        bc[0] = 0.0;
      } else {
        acc_q += -0.5 * log(ct) + log_half_erfc_ref(-cm * sqrt(ct) / 1.4142135623730951) + 0.5 * ct * (vn + (en - cm) * (en - cm));
        acc_prior += lamS[d] * en;
        bc[0] = en - es[d];
        if (mode == -1) { Smu[d] = cm; Stau[d] = ct; Svar[d] = vn; es[d] = en; }
      }
    }
    __syncthreads();
    const double delta = bc[0];
    if (delta != 0.0)
      for (int d2 = threadIdx.x; d2 < KL; d2 += blockDim.x) {
        const int k2 = d2 / L, l2 = d2 - k2 * L;
        c[d2] = fma(-delta, tmat(TT, k, l, k2, l2, K, L, NPL), c[d2]);
      }
    __syncthreads();
  }
  if (mode == -1)
    for (int d = threadIdx.x; d < KL; d += blockDim.x) Sexp[d] = es[d];
  if (threadIdx.x == 0 && mode < 0) { misc[0] = acc_q; misc[1] = acc_prior; }
}

// The same sweep with everything that is not the conditional itself taken off thread 0's serial path (used while
// K*L <= 2 * blockDim.x and the arrays fit 48 KB of shared memory; nmtf_s_sweep_vb_kernel is the general form):
//   * covG_d and covF_d are linear in ES: computed once for every entry in parallel, then maintained like c -- the
//     update of entry (k,l) by delta adds delta PA[pair(k2,k)][l] to covG of the entries (k2,l) and delta
//     QC[pair(l2,l)][k] to covF of the entries (k,l2);
//   * the row of T, PA and QC values an entry needs for the update of step d are loaded before thread 0 starts the
//     conditional of step d, so their L2 latency is hidden behind it;
//   * T_dd, BS, lambdaS sit in shared memory; the q(S) terms of the ELBO (two logarithms and an erfc per entry,
//     bnmtf_vb.py:290-292) and sum lambdaS ES are evaluated by all threads after the sweep (fixed-order block sum).
// Entry reports (mode >= 0) see exactly the from-scratch sums of the general kernel.
static __global__ void __launch_bounds__(512) nmtf_s_sweep_vb_fast_kernel(const double* __restrict__ c_in, const double* __restrict__ TT,
                                       const double* __restrict__ BS, const double* __restrict__ PA, const double* __restrict__ QC,
                                       int K, int L, double* Sexp, double* Svar, double* Smu, double* Stau,
                                       const double* __restrict__ lamS, const double* __restrict__ scal, int mode, double* out_tau,
                                       double* out_mu, double* misc) {
  extern __shared__ double shs[];
  __shared__ double red[32];
  const int KL = K * L, NPL = L * (L + 1) / 2;
  double* c = shs;            // [KL] running c_d
  double* es = c + KL;        // [KL]
  double* cg = es + KL;       // [KL] running covG_d
  double* cf = cg + KL;       // [KL] running covF_d
  double* tdd = cf + KL;      // [KL]
  double* bs = tdd + KL;      // [KL]
  double* lam = bs + KL;      // [KL]
  double* bc = lam + KL;      // [2] the moves of the last two steps
  constexpr int ME = 2;       // entries per thread
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int d = tid; d < KL; d += nt) {
    const int k = d / L, l = d - k * L;
    c[d] = c_in[d]; es[d] = Sexp[d]; tdd[d] = tmat(TT, k, l, k, l, K, L, NPL); bs[d] = BS[d]; lam[d] = lamS[d];
  }
  __syncthreads();
  const double etau = scal[0];
  if (mode != -2) {
    for (int d = tid; d < KL; d += nt) {
      const int k = d / L, l = d - k * L;
      double covG = 0.0, covF = 0.0;
      for (int k2 = 0; k2 < K; ++k2)
        if (k2 != k) covG = fma(es[k2 * L + l], PA[(size_t)pair_index(min(k, k2), max(k, k2), K) * L + l], covG);
      for (int l2 = 0; l2 < L; ++l2)
        if (l2 != l) covF = fma(es[k * L + l2], QC[(size_t)pair_index(min(l, l2), max(l, l2), L) * K + k], covF);
      cg[d] = covG; cf[d] = covF;
    }
    __syncthreads();
    const int d_begin = mode >= 0 ? mode : 0, d_end = mode >= 0 ? mode + 1 : KL;
    // One barrier per step (as nmtf_s_sweep_kernel): the running c_d, covG_d, covF_d are read at step d only, so the move of
    // step d - 1 needs to reach the entries > d; thread 0 applies it to ITS entry d itself and computes the update of step d
    // while the other entries take it.  Same FMAs in the same order on every entry.
    if (tid < 2) bc[tid] = 0.0;
    __syncthreads();
    for (int d = d_begin; d < d_end; ++d) {
      const int k = d / L, l = d - k * L;
      const double dprev = bc[(d + 1) & 1];
      const int kp = d > d_begin ? (d - 1) / L : 0, lp = d > d_begin ? (d - 1) - kp * L : 0;
      if (tid == 0) {
        double cd = c[d], cgd = cg[d], cfd = cf[d];
        if (dprev != 0.0) {
          cd = fma(-dprev, tmat(TT, kp, lp, k, l, K, L, NPL), cd);
          if (l == lp && k != kp) cgd = fma(dprev, PA[(size_t)pair_index(min(kp, k), max(kp, k), K) * L + lp], cgd);
          else if (k == kp && l != lp) cfd = fma(dprev, QC[(size_t)pair_index(min(lp, l), max(lp, l), L) * K + kp], cfd);
        }
        const double ct = etau * bs[d];
        const double cm = (-lam[d] + etau * (cd + es[d] * tdd[d]) - etau * cgd - etau * cfd) / ct;
        double en, vn;
        tn_moments(cm, ct, en, vn);
        if (mode >= 0) {
          out_tau[0] = ct; out_mu[0] = cm;
          bc[d & 1] = 0.0;
        } else {
          bc[d & 1] = en - es[d];
          Smu[d] = cm; Stau[d] = ct; Svar[d] = vn; es[d] = en;
        }
      }
      if (dprev != 0.0 && mode < 0) {   // entries beyond d take the move of step d - 1
#pragma unroll
        for (int q = 0; q < ME; ++q) {
          const int d2 = tid + q * nt;
          if (d2 < KL && d2 > d) {
            const int k2 = d2 / L, l2 = d2 - k2 * L;
            c[d2] = fma(-dprev, tmat(TT, kp, lp, k2, l2, K, L, NPL), c[d2]);
            if (l2 == lp && k2 != kp) cg[d2] = fma(dprev, PA[(size_t)pair_index(min(kp, k2), max(kp, k2), K) * L + lp], cg[d2]);
            else if (k2 == kp && l2 != lp) cf[d2] = fma(dprev, QC[(size_t)pair_index(min(lp, l2), max(lp, l2), L) * K + kp], cf[d2]);
          }
        }
      }
      __syncthreads();
    }
    if (mode == -1)
      for (int d = tid; d < KL; d += nt) Sexp[d] = es[d];
  }
  if (mode < 0) {   // ELBO terms from the (new or stored) parameters of every entry
    double aq = 0.0, ap = 0.0;
    for (int d = tid; d < KL; d += nt) {
      const double ct = Stau[d], cm = Smu[d], en = es[d], vn = Svar[d];
      aq += -0.5 * log(ct) + log_half_erfc_ref(-cm * sqrt(ct) / 1.4142135623730951) + 0.5 * ct * (vn + (en - cm) * (en - cm));
      ap += lam[d] * en;
    }
    aq = block_sum(aq, red);
    ap = block_sum(ap, red);
    if (tid == 0) { misc[0] = aq; misc[1] = ap; }
  }
}

// third term of exp_square_diff (bnmtf_vb.py:323): sum_Omega VarF ((ES EG^T)^2 - ES^2 EG^2^T)
//   = sum_j sum_k C_jk ((sum_l ES_kl EG_jl)^2 - sum_l ES_kl^2 EG_jl^2),  C = masked column sums of VarF.
// T3_BLOCKS blocks over contiguous ranges of j, then a fixed-order fold of their partial sums (nmtf_t3_fold_kernel).
constexpr int T3_BLOCKS = 128;
static __global__ void __launch_bounds__(256) nmtf_t3_kernel(const double* __restrict__ C, const double* __restrict__ EG, const double* __restrict__ ES, int64_t J,
                               int K, int L, double* __restrict__ part) {
  __shared__ double sh[32];
  double a = 0.0;
  const int64_t chunk = (J + gridDim.x - 1) / gridDim.x;
  const int64_t j_end = min(J, (int64_t)(blockIdx.x + 1) * chunk);
  for (int64_t j = (int64_t)blockIdx.x * chunk + threadIdx.x; j < j_end; j += blockDim.x) {
    for (int k = 0; k < K; ++k) {
      double w = 0.0, w2 = 0.0;
      for (int l = 0; l < L; ++l) {
        const double t = ES[k * L + l] * EG[j * L + l];
        w += t; w2 = fma(t, t, w2);
      }
      a = fma(C[j * K + k], w * w - w2, a);
    }
  }
  a = block_sum(a, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = a;
}
static __global__ void nmtf_t3_fold_kernel(const double* __restrict__ part, int n, double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double a = 0.0;
  for (int q = 0; q < n; ++q) a += part[q];
  out[0] = a;
}

// VB lambda updates of both ARD blocks (bnmtf_vb.py:207-213,326-334,383-391): block per column.
// state[4][K+L]: alpha*, beta*, E[lambda], E[log lambda] (F columns first, then G columns).
static __global__ void nmtf_lambda_vb_kernel(const double* __restrict__ F, int64_t I, int K, const double* __restrict__ G, int64_t J, int L,
                                      Hyper h, double* colsum, double* state, double* lamF, double* lamG, int update) {
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
      const int KL = K + L;
      const double alpha = h.alpha0 + (double)n, beta = h.beta0 + a;
      state[0 * KL + c] = alpha; state[1 * KL + c] = beta;
      state[2 * KL + c] = alpha / beta; state[3 * KL + c] = digamma(alpha) - log(beta);
      (isF ? lamF : lamG)[col] = alpha / beta;
    }
  }
}

// update_tau / update_exp_tau / elbo of bnmtf_vb (bnmtf_vb.py:225-229,248-324).  Single block.
// misc: [0] q-terms of S, [1] sum lambdaS*ES, [2] T3.
static __global__ void __launch_bounds__(1024) nmtf_vb_finalize_kernel(const double* __restrict__ rsF, int64_t nF, const double* __restrict__ rsG, int64_t nG, int K,
                                        int L, Hyper h, double sumlog_lamS, const double* __restrict__ colsum,
                                        const double* __restrict__ state, const double* __restrict__ misc, double* scal,
                                        double* stats_out, int update, int64_t I, int64_t J, uint32_t* iter_ctr) {
  __shared__ double sh[32];
  __shared__ double tot[2 * NRS];
  uint32_t iteration = 0;
  if (iter_ctr) { iteration = *iter_ctr; stats_out += (size_t)iteration * 16; }   // graph replays: see finalize_kernel
  for (int sidx = 0; sidx < 2; ++sidx) {
    const double* rs = sidx == 0 ? rsF : rsG;
    const int64_t n = sidx == 0 ? nF : nG;
    for (int f = 0; f < NRS; ++f) {
      double a = 0.0;
      for (int64_t i = threadIdx.x; i < n; i += blockDim.x) a += rs[i * NRS + f];
      a = block_sum(a, sh);
      if (threadIdx.x == 0) tot[sidx * NRS + f] = a;
    }
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const double rss = tot[NRS + RS_RSS];
  for (int i = 0; i < 16; ++i) stats_out[i] = 0.0;
  stats_out[0] = rss; stats_out[1] = tot[NRS + RS_SUM_E]; stats_out[2] = tot[NRS + RS_SUM_RCE];
  // exp_square_diff: residual + (second + fourth term, from the G sweep) + third term
  const double esd = rss + tot[NRS + RS_ESDVAR] + misc[2];
  const double alpha_s = h.alphatau + h.size_Omega / 2.0, beta_s = h.betatau + 0.5 * esd;
  double tau = scal[SC_TAU], exp_logtau = scal[SC_LOGTAU], a_s = scal[SC_ALPHA_S], b_s = scal[SC_BETA_S];
  if (update) {
    a_s = alpha_s; b_s = beta_s; tau = a_s / b_s; exp_logtau = digamma(a_s) - log(b_s);
    scal[SC_TAU] = tau; scal[SC_LOGTAU] = exp_logtau; scal[SC_ALPHA_S] = a_s; scal[SC_BETA_S] = b_s;
  }
  const double log2pi = 1.8378770664093454836;
  double elbo = h.size_Omega / 2.0 * (exp_logtau - log2pi) - tau / 2.0 * esd;
  if (h.ARD) {
    const int KL = K + L;
    for (int blk = 0; blk < 2; ++blk) {
      const int c0 = blk == 0 ? 0 : K, c1 = blk == 0 ? K : K + L;
      const double n = blk == 0 ? (double)I : (double)J;
      double s_ll = 0.0, s_l = 0.0, s_logl = 0.0, s_lcol = 0.0, q = 0.0;
      for (int c = c0; c < c1; ++c) {
        const double al = state[0 * KL + c], be = state[1 * KL + c], el = state[2 * KL + c], ell = state[3 * KL + c];
        s_ll += ell; s_l += el; s_logl += log(el); s_lcol += el * colsum[c];
        q += -al * log(be) + lgamma(al) - (al - 1.0) * ell + be * el;
      }
      elbo += h.alpha0 * log(h.beta0) - lgamma(h.alpha0) + (h.alpha0 - 1.0) * s_ll - h.beta0 * s_l;
      elbo += n * s_logl - s_lcol;
      elbo += q;
    }
  } else {
    elbo += h.sumlog_lamU - tot[RS_PRIOR] + h.sumlog_lamV - tot[NRS + RS_PRIOR];
  }
  elbo += sumlog_lamS - misc[1];
  elbo += h.alphatau * log(h.betatau) - lgamma(h.alphatau) + (h.alphatau - 1.0) * exp_logtau - h.betatau * tau;
  elbo += tot[RS_ELBOQ] + (double)I * K / 2.0 * log2pi + tot[NRS + RS_ELBOQ] + (double)J * L / 2.0 * log2pi
          + misc[0] + (double)K * L / 2.0 * log2pi;
  elbo += -a_s * log(b_s) + lgamma(a_s) - (a_s - 1.0) * exp_logtau + b_s * tau;
  stats_out[3] = tau; stats_out[4] = beta_s; stats_out[5] = esd; stats_out[6] = elbo; stats_out[8] = exp_logtau;
  stats_out[9] = alpha_s;
  if (iter_ctr && update) *iter_ctr = iteration + 1;
}

// nmtf_np.update_S(k,l) (nmtf_np.py:181-188): S_kl <- S_kl * sum_i F_ik num_i / sum_i F_ik den_i with the row parts
// num_i = sum_j M_ij R_ij/Rhat_ij G_jl, den_i = sum_j M_ij G_jl.  One block; writes S_kl (and the new value to out).
static __global__ void nmtf_np_s_update_kernel(const double* __restrict__ F, const double* __restrict__ num, const double* __restrict__ den,
                                        int64_t I, int K, int k, double* S_kl, double* out, int apply) {
  __shared__ double sh[32];
  double a = 0.0, b = 0.0;
  for (int64_t i = threadIdx.x; i < I; i += blockDim.x) {
    const double f = F[i * K + k];
    a = fma(f, num[i], a); b = fma(f, den[i], b);
  }
  a = block_sum(a, sh); b = block_sum(b, sh);
  if (threadIdx.x == 0) {
    const double v = S_kl[0] * a / b;
    if (out) out[0] = v;
    if (apply) S_kl[0] = v;
  }
}

// ---- nmtf_np.update_S, the K*L sequential updates in ONE pass each over stored predictions -------------------------------
// The reference recomputes R_pred = F S G^T for every (k,l) (nmtf_np.py:181-188).  An update of S_kl by delta changes the
// prediction of an observed entry by delta F_ik G_jl, and the denominator sum_ij M_ij F_ik G_jl = sum_i F_ik (M G)_il does
// not depend on S at all.  So the predictions at the observed entries are kept in an array of the row layout (8 bytes per
// slot), (M G) is taken once per iteration, and the pass of entry (k,l) first applies the move of the entry before it and
// then accumulates the row parts of its numerator: 28 bytes of memory traffic per observed entry and pass instead of a
// K-column gather pass.  One warp per row, one entry per lane and step, whole rows of the row-major G within reach of L1.
//
// predictions with the current F, S, G: p_ij = sum_l (F_i S)_l G_jl
static __global__ void __launch_bounds__(256) nmtf_np_pred_kernel(const uint32_t* __restrict__ cols, const int64_t* __restrict__ rowstart,
                                                                  int n_loc, int64_t row0, uint32_t padcol, const double* __restrict__ F_rm, int K,
                                                                  const double* __restrict__ S, int L, const double* __restrict__ G_rm,
                                                                  double* __restrict__ phat) {
  __shared__ double fs_all[8][32];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + w;
  if (row >= n_loc) return;
  double* fs = fs_all[w];
  if (lane < L) {
    double a = 0.0;
    for (int k = 0; k < K; ++k) a = fma(F_rm[(size_t)(row0 + row) * K + k], S[k * L + lane], a);
    fs[lane] = a;
  }
  __syncwarp();
  const int64_t beg = rowstart[row], end = rowstart[row + 1];
  for (int64_t q = beg + lane; q < end; q += 32) {
    const uint32_t j = cols[q];
    double pr = 1.0;                                    // padded slots: never read back as data
    if (j != padcol) {
      pr = 0.0;
      for (int l = 0; l < L; ++l) pr = fma(fs[l], __ldg(G_rm + (size_t)j * L + l), pr);
    }
    phat[q] = pr;
  }
}

// pass of entry (k,l): p_ij += delta_prev F_i,kprev G_j,lprev (kprev < 0: nothing to apply), num_i = sum_j R_ij / p_ij G_jl
static __global__ void __launch_bounds__(256) nmtf_np_s_pass_kernel(const double* __restrict__ vals, const uint32_t* __restrict__ cols,
                                                                    const int64_t* __restrict__ rowstart, int n_loc, int64_t row0, uint32_t padcol,
                                                                    const double* __restrict__ F_rm, int K, const double* __restrict__ G_rm, int L,
                                                                    double* __restrict__ phat, int kprev, int lprev, const double* __restrict__ delta_prev,
                                                                    int l, double* __restrict__ num_out) {
  const int row = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (row >= n_loc) return;
  const bool apply = kprev >= 0;
  const double fprev = apply ? F_rm[(size_t)(row0 + row) * K + kprev] * delta_prev[0] : 0.0;
  const int64_t beg = rowstart[row], end = rowstart[row + 1];
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t q0 = beg + lane; q0 < end; q0 += 128) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t q = q0 + 32 * u;
      if (q < end) {
        const uint32_t j = cols[q];
        if (j != padcol) {
          double pr = phat[q];
          if (apply) { pr = fma(fprev, __ldg(G_rm + (size_t)j * L + lprev), pr); phat[q] = pr; }
          acc[u] = fma(vals[q] / pr, __ldg(G_rm + (size_t)j * L + l), acc[u]);
        }
      }
    }
  }
  const double t = warp_sum((acc[0] + acc[1]) + (acc[2] + acc[3]));
  if (lane == 0) num_out[row0 + row] = t;
}

// S_kl <- S_kl * sum_i F_ik num_i / sum_i F_ik (M G)_il; the move goes to delta_out for the next pass.  One block.
static __global__ void nmtf_np_s_apply_kernel(const double* __restrict__ F, const double* __restrict__ num, const double* __restrict__ MG,
                                              int64_t I, int K, int L, int k, int l, double* S_kl, double* delta_out) {
  __shared__ double sh[32];
  double a = 0.0, b = 0.0;
  for (int64_t i = threadIdx.x; i < I; i += blockDim.x) {
    const double f = F[i * K + k];
    a = fma(f, num[i], a); b = fma(f, MG[i * L + l], b);
  }
  a = block_sum(a, sh); b = block_sum(b, sh);
  if (threadIdx.x == 0) {
    const double old = S_kl[0], v = old * a / b;
    S_kl[0] = v;
    delta_out[0] = v - old;
  }
}

static __global__ void init_vector_kernel(double* __restrict__ X, int64_t n, const double* __restrict__ lam, int random, uint64_t seed,
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
