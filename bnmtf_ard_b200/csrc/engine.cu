// Host side of the engine: model handle, packing of R/M, the iteration loop and
// the C ABI declared in include/bnmtf_b200.h.  No torch types anywhere.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <vector>

#include "../../include/bnmtf_b200.h"
#include "kernels.cuh"
#include "kernels_v2.cuh"
#include "sweep4.h"
#include "sweep5.h"
#include "kernels_nmtf.cuh"
#include "kernels_kmeans.cuh"
#include "kernels_summary.cuh"
#include "kernels_batch.cuh"
#include "kernels_long.cuh"

using namespace bnmtf;

static thread_local std::string g_err;
const char* bnmtf_last_error(void) { return g_err.c_str(); }
int bnmtf_abi_version(void) { return 1; }

#define CK(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) {                                                                         \
      char buf_[512];                                                                                \
      snprintf(buf_, sizeof buf_, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      g_err = buf_;                                                                                  \
      return 1;                                                                                      \
    }                                                                                                \
  } while (0)

#define FAIL(...)                                  \
  do {                                             \
    char buf_[512];                                \
    snprintf(buf_, sizeof buf_, __VA_ARGS__);      \
    g_err = buf_;                                  \
    return 1;                                      \
  } while (0)

static inline int64_t roundup(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

int bnmtf_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// ---- NCCL, loaded lazily (only multi-GPU jobs need it) ------------------------
namespace {
struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load() {
    if (lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) { lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) return false;
    GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
    AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
    return GetUniqueId && CommInitRank && AllGather && CommDestroy && GetErrorString;
  }
};
NcclApi g_nccl;
constexpr int NCCL_FLOAT64 = 8;

// one sweep direction: rows of this side are updated, columns index the other factor
struct Side {
  int64_t n_glob = 0, n_loc = 0, row0 = 0, n_shard = 0, n_padglob = 0;
  int64_t m_glob = 0;
  int64_t nslots = 0, nnz = 0;
  int tpr = 32, npt = 2;
  uint32_t padcol = 0;
  double* vals = nullptr;
  uint32_t* cols = nullptr;
  int64_t* rowstart = nullptr;
  double *val = nullptr, *var = nullptr, *mu = nullptr, *tau = nullptr;  // [n_padglob][K]
  double* km = nullptr;  // [K][ld][2]
  double* kmB = nullptr;  // [ceil(K/2)][ld][2]: block-major copy of the values, read by the blocked cluster sweep of the other side
  double* kmBq = nullptr; // same layout, Var + E^2 (VB)
  int64_t ld = 0;
  double* lam_rm = nullptr;
  double* rowstats = nullptr;  // [n_padglob][NRS]
  double* trace[2] = {nullptr, nullptr};
  size_t trace_cap[2] = {0, 0};
  bool ipc_arrays = false;     // val / var / rowstats were re-allocated outside the pool for the peer exchange
  int occ[4] = {0, 0, 0, 0};  // cached resident blocks/SM of the sweep kernel per method
  // dense block kept between create() and pack()
  const double* dR = nullptr;
  const uint8_t* dM = nullptr;
  bool own_dense = false;
  int* rowlen_dev = nullptr;
  int64_t maxlen_local = 0, maxseg_local = 0;
  bool packed = false;
  // v2 (cluster) layout
  bool v2 = false;
  int cs = 8, tw = 0, rs = 0, npt2 = 0, v2div = 1;
  int64_t nsets = 0, tslots = 0;
  double* tvals = nullptr;
  uint16_t* tcols = nullptr;
  int nclusters[4] = {0, 0, 0, 0};
  bool v4 = false;                       // blocked cluster sweep (kernels_v4.cuh) serves this side's Gibbs / ICM / VB sweeps
  bool v5 = false;                       // per-row-chain cluster sweep (kernels_v5.cuh) serves them
  int nclusters4[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};   // resident clusters per method, without / with the parameter outputs
  int nclusters5p[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};  // the same for the kernel with the projection epilogue
  bool long_rows = false;      // rows beyond the register-resident configurations: residual in global scratch (kernels_long.cuh)
  double* escr = nullptr;      // [nslots]
  int occ_long[4] = {0, 0, 0, 0};
  // factor geometry: kf columns; sweeps on this side gather from ysrc ([kf][ysrc_ld][NV])
  int kf = 0;
  const double* ysrc = nullptr;
  int64_t ysrc_ld = 0;
  double* ycomb = nullptr;   // NMTF: W = G S^T (F side) / Z = F S (G side), owned
  double* ycombB = nullptr;  // the same in the block-major layout [ceil(kf/2)][ysrc_ld][2] the per-row-chain cluster sweep reads
  double* ycombBq = nullptr; // VB: the second moments of W / Z, block-major
  double* lamk = nullptr;    // ARD prior of this side's columns (not owned for NMF)
};
constexpr int V2_CS = 8;
}  // namespace

struct bnmtf_nmf_s {
  int device = 0;
  int64_t I = 0, J = 0;
  int K = 0;
  Side s[2];
  Hyper hyper{};
  double rmean = 0.0;
  double* scal = nullptr;      // [SC_COUNT]
  double* colsum = nullptr;    // [2K]
  double* lamstate = nullptr;  // [4K]
  double* lamk = nullptr;      // [K]
  double* stats_dev = nullptr; // [iters chunk][16]
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  int num_sms = 148;
  int rank = 0, world = 1;
  NcclComm comm = nullptr;
  int64_t launches = 0;
  bool time_sweeps = false;
  double sweep_ms[2] = {0, 0};
  int64_t sweep_n[2] = {0, 0};
  int km_mode = -1;            // NV layout currently held by the km copies
  // launch-bound runs: the row-team sweeps leave the gather copy / trace slot of the factor they update themselves
  double* fold_trace[2] = {nullptr, nullptr};   // graph replays with traces: staging buffers the sweeps write their slot of
  double* fold_lam_trace = nullptr;             // ... and the lambda kernel its slot of all_lambdak
  double* gmat = nullptr;              // random-material scratch of the blocked cluster sweep (kernels_v4.cuh)
  int* set_counter = nullptr;          // kernels_v5.cuh: work queue of row sets
  // multi-GPU peer exchange (bnmtf_nmf_peer_export / _import): the per-row-chain sweep writes its rows into the other
  // ranks' arrays; a device-side barrier replaces the all-gathers
  static constexpr int MAXPEER = 16;
  bool peer_ok = false;
  int* peer_flags = nullptr;           // [MAXPEER] this rank's flag array (one slot per source rank)
  void* peer_open[MAXPEER][7] = {};    // opened IPC pointers of the other ranks: val/var/rowstats of both sides, flags
  void** peer_tab_dev = nullptr;       // device copy: [7][MAXPEER] (peers compacted: own rank left out) + flags by rank
  int peer_epoch = 0;
  // scratch of run() that stays with the handle between calls (grow-only): a cross-validation fold of 200 iterations no
  // longer pays cudaMalloc / cudaFree of the trace chunks, the statistics and iterations + 1 cudaEventCreate per call
  size_t stats_cap = 0, lam_trace_cap = 0, lam_trace2_cap = 0;
  double* lam_trace = nullptr;         // [iterations][K]
  double* lam_trace2 = nullptr;        // tri-factorisation: [iterations][L]
  double* trS[2] = {nullptr, nullptr}; // tri-factorisation: trace chunks of S
  size_t trS_cap[2] = {0, 0};
  std::vector<cudaEvent_t> ev_pool;    // per-iteration timing events
  cudaEvent_t chunk_ev[4] = {nullptr, nullptr, nullptr, nullptr};   // trace chunks: done[2], copied[2]
  size_t gmat_doubles = 0;
  uint32_t* iter_dev = nullptr;        // device-side iteration index (graph replays of the run loop)
  const uint32_t* iter_ptr = nullptr;  // == iter_dev while a graph-mode run is in progress
  // ---- tri-factorisation (R ~ F S G^T): side 0 = F (I x K), side 1 = G (J x L)
  bool nmtf = false;
  int L = 0;
  double *S = nullptr, *Svar = nullptr, *Smu = nullptr, *Stau = nullptr, *lamS = nullptr;   // [K][L]
  double* lamG = nullptr;      // [L] (lamk holds lambdaFk)
  double *Prow = nullptr, *Hrow = nullptr;   // [I_pad][L], [I_pad][L(L+1)/2]
  double* np_delta = nullptr;                // NMTF NP: the move of the S entry updated last (device scalar)
  double *cvec = nullptr, *TT = nullptr;     // [K*L], [K(K+1)/2][L(L+1)/2]
  double* Tpart = nullptr;                   // [RT_RC][max tensor]: partial tiles of nmtf_reduce_T
  double* sumlog_lamS_dev = nullptr;
  double sumlog_lamS = 0.0;
  // BNMTF VB
  double* proj_out3 = nullptr;           // third output of the project modes (transient)
  bool fuse_proj = false;                // the next cluster F sweep also projects its final residual onto G (P_il, transient)
  int cov_side = -1;                     // side whose VB sweep carries the covariance term (transient)
  const double* covA = nullptr; const double* covS = nullptr;
  int covL = 0, cov_sk = 0, cov_sl = 0;
  double *Arow = nullptr, *Brow = nullptr, *Crow = nullptr;   // [I_pad][L], [I_pad][L], [J_pad][K]
  double *BS = nullptr, *PA = nullptr, *QC = nullptr;         // [K*L], [K(K+1)/2][L], [L(L+1)/2][K]
  double* lamstateG = nullptr;           // [4][L]  (lamstate holds the F block [4][K])
  double* vbmisc = nullptr;              // [8]: S q-term, sum lambdaS*ES, T3
  // ---- pinned host staging of the draw traces (TraceDrain), kept between runs
  char* pin[2] = {nullptr, nullptr};
  size_t pin_bytes = 0;
  // ---- on-device posterior summaries (engine_summary.inc)
  int sum_burn = 0, sum_thin = 0;        // thinning <= 0: off
  int64_t sum_count = 0;
  double* sum_fac[2] = {nullptr, nullptr};
  double *sum_S = nullptr, *sum_tau = nullptr;
  double* sum_lam[2] = {nullptr, nullptr};
};
static int summary_reset(bnmtf_nmf_s* h);
static int summary_accumulate(bnmtf_nmf_s* h, int it, const double* stats_it);
static void summary_free(bnmtf_nmf_s* h);

// ---- packing -------------------------------------------------------------------
static void choose_team(int64_t maxlen, int* tpr, int* npt) {
  // entries per thread (register-resident residual) and threads per row team;
  // 16 x 512 and 32 x 512 are the largest register-resident configurations
  int n = maxlen > 256 ? 16 : maxlen > 128 ? 8 : maxlen > 64 ? 4 : 2;
  int64_t t = roundup((maxlen + n - 1) / n, 32);
  if (t < 32) t = 32;
  if (t > 512) { n = 32; t = roundup((maxlen + n - 1) / n, 32); }
  *tpr = (int)t;
  *npt = n;
}

static int v2_tile_width(int64_t m) { return (int)roundup((m + V2_CS - 1) / V2_CS, 32); }

// phase 1: observed entries per row and the longest (row, tile) segment of a dense block
static int count_block(bnmtf_nmf_s* h, Side& sd, int64_t n, int64_t m, int64_t ld) {
  CK(cudaMalloc(&sd.rowlen_dev, sizeof(int) * std::max<int64_t>(n + 1, 2)));
  CK(cudaMemsetAsync(sd.rowlen_dev, 0, sizeof(int) * std::max<int64_t>(n + 1, 2), h->stream));
  sd.tw = v2_tile_width(m);
  if (n > 0) {
    count_rows_tiles_kernel<<<(unsigned)((n + 7) / 8), 256, 0, h->stream>>>(sd.dM, n, m, ld, sd.tw, sd.rowlen_dev, sd.rowlen_dev + n);
    h->launches++;
  }
  std::vector<int> rowlen(n + 1);
  CK(cudaMemcpyAsync(rowlen.data(), sd.rowlen_dev, sizeof(int) * (n + 1), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  int64_t maxlen = 1, nnz = 0;
  for (int64_t i = 0; i < n; ++i) { maxlen = std::max<int64_t>(maxlen, rowlen[i]); nnz += rowlen[i]; }
  sd.maxlen_local = maxlen; sd.maxseg_local = std::max(1, rowlen[n]); sd.nnz = nnz;
  return 0;
}

// phase 2: pack the dense block (n x m, row-major, leading dim ld) into padded rows (v1
// layout) and, for long rows, into cluster tiles (v2 layout).  maxlen / maxseg are the
// GLOBAL maxima over all ranks so that the layout -- and with it every summation
// order -- does not depend on how the matrix is sharded.
static int pack_block(bnmtf_nmf_s* h, Side& sd, int64_t n, int64_t m, int64_t ld, uint32_t padcol, int64_t maxlen,
                      int64_t maxseg) {
  choose_team(maxlen, &sd.tpr, &sd.npt);
  if ((int64_t)sd.tpr * sd.npt < maxlen || sd.tpr > 512) {   // longer than 32 x 512 entries: streamed rows
    sd.long_rows = true; sd.tpr = LONG_T; sd.npt = 0;
  }
  std::vector<int> rowlen(n);
  CK(cudaMemcpyAsync(rowlen.data(), sd.rowlen_dev, sizeof(int) * n, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  std::vector<int64_t> rowstart(n + 1);
  rowstart[0] = 0;
  for (int64_t i = 0; i < n; ++i) rowstart[i + 1] = rowstart[i] + roundup(rowlen[i], sd.tpr);
  sd.nslots = rowstart[n];
  sd.padcol = padcol;
  CK(cudaMalloc(&sd.rowstart, sizeof(int64_t) * (n + 1)));
  CK(cudaMalloc(&sd.vals, sizeof(double) * std::max<int64_t>(sd.nslots, 1)));
  CK(cudaMalloc(&sd.cols, sizeof(uint32_t) * std::max<int64_t>(sd.nslots, 1)));
  if (sd.long_rows) CK(cudaMalloc(&sd.escr, sizeof(double) * std::max<int64_t>(sd.nslots, 1)));
  CK(cudaMemcpyAsync(sd.rowstart, rowstart.data(), sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, h->stream));
  if (n > 0) {
    pack_rows_kernel<<<(unsigned)((n + 7) / 8), 256, 0, h->stream>>>(sd.dR, sd.dM, n, m, ld, sd.rowstart, padcol, sd.vals, sd.cols);
    h->launches++;
  }
  // v2 eligibility depends on global quantities only
  const char* mode = getenv("BNMTF_SWEEP");
  const bool force_v1 = mode && strcmp(mode, "v1") == 0;
  int npt2 = (int)roundup((maxseg + 31) / 32, 2);
  if (npt2 < 4) npt2 = 4;
  if (npt2 > 4 && npt2 < 8) npt2 = 8;
  if (npt2 > 8 && npt2 < 12) npt2 = 12;
  if (npt2 > 12 && npt2 < 16) npt2 = 16;
  const int wpc_full = npt2 <= 16 ? 32 : (npt2 <= 18 ? 28 : Sweep2Cfg<20>::WPC);   // Sweep2Cfg<NPT>::WPC
  const int div = 2;                                               // CTAs (of different clusters) per SM: measured best (1 and 3: r01 profiles)
  int rs = wpc_full / div - 1;                                         // rows per set (warp 0 is the control warp)
  sd.v2div = div;
  const int cs = V2_CS;
  const size_t smem_vb = sweep2_smem_doubles(sd.tw, 2, rs, sd.kf, cs) * sizeof(double);
  sd.v2 = !force_v1 && m >= 2048 && maxseg >= 64 && npt2 <= 20 && sd.kf >= 3 && smem_vb <= 200 * 1024 && (sd.tw + V2_PAD) * 16 < 65536;
  {
    const char* v4env = getenv("BNMTF_V4");
    const int rs4 = npt2 == 20 ? sweep4_wpc() / 2 - 1 : rs;    // the blocked sweep may trade rows per set for registers (20 entries per lane only)
    sd.v4 = sd.v2 && !h->nmtf && npt2 % 4 == 0 && sweep4_fits(sd.tw, rs4, sd.kf) && (v4env && atoi(v4env) == 1);   // opt-in: measured 8.1 - 9.2 ms per sweep at C3 against 8.5 for kernels_v2.cuh, depending on the build
    if (sd.v4) rs = rs4;
    // per-row-chain cluster sweep (kernels_v5.cuh): the default for the two-factor models when it fits; BNMTF_V5=0 falls back
    const char* v5env = getenv("BNMTF_V5");
    const int npt5 = npt2 == 18 ? 20 : npt2;
    // (tri-factorisation: the Gibbs / ICM sweeps of F and G; the F sweep carries the projection onto the L columns of G)
    const int proj_k2 = (h->nmtf && &sd == &h->s[0]) ? h->L : 0;
    const int cov_l = h->nmtf ? (&sd == &h->s[0] ? h->L : h->K) : 0;      // VB covariance term (bnmtf_vb.py:343,371)
    sd.v5 = sd.v2 && !sd.v4 && !(v5env && atoi(v5env) == 0) && sweep5_fits(npt5, sd.tw, sd.kf, proj_k2, cov_l);
    if (sd.v5) { npt2 = npt5; rs = sweep5_rows_per_set(npt5); }
  }
  if (sd.v2 && n > 0) {
    sd.cs = cs; sd.npt2 = npt2; sd.rs = rs;
    sd.nsets = (n + rs - 1) / rs;
    sd.tslots = sd.nsets * cs * (int64_t)npt2 * rs * 32;
    CK(cudaMalloc(&sd.tvals, sizeof(double) * sd.tslots));
    CK(cudaMalloc(&sd.tcols, sizeof(uint16_t) * sd.tslots));
    fill_pads_kernel<<<1024, 256, 0, h->stream>>>(sd.tvals, sd.tcols, sd.tslots, (uint16_t)sd.tw);
    pack_tiles_kernel<<<(unsigned)((n * cs + 7) / 8), 256, 0, h->stream>>>(sd.dR, sd.dM, n, m, ld, cs, sd.tw, rs, npt2,
                                                                             sd.tvals, sd.tcols);
    h->launches += 2;
  } else {
    sd.v2 = false; sd.v4 = false; sd.v5 = false;
  }
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  CK(cudaFree(sd.rowlen_dev)); sd.rowlen_dev = nullptr;
  if (sd.own_dense) { cudaFree((void*)sd.dR); cudaFree((void*)sd.dM); }
  sd.dR = nullptr; sd.dM = nullptr; sd.own_dense = false;
  sd.packed = true;
  return 0;
}

static int alloc_side_state(bnmtf_nmf_s* h, Side& sd, int K) {
  sd.kf = K;
  const size_t nb = sizeof(double) * sd.n_padglob * K;
  CK(cudaMalloc(&sd.val, nb)); CK(cudaMalloc(&sd.var, nb)); CK(cudaMalloc(&sd.mu, nb)); CK(cudaMalloc(&sd.tau, nb));
  CK(cudaMemsetAsync(sd.val, 0, nb, h->stream)); CK(cudaMemsetAsync(sd.var, 0, nb, h->stream));
  CK(cudaMemsetAsync(sd.mu, 0, nb, h->stream)); CK(cudaMemsetAsync(sd.tau, 0, nb, h->stream));
  sd.ld = roundup(std::max(sd.n_padglob, sd.n_glob + 1) + 16 * 32, 16);  // >= CS * tile width (CS <= 16)
  CK(cudaMalloc(&sd.km, sizeof(double) * 2 * sd.ld * K));
  CK(cudaMemsetAsync(sd.km, 0, sizeof(double) * 2 * sd.ld * K, h->stream));
  const size_t nbB = sizeof(double) * 2 * sd.ld * ((K + 1) / 2);
  CK(cudaMalloc(&sd.kmB, nbB)); CK(cudaMalloc(&sd.kmBq, nbB));
  CK(cudaMemsetAsync(sd.kmB, 0, nbB, h->stream)); CK(cudaMemsetAsync(sd.kmBq, 0, nbB, h->stream));
  CK(cudaMalloc(&sd.rowstats, sizeof(double) * sd.n_padglob * NRS));
  CK(cudaMemsetAsync(sd.rowstats, 0, sizeof(double) * sd.n_padglob * NRS, h->stream));
  return 0;
}

static int nmtf_alloc(bnmtf_nmf_s* h);
// Size limits of the tri-factorisation handle (the reference has none; documented in README.md / INTEGRATION.md and
// checked when the handle is created, never in the middle of run()): the Gram kernel H_i = sum G_j G_j^T keeps an
// L x L block per warp (L <= 32), the single-CTA S sweeps keep up to 8 K*L doubles in shared memory (227 KB opt-in).
constexpr int NMTF_MAX_L = 48;      // Gram kernel on the FP64 tensor core: up to six 8-column blocks
constexpr size_t NMTF_S_SMEM_MAX = 227 * 1024;
constexpr int NMTF_MAX_KL = (int)((NMTF_S_SMEM_MAX / sizeof(double) - 2) / (2 + TN_PRE_DOUBLES));   // 3631
// ---- stream-ordered memory and copies ----------------------------------------------
// Every allocation comes from the device's stream-ordered pool (cudaMallocAsync) and every
// "synchronous" copy runs on the calling handle's stream followed by a stream synchronise:
// no call in this library synchronises the whole device, so handles driven from different
// host threads (the cross-validation folds) overlap on the GPU.
static thread_local cudaStream_t g_stream = cudaStreamPerThread;

template <class T>
static cudaError_t pool_malloc(T** p, size_t bytes) { return cudaMallocAsync((void**)p, bytes, g_stream); }
static cudaError_t pool_free(void* p) { return p ? cudaFreeAsync(p, g_stream) : cudaSuccess; }
static cudaError_t copy_sync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, kind, g_stream);
  return e != cudaSuccess ? e : cudaStreamSynchronize(g_stream);
}
static cudaError_t copy2d_sync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
                               cudaMemcpyKind kind) {
  cudaError_t e = cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, kind, g_stream);
  return e != cudaSuccess ? e : cudaStreamSynchronize(g_stream);
}
// keep up to 8 GB of freed blocks cached in the pool (default: released at every synchronise)
static cudaError_t pool_setup(int device) {
  static bool done[64] = {false};
  static std::mutex mu;                       // handles are created from several host threads (cross-validation folds)
  std::lock_guard<std::mutex> lk(mu);
  if (device < 0 || device >= 64 || done[device]) return cudaSuccess;
  cudaMemPool_t pool;
  cudaError_t e = cudaDeviceGetDefaultMemPool(&pool, device);
  if (e != cudaSuccess) return e;
  uint64_t keep = (uint64_t)8 << 30;
  e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  done[device] = true;
  return e;
}
// Grow the device's stream-ordered pool to `bytes` in one step: many small first-time allocations (the
// handles of a batch of models) otherwise grow it a few megabytes at a time, a driver call each.
int bnmtf_pool_reserve(int device, int64_t bytes) {
  if (bytes <= 0) return 0;
  CK(cudaSetDevice(device));
  CK(pool_setup(device));
  cudaMemPool_t pool;
  CK(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t reserved = 0;
  CK(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved));
  uint64_t used = 0;
  CK(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used));
  const uint64_t idle = reserved > used ? reserved - used : 0;
  if (idle >= (uint64_t)bytes) return 0;
  void* p = nullptr;
  if (cudaMallocAsync(&p, (size_t)bytes - idle, cudaStreamPerThread) != cudaSuccess) { cudaGetLastError(); return 0; }   // best effort
  cudaFreeAsync(p, cudaStreamPerThread);
  CK(cudaStreamSynchronize(cudaStreamPerThread));
  return 0;
}

// (arrays shared with other processes through CUDA IPC cannot live in the pool: plain device allocations)
static cudaError_t ipc_malloc(void** p, size_t bytes) { return cudaMalloc(p, bytes); }
static cudaError_t ipc_free(void* p) { return p ? cudaFree(p) : cudaSuccess; }

#define cudaMalloc pool_malloc
#define cudaFree pool_free
#define cudaMemcpy copy_sync
#define cudaMemcpy2D copy2d_sync
// entry of every call on a handle: its device and its stream become current for this thread
#define ENTER(h)                          \
  do {                                    \
    CK(cudaSetDevice((h)->device));       \
    g_stream = (h)->stream;               \
  } while (0)

unsigned long long* bnmtf_prof_last = nullptr;   // timing experiments only (BNMTF_V2_DBG & 2)

static int create_dev_impl(bnmtf_nmf_t* out, int device, int64_t I, int64_t J, int K, int L, bool nmtf, const double* Rrow_dev,
                           const uint8_t* Mrow_dev, const double* Rcol_dev, const uint8_t* Mcol_dev, int64_t row0,
                           int64_t row1, int64_t shard_rows, int64_t col0, int64_t col1, int64_t shard_cols);

int bnmtf_nmf_create_dev(bnmtf_nmf_t* out, int device, int64_t I, int64_t J, int K, const double* Rrow_dev,
                         const uint8_t* Mrow_dev, const double* Rcol_dev, const uint8_t* Mcol_dev, int64_t row0,
                         int64_t row1, int64_t shard_rows, int64_t col0, int64_t col1, int64_t shard_cols) {
  return create_dev_impl(out, device, I, J, K, K, false, Rrow_dev, Mrow_dev, Rcol_dev, Mcol_dev, row0, row1, shard_rows, col0,
                         col1, shard_cols);
}

static int create_dev_impl(bnmtf_nmf_t* out, int device, int64_t I, int64_t J, int K, int L, bool nmtf, const double* Rrow_dev,
                           const uint8_t* Mrow_dev, const double* Rcol_dev, const uint8_t* Mcol_dev, int64_t row0,
                           int64_t row1, int64_t shard_rows, int64_t col0, int64_t col1, int64_t shard_cols) {
  if (!out) FAIL("out is NULL");
  *out = nullptr;
  if (I <= 0 || J <= 0 || K <= 0 || L <= 0) FAIL("bad shape I=%lld J=%lld K=%d L=%d", (long long)I, (long long)J, K, L);
  if (nmtf && L > NMTF_MAX_L) FAIL("tri-factorisation: L = %d is not supported (L <= %d; K is only bound by K*L <= %d)", L, NMTF_MAX_L, NMTF_MAX_KL);
  if (nmtf && (int64_t)K * L > NMTF_MAX_KL) FAIL("tri-factorisation: K*L = %lld is not supported (K*L <= %d)", (long long)K * L, NMTF_MAX_KL);
  if (I >= (int64_t)0xFFFFFFF0ll || J >= (int64_t)0xFFFFFFF0ll) FAIL("dimension too large for 32-bit column indices");
  if (row0 < 0 || row1 < row0 || row1 > I || col0 < 0 || col1 < col0 || col1 > J) FAIL("bad shard bounds");
  if (shard_rows < row1 - row0 || shard_cols < col1 - col0) FAIL("shard size smaller than the owned range");
  CK(cudaSetDevice(device));
  CK(pool_setup(device));
  int cc_major = 0, cc_minor = 0, sms = 0;
  CK(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device));
  CK(cudaDeviceGetAttribute(&cc_minor, cudaDevAttrComputeCapabilityMinor, device));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  if (cc_major < 10) FAIL("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, cc_major, cc_minor);
  bnmtf_nmf_s* h = new bnmtf_nmf_s();
  h->device = device; h->I = I; h->J = J; h->K = K; h->L = L; h->nmtf = nmtf;
  h->num_sms = sms;
  {  // handle streams run at the highest priority; bnmtf_batch_run drives its long launches at the lowest, so the
     // short kernels of a model being set up are never queued behind a batch that occupies the GPU
    int least = 0, greatest = 0;
    CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    CK(cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, greatest));
    CK(cudaStreamCreateWithPriority(&h->copy_stream, cudaStreamNonBlocking, greatest));
  }
  g_stream = h->stream;
  const char* ts = getenv("BNMTF_TIME_SWEEPS");
  h->time_sweeps = ts && atoi(ts) != 0;

  Side& su = h->s[BNMTF_SIDE_U];
  Side& sv = h->s[BNMTF_SIDE_V];
  su.kf = K; sv.kf = L;
  su.n_glob = I; su.m_glob = J; su.row0 = row0; su.n_loc = row1 - row0; su.n_shard = shard_rows;
  sv.n_glob = J; sv.m_glob = I; sv.row0 = col0; sv.n_loc = col1 - col0; sv.n_shard = shard_cols;
  su.n_padglob = std::max(I, ((I + shard_rows - 1) / shard_rows) * shard_rows);
  sv.n_padglob = std::max(J, ((J + shard_cols - 1) / shard_cols) * shard_cols);

  // U side: rows [row0,row1) x J (the caller's block, read again by bnmtf_nmf_pack)
  su.dR = Rrow_dev; su.dM = Mrow_dev; su.own_dense = false;
  if (count_block(h, su, su.n_loc, J, J)) return 1;
  // V side: columns [col0,col1): transposed copy of the I x Jloc block, owned by the handle
  {
    const int64_t jl = sv.n_loc;
    double* Rt = nullptr; uint8_t* Mt = nullptr;
    CK(cudaMalloc(&Rt, sizeof(double) * std::max<int64_t>(jl * I, 1)));
    CK(cudaMalloc(&Mt, std::max<int64_t>(jl * I, 1)));
    if (jl > 0) {
      dim3 blk(32, 8), grd((unsigned)((jl + 31) / 32), (unsigned)((I + 31) / 32));
      transpose_kernel<double><<<grd, blk, 0, h->stream>>>(Rcol_dev, I, jl, jl, Rt, I);
      transpose_kernel<uint8_t><<<grd, blk, 0, h->stream>>>(Mcol_dev, I, jl, jl, Mt, I);
      h->launches += 2;
    }
    sv.dR = Rt; sv.dM = Mt; sv.own_dense = true;
    if (count_block(h, sv, jl, I, I)) return 1;
  }
  su.kf = K; sv.kf = L;
  if (alloc_side_state(h, su, K) || alloc_side_state(h, sv, L)) return 1;
  su.lamk = nullptr; sv.lamk = nullptr;
  CK(cudaMalloc(&h->scal, sizeof(double) * SC_COUNT));
  CK(cudaMalloc(&h->colsum, sizeof(double) * 2 * (K + L)));
  CK(cudaMalloc(&h->lamstate, sizeof(double) * 4 * K));
  CK(cudaMalloc(&h->lamk, sizeof(double) * K));
  CK(cudaMemsetAsync(h->scal, 0, sizeof(double) * SC_COUNT, h->stream));
  CK(cudaMemsetAsync(h->colsum, 0, sizeof(double) * 2 * (K + L), h->stream));
  CK(cudaMemsetAsync(h->lamstate, 0, sizeof(double) * 4 * K, h->stream));
  CK(cudaMemsetAsync(h->lamk, 0, sizeof(double) * K, h->stream));
  if (!nmtf) {
    su.lamk = h->lamk; sv.lamk = h->lamk;
    su.ysrc = sv.km; su.ysrc_ld = sv.ld; sv.ysrc = su.km; sv.ysrc_ld = su.ld;
  } else if (nmtf_alloc(h)) {
    return 1;
  }
  h->hyper.ARD = 1; h->hyper.alphatau = h->hyper.betatau = h->hyper.alpha0 = h->hyper.beta0 = 1.0;
  h->hyper.size_Omega = (double)su.nnz;
  CK(cudaStreamSynchronize(h->stream));
  *out = h;
  return 0;
}

int bnmtf_nmf_layout_stats(bnmtf_nmf_t h, int64_t* out4) {
  if (!h || !out4) FAIL("bad argument");
  out4[0] = h->s[0].maxlen_local; out4[1] = h->s[0].maxseg_local;
  out4[2] = h->s[1].maxlen_local; out4[3] = h->s[1].maxseg_local;
  return 0;
}

int bnmtf_nmf_pack(bnmtf_nmf_t h, const int64_t* global4) {
  if (!h) FAIL("NULL handle");
  ENTER(h);
  Side& su = h->s[0]; Side& sv = h->s[1];
  if (su.packed) FAIL("already packed");
  int64_t g[4] = {su.maxlen_local, su.maxseg_local, sv.maxlen_local, sv.maxseg_local};
  if (global4) for (int q = 0; q < 4; ++q) g[q] = std::max(g[q], global4[q]);
  if (pack_block(h, su, su.n_loc, h->J, h->J, (uint32_t)h->J, g[0], g[1])) return 1;
  if (pack_block(h, sv, sv.n_loc, h->I, h->I, (uint32_t)h->I, g[2], g[3])) return 1;
  // Models whose draws are megabytes per iteration deliver their traces through two pinned staging slots (TraceDrain).
  // cudaHostAlloc takes milliseconds per slot and stalls the launches of the run that first needs them (measured: 8 ms
  // of idle GPU in a 20-iteration run at 20000 x 20000, K = 50), so such handles get their slots here, at set-up time.
  {
    const size_t per_it = sizeof(double) * ((size_t)su.n_glob * su.kf + (size_t)sv.n_glob * sv.kf + (h->nmtf ? (size_t)h->K * h->L : 0));
    if (per_it >= ((size_t)2 << 20) && !getenv("BNMTF_NO_PINNED_TRACES") && !getenv("BNMTF_TRACE_CHUNK_BYTES")) {
      const size_t slot = std::max<size_t>((size_t)16 << 20, per_it);
      if (h->pin_bytes < slot) {
        for (int b = 0; b < 2; ++b) { if (h->pin[b]) cudaFreeHost(h->pin[b]); h->pin[b] = nullptr; }
        h->pin_bytes = 0;
        cudaError_t e = cudaSuccess;
        for (int b = 0; b < 2 && e == cudaSuccess; ++b) e = cudaHostAlloc((void**)&h->pin[b], slot, cudaHostAllocDefault);
        if (e == cudaSuccess) h->pin_bytes = slot;
        else { for (int b = 0; b < 2; ++b) { if (h->pin[b]) cudaFreeHost(h->pin[b]); h->pin[b] = nullptr; } cudaGetLastError(); }   // the run allocates (and reports) later
      }
    }
  }
  return 0;
}

static int create_host_impl(bnmtf_nmf_t* out, int device, int64_t I, int64_t J, int K, int L, bool nmtf, const double* R,
                            const uint8_t* M, int64_t row0, int64_t row1, int64_t shard_rows, int64_t col0, int64_t col1,
                            int64_t shard_cols);
int bnmtf_nmf_create(bnmtf_nmf_t* out, int device, int64_t I, int64_t J, int K, const double* R, const uint8_t* M,
                     int64_t row0, int64_t row1, int64_t shard_rows, int64_t col0, int64_t col1, int64_t shard_cols) {
  return create_host_impl(out, device, I, J, K, K, false, R, M, row0, row1, shard_rows, col0, col1, shard_cols);
}
static int create_host_impl(bnmtf_nmf_t* out, int device, int64_t I, int64_t J, int K, int L, bool nmtf, const double* R,
                            const uint8_t* M, int64_t row0, int64_t row1, int64_t shard_rows, int64_t col0, int64_t col1,
                            int64_t shard_cols) {
  if (!R || !M) FAIL("R or M is NULL");
  if (I <= 0 || J <= 0) FAIL("bad shape");
  CK(cudaSetDevice(device));
  CK(pool_setup(device));
  g_stream = cudaStreamPerThread;
  const int64_t nr = row1 - row0, nc = col1 - col0;
  const bool full = (nr == I && nc == J);
  double *Rrow = nullptr, *Rcol = nullptr;
  uint8_t *Mrow = nullptr, *Mcol = nullptr;
  CK(cudaMalloc(&Rrow, sizeof(double) * std::max<int64_t>(nr * J, 1)));
  CK(cudaMalloc(&Mrow, std::max<int64_t>(nr * J, 1)));
  if (nr > 0) {
    CK(cudaMemcpy(Rrow, R + row0 * J, sizeof(double) * nr * J, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(Mrow, M + row0 * J, (size_t)nr * J, cudaMemcpyHostToDevice));
  }
  if (full) {
    Rcol = Rrow; Mcol = Mrow;
  } else {
    CK(cudaMalloc(&Rcol, sizeof(double) * std::max<int64_t>(I * nc, 1)));
    CK(cudaMalloc(&Mcol, std::max<int64_t>(I * nc, 1)));
    if (nc > 0) {
      CK(cudaMemcpy2D(Rcol, sizeof(double) * nc, R + col0, sizeof(double) * J, sizeof(double) * nc, I, cudaMemcpyHostToDevice));
      CK(cudaMemcpy2D(Mcol, nc, M + col0, J, nc, I, cudaMemcpyHostToDevice));
    }
  }
  int rc = create_dev_impl(out, device, I, J, K, L, nmtf, Rrow, Mrow, Rcol, Mcol, row0, row1, shard_rows, col0, col1, shard_cols);
  if (!full) { cudaFree(Rcol); cudaFree(Mcol); }
  if (rc) { cudaFree(Rrow); cudaFree(Mrow); return rc; }
  (*out)->s[0].own_dense = true;   // released by bnmtf_nmf_pack
  return 0;
}

static void free_side(Side& sd) {
  cudaFree(sd.vals); cudaFree(sd.cols); cudaFree(sd.rowstart);
  if (sd.ipc_arrays) { ipc_free(sd.val); ipc_free(sd.var); ipc_free(sd.rowstats); }
  else { cudaFree(sd.val); cudaFree(sd.var); cudaFree(sd.rowstats); }
  cudaFree(sd.mu); cudaFree(sd.tau);
  cudaFree(sd.km); cudaFree(sd.kmB); cudaFree(sd.kmBq); cudaFree(sd.lam_rm);
  cudaFree(sd.trace[0]); cudaFree(sd.trace[1]);
  cudaFree(sd.tvals); cudaFree(sd.tcols); cudaFree(sd.rowlen_dev); cudaFree(sd.ycomb); cudaFree(sd.ycombB); cudaFree(sd.ycombBq); cudaFree(sd.escr);
  if (sd.own_dense) { cudaFree((void*)sd.dR); cudaFree((void*)sd.dM); }
}

int bnmtf_nmf_destroy(bnmtf_nmf_t h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  g_stream = h->stream;
  cudaStreamSynchronize(h->copy_stream);
  cudaStreamSynchronize(h->stream);
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  free_side(h->s[0]); free_side(h->s[1]);
  cudaFree(h->scal); cudaFree(h->colsum); cudaFree(h->lamstate); cudaFree(h->lamk); cudaFree(h->stats_dev);
  cudaFree(h->S); cudaFree(h->Svar); cudaFree(h->Smu); cudaFree(h->Stau); cudaFree(h->lamS); cudaFree(h->lamG);
  cudaFree(h->Prow); cudaFree(h->Hrow); cudaFree(h->np_delta); cudaFree(h->cvec); cudaFree(h->TT); cudaFree(h->Tpart);
  cudaFree(h->Arow); cudaFree(h->Brow); cudaFree(h->Crow); cudaFree(h->BS); cudaFree(h->PA); cudaFree(h->QC);
  cudaFree(h->lamstateG); cudaFree(h->vbmisc); cudaFree(h->iter_dev); cudaFree(h->gmat); cudaFree(h->set_counter); cudaFree(h->lam_trace); cudaFree(h->lam_trace2); cudaFree(h->trS[0]); cudaFree(h->trS[1]);
  for (auto& e : h->ev_pool) cudaEventDestroy(e);
  for (int q = 0; q < 4; ++q) if (h->chunk_ev[q]) cudaEventDestroy(h->chunk_ev[q]);
  for (int r = 0; r < bnmtf_nmf_s::MAXPEER; ++r) for (int q = 0; q < 7; ++q) if (h->peer_open[r][q]) cudaIpcCloseMemHandle(h->peer_open[r][q]);
  ipc_free(h->peer_flags); cudaFree(h->peer_tab_dev);
  summary_free(h);
  for (int b = 0; b < 2; ++b) if (h->pin[b]) cudaFreeHost(h->pin[b]);
  cudaStreamSynchronize(h->stream);
  g_stream = cudaStreamPerThread;
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  delete h;
  return 0;
}

int bnmtf_nmf_data_stats(bnmtf_nmf_t h, double* out) {
  if (!h || !out) FAIL("NULL argument");
  if (!h->s[0].packed) FAIL("call bnmtf_nmf_pack first");
  ENTER(h);
  Side& su = h->s[0];
  const int nb = 256;
  double* part = nullptr;
  CK(cudaMalloc(&part, sizeof(double) * nb * 3));
  data_stats_kernel<<<nb, 256, 0, h->stream>>>(su.vals, su.cols, su.nslots, su.padcol, h->rmean, part);
  h->launches++;
  std::vector<double> hp(nb * 3);
  CK(cudaMemcpyAsync(hp.data(), part, sizeof(double) * nb * 3, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaFree(part));
  out[0] = out[1] = out[2] = 0.0;
  for (int b = 0; b < nb; ++b) { out[0] += hp[b * 3]; out[1] += hp[b * 3 + 1]; out[2] += hp[b * 3 + 2]; }
  return 0;
}

int bnmtf_nmf_set_data_stats(bnmtf_nmf_t h, double size_Omega, double mean_R) {
  if (!h) FAIL("NULL handle");
  h->hyper.size_Omega = size_Omega;
  h->rmean = mean_R;
  return 0;
}

int bnmtf_nccl_unique_id(uint8_t* id128) {
  if (!id128) FAIL("NULL argument");
  if (!g_nccl.load()) FAIL("cannot load libnccl.so.2: %s", dlerror());
  NcclUniqueId id;
  int rc = g_nccl.GetUniqueId(&id);
  if (rc != 0) FAIL("ncclGetUniqueId: %s", g_nccl.GetErrorString(rc));
  memcpy(id128, id.internal, 128);
  return 0;
}

int bnmtf_nmf_set_comm(bnmtf_nmf_t h, int rank, int world, const uint8_t* nccl_id128) {
  if (!h) FAIL("NULL handle");
  if (world <= 1) { h->rank = 0; h->world = 1; return 0; }
  if (!nccl_id128) FAIL("NULL nccl id");
  if (!g_nccl.load()) FAIL("cannot load libnccl.so.2: %s", dlerror());
  ENTER(h);
  for (int sidx = 0; sidx < 2; ++sidx) {
    const Side& sd = h->s[sidx];
    if (sd.n_shard * world != sd.n_padglob) FAIL("shard size %lld x world %d != padded extent %lld", (long long)sd.n_shard, world, (long long)sd.n_padglob);
    if (sd.row0 != sd.n_shard * rank) FAIL("rank %d does not own the shard starting at %lld", rank, (long long)sd.row0);
  }
  NcclUniqueId id;
  memcpy(id.internal, nccl_id128, 128);
  int rc = g_nccl.CommInitRank(&h->comm, world, id, rank);
  if (rc != 0) FAIL("ncclCommInitRank: %s", g_nccl.GetErrorString(rc));
  h->rank = rank; h->world = world;
  return 0;
}

// ---- peer exchange: CUDA IPC handles of the arrays the sweeps of the other ranks write into
static constexpr int PEER_NBUF = 7;     // U: val, var, rowstats; V: val, var, rowstats; flags
int bnmtf_nmf_peer_export(bnmtf_nmf_t h, uint8_t* blob, int64_t blob_bytes) {
  if (!h || !blob) FAIL("bad argument");
  if (blob_bytes < (int64_t)(PEER_NBUF * sizeof(cudaIpcMemHandle_t))) FAIL("peer blob too small");
  if (h->nmtf) FAIL("the peer exchange serves the two-factor models");
  ENTER(h);
  if (!h->peer_flags) {
    CK(ipc_malloc((void**)&h->peer_flags, sizeof(int) * bnmtf_nmf_s::MAXPEER));
    CK(cudaMemsetAsync(h->peer_flags, 0, sizeof(int) * bnmtf_nmf_s::MAXPEER, h->stream));
  }
  for (int sidx = 0; sidx < 2; ++sidx) {   // move the shared arrays out of the stream-ordered pool
    Side& sd = h->s[sidx];
    if (sd.ipc_arrays) continue;
    if (!sd.val || !sd.var || !sd.rowstats) FAIL("peer export before the factors were allocated");
    double** arr[3] = {&sd.val, &sd.var, &sd.rowstats};
    const size_t bytes[3] = {sizeof(double) * sd.n_padglob * sd.kf, sizeof(double) * sd.n_padglob * sd.kf, sizeof(double) * sd.n_padglob * NRS};
    for (int q = 0; q < 3; ++q) {
      void* fresh = nullptr;
      CK(ipc_malloc(&fresh, bytes[q]));
      CK(cudaMemcpyAsync(fresh, *arr[q], bytes[q], cudaMemcpyDeviceToDevice, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      cudaFree(*arr[q]);
      *arr[q] = (double*)fresh;
    }
    sd.ipc_arrays = true;
  }
  CK(cudaStreamSynchronize(h->stream));
  void* bufs[PEER_NBUF] = {h->s[0].val, h->s[0].var, h->s[0].rowstats, h->s[1].val, h->s[1].var, h->s[1].rowstats, h->peer_flags};
  cudaIpcMemHandle_t* out = reinterpret_cast<cudaIpcMemHandle_t*>(blob);
  for (int q = 0; q < PEER_NBUF; ++q) {
    if (!bufs[q]) FAIL("peer export before the factors were allocated");
    CK(cudaIpcGetMemHandle(&out[q], bufs[q]));
  }
  return 0;
}

int bnmtf_nmf_peer_import(bnmtf_nmf_t h, const uint8_t* blobs, int64_t blob_bytes) {
  if (!h || !blobs) FAIL("bad argument");
  if (h->world <= 1) return 0;
  if (h->world > bnmtf_nmf_s::MAXPEER) FAIL("at most %d ranks", bnmtf_nmf_s::MAXPEER);
  if (blob_bytes < (int64_t)(PEER_NBUF * sizeof(cudaIpcMemHandle_t))) FAIL("peer blob too small");
  ENTER(h);
  // table on the device: rows 0..5 = val/var/rowstats of U, V with the peers compacted (own rank left out); row 6 = the
  // flag arrays indexed by rank (own slot unused)
  std::vector<void*> tab((size_t)PEER_NBUF * bnmtf_nmf_s::MAXPEER, nullptr);
  int np = 0;
  for (int r = 0; r < h->world; ++r) {
    if (r == h->rank) continue;
    const cudaIpcMemHandle_t* hd = reinterpret_cast<const cudaIpcMemHandle_t*>(blobs + (size_t)r * blob_bytes);
    for (int q = 0; q < PEER_NBUF; ++q) {
      void* ptr = nullptr;
      CK(cudaIpcOpenMemHandle(&ptr, hd[q], cudaIpcMemLazyEnablePeerAccess));
      h->peer_open[r][q] = ptr;
      if (q < 6) tab[(size_t)q * bnmtf_nmf_s::MAXPEER + np] = ptr;
      else tab[(size_t)q * bnmtf_nmf_s::MAXPEER + r] = ptr;
    }
    ++np;
  }
  if (!h->peer_tab_dev) CK(cudaMalloc(&h->peer_tab_dev, sizeof(void*) * tab.size()));
  CK(cudaMemcpy(h->peer_tab_dev, tab.data(), sizeof(void*) * tab.size(), cudaMemcpyHostToDevice));
  h->peer_ok = true;
  return 0;
}

int bnmtf_nmf_set_hyper(bnmtf_nmf_t h, int ARD, double alphatau, double betatau, double alpha0, double beta0,
                        const double* lambdaU, const double* lambdaV) {
  if (!h) FAIL("NULL handle");
  ENTER(h);
  h->hyper.ARD = ARD; h->hyper.alphatau = alphatau; h->hyper.betatau = betatau;
  h->hyper.alpha0 = alpha0; h->hyper.beta0 = beta0;
  h->hyper.sumlog_lamU = h->hyper.sumlog_lamV = 0.0;
  const double* src[2] = {lambdaU, lambdaV};
  for (int sidx = 0; sidx < 2; ++sidx) {
    Side& sd = h->s[sidx];
    if (!ARD) {
      if (!src[sidx]) FAIL("lambda%c is required when ARD is off", sidx ? 'V' : 'U');
      if (!sd.lam_rm) CK(cudaMalloc(&sd.lam_rm, sizeof(double) * sd.n_glob * sd.kf));
      CK(cudaMemcpy(sd.lam_rm, src[sidx], sizeof(double) * sd.n_glob * sd.kf, cudaMemcpyHostToDevice));
      double sl = 0.0;
      for (int64_t q = 0; q < sd.n_glob * sd.kf; ++q) sl += log(src[sidx][q]);
      (sidx ? h->hyper.sumlog_lamV : h->hyper.sumlog_lamU) = sl;
    }
  }
  return 0;
}

static double* field_ptr(Side& sd, int field) {
  switch (field) {
    case BNMTF_F_VALUE: return sd.val;
    case BNMTF_F_VAR: return sd.var;
    case BNMTF_F_MU: return sd.mu;
    case BNMTF_F_TAU: return sd.tau;
  }
  return nullptr;
}

int bnmtf_nmf_set_factor(bnmtf_nmf_t h, int side, int field, const double* src) {
  if (!h || !src || side < 0 || side > 1) FAIL("bad argument");
  double* p = field_ptr(h->s[side], field);
  if (!p) FAIL("bad field %d", field);
  ENTER(h);
  CK(cudaMemcpy(p, src, sizeof(double) * h->s[side].n_glob * h->s[side].kf, cudaMemcpyHostToDevice));
  h->km_mode = -1;
  return 0;
}

int bnmtf_nmf_get_factor(bnmtf_nmf_t h, int side, int field, double* dst) {
  if (!h || !dst || side < 0 || side > 1) FAIL("bad argument");
  double* p = field_ptr(h->s[side], field);
  if (!p) FAIL("bad field %d", field);
  ENTER(h);
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaMemcpy(dst, p, sizeof(double) * h->s[side].n_glob * h->s[side].kf, cudaMemcpyDeviceToHost));
  return 0;
}

int bnmtf_nmf_set_scalars(bnmtf_nmf_t h, double tau, const double* lambdak) {
  if (!h) FAIL("NULL handle");
  ENTER(h);
  CK(cudaMemcpy(h->scal + SC_TAU, &tau, sizeof(double), cudaMemcpyHostToDevice));
  if (lambdak) {
    CK(cudaMemcpy(h->lamk, lambdak, sizeof(double) * h->K, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->lamstate + 2 * h->K, lambdak, sizeof(double) * h->K, cudaMemcpyHostToDevice));
  }
  return 0;
}

int bnmtf_nmf_get_scalars(bnmtf_nmf_t h, double* tau, double* lambdak) {
  if (!h) FAIL("NULL handle");
  ENTER(h);
  CK(cudaStreamSynchronize(h->stream));
  if (tau) CK(cudaMemcpy(tau, h->scal + SC_TAU, sizeof(double), cudaMemcpyDeviceToHost));
  if (lambdak) CK(cudaMemcpy(lambdak, h->lamk, sizeof(double) * h->K, cudaMemcpyDeviceToHost));
  return 0;
}

int bnmtf_nmf_set_vb_scalars(bnmtf_nmf_t h, double exp_tau, double exp_logtau, double alpha_s, double beta_s,
                             const double* alphak_s, const double* betak_s, const double* exp_lambdak,
                             const double* exp_loglambdak) {
  if (!h) FAIL("NULL handle");
  ENTER(h);
  double sc[SC_COUNT] = {0};
  sc[SC_TAU] = exp_tau; sc[SC_LOGTAU] = exp_logtau; sc[SC_ALPHA_S] = alpha_s; sc[SC_BETA_S] = beta_s;
  CK(cudaMemcpy(h->scal, sc, sizeof(sc), cudaMemcpyHostToDevice));
  const double* src[4] = {alphak_s, betak_s, exp_lambdak, exp_loglambdak};
  for (int f = 0; f < 4; ++f)
    if (src[f]) CK(cudaMemcpy(h->lamstate + f * h->K, src[f], sizeof(double) * h->K, cudaMemcpyHostToDevice));
  if (exp_lambdak) CK(cudaMemcpy(h->lamk, exp_lambdak, sizeof(double) * h->K, cudaMemcpyHostToDevice));
  return 0;
}

int bnmtf_nmf_get_vb_scalars(bnmtf_nmf_t h, double* scal4, double* lamstate4K) {
  if (!h) FAIL("NULL handle");
  ENTER(h);
  CK(cudaStreamSynchronize(h->stream));
  if (scal4) CK(cudaMemcpy(scal4, h->scal, sizeof(double) * 4, cudaMemcpyDeviceToHost));
  if (lamstate4K) CK(cudaMemcpy(lamstate4K, h->lamstate, sizeof(double) * 4 * h->K, cudaMemcpyDeviceToHost));
  return 0;
}

// ---- launches --------------------------------------------------------------------
static int refresh_km(bnmtf_nmf_s* h, int side, int method) {
  Side& sd = h->s[side];
  // the blocked cluster sweep of the other side reads the block-major copies: both layouts from one launch
  // (tri-factorisation: only the projection epilogue of the F sweep reads a factor directly, G)
  const bool blockmajor = (h->s[1 - side].v4 || h->s[1 - side].v5) && (h->nmtf ? side == 1 : true) && method != BNMTF_NP;
  if (blockmajor) {
    dim3 g3((unsigned)((sd.ld + 255) / 256), (unsigned)((sd.kf + 1) / 2));
    if (method == BNMTF_VB) rm_to_km_both_kernel<true><<<g3, 256, 0, h->stream>>>(sd.val, sd.var, sd.km, sd.kmB, sd.kmBq, sd.n_glob, sd.kf, sd.ld);
    else rm_to_km_both_kernel<false><<<g3, 256, 0, h->stream>>>(sd.val, sd.var, sd.km, sd.kmB, sd.kmBq, sd.n_glob, sd.kf, sd.ld);
    h->launches++;
    return 0;
  }
  dim3 blk(32, 8), grd((unsigned)((sd.ld + 31) / 32), (unsigned)((sd.kf + 31) / 32));
  if (method == BNMTF_VB)
    rm_to_km_kernel<2><<<grd, blk, 0, h->stream>>>(sd.val, sd.var, sd.km, sd.n_glob, sd.kf, sd.ld);
  else
    rm_to_km_kernel<1><<<grd, blk, 0, h->stream>>>(sd.val, sd.var, sd.km, sd.n_glob, sd.kf, sd.ld);
  h->launches++;
  return 0;
}

template <int NPT, int MODE>
static int launch_sweep_t(bnmtf_nmf_s* h, const SweepParams& p, int n_loc, int* occ_cache) {
  const int tpr = p.tpr;
  const int teams = tpr >= 128 ? 1 : 128 / tpr;
  const int block = tpr * teams;
  const int nw = tpr / 32;
  const size_t smem = sizeof(double) * teams * (4 * (size_t)p.K + 3 * nw + 2 + 2 * (size_t)p.covL + (MODE == M_GIBBS ? TN_PRE_DOUBLES * (size_t)p.K : 0));
  auto kern = sweep_kernel<NPT, MODE>;
  if (*occ_cache == 0) {
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, block, smem));
    if (occ < 1) FAIL("sweep kernel does not fit: block %d smem %zu", block, smem);
    *occ_cache = occ;
  }
  int64_t need = ((int64_t)n_loc + teams - 1) / teams;
  int grid = (int)std::min<int64_t>(need, (int64_t)(*occ_cache) * h->num_sms);
  if (grid < 1) return 0;
  kern<<<grid, block, smem, h->stream>>>(p);
  h->launches++;
  return 0;
}

template <int MODE>
static int launch_sweep_m(bnmtf_nmf_s* h, const SweepParams& p, int n_loc, int npt, int* occ) {
  switch (npt) {
    case 2: return launch_sweep_t<2, MODE>(h, p, n_loc, occ);
    case 4: return launch_sweep_t<4, MODE>(h, p, n_loc, occ);
    case 8: return launch_sweep_t<8, MODE>(h, p, n_loc, occ);
    case 16: return launch_sweep_t<16, MODE>(h, p, n_loc, occ);
    case 32: return launch_sweep_t<32, MODE>(h, p, n_loc, occ);
  }
  FAIL("unsupported entries-per-thread %d", npt);
}

template <int NPT, int MODE, int DIV, int CS = V2_CS>
static int launch_sweep2_t(bnmtf_nmf_s* h, const Sweep2Params& p_in, Side& sd, int method) {
  constexpr int NV = (MODE == M_VB) ? 2 : 1;
  auto kern = sweep2_kernel<NPT, MODE, CS, DIV>;
  // tile ring: a fourth stage when it does not cost a resident CTA (228 KB per SM, 1 KB reserved per CTA, 227 KB per CTA)
  Sweep2Params p = p_in;
  const size_t smem3 = sweep2_smem_doubles(p.tw, NV, p.rs, p.K, CS, p.K2, MODE == M_GIBBS, 3) * sizeof(double);
  const size_t smem4 = sweep2_smem_doubles(p.tw, NV, p.rs, p.K, CS, p.K2, MODE == M_GIBBS, 4) * sizeof(double);
  const size_t per_sm = 228 * 1024;
  const int ctas3 = (int)std::min<size_t>(DIV, per_sm / (smem3 + 1024)), ctas4 = (int)std::min<size_t>(DIV, per_sm / (smem4 + 1024));
  p.nst = (smem4 <= 227 * 1024 && ctas4 >= ctas3) ? 4 : 3;
  if (getenv("BNMTF_V2_NST")) p.nst = atoi(getenv("BNMTF_V2_NST")) == 4 && smem4 <= 227 * 1024 ? 4 : 3;
  const size_t smem = p.nst == 4 ? smem4 : smem3;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3((p.rs + 1) * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = h->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (sd.nclusters[method] == 0) {
    // one fixed opt-in limit per instantiation: both sides (and other handles) share the kernel
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (CS > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cfg.gridDim = dim3(CS * 64);
    int ncl = 0;
    CK(cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg));
    if (ncl < 1) FAIL("cluster sweep kernel does not fit (smem %zu bytes, %d threads)", smem, (p.rs + 1) * 32);
    if (getenv("BNMTF_V2_NCL")) ncl = std::max(1, atoi(getenv("BNMTF_V2_NCL")));   // experiments: clusters are independent, a larger grid queues
    if (getenv("BNMTF_V2_DBG")) fprintf(stderr, "sweep2<%d,%d,%d>: %d clusters of %d x %d threads, %zu bytes of shared memory, %d tile stages\n", NPT, MODE, DIV, ncl, CS, (p.rs + 1) * 32, smem, p.nst);
    sd.nclusters[method] = ncl;
  }
  const int ncl = (int)std::min<int64_t>(sd.nclusters[method], p.nsets);
  if (ncl < 1) return 0;
  cfg.gridDim = dim3(CS * ncl);
  CK(cudaLaunchKernelEx(&cfg, kern, p));
  h->launches++;
  return 0;
}

template <int MODE, int DIV>
static int launch_sweep2_d(bnmtf_nmf_s* h, const Sweep2Params& p, Side& sd, int method) {
  switch (sd.npt2) {
    case 4: return launch_sweep2_t<4, MODE, DIV>(h, p, sd, method);
    case 8: return launch_sweep2_t<8, MODE, DIV>(h, p, sd, method);
    case 12: return launch_sweep2_t<12, MODE, DIV>(h, p, sd, method);
    case 16: return launch_sweep2_t<16, MODE, DIV>(h, p, sd, method);
    case 18: return launch_sweep2_t<18, MODE, DIV>(h, p, sd, method);
    case 20: return launch_sweep2_t<20, MODE, DIV>(h, p, sd, method);
  }
  FAIL("unsupported v2 entries-per-lane %d", sd.npt2);
}

template <int MODE>
static int launch_sweep2_m(bnmtf_nmf_s* h, const Sweep2Params& p, Side& sd, int method) {
  return launch_sweep2_d<MODE, 2>(h, p, sd, method);
}

// argument block of the row-team sweep (v1 layout) on `side`
// Both sides on the row-team kernel, one GPU, no streamed rows: the sweeps write the gather copies themselves
// (SweepParams::km_out) and the run loops skip the transposes behind them.  BNMTF_NO_FOLD=1 keeps the separate launches.
static bool fold_ok(const bnmtf_nmf_s* h) {
  static const bool off = [] { const char* e = getenv("BNMTF_NO_FOLD"); return e && atoi(e) != 0; }();
  return !off && !h->nmtf && h->world <= 1 && !h->s[0].v2 && !h->s[1].v2 && !h->s[0].long_rows && !h->s[1].long_rows;
}

static void fill_sweep_params(bnmtf_nmf_s* h, int method, int side, int single_k, uint64_t seed, uint32_t iteration,
                              double* out_tau, double* out_mu, bool want_mu, const double* proj_y2, int64_t proj_ld2,
                              int proj_k2, SweepParams& p, bool fold) {
  Side& sd = h->s[side];
  const bool mu_out = (method == BNMTF_VB) || want_mu;
  p.vals = sd.vals; p.cols = sd.cols; p.rowstart = sd.rowstart;
  p.n_loc = (int)sd.n_loc; p.row0 = sd.row0; p.K = sd.kf; p.tpr = sd.tpr; p.padcol = sd.padcol;
  p.Ykm = sd.ysrc; p.ldy = sd.ysrc_ld;
  p.Ykm2 = proj_y2; p.ldy2 = proj_ld2; p.K2 = proj_k2; p.out3 = h->proj_out3;
  if (h->cov_side == side && method == BNMTF_VB && single_k >= -2) {
    p.covA = h->covA; p.covS = h->covS; p.covL = h->covL; p.cov_sk = h->cov_sk; p.cov_sl = h->cov_sl;
  }
  p.Xval = sd.val; p.Xvar = sd.var;
  p.Xmu = mu_out ? sd.mu : nullptr; p.Xtau = mu_out ? sd.tau : nullptr;
  p.lam_rm = h->hyper.ARD ? nullptr : sd.lam_rm;
  p.lamk = (h->hyper.ARD && method != BNMTF_NP) ? sd.lamk : nullptr;
  p.scal = h->scal;
  p.out_tau = out_tau; p.out_mu = out_mu;
  p.rowstats = sd.rowstats + (size_t)sd.row0 * NRS;
  p.rmean = h->rmean;
  p.seed = seed; p.iteration = iteration; p.purpose = side == 0 ? RNG_TN_U : RNG_TN_V;
  p.iter_ptr = h->iter_ptr;
  p.single_k = single_k;
  if (single_k == -1 && fold) {
    p.km_out = sd.km; p.km_ld = sd.ld;
    p.trace_out = h->fold_trace[side]; p.trace_n = sd.n_glob;
  }
}

// streamed rows (kernels_long.cuh): one CTA per row
static int launch_sweep_long(bnmtf_nmf_s* h, int method, const SweepParams& p, Side& sd) {
  const size_t smem = sizeof(double) * (4 * (size_t)p.K + 3 * (LONG_T / 32) + 2 + 2 * (size_t)p.covL + (method == BNMTF_GIBBS ? TN_PRE_DOUBLES * (size_t)p.K : 0));
  if (smem > 200 * 1024) FAIL("too many factors (%d) for the streamed row sweep", p.K);
#define LONGK(MODE)                                                                                              \
  do {                                                                                                           \
    auto kern = sweep_long_kernel<MODE>;                                                                         \
    if (sd.occ_long[method] == 0) {                                                                              \
      if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
      int occ = 0;                                                                                               \
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, LONG_T, smem));                               \
      if (occ < 1) FAIL("streamed row sweep does not fit (smem %zu)", smem);                                     \
      sd.occ_long[method] = occ;                                                                                 \
    }                                                                                                            \
    const int grid = (int)std::min<int64_t>(p.n_loc, (int64_t)sd.occ_long[method] * h->num_sms);                 \
    if (grid >= 1) { kern<<<grid, LONG_T, smem, h->stream>>>(p, sd.escr); h->launches++; }                       \
  } while (0)
  switch (method) {
    case BNMTF_GIBBS: LONGK(M_GIBBS); break;
    case BNMTF_ICM: LONGK(M_ICM); break;
    case BNMTF_VB: LONGK(M_VB); break;
    case BNMTF_NP: LONGK(M_NP); break;
    default: FAIL("unknown method %d", method);
  }
#undef LONGK
  return 0;
}

// side: which factor is updated.  single_k: see SweepParams.
static int launch_sweep(bnmtf_nmf_s* h, int method, int side, int single_k, uint64_t seed, uint32_t iteration,
                        double* out_tau, double* out_mu, bool want_mu, const double* proj_y2 = nullptr,
                        int64_t proj_ld2 = 0, int proj_k2 = 0) {
  Side& sd = h->s[side];
  if (!sd.packed) FAIL("call bnmtf_nmf_pack first");
  const bool mu_out = (method == BNMTF_VB) || want_mu;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  const bool timed = h->time_sweeps && single_k == -1;
  if (timed) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, h->stream); }
  int rc;
  // (tri-factorisation VB: the per-row-chain kernel carries the covariance term; the lockstep cluster kernel does not)
  const bool nmtf_vb = h->nmtf && method == BNMTF_VB;
  if (sd.v2 && single_k == -1 && method != BNMTF_NP && (!nmtf_vb || (sd.v5 && h->cov_side == side && !getenv("BNMTF_VB_V1")))) {
    Sweep2Params p{};
    p.tvals = sd.tvals; p.tcols = sd.tcols; p.nsets = (int)sd.nsets; p.n_loc = (int)sd.n_loc; p.row0 = sd.row0;
    p.K = sd.kf; p.tw = sd.tw; p.rs = sd.rs; p.Ykm = sd.ysrc; p.ldy = sd.ysrc_ld;
    p.Xval = sd.val; p.Xvar = sd.var; p.Xmu = mu_out ? sd.mu : nullptr; p.Xtau = mu_out ? sd.tau : nullptr;
    p.lam_rm = h->hyper.ARD ? nullptr : sd.lam_rm;
    p.lamk = h->hyper.ARD ? sd.lamk : nullptr;
    p.scal = h->scal;
    p.rowstats = sd.rowstats + (size_t)sd.row0 * NRS;
    p.rmean = h->rmean; p.seed = seed; p.iteration = iteration; p.purpose = side == 0 ? RNG_TN_U : RNG_TN_V;
    p.iter_ptr = h->iter_ptr;
    if (h->fuse_proj && side == 0 && h->nmtf) {   // NMTF: P_il = sum_j e_ij G_jl from the residual the F sweep ends with
      p.Ykm2 = h->s[1].km; p.ldy2 = h->s[1].ld; p.Yproj = h->s[1].kmB; p.ldp = h->s[1].ld; p.K2 = h->L; p.proj_out = h->Prow;
    }
    if (nmtf_vb) {                                // the VB F sweep asks for P through the projection arguments of the call
      if (proj_y2 != nullptr) { p.Yproj = h->s[1].kmB; p.ldp = h->s[1].ld; p.K2 = proj_k2; p.proj_out = out_mu; }
      p.covA = h->covA; p.covS = h->covS; p.covL = h->covL; p.cov_sk = h->cov_sk; p.cov_sl = h->cov_sl;
    }
    p.dbg = getenv("BNMTF_V2_DBG") ? atoi(getenv("BNMTF_V2_DBG")) : 0;
    static unsigned long long* prof_buf = nullptr;   // timing experiments only
    if (p.dbg & 2) {
      if (!prof_buf) CK(cudaMallocManaged(&prof_buf, sizeof(unsigned long long) * 32));
      CK(cudaStreamSynchronize(h->stream));
      memset(prof_buf, 0, sizeof(unsigned long long) * 32);
      p.prof = prof_buf;
      bnmtf_prof_last = prof_buf;
    }
    rc = 0;
    if (sd.v4 || sd.v5) {
      if (h->nmtf) { p.Ykm = sd.ycombB; p.Ykm2 = sd.ycombBq; p.ldy = sd.ysrc_ld; }   // W = G S^T / Z = F S, block-major (nmtf_make_W / nmtf_make_Z)
      else { p.Ykm = h->s[1 - side].kmB; p.Ykm2 = h->s[1 - side].kmBq; p.ldy = h->s[1 - side].ld; }
      if (method != BNMTF_GIBBS && method != BNMTF_ICM && method != BNMTF_VB) FAIL("unknown method %d", method);
      if (method == BNMTF_GIBBS) {
        const size_t need = sd.v5 ? sweep5_gmat_doubles(sd.rs, sd.kf) : sweep4_gmat_doubles(sd.rs, sd.kf);
        if (h->gmat_doubles < need) {
          CK(cudaStreamSynchronize(h->stream));
          cudaFree(h->gmat); h->gmat = nullptr; h->gmat_doubles = 0;
          CK(cudaMalloc(&h->gmat, sizeof(double) * need));
          h->gmat_doubles = need;
        }
        p.gmat = h->gmat;
      }
      if (sd.v5 && h->peer_ok && h->world > 1) {
        p.peer_val = reinterpret_cast<double* const*>(h->peer_tab_dev + (size_t)(side * 3 + 0) * bnmtf_nmf_s::MAXPEER);
        p.peer_var = reinterpret_cast<double* const*>(h->peer_tab_dev + (size_t)(side * 3 + 1) * bnmtf_nmf_s::MAXPEER);
        p.peer_rowstats = reinterpret_cast<double* const*>(h->peer_tab_dev + (size_t)(side * 3 + 2) * bnmtf_nmf_s::MAXPEER);
        p.npeers = h->world - 1;
      }
      if (sd.v5) {
        if (!h->set_counter) CK(cudaMalloc(&h->set_counter, sizeof(int)));
        CK(cudaMemsetAsync(h->set_counter, 0, sizeof(int), h->stream));
        p.set_counter = h->set_counter;
      }
      std::string err;
      rc = sd.v5 ? launch_sweep5(sd.npt2, method, p, h->stream,
                                 p.K2 > 0 ? &sd.nclusters5p[method][p.Xmu != nullptr ? 1 : 0] : &sd.nclusters4[method][p.Xmu != nullptr ? 1 : 0], &err)   // (one kernel per slot: VB sweeps of a tri-factorisation side always carry the covariance term)
                 : launch_sweep4(sd.npt2, method, p, h->stream, &sd.nclusters4[method][p.Xmu != nullptr ? 1 : 0], &err);
      if (rc) { g_err = err; return 1; }
      h->launches++;
    } else
    switch (method) {
      case BNMTF_GIBBS: rc = launch_sweep2_m<M_GIBBS>(h, p, sd, method); break;
      case BNMTF_ICM: rc = launch_sweep2_m<M_ICM>(h, p, sd, method); break;
      case BNMTF_VB: rc = launch_sweep2_m<M_VB>(h, p, sd, method); break;
      default: FAIL("unknown method %d", method);
    }
  } else {
    SweepParams p{};
    fill_sweep_params(h, method, side, single_k, seed, iteration, out_tau, out_mu, want_mu, proj_y2, proj_ld2, proj_k2, p, fold_ok(h));
    if (sd.long_rows) {
      rc = launch_sweep_long(h, method, p, sd);
    } else
    switch (method) {
      case BNMTF_GIBBS: rc = launch_sweep_m<M_GIBBS>(h, p, p.n_loc, sd.npt, &sd.occ[method]); break;
      case BNMTF_ICM: rc = launch_sweep_m<M_ICM>(h, p, p.n_loc, sd.npt, &sd.occ[method]); break;
      case BNMTF_VB: rc = launch_sweep_m<M_VB>(h, p, p.n_loc, sd.npt, &sd.occ[method]); break;
      case BNMTF_NP: rc = launch_sweep_m<M_NP>(h, p, p.n_loc, sd.npt, &sd.occ[method]); break;
      default: FAIL("unknown method %d", method);
    }
  }
  if (getenv("BNMTF_V2_DBG") && (atoi(getenv("BNMTF_V2_DBG")) & 2) && sd.v2 && single_k == -1 && method != BNMTF_NP) {
    // phase clocks of cluster 0 / CTA 0 (control warp: slots 0-4, first compute warp: 8-15), in cycles per set
    cudaStreamSynchronize(h->stream);
    unsigned long long* pb = bnmtf_prof_last;
    if (pb && pb[16] > 0 && sd.v5) {
      fprintf(stderr, "[v5 prof side %d] rows %llu | start %llu | resid: tilewait %llu pass %llu | meet %llu table %llu | update: tilewait %llu pass %llu reduce+send %llu recwait %llu cond %llu | tail %llu (cycles/row) | warp loop %llu, before it %llu, CTA %llu cycles\n",
              side, pb[16], pb[0] / pb[16], pb[1] / pb[16], pb[2] / pb[16], pb[3] / pb[16], pb[4] / pb[16], pb[5] / pb[16], pb[6] / pb[16], pb[7] / pb[16],
              pb[8] / pb[16], pb[9] / pb[16], pb[10] / pb[16], pb[20], pb[21], pb[22]);
    } else
    if (pb && pb[16] > 0 && sd.v4) {
      fprintf(stderr, "[v4 prof side %d] sets %llu | ctrl: start+phaseA %llu prep %llu slotwait %llu recwait %llu cond %llu book %llu tail %llu | comp: start %llu phaseA %llu statx %llu corr %llu tilewait %llu dot %llu reduce %llu deltawait %llu tail %llu (cycles/set)\n",
              side, pb[16], pb[0] / pb[16], pb[1] / pb[16], pb[2] / pb[16], pb[3] / pb[16], pb[4] / pb[16], pb[5] / pb[16], pb[6] / pb[16], pb[8] / pb[16],
              pb[9] / pb[16], pb[10] / pb[16], pb[11] / pb[16], pb[12] / pb[16], pb[13] / pb[16], pb[14] / pb[16], pb[15] / pb[16], pb[17] / pb[16]);
    } else
    if (pb && pb[16] > 0) {
      fprintf(stderr, "[v3 prof side %d] sets %llu | ctrl: pre %llu wait_pass %llu cbar %llu update %llu sync %llu | comp: start %llu resid %llu tilewait %llu pass %llu reduce %llu cbar %llu deltawait %llu tail %llu (cycles/set) | fallback lanes %llu, columns with a fallback %llu of %llu\n",
              side, pb[16], pb[0] / pb[16], pb[1] / pb[16], pb[2] / pb[16], pb[3] / pb[16], pb[4] / pb[16], pb[8] / pb[16],
              pb[9] / pb[16], pb[10] / pb[16], pb[11] / pb[16], pb[12] / pb[16], pb[13] / pb[16], pb[14] / pb[16], pb[15] / pb[16], pb[20], pb[21], pb[22]);
    }
  }
  if (timed) {
    cudaEventRecord(e1, h->stream);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    h->sweep_ms[side] += ms; h->sweep_n[side]++;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
  return rc;
}

static int launch_lambda(bnmtf_nmf_s* h, int method, uint64_t seed, uint32_t iteration, int update) {
  Side& su = h->s[0]; Side& sv = h->s[1];
  const int K = h->K;
#define LK(MODE)                                                                                                      \
  lambda_kernel<MODE><<<K, 1024, 0, h->stream>>>(su.km, su.ld, h->I, sv.km, sv.ld, h->J, K, h->hyper, h->colsum,       \
                                                h->lamstate, h->lamk, seed, iteration, update, h->iter_ptr, h->fold_lam_trace)
  switch (method) {
    case BNMTF_GIBBS: LK(M_GIBBS); break;
    case BNMTF_ICM: LK(M_ICM); break;
    case BNMTF_VB: LK(M_VB); break;
    default: return 0;
  }
#undef LK
  h->launches++;
  return 0;
}

static int launch_finalize(bnmtf_nmf_s* h, int method, double* stats_out, uint64_t seed, uint32_t iteration, int update) {
  Side& su = h->s[0]; Side& sv = h->s[1];
#define FK(MODE)                                                                                                    \
  finalize_kernel<MODE><<<1, 1024, 0, h->stream>>>(su.rowstats, su.n_glob, sv.rowstats, sv.n_glob, h->K, h->hyper,   \
                                                   h->colsum, h->lamstate, h->scal, stats_out, seed, iteration,     \
                                                   update, h->I, h->J, h->iter_ptr ? h->iter_dev : nullptr)
  switch (method) {
    case BNMTF_GIBBS: FK(M_GIBBS); break;
    case BNMTF_ICM: FK(M_ICM); break;
    case BNMTF_VB: FK(M_VB); break;
    case BNMTF_NP: FK(M_NP); break;
    default: FAIL("unknown method %d", method);
  }
#undef FK
  h->launches++;
  return 0;
}

// after a sweep on `side`: make every rank hold the whole factor (and row stats)
static int peer_sync_end_of_iteration(bnmtf_nmf_s* h) {
  if (h->world <= 1 || !h->peer_ok || !(h->s[0].v5 || h->s[1].v5)) return 0;
  std::string err;
  if (launch_peer_barrier(reinterpret_cast<int* const*>(h->peer_tab_dev + (size_t)6 * bnmtf_nmf_s::MAXPEER), h->peer_flags, h->rank,
                          h->world, ++h->peer_epoch, h->stream, &err)) { g_err = err; return 1; }
  h->launches++;
  return 0;
}

static int allgather_side(bnmtf_nmf_s* h, int method, int side, bool stats) {
  if (h->world <= 1) return 0;
  Side& sd = h->s[side];
  if (sd.v5 && h->peer_ok && method != BNMTF_NP) {
    // the sweep wrote its rows (values, VB variances, row statistics) into every rank's arrays: only the barrier is left
    std::string err;
    if (launch_peer_barrier(reinterpret_cast<int* const*>(h->peer_tab_dev + (size_t)6 * bnmtf_nmf_s::MAXPEER), h->peer_flags, h->rank,
                            h->world, ++h->peer_epoch, h->stream, &err)) { g_err = err; return 1; }
    h->launches++;
    return 0;
  }
  const size_t cnt = (size_t)sd.n_shard * sd.kf;
  double* fields[4] = {sd.val, method == BNMTF_VB ? sd.var : nullptr, nullptr, nullptr};
  for (double* f : fields) {
    if (!f) continue;
    int rc = g_nccl.AllGather(f + (size_t)sd.row0 * sd.kf, f, cnt, NCCL_FLOAT64, h->comm, h->stream);
    if (rc != 0) FAIL("ncclAllGather: %s", g_nccl.GetErrorString(rc));
  }
  if (stats) {
    int rc = g_nccl.AllGather(sd.rowstats + (size_t)sd.row0 * NRS, sd.rowstats, (size_t)sd.n_shard * NRS, NCCL_FLOAT64,
                              h->comm, h->stream);
    if (rc != 0) FAIL("ncclAllGather: %s", g_nccl.GetErrorString(rc));
  }
  return 0;
}

int bnmtf_nmf_init_factors(bnmtf_nmf_t h, int method, int random, uint64_t seed) {
  if (!h) FAIL("NULL handle");
  ENTER(h);
  for (int sidx = 0; sidx < 2; ++sidx) {
    Side& sd = h->s[sidx];
    const int64_t n = sd.n_glob * sd.kf;
    double* dst = method == BNMTF_VB ? sd.mu : sd.val;
    init_factor_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(
        dst, sd.n_glob, sd.kf, h->hyper.ARD ? nullptr : sd.lam_rm, h->hyper.ARD ? sd.lamk : nullptr, random, seed,
        sidx == 0 ? RNG_INIT_U : RNG_INIT_V);
    h->launches++;
    if (method == BNMTF_VB) {
      vb_init_moments_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(sd.mu, sd.tau, sd.val, sd.var, n);
      h->launches++;
    }
  }
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  h->km_mode = -1;
  return 0;
}

int bnmtf_nmf_column_params(bnmtf_nmf_t h, int method, int side, int k, double* out_tau, double* out_mu) {
  if (!h || !out_tau || !out_mu || side < 0 || side > 1) FAIL("bad argument");
  if (k < 0 || k >= h->s[side].kf) FAIL("k out of range");
  if (h->nmtf) FAIL("use bnmtf_nmtf_column_params for a tri-factorisation handle");
  if (h->world > 1) FAIL("column_params is single-GPU only");
  ENTER(h);
  Side& sd = h->s[side];
  double *dt = nullptr, *dm = nullptr;
  CK(cudaMalloc(&dt, sizeof(double) * sd.n_glob));
  CK(cudaMalloc(&dm, sizeof(double) * sd.n_glob));
  refresh_km(h, 0, method); refresh_km(h, 1, method);
  h->km_mode = method == BNMTF_VB ? 2 : 1;
  if (launch_sweep(h, method, side, k, 0, 0, dt, dm, false)) return 1;
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  CK(cudaMemcpy(out_tau, dt, sizeof(double) * sd.n_glob, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(out_mu, dm, sizeof(double) * sd.n_glob, cudaMemcpyDeviceToHost));
  cudaFree(dt); cudaFree(dm);
  return 0;
}

int bnmtf_nmf_stats(bnmtf_nmf_t h, int method, double* out) {
  if (!h || !out) FAIL("bad argument");
  ENTER(h);
  refresh_km(h, 0, method); refresh_km(h, 1, method);
  h->km_mode = method == BNMTF_VB ? 2 : 1;
  if (method == BNMTF_VB && launch_sweep(h, method, 0, -2, 0, 0, nullptr, nullptr, false)) return 1;
  if (method == BNMTF_VB && allgather_side(h, method, 0, true)) return 1;
  if (launch_sweep(h, method, 1, -2, 0, 0, nullptr, nullptr, false)) return 1;
  if (allgather_side(h, method, 1, true)) return 1;
  if (launch_lambda(h, method, 0, 0, 0)) return 1;
  double* sdev = nullptr;
  CK(cudaMalloc(&sdev, sizeof(double) * BNMTF_NSTATS));
  if (launch_finalize(h, method, sdev, 0, 0, 0)) return 1;
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  CK(cudaMemcpy(out, sdev, sizeof(double) * BNMTF_NSTATS, cudaMemcpyDeviceToHost));
  cudaFree(sdev);
  return 0;
}

// ---- draw traces to caller memory -------------------------------------------------------
// all_U / all_V of a long run are gigabytes of pageable (usually never touched) caller memory.  A device-to-
// host copy straight into it is staged by the driver and blocks the calling thread, which then cannot keep
// the GPU fed, and the first touch of every page is paid inside that copy.  Instead each chunk of iterations
// goes device -> pinned slot (asynchronous, PCIe speed) on the copy stream, and a host worker thread moves
// it from the pinned slot into the caller's arrays while the GPU runs the next chunks.  Two pinned slots,
// kept with the handle between runs.
namespace {
class TraceDrain {
 public:
  struct Part { const void* dev; char* host; size_t bytes; };
  bool active = false;

  int start(bnmtf_nmf_s* h, int nchunks, size_t slot_bytes) {
    h_ = h;
    slot_bytes_ = slot_bytes;   // the worker allocates the pinned slots (tens of milliseconds) while the caller enqueues the first chunk
    ready_.resize(nchunks, nullptr);
    for (auto& e : ready_) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    jobs_.assign(nchunks, Job{});
    submitted_ = 0; done_ = 0; stop_ = false; slots_ready_ = false; active = true;
    worker_ = std::thread([this] { this->loop(); });
    return 0;
  }

  // chunk c (in order): its parts are complete on the device once `dev_ready` has fired.  Returns after the
  // copies into pinned slot c & 1 are enqueued on the copy stream; `reusable` is recorded behind them.
  int submit(int c, cudaEvent_t dev_ready, const Part* parts, int nparts, cudaEvent_t reusable) {
    {  // the slot exists, and is free once the worker has moved chunk c - 2 out of it
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [&] { return slots_ready_ && done_ >= c - 1; });
    }
    if (failed_) FAIL("pinned host memory for the draw traces: %s", cudaGetErrorString((cudaError_t)err_.load()));
    cudaStream_t cs = h_->copy_stream;
    CK(cudaStreamWaitEvent(cs, dev_ready, 0));
    Job j{};
    char* slot = h_->pin[c & 1];
    size_t off = 0;
    for (int q = 0; q < nparts && q < 3; ++q) {
      if (off + parts[q].bytes > h_->pin_bytes) FAIL("trace chunk larger than its pinned slot");
      CK(cudaMemcpyAsync(slot + off, parts[q].dev, parts[q].bytes, cudaMemcpyDeviceToHost, cs));
      j.src[q] = slot + off; j.dst[q] = parts[q].host; j.bytes[q] = parts[q].bytes;
      off += parts[q].bytes;
    }
    j.n = std::min(nparts, 3);
    CK(cudaEventRecord(ready_[c], cs));
    CK(cudaEventRecord(reusable, cs));
    {
      std::lock_guard<std::mutex> lk(mu_);
      jobs_[c] = j;
      submitted_ = c + 1;
    }
    cv_.notify_all();
    return 0;
  }

  // all submitted chunks are in the caller's arrays when this returns
  int finish() {
    if (!active) return 0;
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    if (worker_.joinable()) worker_.join();
    for (auto& e : ready_) if (e) cudaEventDestroy(e);
    ready_.clear();
    active = false;
    if (failed_) FAIL("copying the draw traces to host memory failed: %s", cudaGetErrorString((cudaError_t)err_.load()));
    return 0;
  }

  ~TraceDrain() { finish(); }

 private:
  struct Job { const char* src[3]; char* dst[3]; size_t bytes[3]; int n; };
  // the destination pages are usually untouched: six threads take the page faults side by side
  static void spread_memcpy(char* dst, const char* src, size_t bytes) {
    constexpr int NT = 6;
    if (bytes < ((size_t)4 << 20)) { memcpy(dst, src, bytes); return; }
    const size_t piece = (bytes / NT + 4095) & ~(size_t)4095;
    std::thread th[NT - 1];
    for (int q = 1; q < NT; ++q) {
      const size_t a = std::min(bytes, piece * q), b = std::min(bytes, piece * (q + 1));
      th[q - 1] = std::thread([=] { if (b > a) memcpy(dst + a, src + a, b - a); });
    }
    memcpy(dst, src, std::min(bytes, piece));
    for (auto& t : th) t.join();
  }
  void loop() {
    cudaSetDevice(h_->device);
    if (h_->pin_bytes < slot_bytes_) {
      for (int b = 0; b < 2; ++b) { if (h_->pin[b]) cudaFreeHost(h_->pin[b]); h_->pin[b] = nullptr; }
      h_->pin_bytes = 0;
      cudaError_t e = cudaSuccess;
      for (int b = 0; b < 2 && e == cudaSuccess; ++b) e = cudaHostAlloc((void**)&h_->pin[b], slot_bytes_, cudaHostAllocDefault);
      if (e == cudaSuccess) h_->pin_bytes = slot_bytes_; else { err_ = (int)e; failed_ = true; }
    }
    {
      std::lock_guard<std::mutex> lk(mu_);
      slots_ready_ = true;
    }
    cv_.notify_all();
    for (int i = 0;; ++i) {
      Job j;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return submitted_ > i || stop_; });
        if (submitted_ <= i) return;
        j = jobs_[i];
      }
      cudaError_t e = cudaEventSynchronize(ready_[i]);
      if (e != cudaSuccess) { err_ = (int)e; failed_ = true; }
      else for (int q = 0; q < j.n; ++q) spread_memcpy(j.dst[q], j.src[q], j.bytes[q]);
      {
        std::lock_guard<std::mutex> lk(mu_);
        done_ = i + 1;
      }
      cv_.notify_all();
    }
  }
  bnmtf_nmf_s* h_ = nullptr;
  std::vector<cudaEvent_t> ready_;
  std::vector<Job> jobs_;
  std::thread worker_;
  std::mutex mu_;
  std::condition_variable cv_;
  int submitted_ = 0, done_ = 0;
  size_t slot_bytes_ = 0;
  bool slots_ready_ = false;
  bool stop_ = false;                        // guarded by mu_
  std::atomic<bool> failed_{false};          // written by the worker, read by submit() / finish() on the caller's thread
  std::atomic<int> err_{(int)cudaSuccess};
};

// iterations per device-side trace chunk: everything at once while it fits 256 MB (the run loop can then be
// replayed as a graph), else 16 MB chunks (at least one iteration) so that the last chunk's way to the caller's
// arrays is short
int trace_chunk_iterations(int iterations, size_t per_it) {
  per_it = std::max<size_t>(per_it, 1);
  size_t whole = (size_t)256 << 20, piece = (size_t)16 << 20;
  if (const char* e = getenv("BNMTF_TRACE_CHUNK_BYTES")) whole = piece = (size_t)std::max(1ll, atoll(e));   // tests: chunked traces of small models
  if ((size_t)iterations * per_it <= whole) return iterations;
  return (int)std::max<size_t>(1, std::min<size_t>((size_t)iterations, piece / per_it));
}
}  // namespace

// grow-only device scratch kept on the handle between runs
static int ensure_dev(double** ptr, size_t* cap, size_t bytes) {
  if (*ptr && *cap >= bytes) return 0;
  cudaFree(*ptr); *ptr = nullptr; *cap = 0;
  CK(cudaMalloc(ptr, std::max<size_t>(bytes, 8)));
  *cap = bytes;
  return 0;
}
static int ensure_events(bnmtf_nmf_s* h, size_t n) {
  while (h->ev_pool.size() < n) {
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    h->ev_pool.push_back(e);
  }
  for (int q = 0; q < 4; ++q) if (!h->chunk_ev[q]) CK(cudaEventCreate(&h->chunk_ev[q]));
  return 0;
}

// Launch-bound run loops: iteration 0 eagerly (first-use attribute / occupancy calls happen here), iteration 1 captured
// into a CUDA graph, the graph replayed for the rest.  While this runs h->iter_ptr is set: kernels read the iteration
// index (Philox counters, trace and statistics slots) from the device and the finalize kernel advances it.
static bool graph_mode_allowed(const bnmtf_nmf_s* h, int iterations) {
  const char* vdbg = getenv("BNMTF_V2_DBG");
  return h->world <= 1 && iterations >= 8 && h->sum_thin <= 0 && !h->time_sweeps && !getenv("BNMTF_NO_GRAPH") &&
         !(vdbg && (atoi(vdbg) & 2));
}

template <class F>
static int run_graph_replays(bnmtf_nmf_s* h, int iterations, std::vector<cudaEvent_t>& ev, F&& one_iteration) {
  if (!h->iter_dev) CK(cudaMalloc(&h->iter_dev, sizeof(uint32_t)));
  CK(cudaMemsetAsync(h->iter_dev, 0, sizeof(uint32_t), h->stream));
  h->iter_ptr = h->iter_dev;
  int rc = one_iteration();
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t gexec = nullptr;
  cudaError_t ce = cudaSuccess;
  if (!rc) {
    ce = cudaEventRecord(ev[1], h->stream);
    const int64_t l0 = h->launches;
    if (ce == cudaSuccess) ce = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
    if (ce == cudaSuccess) {
      rc = one_iteration();
      ce = cudaStreamEndCapture(h->stream, &graph);
    }
    const int64_t per_it = h->launches - l0;
    if (!rc && ce == cudaSuccess) ce = cudaGraphInstantiate(&gexec, graph, 0);
    if (!rc && ce == cudaSuccess) {
      for (int it = 1; it < iterations && ce == cudaSuccess; ++it) {
        ce = cudaGraphLaunch(gexec, h->stream);
        if (ce == cudaSuccess) ce = cudaEventRecord(ev[it + 1], h->stream);
      }
      h->launches += per_it * (int64_t)(iterations - 2);
    }
    if (!rc && ce != cudaSuccess && gexec == nullptr) {
      // capture or instantiation failed (nothing of the captured iteration has run): the eager loop serves the rest;
      // the iteration index stays on the device, so the chain is the same
      cudaGetLastError();
      ce = cudaSuccess;
      for (int it = 1; it < iterations && !rc && ce == cudaSuccess; ++it) {
        rc = one_iteration();
        if (!rc) ce = cudaEventRecord(ev[it + 1], h->stream);
      }
    }
    if (gexec) cudaGraphExecDestroy(gexec);
    if (graph) cudaGraphDestroy(graph);
  }
  h->iter_ptr = nullptr;
  if (rc) return 1;
  if (ce != cudaSuccess) FAIL("graph replay of the iteration failed: %s", cudaGetErrorString(ce));
  return 0;
}

// run(iterations): see include/bnmtf_b200.h
int bnmtf_nmf_run(bnmtf_nmf_t h, int method, int iterations, uint64_t seed, double* all_U, double* all_V,
                  double* all_lambdak, double* iter_stats, double* iter_seconds) {
  if (!h) FAIL("NULL handle");
  if (iterations < 0) FAIL("negative iteration count");
  if (method < 0 || method > 3) FAIL("unknown method %d", method);
  if (h->nmtf) FAIL("use bnmtf_nmtf_run for a tri-factorisation handle");
  ENTER(h);
  const int K = h->K;
  Side& su = h->s[0]; Side& sv = h->s[1];
  const int want_mu = getenv("BNMTF_TRACE_PARAMS") && atoi(getenv("BNMTF_TRACE_PARAMS")) != 0;
  h->sweep_ms[0] = h->sweep_ms[1] = 0; h->sweep_n[0] = h->sweep_n[1] = 0;
  if (iterations == 0) return 0;
  const bool run_timing = getenv("BNMTF_RUN_TIMING") != nullptr;   // host wall clock of the stages of this call, to stderr
  const auto wall0 = std::chrono::steady_clock::now();
  auto wall_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count(); };
  double t_setup = 0, t_enq = 0, t_sync = 0, t_drain = 0;

  // traces are staged on the device in chunks and copied out while the next chunk runs
  const size_t per_it = sizeof(double) * ((size_t)su.n_glob + sv.n_glob) * K;
  const bool tr = (all_U != nullptr) || (all_V != nullptr);
  int chunk = iterations;
  if (tr) chunk = trace_chunk_iterations(iterations, per_it);
  if (tr) {
    for (int b = 0; b < 2; ++b) {
      if (ensure_dev(&su.trace[b], &su.trace_cap[b], sizeof(double) * su.n_glob * K * chunk)) return 1;
      if (ensure_dev(&sv.trace[b], &sv.trace_cap[b], sizeof(double) * sv.n_glob * K * chunk)) return 1;
    }
  }
  TraceDrain drain;   // joins its worker on every way out of this function
  if (ensure_dev(&h->stats_dev, &h->stats_cap, sizeof(double) * BNMTF_NSTATS * iterations)) return 1;
  if (ensure_dev(&h->lam_trace, &h->lam_trace_cap, sizeof(double) * K * iterations)) return 1;
  double* lam_trace = h->lam_trace;
  CK(cudaMemsetAsync(lam_trace, 0, sizeof(double) * K * iterations, h->stream));

  if (ensure_events(h, (size_t)iterations + 1)) return 1;
  std::vector<cudaEvent_t> ev(h->ev_pool.begin(), h->ev_pool.begin() + iterations + 1);
  cudaEvent_t chunk_done[2] = {h->chunk_ev[0], h->chunk_ev[1]}, chunk_copied[2] = {h->chunk_ev[2], h->chunk_ev[3]};

  // peer exchange: no rank's first sweep may write into another rank's arrays before that rank has set up its own state
  // for this run (the uploads of set_factor are complete when the caller gets here)
  if (peer_sync_end_of_iteration(h)) return 1;
  refresh_km(h, 0, method); refresh_km(h, 1, method);
  h->km_mode = method == BNMTF_VB ? 2 : 1;
  if (summary_reset(h)) return 1;
  CK(cudaEventRecord(ev[0], h->stream));
  t_setup = wall_ms();

  const int nchunks = (iterations + chunk - 1) / chunk;
  if (tr && nchunks > 1 && !getenv("BNMTF_NO_PINNED_TRACES") && drain.start(h, nchunks, per_it * (size_t)chunk)) return 1;
  auto copy_chunk = [&](int c) -> int {
    const int b = c & 1;
    const int it0 = c * chunk, n_it = std::min(chunk, iterations - it0);
    const size_t bU = sizeof(double) * su.n_glob * K * n_it, bV = sizeof(double) * sv.n_glob * K * n_it;
    if (drain.active) {
      TraceDrain::Part parts[2];
      int np = 0;
      if (all_U) parts[np++] = {su.trace[b], (char*)(all_U + (size_t)it0 * su.n_glob * K), bU};
      if (all_V) parts[np++] = {sv.trace[b], (char*)(all_V + (size_t)it0 * sv.n_glob * K), bV};
      return drain.submit(c, chunk_done[b], parts, np, chunk_copied[b]);
    }
    CK(cudaStreamWaitEvent(h->copy_stream, chunk_done[b], 0));
    if (all_U) CK(cudaMemcpyAsync(all_U + (size_t)it0 * su.n_glob * K, su.trace[b], bU, cudaMemcpyDeviceToHost, h->copy_stream));
    if (all_V) CK(cudaMemcpyAsync(all_V + (size_t)it0 * sv.n_glob * K, sv.trace[b], bV, cudaMemcpyDeviceToHost, h->copy_stream));
    CK(cudaEventRecord(chunk_copied[b], h->copy_stream));
    return 0;
  };

  // Small problems are launch-bound (~9 kernels of a few microseconds per iteration): replay the iteration as a graph
  const bool use_graph = graph_mode_allowed(h, iterations) && nchunks == 1;
  if (use_graph) {
    const bool lam_tr = h->hyper.ARD && method != BNMTF_NP;
    // launch-bound models: the sweeps write the gather copies and their trace slots, the lambda kernel its own (4 launches
    // per iteration instead of 7)
    const bool fold = fold_ok(h);
    struct FoldReset { bnmtf_nmf_s* h; ~FoldReset() { h->fold_trace[0] = h->fold_trace[1] = nullptr; h->fold_lam_trace = nullptr; } } fold_reset{h};
    if (fold) {
      if (all_U) h->fold_trace[0] = su.trace[0];
      if (all_V) h->fold_trace[1] = sv.trace[0];
      if (lam_tr) h->fold_lam_trace = lam_trace;
    }
    auto one_iteration = [&]() -> int {
      if (launch_lambda(h, method, seed, 0, 1)) return 1;                                  // bnmf_gibbs.py:149-152
      if (launch_sweep(h, method, 0, -1, seed, 0, nullptr, nullptr, want_mu)) return 1;    // :154-158
      if (!fold) refresh_km(h, 0, method);
      if (launch_sweep(h, method, 1, -1, seed, 0, nullptr, nullptr, want_mu)) return 1;    // :160-164
      if (!fold) refresh_km(h, 1, method);
      if (method == BNMTF_VB && launch_lambda(h, method, seed, 0, 0)) return 1;
      if ((tr || lam_tr) && !(fold && (!all_U || h->fold_trace[0]) && (!all_V || h->fold_trace[1]))) {
        const int64_t nU = su.n_glob * K, nV = sv.n_glob * K;
        const unsigned grid = (unsigned)std::min<int64_t>((std::max(nU, nV) + 255) / 256, (int64_t)h->num_sms * 4);
        trace_copy_kernel<<<std::max(grid, 1u), 256, 0, h->stream>>>(tr ? su.trace[0] : nullptr, su.val, nU, tr ? sv.trace[0] : nullptr,
                                                                    sv.val, nV, lam_tr ? lam_trace : nullptr, h->lamk, K, h->iter_dev);
        h->launches++;
      }
      return launch_finalize(h, method, h->stats_dev, seed, 0, 1);                         // :166-167,175
    };
    if (run_graph_replays(h, iterations, ev, one_iteration)) return 1;
    if (tr) CK(cudaEventRecord(chunk_done[0], h->stream));
  }
  for (int c = 0; c < nchunks && !use_graph; ++c) {
    const int b = c & 1;
    if (tr && c >= 2) CK(cudaStreamWaitEvent(h->stream, chunk_copied[b], 0));
    const int it0 = c * chunk, it1 = std::min(iterations, it0 + chunk);
    for (int it = it0; it < it1; ++it) {
      // lambda_k (ARD)                                   bnmf_gibbs.py:149-152
      if (launch_lambda(h, method, seed, (uint32_t)it, 1)) return 1;
      if (h->hyper.ARD && method != BNMTF_NP)
        CK(cudaMemcpyAsync(lam_trace + (size_t)it * K, h->lamk, sizeof(double) * K, cudaMemcpyDeviceToDevice, h->stream));
      // U sweep                                           bnmf_gibbs.py:154-158
      if (launch_sweep(h, method, 0, -1, seed, (uint32_t)it, nullptr, nullptr, want_mu)) return 1;
      if (allgather_side(h, method, 0, method == BNMTF_VB)) return 1;
      if (!fold_ok(h)) refresh_km(h, 0, method);
      // V sweep                                           bnmf_gibbs.py:160-164
      if (launch_sweep(h, method, 1, -1, seed, (uint32_t)it, nullptr, nullptr, want_mu)) return 1;
      if (allgather_side(h, method, 1, true)) return 1;
      if (!fold_ok(h)) refresh_km(h, 1, method);
      // tau, statistics                                   bnmf_gibbs.py:166-167,175
      if (method == BNMTF_VB && launch_lambda(h, method, seed, (uint32_t)it, 0)) return 1;
      if (launch_finalize(h, method, h->stats_dev + (size_t)it * BNMTF_NSTATS, seed, (uint32_t)it, 1)) return 1;
      if (summary_accumulate(h, it, h->stats_dev + (size_t)it * BNMTF_NSTATS)) return 1;
      if (tr) {
        const int slot = it - it0;
        CK(cudaMemcpyAsync(su.trace[b] + (size_t)slot * su.n_glob * K, su.val, sizeof(double) * su.n_glob * K, cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaMemcpyAsync(sv.trace[b] + (size_t)slot * sv.n_glob * K, sv.val, sizeof(double) * sv.n_glob * K, cudaMemcpyDeviceToDevice, h->stream));
      }
      // peer exchange: the other ranks' next U sweep writes into this rank's arrays -- not before everything that reads
      // this iteration's factors here (trace copy, posterior sums, statistics) has run
      if (peer_sync_end_of_iteration(h)) return 1;
      CK(cudaEventRecord(ev[it + 1], h->stream));
    }
    if (tr) {
      CK(cudaEventRecord(chunk_done[b], h->stream));
      if (c >= 1 && copy_chunk(c - 1)) return 1;
    }
  }
  if (tr && copy_chunk(nchunks - 1)) return 1;
  t_enq = wall_ms();
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaStreamSynchronize(h->copy_stream));
  CK(cudaGetLastError());
  t_sync = wall_ms();
  if (iter_stats) CK(cudaMemcpy(iter_stats, h->stats_dev, sizeof(double) * BNMTF_NSTATS * iterations, cudaMemcpyDeviceToHost));
  if (drain.finish()) return 1;
  t_drain = wall_ms();
  if (all_lambdak) CK(cudaMemcpy(all_lambdak, lam_trace, sizeof(double) * K * iterations, cudaMemcpyDeviceToHost));
  if (iter_seconds) {
    for (int it = 0; it < iterations; ++it) {
      float ms = 0; CK(cudaEventElapsedTime(&ms, ev[0], ev[it + 1]));
      iter_seconds[it] = ms * 1e-3;
    }
  }
  if (run_timing) {
    float dev_ms = 0; cudaEventElapsedTime(&dev_ms, ev[0], ev[iterations]);
    fprintf(stderr, "[bnmtf_nmf_run] %d iterations, %d trace chunk(s): set-up %.1f ms, loop enqueued at %.1f, streams idle at %.1f, traces delivered at %.1f, "
            "device loop %.1f ms\n", iterations, tr ? nchunks : 0, t_setup, t_enq, t_sync, t_drain, dev_ms);
  }
  return 0;
}

int64_t bnmtf_nmf_launch_count(bnmtf_nmf_t h) { return h ? h->launches : 0; }

int bnmtf_nmf_layout_info(bnmtf_nmf_t h, int64_t* out8) {  // out8: 16 values
  if (!h || !out8) FAIL("bad argument");
  out8[0] = h->s[0].tpr; out8[1] = h->s[0].npt; out8[2] = h->s[1].tpr; out8[3] = h->s[1].npt;
  out8[4] = h->s[0].nslots; out8[5] = h->s[1].nslots; out8[6] = h->s[0].nnz; out8[7] = h->s[1].nnz;
  for (int q = 0; q < 2; ++q) {
    const Side& sd = h->s[q];
    out8[8 + 4 * q] = sd.v2 ? (sd.v5 ? 5 : (sd.v4 ? 4 : 1)) : 0;   // 1: cluster sweep (kernels_v2.cuh), 4: blocked (kernels_v4.cuh), 5: per-row chains (kernels_v5.cuh)
    out8[9 + 4 * q] = sd.v2 ? sd.npt2 : 0;
    out8[10 + 4 * q] = sd.v2 ? sd.rs : 0; out8[11 + 4 * q] = sd.v2 ? sd.tslots : 0;
  }
  return 0;
}

int bnmtf_nmf_time_sweeps(bnmtf_nmf_t h, int on) {
  if (!h) FAIL("NULL handle");
  h->time_sweeps = on != 0;
  h->sweep_ms[0] = h->sweep_ms[1] = 0; h->sweep_n[0] = h->sweep_n[1] = 0;
  return 0;
}

int bnmtf_nmf_sweep_times(bnmtf_nmf_t h, double* out4) {
  if (!h || !out4) FAIL("bad argument");
  out4[0] = h->sweep_ms[0]; out4[1] = h->sweep_ms[1]; out4[2] = (double)h->sweep_n[0]; out4[3] = (double)h->sweep_n[1];
  return 0;
}

// ---- distributions ---------------------------------------------------------------
namespace {
struct DevBuf {
  double* p = nullptr;
  ~DevBuf() { cudaFree(p); }
};
}  // namespace

#define DIST_PROLOGUE(NIN)                                                                          \
  if (n < 0) FAIL("negative n");                                                                    \
  if (n == 0) return 0;                                                                             \
  CK(cudaSetDevice(device));                                                                        \
  CK(pool_setup(device));                                                                           \
  g_stream = cudaStreamPerThread;                                                                   \
  DevBuf in[2], o[2];                                                                               \
  for (int q = 0; q < NIN; ++q) CK(cudaMalloc(&in[q].p, sizeof(double) * n));                       \
  for (int q = 0; q < 2; ++q) CK(cudaMalloc(&o[q].p, sizeof(double) * n));                          \
  const unsigned grid = (unsigned)((n + 255) / 256)

int bnmtf_tn_draw(int device, uint64_t seed, int64_t n, const double* mu, const double* tau, double* out) {
  DIST_PROLOGUE(2);
  CK(cudaMemcpy(in[0].p, mu, sizeof(double) * n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(in[1].p, tau, sizeof(double) * n, cudaMemcpyHostToDevice));
  api_tn_draw_kernel<<<grid, 256, 0, g_stream>>>(in[0].p, in[1].p, o[0].p, n, seed);
  CK(cudaStreamSynchronize(g_stream)); CK(cudaGetLastError());
  CK(cudaMemcpy(out, o[0].p, sizeof(double) * n, cudaMemcpyDeviceToHost));
  return 0;
}

int bnmtf_tn_moments(int device, int64_t n, const double* mu, const double* tau, double* exp_out, double* var_out) {
  DIST_PROLOGUE(2);
  CK(cudaMemcpy(in[0].p, mu, sizeof(double) * n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(in[1].p, tau, sizeof(double) * n, cudaMemcpyHostToDevice));
  api_tn_moments_kernel<<<grid, 256, 0, g_stream>>>(in[0].p, in[1].p, o[0].p, o[1].p, n);
  CK(cudaStreamSynchronize(g_stream)); CK(cudaGetLastError());
  if (exp_out) CK(cudaMemcpy(exp_out, o[0].p, sizeof(double) * n, cudaMemcpyDeviceToHost));
  if (var_out) CK(cudaMemcpy(var_out, o[1].p, sizeof(double) * n, cudaMemcpyDeviceToHost));
  return 0;
}

int bnmtf_gamma_draw(int device, uint64_t seed, int64_t n, const double* alpha, const double* beta, double* out) {
  DIST_PROLOGUE(2);
  CK(cudaMemcpy(in[0].p, alpha, sizeof(double) * n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(in[1].p, beta, sizeof(double) * n, cudaMemcpyHostToDevice));
  api_gamma_draw_kernel<<<grid, 256, 0, g_stream>>>(in[0].p, in[1].p, o[0].p, n, seed);
  CK(cudaStreamSynchronize(g_stream)); CK(cudaGetLastError());
  CK(cudaMemcpy(out, o[0].p, sizeof(double) * n, cudaMemcpyDeviceToHost));
  return 0;
}

int bnmtf_exp_draw(int device, uint64_t seed, int64_t n, const double* lambda, double* out) {
  DIST_PROLOGUE(1);
  CK(cudaMemcpy(in[0].p, lambda, sizeof(double) * n, cudaMemcpyHostToDevice));
  api_exp_draw_kernel<<<grid, 256, 0, g_stream>>>(in[0].p, o[0].p, n, seed);
  CK(cudaStreamSynchronize(g_stream)); CK(cudaGetLastError());
  CK(cudaMemcpy(out, o[0].p, sizeof(double) * n, cudaMemcpyDeviceToHost));
  return 0;
}

#include "engine_nmtf.inc"
#include "engine_summary.inc"
#include "engine_batch.inc"
