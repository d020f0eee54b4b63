// Launchers of the per-row-chain cluster sweep (kernels_v5.cuh); a translation unit of its own so that the kernel builds
// in parallel with the rest of the engine.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "sweep5.h"
#include "kernels_v5.cuh"

namespace bnmtf {

// Geometry per entries-per-lane class.  20 / 16 entries: 11 row warps + producer, two CTAs per SM (80 registers);
// 12 and fewer: 15 row warps (64 registers).  BNMTF_V5_WIDE=1 (experiment, 20 entries): one CTA of 23 row warps per SM
// with a three-slot tile ring.
static bool v5_wide() {
  static const bool w = [] { const char* e = getenv("BNMTF_V5_WIDE"); return e && atoi(e) == 1; }();
  return w;
}
// BNMTF_V5_RS=9|10|11: rows per CTA at 20 entries per lane (96 / 88 / 80 registers per thread)
static int v5_rs20() {
  static const int v = [] { const char* e = getenv("BNMTF_V5_RS"); const int q = e ? atoi(e) : 11; return (q == 9 || q == 10 || q == 11) ? q : 11; }();
  return v;
}
int sweep5_rows_per_set(int npt) {
  if (npt == 20) return v5_wide() ? 23 : v5_rs20();
  if (npt == 16) return 11;
  if (npt == 4 || npt == 8 || npt == 12) return 15;
  return 0;
}
static int v5_nslot(int npt) { return (npt == 20 && v5_wide()) ? 3 : 2; }

constexpr int V5_MAX_CLUSTERS = 64;   // upper bound of co-resident clusters (148 SMs x 2 CTAs / 8 = 37)
size_t sweep5_gmat_doubles(int rs, int K) { return (size_t)V5_MAX_CLUSTERS * 8 * rs * (size_t)(2 * ((K + 1) / 2)) * TN_PRE_DOUBLES; }

bool sweep5_fits(int npt, int tw, int K, int K2, int covL) {
  const int rs = sweep5_rows_per_set(npt);
  if (rs == 0 || K < 2 || K > 64) return false;          // a lane keeps the row's values of columns lane and lane + 32
  const size_t per_cta = (npt == 20 && v5_wide()) ? (size_t)227 * 1024 : ((size_t)228 * 1024 - 2 * 1024) / 2;
#ifdef BNMTF_V2_PROF
  const bool vb = false;     // timing builds carry the phase clocks in shared memory: sized for the sampled methods only
#else
  const bool vb = true;
#endif
  if (K2 < 0 || K2 > 64 || covL < 0 || covL > 64) return false;
  return sweep5_layout(tw, rs, K, 8, v5_nslot(npt), 4, vb, K2, covL).total * sizeof(double) <= per_cta;
}

#define CK5(call)                                                                                     \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess) {                                                                          \
      char buf_[512];                                                                                 \
      snprintf(buf_, sizeof buf_, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      *err = buf_;                                                                                    \
      return 1;                                                                                       \
    }                                                                                                 \
  } while (0)

// cudaOccupancyMaxActiveClusters under-reports kernels that allocate tensor memory (15 clusters where 33 run, measured in
// round 1); the co-resident cluster count is asked of a stand-in with the same block size and shared memory instead.  A
// count that is too high would only cost a tail: clusters are independent, the ones that do not fit run afterwards.
template <int RS, int MINB>
__global__ void __launch_bounds__((RS + 1) * 32, MINB) sweep5_occupancy_probe(int* out) {
  extern __shared__ double probe_sm[];
  if (out && threadIdx.x == 0) out[blockIdx.x] = (int)probe_sm[0];
}

template <int NPT, int MODE, int RS, int NSLOT, int MINB, bool PROJ = false, bool COV = false>
static int launch_w(const Sweep2Params& p, cudaStream_t stream, int* ncl_cache, std::string* err) {
  constexpr int CS = 8;
  if (PROJ && (p.K2 < 1 || p.K2 > 64 || p.Yproj == nullptr || p.proj_out == nullptr)) { *err = "per-row cluster sweep: bad projection arguments"; return 1; }
  if (!PROJ && p.K2 != 0) { *err = "per-row cluster sweep: projection requested from a kernel without the epilogue"; return 1; }
  if (COV && (p.covL < 1 || p.covL > 64 || p.covA == nullptr || p.covS == nullptr)) { *err = "per-row cluster sweep: bad covariance arguments"; return 1; }
  if (!COV && p.covA != nullptr) { *err = "per-row cluster sweep: covariance term requested from a kernel without it"; return 1; }
  if (p.rs != RS) { *err = "per-row cluster sweep: the side was packed for another number of rows per set"; return 1; }
  if (MODE == M_GIBBS && p.gmat == nullptr) { *err = "per-row cluster sweep: no random-material scratch"; return 1; }
  auto kern = sweep5_kernel<NPT, MODE, CS, RS, NSLOT, MINB, PROJ, COV>;
  const size_t smem = sweep5_layout(p.tw, RS, p.K, CS, NSLOT, 4, MODE == M_VB, PROJ ? p.K2 : 0, COV ? p.covL : 0).total * sizeof(double);
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3((RS + 1) * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (*ncl_cache == 0) {
    CK5(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    cfg.gridDim = dim3(CS * V5_MAX_CLUSTERS);
    int ncl = 0;
    auto probe = sweep5_occupancy_probe<RS, MINB>;
    CK5(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK5(cudaOccupancyMaxActiveClusters(&ncl, probe, &cfg));
    if (ncl < 1) { *err = "per-row cluster sweep does not fit"; return 1; }
    if (getenv("BNMTF_V2_NCL")) ncl = std::max(1, atoi(getenv("BNMTF_V2_NCL")));
    ncl = std::min(ncl, V5_MAX_CLUSTERS);
    if (getenv("BNMTF_V2_DBG")) fprintf(stderr, "sweep5<%d,%d,%d,%d,%d>: %d clusters of %d x %d threads, %zu bytes of shared memory\n", NPT, MODE, RS, NSLOT, MINB, ncl, CS, (RS + 1) * 32, smem);
    *ncl_cache = ncl;
  }
  const int ncl = std::min(*ncl_cache, p.nsets);
  if (ncl < 1) return 0;
  cfg.gridDim = dim3(CS * ncl);
  CK5(cudaLaunchKernelEx(&cfg, kern, p));
  return 0;
}

// the tri-factorisation sweeps: projection epilogue (F sweep) and / or the VB covariance term
template <int MODE, bool PROJ, bool COV>
static int launch_p(int npt, const Sweep2Params& p, cudaStream_t stream, int* ncl_cache, std::string* err) {
  switch (npt) {
    case 4: return launch_w<4, MODE, 15, 2, 2, PROJ, COV>(p, stream, ncl_cache, err);
    case 8: return launch_w<8, MODE, 15, 2, 2, PROJ, COV>(p, stream, ncl_cache, err);
    case 12: return launch_w<12, MODE, 15, 2, 2, PROJ, COV>(p, stream, ncl_cache, err);
    case 16: return launch_w<16, MODE, 11, 2, 2, PROJ, COV>(p, stream, ncl_cache, err);
    case 20:
      if (v5_wide() || v5_rs20() != 11) break;
      return launch_w<20, MODE, 11, 2, 2, PROJ, COV>(p, stream, ncl_cache, err);
  }
  *err = "unsupported entries per lane for the per-row cluster sweep of the tri-factorisation";
  return 1;
}

template <int MODE>
static int launch_m(int npt, const Sweep2Params& p, cudaStream_t stream, int* ncl_cache, std::string* err) {
  switch (npt) {
    case 4: return launch_w<4, MODE, 15, 2, 2>(p, stream, ncl_cache, err);
    case 8: return launch_w<8, MODE, 15, 2, 2>(p, stream, ncl_cache, err);
    case 12: return launch_w<12, MODE, 15, 2, 2>(p, stream, ncl_cache, err);
    case 16: return launch_w<16, MODE, 11, 2, 2>(p, stream, ncl_cache, err);
    case 20:
      if (v5_wide()) return launch_w<20, MODE, 23, 3, 1>(p, stream, ncl_cache, err);
      if (v5_rs20() == 10) return launch_w<20, MODE, 10, 2, 2>(p, stream, ncl_cache, err);
      if (v5_rs20() == 9) return launch_w<20, MODE, 9, 2, 2>(p, stream, ncl_cache, err);
      return launch_w<20, MODE, 11, 2, 2>(p, stream, ncl_cache, err);
  }
  *err = "unsupported entries per lane for the per-row cluster sweep";
  return 1;
}

int launch_peer_barrier(int* const* peer_flags_dev, int* my_flags, int rank, int world, int epoch, cudaStream_t stream, std::string* err) {
  peer_barrier_kernel<<<1, 32, 0, stream>>>(peer_flags_dev, my_flags, rank, world, epoch);
  CK5(cudaGetLastError());
  return 0;
}

int launch_sweep5(int npt, int method, const Sweep2Params& p, cudaStream_t stream, int* ncl_cache, std::string* err) {
  if (p.K2 > 0 && method != 2) {
    if (method == 0) return launch_p<M_GIBBS, true, false>(npt, p, stream, ncl_cache, err);
    if (method == 1) return launch_p<M_ICM, true, false>(npt, p, stream, ncl_cache, err);
  }
  if (method == 2 && p.covA != nullptr)
    return p.K2 > 0 ? launch_p<M_VB, true, true>(npt, p, stream, ncl_cache, err) : launch_p<M_VB, false, true>(npt, p, stream, ncl_cache, err);
  switch (method) {
    case 0: return launch_m<M_GIBBS>(npt, p, stream, ncl_cache, err);
    case 1: return launch_m<M_ICM>(npt, p, stream, ncl_cache, err);
    case 2: return launch_m<M_VB>(npt, p, stream, ncl_cache, err);
  }
  *err = "unknown method";
  return 1;
}

}  // namespace bnmtf
