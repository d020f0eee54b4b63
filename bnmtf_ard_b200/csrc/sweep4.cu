// Launchers of the blocked cluster sweep (kernels_v4.cuh); a translation unit of its own so that the kernel builds in
// parallel with the rest of the engine.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "sweep4.h"
#include "kernels_v4.cuh"

namespace bnmtf {

// BNMTF_V4_WPC: experiments with the register / rows-in-flight balance (kernels_v4.cuh)
int sweep4_wpc() {
  static const int wpc = [] {
    const char* e = getenv("BNMTF_V4_WPC");
    const int v = e ? atoi(e) : 24;
    return (v == 20 || v == 22 || v == 24) ? v : 24;
  }();
  return wpc;
}

constexpr int V4_MAX_CLUSTERS = 64;   // upper bound of co-resident clusters (148 SMs x 2 CTAs / 8 = 37)
size_t sweep4_gmat_doubles(int rs, int K) { return (size_t)V4_MAX_CLUSTERS * 8 * rs * (size_t)(2 * ((K + 1) / 2)) * TN_PRE_DOUBLES; }

bool sweep4_fits(int tw, int rs, int K) {
  return sweep4_layout(tw, rs, K, 8, true, true, true).total * sizeof(double) <= (size_t)227 * 1024;
}

#define CK4(call)                                                                                     \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess) {                                                                          \
      char buf_[512];                                                                                 \
      snprintf(buf_, sizeof buf_, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      *err = buf_;                                                                                    \
      return 1;                                                                                       \
    }                                                                                                 \
  } while (0)

template <int NPT, int MODE, int WPC>
static int launch_w(const Sweep2Params& p, cudaStream_t stream, int* ncl_cache, std::string* err) {
  constexpr int CS = 8;
  if (p.rs != WPC / 2 - 1) { *err = "blocked cluster sweep: the side was packed for another number of rows per set"; return 1; }
  if (MODE == M_GIBBS && p.gmat == nullptr) { *err = "blocked cluster sweep: no random-material scratch"; return 1; }
  auto kern = sweep4_kernel<NPT, MODE, CS, WPC>;
  const bool want_mu = p.Xmu != nullptr;
  const size_t smem = sweep4_layout(p.tw, p.rs, p.K, CS, MODE == M_VB, want_mu, MODE == M_GIBBS).total * sizeof(double);
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3((p.rs + 1) * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (*ncl_cache == 0) {
    CK4(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    cfg.gridDim = dim3(CS * 64);
    int ncl = 0;
    CK4(cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg));
    if (ncl < 1) { *err = "blocked cluster sweep does not fit"; return 1; }
    if (getenv("BNMTF_V2_NCL")) ncl = std::max(1, atoi(getenv("BNMTF_V2_NCL")));
    ncl = std::min(ncl, V4_MAX_CLUSTERS);
    if (getenv("BNMTF_V2_DBG")) fprintf(stderr, "sweep4<%d,%d,%d>: %d clusters of %d x %d threads, %zu bytes of shared memory\n", NPT, MODE, WPC, ncl, CS, (p.rs + 1) * 32, smem);
    *ncl_cache = ncl;
  }
  const int ncl = std::min(*ncl_cache, p.nsets);
  if (ncl < 1) return 0;
  cfg.gridDim = dim3(CS * ncl);
  CK4(cudaLaunchKernelEx(&cfg, kern, p));
  return 0;
}

template <int NPT, int MODE>
static int launch_t(const Sweep2Params& p, cudaStream_t stream, int* ncl_cache, std::string* err) {
  if constexpr (NPT == 20) {
    switch (sweep4_wpc()) {
      case 20: return launch_w<NPT, MODE, 20>(p, stream, ncl_cache, err);
      case 22: return launch_w<NPT, MODE, 22>(p, stream, ncl_cache, err);
    }
    return launch_w<NPT, MODE, 24>(p, stream, ncl_cache, err);
  } else {
    return launch_w<NPT, MODE, Sweep2Cfg<NPT>::WPC>(p, stream, ncl_cache, err);   // 16 entries per lane and fewer: 32 warps per SM
  }
}

template <int MODE>
static int launch_m(int npt, const Sweep2Params& p, cudaStream_t stream, int* ncl_cache, std::string* err) {
  switch (npt) {
    case 4: return launch_t<4, MODE>(p, stream, ncl_cache, err);
    case 8: return launch_t<8, MODE>(p, stream, ncl_cache, err);
    case 12: return launch_t<12, MODE>(p, stream, ncl_cache, err);
    case 16: return launch_t<16, MODE>(p, stream, ncl_cache, err);
    case 20: return launch_t<20, MODE>(p, stream, ncl_cache, err);
  }
  *err = "unsupported entries per lane for the blocked cluster sweep";
  return 1;
}

int launch_sweep4(int npt, int method, const Sweep2Params& p, cudaStream_t stream, int* ncl_cache, std::string* err) {
  switch (method) {
    case 0: return launch_m<M_GIBBS>(npt, p, stream, ncl_cache, err);
    case 1: return launch_m<M_ICM>(npt, p, stream, ncl_cache, err);
    case 2: return launch_m<M_VB>(npt, p, stream, ncl_cache, err);
  }
  *err = "unknown method";
  return 1;
}

}  // namespace bnmtf
