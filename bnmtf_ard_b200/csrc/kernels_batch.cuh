// Batched launches for many small models (BASELINE configs[4]: the 10 folds x
// K-range x methods of a cross-validation sweep run concurrently, one launch per update).
//
// Every kernel here is the body of the single-model kernel of kernels.cuh with blockIdx.y (or
// blockIdx.z) selecting the model: the device code that touches a model's data is the same, so a
// model run in a batch follows bit for bit the chain it follows alone.  The per-model argument
// blocks live in device memory, written once per run (the iteration index is read from the
// model's device-side counter, as in the graph replays of the single-model run loop).
#pragma once
#include "kernels.cuh"
#include "kernels_summary.cuh"

namespace bnmtf {

struct LambdaArgs {
  const double* Ukm; int64_t ldu, I;
  const double* Vkm; int64_t ldv, J;
  int K, update;
  Hyper h;
  double *colsum, *lamstate, *lamk;
  uint64_t seed;
  const uint32_t* iter_ptr;
};

struct FinalizeArgs {
  const double* rsU; int64_t nU;
  const double* rsV; int64_t nV;
  int K, update;
  Hyper h;
  const double *colsum, *lamstate;
  double *scal, *stats_out;
  uint64_t seed;
  int64_t I, J;
  uint32_t* iter_ctr;
};

struct KmArgs {
  const double *Xval, *Xvar;
  double* Ykm;
  int64_t n, ld;
  int K;
};

// running sums of the draws over range(burn_in, iterations, thinning) (engine_summary.inc); launched after the
// finalize kernel has advanced the iteration counter, so the iteration that just ended is *iter_ptr - 1
struct SumBatchArgs {
  SumArgs a;
  const uint32_t* iter_ptr;
  const double* stats_base;   // [iterations][16]; tau of iteration it is stats_base[it * 16 + tau_slot]
  int tau_seg, tau_slot;      // segment of `a` whose source moves with the iteration (-1: none)
  int burn, thin;
};

template <class T>
__device__ __forceinline__ void load_args(T& dst, const T* __restrict__ src) {
  static_assert(sizeof(T) % 4 == 0, "argument blocks are copied word by word");
  for (int i = threadIdx.x + blockDim.x * threadIdx.y; i < (int)(sizeof(T) / 4); i += blockDim.x * blockDim.y)
    reinterpret_cast<uint32_t*>(&dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
  __syncthreads();
}

// grid (max blocks per model, models); blocks beyond a model's share leave at once.
// (min blocks = 1: without it ptxas settles on 64 registers for some instantiations and spills the row.)
template <int NPT, int MODE>
__global__ void __launch_bounds__(SweepCfg<NPT>::MAXT, 1) sweep_batch_kernel(const SweepParams* __restrict__ pp) {
  __shared__ SweepParams sp;
  const SweepParams* g = pp + blockIdx.y;
  const int teams = blockDim.x / g->tpr;
  const int need = (g->n_loc + teams - 1) / teams;
  if ((int)blockIdx.x >= need) return;
  load_args(sp, g);
  sweep_body<NPT, MODE>(sp, (int)blockIdx.x, need);
}

// grid (max K, models), 256 threads
template <int MODE>
__global__ void lambda_batch_kernel(const LambdaArgs* __restrict__ aa) {
  __shared__ LambdaArgs a;
  if ((int)blockIdx.x >= aa[blockIdx.y].K) return;
  load_args(a, aa + blockIdx.y);
  lambda_body<MODE>(a.Ukm, a.ldu, a.I, a.Vkm, a.ldv, a.J, a.K, a.h, a.colsum, a.lamstate, a.lamk, a.seed, 0u, a.update,
                    a.iter_ptr, (int)blockIdx.x);
}

// grid (models), 1024 threads
template <int MODE>
__global__ void __launch_bounds__(1024) finalize_batch_kernel(const FinalizeArgs* __restrict__ aa) {
  __shared__ FinalizeArgs a;
  load_args(a, aa + blockIdx.x);
  finalize_body<MODE>(a.rsU, a.nU, a.rsV, a.nV, a.K, a.h, a.colsum, a.lamstate, a.scal, a.stats_out, a.seed, 0u, a.update,
                      a.I, a.J, a.iter_ctr);
}

// grid (max ld / 32, max K / 32, models), block (32, 8)
template <int NV>
__global__ void rm_to_km_batch_kernel(const KmArgs* __restrict__ aa) {
  __shared__ KmArgs a;
  const KmArgs* g = aa + blockIdx.z;
  if ((int64_t)blockIdx.x * 32 >= g->ld || (int)blockIdx.y * 32 >= g->K) return;
  load_args(a, g);
  rm_to_km_body<NV>(a.Xval, a.Xvar, a.Ykm, a.n, a.K, a.ld, (int)blockIdx.x, (int)blockIdx.y);
}

// grid (blocks per model, models), 256 threads
static __global__ void summary_accum_batch_kernel(const SumBatchArgs* __restrict__ aa) {
  const SumBatchArgs& g = aa[blockIdx.y];
  const int it = (int)*g.iter_ptr - 1;
  if (g.thin <= 0 || it < g.burn || (it - g.burn) % g.thin != 0) return;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int s = 0; s < g.a.nseg; ++s) {
    double* __restrict__ d = g.a.seg[s].dst;
    const double* __restrict__ x = g.a.seg[s].src;
    if (s == g.tau_seg) x = g.stats_base + (size_t)it * 16 + g.tau_slot;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < g.a.seg[s].n; i += stride) d[i] += x[i];
  }
}

}  // namespace bnmtf
