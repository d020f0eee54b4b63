// Host interface of the per-row-chain cluster sweep (kernels_v5.cuh), compiled in its own translation unit (sweep5.cu).
#pragma once
#include <string>
#include <cuda_runtime.h>
#include "kernels_v2.cuh"

namespace bnmtf {
// rows per set (= row warps per CTA) the kernel runs with for `npt` entries per lane; 0 when npt is not served
int sweep5_rows_per_set(int npt);
// true when the kernel serves K columns and its shared memory (worst case: VB) fits one CTA; K2 > 0: also with the
// projection epilogue onto K2 columns of a second factor (tri-factorisation F sweep)
// covL > 0: also with the covariance area of the tri-factorisation VB sweeps
bool sweep5_fits(int npt, int tw, int K, int K2 = 0, int covL = 0);
// doubles of random-material scratch a Gibbs launch needs (Sweep2Params::gmat), for the largest grid
size_t sweep5_gmat_doubles(int rs, int K);
// Launch one sweep on `stream`; p.rs must be sweep5_rows_per_set(npt) as chosen when the side was packed.  method:
// BNMTF_GIBBS / ICM / VB (0 / 1 / 2).  p.K2 > 0: p.Yproj is a block-major second factor [ceil(K2/2)][ldp][2] and
// the final residual is projected onto it (p.proj_out, indexed by global row); p.covA != nullptr (VB): the covariance
// term of the tri-factorisation updates.  *ncl_cache (0 on first use) caches the number of co-resident clusters.
// Returns 0, or 1 with *err set.
int launch_sweep5(int npt, int method, const Sweep2Params& p, cudaStream_t stream, int* ncl_cache, std::string* err);
// device-side barrier of the ranks behind a sweep that wrote its rows to the peers (kernels_v5.cuh: peer_barrier_kernel)
int launch_peer_barrier(int* const* peer_flags_dev, int* my_flags, int rank, int world, int epoch, cudaStream_t stream, std::string* err);
}  // namespace bnmtf
