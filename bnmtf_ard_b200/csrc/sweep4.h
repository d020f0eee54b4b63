// Host interface of the blocked cluster sweep (kernels_v4.cuh), compiled in its own translation unit (sweep4.cu).
#pragma once
#include <string>
#include <cuda_runtime.h>
#include "kernels_v2.cuh"

namespace bnmtf {
// warps per SM of the kernel (two CTAs of wpc / 2 warps; wpc / 2 - 1 rows per CTA): 24, 22 or 20
int sweep4_wpc();
// true when the kernel's shared memory (worst case: VB with parameter outputs) fits one CTA
bool sweep4_fits(int tw, int rs, int K);
// doubles of random-material scratch a Gibbs launch needs (Sweep2Params::gmat), for the largest grid
size_t sweep4_gmat_doubles(int rs, int K);
// Launch one sweep on `stream`; p.rs must be sweep4_wpc() / 2 - 1 as chosen when the side was packed.  method:
// BNMTF_GIBBS / ICM / VB (0 / 1 / 2).  *ncl_cache (0 on first use) caches the number of co-resident clusters.
// Returns 0, or 1 with *err set.
int launch_sweep4(int npt, int method, const Sweep2Params& p, cudaStream_t stream, int* ncl_cache, std::string* err);
}  // namespace bnmtf
