// Micro-benchmark: tensor memory (TMEM) as a per-warp scratch area for fp64 residual segments
// (tcgen05.st / tcgen05.ld round trips, no MMA).  Not part of the product; it answers whether the
// sweep kernels could keep their residuals in TMEM instead of registers.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tmem tmem.cu && ./tmem
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void tm_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tm_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// NCH chunks of 8 columns (= 4 doubles per lane) per warp; COLS = columns allocated per CTA
template <int NCH, int COLS>
__global__ void __launch_bounds__(384, 1) k_tmem(int iters, int with_lds, uint32_t* out, long long* clk, int* bad) {
  __shared__ uint32_t base_sh;
  __shared__ double tile[2048];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int q = threadIdx.x; q < 2048; q += blockDim.x) tile[q] = 1.0;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&base_sh)), "n"(COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = base_sh;
  // warp w may touch lanes 32*(w%4)..+31 only; warps sharing a lane quarter take different columns
  const uint32_t taddr = base + ((uint32_t)(warp & 3) * 32u << 16) + (uint32_t)(warp >> 2) * (NCH * 8);
  uint32_t r[8];
  for (int c = 0; c < NCH; ++c) {
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = (uint32_t)(threadIdx.x * 1000 + c * 8 + i);
    tm_st8(taddr + c * 8, r);
  }
  tm_wait_st();
  __syncthreads();
  double acc = 0.0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (with_lds & 2) {   // all loads of the segment in flight, one wait
      uint32_t rr[NCH][8];
#pragma unroll
      for (int c = 0; c < NCH; ++c) tm_ld8(taddr + c * 8, rr[c]);
      tm_wait_ld();
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
#pragma unroll
        for (int i = 0; i < 8; ++i) rr[c][i] += 1u;
        if (with_lds & 1) {
#pragma unroll
          for (int i = 0; i < 4; ++i) acc += tile[(rr[c][2 * i] * 17u + lane) & 2047];
        }
        tm_st8(taddr + c * 8, rr[c]);
      }
    } else {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        tm_ld8(taddr + c * 8, r);
        tm_wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] += 1u;
        if (with_lds & 1) {
#pragma unroll
          for (int i = 0; i < 4; ++i) acc += tile[(r[2 * i] * 17u + lane) & 2047];
        }
        tm_st8(taddr + c * 8, r);
      }
    }
    tm_wait_st();
  }
  const long long t1 = clock64();
  int nbad = 0;
  for (int c = 0; c < NCH; ++c) {
    tm_ld8(taddr + c * 8, r);
    tm_wait_ld();
#pragma unroll
    for (int i = 0; i < 8; ++i) nbad += (r[i] != (uint32_t)(threadIdx.x * 1000 + c * 8 + i) + (uint32_t)iters);
  }
  if (nbad) atomicAdd(bad, nbad);
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
  if (acc == 123.456) out[0] = 1;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(COLS));
}

template <int NCH, int COLS>
int run(int blocks_per_sm_note, int with_lds) {
  uint32_t* out; long long* clk; int* bad;
  CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&clk, 64)); CK(cudaMalloc(&bad, 4));
  CK(cudaMemset(bad, 0, 4));
  const int iters = 2000;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_tmem<NCH, COLS><<<148, 384>>>(10, with_lds, out, clk, bad);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  k_tmem<NCH, COLS><<<148, 384>>>(iters, with_lds, out, clk, bad);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  long long c; int b;
  CK(cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&b, bad, 4, cudaMemcpyDeviceToHost));
  const double bytes_per_sm_iter = 12.0 * 32 * NCH * 8 * 4;   // one direction
  printf("NCH=%d (%d doubles/lane) lds=%d: %.3f ms, %.1f cycles/iteration (12 warps), TMEM %.1f B/cycle/SM each way, mismatches %d\n",
         NCH, NCH * 4, with_lds, ms, (double)c / iters, bytes_per_sm_iter / ((double)c / iters), b);
  return 0;
}

int main() {
  if (run<5, 128>(1, 0)) return 1;    // 40 columns per warp = 20 doubles per lane (one residual segment)
  if (run<5, 128>(1, 2)) return 1;    // same, all loads in flight
  if (run<10, 256>(1, 0)) return 1;   // two segments
  if (run<10, 256>(1, 2)) return 1;
  if (run<5, 128>(1, 1)) return 1;    // with shared-memory gathers in between
  if (run<5, 128>(1, 3)) return 1;
  return 0;
}
