// Latency micro-benchmarks used to reason about the sweep kernels' dependency chain (not part of the product).
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__global__ void k_dfma(double* out, long long* t, double x, double y) {
  double a = x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) a = fma(a, y, x);
  }
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) t[0] = (t1 - t0) / 1024;
}
__global__ void k_rsqrt(double* out, long long* t, double x) {
  double a = x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) a = rsqrt(a) + x;
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) t[1] = (t1 - t0) / 256;
}
__global__ void k_special(double* out, long long* t, double x) {
  double a = x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) a = erfc(a) * 0.5 + 0.25;
  long long t1 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) a = normcdfinv(a * 0.5 + 0.25) * 0.1 + 0.5;
  long long t2 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) a = log(a * 0.5 + 1.25);
  long long t3 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) a = sqrt(a + 1.5);
  long long t4 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) a = 1.5 / (a + 1.5);
  long long t5 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) { t[2] = (t1 - t0) / 64; t[3] = (t2 - t1) / 64; t[4] = (t3 - t2) / 64; t[5] = (t4 - t3) / 64; t[6] = (t5 - t4) / 64; }
}
__global__ void k_lds(double* out, long long* t) {
  __shared__ double sh[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = (double)((i * 7 + 1) % 1024);
  __syncthreads();
  double a = sh[threadIdx.x];
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) a = sh[((int)a) & 1023];
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) t[7] = (t1 - t0) / 256;
}
__global__ void __cluster_dims__(8, 1, 1) k_cluster(double* out, long long* t) {
  __shared__ double sh[64];
  cg::cluster_group cl = cg::this_cluster();
  cl.sync();
  long long t0 = clock64();
  for (int i = 0; i < 64; ++i) {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  long long t1 = clock64();
  double* remote = cl.map_shared_rank(sh, (cl.block_rank() + 1) % 8);
  for (int i = 0; i < 64; ++i) {
    if (threadIdx.x < 8) remote[threadIdx.x] = (double)i;
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  long long t2 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = sh[threadIdx.x & 7];
  if (threadIdx.x == 0 && blockIdx.x == 0) { t[8] = (t1 - t0) / 64; t[9] = (t2 - t1) / 64; }
  cl.sync();
}
int main() {
  double* out; long long* t;
  cudaMalloc(&out, 8 * 1024 * sizeof(double)); cudaMallocManaged(&t, 16 * sizeof(long long));
  for (int i = 0; i < 16; ++i) t[i] = 0;
  k_dfma<<<1, 32>>>(out, t, 0.999, 1.0000001);
  k_rsqrt<<<1, 32>>>(out, t, 1.3);
  k_special<<<1, 32>>>(out, t, 0.3);
  k_lds<<<1, 32>>>(out, t);
  k_cluster<<<8, 128>>>(out, t);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  printf("cycles: dfma %lld rsqrt+add %lld erfc %lld normcdfinv %lld log %lld sqrt %lld div %lld lds-dependent %lld cluster-barrier %lld dsmem+barrier %lld\n",
         t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8], t[9]);
  return 0;
}
