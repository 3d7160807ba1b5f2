// Latency microbenchmarks on sm_100a: dependent DFMA / rcp / rsqrt / sqrt / div chains, LDS round trip,
// __syncthreads, cluster.sync.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 lat.cu -o lat
#include <cstdio>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#define N 1024
__global__ void k_dfma(double* out, long long* t, double a, double b) {
  double x = a;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, b, a);
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) t[0] = t1 - t0;
}
__global__ void k_rcp(double* out, long long* t, double a) {
  double x = a; long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) x = __drcp_rn(x) + 1.0;
  long long t1 = clock64(); out[threadIdx.x] = x; if (threadIdx.x == 0) t[1] = t1 - t0;
}
__global__ void k_rsqrt(double* out, long long* t, double a) {
  double x = a; long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) x = rsqrt(x) + 1.0;
  long long t1 = clock64(); out[threadIdx.x] = x; if (threadIdx.x == 0) t[2] = t1 - t0;
}
__global__ void k_div(double* out, long long* t, double a) {
  double x = a; long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) x = 1.0 / x + 1.0;
  long long t1 = clock64(); out[threadIdx.x] = x; if (threadIdx.x == 0) t[3] = t1 - t0;
}
__global__ void k_lds(double* out, long long* t) {
  __shared__ int idx[256];
  idx[threadIdx.x] = (threadIdx.x + 1) & 255; __syncthreads();
  int j = threadIdx.x; long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) j = idx[j];
  long long t1 = clock64(); out[threadIdx.x] = j; if (threadIdx.x == 0) t[4] = t1 - t0;
}
__global__ void k_sync(double* out, long long* t, int slot) {
  long long t0 = clock64();
  for (int i = 0; i < N; ++i) __syncthreads();
  long long t1 = clock64(); if (threadIdx.x == 0) t[slot] = t1 - t0;
}
__global__ void k_sts_sync(double* out, long long* t, int slot) {  // sts + sync + lds dependent round
  __shared__ double buf[512];
  double x = threadIdx.x; long long t0 = clock64();
  for (int i = 0; i < N; ++i) { buf[threadIdx.x] = x; __syncthreads(); x = buf[(threadIdx.x + 33) % blockDim.x] + 1.0; __syncthreads(); }
  long long t1 = clock64(); out[threadIdx.x] = x; if (threadIdx.x == 0) t[slot] = t1 - t0;
}
__global__ void __cluster_dims__(4, 1, 1) k_csync(long long* t, int slot) {
  cg::cluster_group c = cg::this_cluster();
  long long t0 = clock64();
  for (int i = 0; i < 256; ++i) c.sync();
  long long t1 = clock64(); if (threadIdx.x == 0 && blockIdx.x == 0) t[slot] = t1 - t0;
}
__global__ void k_empty() {}
int main() {
  double* out; long long* t; cudaMalloc(&out, 4096 * 8); cudaMallocManaged(&t, 64 * 8);
  k_dfma<<<1, 32>>>(out, t, 1.0000001, 0.9999999); k_rcp<<<1, 32>>>(out, t, 1.5); k_rsqrt<<<1, 32>>>(out, t, 1.5); k_div<<<1, 32>>>(out, t, 1.5);
  k_lds<<<1, 256>>>(out, t); k_sync<<<1, 256>>>(out, t, 5); k_sync<<<1, 512>>>(out, t, 6); k_sts_sync<<<1, 256>>>(out, t, 7); k_sts_sync<<<1, 512>>>(out, t, 8);
  k_csync<<<4, 256>>>(t, 9);
  cudaDeviceSynchronize();
  printf("dfma dep %.1f cyc | drcp+add %.1f | rsqrt+add %.1f | div+add %.1f | lds dep %.1f | sync256 %.1f | sync512 %.1f | sts-sync-lds-sync 256: %.1f 512: %.1f | cluster.sync(4) %.1f\n",
         t[0] / (double)N, t[1] / (double)N, t[2] / (double)N, t[3] / (double)N, t[4] / (double)N, t[5] / (double)N, t[6] / (double)N, t[7] / (double)N, t[8] / (double)N, t[9] / 256.0);
  // back-to-back launch cost of empty kernels (device-side gap)
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 100; ++i) k_empty<<<128, 512>>>();
  cudaEventRecord(e0); for (int i = 0; i < 2000; ++i) k_empty<<<128, 512>>>(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); printf("empty kernel launch-to-launch %.2f us\n", ms * 1000 / 2000);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
