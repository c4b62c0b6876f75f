// Latency / single-warp issue-rate microbenchmarks that shaped k_block (profiles/r1_notes.md):
//   * FP64 FMA issue interval from one warp with 32 independent accumulators, and dependent latency
//   * redux.sync, shfl, vote, LDS broadcast, STS->LDS round trip, bar.sync (128 threads)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu && ./lat
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma_indep(double* out, long long* cyc, int iters) {
   double a[32];
   for (int i = 0; i < 32; ++i) a[i] = threadIdx.x + i;
   double l = out[0], v = out[1];
   long long t0 = clock64();
   for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 32; ++i) a[i] = fma(-l, v, a[i]);
   }
   long long t1 = clock64();
   double s = 0;
   for (int i = 0; i < 32; ++i) s += a[i];
   out[2 + threadIdx.x + blockDim.x * blockIdx.x] = s;
   if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dfma_dep(double* out, long long* cyc, int iters) {
   double a = threadIdx.x, l = out[0], v = out[1];
   long long t0 = clock64();
   for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 32; ++i) a = fma(-l, a, v);
   }
   long long t1 = clock64();
   out[2 + threadIdx.x] = a;
   if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_rcp_dep(double* out, long long* cyc, int iters) {
   double a = out[0] + threadIdx.x + 1.5;
   long long t0 = clock64();
   for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a = 1.0 / a + 1.25;
   }
   long long t1 = clock64();
   out[2 + threadIdx.x] = a;
   if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_redux(unsigned* out, long long* cyc, int iters) {
   unsigned x = threadIdx.x * 2654435761u + out[0];
   long long t0 = clock64();
   for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x = __reduce_max_sync(0xffffffffu, x + threadIdx.x) ^ (x >> 3);
   }
   long long t1 = clock64();
   out[1 + threadIdx.x] = x;
   if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl(double* out, long long* cyc, int iters) {
   double x = out[0] + threadIdx.x;
   int q = (int)out[1];
   long long t0 = clock64();
   for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { x = __shfl_sync(0xffffffffu, x, q) + 1.0; q = (q + 7) & 31; }
   }
   long long t1 = clock64();
   out[2 + threadIdx.x] = x;
   if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_vote(unsigned* out, long long* cyc, int iters) {
   unsigned x = threadIdx.x + out[0];
   long long t0 = clock64();
   for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x += __any_sync(0xffffffffu, (x & 31u) == 77u) ? 3u : 1u;
   }
   long long t1 = clock64();
   out[1 + threadIdx.x] = x;
   if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_sts_lds(double* out, long long* cyc, int iters) {
   __shared__ double s[64];
   double x = out[0] + threadIdx.x;
   int q = 0;
   long long t0 = clock64();
   for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
         if ((int)threadIdx.x == q) s[q] = x;
         __syncwarp();
         x += s[q];
         __syncwarp();
         q = (q + 5) & 31;
      }
   }
   long long t1 = clock64();
   out[2 + threadIdx.x] = x;
   if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_bar128(double* out, long long* cyc, int iters) {
   __shared__ double s[128];
   double x = out[0] + threadIdx.x;
   long long t0 = clock64();
   for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
         s[threadIdx.x] = x;
         __syncthreads();
         x += s[(threadIdx.x + 37) & 127];
         __syncthreads();
      }
   }
   long long t1 = clock64();
   out[2 + threadIdx.x] = x;
   if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
   double* d; long long* c; unsigned* u;
   cudaMalloc(&d, 1 << 20); cudaMalloc(&c, 64); cudaMalloc(&u, 1 << 16);
   double h[2] = {1e-9, 0.5};
   cudaMemcpy(d, h, 16, cudaMemcpyHostToDevice);
   cudaMemset(u, 0, 1 << 16);
   long long cy; const int it = 2000;
   for (int warps = 1; warps <= 8; warps *= 2) {
      k_dfma_indep<<<1, 32 * warps>>>(d, c, it); cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
      printf("DFMA independent x32, %d warp(s) on one SM: %.2f cycles per DFMA per warp\n", warps, (double)cy / (32.0 * it));
   }
   cudaMemcpy(d, h, 16, cudaMemcpyHostToDevice);
   k_dfma_dep<<<1, 32>>>(d, c, it); cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
   printf("DFMA dependent chain: %.2f cycles\n", (double)cy / (32.0 * it));
   cudaMemcpy(d, h, 16, cudaMemcpyHostToDevice);
   k_rcp_dep<<<1, 32>>>(d, c, it); cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
   printf("1.0/x + c dependent: %.2f cycles\n", (double)cy / (8.0 * it));
   k_redux<<<1, 32>>>(u, c, it); cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
   printf("redux.sync.max dependent (+2 alu): %.2f cycles\n", (double)cy / (8.0 * it));
   h[1] = 3; cudaMemcpy(d, h, 16, cudaMemcpyHostToDevice);
   k_shfl<<<1, 32>>>(d, c, it); cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
   printf("shfl.idx f64 + DADD dependent: %.2f cycles\n", (double)cy / (8.0 * it));
   k_vote<<<1, 32>>>(u, c, it); cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
   printf("vote.any + alu dependent: %.2f cycles\n", (double)cy / (8.0 * it));
   k_sts_lds<<<1, 32>>>(d, c, it); cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
   printf("STS (one lane) -> syncwarp -> LDS broadcast -> DADD -> syncwarp: %.2f cycles\n", (double)cy / (8.0 * it));
   k_bar128<<<1, 128>>>(d, c, it); cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
   printf("STS, bar.sync(128), LDS, DADD, bar.sync: %.2f cycles\n", (double)cy / (8.0 * it));
   printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
   return 0;
}
