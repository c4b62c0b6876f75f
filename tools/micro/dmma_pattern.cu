// Microbenchmark: DMMA m8n8k4 issue patterns as used by the Schur-update kernel.
//  v1: 4x4 register outer product (af[4], bf[4], acc[4][4]) with operands in registers only
//  v2: same, with the 8 fragment loads per k-step from shared memory (LDS.64), double buffered
//  v3: 2x8 pattern   v4: 4x8 pattern (acc 64 doubles)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int MI, int NJ>
__global__ void k_reg(double* out, int iters) {
   double acc[MI][NJ][2], af[MI], bf[NJ];
   for (int i = 0; i < MI; ++i) for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0;
   for (int i = 0; i < MI; ++i) af[i] = threadIdx.x * 1e-3 + i;
   for (int j = 0; j < NJ; ++j) bf[j] = 1.0 + threadIdx.x * 1e-4 + j;
   for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
         for (int j = 0; j < NJ; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
#pragma unroll
      for (int i = 0; i < MI; ++i) af[i] += 1e-9;   // keep operands changing
   }
   double s = 0;
   for (int i = 0; i < MI; ++i) for (int j = 0; j < NJ; ++j) s += acc[i][j][0] + acc[i][j][1];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MI, int NJ>
__global__ void k_smem(double* out, int iters) {
   extern __shared__ double sm[];
   const int LDT = 132;
   for (int i = threadIdx.x; i < 2 * 32 * LDT; i += blockDim.x) sm[i] = i * 1e-6;
   __syncthreads();
   const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t4 = lane & 3;
   const double* as = sm + (warp & 3) * 32 % (128 - 8 * MI + 1) + g;
   const double* bs = sm + 32 * LDT + (warp >> 2) * 32 % (128 - 8 * NJ + 1) + g;
   double acc[MI][NJ][2];
   for (int i = 0; i < MI; ++i) for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0;
   for (int it = 0; it < iters; ++it) {
      double af[2][MI], bf[2][NJ];
#pragma unroll
      for (int i = 0; i < MI; ++i) af[0][i] = as[t4 * LDT + i * 8];
#pragma unroll
      for (int j = 0; j < NJ; ++j) bf[0][j] = -bs[t4 * LDT + j * 8];
#pragma unroll
      for (int ks = 0; ks < 32; ks += 4) {
         const int cur = (ks >> 2) & 1, nxt = cur ^ 1;
         if (ks + 4 < 32) {
#pragma unroll
            for (int i = 0; i < MI; ++i) af[nxt][i] = as[(ks + 4 + t4) * LDT + i * 8];
#pragma unroll
            for (int j = 0; j < NJ; ++j) bf[nxt][j] = -bs[(ks + 4 + t4) * LDT + j * 8];
         }
#pragma unroll
         for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) dmma(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
      }
   }
   double s = 0;
   for (int i = 0; i < MI; ++i) for (int j = 0; j < NJ; ++j) s += acc[i][j][0] + acc[i][j][1];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> double timeit(F f) {
   cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
   f(); cudaDeviceSynchronize();
   cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
   float ms; cudaEventElapsedTime(&ms, e0, e1); return ms * 1e-3;
}
int main() {
   double* out; cudaMalloc(&out, sizeof(double) * 148 * 1024);
   const int blocks = 148;
   cudaFuncSetAttribute(k_smem<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
   cudaFuncSetAttribute(k_smem<4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
   cudaFuncSetAttribute(k_smem<2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
   for (int warps : {4, 8, 16}) {
      int threads = warps * 32, iters = 4000;
      double t;
      t = timeit([&] { k_reg<4, 4><<<blocks, threads>>>(out, iters); });
      printf("warps %2d reg 4x4: %.2f TF/s", warps, 2.0 * 256 * 16 * iters * (double)warps * blocks / t / 1e12);
      t = timeit([&] { k_reg<4, 8><<<blocks, threads>>>(out, iters); });
      printf("  reg 4x8: %.2f", 2.0 * 256 * 32 * iters * (double)warps * blocks / t / 1e12);
      t = timeit([&] { k_reg<2, 8><<<blocks, threads>>>(out, iters); });
      printf("  reg 2x8: %.2f", 2.0 * 256 * 16 * iters * (double)warps * blocks / t / 1e12);
      int it2 = 500;
      t = timeit([&] { k_smem<4, 4><<<blocks, threads, 68000>>>(out, it2); });
      printf(" | smem 4x4: %.2f", 2.0 * 256 * 16 * 8 * it2 * (double)warps * blocks / t / 1e12);
      if (warps <= 8) {
         t = timeit([&] { k_smem<4, 8><<<blocks, threads, 68000>>>(out, it2); });
         printf("  smem 4x8: %.2f", 2.0 * 256 * 32 * 8 * it2 * (double)warps * blocks / t / 1e12);
      }
      t = timeit([&] { k_smem<2, 8><<<blocks, threads, 68000>>>(out, it2); });
      printf("  smem 2x8: %.2f\n", 2.0 * 256 * 16 * 8 * it2 * (double)warps * blocks / t / 1e12);
   }
   return 0;
}
