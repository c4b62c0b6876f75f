// Microbenchmark: FP64 throughput of DFMA vs DMMA (mma.sync f64 shapes) on B200, per warps/SM.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double* out, int iters) {
   double a[16];
   for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
   double b = 1.000001, c = 1e-9;
   for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
   }
   double s = 0;
   for (int i = 0; i < 16; ++i) s += a[i];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma884(double* out, int iters) {
   double acc[16][2];
   for (int i = 0; i < 16; ++i) { acc[i][0] = 0; acc[i][1] = 0; }
   double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
   for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i)
         asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                      : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a), "d"(b));
   }
   double s = 0;
   for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma16816(double* out, int iters) {
   double acc[8][4];
   for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = 0;
   double a[8], b[4];
   for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
   for (int i = 0; i < 4; ++i) b[i] = 1.0 + threadIdx.x * 1e-4 + i;
   for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
         asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                      : "+d"(acc[i][0]), "+d"(acc[i][1]), "+d"(acc[i][2]), "+d"(acc[i][3])
                      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                        "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
   }
   double s = 0;
   for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
double timeit(F f) {
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0); cudaEventCreate(&e1);
   f(); cudaDeviceSynchronize();
   cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
   float ms; cudaEventElapsedTime(&ms, e0, e1);
   return ms * 1e-3;
}

int main() {
   double* out; cudaMalloc(&out, sizeof(double) * 148 * 1024 * 4);
   int iters = 20000;
   for (int warps : {1, 2, 4, 8, 16, 32}) {
      int threads = warps * 32, blocks = 148;
      double t1 = timeit([&] { k_dfma<<<blocks, threads>>>(out, iters); });
      double f1 = 2.0 * 16 * iters * (double)threads * blocks / t1 / 1e12;
      double t2 = timeit([&] { k_dmma884<<<blocks, threads>>>(out, iters); });
      double f2 = 2.0 * 256 * 16 * iters * (double)warps * blocks / t2 / 1e12;
      double t3 = timeit([&] { k_dmma16816<<<blocks, threads>>>(out, iters / 4); });
      double f3 = 2.0 * 16 * 8 * 16 * 8 * (iters / 4) * (double)warps * blocks / t3 / 1e12;
      printf("warps/SM %2d: DFMA %.2f TF/s  DMMA.m8n8k4 %.2f TF/s  mma.m16n8k16 %.2f TF/s\n", warps, f1, f2, f3);
   }
   return 0;
}
