// Dense front factorisation kernels (included by engine.cu).
//
// One elimination step of a front = four launches, batched over all fronts of a level bucket:
//   k_diag    one CTA per front: close the previous panel (count accepted pivots, park failed
//             columns behind the candidates), then factor the next 32x32 diagonal block with
//             complete pivoting (block_ldlt.hxx:290-414) while accumulating L11^{-1}.
//   k_apply   rows below the block: X = A21 P L11^{-T} (as a product with the explicit inverse),
//             L21 = X D^{-1} written IN PLACE, X = L21 D kept in Wld for the Schur updates, the
//             a-posteriori test |l_ij| <= 1/u (ldlt_app.cxx:304-323) by warp shuffle + atomicMin.
//   k_commit  the block's own rows, D, pivot lists, row permutation of finished columns; restores
//             failed columns from the saved originals (CopyBackup, ldlt_app.cxx:423).
//   k_update  C -= L (LD)^T on FP64 tensor cores (mma.sync m8n8k4 -> DMMA), lower tiles only.
// Two-level blocking: panels of NB=32 update only the current outer block of NBO=128 columns
// (mode 0); the rest of the front is updated once per outer block with k <= 128 (mode 1), which
// cuts the HBM traffic of the trailing matrix by 4.
#pragma once

constexpr int TM = 128;          // update tile rows
constexpr int TN = 128;          // update tile cols
constexpr int BK = 32;           // k chunk staged in shared memory
constexpr int LDT = TM + 4;      // padded smem leading dimension: t4*LDT + g distinct mod 16 per half warp

// ------------------------------------------------------------------------------------------
// K1
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_diag(FrontDev* fr, const int* list, Params prm, int mode) {
   FrontDev& f = fr[list[blockIdx.x]];
   const int tid = threadIdx.x, lane = tid & 31, grp = tid >> 5;
   const bool new_outer = mode & 1, finish_only = mode & 2;
   __shared__ double sa[NB * LDS], sld[NB * LDS], s0[NB * LDS], sinv[NB * LDS];
   __shared__ double sd[2 * NB + 2];
   __shared__ unsigned long long skey[8];
   __shared__ int sptype[NB], slperm[NB];

   if (f.flag < 0) return;
   int done = f.done, oend = f.oend, gend = f.gend, w = f.w, jstart = f.jstart, jend0 = f.jend0;
   const int fail = f.fail;
   __syncthreads();
   if (w > 0) {
      const int nelim = compute_nelim(fail, w, f.ibuf + NB);
      const int nfail = w - nelim;
      done += nelim;
      for (int q = 0; q < nfail; ++q) {
         int a = done + q, b = oend - 1 - q;
         if (a < b) block_sym_swap(f, a, b);
      }
      oend -= nfail;
      if (prm.posdef && nfail > 0) {
         if (tid == 0) { f.flag = SSIDS_ERROR_NOT_POS_DEF; f.w = 0; }
         return;
      }
   }
   if (new_outer || finish_only) {
      const int nfo = jend0 - oend;            // failed columns parked at [oend, jend0)
      for (int q = 0; q < nfo; ++q) {
         int a = oend + q, b = gend - 1 - q;
         if (a < b) block_sym_swap(f, a, b);
      }
      gend -= nfo;
      jstart = done;
      jend0 = min(done + NBO, gend);
      oend = jend0;
   }
   w = min(NB, oend - done);
   if (w < 0) w = 0;
   if (finish_only) w = 0;
   __syncthreads();
   if (tid == 0) {
      f.done = done; f.oend = oend; f.gend = gend; f.jstart = jstart; f.jend0 = jend0;
      f.w = w; f.fail = w; f.nelim = done;
   }
   if (w == 0) return;

   for (int i = tid; i < NB * LDS; i += 256) { sa[i] = 0.0; sld[i] = 0.0; sinv[i] = 0.0; }
   __syncthreads();
   {
      double v4[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
         int i = tid + q * 256, r = i & 31, c = i >> 5;
         v4[q] = (r >= c && r < w) ? f.L[(size_t)(done + c) * f.ldl + done + r] : 0.0;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
         int i = tid + q * 256, r = i & 31, c = i >> 5;
         if (r >= c && r < w) { sa[r + c * LDS] = v4[q]; sa[c + r * LDS] = v4[q]; }
      }
   }
   if (tid < NB) { slperm[tid] = tid; sptype[tid] = 0; sinv[tid + tid * LDS] = 1.0; }
   for (int i = tid; i < 2 * NB + 2; i += 256) sd[i] = 0.0;
   __syncthreads();
   for (int i = tid; i < NB * LDS; i += 256) s0[i] = sa[i];
   __syncthreads();

   int err = 0;
   const int c0 = grp * 4;      // this thread's 4 columns; its row is `lane`
   for (int p = 0; p < w;) {
      int t = p, mm = p, pivsiz = 1;
      double a11 = 0, a21 = 0, a22 = 0, detscale = 0, detpiv = 0;
      if (prm.posdef) {
         a11 = sa[p + p * LDS];
         if (!(a11 > 0.0)) { err = SSIDS_ERROR_NOT_POS_DEF; break; }
      } else {
         // argmax |a(r,c)| over the remaining lower triangle with one 64-bit integer key per
         // candidate: the bit pattern of |v| (monotone for non-negative doubles) with its low 10
         // bits replaced by 1023 - (c*32 + r), so that ties go to the smallest column, then row
         unsigned long long key = 0ull;
         if (lane >= p && lane < w) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
               int c = c0 + j;
               if (c >= p && c <= lane) {
                  unsigned long long kk = ((unsigned long long)__double_as_longlong(fabs(sa[lane + c * LDS])) & ~0x3FFull) |
                                          (unsigned long long)(1023 - (c * 32 + lane));
                  key = kk > key ? kk : key;
               }
            }
         }
         {
            unsigned hi = (unsigned)(key >> 32);
            unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
            unsigned lo = (hi == mhi) ? (unsigned)key : 0u;
            unsigned mlo = __reduce_max_sync(0xffffffffu, lo);
            key = ((unsigned long long)mhi << 32) | mlo;
         }
         if (lane == 0) skey[grp] = key;
         __syncthreads();
         key = skey[0];
#pragma unroll
         for (int g = 1; g < 8; ++g) { unsigned long long ok = skey[g]; key = ok > key ? ok : key; }
         const double best = __longlong_as_double((long long)(key & ~0x3FFull));
         {
            int idx = 1023 - (int)(key & 0x3FFull);
            t = idx & 31;
            mm = idx >> 5;
         }
         if (best < prm.small_) {
            if (!prm.action) { err = SSIDS_ERROR_SINGULAR; break; }
            __syncthreads();
            for (int i = tid; i < NB * NB; i += 256) {
               int r = i & 31, c = i >> 5;
               if (r >= p && c >= p && r < w && c < w) { sa[r + c * LDS] = 0.0; }
               if (r >= c && c >= p && r < w) sld[r + c * LDS] = 0.0;
            }
            if (tid >= p && tid < w) { sd[2 * tid] = 0.0; sd[2 * tid + 1] = 0.0; sptype[tid] = 0; }
            __syncthreads();
            break;
         }
         if (t == mm) {
            pivsiz = 1;
            a11 = sa[t + t * LDS];
         } else {
            a11 = sa[mm + mm * LDS];
            a22 = sa[t + t * LDS];
            a21 = sa[t + mm * LDS];
            detscale = 1.0 / fabs(a21);
            detpiv = (a11 * detscale) * a22 - fabs(a21);
            if (fabs(detpiv) >= fabs(a21) / 2) {
               pivsiz = 2;
            } else if (fabs(a11) > fabs(a22)) {
               pivsiz = 1; t = mm;
            } else {
               pivsiz = 1; a11 = a22; mm = t;
            }
         }
      }
      __syncthreads();
      // symmetric swaps bringing the pivot(s) to p (and p+1); sinv follows (M' = P M P^T)
      for (int i = 0; i < pivsiz; ++i) {
         const int dst = p + i;
         const int src = (pivsiz == 1) ? t : (i == 0 ? mm : t);
         if (src != dst) {
            if (tid < NB) {
               double s = sa[dst + tid * LDS]; sa[dst + tid * LDS] = sa[src + tid * LDS]; sa[src + tid * LDS] = s;
            } else if (tid < 2 * NB) {
               int q = tid - NB;
               double s = sinv[dst + q * LDS]; sinv[dst + q * LDS] = sinv[src + q * LDS]; sinv[src + q * LDS] = s;
            } else if (tid < 3 * NB) {
               int q = tid - 2 * NB;
               if (q < p) { double s = sld[dst + q * LDS]; sld[dst + q * LDS] = sld[src + q * LDS]; sld[src + q * LDS] = s; }
            } else if (tid == 3 * NB) {
               int q = slperm[dst]; slperm[dst] = slperm[src]; slperm[src] = q;
            }
            __syncthreads();
            if (tid < NB) {
               double s = sa[tid + dst * LDS]; sa[tid + dst * LDS] = sa[tid + src * LDS]; sa[tid + src * LDS] = s;
            } else if (tid < 2 * NB) {
               int q = tid - NB;
               double s = sinv[q + dst * LDS]; sinv[q + dst * LDS] = sinv[q + src * LDS]; sinv[q + src * LDS] = s;
            }
            __syncthreads();
         }
      }
      if (pivsiz == 1) {
         const double d11 = 1.0 / sa[p + p * LDS];
         __syncthreads();
         if (tid > p && tid < w) {
            double ldr = sa[tid + p * LDS];
            sld[tid + p * LDS] = ldr;
            sa[tid + p * LDS] = ldr * d11;
         }
         __syncthreads();
         if (lane > p && lane < w) {
            const double lr = sa[lane + p * LDS];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
               int c = c0 + j;
               if (c > p) { if (c < w) sa[lane + c * LDS] -= lr * sld[c + p * LDS]; }
               else sinv[lane + c * LDS] -= lr * sinv[p + c * LDS];
            }
         }
         __syncthreads();
         if (tid == 0) { sd[2 * p] = d11; sd[2 * p + 1] = 0.0; sptype[p] = 0; sa[p + p * LDS] = 1.0; }
         if (tid > p && tid < w) sa[p + tid * LDS] = 0.0;
         p += 1;
      } else {
         const double d11 = (a22 * detscale) / detpiv, d22 = (a11 * detscale) / detpiv;
         const double d21 = (-a21 * detscale) / detpiv;
         if (tid > p + 1 && tid < w) {
            double w1 = sa[tid + p * LDS], w2 = sa[tid + (p + 1) * LDS];
            sld[tid + p * LDS] = w1;
            sld[tid + (p + 1) * LDS] = w2;
            sa[tid + p * LDS] = d11 * w1 + d21 * w2;
            sa[tid + (p + 1) * LDS] = d21 * w1 + d22 * w2;
         }
         __syncthreads();
         if (lane > p + 1 && lane < w) {
            const double l1 = sa[lane + p * LDS], l2 = sa[lane + (p + 1) * LDS];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
               int c = c0 + j;
               if (c > p + 1) { if (c < w) sa[lane + c * LDS] -= l1 * sld[c + p * LDS] + l2 * sld[c + (p + 1) * LDS]; }
               else sinv[lane + c * LDS] -= l1 * sinv[p + c * LDS] + l2 * sinv[p + 1 + c * LDS];
            }
         }
         __syncthreads();
         if (tid == 0) {
            sd[2 * p] = d11; sd[2 * p + 1] = d21; sd[2 * p + 2] = DINF; sd[2 * p + 3] = d22;
            sptype[p] = 1; sptype[p + 1] = 2;
            sa[p + p * LDS] = 1.0; sa[p + 1 + p * LDS] = 0.0; sa[p + (p + 1) * LDS] = 0.0;
            sa[p + 1 + (p + 1) * LDS] = 1.0;
            sld[p + 1 + p * LDS] = 0.0;
         }
         if (tid > p + 1 && tid < w) { sa[p + tid * LDS] = 0.0; sa[p + 1 + tid * LDS] = 0.0; }
         p += 2;
      }
      __syncthreads();
   }
   if (err) {
      if (tid == 0) { f.flag = err; f.w = 0; }
      return;
   }
   __syncthreads();
   // publish
   double* L11 = f.dbuf;
   double* LD11 = f.dbuf + NB * NB;
   double* A0p = f.dbuf + 2 * NB * NB;
   double* Linv = f.dbuf + 3 * NB * NB;      // row-major: Linv[k*NB + i] = (L11^{-1})(k,i)
   double* dk = f.dbuf + 4 * NB * NB;
   const int kofs = done - jstart;
   for (int idx = tid; idx < NB * NB; idx += 256) {
      const int i = idx & 31, k = idx >> 5;
      if (i < w && k < w) {
         double lv = i > k ? sa[i + k * LDS] : (i == k ? 1.0 : 0.0);
         double ldv = i > k ? sld[i + k * LDS] : 0.0;
         L11[i + k * NB] = lv;
         LD11[i + k * NB] = ldv;
         A0p[i + k * NB] = s0[slperm[i] + slperm[k] * LDS];
         f.Wld[(size_t)(kofs + k) * f.m + done + i] = ldv;
      }
      // transpose while writing: thread idx -> (k = idx&31 row of Linv, i = idx>>5)
      {
         const int kk = idx & 31, ii = idx >> 5;
         Linv[kk * NB + ii] = (kk < w && ii <= kk) ? sinv[kk + ii * LDS] : 0.0;
      }
   }
   for (int i = tid; i < 2 * NB + 2; i += 256) dk[i] = sd[i];
   if (tid < NB) { f.ibuf[tid] = slperm[tid]; f.ibuf[NB + tid] = sptype[tid]; }
   if (tid == 0) {
      int permuted = 0;
      for (int k = 0; k < w; ++k) if (slperm[k] != k) permuted = 1;
      f.ibuf[2 * NB] = permuted;
   }
}

// ------------------------------------------------------------------------------------------
// K2
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_apply(FrontDev* fr, const int* list, Params prm) {
   FrontDev& f = fr[list[blockIdx.x]];
   const int w = f.w;
   if (w == 0 || f.flag < 0) return;
   const int done = f.done, m = f.m;
   const int r0 = done + w + blockIdx.y * 128;
   if (r0 >= m) return;
   __shared__ double sinv[NB * NB];
   __shared__ double sdk[2 * NB + 2];
   __shared__ int slperm[NB], sptype[NB];
   for (int i = threadIdx.x; i < NB * NB; i += 128) sinv[i] = f.dbuf[3 * NB * NB + i];
   for (int i = threadIdx.x; i < 2 * NB + 2; i += 128) sdk[i] = f.dbuf[4 * NB * NB + i];
   if (threadIdx.x < NB) { slperm[threadIdx.x] = f.ibuf[threadIdx.x]; sptype[threadIdx.x] = f.ibuf[NB + threadIdx.x]; }
   __syncthreads();
   const int r = r0 + threadIdx.x;
   int myfail = w;
   if (r < m) {
      double a[NB], x[NB];
      const size_t ldl = f.ldl;
      const double* Lr = f.L + r;
#pragma unroll
      for (int k = 0; k < NB; ++k) a[k] = (k < w) ? Lr[(size_t)(done + slperm[k]) * ldl] : 0.0;
      double* Wor = f.Wo + r;
#pragma unroll
      for (int k = 0; k < NB; ++k) if (k < w) Wor[(size_t)k * m] = a[k];
#pragma unroll
      for (int k = 0; k < NB; ++k) {
         double s = 0.0;
#pragma unroll
         for (int i = 0; i <= k; ++i) s += a[i] * sinv[k * NB + i];
         x[k] = s;
      }
      const int kofs = done - f.jstart;
      double lv[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k) {
         lv[k] = 0.0;
         if (k < w) {
            int pt = sptype[k];
            double l;
            if (pt == 0) {
               double d11 = sdk[2 * k];
               if (d11 == 0.0) l = (fabs(x[k]) < prm.small_) ? 0.0 : DINF * x[k];
               else l = x[k] * d11;
            } else if (pt == 1) {
               l = sdk[2 * k] * x[k] + sdk[2 * k + 1] * x[(k + 1) & (NB - 1)];
            } else {
               l = sdk[2 * k - 1] * x[(k - 1) & (NB - 1)] + sdk[2 * k + 1] * x[k];
            }
            lv[k] = l;
            if (!prm.posdef && (!(fabs(l) <= prm.lim)) && k < myfail) myfail = k;
         }
      }
      double* Lw = f.L + (size_t)done * ldl + r;
      double* Wr = f.Wld + (size_t)kofs * m + r;
#pragma unroll
      for (int k = 0; k < NB; ++k) if (k < w) { Lw[(size_t)k * ldl] = lv[k]; Wr[(size_t)k * m] = x[k]; }
   }
#pragma unroll
   for (int off = 16; off > 0; off >>= 1) myfail = min(myfail, __shfl_xor_sync(0xffffffffu, myfail, off));
   if ((threadIdx.x & 31) == 0 && myfail < w) atomicMin(&f.fail, myfail);
}

// ------------------------------------------------------------------------------------------
// K3.  blockIdx.y == 0: the block's own rows, D, pivot lists.  Then cdiv(n,128) chunks for the
// row permutation of finished columns, then cdiv(m,128) chunks restoring failed columns.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_commit(FrontDev* fr, const int* list, int ncolchunks) {
   FrontDev& f = fr[list[blockIdx.x]];
   const int w = f.w;
   if (w == 0 || f.flag < 0) return;
   const int done = f.done, m = f.m;
   const int* lperm = f.ibuf;
   const int nelim = compute_nelim(f.fail, w, f.ibuf + NB);
   const size_t ldl = f.ldl;
   const int y = blockIdx.y;
   if (y == 0) {
      const double* L11 = f.dbuf;
      const double* A0p = f.dbuf + 2 * NB * NB;
      for (int idx = threadIdx.x; idx < NB * NB; idx += 128) {
         const int i = idx & 31, k = idx >> 5;
         if (i < w && k <= i)
            f.L[(size_t)(done + k) * ldl + done + i] = k < nelim ? (i == k ? 1.0 : L11[i + k * NB]) : A0p[i + k * NB];
      }
      const double* dk = f.dbuf + 4 * NB * NB;
      for (int k = threadIdx.x; k < nelim; k += 128) {
         f.D[2 * (done + k)] = dk[2 * k];
         f.D[2 * (done + k) + 1] = dk[2 * k + 1];
      }
      if (f.ibuf[2 * NB] && threadIdx.x < 64) {
         int* arr = threadIdx.x < 32 ? f.perm : f.orig;
         const int q = threadIdx.x & 31;
         int v = q < w ? arr[done + lperm[q]] : 0;
         __syncwarp();
         if (q < w) arr[done + q] = v;
      }
   } else if (y <= ncolchunks) {
      if (!f.ibuf[2 * NB]) return;
      const int c = (y - 1) * 128 + threadIdx.x;
      if (c >= done) return;
      double tmp[NB];
#pragma unroll
      for (int q = 0; q < NB; ++q) tmp[q] = q < w ? f.L[(size_t)c * ldl + done + lperm[q]] : 0.0;
#pragma unroll
      for (int q = 0; q < NB; ++q) if (q < w) f.L[(size_t)c * ldl + done + q] = tmp[q];
   } else {
      if (nelim == w) return;
      const int r = done + w + (y - 1 - ncolchunks) * 128 + threadIdx.x;
      if (r >= m) return;
      for (int k = nelim; k < w; ++k) f.L[(size_t)(done + k) * ldl + r] = f.Wo[(size_t)k * m + r];
   }
}

// ------------------------------------------------------------------------------------------
// K4: FP64 tensor-core Schur update.  C(r,c) -= sum_k A(r,k) B(c,k) over the lower triangle of
// the region cols [cbeg, cend), rows >= col, with A = L(:, k0:k0+kc) and B = Wld(:, kofs:kofs+kc).
// CTA = 256 threads = 8 warps (4 x 2); CTA tile 128 x 128; warp tile 32 x 64 = 4 x 8 m8n8k4 MMAs.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                : "+d"(d0), "+d"(d1)
                : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc, bool valid) {
   unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
   int sz = valid ? 8 : 0;
   asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__global__ void __launch_bounds__(256, 1) k_update(FrontDev* fr, const int* list, int mode) {
   const FrontDev& f = fr[list[blockIdx.x]];
   const int w = f.w;
   if (f.flag < 0) return;
   const int nel_cur = w > 0 ? compute_nelim(f.fail, w, f.ibuf + NB) : 0;
   int k0, kc, cbeg, cend;
   const int m = f.m;
   if (mode == 0) {
      k0 = f.done; kc = nel_cur; cbeg = f.done + nel_cur; cend = f.jend0;
   } else {
      k0 = f.jstart; kc = f.done + nel_cur - f.jstart; cbeg = f.jend0; cend = m;
   }
   if (kc <= 0 || cbeg >= cend) return;
   const int kofs = k0 - f.jstart;
   const int ntr = (m - cbeg + TM - 1) / TM;
   const int ntc = (cend - cbeg + TN - 1) / TN;
   const int ntiles = ntr * ntc;
   extern __shared__ double smem[];
   double* As = smem;                       // [2][BK][LDT]
   double* Bs = smem + 2 * BK * LDT;        // [2][BK][LDT]
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const int wr = warp & 3, wc = warp >> 2;  // warp tile origin: rows wr*32, cols wc*64
   const int g = lane >> 2, t4 = lane & 3;
   const size_t ldl = f.ldl;
   const double* Abase = f.L + (size_t)k0 * ldl;
   const double* Bbase = f.Wld + (size_t)kofs * m;
   for (int id = blockIdx.y; id < ntiles; id += gridDim.y) {
      const int tj = id / ntr, ti = id % ntr;
      const int rbase = cbeg + ti * TM, cbase = cbeg + tj * TN;
      if (rbase + TM - 1 < cbase) continue;   // tile entirely above the diagonal
      double acc[4][8][2];
      const int nchunk = (kc + BK - 1) / BK;
      // stage loader: 256 threads move BK x 128 doubles of A and of B (16 + 16 cp.async each)
      auto load_stage = [&](int chunk, int buf) {
         const int kb = chunk * BK;
         double* as = As + buf * BK * LDT;
         double* bs = Bs + buf * BK * LDT;
#pragma unroll
         for (int q = 0; q < 16; ++q) {
            const int e = q * 256 + tid;
            const int kk = e >> 7, rr = e & 127;
            const bool kv = kb + kk < kc;
            const int ra = rbase + rr, cb = cbase + rr;
            cp_async8(as + kk * LDT + rr, Abase + (size_t)(kb + kk) * ldl + ra, kv && ra < m);
            cp_async8(bs + kk * LDT + rr, Bbase + (size_t)(kb + kk) * m + cb, kv && cb < cend);
         }
         cp_async_commit();
      };
      __syncthreads();   // previous tile's smem reads are done
      load_stage(0, 0);
      // accumulators start as the C tile (one batch of independent loads, overlapped with the
      // first stage); the product is subtracted by feeding -B to the MMA
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
         for (int h = 0; h < 2; ++h) {
            const int c = cbase + wc * 64 + j * 8 + 2 * t4 + h;
            const bool cv = c < cend;
            const double* colp = (c < f.n) ? (f.L + (size_t)c * ldl) : (f.C + (size_t)(c - f.n) * f.ldc - f.n);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
               const int r = rbase + wr * 32 + i * 8 + g;
               acc[i][j][h] = (cv && r < m && r >= c) ? colp[r] : 0.0;
            }
         }
      }
      for (int ch = 0; ch < nchunk; ++ch) {
         const int buf = ch & 1;
         if (ch + 1 < nchunk) { load_stage(ch + 1, buf ^ 1); cp_async_wait<1>(); }
         else cp_async_wait<0>();
         __syncthreads();
         const double* as = As + buf * BK * LDT + wr * 32 + g;
         const double* bs = Bs + buf * BK * LDT + wc * 64 + g;
#pragma unroll
         for (int ks = 0; ks < BK; ks += 4) {
            double af[4], bf[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = as[(ks + t4) * LDT + i * 8];
#pragma unroll
            for (int j = 0; j < 8; ++j) bf[j] = -bs[(ks + t4) * LDT + j * 8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
               for (int j = 0; j < 8; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
         }
         __syncthreads();
      }
      // epilogue: store only
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
         for (int h = 0; h < 2; ++h) {
            const int c = cbase + wc * 64 + j * 8 + 2 * t4 + h;
            if (c >= cend) continue;
            double* colp = (c < f.n) ? (f.L + (size_t)c * ldl) : (f.C + (size_t)(c - f.n) * f.ldc - f.n);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
               const int r = rbase + wr * 32 + i * 8 + g;
               if (r < m && r >= c) colp[r] = acc[i][j][h];
            }
         }
      }
   }
}
