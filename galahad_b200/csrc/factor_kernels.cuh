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
constexpr int TN = 64;           // update tile cols
constexpr int BK = 32;           // k chunk staged in shared memory
constexpr int LDA = TM + 4;      // padded smem leading dimensions: t4*LD + g distinct mod 16 per half warp
constexpr int LDB = TN + 4;

// ------------------------------------------------------------------------------------------
// K1
// ------------------------------------------------------------------------------------------
// One warp per front; lane r keeps row r of the (full, symmetric) block in registers, so a pivot
// step is a register scan, two warp reductions, one shared-memory broadcast of the pivot column
// and 32 FMAs -- no block barriers (the 128/256-thread shared-memory versions spent 2400 cycles
// per pivot in barrier / dependent-issue stalls, profiles/r1_notes.md).  Rows and columns are
// never moved: eliminated indices leave the `alive` mask and the pivot sequence is recorded.
__device__ __forceinline__ double pick32(const double (&a)[NB], int q) {
   switch (q) {   // q is warp-uniform: one indirect branch
#define B200_PICK(i) case i: return a[i];
      B200_PICK(0) B200_PICK(1) B200_PICK(2) B200_PICK(3) B200_PICK(4) B200_PICK(5) B200_PICK(6) B200_PICK(7)
      B200_PICK(8) B200_PICK(9) B200_PICK(10) B200_PICK(11) B200_PICK(12) B200_PICK(13) B200_PICK(14) B200_PICK(15)
      B200_PICK(16) B200_PICK(17) B200_PICK(18) B200_PICK(19) B200_PICK(20) B200_PICK(21) B200_PICK(22) B200_PICK(23)
      B200_PICK(24) B200_PICK(25) B200_PICK(26) B200_PICK(27) B200_PICK(28) B200_PICK(29) B200_PICK(30) B200_PICK(31)
#undef B200_PICK
   }
   return 0.0;
}

__global__ void __launch_bounds__(32) k_diag(FrontDev* fr, const int* list, int slot0, Params prm, int mode) {
   FrontDev& f = fr[front_slot(list, slot0)];
   const int lane = threadIdx.x;
   const bool new_outer = mode & 1, finish_only = mode & 2;
   __shared__ double s0[NB * LDS];          // original block, full symmetric
   __shared__ double sLc[NB * LDS];         // sLc[r + k*LDS]: multiplier of (original) row r for pivot k
   __shared__ double sLd[NB][NB];           // sLd[k][r]: the same times D (pivot column before scaling), contiguous per pivot
   __shared__ double sL[NB * LDS], sinv[NB * LDS];   // L11 in pivot order and its inverse
   __shared__ double sd[2 * NB + 2];
   __shared__ int sptype[NB], spord[NB];

   if (f.flag < 0) return;
   int done = f.done, oend = f.oend, gend = f.gend, w = f.w, jstart = f.jstart, jend0 = f.jend0;
   const int fail = f.fail;
   __syncwarp();
   if (w > 0) {
      const int nelim = compute_nelim(fail, w, f.p2mask);
      const int nfail = w - nelim;
      done += nelim;
      for (int q = 0; q < nfail; ++q) {
         int a = done + q, b = oend - 1 - q;
         if (a < b) block_sym_swap(f, a, b);
      }
      oend -= nfail;
      if (prm.posdef && nfail > 0) {
         if (lane == 0) { f.flag = SSIDS_ERROR_NOT_POS_DEF; f.w = 0; }
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
   __syncwarp();
   if (lane == 0) {
      f.done = done; f.oend = oend; f.gend = gend; f.jstart = jstart; f.jend0 = jend0;
      f.w = w; f.fail = w; f.nelim = done;
   }
   if (w == 0) return;

   // load: coalesced lower triangle -> shared (both triangles), then one row per lane
   {
      const double* Lb = f.L + (size_t)done * f.ldl + done + lane;
      double v[NB];
#pragma unroll
      for (int c = 0; c < NB; ++c) v[c] = (c < w && lane >= c && lane < w) ? Lb[(size_t)c * f.ldl] : 0.0;
#pragma unroll
      for (int c = 0; c < NB; ++c) {
         if (lane >= c) { s0[lane + c * LDS] = v[c]; s0[c + lane * LDS] = v[c]; }
         sLc[lane + c * LDS] = 0.0;
         sLd[c][lane] = 0.0;
      }
   }
   for (int i = lane; i < 2 * NB + 2; i += 32) sd[i] = 0.0;
   sptype[lane] = 0;
   spord[lane] = lane;
   __syncwarp();
   double a[NB];
#pragma unroll
   for (int c = 0; c < NB; ++c) a[c] = s0[lane + c * LDS];
   double dg = s0[lane + lane * LDS];
   unsigned alive = (w == NB) ? 0xffffffffu : ((1u << w) - 1u);
   int np = 0, err = 0;
   while (alive) {
      int q1, q2 = -1;
      double d11 = 0.0, d21 = 0.0, d22 = 0.0;
      if (prm.posdef) {
         q1 = __ffs(alive) - 1;
         const double a11 = __shfl_sync(0xffffffffu, dg, q1);
         if (!(a11 > 0.0)) { err = SSIDS_ERROR_NOT_POS_DEF; break; }
         d11 = 1.0 / a11;
      } else {
         // largest |a(r,c)| over the alive rows / columns; compared on the high word of the double
         // (sign cleared, 15 mantissa bits): ties -> smallest column, then row
         // per-lane keys (|a| high word with the low 5 bits replaced by 31-c), reduced by a max tree
         const bool rowalive = (alive >> lane) & 1u;
         unsigned kb[NB];
#pragma unroll
         for (int c = 0; c < NB; ++c)
            kb[c] = ((alive >> c) & 1u) ? ((((unsigned)__double2hiint(a[c])) & 0x7fffffe0u) | (unsigned)(31 - c)) : 0u;
#pragma unroll
         for (int st = NB / 2; st > 0; st >>= 1)
#pragma unroll
            for (int c = 0; c < st; ++c) kb[c] = max(kb[c], kb[c + st]);
         const unsigned bk = rowalive ? kb[0] : 0u;
         const int bc = 31 - (int)(bk & 31u);
         const unsigned mk = __reduce_max_sync(0xffffffffu, bk & ~31u);
         const int lo = min(lane, bc), hi = max(lane, bc);
         const unsigned idx = (rowalive && (bk & ~31u) == mk) ? (unsigned)(lo * 32 + hi) : 0xffffu;
         const unsigned mi = __reduce_min_sync(0xffffffffu, idx);
         const int mm = (int)(mi >> 5), t = (int)(mi & 31u);
         const double a21 = __shfl_sync(0xffffffffu, pick32(a, mm), t);   // A(t, mm)
         if (fabs(a21) < prm.small_) {
            if (!prm.action) { err = SSIDS_ERROR_SINGULAR; break; }
            // everything left is numerically zero: record zero pivots (block_ldlt.hxx:303-318)
            while (alive) {
               const int q = __ffs(alive) - 1;
               if (lane == 0) { spord[np] = q; sptype[np] = 0; sd[2 * np] = 0.0; sd[2 * np + 1] = 0.0; }
               alive &= ~(1u << q);
               ++np;
            }
            break;
         }
         const double a11 = __shfl_sync(0xffffffffu, dg, mm), a22 = __shfl_sync(0xffffffffu, dg, t);
         if (t == mm) {
            q1 = t;
            d11 = 1.0 / a11;
         } else {
            const double detscale = 1.0 / fabs(a21);
            const double detpiv = (a11 * detscale) * a22 - fabs(a21);
            if (fabs(detpiv) >= fabs(a21) / 2) {
               q1 = mm; q2 = t;
               d11 = (a22 * detscale) / detpiv;
               d22 = (a11 * detscale) / detpiv;
               d21 = (-a21 * detscale) / detpiv;
            } else if (fabs(a11) > fabs(a22)) {
               q1 = mm; d11 = 1.0 / a11;
            } else {
               q1 = t; d11 = 1.0 / a22;
            }
         }
      }
      if (q2 < 0) {
         alive &= ~(1u << q1);
         const bool upd = (alive >> lane) & 1u;
         const double ldr = upd ? pick32(a, q1) : 0.0;
         const double l = ldr * d11;
         sLd[np][lane] = ldr;
         sLc[lane + np * LDS] = l;
         __syncwarp();
         const double2* ldv = reinterpret_cast<const double2*>(&sLd[np][0]);
#pragma unroll
         for (int c2 = 0; c2 < NB / 2; ++c2) {
            const double2 v = ldv[c2];
            a[2 * c2] -= l * v.x;
            a[2 * c2 + 1] -= l * v.y;
         }
         dg -= l * ldr;
         if (lane == 0) { spord[np] = q1; sptype[np] = 0; sd[2 * np] = d11; sd[2 * np + 1] = 0.0; }
         np += 1;
      } else {
         alive &= ~((1u << q1) | (1u << q2));
         const bool upd = (alive >> lane) & 1u;
         const double w1 = upd ? pick32(a, q1) : 0.0, w2 = upd ? pick32(a, q2) : 0.0;
         const double l1 = d11 * w1 + d21 * w2, l2 = d21 * w1 + d22 * w2;
         sLd[np][lane] = w1;
         sLd[np + 1][lane] = w2;
         sLc[lane + np * LDS] = l1;
         sLc[lane + (np + 1) * LDS] = l2;
         __syncwarp();
         const double2* ldv1 = reinterpret_cast<const double2*>(&sLd[np][0]);
         const double2* ldv2 = reinterpret_cast<const double2*>(&sLd[np + 1][0]);
#pragma unroll
         for (int c2 = 0; c2 < NB / 2; ++c2) {
            const double2 v1 = ldv1[c2], v2 = ldv2[c2];
            a[2 * c2] -= l1 * v1.x + l2 * v2.x;
            a[2 * c2 + 1] -= l1 * v1.y + l2 * v2.y;
         }
         dg -= l1 * w1 + l2 * w2;
         if (lane == 0) {
            spord[np] = q1; spord[np + 1] = q2;
            sptype[np] = 1; sptype[np + 1] = 2;
            sd[2 * np] = d11; sd[2 * np + 1] = d21; sd[2 * np + 2] = DINF; sd[2 * np + 3] = d22;
         }
         np += 2;
      }
   }
   if (err) {
      if (lane == 0) { f.flag = err; f.w = 0; }
      return;
   }
   __syncwarp();
   // L11 in pivot order (unit lower), then its inverse: lane j solves L11 z = e_j
#pragma unroll 4
   for (int k = 0; k < NB; ++k) {
      double lv = 0.0;
      if (lane < w && k < w) lv = lane > k ? sLc[spord[lane] + k * LDS] : (lane == k ? 1.0 : 0.0);
      sL[lane + k * LDS] = lv;
      sinv[lane + k * LDS] = 0.0;
   }
   __syncwarp();
   if (lane < w) {
      const int j = lane;
      sinv[j + j * LDS] = 1.0;
      for (int i = j + 1; i < w; ++i) {
         double s = 0.0;
         for (int t = j; t < i; ++t) s += sL[i + t * LDS] * sinv[t + j * LDS];
         sinv[i + j * LDS] = -s;
      }
   }
   __syncwarp();
   // publish
   double* L11 = f.dbuf;
   double* LD11 = f.dbuf + NB * NB;
   double* A0p = f.dbuf + 2 * NB * NB;
   double* Linv = f.dbuf + 3 * NB * NB;      // row-major: Linv[k*NB + i] = (L11^{-1})(k,i)
   double* dk = f.dbuf + 4 * NB * NB;
   const int kofs = done - jstart;
   const int pl = spord[lane];
   for (int k = 0; k < NB; ++k) {
      const int i = lane;
      if (i < w && k < w) {
         const double ldv = i > k ? sLd[k][pl] : 0.0;
         L11[i + k * NB] = sL[i + k * LDS];
         LD11[i + k * NB] = ldv;
         A0p[i + k * NB] = s0[pl + spord[k] * LDS];
         f.Wld[(size_t)(kofs + k) * f.ldl + done + i] = ldv;
      }
      // Linv row k, column `lane`
      Linv[k * NB + lane] = (k < w && lane <= k) ? sinv[k + lane * LDS] : 0.0;
   }
   for (int i = lane; i < 2 * NB + 2; i += 32) dk[i] = sd[i];
   f.ibuf[lane] = spord[lane];
   f.ibuf[NB + lane] = sptype[lane];
   {
      const bool moved = lane < w && spord[lane] != lane;
      const bool second = lane < w && sptype[lane] == 2;
      const unsigned pm = __ballot_sync(0xffffffffu, moved), m2 = __ballot_sync(0xffffffffu, second);
      if (lane == 0) { f.ibuf[2 * NB] = pm != 0; f.p2mask = m2; }
   }
}

// ------------------------------------------------------------------------------------------
// K2
// ------------------------------------------------------------------------------------------
constexpr int AR = 64;    // rows per CTA in k_apply

__global__ void __launch_bounds__(256) k_apply(FrontDev* fr, const int* list, int slot0, Params prm) {
   FrontDev& f = fr[front_slot(list, slot0)];
   const int w = f.w;
   if (w == 0 || f.flag < 0) return;
   const int done = f.done, m = f.m;
   const int r0 = done + w + blockIdx.y * AR;
   if (r0 >= m) return;
   __shared__ double sinv[NB * NB];
   __shared__ double sdk[2 * NB + 2];
   __shared__ int slperm[NB], sptype[NB];
   __shared__ double at[AR][NB + 1], xt[AR][NB + 1];
   const int tid = threadIdx.x;
   for (int i = tid; i < NB * NB; i += 256) sinv[i] = f.dbuf[3 * NB * NB + i];
   for (int i = tid; i < 2 * NB + 2; i += 256) sdk[i] = f.dbuf[4 * NB * NB + i];
   if (tid < NB) { slperm[tid] = f.ibuf[tid]; sptype[tid] = f.ibuf[NB + tid]; }
   __syncthreads();
   const int rl = tid & (AR - 1), kq = tid >> 6;      // row within the tile, group of 8 columns
   const int r = r0 + rl;
   const bool rv = r < m;
   const size_t ldl = f.ldl;
   const int kofs = done - f.jstart;
   {
      double a8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
         const int k = kq * 8 + j;
         a8[j] = (rv && k < w) ? f.L[(size_t)(done + slperm[k]) * ldl + r] : 0.0;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
         const int k = kq * 8 + j;
         at[rl][k] = a8[j];
         if (rv && k < w) f.Wo[(size_t)k * ldl + r] = a8[j];
      }
   }
   __syncthreads();
   {
      double x8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x8[j] = 0.0;
      const int kmax = kq * 8 + 7;
      for (int i = 0; i <= kmax; ++i) {
         const double ai = at[rl][i];
#pragma unroll
         for (int j = 0; j < 8; ++j) x8[j] += ai * sinv[(kq * 8 + j) * NB + i];   // Linv(k,i) = 0 for i > k
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) xt[rl][kq * 8 + j] = x8[j];
   }
   __syncthreads();
   int myfail = w;
   {
      double l8[8], x8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
         const int k = kq * 8 + j;
         l8[j] = 0.0;
         x8[j] = xt[rl][k];
         if (k < w) {
            const int pt = sptype[k];
            double l;
            if (pt == 0) {
               const double d11 = sdk[2 * k];
               if (d11 == 0.0) l = (fabs(x8[j]) < prm.small_) ? 0.0 : DINF * x8[j];
               else l = x8[j] * d11;
            } else if (pt == 1) {
               l = sdk[2 * k] * x8[j] + sdk[2 * k + 1] * xt[rl][(k + 1) & (NB - 1)];
            } else {
               l = sdk[2 * k - 1] * xt[rl][(k - 1) & (NB - 1)] + sdk[2 * k + 1] * x8[j];
            }
            l8[j] = l;
            if (rv && !prm.posdef && (!(fabs(l) <= prm.lim)) && k < myfail) myfail = k;
         }
      }
      if (rv) {
#pragma unroll
         for (int j = 0; j < 8; ++j) {
            const int k = kq * 8 + j;
            if (k < w) {
               f.L[(size_t)(done + k) * ldl + r] = l8[j];
               f.Wld[(size_t)(kofs + k) * ldl + r] = x8[j];
            }
         }
      }
   }
#pragma unroll
   for (int off = 16; off > 0; off >>= 1) myfail = min(myfail, __shfl_xor_sync(0xffffffffu, myfail, off));
   if ((tid & 31) == 0 && myfail < w) atomicMin(&f.fail, myfail);
}

// ------------------------------------------------------------------------------------------
// K3.  blockIdx.y == 0: the block's own rows, D, pivot lists.  Then cdiv(n,128) chunks for the
// row permutation of finished columns, then cdiv(m,128) chunks restoring failed columns.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_commit(FrontDev* fr, const int* list, int slot0, int ncolchunks) {
   FrontDev& f = fr[front_slot(list, slot0)];
   const int w = f.w;
   if (w == 0 || f.flag < 0) return;
   const int done = f.done, m = f.m;
   const int* lperm = f.ibuf;
   const int nelim = compute_nelim(f.fail, w, f.p2mask);
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
      for (int k = nelim; k < w; ++k) f.L[(size_t)(done + k) * ldl + r] = f.Wo[(size_t)k * ldl + r];
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

__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc, int src_bytes) {
   unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
   asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gsrc), "r"(src_bytes));
}

// CTA = 256 threads = 8 warps (4 x 2); CTA tile 128 x 64; warp tile 32 x 32 = 4 x 4 m8n8k4 MMAs;
// two CTAs per SM.  ncu (profiles/r1_notes.md) showed the kernel bound by per-warp instruction
// issue around the MMAs, so interior tiles take a path without bounds predicates and with
// strength-reduced addresses, and tile coordinates advance incrementally.
__global__ void __launch_bounds__(256, 2) k_update(FrontDev* fr, const int* list, int slot0, int mode) {
   const FrontDev& f = fr[front_slot(list, slot0)];
   const int w = f.w;
   if (f.flag < 0) return;
   const int nel_cur = w > 0 ? compute_nelim(f.fail, w, f.p2mask) : 0;
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
   extern __shared__ double smem[];
   double* As = smem;                       // [2][BK][LDA]
   double* Bs = smem + 2 * BK * LDA;        // [2][BK][LDB]
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const int wr = warp & 3, wc = warp >> 2;  // 4 x 2 warps; warp tile origin: rows wr*32, cols wc*32
   const int g = lane >> 2, t4 = lane & 3;
   const size_t ldl = f.ldl;
   const int fn = f.n, ldc = f.ldc;
   double* const Lp = f.L;
   double* const Cp = f.C;
   const double* Abase = Lp + (size_t)k0 * ldl;
   const double* Bbase = f.Wld + (size_t)kofs * ldl;
   const bool vec16 = (cbeg & 1) == 0;      // 16-byte cp.async needs even row offsets (ldl is even)
   const int nchunk = (kc + BK - 1) / BK;
   // tile (ti,tj) touches the lower triangle iff tj <= 2*ti+1 (TM = 2*TN, shared origin cbeg);
   // tiles are numbered row by row; a CTA walks its tiles incrementally
   int ti = 0, tj = blockIdx.y;
   for (;;) {
      // normalise (ti,tj)
      while (ti < ntr) { const int wd = min(ntc, 2 * ti + 2); if (tj < wd) break; tj -= wd; ++ti; }
      if (ti >= ntr) break;
      const int rbase = cbeg + ti * TM, cbase = cbeg + tj * TN;
      const bool full = (rbase + TM <= m) && (cbase + TN <= cend) && vec16;
      const bool below = cbase + TN - 1 <= rbase;     // tile entirely below the diagonal
      double acc[4][4][2];
      auto load_stage = [&](int chunk, int buf) {
         const int kb = chunk * BK;
         double* as = As + buf * BK * LDA;
         double* bs = Bs + buf * BK * LDB;
         if (full && kb + BK <= kc) {
            // thread -> (k row = q*4 + tid/64, piece (tid%64)*2) for A; (q*8 + tid/32, (tid%32)*2) for B
            const double* pa = Abase + (size_t)(kb + (tid >> 6)) * ldl + rbase + (tid & 63) * 2;
            double* sa_ = as + (tid >> 6) * LDA + (tid & 63) * 2;
            const size_t sta = 4 * ldl;
#pragma unroll
            for (int q = 0; q < 8; ++q) cp_async16(sa_ + q * 4 * LDA, pa + q * sta, 16);
            const double* pb = Bbase + (size_t)(kb + (tid >> 5)) * ldl + cbase + (tid & 31) * 2;
            double* sb_ = bs + (tid >> 5) * LDB + (tid & 31) * 2;
            const size_t stb = 8 * ldl;
#pragma unroll
            for (int q = 0; q < 4; ++q) cp_async16(sb_ + q * 8 * LDB, pb + q * stb, 16);
         } else if (vec16) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
               const int e = q * 256 + tid, kk = e >> 6, rr = (e & 63) * 2;
               const int ra = rbase + rr;
               const int nb = (kb + kk < kc) ? (ra + 1 < m ? 16 : (ra < m ? 8 : 0)) : 0;
               cp_async16(as + kk * LDA + rr, nb ? Abase + (size_t)(kb + kk) * ldl + ra : Abase, nb);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
               const int e = q * 256 + tid, kk = e >> 5, rr = (e & 31) * 2;
               const int cb = cbase + rr;
               const int nb = (kb + kk < kc) ? (cb + 1 < cend ? 16 : (cb < cend ? 8 : 0)) : 0;
               cp_async16(bs + kk * LDB + rr, nb ? Bbase + (size_t)(kb + kk) * ldl + cb : Bbase, nb);
            }
         } else {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
               const int e = q * 256 + tid, kk = e >> 7, rr = e & 127;
               const int ra = rbase + rr;
               const bool ok = kb + kk < kc && ra < m;
               cp_async8(as + kk * LDA + rr, ok ? Abase + (size_t)(kb + kk) * ldl + ra : Abase, ok);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
               const int e = q * 256 + tid, kk = e >> 6, rr = e & 63;
               const int cb = cbase + rr;
               const bool ok = kb + kk < kc && cb < cend;
               cp_async8(bs + kk * LDB + rr, ok ? Bbase + (size_t)(kb + kk) * ldl + cb : Bbase, ok);
            }
         }
         cp_async_commit();
      };
      __syncthreads();   // previous tile's smem reads are done
      load_stage(0, 0);
      // accumulators start as the C tile; the product is subtracted by feeding -B to the MMA
      const int r0 = rbase + wr * 32 + g, c0 = cbase + wc * 32 + 2 * t4;
      if (full && below) {
#pragma unroll
         for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
               const int c = c0 + j * 8 + h;
               const double* colp = (c < fn) ? (Lp + (size_t)c * ldl + r0) : (Cp + (size_t)(c - fn) * ldc + (r0 - fn));
#pragma unroll
               for (int i = 0; i < 4; ++i) acc[i][j][h] = colp[i * 8];
            }
      } else {
#pragma unroll
         for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
               const int c = c0 + j * 8 + h;
               const bool cv = c < cend;
               const double* colp = (c < fn) ? (Lp + (size_t)c * ldl) : (Cp + (size_t)(c - fn) * ldc - fn);
#pragma unroll
               for (int i = 0; i < 4; ++i) {
                  const int r = r0 + i * 8;
                  acc[i][j][h] = (cv && r < m && r >= c) ? colp[r] : 0.0;
               }
            }
      }
      for (int ch = 0; ch < nchunk; ++ch) {
         const int buf = ch & 1;
         if (ch + 1 < nchunk) { load_stage(ch + 1, buf ^ 1); cp_async_wait<1>(); }
         else cp_async_wait<0>();
         __syncthreads();
         const double* as = As + buf * BK * LDA + wr * 32 + g + t4 * LDA;
         const double* bs = Bs + buf * BK * LDB + wc * 32 + g + t4 * LDB;
         double af[2][4], bf[2][4];
#pragma unroll
         for (int i = 0; i < 4; ++i) { af[0][i] = as[i * 8]; bf[0][i] = bs[i * 8]; }
#pragma unroll
         for (int ks = 0; ks < BK; ks += 4) {
            const int cur = (ks >> 2) & 1, nxt = cur ^ 1;
            if (ks + 4 < BK) {
#pragma unroll
               for (int i = 0; i < 4; ++i) {
                  af[nxt][i] = as[(ks + 4) * LDA + i * 8];
                  bf[nxt][i] = bs[(ks + 4) * LDB + i * 8];
               }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
               for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], af[cur][i], -bf[cur][j]);
         }
         __syncthreads();
      }
      // epilogue: store only
      if (full && below) {
#pragma unroll
         for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
               const int c = c0 + j * 8 + h;
               double* colp = (c < fn) ? (Lp + (size_t)c * ldl + r0) : (Cp + (size_t)(c - fn) * ldc + (r0 - fn));
#pragma unroll
               for (int i = 0; i < 4; ++i) colp[i * 8] = acc[i][j][h];
            }
      } else {
#pragma unroll
         for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
               const int c = c0 + j * 8 + h;
               if (c >= cend) continue;
               double* colp = (c < fn) ? (Lp + (size_t)c * ldl) : (Cp + (size_t)(c - fn) * ldc - fn);
#pragma unroll
               for (int i = 0; i < 4; ++i) {
                  const int r = r0 + i * 8;
                  if (r < m && r >= c) colp[r] = acc[i][j][h];
               }
            }
      }
      tj += gridDim.y;
   }
}
