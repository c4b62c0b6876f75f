// Dense front factorisation kernels (included by engine.cu).
//
// A front is eliminated PB = 128 columns at a time; one block step = four launches, batched over
// all fronts of a level bucket:
//   k_block   one CTA per front: close the previous block (count accepted pivots, park failed
//             columns behind the candidates), then factorise the next 128x128 diagonal block
//             entirely in shared memory: sub-panels of 32 columns, complete pivoting inside each
//             32x32 diagonal sub-block by ONE WARP holding the sub-block in registers
//             (block_ldlt.hxx:290-414), forward substitution + a-posteriori test for the rows of
//             the block below it, FP64 FMA update of the rest of the block.
//   k_apply   rows below the block: X = A21 P L11^{-T} by forward substitution (one thread per
//             row, 32 columns in registers at a time), L21 = X D^{-1} written IN PLACE, X = L21 D
//             kept in Wld for the Schur update, a-posteriori test |l_ij| <= 1/u
//             (ldlt_app.cxx:304-323) by warp shuffle + atomicMin.
//   k_commit  the block's own rows, D, pivot lists, row permutation of finished columns; restores
//             failed columns from the saved originals (CopyBackup, ldlt_app.cxx:423).
//   k_update  C -= L (LD)^T on FP64 tensor cores (mma.sync m8n8k4 -> DMMA), lower tiles only,
//             k <= 128.
#pragma once

constexpr int TM = 128;          // update tile rows
constexpr int TN = 64;           // update tile cols
constexpr int BK = 32;           // k chunk staged in shared memory
constexpr int LDA = TM + 4;      // padded smem leading dimensions: t4*LD + g distinct mod 16 per half warp
constexpr int LDB = TN + 4;
constexpr int XLD = 100;         // leading dimension of the k-major in-block buffers (rows <= 96), = 4 mod 16

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


// ------------------------------------------------------------------------------------------
// Complete pivoting on a symmetric block of order wj <= 32 by one warp.  Lane r holds row r of the
// (full) block in registers; rows and columns are never moved: eliminated indices leave the
// `alive` mask and the pivot sequence is recorded.  A pivot step = one packed-key max reduction
// (magnitude, column, row in 32 bits), the pivot row(s) broadcast through shared memory, 32 FMAs.
// Outputs (shared): spord[k] original index of pivot k, sptype[k] (0 = 1x1, 1/2 = first/second of
// a 2x2), sd = D^{-1} in SSIDS's format, sdd = D itself (for L*D), sLc[r + k*LDS] = multiplier of
// original row r for pivot k.  Returns 0 or an SSIDS error flag.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_zero_pivots(unsigned& alive, int& np, int lane, double* sLc, double* sd, double* sdd,
                                                int* spord, int* sptype) {
   // everything left is numerically zero: record zero pivots (block_ldlt.hxx:303-318)
   while (alive) {
      const int q = __ffs(alive) - 1;
      if (lane == 0) { spord[np] = q; sptype[np] = 0; sd[2 * np] = 0.0; sd[2 * np + 1] = 0.0; sdd[2 * np] = 0.0; sdd[2 * np + 1] = 0.0; }
      sLc[lane + np * LDS] = 0.0;
      alive &= ~(1u << q);
      ++np;
   }
}

// reciprocal without the special-case branches of the IEEE division: MUFU seed + three Newton steps
// (1-2 ulp; the value is stored as D^{-1} and used consistently, so the factorisation stays exact for it)
__device__ __forceinline__ double fast_rcp(double x) {
   double r;
   asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
   r = fma(fma(-x, r, 1.0), r, r);
   r = fma(fma(-x, r, 1.0), r, r);
   r = fma(fma(-x, r, 1.0), r, r);
   return r;
}
__device__ __forceinline__ void bar128() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ double sel8(const double (&a)[8], int i) {
   const double s0 = (i & 1) ? a[1] : a[0], s1 = (i & 1) ? a[3] : a[2], s2 = (i & 1) ? a[5] : a[4], s3 = (i & 1) ? a[7] : a[6];
   const double t0 = (i & 2) ? s1 : s0, t1 = (i & 2) ? s3 : s2;
   return (i & 4) ? t1 : t0;
}

// Pivoting on a symmetric block of order wj <= 32 by FOUR warps (threads 0..127 of the CTA, named
// barrier 1).  Thread (wq, lane) holds A(lane, 8*wq .. 8*wq+7) in registers, so a pivot step is 8 FMAs
// per thread; every warp keeps its own copy of the diagonal (and of its reciprocal), so that the
// pivot candidate is found without talking to the other warps.  Rows and columns are never moved:
// eliminated indices leave the `alive` mask and the pivot sequence is recorded.
//   fast step   the largest diagonal entry as a 1x1 pivot when all multipliers of the block stay
//               <= 2 in modulus (checked on the multipliers by a vote): one warp reduction, one
//               shared-memory broadcast of the pivot column, one barrier
//   general     complete pivoting with 1x1 / 2x2 pivots exactly as block_ldlt.hxx:290-414
// Outputs (shared): spord[k] original index of pivot k, sptype[k] (0 = 1x1, 1/2 = first/second of
// a 2x2), sd = D^{-1} in SSIDS's format, sdd = D itself (diagonal, off-diagonal), sLc[r + k*LDS] =
// multiplier of original row r for pivot k.  Returns 0 or an SSIDS error flag (uniform).
// sv: 2*NB + 2 doubles (two broadcast columns, reciprocal and value of the pivot); skey: 8 ints.
__device__ __forceinline__ int ldlt32_4w(double (&a)[8], double dg, int wj, const Params& prm, int lane, int wq,
                                         double* svbase, unsigned* skey, double* sLc, double* sd, double* sdd,
                                         int* spord, int* sptype) {
   unsigned alive = (wj == NB) ? 0xffffffffu : ((1u << wj) - 1u);
   int np = 0;
   double rdg = 1.0 / dg;
   // a pivot must be at least `small` (and a normal number): compared on the high word
   const unsigned small_hi_pos = max((unsigned)__double2hiint(prm.small_) & 0x7fffffffu, 0x00100000u);
   const bool rec = wq == 0 && lane == 0;       // the thread that records the pivot sequence
   if (wq == 0) { sptype[lane] = 0; sd[lane] = 0.0; sd[NB + lane] = 0.0; sdd[lane] = 0.0; sdd[NB + lane] = 0.0; }
   const int cbase = 8 * wq;
   auto next_q = [&](unsigned al) -> int {
      if (prm.posdef) return al ? __ffs(al) - 1 : 0;
      const unsigned dkey = ((al >> lane) & 1u) ? ((((unsigned)__double2hiint(dg)) & 0x7fffffe0u) | (unsigned)(31 - lane)) : 0u;
      return 31 - (int)(__reduce_max_sync(0xffffffffu, dkey) & 31u);
   };
   auto zero_pivots = [&]() {
      // everything left is numerically zero: record zero pivots (block_ldlt.hxx:303-318)
      while (alive) {
         const int q = __ffs(alive) - 1;
         if (rec) { spord[np] = q; sptype[np] = 0; sd[2 * np] = 0.0; sd[2 * np + 1] = 0.0; sdd[2 * np] = 0.0; sdd[2 * np + 1] = 0.0; }
         if (wq == 0) sLc[lane + np * LDS] = 0.0;
         alive &= ~(1u << q);
         ++np;
      }
   };
   long long tk[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = 0;
#define TK(i) do { if (prm.debug) { long long t_ = clock64(); if (i > 0) tk[i] += t_ - tlast; tlast = t_; } } while (0)
   int q = next_q(alive);
   bar128();
   int it = 0;
   while (alive) {
      // the broadcast buffers alternate between steps: the barrier after this step's write also orders the
      // reads of the step before the previous one, so no barrier is needed at the end of a step
      double* const sv = svbase + (it & 1) * (2 * NB);
      ++it;
      const bool rowalive = (alive >> lane) & 1u;
      // ---- fast step
      {
         TK(0);
         if (wq == (q >> 3)) {
            sv[lane] = sel8(a, q & 7);
         }
         TK(1);
         bar128();
         TK(2);
         // every warp holds the diagonal and its reciprocal: no division between the broadcast and its use
         const double vown = sv[lane], d11 = __shfl_sync(0xffffffffu, rdg, q), aqq = __shfl_sync(0xffffffffu, dg, q);
         const unsigned alive2 = alive & ~(1u << q);
         const bool upd = (alive2 >> lane) & 1u;
         const double ldr = upd ? vown : 0.0;
         const double l = ldr * d11;
         const unsigned qh = (unsigned)__double2hiint(aqq), lh = (unsigned)__double2hiint(l) & 0x7fffffffu;
         const bool badq = prm.posdef ? ((qh >> 31) != 0u || (qh & 0x7fffffffu) < 0x00100000u) : ((qh & 0x7fffffffu) < small_hi_pos);
         const bool badl = !prm.posdef && lh > 0x40000000u;
         const bool anyb = __any_sync(0xffffffffu, badq || badl);
         TK(3);
         if (!anyb) {
            alive = alive2;
            dg -= l * ldr;
            const int qn = next_q(alive);              // overlaps the update below
            TK(4);
            rdg = fast_rcp(dg);
            const double2* v2 = reinterpret_cast<const double2*>(sv + cbase);
#pragma unroll
            for (int j2 = 0; j2 < 4; ++j2) {
               const double2 v = v2[j2];
               a[2 * j2] -= l * v.x;
               a[2 * j2 + 1] -= l * v.y;
            }
            if (wq == 0) sLc[lane + np * LDS] = l;
            if (rec) { spord[np] = q; sd[2 * np] = d11; sdd[2 * np] = aqq; }
            np += 1;
            q = qn;
            TK(5);
            continue;
         }
         if (prm.posdef) return SSIDS_ERROR_NOT_POS_DEF;
      }
      // ---- general step: complete pivoting (block_ldlt.hxx:290-414)
      int q1, q2 = -1, m1, t1;
      double d11 = 0.0, d21 = 0.0, d22 = 0.0, a11, a21 = 0.0, a22 = 0.0;
      {
         // largest |a(r,c)| over the alive rows / columns, compared on the high word of the double
         // (sign cleared, 15 mantissa bits, the low 5 bits carry 31-col); ties -> smallest column,
         // then smallest row
         unsigned kb = 0u;
#pragma unroll
         for (int j = 0; j < 8; ++j) {
            const int c = cbase + j;
            const unsigned k = ((alive >> c) & 1u) ? ((((unsigned)__double2hiint(a[j])) & 0x7fffffe0u) | (unsigned)(31 - c)) : 0u;
            kb = max(kb, k);
         }
         const unsigned bk = rowalive ? kb : 0u;
         const int bc = 31 - (int)(bk & 31u);
         const unsigned wmk = __reduce_max_sync(0xffffffffu, bk & ~31u);
         if (lane == 0) skey[wq] = wmk;
         bar128();
         const unsigned mk = max(max(skey[0], skey[1]), max(skey[2], skey[3]));
         if ((mk >> 20) == 0u) {
            if (!prm.action) return SSIDS_ERROR_SINGULAR;
            zero_pivots();
            break;
         }
         const int lo = min(lane, bc), hi = max(lane, bc);
         const unsigned idx = (rowalive && (bk & ~31u) == mk) ? (unsigned)(lo * 32 + hi) : 0xffffu;
         const unsigned wmi = __reduce_min_sync(0xffffffffu, idx);
         if (lane == 0) skey[4 + wq] = wmi;
         bar128();
         const unsigned mi = min(min(skey[4], skey[5]), min(skey[6], skey[7]));
         m1 = (int)(mi >> 5);
         t1 = (int)(mi & 31u);
         q1 = m1;
      }
      // columns m1 and t1 of the current Schur complement -> shared
      if (wq == (m1 >> 3)) sv[lane] = sel8(a, m1 & 7);
      if (wq == (t1 >> 3)) sv[NB + lane] = sel8(a, t1 & 7);
      bar128();
      a11 = sv[m1];
      int voff = 0;
      a21 = sv[t1];
      a22 = sv[NB + t1];
      if (fabs(a21) < prm.small_) {
         // the largest entry is below `small`: all remaining pivots are zero pivots
         if (!prm.action) return SSIDS_ERROR_SINGULAR;
         zero_pivots();
         break;
      }
      if (t1 == m1) {
         d11 = 1.0 / a11;
      } else {
         const double detscale = 1.0 / fabs(a21);
         const double detpiv = (a11 * detscale) * a22 - fabs(a21);
         if (fabs(detpiv) >= fabs(a21) / 2) {
            q2 = t1;
            const double rdet = 1.0 / detpiv;
            d11 = (a22 * detscale) * rdet;
            d22 = (a11 * detscale) * rdet;
            d21 = (-a21 * detscale) * rdet;
         } else if (fabs(a11) > fabs(a22)) {
            d11 = 1.0 / a11;
         } else {
            q1 = t1; voff = NB; a11 = a22;
            d11 = 1.0 / a22;
         }
      }
      if (q2 < 0) {
         alive &= ~(1u << q1);
         const bool upd = (alive >> lane) & 1u;
         const double ldr = upd ? sv[voff + lane] : 0.0;
         const double l = ldr * d11;
         if (wq == 0) sLc[lane + np * LDS] = l;
         const double2* v2 = reinterpret_cast<const double2*>(sv + voff + cbase);
#pragma unroll
         for (int j2 = 0; j2 < 4; ++j2) {
            const double2 v = v2[j2];
            a[2 * j2] -= l * v.x;
            a[2 * j2 + 1] -= l * v.y;
         }
         dg -= l * ldr;
         if (rec) { spord[np] = q1; sptype[np] = 0; sd[2 * np] = d11; sd[2 * np + 1] = 0.0; sdd[2 * np] = a11; sdd[2 * np + 1] = 0.0; }
         np += 1;
      } else {
         alive &= ~((1u << q1) | (1u << q2));
         const bool upd = (alive >> lane) & 1u;
         const double w1 = upd ? sv[lane] : 0.0, w2 = upd ? sv[NB + lane] : 0.0;
         const double l1 = d11 * w1 + d21 * w2, l2 = d21 * w1 + d22 * w2;
         if (wq == 0) { sLc[lane + np * LDS] = l1; sLc[lane + (np + 1) * LDS] = l2; }
         const double2* v1p = reinterpret_cast<const double2*>(sv + cbase);
         const double2* v2p = reinterpret_cast<const double2*>(sv + NB + cbase);
#pragma unroll
         for (int j2 = 0; j2 < 4; ++j2) {
            const double2 v1 = v1p[j2], v2 = v2p[j2];
            a[2 * j2] -= l1 * v1.x + l2 * v2.x;
            a[2 * j2 + 1] -= l1 * v1.y + l2 * v2.y;
         }
         dg -= l1 * w1 + l2 * w2;
         if (rec) {
            spord[np] = q1; spord[np + 1] = q2;
            sptype[np] = 1; sptype[np + 1] = 2;
            sd[2 * np] = d11; sd[2 * np + 1] = d21; sd[2 * np + 2] = DINF; sd[2 * np + 3] = d22;
            sdd[2 * np] = a11; sdd[2 * np + 1] = a21; sdd[2 * np + 2] = a22; sdd[2 * np + 3] = a21;
         }
         np += 2;
      }
      rdg = 1.0 / dg;
      q = next_q(alive);
   }
   if (prm.debug && lane == 0 && wq == 0 && blockIdx.x == 0) printf("[pivot4] np=%d sel+sts %lld bar %lld lds+shfl+test+vote %lld dg+redux %lld rcp+update+book %lld bar %lld\n", np, tk[1], tk[2], tk[3], tk[4], tk[5], tk[6]);
#undef TK
   return 0;
}

constexpr int LDBB = NB + 4;       // k-major 32-column operand, = 4 mod 16: conflict-free MMA fragments
__host__ __device__ inline size_t block_smem_bytes(int wcap) {
   // sA wcap x (wcap+4) | sX, sXa NB x XLD each (only when the block has more than one sub-panel; sLc
   // aliases sXa) | sLj NB x NB | sInv NB x LDBB | srow | sdk, sdd | ints
   const size_t xbuf = wcap > NB ? 2 * (size_t)NB * XLD : (size_t)NB * LDS;
   return sizeof(double) * ((size_t)wcap * (wcap + 4) + xbuf + NB * NB + NB * LDBB + 4 * NB + 2 + 2 * (2 * PB + 2)) +
          sizeof(int) * (2 * PB + 2 * NB + 16);
}

// KIND 0: blocks of at most one sub-panel (fronts with <= 32 candidate columns, the bulk of the tree): no
// in-block solve / update code, half the registers, two CTAs per SM.  KIND 1: two sub-panels, register
// budget of two CTAs per SM (the many 33..64-column fronts of banded / KKT problems).
// Programmatic dependent launch: a kernel launched with the stream-serialisation attribute becomes resident while its
// predecessor in the stream is still running and waits here for it to complete (memory included); it lets its own
// successor do the same at once.  Without the attribute both instructions do nothing.
__device__ __forceinline__ void pdl_enter() {
   asm volatile("griddepcontrol.wait;" ::: "memory");
   asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// the two halves on their own (solve kernels): let the successor become resident at once, do everything that does not
// depend on the predecessor (staging L), then wait
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <int KIND>      // 0: blocks of <= 32 columns, 1: <= 64 columns (both two CTAs per SM), 2: up to 128 columns
__global__ void __launch_bounds__(256, KIND < 2 ? 2 : 1) k_block(FrontDev* fr, const int* list, int slot0, Params prm, int finish_only,
                                                              int wcap, int par, int la) {
   constexpr bool SMALL = KIND == 0;
   FrontDev& f = fr[front_slot(list, slot0)];
   // the front record is read through shared memory: one round trip instead of a chain of dependent loads
   __shared__ FrontDev sf;
   static_assert(sizeof(FrontDev) % 8 == 0, "FrontDev is copied in 8-byte words");
   if (threadIdx.x < sizeof(FrontDev) / 8) reinterpret_cast<double*>(&sf)[threadIdx.x] = reinterpret_cast<const double*>(&f)[threadIdx.x];
   __syncthreads();
   if (sf.flag < 0) return;
   // la bit 1: the front belongs to k_front on this pass (a front it rolled back waits for the retry rounds)
   if (sf.fused && (la & 2)) return;
   // look-ahead schedule (la bit 0): a front that saw a failed pivot waits for the synchronous retry rounds, because
   // parking failed columns moves entries that a concurrent bulk update may be working on
   if (sf.stall && (la & 1)) {
      if (threadIdx.x == 0) f.urec[par][1] = 0;   // this parity's record belongs to a block that is already applied
      return;
   }
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const int g = lane >> 2, t4 = lane & 3;
   extern __shared__ __align__(16) double smem[];
   const int sld = wcap + 4;
   double* sA = smem;                        // element (r,c) of the block at sA[r + c*sld]
   double* sX = sA + (size_t)wcap * sld;     // X(r1+i, k) of the current sub-panel at sX[k*XLD + i]
   double* sXa = wcap > NB ? sX + NB * XLD : sX;   // permuted rows below the sub-panel, k-major
   double* sLc = sXa;                        // multipliers of the pivot loop (dead before sXa is filled)
   double* sLj = sXa + (wcap > NB ? NB * XLD : NB * LDS);   // L_jj in pivot order, sLj[i + k*NB]
   double* sInv = sLj + NB * NB;             // sInv[k*LDBB + c] = inv(L_jj)(c,k)
   double* srow = sInv + NB * LDBB;
   double* sdk = srow + 4 * NB + 2;          // D^{-1} of the block (SSIDS format)
   double* sdd = sdk + 2 * PB + 2;           // D of the block: diagonal, off-diagonal
   int* slperm = reinterpret_cast<int*>(sdd + 2 * PB + 2);
   int* sptype = slperm + PB;
   int* spord = sptype + PB;
   int* sptj = spord + NB;
   int* smisc = sptj + NB;                   // 0: error flag, 1: first failed column of the sub-panel, 2: moved

   const bool dbg = prm.debug && blockIdx.x == 0 && tid == 0;
   long long tq[16];
   int ntq = 0;
#define B200_TICK() do { if (dbg && ntq < 16) tq[ntq++] = clock64(); } while (0)
   B200_TICK();
   // ---- close the previous block
   int done = sf.done, gend = sf.gend, w = sf.w;
   if (w > 0) {
      const int nelim = compute_nelim(sf.fail, w, sf.p2mask);
      const int nfail = w - nelim;
      if (prm.posdef && nfail > 0) {
         __syncthreads();
         if (tid == 0) { f.flag = SSIDS_ERROR_NOT_POS_DEF; f.w = 0; }
         return;
      }
      if (nfail > 0 && (la & 1)) {
         __syncthreads();
         if (tid == 0) { f.stall = 1; f.urec[par][1] = 0; }
         return;
      }
      done += nelim;
      for (int q = 0; q < nfail; ++q) {
         const int a = done + q, b = gend - 1 - q;
         if (a < b) block_sym_swap(f, a, b);
      }
      gend -= nfail;
   }
   // everything the rest of the kernel needs from the front record (the stores below may alias it)
   const size_t ldl = sf.ldl;
   const int wc = sf.wc;
   double* const fL = sf.L;
   double* const A0g = sf.dbuf + (size_t)PB * wc;
   double* const dkg = sf.dbuf + 2 * (size_t)PB * wc;
   double* const Linvg = dkg + 2 * PB + 2;
   int* const ibuf = sf.ibuf;
   __syncthreads();
   w = finish_only ? 0 : max(0, min(wcap, gend - done));   // wcap <= PB: the launch's shared-memory capacity
   if (tid == 0) {
      f.stall = 0;
      f.urec[par][1] = 0;            // no trailing update for this parity until k_commit says so
      f.done = done; f.gend = gend; f.w = w; f.fail = w; f.nelim = done;
      f.p2mask[0] = f.p2mask[1] = f.p2mask[2] = f.p2mask[3] = 0u;
   }
   if (w == 0) return;

   // ---- load the lower triangle (asynchronous copies), then keep a copy for the restore path
   {
      const double* Lb = fL + (size_t)done * ldl + done;
      for (int c = warp; c < w; c += 8)
         for (int r = lane + (c & ~31); r < w; r += 32)
            if (r >= c) cp_async8(sA + r + c * sld, Lb + (size_t)c * ldl + r, true);
      cp_async_commit();
      if (tid < 8) smisc[tid] = 0;
      cp_async_wait<0>();
      __syncthreads();
      for (int c = warp; c < w; c += 8)
         for (int r = lane + (c & ~31); r < w; r += 32)
            if (r >= c) A0g[r + c * PB] = sA[r + c * sld];
   }
   B200_TICK();   // 1: closed + loaded

   int fail_blk = w;
   unsigned anymoved = 0;
   for (int c0 = 0; c0 < w; c0 += NB) {
      const int wj = min(NB, w - c0), r1 = c0 + NB;
      const int nr = max(0, w - r1);            // rows of the block below the sub-panel
      // a. pivoting on the diagonal sub-block (warps 0-3)
      if (warp < 4) {
         double a[8];
#pragma unroll
         for (int j = 0; j < 8; ++j) {
            const int c = 8 * warp + j;
            double v = 0.0;
            if (c < wj && lane < wj) v = (c <= lane) ? sA[(c0 + lane) + (c0 + c) * sld] : sA[(c0 + c) + (c0 + lane) * sld];
            a[j] = v;
         }
         const double dg = lane < wj ? sA[(c0 + lane) + (c0 + lane) * sld] : 1.0;
         const int err = ldlt32_4w(a, dg, wj, prm, lane, warp, srow, reinterpret_cast<unsigned*>(smisc + 8), sLc,
                                   sdk + 2 * c0, sdd + 2 * c0, spord, sptj);
         bar128();
         if (warp == 0) {
            const bool mv = lane < wj && spord[lane] != lane;
            const unsigned pm = __ballot_sync(0xffffffffu, mv);
            if (lane == 0) { smisc[0] = err; smisc[1] = wj; smisc[2] = pm != 0u; }
         }
      }
      __syncthreads();
      if (c0 == 0) B200_TICK();   // 2: first sub-panel pivoted
      if (smisc[0] != 0) {
         if (tid == 0) { f.flag = smisc[0]; f.w = 0; }
         return;
      }
      const bool moved = smisc[2] != 0;
      anymoved |= moved ? 1u : 0u;
      // b. rows of the sub-block in the earlier columns of the block follow the pivot order; L_jj
      {
         double tmp[KIND == 0 ? 1 : (KIND == 1 ? NB / 8 : (PB - NB) / 8)];
         if (!SMALL && moved) {
#pragma unroll
            for (int s = 0; s < (int)(sizeof(tmp) / sizeof(double)); ++s) {
               const int c = warp + 8 * s;
               tmp[s] = (c < c0 && lane < wj) ? sA[(c0 + spord[lane]) + c * sld] : 0.0;
            }
         }
         double lv[NB / 8];
#pragma unroll
         for (int s = 0; s < NB / 8; ++s) {
            const int k = warp + 8 * s;
            lv[s] = 0.0;
            if (lane < wj && k < wj) lv[s] = lane > k ? sLc[spord[lane] + k * LDS] : (lane == k ? 1.0 : 0.0);
         }
         __syncthreads();          // sLc (aliased by sXa) and the old rows have been read
         if (!SMALL && moved) {
#pragma unroll
            for (int s = 0; s < (int)(sizeof(tmp) / sizeof(double)); ++s) {
               const int c = warp + 8 * s;
               if (c < c0 && lane < wj) sA[(c0 + lane) + c * sld] = tmp[s];
            }
         }
#pragma unroll
         for (int s = 0; s < NB / 8; ++s) {
            const int k = warp + 8 * s;
            sLj[lane + k * NB] = lv[s];
            sA[(c0 + lane) + (c0 + k) * sld] = lv[s];
         }
         if (tid < NB) {
            slperm[c0 + tid] = c0 + (tid < wj ? spord[tid] : tid);
            sptype[c0 + tid] = tid < wj ? sptj[tid] : 0;
         }
         // rows below the sub-panel, columns in pivot order, k-major: the A operand of X = A inv(L_jj)^T
         for (int k = warp; k < NB; k += 8) {
            const int src = c0 + spord[k];
            for (int i = lane; i < ((nr + 7) & ~7); i += 32) sXa[k * XLD + i] = i < nr ? sA[(r1 + i) + src * sld] : 0.0;
         }
      }
      __syncthreads();
      if (c0 == 0) B200_TICK();   // 3: phase b
      // b2. inverse of L_jj (one warp): needed by k_apply and by the rows below
      if (warp == 0) {
         double z[NB];
#pragma unroll
         for (int i = 0; i < NB; ++i) z[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
         for (int k = 0; k < NB - 1; ++k) {
            const double zk = z[k];
#pragma unroll
            for (int i = k + 1; i < NB; ++i) z[i] -= zk * sLj[i + k * NB];
         }
         // lane c holds column c of the inverse: row c of the k-major operand
         double2* d2 = reinterpret_cast<double2*>(Linvg + (size_t)(c0 / NB) * NB * NB + lane * NB);
         double2* s2 = reinterpret_cast<double2*>(sInv + lane * LDBB);
#pragma unroll
         for (int i2 = 0; i2 < NB / 2; ++i2) {
            const double2 v = make_double2(z[2 * i2], z[2 * i2 + 1]);
            d2[i2] = v;
            s2[i2] = v;
         }
      }
      if (SMALL || r1 >= w) break;
      __syncthreads();
      if (c0 == 0) B200_TICK();   // 4: inverse
      if constexpr (!SMALL) {
      // c. rows below: X = A inv(L_jj)^T on the tensor pipe.  Warp = column tile ct (8 columns) x
      //    every second row tile; inv(L_jj)(c,k) = 0 for k > c bounds the k loop.
      {
         const int ct = warp & 3, rt0 = warp >> 2, nrt = (nr + 7) >> 3;
         double acc[6][2];
#pragma unroll
         for (int i = 0; i < 6; ++i) acc[i][0] = acc[i][1] = 0.0;
         const double* as = sXa + t4 * XLD + rt0 * 8 + g;
         const double* bs = sInv + t4 * LDBB + ct * 8 + g;
         const int kend = 8 * (ct + 1);
         for (int ks = 0; ks < kend; ks += 4) {
            const double b = bs[ks * LDBB];
#pragma unroll
            for (int i = 0; i < 6; ++i)
               if (rt0 + 2 * i < nrt) dmma(acc[i][0], acc[i][1], as[ks * XLD + 16 * i], b);
         }
#pragma unroll
         for (int i = 0; i < 6; ++i)
            if (rt0 + 2 * i < nrt) {
               sX[(ct * 8 + 2 * t4) * XLD + (rt0 + 2 * i) * 8 + g] = acc[i][0];
               sX[(ct * 8 + 2 * t4 + 1) * XLD + (rt0 + 2 * i) * 8 + g] = acc[i][1];
            }
      }
      __syncthreads();
      // c2. L = X D^{-1}, a-posteriori test; thread = (row, half of the columns)
      {
         const int i = tid & 127, kq = tid >> 7;
         if (i < nr) {
            int myfail = NB;
#pragma unroll
            for (int q = NB / 2 - 1; q >= 0; --q) {
               const int k = kq * (NB / 2) + q;
               const double xv = sX[k * XLD + i];
               const int pt = sptj[k];
               double l;
               if (pt == 0) {
                  const double d11 = sdk[2 * (c0 + k)];
                  if (d11 == 0.0) l = (fabs(xv) < prm.small_) ? 0.0 : DINF * xv;
                  else l = xv * d11;
               } else if (pt == 1) {
                  l = sdk[2 * (c0 + k)] * xv + sdk[2 * (c0 + k) + 1] * sX[(k + 1) * XLD + i];
               } else {
                  l = sdk[2 * (c0 + k) - 1] * sX[(k - 1) * XLD + i] + sdk[2 * (c0 + k) + 1] * xv;
               }
               if (!prm.posdef && !(fabs(l) <= prm.lim)) myfail = min(myfail, k);
               sA[(r1 + i) + (c0 + k) * sld] = l;
            }
            if (myfail < NB) atomicMin(&smisc[1], myfail);
         }
      }
      __syncthreads();
      if (c0 == 0) B200_TICK();   // 5: phase c
      if (smisc[1] < NB) {
         // a column failed inside the block: the block ends with this sub-panel; the columns after
         // it stay candidates and are treated as ordinary rows by k_apply
         fail_blk = c0 + smisc[1];
         w = r1;
         break;
      }
      // e. update the rest of the block on the tensor pipe: A(r,c) -= sum_k L(r,k) X(c,k); work items
      //    are strips of 8 rows x 32 columns that touch the lower triangle, dealt round robin to warps
      {
         const int nrt = (nr + 7) >> 3;
         int item = 0;
         for (int st = 0; st < nrt; ++st)
            for (int cb = 0; cb <= ((8 * st + 7) >> 5); ++cb, ++item) {
               if ((item & 7) != warp) continue;
               double acc[4][2];
#pragma unroll
               for (int jt = 0; jt < 4; ++jt) acc[jt][0] = acc[jt][1] = 0.0;
               const double* as = sA + (r1 + 8 * st + g) + (size_t)(c0 + t4) * sld;
               const double* bs = sX + t4 * XLD + cb * 32 + g;
#pragma unroll
               for (int ks = 0; ks < NB; ks += 4) {
                  const double av = as[ks * sld];
#pragma unroll
                  for (int jt = 0; jt < 4; ++jt) dmma(acc[jt][0], acc[jt][1], av, bs[ks * XLD + jt * 8]);
               }
               double* cs = sA + (r1 + 8 * st + g) + (size_t)(r1 + cb * 32 + 2 * t4) * sld;
#pragma unroll
               for (int jt = 0; jt < 4; ++jt) {
                  cs[(jt * 8) * sld] -= acc[jt][0];
                  cs[(jt * 8 + 1) * sld] -= acc[jt][1];
               }
            }
      }
      __syncthreads();
      if (c0 == 0) B200_TICK();   // 6: phase e
      }
   }
   __syncthreads();
   B200_TICK();      // all sub-panels

   // ---- publish: L11 (unit lower, pivot order) straight into the front, D^{-1}, pivot lists.  The
   // original block stays in A0 for the columns k_commit may have to restore.
   {
      double* Lb = fL + (size_t)done * ldl + done;
      for (int k = warp; k < w; k += 8)
         for (int i = k + lane; i < w; i += 32) Lb[(size_t)k * ldl + i] = i == k ? 1.0 : sA[i + k * sld];
      for (int i = tid; i < 2 * w + 2; i += 256) dkg[i] = i < 2 * w ? sdk[i] : 0.0;
      if (tid < w) { ibuf[tid] = slperm[tid]; ibuf[PB + tid] = sptype[tid]; }
      if (tid < 4) {
         unsigned mask = 0u;
         for (int b = 0; b < 32; ++b) { const int k = tid * 32 + b; if (k < w && sptype[k] == 2) mask |= 1u << b; }
         f.p2mask[tid] = mask;
      }
      if (tid == 0) { ibuf[2 * PB] = (int)anymoved; f.w = w; f.fail = fail_blk; }
   }
   B200_TICK();      // published
   if (dbg) {
      printf("[k_block] w=%d ticks:", w);
      for (int i = 1; i < ntq; ++i) printf(" %lld", tq[i] - tq[i - 1]);
      printf("\n");
   }
#undef B200_TICK
}

// ------------------------------------------------------------------------------------------
// K2: rows below the block, X L11^T = A21 P as a blocked triangular solve on the FP64 tensor pipe.
// CTA = 64 rows x the whole block (w <= 128 columns), 8 warps.  Stage j (32 columns):
//    Y   = A_j - sum_{i<j} X_i L_ji^T        DMMA, X_i and the L11 blocks in shared memory
//    X_j = Y inv(L_jj)^T                      DMMA with the explicit 32x32 inverses from k_block
//    L_j = X_j D^{-1} (in place), X_j kept in Wld, a-posteriori test |l| <= 1/u
// ------------------------------------------------------------------------------------------

constexpr int AR = 64;             // rows per CTA in k_apply
constexpr int LDX = AR + 4;        // k-major smem leading dimension, = 4 mod 16: conflict-free fragments
__host__ __device__ inline size_t apply_smem_bytes(int wcap) {
   return sizeof(double) * ((size_t)wcap * LDX + NB * LDX + 2 * PB + 2) + sizeof(int) * 2 * PB;
}

__global__ void __launch_bounds__(256, 2) k_apply(FrontDev* fr, const int* list, int slot0, Params prm, int wcap) {
   FrontDev& f = fr[front_slot(list, slot0)];
   const int w = f.w;
   if (w == 0 || f.flag < 0 || f.stall) return;
   double* const Wldp = f.Wld;     // X ring: pivot p of the front in column p mod xw
   const int xw = f.xw;
   const int done = f.done, m = f.m;
   // a CTA walks row tiles blockIdx.y, blockIdx.y + gridDim.y, ...: the L11 blocks and inverses are staged once, the next
   // tile's rows are pulled towards L2 while the current tile is solved
   const int ntile = (m - (done + w) + AR - 1) / AR;
   if ((int)blockIdx.y >= ntile) return;
   extern __shared__ __align__(16) double smem[];
   double* Xs = smem;                                    // Xs[k*LDX + r]: A, then X, k < wcap
   double* Ys = Xs + (size_t)wcap * LDX;                 // Ys[k*LDX + r], k < NB
   double* sdk = Ys + NB * LDX;
   int* slperm = reinterpret_cast<int*>(sdk + 2 * PB + 2);
   int* sptype = slperm + PB;
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const int nb = (w + NB - 1) / NB, wr = nb * NB;
   {
      const double* dkg = f.dbuf + 2 * (size_t)PB * f.wc;
      for (int i = tid; i < 2 * w + 2; i += 256) sdk[i] = dkg[i];
      for (int i = tid; i < wr; i += 256) { slperm[i] = i < w ? f.ibuf[i] : 0; sptype[i] = i < w ? f.ibuf[PB + i] : 0; }
   }
   __syncthreads();
   const size_t ldl = f.ldl;
   double* const Lp = f.L;
   const double* const L11 = Lp + (size_t)done * ldl + done;
   const double* const Linv = f.dbuf + 2 * (size_t)PB * f.wc + 2 * PB + 2;
   double* const Wo = f.Wo;
   int myfail = w;
   for (int tile = blockIdx.y; tile < ntile; tile += gridDim.y) {
   const int r0 = done + w + tile * AR;
   if (tile != (int)blockIdx.y) __syncthreads();      // the previous tile's last pass still reads Xs
   {
      // A tile (8-byte copies: the row offset may be odd)
      for (int e = tid; e < wr * AR; e += 256) {
         const int k = e >> 6, rr = e & (AR - 1);
         const bool ok = k < w && r0 + rr < m;
         cp_async8(Xs + k * LDX + rr, ok ? Lp + (size_t)(done + slperm[k]) * ldl + r0 + rr : Lp, ok);
      }
      if (tile + (int)gridDim.y < ntile) {
         // the next tile of this CTA towards L2: 64 rows of a column are 4 or 5 lines of 128 bytes
         const int rn = r0 + (int)gridDim.y * AR;
         for (int e = tid; e < w * 5; e += 256) {
            const int k = e / 5, q = e - k * 5;
            const int rr = min(rn + q * 16, m - 1);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(Lp + (size_t)(done + slperm[k]) * ldl + rr));
         }
      }
      cp_async_commit();
      cp_async_wait<0>();
   }
   __syncthreads();
   // keep the permuted originals for the restore path (the pointer is read once: a store through it could alias the
   // front record, and the reload after every store would serialise the loop on L2 round trips)
   for (int e = tid; e < w * AR; e += 256) {
      const int k = e >> 6, rr = e & (AR - 1);
      if (r0 + rr < m) Wo[(size_t)k * ldl + r0 + rr] = Xs[k * LDX + rr];
   }
   const int rg = warp & 1, cg = warp >> 1;      // warp tile: rows rg*32..+31, columns cg*8..+7 of the stage
   const int g = lane >> 2, t4 = lane & 3;
   const int srow_ = tid & (AR - 1), skq = tid >> 6;   // scaling pass: row, group of 8 columns
   const bool rv = r0 + srow_ < m;
   for (int j = 0; j < nb; ++j) {
      const int c0 = j * NB;
      double acc[4][2];
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4)
#pragma unroll
         for (int h = 0; h < 2; ++h) acc[i4][h] = Xs[(c0 + cg * 8 + 2 * t4 + h) * LDX + rg * 32 + i4 * 8 + g];
      // The B operands -- this warp's 8 columns of the L11 blocks (j, i), i < j, which k_block wrote into the front in
      // place, and of inv(L_jj) -- come straight from global memory as MMA fragments (one double per lane and k-step;
      // every CTA of the front reads the same few KB, L1 / L2 hits): no shared-memory copy, two CTAs per SM
      const int bc = j * NB + cg * 8 + g;               // column of the block row j this lane feeds
      const double* const lrow = L11 + bc;
      double binv[NB / 4], bl[2][NB / 4];
      // block (j, 0) and the inverse are requested before anything else of the stage, block (j, i + 1) while block
      // (j, i) is multiplied: the L2 round trips overlap the MMAs
      if (j > 0) {
#pragma unroll
         for (int q = 0; q < NB / 4; ++q) bl[0][q] = bc < w ? __ldg(lrow + (size_t)(q * 4 + t4) * ldl) : 0.0;
      }
#pragma unroll
      for (int q = 0; q < NB / 4; ++q) binv[q] = __ldg(Linv + (size_t)j * NB * NB + (q * 4 + t4) * NB + cg * 8 + g);
#pragma unroll
      for (int i = 0; i < PB / NB - 1; ++i) {
         if (i < j) {
            if (i + 1 < j) {
#pragma unroll
               for (int q = 0; q < NB / 4; ++q) bl[(i + 1) & 1][q] = bc < w ? __ldg(lrow + (size_t)((i + 1) * NB + q * 4 + t4) * ldl) : 0.0;
            }
            const double* as = Xs + (i * NB + t4) * LDX + rg * 32 + g;
#pragma unroll
            for (int ks = 0; ks < NB; ks += 4) {
               const double b = -bl[i & 1][ks >> 2];
#pragma unroll
               for (int i4 = 0; i4 < 4; ++i4) dmma(acc[i4][0], acc[i4][1], as[ks * LDX + i4 * 8], b);
            }
         }
      }
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4)
#pragma unroll
         for (int h = 0; h < 2; ++h) Ys[(cg * 8 + 2 * t4 + h) * LDX + rg * 32 + i4 * 8 + g] = acc[i4][h];
      __syncthreads();
      double x[4][2];
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) x[i4][0] = x[i4][1] = 0.0;
      {
         const double* as = Ys + t4 * LDX + rg * 32 + g;
         const int kend = 8 * (cg + 1);            // inv(L_jj)(c,k) = 0 for k > c
#pragma unroll
         for (int ks = 0; ks < NB; ks += 4) {
            if (ks < kend) {
#pragma unroll
               for (int i4 = 0; i4 < 4; ++i4) dmma(x[i4][0], x[i4][1], as[ks * LDX + i4 * 8], binv[ks >> 2]);
            }
         }
      }
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4)
#pragma unroll
         for (int h = 0; h < 2; ++h) Xs[(c0 + cg * 8 + 2 * t4 + h) * LDX + rg * 32 + i4 * 8 + g] = x[i4][h];
      __syncthreads();
      // D^{-1}, a-posteriori test, results to global (coalesced along the rows)
      if (rv) {
         const int r = r0 + srow_;
#pragma unroll
         for (int q = 7; q >= 0; --q) {
            const int kk = c0 + skq * 8 + q;
            if (kk < w) {
               const double xv = Xs[kk * LDX + srow_];
               const int pt = sptype[kk];
               double l;
               if (pt == 0) {
                  const double d11 = sdk[2 * kk];
                  if (d11 == 0.0) l = (fabs(xv) < prm.small_) ? 0.0 : DINF * xv;
                  else l = xv * d11;
               } else if (pt == 1) {
                  l = sdk[2 * kk] * xv + sdk[2 * kk + 1] * Xs[(kk + 1) * LDX + srow_];
               } else {
                  l = sdk[2 * kk - 1] * Xs[(kk - 1) * LDX + srow_] + sdk[2 * kk + 1] * xv;
               }
               if (!prm.posdef && !(fabs(l) <= prm.lim)) myfail = min(myfail, kk);
               Lp[(size_t)(done + kk) * ldl + r] = l;
               Wldp[(size_t)((done + kk) % xw) * ldl + r] = xv;
            }
         }
      }
   }
   }   // row tiles
#pragma unroll
   for (int off = 16; off > 0; off >>= 1) myfail = min(myfail, __shfl_xor_sync(0xffffffffu, myfail, off));
   if (lane == 0 && myfail < w) atomicMin(&f.fail, myfail);
}

// ------------------------------------------------------------------------------------------
// K3.  blockIdx.y == 0: the block's own rows, D, pivot lists.  Then `ncolchunks` chunks of 64
// finished columns whose rows [done, done+w) follow the pivot order, then chunks of 256 rows
// restoring failed columns.
// ------------------------------------------------------------------------------------------
// Update records (thread 0 of the first CTA).  urec[par]: this block's pivots and the end of its strip.  With deferred
// updates (G > 1 blocks per group, look-ahead chunks) the block's own update stops at the end of the group's columns and
// the columns behind them wait for the group's last block, whose record brec[gpar] covers the pivots of the whole group;
// until then `pend` says what is owed (a front that stalls on a failed pivot is paid by the chunk's closing flush).
__global__ void __launch_bounds__(256) k_commit(FrontDev* fr, const int* list, int slot0, int ncolchunks, int par, int gpos, int G,
                                                int gtag) {
   pdl_enter();
   FrontDev& f = fr[front_slot(list, slot0)];
   const int w = f.w;
   if (w == 0 || f.flag < 0 || f.stall) return;
   const int done = f.done, m = f.m;
   const int* lperm = f.ibuf;
   const int nelim = compute_nelim(f.fail, w, f.p2mask);
   const size_t ldl = f.ldl;
   const int y = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   if (y == 0) {
      // what the trailing updates of this block need: first pivot column, accepted pivots, end of the strip
      if (tid == 0) {
         const int kend = done + nelim;
         const int gk0 = gpos == 0 ? done : f.gk0;
         if (gpos == 0) f.gk0 = done;
         const bool last = gpos == G - 1 || done + w >= f.gend;       // no further block of this group for this front
         // G == 1: one region behind the accepted pivots (failed columns included), whose first columns -- the next
         // block's -- form the part the next block waits for
         const int c1 = G == 1 ? kend : min(m, max(done + w, gk0 + G * PB));
         const int gse = G == 1 ? min(m, done + w + PB) : min(m, c1 + G * PB);
         f.urec[par][0] = done; f.urec[par][2] = c1; f.urec[par][1] = nelim;
         if (g_upd_debug)
            printf("[k_commit] front %d m %d n %d done %d w %d nelim %d gend %d gpos %d G %d tag %d gk0 %d c1 %d last %d\n", f.slot, m, f.n, done, w,
                   nelim, f.gend, gpos, G, gtag, gk0, c1, (int)last);
         if (last) {
            const int gp = (gtag - 1) & 1;
            f.brec[gp][0] = gk0; f.brec[gp][1] = kend; f.brec[gp][2] = c1; f.brec[gp][3] = gse; f.brec[gp][4] = gtag;
            f.pend[3] = 0;
         } else {
            f.pend[0] = gk0; f.pend[1] = kend; f.pend[2] = c1; f.pend[3] = 1;
         }
      }
      double* const Wldp = f.Wld;
      const int xw = f.xw;
      const double* dk = f.dbuf + 2 * (size_t)PB * f.wc;
      if (nelim < w) {
         // failed columns get their original values back (block part from A0 in pivot order; CopyBackup,
         // ldlt_app.cxx:423) ...
         const double* A0 = f.dbuf + (size_t)PB * f.wc;
         for (int k = nelim + warp; k < w; k += 8) {
            const int pk = lperm[k];
            for (int i = k + lane; i < w; i += 32) {
               const int pi = lperm[i];
               f.L[(size_t)(done + k) * ldl + done + i] = A0[max(pi, pk) + (size_t)min(pi, pk) * PB];
            }
         }
         // ... and the trailing update needs X = L D for their rows (calcLD, calc_ld.hxx:42-118)
         const int* ptype = f.ibuf + PB;
         for (int k = warp; k < nelim; k += 8) {
            const int pt = ptype[k];
            double d0, d1;          // X(i,k) = L(i,k) d0 + L(i,partner) d1
            int kn = k;
            if (pt == 0) {
               d0 = dk[2 * k] != 0.0 ? 1.0 / dk[2 * k] : 0.0; d1 = 0.0;
            } else {
               const int kf = pt == 1 ? k : k - 1;
               const double i11 = dk[2 * kf], i21 = dk[2 * kf + 1], i22 = dk[2 * kf + 3];
               const double det = i11 * i22 - i21 * i21;
               d0 = (pt == 1 ? i22 : i11) / det;
               d1 = -i21 / det;
               kn = pt == 1 ? k + 1 : k - 1;
            }
            for (int i = nelim + lane; i < w; i += 32) {
               double x = f.L[(size_t)(done + k) * ldl + done + i] * d0;
               if (pt != 0) x += f.L[(size_t)(done + kn) * ldl + done + i] * d1;
               Wldp[(size_t)((done + k) % xw) * ldl + done + i] = x;
            }
         }
      }
      for (int k = tid; k < nelim; k += 256) {
         f.D[2 * (done + k)] = dk[2 * k];
         f.D[2 * (done + k) + 1] = dk[2 * k + 1];
      }
      if (f.ibuf[2 * PB]) {
         int* arr = tid < 128 ? f.perm : f.orig;
         const int q = tid & 127;
         const int v = q < w ? arr[done + lperm[q]] : 0;
         __syncthreads();
         if (q < w) arr[done + q] = v;
      }
   } else if (y <= ncolchunks) {
      if (!f.ibuf[2 * PB]) return;
      // rows [done, done+w) of 8 finished columns per warp follow the pivot order: all loads, then all stores
      const int cb = (y - 1) * 64 + warp * 8;
      if (cb >= done) return;
      int lp[PB / 32];
#pragma unroll
      for (int q = 0; q < PB / 32; ++q) { const int i = lane + 32 * q; lp[q] = i < w ? lperm[i] : 0; }
      double tmp[8][PB / 32];
#pragma unroll
      for (int s = 0; s < 8; ++s) {
         const double* col = f.L + (size_t)(cb + s) * ldl + done;
#pragma unroll
         for (int q = 0; q < PB / 32; ++q) tmp[s][q] = (cb + s < done && lane + 32 * q < w) ? col[lp[q]] : 0.0;
      }
      __syncwarp();
#pragma unroll
      for (int s = 0; s < 8; ++s) {
         double* col = f.L + (size_t)(cb + s) * ldl + done;
#pragma unroll
         for (int q = 0; q < PB / 32; ++q) if (cb + s < done && lane + 32 * q < w) col[lane + 32 * q] = tmp[s][q];
      }
   } else {
      if (nelim == w) return;
      const int r = done + w + (y - 1 - ncolchunks) * 256 + tid;
      if (r >= m) return;
      for (int k = nelim; k < w; ++k) f.L[(size_t)(done + k) * ldl + r] = f.Wo[(size_t)k * ldl + r];
   }
}

// ------------------------------------------------------------------------------------------
// K4: FP64 tensor-core Schur update of the small and medium fronts (and of the synchronous retry rounds):
// C(r,c) -= sum_k A(r,k) B(c,k) over the lower triangle of the columns behind the accepted pivots of the block
// urec[par], with A = L(:, k0:k0+kc) and B = the pivots' columns of the X ring.  The large fronts of the look-ahead
// chunks go through k_upd3 (upd_kernels.cuh).
// CTA = 256 threads = 8 warps (4 x 2); CTA tile 128 x 64; warp tile 32 x 32 = 4 x 4 m8n8k4 MMAs;
// two CTAs per SM.  ncu (profiles/r1_notes.md) showed the kernel bound by per-warp instruction
// issue around the MMAs, so interior tiles take a path without bounds predicates and with
// strength-reduced addresses, and tile coordinates advance incrementally.
// Work distribution: with `map` (one int3 per CTA: front offset, part, parts) CTAs are dealt to the fronts
// of the launch in proportion to their tile counts; without it the grid is (fronts, parts).
__global__ void __launch_bounds__(256, 2) k_update(FrontDev* fr, const int* list, int slot0, const int* __restrict__ map, int par) {
   int fq, part, nparts;
   if (map) { fq = map[3 * blockIdx.x]; part = map[3 * blockIdx.x + 1]; nparts = map[3 * blockIdx.x + 2]; }
   else { fq = blockIdx.x; part = blockIdx.y; nparts = gridDim.y; }
   const FrontDev& f = fr[list ? list[fq] : slot0 + fq];
   if (f.flag < 0) return;
   // pivots accepted in the block (recorded by k_commit); the region starts right behind them, so columns
   // of the block that failed (restored by k_commit) receive the update as well
   const int kc = f.urec[par][1];
   const int m = f.m;
   const int k0 = f.urec[par][0], cbeg = k0 + kc, cend = m;
   if (kc <= 0 || cbeg >= cend) return;
   // The tile grid starts at the even index org <= cbeg: every row offset of the operands and of C is then a
   // multiple of 16 bytes (ldl, ldc are even, C is shifted by n & 1), so all operand loads are 16-byte copies.
   // When cbeg is odd the grid has one extra leading row and column, masked in the first row / column of tiles.
   const int org = cbeg & ~1;
   const int ntr = (m - org + TM - 1) / TM;
   const int ntc = (cend - org + TN - 1) / TN;
   extern __shared__ double smem[];
   double* As = smem;                       // [2][BK][LDA]
   double* Bs = smem + 2 * BK * LDA;        // [2][BK][LDB]
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const int wr = warp & 3, wc = warp >> 2;  // 4 x 2 warps; warp tile origin: rows wr*32, cols wc*32
   const int g = lane >> 2, t4 = lane & 3;
   const size_t ldl = f.ldl;
   const int fn = f.n, ldc = f.ldc, xw = f.xw;
   double* const Lp = f.L;
   double* const Cp = f.C;
   const double* Abase = Lp + (size_t)k0 * ldl;
   const double* const Xp = f.Wld;          // pivot p in column p mod xw
   const bool vec16 = true;
   const int rclamp = (int)ldl - 2;         // last even row index inside a column of L / X
   const int nchunk = (kc + BK - 1) / BK;
   // tile (ti,tj) touches the lower triangle iff tj <= 2*ti+1 (TM = 2*TN, shared origin cbeg);
   // tiles are numbered row by row; a CTA walks its tiles incrementally.  The first k-chunk of the
   // NEXT tile is requested while the last chunk of the current one is multiplied, and its C tile
   // is pulled towards L2.
   auto normalise = [&](int& ti, int& tj) -> bool {
      while (ti < ntr) { const int wd = min(ntc, 2 * ti + 2); if (tj < wd) break; tj -= wd; ++ti; }
      return ti < ntr;
   };
   auto load_stage = [&](int rbase, int cbase, bool full, int chunk, int buf) {
      const int kb = chunk * BK;
      double* as = As + buf * BK * LDA;
      double* bs = Bs + buf * BK * LDB;
      const int klim = kc - kb;                             // valid k rows of the chunk
      if (full && klim >= BK) {
         // thread -> (k row = q*4 + tid/64, piece (tid%64)*2) for A; (q*8 + tid/32, (tid%32)*2) for B
         const double* pa = Abase + (size_t)(kb + (tid >> 6)) * ldl + rbase + (tid & 63) * 2;
         double* sa_ = as + (tid >> 6) * LDA + (tid & 63) * 2;
         const size_t sta = 4 * ldl;
#pragma unroll
         for (int q = 0; q < 8; ++q) cp_async16(sa_ + q * 4 * LDA, pa + q * sta, 16);
         double* sb_ = bs + (tid >> 5) * LDB + (tid & 31) * 2;
         const int kr = k0 + kb + (tid >> 5), cc = cbase + (tid & 31) * 2;
#pragma unroll
         for (int q = 0; q < 4; ++q) cp_async16(sb_ + q * 8 * LDB, Xp + (size_t)((kr + q * 8) % xw) * ldl + cc, 16);
      } else {
         // edge tiles and short k-chunks: the same 16-byte copies with the row index clamped into the column
         // (rows / columns beyond the front are never stored, and every output element only sees its own row of
         // A and its own column of B); k-rows beyond kc are zero-filled
#pragma unroll
         for (int q = 0; q < 8; ++q) {
            const int e = q * 256 + tid, kk = e >> 6, rr = (e & 63) * 2;
            const int ra = min(rbase + rr, rclamp);
            const int nb = (kk < klim) ? 16 : 0;
            cp_async16(as + kk * LDA + rr, nb ? Abase + (size_t)(kb + kk) * ldl + ra : Abase, nb);
         }
#pragma unroll
         for (int q = 0; q < 4; ++q) {
            const int e = q * 256 + tid, kk = e >> 5, rr = (e & 31) * 2;
            const int cb = min(cbase + rr, rclamp);
            const int nb = (kk < klim) ? 16 : 0;
            cp_async16(bs + kk * LDB + rr, nb ? Xp + (size_t)((k0 + kb + kk) % xw) * ldl + cb : Xp, nb);
         }
      }
      cp_async_commit();
   };
   int ti = 0, tj = part;
   if (!normalise(ti, tj)) return;
   int buf = 0;
   {
      const int rbase = org + ti * TM, cbase = org + tj * TN;
      load_stage(rbase, cbase, (rbase + TM <= m) && (cbase + TN <= cend) && vec16, 0, 0);
   }
   for (;;) {
      const int rbase = org + ti * TM, cbase = org + tj * TN;
      const bool fullg = (rbase + TM <= m) && (cbase + TN <= cend);      // geometry: operand loads
      const bool full = fullg && rbase >= cbeg && cbase >= cbeg;         // and no masked leading row / column
      const bool below = cbase + TN - 1 <= rbase;     // tile entirely below the diagonal
      // a warp whose 32 x 32 part of the tile lies strictly above the diagonal or outside the front has nothing to compute
      const bool above = rbase + wr * 32 + 31 < cbase + wc * 32 || rbase + wr * 32 >= m || cbase + wc * 32 >= cend;   // or outside the front
      int nti = ti, ntj = tj + nparts;
      const bool has_next = normalise(nti, ntj);
      const int nrbase = org + nti * TM, ncbase = org + ntj * TN;
      double acc[4][4][2];
      // accumulators start as the C tile; the product is subtracted by feeding -B to the MMA.  The MMA computes
      // the TRANSPOSE of the tile (B-operand rows are the tile's columns), so that a thread's two accumulator
      // elements are two consecutive rows of one column: one 16-byte access instead of two 8-byte ones.
      // acc[i][j][h] = C(r0 + 8 i + h, c0 + 8 j)
      const int r0 = rbase + wr * 32 + 2 * t4, c0 = cbase + wc * 32 + g;
      if (full && below) {
#pragma unroll
         for (int j = 0; j < 4; ++j) {
            const int c = c0 + j * 8;
            const double* colp = (c < fn) ? (Lp + (size_t)c * ldl + r0) : (Cp + (size_t)(c - fn) * ldc + (r0 - fn));
#pragma unroll
            for (int i = 0; i < 4; ++i) {
               const double2 v = *reinterpret_cast<const double2*>(colp + i * 8);
               acc[i][j][0] = v.x; acc[i][j][1] = v.y;
            }
         }
      } else {
#pragma unroll
         for (int j = 0; j < 4; ++j) {
            const int c = c0 + j * 8;
            const bool cv = c < cend && c >= cbeg;
            const double* colp = (c < fn) ? (Lp + (size_t)c * ldl) : (Cp + (size_t)(c - fn) * ldc - fn);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
               for (int h = 0; h < 2; ++h) {
                  const int r = r0 + i * 8 + h;
                  acc[i][j][h] = (cv && r < m && r >= c) ? colp[r] : 0.0;
               }
         }
      }
      for (int ch = 0; ch < nchunk; ++ch) {
         if (ch + 1 < nchunk) {
            load_stage(rbase, cbase, fullg, ch + 1, buf ^ 1);
            cp_async_wait<1>();
         } else if (has_next) {
            load_stage(nrbase, ncbase, (nrbase + TM <= m) && (ncbase + TN <= cend) && vec16, 0, buf ^ 1);
            {
               // next C tile towards L2: 64 columns x 128 rows = 8 lines of 128 B per column
               const int c = ncbase + (tid >> 2), r = nrbase + (tid & 3) * 32;
               if (c < cend && r < m) {
                  const double* pp = (c < fn) ? (Lp + (size_t)c * ldl + r) : (Cp + (size_t)(c - fn) * ldc + (r - fn));
                  asm volatile("prefetch.global.L2 [%0];" ::"l"(pp));
                  asm volatile("prefetch.global.L2 [%0];" ::"l"(pp + 16));
               }
            }
            cp_async_wait<1>();
         } else {
            cp_async_wait<0>();
         }
         __syncthreads();
         const double* as = As + buf * BK * LDA + wr * 32 + g + t4 * LDA;
         const double* bs = Bs + buf * BK * LDB + wc * 32 + g + t4 * LDB;
         double af[2][4], bf[2][4];
         if (!above) {
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
               for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], -bf[cur][j], af[cur][i]);
         }
         }
         __syncthreads();
         buf ^= 1;
      }
      // epilogue: store only
      if (full && below) {
#pragma unroll
         for (int j = 0; j < 4; ++j) {
            const int c = c0 + j * 8;
            double* colp = (c < fn) ? (Lp + (size_t)c * ldl + r0) : (Cp + (size_t)(c - fn) * ldc + (r0 - fn));
#pragma unroll
            for (int i = 0; i < 4; ++i) *reinterpret_cast<double2*>(colp + i * 8) = make_double2(acc[i][j][0], acc[i][j][1]);
         }
      } else {
#pragma unroll
         for (int j = 0; j < 4; ++j) {
            const int c = c0 + j * 8;
            if (c >= cend || c < cbeg) continue;
            double* colp = (c < fn) ? (Lp + (size_t)c * ldl) : (Cp + (size_t)(c - fn) * ldc - fn);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
               for (int h = 0; h < 2; ++h) {
                  const int r = r0 + i * 8 + h;
                  if (r < m && r >= c) colp[r] = acc[i][j][h];
               }
         }
      }
      if (!has_next) break;
      ti = nti; tj = ntj;
   }
}
