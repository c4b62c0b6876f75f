// k_upd3: the trailing (Schur-complement) update of the large fronts, FP64 tensor pipe, warp-specialised.
//
//    C(r, c) -= sum_{p in [ka, kb)} L(r, p) X(c, p)        ca <= c < cb,  c <= r < m
//
// (Schur update of ldlt_app.cxx:1081-1169 `Block::update` / `form_contrib`, with calcLD fused: X = L D is what
// k_apply left in the front's X ring, slot p mod xw for pivot p.)
//
// Shape: one CTA per SM, 288 threads.  Eight consumer warps (2 x 4, warp tile 64 x 32 of m8n8k4 DMMAs, CTA tile
// 128 x 128, 64 accumulator doubles per thread) never meet at a CTA-wide barrier: each waits for a k-chunk of the
// operands on that stage's `full` mbarrier, multiplies, and releases the stage on its `empty` mbarrier.  The ninth
// warp is the producer: per k-chunk of UBK pivots it arms the stage's `full` barrier with the byte count and issues one
// bulk copy (cp.async.bulk, 1 KB) per pivot and operand -- a pivot's 128 rows of L and of X are contiguous -- into rows
// padded to LD = 132 doubles (conflict-free fragment loads); it also pulls the next C tile towards L2.  Consumers drift
// up to UNST stages apart, so one warp's tile prologue / epilogue (C tile in, C tile out) overlaps its neighbours' MMAs:
// tools/micro/upd3_bench.cu measures 27.8 TF/s at k = 128 and 32 TF/s at k = 512 on t = 6016 where the previous kernel
// (CTA-wide barriers per k-chunk) reaches 21.6 / 28.4 and cuBLAS Dgemm 29.2 / 34.3 (full square).
//
// Work: tiles of the region's lower triangle, numbered column by column from the even origin org = ca & ~1 (all operand
// rows are then 16-byte aligned); a launch handles a slice [lo, hi) of every front's tile list, a
// CTA (map entry: front, part, parts) walks tiles lo + part, lo + part + parts, ...
//
// kind 0 (STRIP): pivots of the block urec[par] on the columns behind it up to urec[par][2].
// kind 1 (BULK) : pivots of the whole group brec[par] on the columns from brec[par][2]; slice (sa, sb, sc, z):
//                 bound0 = max(tiles of the columns up to brec[par][3], T z / 256);  sa < 0: tiles [0, bound0),
//                 else [bound0 + R sa / sc, bound0 + R sb / sc) with R = T - bound0.
// kind 2 (FLUSH): what a stalled front still owes (pend), all tiles.
#pragma once

namespace upd3 {

constexpr int UTM = 128, UTN = 128, UBK = 16, UNST = 6;
constexpr int ULDA = UTM + 4, ULDB = UTN + 4;
constexpr int USTAGE = UBK * (ULDA + ULDB);                       // doubles
constexpr int SMEM = UNST * USTAGE * 8 + 2 * UNST * 8 + 64;      // bytes
constexpr int THREADS = 288;
// The launches on a front's critical path (a block's strip, the next group's columns of a deferred update) are a few
// dozen 128 x 128 tiles: one tile per CTA at the rate of ONE SM (18 us per 128 pivots).  They take 64 x 64 tiles instead
// (kt_upd3<64>): four times the CTAs, two per SM.
constexpr int STM = 64;
template <int TM> struct Shape {
   static constexpr int TN = TM, LDA = TM + 4, LDB = TM + 4;
   static constexpr int STAGE = UBK * (LDA + LDB);
   static constexpr int SMEM = UNST * STAGE * 8 + 2 * UNST * 8 + 64;
   static constexpr int RI = TM / 16, CJ = TM / 32;     // 8 x 8 MMA tiles per warp: rows, columns (warps 2 x 4)
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) {
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
   asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
   asm volatile(
       "{\n"
       ".reg .pred p;\n"
       "WAIT_%=:\n"
       "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
       "@p bra DONE_%=;\n"
       "bra WAIT_%=;\n"
       "DONE_%=:\n"
       "}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
// Operand rows are re-read by every tile of their tile row / column, the C tile is touched once: the bulk copies carry an
// L2 evict-last policy and the C accesses an evict-first one, so that the C stream does not push the operands out of L2
// (ncu on the longest launch of C2: 2.4 GB of DRAM traffic for 0.6 GB of algorithmic bytes without the hints).
__device__ __forceinline__ unsigned long long policy_evict_last() {
   unsigned long long p;
   asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
   return p;
}
__device__ __forceinline__ unsigned long long policy_evict_first() {
   unsigned long long p;
   asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
   return p;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar, unsigned long long pol) {
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
                "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                : "memory");
}
__device__ __forceinline__ double2 ld_c2(const double* p, unsigned long long pol) {
   double2 v;
   asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
   return v;
}
__device__ __forceinline__ void st_c2(double* p, double a, double b, unsigned long long pol) {
   asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(a), "d"(b), "l"(pol) : "memory");
}

// tile index -> (column tj, row ti) of a lower triangle with ntr tile rows, numbered column by column
__device__ __forceinline__ void tile_of(int idx, int ntr, int& tj, int& ti) {
   // column tj starts at s(tj) = tj ntr - tj (tj - 1) / 2
   const float b = (float)ntr + 0.5f;
   int c = (int)(b - sqrtf(fmaxf(b * b - 2.0f * (float)idx, 0.0f)));
   c = max(0, min(c, ntr - 1));
   while (c > 0 && c * ntr - c * (c - 1) / 2 > idx) --c;
   while (c + 1 < ntr && (c + 1) * ntr - (c + 1) * c / 2 <= idx) ++c;
   tj = c;
   ti = c + idx - (c * ntr - c * (c - 1) / 2);
}
__host__ __device__ __forceinline__ long long tiles_upto(int ncols, int ntr) {   // tiles in the first ncols tile columns
   const long long c = ncols < ntr ? ncols : ntr;
   return c * ntr - c * (c - 1) / 2;
}

}  // namespace upd3

template <int TM>
__global__ void __launch_bounds__(upd3::THREADS, TM == 128 ? 1 : 2) kt_upd3(FrontDev* fr, const int* list, int slot0, const int* __restrict__ map,
                                                                          int kind, int par, int tag, int sa, int sb, int sc, int z) {
   using namespace upd3;
   using SH = Shape<TM>;
   constexpr int UTM = TM, UTN = TM, ULDA = SH::LDA, ULDB = SH::LDB, USTAGE = SH::STAGE, RI = SH::RI, CJ = SH::CJ;
   pdl_enter();
   const int fq = map[3 * blockIdx.x], part = map[3 * blockIdx.x + 1], nparts = map[3 * blockIdx.x + 2];
   const FrontDev& f = fr[list ? list[fq] : slot0 + fq];
   if (f.flag < 0) return;
   // the update this launch owes the front
   int ka, kb, ca, cb, gse;
   const int m = f.m;
   if (kind == 0) {
      ka = f.urec[par][0]; kb = ka + f.urec[par][1]; ca = kb; cb = f.urec[par][2]; gse = cb;
      if (f.urec[par][1] <= 0) return;
   } else if (kind == 1 || kind == 3) {
      if (f.brec[par][4] != tag) return;              // this group never reached the front (or it stalled before)
      ka = f.brec[par][0]; kb = f.brec[par][1]; ca = f.brec[par][2]; cb = m; gse = f.brec[par][3];
      if (kind == 3) {
         // GSTRIP: the part of the group's update the next group's blocks wait for -- the 128-wide tile columns (BULK's
         // numbering) that hold the columns up to gse; BULK launches with sa == -2 start behind them
         const int org128 = ca & ~1;
         cb = min(m, org128 + (max(gse, ca) - org128 + upd3::UTN - 1) / upd3::UTN * upd3::UTN);
      }
   } else {
      if (!f.pend[3]) return;
      ka = f.pend[0]; kb = f.pend[1]; ca = f.pend[2]; cb = m; gse = m;
   }
   if (kb <= ka || ca >= cb) return;
   const int org = ca & ~1;
   const int ntr = (m - org + UTM - 1) / UTM;
   const int ntc = (cb - org + UTN - 1) / UTN;
   const int T = (int)tiles_upto(ntc, ntr);
   int lo = 0, hi = T;
   if (kind == 1) {
      const int ngs = (int)tiles_upto((max(gse, ca) - org + UTN - 1) / UTN, ntr);
      const int bound0 = min(T, max(ngs, (int)(((long long)T * z + 255) >> 8)));
      if (sa < 0) { hi = bound0; if (sa == -2) lo = min(ngs, bound0); }
      else {
         const long long R = T - bound0;
         lo = bound0 + (int)(R * sa / sc);
         hi = bound0 + (int)(R * sb / sc);
      }
   }
   if (g_upd_debug && part == 0 && threadIdx.x == 0)
      printf("[k_upd3] kind %d par %d tag %d front %d m %d n %d: pivots [%d,%d) cols [%d,%d) tiles [%d,%d) of %d parts %d\n", kind, par, tag, f.slot, m,
             f.n, ka, kb, ca, cb, lo, hi, T, nparts);
   const int first = lo + part;
   if (first >= hi) return;
   const int my_tiles = (hi - first + nparts - 1) / nparts;
   const int kc = kb - ka;
   const int nchunk = (kc + UBK - 1) / UBK;

   extern __shared__ __align__(16) double smem[];
   unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + UNST * USTAGE);
   unsigned long long* empty = full + UNST;
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   if (tid == 0) {
      for (int s = 0; s < UNST; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 8); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
   }
   __syncthreads();
   const size_t ldl = f.ldl;
   const int fn = f.n, ldc = f.ldc, xw = f.xw;
   double* const Lp = f.L;
   double* const Cp = f.C;
   const double* const Xp = f.Wld;
   const int ldl_i = (int)ldl;

   if (warp == 8) {
      // ---------------------------------------------------------------- producer
      const unsigned long long pol_keep = policy_evict_last();
      int it = 0;
      int tj, ti;
      tile_of(first, ntr, tj, ti);
      for (int tl = 0; tl < my_tiles; ++tl) {
         const int rb = org + ti * UTM, cbase = org + tj * UTN;
         // next tile (column-major walk)
         int ntj = tj, nti = ti + nparts;
         while (ntj < ntr && nti >= ntr) { nti = nti - ntr + ntj + 1; ++ntj; }
         if (tl + 1 < my_tiles) {
            // its C tile towards L2: 128 columns x 8 lines of 128 B
            const int nrb = org + nti * UTM, ncb = org + ntj * UTN;
            for (int q = lane; q < UTN * (UTM / 16); q += 32) {
               const int c = ncb + q / (UTM / 16), r = nrb + (q % (UTM / 16)) * 16;
               if (c < cb && r < m && r + 15 >= c) {
                  const double* pp = (c < fn) ? (Lp + (size_t)c * ldl + r) : (Cp + (size_t)(c - fn) * ldc + (r - fn));
                  asm volatile("prefetch.global.L2 [%0];" ::"l"(pp));
               }
            }
         }
         const unsigned bytesA = (unsigned)min(UTM, ldl_i - rb) * 8u, bytesB = (unsigned)min(UTN, ldl_i - cbase) * 8u;
         for (int ch = 0; ch < nchunk; ++ch, ++it) {
            const int s = it % UNST;
            const unsigned ph = (unsigned)(it / UNST) & 1u;
            if (it >= UNST) mbar_wait(empty + s, ph ^ 1u);
            double* as = smem + s * USTAGE;
            double* bs = as + UBK * ULDA;
            const int k0 = ch * UBK, nk = min(UBK, kc - k0);
            if (nk < UBK) {
               // short last chunk: the pivots up to the next multiple of 4 are zero rows (generic-proxy stores, ordered
               // before the consumers' reads by the barrier, and before later bulk writes by the proxy fence)
               const int nz = ((nk + 3) & ~3) - nk;
               for (int e = lane; e < nz * ULDA; e += 32) { as[nk * ULDA + e] = 0.0; bs[nk * ULDB + e] = 0.0; }
               asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            __syncwarp();
            if (lane == 0) mbar_expect_tx(full + s, (unsigned)nk * (bytesA + bytesB));
            __syncwarp();
            for (int q = lane; q < 2 * UBK; q += 32) {
               const int kk = q & (UBK - 1);
               if (kk < nk) {
                  const int p = ka + k0 + kk;
                  if (q < UBK) bulk_g2s(as + kk * ULDA, Lp + (size_t)p * ldl + rb, bytesA, full + s, pol_keep);
                  else bulk_g2s(bs + kk * ULDB, Xp + (size_t)(p % xw) * ldl + cbase, bytesB, full + s, pol_keep);
               }
            }
         }
         tj = ntj; ti = nti;
      }
      return;
   }

   // ------------------------------------------------------------------- consumers: 2 x 4 warps, warp tile TM/2 x TN/4
   const int wr = warp & 1, wc = warp >> 1, g = lane >> 2, t4 = lane & 3;
   const unsigned long long pol_stream = policy_evict_first();
   int it = 0;
   int tj, ti;
   tile_of(first, ntr, tj, ti);
   for (int tl = 0; tl < my_tiles; ++tl) {
      const int rb = org + ti * UTM, cbase = org + tj * UTN;
      // accumulators hold the transposed tile: acc[i][j][h] = C(r0 + 8 i + h, c0 + 8 j), so that a thread's pair is two
      // consecutive rows of one column (16-byte accesses)
      const int r0 = rb + wr * (UTM / 2) + 2 * t4, c0 = cbase + wc * (UTN / 4) + g;
      const int wrow = rb + wr * (UTM / 2), wcol = cbase + wc * (UTN / 4);
      // nothing to compute for a warp whose part lies strictly above the diagonal or outside the region
      const bool idle = wrow + UTM / 2 - 1 < wcol || wrow >= m || wcol >= cb || wcol + UTN / 4 - 1 < ca;
      const bool interior = ti > tj && rb + UTM <= m && cbase + UTN <= cb && cbase >= ca;
      double acc[RI][CJ][2];
      if (!idle) {
         if (interior) {
#pragma unroll
            for (int j = 0; j < CJ; ++j) {
               const int c = c0 + j * 8;
               const double* colp = (c < fn) ? (Lp + (size_t)c * ldl + r0) : (Cp + (size_t)(c - fn) * ldc + (r0 - fn));
#pragma unroll
               for (int i = 0; i < RI; ++i) {
                  const double2 v = ld_c2(colp + i * 8, pol_stream);
                  acc[i][j][0] = v.x; acc[i][j][1] = v.y;
               }
            }
         } else {
#pragma unroll
            for (int j = 0; j < CJ; ++j) {
               const int c = c0 + j * 8;
               const bool cv = c < cb && c >= ca;
               const double* colp = (c < fn) ? (Lp + (size_t)c * ldl) : (Cp + (size_t)(c - fn) * ldc - fn);
#pragma unroll
               for (int i = 0; i < RI; ++i)
#pragma unroll
                  for (int h = 0; h < 2; ++h) {
                     const int r = r0 + i * 8 + h;
                     acc[i][j][h] = (cv && r < m && r >= c) ? colp[r] : 0.0;
                  }
            }
         }
      }
      for (int ch = 0; ch < nchunk; ++ch, ++it) {
         const int s = it % UNST;
         const unsigned ph = (unsigned)(it / UNST) & 1u;
         mbar_wait(full + s, ph);
         if (!idle) {
            const double* as = smem + s * USTAGE + wr * (UTM / 2) + g + t4 * ULDA;
            const double* bs = smem + s * USTAGE + UBK * ULDA + wc * (UTN / 4) + g + t4 * ULDB;
            const int nk4 = (min(UBK, kc - ch * UBK) + 3) >> 2;
            if (nk4 == UBK / 4) {
               double af[2][RI], bf[2][CJ];
#pragma unroll
               for (int i = 0; i < RI; ++i) af[0][i] = as[i * 8];
#pragma unroll
               for (int j = 0; j < CJ; ++j) bf[0][j] = bs[j * 8];
#pragma unroll
               for (int ks = 0; ks < UBK / 4; ++ks) {
                  const int cur = ks & 1, nxt = cur ^ 1;
                  if (ks + 1 < UBK / 4) {
#pragma unroll
                     for (int i = 0; i < RI; ++i) af[nxt][i] = as[(ks + 1) * 4 * ULDA + i * 8];
#pragma unroll
                     for (int j = 0; j < CJ; ++j) bf[nxt][j] = bs[(ks + 1) * 4 * ULDB + j * 8];
                  }
#pragma unroll
                  for (int i = 0; i < RI; ++i)
#pragma unroll
                     for (int j = 0; j < CJ; ++j) dmma(acc[i][j][0], acc[i][j][1], -bf[cur][j], af[cur][i]);
               }
            } else {
               // short last chunk of the pivot range
#pragma unroll 1
               for (int ks = 0; ks < nk4; ++ks) {
                  double af[RI], bf[CJ];
#pragma unroll
                  for (int i = 0; i < RI; ++i) af[i] = as[ks * 4 * ULDA + i * 8];
#pragma unroll
                  for (int j = 0; j < CJ; ++j) bf[j] = -bs[ks * 4 * ULDB + j * 8];
#pragma unroll
                  for (int i = 0; i < RI; ++i)
#pragma unroll
                     for (int j = 0; j < CJ; ++j) dmma(acc[i][j][0], acc[i][j][1], bf[j], af[i]);
               }
            }
         }
         __syncwarp();
         if (lane == 0) mbar_arrive(empty + s);
      }
      if (!idle) {
         if (interior) {
#pragma unroll
            for (int j = 0; j < CJ; ++j) {
               const int c = c0 + j * 8;
               double* colp = (c < fn) ? (Lp + (size_t)c * ldl + r0) : (Cp + (size_t)(c - fn) * ldc + (r0 - fn));
#pragma unroll
               for (int i = 0; i < RI; ++i) st_c2(colp + i * 8, acc[i][j][0], acc[i][j][1], pol_stream);
            }
         } else {
#pragma unroll
            for (int j = 0; j < CJ; ++j) {
               const int c = c0 + j * 8;
               if (c >= cb || c < ca) continue;
               double* colp = (c < fn) ? (Lp + (size_t)c * ldl) : (Cp + (size_t)(c - fn) * ldc - fn);
#pragma unroll
               for (int i = 0; i < RI; ++i)
#pragma unroll
                  for (int h = 0; h < 2; ++h) {
                     const int r = r0 + i * 8 + h;
                     if (r < m && r >= c) colp[r] = acc[i][j][h];
                  }
            }
         }
      }
      int ntj = tj, nti = ti + nparts;
      while (ntj < ntr && nti >= ntr) { nti = nti - ntr + ntj + 1; ++ntj; }
      tj = ntj; ti = nti;
   }
}
