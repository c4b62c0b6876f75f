// Small fronts in one launch (included by engine.cu after factor_kernels.cuh).
//
// The bottom of an assembly tree is thousands of fronts with a few dozen pivot columns and a few hundred rows.  Taken
// through the general path (k_setup, k_scatter_a, k_extend_add per child ordinal, k_block, k_apply, k_commit, k_update)
// a front's contribution block is zero-filled, read and written by the extend-add and read and written again by the
// update, i.e. the level is bound by HBM traffic it does not need.  k_front does the whole front in one CTA:
//   1. the candidate columns are assembled in shared memory (A entries, then the children's entries that land in them,
//      child by child in the order of the general path: assemble_pre, kernels/assemble.hxx:142-345);
//   2. panels of 32 columns: pivoting on the diagonal block by four warps (ldlt32_4w, block_ldlt.hxx:290-414), the rows
//      below as X = A P inv(L_jj)^T on the FP64 tensor pipe in place, L = X D^{-1} with the a-posteriori test
//      (ldlt_app.cxx:304-323), right-looking update of the next panel;
//   3. L, D^{-1} and the pivot lists go to global memory, and every entry of the contribution block is written exactly
//      once: the children's entries (gathered through inverse row maps, children in order: assemble_post,
//      kernels/assemble.hxx:350-437) minus L (L D)^T from the shared-memory panel.
// A front in which any pivot fails the a-posteriori test is rolled back to its assembled state (columns re-assembled,
// contribution block = the children's entries) and left to the general path's retry rounds, which treat it exactly as
// they would have from the start.  Results are bit-reproducible run to run (fixed summation orders, no atomics).
#pragma once

namespace ffront {
constexpr int MAXN = 64;      // candidate columns (delays included) of a front taken by k_front: two panels
constexpr int MAXCH = 16;     // children
constexpr size_t SMEM_CAP = 200 * 1024;

// leading dimension of the shared panel: >= m + 8 (row tiles of 8 may start at any row), = 4 mod 16 (conflict-free fragments)
__host__ __device__ inline int ldp_of(int m) { return ((m + 4 + 15) / 16) * 16 + 4; }
__host__ __device__ inline int tpad_of(int m, int n) { return ((m - n + 7) & ~7) + 8; }

struct Layout {
   size_t oP, oLc, oLj, oInv, oFwd, oRow, oSd, oSdd, oCa, oCb, oCptr, oKp, oPord, oPtj, oPorig, oPtype, oPerm0, oKey, oMisc, oChild, oInvmap, total;
};
__host__ __device__ inline Layout layout(int m, int n, int nch) {
   Layout L;
   const int npad = (n + NB - 1) / NB * NB;
   size_t o = 0;
   auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 15) & ~(size_t)15; return at; };
   L.oP = take(sizeof(double) * (size_t)npad * ldp_of(m));
   L.oLc = take(sizeof(double) * NB * LDBB);      // multipliers of the pivot loop (NB x LDS), then inv(L_jj) (NB x LDBB)
   L.oInv = L.oLc;
   L.oLj = take(sizeof(double) * NB * NB);
   L.oFwd = take(sizeof(short) * (size_t)(m + 8));  // child row -> parent row of the child being assembled
   L.oRow = take(sizeof(double) * (4 * NB + 2));
   L.oSd = take(sizeof(double) * (2 * MAXN + 2));
   L.oSdd = take(sizeof(double) * (2 * MAXN + 2));
   L.oCa = take(sizeof(double) * MAXN);
   L.oCb = take(sizeof(double) * MAXN);
   L.oCptr = take(sizeof(double*) * 2 * MAXCH);
   L.oKp = take(sizeof(int) * MAXN);
   L.oPord = take(sizeof(int) * NB);
   L.oPtj = take(sizeof(int) * NB);
   L.oPorig = take(sizeof(int) * MAXN);
   L.oPtype = take(sizeof(int) * MAXN);
   L.oPerm0 = take(sizeof(int) * MAXN);
   L.oKey = take(sizeof(int) * 16);
   L.oMisc = take(sizeof(int) * 16);
   L.oChild = take(sizeof(int) * 8 * MAXCH);
   L.oInvmap = take(sizeof(short) * (size_t)(nch > 0 ? nch : 1) * tpad_of(m, n));
   L.total = o;
   return L;
}
__host__ __device__ inline size_t smem_bytes(int m, int n, int nch) { return layout(m, n, nch).total; }
}   // namespace ffront

// child-local trailing index x (0 <= x < t) -> parent-local row: delayed columns first, then the row list through rel
__device__ __forceinline__ int front_map(int x, int nd, int doff, const int* __restrict__ relc, int pncol, int pdel) {
   if (x < nd) return pncol + doff + x;
   const int d = relc[x - nd];
   return d >= pncol ? d + pdel : d;
}

// Candidate columns of the front into the shared panel sP (column-major, ld = ldp): zero, A, children in order.
// Per child, in shared memory: ints sci = {ldl, ldc, n, nelim, t, nd, npm, doff}, pointers scp = {L, C}, srel = offset of its
// relative indices.  A child's rows are first mapped to parent rows in shared memory (sfwd); then a warp takes 32 rows x 8
// columns at a time with all eight loads of a thread in flight at once (the level is bound by memory latency otherwise).
template <int NW>
__device__ __noinline__ void front_assemble(double* sP, int ldp, int npad, int pncol, int pdel, int64_t e0, int64_t e1,
                                            const int* __restrict__ asrc, const int* __restrict__ acol, const int* __restrict__ arow,
                                            const double* __restrict__ val, int nch, const int* sci, const double* const* scp,
                                            const int* __restrict__ rel, const int64_t* srel, short* sfwd) {
   constexpr int T = NW * 32;
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   for (int i = tid; i < npad * ldp; i += T) sP[i] = 0.0;
   __syncthreads();
   for (int64_t e = e0 + tid; e < e1; e += T) {
      int r = arow[e];
      const int c = acol[e];
      if (r >= pncol) r += pdel;
      sP[(size_t)c * ldp + r] = val[asrc[e]];
   }
   __syncthreads();
   for (int k = 0; k < nch; ++k) {
      const int* q = sci + 8 * k;
      const int cldl = q[0], cldc = q[1], cn = q[2], cne = q[3], t = q[4], nd = q[5], npm = q[6], doff = q[7];
      const double* __restrict__ cL = scp[2 * k];
      const double* __restrict__ cC = scp[2 * k + 1];
      const int* relc = rel + srel[k];
      const int ncp = nd + npm;          // child columns that land in candidate columns of the parent
      if (ncp == 0) continue;
      for (int x = tid; x < t; x += T) sfwd[x] = (short)front_map(x, nd, doff, relc, pncol, pdel);
      __syncthreads();
      const int nrc = (t + 31) >> 5, ncb = (ncp + 7) >> 3;
      for (int item = warp; item < nrc * ncb; item += NW) {
         const int rc = item % nrc, cc0 = 8 * (item / nrc);
         if (32 * rc + 31 < cc0) continue;           // the whole chunk lies above the diagonal
         const int r = 32 * rc + lane;
         double v[8];
#pragma unroll
         for (int j = 0; j < 8; ++j) {
            const int cc = cc0 + j, gc = cne + cc;     // gc: column of the child front
            v[j] = 0.0;
            if (cc < ncp && r < t && r >= cc)
               v[j] = gc < cn ? cL[(size_t)gc * cldl + cne + r] : cC[(size_t)(gc - cn) * cldc + (cne - cn) + r];
         }
         const int dr = r < t ? sfwd[r] : 0;
#pragma unroll
         for (int j = 0; j < 8; ++j) {
            const int cc = cc0 + j;
            if (cc < ncp && r < t && r >= cc) {
               const int dc = sfwd[cc];
               const int rr = dr > dc ? dr : dc, cl = dr > dc ? dc : dr;
               sP[(size_t)cl * ldp + rr] += v[j];
            }
         }
      }
      __syncthreads();
   }
}

// NP = number of 32-column panels the launch's fronts may have (1: n <= 32, 2: n <= 64); NW = warps per CTA (4 for the
// smallest fronts, where the number of fronts in flight per SM counts, 8 otherwise)
template <int NP, int NW>
__global__ void __launch_bounds__(NW * 32, NW == 4 ? 4 : 2)
k_front(FrontDev* fr, const SlotStatic* __restrict__ ss, const int* __restrict__ list, const int* __restrict__ child_idx,
        const int* __restrict__ child_doff, const int* __restrict__ rel, const int64_t* __restrict__ aptr,
        const int* __restrict__ asrc, const int* __restrict__ acol, const int* __restrict__ arow,
        const double* __restrict__ val, Params prm) {
   using namespace ffront;
   constexpr int T = NW * 32;
   const int slot = list[blockIdx.x];
   FrontDev& f = fr[slot];
   __shared__ FrontDev sf;
   __shared__ SlotStatic sst;
   __shared__ int64_t srel[MAXCH];
   static_assert(sizeof(SlotStatic) % 8 == 0, "SlotStatic is copied in 8-byte words");
   static_assert(sizeof(FrontDev) / 8 + sizeof(SlotStatic) / 8 <= 64, "front records are copied by the first two warps");
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const int g = lane >> 2, t4 = lane & 3;
   if (tid < (int)(sizeof(FrontDev) / 8)) reinterpret_cast<double*>(&sf)[tid] = reinterpret_cast<const double*>(&f)[tid];
   if (tid >= 64 && tid < 64 + (int)(sizeof(SlotStatic) / 8))
      reinterpret_cast<double*>(&sst)[tid - 64] = reinterpret_cast<const double*>(&ss[slot])[tid - 64];
   __syncthreads();
   const int m = sf.m, n = sf.n, pncol = sf.ncol, pdel = sf.ndelay_in;
   const int nch = sst.child_end - sst.child_begin;
   const bool dbg = prm.debug && blockIdx.x == gridDim.x / 2 && tid == 0;
   long long tq[20];
   int ntq = 0;
#define FT() do { if (dbg && ntq < 20) tq[ntq++] = clock64(); } while (0)
   FT();
   const int npad = (n + NB - 1) / NB * NB, npan = npad / NB;
   const int ldp = ldp_of(m), t = m - n, tpad = tpad_of(m, n);
   extern __shared__ __align__(16) unsigned char fsm[];
   const Layout lo = layout(m, n, nch);
   double* sP = reinterpret_cast<double*>(fsm + lo.oP);
   double* sLc = reinterpret_cast<double*>(fsm + lo.oLc);
   double* sLj = reinterpret_cast<double*>(fsm + lo.oLj);
   double* sInv = reinterpret_cast<double*>(fsm + lo.oInv);     // aliases sLc: the multipliers are dead when the inverse is formed
   short* sfwd = reinterpret_cast<short*>(fsm + lo.oFwd);
   double* srow = reinterpret_cast<double*>(fsm + lo.oRow);
   double* sd = reinterpret_cast<double*>(fsm + lo.oSd);
   double* sdd = reinterpret_cast<double*>(fsm + lo.oSdd);
   double* sca = reinterpret_cast<double*>(fsm + lo.oCa);
   double* scb = reinterpret_cast<double*>(fsm + lo.oCb);
   const double** scp = reinterpret_cast<const double**>(fsm + lo.oCptr);
   int* skp = reinterpret_cast<int*>(fsm + lo.oKp);
   int* spord = reinterpret_cast<int*>(fsm + lo.oPord);
   int* sptj = reinterpret_cast<int*>(fsm + lo.oPtj);
   int* sporig = reinterpret_cast<int*>(fsm + lo.oPorig);
   int* sptype = reinterpret_cast<int*>(fsm + lo.oPtype);
   int* sperm0 = reinterpret_cast<int*>(fsm + lo.oPerm0);
   unsigned* skey = reinterpret_cast<unsigned*>(fsm + lo.oKey);
   int* smisc = reinterpret_cast<int*>(fsm + lo.oMisc);     // 0: error flag, 2: moved, 3: any 2x2 pivot
   int* sci = reinterpret_cast<int*>(fsm + lo.oChild);
   short* sinv = reinterpret_cast<short*>(fsm + lo.oInvmap);

   // ---- children: shapes, pointers, the number of their rows that map into candidate columns (static: SlotStatic::npm)
   if (tid < nch) {
      const int cs = child_idx[sst.child_begin + tid];
      const FrontDev& fc = fr[cs];
      const int cne = fc.nelim, cn = fc.n, cm = fc.m;
      int* q = sci + 8 * tid;
      q[0] = fc.ldl; q[1] = fc.ldc; q[2] = cn; q[3] = cne; q[4] = cm - cne; q[5] = cn - cne; q[6] = ss[cs].npm;
      q[7] = child_doff[sst.child_begin + tid];
      scp[2 * tid] = fc.L; scp[2 * tid + 1] = fc.C;
      srel[tid] = ss[cs].rel_off;
   }
   if (tid >= 32 && tid < 48) smisc[tid - 32] = 0;
   if (tid >= 64 && tid < 64 + MAXN) {
      const int j = tid - 64;
      sperm0[j] = j < pncol ? sst.first_piv + j : 0;
      sporig[j] = j; sptype[j] = 0; sca[j] = 0.0; scb[j] = 0.0; skp[j] = j;
   }
   __syncthreads();
   // delayed columns of the children keep their global pivot indices (k_setup)
   for (int k = 0; k < nch; ++k) {
      const int* q = sci + 8 * k;
      const int nd = q[5];
      if (nd > 0) {
         const int* cperm = fr[child_idx[sst.child_begin + k]].perm;
         for (int d = tid; d < nd; d += T) sperm0[pncol + q[7] + d] = cperm[q[3] + d];
      }
   }
   const int64_t e0 = aptr[slot], e1 = aptr[slot + 1];
   FT();      // 1: records, children
   front_assemble<NW>(sP, ldp, npad, pncol, pdel, e0, e1, asrc, acol, arow, val, nch, sci, scp, rel, srel, sfwd);
   FT();      // 2: assembled

   bool failed = false;
#pragma unroll 1
   for (int pj = 0; pj < NP; ++pj) {
      if (pj >= npan) break;
      const int c0 = pj * NB, wj = min(NB, n - c0), b = c0 + wj;
      // a. pivoting on the diagonal block (warps 0-3)
      if (warp < 4) {
         double a[8];
#pragma unroll
         for (int j = 0; j < 8; ++j) {
            const int c = 8 * warp + j;
            double v = 0.0;
            if (c < wj && lane < wj) v = (c <= lane) ? sP[(size_t)(c0 + c) * ldp + c0 + lane] : sP[(size_t)(c0 + lane) * ldp + c0 + c];
            a[j] = v;
         }
         const double dg = lane < wj ? sP[(size_t)(c0 + lane) * ldp + c0 + lane] : 1.0;
         const int err = ldlt32_4w(a, dg, wj, prm, lane, warp, srow, skey, sLc, sd + 2 * c0, sdd + 2 * c0, spord, sptj);
         bar128();
         if (warp == 0) {
            const bool mv = lane < wj && spord[lane] != lane;
            const unsigned pm = __ballot_sync(0xffffffffu, mv);
            const unsigned p2 = __ballot_sync(0xffffffffu, lane < wj && sptj[lane] != 0);
            if (lane == 0) { smisc[0] = err; smisc[2] = pm != 0u; if (p2) smisc[3] = 1; }
         }
      }
      __syncthreads();
      FT();   // pivoted
      if (smisc[0] != 0) {
         if (tid == 0) { f.flag = smisc[0]; f.w = 0; }
         return;
      }
      const bool moved = smisc[2] != 0;
      // b. L_jj in pivot order; rows of the block in the earlier panel follow the pivot order; coefficients of L = X D^{-1}
      {
         double tmp[NB / NW];
         if (NP > 1 && pj > 0 && moved) {
#pragma unroll
            for (int s = 0; s < NB / NW; ++s) {
               const int c = warp + NW * s;
               tmp[s] = lane < wj ? sP[(size_t)c * ldp + c0 + spord[lane]] : 0.0;
            }
         }
         double lv[NB / NW];
#pragma unroll
         for (int s = 0; s < NB / NW; ++s) {
            const int k = warp + NW * s;
            lv[s] = 0.0;
            if (lane < wj && k < wj) lv[s] = lane > k ? sLc[spord[lane] + k * LDS] : (lane == k ? 1.0 : 0.0);
         }
         __syncthreads();
         if (NP > 1 && pj > 0 && moved) {
#pragma unroll
            for (int s = 0; s < NB / NW; ++s) {
               const int c = warp + NW * s;
               if (lane < wj) sP[(size_t)c * ldp + c0 + lane] = tmp[s];
            }
         }
#pragma unroll
         for (int s = 0; s < NB / NW; ++s) {
            const int k = warp + NW * s;
            sLj[lane + k * NB] = lv[s];
         }
         if (tid < NB) {
            const int k = tid;
            const int pt = k < wj ? sptj[k] : 0;
            sporig[c0 + k] = c0 + (k < wj ? spord[k] : k);
            sptype[c0 + k] = pt;
            const double* dk = sd + 2 * (c0 + k);
            double ca = 0.0, cb = 0.0;
            int kp = c0 + k;
            if (k < wj) {
               if (pt == 0) ca = dk[0];
               else if (pt == 1) { ca = dk[0]; cb = dk[1]; kp = c0 + k + 1; }
               else { ca = dk[1]; cb = dk[-1]; kp = c0 + k - 1; }
            }
            sca[c0 + k] = ca; scb[c0 + k] = cb; skp[c0 + k] = kp;
         }
      }
      __syncthreads();
      if (tid < NB && tid >= wj) spord[tid] = tid;      // padding columns of a short panel are their own pivots
      // b2. inverse of L_jj (one warp)
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
         double2* s2 = reinterpret_cast<double2*>(sInv + lane * LDBB);
#pragma unroll
         for (int i2 = 0; i2 < NB / 2; ++i2) s2[i2] = make_double2(z[2 * i2], z[2 * i2 + 1]);
      }
      __syncthreads();
      FT();   // L_jj, inverse
      // c. rows below the diagonal block: X = A P inv(L_jj)^T in place, a warp per tile of 8 rows; without 2x2 pivots the
      //    a-posteriori test on L = X D^{-1} (c2) is taken on the accumulators
      int bad = 0;
      const bool regtest = !prm.posdef && smisc[3] == 0;
      {
         const int nrt = (m - b + 7) >> 3;
         double cat[4][2];
#pragma unroll
         for (int ct = 0; ct < 4; ++ct) { cat[ct][0] = sca[c0 + ct * 8 + 2 * t4]; cat[ct][1] = sca[c0 + ct * 8 + 2 * t4 + 1]; }
         int so[NB / 4];
#pragma unroll
         for (int q = 0; q < NB / 4; ++q) so[q] = (c0 + spord[4 * q + t4]) * ldp + g;
         for (int rt = warp; rt < nrt; rt += NW) {
            const int rb = b + 8 * rt;
            double af[NB / 4];
#pragma unroll
            for (int q = 0; q < NB / 4; ++q) af[q] = sP[so[q] + rb];
            double acc[4][2];
#pragma unroll
            for (int ct = 0; ct < 4; ++ct) acc[ct][0] = acc[ct][1] = 0.0;
#pragma unroll
            for (int q = 0; q < NB / 4; ++q) {
#pragma unroll
               for (int ct = 0; ct < 4; ++ct)
                  if (4 * q < 8 * (ct + 1)) dmma(acc[ct][0], acc[ct][1], af[q], sInv[(4 * q + t4) * LDBB + ct * 8 + g]);
            }
            __syncwarp();
#pragma unroll
            for (int ct = 0; ct < 4; ++ct) {
               sP[(size_t)(c0 + ct * 8 + 2 * t4) * ldp + rb + g] = acc[ct][0];
               sP[(size_t)(c0 + ct * 8 + 2 * t4 + 1) * ldp + rb + g] = acc[ct][1];
            }
            if (regtest) {
#pragma unroll
               for (int ct = 0; ct < 4; ++ct)
#pragma unroll
                  for (int h = 0; h < 2; ++h) {
                     const double ca = cat[ct][h], xv = acc[ct][h];
                     if (!(fabs(ca * xv) <= prm.lim) || (ca == 0.0 && !(fabs(xv) < prm.small_))) bad = 1;
                  }
            }
         }
      }
      __syncthreads();
      // the diagonal block itself: L_jj (unit lower, pivot order)
      for (int k = warp; k < wj; k += NW)
         if (lane < wj) sP[(size_t)(c0 + k) * ldp + c0 + lane] = sLj[lane + k * NB];
      // c2. a-posteriori test on L = X D^{-1} with 2x2 pivots in the front
      if (!prm.posdef && !regtest) {
         for (int r = b + tid; r < m; r += T) {
            for (int k = 0; k < wj; ++k) {
               const int kk = c0 + k;
               const double xv = sP[(size_t)kk * ldp + r];
               const double ca = sca[kk];
               double l = ca * xv;
               if (sptype[kk] != 0) l += scb[kk] * sP[(size_t)skp[kk] * ldp + r];
               else if (ca == 0.0 && !(fabs(xv) < prm.small_)) l = DINF;
               if (!(fabs(l) <= prm.lim)) bad = 1;
            }
         }
      }
      if (__syncthreads_or(bad)) { failed = true; break; }
      FT();   // X, test
      // e. right-looking update of the next panel: A(r, c) -= sum_k L(r, k) X(c, k), c in [b, n), r >= b
      if (NP > 1 && pj + 1 < npan) {
         const int nrb = (m - b + 7) >> 3, njt = (n - b + 7) >> 3;
         const bool any2 = smisc[3] != 0;
         for (int ib = warp; ib < nrb; ib += NW) {
            const int rb = b + 8 * ib;
            double lf[NB / 4];
#pragma unroll
            for (int q = 0; q < NB / 4; ++q) {
               const int k = c0 + 4 * q + t4;
               double l = sca[k] * sP[(size_t)k * ldp + rb + g];
               if (any2) l += scb[k] * sP[(size_t)skp[k] * ldp + rb + g];
               lf[q] = l;
            }
            for (int jt = 0; jt < njt; ++jt) {
               const int cb0 = b + 8 * jt;
               double* cp = sP + (size_t)(cb0 + g) * ldp + rb + 2 * t4;
               double a0 = cp[0], a1 = cp[1];
#pragma unroll
               for (int q = 0; q < NB / 4; ++q) dmma(a0, a1, -sP[(size_t)(c0 + 4 * q + t4) * ldp + cb0 + g], lf[q]);
               __syncwarp();
               if (cb0 + g < n) { cp[0] = a0; cp[1] = a1; }       // padding columns stay zero
            }
         }
         __syncthreads();
      }
   }

   if (failed) {
      // ---- roll back: the assembled columns to global memory, the contribution block without the update
      __syncthreads();
      front_assemble<NW>(sP, ldp, npad, pncol, pdel, e0, e1, asrc, acol, arow, val, nch, sci, scp, rel, srel, sfwd);
      for (int c = warp; c < n; c += NW) {
         double* col = sf.L + (size_t)c * sf.ldl;
         for (int r = lane; r < m; r += 32) col[r] = r >= c ? sP[(size_t)c * ldp + r] : 0.0;
      }
      for (int j = tid; j < n; j += T) { sf.perm[j] = sperm0[j]; sf.orig[j] = j; }
   } else {
      // ---- L, D^{-1}, pivot lists
      const bool any2 = smisc[3] != 0;
      for (int k = warp; k < n; k += NW) {
         const int bk = min(n, (k / NB + 1) * NB);      // end of the diagonal block of k's panel
         double* col = sf.L + (size_t)k * sf.ldl;
         const double ca = sca[k], cb = scb[k];
         const double* xk = sP + (size_t)k * ldp;
         const double* xp = sP + (size_t)skp[k] * ldp;
         for (int r = lane; r < m; r += 32) {
            double v;
            if (r < k) v = 0.0;
            else if (r < bk) v = xk[r];
            else { v = ca * xk[r]; if (any2) v += cb * xp[r]; }
            col[r] = v;
         }
      }
      for (int i = tid; i < 2 * n; i += T) sf.D[i] = sd[i];
      for (int j = tid; j < n; j += T) { sf.perm[j] = sperm0[sporig[j]]; sf.orig[j] = sporig[j]; }
   }
   if (tid == 0) {
      const int ne = failed ? 0 : n;
      f.done = ne; f.gend = n; f.w = 0; f.fail = 0; f.nelim = ne; f.stall = 0;
      f.urec[0][1] = 0; f.urec[1][1] = 0;
      f.p2mask[0] = f.p2mask[1] = f.p2mask[2] = f.p2mask[3] = 0u;
   }
   FT();      // L, D written
   if (t <= 0) return;

   // ---- contribution block: every entry written once = children's entries (in child order) - L (L D)^T
   for (int i = tid; i < nch * tpad; i += T) sinv[i] = (short)-1;
   __syncthreads();
   for (int k = 0; k < nch; ++k) {
      const int* q = sci + 8 * k;
      const int tt = q[4], nd = q[5], npm = q[6];
      const int* relc = rel + srel[k];
      for (int x = nd + npm + tid; x < tt; x += T) sinv[k * tpad + relc[x - nd] - pncol] = (short)(x - nd);
   }
   __syncthreads();
   FT();      // inverse maps
   {
      // A warp takes 8 rows x four tiles of 8 columns at a time (item (ib, jg): rows 8 ib .., column tiles 4 jg .. 4 jg + 3),
      // items of a row block dealt to the warps round robin with a start that rotates with the row block.  The children's
      // entries of the NEXT item are requested (first child straight into the accumulators, second into vb) before the
      // MMAs of the current one.  Interior items (all four tiles below the diagonal, all rows inside the front) take a
      // path without bounds predicates, with pointer increments instead of index arithmetic: the phase is bound by
      // instruction issue, not by the tensor pipe.
      const int ntile = (t + 7) >> 3;
      const bool any2 = smisc[3] != 0;
      const int kq = failed ? 0 : npad / 4;
      double* const Cp = sf.C;
      const int ldc = sf.ldc;
      const bool two = nch >= 2;
      const bool c16 = (reinterpret_cast<size_t>(Cp) & 15) == 0;
      const short* const iva = sinv;
      const short* const ivb = two ? sinv + tpad : sinv;
      const double* __restrict__ const Ca = nch >= 1 ? scp[1] : sP;     // never dereferenced without children
      const double* __restrict__ const Cb2 = two ? scp[3] : sP;
      const int lda = nch >= 1 ? sci[1] : 0, ldb = two ? sci[9] : 0;
      const int ldp4 = 4 * ldp;
      const double* const xb = sP + (size_t)t4 * ldp + n + g;           // operand base: + q * ldp4 + 8 * (tile index)
      const bool lean = !failed && !any2;
      double caq[NP * NB / 4];
#pragma unroll
      for (int q = 0; q < NP * NB / 4; ++q) caq[q] = q < kq ? sca[4 * q + t4] : 0.0;
      auto interior = [&](int ib, int jg) { return 4 * jg + 3 < ib && 8 * ib + 7 < t; };
      // items of this warp: in row block ib the items jg = (warp - 3 ib) mod NW, + NW, ...
      int ib = 0, jg = 0;
      auto settle = [&]() {       // (ib, jg) -> the first item of this warp at or after it
         while (ib < ntile && jg > (ib >> 2)) { ++ib; jg = (warp - 3 * ib) & (NW - 1); }
      };
      if (lean && nch <= 2) {
         // ---- pass 1: interior items, software-pipelined
         auto settle_int = [&]() {
            for (;;) {
               settle();
               if (ib >= ntile || interior(ib, jg)) break;
               jg += NW;
            }
         };
         auto gather_fast = [&](int gi, int gj, double (&va)[4][2], double (&vb)[4][2]) {
            const int i0 = 8 * gi + 2 * t4;
            if (nch == 0) {
#pragma unroll
               for (int u = 0; u < 4; ++u) va[u][0] = va[u][1] = 0.0;
               return;
            }
            const int ia0 = iva[i0], ia1 = iva[i0 + 1];
            const short* ja_p = iva + 32 * gj + g;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
               const int ja = ja_p[8 * u];
               const double* __restrict__ pc = Ca + (size_t)ja * lda;
               va[u][0] = (ja >= 0 && ia0 >= ja) ? pc[ia0] : 0.0;
               va[u][1] = (ja >= 0 && ia1 >= ja) ? pc[ia1] : 0.0;
            }
            if (two) {
               const int ib0 = ivb[i0], ib1 = ivb[i0 + 1];
               const short* jb_p = ivb + 32 * gj + g;
#pragma unroll
               for (int u = 0; u < 4; ++u) {
                  const int jb = jb_p[8 * u];
                  const double* __restrict__ pc = Cb2 + (size_t)jb * ldb;
                  vb[u][0] = (jb >= 0 && ib0 >= jb) ? pc[ib0] : 0.0;
                  vb[u][1] = (jb >= 0 && ib1 >= jb) ? pc[ib1] : 0.0;
               }
            }
         };
         auto process_fast = [&](int cib, int cjg, double (&acc)[4][2], double (&vb)[4][2]) {
            const int i0 = 8 * cib + 2 * t4;
            if (two) {
#pragma unroll
               for (int u = 0; u < 4; ++u) { acc[u][0] += vb[u][0]; acc[u][1] += vb[u][1]; }
            }
            const double* lp = xb + 8 * cib;
            const double* xp = xb + 32 * cjg;
            double lf[NP * NB / 4];
#pragma unroll
            for (int q = 0; q < NP * NB / 4; ++q) lf[q] = (NP == 1 || q < kq) ? caq[q] * lp[(size_t)q * ldp4] : 0.0;
#pragma unroll
            for (int q = 0; q < NP * NB / 4; ++q) {
               if (NP == 1 || q < kq) {
#pragma unroll
#pragma unroll
                  for (int u = 0; u < 4; ++u) dmma(acc[u][0], acc[u][1], -xp[8 * u], lf[q]);
                  xp += ldp4;
               }
            }
            double* cp = Cp + (size_t)(32 * cjg + g) * ldc + i0;
            if (c16) {
#pragma unroll
               for (int u = 0; u < 4; ++u) { *reinterpret_cast<double2*>(cp) = make_double2(acc[u][0], acc[u][1]); cp += 8 * (size_t)ldc; }
            } else {
#pragma unroll
               for (int u = 0; u < 4; ++u) { cp[0] = acc[u][0]; cp[1] = acc[u][1]; cp += 8 * (size_t)ldc; }
            }
         };
         ib = 0; jg = warp;
         settle_int();
         double accA[4][2], vbA[4][2], accB[4][2], vbB[4][2];
         if (ib < ntile) gather_fast(ib, jg, accA, vbA);
         while (ib < ntile) {
            int cib = ib, cjg = jg;
            jg += NW; settle_int();
            if (ib < ntile) gather_fast(ib, jg, accB, vbB);
            process_fast(cib, cjg, accA, vbA);
            if (ib >= ntile) break;
            cib = ib; cjg = jg;
            jg += NW; settle_int();
            if (ib < ntile) gather_fast(ib, jg, accA, vbA);
            process_fast(cib, cjg, accB, vbB);
         }
      }
      // ---- pass 2: the items pass 1 left (tiles on the diagonal, the last rows; everything for fronts with 2x2 pivots,
      //      more than two children, or rolled back)
      const bool did_interior = lean && nch <= 2;
      FT();   // interior items
      ib = 0; jg = warp;
      settle();
#pragma unroll 1
      for (; ib < ntile; jg += NW, settle()) {
         if (ib >= ntile) break;
         if (did_interior && interior(ib, jg)) continue;
         const int rb = n + 8 * ib;
         const int jt1 = min(4 * jg + 4, ib + 1);
         const int i0 = 8 * ib + 2 * t4;
         double acc[4][2];
#pragma unroll
         for (int u = 0; u < 4; ++u) acc[u][0] = acc[u][1] = 0.0;
         for (int k = 0; k < nch; ++k) {          // children in order
            const short* iv = sinv + k * tpad;
            const int ic0 = iv[i0], ic1 = iv[i0 + 1];
            const double* __restrict__ Cc = scp[2 * k + 1];
            const int ldk = sci[8 * k + 1];
            double vc[4][2];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
               const int jt = 4 * jg + u;
               const int jc = jt < jt1 ? iv[8 * jt + g] : -1;
               vc[u][0] = (jc >= 0 && ic0 >= jc) ? Cc[(size_t)jc * ldk + ic0] : 0.0;
               vc[u][1] = (jc >= 0 && ic1 >= jc) ? Cc[(size_t)jc * ldk + ic1] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) { acc[u][0] += vc[u][0]; acc[u][1] += vc[u][1]; }
         }
         double lf[NP * NB / 4];
#pragma unroll
         for (int q = 0; q < NP * NB / 4; ++q) {
            double l = 0.0;
            if (q < kq) {
               const int k = 4 * q + t4;
               l = sca[k] * sP[(size_t)k * ldp + rb + g];
               if (any2) l += scb[k] * sP[(size_t)skp[k] * ldp + rb + g];
            }
            lf[q] = l;
         }
         __syncwarp();
#pragma unroll
         for (int u = 0; u < 4; ++u) {
            const int jt = 4 * jg + u;
            if (jt < jt1) {
               const int j = 8 * jt + g;
#pragma unroll
               for (int q = 0; q < NP * NB / 4; ++q)
                  if (q < kq) dmma(acc[u][0], acc[u][1], -sP[(size_t)(4 * q + t4) * ldp + n + j], lf[q]);
               if (j < t) {
                  double* cp = Cp + (size_t)j * ldc;
                  if (i0 < t && i0 >= j) cp[i0] = acc[u][0];
                  if (i0 + 1 < t && i0 + 1 >= j) cp[i0 + 1] = acc[u][1];
               }
            }
         }
      }
   }
   FT();
   if (dbg) {
      printf("[k_front] m=%d n=%d nch=%d ticks:", m, n, nch);
      for (int i = 1; i < ntq; ++i) printf(" %lld", tq[i] - tq[i - 1]);
      printf("\n");
   }
#undef FT
}
