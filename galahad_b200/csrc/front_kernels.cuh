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
constexpr int THREADS = 256;
constexpr size_t SMEM_CAP = 200 * 1024;

// leading dimension of the shared panel: >= m + 8 (row tiles of 8 may start at any row), = 4 mod 16 (conflict-free fragments)
__host__ __device__ inline int ldp_of(int m) { return ((m + 4 + 15) / 16) * 16 + 4; }
__host__ __device__ inline int tpad_of(int m, int n) { return ((m - n + 7) & ~7) + 8; }

struct Layout {
   size_t oP, oLc, oLj, oInv, oRow, oSd, oSdd, oCa, oCb, oCptr, oKp, oPord, oPtj, oPorig, oPtype, oPerm0, oKey, oMisc, oChild, oInvmap, total;
};
__host__ __device__ inline Layout layout(int m, int n, int nch) {
   Layout L;
   const int npad = (n + NB - 1) / NB * NB;
   size_t o = 0;
   auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 15) & ~(size_t)15; return at; };
   L.oP = take(sizeof(double) * (size_t)npad * ldp_of(m));
   L.oLc = take(sizeof(double) * NB * LDS);
   L.oLj = take(sizeof(double) * NB * NB);
   L.oInv = take(sizeof(double) * NB * LDBB);
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

// per child, in shared memory: ints {ldl, ldc, n, nelim, t, nd, npm, doff} and pointers {L, C}; rel offset via SlotStatic
struct FrontChildren {
   int* ci;             // 8 ints per child
   const double** cp;   // 2 pointers per child
   const int* const* relp;
};

// child-local trailing index x (0 <= x < t) -> parent-local row: delayed columns first, then the row list through rel
__device__ __forceinline__ int front_map(int x, int nd, int doff, const int* __restrict__ relc, int pncol, int pdel) {
   if (x < nd) return pncol + doff + x;
   const int d = relc[x - nd];
   return d >= pncol ? d + pdel : d;
}

// Candidate columns of the front into the shared panel sP (column-major, ld = ldp): zero, A, children in order.
__device__ __noinline__ void front_assemble(double* sP, int ldp, int npad, int pncol, int pdel, int64_t e0, int64_t e1,
                                            const int* __restrict__ asrc, const int* __restrict__ acol, const int* __restrict__ arow,
                                            const double* __restrict__ val, int nch, const int* sci, const double* const* scp,
                                            const int* __restrict__ rel, const int64_t* srel) {
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   for (int i = tid; i < npad * ldp; i += ffront::THREADS) sP[i] = 0.0;
   __syncthreads();
   for (int64_t e = e0 + tid; e < e1; e += ffront::THREADS) {
      int r = arow[e];
      const int c = acol[e];
      if (r >= pncol) r += pdel;
      sP[(size_t)c * ldp + r] = val[asrc[e]];
   }
   __syncthreads();
   for (int k = 0; k < nch; ++k) {
      const int* q = sci + 8 * k;
      const int cldl = q[0], cldc = q[1], cn = q[2], cne = q[3], t = q[4], nd = q[5], npm = q[6], doff = q[7];
      const double* cL = scp[2 * k];
      const double* cC = scp[2 * k + 1];
      const int* relc = rel + srel[k];
      const int ncp = nd + npm;          // child columns that land in candidate columns of the parent
      for (int cc = warp; cc < ncp; cc += 8) {
         const int dc = front_map(cc, nd, doff, relc, pncol, pdel);
         const int gc = cne + cc;        // column of the child front
         const double* colp = gc < cn ? cL + (size_t)gc * cldl + cne : cC + (size_t)(gc - cn) * cldc + (cne - cn);
         for (int r = cc + lane; r < t; r += 32) {
            const double v = colp[r];
            const int dr = front_map(r, nd, doff, relc, pncol, pdel);
            const int rr = dr > dc ? dr : dc, cl = dr > dc ? dc : dr;
            sP[(size_t)cl * ldp + rr] += v;
         }
      }
      __syncthreads();
   }
}

// NP = number of 32-column panels the launch's fronts may have (1: n <= 32, 2: n <= 64)
template <int NP>
__global__ void __launch_bounds__(ffront::THREADS, 2)
k_front(FrontDev* fr, const SlotStatic* __restrict__ ss, const int* __restrict__ list, const int* __restrict__ child_idx,
        const int* __restrict__ child_doff, const int* __restrict__ rel, const int64_t* __restrict__ aptr,
        const int* __restrict__ asrc, const int* __restrict__ acol, const int* __restrict__ arow,
        const double* __restrict__ val, Params prm) {
   using namespace ffront;
   const int slot = list[blockIdx.x];
   FrontDev& f = fr[slot];
   __shared__ FrontDev sf;
   __shared__ SlotStatic sst;
   __shared__ int64_t srel[MAXCH];
   static_assert(sizeof(SlotStatic) % 8 == 0, "SlotStatic is copied in 8-byte words");
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const int g = lane >> 2, t4 = lane & 3;
   if (tid < (int)(sizeof(FrontDev) / 8)) reinterpret_cast<double*>(&sf)[tid] = reinterpret_cast<const double*>(&f)[tid];
   if (tid >= 128 && tid < 128 + (int)(sizeof(SlotStatic) / 8))
      reinterpret_cast<double*>(&sst)[tid - 128] = reinterpret_cast<const double*>(&ss[slot])[tid - 128];
   __syncthreads();
   const int m = sf.m, n = sf.n, pncol = sf.ncol, pdel = sf.ndelay_in;
   const int nch = sst.child_end - sst.child_begin;
   const int npad = (n + NB - 1) / NB * NB, npan = npad / NB;
   const int ldp = ldp_of(m), t = m - n, tpad = tpad_of(m, n);
   extern __shared__ __align__(16) unsigned char fsm[];
   const Layout lo = layout(m, n, nch);
   double* sP = reinterpret_cast<double*>(fsm + lo.oP);
   double* sLc = reinterpret_cast<double*>(fsm + lo.oLc);
   double* sLj = reinterpret_cast<double*>(fsm + lo.oLj);
   double* sInv = reinterpret_cast<double*>(fsm + lo.oInv);
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

   // ---- children: shapes, pointers, the number of their rows that map into candidate columns
   if (tid < nch) {
      const int cs = child_idx[sst.child_begin + tid];
      const FrontDev& fc = fr[cs];
      const int cne = fc.nelim, cn = fc.n, cm = fc.m;
      const int tt = cm - cne, nd = cn - cne;
      const int64_t ro = ss[cs].rel_off;
      const int* relc = rel + ro;
      int lo2 = 0, hi = tt - nd;                 // rel is ascending: first entry >= pncol
      while (lo2 < hi) { const int mid = (lo2 + hi) >> 1; if (relc[mid] < pncol) lo2 = mid + 1; else hi = mid; }
      int* q = sci + 8 * tid;
      q[0] = fc.ldl; q[1] = fc.ldc; q[2] = cn; q[3] = cne; q[4] = tt; q[5] = nd; q[6] = lo2; q[7] = child_doff[sst.child_begin + tid];
      scp[2 * tid] = fc.L; scp[2 * tid + 1] = fc.C;
      srel[tid] = ro;
   }
   if (tid < 16) smisc[tid] = 0;
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
         for (int d = tid; d < nd; d += THREADS) sperm0[pncol + q[7] + d] = cperm[q[3] + d];
      }
   }
   const int64_t e0 = aptr[slot], e1 = aptr[slot + 1];
   front_assemble(sP, ldp, npad, pncol, pdel, e0, e1, asrc, acol, arow, val, nch, sci, scp, rel, srel);

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
      if (smisc[0] != 0) {
         if (tid == 0) { f.flag = smisc[0]; f.w = 0; }
         return;
      }
      const bool moved = smisc[2] != 0;
      // b. L_jj in pivot order; rows of the block in the earlier panel follow the pivot order; coefficients of L = X D^{-1}
      {
         double tmp[NB / 8];
         if (NP > 1 && pj > 0 && moved) {
#pragma unroll
            for (int s = 0; s < NB / 8; ++s) {
               const int c = warp + 8 * s;
               tmp[s] = lane < wj ? sP[(size_t)c * ldp + c0 + spord[lane]] : 0.0;
            }
         }
         double lv[NB / 8];
#pragma unroll
         for (int s = 0; s < NB / 8; ++s) {
            const int k = warp + 8 * s;
            lv[s] = 0.0;
            if (lane < wj && k < wj) lv[s] = lane > k ? sLc[spord[lane] + k * LDS] : (lane == k ? 1.0 : 0.0);
         }
         __syncthreads();
         if (NP > 1 && pj > 0 && moved) {
#pragma unroll
            for (int s = 0; s < NB / 8; ++s) {
               const int c = warp + 8 * s;
               if (lane < wj) sP[(size_t)c * ldp + c0 + lane] = tmp[s];
            }
         }
#pragma unroll
         for (int s = 0; s < NB / 8; ++s) {
            const int k = warp + 8 * s;
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
      // c. rows below the diagonal block: X = A P inv(L_jj)^T in place, a warp per tile of 8 rows
      {
         const int nrt = (m - b + 7) >> 3;
         for (int rt = warp; rt < nrt; rt += 8) {
            const int rb = b + 8 * rt;
            double af[NB / 4];
#pragma unroll
            for (int q = 0; q < NB / 4; ++q) af[q] = sP[(size_t)(c0 + spord[4 * q + t4]) * ldp + rb + g];
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
         }
      }
      __syncthreads();
      // the diagonal block itself: L_jj (unit lower, pivot order)
      for (int k = warp; k < wj; k += 8)
         if (lane < wj) sP[(size_t)(c0 + k) * ldp + c0 + lane] = sLj[lane + k * NB];
      // c2. a-posteriori test on L = X D^{-1}
      int bad = 0;
      if (!prm.posdef) {
         for (int r = b + tid; r < m; r += THREADS) {
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
      // e. right-looking update of the next panel: A(r, c) -= sum_k L(r, k) X(c, k), c in [b, n), r >= b
      if (NP > 1 && pj + 1 < npan) {
         const int nrb = (m - b + 7) >> 3, njt = (n - b + 7) >> 3;
         const bool any2 = smisc[3] != 0;
         for (int ib = warp; ib < nrb; ib += 8) {
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
               if (cb0 + g < npad) { cp[0] = a0; cp[1] = a1; }
            }
         }
         __syncthreads();
      }
   }

   if (failed) {
      // ---- roll back: the assembled columns to global memory, the contribution block without the update
      __syncthreads();
      front_assemble(sP, ldp, npad, pncol, pdel, e0, e1, asrc, acol, arow, val, nch, sci, scp, rel, srel);
      for (int c = warp; c < n; c += 8) {
         double* col = sf.L + (size_t)c * sf.ldl;
         for (int r = lane; r < m; r += 32) col[r] = r >= c ? sP[(size_t)c * ldp + r] : 0.0;
      }
      for (int j = tid; j < n; j += THREADS) { sf.perm[j] = sperm0[j]; sf.orig[j] = j; }
   } else {
      // ---- L, D^{-1}, pivot lists
      const bool any2 = smisc[3] != 0;
      for (int k = warp; k < n; k += 8) {
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
      for (int i = tid; i < 2 * n; i += THREADS) sf.D[i] = sd[i];
      for (int j = tid; j < n; j += THREADS) { sf.perm[j] = sperm0[sporig[j]]; sf.orig[j] = sporig[j]; }
   }
   if (tid == 0) {
      const int ne = failed ? 0 : n;
      f.done = ne; f.gend = n; f.w = 0; f.fail = 0; f.nelim = ne; f.stall = 0;
      f.urec[0][1] = 0; f.urec[1][1] = 0;
      f.p2mask[0] = f.p2mask[1] = f.p2mask[2] = f.p2mask[3] = 0u;
   }
   if (t <= 0) return;

   // ---- contribution block: every entry written once = children's entries (in child order) - L (L D)^T
   for (int i = tid; i < nch * tpad; i += THREADS) sinv[i] = (short)-1;
   __syncthreads();
   for (int k = 0; k < nch; ++k) {
      const int* q = sci + 8 * k;
      const int tt = q[4], nd = q[5], npm = q[6];
      const int* relc = rel + srel[k];
      for (int x = nd + npm + tid; x < tt; x += THREADS) sinv[k * tpad + relc[x - nd] - pncol] = (short)(x - nd);
   }
   __syncthreads();
   {
      const int ntile = (t + 7) >> 3;
      const bool any2 = smisc[3] != 0;
      const int kq = failed ? 0 : npad / 4;
      double* const Cp = sf.C;
      const int ldc = sf.ldc;
      int item = 0;
      for (int ib = 0; ib < ntile; ++ib) {
         for (int jg = 0; jg <= (ib >> 2); ++jg, ++item) {
            if ((item & 7) != warp) continue;
            const int rb = n + 8 * ib;
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
            const int jt1 = min(4 * jg + 4, ib + 1);
            const int i0 = 8 * ib + 2 * t4;
            for (int jt = 4 * jg; jt < jt1; ++jt) {
               const int j = 8 * jt + g;
               double a0 = 0.0, a1 = 0.0;
               for (int k = 0; k < nch; ++k) {
                  const short* iv = sinv + k * tpad;
                  const int jj = iv[j];
                  if (jj < 0) continue;
                  const int ii0 = iv[i0], ii1 = iv[i0 + 1];
                  const double* cc = scp[2 * k + 1] + (size_t)jj * sci[8 * k + 1];
                  if (ii0 >= jj) a0 += cc[ii0];
                  if (ii1 >= jj) a1 += cc[ii1];
               }
               __syncwarp();
#pragma unroll
               for (int q = 0; q < NP * NB / 4; ++q)
                  if (q < kq) dmma(a0, a1, -sP[(size_t)(4 * q + t4) * ldp + n + j], lf[q]);
               if (j < t) {
                  double* cp = Cp + (size_t)j * ldc;
                  if (i0 < t && i0 >= j) cp[i0] = a0;
                  if (i0 + 1 < t && i0 + 1 >= j) cp[i0 + 1] = a1;
               }
            }
         }
      }
   }
}
