// b200ldlt numeric engine: multifrontal LDL^T on one B200 (sm_100a), level-scheduled.
//
// Reference behaviour restated here (not ported): src/ssids/cpu/NumericSubtree.hxx:53-275
// (assemble -> factor -> contribution, inertia count), kernels/assemble.hxx:142-437,
// kernels/ldlt_app.cxx (a-posteriori threshold pivoting), kernels/block_ldlt.hxx:290-414
// (complete pivoting in the diagonal block), kernels/ldlt_tpp.cxx:166-270 (TPP finish) and the
// solves NumericSubtree.hxx:280-400.  The schedule is our own: the assembly tree is cut into depth
// levels; all fronts of a level are assembled and factorised by batched kernels; every front is
// eliminated 32 columns at a time (diag-block factor -> apply+a-posteriori test -> commit ->
// trailing update); columns that fail are swapped behind the candidates, retried by a TPP kernel
// and otherwise delayed to the parent.  Progress counters live in device memory, the host reads
// them once per level.  Extend-add is atomic-free (one launch per child ordinal) so results are
// bit-reproducible run to run.
//
// Data layout in HBM: per front, L is m x n column-major (ld = m rounded
// up to 2), D^{-1} is 2n doubles in SSIDS's format (ldlt_app.cxx:324-329), perm is n ints, the
// contribution block C is (m-n)^2 column-major and lives one level (ping-pong arenas).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <string>
#include <vector>

#include "engine.hpp"

namespace b200 {

constexpr int NB = 32;          // sub-panel width: complete pivoting inside 32x32 diagonal sub-blocks
constexpr int PB = 128;         // block: columns eliminated between two trailing updates
constexpr int LDS = NB + 1;     // padded shared leading dimension
constexpr int EA_TR = 256;      // extend-add tile rows
constexpr int EA_TC = 8;        // extend-add tile cols
#define DINF (__longlong_as_double(0x7ff0000000000000LL))

struct FrontDev {
   double* L;
   double* C;
   double* D;
   int* perm;
   int* orig;      // orig[r] = local index variable r had when the front was assembled (r < n)
   double* Wld;    // ldl x xw  X = L*D of the pivots (front-row indexed): a ring over the pivot index, pivot p of the front
                   //           lives in column p mod xw, so that a trailing update of any rank reads one contiguous pivot range
   double* Wo;     // ldl x wc  permuted original columns of the block (restore on failure)
   double* dbuf;   // L11[PB*wc] | A0[PB*wc] | dk[2*PB+2] | Linv[NB*wc] (inverses of the 32x32 diagonal blocks of L11)
   int* ibuf;      // lperm[PB] | ptype[PB] | permuted
   int m, n, ldl, ldc;
   int ncol, nrow, ndelay_in, first_piv;
   int done;       // columns eliminated so far (L columns [0,done) are final)
   int gend;       // candidates of the front are [done, gend); [gend, n) failed this pass
   int w, fail;    // current block width, first failed column of the block (atomicMin)
   int nelim, flag, slot;
   int wc;         // block workspace width: min(n, PB) rounded up to a multiple of NB
   unsigned p2mask[4];  // bit k set: column k of the current block is the second column of a 2x2 pivot
   int state;           // 0 absent on this rank, 2 factorised here, 3 imported from another rank (child data only)
   int active;          // 1 while this front is being assembled / factorised by the current call
   int part_kind;       // stage of the front in the proportional mapping (multi-GPU factor / solve phases); 0 on one GPU
   int stall;           // look-ahead schedule: a failed pivot was seen, the front waits for the synchronous retry rounds
   int urec[2][4];      // per block parity: {first pivot column k0, accepted pivots kc, end of the strip columns, -}
                        // written by k_commit, read by the trailing updates of that block
   int xw;              // columns of the X ring
   int gk0;             // deferred (rank G*PB) updates: first pivot of the current group of blocks
   int brec[2][6];      // per group parity: {first pivot, end pivot, first column, end of the next group's columns, tag, -}:
                        // the group's trailing update, written by k_commit of the group's last block
   int pend[4];         // {first pivot, end pivot, first column, flag}: deferred update not yet owed by a bulk launch
   int fused;           // 1: assembled and factorised by k_front (front_kernels.cuh); the general kernels of the level pass it by
};

struct SlotStatic {
   int ncol, nrow, parent, first_piv;   // parent slot, -1 for roots
   int64_t rl_off, rel_off;
   int child_begin, child_end;
   int npm, pad_;                       // npm: rows below the pivots that are candidate columns of the parent (rel < parent's ncol)
};

struct Params {
   double u, small_, lim;
   int action, posdef;
   int debug;      // B200LDLT_TRACE=2: k_block prints its phase timings (clock64) for the first front of a launch
};

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double* felem(const FrontDev& f, int r, int c) {
   // element (r,c), r >= c, of the front: candidate columns live in L, the rest in C
   return (c < f.n) ? (f.L + (size_t)c * f.ldl + r) : (f.C + (size_t)(c - f.n) * f.ldc + (r - f.n));
}

__device__ __forceinline__ int compute_nelim(int fail, int w, const unsigned* p2mask) {
   int nelim = fail < w ? fail : w;
   if (nelim < w && nelim > 0 && ((p2mask[nelim >> 5] >> (nelim & 31)) & 1u)) nelim -= 1;   // never split a 2x2 (Column::adjust)
   return nelim;
}
// front handled by this CTA: bucket-contiguous slots, or an explicit list on the rare retry paths
__device__ __forceinline__ int front_slot(const int* list, int slot0) {
   return list ? list[blockIdx.x] : slot0 + (int)blockIdx.x;
}

// ------------------------------------------------------------------------------------------
// assembly kernels
// ------------------------------------------------------------------------------------------
__global__ void k_setup(FrontDev* fr, const SlotStatic* ss, const int* child_idx,
                        const int* child_doff, int slot0, int nslots) {
   int s = slot0 + blockIdx.x;
   if (s >= slot0 + nslots) return;
   FrontDev& f = fr[s];
   if (!f.active || f.fused) return;
   const SlotStatic st = ss[s];
   for (int j = threadIdx.x; j < st.ncol; j += blockDim.x) f.perm[j] = st.first_piv + j;
   for (int j = threadIdx.x; j < f.n; j += blockDim.x) f.orig[j] = j;
   for (int ci = st.child_begin; ci < st.child_end; ++ci) {
      const FrontDev& fc = fr[child_idx[ci]];
      int nd = fc.n - fc.nelim, off = child_doff[ci];
      for (int d = threadIdx.x; d < nd; d += blockDim.x)
         f.perm[st.ncol + off + d] = fc.perm[fc.nelim + d];
   }
}

__global__ void k_scatter_a(const FrontDev* fr, const int* aslot, const int* asrc, const int* acol,
                            const int* arow, const double* __restrict__ val, int64_t e0, int64_t e1) {
   int64_t e = e0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (e >= e1) return;
   const FrontDev& f = fr[aslot[e]];
   if (!f.active || f.fused) return;
   int r = arow[e], c = acol[e];
   if (r >= f.ncol) r += f.ndelay_in;
   f.L[(size_t)c * f.ldl + r] = val[asrc[e]];
}

// children -> parent, one launch per child ordinal so that no two CTAs touch the same parent entry
__global__ void k_extend_add(const FrontDev* fr, const SlotStatic* ss, const int* rel,
                             const int* list, const int* doffs, const int64_t* tileptr, int nlist) {
   __shared__ int sdc[EA_TC];
   int64_t b = blockIdx.x;
   int lo = 0, hi = nlist;   // find i: tileptr[i] <= b < tileptr[i+1]
   while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (tileptr[mid] <= b) lo = mid; else hi = mid;
   }
   const int cs = list[lo];
   const FrontDev& fc = fr[cs];
   const SlotStatic sc = ss[cs];
   const FrontDev& fp = fr[sc.parent];
   const int doff = doffs[lo];
   const int t = fc.m - fc.nelim, nd = fc.n - fc.nelim;
   const int ntr = (t + EA_TR - 1) / EA_TR;
   const int tile = (int)(b - tileptr[lo]);
   const int tr = tile % ntr, tc = tile / ntr;
   const int c0 = tc * EA_TC;
   if (tr * EA_TR + EA_TR - 1 < c0) return;   // tile entirely above the diagonal
   const int* relc = rel + sc.rel_off;
   const int pncol = fp.ncol, pdel = fp.ndelay_in;
   if (threadIdx.x < EA_TC) {
      int cc = c0 + threadIdx.x, d = 0;
      if (cc < t) {
         if (cc < nd) d = pncol + doff + cc;
         else { d = relc[cc - nd]; if (d >= pncol) d += pdel; }
      }
      sdc[threadIdx.x] = d;
   }
   __syncthreads();
   const int r = tr * EA_TR + threadIdx.x;
   if (r >= t) return;
   int dr;
   if (r < nd) dr = pncol + doff + r;
   else { dr = relc[r - nd]; if (dr >= pncol) dr += pdel; }
   const int cend = min(min(c0 + EA_TC, t), r + 1);
   for (int cc = c0; cc < cend; ++cc) {
      double v = *felem(fc, fc.nelim + r, fc.nelim + cc);
      int dc = sdc[cc - c0];
      int rr = dr > dc ? dr : dc, cl = dr > dc ? dc : dr;
      double* p = felem(fp, rr, cl);
      *p += v;
   }
}

// symmetric swap of two candidate columns/rows (every touched column lives in L); whole CTA
__device__ void block_sym_swap(const FrontDev& f, int a, int b) {
   if (a == b) return;
   if (a > b) { int s = a; a = b; b = s; }
   double* L = f.L;
   const size_t ld = f.ldl;
   for (int c = threadIdx.x; c < a; c += blockDim.x) {
      double t = L[a + c * ld]; L[a + c * ld] = L[b + c * ld]; L[b + c * ld] = t;
   }
   if (threadIdx.x == 0) {
      double t = L[a + a * ld]; L[a + a * ld] = L[b + b * ld]; L[b + b * ld] = t;
      int s = f.perm[a]; f.perm[a] = f.perm[b]; f.perm[b] = s;
      s = f.orig[a]; f.orig[a] = f.orig[b]; f.orig[b] = s;
   }
   for (int c = a + 1 + threadIdx.x; c < b; c += blockDim.x) {
      double t = L[c + a * ld]; L[c + a * ld] = L[b + c * ld]; L[b + c * ld] = t;
   }
   for (int r = b + 1 + threadIdx.x; r < f.m; r += blockDim.x) {
      double t = L[r + a * ld]; L[r + a * ld] = L[r + b * ld]; L[r + b * ld] = t;
   }
   __syncthreads();
}

__device__ int g_upd_debug = 0;     // B200LDLT_DEBUG_UPD=1: k_commit and k_upd3 print what they record / apply (diagnostic)
#include "factor_kernels.cuh"
#include "upd_kernels.cuh"
#include "front_kernels.cuh"
constexpr int UPD_SMEM = 2 * BK * (LDA + LDB) * (int)sizeof(double);

// ------------------------------------------------------------------------------------------
// TPP finish (ldlt_tpp.cxx:166-270), one CTA per front, right-looking on the whole trailing
// matrix.  Rare path: only fronts whose a-posteriori pass left columns behind come here.
// ------------------------------------------------------------------------------------------
__device__ double block_max(double v, double* sred) {
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
   for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
   __syncthreads();
   if (lane == 0) sred[wid] = v;
   __syncthreads();
   double r = sred[0];
   for (int i = 1; i < nw; ++i) r = fmax(r, sred[i]);
   return r;
}

// max |.| over row/col `col` of the trailing matrix from `from`, excluding `excl` and the diagonal
__device__ double tpp_rc_max(const FrontDev& f, int from, int col, int excl, double* sred) {
   double v = 0.0;
   for (int c = from + threadIdx.x; c < col; c += blockDim.x)
      if (c != excl) v = fmax(v, fabs(*felem(f, col, c)));
   for (int r = col + 1 + threadIdx.x; r < f.m; r += blockDim.x)
      if (r != excl) v = fmax(v, fabs(*felem(f, r, col)));
   return block_max(v, sred);
}

__device__ void tpp_elim1(const FrontDev& f, int ne, double d11) {
   double* l1 = f.Wld;   // column 0 of the workspace, front-row indexed
   const int m = f.m;
   for (int r = ne + 1 + threadIdx.x; r < m; r += blockDim.x) {
      double* p = f.L + (size_t)ne * f.ldl + r;
      double v = *p;
      l1[r] = v;
      *p = v * d11;
   }
   __syncthreads();
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
   for (int c = ne + 1 + wid; c < m; c += nw) {
      double b = l1[c];
      if (b == 0.0) continue;
      for (int r = c + lane; r < m; r += 32) *felem(f, r, c) -= f.L[(size_t)ne * f.ldl + r] * b;
   }
   __syncthreads();
   if (threadIdx.x == 0) { f.L[(size_t)ne * f.ldl + ne] = 1.0; f.D[2 * ne] = d11; f.D[2 * ne + 1] = 0.0; }
   __syncthreads();
}

__global__ void __launch_bounds__(256) k_tpp(FrontDev* fr, const int* list, Params prm) {
   FrontDev& f = fr[list[blockIdx.x]];
   __shared__ double sred[32];
   __shared__ int sidx[8];
   if (f.flag < 0) return;
   int ne = f.done;
   const int n = f.n, m = f.m;
   const size_t ldl = f.ldl;
   double* l1 = f.Wld;
   double* l2 = f.Wld + f.ldl;
   while (ne < n) {
      // zero column at ne?
      {
         double v = fabs(f.L[ne * ldl + ne]);
         if (threadIdx.x != 0) v = 0.0;
         v = fmax(v, tpp_rc_max(f, ne, ne, -1, sred));
         v = block_max(v, sred);
         if (v < prm.small_) {
            if (!prm.action) { if (threadIdx.x == 0) f.flag = SSIDS_ERROR_SINGULAR; break; }
            for (int r = ne + threadIdx.x; r < m; r += blockDim.x) f.L[ne * ldl + r] = 0.0;
            __syncthreads();
            if (threadIdx.x == 0) { f.L[ne * ldl + ne] = 1.0; f.D[2 * ne] = 0.0; f.D[2 * ne + 1] = 0.0; }
            ne++;
            __syncthreads();
            continue;
         }
      }
      bool found = false;
      for (int p = ne + 1; p < n && !found; ++p) {
         // zero column p?
         {
            double v = (threadIdx.x == 0) ? fabs(f.L[p * ldl + p]) : 0.0;
            v = fmax(v, tpp_rc_max(f, ne, p, -1, sred));
            v = block_max(v, sred);
            if (v < prm.small_) {
               if (!prm.action) { if (threadIdx.x == 0) f.flag = SSIDS_ERROR_SINGULAR; ne = n + 1; break; }
               block_sym_swap(f, ne, p);
               for (int r = ne + threadIdx.x; r < m; r += blockDim.x) f.L[ne * ldl + r] = 0.0;
               __syncthreads();
               if (threadIdx.x == 0) { f.L[ne * ldl + ne] = 1.0; f.D[2 * ne] = 0.0; f.D[2 * ne + 1] = 0.0; }
               ne++;
               found = true;
               __syncthreads();
               break;
            }
         }
         // t = argmax |A(p, ne..p-1)| (first maximum)
         double best = -1.0;
         int bt = ne;
         for (int c = ne + threadIdx.x; c < p; c += blockDim.x) {
            double v = fabs(f.L[c * ldl + p]);
            if (v > best) { best = v; bt = c; }
         }
         double gbest = block_max(best, sred);
         __syncthreads();
         if (threadIdx.x == 0) sidx[0] = 1 << 30;
         __syncthreads();
         if (best == gbest && best >= 0.0) atomicMin(&sidx[0], bt);
         __syncthreads();
         const int t = sidx[0];
         double maxt = tpp_rc_max(f, ne, t, p, sred);
         double maxp = tpp_rc_max(f, ne, p, t, sred);
         double a11 = f.L[t * ldl + t], a21 = f.L[t * ldl + p], a22 = f.L[p * ldl + p];
         double maxpiv = fmax(fabs(a11), fmax(fabs(a21), fabs(a22)));
         bool ok2 = false;
         double d11 = 0, d21 = 0, d22 = 0;
         if (maxpiv >= prm.small_) {
            double detscale = 1.0 / maxpiv;
            double detpiv0 = (a11 * detscale) * a22, detpiv1 = (a21 * detscale) * a21;
            double detpiv = detpiv0 - detpiv1;
            if (!(fabs(detpiv) < fmax(prm.small_, fmax(fabs(detpiv0 / 2), fabs(detpiv1 / 2))))) {
               d11 = (a22 * detscale) / detpiv; d21 = (-a21 * detscale) / detpiv; d22 = (a11 * detscale) / detpiv;
               if (fmax(maxp, maxt) < prm.small_) ok2 = true;
               else {
                  double x1 = fabs(d11) * maxt + fabs(d21) * maxp, x2 = fabs(d21) * maxt + fabs(d22) * maxp;
                  ok2 = (prm.u * fmax(x1, x2) < 1.0);
               }
            }
         }
         __syncthreads();
         if (ok2) {
            block_sym_swap(f, ne, t);
            block_sym_swap(f, ne + 1, p);
            for (int r = ne + 2 + threadIdx.x; r < m; r += blockDim.x) {
               double w1 = f.L[ne * ldl + r], w2 = f.L[(ne + 1) * ldl + r];
               l1[r] = w1; l2[r] = w2;
               f.L[ne * ldl + r] = d11 * w1 + d21 * w2;
               f.L[(ne + 1) * ldl + r] = d21 * w1 + d22 * w2;
            }
            __syncthreads();
            const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
            for (int c = ne + 2 + wid; c < m; c += nw) {
               double b1 = l1[c], b2 = l2[c];
               for (int r = c + lane; r < m; r += 32)
                  *felem(f, r, c) -= f.L[ne * ldl + r] * b1 + f.L[(ne + 1) * ldl + r] * b2;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
               f.L[ne * ldl + ne] = 1.0; f.L[ne * ldl + ne + 1] = 0.0; f.L[(ne + 1) * ldl + ne + 1] = 1.0;
               f.D[2 * ne] = d11; f.D[2 * ne + 1] = d21; f.D[2 * ne + 2] = DINF; f.D[2 * ne + 3] = d22;
            }
            ne += 2;
            found = true;
            __syncthreads();
            break;
         }
         maxp = fmax(maxp, fabs(a21));
         if (fabs(a22) >= prm.u * maxp) {
            block_sym_swap(f, ne, p);
            double dd = 1.0 / f.L[ne * ldl + ne];
            __syncthreads();
            tpp_elim1(f, ne, dd);
            ne += 1;
            found = true;
            break;
         }
      }
      if (ne > n) break;
      if (!found) {
         double maxp = tpp_rc_max(f, ne, ne, -1, sred);
         double app = f.L[ne * ldl + ne];
         __syncthreads();
         if (fabs(app) >= prm.u * maxp) {
            tpp_elim1(f, ne, 1.0 / app);
            ne += 1;
         } else {
            break;
         }
      }
   }
   __syncthreads();
   if (ne > n) ne = n;
   if (threadIdx.x == 0) { f.done = ne; f.nelim = ne; }
}

__global__ void k_collect(const FrontDev* fr, int slot0, int nslots, int* out) {
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= nslots) return;
   out[3 * i] = fr[slot0 + i].done;
   out[3 * i + 1] = fr[slot0 + i].flag;
   out[3 * i + 2] = fr[slot0 + i].gend;
}

// inertia: counters[0..2] = num_neg, num_two, num_zero  (NumericSubtree.hxx:242-274)
__global__ void k_stats(const FrontDev* fr, int nslots, int* counters) {
   const int lane = threadIdx.x & 31;
   const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
   if (s >= nslots) return;
   const FrontDev& f = fr[s];
   if (f.state != 2) return;
   int neg = 0, two = 0, zero = 0;
   for (int i = lane; i < f.nelim; i += 32) {
      double a11 = f.D[2 * i];
      if (isinf(a11) && i > 0) continue;   // second column of a 2x2
      if (i + 1 < f.nelim && isinf(f.D[2 * i + 2])) {
         double a21 = f.D[2 * i + 1], a22 = f.D[2 * i + 3];
         double det = a11 * a22 - a21 * a21, tr = a11 + a22;
         two++;
         if (det < 0) neg++;
         else if (tr < 0) neg += 2;
      } else {
         if (a11 == 0.0) zero++;
         if (a11 < 0.0) neg++;
      }
   }
#pragma unroll
   for (int off = 16; off > 0; off >>= 1) {
      neg += __shfl_xor_sync(0xffffffffu, neg, off);
      two += __shfl_xor_sync(0xffffffffu, two, off);
      zero += __shfl_xor_sync(0xffffffffu, zero, off);
   }
   if (lane == 0) {
      if (neg) atomicAdd(&counters[0], neg);
      if (two) atomicAdd(&counters[1], two);
      if (zero) atomicAdd(&counters[2], zero);
   }
}

#include "solve_kernels.cuh"

static inline size_t step_smem_fwd(int kbc) { return (size_t)(kbc * (kbc + 1) + kbc + 256) * 8; }
// the instantiations with t threads per CTA (64: very small fronts, 192: blocks of 65..96 pivots)
static inline size_t step_smem_fwd_t(int kbc, int t) { return (size_t)(kbc * (kbc + 1) + kbc + t) * 8; }
static inline size_t step_smem_bwd_t(int kbc, int t) { return (size_t)(kbc * (kbc + 1) + std::max(kbc, t * 32 / kbc) + 32) * 8; }
static inline size_t chain_smem() {
   return (size_t)(KB * (KB + 1) + KB + 2 * KB + 32 + 2 * KB + 4 * 32 * ILD + KB + 4 * KB + 32) * 8;
}
static inline int step_grid_y(int rows, int rc, int slots, int fronts) {
   return std::max(1, std::min((rows + rc - 1) / rc, slots / std::max(1, fronts)));
}
static inline size_t step_smem_bwd(int kbc) { return (size_t)(kbc * (kbc + 1) + std::max(kbc, SE / kbc) + 32) * 8; }

// SLS value gather on the device (SLS_factorize, sls.f90:4106-4147): VAL(MAPS(l)) = val(l) for the first entry of
// every position, duplicates added in input order, positions without an entry (missing diagonals) stay zero.
// set_*: the (source, destination) pairs with MAPS > 0; dup_*: per destination with duplicates, its sources in
// input order (one thread each: the summation order is the reference's).
__global__ void k_gather_set(double* __restrict__ val, const double* __restrict__ uval, const int* __restrict__ src,
                             const int* __restrict__ dst, int nset) {
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < nset) val[dst[i]] = uval[src[i]];
}
__global__ void k_gather_dup(double* __restrict__ val, const double* __restrict__ uval, const int* __restrict__ dup_dst,
                             const int* __restrict__ dup_ptr, const int* __restrict__ dup_src, int ndup) {
   int t = blockIdx.x * blockDim.x + threadIdx.x;
   if (t >= ndup) return;
   double acc = val[dup_dst[t]];
   for (int q = dup_ptr[t]; q < dup_ptr[t + 1]; ++q) acc += uval[dup_src[q]];
   val[dup_dst[t]] = acc;
}

// changed entries of the caller's values: uval[which[i]] = newval[i] (SBLS's interior-point loop changes only the barrier
// diagonal between two factorisations: sbls.f90:143-153 new_h / new_a / new_c)
__global__ void k_update_user_values(double* __restrict__ uval, const int* __restrict__ which, const double* __restrict__ newval, int nchg) {
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < nchg) uval[which[i]] = newval[i];
}

// a_ij <- s_i a_ij s_j on the values of the lower triangle (add_a_block with scaling, assemble.hxx:54-66)
__global__ void k_scale_values(double* __restrict__ val, const int* __restrict__ rowi, const int* __restrict__ coli,
                               const double* __restrict__ scale, int64_t nz) {
   int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (k < nz) val[k] = __dmul_rn(__dmul_rn(scale[rowi[k]], val[k]), scale[coli[k]]);
}

// Iterative refinement on the device (SLS_solve_ir, sls.f90:4751-4947).  The residual uses the caller's entries
// (coordinates + the values of the latest factor_mapped), row by row in input order, without contraction into
// FMAs: the same sequence of operations as the reference's loop over the entries.
__global__ void k_residual(double* __restrict__ r, const double* __restrict__ b, const double* __restrict__ x,
                           const double* __restrict__ uval, const int* __restrict__ rptr, const int* __restrict__ rent,
                           const int* __restrict__ rcol, int n) {
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   double acc = b[i];
   for (int q = rptr[i]; q < rptr[i + 1]; ++q) acc = __dsub_rn(acc, __dmul_rn(uval[rent[q]], x[rcol[q]]));
   r[i] = acc;
}
__global__ void k_axpy1(double* __restrict__ x, const double* __restrict__ r, int n) {
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) x[i] += r[i];
}
// max |v_i| into *out (non-negative doubles order like their bit patterns)
__global__ void k_maxabs(const double* __restrict__ v, int n, unsigned long long* out) {
   double m = 0.0;
   // fmax drops NaN: a non-finite entry must not read as a small residual (the refinement loop would report a NaN
   // solution as converged), so it counts as +inf
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const double a = fabs(v[i]);
      m = fmax(m, a <= 1.7976931348623157e308 ? a : DINF);
   }
#pragma unroll
   for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
   if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

// multi-GPU: the lower triangle of a contribution block (t columns, column j rows j..t-1) to / from a packed trapezoid:
// half the bytes of the full square over NVLink
__global__ void k_pack_lower(double* __restrict__ dst, const double* __restrict__ C, int ldc, int t) {
   const int j = blockIdx.x;
   const double* col = C + (size_t)j * ldc;
   double* out = dst + ((size_t)j * t - (size_t)j * (j - 1) / 2) - j;       // out[r] for r >= j
   for (int r = j + threadIdx.x; r < t; r += blockDim.x) out[r] = col[r];
}
__global__ void k_unpack_lower(double* __restrict__ C, const double* __restrict__ src, int ldc, int t) {
   const int j = blockIdx.x;
   double* col = C + (size_t)j * ldc;
   const double* in = src + ((size_t)j * t - (size_t)j * (j - 1) / 2) - j;
   for (int r = threadIdx.x; r < ldc; r += blockDim.x) col[r] = (r >= j && r < t) ? in[r] : 0.0;
}

// multi-GPU: which pivots did this rank eliminate, and in which stage
__global__ void k_build_masks(const FrontDev* fr, int nslots, char* mine, char* stage) {
   const int lane = threadIdx.x & 31;
   const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
   if (s >= nslots) return;
   const FrontDev& f = fr[s];
   if (f.state != 2) return;
   for (int k = lane; k < f.nelim; k += 32) { mine[f.perm[k]] = 1; stage[f.perm[k]] = (char)f.part_kind; }
}
// dst[0..n) = the entries of xp this rank produced in `stage`, zero elsewhere; dst[n..2n) = 1 where produced.
// Both halves are summed over the ranks; k_merge_stage then takes the entries some rank produced.
__global__ void k_get_stage(double* __restrict__ dst, const double* __restrict__ xp, const char* __restrict__ mine,
                            const char* __restrict__ stage, int st, int n) {
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) {
      const bool p = mine[i] && stage[i] == st;
      dst[i] = p ? xp[i] : 0.0;
      dst[n + i] = p ? 1.0 : 0.0;
   }
}
__global__ void k_merge_stage(double* __restrict__ xp, const double* __restrict__ src, int n) {
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n && src[n + i] != 0.0) xp[i] = src[i];
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
// Launch as a programmatic dependent of the previous kernel in the stream (pdl_enter in the kernel): the grid is resident
// and past its launch latency when its predecessor completes.
template <typename... KArgs, typename... Args>
static inline void launch_dep(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
   cudaLaunchConfig_t cfg = {};
   cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
   cudaLaunchAttribute at[1];
   at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
   at[0].val.programmaticStreamSerializationAllowed = 1;
   cfg.attrs = at;
   cfg.numAttrs = pdl ? 1 : 0;
   cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int block_wc(int n) { return std::max(NB, (std::min(n, PB) + NB - 1) / NB * NB); }
// columns of a front's X ring: every pivot its own column up to 2 * GMAX blocks, where a group's deferred update reads
// one half while the next group's blocks write the other
constexpr int GMAX = 4;          // blocks per deferred trailing update (rank GMAX * PB)
static inline int ring_cols(int n) { return n <= PB ? block_wc(n) : std::min((n + PB - 1) / PB * PB, 2 * GMAX * PB); }
// doubles of block workspace of a front: X ring | Wo | L11 | A0 | dk | Linv | ibuf
static inline size_t front_ws_doubles(size_t ldl, int n) {
   const size_t wc = (size_t)block_wc(n);
   return align_up(ldl * ((size_t)ring_cols(n) + wc) + 2 * (size_t)PB * wc + 2 * PB + 2 + (size_t)NB * wc + (2 * PB + 8) / 2, 32) + 128;
}
static inline int size_class(int m) { return m <= 64 ? 0 : m <= 256 ? 1 : m <= 1024 ? 2 : m <= 4096 ? 3 : 4; }

struct Arena {
   struct Chunk { char* p; size_t size; };
   std::vector<Chunk> chunks;
   size_t cur = 0, off = 0, min_chunk = (size_t)64 << 20;
   int* err = nullptr;
   void* alloc(size_t bytes) {
      bytes = align_up(bytes ? bytes : 1, 256);
      for (;;) {
         if (cur < chunks.size() && off + bytes <= chunks[cur].size) {
            void* p = chunks[cur].p + off;
            off += bytes;
            return p;
         }
         if (cur + 1 < chunks.size()) { ++cur; off = 0; continue; }
         size_t sz = std::max(min_chunk, bytes);
         char* p = nullptr;
         cudaError_t e = cudaMalloc(&p, sz);
         if (e != cudaSuccess) { if (err) *err = (int)e; return nullptr; }
         chunks.push_back({p, sz});
         cur = chunks.size() - 1;
         off = 0;
      }
   }
   void reserve(size_t bytes) {   // make sure the first chunk can hold `bytes`
      if (!chunks.empty()) return;
      bytes = align_up(std::max(bytes, (size_t)256), 256);
      char* p = nullptr;
      if (cudaMalloc(&p, bytes) == cudaSuccess) chunks.push_back({p, bytes});
      else cudaGetLastError();
   }
   void reset() { cur = 0; off = 0; }
   size_t capacity() const { size_t s = 0; for (auto& c : chunks) s += c.size; return s; }
   void release() { for (auto& c : chunks) cudaFree(c.p); chunks.clear(); reset(); }
};

class Plan {
  public:
   const Symbolic* sym = nullptr;
   int device = 0;
   int nnodes = 0, nlevels = 0;
   std::vector<int> lvl_ptr, slot_node, node_slot;
   std::vector<SlotStatic> hss;
   std::vector<int> child_idx;
   std::vector<int64_t> lvl_a_ptr;
   int max_children = 0;
   // device
   SlotStatic* d_ss = nullptr;
   int *d_child_idx = nullptr, *d_rlist = nullptr, *d_rel = nullptr, *d_order = nullptr;
   int *d_aslot = nullptr, *d_asrc = nullptr, *d_acol = nullptr, *d_arow = nullptr;
   int64_t* d_aptr = nullptr;        // per slot: its range of the A scatter map (k_front)
   int64_t est_factor_doubles = 0, est_c_doubles[2] = {0, 0}, est_ws_doubles = 0;
   std::vector<int64_t> lvl_L_doubles, lvl_C_doubles;   // per level, without delays (same rounding as the numeric phase)
   ~Plan() {
      cudaSetDevice(device);
      cudaFree(d_ss); cudaFree(d_child_idx); cudaFree(d_rlist); cudaFree(d_rel); cudaFree(d_order);
      cudaFree(d_aslot); cudaFree(d_asrc); cudaFree(d_acol); cudaFree(d_arow); cudaFree(d_aptr);
   }
};

template <typename T>
static cudaError_t upload(T** d, const std::vector<T>& h) {
   size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
   cudaError_t e = cudaMalloc(d, bytes);
   if (e != cudaSuccess) return e;
   if (!h.empty()) e = cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
   return e;
}

Plan* plan_create(const Symbolic& sym, int device, int* cuda_error) {
   *cuda_error = 0;
   int ndev = 0;
   cudaError_t e = cudaGetDeviceCount(&ndev);
   if (e != cudaSuccess || ndev == 0) { *cuda_error = e != cudaSuccess ? (int)e : (int)cudaErrorNoDevice; cudaGetLastError(); return nullptr; }
   if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
   if ((e = cudaSetDevice(device)) != cudaSuccess) { *cuda_error = (int)e; return nullptr; }
   Plan* P = new Plan();
   P->sym = &sym;
   P->device = device;
   const int nn = sym.nnodes;
   P->nnodes = nn;
   // depth levels, deepest first
   std::vector<int> depth(nn + 1, 0);
   int maxd = 0;
   for (int i = nn - 1; i >= 0; --i) {
      depth[i] = depth[sym.sparent[i]] + 1;
      maxd = std::max(maxd, depth[i]);
   }
   P->nlevels = maxd;
   std::vector<int> lvl(nn);
   P->lvl_ptr.assign(maxd + 1, 0);
   for (int i = 0; i < nn; ++i) { lvl[i] = maxd - depth[i]; P->lvl_ptr[lvl[i] + 1]++; }
   for (int l = 0; l < maxd; ++l) P->lvl_ptr[l + 1] += P->lvl_ptr[l];
   P->slot_node.assign(nn, 0);
   P->node_slot.assign(nn, 0);
   {
      std::vector<int> pos(P->lvl_ptr.begin(), P->lvl_ptr.end() - 1);
      for (int i = 0; i < nn; ++i) { int s = pos[lvl[i]]++; P->slot_node[s] = i; }
      // within a level: size class of the front, then wider fronts first, so that a bucket of the
      // factorisation is a contiguous slot range (no indirection list on the hot path)
      auto nrow_of = [&](int i) { return (int)(sym.rptr[i + 1] - sym.rptr[i]); };
      auto ncol_of = [&](int i) { return sym.sptr[i + 1] - sym.sptr[i]; };
      for (int l = 0; l < maxd; ++l)
         std::sort(P->slot_node.begin() + P->lvl_ptr[l], P->slot_node.begin() + P->lvl_ptr[l + 1], [&](int a, int b) {
            int ca = size_class(nrow_of(a)), cb = size_class(nrow_of(b));
            if (ca != cb) return ca < cb;
            if (ncol_of(a) != ncol_of(b)) return ncol_of(a) > ncol_of(b);
            return a < b;
         });
      for (int s = 0; s < nn; ++s) P->node_slot[P->slot_node[s]] = s;
   }
   // static per-slot data, children (ascending node order), relative indices
   P->hss.assign(nn, SlotStatic());
   std::vector<int> nchild(nn + 1, 0);
   for (int i = 0; i < nn; ++i) nchild[sym.sparent[i]]++;
   std::vector<int> cptr(nn + 1, 0);
   {
      int acc = 0;
      for (int s = 0; s < nn; ++s) { cptr[s] = acc; acc += nchild[P->slot_node[s]]; }
      cptr[nn] = acc;
      P->child_idx.assign(acc, 0);
      std::vector<int> fill(cptr.begin(), cptr.end() - 1);
      for (int i = 0; i < nn; ++i) {
         int p = sym.sparent[i];
         if (p < nn) P->child_idx[fill[P->node_slot[p]]++] = P->node_slot[i];
      }
   }
   std::vector<int> rel;
   int64_t rel_total = 0;
   for (int i = 0; i < nn; ++i) rel_total += (sym.rptr[i + 1] - sym.rptr[i]) - (sym.sptr[i + 1] - sym.sptr[i]);
   rel.assign(rel_total, 0);
   {
      std::vector<int> map(sym.n, 0);
      std::vector<int64_t> rel_off(nn, 0);
      int64_t acc = 0;
      for (int s = 0; s < nn; ++s) {
         int i = P->slot_node[s];
         rel_off[i] = acc;
         acc += (sym.rptr[i + 1] - sym.rptr[i]) - (sym.sptr[i + 1] - sym.sptr[i]);
      }
      for (int s = 0; s < nn; ++s) {
         int i = P->slot_node[s];
         SlotStatic& st = P->hss[s];
         st.ncol = sym.sptr[i + 1] - sym.sptr[i];
         st.nrow = (int)(sym.rptr[i + 1] - sym.rptr[i]);
         st.parent = sym.sparent[i] < nn ? P->node_slot[sym.sparent[i]] : -1;
         st.first_piv = sym.sptr[i];
         st.rl_off = sym.rptr[i];
         st.rel_off = rel_off[i];
         st.child_begin = cptr[s];
         st.child_end = cptr[s + 1];
         P->max_children = std::max(P->max_children, st.child_end - st.child_begin);
      }
      for (int p = 0; p < nn; ++p) {
         int sp = P->node_slot[p];
         const SlotStatic& st = P->hss[sp];
         if (st.child_begin == st.child_end) continue;
         for (int64_t j = sym.rptr[p]; j < sym.rptr[p + 1]; ++j) map[sym.rlist[j]] = (int)(j - sym.rptr[p]);
         for (int ci = st.child_begin; ci < st.child_end; ++ci) {
            int c = P->slot_node[P->child_idx[ci]];
            int cn = sym.sptr[c + 1] - sym.sptr[c];
            int64_t o = rel_off[c];
            int npm = 0;
            for (int64_t j = sym.rptr[c] + cn; j < sym.rptr[c + 1]; ++j) { rel[o++] = map[sym.rlist[j]]; if (map[sym.rlist[j]] < st.ncol) ++npm; }
            P->hss[P->child_idx[ci]].npm = npm;
         }
      }
   }
   // A scatter map, grouped by level
   int64_t na = sym.nptr[nn];
   std::vector<int> aslot(na), asrc(na), acol(na), arow(na);
   std::vector<int64_t> aptr(nn + 1, 0);
   P->lvl_a_ptr.assign(P->nlevels + 1, 0);
   {
      int64_t o = 0;
      for (int l = 0; l < P->nlevels; ++l) {
         P->lvl_a_ptr[l] = o;
         for (int s = P->lvl_ptr[l]; s < P->lvl_ptr[l + 1]; ++s) {
            int i = P->slot_node[s];
            int64_t blkm = sym.rptr[i + 1] - sym.rptr[i];
            aptr[s] = o;
            for (int64_t e2 = sym.nptr[i]; e2 < sym.nptr[i + 1]; ++e2) {
               int64_t dest = sym.nlist[2 * e2 + 1];
               aslot[o] = s;
               asrc[o] = (int)sym.nlist[2 * e2];
               acol[o] = (int)(dest / blkm);
               arow[o] = (int)(dest % blkm);
               ++o;
            }
         }
      }
      P->lvl_a_ptr[P->nlevels] = o;
      aptr[nn] = o;
   }
   // size estimates without delays
   for (int l = 0; l < P->nlevels; ++l) {
      int64_t c = 0, ws = 0;
      for (int s = P->lvl_ptr[l]; s < P->lvl_ptr[l + 1]; ++s) {
         const SlotStatic& st = P->hss[s];
         int64_t ldl = align_up(st.nrow, 2), cm = st.nrow - st.ncol;
         P->est_factor_doubles += ldl * st.ncol + 32;
         c += align_up(cm, 2) * cm + 32;
         ws += (int64_t)front_ws_doubles(align_up(st.nrow, 2), st.ncol);
      }
      P->est_c_doubles[l & 1] = std::max(P->est_c_doubles[l & 1], c);
      {
         int64_t tl = 0, tc = 0;
         for (int s = P->lvl_ptr[l]; s < P->lvl_ptr[l + 1]; ++s) {
            const SlotStatic& st2 = P->hss[s];
            tl += (int64_t)align_up((size_t)align_up(st2.nrow, 2) * st2.ncol, 32);
            tc += (int64_t)align_up((size_t)align_up(st2.nrow - st2.ncol, 2) * (st2.nrow - st2.ncol), 32);
         }
         P->lvl_L_doubles.push_back(tl);
         P->lvl_C_doubles.push_back(tc);
      }
      P->est_ws_doubles = std::max(P->est_ws_doubles, ws);
   }
   std::vector<int> order(sym.order);
   bool ok = upload(&P->d_ss, P->hss) == cudaSuccess && upload(&P->d_child_idx, P->child_idx) == cudaSuccess &&
             upload(&P->d_rlist, sym.rlist) == cudaSuccess && upload(&P->d_rel, rel) == cudaSuccess &&
             upload(&P->d_order, order) == cudaSuccess && upload(&P->d_aslot, aslot) == cudaSuccess &&
             upload(&P->d_asrc, asrc) == cudaSuccess && upload(&P->d_acol, acol) == cudaSuccess &&
             upload(&P->d_arow, arow) == cudaSuccess && upload(&P->d_aptr, aptr) == cudaSuccess;
   if (!ok) {
      *cuda_error = (int)cudaGetLastError();
      if (*cuda_error == 0) *cuda_error = (int)cudaErrorMemoryAllocation;
      delete P;
      return nullptr;
   }
   return P;
}

void plan_destroy(Plan* p) { delete p; }
int plan_device(const Plan* p) { return p->device; }

class Numeric {
  public:
   Plan* P;
   bool posdef = false, have = false;
   std::vector<FrontDev> hfr;
   FrontDev* d_fr = nullptr;
   bool inited = false, init_failed = false;   // init() ran to its end / stopped at a failed allocation
   FrontDev* pin_fr = nullptr;       // pinned staging, nslots entries
   int* pin_i = nullptr;             // pinned ints: collect + lists staging
   size_t pin_i_cap = 0;
   int *d_child_doff = nullptr, *d_list = nullptr, *d_collect = nullptr, *d_counters = nullptr;
   int64_t* d_tileptr = nullptr;
   int *d_updmap = nullptr, *pin_updmap = nullptr;   // CTA -> (front, part, parts) maps of the trailing updates
   static constexpr size_t UPDMAP_CAP = (size_t)3 << 20;   // ints
   int64_t* d_xw_off = nullptr;
   int64_t* d_part_off = nullptr;
   // chain kernels (solve_kernels.cuh): role tables per level and direction, block flags, ticket counters
   int2* d_roles = nullptr;
   size_t roles_cap = 0;
   double* d_cy = nullptr;          // blocks published by the chain kernels of the running sweep (indexed like d_xw)
   size_t cy_cap = 0;
   bool have_chain = false;
   unsigned long long* d_chain_counter = nullptr;   // [2 * nlevels]
   int* d_chain_err = nullptr;
   unsigned long long* d_chain_dbg = nullptr;       // B200LDLT_SOLVE_TRACE: four time stamps per chain role
   // Several right-hand sides: the columns are solved on up to SOLVE_STREAMS streams at once (8: 10 % over 4 on eight columns), each with a set of solve
   // workspaces of its own (the sweeps read the workspace pointers and the stream from the members below, which the host
   // swaps between enqueues: a solve of one column is bound by dependent launches, not by the machine)
   struct SolveWs {
      cudaStream_t st = nullptr;
      double *x = nullptr, *xp = nullptr, *xw = nullptr, *cy = nullptr, *part = nullptr, *tail = nullptr;
      int* ticket = nullptr;
      unsigned long long* counter = nullptr;
      cudaEvent_t done = nullptr;
   };
   static constexpr int SOLVE_STREAMS = 8;
   SolveWs extra_ws[SOLVE_STREAMS - 1];
   int n_extra_ws = 0;
   size_t extra_sig[6] = {0, 0, 0, 0, 0, 0};     // the capacities the extra sets were allocated for
   cudaEvent_t ev_ws_fork = nullptr;
   SolveWs current_ws() const {
      SolveWs w;
      w.st = st; w.x = d_x; w.xp = d_xp; w.xw = d_xw; w.cy = d_cy; w.part = d_part; w.tail = d_tail; w.ticket = d_ticket; w.counter = d_chain_counter;
      return w;
   }
   void use_ws(const SolveWs& w) {
      st = w.st; d_x = w.x; d_xp = w.xp; d_xw = w.xw; d_cy = w.cy; d_part = w.part; d_tail = w.tail; d_ticket = w.ticket; d_chain_counter = w.counter;
   }
   void free_extra_ws() {
      for (int i = 0; i < n_extra_ws; ++i) {
         SolveWs& w = extra_ws[i];
         cudaFree(w.x); cudaFree(w.xp); cudaFree(w.xw); cudaFree(w.cy); cudaFree(w.part); cudaFree(w.tail); cudaFree(w.ticket); cudaFree(w.counter);
         if (w.st) cudaStreamDestroy(w.st);
         if (w.done) cudaEventDestroy(w.done);
         w = SolveWs();
      }
      n_extra_ws = 0;
   }
   int ensure_extra_ws(int want);
   // B200LDLT_SOLVE_TIMELINE: an event after every launch of a solve, printed per launch at its end
   bool solve_tl = getenv("B200LDLT_SOLVE_TIMELINE") != nullptr;
   std::vector<cudaEvent_t> stl_ev;
   std::vector<std::string> stl_name;
   void smark(const char* what, int l) {
      if (!solve_tl) return;
      cudaEvent_t e;
      cudaEventCreate(&e);
      cudaEventRecord(e, st);
      stl_ev.push_back(e);
      stl_name.push_back(std::string(what) + " " + std::to_string(l));
   }
   // chain fronts: inverses of the 128 x 128 diagonal blocks (k_chain_inverses, once per factorisation) and the partial
   // sums of the rows below the pivots for the backward sweep (k_bwd_tail)
   double* d_minv = nullptr; size_t minv_cap = 0;   // blocks of KB x KB doubles
   int* d_minv_off = nullptr;                       // per slot: first block, -1 = none
   int* d_mflag = nullptr; size_t mflag_cap = 0;
   int2* d_blist = nullptr;
   bool have_minv = false;
   double* d_tail = nullptr; size_t tail_cap = 0;
   int64_t* d_tail_off = nullptr;                   // per slot: offset into d_tail, -1 = none
   int4* d_troles = nullptr; size_t troles_cap = 0;
   std::vector<int64_t> troles_off;
   std::vector<int> troles_n;
   bool have_tail = false;
   void chain_trace(const char* what, int l, int nroles, int off);
   std::vector<int64_t> roles_fwd_off, roles_bwd_off;
   std::vector<int> roles_fwd_n, roles_bwd_n;
   bool use_fused = !(getenv("B200LDLT_FUSED") && atoi(getenv("B200LDLT_FUSED")) == 0);
   bool use_chain = !(getenv("B200LDLT_CHAIN") && atoi(getenv("B200LDLT_CHAIN")) == 0);
   // chain kernels: 32 x 32 diagonal sub-triangles through explicit inverses formed while a role waits (solve_kernels.cuh)
   // chain kernels of the solve: 0 substitution, 1 the 32 x 32 sub-triangles through their inverses, 2 the whole 128 x 128
   // diagonal block through its inverse and one warp polling per 32 published values
   int use_solve_inv = getenv("B200LDLT_SOLVE_INV") ? atoi(getenv("B200LDLT_SOLVE_INV")) : 2;
   // fused levels of very small fronts (at most SMALL_T_MAXM rows) in CTAs of 64 threads
   bool use_st64 = !(getenv("B200LDLT_SOLVE_T64") && atoi(getenv("B200LDLT_SOLVE_T64")) == 0);
   bool split_classes = !(getenv("B200LDLT_SOLVE_SPLIT") && atoi(getenv("B200LDLT_SOLVE_SPLIT")) == 0);
   // levels with a front of this many rows go through the chain kernels whatever their pivot counts
   int chain_tall_m = getenv("B200LDLT_CHAIN_TALL") ? atoi(getenv("B200LDLT_CHAIN_TALL")) : 4096;
   int* d_ticket = nullptr;         // per slot: arrival counter of the backward step's CTAs (left at zero by the last one)
   int* d_slist = nullptr;
   std::vector<int> slist, lvl_steps, lvl_maxm, lvl_maxne;
   std::vector<int> lvl_ngt32, lvl_ngt64;     // fronts of the level with more than 32 / 64 eliminated columns (a prefix of its sorted list)
   // the first `total` fronts of level l's list by width class (128 / 64 / 32 columns of triangle per CTA): a class is split
   // off when it is the minority of what is left, else it takes the narrower fronts along
   void split_by_width(int l, int kb0, int total, int& n128, int& n64, int& n32) const {
      n128 = kb0 > 64 ? std::min(lvl_ngt64[l], total) : 0;
      if (kb0 > 64 && (!split_classes || 2 * n128 > total)) n128 = total;
      const int rest = total - n128;
      n64 = kb0 > 32 ? std::min(std::max(0, lvl_ngt32[l] - (kb0 > 64 ? lvl_ngt64[l] : 0)), rest) : 0;
      if (kb0 > 32 && kb0 <= 64 && n64 == 0) n64 = rest;
      if (n64 > 0 && (!split_classes || 2 * n64 > rest)) n64 = rest;
      n32 = rest - n64;
   }
   std::vector<std::vector<int>> lvl_nact, lvl_actm;
   int64_t part_total = 0;
   double* d_part = nullptr;
   size_t part_cap = 0;
   std::vector<int> child_doff;
   double* d_val = nullptr;
   // SLS value map (set_value_map): device copies and the user's values of the current factorisation
   double* d_uval = nullptr;
   int *d_vm_src = nullptr, *d_vm_dst = nullptr, *d_vm_dup_dst = nullptr, *d_vm_dup_ptr = nullptr, *d_vm_dup_src = nullptr;
   int vm_ne = 0, vm_nset = 0, vm_ndup = 0;
   int set_value_map(int ne, const int* maps);
   // scaling of the factorisation (variable order), the A pattern as (row, column) per entry
   double* d_scale = nullptr;
   int *d_a_rowi = nullptr, *d_a_coli = nullptr;
   bool scaled = false;
   int set_scaling(const double* scale);
   // refinement: the caller's entries by row (set_user_matrix), work vectors
   int *d_um_rptr = nullptr, *d_um_rent = nullptr, *d_um_rcol = nullptr;
   double *d_rb = nullptr, *d_rx = nullptr, *d_rr = nullptr;
   unsigned long long* d_rnorm = nullptr;
   int um_ne = -1;
   int set_user_matrix(int ne, const int* row, const int* col);
   int solve_refined(double* x, int max_ref, double stop_abs, double stop_rel, int* nref, double* seconds);
   int gather_values(int ne, const double* uval);
   int update_values(int nchg, const int* which, const double* newval);
   int *d_chg_idx = nullptr;
   double* d_chg_val = nullptr;
   size_t chg_cap = 0;
   bool uval_valid = false;       // d_uval holds the values of a complete factor_mapped call
   double *d_x = nullptr, *d_xp = nullptr, *d_xw = nullptr;
   size_t x_cap = 0, xw_cap = 0;
   int64_t xw_total = 0;
   Arena fac, dar, par, car[2], war, bar;   // bar: contribution blocks that cross a part boundary + imports
   std::vector<char> active, boundary;
   std::vector<char> live;                 // per slot, during a level: active and not taken by k_front
   std::vector<char> lvl_fusable;          // per level: every front fits k_front when no column is delayed into it
   int *d_flist = nullptr, *pin_flist = nullptr;   // fronts of a level taken by k_front: [n <= 32 | n <= 64]
   std::vector<int> part_kind, owner;      // per slot; owner -1 = root part
   int my_rank = 0, nranks = 1;
   cudaStream_t st = nullptr;
   static constexpr int NAUX = 4;
   cudaStream_t aux[NAUX] = {nullptr, nullptr, nullptr, nullptr};
   cudaEvent_t ev_fork = nullptr, ev_join[NAUX] = {nullptr, nullptr, nullptr, nullptr};
   // zero-fill of the next level's L and C regions on a side stream while this level is factorised
   cudaStream_t zs = nullptr;
   cudaEvent_t ev_ea = nullptr, ev_zero = nullptr;
   char *prez_c = nullptr, *prez_l = nullptr;
   size_t prez_c_bytes = 0, prez_l_bytes = 0;
   // look-ahead: bulk-update stream and events per chunk stream (0: st, 1..NAUX: aux[g-1])
   cudaStream_t bulk[NAUX + 1] = {nullptr, nullptr, nullptr, nullptr, nullptr};
   cudaEvent_t ev_commit[NAUX + 1][2] = {}, ev_bulk[NAUX + 1][2] = {};
   cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
   int cuda_err = 0;
   int64_t launches = 0;
   FactorStats last;
   std::vector<cudaEvent_t> prof_ev;
   size_t prof_used = 0;
   double prof_flops = 0;
   explicit Numeric(Plan* p) : P(p) {}
   ~Numeric() {
      cudaSetDevice(P->device);
      fac.release(); dar.release(); par.release(); car[0].release(); car[1].release(); war.release(); bar.release();
      cudaFree(d_fr); cudaFree(d_child_doff); cudaFree(d_list); cudaFree(d_collect); cudaFree(d_counters);
      free_extra_ws();
      if (ev_ws_fork) cudaEventDestroy(ev_ws_fork);
      cudaFree(d_root_mask); cudaFree(d_mine_mask); cudaFree(d_tileptr); cudaFree(d_xw_off); cudaFree(d_part_off); cudaFree(d_ticket); cudaFree(d_roles); cudaFree(d_cy); cudaFree(d_chain_counter); cudaFree(d_chain_err); cudaFree(d_chain_dbg); cudaFree(d_minv); cudaFree(d_minv_off); cudaFree(d_mflag); cudaFree(d_blist); cudaFree(d_tail); cudaFree(d_tail_off); cudaFree(d_troles); cudaFree(d_slist); cudaFree(d_scale); cudaFree(d_a_rowi); cudaFree(d_a_coli); cudaFree(d_um_rptr); cudaFree(d_um_rent); cudaFree(d_um_rcol); cudaFree(d_rb); cudaFree(d_rx); cudaFree(d_rr); cudaFree(d_rnorm);
      cudaFree(d_chg_idx); cudaFree(d_chg_val);
      cudaFree(d_part); cudaFree(d_val); cudaFree(d_uval); cudaFree(d_vm_src); cudaFree(d_vm_dst); cudaFree(d_vm_dup_dst); cudaFree(d_vm_dup_ptr); cudaFree(d_vm_dup_src); cudaFree(d_x); cudaFree(d_xp); cudaFree(d_xw);
      if (pin_fr) cudaFreeHost(pin_fr);
      if (pin_i) cudaFreeHost(pin_i);
      if (pin_updmap) cudaFreeHost(pin_updmap);
      if (pin_flist) cudaFreeHost(pin_flist);
      cudaFree(d_flist);
      cudaFree(d_updmap);
      if (st) cudaStreamDestroy(st);
      for (int g = 0; g < NAUX; ++g) { if (aux[g]) cudaStreamDestroy(aux[g]); if (ev_join[g]) cudaEventDestroy(ev_join[g]); }
      if (ev_fork) cudaEventDestroy(ev_fork);
      if (zs) cudaStreamDestroy(zs);
      if (ev_ea) cudaEventDestroy(ev_ea);
      if (ev_zero) cudaEventDestroy(ev_zero);
      for (int g = 0; g <= NAUX; ++g) {
         if (bulk[g]) cudaStreamDestroy(bulk[g]);
         for (int q = 0; q < 2; ++q) { if (ev_commit[g][q]) cudaEventDestroy(ev_commit[g][q]); if (ev_bulk[g][q]) cudaEventDestroy(ev_bulk[g][q]); }
      }
      if (ev0) cudaEventDestroy(ev0);
      if (ev1) cudaEventDestroy(ev1);
      if (ev2) cudaEventDestroy(ev2);
   }
   int init();
   int nstages = 1;
   int factor(const double* val, bool on_device, const FactorOptions& o, FactorStats& stats, int phase = 0);
   int set_partition(int rank, int nr);
   int finish_factor(FactorStats& stats, int flag);
   int solve(int job, int nrhs, double* x, int ldx, bool on_device, double* seconds, int64_t* nl);
   int solve_prepare();
   int build_solve_tables();
   bool tables_dirty = true;
   std::vector<int> child_doff_final;
   void sweep_fwd(int kind, int64_t& nlaunch);
   void sweep_bwd(int kind, bool diag, int64_t& nlaunch);
   int64_t contrib_bytes(int node) const;
   int contrib_pack(int node, char* dst);
   int contrib_unpack(int node, const char* src, int64_t bytes);
   int64_t xtail_bytes(int node) const;
   int xtail_pack(int node, char* dst);
   int xtail_unpack(int node, const char* src);
   int dist_solve(int step, double* d_xio, const double* src);
   char *d_root_mask = nullptr, *d_mine_mask = nullptr;
};

#define CU(x)                                                          \
   do {                                                                \
      cudaError_t e_ = (x);                                            \
      if (e_ != cudaSuccess) {                                         \
         cuda_err = (int)e_;                                           \
         cudaGetLastError();                                           \
         return SSIDS_ERROR_CUDA_UNKNOWN;                              \
      }                                                                \
   } while (0)

int Numeric::init() {
   if (inited) return 0;
   if (init_failed) return SSIDS_ERROR_ALLOCATION;   // a half-built handle is never used: free it and start again
   init_failed = true;                               // cleared by the last line; any CU() failure below leaves it set
   CU(cudaSetDevice(P->device));
   CU(cudaFuncSetAttribute(k_update, cudaFuncAttributeMaxDynamicSharedMemorySize, UPD_SMEM));
   CU(cudaFuncSetAttribute(kt_upd3<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, upd3::SMEM));
   CU(cudaFuncSetAttribute(kt_upd3<upd3::STM>, cudaFuncAttributeMaxDynamicSharedMemorySize, upd3::Shape<upd3::STM>::SMEM));
   { const int dbg = getenv("B200LDLT_DEBUG_UPD") ? 1 : 0; CU(cudaMemcpyToSymbol(g_upd_debug, &dbg, sizeof(int))); }
   CU(cudaFuncSetAttribute(k_block<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)block_smem_bytes(PB)));
   CU(cudaFuncSetAttribute(k_block<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)block_smem_bytes(2 * NB)));
   CU(cudaFuncSetAttribute(k_block<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)block_smem_bytes(NB)));
   CU(cudaFuncSetAttribute(k_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)apply_smem_bytes(PB)));
   CU(cudaFuncSetAttribute(k_front<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ffront::SMEM_CAP));
   CU(cudaFuncSetAttribute(k_front<2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ffront::SMEM_CAP));
   CU(cudaFuncSetAttribute(k_front<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ffront::SMEM_CAP));
   CU(cudaFuncSetAttribute(k_front<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ffront::SMEM_CAP));
   CU(cudaFuncSetAttribute(k_fwd_step<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_smem_fwd(128)));
   CU(cudaFuncSetAttribute(k_fwd_step<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_smem_fwd(128)));
   CU(cudaFuncSetAttribute(k_fwd_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chain_smem()));
   CU(cudaFuncSetAttribute(k_bwd_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chain_smem()));
   CU(cudaFuncSetAttribute(k_chain_inverses, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chain_inv_smem()));
   CU(cudaFuncSetAttribute(k_fwd_step<96, true, 192>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_smem_fwd_t(96, 192)));
   CU(cudaFuncSetAttribute(k_bwd_step<96, true, 192>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_smem_bwd_t(96, 192)));
   CU(cudaFuncSetAttribute(k_bwd_step<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_smem_bwd(128)));
   CU(cudaFuncSetAttribute(k_bwd_step<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_smem_bwd(128)));
   const int nn = std::max(P->nnodes, 1);
   CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
   for (int g = 0; g < NAUX; ++g) {
      CU(cudaStreamCreateWithFlags(&aux[g], cudaStreamNonBlocking));
      CU(cudaEventCreateWithFlags(&ev_join[g], cudaEventDisableTiming));
   }
   CU(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
   CU(cudaStreamCreateWithFlags(&zs, cudaStreamNonBlocking));
   CU(cudaEventCreateWithFlags(&ev_ea, cudaEventDisableTiming));
   CU(cudaEventCreateWithFlags(&ev_zero, cudaEventDisableTiming));
   for (int g = 0; g <= NAUX; ++g) {
      if (!bulk[g]) CU(cudaStreamCreateWithFlags(&bulk[g], cudaStreamNonBlocking));
      for (int q = 0; q < 2; ++q) {
         CU(cudaEventCreateWithFlags(&ev_commit[g][q], cudaEventDisableTiming));
         CU(cudaEventCreateWithFlags(&ev_bulk[g][q], cudaEventDisableTiming));
      }
   }
   CU(cudaEventCreate(&ev0));
   CU(cudaEventCreate(&ev1));
   CU(cudaEventCreate(&ev2));
   CU(cudaMalloc(&d_fr, sizeof(FrontDev) * nn));
   CU(cudaMemset(d_fr, 0, sizeof(FrontDev) * nn));
   CU(cudaMallocHost(&pin_fr, sizeof(FrontDev) * nn));
   pin_i_cap = (size_t)6 * nn + 64;
   CU(cudaMallocHost(&pin_i, sizeof(int) * pin_i_cap));
   CU(cudaMalloc(&d_child_doff, sizeof(int) * std::max<size_t>(P->child_idx.size(), 1)));
   CU(cudaMalloc(&d_list, sizeof(int) * 2 * nn));
   CU(cudaMalloc(&d_collect, sizeof(int) * 3 * nn));
   CU(cudaMalloc(&d_counters, sizeof(int) * 8));
   CU(cudaMalloc(&d_tileptr, sizeof(int64_t) * (nn + 1)));
   CU(cudaMalloc(&d_updmap, sizeof(int) * UPDMAP_CAP));
   CU(cudaMallocHost(&pin_updmap, sizeof(int) * UPDMAP_CAP));
   CU(cudaMalloc(&d_xw_off, sizeof(int64_t) * (nn + 1)));
   CU(cudaMalloc(&d_part_off, sizeof(int64_t) * (nn + 1)));
   CU(cudaMalloc(&d_ticket, sizeof(int) * (nn + 1)));
   CU(cudaMalloc(&d_chain_counter, sizeof(unsigned long long) * 2 * (P->nlevels + 1)));
   CU(cudaMalloc(&d_chain_err, sizeof(int)));
   if (getenv("B200LDLT_SOLVE_TRACE")) CU(cudaMalloc(&d_chain_dbg, sizeof(unsigned long long) * 16 * 4096));
   CU(cudaMemset(d_chain_err, 0, sizeof(int)));
   CU(cudaMemset(d_ticket, 0, sizeof(int) * (nn + 1)));
   CU(cudaMalloc(&d_slist, sizeof(int) * nn));
   CU(cudaMalloc(&d_flist, sizeof(int) * nn));
   CU(cudaMallocHost(&pin_flist, sizeof(int) * nn));
   CU(cudaMalloc(&d_val, sizeof(double) * std::max<int64_t>(P->sym->nz, 1)));
   CU(cudaDeviceSynchronize());   // the null-stream memsets above are not ordered with the non-blocking streams
   child_doff.assign(P->child_idx.size(), 0);
   hfr.assign(nn, FrontDev());
   fac.err = dar.err = par.err = car[0].err = car[1].err = war.err = bar.err = &cuda_err;
   active.assign(nn, 1);
   live.assign(nn, 1);
   boundary.assign(nn, 0);
   part_kind.assign(nn, 0);
   owner.assign(nn, 0);
   lvl_fusable.assign(P->nlevels + 1, 0);
   for (int l = 0; l < P->nlevels; ++l) {
      bool all = true;
      for (int s = P->lvl_ptr[l]; s < P->lvl_ptr[l + 1] && all; ++s) {
         const SlotStatic& ss = P->hss[s];
         const int nchs = ss.child_end - ss.child_begin;
         all = ss.ncol <= ffront::MAXN && nchs <= ffront::MAXCH && ffront::smem_bytes(ss.nrow, ss.ncol, nchs) <= ffront::SMEM_CAP;
      }
      lvl_fusable[l] = all ? 1 : 0;
   }
   init_failed = false;
   inited = true;
   return 0;
}

int Numeric::factor(const double* val, bool on_device, const FactorOptions& o, FactorStats& stats, int phase) {
   stats = FactorStats();
   int rc = init();
   if (rc) { stats.cuda_error = cuda_err; return stats.flag = rc; }
   CU(cudaSetDevice(P->device));
   const Symbolic& sym = *P->sym;
   const int nn = P->nnodes;
   have = false;
   posdef = o.posdef != 0;
   launches = 0;
   prof_used = 0;
   prof_flops = 0;
   Params prm;
   prm.u = std::min(std::max(o.u, 0.0), 0.5);
   prm.small_ = o.small_;
   prm.lim = prm.u > 0 ? 1.0 / prm.u : INFINITY;
   prm.action = o.action;
   prm.posdef = posdef;
   { const char* tr = getenv("B200LDLT_TRACE"); prm.debug = tr && atoi(tr) >= 2; }
   double mult = std::max(o.multiplier, 1.0);
   // first reservations of the arenas (they grow on demand afterwards): the whole tree on one GPU, this rank's fronts
   // only under a partition -- C5 on 8 GPUs ran rank 0 out of memory when every rank reserved the whole tree's 81 GB
   int64_t estF = P->est_factor_doubles, estC[2] = {P->est_c_doubles[0], P->est_c_doubles[1]}, estW = P->est_ws_doubles;
   if (phase >= 1 && nranks > 1) {
      estF = 0; estC[0] = estC[1] = 0; estW = 0;
      for (int l = 0; l < P->nlevels; ++l) {
         int64_t c = 0, ws = 0;
         for (int s = P->lvl_ptr[l]; s < P->lvl_ptr[l + 1]; ++s) {
            if (owner[s] != my_rank) continue;
            const SlotStatic& ss = P->hss[s];
            const int64_t ldl = align_up(ss.nrow, 2), cm = ss.nrow - ss.ncol;
            estF += ldl * ss.ncol + 32;
            if (!boundary[s]) c += align_up(cm, 2) * cm + 32;
            ws += (int64_t)front_ws_doubles(align_up(ss.nrow, 2), ss.ncol);
         }
         estC[l & 1] = std::max(estC[l & 1], c);
         estW = std::max(estW, ws);
      }
   }
   fac.reserve((size_t)(estF * mult) * 8 + (1 << 20));
   dar.reserve((size_t)(2.0 * sym.n * mult) * 8 + 64 * (size_t)nn + (1 << 16));
   par.reserve((size_t)(2.0 * sym.n * mult) * 4 + 512 * (size_t)nn + (1 << 16));
   car[0].reserve((size_t)(estC[0] * mult) * 8 + (1 << 16));
   car[1].reserve((size_t)(estC[1] * mult) * 8 + (1 << 16));
   war.reserve((size_t)(estW * mult) * 8 + (1 << 16));
   if (phase <= 1) CU(cudaMemsetAsync(d_fr, 0, sizeof(FrontDev) * std::max(nn, 1), st));
   // phase 0: everything (single GPU).  phase k >= 1: this rank's fronts of stage k-1 of the proportional
   // mapping, on top of what the earlier phases and the imports left in place.
   if (phase <= 1) { fac.reset(); dar.reset(); par.reset(); bar.reset(); }
   for (int s = 0; s < nn; ++s) {
      active[s] = phase == 0 ? 1 : ((owner[s] == my_rank && part_kind[s] == phase - 1) ? 1 : 0);
      // phase 0 overwrites every record at its level (children before parents); the staged phases need the fronts of
      // other ranks and later stages empty
      if (phase == 1) { hfr[s] = FrontDev(); }
   }

   CU(cudaEventRecord(ev0, st));
   if (phase >= 2) {
      // values are already resident from phase 1
   } else if (on_device) {
      if (val != d_val) CU(cudaMemcpyAsync(d_val, val, sizeof(double) * sym.nz, cudaMemcpyDeviceToDevice, st));
   } else {
      CU(cudaMemcpyAsync(d_val, val, sizeof(double) * sym.nz, cudaMemcpyHostToDevice, st));
   }
   if (phase <= 1 && scaled && sym.nz > 0) {
      k_scale_values<<<(unsigned)((sym.nz + 255) / 256), 256, 0, st>>>(d_val, d_a_rowi, d_a_coli, d_scale, sym.nz);
      ++launches;
   }
   CU(cudaEventRecord(ev1, st));

   int flag = 0, num_delay = 0, maxfront = 0, nfp = 0, nsp = 0;
   // B200LDLT_TRACE=1: print the progress counters of every front after each block step (debugging aid;
   // synchronises the stream)
   const bool trace = getenv("B200LDLT_TRACE") != nullptr;
   const int upd_share = getenv("B200LDLT_UPD_SHARE") ? atoi(getenv("B200LDLT_UPD_SHARE")) : 296;
   const bool lookahead = !(getenv("B200LDLT_LOOKAHEAD") && atoi(getenv("B200LDLT_LOOKAHEAD")) == 0);
   // large fronts: trailing updates through k_upd3, deferred over defer_g blocks for fronts of defer_min_m rows and more
   const bool use_upd3 = !(getenv("B200LDLT_UPD3") && atoi(getenv("B200LDLT_UPD3")) == 0);
   const bool use_pdl = !(getenv("B200LDLT_PDL") && atoi(getenv("B200LDLT_PDL")) == 0);
   // small fronts (at most 64 candidate columns, panel in shared memory) in one launch per level (front_kernels.cuh)
   const bool use_front = !(getenv("B200LDLT_FRONT") && atoi(getenv("B200LDLT_FRONT")) == 0);
   // the launches on a front's critical path (strips, the next group's columns of a deferred update) in 64 x 64 tiles
   const bool small_strips = !(getenv("B200LDLT_SMALL_STRIPS") && atoi(getenv("B200LDLT_SMALL_STRIPS")) == 0);
   const int big_groups = getenv("B200LDLT_BIG_GROUPS") ? atoi(getenv("B200LDLT_BIG_GROUPS")) : 1;
   const int big_groups_minf = getenv("B200LDLT_BIG_GROUPS_MINF") ? atoi(getenv("B200LDLT_BIG_GROUPS_MINF")) : 4;
   const int small_strip_tiles = getenv("B200LDLT_SMALL_STRIP_TILES") ? atoi(getenv("B200LDLT_SMALL_STRIP_TILES")) : 148;
   const int front_small_m = getenv("B200LDLT_FRONT_SMALLM") ? atoi(getenv("B200LDLT_FRONT_SMALLM")) : 96;
   const int upd3_min_m = getenv("B200LDLT_UPD3_MINM") ? atoi(getenv("B200LDLT_UPD3_MINM")) : 1800;
   const int defer_g = getenv("B200LDLT_DEFER") ? std::max(1, atoi(getenv("B200LDLT_DEFER"))) : 4;
   const int defer_min_m = getenv("B200LDLT_DEFER_MINM") ? atoi(getenv("B200LDLT_DEFER_MINM")) : 2048;
   const int max_groups = getenv("B200LDLT_GROUPS") ? std::max(1, std::min((int)NAUX, atoi(getenv("B200LDLT_GROUPS")))) : (int)NAUX;
   auto trace_fronts = [&](const char* what, int slot_first, int count) {
      std::vector<FrontDev> tmp(count);
      cudaStreamSynchronize(st);
      cudaMemcpy(tmp.data(), d_fr + slot_first, sizeof(FrontDev) * count, cudaMemcpyDeviceToHost);
      for (int q = 0; q < count; ++q)
         fprintf(stderr, "[b200ldlt] %s slot %d m=%d n=%d done=%d gend=%d w=%d fail=%d flag=%d\n", what, slot_first + q,
                 tmp[q].m, tmp[q].n, tmp[q].done, tmp[q].gend, tmp[q].w, tmp[q].fail, tmp[q].flag);
   };
   std::vector<int> lvl_list, tpp_list;
   std::vector<int64_t> tileptr;
   size_t updmap_used = 0;
   prez_c = prez_l = nullptr;
   const bool lvl_times = getenv("B200LDLT_LEVELTIMES") != nullptr;   // wall time per level (each level ends with a sync)
   const bool timeline = getenv("B200LDLT_TIMELINE") != nullptr;
   std::vector<cudaEvent_t> tl_ev;
   std::vector<std::string> tl_name;
   for (int l = 0; l < P->nlevels && flag >= 0; ++l) {
      const int s0 = P->lvl_ptr[l], s1 = P->lvl_ptr[l + 1], ns = s1 - s0;
      const auto lvl_t0 = std::chrono::steady_clock::now();
      const int64_t lvl_l0 = launches;
      Arena& ca = car[l & 1];
      ca.reset();
      war.reset();
      updmap_used = 0;   // the previous level ended with a stream synchronisation
      // sizes with delays
      size_t totL = 0, totC = 0, totD = 0, totP = 0, totW = 0;
      std::vector<int> flist[4];
      size_t fsmem[4] = {0, 0, 0, 0};
      int ngeneral = 0;
      for (int s = s0; s < s1; ++s) live[s] = 0;
      int nactive_lvl = 0;
      for (int s = s0; s < s1; ++s) nactive_lvl += active[s];
      if (nactive_lvl == 0) continue;
      for (int s = s0; s < s1; ++s) {
         if (!active[s]) continue;
         const SlotStatic& ss = P->hss[s];
         FrontDev& f = hfr[s];
         int nd = 0;
         for (int ci = ss.child_begin; ci < ss.child_end; ++ci) {
            const FrontDev& fc = hfr[P->child_idx[ci]];
            child_doff[ci] = nd;
            nd += fc.n - fc.nelim;
         }
         f = FrontDev();
         f.slot = s;
         f.ncol = ss.ncol; f.nrow = ss.nrow; f.first_piv = ss.first_piv;
         f.ndelay_in = nd;
         f.m = ss.nrow + nd; f.n = ss.ncol + nd;
         f.ldl = (int)align_up(f.m, 2);
         f.ldc = (int)align_up(f.m - f.n, 2);
         f.done = 0; f.gend = f.n; f.w = 0; f.fail = 0; f.nelim = 0; f.flag = 0; f.wc = block_wc(f.n);
         f.state = 2; f.active = 1; f.part_kind = part_kind[s];
         {
            const int nchs = ss.child_end - ss.child_begin;
            if (use_front && f.n <= ffront::MAXN && nchs <= ffront::MAXCH && ffront::smem_bytes(f.m, f.n, nchs) <= ffront::SMEM_CAP) {
               f.fused = 1;
               // buckets: one or two panels x four warps (fronts of at most front_small_m rows) or eight
               const int np = (f.n <= NB ? 0 : 1) + (f.m <= front_small_m ? 2 : 0);
               flist[np].push_back(s);
               fsmem[np] = std::max(fsmem[np], ffront::smem_bytes(f.m, f.n, nchs));
            }
            live[s] = f.fused ? 0 : 1;
            if (!f.fused) ++ngeneral;
         }
         totL += align_up((size_t)f.ldl * f.n, 32);
         if (!boundary[s]) totC += align_up((size_t)f.ldc * (f.m - f.n) + 1, 32);
         totD += align_up(2 * (size_t)f.n + 2, 32);
         totP += 2 * align_up((size_t)f.n + 1, 32);
         totW += front_ws_doubles(f.ldl, f.n);
         maxfront = std::max(maxfront, f.m);
      }
      double* Lb = (double*)fac.alloc(totL * 8);
      double* Cb = (double*)ca.alloc(totC * 8);
      double* Db = (double*)dar.alloc(totD * 8);
      int* Pb = (int*)par.alloc(totP * 4);
      double* Wb = (double*)war.alloc(totW * 8);
      if (!Lb || !Cb || !Db || !Pb || !Wb) { stats.cuda_error = cuda_err; return stats.flag = SSIDS_ERROR_ALLOCATION; }
      {
         size_t oL = 0, oC = 0, oD = 0, oP = 0, oW = 0;
         for (int s = s0; s < s1; ++s) {
            if (!active[s]) continue;
            FrontDev& f = hfr[s];
            f.L = Lb + oL; oL += align_up((size_t)f.ldl * f.n, 32);
            if (boundary[s]) {
               // the contribution block crosses a part boundary: it must outlive the level arenas
               size_t cb = align_up((size_t)f.ldc * (f.m - f.n) + 1, 32) * 8;
               f.C = (double*)bar.alloc(cb);
               if (!f.C) { stats.cuda_error = cuda_err; return stats.flag = SSIDS_ERROR_ALLOCATION; }
               CU(cudaMemsetAsync(f.C, 0, cb, st));
               f.C += f.n & 1;
            } else {
               // C(r, c) sits at f.C + (c - n) ldc + (r - n): shifted by one double when n is odd, so that even rows r
               // are 16-byte aligned in C exactly as they are in L (k_update's paired accesses)
               f.C = Cb + oC + (f.n & 1); oC += align_up((size_t)f.ldc * (f.m - f.n) + 1, 32);
            }
            f.D = Db + oD; oD += align_up(2 * (size_t)f.n + 2, 32);
            f.perm = Pb + oP; f.orig = f.perm + align_up((size_t)f.n + 1, 32); oP += 2 * align_up((size_t)f.n + 1, 32);
            double* wp = Wb + oW;
            oW += front_ws_doubles(f.ldl, f.n);
            f.xw = ring_cols(f.n);
            f.Wld = wp; f.Wo = wp + (size_t)f.ldl * f.xw;
            f.dbuf = f.Wo + (size_t)f.ldl * f.wc;
            f.ibuf = (int*)(f.dbuf + 2 * (size_t)PB * f.wc + 2 * PB + 2 + (size_t)NB * f.wc);
         }
      }
      const auto lvl_t1 = std::chrono::steady_clock::now();
      memcpy(pin_fr + s0, hfr.data() + s0, sizeof(FrontDev) * ns);
      CU(cudaMemcpyAsync(d_fr + s0, pin_fr + s0, sizeof(FrontDev) * ns, cudaMemcpyHostToDevice, st));
      {
         // what the previous level's side stream has already zeroed (sizes without delays) is skipped
         size_t zl = 0, zc = 0;
         if (prez_l || prez_c) CU(cudaStreamWaitEvent(st, ev_zero, 0));
         if (prez_l && prez_l == (char*)Lb) zl = std::min(prez_l_bytes, totL * 8);
         if (prez_c && prez_c == (char*)Cb) zc = std::min(prez_c_bytes, totC * 8);
         prez_l = prez_c = nullptr;
         // k_front writes every entry of its fronts' L and of the lower triangle of their contribution blocks: a level
         // it takes whole needs no zero-fill
         if (ngeneral > 0 && totL * 8 > zl) CU(cudaMemsetAsync((char*)Lb + zl, 0, totL * 8 - zl, st));
         if (ngeneral > 0 && totC * 8 > zc) CU(cudaMemsetAsync((char*)Cb + zc, 0, totC * 8 - zc, st));
      }
      CU(cudaMemsetAsync(Db, 0, totD * 8, st));
      const int cb = P->hss[s0].child_begin, ce = P->hss[s1 - 1].child_end;
      if (ce > cb)
         CU(cudaMemcpyAsync(d_child_doff + cb, child_doff.data() + cb, sizeof(int) * (ce - cb), cudaMemcpyHostToDevice, st));
      {
         int nf = 0;
         for (int b = 0; b < 4; ++b) {
            if (!flist[b].empty()) memcpy(pin_flist + s0 + nf, flist[b].data(), sizeof(int) * flist[b].size());
            nf += (int)flist[b].size();
         }
         if (nf > 0) CU(cudaMemcpyAsync(d_flist + s0, pin_flist + s0, sizeof(int) * nf, cudaMemcpyHostToDevice, st));
         int off = 0;
         for (int b = 0; b < 4; ++b) {
            const int nb = (int)flist[b].size();
            if (nb == 0) continue;
            const int* lst = d_flist + s0 + off;
#define B200_FRONT_ARGS d_fr, P->d_ss, lst, P->d_child_idx, d_child_doff, P->d_rel, P->d_aptr, P->d_asrc, P->d_acol, P->d_arow, d_val, prm
            if (b == 0) k_front<1, 8><<<nb, 256, fsmem[b], st>>>(B200_FRONT_ARGS);
            else if (b == 1) k_front<2, 8><<<nb, 256, fsmem[b], st>>>(B200_FRONT_ARGS);
            else if (b == 2) k_front<1, 4><<<nb, 128, fsmem[b], st>>>(B200_FRONT_ARGS);
            else k_front<2, 4><<<nb, 128, fsmem[b], st>>>(B200_FRONT_ARGS);
#undef B200_FRONT_ARGS
            ++launches;
            off += nb;
         }
      }
      if (ngeneral > 0) {
         k_setup<<<ns, 128, 0, st>>>(d_fr, P->d_ss, P->d_child_idx, d_child_doff, s0, ns);
         ++launches;
      }
      if (ngeneral > 0) {
         int64_t e0 = P->lvl_a_ptr[l], e1 = P->lvl_a_ptr[l + 1];
         if (e1 > e0) {
            k_scatter_a<<<(unsigned)((e1 - e0 + 255) / 256), 256, 0, st>>>(d_fr, P->d_aslot, P->d_asrc, P->d_acol, P->d_arow, d_val, e0, e1);
            ++launches;
         }
      }
      // extend-add, one launch per child ordinal.  The staging lists are consumed by async copies,
      // so each ordinal gets its own region of the pinned buffer.
      {
         int maxc = 0;
         for (int s = s0; s < s1; ++s) if (live[s]) maxc = std::max(maxc, P->hss[s].child_end - P->hss[s].child_begin);
         size_t pin_off = 0;
         for (int k = 0; k < maxc; ++k) {
            lvl_list.clear();
            std::vector<int> doffs;
            tileptr.clear();
            tileptr.push_back(0);
            for (int s = s0; s < s1; ++s) {
               if (!live[s]) continue;
               const SlotStatic& ss = P->hss[s];
               if (ss.child_end - ss.child_begin <= k) continue;
               int cs = P->child_idx[ss.child_begin + k];
               const FrontDev& fc = hfr[cs];
               int t = fc.m - fc.nelim;
               if (t <= 0) continue;
               int64_t nt = (int64_t)((t + EA_TR - 1) / EA_TR) * ((t + EA_TC - 1) / EA_TC);
               lvl_list.push_back(cs);
               doffs.push_back(child_doff[ss.child_begin + k]);
               tileptr.push_back(tileptr.back() + nt);
            }
            int nl = (int)lvl_list.size();
            if (nl == 0) continue;
            // staging: [list nl][doffs nl] ints then tileptr (nl+1) int64
            size_t need = 2 * (size_t)nl + 2 * (size_t)(nl + 1) + 4;
            if (pin_off + need > pin_i_cap) {   // drain and reuse
               CU(cudaStreamSynchronize(st));
               pin_off = 0;
            }
            int* pl = pin_i + pin_off;
            memcpy(pl, lvl_list.data(), sizeof(int) * nl);
            memcpy(pl + nl, doffs.data(), sizeof(int) * nl);
            size_t toff = align_up(pin_off + 2 * (size_t)nl, 2);
            int64_t* pt = (int64_t*)(pin_i + toff);
            memcpy(pt, tileptr.data(), sizeof(int64_t) * (nl + 1));
            pin_off = toff + 2 * (size_t)(nl + 1);
            CU(cudaMemcpyAsync(d_list, pl, sizeof(int) * 2 * nl, cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(d_tileptr, pt, sizeof(int64_t) * (nl + 1), cudaMemcpyHostToDevice, st));
            int64_t total = tileptr.back();
            if (total > 0x7fffffffLL) { return stats.flag = SSIDS_ERROR_UNKNOWN; }
            k_extend_add<<<(unsigned)total, EA_TR, 0, st>>>(d_fr, P->d_ss, P->d_rel, d_list, d_list + nl, d_tileptr, nl);
            ++launches;
         }
         // the staging regions stay untouched until the level's closing synchronisation (collect)
      }
      // The children's contribution blocks (the other arena) are dead once extend-add has run: zero the next
      // level's C region there, and its L region behind this level's, beside the factorisation kernels.
      if (phase == 0 && l + 1 < P->nlevels && !trace && !(use_front && lvl_fusable[l + 1])) {
         Arena& na2 = car[(l + 1) & 1];
         const size_t cbytes = (size_t)P->lvl_C_doubles[l + 1] * 8, lbytes = (size_t)P->lvl_L_doubles[l + 1] * 8;
         CU(cudaEventRecord(ev_ea, st));
         CU(cudaStreamWaitEvent(zs, ev_ea, 0));
         if (cbytes > 0 && !na2.chunks.empty() && cbytes <= na2.chunks[0].size) {
            prez_c = na2.chunks[0].p; prez_c_bytes = cbytes;
            CU(cudaMemsetAsync(prez_c, 0, cbytes, zs));
         }
         if (lbytes > 0 && fac.cur < fac.chunks.size() && fac.off + align_up(lbytes, 256) <= fac.chunks[fac.cur].size) {
            prez_l = fac.chunks[fac.cur].p + fac.off; prez_l_bytes = lbytes;
            CU(cudaMemsetAsync(prez_l, 0, lbytes, zs));
         }
         CU(cudaEventRecord(ev_zero, zs));
      }
      // factor: fronts bucketed by size class, sorted by candidate count (descending) inside.
      // Each bucket runs rounds of blocks (PB columns, four launches each); the
      // device keeps the progress counters, the host only needs them once per round.
      {
         std::vector<int> order(ns);
         for (int i = 0; i < ns; ++i) order[i] = s0 + i;
         auto cls = [&](int s) { return size_class(P->hss[s].nrow); };   // static class: slots are pre-sorted by it
         auto prof_begin = [&]() {
            if (prof_used + 2 > prof_ev.size()) {
               size_t old = prof_ev.size();
               prof_ev.resize(old + 256);
               for (size_t q = old; q < prof_ev.size(); ++q) cudaEventCreate(&prof_ev[q]);
            }
            cudaEventRecord(prof_ev[prof_used], st);
         };
         auto prof_end = [&]() { cudaEventRecord(prof_ev[prof_used + 1], st); prof_used += 2; };   // profile mode is single-stream
         // B200LDLT_TIMELINE=1: an event after every launch of the large-front chunks, printed per level (diagnostic)
         auto mark = [&](cudaStream_t sm, const char* what, int a, int b) {
            if (!timeline) return;
            cudaEvent_t e;
            cudaEventCreate(&e);
            cudaEventRecord(e, sm);
            tl_ev.push_back(e);
            char buf[64];
            snprintf(buf, sizeof buf, "%s %s %d %d", sm == st ? "main" : "side", what, a, b);
            tl_name.push_back(buf);
         };
         // one chunk = a contiguous slot range of one size class, processed block by block on stream sx
         int big_split = 1;      // chunks a class of large fronts is split into (they share the SMs)
         auto run_chunk = [&](int j0, int j1, cudaStream_t sx, bool single, int sidx) -> int {
            int maxm = 0, maxn = 0;
            for (int q = j0; q < j1; ++q)
               if (live[order[q]]) { maxm = std::max(maxm, hfr[order[q]].m); maxn = std::max(maxn, hfr[order[q]].n); }
            if (maxm == 0) return 0;
            const int* lst = nullptr;
            const int b0 = s0 + j0;   // first slot of the bucket
            const int ncolchunks = (maxn + 63) / 64;
            const int ycommit = 1 + ncolchunks + (maxm + 255) / 256;
            const int nouter = (maxn + PB - 1) / PB;
            // Large fronts (size class >= 3) take their trailing updates through k_upd3; with several blocks per front
            // the schedule looks ahead: the part of an update the next block needs runs on this stream, the rest in slices
            // on the chunk's second stream beside the next blocks' factorisation kernels (k_block: one CTA per front).
            const bool big = (single || big_split > 1) && cls(order[j0]) >= 3 && use_upd3 && maxm >= upd3_min_m;
            const bool la = big && lookahead && !o.profile && !trace && nouter >= 2;
            // blocks per trailing update of the columns behind the group (rank G * PB): C is read and written once per G blocks
            const int G = (big && nouter >= 2 && maxm >= defer_min_m) ? std::max(1, std::min(defer_g, (int)GMAX)) : 1;
            const int ngroups = (nouter + G - 1) / G;
            auto nblocks_of = [&](const FrontDev& ff) { return (ff.n + PB - 1) / PB; };
            // work maps: CTAs in proportion to the fronts' tile counts (shapes without failed pivots; every front of
            // a launch keeps at least one CTA, so a front that lags behind is still updated correctly, only slower)
            std::vector<int> map_off(nouter + 1, 0), map_na(nouter, 0);            // k_update / strips, per block
            std::vector<int> bmap_off(ngroups + 1, 0);                              // k_upd3 bulk, per group
            std::vector<int> gmap_off(ngroups + 1, 0);                              // the next group's columns of a group's update (GSTRIP)
            // tile size of the launches on the critical path, per block / group: 64 when the launch would be less than one
            // wave of 128 x 128 tiles (then a tile's latency, not the throughput, is what the next block waits for)
            std::vector<int> strip_tile(nouter, 128), gstrip_on(ngroups, 0);
            std::vector<int64_t> gtiles(ngroups, 0);
            {
               std::vector<int> hmap;
               std::vector<int64_t> tl;
               auto deal = [&](int na, int64_t total, int64_t target) {
                  for (int q = 0; q < na; ++q) {
                     int parts = total > 0 ? (int)((tl[q] * target + total / 2) / total) : 1;
                     parts = std::max(1, (int)std::min<int64_t>(parts, std::max<int64_t>(tl[q], 1)));
                     // a front without work in the static picture may have some after a failed pivot (a shortened
                     // block leaves columns of its group behind): a few CTAs that normally exit at once
                     if (big && tl[q] == 0) parts = 4;
                     for (int pp = 0; pp < parts; ++pp) { hmap.push_back(q); hmap.push_back(pp); hmap.push_back(parts); }
                  }
               };
               for (int ob = 0; ob < nouter; ++ob) {
                  int na = 0;
                  for (int q = j0; q < j1; ++q) if (live[order[q]] && nblocks_of(hfr[order[q]]) > ob) na = q - j0 + 1;
                  map_na[ob] = na;
                  tl.assign(na, 0);
                  int64_t total = 0;
                  const int g = ob / G;
                  int stile = 128;
                  if (big && G > 1 && small_strips) {
                     int64_t t128 = 0;
                     for (int q = 0; q < na; ++q) {
                        const FrontDev& ff = hfr[order[j0 + q]];
                        if (!live[order[j0 + q]] || ff.n <= ob * PB) continue;
                        const int kend = std::min(ff.n, (ob + 1) * PB);
                        const int c1 = std::min(ff.m, std::max(kend, (g + 1) * G * PB));
                        const int org = kend & ~1;
                        if (c1 > kend) t128 += upd3::tiles_upto((c1 - org + 127) / 128, (ff.m - org + 127) / 128);
                     }
                     if (t128 < small_strip_tiles) stile = upd3::STM;
                  }
                  strip_tile[ob] = stile;
                  for (int q = 0; q < na; ++q) {
                     const FrontDev& ff = hfr[order[j0 + q]];
                     if (!live[order[j0 + q]] || ff.n <= ob * PB) continue;
                     const int kend = std::min(ff.n, (ob + 1) * PB);
                     int64_t cnt = 0;
                     if (!big) {
                        const int t = ff.m - kend;
                        if (t <= 0) continue;
                        const int ntr = (t + TM - 1) / TM, ntc = (t + TN - 1) / TN;
                        for (int ti = 0; ti < ntr; ++ti) cnt += std::min(ntc, 2 * ti + 2);
                     } else if (G > 1) {
                        // the block's own update: the columns behind it up to the end of the group
                        const int c1 = std::min(ff.m, std::max(kend, (g + 1) * G * PB));
                        const int org = kend & ~1;
                        if (c1 > kend) cnt = upd3::tiles_upto((c1 - org + stile - 1) / stile, (ff.m - org + stile - 1) / stile);
                     }
                     tl[q] = cnt;
                     total += cnt;
                  }
                  // chunks that share the GPU with other streams keep to about half of the CTA slots, so that
                  // another chunk's block factorisation finds a free SM while this update runs
                  const int64_t target = big ? (stile == 128 ? 148 : 296) / big_split : std::min<int64_t>(std::max<int64_t>(total, 1), (single ? 2 : 1) * upd_share);
                  if (!big || G > 1) deal(na, total, std::min<int64_t>(target, std::max<int64_t>(total, 1)));
                  map_off[ob + 1] = (int)hmap.size();
               }
               if (big) {
                  for (int g = 0; g < ngroups; ++g) {
                     const int na = map_na[g * G];
                     tl.assign(na, 0);
                     int64_t total = 0;
                     for (int q = 0; q < na; ++q) {
                        const FrontDev& ff = hfr[order[j0 + q]];
                        if (!live[order[j0 + q]] || ff.n <= g * G * PB) continue;
                        const int kend = std::min(ff.n, (g + 1) * G * PB);
                        const int c1 = G == 1 ? kend : std::min(ff.m, std::max(kend, (g + 1) * G * PB));
                        if (c1 >= ff.m) continue;
                        const int ntr = (ff.m - (c1 & ~1) + upd3::UTM - 1) / upd3::UTM;
                        tl[q] = upd3::tiles_upto(ntr, ntr);
                        total += tl[q];
                     }
                     gtiles[g] = total;
                     // the SMs of the next blocks' k_block CTAs (one per front) stay free
                     const int nkb = g + 1 < ngroups ? map_na[(g + 1) * G] : 0;
                     const int64_t target = std::max(64, 148 - std::min(nkb, 84)) / big_split;
                     deal(na, total, std::min<int64_t>(target, std::max<int64_t>(total, 1)));
                     bmap_off[g + 1] = (int)hmap.size();
                  }
                  for (int g = 0; g <= ngroups; ++g) if (bmap_off[g] < map_off[nouter]) bmap_off[g] = map_off[nouter];
                  gmap_off[0] = (int)hmap.size();
                  for (int g = 0; g < ngroups; ++g) {
                     if (small_strips && g + 1 < ngroups) {
                        const int na = map_na[g * G];
                        const int stile = upd3::STM;
                        tl.assign(na, 0);
                        int64_t total = 0, t128 = 0;
                        for (int q = 0; q < na; ++q) {
                           const FrontDev& ff = hfr[order[j0 + q]];
                           if (!live[order[j0 + q]] || ff.n <= g * G * PB) continue;
                           const int kend = std::min(ff.n, (g + 1) * G * PB);
                           const int c1 = G == 1 ? kend : std::min(ff.m, std::max(kend, (g + 1) * G * PB));
                           if (c1 >= ff.m) continue;
                           const int gse = G == 1 ? std::min(ff.m, kend + PB) : std::min(ff.m, c1 + G * PB);
                           const int org = c1 & ~1;
                           const int cb = std::min(ff.m, org + (gse - org + 127) / 128 * 128);
                           tl[q] = upd3::tiles_upto((cb - org + stile - 1) / stile, (ff.m - org + stile - 1) / stile);
                           total += tl[q];
                           t128 += upd3::tiles_upto((cb - org + 127) / 128, (ff.m - org + 127) / 128);
                        }
                        if (t128 < small_strip_tiles) {
                           gstrip_on[g] = 1;
                           deal(na, total, std::min<int64_t>(296, std::max<int64_t>(total, 1)));
                        }
                     }
                     gmap_off[g + 1] = (int)hmap.size();
                  }
               }
               if (updmap_used + hmap.size() > UPDMAP_CAP) {   // staging in use by earlier copies: drain first
                  CU(cudaStreamSynchronize(st));
                  for (int g2 = 0; g2 < NAUX; ++g2) CU(cudaStreamSynchronize(aux[g2]));
                  updmap_used = 0;
               }
               if (hmap.size() > UPDMAP_CAP) return SSIDS_ERROR_ALLOCATION;
               memcpy(pin_updmap + updmap_used, hmap.data(), sizeof(int) * hmap.size());
               if (!hmap.empty())
                  CU(cudaMemcpyAsync(d_updmap + updmap_used, pin_updmap + updmap_used, sizeof(int) * hmap.size(), cudaMemcpyHostToDevice, sx));
               for (int ob = 0; ob <= nouter; ++ob) map_off[ob] += (int)updmap_used;
               for (int g = 0; g <= ngroups; ++g) bmap_off[g] += (int)updmap_used;
               for (int g = 0; g <= ngroups; ++g) gmap_off[g] += (int)updmap_used;
               updmap_used += hmap.size();
            }
            // profile mode (bench roofline): every trailing-update launch between two events, its algorithmic flops
            auto prof_region = [&](int ob0, int ob1, bool strip_only, bool bulk_only) {
               // flops of the launch that applies the pivots of blocks [ob0, ob1] of every front to its region
               const int g = ob0 / G;
               for (int q = j0; q < j1; ++q) {
                  const FrontDev& ff = hfr[order[q]];
                  if (!live[order[q]] || ff.n <= ob0 * PB) continue;
                  const double ka = ob0 * PB, kb = std::min<double>(ff.n, (ob1 + 1) * PB);
                  const double c1 = G == 1 ? kb : std::min<double>(ff.m, std::max<double>(kb, (g + 1) * G * PB));
                  double ca, cb;
                  if (strip_only) { ca = kb; cb = c1; }
                  else if (bulk_only) { ca = c1; cb = ff.m; }
                  else { ca = kb; cb = ff.m; }
                  if (cb > ca && kb > ka) prof_flops += (kb - ka) * (cb - ca) * (2.0 * ff.m - ca - cb + 1.0);
               }
            };
            int slices_left = 0, slice_ns = 1, slice_tag = 0, slice_z = 0, slice_g = 0;   // window slices of the previous group's update
            for (int ob = 0; ob < nouter; ++ob) {
               const int na = map_na[ob];   // slots [b0, b0+na) cover every front that still has a block number ob
               if (na == 0) break;
               const int par = ob & 1;
               const int g = ob / G, gpos = ob % G;
               const bool glast = gpos == G - 1 || ob == nouter - 1;
               int maxw = 0;
               for (int q = 0; q < na; ++q) if (live[order[j0 + q]]) maxw = std::max(maxw, std::min(PB, hfr[order[j0 + q]].n - ob * PB));
               const int wcap = (maxw + NB - 1) / NB * NB;
               // blocks shortened by failed pivots leave `done` behind ob*PB, so the row grid covers the
               // whole front; CTAs above the block exit at once
               // k_apply: about two CTAs per SM over the fronts of the launch, each walking its share of the row tiles
               int yapply = std::max(1, (maxm + AR - 1) / AR);
               {
                  const int ntile_max = yapply;
                  yapply = std::max(1, std::min(ntile_max, 296 / std::max(na, 1)));
                  const int rounds = (ntile_max + yapply - 1) / yapply;
                  yapply = (ntile_max + rounds - 1) / rounds;
               }
               bool slice_running = false;
               if (slices_left > 0) {
                  // one slice of the previous group's update beside this block's factorisation
                  const int s = slice_ns - slices_left;      // 1 .. ns-1
                  const int nct = (bmap_off[slice_g + 1] - bmap_off[slice_g]) / 3;
                  if (nct > 0) {
                     CU(cudaEventRecord(ev_commit[sidx][0], sx));
                     CU(cudaStreamWaitEvent(bulk[sidx], ev_commit[sidx][0], 0));
                     kt_upd3<128><<<nct, upd3::THREADS, upd3::SMEM, bulk[sidx]>>>(d_fr, lst, b0, d_updmap + bmap_off[slice_g], 1, slice_g & 1, slice_tag,
                                                                            s - 1, s, slice_ns - 1, slice_z);
                     CU(cudaEventRecord(ev_bulk[sidx][0], bulk[sidx]));
                     mark(bulk[sidx], "slice", ob, s);
                     ++launches;
                     slice_running = true;
                  }
                  --slices_left;
               }
               // a front that sees a failed pivot stalls (instead of parking the failed columns at once) whenever other
               // work on it may be outstanding: a concurrent slice, or the deferred update of its group
               const int stall_on_fail = ((la || G > 1) ? 1 : 0) | 2;     // bit 1: fronts of k_front are passed by
               if (wcap == NB) k_block<0><<<na, 256, block_smem_bytes(wcap), sx>>>(d_fr, lst, b0, prm, 0, wcap, par, stall_on_fail);
               else if (wcap == 2 * NB) k_block<1><<<na, 256, block_smem_bytes(wcap), sx>>>(d_fr, lst, b0, prm, 0, wcap, par, stall_on_fail);
               else k_block<2><<<na, 256, block_smem_bytes(wcap), sx>>>(d_fr, lst, b0, prm, 0, wcap, par, stall_on_fail);
               if (big) mark(sx, "block", ob, na);
               if (slice_running) CU(cudaStreamWaitEvent(sx, ev_bulk[sidx][0], 0));
               k_apply<<<dim3(na, yapply), 256, apply_smem_bytes(wcap), sx>>>(d_fr, lst, b0, prm, wcap);
               if (big) mark(sx, "apply", ob, yapply);
               launch_dep(use_pdl && !timeline, k_commit, dim3(na, ycommit), dim3(256), 0, sx, d_fr, lst, b0, ncolchunks, par, gpos, G, g + 1);
               if (big) mark(sx, "commit", ob, 0);
               launches += 3;
               const int nct = (map_off[ob + 1] - map_off[ob]) / 3;
               if (!big) {
                  if (o.profile == 1) { prof_region(ob, ob, false, false); prof_begin(); }
                  k_update<<<nct, 256, UPD_SMEM, sx>>>(d_fr, lst, b0, d_updmap + map_off[ob], par);
                  if (o.profile == 1) prof_end();
                  ++launches;
               } else {
                  if (G > 1 && nct > 0) {
                     if (o.profile) { prof_region(ob, ob, true, false); prof_begin(); }
                     if (strip_tile[ob] == upd3::STM)
                        launch_dep(use_pdl && !timeline && !o.profile, kt_upd3<upd3::STM>, dim3(nct), dim3(upd3::THREADS), upd3::Shape<upd3::STM>::SMEM, sx,
                                   d_fr, lst, b0, (const int*)(d_updmap + map_off[ob]), 0, par, 0, 0, 0, 1, 0);
                     else
                        launch_dep(use_pdl && !timeline && !o.profile, kt_upd3<128>, dim3(nct), dim3(upd3::THREADS), upd3::SMEM, sx, d_fr, lst, b0,
                                   (const int*)(d_updmap + map_off[ob]), 0, par, 0, 0, 0, 1, 0);
                     mark(sx, "strip", ob, nct);
                     if (o.profile) prof_end();
                     ++launches;
                  }
                  if (glast) {
                     const int nctb = (bmap_off[g + 1] - bmap_off[g]) / 3;
                     // slices beside the next group's blocks, as many tiles as fit behind their k_block launches
                     const int nnext = (la && ob + 1 < nouter) ? std::min(G, nouter - (ob + 1)) : 0;
                     int z = 256;
                     if (nnext > 0 && nctb > 0 && gtiles[g] > 0) {
                        const double tile_us = 20.0 * G;                         // 128 x 128 x (G PB) at the kernel's rate
                        const int per_window = std::max(1, (int)(66.0 / tile_us));
                        const double cap = (double)nnext * nctb * per_window;
                        z = (int)std::ceil(256.0 * std::max(0.0, 1.0 - cap / (double)gtiles[g]));
                        z = std::max(0, std::min(256, z));
                     }
                     const int nctg = (gmap_off[g + 1] - gmap_off[g]) / 3;
                     if (nctb > 0) {
                        if (o.profile) { prof_region(g * G, ob, false, true); prof_begin(); }
                        if (nctg > 0) {
                           // the next group's columns first, in small tiles; the bulk launches start behind them (sa = -2)
                           launch_dep(use_pdl && !timeline && !o.profile, kt_upd3<upd3::STM>, dim3(nctg), dim3(upd3::THREADS), upd3::Shape<upd3::STM>::SMEM, sx,
                                      d_fr, lst, b0, (const int*)(d_updmap + gmap_off[g]), 3, g & 1, g + 1, 0, 0, 1, 0);
                           mark(sx, "gstrip", ob, nctg);
                           ++launches;
                        }
                        if (nctg == 0 || z > 0)
                        launch_dep(use_pdl && !timeline && !o.profile, kt_upd3<128>, dim3(nctb), dim3(upd3::THREADS), upd3::SMEM, sx, d_fr, lst, b0,
                                   (const int*)(d_updmap + bmap_off[g]), 1, g & 1, g + 1, nctg > 0 ? -2 : -1, 0, 1, z);
                        mark(sx, "slice0", ob, z);
                        if (o.profile) prof_end();
                        ++launches;
                     }
                     if (nnext > 0 && z < 256) { slices_left = nnext; slice_ns = nnext + 1; slice_tag = g + 1; slice_z = z; slice_g = g; }
                     else if (nnext > 0 && nctb > 0) {
                        // everything already ran in slice 0 (bound0 = T whenever z = 256): nothing left for the windows
                        slices_left = 0;
                     }
                  }
               }
               if (trace) trace_fronts("block", b0, na);
            }
            if (big && G > 1) {
               // what stalled fronts still owe (deferred pivots of an unfinished group)
               int nf = 0;
               std::vector<int> fm;
               int nfl = 0;
               for (int q = j0; q < j1; ++q) if (live[order[q]] && nblocks_of(hfr[order[q]]) >= 2) ++nfl;
               const int fparts = std::max(1, 148 / std::max(nfl, 1));
               for (int q = j0; q < j1; ++q)
                  if (live[order[q]] && nblocks_of(hfr[order[q]]) >= 2)
                     for (int pp = 0; pp < fparts; ++pp) { fm.push_back(q - j0); fm.push_back(pp); fm.push_back(fparts); ++nf; }
               if (nf > 0 && updmap_used + fm.size() <= UPDMAP_CAP) {
                  memcpy(pin_updmap + updmap_used, fm.data(), sizeof(int) * fm.size());
                  CU(cudaMemcpyAsync(d_updmap + updmap_used, pin_updmap + updmap_used, sizeof(int) * fm.size(), cudaMemcpyHostToDevice, sx));
                  kt_upd3<128><<<nf, upd3::THREADS, upd3::SMEM, sx>>>(d_fr, lst, b0, d_updmap + updmap_used, 2, 0, 0, 0, 0, 1, 0);
                  updmap_used += fm.size();
                  ++launches;
               } else if (nf > 0) return SSIDS_ERROR_ALLOCATION;
            }
            k_block<0><<<j1 - j0, 256, block_smem_bytes(NB), sx>>>(d_fr, lst, b0, prm, 1, NB, 0, 2);
            ++launches;
            return 0;
         };
         int i0 = 0;
         bool forked = false;
         while (i0 < ns) {
            int c = cls(order[i0]), i1 = i0;
            int nact = 0;
            double work = 0;
            while (i1 < ns && cls(order[i1]) == c) {
               if (live[order[i1]]) { ++nact; work += (double)hfr[order[i1]].n * hfr[order[i1]].m * hfr[order[i1]].m; }
               ++i1;
            }
            // a few large fronts: split them over streams so that one group's block factorisation
            // overlaps another group's trailing update
            int G = 1;
            // classes of many medium fronts: stream groups and the small-front kernel (one group's block factorisation
            // beside another's update); classes with fronts of upd3_min_m rows and more: one chunk, k_upd3
            int cls_maxm = 0;
            for (int q = i0; q < i1; ++q) if (live[order[q]]) cls_maxm = std::max(cls_maxm, hfr[order[q]].m);
            const bool cls_big = c >= 3 && use_upd3 && cls_maxm >= upd3_min_m;
            if (!o.profile && !trace && nact >= 2 && nact <= 128 && c >= 2 && !cls_big) G = std::min(nact, max_groups);
            // several large fronts: groups on separate streams as well, each with its own look-ahead schedule
            big_split = 1;
            if (!o.profile && !trace && !timeline && cls_big && big_groups > 1 && nact >= big_groups_minf) {
               G = std::min({nact, big_groups, max_groups});
               big_split = G;
            }
            if (G == 1) {
               int rc2 = run_chunk(i0, i1, st, true, 0);
               if (rc2) return stats.flag = rc2;
            } else {
               if (!forked) { CU(cudaEventRecord(ev_fork, st)); forked = true; }
               int j0 = i0;
               double acc = 0;
               for (int g2 = 0; g2 < G; ++g2) {
                  int j1 = j0;
                  const double goal = work * (g2 + 1) / G;
                  while (j1 < i1 && (acc < goal || g2 == G - 1)) {
                     if (live[order[j1]]) acc += (double)hfr[order[j1]].n * hfr[order[j1]].m * hfr[order[j1]].m;
                     ++j1;
                  }
                  if (g2 == G - 1) j1 = i1;
                  if (j1 > j0) {
                     CU(cudaStreamWaitEvent(aux[g2], ev_fork, 0));
                     int rc2 = run_chunk(j0, j1, aux[g2], false, g2 + 1);
                     if (rc2) return stats.flag = rc2;
                     CU(cudaEventRecord(ev_join[g2], aux[g2]));
                     CU(cudaStreamWaitEvent(st, ev_join[g2], 0));
                  }
                  j0 = j1;
               }
            }
            i0 = i1;
         }
      }
      // read back progress; fronts whose scheduled panels did not exhaust their candidates
      // (only possible after failures) get further rounds
      auto collect = [&]() -> int {
         k_collect<<<(ns + 255) / 256, 256, 0, st>>>(d_fr, s0, ns, d_collect);
         ++launches;
         if (cudaMemcpyAsync(pin_i, d_collect, sizeof(int) * 3 * ns, cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
         if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
         for (int i = 0; i < ns; ++i) {
            if (!active[s0 + i]) continue;
            FrontDev& f = hfr[s0 + i];
            f.done = pin_i[3 * i];
            f.nelim = f.done;
            f.flag = pin_i[3 * i + 1];
            f.gend = pin_i[3 * i + 2];
            if (f.flag < 0) flag = f.flag;
         }
         return 0;
      };
      const auto lvl_t2 = std::chrono::steady_clock::now();
      if (collect()) { cuda_err = (int)cudaGetLastError(); stats.cuda_error = cuda_err; return stats.flag = SSIDS_ERROR_CUDA_UNKNOWN; }
      const auto lvl_t3 = std::chrono::steady_clock::now();
      if (flag < 0) break;
      for (int round = 0; round < 1000; ++round) {
         tpp_list.clear();
         for (int i = 0; i < ns; ++i) if (active[s0 + i] && hfr[s0 + i].done < hfr[s0 + i].gend) tpp_list.push_back(s0 + i);
         if (tpp_list.empty()) break;
         int nt = (int)tpp_list.size(), maxm = 0, maxn = 0;
         for (int s : tpp_list) { maxm = std::max(maxm, hfr[s].m); maxn = std::max(maxn, hfr[s].n); }
         memcpy(pin_i, tpp_list.data(), sizeof(int) * nt);
         CU(cudaMemcpyAsync(d_list, pin_i, sizeof(int) * nt, cudaMemcpyHostToDevice, st));
         const int ncolchunks = (maxn + 63) / 64;
         const int wcap = (std::min(maxn, PB) + NB - 1) / NB * NB;
         k_block<2><<<nt, 256, block_smem_bytes(wcap), st>>>(d_fr, d_list, 0, prm, 0, wcap, 0, 0);
         k_apply<<<dim3(nt, std::max(1, (maxm + AR - 1) / AR)), 256, apply_smem_bytes(wcap), st>>>(d_fr, d_list, 0, prm, wcap);
         k_commit<<<dim3(nt, 1 + ncolchunks + (maxm + 255) / 256), 256, 0, st>>>(d_fr, d_list, 0, ncolchunks, 0, 0, 1, 0);
         k_update<<<dim3(nt, 148), 256, UPD_SMEM, st>>>(d_fr, d_list, 0, nullptr, 0);
         if (trace) for (int s : tpp_list) trace_fronts("retry", s, 1);
         k_block<0><<<nt, 256, block_smem_bytes(NB), st>>>(d_fr, d_list, 0, prm, 1, NB, 0, 0);
         launches += 5;
         if (collect()) { cuda_err = (int)cudaGetLastError(); stats.cuda_error = cuda_err; return stats.flag = SSIDS_ERROR_CUDA_UNKNOWN; }
         if (flag < 0) break;
      }
      if (flag < 0) break;
      tpp_list.clear();
      for (int i = 0; i < ns; ++i) {
         if (!active[s0 + i]) continue;
         FrontDev& f = hfr[s0 + i];
         if (f.done < f.n) { tpp_list.push_back(s0 + i); nfp += f.n - f.done; }
      }
      if (!tpp_list.empty() && !posdef) {
         int nt = (int)tpp_list.size();
         memcpy(pin_i, tpp_list.data(), sizeof(int) * nt);
         CU(cudaMemcpyAsync(d_list, pin_i, sizeof(int) * nt, cudaMemcpyHostToDevice, st));
         k_tpp<<<nt, 256, 0, st>>>(d_fr, d_list, prm);
         ++launches;
         if (collect()) { cuda_err = (int)cudaGetLastError(); stats.cuda_error = cuda_err; return stats.flag = SSIDS_ERROR_CUDA_UNKNOWN; }
         for (int i = 0; i < ns; ++i) if (active[s0 + i]) nsp += hfr[s0 + i].n - hfr[s0 + i].done;
      } else if (!tpp_list.empty() && posdef) {
         flag = SSIDS_ERROR_NOT_POS_DEF;
      }
      if (flag < 0) break;
      for (int s = s0; s < s1; ++s) if (active[s]) { num_delay += hfr[s].n - hfr[s].nelim; hfr[s].active = 0; }
      if (timeline && !tl_ev.empty()) {
         cudaDeviceSynchronize();
         float prev_main = 0, prev_side = 0;
         for (size_t q = 0; q < tl_ev.size(); ++q) {
            float t = 0;
            cudaEventElapsedTime(&t, tl_ev[0], tl_ev[q]);
            const bool side = tl_name[q][0] == 's';
            fprintf(stderr, "[timeline] L%d %8.1f us (+%7.1f) %s\n", l, t * 1e3, (t - (side ? prev_side : prev_main)) * 1e3, tl_name[q].c_str());
            (side ? prev_side : prev_main) = t;
         }
         for (cudaEvent_t e : tl_ev) cudaEventDestroy(e);
         tl_ev.clear(); tl_name.clear();
      }
      if (lvl_times)
         fprintf(stderr, "[b200ldlt] level %d fronts %d launches %lld wall %.3f ms  (host: sizes %.3f, enqueue %.3f, wait %.3f)\n", l, ns,
                 (long long)(launches - lvl_l0), std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - lvl_t0).count(),
                 std::chrono::duration<double, std::milli>(lvl_t1 - lvl_t0).count(), std::chrono::duration<double, std::milli>(lvl_t2 - lvl_t1).count(),
                 std::chrono::duration<double, std::milli>(lvl_t3 - lvl_t2).count());
   }
   if (flag >= 0) {
      CU(cudaMemsetAsync(d_counters, 0, sizeof(int) * 8, st));
      if (!posdef && nn > 0) {
         k_stats<<<(nn + 7) / 8, 256, 0, st>>>(d_fr, nn, d_counters);
         ++launches;
      }
      CU(cudaMemcpyAsync(pin_i, d_counters, sizeof(int) * 8, cudaMemcpyDeviceToHost, st));
   }
   CU(cudaEventRecord(ev2, st));
   CU(cudaStreamSynchronize(st));
   {
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) { cuda_err = (int)e; stats.cuda_error = cuda_err; return stats.flag = SSIDS_ERROR_CUDA_UNKNOWN; }
   }
   float ms_all = 0, ms_dev = 0;
   cudaEventElapsedTime(&ms_all, ev0, ev2);
   cudaEventElapsedTime(&ms_dev, ev1, ev2);
   stats.t_total = ms_all * 1e-3;
   stats.t_device = ms_dev * 1e-3;
   stats.launches = launches;
   if (o.profile) {
      double ms_sum = 0;
      for (size_t q = 0; q + 1 < prof_used; q += 2) { float t = 0; cudaEventElapsedTime(&t, prof_ev[q], prof_ev[q + 1]); ms_sum += t; }
      stats.upd_seconds = ms_sum * 1e-3;
      stats.upd_flops = prof_flops;
      stats.upd_launches = (int64_t)(prof_used / 2);
   }
   stats.maxfront = maxfront;
   stats.num_delay = num_delay;
   stats.not_first_pass = nfp;
   stats.not_second_pass = nsp;
   stats.factor_bytes = (int64_t)(fac.capacity() + dar.capacity() + par.capacity());
   if (flag < 0) return stats.flag = flag;
   stats.num_neg = pin_i[0];
   stats.num_two = pin_i[1];
   stats.num_zero = pin_i[2];
   if (stats.num_zero > 0) flag = SSIDS_WARNING_FACT_SINGULAR;
   tables_dirty = true;      // solve offsets, level lists and chain roles are rebuilt by the next solve
   child_doff_final = child_doff;
   if (d_mine_mask && nn > 0) {
      CU(cudaMemsetAsync(d_mine_mask, 0, std::max(sym.n, 1), st));   // on st: the null stream is not ordered with it
      CU(cudaMemsetAsync(d_root_mask, 0, std::max(sym.n, 1), st));
      k_build_masks<<<(nn + 7) / 8, 256, 0, st>>>(d_fr, nn, d_mine_mask, d_root_mask);
      CU(cudaStreamSynchronize(st));
   }
   have = true;
   last = stats;
   last.flag = flag;
   return stats.flag = flag;
}

// Solve tables (work-vector offsets, per-level front lists, chain roles): host work that depends on the pivot
// counts of the finished factorisation; built by the first solve after a factorisation, not inside its timing.
int Numeric::build_solve_tables() {
   const int nn = P->nnodes;
   const std::vector<int>& child_doff = child_doff_final;
      std::vector<int64_t> off(nn + 1, 0);
      for (int s = 0; s < nn; ++s) off[s + 1] = off[s] + align_up(hfr[s].m, 2);
      xw_total = off[nn];
      CU(cudaMemcpy(d_xw_off, off.data(), sizeof(int64_t) * (nn + 1), cudaMemcpyHostToDevice));
      // partial-dot workspace of the backward solve: (row chunks + 1) x KB per front
      std::vector<int64_t> poff(nn + 1, 0);
      for (int s = 0; s < nn; ++s) poff[s + 1] = poff[s] + (int64_t)((hfr[s].m + 63) / 64 + 1) * KB;   // chunks of >= SE / KB = 64 rows
      part_total = poff[nn];
      CU(cudaMemcpy(d_part_off, poff.data(), sizeof(int64_t) * (nn + 1), cudaMemcpyHostToDevice));
      // per level: fronts sorted by eliminated columns (descending) so that the fronts active in
      // column block `step` form a prefix
      slist.assign(nn, 0);
      lvl_steps.assign(P->nlevels, 0);
      lvl_maxne.assign(P->nlevels, 0);
      lvl_maxm.assign(P->nlevels, 0);
      lvl_ngt32.assign(P->nlevels, 0);
      lvl_ngt64.assign(P->nlevels, 0);
      lvl_nact.assign(P->nlevels, std::vector<int>());
      lvl_actm.assign(P->nlevels, std::vector<int>());
      for (int l = 0; l < P->nlevels; ++l) {
         const int a0 = P->lvl_ptr[l], a1 = P->lvl_ptr[l + 1];
         for (int s = a0; s < a1; ++s) slist[s] = s;
         std::sort(slist.begin() + a0, slist.begin() + a1, [&](int x, int y) {
            const int ex = hfr[x].state == 2 ? hfr[x].nelim : 0, ey = hfr[y].state == 2 ? hfr[y].nelim : 0;
            if (ex != ey) return ex > ey;
            return x < y;
         });
         auto ne_of = [&](int s) { return hfr[s].state == 2 ? hfr[s].nelim : 0; };
         const int steps = a1 > a0 ? (ne_of(slist[a0]) + KB - 1) / KB : 0;
         lvl_steps[l] = steps;
         lvl_maxne[l] = a1 > a0 ? ne_of(slist[a0]) : 0;
         for (int s = a0; s < a1; ++s) if (hfr[s].state == 2) lvl_maxm[l] = std::max(lvl_maxm[l], hfr[s].m);
         lvl_ngt32[l] = lvl_ngt64[l] = 0;
         for (int q = a0; q < a1; ++q) {
            if (ne_of(slist[q]) > 64) ++lvl_ngt64[l];
            if (ne_of(slist[q]) > 32) ++lvl_ngt32[l];
         }
         lvl_nact[l].assign(steps, 0);
         lvl_actm[l].assign(steps, 0);
         for (int st2 = 0; st2 < steps; ++st2) {
            int na = 0, mm = 0;
            while (a0 + na < a1 && (ne_of(slist[a0 + na]) + KB - 1) / KB > st2) { mm = std::max(mm, hfr[slist[a0 + na]].m); ++na; }
            lvl_nact[l][st2] = na;
            lvl_actm[l][st2] = mm;
         }
      }
      if (nn > 0) CU(cudaMemcpy(d_slist, slist.data(), sizeof(int) * nn, cudaMemcpyHostToDevice));
      // chain kernels for the levels that hold fronts of more than KB eliminated columns
      {
         std::vector<int2> roles;
         roles_fwd_off.assign(P->nlevels, 0); roles_bwd_off.assign(P->nlevels, 0);
         roles_fwd_n.assign(P->nlevels, 0); roles_bwd_n.assign(P->nlevels, 0);
         for (int l = 0; l < P->nlevels; ++l) {
            // ... and for the levels with a tall front (one block of pivots over thousands of rows, the separator chains of a
            // KKT order): a level launch gives such a front one CTA, which then streams megabytes of L alone
            if (!use_chain || (lvl_maxne[l] <= KB && lvl_maxm[l] < chain_tall_m)) continue;
            const int a0 = P->lvl_ptr[l], a1 = P->lvl_ptr[l + 1];
            auto ne_of = [&](int s) { return hfr[s].state == 2 ? hfr[s].nelim : 0; };
            int maxr = 0;
            for (int s = a0; s < a1; ++s)
               if (ne_of(s) > 0) maxr = std::max(maxr, (ne_of(s) + KB - 1) / KB + (hfr[s].m - ne_of(s) + CH - 1) / CH);
            roles_fwd_off[l] = (int64_t)roles.size();
            for (int i = 0; i < maxr; ++i)
               for (int q = a0; q < a1; ++q) {
                  const int s = slist[q], ne = ne_of(s);
                  if (ne > 0 && i < (ne + KB - 1) / KB + (hfr[s].m - ne + CH - 1) / CH) roles.push_back(make_int2(s, i));
               }
            roles_fwd_n[l] = (int)(roles.size() - roles_fwd_off[l]);
            roles_bwd_off[l] = (int64_t)roles.size();
            for (int i = 0; i < maxr; ++i)
               for (int q = a0; q < a1; ++q) {
                  const int s = slist[q], nb = (ne_of(s) + KB - 1) / KB;
                  if (i < nb) roles.push_back(make_int2(s, nb - 1 - i));
               }
            roles_bwd_n[l] = (int)(roles.size() - roles_bwd_off[l]);
         }
         if (roles_cap < roles.size() + 1) {
            cudaFree(d_roles); d_roles = nullptr;
            CU(cudaMalloc(&d_roles, sizeof(int2) * (roles.size() + 1)));
            roles_cap = roles.size() + 1;
         }
         if (!roles.empty()) CU(cudaMemcpy(d_roles, roles.data(), sizeof(int2) * roles.size(), cudaMemcpyHostToDevice));
         CU(cudaMemsetAsync(d_chain_counter, 0, sizeof(unsigned long long) * 2 * (P->nlevels + 1), st));
         have_chain = !roles.empty();
         // the diagonal blocks' inverses and the tail chunks of the backward sweep for the same fronts
         have_minv = have_tail = false;
         troles_off.assign(P->nlevels, 0);
         troles_n.assign(P->nlevels, 0);
         if (have_chain && use_solve_inv > 1) {
            std::vector<int2> blist;
            std::vector<int> moff(std::max(nn, 1), -1);
            std::vector<int64_t> toff(std::max(nn, 1), -1);
            std::vector<int4> troles;
            int64_t tail_total = 0;
            for (int l = 0; l < P->nlevels; ++l) {
               if (roles_fwd_n[l] == 0) continue;
               troles_off[l] = (int64_t)troles.size();
               for (int q = P->lvl_ptr[l]; q < P->lvl_ptr[l + 1]; ++q) {
                  const int s = slist[q];
                  const int ne = hfr[s].state == 2 ? hfr[s].nelim : 0;
                  if (ne <= 0) continue;
                  const int nb = (ne + KB - 1) / KB, ntc = (hfr[s].m - ne + TAIL_ROWS - 1) / TAIL_ROWS;
                  if (nb > 1) {       // a single block is no chain: it keeps the sub-triangle path (and no 128 KB of inverse)
                     moff[s] = (int)blist.size();
                     for (int c = 0; c < nb; ++c) blist.push_back(make_int2(s, c));
                  }
                  // (a single block is worth the extra launch's partial sums only if its rows are many: the role would
                  // stream them alone, 20 MB through one SM for a 128-column front of 20 000 rows)
                  if (ntc > 0 && (nb > 1 || ntc >= 4)) {
                     toff[s] = tail_total;
                     tail_total += (int64_t)ntc * nb * KB;
                     for (int c = nb - 1; c >= 0; --c)
                        for (int t = 0; t < ntc; ++t) troles.push_back(make_int4(s, c, t, 0));
                  }
               }
               troles_n[l] = (int)(troles.size() - troles_off[l]);
            }
            const char* mb = getenv("B200LDLT_SOLVE_INV_MB");
            const size_t budget = (size_t)(mb ? atoll(mb) : 4096) << 20;
            if (!blist.empty() && blist.size() * KB * KB * 8 <= budget) {
               if (!d_minv_off) CU(cudaMalloc(&d_minv_off, sizeof(int) * std::max(nn, 1)));
               if (minv_cap < blist.size()) {
                  cudaFree(d_minv); cudaFree(d_blist); d_minv = nullptr; d_blist = nullptr;
                  if (cudaMalloc(&d_minv, blist.size() * KB * KB * 8) != cudaSuccess) { cudaGetLastError(); d_minv = nullptr; minv_cap = 0; }
                  else { CU(cudaMalloc(&d_blist, sizeof(int2) * blist.size())); minv_cap = blist.size(); }
               }
               if (mflag_cap < blist.size()) {
                  cudaFree(d_mflag); d_mflag = nullptr;
                  CU(cudaMalloc(&d_mflag, sizeof(int) * blist.size()));
                  mflag_cap = blist.size();
               }
               if (d_minv) {
                  CU(cudaMemcpy(d_minv_off, moff.data(), sizeof(int) * std::max(nn, 1), cudaMemcpyHostToDevice));
                  CU(cudaMemcpy(d_blist, blist.data(), sizeof(int2) * blist.size(), cudaMemcpyHostToDevice));
                  k_chain_inverses<<<(int)blist.size(), 256, chain_inv_smem(), st>>>(d_fr, d_blist, d_minv, d_mflag);
                  have_minv = true;
               }
            }
            if (!troles.empty()) {
               if (!d_tail_off) CU(cudaMalloc(&d_tail_off, sizeof(int64_t) * std::max(nn, 1)));
               if (tail_cap < (size_t)tail_total) {
                  cudaFree(d_tail); d_tail = nullptr;
                  tail_cap = 0;
                  if (cudaMalloc(&d_tail, sizeof(double) * (size_t)tail_total) == cudaSuccess) tail_cap = (size_t)tail_total;
                  else { cudaGetLastError(); d_tail = nullptr; }      // no room: the chain roles sum their rows themselves
               }
               if (troles_cap < troles.size()) {
                  cudaFree(d_troles); d_troles = nullptr;
                  CU(cudaMalloc(&d_troles, sizeof(int4) * troles.size()));
                  troles_cap = troles.size();
               }
               CU(cudaMemcpy(d_tail_off, toff.data(), sizeof(int64_t) * std::max(nn, 1), cudaMemcpyHostToDevice));
               CU(cudaMemcpy(d_troles, troles.data(), sizeof(int4) * troles.size(), cudaMemcpyHostToDevice));
               have_tail = d_tail != nullptr;
            }
         }
      }
      // child delay offsets must describe the final state for the solve
      if (!child_doff.empty())
         CU(cudaMemcpy(d_child_doff, child_doff.data(), sizeof(int) * child_doff.size(), cudaMemcpyHostToDevice));
   tables_dirty = false;
   return 0;
}

int Numeric::solve_prepare() {
   const int n = P->sym->n;
   if (tables_dirty) { int rc = build_solve_tables(); if (rc) return rc; }
   size_t need = (size_t)std::max(n, 1);
   if (x_cap < need) {
      cudaFree(d_x); cudaFree(d_xp);
      d_x = d_xp = nullptr;
      CU(cudaMalloc(&d_x, need * 8));
      CU(cudaMalloc(&d_xp, need * 8));
      x_cap = need;
   }
   if (xw_cap < 2 * ((size_t)xw_total + 2)) {   // two halves: work vectors | assembly-order accumulators
      cudaFree(d_xw);
      d_xw = nullptr;
      CU(cudaMalloc(&d_xw, 2 * ((size_t)xw_total + 2) * 8));
      xw_cap = 2 * ((size_t)xw_total + 2);
   }
   if (have_chain && cy_cap < (size_t)xw_total + 2) {
      cudaFree(d_cy);
      d_cy = nullptr;
      CU(cudaMalloc(&d_cy, ((size_t)xw_total + 2) * 8));
      cy_cap = (size_t)xw_total + 2;
   }
   if (part_cap < (size_t)part_total + 2) {
      cudaFree(d_part);
      d_part = nullptr;
      CU(cudaMalloc(&d_part, ((size_t)part_total + 2) * 8));
      part_cap = (size_t)part_total + 2;
   }
   return 0;
}

// launch helpers: instantiation by the widest block of the launch
static inline bool solve_pdl() { return !(getenv("B200LDLT_SOLVE_PDL") && atoi(getenv("B200LDLT_SOLVE_PDL")) == 0); }
// ST: the 64-thread instantiations, for FUSED levels whose fronts have at most SMALL_T_MAXM rows and 64 pivots (six CTAs per SM
// instead of two: a level of thousands of tiny fronts is bound by how many of them are resident)
constexpr int SMALL_T = 64, SMALL_T_MAXM = 96;
template <bool FUSED>
static void launch_fwd_step(int kbmax, dim3 grid, cudaStream_t st, const FrontDev* fr, const int* list, const int64_t* xw_off, double* xw,
                            double* xy, int sp, int kind, const GatherArgs& ga, bool st64 = false) {
   // levels of one launch each follow one another as programmatic dependents: the next level stages its L while this one runs
   const bool dep = FUSED && solve_pdl();
   if (FUSED && st64 && kbmax <= 32)
      launch_dep(dep, k_fwd_step<32, true, SMALL_T>, grid, dim3(SMALL_T), step_smem_fwd_t(32, SMALL_T), st, fr, list, xw_off, xw, xy, sp, kind, ga);
   else if (FUSED && st64 && kbmax <= 64)
      launch_dep(dep, k_fwd_step<64, true, SMALL_T>, grid, dim3(SMALL_T), step_smem_fwd_t(64, SMALL_T), st, fr, list, xw_off, xw, xy, sp, kind, ga);
   else if (FUSED && kbmax > 64 && kbmax <= 96)
      launch_dep(dep, k_fwd_step<96, true, 192>, grid, dim3(192), step_smem_fwd_t(96, 192), st, fr, list, xw_off, xw, xy, sp, kind, ga);
   else if (kbmax <= 32) launch_dep(dep, k_fwd_step<32, FUSED>, grid, dim3(256), step_smem_fwd(32), st, fr, list, xw_off, xw, xy, sp, kind, ga);
   else if (kbmax <= 64) launch_dep(dep, k_fwd_step<64, FUSED>, grid, dim3(256), step_smem_fwd(64), st, fr, list, xw_off, xw, xy, sp, kind, ga);
   else launch_dep(dep, k_fwd_step<128, FUSED>, grid, dim3(256), step_smem_fwd(128), st, fr, list, xw_off, xw, xy, sp, kind, ga);
}
template <bool FUSED>
static void launch_bwd_step(int kbmax, dim3 grid, cudaStream_t st, const FrontDev* fr, const int* list, const int64_t* xw_off, double* xw,
                            const int64_t* part_off, double* part, int* ticket, int sp, int kind, const GatherArgs& ga, bool st64 = false) {
   const bool dep = FUSED && solve_pdl();
   if (FUSED && st64 && kbmax <= 32)
      launch_dep(dep, k_bwd_step<32, true, SMALL_T>, grid, dim3(SMALL_T), step_smem_bwd_t(32, SMALL_T), st, fr, list, xw_off, xw, part_off, part, ticket, sp, kind, ga);
   else if (FUSED && st64 && kbmax <= 64)
      launch_dep(dep, k_bwd_step<64, true, SMALL_T>, grid, dim3(SMALL_T), step_smem_bwd_t(64, SMALL_T), st, fr, list, xw_off, xw, part_off, part, ticket, sp, kind, ga);
   else if (FUSED && kbmax > 64 && kbmax <= 96)
      launch_dep(dep, k_bwd_step<96, true, 192>, grid, dim3(192), step_smem_bwd_t(96, 192), st, fr, list, xw_off, xw, part_off, part, ticket, sp, kind, ga);
   else if (kbmax <= 32) launch_dep(dep, k_bwd_step<32, FUSED>, grid, dim3(256), step_smem_bwd(32), st, fr, list, xw_off, xw, part_off, part, ticket, sp, kind, ga);
   else if (kbmax <= 64) launch_dep(dep, k_bwd_step<64, FUSED>, grid, dim3(256), step_smem_bwd(64), st, fr, list, xw_off, xw, part_off, part, ticket, sp, kind, ga);
   else launch_dep(dep, k_bwd_step<128, FUSED>, grid, dim3(256), step_smem_bwd(128), st, fr, list, xw_off, xw, part_off, part, ticket, sp, kind, ga);
}
static inline int step_slots(int kbmax, bool bwd) { return kbmax <= 64 ? (bwd ? 3 : 2) * 148 : 148; }

// B200LDLT_SOLVE_TRACE: when each role of a chain launch started, had its inverses, saw its last dependency and published
void Numeric::chain_trace(const char* what, int l, int nroles, int off) {
   std::vector<unsigned long long> t(16 * (size_t)nroles);
   std::vector<int2> rl(nroles);
   cudaStreamSynchronize(st);
   cudaMemcpy(t.data(), d_chain_dbg, sizeof(unsigned long long) * t.size(), cudaMemcpyDeviceToHost);
   cudaMemcpy(rl.data(), d_roles + off, sizeof(int2) * nroles, cudaMemcpyDeviceToHost);
   unsigned long long t0 = ~0ULL, t3 = 0;
   for (int i = 0; i < nroles; ++i) { t0 = std::min(t0, t[16 * i]); t3 = std::max(t3, t[16 * i + 3]); }
   fprintf(stderr, "[b200ldlt] %s chain level %d: %d roles, %.1f us\n", what, l, nroles, (t3 - t0) * 1e-3);
   if (nroles > 80) return;
   const bool bwd = what[0] == 'b';
   for (int i = 0; i < nroles; ++i) {
      fprintf(stderr, "   cta %3d role (%d,%d): start %.1f inv %.1f deps %.1f done %.1f", i, rl[i].x, rl[i].y, (t[16 * i] - t0) * 1e-3,
              (t[16 * i + 1] - t0) * 1e-3, (t[16 * i + 2] - t0) * 1e-3, (t[16 * i + 3] - t0) * 1e-3);
      if (bwd) fprintf(stderr, "  reduced %.1f staged %.1f applied %.1f full %d", (t[16 * i + 4] - t0) * 1e-3, (t[16 * i + 5] - t0) * 1e-3,
                       (t[16 * i + 6] - t0) * 1e-3, (int)t[16 * i + 7]);
      if (bwd) { fprintf(stderr, "  warps"); for (int w = 0; w < 8; ++w) fprintf(stderr, " %.1f", (t[16 * i + 8 + w] - t0) * 1e-3); }
      fprintf(stderr, "\n");
   }
}

// forward sweep over the fronts of `kind` (-1: all)
void Numeric::sweep_fwd(int kind, int64_t& nlaunch) {
   if (have_chain) cudaMemsetAsync(d_cy, 0xff, (size_t)xw_total * 8, st);   // nothing published yet (CY_EMPTY)
   double* d_xa = d_xw + xw_cap / 2;
   GatherArgs ga{P->d_ss, P->d_child_idx, d_child_doff, P->d_rel, P->d_rlist, d_xp, posdef ? 1 : 0, 0};
   for (int l = 0; l < P->nlevels; ++l) {
      const int s0 = P->lvl_ptr[l], ns = P->lvl_ptr[l + 1] - s0;
      if (ns == 0) continue;
      const int kb0 = std::min(KB, lvl_maxne[l]), rows0 = lvl_steps[l] > 0 ? std::max(0, lvl_actm[l][0] - 1) : 0;
      if (use_fused && roles_fwd_n[l] == 0 && lvl_steps[l] == 1 &&
          step_grid_y(rows0, SE / (kb0 <= 32 ? 32 : kb0 <= 64 ? 64 : 128), step_slots(kb0, false), ns) == 1) {
         // every front of the level is one block and one CTA: gather, solve, update and store in one launch
         // one launch per width class when the wide fronts are a minority (the list is sorted by width): the narrow ones then
         // run two or three CTAs to the SM instead of one
         const bool st64 = use_st64 && lvl_maxm[l] <= SMALL_T_MAXM;
         int n128, n64, n32;
         split_by_width(l, kb0, ns, n128, n64, n32);
         if (n128 > 0) { launch_fwd_step<true>(kb0, dim3(n128, 1), st, d_fr, d_slist + s0, d_xw_off, d_xw, d_xa, 0, kind, ga, st64); ++nlaunch; smark("fwd_fused", l); }
         if (n64 > 0) { launch_fwd_step<true>(64, dim3(n64, 1), st, d_fr, d_slist + s0 + n128, d_xw_off, d_xw, d_xa, 0, kind, ga, st64); ++nlaunch; smark("fwd_fused", l); }
         if (n32 > 0) { launch_fwd_step<true>(32, dim3(n32, 1), st, d_fr, d_slist + s0 + n128 + n64, d_xw_off, d_xw, d_xa, 0, kind, ga, st64); ++nlaunch; smark("fwd_fused", l); }
         continue;
      }
      k_fwd_gather<<<ns, lvl_maxm[l] > 1024 ? 1024 : 256, 0, st>>>(d_fr, P->d_ss, P->d_child_idx, d_child_doff, P->d_rel, d_xw_off, d_xw, d_xa, d_xp, s0, kind);
      ++nlaunch;
      smark("k_fwd_gather", l);
      if (roles_fwd_n[l] > 0) {
         launch_dep(solve_pdl(), k_fwd_chain, dim3(roles_fwd_n[l]), dim3(256), chain_smem(), st, d_fr, d_roles + roles_fwd_off[l], roles_fwd_n[l],
                    d_chain_counter + 2 * l, d_xw_off, d_xw, d_cy, d_xp, posdef ? 1 : 0, kind, d_chain_err, use_solve_inv,
                    have_minv ? d_minv : nullptr, d_minv_off, d_mflag, d_chain_dbg);
         if (d_chain_dbg) chain_trace("fwd", l, roles_fwd_n[l], roles_fwd_off[l]);
         ++nlaunch;
         smark("k_fwd_chain", l);
         continue;      // the owners stored their pivots themselves
      }
      for (int sp = 0; sp < lvl_steps[l]; ++sp) {
         const int na = lvl_nact[l][sp];
         if (na == 0) break;
         const int kbmax = std::min(KB, lvl_maxne[l] - sp * KB);
         const int rows = std::max(0, lvl_actm[l][sp] - sp * KB - 1);
         // one wave of CTAs: a front's row chunks are dealt to at most `slots / fronts` CTAs
         const int rc = SE / (kbmax <= 32 ? 32 : kbmax <= 64 ? 64 : 128);
         launch_fwd_step<false>(kbmax, dim3(na, step_grid_y(rows, rc, step_slots(kbmax, false), na)), st, d_fr, d_slist + s0, d_xw_off, d_xw, d_xa, sp, kind, ga);
         ++nlaunch;
         smark("fwd_step", l);
      }
      k_fwd_store<<<ns, 256, 0, st>>>(d_fr, d_xw_off, d_xa, d_xp, s0, posdef ? 1 : 0, kind);
      ++nlaunch;
      smark("k_fwd_store", l);
   }
}

void Numeric::sweep_bwd(int kind, bool diag, int64_t& nlaunch) {
   if (have_chain) cudaMemsetAsync(d_cy, 0xff, (size_t)xw_total * 8, st);
   GatherArgs ga{P->d_ss, P->d_child_idx, d_child_doff, P->d_rel, P->d_rlist, d_xp, posdef ? 1 : 0, diag ? 1 : 0};
   for (int l = P->nlevels - 1; l >= 0; --l) {
      const int s0 = P->lvl_ptr[l], ns = P->lvl_ptr[l + 1] - s0;
      if (ns == 0) continue;
      const int kb0 = std::min(KB, lvl_maxne[l]), rows0 = lvl_steps[l] > 0 ? std::max(0, lvl_actm[l][0] - 1) : 0;
      if (use_fused && roles_bwd_n[l] == 0 && lvl_steps[l] == 1 &&
          step_grid_y(rows0, SE / (kb0 <= 32 ? 32 : kb0 <= 64 ? 64 : 128), step_slots(kb0, true), ns) == 1) {
         const bool st64 = use_st64 && lvl_maxm[l] <= SMALL_T_MAXM;
         const int na = lvl_nact[l][0];
         int n128, n64, n32;
         split_by_width(l, kb0, na, n128, n64, n32);
         if (n128 > 0) { launch_bwd_step<true>(kb0, dim3(n128, 1), st, d_fr, d_slist + s0, d_xw_off, d_xw, d_part_off, d_part, d_ticket, 0, kind, ga, st64); ++nlaunch; smark("bwd_fused", l); }
         if (n64 > 0) { launch_bwd_step<true>(64, dim3(n64, 1), st, d_fr, d_slist + s0 + n128, d_xw_off, d_xw, d_part_off, d_part, d_ticket, 0, kind, ga, st64); ++nlaunch; smark("bwd_fused", l); }
         if (n32 > 0) { launch_bwd_step<true>(32, dim3(n32, 1), st, d_fr, d_slist + s0 + n128 + n64, d_xw_off, d_xw, d_part_off, d_part, d_ticket, 0, kind, ga, st64); ++nlaunch; smark("bwd_fused", l); }
         continue;
      }
      {
         const int gy = std::max(1, std::min((lvl_maxm[l] + 255) / 256, 592 / ns));
         k_bwd_gather<<<dim3(ns, gy), 256, 0, st>>>(d_fr, P->d_ss, P->d_rlist, d_xw_off, d_xw, d_xp, s0, diag ? 1 : 0, posdef ? 1 : 0, kind);
         ++nlaunch;
         smark("k_bwd_gather", l);
      }
      if (roles_bwd_n[l] > 0) {
         if (have_tail && troles_n[l] > 0) {
            k_bwd_tail<<<troles_n[l], 256, 0, st>>>(d_fr, d_troles + troles_off[l], d_xw_off, d_xw, d_tail, d_tail_off, kind);
            ++nlaunch;
            smark("k_bwd_tail", l);
         }
         launch_dep(solve_pdl(), k_bwd_chain, dim3(roles_bwd_n[l]), dim3(256), chain_smem(), st, d_fr, d_roles + roles_bwd_off[l], roles_bwd_n[l],
                    d_chain_counter + 2 * l + 1, d_xw_off, d_xw, d_cy, d_xp, kind, d_chain_err, use_solve_inv, have_minv ? d_minv : nullptr, d_minv_off,
                    d_mflag, have_tail ? d_tail : nullptr, d_tail_off, d_chain_dbg);
         if (d_chain_dbg) chain_trace("bwd", l, roles_bwd_n[l], roles_bwd_off[l]);
         ++nlaunch;
         smark("k_bwd_chain", l);
         continue;
      }
      for (int sp = lvl_steps[l] - 1; sp >= 0; --sp) {
         const int na = lvl_nact[l][sp];
         if (na == 0) continue;
         const int kbmax = std::min(KB, lvl_maxne[l] - sp * KB);
         const int rows = std::max(0, lvl_actm[l][sp] - sp * KB - 1);
         const int rd = SE / (kbmax <= 32 ? 32 : kbmax <= 64 ? 64 : 128);
         launch_bwd_step<false>(kbmax, dim3(na, step_grid_y(rows, rd, step_slots(kbmax, true), na)), st, d_fr, d_slist + s0, d_xw_off, d_xw, d_part_off, d_part,
                                d_ticket, sp, kind, ga);
         ++nlaunch;
         smark("bwd_step", l);
      }
      k_bwd_store<<<ns, 256, 0, st>>>(d_fr, d_xw_off, d_xw, d_xp, s0, kind);
      ++nlaunch;
      smark("k_bwd_store", l);
   }
}

// up to `want` extra sets of solve workspaces, as many as fit in a quarter of the free memory; returns how many there are
int Numeric::ensure_extra_ws(int want) {
   const int nn = P->nnodes;
   const size_t sig[6] = {x_cap, xw_cap, cy_cap, part_cap, tail_cap, (size_t)nn};
   bool same = true;
   for (int i = 0; i < 6; ++i) same = same && sig[i] == extra_sig[i];
   if (!same) free_extra_ws();
   for (int i = 0; i < 6; ++i) extra_sig[i] = sig[i];
   want = std::min(want, SOLVE_STREAMS - 1);
   const size_t bytes = 8 * (2 * x_cap + xw_cap + cy_cap + part_cap + tail_cap) + 4 * (size_t)(nn + 1) + 16 * (size_t)(P->nlevels + 1) + 1024;
   while (n_extra_ws < want) {
      size_t fr_b = 0, tot_b = 0;
      if (cudaMemGetInfo(&fr_b, &tot_b) != cudaSuccess || bytes > fr_b / 4) break;
      SolveWs w;
      bool ok = cudaStreamCreateWithFlags(&w.st, cudaStreamNonBlocking) == cudaSuccess &&
                cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming) == cudaSuccess &&
                cudaMalloc(&w.x, 8 * std::max<size_t>(x_cap, 1)) == cudaSuccess && cudaMalloc(&w.xp, 8 * std::max<size_t>(x_cap, 1)) == cudaSuccess &&
                cudaMalloc(&w.xw, 8 * std::max<size_t>(xw_cap, 1)) == cudaSuccess && cudaMalloc(&w.cy, 8 * std::max<size_t>(cy_cap, 1)) == cudaSuccess &&
                cudaMalloc(&w.part, 8 * std::max<size_t>(part_cap, 1)) == cudaSuccess && cudaMalloc(&w.tail, 8 * std::max<size_t>(tail_cap, 1)) == cudaSuccess &&
                cudaMalloc(&w.ticket, 4 * (size_t)(nn + 1)) == cudaSuccess &&
                cudaMalloc(&w.counter, sizeof(unsigned long long) * 2 * (P->nlevels + 1)) == cudaSuccess;
      if (ok) ok = cudaMemset(w.ticket, 0, 4 * (size_t)(nn + 1)) == cudaSuccess &&
                   cudaMemset(w.counter, 0, sizeof(unsigned long long) * 2 * (P->nlevels + 1)) == cudaSuccess;
      extra_ws[n_extra_ws++] = w;
      if (!ok) {           // out of memory half way: give this set back, keep the ones before it
         cudaGetLastError();
         SolveWs& b = extra_ws[n_extra_ws - 1];
         cudaFree(b.x); cudaFree(b.xp); cudaFree(b.xw); cudaFree(b.cy); cudaFree(b.part); cudaFree(b.tail); cudaFree(b.ticket); cudaFree(b.counter);
         if (b.st) cudaStreamDestroy(b.st);
         if (b.done) cudaEventDestroy(b.done);
         b = SolveWs();
         --n_extra_ws;
         break;
      }
   }
   return std::min(n_extra_ws, want);
}

int Numeric::solve(int job, int nrhs, double* x, int ldx, bool on_device, double* seconds, int64_t* nl) {
   if (!have) return SSIDS_ERROR_CALL_SEQUENCE;
   if (job < 0 || job > 4) return SSIDS_ERROR_JOB_OOR;
   const Symbolic& sym = *P->sym;
   const int n = sym.n, nn = P->nnodes;
   if (ldx < n || nrhs < 1) return SSIDS_ERROR_X_SIZE;
   if (posdef && (job == 2 || job == 4)) return SSIDS_ERROR_JOB_OOR;   // ssids.f90:1180-1190
   CU(cudaSetDevice(P->device));
   int64_t nlaunch = 0;
   int rc = solve_prepare();
   if (rc) return rc;
   CU(cudaEventRecord(ev0, st));
   const int nb = (n + 255) / 256;
   // columns j, j + nws, ... on workspace set j mod nws (set 0 = the members, on the engine's stream)
   const int max_streams = getenv("B200LDLT_SOLVE_STREAMS") ? std::max(1, atoi(getenv("B200LDLT_SOLVE_STREAMS"))) : SOLVE_STREAMS;
   const SolveWs main_ws = current_ws();
   int nws = 1;
   if (nrhs > 1 && max_streams > 1 && nranks <= 1 && !solve_tl && !d_chain_dbg) nws = 1 + ensure_extra_ws(std::min(nrhs, max_streams) - 1);
   if (nws > 1) {
      if (!ev_ws_fork) CU(cudaEventCreateWithFlags(&ev_ws_fork, cudaEventDisableTiming));
      CU(cudaEventRecord(ev_ws_fork, st));
      for (int w = 0; w + 1 < nws; ++w) {
         CU(cudaStreamWaitEvent(extra_ws[w].st, ev_ws_fork, 0));
         // the ticket counters count modulo the roles of a level, which a new factorisation may have changed
         CU(cudaMemsetAsync(extra_ws[w].counter, 0, sizeof(unsigned long long) * 2 * (P->nlevels + 1), extra_ws[w].st));
      }
   }
   // groups of nws columns: everything of a group is enqueued before its results are copied back, because a copy to
   // pageable host memory holds the host until the column's stream has drained
   // (in a lambda: a failing CUDA call must not leave the members pointing at another column's workspaces)
   const int rc_cols = [&]() -> int {
   for (int j0 = 0; j0 < nrhs; j0 += nws) {
      for (int j = j0; j < std::min(nrhs, j0 + nws); ++j) {
         if (nws > 1) use_ws(j % nws == 0 ? main_ws : extra_ws[j % nws - 1]);
         double* xj = x + (size_t)j * ldx;
         double* dx = d_x;
         if (on_device) dx = xj;
         else CU(cudaMemcpyAsync(d_x, xj, sizeof(double) * n, cudaMemcpyHostToDevice, st));
         if (n > 0) { k_permute_in<<<nb, 256, 0, st>>>(dx, d_xp, P->d_order, (scaled && (job == 0 || job == 1)) ? d_scale : nullptr, n); ++nlaunch; }
         const bool fwd = (job == 0 || job == 1), bwd = (job == 0 || job == 3 || job == 4);
         const bool diag = !posdef && (job == 0 || job == 2 || job == 4);
         // posdef: LL^T semantics with L_chol = L sqrt(D); sqrt(D^-1) is folded into each sweep so
         // that the partial solves match SSIDS's posdef jobs
         if (fwd) sweep_fwd(-1, nlaunch);
         if (diag && !bwd && nn > 0) { k_solve_diag<<<(nn + 7) / 8, 256, 0, st>>>(d_fr, d_xp, nn); ++nlaunch; }
         if (bwd) sweep_bwd(-1, diag, nlaunch);
         if (n > 0) { k_permute_out<<<nb, 256, 0, st>>>(dx, d_xp, P->d_order, (scaled && (job == 0 || job == 3 || job == 4)) ? d_scale : nullptr, n); ++nlaunch; }
      }
      if (!on_device)
         for (int j = j0; j < std::min(nrhs, j0 + nws); ++j) {
            if (nws > 1) use_ws(j % nws == 0 ? main_ws : extra_ws[j % nws - 1]);
            CU(cudaMemcpyAsync(x + (size_t)j * ldx, d_x, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
         }
   }
   return 0;
   }();
   if (nws > 1) use_ws(main_ws);
   if (rc_cols) return rc_cols;
   if (nws > 1) {
      for (int w = 0; w + 1 < nws; ++w) {
         CU(cudaEventRecord(extra_ws[w].done, extra_ws[w].st));
         CU(cudaStreamWaitEvent(st, extra_ws[w].done, 0));
      }
   }
   CU(cudaEventRecord(ev1, st));
   pin_i[pin_i_cap - 1] = 0;
   CU(cudaMemcpyAsync(pin_i + pin_i_cap - 1, d_chain_err, sizeof(int), cudaMemcpyDeviceToHost, st));
   CU(cudaStreamSynchronize(st));
   {
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) { cuda_err = (int)e; return SSIDS_ERROR_CUDA_UNKNOWN; }
   }
   if (pin_i[pin_i_cap - 1] != 0) {   // a chain kernel gave up waiting for a block (solve_kernels.cuh): never expected
      CU(cudaMemsetAsync(d_chain_err, 0, sizeof(int), st));
      return SSIDS_ERROR_CUDA_UNKNOWN;
   }
   float ms = 0;
   cudaEventElapsedTime(&ms, ev0, ev1);
   if (seconds) *seconds = ms * 1e-3;
   if (nl) *nl = nlaunch;
   if (solve_tl && !stl_ev.empty()) {
      cudaEvent_t prev = ev0;
      for (size_t i = 0; i < stl_ev.size(); ++i) {
         float d = 0;
         cudaEventElapsedTime(&d, prev, stl_ev[i]);
         fprintf(stderr, "[b200ldlt] solve %-22s %8.1f us\n", stl_name[i].c_str(), d * 1e3);
         prev = stl_ev[i];
      }
      for (cudaEvent_t e : stl_ev) cudaEventDestroy(e);
      stl_ev.clear();
      stl_name.clear();
   }
   return 0;
}

// ---------------------------------------------------------------------------------------------
// multi-GPU: subtree partition, boundary fronts, phased solve (one process per GPU; the host side
// moves the packed boundary data between ranks with NCCL, galahad_b200/dist.py)
// ---------------------------------------------------------------------------------------------
int Numeric::set_partition(int rank, int nr) {
   int rc = init();
   if (rc) return rc;
   const Symbolic& sym = *P->sym;
   const int nn = P->nnodes;
   my_rank = rank;
   nranks = nr;
   nstages = std::max(1, sym.nstages);
   if ((int)sym.map_owner.size() < nn || (int)sym.map_stage.size() < nn) return SSIDS_ERROR_CALL_SEQUENCE;
   for (int node = 0; node < nn; ++node) {
      const int s = P->node_slot[node];
      owner[s] = sym.map_owner[node];
      part_kind[s] = sym.map_stage[node];
      const int par = sym.sparent[node];
      // the parent is factorised by a later call: the contribution block must outlive the level arenas
      boundary[s] = (par < nn && sym.map_stage[par] > sym.map_stage[node]) ? 1 : 0;
   }
   // which pivots this rank eliminated, and in which stage, is recorded after each factorisation phase
   // (k_build_masks): delayed pivots move to the front that finally eliminates them
   cudaFree(d_root_mask); cudaFree(d_mine_mask);
   d_root_mask = d_mine_mask = nullptr;
   CU(cudaMalloc(&d_root_mask, std::max(sym.n, 1)));
   CU(cudaMalloc(&d_mine_mask, std::max(sym.n, 1)));
   CU(cudaMemsetAsync(d_root_mask, 0, std::max(sym.n, 1), st));
   CU(cudaMemsetAsync(d_mine_mask, 0, std::max(sym.n, 1), st));
   CU(cudaStreamSynchronize(st));
   return 0;
}

// packed boundary front: 8 x int64 header {m, n, nelim, ldl, ldc, 0,0,0} | perm[n-nelim] (padded to 8 B)
// | L(:, nelim:n) (ldl x (n-nelim)) | C (ldc x (m-n))
// wire format of a boundary front (contrib_type, contrib.f90:19-33): header | pivot indices of the delayed columns | the
// delayed columns of L | the LOWER triangle of the contribution block, packed by columns
static inline int64_t packed_lower(int64_t t) { return t * (t + 1) / 2; }
int64_t Numeric::contrib_bytes(int node) const {
   const FrontDev& f = hfr[P->node_slot[node]];
   const int64_t nd = f.n - f.nelim;
   return 64 + (int64_t)align_up(nd * 4, 8) + 8 * (int64_t)f.ldl * nd + 8 * packed_lower(f.m - f.n);
}

int Numeric::contrib_pack(int node, char* dst) {
   CU(cudaSetDevice(P->device));
   const FrontDev& f = hfr[P->node_slot[node]];
   const int64_t nd = f.n - f.nelim;
   int64_t hdr[8] = {f.m, f.n, f.nelim, f.ldl, f.ldc, 0, 0, 0};
   CU(cudaMemcpyAsync(dst, hdr, 64, cudaMemcpyHostToDevice, st));
   char* p = dst + 64;
   if (nd > 0) {
      CU(cudaMemcpyAsync(p, f.perm + f.nelim, nd * 4, cudaMemcpyDeviceToDevice, st));
      p += align_up(nd * 4, 8);
      CU(cudaMemcpyAsync(p, f.L + (size_t)f.nelim * f.ldl, 8 * (size_t)f.ldl * nd, cudaMemcpyDeviceToDevice, st));
      p += 8 * (size_t)f.ldl * nd;
   }
   if (f.m > f.n) k_pack_lower<<<f.m - f.n, 256, 0, st>>>((double*)p, f.C, f.ldc, f.m - f.n);
   CU(cudaStreamSynchronize(st));
   return 0;
}

int Numeric::contrib_unpack(int node, const char* src, int64_t bytes) {
   CU(cudaSetDevice(P->device));
   const int s = P->node_slot[node];
   int64_t hdr[8];
   CU(cudaMemcpy(hdr, src, 64, cudaMemcpyDeviceToHost));
   FrontDev f = FrontDev();
   f.slot = s;
   f.m = (int)hdr[0]; f.n = (int)hdr[1]; f.nelim = (int)hdr[2]; f.ldl = (int)hdr[3]; f.ldc = (int)hdr[4];
   f.ncol = P->hss[s].ncol; f.nrow = P->hss[s].nrow; f.first_piv = P->hss[s].first_piv;
   f.ndelay_in = f.m - f.nrow;
   f.done = f.gend = f.nelim;
   f.state = 3; f.part_kind = 0;
   const int64_t nd = f.n - f.nelim, t = f.m - f.n;
   const int64_t head = 64 + (int64_t)align_up(nd * 4, 8) + 8 * (int64_t)f.ldl * nd;
   if (head + 8 * packed_lower(t) != bytes) return SSIDS_ERROR_UNKNOWN;
   // header, pivot indices and delayed columns as they travelled; the contribution block expanded to its square
   char* buf = (char*)bar.alloc((size_t)(head + 8 * (int64_t)f.ldc * t + 16));
   if (!buf) return SSIDS_ERROR_ALLOCATION;
   CU(cudaMemcpyAsync(buf, src, (size_t)head, cudaMemcpyDeviceToDevice, st));
   char* p = buf + 64;
   f.perm = (int*)p - f.nelim;                       // only perm[nelim..n) is ever read
   f.orig = nullptr;
   p += align_up(nd * 4, 8);
   f.L = (double*)p - (size_t)f.nelim * f.ldl;      // only columns nelim..n are ever read
   p += 8 * (size_t)f.ldl * nd;
   f.C = (double*)p;
   if (t > 0) k_unpack_lower<<<(unsigned)t, 256, 0, st>>>(f.C, (const double*)(src + head), f.ldc, (int)t);
   hfr[s] = f;
   CU(cudaMemcpyAsync(d_fr + s, &hfr[s], sizeof(FrontDev), cudaMemcpyHostToDevice, st));
   CU(cudaStreamSynchronize(st));
   return 0;
}

int64_t Numeric::xtail_bytes(int node) const {
   const FrontDev& f = hfr[P->node_slot[node]];
   return 8 * (int64_t)(f.m - f.nelim);
}
int Numeric::xtail_pack(int node, char* dst) {
   CU(cudaSetDevice(P->device));
   const int s = P->node_slot[node];
   const FrontDev& f = hfr[s];
   std::vector<int64_t> off(1);
   CU(cudaMemcpy(off.data(), d_xw_off + s, 8, cudaMemcpyDeviceToHost));
   if (f.m > f.nelim) CU(cudaMemcpyAsync(dst, d_xw + off[0] + f.nelim, 8 * (size_t)(f.m - f.nelim), cudaMemcpyDeviceToDevice, st));
   CU(cudaStreamSynchronize(st));
   return 0;
}
int Numeric::xtail_unpack(int node, const char* src) {
   CU(cudaSetDevice(P->device));
   const int s = P->node_slot[node];
   const FrontDev& f = hfr[s];
   std::vector<int64_t> off(1);
   CU(cudaMemcpy(off.data(), d_xw_off + s, 8, cudaMemcpyDeviceToHost));
   if (f.m > f.nelim) CU(cudaMemcpyAsync(d_xw + off[0] + f.nelim, src, 8 * (size_t)(f.m - f.nelim), cudaMemcpyDeviceToDevice, st));
   CU(cudaStreamSynchronize(st));
   return 0;
}

// step 1: permute b in, forward over my leaf parts.  step 2 (root owner): forward, D, backward over
// the root part.  step 3: take the root part's entries from `src` (the owner's pivot-order vector),
// then D + backward over my leaf parts.  step 4: scatter my entries to x (zero elsewhere).
int Numeric::dist_solve(int step, double* d_xio, const double* src) {
   if (!have) return SSIDS_ERROR_CALL_SEQUENCE;
   CU(cudaSetDevice(P->device));
   const int n = P->sym->n, nb = (n + 255) / 256;
   int64_t nl = 0;
   int rc = solve_prepare();
   if (rc) return rc;
   (void)src;
   if (step == 1) {                                   // right-hand side into pivot order
      k_permute_in<<<nb, 256, 0, st>>>(d_xio, d_xp, P->d_order, scaled ? d_scale : nullptr, n);
   } else if (step >= 100 && step < 200) {            // forward sweep over this rank's fronts of stage step - 100
      sweep_fwd(step - 100, nl);
   } else if (step >= 200 && step < 300) {            // diagonal + backward sweep of stage step - 200
      sweep_bwd(step - 200, !posdef, nl);
   } else if (step == 4) {                            // the whole solution in the caller's order
      k_permute_out<<<nb, 256, 0, st>>>(d_xio, d_xp, P->d_order, scaled ? d_scale : nullptr, n);
   } else {
      return SSIDS_ERROR_JOB_OOR;
   }
   pin_i[pin_i_cap - 1] = 0;
   CU(cudaMemcpyAsync(pin_i + pin_i_cap - 1, d_chain_err, sizeof(int), cudaMemcpyDeviceToHost, st));
   CU(cudaStreamSynchronize(st));
   cudaError_t e = cudaGetLastError();
   if (e != cudaSuccess) { cuda_err = (int)e; return SSIDS_ERROR_CUDA_UNKNOWN; }
   if (pin_i[pin_i_cap - 1] != 0) {   // a chain kernel gave up waiting for a block, as in Numeric::solve
      CU(cudaMemsetAsync(d_chain_err, 0, sizeof(int), st));
      return SSIDS_ERROR_CUDA_UNKNOWN;
   }
   return 0;
}

Numeric* numeric_create(Plan* p) { return new Numeric(p); }
int numeric_set_partition(Numeric* n, int rank, int nranks) { return n->set_partition(rank, nranks); }
int numeric_factor_phase(Numeric* n, const double* val, const FactorOptions& o, FactorStats& s, int phase) {
   int rc = n->factor(val, false, o, s, phase);
   s.cuda_error = n->cuda_err;
   return rc;
}
int64_t numeric_contrib_bytes(Numeric* n, int node) { return n->contrib_bytes(node); }
int numeric_contrib_pack(Numeric* n, int node, void* dst) { return n->contrib_pack(node, (char*)dst); }
int numeric_contrib_unpack(Numeric* n, int node, const void* src, int64_t bytes) { return n->contrib_unpack(node, (const char*)src, bytes); }
int64_t numeric_xtail_bytes(Numeric* n, int node) { return n->xtail_bytes(node); }
int numeric_xtail_pack(Numeric* n, int node, void* dst) { return n->xtail_pack(node, (char*)dst); }
int numeric_xtail_unpack(Numeric* n, int node, const void* src) { return n->xtail_unpack(node, (const char*)src); }
int numeric_dist_solve(Numeric* n, int step, double* d_x, const double* src) { return n->dist_solve(step, d_x, src); }
double* numeric_xp(Numeric* n) { return n->d_xp; }
int numeric_xp_get_stage(Numeric* N, int stage, void* d_dst, int n) {
   cudaSetDevice(N->P->device);
   if (!N->d_mine_mask || !N->d_xp) return SSIDS_ERROR_CALL_SEQUENCE;
   k_get_stage<<<(n + 255) / 256, 256, 0, N->st>>>((double*)d_dst, N->d_xp, N->d_mine_mask, N->d_root_mask, stage, n);
   return cudaStreamSynchronize(N->st) == cudaSuccess ? 0 : SSIDS_ERROR_CUDA_UNKNOWN;
}
int numeric_xp_merge_stage(Numeric* N, int stage, const void* d_src, int n) {
   (void)stage;
   cudaSetDevice(N->P->device);
   if (!N->d_xp) return SSIDS_ERROR_CALL_SEQUENCE;
   k_merge_stage<<<(n + 255) / 256, 256, 0, N->st>>>(N->d_xp, (const double*)d_src, n);
   return cudaStreamSynchronize(N->st) == cudaSuccess ? 0 : SSIDS_ERROR_CUDA_UNKNOWN;
}
int numeric_xp_get(Numeric* N, void* d_dst, int n) {
   cudaSetDevice(N->P->device);
   return cudaMemcpy(d_dst, N->d_xp, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice) == cudaSuccess ? 0 : SSIDS_ERROR_CUDA_UNKNOWN;
}
void numeric_destroy(Numeric* n) { delete n; }
int Numeric::set_value_map(int ne, const int* maps) {
   int rc = init();
   if (rc) return rc;
   CU(cudaSetDevice(P->device));
   const int64_t nz = P->sym->nz;
   std::vector<int> src, dst;
   std::vector<std::vector<int>> dups;        // per destination, in input order
   std::vector<int> dup_of(std::max<int64_t>(nz, 1), -1), dup_dst;
   for (int l = 0; l < ne; ++l) {
      const int k = maps[l];
      if (k > 0) { if (k > nz) return SSIDS_ERROR_VAL; src.push_back(l); dst.push_back(k - 1); }
      else if (k < 0) {
         const int t = -k - 1;
         if (t >= nz) return SSIDS_ERROR_VAL;
         if (dup_of[t] < 0) { dup_of[t] = (int)dups.size(); dups.emplace_back(); dup_dst.push_back(t); }
         dups[dup_of[t]].push_back(l);
      }
   }
   std::vector<int> dptr(dups.size() + 1, 0), dsrc;
   for (size_t t = 0; t < dups.size(); ++t) { dptr[t + 1] = dptr[t] + (int)dups[t].size(); dsrc.insert(dsrc.end(), dups[t].begin(), dups[t].end()); }
   cudaFree(d_uval); cudaFree(d_vm_src); cudaFree(d_vm_dst); cudaFree(d_vm_dup_dst); cudaFree(d_vm_dup_ptr); cudaFree(d_vm_dup_src);
   d_uval = nullptr; d_vm_src = d_vm_dst = d_vm_dup_dst = d_vm_dup_ptr = d_vm_dup_src = nullptr;
   vm_ne = ne; vm_nset = (int)src.size(); vm_ndup = (int)dups.size();
   uval_valid = false;
   CU(cudaMalloc(&d_uval, sizeof(double) * std::max(ne, 1)));
   CU(cudaMalloc(&d_vm_src, sizeof(int) * std::max<size_t>(src.size(), 1)));
   CU(cudaMalloc(&d_vm_dst, sizeof(int) * std::max<size_t>(dst.size(), 1)));
   CU(cudaMalloc(&d_vm_dup_dst, sizeof(int) * std::max<size_t>(dup_dst.size(), 1)));
   CU(cudaMalloc(&d_vm_dup_ptr, sizeof(int) * dptr.size()));
   CU(cudaMalloc(&d_vm_dup_src, sizeof(int) * std::max<size_t>(dsrc.size(), 1)));
   if (!src.empty()) {
      CU(cudaMemcpy(d_vm_src, src.data(), sizeof(int) * src.size(), cudaMemcpyHostToDevice));
      CU(cudaMemcpy(d_vm_dst, dst.data(), sizeof(int) * dst.size(), cudaMemcpyHostToDevice));
   }
   if (!dup_dst.empty()) {
      CU(cudaMemcpy(d_vm_dup_dst, dup_dst.data(), sizeof(int) * dup_dst.size(), cudaMemcpyHostToDevice));
      CU(cudaMemcpy(d_vm_dup_src, dsrc.data(), sizeof(int) * dsrc.size(), cudaMemcpyHostToDevice));
   }
   CU(cudaMemcpy(d_vm_dup_ptr, dptr.data(), sizeof(int) * dptr.size(), cudaMemcpyHostToDevice));
   return 0;
}

// user's values (host, ne of them) -> d_val through the value map; the factorisation then runs on d_val in place
int Numeric::gather_values(int ne, const double* uval) {
   if (!d_uval || ne != vm_ne) return SSIDS_ERROR_CALL_SEQUENCE;
   CU(cudaSetDevice(P->device));
   const int64_t nz = P->sym->nz;
   if (ne > 0 && uval) CU(cudaMemcpyAsync(d_uval, uval, sizeof(double) * ne, cudaMemcpyHostToDevice, st));
   if (uval) uval_valid = true;
   if (nz > 0) CU(cudaMemsetAsync(d_val, 0, sizeof(double) * nz, st));
   if (vm_nset > 0) k_gather_set<<<(vm_nset + 255) / 256, 256, 0, st>>>(d_val, d_uval, d_vm_src, d_vm_dst, vm_nset);
   if (vm_ndup > 0) k_gather_dup<<<(vm_ndup + 255) / 256, 256, 0, st>>>(d_val, d_uval, d_vm_dup_dst, d_vm_dup_ptr, d_vm_dup_src, vm_ndup);
   return 0;
}

// nchg changed entries of the caller's values (0-based positions in the caller's arrays) into the resident copy
int Numeric::update_values(int nchg, const int* which, const double* newval) {
   if (!d_uval || !uval_valid) return SSIDS_ERROR_CALL_SEQUENCE;       // needs a complete factor_mapped first
   if (nchg == 0) return 0;
   for (int i = 0; i < nchg; ++i) if (which[i] < 0 || which[i] >= vm_ne) return SSIDS_ERROR_CALL_SEQUENCE;
   CU(cudaSetDevice(P->device));
   if (chg_cap < (size_t)nchg) {
      cudaFree(d_chg_idx); cudaFree(d_chg_val);
      d_chg_idx = nullptr; d_chg_val = nullptr; chg_cap = 0;
      CU(cudaMalloc(&d_chg_idx, sizeof(int) * nchg));
      CU(cudaMalloc(&d_chg_val, sizeof(double) * nchg));
      chg_cap = (size_t)nchg;
   }
   CU(cudaMemcpyAsync(d_chg_idx, which, sizeof(int) * nchg, cudaMemcpyHostToDevice, st));
   CU(cudaMemcpyAsync(d_chg_val, newval, sizeof(double) * nchg, cudaMemcpyHostToDevice, st));
   k_update_user_values<<<(nchg + 255) / 256, 256, 0, st>>>(d_uval, d_chg_idx, d_chg_val, nchg);
   return 0;
}

// scale: n factors in variable order for the next factorisations and their solves, or null for none
int Numeric::set_scaling(const double* scale) {
   int rc = init();
   if (rc) return rc;
   CU(cudaSetDevice(P->device));
   const Symbolic& sym = *P->sym;
   if (!scale) { scaled = false; return 0; }
   if (!d_scale) CU(cudaMalloc(&d_scale, sizeof(double) * std::max(sym.n, 1)));
   if (!d_a_rowi) {
      std::vector<int> coli((size_t)std::max<int64_t>(sym.nz, 1));
      for (int c = 0; c < sym.n; ++c)
         for (int64_t j = sym.ptr[c]; j < sym.ptr[c + 1]; ++j) coli[j] = c;
      CU(cudaMalloc(&d_a_rowi, sizeof(int) * std::max<int64_t>(sym.nz, 1)));
      CU(cudaMalloc(&d_a_coli, sizeof(int) * std::max<int64_t>(sym.nz, 1)));
      if (sym.nz > 0) {
         CU(cudaMemcpy(d_a_rowi, sym.row.data(), sizeof(int) * sym.nz, cudaMemcpyHostToDevice));
         CU(cudaMemcpy(d_a_coli, coli.data(), sizeof(int) * sym.nz, cudaMemcpyHostToDevice));
      }
   }
   CU(cudaMemcpy(d_scale, scale, sizeof(double) * sym.n, cudaMemcpyHostToDevice));
   scaled = true;
   return 0;
}

int Numeric::set_user_matrix(int ne, const int* row, const int* col) {
   int rc = init();
   if (rc) return rc;
   CU(cudaSetDevice(P->device));
   const int n = P->sym->n;
   std::vector<int> rptr(n + 2, 0);
   for (int l = 0; l < ne; ++l) {
      const int i = row[l], j = col[l];
      if (std::min(i, j) < 1 || std::max(i, j) > n) continue;
      ++rptr[i + 1];
      if (i != j) ++rptr[j + 1];
   }
   for (int i = 0; i <= n; ++i) rptr[i + 1] += rptr[i];
   std::vector<int> rent(std::max(rptr[n + 1], 1)), rcol(std::max(rptr[n + 1], 1)), fill(rptr.begin() + 1, rptr.end());
   for (int l = 0; l < ne; ++l) {          // input order within every row: the order of the reference's loop
      const int i = row[l], j = col[l];
      if (std::min(i, j) < 1 || std::max(i, j) > n) continue;
      rent[fill[i - 1]] = l; rcol[fill[i - 1]++] = j - 1;
      if (i != j) { rent[fill[j - 1]] = l; rcol[fill[j - 1]++] = i - 1; }
   }
   cudaFree(d_um_rptr); cudaFree(d_um_rent); cudaFree(d_um_rcol); cudaFree(d_rb); cudaFree(d_rx); cudaFree(d_rr); cudaFree(d_rnorm);
   d_um_rptr = d_um_rent = d_um_rcol = nullptr; d_rb = d_rx = d_rr = nullptr; d_rnorm = nullptr;
   CU(cudaMalloc(&d_um_rptr, sizeof(int) * (n + 1)));
   CU(cudaMalloc(&d_um_rent, sizeof(int) * rent.size()));
   CU(cudaMalloc(&d_um_rcol, sizeof(int) * rcol.size()));
   CU(cudaMalloc(&d_rb, sizeof(double) * std::max(n, 1)));
   CU(cudaMalloc(&d_rx, sizeof(double) * std::max(n, 1)));
   CU(cudaMalloc(&d_rr, sizeof(double) * std::max(n, 1)));
   CU(cudaMalloc(&d_rnorm, sizeof(unsigned long long)));
   CU(cudaMemcpy(d_um_rptr, rptr.data() + 1, sizeof(int) * (n + 1), cudaMemcpyHostToDevice));
   CU(cudaMemcpy(d_um_rent, rent.data(), sizeof(int) * rent.size(), cudaMemcpyHostToDevice));
   CU(cudaMemcpy(d_um_rcol, rcol.data(), sizeof(int) * rcol.size(), cudaMemcpyHostToDevice));
   um_ne = ne;
   return 0;
}

// x: b on entry, the refined solution on return (host).  x_0 = 0, r_0 = b; repeat { solve A dx = r; x += dx; r = b - A x }
// until max |r_i| < max(stop_abs, stop_rel max |b_i|) or max_ref refinements (sls.f90:4806-4945)
int Numeric::solve_refined(double* x, int max_ref, double stop_abs, double stop_rel, int* nref, double* seconds) {
   if (!have) return SSIDS_ERROR_CALL_SEQUENCE;
   if (um_ne < 0 || um_ne != vm_ne || !d_uval) return SSIDS_ERROR_CALL_SEQUENCE;   // needs set_user_matrix + factor_mapped
   CU(cudaSetDevice(P->device));
   const int n = P->sym->n, nb = (n + 255) / 256;
   double bnorm = 0.0;
   for (int i = 0; i < n; ++i) bnorm = std::max(bnorm, std::fabs(x[i]));
   const double stop = std::max(stop_abs, stop_rel * bnorm);
   CU(cudaMemcpyAsync(d_rb, x, sizeof(double) * n, cudaMemcpyHostToDevice, st));
   CU(cudaMemcpyAsync(d_rr, d_rb, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
   CU(cudaMemsetAsync(d_rx, 0, sizeof(double) * n, st));
   CU(cudaStreamSynchronize(st));
   double total = 0.0;
   int it = 0, flag = 0;
   for (it = 0; it <= max_ref; ++it) {
      double secs = 0.0;
      int64_t nl = 0;
      flag = solve(0, 1, d_rr, n, true, &secs, &nl);
      total += secs;
      if (flag < 0) break;
      k_axpy1<<<nb, 256, 0, st>>>(d_rx, d_rr, n);
      if (it == max_ref) break;
      k_residual<<<nb, 256, 0, st>>>(d_rr, d_rb, d_rx, d_uval, d_um_rptr, d_um_rent, d_um_rcol, n);
      CU(cudaMemsetAsync(d_rnorm, 0, sizeof(unsigned long long), st));
      k_maxabs<<<148, 256, 0, st>>>(d_rr, n, d_rnorm);
      unsigned long long bits = 0;
      CU(cudaMemcpyAsync(&bits, d_rnorm, sizeof(bits), cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      double rn;
      std::memcpy(&rn, &bits, sizeof(rn));
      if (rn < stop) break;
   }
   if (nref) *nref = std::min(it, max_ref);
   if (flag >= 0) {
      CU(cudaMemcpyAsync(x, d_rx, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
   }
   if (seconds) *seconds = total;
   return flag;
}

int numeric_set_scaling(Numeric* n, const double* scale) { return n->set_scaling(scale); }
int numeric_set_user_matrix(Numeric* n, int ne, const int* row, const int* col) { return n->set_user_matrix(ne, row, col); }
int numeric_solve_refined(Numeric* n, double* x, int max_ref, double stop_abs, double stop_rel, int* nref, double* seconds) {
   return n->solve_refined(x, max_ref, stop_abs, stop_rel, nref, seconds);
}
int numeric_set_value_map(Numeric* n, int ne, const int* maps) { return n->set_value_map(ne, maps); }
int numeric_factor_mapped(Numeric* n, int ne, const double* uval, const FactorOptions& o, FactorStats& s) {
   int rc = n->gather_values(ne, uval);
   if (rc) return s.flag = rc;
   return n->factor(n->d_val, true, o, s);
}
int numeric_factor_changed(Numeric* n, int nchg, const int* which, const double* newval, const FactorOptions& o, FactorStats& s) {
   int rc = n->update_values(nchg, which, newval);
   if (rc == 0) rc = n->gather_values(n->vm_ne, nullptr);     // from the resident values
   if (rc) { s = FactorStats(); return s.flag = rc; }
   return n->factor(n->d_val, true, o, s);
}
int numeric_factor(Numeric* n, const double* val, bool on_device, const FactorOptions& o, FactorStats& s) {
   int rc = n->factor(val, on_device, o, s);
   s.cuda_error = n->cuda_err;
   return rc;
}
int numeric_solve(Numeric* n, int job, int nrhs, double* x, int ldx, bool on_device, double* seconds, int64_t* nl) {
   return n->solve(job, nrhs, x, ldx, on_device, seconds, nl);
}
bool numeric_is_posdef(const Numeric* n) { return n->posdef; }
bool numeric_has_factors(const Numeric* n) { return n->have; }

int numeric_node_info(Numeric* N, int node, int* m, int* n, int* nelim) {
   if (!N->have || node < 0 || node >= N->P->nnodes) return -1;
   const FrontDev& f = N->hfr[N->P->node_slot[node]];
   *m = f.m; *n = f.n; *nelim = f.nelim;
   return 0;
}
int numeric_node_copy(Numeric* N, int node, double* L, double* d, int* perm) {
   if (!N->have || node < 0 || node >= N->P->nnodes) return -1;
   cudaSetDevice(N->P->device);
   const FrontDev& f = N->hfr[N->P->node_slot[node]];
   if (f.state != 2) return -3;      // absent on this rank, or imported (only its trailing block lives here)
   if (L && f.nelim > 0)
      if (cudaMemcpy2D(L, sizeof(double) * f.m, f.L, sizeof(double) * f.ldl, sizeof(double) * f.m, f.nelim, cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
   if (d && f.n > 0 && cudaMemcpy(d, f.D, sizeof(double) * 2 * f.n, cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
   if (perm && f.n > 0 && cudaMemcpy(perm, f.perm, sizeof(int) * f.n, cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
   return 0;
}

// piv_order / d in pivot-position indexing (enquire, NumericSubtree.hxx:418-462)
int numeric_enquire(Numeric* N, int* piv_order, double* d) {
   if (!N->have) return SSIDS_ERROR_CALL_SEQUENCE;
   cudaSetDevice(N->P->device);
   const int nn = N->P->nnodes;
   std::vector<double> dd;
   std::vector<int> pp;
   int piv = 0;
   for (int node = 0; node < nn; ++node) {
      const FrontDev& f = N->hfr[N->P->node_slot[node]];
      if (f.nelim == 0) continue;
      dd.resize(2 * (size_t)f.n + 2);
      pp.resize(f.n);
      if (cudaMemcpy(dd.data(), f.D, sizeof(double) * (2 * (size_t)f.n + 2), cudaMemcpyDeviceToHost) != cudaSuccess) return SSIDS_ERROR_CUDA_UNKNOWN;
      if (cudaMemcpy(pp.data(), f.perm, sizeof(int) * f.n, cudaMemcpyDeviceToHost) != cudaSuccess) return SSIDS_ERROR_CUDA_UNKNOWN;
      for (int i = 0; i < f.nelim;) {
         if (i + 1 == f.nelim || std::isfinite(dd[2 * i + 2])) {
            if (piv_order) piv_order[pp[i]] = piv + 1;
            if (d) { d[2 * piv] = dd[2 * i]; d[2 * piv + 1] = 0.0; }
            piv += 1; i += 1;
         } else {
            if (piv_order) { piv_order[pp[i]] = -(piv + 1); piv_order[pp[i + 1]] = -(piv + 2); }
            if (d) { d[2 * piv] = dd[2 * i]; d[2 * piv + 1] = dd[2 * i + 1]; d[2 * piv + 2] = dd[2 * i + 3]; d[2 * piv + 3] = 0.0; }
            piv += 2; i += 2;
         }
      }
   }
   return 0;
}

// overwrite D^{-1} (alter, NumericSubtree.hxx:465-477); the 2x2 tags are kept
int numeric_alter(Numeric* N, const double* d) {
   if (!N->have) return SSIDS_ERROR_CALL_SEQUENCE;
   if (N->posdef) return SSIDS_ERROR_NOT_LDLT;
   cudaSetDevice(N->P->device);
   const int nn = N->P->nnodes;
   std::vector<double> dd;
   int piv = 0;
   for (int node = 0; node < nn; ++node) {
      const FrontDev& f = N->hfr[N->P->node_slot[node]];
      if (f.nelim == 0) continue;
      dd.resize(2 * (size_t)f.n + 2);
      if (cudaMemcpy(dd.data(), f.D, sizeof(double) * (2 * (size_t)f.n + 2), cudaMemcpyDeviceToHost) != cudaSuccess) return SSIDS_ERROR_CUDA_UNKNOWN;
      for (int i = 0; i < f.nelim;) {
         if (i + 1 == f.nelim || std::isfinite(dd[2 * i + 2])) {
            dd[2 * i] = d[2 * piv];
            piv += 1; i += 1;
         } else {
            dd[2 * i] = d[2 * piv]; dd[2 * i + 1] = d[2 * piv + 1]; dd[2 * i + 3] = d[2 * piv + 2];
            piv += 2; i += 2;
         }
      }
      if (cudaMemcpy(f.D, dd.data(), sizeof(double) * (2 * (size_t)f.n + 2), cudaMemcpyHostToDevice) != cudaSuccess) return SSIDS_ERROR_CUDA_UNKNOWN;
   }
   return 0;
}

}  // namespace b200
