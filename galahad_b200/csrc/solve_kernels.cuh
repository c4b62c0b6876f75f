// Level-scheduled triangular solves (included by engine.cu).
//
// Reference behaviour: NumericSubtree.hxx:280-400 (gather, trsv/gemv per node, scatter) and
// ldlt_app.cxx:2537-2588.  Here every level of the (fused) assembly tree is processed by batched
// kernels; a front's L is walked in column blocks of KB, ONE launch per block and direction:
//   forward  (k_fwd_step): every CTA of a front stages the block's KB x KB triangle in shared memory and
//            solves it (redundantly: the result is needed by all of them and there is no inter-CTA wait),
//            then applies the block's columns to its own chunk of the rows below, whose entries of L it
//            requested before touching the triangle;
//   backward (k_bwd_step): every CTA forms the partial dot products of its chunk of rows with the block's
//            columns; the CTA that finishes last (ticket counter) sums the partials in chunk order and solves
//            the transposed triangle.
// Levels with fronts of more than KB pivots (or with one very tall front) take the chain kernels further down instead: one
// launch per level and direction, the blocks of a front passed from CTA to CTA through global memory, the diagonal blocks
// through inverses computed once per factorisation (k_chain_inverses), the rows below the pivots of the backward sweep
// in k_bwd_tail (profiles/r2_notes.md part 7).  Consecutive launches overlap as programmatic dependents (pdl_*).
// Everything here is bound by memory latency, not bandwidth (ncu launch list, profiles/r1_notes.md part 3):
// each thread therefore has all of its loads (32 per thread) in flight at once, and dependent work runs from
// shared memory.  No floating-point atomics: children's contributions are gathered by the parent in child
// order and partial dot products are summed in chunk order, so a solve is bit-reproducible.
#pragma once

constexpr int KB = 128;      // solve column block
constexpr int SE = 8192;     // entries of L per CTA in the update / dot phases: rows x block columns, 32 per thread

__device__ __forceinline__ int row_gidx(const FrontDev& f, const int* rl, int r) {
   return r < f.n ? f.perm[r] : rl[f.ncol + (r - f.n)];
}

// scale: the factorisation's scaling in variable order or null (fkeep.F90:247-310: scaled on the way in for the
// jobs that start with the forward sweep, on the way out for those that end with the backward sweep)
__global__ void k_permute_in(const double* __restrict__ x, double* __restrict__ xp, const int* __restrict__ order,
                             const double* __restrict__ scale, int n) {
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n && order[i] >= 0) xp[order[i]] = scale ? x[i] * scale[i] : x[i];
}
__global__ void k_permute_out(double* __restrict__ x, const double* __restrict__ xp, const int* __restrict__ order,
                              const double* __restrict__ scale, int n) {
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n && order[i] >= 0) x[i] = scale ? xp[order[i]] * scale[i] : xp[order[i]];
}

// xw = [b(pivot rows); 0] + sum over children (in child order) of their update vectors
__device__ __forceinline__ void fwd_gather_front(const FrontDev* __restrict__ fr, const SlotStatic* __restrict__ ss,
                                                 const int* __restrict__ child_idx, const int* __restrict__ child_doff,
                                                 const int* __restrict__ rel, const int64_t* __restrict__ xw_off,
                                                 double* XW, double* XA, const double* __restrict__ x, int s) {
   const FrontDev& f = fr[s];
   const SlotStatic st = ss[s];
   double* xw = XW + xw_off[s];
   double* acc = XA + xw_off[s];   // indexed by the parent's ASSEMBLY order
   const int m = f.m, ne = f.nelim;
   const bool has_children = st.child_end > st.child_begin;
   // four rows per thread and round, the loads of a round issued together: a tall front (thousands of rows, one CTA) is bound
   // by the latency of these dependent loads, not by their number
   const int nt = blockDim.x;
   if (has_children) {
      for (int r = threadIdx.x; r < m; r += nt) acc[r] = 0.0;
      __syncthreads();
      for (int ci = st.child_begin; ci < st.child_end; ++ci) {
         const int cs = child_idx[ci];
         const FrontDev& fc = fr[cs];
         const int t = fc.m - fc.nelim, nd = fc.n - fc.nelim, doff = child_doff[ci];
         const double* xc = XW + xw_off[cs] + fc.nelim;
         const int* relc = rel + ss[cs].rel_off;
         for (int r0 = threadIdx.x; r0 < t; r0 += 4 * nt) {
            int d[4];
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
               const int rr = r0 + u * nt;
               d[u] = (rr < t && rr >= nd) ? relc[rr - nd] : 0;
               v[u] = rr < t ? xc[rr] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
               const int rr = r0 + u * nt;
               if (rr < t) {
                  int dd = d[u];
                  if (rr < nd) dd = f.ncol + doff + rr;
                  else if (dd >= f.ncol) dd += f.ndelay_in;
                  acc[dd] += v[u];
               }
            }
         }
         __syncthreads();
      }
   }
   // pivoting permuted the candidate rows after assembly: orig[] maps them back
   const int fn = f.n;
   for (int r0 = threadIdx.x; r0 < m; r0 += 4 * nt) {
      int pi[4], oi[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
         const int r = r0 + u * nt;
         pi[u] = r < ne ? f.perm[r] : -1;
         oi[u] = (has_children && r < fn) ? f.orig[r] : r;
      }
      double v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = pi[u] >= 0 ? x[pi[u]] : 0.0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
         const int r = r0 + u * nt;
         if (r < m) xw[r] = has_children ? v[u] + acc[oi[u]] : v[u];
      }
   }
}

__global__ void __launch_bounds__(1024) k_fwd_gather(const FrontDev* __restrict__ fr, const SlotStatic* __restrict__ ss,
                                                     const int* __restrict__ child_idx, const int* __restrict__ child_doff,
                                                     const int* __restrict__ rel, const int64_t* __restrict__ xw_off,
                                                     double* XW, double* XA, const double* __restrict__ x, int slot0, int kind) {
   pdl_launch_dependents();
   const int s = slot0 + blockIdx.x;
   const FrontDev& f = fr[s];
   if (f.state != 2 || (kind >= 0 && f.part_kind != kind)) return;
   fwd_gather_front(fr, ss, child_idx, child_doff, rel, xw_off, XW, XA, x, s);
}

// stage the lower triangle of L(k0:k0+kb, k0:k0+kb) in shared memory (column-major, leading dimension sld)
__device__ __forceinline__ void stage_triangle(double* S, int sld, const double* __restrict__ Lb, size_t ldl, int kb) {
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
   for (int c = wid; c < kb; c += nw)
      for (int r = (c & ~31) + lane; r < kb; r += 32) cp_async8(S + c * sld + r, Lb + (size_t)c * ldl + r, true);
   cp_async_commit();
}

// xs(0:kb) <- L^{-1} xs, L unit lower triangular in shared memory; 32 x 32 sub-triangles by one warp with the
// row in registers and the solution passed by shuffles, the rows below them by all threads
__device__ __forceinline__ void smem_solve_lower(const double* S, int sld, int kb, double* xs) {
   const int lane = threadIdx.x & 31;
   for (int s0 = 0; s0 < kb; s0 += 32) {
      const int sb = min(32, kb - s0);
      if (threadIdx.x < 32) {
         double lrow[32];
         const double* Sb = S + s0 * sld + s0 + lane;
#pragma unroll
         for (int j = 0; j < 32; ++j) lrow[j] = (j < lane && lane < sb) ? Sb[j * sld] : 0.0;
         double xi = lane < sb ? xs[s0 + lane] : 0.0;
#pragma unroll
         for (int j = 0; j < 32; ++j) {
            const double yj = __shfl_sync(0xffffffffu, xi, j);
            xi -= lrow[j] * yj;
         }
         if (lane < sb) xs[s0 + lane] = xi;
      }
      __syncthreads();
      const int r = s0 + sb + threadIdx.x;
      if (r < kb) {
         const double* Sr = S + s0 * sld + r;
         double a0 = 0.0, a1 = 0.0;
#pragma unroll 8
         for (int j = 0; j < sb; j += 2) { a0 += Sr[j * sld] * xs[s0 + j]; if (j + 1 < sb) a1 += Sr[(j + 1) * sld] * xs[s0 + j + 1]; }
         xs[r] -= a0 + a1;
      }
      __syncthreads();
   }
}

// xs(0:kb) <- L^{-T} xs, sub-blocks last to first; dots = 32 doubles of scratch
__device__ __forceinline__ void smem_solve_lower_t(const double* S, int sld, int kb, double* xs, double* dots) {
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   const int nsb = (kb + 31) / 32;
   for (int sbi = nsb - 1; sbi >= 0; --sbi) {
      const int s0 = sbi * 32, sb = min(32, kb - s0);
      // the rows of the block already solved: columns s0..s0+sb, rows s0+sb..kb (contiguous down a column)
      for (int j = wid; j < sb; j += (int)(blockDim.x >> 5)) {
         const double* col = S + (s0 + j) * sld;
         double acc = 0.0;
         for (int r = s0 + sb + lane; r < kb; r += 32) acc += col[r] * xs[r];
#pragma unroll
         for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
         if (lane == 0) dots[j] = acc;
      }
      __syncthreads();
      if (threadIdx.x < 32) {
         // lane i holds column i of the sub-triangle (rows below the diagonal) = row i of L^T
         double lcol[32];
         const double* Sb = S + (s0 + lane) * sld + s0;
#pragma unroll
         for (int j = 0; j < 32; ++j) lcol[j] = (j > lane && j < sb && lane < sb) ? Sb[j] : 0.0;
         double v = lane < sb ? xs[s0 + lane] - dots[lane] : 0.0;
#pragma unroll
         for (int j = 31; j >= 0; --j) {
            const double xj = __shfl_sync(0xffffffffu, v, j);
            v -= lcol[j] * xj;
         }
         if (lane < sb) xs[s0 + lane] = v;
      }
      __syncthreads();
   }
}

// ---- chain kernels: the 32 x 32 diagonal sub-triangles through explicit inverses.  A chain role stages its triangle long
// before the block it depends on is published, so it inverts the sub-triangles while it waits (warp w: sub-triangle w,
// lane c: column c of the inverse by forward substitution on e_c); the solve on the chain is then four 32 x 32
// matrix-vector products instead of 128 dependent substitution steps.  A sub-triangle whose inverse has an entry above
// INV_MAX in modulus keeps the substitution (flag 0): the inverse is only used where it is as harmless as the
// substitution itself (multipliers of well-conditioned blocks); results stay bit-reproducible run to run.
constexpr int ILD = 33;
constexpr double INV_MAX = 8.0;
__device__ __forceinline__ void smem_invert_subtriangles(const double* S, int sld, int kb, double* Inv, int* iflag) {
   const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
   if (w < 4 && 32 * w < kb) {
      const int s0 = 32 * w, sb = min(32, kb - s0);
      double z[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) z[i] = (i == lane) ? 1.0 : 0.0;
      const double* T = S + s0 * sld + s0;
#pragma unroll
      for (int k = 0; k < 31; ++k) {
         const double zk = z[k];
         const double* col = T + k * sld;
#pragma unroll
         for (int i = k + 1; i < 32; ++i) z[i] -= zk * ((i < sb && k < sb) ? col[i] : 0.0);
      }
      double mx = 0.0;
      double* dst = Inv + w * 32 * ILD + lane;
#pragma unroll
      for (int i = 0; i < 32; ++i) { dst[i * ILD] = z[i]; mx = fmax(mx, fabs(z[i])); }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      if (lane == 0) iflag[w] = (mx <= INV_MAX) ? 1 : 0;
   }
}

// smem_solve_lower with the sub-triangles through their inverses where iflag allows
__device__ __forceinline__ void smem_solve_lower_inv(const double* S, int sld, int kb, double* xs, const double* Inv, const int* iflag) {
   const int lane = threadIdx.x & 31;
   for (int s0 = 0; s0 < kb; s0 += 32) {
      const int sb = min(32, kb - s0);
      if (threadIdx.x < 32) {
         if (iflag[s0 >> 5]) {
            const double* Ir = Inv + (s0 >> 5) * 32 * ILD + lane * ILD;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
               a0 += Ir[j] * xs[s0 + j]; a1 += Ir[j + 1] * xs[s0 + j + 1]; a2 += Ir[j + 2] * xs[s0 + j + 2]; a3 += Ir[j + 3] * xs[s0 + j + 3];
            }
            __syncwarp();
            if (lane < sb) xs[s0 + lane] = (a0 + a1) + (a2 + a3);
         } else {
            double lrow[32];
            const double* Sb = S + s0 * sld + s0 + lane;
#pragma unroll
            for (int j = 0; j < 32; ++j) lrow[j] = (j < lane && lane < sb) ? Sb[j * sld] : 0.0;
            double xi = lane < sb ? xs[s0 + lane] : 0.0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
               const double yj = __shfl_sync(0xffffffffu, xi, j);
               xi -= lrow[j] * yj;
            }
            if (lane < sb) xs[s0 + lane] = xi;
         }
      }
      __syncthreads();
      const int r = s0 + sb + threadIdx.x;
      if (r < kb) {
         const double* Sr = S + s0 * sld + r;
         double a0 = 0.0, a1 = 0.0;
#pragma unroll 8
         for (int j = 0; j < sb; j += 2) { a0 += Sr[j * sld] * xs[s0 + j]; if (j + 1 < sb) a1 += Sr[(j + 1) * sld] * xs[s0 + j + 1]; }
         xs[r] -= a0 + a1;
      }
      __syncthreads();
   }
}

// smem_solve_lower_t with the sub-triangles through their (transposed) inverses where iflag allows
__device__ __forceinline__ void smem_solve_lower_t_inv(const double* S, int sld, int kb, double* xs, double* dots, const double* Inv,
                                                       const int* iflag) {
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   const int nsb = (kb + 31) / 32;
   for (int sbi = nsb - 1; sbi >= 0; --sbi) {
      const int s0 = sbi * 32, sb = min(32, kb - s0);
      for (int j = wid; j < sb; j += (int)(blockDim.x >> 5)) {
         const double* col = S + (s0 + j) * sld;
         double acc = 0.0;
         for (int r = s0 + sb + lane; r < kb; r += 32) acc += col[r] * xs[r];
#pragma unroll
         for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
         if (lane == 0) dots[j] = acc;
      }
      __syncthreads();
      if (threadIdx.x < 32) {
         if (iflag[sbi]) {
            // x_i = sum_{j >= i} inv(j, i) (b_j - dots_j): column i of the inverse, conflict-free down the rows
            const double* Ic = Inv + sbi * 32 * ILD + lane;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
               const double v0 = j < sb ? xs[s0 + j] - dots[j] : 0.0, v1 = j + 1 < sb ? xs[s0 + j + 1] - dots[j + 1] : 0.0;
               const double v2 = j + 2 < sb ? xs[s0 + j + 2] - dots[j + 2] : 0.0, v3 = j + 3 < sb ? xs[s0 + j + 3] - dots[j + 3] : 0.0;
               a0 += Ic[j * ILD] * v0; a1 += Ic[(j + 1) * ILD] * v1; a2 += Ic[(j + 2) * ILD] * v2; a3 += Ic[(j + 3) * ILD] * v3;
            }
            __syncwarp();
            if (lane < sb) xs[s0 + lane] = (a0 + a1) + (a2 + a3);
         } else {
            double lcol[32];
            const double* Sb = S + (s0 + lane) * sld + s0;
#pragma unroll
            for (int j = 0; j < 32; ++j) lcol[j] = (j > lane && j < sb && lane < sb) ? Sb[j] : 0.0;
            double v = lane < sb ? xs[s0 + lane] - dots[lane] : 0.0;
#pragma unroll
            for (int j = 31; j >= 0; --j) {
               const double xj = __shfl_sync(0xffffffffu, v, j);
               v -= lcol[j] * xj;
            }
            if (lane < sb) xs[s0 + lane] = v;
         }
      }
      __syncthreads();
   }
}

// ---- the whole 128 x 128 diagonal block through its inverse (chain kernels, use_inv = 2).  With the four sub-triangle
// inverses in Inv, the blocks of M = inv(L_bb) below the diagonal replace those of L in S, block row by block row:
//    T_j = sum_{k=j}^{i-1} L_ik M_kj,   M_ij = - M_ii T_j        (j < i; FP64 MMA, 8 x 8 tiles dealt to the warps)
// Row i reads only L's blocks of row i and finished rows of M, so the update is in place.  The solve on the chain is then
// ONE matrix-vector product (64 multiply-adds per thread, two threads per row) instead of four dependent sub-block steps.
// Used only where it is harmless: every sub-triangle inverse within INV_MAX and every entry of M below the diagonal blocks
// within INV_MAX_OFF; otherwise *flag128 = 0, the caller stages L again and keeps the sub-block path.
constexpr int TLD = 33;
constexpr double INV_MAX_OFF = 8.0;
__device__ __forceinline__ void smem_invert_block128(double* S, int sld, const double* Inv, const int* iflag, double* T, int* flag128) {
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, g = lane >> 2, t4 = lane & 3;
   if (threadIdx.x == 0) *flag128 = (iflag[0] && iflag[1] && iflag[2] && iflag[3]) ? 1 : 0;
   __syncthreads();
   if (!*flag128) return;
   double mx = 0.0;
   for (int i = 1; i < 4; ++i) {
      for (int tile = wid; tile < 16 * i; tile += 8) {
         const int j = tile >> 4, tr = (tile >> 2) & 3, tc = tile & 3;
         double c0 = 0.0, c1 = 0.0;
         for (int k = j; k < i; ++k) {
            const double* Lik = S + (size_t)(32 * k) * sld + 32 * i + 8 * tr + g;                   // (r, c) at c * sld + r
            const double* Mkj = (k == j) ? Inv + j * 32 * ILD + 8 * tc + g                          // (r, c) at r * ILD + c
                                         : S + (size_t)(32 * j + 8 * tc + g) * sld + 32 * k;
            const int bstride = (k == j) ? ILD : 1;
#pragma unroll
            for (int ks = 0; ks < 32; ks += 4) dmma(c0, c1, Lik[(size_t)(ks + t4) * sld], Mkj[(ks + t4) * bstride]);
         }
         double* Tj = T + j * 32 * TLD + (8 * tr + g) * TLD + 8 * tc + 2 * t4;
         Tj[0] = c0;
         Tj[1] = c1;
      }
      __syncthreads();
      for (int tile = wid; tile < 16 * i; tile += 8) {
         const int j = tile >> 4, tr = (tile >> 2) & 3, tc = tile & 3;
         double c0 = 0.0, c1 = 0.0;
         const double* Mii = Inv + i * 32 * ILD + (8 * tr + g) * ILD;
         const double* Tj = T + j * 32 * TLD + 8 * tc + g;
#pragma unroll
         for (int ks = 0; ks < 32; ks += 4) dmma(c0, c1, Mii[ks + t4], Tj[(ks + t4) * TLD]);
         double* Mij = S + (size_t)(32 * j + 8 * tc + 2 * t4) * sld + 32 * i + 8 * tr + g;
         Mij[0] = -c0;
         Mij[sld] = -c1;
         mx = fmax(mx, fmax(fabs(c0), fabs(c1)));
         if (c0 != c0 || c1 != c1) mx = INFINITY;
      }
      __syncthreads();
   }
#pragma unroll
   for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
   if (lane == 0 && !(mx <= INV_MAX_OFF)) *flag128 = 0;
   __syncthreads();
}

// rows / columns kb..KB-1 of a short last block: zeros, so that the padded part of the inverse is the identity
__device__ __forceinline__ void smem_pad_block128(double* S, int sld, int kb, double* Inv, int* iflag) {
   for (int e = threadIdx.x; e < KB * KB; e += blockDim.x) {
      const int c = e >> 7, r = e & (KB - 1);
      if (r >= kb || c >= kb) S[c * sld + r] = 0.0;
   }
   // sub-triangles wholly behind kb are not visited by smem_invert_subtriangles
   for (int w = (kb + 31) / 32; w < 4; ++w) {
      for (int e = threadIdx.x; e < 32 * 32; e += blockDim.x) Inv[w * 32 * ILD + (e >> 5) * ILD + (e & 31)] = ((e >> 5) == (e & 31)) ? 1.0 : 0.0;
      if (threadIdx.x == 0) iflag[w] = 1;
   }
}

// xs <- M xs (forward) with M = inv(L_bb) in S (block lower triangular, zeros above the diagonal of the diagonal blocks).
// Thread (row r, half h) takes columns 16 h .. 16 h + 15 of every 32-block; part: 2 x KB doubles of scratch.
__device__ __forceinline__ void smem_apply_inverse(const double* S, int sld, double* xs, double* part) {
   const int r = threadIdx.x & (KB - 1), h = threadIdx.x >> 7, ib = r >> 5;
   double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
   for (int jb = 0; jb <= ib; ++jb) {
      const double* Sr = S + (size_t)(32 * jb + 16 * h) * sld + r;
      const double* xv = xs + 32 * jb + 16 * h;
#pragma unroll
      for (int c = 0; c < 16; c += 4) {
         a0 += Sr[(size_t)c * sld] * xv[c]; a1 += Sr[(size_t)(c + 1) * sld] * xv[c + 1];
         a2 += Sr[(size_t)(c + 2) * sld] * xv[c + 2]; a3 += Sr[(size_t)(c + 3) * sld] * xv[c + 3];
      }
   }
   part[h * KB + r] = (a0 + a1) + (a2 + a3);
   __syncthreads();
   if (threadIdx.x < KB) xs[threadIdx.x] = part[threadIdx.x] + part[KB + threadIdx.x];
}

// xs <- M^T xs (backward): thread (column c, half h) takes rows 16 h .. 16 h + 15 of every 32-block at and below its own
__device__ __forceinline__ void smem_apply_inverse_t(const double* S, int sld, double* xs, double* part) {
   const int c = threadIdx.x & (KB - 1), h = threadIdx.x >> 7, cb = c >> 5;
   double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
   for (int ib = cb; ib < 4; ++ib) {
      const double* Sc = S + (size_t)c * sld + 32 * ib + 16 * h;
      const double* xv = xs + 32 * ib + 16 * h;
#pragma unroll
      for (int rr = 0; rr < 16; rr += 4) { a0 += Sc[rr] * xv[rr]; a1 += Sc[rr + 1] * xv[rr + 1]; a2 += Sc[rr + 2] * xv[rr + 2]; a3 += Sc[rr + 3] * xv[rr + 3]; }
   }
   part[h * KB + c] = (a0 + a1) + (a2 + a3);
   __syncthreads();
   if (threadIdx.x < KB) xs[threadIdx.x] = part[threadIdx.x] + part[KB + threadIdx.x];
}

// The inverses of the 128 x 128 diagonal blocks of the chain fronts, once per factorisation (first solve after it): one CTA
// per block.  Minv: KB x KB doubles per block, column-major, block lower triangular with explicit zeros above the diagonal
// of the 32 x 32 diagonal blocks (what stage_triangle copies); mflag = 0 where a bound failed (the chain keeps L there).
__global__ void __launch_bounds__(256, 1) k_chain_inverses(const FrontDev* __restrict__ fr, const int2* __restrict__ blist, double* Minv,
                                                           int* mflag) {
   constexpr int SLD = KB + 1;
   const int s = blist[blockIdx.x].x, c = blist[blockIdx.x].y;
   const FrontDev& f = fr[s];
   __shared__ int iflag[4];
   __shared__ int flag128;
   const int kb = f.state == 2 ? min(KB, f.nelim - c * KB) : 0;
   if (kb <= 0) { if (threadIdx.x == 0) mflag[blockIdx.x] = 0; return; }
   extern __shared__ double sm[];
   double* S = sm;
   double* Inv = S + KB * SLD;
   double* T = Inv + 4 * 32 * ILD;
   const size_t ldl = f.ldl;
   stage_triangle(S, SLD, f.L + (size_t)(c * KB) * ldl + c * KB, ldl, kb);
   cp_async_wait<0>();
   __syncthreads();
   if (kb < KB) { smem_pad_block128(S, SLD, kb, Inv, iflag); __syncthreads(); }
   smem_invert_subtriangles(S, SLD, kb, Inv, iflag);
   __syncthreads();
   smem_invert_block128(S, SLD, Inv, iflag, T, &flag128);
   if (threadIdx.x == 0) mflag[blockIdx.x] = flag128;
   if (!flag128) return;
   double* out = Minv + (size_t)blockIdx.x * KB * KB;
   for (int e = threadIdx.x; e < KB * KB; e += blockDim.x) {
      const int col = e >> 7, r = e & (KB - 1), cb = col >> 5, rb = r >> 5;
      if (rb < cb) continue;
      out[e] = rb == cb ? Inv[cb * 32 * ILD + (r & 31) * ILD + (col & 31)] : S[col * SLD + r];
   }
}
static inline size_t chain_inv_smem() { return (size_t)(KB * (KB + 1) + 4 * 32 * ILD + 3 * 32 * TLD) * 8; }

// Backward sweep, chain fronts: the rows below the pivots (the bulk of L's bytes in a front that is not a root) leave the
// chain: CTA (front, column block c, chunk q of TAIL_ROWS rows) leaves sum_r L(r, col) x(r) over its rows for the 128
// columns of the block in TAIL[tail_off + (q nb + c) KB + col]; the chain role subtracts the chunks in order.
constexpr int TAIL_ROWS = 512;
__global__ void __launch_bounds__(256) k_bwd_tail(const FrontDev* __restrict__ fr, const int4* __restrict__ troles, const int64_t* __restrict__ xw_off,
                                                  const double* __restrict__ XW, double* TAIL, const int64_t* __restrict__ tail_off, int kind) {
   pdl_launch_dependents();
   const int4 tr = troles[blockIdx.x];
   const int s = tr.x, c = tr.y, q = tr.z;
   const FrontDev& f = fr[s];
   if (f.state != 2 || (kind >= 0 && f.part_kind != kind)) return;
   const int ne = f.nelim, m = f.m, nb = (ne + KB - 1) / KB;
   const int cbase = c * KB, kb = min(KB, ne - cbase);
   const int r0 = ne + q * TAIL_ROWS, r1 = min(m, r0 + TAIL_ROWS);
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   const size_t ldl = f.ldl;
   const double* __restrict__ xw = XW + xw_off[s];
   const double* __restrict__ Lc = f.L + (size_t)(cbase + 16 * wid) * ldl;
   double acc[16];
#pragma unroll
   for (int j = 0; j < 16; ++j) acc[j] = 0.0;
   for (int r = r0 + lane; r < r1; r += 64) {
      const bool v1 = r + 32 < r1;
      const double xa = xw[r], xb = v1 ? xw[r + 32] : 0.0;
      double la[16], lb[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
         const bool cv = 16 * wid + j < kb;
         la[j] = cv ? Lc[(size_t)j * ldl + r] : 0.0;
         lb[j] = (cv && v1) ? Lc[(size_t)j * ldl + r + 32] : 0.0;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] += la[j] * xa + lb[j] * xb;
   }
   double* out = TAIL + tail_off[s] + ((size_t)q * nb + c) * KB + 16 * wid;
#pragma unroll
   for (int j = 0; j < 16; ++j) {
      double a = acc[j];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
      if (lane == 0) out[j] = a;
   }
}

// Forward step for column block `step` of every front of the launch.  KBC = 32 / 64 / 128 bounds the block
// width of the launch; a CTA covers RC = SE / KBC rows below the block, NG = 256 / RC threads per row.
// The solved block goes to XY (the front's slice of the gather accumulators, dead by now), never back into
// xw: a CTA scheduled late must still find the unsolved right-hand side there.
// FUSED (levels whose fronts have one block and one CTA each): the CTA also gathers the children's update vectors
// first and scatters the solved pivots into x afterwards, i.e. the whole level is this one launch.
struct GatherArgs {
   const SlotStatic* ss; const int* child_idx; const int* child_doff; const int* rel; const int* rlist;
   double* x; int posdef_scale, do_diag;
};
// T: threads per CTA (256; 64 for the FUSED levels of very small fronts, where six times as many fronts fit on an SM;
// 192 with KBC = 96 for the FUSED levels of 65..96 pivots: two CTAs per SM where the 128-column instantiation has one)
template <int KBC, bool FUSED, int T = 256>
__global__ void __launch_bounds__(T, T == 64 ? 6 : (KBC <= 96 ? 2 : 1)) k_fwd_step(const FrontDev* __restrict__ fr, const int* __restrict__ list,
                                                  const int64_t* __restrict__ xw_off, double* XW, double* XY, int step, int kind, GatherArgs ga) {
   constexpr int RC = T * 32 / KBC, NG = T / RC, SLD = KBC + 1;
   pdl_launch_dependents();
   const int s = list[blockIdx.x];
   const FrontDev& f = fr[s];
   if (f.state != 2 || (kind >= 0 && f.part_kind != kind)) return;
   const int ne = f.nelim, k0 = step * KB;
   if (FUSED && ne == 0) { pdl_wait(); fwd_gather_front(fr, ga.ss, ga.child_idx, ga.child_doff, ga.rel, xw_off, XW, XY, ga.x, s); return; }
   if (k0 >= ne) return;
   const int kb = min(KB, ne - k0), m = f.m;
   const int nchunk = (m - (k0 + kb) + RC - 1) / RC;      // chunks of RC rows below the block; CTA y takes y, y + gridDim.y, ...
   if (blockIdx.y > 0 && (int)blockIdx.y >= nchunk) return;
   extern __shared__ double sm[];
   double* S = sm;                    // KBC x SLD
   double* xs = sm + KBC * SLD;       // KBC
   double* red = xs + KBC;            // T
   double* xw = XW + xw_off[s];
   const size_t ldl = f.ldl;
   const double* __restrict__ Lb = f.L + (size_t)k0 * ldl + k0;
   stage_triangle(S, SLD, Lb, ldl, kb);
   // this CTA's part of the rows below: thread (row rl, column group g) holds 32 entries of L
   const int rl = threadIdx.x & (RC - 1), g = threadIdx.x / RC;
   const double* __restrict__ Lg = f.L + (size_t)(k0 + g * (KBC / NG)) * ldl;
   double lreg[KBC / NG];
   int r = k0 + kb + blockIdx.y * RC + rl;
#pragma unroll
   for (int j = 0; j < KBC / NG; ++j) lreg[j] = (r < m && g * (KBC / NG) + j < kb) ? Lg[(size_t)j * ldl + r] : 0.0;
   pdl_wait();       // the level before (a launch with programmatic serialisation starts while it is still running)
   if (FUSED) {      // the triangle and this thread's entries of L are on their way meanwhile
      fwd_gather_front(fr, ga.ss, ga.child_idx, ga.child_doff, ga.rel, xw_off, XW, XY, ga.x, s);
      __syncthreads();
   }
   for (int i = threadIdx.x; i < KBC; i += T) xs[i] = i < kb ? xw[k0 + i] : 0.0;
   cp_async_wait<0>();
   __syncthreads();
   smem_solve_lower(S, SLD, kb, xs);
   if (FUSED) {
      for (int i = threadIdx.x; i < kb; i += T) ga.x[f.perm[i]] = ga.posdef_scale ? xs[i] * sqrt(f.D[2 * i]) : xs[i];
   } else if (blockIdx.y == 0) {
      double* xy = XY + xw_off[s];
      for (int i = threadIdx.x; i < kb; i += T) xy[k0 + i] = xs[i];
   }
   double xg[KBC / NG];
#pragma unroll
   for (int j = 0; j < KBC / NG; ++j) xg[j] = xs[g * (KBC / NG) + j];
   for (int ch = blockIdx.y; ch < nchunk; ch += gridDim.y) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < KBC / NG; ++j) acc += lreg[j] * xg[j];
      const int rcur = r;
      r += gridDim.y * RC;
      if (ch + (int)gridDim.y < nchunk) {
#pragma unroll
         for (int j = 0; j < KBC / NG; ++j) lreg[j] = (r < m && g * (KBC / NG) + j < kb) ? Lg[(size_t)j * ldl + r] : 0.0;
      }
      if (NG > 1) {
         __syncthreads();
         red[threadIdx.x] = acc;
         __syncthreads();
         if (g == 0) {
#pragma unroll
            for (int q = 1; q < NG; ++q) acc += red[q * RC + rl];
         }
      }
      if (g == 0 && rcur < m) xw[rcur] -= acc;
   }
}

// ------------------------------------------------------------------------------------------
// Chain kernels: ONE launch per level and direction for levels that hold fronts of more than KB eliminated
// columns.  A front's work is cut into roles, dealt to CTAs through a ticket counter in an order in which a
// role only ever waits for roles with smaller tickets (so a waiting CTA can never keep its producer off the
// GPU):
//   forward : role c < nb owns block row c of the front: it streams L(c, b) for b < c as the solved blocks y_b
//             are published, then solves its own triangle from shared memory and publishes y_c; roles >= nb take
//             128 rows each of the part below the pivots;
//   backward: role c owns block column c: dot products with the rows below the pivots (known), then with the
//             blocks b > c as their solutions appear, last first; then the transposed triangle.
// The entries of L needed next are always requested BEFORE the block they multiply is polled, so that the chain
// between two consecutive blocks is: values arrive -> 64 FMAs per thread from registers -> reduction -> triangle.
// A lost producer cannot hang the GPU: a poll gives up after ~10 s and raises *err.
constexpr int CH = 128;      // rows of a role = block width

// Solved blocks travel between the CTAs of a launch through CY, a vector the host fills with the bit pattern
// CY_EMPTY (all ones, a NaN no computation produces: the publisher canonicalises) before each sweep.  The
// consumer polls the values themselves: 8-byte stores are atomic, so no flag and no fence are needed, and the
// chain saves one L2 round trip per block.
constexpr unsigned long long CY_EMPTY = 0xffffffffffffffffULL;
__device__ __forceinline__ double cy_load(const double* p) {
   return __longlong_as_double(*((volatile const long long*)p));
}
__device__ __forceinline__ bool cy_empty(double v) { return (unsigned long long)__double_as_longlong(v) == CY_EMPTY; }
__device__ __forceinline__ void cy_store(double* p, double v) {
   if (cy_empty(v)) v = __longlong_as_double(0x7ff8000000000000LL);
   *((volatile double*)p) = v;
}
// spin until buf[0..N) (values this thread needs, at stride `stride` from p) are all published
template <int N>
__device__ __forceinline__ void cy_wait(double (&buf)[N], const double* p, int stride, int nvalid, int* err) {
   long long t0 = 0;
   for (;;) {
      bool ok = true;
#pragma unroll
      for (int j = 0; j < N; ++j) { buf[j] = j * stride < nvalid ? cy_load(p + j * stride) : 0.0; ok = ok && !cy_empty(buf[j]); }
      if (ok) break;
      if (t0 == 0) { if (*((volatile int*)err)) break; t0 = clock64(); }
      else if (clock64() - t0 > 20000000000LL) { *err = 1; break; }
   }
}

// One warp polls 32 consecutive values (lane l: p[l]) with a single coalesced load per round and leaves them in shared
// memory: the roles of a launch all wait for the same block, and 256 threads polling 32 values each from every CTA made
// the few cache lines of CY the busiest spot of L2.
__device__ __forceinline__ void cy_wait_warp(double* dst, const double* p, int nvalid, int* err) {
   const int lane = threadIdx.x & 31;
   long long t0 = 0;
   double v;
   for (;;) {
      v = lane < nvalid ? cy_load(p + lane) : 0.0;
      if (__all_sync(0xffffffffu, !cy_empty(v))) break;
      if (t0 == 0) { if (*((volatile int*)err)) break; t0 = clock64(); }
      else if (clock64() - t0 > 20000000000LL) { *err = 1; break; }
   }
   dst[lane] = v;
}

__device__ __forceinline__ unsigned long long chain_now() {
   unsigned long long t;
   asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
   return t;
}
// dbg (B200LDLT_SOLVE_TRACE), 16 slots per role: {start, block staged / sub-inverses done, last dependency arrived, published;
// backward: reduced, staged, applied, whether the block inverse was used, loop exit of each warp} in ns
#define CHAIN_TICK(k) do { if (dbg && threadIdx.x == 0) dbg[16 * (size_t)ticket + (k)] = chain_now(); } while (0)

__global__ void __launch_bounds__(256, 1) k_fwd_chain(const FrontDev* __restrict__ fr, const int2* __restrict__ roles, int nroles,
                                                      unsigned long long* counter, const int64_t* __restrict__ xw_off, double* XW,
                                                      double* CY, double* x, int posdef_scale, int kind, int* err, int use_inv,
                                                      const double* __restrict__ Minv, const int* __restrict__ minv_off,
                                                      const int* __restrict__ mflag, unsigned long long* dbg) {
   constexpr int SLD = KB + 1;
   __shared__ int2 role;
   __shared__ int iflag[4];
   __shared__ int ticket;
   pdl_launch_dependents();
   if (threadIdx.x == 0) { ticket = (int)(atomicAdd(counter, 1ULL) % (unsigned long long)nroles); role = roles[ticket]; }
   __syncthreads();
   const int s = role.x, idx = role.y;
   const FrontDev& f = fr[s];
   if (f.state != 2 || (kind >= 0 && f.part_kind != kind)) return;
   const int ne = f.nelim, m = f.m, nb = (ne + KB - 1) / KB;
   const bool owner = idx < nb;
   const int rbase = owner ? idx * KB : ne + (idx - nb) * CH;
   const int rend = min(owner ? ne : m, rbase + CH);
   const int nblk = owner ? idx : nb;
   extern __shared__ double sm[];
   double* S = sm;                    // KB x SLD
   double* xs = sm + KB * SLD;        // KB
   double* red = xs + KB;             // 2 x CH
   double* Inv = red + 2 * CH + 32;   // 4 x 32 x ILD
   double* ybuf = Inv + 4 * 32 * ILD + KB;  // 2 x KB: the published block, double buffered
   double* part = ybuf + 2 * KB;            // 2 x KB
   double* xw = XW + xw_off[s];
   const double* cy = CY + xw_off[s];
   const size_t ldl = f.ldl;
   // the block's precomputed inverse (k_chain_inverses) where it exists and passed its bounds, else L itself
   const int mq = (owner && use_inv > 1 && Minv != nullptr && minv_off[s] >= 0) ? minv_off[s] + idx : -1;
   const bool full = mq >= 0 && mflag[mq] != 0;
   if (owner) {
      if (full) stage_triangle(S, SLD, Minv + (size_t)mq * KB * KB, KB, KB);
      else stage_triangle(S, SLD, f.L + (size_t)rbase * ldl + rbase, ldl, rend - rbase);
   }
   const int i = threadIdx.x & (CH - 1), g = threadIdx.x >> 7;      // row of the role, column half of a half block
   const bool rv = rbase + i < rend;
   pdl_wait();       // the gather of this level
   const double x0 = (g == 0 && rv) ? xw[rbase + i] : 0.0;
   const double* __restrict__ Lr = f.L + rbase + i;
   double la[32], lb[32], yv[32], acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
   // half h of block b: columns KB b + 64 h + 32 g + j
   auto load_half = [&](double (&buf)[32], int b, int h) {
      const int c0 = KB * b + 64 * h + 32 * g;
#pragma unroll
      for (int j = 0; j < 32; ++j) buf[j] = (rv && c0 + j < ne) ? Lr[(size_t)(c0 + j) * ldl] : 0.0;
   };
   auto use_half = [&](const double (&buf)[32], int b, int h) {
      const int c0 = KB * b + 64 * h + 32 * g;
      cy_wait(yv, cy + c0, 1, ne - c0, err);
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
         acc0 += buf[j] * yv[j]; acc1 += buf[j + 1] * yv[j + 1]; acc2 += buf[j + 2] * yv[j + 2]; acc3 += buf[j + 3] * yv[j + 3];
      }
   };
   CHAIN_TICK(0);
   if (nblk > 0) load_half(la, 0, 0);
   if (owner && use_inv && !full) {
      // the sub-triangles' inverses while the blocks this role depends on are being solved elsewhere
      cp_async_wait<0>();
      __syncthreads();
      smem_invert_subtriangles(S, SLD, rend - rbase, Inv, iflag);
   }
   CHAIN_TICK(1);
   if (use_inv > 1) {
      const int wid = threadIdx.x >> 5;
      for (int b = 0; b < nblk; ++b) {
         load_half(lb, b, 1);
         if (wid < 4) cy_wait_warp(ybuf + (b & 1) * KB + 32 * wid, cy + KB * b + 32 * wid, ne - (KB * b + 32 * wid), err);
         __syncthreads();
         const double* yb = ybuf + (b & 1) * KB + 32 * g;
#pragma unroll
         for (int j = 0; j < 32; j += 4) {
            acc0 += la[j] * yb[j]; acc1 += la[j + 1] * yb[j + 1]; acc2 += la[j + 2] * yb[j + 2]; acc3 += la[j + 3] * yb[j + 3];
         }
         if (b + 1 < nblk) load_half(la, b + 1, 0);
#pragma unroll
         for (int j = 0; j < 32; j += 4) {
            acc0 += lb[j] * yb[64 + j]; acc1 += lb[j + 1] * yb[64 + j + 1]; acc2 += lb[j + 2] * yb[64 + j + 2]; acc3 += lb[j + 3] * yb[64 + j + 3];
         }
      }
   } else {
      for (int b = 0; b < nblk; ++b) {
         load_half(lb, b, 1);
         use_half(la, b, 0);
         if (b + 1 < nblk) load_half(la, b + 1, 0);
         use_half(lb, b, 1);
      }
   }
   CHAIN_TICK(2);
   red[g * CH + i] = (acc0 + acc1) + (acc2 + acc3);
   __syncthreads();
   if (!owner) {
      if (g == 0 && rv) xw[rbase + i] = x0 - (red[i] + red[CH + i]);
      return;
   }
   const int kb = rend - rbase;
   if (g == 0) xs[i] = rv ? x0 - (red[i] + red[CH + i]) : 0.0;
   cp_async_wait<0>();
   __syncthreads();
   if (full) smem_apply_inverse(S, SLD, xs, part);
   else if (use_inv) smem_solve_lower_inv(S, SLD, kb, xs, Inv, iflag);
   else smem_solve_lower(S, SLD, kb, xs);
   if (threadIdx.x < kb) {
      const double v = xs[threadIdx.x];
      cy_store(CY + xw_off[s] + rbase + threadIdx.x, v);
      // the solved pivots go straight into x (L_chol = L sqrt(D) for posdef; D is stored inverted)
      x[f.perm[rbase + threadIdx.x]] = posdef_scale ? v * sqrt(f.D[2 * (rbase + threadIdx.x)]) : v;
   }
   CHAIN_TICK(3);
}

__global__ void __launch_bounds__(256, 1) k_bwd_chain(const FrontDev* __restrict__ fr, const int2* __restrict__ roles, int nroles,
                                                      unsigned long long* counter, const int64_t* __restrict__ xw_off, double* XW,
                                                      double* CY, double* x, int kind, int* err, int use_inv,
                                                      const double* __restrict__ Minv, const int* __restrict__ minv_off,
                                                      const int* __restrict__ mflag, const double* __restrict__ TAIL,
                                                      const int64_t* __restrict__ tail_off, unsigned long long* dbg) {
   constexpr int SLD = KB + 1;
   __shared__ int2 role;
   __shared__ int iflag[4];
   __shared__ int ticket;
   pdl_launch_dependents();
   if (threadIdx.x == 0) { ticket = (int)(atomicAdd(counter, 1ULL) % (unsigned long long)nroles); role = roles[ticket]; }
   __syncthreads();
   const int s = role.x, c = role.y;
   const FrontDev& f = fr[s];
   if (f.state != 2 || (kind >= 0 && f.part_kind != kind)) return;
   const int ne = f.nelim, m = f.m, nb = (ne + KB - 1) / KB;
   const int cbase = c * KB, kb = min(KB, ne - cbase);
   const int ntail = (TAIL != nullptr && tail_off[s] >= 0) ? 0 : (m - ne + CH - 1) / CH;
   extern __shared__ double sm[];
   double* S = sm;                    // KB x SLD
   double* xs = sm + KB * SLD;        // KB
   double* dots = xs + KB;            // 32 + KB
   double* Inv = dots + 32 + KB + 2 * KB;   // 4 x 32 x ILD (behind the forward kernel's layout: one size for both)
   double* part = Inv + 4 * 32 * ILD;       // 2 x KB
   double* xw = XW + xw_off[s];
   const double* cy = CY + xw_off[s];
   const size_t ldl = f.ldl;
   const int mq = (use_inv > 1 && Minv != nullptr && minv_off[s] >= 0) ? minv_off[s] + c : -1;
   const bool full = mq >= 0 && mflag[mq] != 0;
   if (full) stage_triangle(S, SLD, Minv + (size_t)mq * KB * KB, KB, KB);
   else stage_triangle(S, SLD, f.L + (size_t)cbase * ldl + cbase, ldl, kb);
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   // the rows below the pivots: partial sums of k_bwd_tail (chunks of TAIL_ROWS rows, added in chunk order), else by this role
   const bool tailed = TAIL != nullptr && tail_off[s] >= 0;
   pdl_wait();       // the gather (and the tail sums) of this level
   double x0 = threadIdx.x < kb ? xw[cbase + threadIdx.x] : 0.0;
   if (tailed && threadIdx.x < kb) {
      const int ntc = (m - ne + TAIL_ROWS - 1) / TAIL_ROWS;
      const double* tp = TAIL + tail_off[s] + cbase + threadIdx.x;
      for (int q = 0; q < ntc; ++q) x0 -= tp[(size_t)q * nb * KB];
   }
   // warp wid takes columns cbase + 16 wid + q; a half chunk is 64 rows: two loads per lane and column
   const double* __restrict__ Lc = f.L + (size_t)(cbase + 16 * wid) * ldl + lane;
   double va[16][2], vb[16][2], acc[16];
#pragma unroll
   for (int q = 0; q < 16; ++q) acc[q] = 0.0;
   // half chunk hh: chunks 0..ntail-1 are the rows below the pivots, then the blocks nb-1, nb-2, ..., c+1
   auto rows_of = [&](int hh, int& r0, int& r1, int& b) {
      const int q = hh >> 1;
      int rb, rl;
      if (q < ntail) { rb = ne + CH * q; rl = m; b = -1; }
      else { b = nb - 1 - (q - ntail); rb = KB * b; rl = ne; }
      r0 = rb + 64 * (hh & 1);
      r1 = min(rl, rb + CH);
   };
   auto load_half = [&](double (&buf)[16][2], int hh) {
      int r0, r1, b;
      rows_of(hh, r0, r1, b);
#pragma unroll
      for (int q = 0; q < 16; ++q)
#pragma unroll
         for (int u = 0; u < 2; ++u)
            buf[q][u] = (16 * wid + q < kb && r0 + lane + 32 * u < r1) ? Lc[(size_t)q * ldl + r0 + 32 * u] : 0.0;
   };
   auto use_half = [&](const double (&buf)[16][2], int hh) {
      int r0, r1, b;
      rows_of(hh, r0, r1, b);
      double xv[2];
      if (b >= 0) cy_wait(xv, cy + r0 + lane, 32, r1 - (r0 + lane), err);          // a block solved by this launch
      else {
#pragma unroll
         for (int u = 0; u < 2; ++u) xv[u] = (r0 + lane + 32 * u < r1) ? xw[r0 + lane + 32 * u] : 0.0;
      }
#pragma unroll
      for (int q = 0; q < 16; ++q) acc[q] += buf[q][0] * xv[0] + buf[q][1] * xv[1];
   };
   const int H = 2 * (ntail + nb - 1 - c);
   CHAIN_TICK(0);
   if (H > 0) load_half(va, 0);
   if (use_inv && !full) {
      cp_async_wait<0>();
      __syncthreads();
      smem_invert_subtriangles(S, SLD, kb, Inv, iflag);
   }
   CHAIN_TICK(1);
   if (use_inv > 1) {
      // one warp per 32 rows fetches (or waits for) the chunk's x into shared memory, one barrier per chunk of 128 rows
      double* ybuf = part + 2 * KB;         // 2 x KB, double buffered
      for (int k = 0; 2 * k < H; ++k) {
         int r0, r1, b;
         rows_of(2 * k, r0, r1, b);
         load_half(vb, 2 * k + 1);
         if (wid < 4) {
            const int r = r0 + 32 * wid;
            if (b >= 0) cy_wait_warp(ybuf + (k & 1) * KB + 32 * wid, cy + r, r1 - r, err);
            else ybuf[(k & 1) * KB + 32 * wid + lane] = (r + lane < r1) ? xw[r + lane] : 0.0;
         }
         __syncthreads();
         const double* yb = ybuf + (k & 1) * KB + lane;
         {
            const double y0 = yb[0], y1 = yb[32];
#pragma unroll
            for (int q = 0; q < 16; ++q) acc[q] += va[q][0] * y0 + va[q][1] * y1;
         }
         if (2 * k + 2 < H) load_half(va, 2 * k + 2);
         {
            const double y0 = yb[64], y1 = yb[96];
#pragma unroll
            for (int q = 0; q < 16; ++q) acc[q] += vb[q][0] * y0 + vb[q][1] * y1;
         }
      }
   } else {
      for (int hh = 0; hh < H; hh += 2) {
         load_half(vb, hh + 1);
         use_half(va, hh);
         if (hh + 2 < H) load_half(va, hh + 2);
         use_half(vb, hh + 1);
      }
   }
   CHAIN_TICK(2);
   if (dbg && lane == 0) dbg[16 * (size_t)ticket + 8 + wid] = chain_now();
   // transposing reduction of the 16 sums over the 32 lanes: each round halves the sums a lane carries (16 shuffles in all
   // instead of 80); lanes 2 q and 2 q + 1 end with column q
#define B200_FOLD(OFF, N)                                                                         \
   {                                                                                              \
      const bool up = (lane & (OFF)) != 0;                                                        \
      _Pragma("unroll") for (int i = 0; i < (N); ++i) {                                           \
         const double keep = up ? acc[i + (N)] : acc[i], send = up ? acc[i] : acc[i + (N)];       \
         acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, (OFF));                               \
      }                                                                                           \
   }
   B200_FOLD(16, 8) B200_FOLD(8, 4) B200_FOLD(4, 2) B200_FOLD(2, 1)
#undef B200_FOLD
   acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 1);
   if ((lane & 1) == 0) dots[32 + 16 * wid + (lane >> 1)] = acc[0];
   __syncthreads();
   CHAIN_TICK(4);
   if (threadIdx.x < KB) xs[threadIdx.x] = threadIdx.x < kb ? x0 - dots[32 + threadIdx.x] : 0.0;
   cp_async_wait<0>();
   __syncthreads();
   CHAIN_TICK(5);
   if (full) smem_apply_inverse_t(S, SLD, xs, part);
   else if (use_inv) smem_solve_lower_t_inv(S, SLD, kb, xs, dots, Inv, iflag);
   else smem_solve_lower_t(S, SLD, kb, xs, dots);
   CHAIN_TICK(6);
   if (threadIdx.x == 0 && dbg) dbg[16 * (size_t)ticket + 7] = (unsigned long long)(full ? 1 : 0);
   if (threadIdx.x < kb) {
      cy_store(CY + xw_off[s] + cbase + threadIdx.x, xs[threadIdx.x]);
      x[f.perm[cbase + threadIdx.x]] = xs[threadIdx.x];
   }
   CHAIN_TICK(3);
}

// XY: the solved blocks (k_fwd_step)
__global__ void __launch_bounds__(256) k_fwd_store(const FrontDev* __restrict__ fr, const int64_t* __restrict__ xw_off,
                                                   const double* __restrict__ XW, double* x, int slot0, int posdef_scale, int kind) {
   const int s = slot0 + blockIdx.x;
   const FrontDev& f = fr[s];
   if (f.state != 2 || (kind >= 0 && f.part_kind != kind)) return;
   const double* xw = XW + xw_off[s];
   for (int k = threadIdx.x; k < f.nelim; k += blockDim.x) {
      double v = xw[k];
      if (posdef_scale) v *= sqrt(f.D[2 * k]);   // L_chol = L sqrt(D); D is stored inverted
      x[f.perm[k]] = v;
   }
}

__global__ void k_solve_diag(const FrontDev* __restrict__ fr, double* x, int nslots) {
   const int lane = threadIdx.x & 31;
   const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
   if (s >= nslots) return;
   const FrontDev& f = fr[s];
   if (f.state != 2) return;
   for (int i = lane; i < f.nelim; i += 32) {
      double a11 = f.D[2 * i];
      if (isinf(a11) && i > 0) continue;
      if (i + 1 < f.nelim && isinf(f.D[2 * i + 2])) {
         double d21 = f.D[2 * i + 1], d22 = f.D[2 * i + 3];
         double x1 = x[f.perm[i]], x2 = x[f.perm[i + 1]];
         x[f.perm[i]] = a11 * x1 + d21 * x2;
         x[f.perm[i + 1]] = d21 * x1 + d22 * x2;
      } else {
         x[f.perm[i]] *= a11;
      }
   }
}

// backward: xw(r >= ne) = x(ancestors), xw(r < ne) = D^{-1} y (or sqrt scaling for posdef)
__device__ __forceinline__ void bwd_gather_front(const FrontDev& f, const int* __restrict__ rl, double* xw, const double* __restrict__ x,
                                                 int do_diag, int posdef_scale, int rfirst, int rstride) {
   const int m = f.m, ne = f.nelim;
   // the rows below the pivots four at a time (index loads, then value loads), the pivots one by one
   for (int r0 = ne + rfirst; r0 < m; r0 += 4 * rstride) {
      int gi[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int r = r0 + u * rstride; gi[u] = r < m ? row_gidx(f, rl, r) : -1; }
      double v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = gi[u] >= 0 ? x[gi[u]] : 0.0;
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int r = r0 + u * rstride; if (r < m) xw[r] = v[u]; }
   }
   for (int r = rfirst; r < ne; r += rstride) {
      double v;
      {
         v = x[f.perm[r]];
         if (posdef_scale) v *= sqrt(f.D[2 * r]);
         else if (do_diag) {
            double a11 = f.D[2 * r];
            if (isinf(a11) && r > 0) v = f.D[2 * r - 1] * x[f.perm[r - 1]] + f.D[2 * r + 1] * v;
            else if (r + 1 < ne && isinf(f.D[2 * r + 2])) v = a11 * v + f.D[2 * r + 1] * x[f.perm[r + 1]];
            else v = a11 * v;
         }
      }
      xw[r] = v;
   }
}
// grid (fronts of the level, row chunks)
__global__ void __launch_bounds__(256) k_bwd_gather(const FrontDev* __restrict__ fr, const SlotStatic* __restrict__ ss,
                                                    const int* __restrict__ rlist, const int64_t* __restrict__ xw_off,
                                                    double* XW, const double* __restrict__ x, int slot0, int do_diag,
                                                    int posdef_scale, int kind) {
   pdl_launch_dependents();
   const int s = slot0 + blockIdx.x;
   const FrontDev& f = fr[s];
   if (f.state != 2 || (kind >= 0 && f.part_kind != kind)) return;
   bwd_gather_front(f, rlist + ss[s].rl_off, XW + xw_off[s], x, do_diag, posdef_scale, blockIdx.y * blockDim.x + threadIdx.x,
                    gridDim.y * blockDim.x);
}

// Backward step for column block `step`: PART[front][chunk][j] = sum_{r in chunk} L(r, k0+j) xw(r) by every CTA
// (one warp per group of KBC / 8 columns, lanes down the rows), then the last CTA to arrive computes
// x(block) = L(block,block)^{-T} (xw(block) - sum of the partials, in chunk order).
template <int KBC, bool FUSED, int T = 256>
__global__ void __launch_bounds__(T, T == 64 ? 6 : (KBC <= 64 ? 3 : (KBC <= 96 ? 2 : 1))) k_bwd_step(const FrontDev* __restrict__ fr, const int* __restrict__ list,
                                                  const int64_t* __restrict__ xw_off, double* XW,
                                                  const int64_t* __restrict__ part_off, double* PART, int* ticket, int step, int kind, GatherArgs ga) {
   constexpr int RD = T * 32 / KBC, CW = KBC / (T / 32), RL = RD / 32, SLD = KBC + 1;
   pdl_launch_dependents();
   const int s = list[blockIdx.x];
   const FrontDev& f = fr[s];
   if (f.state != 2 || (kind >= 0 && f.part_kind != kind)) return;
   const int ne = f.nelim, k0 = step * KB;
   if (k0 >= ne) return;
   const int kb = min(KB, ne - k0), m = f.m;
   const int nchunk = (m - (k0 + kb) + RD - 1) / RD;
   if ((int)blockIdx.y >= max(nchunk, 1)) return;
   extern __shared__ double sm[];
   double* S = sm;                    // KBC x SLD
   double* xs = sm + KBC * SLD;       // max(KBC, RD)
   double* dots = xs + (KBC > RD ? KBC : RD);   // 32
   __shared__ int is_last;
   double* xw = XW + xw_off[s];
   double* part = PART + part_off[s];
   const size_t ldl = f.ldl;
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   if (FUSED) {
      // this front's part of L towards L2 while the level before is still running
      const int nlines = (m - k0 + 15) / 16;
      for (int e = threadIdx.x; e < kb * nlines; e += T) {
         const int j = e / nlines, q = e - j * nlines;
         asm volatile("prefetch.global.L2 [%0];" ::"l"(f.L + (size_t)(k0 + j) * ldl + min(k0 + 16 * q, m - 1)));
      }
   }
   pdl_wait();
   if (FUSED) {
      bwd_gather_front(f, ga.rlist + ga.ss[s].rl_off, xw, ga.x, ga.do_diag, ga.posdef_scale, threadIdx.x, T);
      __syncthreads();
   }
   if (nchunk > 0) {
      const int nctas = min((int)gridDim.y, nchunk);      // CTA y takes chunks y, y + gridDim.y, ...
      for (int ch = blockIdx.y; ch < nchunk; ch += gridDim.y) {
         const int r0 = k0 + kb + ch * RD, r1 = min(m, r0 + RD);
         double v[CW][RL];
#pragma unroll
         for (int c = 0; c < CW; ++c) {
            const int j = wid * CW + c;
            const double* __restrict__ col = f.L + (size_t)(k0 + j) * ldl + r0;
#pragma unroll
            for (int u = 0; u < RL; ++u) v[c][u] = (j < kb && r0 + lane + 32 * u < r1) ? col[lane + 32 * u] : 0.0;
         }
         __syncthreads();
         for (int i = threadIdx.x; i < RD; i += T) xs[i] = (r0 + i < r1) ? xw[r0 + i] : 0.0;
         __syncthreads();
         double* pc = part + (size_t)ch * KB;
#pragma unroll
         for (int c = 0; c < CW; ++c) {
            double acc = 0.0;
#pragma unroll
            for (int u = 0; u < RL; ++u) acc += v[c][u] * xs[lane + 32 * u];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            if (lane == 0 && wid * CW + c < kb) pc[wid * CW + c] = acc;
         }
      }
      if (nctas > 1) {
         __threadfence();
         __syncthreads();
         if (threadIdx.x == 0) {
            const int t = atomicAdd(ticket + s, 1);
            is_last = (t == nctas - 1);
            if (is_last) ticket[s] = 0;       // ready for the next step
         }
         __syncthreads();
         if (!is_last) return;
         __threadfence();
      } else {
         __syncthreads();
      }
   }
   stage_triangle(S, SLD, f.L + (size_t)k0 * ldl + k0, ldl, kb);
   for (int i = threadIdx.x; i < kb; i += T) {
      double vv = xw[k0 + i];
      for (int c = 0; c < nchunk; ++c) vv -= __ldcg(part + (size_t)c * KB + i);
      xs[i] = vv;
   }
   cp_async_wait<0>();
   __syncthreads();
   smem_solve_lower_t(S, SLD, kb, xs, dots);
   if (FUSED) { for (int i = threadIdx.x; i < kb; i += T) ga.x[f.perm[i]] = xs[i]; }
   else for (int i = threadIdx.x; i < kb; i += T) xw[k0 + i] = xs[i];
}

__global__ void __launch_bounds__(256) k_bwd_store(const FrontDev* __restrict__ fr, const int64_t* __restrict__ xw_off,
                                                   const double* __restrict__ XW, double* x, int slot0, int kind) {
   const int s = slot0 + blockIdx.x;
   const FrontDev& f = fr[s];
   if (f.state != 2 || (kind >= 0 && f.part_kind != kind)) return;
   const double* xw = XW + xw_off[s];
   for (int k = threadIdx.x; k < f.nelim; k += blockDim.x) x[f.perm[k]] = xw[k];
}
