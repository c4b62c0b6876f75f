// Level-scheduled triangular solves (included by engine.cu).
//
// Reference behaviour: NumericSubtree.hxx:280-400 (gather, trsv/gemv per node, scatter) and
// ldlt_app.cxx:2537-2588.  Here every level of the (fused) assembly tree is processed by batched
// kernels; a front's L is walked in column blocks of KB: a single CTA solves the KB x KB triangle,
// then all rows below are updated (forward) / reduced (backward) by many CTAs, so the big fronts
// stream L at HBM speed instead of through one SM.  No atomics: children's contributions are
// gathered by the parent in child order, partial dot products are summed in chunk order, so a
// solve is bit-reproducible.
#pragma once

constexpr int KB = 256;      // solve column block
constexpr int RC = 64;       // rows per CTA in the forward update
constexpr int RD = 128;      // rows per CTA in the backward dot products

__device__ __forceinline__ int row_gidx(const FrontDev& f, const int* rl, int r) {
   return r < f.n ? f.perm[r] : rl[f.ncol + (r - f.n)];
}

__global__ void k_permute_in(const double* __restrict__ x, double* __restrict__ xp, const int* __restrict__ order, int n) {
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n && order[i] >= 0) xp[order[i]] = x[i];
}
__global__ void k_permute_out(double* __restrict__ x, const double* __restrict__ xp, const int* __restrict__ order, int n) {
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n && order[i] >= 0) x[i] = xp[order[i]];
}

// multi-GPU: keep only the entries this rank produced (mask[p] != 0 for its pivots), zero elsewhere
__global__ void k_permute_out_masked(double* __restrict__ x, const double* __restrict__ xp, const int* __restrict__ order,
                                     const char* __restrict__ mask, int n) {
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) x[i] = (order[i] >= 0 && mask[order[i]]) ? xp[order[i]] : 0.0;
}
// multi-GPU: take the root part's solution entries from the owner's pivot-order vector
__global__ void k_merge_masked(double* __restrict__ xp, const double* __restrict__ src, const char* __restrict__ mine, int n) {
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n && !mine[i]) xp[i] = src[i];   // everything this rank did not eliminate itself (its ancestors are among them)
}

// xw = [b(pivot rows); 0] + sum over children (in child order) of their update vectors
__global__ void __launch_bounds__(256) k_fwd_gather(const FrontDev* __restrict__ fr, const SlotStatic* __restrict__ ss,
                                                    const int* __restrict__ child_idx, const int* __restrict__ child_doff,
                                                    const int* __restrict__ rel, const int64_t* __restrict__ xw_off,
                                                    double* XW, double* XA, const double* __restrict__ x, int slot0, int kind) {
   const int s = slot0 + blockIdx.x;
   const FrontDev& f = fr[s];
   if (f.state != 2 || (kind >= 0 && f.part_kind != kind)) return;
   const SlotStatic st = ss[s];
   double* xw = XW + xw_off[s];
   double* acc = XA + xw_off[s];   // indexed by the parent's ASSEMBLY order
   const int m = f.m, ne = f.nelim;
   const bool has_children = st.child_end > st.child_begin;
   if (has_children) {
      for (int r = threadIdx.x; r < m; r += blockDim.x) acc[r] = 0.0;
      __syncthreads();
      for (int ci = st.child_begin; ci < st.child_end; ++ci) {
         const int cs = child_idx[ci];
         const FrontDev& fc = fr[cs];
         const int t = fc.m - fc.nelim, nd = fc.n - fc.nelim, doff = child_doff[ci];
         const double* xc = XW + xw_off[cs] + fc.nelim;
         const int* relc = rel + ss[cs].rel_off;
         for (int rr = threadIdx.x; rr < t; rr += blockDim.x) {
            int d;
            if (rr < nd) d = f.ncol + doff + rr;
            else { d = relc[rr - nd]; if (d >= f.ncol) d += f.ndelay_in; }
            acc[d] += xc[rr];
         }
         __syncthreads();
      }
   }
   // pivoting permuted the candidate rows after assembly: orig[] maps them back
   for (int r = threadIdx.x; r < m; r += blockDim.x) {
      double v = r < ne ? x[f.perm[r]] : 0.0;
      if (has_children) v += acc[r < f.n ? f.orig[r] : r];
      xw[r] = v;
   }
}

// solve the KB x KB unit-lower triangle of block `step`; one CTA per front
__global__ void __launch_bounds__(256) k_fwd_tri(const FrontDev* __restrict__ fr, const int* __restrict__ list,
                                                 const int64_t* __restrict__ xw_off, double* XW, int step, int kind) {
   const int s = list[blockIdx.x];
   const FrontDev& f = fr[s];
   if (f.state != 2 || (kind >= 0 && f.part_kind != kind)) return;
   const int ne = f.nelim, k0 = step * KB;
   if (k0 >= ne) return;
   const int kb = min(KB, ne - k0);
   double* xw = XW + xw_off[s];
   const size_t ldl = f.ldl;
   const double* __restrict__ L = f.L;
   __shared__ double ys[KB];
   for (int i = threadIdx.x; i < kb; i += 256) ys[i] = xw[k0 + i];
   __syncthreads();
   const int lane = threadIdx.x & 31;
   for (int s0 = 0; s0 < kb; s0 += 32) {
      const int sb = min(32, kb - s0);
      if (threadIdx.x < 32) {
         // lane i holds row i of the 32x32 sub-triangle in registers
         double lrow[32];
         const double* Lb = L + (size_t)(k0 + s0) * ldl + k0 + s0 + lane;
#pragma unroll
         for (int j = 0; j < 32; ++j) lrow[j] = (j < lane && lane < sb) ? Lb[(size_t)j * ldl] : 0.0;
         double xi = lane < sb ? ys[s0 + lane] : 0.0;
#pragma unroll
         for (int j = 0; j < 32; ++j) {
            double yj = __shfl_sync(0xffffffffu, xi, j);
            xi -= lrow[j] * yj;
         }
         if (lane < sb) ys[s0 + lane] = xi;
      }
      __syncthreads();
      // rows of this KB block below the sub-block
      const int r = s0 + sb + threadIdx.x;
      if (r < kb) {
         const double* Lr = L + (size_t)(k0 + s0) * ldl + k0 + r;
         double acc = 0.0;
#pragma unroll 8
         for (int j = 0; j < sb; ++j) acc += Lr[(size_t)j * ldl] * ys[s0 + j];
         ys[r] -= acc;
      }
      __syncthreads();
   }
   for (int i = threadIdx.x; i < kb; i += 256) xw[k0 + i] = ys[i];
}

// xw(r) -= L(r, block) y(block) for the rows below block `step`; RC rows per CTA, 4 threads per row
__global__ void __launch_bounds__(256) k_fwd_upd(const FrontDev* __restrict__ fr, const int* __restrict__ list,
                                                 const int64_t* __restrict__ xw_off, double* XW, int step, int kind) {
   const int s = list[blockIdx.x];
   const FrontDev& f = fr[s];
   if (f.state != 2 || (kind >= 0 && f.part_kind != kind)) return;
   const int ne = f.nelim, k0 = step * KB;
   if (k0 >= ne) return;
   const int kb = min(KB, ne - k0);
   const int r0 = k0 + kb + blockIdx.y * RC;
   if (r0 >= f.m) return;
   double* xw = XW + xw_off[s];
   __shared__ double ys[KB];
   __shared__ double part[4][RC];
   for (int i = threadIdx.x; i < kb; i += 256) ys[i] = xw[k0 + i];
   __syncthreads();
   const int rl = threadIdx.x & (RC - 1), q = threadIdx.x / RC;
   const int r = r0 + rl;
   double acc = 0.0;
   if (r < f.m) {
      const int j0 = q * (KB / 4), j1 = min(kb, j0 + KB / 4);
      const double* __restrict__ Lr = f.L + (size_t)(k0 + j0) * f.ldl + r;
      const size_t ldl = f.ldl;
      int j = j0;
      for (; j + 8 <= j1; j += 8) {
         double v[8];
#pragma unroll
         for (int u = 0; u < 8; ++u) v[u] = Lr[(size_t)(j - j0 + u) * ldl];
#pragma unroll
         for (int u = 0; u < 8; ++u) acc += v[u] * ys[j + u];
      }
      for (; j < j1; ++j) acc += Lr[(size_t)(j - j0) * ldl] * ys[j];
   }
   part[q][rl] = acc;
   __syncthreads();
   if (q == 0 && r < f.m) xw[r] -= (part[0][rl] + part[1][rl]) + (part[2][rl] + part[3][rl]);
}

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
__global__ void __launch_bounds__(256) k_bwd_gather(const FrontDev* __restrict__ fr, const SlotStatic* __restrict__ ss,
                                                    const int* __restrict__ rlist, const int64_t* __restrict__ xw_off,
                                                    double* XW, const double* __restrict__ x, int slot0, int do_diag,
                                                    int posdef_scale, int kind) {
   const int s = slot0 + blockIdx.x;
   const FrontDev& f = fr[s];
   if (f.state != 2 || (kind >= 0 && f.part_kind != kind)) return;
   const int* rl = rlist + ss[s].rl_off;
   double* xw = XW + xw_off[s];
   const int m = f.m, ne = f.nelim;
   for (int r = threadIdx.x; r < m; r += blockDim.x) {
      double v;
      if (r >= ne) v = x[row_gidx(f, rl, r)];
      else {
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

// partial dot products of the rows below block `step` with the block's columns:
// PART[front][chunk][j] = sum_{r in chunk} L(r, k0+j) xw(r)
__global__ void __launch_bounds__(256) k_bwd_dots(const FrontDev* __restrict__ fr, const int* __restrict__ list,
                                                  const int64_t* __restrict__ xw_off, const double* __restrict__ XW,
                                                  const int64_t* __restrict__ part_off, double* PART, int step, int kind) {
   const int s = list[blockIdx.x];
   const FrontDev& f = fr[s];
   if (f.state != 2 || (kind >= 0 && f.part_kind != kind)) return;
   const int ne = f.nelim, k0 = step * KB;
   if (k0 >= ne) return;
   const int kb = min(KB, ne - k0);
   const int r0 = k0 + kb + blockIdx.y * RD;
   if (r0 >= f.m) return;
   const int r1 = min(f.m, r0 + RD);
   const double* xw = XW + xw_off[s];
   double* part = PART + part_off[s] + (size_t)blockIdx.y * KB;
   __shared__ double xs[RD];
   for (int i = threadIdx.x; i < RD; i += 256) xs[i] = (r0 + i < r1) ? xw[r0 + i] : 0.0;
   __syncthreads();
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   const size_t ldl = f.ldl;
   for (int j = wid; j < kb; j += 8) {
      const double* __restrict__ col = f.L + (size_t)(k0 + j) * ldl + r0;
      double v[RD / 32];
#pragma unroll
      for (int u = 0; u < RD / 32; ++u) v[u] = (r0 + lane + 32 * u < r1) ? col[lane + 32 * u] : 0.0;
      double acc = 0.0;
#pragma unroll
      for (int u = 0; u < RD / 32; ++u) acc += v[u] * xs[lane + 32 * u];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      if (lane == 0) part[j] = acc;
   }
}

// x(block) = L(block,block)^{-T} (xw(block) - sum of partial dots); one CTA per front
__global__ void __launch_bounds__(256) k_bwd_tri(const FrontDev* __restrict__ fr, const int* __restrict__ list,
                                                 const int64_t* __restrict__ xw_off, double* XW,
                                                 const int64_t* __restrict__ part_off, const double* __restrict__ PART, int step, int kind) {
   const int s = list[blockIdx.x];
   const FrontDev& f = fr[s];
   if (f.state != 2 || (kind >= 0 && f.part_kind != kind)) return;
   const int ne = f.nelim, k0 = step * KB;
   if (k0 >= ne) return;
   const int kb = min(KB, ne - k0);
   double* xw = XW + xw_off[s];
   const double* part = PART + part_off[s];
   const int nchunk = (f.m - (k0 + kb) + RD - 1) / RD;
   const size_t ldl = f.ldl;
   const double* __restrict__ L = f.L;
   __shared__ double ys[KB];
   for (int i = threadIdx.x; i < kb; i += 256) {
      double v = xw[k0 + i];
      for (int c = 0; c < nchunk; ++c) v -= part[(size_t)c * KB + i];
      ys[i] = v;
   }
   __syncthreads();
   const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
   const int nsb = (kb + 31) / 32;
   for (int sbi = nsb - 1; sbi >= 0; --sbi) {
      const int s0 = sbi * 32, sb = min(32, kb - s0);
      // subtract the already solved rows of this KB block: columns s0..s0+sb, rows s0+sb..kb
      __shared__ double dots[32];
      for (int j = wid; j < sb; j += 8) {
         const double* col = L + (size_t)(k0 + s0 + j) * ldl + k0;
         double acc = 0.0;
         for (int r = s0 + sb + lane; r < kb; r += 32) acc += col[r] * ys[r];
#pragma unroll
         for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
         if (lane == 0) dots[j] = acc;
      }
      __syncthreads();
      if (threadIdx.x < 32) {
         // lane i holds column i of the sub-triangle (rows below the diagonal) = row i of L^T
         double lcol[32];
         const double* Lb = L + (size_t)(k0 + s0 + lane) * ldl + k0 + s0;
#pragma unroll
         for (int j = 0; j < 32; ++j) lcol[j] = (j > lane && j < sb && lane < sb) ? Lb[j] : 0.0;
         double v = lane < sb ? ys[s0 + lane] - dots[lane] : 0.0;
#pragma unroll
         for (int j = 31; j >= 0; --j) {
            double xj = __shfl_sync(0xffffffffu, v, j);
            v -= lcol[j] * xj;
         }
         if (lane < sb) ys[s0 + lane] = v;
      }
      __syncthreads();
   }
   for (int i = threadIdx.x; i < kb; i += 256) xw[k0 + i] = ys[i];
}

__global__ void __launch_bounds__(256) k_bwd_store(const FrontDev* __restrict__ fr, const int64_t* __restrict__ xw_off,
                                                   const double* __restrict__ XW, double* x, int slot0, int kind) {
   const int s = slot0 + blockIdx.x;
   const FrontDev& f = fr[s];
   if (f.state != 2 || (kind >= 0 && f.part_kind != kind)) return;
   const double* xw = XW + xw_off[s];
   for (int k = threadIdx.x; k < f.nelim; k += blockDim.x) x[f.perm[k]] = xw[k];
}
