/* TEST INFRASTRUCTURE ONLY (oracle).  Plain-C restatement of the reference's numeric path:
 * multifrontal LDL^T with a-posteriori threshold pivoting (APP), a threshold-partial-pivoting
 * (TPP) finish, delayed pivots, inertia count, and the forward / diagonal / backward solves.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may build, load or call
 * this file; the product (galahad_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks it against (i) the reference's own
 * known-answer systems (src/sls/slst.f90:29-51, src/ssids/ssidss.data, examples/bcsstk01) and
 * (ii) the reference's C++ engine compiled unmodified into oracle/_ref (inertia exact, solutions to
 * rounding) on seeded random SPD / indefinite / KKT matrices.
 *
 * What each part follows in /root/reference:
 *   assemble()        src/ssids/cpu/kernels/assemble.hxx:142-345 (assemble_pre: zeroed front, A
 *                     scatter via nlist, delayed columns appended after the node's own columns,
 *                     child contributions) and :350-437 (assemble_post, contribution block part)
 *   block_ldlt()      src/ssids/cpu/kernels/block_ldlt.hxx:290-414 (complete pivoting inside the
 *                     diagonal block; test_2x2 :214-218; zero-pivot rule :303-318)
 *   app_pass()        src/ssids/cpu/kernels/ldlt_app.cxx: apply_pivot :331 (L21 = A21 L11^-T D^-1,
 *                     zero pivots -> 0 or inf), check_threshold :304 (|l_ij| > 1/u fails the
 *                     column), Column::adjust :111 (never split a 2x2), Block::update :1081
 *                     (trailing -= L (LD)^T with failed columns restored from the backup),
 *                     failed columns moved behind the candidates :2379-2474
 *   tpp_pass()        src/ssids/cpu/kernels/ldlt_tpp.cxx:166-270 (+ test_2x2 :89-119) and the
 *                     Schur update for its pivots, src/ssids/cpu/factor.hxx:73-110
 *   inertia()         src/ssids/cpu/NumericSubtree.hxx:242-274
 *   solve_*()         src/ssids/cpu/NumericSubtree.hxx:280-400 and ldlt_app.cxx:2537-2588
 *   enquire()/alter() src/ssids/cpu/NumericSubtree.hxx:418-477
 * D is stored inverted, d[2k] d[2k+1], with d[2k+2] = +inf tagging the second column of a 2x2
 * (ldlt_app.cxx:324-329).  The reference's own blocking (256 outer / 32 inner, task DAG) is an
 * implementation schedule, not part of the result contract: pivot order may differ, inertia and
 * the solution may not.  This restatement uses one 32-column panel at a time, the schedule the
 * CUDA engine also uses, so that per-front intermediate results are comparable.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NB 32

typedef struct {
   int m, n;        /* rows / candidate columns including incoming delays */
   int ncol, nrow;  /* symbolic sizes */
   int ndelay_in, nelim;
   double *L;       /* m x nelim (ld = m) after compaction; full m x m front while active */
   int ldl;
   double *F;       /* working square front (m x m, lower), freed once the parent has consumed it */
   double *d;       /* 2*n */
   int *perm;       /* n global pivot-order indices of the candidate columns (permuted) */
} onode;

typedef struct {
   int n, nnodes, posdef;
   const int *sptr, *sparent, *rlist;
   const int64_t *rptr;
   onode *nodes;
   int num_neg, num_two, num_zero, num_delay, maxfront, flag;
   int not_first_pass, not_second_pass;
} ofact;

static double dabs(double x) { return fabs(x); }

/* symmetric swap of front indices a < b on lower storage, all columns (incl. finished L columns) */
static void sym_swap(double *F, int ld, int m, int a, int b) {
   double t;
   int c, r;
   if (a == b) return;
   if (a > b) { int s = a; a = b; b = s; }
   for (c = 0; c < a; ++c) { t = F[a + (size_t)c * ld]; F[a + (size_t)c * ld] = F[b + (size_t)c * ld]; F[b + (size_t)c * ld] = t; }
   t = F[a + (size_t)a * ld]; F[a + (size_t)a * ld] = F[b + (size_t)b * ld]; F[b + (size_t)b * ld] = t;
   for (c = a + 1; c < b; ++c) { t = F[c + (size_t)a * ld]; F[c + (size_t)a * ld] = F[b + (size_t)c * ld]; F[b + (size_t)c * ld] = t; }
   for (r = b + 1; r < m; ++r) { t = F[r + (size_t)a * ld]; F[r + (size_t)a * ld] = F[r + (size_t)b * ld]; F[r + (size_t)b * ld] = t; }
}

/* Complete-pivoting LDL^T of a w x w block held FULL symmetric in a (ld NB).  On exit the strict
 * lower part holds L11, ldm the matching L*D entries, d the inverted pivots, ptype 0/1/2 =
 * 1x1 / first / second column of a 2x2, lperm the local permutation.  Returns 0, or -1 when the
 * block is numerically zero and action is false. */
static int block_ldlt(int w, double *a, double *ldm, double *d, int *ptype, int *lperm,
                      double u, double small, int action, int posdef, int *notposdef) {
   int p, r, c, i;
   for (i = 0; i < w; ++i) lperm[i] = i;
   for (p = 0; p < w;) {
      int t = p, mm = p, pivsiz;
      double best = -1.0, a11 = 0, a21 = 0, a22 = 0, detscale = 0, detpiv = 0;
      if (posdef) {
         if (!(a[p + p * NB] > 0.0)) { *notposdef = 1; return -2; }
         best = a[p + p * NB];
      } else {
         /* largest entry of the remaining lower triangle; ties: smallest column, then row */
         for (c = p; c < w; ++c)
            for (r = c; r < w; ++r) {
               double v = dabs(a[r + c * NB]);
               if (v > best) { best = v; t = r; mm = c; }
            }
         if (best < small) {
            if (!action) return -1;
            for (; p < w; ++p) {
               d[2 * p] = 0.0; d[2 * p + 1] = 0.0; ptype[p] = 0;
               for (r = p; r < w; ++r) { a[r + p * NB] = 0.0; a[p + r * NB] = 0.0; ldm[r + p * NB] = 0.0; }
            }
            break;
         }
      }
      if (t == mm) {
         pivsiz = 1; a11 = a[t + t * NB];
      } else {
         a11 = a[mm + mm * NB]; a22 = a[t + t * NB]; a21 = a[t + mm * NB];
         detscale = 1.0 / dabs(a21);
         detpiv = (a11 * detscale) * a22 - dabs(a21);
         if (dabs(detpiv) >= dabs(a21) / 2) {
            pivsiz = 2;
         } else if (dabs(a11) > dabs(a22)) {
            pivsiz = 1; t = mm;
         } else {
            pivsiz = 1; a11 = a22; mm = t;
         }
      }
      if (pivsiz == 1) {
         /* bring t to position p (full symmetric storage: swap rows then columns) */
         if (t != p) {
            double s; int q;
            for (q = 0; q < w; ++q) { s = a[p + q * NB]; a[p + q * NB] = a[t + q * NB]; a[t + q * NB] = s; }
            for (q = 0; q < w; ++q) { s = a[q + p * NB]; a[q + p * NB] = a[q + t * NB]; a[q + t * NB] = s; }
            for (q = 0; q < p; ++q) { s = ldm[p + q * NB]; ldm[p + q * NB] = ldm[t + q * NB]; ldm[t + q * NB] = s; }
            i = lperm[p]; lperm[p] = lperm[t]; lperm[t] = i;
         }
         {
            double d11 = 1.0 / a[p + p * NB];
            for (r = p + 1; r < w; ++r) { ldm[r + p * NB] = a[r + p * NB]; a[r + p * NB] *= d11; }
            for (c = p + 1; c < w; ++c)
               for (r = p + 1; r < w; ++r) a[r + c * NB] -= a[r + p * NB] * ldm[c + p * NB];
            d[2 * p] = d11; d[2 * p + 1] = 0.0; ptype[p] = 0;
            a[p + p * NB] = 1.0;
            for (c = p + 1; c < w; ++c) a[p + c * NB] = 0.0;
         }
         p += 1;
      } else {
         /* 2x2 pivot on (mm, t), t > mm: mm -> p, t -> p+1 */
         double s; int q, src;
         for (i = 0; i < 2; ++i) {
            int dst = p + i; src = i == 0 ? mm : t;
            if (src != dst) {
               for (q = 0; q < w; ++q) { s = a[dst + q * NB]; a[dst + q * NB] = a[src + q * NB]; a[src + q * NB] = s; }
               for (q = 0; q < w; ++q) { s = a[q + dst * NB]; a[q + dst * NB] = a[q + src * NB]; a[q + src * NB] = s; }
               for (q = 0; q < p; ++q) { s = ldm[dst + q * NB]; ldm[dst + q * NB] = ldm[src + q * NB]; ldm[src + q * NB] = s; }
               q = lperm[dst]; lperm[dst] = lperm[src]; lperm[src] = q;
               if (i == 0 && t == p) t = mm; /* t was displaced by the first swap */
            }
         }
         {
            double d11 = (a22 * detscale) / detpiv, d22 = (a11 * detscale) / detpiv;
            double d21 = (-a21 * detscale) / detpiv;
            for (r = p + 2; r < w; ++r) {
               double w1 = a[r + p * NB], w2 = a[r + (p + 1) * NB];
               ldm[r + p * NB] = w1; ldm[r + (p + 1) * NB] = w2;
               a[r + p * NB] = d11 * w1 + d21 * w2;
               a[r + (p + 1) * NB] = d21 * w1 + d22 * w2;
            }
            for (c = p + 2; c < w; ++c)
               for (r = p + 2; r < w; ++r)
                  a[r + c * NB] -= a[r + p * NB] * ldm[c + p * NB] + a[r + (p + 1) * NB] * ldm[c + (p + 1) * NB];
            d[2 * p] = d11; d[2 * p + 1] = d21; d[2 * p + 2] = INFINITY; d[2 * p + 3] = d22;
            ptype[p] = 1; ptype[p + 1] = 2;
            a[p + p * NB] = 1.0; a[p + 1 + p * NB] = 0.0; a[p + (p + 1) * NB] = 0.0; a[p + 1 + (p + 1) * NB] = 1.0;
            ldm[p + 1 + p * NB] = 0.0;
            for (c = p + 2; c < w; ++c) { a[p + c * NB] = 0.0; a[p + 1 + c * NB] = 0.0; }
         }
         p += 2;
      }
   }
   return 0;
}

/* One APP sweep over the candidate columns [*done, n) of front F (m x m, lower, ld m). */
static int app_pass(int m, int n, double *F, int ld, double *d, int *perm, int *done_io,
                    double u, double small, int action, int posdef, int *flag) {
   int done = *done_io, end = n;
   double *a = (double *)malloc(sizeof(double) * NB * NB), *a0 = (double *)malloc(sizeof(double) * NB * NB);
   double *ldm = (double *)calloc(NB * NB, sizeof(double));
   double *Wl = (double *)malloc(sizeof(double) * (size_t)m * NB), *Wld = (double *)malloc(sizeof(double) * (size_t)m * NB);
   double *Wo = (double *)malloc(sizeof(double) * (size_t)m * NB);
   double dk[2 * NB + 2], tmpv[NB];
   int ptype[NB], lperm[NB], tmpi[NB];
   const double lim = (u > 0) ? 1.0 / u : INFINITY;
   while (done < end) {
      int w = end - done < NB ? end - done : NB, i, k, r, c, fail, nelim, nfail, notpd = 0, rc;
      for (c = 0; c < w; ++c)
         for (r = c; r < w; ++r) {
            double v = F[done + r + (size_t)(done + c) * ld];
            a[r + c * NB] = v; a[c + r * NB] = v;
         }
      memcpy(a0, a, sizeof(double) * NB * NB);
      memset(ldm, 0, sizeof(double) * NB * NB);
      rc = block_ldlt(w, a, ldm, dk, ptype, lperm, u, small, action, posdef, &notpd);
      if (rc == -1) { *flag = -5; break; }
      if (rc == -2) { *flag = -6; break; }
      /* rows of the block itself: L11 and (L D)11 */
      for (i = 0; i < w; ++i)
         for (k = 0; k < w; ++k) {
            Wl[done + i + (size_t)k * m] = i > k ? a[i + k * NB] : (i == k ? 1.0 : 0.0);
            Wld[done + i + (size_t)k * m] = i > k ? ldm[i + k * NB] : 0.0;
         }
      /* apply to the rows below and test a posteriori */
      fail = w;
      for (r = done + w; r < m; ++r) {
         double x[NB];
         for (k = 0; k < w; ++k) { x[k] = F[r + (size_t)(done + lperm[k]) * ld]; Wo[r + (size_t)k * m] = x[k]; }
         for (k = 0; k < w; ++k) {
            double s = x[k];
            for (i = 0; i < k; ++i) s -= x[i] * a[k + i * NB];
            x[k] = s;
         }
         for (k = 0; k < w;) {
            if (ptype[k] == 0) {
               double d11 = dk[2 * k], l;
               if (d11 == 0.0) l = (dabs(x[k]) < small) ? 0.0 : INFINITY * x[k];
               else l = x[k] * d11;
               Wl[r + (size_t)k * m] = l; Wld[r + (size_t)k * m] = x[k];
               if (!posdef && dabs(l) > lim && k < fail) fail = k;
               if (l != l && k < fail) fail = k;
               k += 1;
            } else {
               double d11 = dk[2 * k], d21 = dk[2 * k + 1], d22 = dk[2 * k + 3];
               double l1 = d11 * x[k] + d21 * x[k + 1], l2 = d21 * x[k] + d22 * x[k + 1];
               Wl[r + (size_t)k * m] = l1; Wl[r + (size_t)(k + 1) * m] = l2;
               Wld[r + (size_t)k * m] = x[k]; Wld[r + (size_t)(k + 1) * m] = x[k + 1];
               if ((dabs(l1) > lim || l1 != l1) && k < fail) fail = k;
               if ((dabs(l2) > lim || l2 != l2) && k + 1 < fail) fail = k + 1;
               k += 2;
            }
         }
      }
      nelim = fail;
      if (nelim < w && ptype[nelim] == 2) nelim -= 1;
      /* commit: rows of finished L columns and the pivot list follow the local permutation */
      for (c = 0; c < done; ++c) {
         for (i = 0; i < w; ++i) tmpv[i] = F[done + lperm[i] + (size_t)c * ld];
         for (i = 0; i < w; ++i) F[done + i + (size_t)c * ld] = tmpv[i];
      }
      for (i = 0; i < w; ++i) tmpi[i] = perm[done + lperm[i]];
      for (i = 0; i < w; ++i) perm[done + i] = tmpi[i];
      for (k = 0; k < w; ++k) {
         for (i = k; i < w; ++i)
            F[done + i + (size_t)(done + k) * ld] = k < nelim ? (i == k ? 1.0 : a[i + k * NB]) : a0[lperm[i] + lperm[k] * NB];
         for (r = done + w; r < m; ++r)
            F[r + (size_t)(done + k) * ld] = k < nelim ? Wl[r + (size_t)k * m] : Wo[r + (size_t)k * m];
      }
      for (k = 0; k < nelim; ++k) { d[2 * (done + k)] = dk[2 * k]; d[2 * (done + k) + 1] = dk[2 * k + 1]; }
      /* trailing update, lower triangle, rows/cols >= done+nelim */
      if (nelim > 0)
         for (c = done + nelim; c < m; ++c)
            for (k = 0; k < nelim; ++k) {
               double b = Wld[c + (size_t)k * m];
               if (b == 0.0) continue;
               for (r = c; r < m; ++r) F[r + (size_t)c * ld] -= Wl[r + (size_t)k * m] * b;
            }
      done += nelim;
      nfail = w - nelim;
      for (i = 0; i < nfail; ++i) {
         int p = done + i, q = end - 1 - i;
         if (p < q) { int s; sym_swap(F, ld, m, p, q); s = perm[p]; perm[p] = perm[q]; perm[q] = s; }
      }
      end -= nfail;
      if (posdef && nfail > 0) { *flag = -6; break; }
   }
   free(a); free(a0); free(ldm); free(Wl); free(Wld); free(Wo);
   *done_io = done;
   return 0;
}

/* max |.| over row/column `col` of the trailing symmetric matrix starting at `from`, skipping
 * index `excl` and the diagonal (ldlt_tpp.cxx find_rc_abs_max_exclude) */
static double rc_abs_max_excl(const double *F, int ld, int m, int from, int col, int excl) {
   double best = 0.0;
   int c, r;
   for (c = from; c < col; ++c) if (c != excl) { double v = dabs(F[col + (size_t)c * ld]); if (v > best) best = v; }
   for (r = col + 1; r < m; ++r) if (r != excl) { double v = dabs(F[r + (size_t)col * ld]); if (v > best) best = v; }
   return best;
}

static int col_small(const double *F, int ld, int m, int from, int col, double small) {
   int c, r;
   for (c = from; c < col; ++c) if (dabs(F[col + (size_t)c * ld]) >= small) return 0;
   for (r = col; r < m; ++r) if (dabs(F[r + (size_t)col * ld]) >= small) return 0;
   return 1;
}

/* TPP over candidate columns [*done, n); right-looking on the WHOLE trailing matrix so that the
 * contribution block is updated in the same sweep (factor.hxx:92-102 does it as one GEMM after). */
static void tpp_pass(int m, int n, double *F, int ld, double *d, int *perm, int *done_io,
                     double u, double small, int action, int *flag) {
   int ne = *done_io;
   double *l1 = (double *)malloc(sizeof(double) * m), *l2 = (double *)malloc(sizeof(double) * m);
   while (ne < n) {
      int p, found = 0, r, c, s;
      if (col_small(F, ld, m, ne, ne, small)) {
         if (!action) { *flag = -5; break; }
         for (r = ne; r < m; ++r) F[r + (size_t)ne * ld] = 0.0;
         F[ne + (size_t)ne * ld] = 1.0;
         d[2 * ne] = 0.0; d[2 * ne + 1] = 0.0; ne++;
         continue;
      }
      for (p = ne + 1; p < n && !found; ++p) {
         int t = ne;
         double best = -1.0, maxt, maxp, a11, a21, a22, maxpiv, detscale, detpiv0, detpiv1, detpiv;
         if (col_small(F, ld, m, ne, p, small)) {
            if (!action) { *flag = -5; goto out; }
            sym_swap(F, ld, m, ne, p); s = perm[ne]; perm[ne] = perm[p]; perm[p] = s;
            for (r = ne; r < m; ++r) F[r + (size_t)ne * ld] = 0.0;
            F[ne + (size_t)ne * ld] = 1.0;
            d[2 * ne] = 0.0; d[2 * ne + 1] = 0.0; ne++;
            found = 1; break;
         }
         for (c = ne; c < p; ++c) { double v = dabs(F[p + (size_t)c * ld]); if (v > best) { best = v; t = c; } }
         maxt = rc_abs_max_excl(F, ld, m, ne, t, p);
         maxp = rc_abs_max_excl(F, ld, m, ne, p, t);
         a11 = F[t + (size_t)t * ld]; a21 = F[p + (size_t)t * ld]; a22 = F[p + (size_t)p * ld];
         maxpiv = fmax(dabs(a11), fmax(dabs(a21), dabs(a22)));
         if (maxpiv >= small) {
            detscale = 1 / maxpiv;
            detpiv0 = (a11 * detscale) * a22; detpiv1 = (a21 * detscale) * a21; detpiv = detpiv0 - detpiv1;
            if (!(dabs(detpiv) < fmax(small, fmax(dabs(detpiv0 / 2), dabs(detpiv1 / 2))))) {
               double d11 = (a22 * detscale) / detpiv, d21 = (-a21 * detscale) / detpiv, d22 = (a11 * detscale) / detpiv;
               int ok;
               if (fmax(maxp, maxt) < small) ok = 1;
               else {
                  double x1 = dabs(d11) * maxt + dabs(d21) * maxp, x2 = dabs(d21) * maxt + dabs(d22) * maxp;
                  ok = (u * fmax(x1, x2) < 1.0);
               }
               if (ok) {
                  sym_swap(F, ld, m, ne, t); s = perm[ne]; perm[ne] = perm[t]; perm[t] = s;
                  sym_swap(F, ld, m, ne + 1, p); s = perm[ne + 1]; perm[ne + 1] = perm[p]; perm[p] = s;
                  for (r = ne + 2; r < m; ++r) {
                     double w1 = F[r + (size_t)ne * ld], w2 = F[r + (size_t)(ne + 1) * ld];
                     l1[r] = w1; l2[r] = w2;
                     F[r + (size_t)ne * ld] = d11 * w1 + d21 * w2;
                     F[r + (size_t)(ne + 1) * ld] = d21 * w1 + d22 * w2;
                  }
                  for (c = ne + 2; c < m; ++c)
                     for (r = c; r < m; ++r)
                        F[r + (size_t)c * ld] -= F[r + (size_t)ne * ld] * l1[c] + F[r + (size_t)(ne + 1) * ld] * l2[c];
                  F[ne + (size_t)ne * ld] = 1.0; F[ne + 1 + (size_t)ne * ld] = 0.0; F[ne + 1 + (size_t)(ne + 1) * ld] = 1.0;
                  d[2 * ne] = d11; d[2 * ne + 1] = d21; d[2 * ne + 2] = INFINITY; d[2 * ne + 3] = d22;
                  ne += 2; found = 1; break;
               }
            }
         }
         maxp = fmax(maxp, dabs(F[p + (size_t)t * ld]));
         if (dabs(F[p + (size_t)p * ld]) >= u * maxp) {
            double d11;
            sym_swap(F, ld, m, ne, p); s = perm[ne]; perm[ne] = perm[p]; perm[p] = s;
            d11 = 1 / F[ne + (size_t)ne * ld];
            for (r = ne + 1; r < m; ++r) { l1[r] = F[r + (size_t)ne * ld]; F[r + (size_t)ne * ld] *= d11; }
            for (c = ne + 1; c < m; ++c)
               for (r = c; r < m; ++r) F[r + (size_t)c * ld] -= F[r + (size_t)ne * ld] * l1[c];
            F[ne + (size_t)ne * ld] = 1.0;
            d[2 * ne] = d11; d[2 * ne + 1] = 0.0; ne += 1; found = 1; break;
         }
      }
      if (!found) {
         double maxp = rc_abs_max_excl(F, ld, m, ne, ne, -1);
         if (dabs(F[ne + (size_t)ne * ld]) >= u * maxp) {
            double d11 = 1 / F[ne + (size_t)ne * ld];
            for (r = ne + 1; r < m; ++r) { l1[r] = F[r + (size_t)ne * ld]; F[r + (size_t)ne * ld] *= d11; }
            for (c = ne + 1; c < m; ++c)
               for (r = c; r < m; ++r) F[r + (size_t)c * ld] -= F[r + (size_t)ne * ld] * l1[c];
            F[ne + (size_t)ne * ld] = 1.0;
            d[2 * ne] = d11; d[2 * ne + 1] = 0.0; ne += 1;
         } else break; /* out of pivots: the rest is delayed */
      }
   }
out:
   free(l1); free(l2);
   *done_io = ne;
}

void oracle_free(ofact *f) {
   int i;
   if (!f) return;
   for (i = 0; i < f->nnodes; ++i) { free(f->nodes[i].L); free(f->nodes[i].F); free(f->nodes[i].d); free(f->nodes[i].perm); }
   free(f->nodes); free(f);
}

/* All index arrays 0-based; sparent[node] == nnodes for roots; nlist = (src, col*nrow+row) pairs.
 * stats: flag, num_delay, num_neg, num_two, num_zero, maxfront, not_first_pass, not_second_pass */
ofact *oracle_factor(int n, int nnodes, const int *sptr, const int *sparent, const int64_t *rptr,
                     const int *rlist, const int64_t *nptr, const int64_t *nlist, const double *val,
                     int posdef, double u, double small, int action, int *stats) {
   ofact *f = (ofact *)calloc(1, sizeof(ofact));
   int *map = (int *)malloc(sizeof(int) * (n > 0 ? n : 1));
   int *chead = (int *)malloc(sizeof(int) * (nnodes + 1)), *cnext = (int *)malloc(sizeof(int) * (nnodes + 1));
   int ni, i;
   f->n = n; f->nnodes = nnodes; f->posdef = posdef; f->sptr = sptr; f->sparent = sparent; f->rptr = rptr; f->rlist = rlist;
   f->nodes = (onode *)calloc(nnodes > 0 ? nnodes : 1, sizeof(onode));
   if (u > 0.5) u = 0.5;
   if (u < 0) u = 0;
   for (i = 0; i <= nnodes; ++i) chead[i] = -1;
   for (ni = nnodes - 1; ni >= 0; --ni) { cnext[ni] = chead[sparent[ni]]; chead[sparent[ni]] = ni; }
   for (ni = 0; ni < nnodes && f->flag >= 0; ++ni) {
      onode *nd = &f->nodes[ni];
      int ncol = sptr[ni + 1] - sptr[ni], nrow = (int)(rptr[ni + 1] - rptr[ni]);
      int ndelay = 0, c, m, nn, doff, done, r, cc, first_done;
      int64_t e;
      const int *rl = rlist + rptr[ni];
      for (c = chead[ni]; c != -1; c = cnext[c]) ndelay += f->nodes[c].n - f->nodes[c].nelim;
      m = nrow + ndelay; nn = ncol + ndelay;
      nd->m = m; nd->n = nn; nd->ncol = ncol; nd->nrow = nrow; nd->ndelay_in = ndelay; nd->ldl = m;
      nd->F = (double *)calloc((size_t)m * m > 0 ? (size_t)m * m : 1, sizeof(double));
      nd->d = (double *)calloc(2 * (size_t)nn + 2, sizeof(double));
      nd->perm = (int *)malloc(sizeof(int) * (nn > 0 ? nn : 1));
      if (m > f->maxfront) f->maxfront = m;
      for (i = 0; i < ncol; ++i) nd->perm[i] = sptr[ni] + i;
      /* A */
      for (e = nptr[ni]; e < nptr[ni + 1]; ++e) {
         int64_t dest = nlist[2 * e + 1];
         int col = (int)(dest / nrow), row = (int)(dest % nrow);
         if (row >= ncol) row += ndelay;
         nd->F[row + (size_t)col * m] = val[nlist[2 * e]];
      }
      /* children: delayed columns go behind the node's own columns, the rest via the row map */
      for (i = 0; i < nrow; ++i) map[rl[i]] = i < ncol ? i : i + ndelay;
      doff = 0;
      for (c = chead[ni]; c != -1; c = cnext[c]) {
         onode *ch = &f->nodes[c];
         int cm = ch->m, cn = ch->n, ce = ch->nelim, nd_c = cn - ce, t = cm - ce;
         const int *crl = rlist + rptr[c] + ch->ncol;
         int *dst = (int *)malloc(sizeof(int) * (t > 0 ? t : 1));
         for (r = 0; r < t; ++r) dst[r] = r < nd_c ? ncol + doff + r : map[crl[r - nd_c]];
         for (r = 0; r < nd_c; ++r) nd->perm[ncol + doff + r] = ch->perm[ce + r];
         for (cc = 0; cc < t; ++cc)
            for (r = cc; r < t; ++r) {
               double v = ch->F[ce + r + (size_t)(ce + cc) * cm];
               int dr = dst[r], dc = dst[cc];
               if (dr < dc) { int s = dr; dr = dc; dc = s; }
               nd->F[dr + (size_t)dc * m] += v;
            }
         doff += nd_c;
         free(dst);
         /* the child keeps only its L columns from here on */
         {
            double *L = (double *)malloc(sizeof(double) * ((size_t)cm * ce > 0 ? (size_t)cm * ce : 1));
            memcpy(L, ch->F, sizeof(double) * (size_t)cm * ce);
            free(ch->F); ch->F = NULL; ch->L = L;
         }
      }
      /* factor */
      done = 0;
      app_pass(m, nn, nd->F, m, nd->d, nd->perm, &done, u, small, action, posdef, &f->flag);
      if (f->flag < 0) break;
      first_done = done;
      f->not_first_pass += nn - done;
      if (done < nn && !posdef) {
         tpp_pass(m, nn, nd->F, m, nd->d, nd->perm, &done, u, small, action, &f->flag);
         if (f->flag < 0) break;
         f->not_second_pass += nn - done;
      }
      (void)first_done;
      nd->nelim = done;
      f->num_delay += nn - done;
   }
   for (ni = 0; ni < nnodes; ++ni) {
      onode *nd = &f->nodes[ni];
      if (nd->F && !nd->L) { /* roots (and everything after an error) */
         size_t sz = (size_t)nd->m * nd->nelim;
         nd->L = (double *)malloc(sizeof(double) * (sz > 0 ? sz : 1));
         memcpy(nd->L, nd->F, sizeof(double) * sz);
         free(nd->F); nd->F = NULL;
      }
   }
   /* inertia (NumericSubtree.hxx:242-274) */
   if (!posdef && f->flag >= 0)
      for (ni = 0; ni < nnodes; ++ni) {
         onode *nd = &f->nodes[ni];
         for (i = 0; i < nd->nelim;) {
            double a11 = nd->d[2 * i], a21 = nd->d[2 * i + 1];
            if (i + 1 == nd->nelim || isfinite(nd->d[2 * i + 2])) {
               if (a11 == 0.0) { f->flag = 7; f->num_zero++; }
               if (a11 < 0.0) f->num_neg++;
               i += 1;
            } else {
               double a22 = nd->d[2 * i + 3], det = a11 * a22 - a21 * a21, tr = a11 + a22;
               f->num_two++;
               if (det < 0) f->num_neg++;
               else if (tr < 0) f->num_neg += 2;
               i += 2;
            }
         }
      }
   stats[0] = f->flag; stats[1] = f->num_delay; stats[2] = f->num_neg; stats[3] = f->num_two;
   stats[4] = f->num_zero; stats[5] = f->maxfront; stats[6] = f->not_first_pass; stats[7] = f->not_second_pass;
   free(map); free(chead); free(cnext);
   return f;
}

static int row_index(const ofact *f, int ni, int r) {
   const onode *nd = &f->nodes[ni];
   return r < nd->n ? nd->perm[r] : f->rlist[f->rptr[ni] + nd->ncol + (r - nd->n)];
}

/* x(ldx,nrhs) in pivot order.  job: 0 all, 1 fwd, 2 diag, 3 bwd, 4 diag+bwd */
int oracle_solve(const ofact *f, int job, int nrhs, double *x, int ldx) {
   int ni, r, k, j;
   if (job < 0 || job > 4) return -11;
   for (j = 0; j < nrhs; ++j) {
      double *xx = x + (size_t)j * ldx;
      if (job == 0 || job == 1)
         for (ni = 0; ni < f->nnodes; ++ni) {
            const onode *nd = &f->nodes[ni];
            for (k = 0; k < nd->nelim; ++k) {
               double xk = xx[row_index(f, ni, k)];
               if (xk == 0.0) continue;
               for (r = k + 1; r < nd->m; ++r) xx[row_index(f, ni, r)] -= nd->L[r + (size_t)k * nd->ldl] * xk;
            }
         }
      if (job == 0 || job == 2 || job == 4)
         for (ni = 0; ni < f->nnodes; ++ni) {
            const onode *nd = &f->nodes[ni];
            for (k = 0; k < nd->nelim;) {
               if (k + 1 == nd->nelim || isfinite(nd->d[2 * k + 2])) {
                  xx[nd->perm[k]] *= nd->d[2 * k];
                  k += 1;
               } else {
                  double d11 = nd->d[2 * k], d21 = nd->d[2 * k + 1], d22 = nd->d[2 * k + 3];
                  double x1 = xx[nd->perm[k]], x2 = xx[nd->perm[k + 1]];
                  xx[nd->perm[k]] = d11 * x1 + d21 * x2;
                  xx[nd->perm[k + 1]] = d21 * x1 + d22 * x2;
                  k += 2;
               }
            }
         }
      if (job == 0 || job == 3 || job == 4)
         for (ni = f->nnodes - 1; ni >= 0; --ni) {
            const onode *nd = &f->nodes[ni];
            for (k = nd->nelim - 1; k >= 0; --k) {
               double s = xx[row_index(f, ni, k)];
               for (r = k + 1; r < nd->m; ++r) s -= nd->L[r + (size_t)k * nd->ldl] * xx[row_index(f, ni, r)];
               xx[row_index(f, ni, k)] = s;
            }
         }
   }
   return 0;
}

/* piv_order[n] (0-based |.|-1 encoding avoided: returns position+1, negative for 2x2), d[2n] */
void oracle_enquire(const ofact *f, int *piv_order, double *d) {
   int ni, i, piv = 0;
   for (ni = 0; ni < f->nnodes; ++ni) {
      const onode *nd = &f->nodes[ni];
      for (i = 0; i < nd->nelim;) {
         if (i + 1 == nd->nelim || isfinite(nd->d[2 * i + 2])) {
            if (piv_order) piv_order[nd->perm[i]] = (piv + 1);
            if (d) { d[2 * piv] = nd->d[2 * i]; d[2 * piv + 1] = 0.0; }
            piv += 1; i += 1;
         } else {
            if (piv_order) { piv_order[nd->perm[i]] = -(piv + 1); piv_order[nd->perm[i + 1]] = -(piv + 2); }
            if (d) { d[2 * piv] = nd->d[2 * i]; d[2 * piv + 1] = nd->d[2 * i + 1]; d[2 * piv + 2] = nd->d[2 * i + 3]; d[2 * piv + 3] = 0.0; }
            piv += 2; i += 2;
         }
      }
   }
}

/* per-node results for fine-grained comparison with the CUDA engine */
void oracle_node_info(const ofact *f, int ni, int *m, int *n, int *nelim) {
   *m = f->nodes[ni].m; *n = f->nodes[ni].n; *nelim = f->nodes[ni].nelim;
}
const double *oracle_node_L(const ofact *f, int ni) { return f->nodes[ni].L; }
const double *oracle_node_d(const ofact *f, int ni) { return f->nodes[ni].d; }
const int *oracle_node_perm(const ofact *f, int ni) { return f->nodes[ni].perm; }
