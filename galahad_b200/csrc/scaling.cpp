// Matching-based scalings on the host (SSIDS options%scaling = 1 / 2, SLS scaling = -1 / -2): the Hungarian algorithm of
// MC64 as SPRAL carries it, and SPRAL's auction algorithm.  Restatement of
//    auction_scale_sym / auction_match / auction_match_core   src/spral/scaling.f90:269-309, 1504-1609, 1351-1489
//    hungarian_scale_sym / hungarian_wrapper        src/spral/scaling.f90:134-170, 597-801
//    hungarian_init_heurisitic / hungarian_match    src/spral/scaling.f90:810-929, 938-1194
//    heap_update / heap_pop / heap_delete           src/spral/scaling.f90:1203-1325
//    match_postproc (square case)                   src/spral/scaling.f90:1611-1640
//    half_to_full                                   src/spral/matrix_util.f90:3167-3302
// The maximum-product matching of |a_ij| is found as a minimum-sum assignment on c_ij = log max_i|a_ij| - log|a_ij| by
// shortest augmenting paths (Dijkstra with a binary heap and a bucket of ties); the duals u, v give the scaling
// s_i = exp((u_i + v_i - log colmax_i) / 2), for which |s_i a_ij s_j| <= 1 with equality on the matching.  Arrays are kept
// 1-based inside (index 0 unused) so that scan orders and tie-breaking are the reference's.  Checked entry by entry against
// the independent Python restatement oracle/scaling_oracle.py (tests/test_scaling_oracle.py).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

#include "analyse.hpp"

namespace b200 {
namespace {

constexpr double RINF = std::numeric_limits<double>::max();       // huge(rinf)
using i64 = int64_t;

// lower triangle by columns (1-based, row / a long enough for the expansion) -> both triangles
void half_to_full(int n, std::vector<int>& row, std::vector<i64>& ptr, std::vector<double>& a) {
   const i64 oldtau = ptr[n + 1] - 1;
   std::vector<int> iw(n + 1, 0);
   int ndiag = 0;
   for (int j = 1; j <= n; ++j) {
      const i64 i1 = ptr[j], i2 = ptr[j + 1] - 1;
      iw[j] += (int)(i2 - i1) + 1;
      for (i64 ii = i1; ii <= i2; ++ii) {
         const int i = row[ii];
         if (i != j) iw[i] += 1; else ndiag += 1;
      }
   }
   const i64 newtau = 2 * oldtau - ndiag;
   i64 ipkp1 = oldtau + 1, ckp1 = newtau + 1;
   for (int j = n; j >= 1; --j) {
      const i64 i1 = ptr[j];
      i64 i2 = ipkp1;
      const int lenk = (int)(i2 - i1);
      i64 jstart = ckp1;
      ipkp1 = i1;
      i2 -= 1;
      for (i64 ii = i2; ii >= i1; --ii) {
         jstart -= 1;
         a[jstart] = a[ii];
         row[jstart] = row[ii];
      }
      ptr[j] = jstart;
      ckp1 -= iw[j];
      iw[j] = lenk;
   }
   for (int j = n; j >= 1; --j) {
      const i64 i1 = ptr[j], i2 = ptr[j] + iw[j] - 1;
      for (i64 ii = i1; ii <= i2; ++ii) {
         const int i = row[ii];
         if (i == j) continue;
         ptr[i] -= 1;
         const i64 ipos = ptr[i];
         a[ipos] = a[ii];
         row[ipos] = j;
      }
   }
   ptr[n + 1] = newtau + 1;
}

// the value of idx has decreased: move it towards the root
void heap_update(int idx, std::vector<int>& q, const std::vector<double>& val, std::vector<int>& l) {
   int pos = l[idx];
   if (pos <= 1) { q[pos] = idx; return; }
   const double v = val[idx];
   while (pos > 1) {
      const int parent_pos = pos / 2, parent_idx = q[parent_pos];
      if (v >= val[parent_idx]) break;
      q[pos] = parent_idx;
      l[parent_idx] = pos;
      pos = parent_pos;
   }
   q[pos] = idx;
   l[idx] = pos;
}

void heap_delete(int pos0, int& qlen, std::vector<int>& q, const std::vector<double>& d, std::vector<int>& l) {
   if (qlen == pos0) { qlen -= 1; return; }
   const int idx = q[qlen];
   const double v = d[idx];
   qlen -= 1;
   int pos = pos0;
   if (pos > 1) {
      for (;;) {
         const int parent = pos / 2, qk = q[parent];
         if (v >= d[qk]) break;
         q[pos] = qk;
         l[qk] = pos;
         pos = parent;
         if (pos <= 1) break;
      }
   }
   q[pos] = idx;
   l[idx] = pos;
   if (pos != pos0) return;
   for (;;) {
      int child = 2 * pos;
      if (child > qlen) break;
      double dk = d[q[child]];
      if (child < qlen) {
         const double dr = d[q[child + 1]];
         if (dk > dr) { child += 1; dk = dr; }
      }
      if (v <= dk) break;
      const int qk = q[child];
      q[pos] = qk;
      l[qk] = pos;
      pos = child;
   }
   q[pos] = idx;
   l[idx] = pos;
}

void init_heuristic(int m, int n, const std::vector<i64>& ptr, const std::vector<int>& row, const std::vector<double>& val, int& num,
                    std::vector<int>& iperm, std::vector<i64>& jperm, std::vector<double>& dualu, std::vector<double>& d,
                    std::vector<i64>& l, std::vector<i64>& search_from) {
   for (int i = 1; i <= m; ++i) { dualu[i] = RINF; l[i] = 0; }
   for (int j = 1; j <= n; ++j)
      for (i64 k = ptr[j]; k < ptr[j + 1]; ++k) {
         const int i = row[k];
         if (val[k] > dualu[i]) continue;
         dualu[i] = val[k];
         iperm[i] = j;
         l[i] = k;
      }
   for (int i = 1; i <= m; ++i) {
      const int j = iperm[i];
      if (j == 0) continue;
      iperm[i] = 0;
      if (jperm[j] != 0) continue;
      if (ptr[j + 1] - ptr[j] > m / 10 && m > 50) continue;       // no cheap assignment from dense columns
      num += 1;
      iperm[i] = j;
      jperm[j] = l[i];
   }
   if (num == std::min(m, n)) return;
   for (int j = 1; j <= n; ++j) { d[j] = 0.0; search_from[j] = ptr[j]; }
   for (int j = 1; j <= n; ++j) {
      if (jperm[j] != 0) continue;
      if (ptr[j] > ptr[j + 1] - 1) continue;
      int i0 = row[ptr[j]];
      double vj = val[ptr[j]] - dualu[i0];
      i64 k0 = ptr[j];
      for (i64 k = ptr[j] + 1; k < ptr[j + 1]; ++k) {
         const int i = row[k];
         const double di = val[k] - dualu[i];
         if (di > vj) continue;
         if (di == vj && di != RINF) {
            if (iperm[i] != 0 || iperm[i0] == 0) continue;        // ties: first unmatched, then first matched row
         }
         vj = di; i0 = i; k0 = k;
      }
      d[j] = vj;
      if (iperm[i0] == 0) {
         num += 1;
         jperm[j] = k0;
         iperm[i0] = j;
         search_from[j] = k0 + 1;
         continue;
      }
      bool augmented = false;
      for (i64 k = k0; k < ptr[j + 1] && !augmented; ++k) {
         const int i = row[k];
         if (val[k] - dualu[i] > vj) continue;
         const int jj = iperm[i];
         for (i64 kk = search_from[jj]; kk < ptr[jj + 1]; ++kk) {
            const int ii = row[kk];
            if (iperm[ii] > 0) continue;
            if (val[kk] - dualu[ii] <= d[jj]) {
               jperm[jj] = kk;
               iperm[ii] = jj;
               search_from[jj] = kk + 1;
               num += 1;
               jperm[j] = k;
               iperm[i] = j;
               search_from[j] = k + 1;
               augmented = true;
               break;
            }
         }
         if (!augmented) search_from[jj] = ptr[jj + 1];
      }
   }
}

// minimum-sum assignment (Duff and Koster); iperm[i] = column matched to row i (negative for the completion of a
// structurally singular matrix), duals u (rows), v (columns)
void hungarian_match(int m, int n, const std::vector<i64>& ptr, const std::vector<int>& row, const std::vector<double>& val,
                     std::vector<int>& iperm, int& num, std::vector<double>& dualu, std::vector<double>& dualv) {
   std::vector<i64> jperm(n + 1, 0), out(n + 1, 0), longwork(m + 1, 0);
   std::vector<int> pr(n + 1, 0), q(m + 2, 0), l(m + 1, 0);
   std::vector<double> d(std::max(m, n) + 1, 0.0);
   iperm.assign(m + 1, 0);
   dualu.assign(m + 1, 0.0);
   dualv.assign(n + 1, 0.0);
   num = 0;
   init_heuristic(m, n, ptr, row, val, num, iperm, jperm, dualu, d, longwork, out);
   if (num != std::min(m, n)) {
      for (int i = 1; i <= m; ++i) { d[i] = RINF; l[i] = 0; }
      i64 isp = -1;
      int jsp = -1;
      for (int jord = 1; jord <= n; ++jord) {
         if (jperm[jord] != 0) continue;
         double dmin = RINF, csp = RINF;
         int qlen = 0, low = m + 1, up = m + 1;
         int j = jord;
         pr[j] = -1;
         for (i64 k = ptr[j]; k < ptr[j + 1]; ++k) {
            const int i = row[k];
            const double dnew = val[k] - dualu[i];
            if (dnew >= csp) continue;
            if (iperm[i] == 0) { csp = dnew; isp = k; jsp = j; }
            else {
               if (dnew < dmin) dmin = dnew;
               d[i] = dnew;
               qlen += 1;
               longwork[qlen] = k;
            }
         }
         int q0 = qlen;
         qlen = 0;
         for (int kk = 1; kk <= q0; ++kk) {
            const i64 k = longwork[kk];
            const int i = row[k];
            if (csp <= d[i]) { d[i] = RINF; continue; }
            if (d[i] <= dmin) { low -= 1; q[low] = i; l[i] = low; }
            else { qlen += 1; l[i] = qlen; heap_update(i, q, d, l); }
            const int jj = iperm[i];
            out[jj] = k;
            pr[jj] = j;
         }
         const int npass = num;
         for (int jdum = 1; jdum <= npass; ++jdum) {
            if (low == up) {
               if (qlen == 0) break;
               int i = q[1];
               if (d[i] >= csp) break;
               dmin = d[i];
               while (qlen > 0) {
                  i = q[1];
                  if (d[i] > dmin) break;
                  heap_delete(1, qlen, q, d, l);
                  low -= 1;
                  q[low] = i;
                  l[i] = low;
               }
            }
            q0 = q[up - 1];
            const double dq0 = d[q0];
            if (dq0 >= csp) break;
            up -= 1;
            j = iperm[q0];
            const double vj = dq0 - val[jperm[j]] + dualu[q0];
            for (i64 k = ptr[j]; k < ptr[j + 1]; ++k) {
               const int i = row[k];
               if (l[i] >= up) continue;
               const double dnew = vj + val[k] - dualu[i];
               if (dnew >= csp) continue;
               if (iperm[i] == 0) { csp = dnew; isp = k; jsp = j; }
               else {
                  const double di = d[i];
                  if (di <= dnew) continue;
                  if (l[i] >= low) continue;
                  d[i] = dnew;
                  if (dnew <= dmin) {
                     const int lpos = l[i];
                     if (lpos != 0) heap_delete(lpos, qlen, q, d, l);
                     low -= 1;
                     q[low] = i;
                     l[i] = low;
                  } else {
                     if (l[i] == 0) { qlen += 1; l[i] = qlen; }
                     heap_update(i, q, d, l);
                  }
                  const int jj = iperm[i];
                  out[jj] = k;
                  pr[jj] = j;
               }
            }
         }
         if (csp != RINF) {
            num += 1;
            int i = row[isp];
            iperm[i] = jsp;
            jperm[jsp] = isp;
            j = jsp;
            for (int jdum = 1; jdum <= num; ++jdum) {
               const int jj = pr[j];
               if (jj == -1) break;
               const i64 k = out[j];
               i = row[k];
               iperm[i] = jj;
               jperm[jj] = k;
               j = jj;
            }
            for (int kk = up; kk <= m; ++kk) {
               i = q[kk];
               dualu[i] = dualu[i] + d[i] - csp;
            }
         }
         for (int kk = low; kk <= m; ++kk) { const int i = q[kk]; d[i] = RINF; l[i] = 0; }
         for (int kk = 1; kk <= qlen; ++kk) { const int i = q[kk]; d[i] = RINF; l[i] = 0; }
      }
   }
   for (int j = 1; j <= n; ++j) {
      const i64 k = jperm[j];
      dualv[j] = k != 0 ? val[k] - dualu[row[k]] : 0.0;
   }
   for (int i = 1; i <= m; ++i) if (iperm[i] == 0) dualu[i] = 0.0;
   if (num == std::min(m, n)) return;
   // structurally singular: complete iperm with negative entries
   std::fill(jperm.begin(), jperm.end(), 0);
   int k = 0;
   for (int i = 1; i <= m; ++i) {
      if (iperm[i] == 0) { k += 1; out[k] = i; }
      else jperm[iperm[i]] = i;
   }
   k = 0;
   for (int j = 1; j <= n; ++j) {
      if (jperm[j] != 0) continue;
      k += 1;
      iperm[(int)out[k]] = -j;
   }
}

}  // namespace

int hungarian_scale_sym(int n, const int64_t* ptr0, const int* row0, const double* val0, double* scaling, bool scale_if_singular,
                        int* match_out, int* matched_out) {
   int flag = 0;
   const i64 nz = n > 0 ? ptr0[n] : 0;
   std::vector<i64> ptr2(n + 2, 0);
   std::vector<int> row2(2 * nz + 2, 0);
   std::vector<double> val2(2 * nz + 2, 0.0);
   // drop explicit zeros, log of the absolute values
   i64 k = 1;
   for (int i = 1; i <= n; ++i) {
      ptr2[i] = k;
      for (i64 j = ptr0[i - 1]; j < ptr0[i]; ++j) {
         if (val0[j] == 0.0) continue;
         row2[k] = row0[j] + 1;
         val2[k] = std::log(std::fabs(val0[j]));
         k += 1;
      }
   }
   ptr2[n + 1] = k;
   half_to_full(n, row2, ptr2, val2);
   std::vector<double> cmax(n + 1, 0.0);
   for (int i = 1; i <= n; ++i) {
      double colmax = -RINF;                                     // maxval of an empty section
      for (i64 j = ptr2[i]; j < ptr2[i + 1]; ++j) colmax = std::max(colmax, val2[j]);
      cmax[i] = colmax;
      for (i64 j = ptr2[i]; j < ptr2[i + 1]; ++j) val2[j] = colmax - val2[j];
   }
   std::vector<int> match;
   std::vector<double> dualu, dualv;
   int matched = 0;
   hungarian_match(n, n, ptr2, row2, val2, match, matched, dualu, dualv);
   if (matched != n) flag = scale_if_singular ? 1 : -2;         // WARNING_SINGULAR / ERROR_SINGULAR
   std::vector<double> rscaling(n + 1, 0.0), cscaling(n + 1, 0.0);
   if (matched == n) {
      for (int i = 1; i <= n; ++i) { rscaling[i] = dualu[i]; cscaling[i] = dualv[i] - cmax[i]; }
      // match_postproc, square: balance the row and column scalings
      double rsum = 0.0, csum = 0.0;
      for (int i = 1; i <= n; ++i) rsum += rscaling[i];
      for (int i = 1; i <= n; ++i) csum += cscaling[i];
      const double adjust = (n > 0) ? (rsum / n - csum / n) / 2 : 0.0;
      for (int i = 1; i <= n; ++i) { rscaling[i] -= adjust; cscaling[i] += adjust; }
      for (int i = 1; i <= n; ++i) scaling[i - 1] = std::exp((rscaling[i] + cscaling[i]) / 2);
      if (match_out) for (int i = 1; i <= n; ++i) match_out[i - 1] = match[i];
      if (matched_out) *matched_out = matched;
      return flag;
   }
   // structurally rank deficient: the matching of the full-rank submatrix, Duff and Pralet's correction for the rest
   std::vector<int> old_to_new(n + 1, 0), new_to_old(n + 1, 0);
   {
      int j = matched + 1, kk = 0;
      for (int i = 1; i <= n; ++i) {
         if (match[i] < 0) { old_to_new[i] = -j; j += 1; }
         else { kk += 1; old_to_new[i] = kk; new_to_old[kk] = i; }
      }
   }
   i64 ne = 0;
   int kk = 0;
   i64 j2 = 1;
   ptr2[1] = 1;
   for (int i = 1; i <= n; ++i) {
      const i64 j1 = j2;
      j2 = ptr2[i + 1];
      if (match[i] < 0) continue;
      kk += 1;
      for (i64 j = j1; j < j2; ++j) {
         const int jj = row2[j];
         if (match[jj] < 0) continue;
         ne += 1;
         row2[ne] = old_to_new[jj];
         val2[ne] = val2[j];
      }
      ptr2[kk + 1] = ne + 1;
   }
   const int nn = kk;
   std::vector<int> cperm;
   int matched2 = 0;
   hungarian_match(nn, nn, ptr2, row2, val2, cperm, matched2, dualu, dualv);
   const double NEG = -RINF;
   for (int i = 1; i <= n; ++i) {
      const int j = old_to_new[i];
      rscaling[i] = j < 0 ? NEG : (dualu[j] + dualv[j] - cmax[i]) / 2;
   }
   std::fill(match.begin(), match.end(), -1);
   for (int i = 1; i <= nn; ++i) match[new_to_old[i]] = cperm[i];
   for (int i = 1; i <= n; ++i) if (match[i] == -1) match[i] = old_to_new[i];
   std::vector<double> cscale(rscaling);
   for (int i = 1; i <= n; ++i)
      for (i64 j = ptr0[i - 1]; j < ptr0[i]; ++j) {
         const int r = row0[j] + 1;
         const double lv = std::log(std::fabs(val0[j]));          // -inf for an explicit zero: never the maximum
         if (cscale[i] == NEG && cscale[r] != NEG) rscaling[i] = std::max(rscaling[i], lv + rscaling[r]);
         if (cscale[r] == NEG && cscale[i] != NEG) rscaling[r] = std::max(rscaling[r], lv + rscaling[i]);
      }
   for (int i = 1; i <= n; ++i) {
      if (cscale[i] != NEG) continue;
      rscaling[i] = rscaling[i] == NEG ? 0.0 : -rscaling[i];
   }
   for (int i = 1; i <= n; ++i) scaling[i - 1] = flag == -2 ? 1.0 : std::exp((rscaling[i] + rscaling[i]) / 2);
   if (match_out) for (int i = 1; i <= n; ++i) match_out[i - 1] = match[i];
   if (matched_out) *matched_out = matched2;
   return flag;
}

// ------------------------------------------------------------------------------------------------------------------
// Matching-based ordering (SSIDS options%ordering = 2): restatement of
//    match_order_metis / mo_scale / mo_match / mo_split     src/spral/match_order.f90:135-208, 406-485, 495-629, 220-396
//    expand_matrix                                           src/ssids/anal.f90:87-138
// A maximum-product matching of |a_ij| (hungarian_match above) is broken into 1- and 2-cycles, matched pairs (i, j) are
// merged into one vertex of a compressed graph, the compressed graph is ordered, and each pair takes two consecutive
// places in the pivot order: the large off-diagonal a_ij is then at hand as a 2x2 pivot when the diagonals are small or
// zero (KKT systems).  The duals of the matching give the scaling that SSIDS keeps for options%scaling = 3.  The
// compressed graph is ordered by this library's nested dissection / minimum degree (ordering.cpp) where the reference
// calls METIS, a package outside its tree.
namespace {

// lower triangle by columns (0-based in) -> both triangles, 1-based, entries of a column in the reference's order
void expand_matrix(int n, const int64_t* ptr0, const int* row0, const double* val0, std::vector<i64>& aptr, std::vector<int>& arow,
                   std::vector<double>& aval) {
   const i64 nz = n > 0 ? ptr0[n] : 0;
   aptr.assign(n + 2, 0);
   arow.assign(2 * nz + 2, 0);
   aval.assign(2 * nz + 2, 0.0);
   for (int j = 1; j <= n; ++j)
      for (i64 kk = ptr0[j - 1]; kk < ptr0[j]; ++kk) {
         const int i = row0[kk] + 1;
         aptr[i] += 1;
         if (j == i) continue;
         aptr[j] += 1;
      }
   for (int j = 2; j <= n; ++j) aptr[j] += aptr[j - 1];
   aptr[n + 1] = aptr[n] + 1;
   for (int j = 1; j <= n; ++j)
      for (i64 kk = ptr0[j - 1]; kk < ptr0[j]; ++kk) {
         const int i = row0[kk] + 1;
         const double atemp = val0[kk];
         const i64 ipos = aptr[i];
         arow[ipos] = j;
         aval[ipos] = atemp;
         aptr[i] = ipos - 1;
         if (j == i) continue;
         const i64 jpos = aptr[j];
         arow[jpos] = i;
         aval[jpos] = atemp;
         aptr[j] = jpos - 1;
      }
   for (int j = 1; j <= n; ++j) aptr[j] += 1;
}

// mo_match: matching and log-scaling of the full matrix of absolute values (explicit zeros gone); ptr2 / row2 / val2 are
// overwritten.  Returns 0, or 1 when the matrix is structurally singular (unmatched rows: perm = -1, scale = -huge).
int mo_match(int n, std::vector<int>& row2, std::vector<i64>& ptr2, std::vector<double>& val2, std::vector<double>& scale,
             std::vector<int>& perm) {
   std::vector<double> cmax(n + 1, 0.0);
   for (int i = 1; i <= n; ++i) {
      double colmax = 0.0;
      for (i64 j = ptr2[i]; j < ptr2[i + 1]; ++j) colmax = std::max(colmax, val2[j]);
      if (colmax != 0.0) colmax = std::log(colmax);
      cmax[i] = colmax;
   }
   for (int i = 1; i <= n; ++i)
      for (i64 j = ptr2[i]; j < ptr2[i + 1]; ++j) val2[j] = cmax[i] - std::log(val2[j]);
   std::vector<int> cperm;
   std::vector<double> du, dv;
   int rank = 0;
   hungarian_match(n, n, ptr2, row2, val2, cperm, rank, du, dv);
   scale.assign(n + 1, 0.0);
   perm.assign(n + 1, 0);
   if (rank == n) {
      for (int i = 1; i <= n; ++i) scale[i] = (du[i] + dv[i] - cmax[i]) / 2;
      for (int i = 1; i <= n; ++i) perm[i] = cperm[i];
      return 0;
   }
   // structurally singular: squeeze the unmatched rows and columns out, match what is left again
   std::vector<int> old_to_new(n + 1, 0), new_to_old(n + 1, 0);
   int k = 0;
   for (int i = 1; i <= n; ++i) {
      if (cperm[i] < 0) old_to_new[i] = -1;
      else { k += 1; old_to_new[i] = k; new_to_old[k] = i; }
   }
   i64 nne = 0;
   k = 0;
   i64 j2 = 1;
   ptr2[1] = 1;
   for (int i = 1; i <= n; ++i) {
      const i64 j1 = j2;
      j2 = ptr2[i + 1];
      if (cperm[i] < 0) continue;
      k += 1;
      for (i64 jl = j1; jl < j2; ++jl) {
         const int jj = row2[jl];
         if (cperm[jj] < 0) continue;
         nne += 1;
         row2[nne] = old_to_new[jj];
         val2[nne] = val2[jl];
      }
      ptr2[k + 1] = nne + 1;
   }
   const int nn = k;
   std::vector<int> cperm2;
   int rank2 = 0;
   hungarian_match(nn, nn, ptr2, row2, val2, cperm2, rank2, du, dv);
   for (int i = 1; i <= n; ++i) {
      const int j = old_to_new[i];
      scale[i] = j < 0 ? -RINF : (du[j] + dv[j] - cmax[i]) / 2;
   }
   for (int i = 1; i <= n; ++i) perm[i] = -1;
   for (int i = 1; i <= nn; ++i) perm[new_to_old[i]] = new_to_old[cperm2[i]];
   return 1;
}

}  // namespace

int match_order(int n, const int64_t* ptr0, const int* row0, const double* val0, int kind, int* order, double* scale_out,
                int* pairs_out, std::vector<int64_t>* cptr_out, std::vector<int>* crow_out) {
   if (n < 0) return SSIDS_ERROR_A_N_OOR;
   if (n == 0) return 0;
   std::vector<i64> aptr;
   std::vector<int> arow;
   std::vector<double> aval;
   expand_matrix(n, ptr0, row0, val0, aptr, arow, aval);
   // explicit zeros out, absolute values (match_order_metis; mo_scale repeats the squeeze on its own copy)
   std::vector<i64> ptr2(n + 2, 0);
   std::vector<int> row2(arow.size(), 0);
   std::vector<double> val2(aval.size(), 0.0);
   i64 k = 1;
   for (int i = 1; i <= n; ++i) {
      ptr2[i] = k;
      for (i64 j = aptr[i]; j < aptr[i + 1]; ++j) {
         if (aval[j] == 0.0) continue;
         row2[k] = arow[j];
         val2[k] = std::fabs(aval[j]);
         k += 1;
      }
   }
   ptr2[n + 1] = k;
   std::vector<double> scale;
   std::vector<int> cperm;
   int flag;
   {
      std::vector<i64> p(ptr2);
      std::vector<int> r(row2);
      std::vector<double> v(val2);
      flag = mo_match(n, r, p, v, scale, cperm);
   }
   // mo_split: cycles of the matching -> singletons and pairs
   std::vector<int> iwork(n + 1, 0);
   for (int i = 1; i <= n; ++i) {
      if (iwork[i] != 0) continue;
      int j = i;
      for (;;) {
         if (cperm[j] == -1) { iwork[j] = -2; break; }          // unmatched
         if (cperm[j] == i) { iwork[j] = -1; break; }           // singleton, or the odd one of a cycle
         const int jj = cperm[j];
         iwork[j] = jj;
         iwork[jj] = j;
         j = cperm[jj];
         if (j == i) break;
      }
   }
   cperm = iwork;
   std::vector<int> old_to_new(n + 1, 0), new_to_old(n + 1, 0);
   int kc = 1;
   for (int i = 1; i <= n; ++i) {
      const int j = cperm[i];
      if (j < i && j > 0) continue;
      old_to_new[i] = kc;
      new_to_old[kc] = i;                                        // the first of a pair
      if (j > 0) old_to_new[j] = kc;
      kc += 1;
   }
   const int ncomp_matched = kc - 1;
   // pattern of the compressed matrix, then its lower triangle
   std::vector<i64> ptr3(n + 2, 0);
   std::vector<int> row3(row2.size(), 0);
   std::fill(iwork.begin(), iwork.end(), 0);
   ptr3[1] = 1;
   int ncomp = 1;
   i64 jj = 1;
   for (int i = 1; i <= n; ++i) {
      const int j = cperm[i];
      if (j < i && j > 0) continue;
      for (int pass = 0; pass < 2; ++pass) {
         const int c = pass == 0 ? i : j;
         if (c <= 0) continue;
         for (i64 kl = ptr2[c]; kl < ptr2[c + 1]; ++kl) {
            const int krow = old_to_new[row2[kl]];
            if (iwork[krow] == i) continue;
            if (krow > ncomp_matched) continue;
            row3[jj] = krow;
            jj += 1;
            iwork[krow] = i;
         }
      }
      ptr3[ncomp + 1] = jj;
      ncomp += 1;
   }
   ncomp -= 1;
   {
      i64 out = 1, j1 = 1;
      for (int i = 1; i <= ncomp; ++i) {
         const i64 j2 = ptr3[i + 1];
         for (i64 kk = j1; kk < j2; ++kk) {
            const int krow = row3[kk];
            if (krow < i) continue;
            row3[out] = krow;
            out += 1;
         }
         ptr3[i + 1] = out;
         j1 = j2;
      }
   }
   // order of the compressed matrix (0-based for the ordering routines)
   std::vector<int64_t> cp(ncomp + 1);
   std::vector<int> cr((size_t)(ptr3[ncomp + 1] - 1));
   for (int i = 0; i <= ncomp; ++i) cp[i] = ptr3[i + 1] - 1;
   for (size_t e = 0; e < cr.size(); ++e) cr[e] = row3[e + 1] - 1;
   std::vector<int> cord(ncomp, 0);
   const int orc = kind == 1 ? amd_order(ncomp, cp.data(), cr.data(), cord.data()) : nd_order(ncomp, cp.data(), cr.data(), cord.data());
   if (orc < 0) return orc;
   std::vector<int> inv(ncomp + 1, 0);
   for (int i = 1; i <= ncomp; ++i) inv[cord[i - 1] + 1] = i;
   int pos = 1;
   for (int i = 1; i <= ncomp; ++i) {
      int j = new_to_old[inv[i]];
      order[j - 1] = pos - 1;
      pos += 1;
      if (cperm[j] > 0) {
         j = cperm[j];
         order[j - 1] = pos - 1;
         pos += 1;
      }
   }
   if (scale_out) for (int i = 1; i <= n; ++i) scale_out[i - 1] = std::exp(scale[i]);
   if (pairs_out) for (int i = 1; i <= n; ++i) pairs_out[i - 1] = cperm[i];
   if (cptr_out) *cptr_out = cp;
   if (crow_out) *crow_out = cr;
   return flag;
}


// ------------------------------------------------------------------------------------------------------------------
// Auction algorithm (options%scaling = 2): each unmatched column bids for its best row, prices (row duals) rise by the
// margin over the second best plus eps; an approximate maximum-product matching in a bounded number of sweeps.
// Options are the reference's defaults (auction_options, scaling.f90:33-38): max_iterations 30000, max_unchanged
// (10, 100, 100), min_proportion (0.90, 0, 0), eps_initial 0.01 -- the last two are default REALs there, hence the floats.
namespace {

void auction_match_core(int m, int n, const std::vector<i64>& ptr, const std::vector<int>& row, const std::vector<double>& val,
                        std::vector<int>& match, std::vector<double>& dualu, std::vector<double>& dualv, int& iterations,
                        int& unmatchable) {
   const int max_iterations = 30000;
   const int max_unchanged[3] = {10, 100, 100};
   const float min_proportion[3] = {0.90f, 0.0f, 0.0f};
   const float eps_initial = 0.01f;
   const double NEGH = -RINF;
   std::vector<int> owner(m + 1, 0), next(n + 1, 0);
   match.assign(n + 1, 0);
   dualu.assign(m + 1, 0.0);
   const int minmn = std::min(m, n);
   int unmatched = minmn, prev = -1, nunchanged = 0, tail = n;
   for (int i = 1; i <= n; ++i) next[i] = i;
   double eps = (double)eps_initial;
   unmatchable = 0;
   int itr = 1;
   for (; itr <= max_iterations; ++itr) {
      if (unmatched == 0) break;
      if (unmatched != prev) nunchanged = 0;
      prev = unmatched;
      nunchanged += 1;
      const float frac = (float)(minmn - unmatched) / (float)minmn;
      bool stop = false;
      for (int k = 0; k < 3; ++k) if (nunchanged >= max_unchanged[k] && frac >= min_proportion[k]) stop = true;
      if (stop) break;
      eps = std::min(1.0, eps + 1.0 / (n + 1));
      int insert = 0;
      for (int cptr = 1; cptr <= tail; ++cptr) {
         const int col = next[cptr];
         if (match[col] != 0) continue;
         if (ptr[col] == ptr[col + 1]) continue;
         i64 j = ptr[col];
         int bestr = row[j];
         double bestu = val[j] - dualu[bestr], bestv = NEGH;
         for (j = ptr[col] + 1; j < ptr[col + 1]; ++j) {
            const double u = val[j] - dualu[row[j]];
            if (u > bestu) { bestv = bestu; bestr = row[j]; bestu = u; }
            else if (u > bestv) bestv = u;
         }
         if (bestv == NEGH) bestv = 0.0;
         if (bestu > 0) {
            dualu[bestr] = dualu[bestr] + bestu - bestv + eps;
            dualv[col] = bestv - eps;
            match[col] = bestr;
            unmatched -= 1;
            const int k = owner[bestr];
            owner[bestr] = col;
            if (k != 0) { match[k] = 0; unmatched += 1; insert += 1; next[insert] = k; }
         } else {
            match[col] = -1;
            unmatched -= 1;
            unmatchable += 1;
         }
      }
      tail = insert;
   }
   iterations = itr - 1;
   for (int j = 1; j <= n; ++j) if (match[j] == -1) match[j] = 0;
}

}  // namespace

int auction_scale_sym(int n, const int64_t* ptr0, const int* row0, const double* val0, double* scaling, int* match_out,
                      int* matched_out, int* iterations_out) {
   const i64 nz = n > 0 ? ptr0[n] : 0;
   std::vector<i64> ptr2(n + 2, 0);
   std::vector<int> row2(2 * nz + 2, 0);
   std::vector<double> val2(2 * nz + 2, 0.0);
   i64 k = 1;
   for (int i = 1; i <= n; ++i) {
      ptr2[i] = k;
      for (i64 j = ptr0[i - 1]; j < ptr0[i]; ++j) {
         if (val0[j] == 0.0) continue;
         row2[k] = row0[j] + 1;
         val2[k] = std::log(std::fabs(val0[j]));
         k += 1;
      }
   }
   ptr2[n + 1] = k;
   half_to_full(n, row2, ptr2, val2);
   std::vector<double> cmax(n + 1, 0.0);
   for (int i = 1; i <= n; ++i) {
      if (ptr2[i + 1] <= ptr2[i]) { cmax[i] = 0.0; continue; }
      double colmax = -RINF;
      for (i64 j = ptr2[i]; j < ptr2[i + 1]; ++j) colmax = std::max(colmax, val2[j]);
      cmax[i] = colmax;
      for (i64 j = ptr2[i]; j < ptr2[i + 1]; ++j) val2[j] = colmax - val2[j];
   }
   const i64 nz2 = ptr2[n + 1] - 1;
   double maxentry = -RINF;
   for (i64 j = 1; j <= nz2; ++j) maxentry = std::max(maxentry, val2[j]);
   maxentry = 2 * maxentry + 1;            // prefers matchings of high cardinality
   for (i64 j = 1; j <= nz2; ++j) val2[j] = maxentry - val2[j];
   std::vector<double> rscaling, cscaling(n + 1, 0.0);
   for (int i = 1; i <= n; ++i) cscaling[i] = -cmax[i];
   std::vector<int> cmatch;
   int iterations = 0, unmatchable = 0;
   auction_match_core(n, n, ptr2, row2, val2, cmatch, rscaling, cscaling, iterations, unmatchable);
   int matched = 0;
   for (int j = 1; j <= n; ++j) if (cmatch[j] != 0) ++matched;
   for (int i = 1; i <= n; ++i) { rscaling[i] = -rscaling[i] + maxentry; cscaling[i] = -cscaling[i] - cmax[i]; }
   if (match_out) {
      for (int i = 0; i < n; ++i) match_out[i] = 0;
      for (int j = 1; j <= n; ++j) if (cmatch[j] != 0) match_out[cmatch[j] - 1] = j;
   }
   // match_postproc, square
   double rsum = 0.0, csum = 0.0;
   for (int i = 1; i <= n; ++i) rsum += rscaling[i];
   for (int i = 1; i <= n; ++i) csum += cscaling[i];
   const double adjust = n > 0 ? (rsum / n - csum / n) / 2 : 0.0;
   for (int i = 1; i <= n; ++i) scaling[i - 1] = std::exp(((rscaling[i] - adjust) + (cscaling[i] + adjust)) / 2);
   if (matched_out) *matched_out = matched;
   if (iterations_out) *iterations_out = iterations;
   return 0;
}

}  // namespace b200
