// Symbolic analysis (CPU).  See analyse.hpp for the contract and the reference lines followed.
// Written 0-based; the virtual root of the elimination tree is vertex n, of the assembly tree
// node nnodes.  Tie-breaking (child order, stable decreasing-cc sort, numbering of merged
// vertices) follows the reference so that order/sptr/sparent/rptr/rlist/nptr/nlist agree
// entry for entry with SSIDS's analyse on the same input.
#include "analyse.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <limits>

namespace b200 {
namespace {

// Lower CSC -> full symmetric pattern.  (anal.f90:37-80 expand_pattern)
void expand_pattern(int n, const int64_t* ptr, const int* row, std::vector<int64_t>& aptr,
                    std::vector<int>& arow) {
   int64_t nz = ptr[n];
   aptr.assign(n + 1, 0);
   arow.assign(2 * nz, 0);
   // counts per column
   std::vector<int64_t> cnt(n, 0);
   for (int j = 0; j < n; ++j)
      for (int64_t kk = ptr[j]; kk < ptr[j + 1]; ++kk) {
         int i = row[kk];
         cnt[i]++;
         if (i != j) cnt[j]++;
      }
   // end pointers; the reference fills each column from its end backwards
   std::vector<int64_t> endp(n);
   int64_t acc = 0;
   for (int j = 0; j < n; ++j) {
      acc += cnt[j];
      endp[j] = acc - 1;
   }
   aptr[n] = acc;
   for (int j = 0; j < n; ++j)
      for (int64_t kk = ptr[j]; kk < ptr[j + 1]; ++kk) {
         int i = row[kk];
         arow[endp[i]--] = j;
         if (i == j) continue;
         arow[endp[j]--] = i;
      }
   for (int j = 0; j < n; ++j) aptr[j] = endp[j] + 1;
   aptr.resize(n + 1);
   aptr[n] = acc;
   arow.resize(acc);
}

// Liu's elimination tree with path compression.  (core_analyse.f90:173-222)
void find_etree(int n, const int64_t* ptr, const int* row, const int* perm, const int* invp,
                std::vector<int>& parent) {
   std::vector<int> vforest(n, n);
   parent.assign(n, n);
   for (int piv = 0; piv < n; ++piv) {
      int col = invp[piv];
      for (int64_t i = ptr[col]; i < ptr[col + 1]; ++i) {
         int j = perm[row[i]];
         if (j >= piv) continue;
         int k = j;
         while (vforest[k] < piv) {
            int l = vforest[k];
            vforest[k] = piv;
            k = l;
         }
         if (vforest[k] == piv) continue;
         parent[k] = piv;
         vforest[k] = piv;
      }
      parent[piv] = n;
   }
}

// Depth-first relabelling; empty columns go last.  (core_analyse.f90:233-357)
void find_postorder(int n, int& realn, const int64_t* ptr, std::vector<int>& perm,
                    std::vector<int>& invp, std::vector<int>& parent) {
   realn = n;
   std::vector<int> chead(n + 1, -1), cnext(n + 1, -1);
   for (int i = n - 1; i >= 0; --i) {
      int j = parent[i];
      cnext[i] = chead[j];
      chead[j] = i;
   }
   std::vector<int> map(n + 1), stack;
   stack.reserve(n + 1);
   stack.push_back(n);
   int id = n;
   while (!stack.empty()) {
      int node = stack.back();
      stack.pop_back();
      map[node] = id--;
      if (node == n) {
         for (int i = chead[node]; i != -1; i = cnext[i])
            if (ptr[invp[i] + 1] - ptr[invp[i]] != 0) stack.push_back(i);
         for (int i = chead[node]; i != -1; i = cnext[i])
            if (ptr[invp[i] + 1] - ptr[invp[i]] == 0) {
               --realn;
               stack.push_back(i);
            }
      } else {
         for (int i = chead[node]; i != -1; i = cnext[i]) stack.push_back(i);
      }
   }
   std::vector<int> tmp(invp.begin(), invp.begin() + n);
   for (int i = 0; i < n; ++i) invp[map[i]] = tmp[i];
   for (int i = 0; i < n; ++i) perm[invp[i]] = i;
   for (int i = 0; i < n; ++i) tmp[i] = map[parent[i]];
   for (int i = 0; i < n; ++i) parent[map[i]] = tmp[i];
}

inline int uf_find(std::vector<int>& vforest, int u) {
   int prev = -1, current = u;
   while (vforest[current] != -1) {
      prev = current;
      current = vforest[current];
      if (vforest[current] != -1) vforest[prev] = vforest[current];
   }
   return current;
}

// Gilbert-Ng-Peyton column counts.  (core_analyse.f90:387-498)
void find_col_counts(int n, const int64_t* ptr, const int* row, const int* perm, const int* invp,
                     const std::vector<int>& parent, std::vector<int>& cc) {
   std::vector<int> first(n + 1);
   cc.assign(n + 1, 0);
   for (int i = 0; i <= n; ++i) first[i] = i;
   for (int i = 0; i < n; ++i) {
      int par = parent[i];
      first[par] = std::min(first[i], first[par]);
      cc[i] = (first[i] == i) ? 1 : 0;
   }
   cc[n] = n + 1;
   std::vector<int> vforest(n + 1, -1), last_p(n + 1, -1), last_nbr(n + 1, -1);
   for (int piv = 0; piv < n; ++piv) {
      int col = invp[piv];
      for (int64_t ii = ptr[col]; ii < ptr[col + 1]; ++ii) {
         int u = perm[row[ii]];
         if (u <= piv) continue;
         if (first[piv] > last_nbr[u]) {
            cc[piv] += 1;
            int pp = last_p[u];
            if (pp != -1) {
               int lca = uf_find(vforest, pp);
               cc[lca] -= 1;
            }
            last_p[u] = piv;
         }
         last_nbr[u] = piv;
      }
      int par = parent[piv];
      cc[par] += cc[piv] - 1;
      vforest[piv] = par;
   }
}

// Relaxed supernodes (nemin rule).  (core_analyse.f90:536-860)
void find_supernodes(int n, int realn, const std::vector<int>& parent, const std::vector<int>& cc,
                     std::vector<int>& sperm, int& nnodes, std::vector<int>& sptr,
                     std::vector<int>& sparent, std::vector<int>& scc, int nemin) {
   const int64_t EZ_HUGE = std::numeric_limits<int64_t>::max();
   std::vector<int> nelim(n + 1, 1), nvert(n + 1, 1), vhead(n + 1, -1), vnext(n + 1, -1);
   std::vector<int> height(n + 1, 1), map(n + 1, 0), npar(n + 1, 0);
   std::vector<char> mark(n + 1, 0);
   std::vector<int64_t> ezero(n + 1, 0);
   ezero[n] = EZ_HUGE;
   nelim[n] = n + 1 + nemin;
   std::vector<int> chead(n + 1, -1), cnext(n + 1, -1), child;
   for (int i = realn - 1; i >= 0; --i) {
      int j = parent[i];
      cnext[i] = chead[j];
      chead[j] = i;
   }
   for (int par = 0; par <= n; ++par) {
      child.clear();
      for (int node = chead[par]; node != -1; node = cnext[node]) child.push_back(node);
      // stable, decreasing column count (sort_by_val / sort_by_val_ms are both stable)
      std::stable_sort(child.begin(), child.end(), [&](int a, int b) { return cc[a] > cc[b]; });
      for (int node : child) {
         bool merge;
         if (ezero[par] == EZ_HUGE) merge = false;
         else
            merge = ((cc[par] == cc[node] - 1) && (nelim[par] == 1)) ||
                    ((nelim[par] < nemin) && (nelim[node] < nemin));
         if (merge) {
            vnext[node] = vhead[par];
            vhead[par] = node;
            ezero[par] += ezero[node] +
                          ((int64_t)cc[par] - 1 + nelim[par] - cc[node] + 1) * nelim[par];
            nelim[par] += nelim[node];
            nvert[par] += nvert[node];
            height[par] = std::max(height[par], height[node]);
            mark[node] = 0;
         } else {
            mark[node] = 1;
         }
      }
   }
   sperm.assign(n, 0);
   sptr.assign(n + 1, 0);
   scc.assign(n, 0);
   nnodes = 0;
   int v = 0;
   std::vector<int> stack;
   for (int node = 0; node < realn; ++node) {
      if (!mark[node]) continue;
      int nn = nnodes++;
      sptr[nn] = v;
      npar[nn] = parent[node];
      scc[nn] = cc[node] + nelim[node] - 1;
      height[parent[node]] = std::max(height[parent[node]], height[node] + 1);
      v += nvert[node];
      int k = v;
      stack.clear();
      stack.push_back(node);
      while (!stack.empty()) {
         int i = stack.back();
         stack.pop_back();
         sperm[i] = --k;
         map[i] = nn;
         if (vnext[i] != -1) stack.push_back(vnext[i]);
         if (vhead[i] != -1) stack.push_back(vhead[i]);
      }
   }
   sptr[nnodes] = v;
   map[n] = nnodes;
   for (int i = realn; i < n; ++i) sperm[i] = i;
   sparent.assign(nnodes, 0);
   for (int node = 0; node < nnodes; ++node) sparent[node] = map[npar[node]];
   sptr.resize(nnodes + 1);
   scc.resize(nnodes);
}

// (core_analyse.f90:1069-1097 apply_perm)
void apply_perm(int n, const std::vector<int>& sperm, std::vector<int>& order,
                std::vector<int>& invp, std::vector<int>& cc) {
   std::vector<int> tmp(n);
   for (int i = 0; i < n; ++i) tmp[i] = cc[i];
   for (int i = 0; i < n; ++i) cc[sperm[i]] = tmp[i];
   for (int i = 0; i < n; ++i) tmp[i] = invp[i];
   for (int i = 0; i < n; ++i) invp[sperm[i]] = tmp[i];
   for (int i = 0; i < n; ++i) order[invp[i]] = i;
}

// Row lists per supernode: own pivots, then children's, then new from A.
// (core_analyse.f90:911-1001)
void find_row_lists(int n, const int64_t* ptr, const int* row, const int* perm, const int* invp,
                    int nnodes, const std::vector<int>& sptr, const std::vector<int>& sparent,
                    const std::vector<int>& scc, std::vector<int64_t>& rptr,
                    std::vector<int>& rlist) {
   std::vector<int> seen(n, -1), chead(nnodes + 1, -1), cnext(nnodes + 1, -1);
   for (int node = nnodes - 1; node >= 0; --node) {
      int p = sparent[node];
      cnext[node] = chead[p];
      chead[p] = node;
   }
   rptr.assign(nnodes + 1, 0);
   int64_t tot = 0;
   for (int node = 0; node < nnodes; ++node) tot += scc[node];
   rlist.assign(tot, 0);
   for (int node = 0; node < nnodes; ++node) {
      rptr[node + 1] = rptr[node] + scc[node];
      int64_t idx = rptr[node];
      for (int piv = sptr[node]; piv < sptr[node + 1]; ++piv) {
         seen[piv] = node;
         rlist[idx++] = piv;
      }
      for (int c = chead[node]; c != -1; c = cnext[c])
         for (int64_t i = rptr[c]; i < rptr[c + 1]; ++i) {
            int j = rlist[i];
            if (j < sptr[node]) continue;
            if (seen[j] == node) continue;
            seen[j] = node;
            rlist[idx++] = j;
         }
      for (int piv = sptr[node]; piv < sptr[node + 1]; ++piv) {
         int col = invp[piv];
         for (int64_t i = ptr[col]; i < ptr[col + 1]; ++i) {
            int j = perm[row[i]];
            if (j < piv) continue;
            if (seen[j] == node) continue;
            seen[j] = node;
            rlist[idx++] = j;
         }
      }
   }
}

// Double-transpose sort of the row lists (ascending within a node).  (core_analyse.f90:1007-1063)
void dbl_tr_sort(int n, int nnodes, const std::vector<int64_t>& rptr, std::vector<int>& rlist) {
   std::vector<int64_t> p(n + 2, 0);
   for (int node = 0; node < nnodes; ++node)
      for (int64_t ii = rptr[node]; ii < rptr[node + 1]; ++ii) p[rlist[ii] + 2]++;
   for (int i = 0; i < n; ++i) p[i + 2] += p[i + 1];
   std::vector<int> col(p[n + 1]);
   for (int node = 0; node < nnodes; ++node)
      for (int64_t ii = rptr[node]; ii < rptr[node + 1]; ++ii) col[p[rlist[ii] + 1]++] = node;
   std::vector<int64_t> np(rptr.begin(), rptr.begin() + nnodes);
   for (int i = 0; i < n; ++i)
      for (int64_t jj = p[i]; jj < p[i + 1]; ++jj) rlist[np[col[jj]]++] = i;
}

// A -> L scatter map.  (anal.f90:1129-1231 build_map)
void build_map(int n, const int64_t* ptr, const int* row, const int* perm, const int* invp,
               int nnodes, const std::vector<int>& sptr, const std::vector<int64_t>& rptr,
               const std::vector<int>& rlist, std::vector<int64_t>& nptr,
               std::vector<int64_t>& nlist) {
   int64_t nz = ptr[n];
   std::vector<int64_t> ptr2(n + 2, 0), origin(nz);
   std::vector<int> row2(nz), map(n, 0);
   for (int i = 0; i < n; ++i)
      for (int64_t jj = ptr[i]; jj < ptr[i + 1]; ++jj) {
         int k = row[jj];
         if (k == i) continue;
         ptr2[k + 2]++;
      }
   for (int i = 0; i < n; ++i) ptr2[i + 2] += ptr2[i + 1];
   for (int i = 0; i < n; ++i)
      for (int64_t jj = ptr[i]; jj < ptr[i + 1]; ++jj) {
         int k = row[jj];
         if (k == i) continue;
         row2[ptr2[k + 1]] = i;
         origin[ptr2[k + 1]] = jj;
         ptr2[k + 1]++;
      }
   nptr.assign(nnodes + 1, 0);
   nlist.assign(2 * nz, 0);
   int64_t pp = 0;
   for (int node = 0; node < nnodes; ++node) {
      int64_t blkm = rptr[node + 1] - rptr[node];
      nptr[node] = pp;
      for (int64_t jj = rptr[node]; jj < rptr[node + 1]; ++jj) map[rlist[jj]] = (int)(jj - rptr[node]);
      for (int j = sptr[node]; j < sptr[node + 1]; ++j) {
         int col = invp[j];
         for (int64_t i = ptr2[col]; i < ptr2[col + 1]; ++i) {
            int k = perm[row2[i]];
            if (k < 0 || k < j) continue;
            nlist[2 * pp] = origin[i];
            nlist[2 * pp + 1] = (int64_t)(j - sptr[node]) * blkm + map[k];
            ++pp;
         }
      }
      for (int j = sptr[node]; j < sptr[node + 1]; ++j) {
         int col = invp[j];
         for (int64_t ii = ptr[col]; ii < ptr[col + 1]; ++ii) {
            int k = perm[row[ii]];
            if (k < 0 || k < j) continue;
            nlist[2 * pp] = ii;
            nlist[2 * pp + 1] = (int64_t)(j - sptr[node]) * blkm + map[k];
            ++pp;
         }
      }
   }
   nptr[nnodes] = pp;
   nlist.resize(2 * pp);
}

}  // namespace

int analyse(int n, const int64_t* ptr, const int* row, const int* order_in, int nemin,
            Symbolic& sym) {
   sym = Symbolic();
   sym.n = n;
   if (n < 0) return sym.flag = SSIDS_ERROR_A_N_OOR;
   sym.nz = n > 0 ? ptr[n] : 0;
   sym.ptr.assign(ptr, ptr + n + 1);
   sym.row.assign(row, row + sym.nz);
   if (nemin < 1) nemin = 32;

   // check_order (anal.f90:147-197): |order| must be a permutation of the pivot positions
   sym.order.assign(n, 0);
   sym.invp.assign(n, -1);
   for (int i = 0; i < n; ++i) {
      int j = order_in[i];
      if (j < 0 || j >= n || sym.invp[j] != -1) return sym.flag = SSIDS_ERROR_ORDER;
      sym.invp[j] = i;
      sym.order[i] = j;
   }
   if (n == 0) {
      sym.sptr.assign(1, 0);
      sym.rptr.assign(1, 0);
      sym.nptr.assign(1, 0);
      sym.part = {0, 0};
      sym.exec_loc = {-1};
      return sym.flag = 0;
   }

   std::vector<int64_t> aptr;
   std::vector<int> arow;
   expand_pattern(n, ptr, row, aptr, arow);

   // basic_analyse (core_analyse.f90:38-151)
   std::vector<int> parent, cc, sperm, scc;
   find_etree(n, aptr.data(), arow.data(), sym.order.data(), sym.invp.data(), parent);
   find_postorder(n, sym.realn, aptr.data(), sym.order, sym.invp, parent);
   if (sym.realn != n) sym.flag = SSIDS_WARNING_ANAL_SINGULAR;
   find_col_counts(n, aptr.data(), arow.data(), sym.order.data(), sym.invp.data(), parent, cc);
   find_supernodes(n, sym.realn, parent, cc, sperm, sym.nnodes, sym.sptr, sym.sparent, scc, nemin);
   apply_perm(n, sperm, sym.order, sym.invp, cc);
   find_row_lists(n, aptr.data(), arow.data(), sym.order.data(), sym.invp.data(), sym.nnodes,
                  sym.sptr, sym.sparent, scc, sym.rptr, sym.rlist);
   // calc_stats (core_analyse.f90:862-905)
   sym.num_factor = 0;
   sym.num_flops = 0;
   for (int node = 0; node < sym.nnodes; ++node) {
      int64_t ne = sym.sptr[node + 1] - sym.sptr[node];
      int64_t m = scc[node] - ne;
      sym.num_factor += (ne * (ne + 1)) / 2 + ne * m;
      for (int64_t j = 1; j <= ne; ++j) sym.num_flops += (m + j) * (m + j);
   }
   dbl_tr_sort(n, sym.nnodes, sym.rptr, sym.rlist);

   // analyse_phase tail (anal.f90:995-1108): variables with no entries get no pivot position
   int rank = sym.sptr[sym.nnodes];
   std::vector<int> perm_for_map(sym.order);
   for (int j = rank; j < n; ++j) perm_for_map[sym.invp[j]] = -1;
   build_map(n, ptr, row, perm_for_map.data(), sym.invp.data(), sym.nnodes, sym.sptr, sym.rptr,
             sym.rlist, sym.nptr, sym.nlist);
   sym.order = perm_for_map;

   std::vector<int> level(sym.nnodes + 1, 0);
   sym.maxfront = 0;
   sym.maxdepth = 0;
   for (int i = sym.nnodes - 1; i >= 0; --i) {
      int blkn = sym.sptr[i + 1] - sym.sptr[i];
      level[i] = level[sym.sparent[i]] + 1;
      sym.maxfront = std::max(sym.maxfront, blkn);
      sym.maxdepth = std::max(sym.maxdepth, level[i]);
   }
   sym.matrix_rank = rank;
   sym.nparts = 1;
   sym.part = {0, sym.nnodes};
   sym.exec_loc = {-1};
   sym.contrib_ptr = {0, 0};
   return sym.flag;
}

}  // namespace b200

namespace b200 {

// Subtree partition for `ndev` equal devices.  Follows the reference's scheme
// (anal.f90:284-459: start from the root subtrees, keep splitting the largest leaf subtree,
// allocate largest-first to the least loaded device as calc_exec_alloc :503-590 does), with one
// deliberate change of objective: the reference runs the top of the tree on all CPUs and so only
// balances the leaves (max_load_inbalance); here the root part runs on the owner of its first
// contribution after the leaves, so we stop at the split count that minimises
// max_i(load_i) + flops(root part), trying at most 2*(ndev) further splits past the balance test
// like the reference's 2*(nregion+ngpu) bound.
void find_subtree_partition(Symbolic& sym, int ndev, double max_load_inbalance, int64_t min_gpu_work) {
   const int nn = sym.nnodes;
   if (ndev < 1) ndev = 1;
   std::vector<int64_t> flops(nn + 1, 0);
   std::vector<int> size(nn + 1, 0);
   for (int node = 0; node < nn; ++node) {
      int64_t m = sym.rptr[node + 1] - sym.rptr[node], nc = sym.sptr[node + 1] - sym.sptr[node];
      for (int64_t j = m - nc + 1; j <= m; ++j) flops[node] += j * j;   // compute_flops, anal.f90:205-224
      size[node] += 1;
      flops[sym.sparent[node]] += flops[node];
      size[sym.sparent[node]] += size[node];
   }
   std::vector<std::vector<int>> children(nn + 1);
   for (int node = 0; node < nn; ++node) children[sym.sparent[node]].push_back(node);
   std::vector<int> roots = children[nn];   // current leaf-subtree roots
   std::vector<char> is_top(nn + 1, 0);
   int64_t top_flops = 0;
   auto evaluate = [&](std::vector<int>& owner, double& balance) -> int64_t {
      std::vector<int> ord(roots.size());
      for (size_t i = 0; i < ord.size(); ++i) ord[i] = (int)i;
      std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return flops[roots[a]] > flops[roots[b]]; });
      std::vector<int64_t> load(ndev, 0);
      owner.assign(roots.size(), 0);
      for (int i : ord) {
         int best = 0;
         for (int d = 1; d < ndev; ++d) if (load[d] < load[best]) best = d;
         owner[i] = best;
         load[best] += flops[roots[i]];
      }
      int64_t mx = 0, tot = 0;
      for (int d = 0; d < ndev; ++d) { mx = std::max(mx, load[d]); tot += load[d]; }
      balance = tot > 0 ? (double)ndev * (double)mx / (double)tot : 1.0;
      return mx + top_flops;
   };
   std::vector<int> owner, best_owner, best_roots;
   std::vector<char> best_top;
   int64_t best_span = -1;
   const int max_tries = (ndev == 1) ? 0 : 16 * ndev;
   for (int it = 0; it <= max_tries; ++it) {
      double bal;
      int64_t span = evaluate(owner, bal);
      if (best_span < 0 || span < best_span) {
         best_span = span; best_owner = owner; best_roots = roots; best_top = is_top;
      }
      (void)max_load_inbalance;
      if (it == max_tries) break;
      // split the largest subtree that still has children and enough work
      int pick = -1;
      for (size_t i = 0; i < roots.size(); ++i) {
         if (children[roots[i]].empty()) continue;
         if (flops[roots[i]] < min_gpu_work) continue;
         if (pick < 0 || flops[roots[i]] > flops[roots[pick]]) pick = (int)i;
      }
      if (pick < 0) break;
      int r = roots[pick];
      roots.erase(roots.begin() + pick);
      is_top[r] = 1;
      top_flops += flops[r];
      for (int c : children[r]) { roots.push_back(c); top_flops -= flops[c]; }
   }
   // emit contiguous parts in node order: leaf subtrees [r-size+1, r], runs of top nodes between
   std::vector<int> leaf_owner(nn, -2);   // owner per node; -1 = root part
   for (size_t i = 0; i < best_roots.size(); ++i) {
      int r = best_roots[i];
      for (int node = r - size[r] + 1; node <= r; ++node) leaf_owner[node] = best_owner[i];
   }
   for (int node = 0; node < nn; ++node) if (leaf_owner[node] == -2) leaf_owner[node] = -1;
   std::vector<char> is_leaf_root(nn, 0);
   for (int r : best_roots) is_leaf_root[r] = 1;
   sym.part.clear();
   sym.exec_loc.clear();
   sym.part.push_back(0);
   for (int node = 0; node < nn; ++node) {
      bool end_here = (node == nn - 1) || is_leaf_root[node] ||
                      (leaf_owner[node] == -1 && leaf_owner[node + 1] != -1);
      if (end_here) {
         sym.exec_loc.push_back(leaf_owner[node]);
         sym.part.push_back(node + 1);
      }
   }
   sym.nparts = (int)sym.exec_loc.size();
   if (nn == 0) { sym.nparts = 1; sym.part = {0, 0}; sym.exec_loc = {-1}; }
}


int equilib_scale_sym(int n, const int64_t* ptr, const int* row, const double* val, double* scaling,
                      int max_iterations, float tol) {
   std::vector<double> maxentry(std::max(n, 1));
   for (int i = 0; i < n; ++i) scaling[i] = 1.0;
   int itr = 1;
   for (; itr <= max_iterations; ++itr) {
      // maximum entry in each row and column; the matrix is symmetric but only half of it is held
      std::fill(maxentry.begin(), maxentry.end(), 0.0);
      for (int c = 0; c < n; ++c)
         for (int64_t j = ptr[c]; j < ptr[c + 1]; ++j) {
            const int r = row[j];
            const double v = std::fabs(scaling[r] * val[j] * scaling[c]);
            maxentry[r] = std::max(maxentry[r], v);
            maxentry[c] = std::max(maxentry[c], v);
         }
      // update the scaling (empty columns keep theirs)
      for (int i = 0; i < n; ++i) if (maxentry[i] > 0) scaling[i] = scaling[i] / std::sqrt(maxentry[i]);
      double worst = 0.0;
      for (int i = 0; i < n; ++i) worst = std::max(worst, std::fabs(1 - maxentry[i]));
      if (n > 0 && worst < (double)tol) break;
      if (n == 0) break;
   }
   return itr - 1;      // inform%iterations
}

void proportional_mapping(Symbolic& sym, int ndev) {
   const int nn = sym.nnodes;
   if (ndev < 1) ndev = 1;
   sym.map_owner.assign(std::max(nn, 1), 0);
   sym.map_stage.assign(std::max(nn, 1), 0);
   sym.nstages = 1;
   if (nn == 0) return;
   std::vector<int64_t> flops(nn + 1, 0);
   std::vector<int> size(nn + 1, 0);
   for (int node = 0; node < nn; ++node) {
      int64_t m = sym.rptr[node + 1] - sym.rptr[node], nc = sym.sptr[node + 1] - sym.sptr[node];
      for (int64_t j = m - nc + 1; j <= m; ++j) flops[node] += j * j;   // compute_flops, anal.f90:205-224
      size[node] += 1;
      flops[sym.sparent[node]] += flops[node];
      size[sym.sparent[node]] += size[node];
   }
   std::vector<std::vector<int>> children(nn + 1);
   for (int node = 0; node < nn; ++node) children[sym.sparent[node]].push_back(node);
   std::vector<char> upper(nn + 1, 0);
   struct Item { int node, r0, g; };
   std::vector<Item> stack;
   stack.push_back({nn, 0, ndev});           // the virtual root holds the forest
   while (!stack.empty()) {
      const Item it = stack.back();
      stack.pop_back();
      if (it.g <= 1 && it.node < nn) {         // one device left: the whole subtree is a leaf part of it
         for (int q = it.node - size[it.node] + 1; q <= it.node; ++q) sym.map_owner[q] = it.r0;
         continue;
      }
      if (it.node < nn) { sym.map_owner[it.node] = it.r0; upper[it.node] = 1; }
      std::vector<int> kids = children[it.node];
      std::stable_sort(kids.begin(), kids.end(), [&](int a, int b) { return flops[a] > flops[b]; });
      int64_t W = 0;
      for (int c : kids) W += flops[c];
      // devices per child: the children whose share is at least half a device are candidates (largest first, at most
      // g of them, the largest always); every candidate gets one device, the others go one at a time to the candidate
      // with the largest flops per device -- the allocation that minimises the largest load per device.  (Rounded
      // shares cascade badly down an unbalanced tree: the 27-point 160^3 grid on 8 devices left one device with 18 %
      // of the flops and two with 2 % each.)
      std::vector<int> alloc(kids.size(), 0);
      {
         size_t ncand = 0;
         for (size_t i = 0; i < kids.size() && (int)ncand < it.g; ++i) {
            const double share = W > 0 ? (double)it.g * (double)flops[kids[i]] / (double)W : 0.0;
            if (i == 0 || share >= 0.5) { alloc[i] = 1; ++ncand; } else break;      // kids are sorted: the rest is smaller
         }
         int left = ncand > 0 ? it.g - (int)ncand : 0;      // a leaf of the tree with several devices keeps them idle
         while (left > 0) {
            size_t best = 0;
            double worst = -1.0;
            for (size_t i = 0; i < ncand; ++i) {
               const double load = (double)flops[kids[i]] / (double)alloc[i];
               if (load > worst) { worst = load; best = i; }
            }
            ++alloc[best];
            --left;
         }
      }
      int r = it.r0;
      for (size_t i = 0; i < kids.size(); ++i) {
         if (alloc[i] > 0) { stack.push_back({kids[i], r, alloc[i]}); r += alloc[i]; }
         else stack.push_back({kids[i], it.r0, 1});      // too small for a device of its own: stays with the owner
      }
   }
   // stages, children before parents (postorder numbering)
   int mx = 0;
   for (int node = 0; node < nn; ++node) {
      int st = 0;
      for (int c : children[node]) st = std::max(st, sym.map_stage[c] + (sym.map_owner[c] != sym.map_owner[node] ? 1 : 0));
      sym.map_stage[node] = st;
      mx = std::max(mx, st);
   }
   (void)upper;
   sym.nstages = mx + 1;
}

}  // namespace b200

namespace b200 {

// Engine-internal relaxation of the assembly tree.  The reference's amalgamation (do_merge,
// core_analyse.f90:806-822) stops at nemin columns, which on nested-dissection orderings leaves
// long chains parent(i) = i+1 of 32..62-column supernodes whose fronts differ by a handful of
// rows (each link of a separator plane brings in 2-3 new boundary rows).  Every link would cost a
// full extend-add of an (m-62)^2 contribution block.  The numeric engine therefore factorises a
// chain as ONE front: node p absorbs its last child c = p-1 when the explicit zeros this adds to L
// stay below `max_zero_frac` of the fused L.  The fused front's rows are cols(chain) ++ rlist(p)
// (the child's contribution rows are a subset of the parent's rows in any assembly tree), so only
// structural zeros are introduced and the factorisation is the same P A P^T = L D L^T.  What SLS
// reports (sptr, sparent, rptr, rlist, num_factor, num_flops, ...) stays the reference's tree.
void fuse_chains(const Symbolic& in, Symbolic& out, double max_zero_frac) {
   const int nn = in.nnodes;
   out = Symbolic();
   out.n = in.n; out.nz = in.nz; out.ptr = in.ptr; out.row = in.row; out.order = in.order; out.invp = in.invp;
   out.num_factor = in.num_factor; out.num_flops = in.num_flops; out.maxfront = in.maxfront;
   out.maxdepth = in.maxdepth; out.matrix_rank = in.matrix_rank; out.realn = in.realn; out.flag = in.flag;
   // group[i] = index of the fused node that original node i belongs to; groups are runs i..top
   std::vector<int> first(nn, 0);          // first original node of the run ending at i
   std::vector<int64_t> zeros(nn, 0);      // explicit zeros accumulated in the run ending at i
   std::vector<char> absorbed(nn, 0);      // node i was absorbed by i+1
   std::vector<int> nchild(nn + 1, 0);
   for (int i = 0; i < nn; ++i) nchild[in.sparent[i]]++;
   for (int p = 0; p < nn; ++p) {
      first[p] = p;
      zeros[p] = 0;
      const int c = p - 1;
      if (c < 0 || in.sparent[c] != p) continue;
      const int64_t gcols = in.sptr[c + 1] - in.sptr[first[c]];
      // a node with siblings is only absorbed while it is small: a large child fused into its parent is
      // eliminated after its siblings instead of beside them (the fronts of a level run concurrently)
      if (nchild[p] > 1 && gcols > 128) continue;
      const int64_t ncol_p = in.sptr[p + 1] - in.sptr[p];
      const int64_t nrow_p = in.rptr[p + 1] - in.rptr[p];
      const int64_t contrib_c = (in.rptr[c + 1] - in.rptr[c]) - (in.sptr[c + 1] - in.sptr[c]);
      const int64_t newz = gcols * (nrow_p - contrib_c);
      const int64_t fused_m = gcols + nrow_p, fused_n = gcols + ncol_p;
      const int64_t fused_entries = fused_m * fused_n - fused_n * (fused_n - 1) / 2;
      if ((double)(zeros[c] + newz) <= max_zero_frac * (double)fused_entries) {
         absorbed[c] = 1;
         first[p] = first[c];
         zeros[p] = zeros[c] + newz;
      }
   }
   std::vector<int> gid(nn, -1);
   int ng = 0;
   for (int i = 0; i < nn; ++i) {
      if (absorbed[i]) continue;            // i is not a top
      for (int q = first[i]; q <= i; ++q) gid[q] = ng;
      ++ng;
   }
   out.nnodes = ng;
   out.sptr.assign(ng + 1, 0);
   out.sparent.assign(ng, ng);
   out.rptr.assign(ng + 1, 0);
   out.nptr.assign(ng + 1, 0);
   std::vector<int> top(ng, 0), bot(ng, 0);
   for (int i = 0; i < nn; ++i) {
      if (absorbed[i]) continue;
      int g = gid[i];
      top[g] = i;
      bot[g] = first[i];
   }
   int64_t rtot = 0;
   for (int g = 0; g < ng; ++g) {
      const int t = top[g], b = bot[g];
      out.sptr[g] = in.sptr[b];
      out.sparent[g] = in.sparent[t] < nn ? gid[in.sparent[t]] : ng;
      const int64_t gcols_below = in.sptr[t] - in.sptr[b];
      out.rptr[g] = rtot;
      rtot += gcols_below + (in.rptr[t + 1] - in.rptr[t]);
   }
   out.sptr[ng] = in.sptr[nn];
   out.rptr[ng] = rtot;
   out.rlist.assign(rtot, 0);
   out.nlist.assign(in.nlist.size(), 0);
   std::vector<int> map(in.n, 0);
   int64_t pp = 0;
   for (int g = 0; g < ng; ++g) {
      const int t = top[g], b = bot[g];
      int64_t o = out.rptr[g];
      for (int piv = in.sptr[b]; piv < in.sptr[t]; ++piv) out.rlist[o++] = piv;
      for (int64_t j = in.rptr[t]; j < in.rptr[t + 1]; ++j) out.rlist[o++] = in.rlist[j];
      const int64_t blkm = out.rptr[g + 1] - out.rptr[g];
      for (int64_t j = out.rptr[g]; j < out.rptr[g + 1]; ++j) map[out.rlist[j]] = (int)(j - out.rptr[g]);
      out.nptr[g] = pp;
      for (int q = b; q <= t; ++q) {
         const int64_t qm = in.rptr[q + 1] - in.rptr[q];
         const int coff = in.sptr[q] - in.sptr[b];
         const int* qrl = in.rlist.data() + in.rptr[q];
         for (int64_t e = in.nptr[q]; e < in.nptr[q + 1]; ++e) {
            const int64_t dest = in.nlist[2 * e + 1];
            const int col = (int)(dest / qm), row = (int)(dest % qm);
            out.nlist[2 * pp] = in.nlist[2 * e];
            out.nlist[2 * pp + 1] = (int64_t)(coff + col) * blkm + map[qrl[row]];
            ++pp;
         }
      }
   }
   out.nptr[ng] = pp;
   out.nparts = 1;
   out.part = {0, ng};
   out.exec_loc = {-1};
   out.contrib_ptr = {0, 0};
}

}  // namespace b200
