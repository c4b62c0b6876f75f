// Fill-reducing orderings for the b200ldlt SLS layer (CPU side: the north star keeps ordering on the host).
//
// In the reference SLS_analyse takes its orderings from packages that a stock build only has as stubs:
// control%ordering >= 0 goes through HSL MC68 (AMD, minimum degree, METIS nested dissection, ...; sls.f90:2395-2470)
// and control%ordering < 0 asks SSIDS for its own (METIS, src/spral/metis4_wrapper.f90).  So no file of the reference
// pins a permutation: what is restated here is the published algorithm of each class, and the parity tests check what
// an ordering must satisfy (a permutation; fill and flops against the natural order and against a geometric nested
// dissection on grids) rather than bits -- DESIGN.md section 2, "parity unpinned".
//
//   amd_order : approximate minimum degree on the quotient graph (Amestoy, Davis and Duff, SIAM J. Matrix Anal. Appl.
//               17 (1996)): element absorption, approximate external degrees through the |Le \ Lp| pass, mass
//               elimination, supervariable detection by hashing, dense rows ordered last.
//   nd_order  : nested dissection by recursive bisection with vertex separators taken from the level structure of a
//               pseudo-peripheral vertex (George and Liu's automatic nested dissection), minimum degree inside the
//               leaves, separators last.
//
// Input everywhere: lower triangle by columns, 0-based, diagonal entries allowed (ignored).  Output: order[i] = position of
// variable i in the pivot sequence, 0-based.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

#include "analyse.hpp"

namespace b200 {

namespace {

// full symmetric adjacency (no diagonal, no duplicates) of the vertices listed in `verts` (local numbering in the
// order of `verts`); glob2loc[v] = local index or -1
void build_adjacency(int n, const int64_t* ptr, const int* row, std::vector<int64_t>& ap, std::vector<int>& ai) {
   std::vector<int64_t> cnt(n + 1, 0);
   for (int c = 0; c < n; ++c)
      for (int64_t k = ptr[c]; k < ptr[c + 1]; ++k) {
         const int r = row[k];
         if (r == c || r < 0 || r >= n) continue;
         ++cnt[r + 1];
         ++cnt[c + 1];
      }
   ap.assign(n + 1, 0);
   for (int i = 0; i < n; ++i) ap[i + 1] = ap[i] + cnt[i + 1];
   ai.assign((size_t)ap[n], 0);
   std::vector<int64_t> pos(ap.begin(), ap.end() - 1);
   for (int c = 0; c < n; ++c)
      for (int64_t k = ptr[c]; k < ptr[c + 1]; ++k) {
         const int r = row[k];
         if (r == c || r < 0 || r >= n) continue;
         ai[pos[r]++] = c;
         ai[pos[c]++] = r;
      }
   // sort and remove duplicates
   std::vector<int64_t> np(n + 1, 0);
   int64_t w = 0;
   for (int i = 0; i < n; ++i) {
      std::sort(ai.begin() + ap[i], ai.begin() + ap[i + 1]);
      const int64_t b = w;
      for (int64_t k = ap[i]; k < ap[i + 1]; ++k)
         if (w == b || ai[w - 1] != ai[k]) ai[w++] = ai[k];
      np[i + 1] = w;
   }
   ap.swap(np);
   ai.resize((size_t)w);
}

// AMD on a graph given by adjacency lists (ap / ai, symmetric, no diagonal).  seq receives the vertices in elimination
// order.
void amd_graph(int n, const std::vector<int64_t>& ap, const std::vector<int>& ai, std::vector<int>& seq) {
   seq.clear();
   seq.reserve(n);
   if (n == 0) return;
   std::vector<std::vector<int>> A(n), E(n), Le(n);     // variable lists, element lists of a variable; variables of an element
   std::vector<int> nv(n, 1), deg(n, 0), state(n, 0);   // state 0 variable, 1 element, 2 absorbed / dead element, 3 merged variable
   std::vector<int> head, nxt(n, -1), prv(n, -1);       // degree buckets
   std::vector<int> mark(n, 0), w(n, 0), sup_next(n, -1), sup_last(n);
   std::vector<int64_t> le_w(n, 0);                      // weighted size of an element
   std::iota(sup_last.begin(), sup_last.end(), 0);
   head.assign(n + 1, -1);
   int mindeg = n;
   auto bucket_add = [&](int i) {
      const int d = std::min(std::max(deg[i], 0), n);
      nxt[i] = head[d]; prv[i] = -1;
      if (head[d] >= 0) prv[head[d]] = i;
      head[d] = i;
      if (d < mindeg) mindeg = d;
   };
   auto bucket_del = [&](int i) {
      const int d = std::min(std::max(deg[i], 0), n);
      if (prv[i] >= 0) nxt[prv[i]] = nxt[i]; else head[d] = nxt[i];
      if (nxt[i] >= 0) prv[nxt[i]] = prv[i];
      nxt[i] = prv[i] = -1;
   };
   // dense rows last (AMD's default: more than max(16, 10 sqrt(n)) neighbours)
   const int dense = std::max(16, (int)(10.0 * std::sqrt((double)n)));
   std::vector<int> dense_list;
   std::vector<char> is_dense(n, 0);
   for (int i = 0; i < n; ++i)
      if (ap[i + 1] - ap[i] > dense) { is_dense[i] = 1; dense_list.push_back(i); state[i] = 2; nv[i] = 0; }
   int nleft = n - (int)dense_list.size();
   for (int i = 0; i < n; ++i) {
      if (is_dense[i]) continue;
      for (int64_t k = ap[i]; k < ap[i + 1]; ++k) if (!is_dense[ai[k]]) A[i].push_back(ai[k]);
      deg[i] = (int)A[i].size();
      bucket_add(i);
   }
   int tag = 0;                                          // mark stamps
   int64_t wflg = 1;
   std::vector<int64_t> wv(n, 0);
   std::vector<int> Lp, hash_bucket_keys;
   std::vector<int> hhead;                               // hash buckets (size n), values = variable index chain through hnext
   std::vector<int> hnext(n, -1);
   hhead.assign(n, -1);
   std::vector<unsigned> hval(n, 0);
   int eliminated = 0;
   while (eliminated < nleft) {
      while (mindeg <= n && head[mindeg] < 0) ++mindeg;
      const int p = head[mindeg];
      bucket_del(p);
      // ---- the new element p: Lp = A_p  U  (U_{e in E_p} L_e)  \ {p}
      ++tag;
      mark[p] = tag;
      Lp.clear();
      int64_t lp_w = 0;
      for (int j : A[p])
         if (state[j] == 0 && nv[j] > 0 && mark[j] != tag) { mark[j] = tag; Lp.push_back(j); lp_w += nv[j]; }
      for (int e : E[p]) {
         if (state[e] != 1) continue;
         for (int j : Le[e])
            if (state[j] == 0 && nv[j] > 0 && mark[j] != tag) { mark[j] = tag; Lp.push_back(j); lp_w += nv[j]; }
         state[e] = 2;                                   // absorbed into p
         std::vector<int>().swap(Le[e]);
      }
      std::vector<int>().swap(A[p]);
      std::vector<int>().swap(E[p]);
      state[p] = 1;
      const int nvp = nv[p];
      eliminated += nvp;
      seq.push_back(p);
      // ---- first pass: w[e] = |L_e \ Lp| (weighted) for the elements touching Lp
      if (wflg > (int64_t)4e18) { std::fill(wv.begin(), wv.end(), 0); wflg = 1; }
      for (int i : Lp) {
         for (int e : E[i]) {
            if (state[e] != 1) continue;
            if (wv[e] < wflg) wv[e] = le_w[e] + wflg;
            wv[e] -= nv[i];
         }
      }
      // ---- second pass: clean the lists of every i in Lp, approximate degree, hash
      hash_bucket_keys.clear();
      for (int i : Lp) {
         bucket_del(i);
         // elements: drop the absorbed ones, add p
         unsigned h = 0;
         int64_t d = 0;
         {
            std::vector<int>& Ei = E[i];
            size_t q = 0;
            for (int e : Ei) {
               if (state[e] != 1) continue;
               const int64_t dext = wv[e] - wflg;
               if (dext > 0) { d += dext; Ei[q++] = e; h += (unsigned)e; }
               // dext == 0: L_e is a subset of Lp, the element is absorbed by p (aggressive absorption)
               else { state[e] = 2; std::vector<int>().swap(Le[e]); }
            }
            Ei.resize(q);
            Ei.push_back(p);
            h += (unsigned)p;
         }
         // variables: drop dead ones and those now reachable through p
         {
            std::vector<int>& Ai = A[i];
            size_t q = 0;
            for (int j : Ai) {
               if (state[j] != 0 || nv[j] <= 0 || mark[j] == tag) continue;
               d += nv[j];
               Ai[q++] = j;
               h += (unsigned)j;
            }
            Ai.resize(q);
         }
         const int64_t dlp = lp_w - nv[i];
         int64_t dnew = std::min<int64_t>((int64_t)deg[i] + dlp, d + dlp);
         dnew = std::min<int64_t>(dnew, (int64_t)nleft - eliminated - nv[i]);
         deg[i] = (int)std::max<int64_t>(dnew, 0);
         hval[i] = h;
      }
      wflg += (int64_t)n + 1;                            // stamps of this round expire
      // ---- mass elimination and supervariable detection
      for (int i : Lp) {
         if (nv[i] <= 0) continue;
         if (A[i].empty() && E[i].size() == 1) {
            // only connected through p: eliminated with it at no extra fill
            eliminated += nv[i];
            lp_w -= nv[i];
            sup_next[sup_last[p]] = i;
            sup_last[p] = sup_last[i];
            nv[i] = 0;
            state[i] = 3;
            std::vector<int>().swap(E[i]);
            continue;
         }
         const int b = (int)(hval[i] % (unsigned)n);
         if (hhead[b] < 0) hash_bucket_keys.push_back(b);
         hnext[i] = hhead[b];
         hhead[b] = i;
      }
      for (int b : hash_bucket_keys) {
         // compare the members of a bucket pairwise (lists are compared as sets through the mark array)
         for (int i = hhead[b]; i >= 0; i = hnext[i]) {
            if (nv[i] <= 0) continue;
            ++tag;
            for (int j : A[i]) mark[j] = tag;
            for (int e : E[i]) mark[e] = tag;
            int prevj = i;
            for (int j = hnext[i]; j >= 0; j = hnext[j]) {
               bool same = nv[j] > 0 && hval[j] == hval[i] && A[j].size() == A[i].size() && E[j].size() == E[i].size();
               if (same) for (int v : A[j]) if (mark[v] != tag) { same = false; break; }
               if (same) for (int e : E[j]) if (mark[e] != tag) { same = false; break; }
               if (same) {
                  // j is indistinguishable from i: merge
                  nv[i] += nv[j];
                  deg[i] = std::max(0, deg[i] - nv[j]);
                  nv[j] = 0;
                  state[j] = 3;
                  sup_next[sup_last[i]] = j;
                  sup_last[i] = sup_last[j];
                  std::vector<int>().swap(A[j]);
                  std::vector<int>().swap(E[j]);
                  hnext[prevj] = hnext[j];               // unlink
               } else {
                  prevj = j;
               }
            }
         }
         hhead[b] = -1;
      }
      // ---- the element's list: the surviving principal variables, and their place in the degree lists
      // (mark stamps were reused by the hash pass: rebuild Lp's membership from nv)
      std::vector<int>& L = Le[p];
      L.clear();
      int64_t lw = 0;
      for (int i : Lp)
         if (nv[i] > 0 && state[i] == 0) { L.push_back(i); lw += nv[i]; bucket_add(i); }
      le_w[p] = lw;
      if (L.empty()) state[p] = 2;
      for (int i : Lp) hnext[i] = -1;
   }
   // expand supervariables: a principal variable is followed by the variables merged into it
   std::vector<int> out;
   out.reserve(n);
   for (int p : seq)
      for (int v = p; v >= 0; v = sup_next[v]) out.push_back(v);
   for (int v : dense_list) out.push_back(v);
   seq.swap(out);
}

}  // namespace

int amd_order(int n, const int64_t* ptr, const int* row, int* order) {
   if (n < 0) return SSIDS_ERROR_A_N_OOR;
   std::vector<int64_t> ap;
   std::vector<int> ai, seq;
   build_adjacency(n, ptr, row, ap, ai);
   amd_graph(n, ap, ai, seq);
   if ((int)seq.size() != n) return SSIDS_ERROR_UNKNOWN;
   std::vector<char> seen(n, 0);
   for (int k = 0; k < n; ++k) {
      const int v = seq[k];
      if (v < 0 || v >= n || seen[v]) return SSIDS_ERROR_UNKNOWN;
      seen[v] = 1;
      order[v] = k;
   }
   return 0;
}

namespace {

// Nested dissection of the subgraph induced by verts (all with part[v] == id), appended to seq.
struct NdWork {
   int n;
   const std::vector<int64_t>* ap;
   const std::vector<int>* ai;
   std::vector<int> part;        // current part id of every vertex (-1: already ordered / separator)
   std::vector<int> level, queue, loc;
   int leaf_size;
};

// BFS over vertices of part `id` from `root`; fills W.level (>= 0 for reached vertices of this BFS, stamped by `stamp`
// in W.loc) and returns the vertices in BFS order
void bfs(NdWork& W, int root, int id, std::vector<int>& ord, std::vector<int>& lvl_of) {
   ord.clear();
   ord.push_back(root);
   lvl_of[root] = 0;
   for (size_t q = 0; q < ord.size(); ++q) {
      const int v = ord[q];
      for (int64_t k = (*W.ap)[v]; k < (*W.ap)[v + 1]; ++k) {
         const int u = (*W.ai)[k];
         if (W.part[u] == id && lvl_of[u] < 0) { lvl_of[u] = lvl_of[v] + 1; ord.push_back(u); }
      }
   }
}

void order_leaf(NdWork& W, const std::vector<int>& verts, std::vector<int>& seq) {
   // minimum degree on the induced subgraph
   const int m = (int)verts.size();
   for (int q = 0; q < m; ++q) W.loc[verts[q]] = q;
   std::vector<int64_t> lp(m + 1, 0);
   std::vector<int> li;
   for (int q = 0; q < m; ++q) {
      const int v = verts[q];
      for (int64_t k = (*W.ap)[v]; k < (*W.ap)[v + 1]; ++k) {
         const int u = (*W.ai)[k];
         if (W.loc[u] >= 0) li.push_back(W.loc[u]);
      }
      lp[q + 1] = (int64_t)li.size();
   }
   std::vector<int> s;
   amd_graph(m, lp, li, s);
   for (int q : s) seq.push_back(verts[q]);
   for (int q = 0; q < m; ++q) W.loc[verts[q]] = -1;
}

void nd_recurse(NdWork& W, std::vector<int> verts, int& next_id, std::vector<int>& seq_front, std::vector<std::vector<int>>& seps) {
   // iterative over a work stack of parts; separators are emitted after both halves, so the result is assembled from
   // a stack of (part, marker) records
   struct Item { std::vector<int> verts; };
   std::vector<Item> stack;
   stack.push_back({std::move(verts)});
   std::vector<int> lvl(W.n, -1), ord, ord2;
   // post-order assembly: pieces[k] = vertices to append; a part's separator goes after its two halves, i.e. the final
   // sequence is built by a recursive structure.  To stay iterative the order is produced in REVERSE: separator first,
   // then the halves (each recursively separator-first); reversing the whole sequence at the end puts separators last.
   std::vector<int>& rev = seq_front;
   (void)seps;
   while (!stack.empty()) {
      std::vector<int> vs = std::move(stack.back().verts);
      stack.pop_back();
      if (vs.empty()) continue;
      const int id = next_id++;
      for (int v : vs) W.part[v] = id;
      if ((int)vs.size() <= W.leaf_size) {
         std::vector<int> s;
         order_leaf(W, vs, s);
         for (auto it = s.rbegin(); it != s.rend(); ++it) rev.push_back(*it);
         for (int v : vs) W.part[v] = -1;
         continue;
      }
      // connected component of the first vertex; the rest of the part goes back on the stack
      for (int v : vs) lvl[v] = -1;
      bfs(W, vs[0], id, ord, lvl);
      if (ord.size() < vs.size()) {
         std::vector<int> rest;
         for (int v : vs) if (lvl[v] < 0) rest.push_back(v);
         for (int v : ord) lvl[v] = -1;
         std::vector<int> comp(ord);
         stack.push_back({std::move(rest)});
         stack.push_back({std::move(comp)});
         for (int v : vs) W.part[v] = -1;
         continue;
      }
      // pseudo-peripheral root: restart from a vertex of the last level while the depth grows
      int root = ord.back(), depth = lvl[ord.back()];
      for (int it = 0; it < 4; ++it) {
         for (int v : ord) lvl[v] = -1;
         bfs(W, root, id, ord2, lvl);
         const int d2 = lvl[ord2.back()];
         ord.swap(ord2);
         if (d2 <= depth) { depth = d2; break; }
         depth = d2;
         root = ord.back();
      }
      if (depth < 2) {
         // (nearly) complete graph: no useful separator
         std::vector<int> s;
         for (int v : ord) lvl[v] = -1;
         order_leaf(W, vs, s);
         for (auto it = s.rbegin(); it != s.rend(); ++it) rev.push_back(*it);
         for (int v : vs) W.part[v] = -1;
         continue;
      }
      // the level that splits the vertices most evenly
      std::vector<int> cnt(depth + 1, 0);
      for (int v : ord) ++cnt[lvl[v]];
      int best = 1;
      int64_t below = cnt[0], bestcost = -1;
      for (int l = 1; l < depth; ++l) {
         const int64_t above = (int64_t)ord.size() - below - cnt[l];
         const int64_t cost = std::llabs(below - above) + 2 * (int64_t)cnt[l];   // balance, with a premium on small separators
         if (bestcost < 0 || cost < bestcost) { bestcost = cost; best = l; }
         below += cnt[l];
      }
      std::vector<int> a, b, sep;
      for (int v : ord) {
         if (lvl[v] < best) a.push_back(v);
         else if (lvl[v] > best) b.push_back(v);
         else {
            // a vertex of the separating level without a neighbour in the next level can join the lower half
            bool touches_up = false;
            for (int64_t k = (*W.ap)[v]; k < (*W.ap)[v + 1] && !touches_up; ++k) {
               const int u = (*W.ai)[k];
               if (W.part[u] == id && lvl[u] == best + 1) touches_up = true;
            }
            (touches_up ? sep : a).push_back(v);
         }
      }
      for (int v : ord) lvl[v] = -1;
      for (int v : sep) { rev.push_back(v); W.part[v] = -1; }
      for (int v : a) W.part[v] = -1;
      for (int v : b) W.part[v] = -1;
      stack.push_back({std::move(a)});
      stack.push_back({std::move(b)});
   }
}

}  // namespace

int nd_order(int n, const int64_t* ptr, const int* row, int* order, int leaf_size) {
   if (n < 0) return SSIDS_ERROR_A_N_OOR;
   std::vector<int64_t> ap;
   std::vector<int> ai;
   build_adjacency(n, ptr, row, ap, ai);
   NdWork W;
   W.n = n;
   W.ap = &ap;
   W.ai = &ai;
   W.part.assign(n, -1);
   W.loc.assign(n, -1);
   W.leaf_size = leaf_size > 0 ? leaf_size : 200;
   std::vector<int> all(n), rev;
   std::iota(all.begin(), all.end(), 0);
   rev.reserve(n);
   int next_id = 0;
   std::vector<std::vector<int>> seps;
   nd_recurse(W, std::move(all), next_id, rev, seps);
   if ((int)rev.size() != n) return SSIDS_ERROR_UNKNOWN;
   std::vector<char> seen(n, 0);
   for (int k = 0; k < n; ++k) {
      const int v = rev[n - 1 - k];
      if (v < 0 || v >= n || seen[v]) return SSIDS_ERROR_UNKNOWN;
      seen[v] = 1;
      order[v] = k;
   }
   return 0;
}

}  // namespace b200
