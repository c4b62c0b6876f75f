// Host side of the SLS interface for solver 'b200ldlt' (include/b200_sls.h).  C++ restatement of the
// solver-neutral logic of src/sls/sls.f90 for the 'ssids' class of back ends:
//   sls_initialize          SLS_initialize / SLS_initialize_solver   sls.f90:817, 959
//   sls_analyse_matrix      SLS_analyse                              sls.f90:2178, 2853-2968, 3115-3150
//   coord_to_sorted_csr     SLS_coord_to_sorted_csr                  sls.f90:8409-8578
//   sls_factorize_matrix    SLS_factorize (value gather, posdef)     sls.f90:3521, 4106-4147, 4273-4297
//   sls_solve_system        SLS_solve_ir (refinement loop)           sls.f90:4692, 4806-4947
//   sls_partial_solve_...   SLS_part_solve                           sls.f90:6551 (disabled for 'ssids'
//                           at :6886-6888; enabled here)
//   control / inform maps   SLS_copy_control_to_ssids :1385-1439, SLS_copy_inform_from_ssids :1737-1786
#include <algorithm>
#include <cctype>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200_sls.h"

namespace {

struct SlsData {
   std::string solver;
   sls_control_type control;
   sls_inform_type inform;
   int n = 0, ne = 0;          // user's matrix
   int type = 0;               // 0 coordinate, 1 sparse_by_rows, 2 dense
   std::vector<int> urow, ucol;   // user's entries as 1-based coordinates (uploaded for the residuals of the refinement loop)
   std::vector<int> maps, ptr, col, order;   // 1-based, as data%MAPS, data%matrix%PTR/COL, data%ORDER
   std::vector<double> val;
   void* akeep = nullptr;
   void* fkeep = nullptr;
   bool analysed = false, factorized = false, must_be_definite = false, map_set = false, scale_set = false;
};

double now() {
   return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

std::string lower(const char* s) {
   std::string r(s ? s : "");
   while (!r.empty() && r.back() == ' ') r.pop_back();
   for (auto& c : r) c = (char)std::tolower((unsigned char)c);
   return r;
}

b200ldlt_options to_solver(const sls_control_type& c) {   // SLS_copy_control_to_ssids
   b200ldlt_options o;
   b200ldlt_default_options(&o);
   o.array_base = 1;
   o.print_level = c.print_level_solver > 0 ? c.print_level_solver : -1;
   o.nemin = c.node_amalgamation;
   o.small_ = c.absolute_pivot_tolerance;
   o.device = c.device;
   if (c.pivot_control == 2) { o.u = 0.0; o.action = 1; }
   else if (c.pivot_control == 3) { o.u = 0.0; o.action = 0; }
   else if (c.pivot_control == 4) { o.u = 0.0; o.action = 1; }
   else { o.u = c.relative_pivot_tolerance; o.action = 1; }
   return o;
}

void from_solver(sls_inform_type& inf, const b200ldlt_inform& s) {   // SLS_copy_inform_from_ssids
   inf.b200ldlt_inform = s;
   inf.status = s.flag;
   if (s.flag >= 0) {
      inf.status = GALAHAD_OK;
      inf.two_by_two_pivots = s.num_two;
      inf.rank = s.matrix_rank;
      inf.negative_eigenvalues = s.num_neg;
      inf.delayed_pivots = s.num_delay;
      inf.entries_in_factors = s.num_factor;
      inf.flops_elimination = s.num_flops;
      inf.max_front_size = s.maxfront;
      inf.max_depth_assembly_tree = s.maxdepth;
      inf.nodes_assembly_tree = s.num_sup;
      return;
   }
   switch (s.flag) {
      case -30: inf.status = GALAHAD_ERROR_ALLOCATE; inf.alloc_status = s.stat; break;
      case -31: inf.status = GALAHAD_ERROR_DEALLOCATE; inf.alloc_status = s.stat; break;
      case -1: case -2: case -3: case -4: case -5: case -6: case -9: case -10: case -12: case -13: case -14: case -15:
         inf.status = GALAHAD_ERROR_RESTRICTIONS; break;
      case -11: inf.status = GALAHAD_ERROR_PERMUTATION; break;
      case -7: case -8: inf.status = GALAHAD_ERROR_INERTIA; break;
      case -32: case GALAHAD_UNAVAILABLE_OPTION: inf.status = GALAHAD_UNAVAILABLE_OPTION; break;
      case GALAHAD_ERROR_UNKNOWN_SOLVER: inf.status = GALAHAD_ERROR_UNKNOWN_SOLVER; break;
      default: inf.status = GALAHAD_ERROR_TECHNICAL;
   }
}

}  // namespace

extern "C" {

// sls.f90:8409-8578; all arrays 1-based as in the Fortran
void sls_coord_to_sorted_csr(int n, int ne, const int row[], const int col[], int map[], int ptr[],
                             int* dup, int* oor, int* upper, int* missing_diagonals) {
   std::vector<int> iw(n + 2, 0);
   *upper = 0;
   for (int i = 2; i <= n + 1; ++i) iw[i] = 1;
   for (int l = 1; l <= ne; ++l) {
      int i = row[l - 1], j = col[l - 1];
      if (i < 1 || i > n || j < 1 || j > n) continue;
      if (i < j) { iw[i + 1]++; (*upper)++; }
      else if (i == j) iw[i + 1]++;
      else iw[j + 1]++;
   }
   iw[1] = 1;
   for (int i = 1; i <= n; ++i) iw[i + 1] += iw[i];
   int size = iw[n + 1] - 1;
   std::vector<int> map2(size + 1, 0);
   for (int i = 1; i <= n; ++i) { map2[iw[i]] = 0; iw[i]++; }
   *oor = 0;
   for (int l = 1; l <= ne; ++l) {
      int i = row[l - 1], j = col[l - 1];
      if (i < 1 || i > n || j < 1 || j > n) { (*oor)++; map[l - 1] = 0; }
      else if (i <= j) { map[l - 1] = iw[i]; map2[iw[i]] = l; iw[i]++; }
      else { map[l - 1] = iw[j]; map2[iw[j]] = l; iw[j]++; }
   }
   for (int i = n; i >= 1; --i) iw[i + 1] = iw[i];
   iw[1] = 1;
   *dup = 0;
   *missing_diagonals = 0;
   ptr[0] = 1;
   std::vector<std::pair<int, int>> ents;   // (column, entry) -- the Fortran quicksorts COLS with ENTS along
   for (int i = 1; i <= n; ++i) {
      int sz = iw[i + 1] - iw[i] - 1;
      int pt = ptr[i - 1];
      if (sz > 0) {
         ents.clear();
         for (int ll = iw[i] + 1; ll <= iw[i + 1] - 1; ++ll) {
            int l = map2[ll];
            ents.push_back({std::max(row[l - 1], col[l - 1]), l});
         }
         // any sort by column gives the same MAP: among duplicates the smallest entry index wins
         std::sort(ents.begin(), ents.end());
         int j_old = i - 1, ll = 0;
         if (ents[0].first > i) { pt++; (*missing_diagonals)++; }
         for (int k = 0; k < sz; ++k) {
            int l = ents[k].second, jj = ents[k].first;
            if (jj > j_old) { j_old = jj; map[l - 1] = pt; ll = l; pt++; }
            else {
               (*dup)++;
               if (l < ll) { map[l - 1] = pt - 1; map[ll - 1] = -pt + 1; ll = l; }
               else map[l - 1] = -pt + 1;
            }
         }
      } else {
         pt++;
      }
      ptr[i] = pt;
   }
}

void sls_initialize(const char solver[], void** data, sls_control_type* control, int* status) {
   SlsData* d = new SlsData();
   d->solver = lower(solver);
   std::memset(control, 0, sizeof(*control));
   control->f_indexing = false;
   control->error = control->warning = control->out = 6;
   control->node_amalgamation = 32;
   control->pivot_control = 1;
   control->ordering = 0;
   control->scaling = 0;
   control->relative_pivot_tolerance = 0.01;
   control->absolute_pivot_tolerance = DBL_EPSILON;
   control->max_iterative_refinements = 0;
   control->acceptable_residual_relative = 10.0 * DBL_EPSILON;
   control->acceptable_residual_absolute = 10.0 * DBL_EPSILON;
   control->multiple_rhs = true;
   control->device = -1;
   std::strcpy(control->prefix, "\"\"");
   std::memset(&d->inform, 0, sizeof(d->inform));
   *data = d;
   if (d->solver != "b200ldlt") {          // SLS_initialize_solver, sls.f90:1054-1055
      *status = d->inform.status = GALAHAD_ERROR_UNKNOWN_SOLVER;
      return;
   }
   control->ordering = -1;                  // as for 'ssids': the solver's default ordering (sls.f90:887-892)
   d->control = *control;
   *status = d->inform.status = GALAHAD_OK;
}

void sls_analyse_matrix_perm(sls_control_type* control, void** data, int* status, int n, const char type[], int ne,
                             const int row[], const int col[], const int ptr[], const int perm[]) {
   SlsData* d = static_cast<SlsData*>(*data);
   double t0 = now();
   d->control = *control;
   if (d->solver != "b200ldlt") { *status = d->inform.status = GALAHAD_ERROR_UNKNOWN_SOLVER; return; }
   std::string ty = lower(type);
   const int base = control->f_indexing ? 1 : 0;
   if (n <= 0 || (ty != "coordinate" && ty != "sparse_by_rows" && ty != "dense")) {
      *status = d->inform.status = GALAHAD_ERROR_RESTRICTIONS;
      return;
   }
   d->n = n;
   d->type = ty == "coordinate" ? 0 : ty == "sparse_by_rows" ? 1 : 2;
   d->must_be_definite = control->pivot_control == 2 || control->pivot_control == 3;   // sls.f90:2227-2228
   // user's lower triangle as 1-based coordinates (sls.f90:2869-2909)
   d->urow.clear(); d->ucol.clear();
   if (d->type == 0) {
      d->ne = ne;
      for (int l = 0; l < ne; ++l) { d->urow.push_back(row[l] - base + 1); d->ucol.push_back(col[l] - base + 1); }
   } else if (d->type == 1) {
      d->ne = ptr[n] - base;
      for (int i = 0; i < n; ++i)
         for (int l = ptr[i] - base; l < ptr[i + 1] - base; ++l) { d->urow.push_back(i + 1); d->ucol.push_back(col[l] - base + 1); }
   } else {
      d->ne = n * (n + 1) / 2;
      for (int i = 1; i <= n; ++i) for (int j = 1; j <= i; ++j) { d->urow.push_back(i); d->ucol.push_back(j); }
   }
   d->maps.assign(d->ne, 0);
   d->ptr.assign(n + 1, 0);
   sls_inform_type& inf = d->inform;
   std::memset(&inf, 0, sizeof(inf));
   inf.entries = d->ne;
   sls_coord_to_sorted_csr(n, d->ne, d->urow.data(), d->ucol.data(), d->maps.data(), d->ptr.data(), &inf.duplicates,
                           &inf.out_of_range, &inf.upper, &inf.missing_diagonals);
   if (d->must_be_definite && inf.missing_diagonals > 0) {   // sls.f90:2326-2329, 2739-2742
      d->analysed = false;
      *status = inf.status = GALAHAD_ERROR_INERTIA;
      return;
   }
   const int nz = d->ptr[n] - 1;
   d->col.assign(nz, 0);
   d->val.assign(nz, 0.0);
   for (int i = 1; i <= n; ++i) d->col[d->ptr[i - 1] - 1] = i;
   for (int l = 0; l < d->ne; ++l) {
      int k = d->maps[l];
      if (k > 0) d->col[k - 1] = std::max(d->urow[l], d->ucol[l]);
   }
   // The order.  PERM if present (sls.f90:2962-2968).  Otherwise control%ordering decides (sls.f90:2262-2268, 2395-2412,
   // 3134-3140): >= 0 is MC68's menu -- 0, 1 and anything above 6 approximate minimum degree, 2 minimum degree, 3 nested
   // dissection, 4 (indefinite 1x1 / 2x2 ordering), 5 and 6 (MC61 profile / band) have nothing behind them here; -1 asks the
   // solver for its own, which for this solver, as for SSIDS with METIS, is nested dissection.  The orderings are this
   // library's (ordering.cpp): the reference's come from packages outside its tree, no permutation is pinned.
   b200ldlt_options o = to_solver(d->control);
   d->order.assign(n, 0);
   if (perm) {
      for (int i = 0; i < n; ++i) d->order[i] = perm[i] - base + 1;
   } else {
      const int oc = d->control.ordering;
      int kind;
      if (oc < -1) {
         // SSIDS options%ordering = -oc.  2 is the matching-based order, which needs the values: SLS_analyse has none to
         // pass (sls.f90:3134-3139), so the reference ends in SSIDS_ERROR_VAL (ssids.f90:230-238) and anything beyond in
         // SSIDS_ERROR_ORDER (:222-227), mapped as every SSIDS flag is.  With values in hand the order is
         // b200ldlt_analyse_matched's.
         b200ldlt_inform e = {};
         e.flag = oc == -2 ? -9 : -8;
         from_solver(inf, e);
         d->analysed = false;
         *status = inf.status;
         return;
      }
      if (oc == -1 || oc == 3) kind = 3;
      else if (oc == 4 || oc == 5 || oc == 6) kind = 0;
      else kind = 1;
      std::vector<int64_t> p64(d->ptr.begin(), d->ptr.end());
      const int rc = kind ? b200ldlt_order(n, p64.data(), d->col.data(), kind, d->order.data(), &o) : GALAHAD_UNAVAILABLE_OPTION;
      if (rc < 0) {
         d->analysed = false;
         *status = inf.status = rc == GALAHAD_UNAVAILABLE_OPTION ? GALAHAD_UNAVAILABLE_OPTION : GALAHAD_ERROR_TECHNICAL;
         return;
      }
   }
   b200ldlt_inform si;
   if (d->akeep) b200ldlt_free(&d->akeep, &d->fkeep);
   b200ldlt_analyse(&d->akeep, n, d->ptr.data(), d->col.data(), d->order.data(), &o, &si);
   int dup = inf.duplicates, oor = inf.out_of_range, up = inf.upper, md = inf.missing_diagonals, ents = inf.entries;
   from_solver(inf, si);
   inf.duplicates = dup; inf.out_of_range = oor; inf.upper = up; inf.missing_diagonals = md; inf.entries = ents;
   inf.real_size_factors = inf.integer_size_factors = -1;   // SSIDS leaves them unset too (sbls.f90:4132)
   d->analysed = inf.status == GALAHAD_OK;
   d->map_set = false;          // new MAPS: uploaded again by the next factorize
   d->scale_set = false;
   d->factorized = false;
   inf.time.clock_analyse = now() - t0;
   inf.time.analyse_external = si.time_analyse;
   inf.time.clock_total += inf.time.clock_analyse;
   *status = inf.status;
}

void sls_analyse_matrix(sls_control_type* control, void** data, int* status, int n, const char type[], int ne,
                        const int row[], const int col[], const int ptr[]) {
   sls_analyse_matrix_perm(control, data, status, n, type, ne, row, col, ptr, nullptr);
}

void sls_reset_control(sls_control_type* control, void** data, int* status) {
   SlsData* d = static_cast<SlsData*>(*data);
   d->control = *control;
   *status = GALAHAD_OK;
}

void sls_factorize_matrix(void** data, int* status, int ne, const double val[]) {
   SlsData* d = static_cast<SlsData*>(*data);
   sls_inform_type& inf = d->inform;
   if (!d->analysed) { *status = inf.status = GALAHAD_ERROR_RESTRICTIONS; return; }
   (void)ne;
   double t0 = now();
   const int n = d->n;
   // pivot_control 2/3 -> must_be_definite -> Cholesky path (sls.f90:2227-2228, 4282)
   const int posdef = (d->control.pivot_control == 2 || d->control.pivot_control == 3) ? 1 : 0;
   b200ldlt_options o = to_solver(d->control);
   b200ldlt_inform si;
   // value gather (sls.f90:4106-4147: diagonals zeroed first, duplicates summed) on the device, through the MAPS
   // uploaded once per analyse
   if (!d->map_set) {
      int mf = b200ldlt_set_value_map(d->akeep, &d->fkeep, d->ne, d->maps.data(), &o, &si);
      if (mf >= 0) mf = b200ldlt_set_user_matrix(d->akeep, &d->fkeep, d->ne, d->urow.data(), d->ucol.data(), &o, &si);
      d->map_set = mf >= 0;
   }
   // control%scaling (sls.f90:3558, 3766-3860, 1405-1413).  3: SLS computes symmetric infinity-norm equilibration factors
   // itself (MC77 with job 0 there, an HSL stub in a stock build) and factorises S A S: here the factors come from
   // the engine's equilibration (b200ldlt_equilib_scale, the restatement of spral/scaling.f90 inf_norm_equilib_sym --
   // the same Ruiz iteration MC77 runs in that norm), computed from the gathered values.  -1: SSIDS's own matching-based
   // scaling by the Hungarian algorithm (options%scaling = 1, sls.f90:1405-1413 -> ssids.f90:928-960), the restatement
   // of spral/scaling.f90 hungarian_scale_sym behind b200ldlt_hungarian_scale; -2: SSIDS's auction scaling
   // (options%scaling = 2, auction_scale_sym).  -3 (the scaling SSIDS saves with its matching-based order): that order
   // cannot be had through SLS_analyse (no values, see above), so as in the reference nothing was saved:
   // SSIDS_ERROR_NO_SAVED_SCALING (-15, ssids.f90:990-995) -> error_restrictions.  1 (MC64) and 2 (MC77 one-norm) have no
   // implementation behind them.
   const int sc = d->control.scaling;
   if (sc == -3) {
      *status = inf.status = GALAHAD_ERROR_RESTRICTIONS;
      d->factorized = false;
      return;
   }
   if (sc != 0 && sc != 3 && sc != -1 && sc != -2) {
      *status = inf.status = GALAHAD_UNAVAILABLE_OPTION;
      d->factorized = false;
      return;
   }
   if (d->map_set && (sc == 3 || sc == -1 || sc == -2)) {
      const int nz = d->ptr[n] - 1;
      std::vector<double> hv(nz, 0.0), sf(n, 1.0);
      for (int l = 0; l < d->ne; ++l) {          // the gather of sls.f90:4106-4147 on the host, for the factors only
         const int k = d->maps[l];
         if (k > 0) hv[k - 1] += val[l]; else if (k < 0) hv[-k - 1] += val[l];
      }
      std::vector<int64_t> p0(n + 1);
      std::vector<int> r0(nz);
      for (int i = 0; i <= n; ++i) p0[i] = d->ptr[i] - 1;
      for (int k = 0; k < nz; ++k) r0[k] = d->col[k] - 1;
      if (sc == 3) b200ldlt_equilib_scale(n, p0.data(), r0.data(), hv.data(), sf.data(), nullptr);
      else if (sc == -2) b200ldlt_auction_scale(n, p0.data(), r0.data(), hv.data(), sf.data(), nullptr, nullptr, nullptr);
      else {
         // the facade holds the UPPER triangle by rows = the lower triangle by columns (sls.f90:8409-8578)
         const int hf = b200ldlt_hungarian_scale(n, p0.data(), r0.data(), hv.data(), sf.data(), nullptr, nullptr, o.action);
         if (hf == -2) {       // structurally singular and pivot_control = 3 (action = .false.): ssids.f90:946-949
            *status = inf.status = GALAHAD_ERROR_RESTRICTIONS;  // SSIDS_ERROR_SINGULAR (-5) through SLS_copy_inform_from_ssids
            d->factorized = false;
            return;
         }
      }
      if (b200ldlt_set_scaling(d->akeep, &d->fkeep, sf.data(), &o, &si) < 0) d->map_set = false;
      d->scale_set = d->map_set;
   } else if (d->map_set && d->scale_set) {
      b200ldlt_set_scaling(d->akeep, &d->fkeep, nullptr, &o, &si);
      d->scale_set = false;
   }
   if (d->map_set) b200ldlt_factor_mapped(d->akeep, &d->fkeep, posdef, d->ne, val, &o, &si);
   // (set_value_map only fails without a CUDA device; si then carries -51 and there is no other path)
   int dup = inf.duplicates, oor = inf.out_of_range, up = inf.upper, md = inf.missing_diagonals, ents = inf.entries;
   sls_time_type tm = inf.time;
   from_solver(inf, si);
   inf.duplicates = dup; inf.out_of_range = oor; inf.upper = up; inf.missing_diagonals = md; inf.entries = ents;
   inf.time = tm;
   inf.num_zero = d->n - si.matrix_rank;
   // (a matrix that is not positive definite where that was required, SSIDS flag -6, reads GALAHAD_ERROR_RESTRICTIONS
   // through the table above exactly as SLS_copy_inform_from_ssids has it, sls.f90:1768-1769)
   inf.real_size_factors = inf.integer_size_factors = -1;
   d->factorized = inf.status == GALAHAD_OK;
   inf.time.clock_factorize = now() - t0;
   inf.time.factorize_external = si.time_factor;
   inf.time.clock_total += inf.time.clock_factorize;
   *status = inf.status;
}

static int solver_solve(SlsData* d, int job, double* x) {
   b200ldlt_options o = to_solver(d->control);
   b200ldlt_inform si;
   int flag = b200ldlt_solve(d->akeep, d->fkeep, job, 1, x, d->n, &o, &si);
   d->inform.b200ldlt_inform.time_solve = si.time_solve;
   d->inform.time.solve_external += si.time_solve;
   return flag;
}

void sls_solve_system(void** data, int* status, int n, double sol[]) {
   SlsData* d = static_cast<SlsData*>(*data);
   sls_inform_type& inf = d->inform;
   if (!d->factorized || n != d->n) { *status = inf.status = GALAHAD_ERROR_RESTRICTIONS; return; }
   double t0 = now();
   inf.iterative_refinements = 0;
   int flag;
   if (d->control.max_iterative_refinements <= 0) {
      flag = solver_solve(d, 0, sol);
   } else {
      // SLS_solve_ir (sls.f90:4751-4947): x = 0; r = b; repeat { solve A dx = r; x += dx; r = b - A x }, on the
      // device: b goes up once, x comes back once, the residual uses the entries uploaded by factorize
      b200ldlt_options o = to_solver(d->control);
      b200ldlt_inform si;
      int nref = 0;
      flag = b200ldlt_solve_refined(d->akeep, d->fkeep, sol, d->control.max_iterative_refinements,
                                    d->control.acceptable_residual_absolute, d->control.acceptable_residual_relative, &nref, &o, &si);
      inf.iterative_refinements = nref;
      d->inform.b200ldlt_inform.time_solve = si.time_solve;
      d->inform.time.solve_external += si.time_solve;
   }
   inf.status = flag >= 0 ? GALAHAD_OK : (flag == -11 ? GALAHAD_ERROR_PERMUTATION : GALAHAD_ERROR_TECHNICAL);
   inf.time.clock_solve = now() - t0;
   inf.time.clock_total += inf.time.clock_solve;
   *status = inf.status;
}

void sls_partial_solve_system(const char part[], void** data, int* status, int n, double sol[]) {
   SlsData* d = static_cast<SlsData*>(*data);
   sls_inform_type& inf = d->inform;
   if (!d->factorized || n != d->n) { *status = inf.status = GALAHAD_ERROR_RESTRICTIONS; return; }
   std::string p = lower(part);
   int job = p == "l" ? B200LDLT_JOB_FWD : p == "d" ? B200LDLT_JOB_DIAG : p == "u" ? B200LDLT_JOB_BWD : -1;
   if (job < 0) { *status = inf.status = GALAHAD_UNAVAILABLE_OPTION; return; }   // 'S' needs sqrt(D): indefinite D has none
   if (job == B200LDLT_JOB_DIAG && (d->control.pivot_control == 2 || d->control.pivot_control == 3)) {
      *status = inf.status = GALAHAD_OK;   // LL^T: the D part is the identity
      return;
   }
   int flag = solver_solve(d, job, sol);
   inf.status = flag >= 0 ? GALAHAD_OK : GALAHAD_ERROR_TECHNICAL;
   *status = inf.status;
}

void sls_information(void** data, sls_inform_type* inform, int* status) {
   SlsData* d = static_cast<SlsData*>(*data);
   *inform = d->inform;
   *status = GALAHAD_OK;
}

void sls_terminate(void** data, sls_control_type* control, sls_inform_type* inform) {
   (void)control;
   if (!data || !*data) return;
   SlsData* d = static_cast<SlsData*>(*data);
   b200ldlt_free(&d->akeep, &d->fkeep);
   if (inform) { *inform = d->inform; inform->status = GALAHAD_OK; }
   delete d;
   *data = nullptr;
}

}  // extern "C"
