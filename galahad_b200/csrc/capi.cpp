// C ABI of the b200ldlt engine (include/b200ldlt.h).  Thin: argument checking, index-base
// shifts, handle ownership, flag bookkeeping.  Mirrors the call-sequence checks of
// src/ssids/ssids.f90 (analyse :148, factor :770/:821-827, solve :1139/:1155-1195,
// enquire :1255/:1299, alter :1347, free :1388).
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/b200ldlt.h"
#include "analyse.hpp"
#include "engine.hpp"

using namespace b200;

namespace {

// An Akeep is shared between the caller's handle and every Fkeep created from it (the Numeric keeps pointers into
// its Plan): the last owner destroys it.  Re-analysing into a live handle or freeing akeep before fkeep therefore
// never leaves a Numeric with a dangling Plan.
struct Akeep {
   std::atomic<int> refs{1};
   Symbolic sym;      // the reference's assembly tree (what SLS reports)
   Symbolic nsym;     // the engine's tree: chains fused (analyse.cpp fuse_chains)
   Plan* plan = nullptr;
   int plan_cuda_error = 0;
   b200ldlt_inform inf_analyse;
   // partition
   std::vector<int64_t> part_flops;
   std::vector<double> match_scale;   // scaling saved by a matching-based order (akeep%scaling, ssids.f90:283-286)
};

struct Fkeep {
   Akeep* ak = nullptr;
   Numeric* num = nullptr;
};

void akeep_release(Akeep* ak) {
   if (!ak) return;
   if (ak->refs.fetch_sub(1) == 1) {
      if (ak->plan) plan_destroy(ak->plan);
      delete ak;
   }
}

void clear_inform(b200ldlt_inform* inf) {
   if (inf) std::memset(inf, 0, sizeof(*inf));
}

int ensure_plan(Akeep* ak, int device) {
   if (ak->plan) return 0;
   ak->plan = plan_create(ak->nsym, device, &ak->plan_cuda_error);
   return ak->plan ? 0 : SSIDS_ERROR_CUDA_UNKNOWN;
}

FactorOptions to_engine(const b200ldlt_options* o, int posdef) {
   FactorOptions f;
   f.posdef = posdef;
   f.action = o->action;
   f.u = o->u;
   f.small_ = o->small_;
   f.multiplier = o->multiplier;
   f.print_level = o->print_level;
   f.profile = o->reserved[0];
   return f;
}

}  // namespace

extern "C" {

void b200ldlt_default_options(b200ldlt_options* o) {
   std::memset(o, 0, sizeof(*o));
   o->array_base = 0;
   o->print_level = -1;
   o->nemin = 32;          // datatypes.f90:21 nemin_default
   o->action = 1;
   o->u = 0.01;
   o->small_ = 1e-20;
   o->multiplier = 1.1;
   o->device = -1;
}

const char* b200ldlt_version(void) { return "b200ldlt 0.1 (sm_100a)"; }

int b200ldlt_analyse_ptr64(void** akeep, int n, const int64_t* ptr, const int* row, int* order,
                           const b200ldlt_options* opt, b200ldlt_inform* inf) {
   clear_inform(inf);
   if (!akeep || !opt || !inf) return SSIDS_ERROR_CALL_SEQUENCE;
   if (n < 0) return inf->flag = SSIDS_ERROR_A_N_OOR;
   if (n > 0 && (!ptr || !row || !order)) return inf->flag = SSIDS_ERROR_A_PTR;
   auto t0 = std::chrono::steady_clock::now();
   const int base = opt->array_base ? 1 : 0;
   if (n > 0 && ptr[0] != base) return inf->flag = SSIDS_ERROR_A_PTR;
   for (int i = 0; i < n; ++i)
      if (ptr[i + 1] < ptr[i]) return inf->flag = SSIDS_ERROR_A_PTR;
   Akeep* ak = new (std::nothrow) Akeep();
   if (!ak) return inf->flag = SSIDS_ERROR_ALLOCATION;
   int flag;
   try {
      const int64_t nz = n > 0 ? ptr[n] - base : 0;
      std::vector<int64_t> p0(n + 1);
      std::vector<int> r0(nz), o0(n);
      for (int i = 0; i <= n; ++i) p0[i] = ptr[i] - base;
      for (int64_t k = 0; k < nz; ++k) {
         r0[k] = row[k] - base;
         if (r0[k] < 0 || r0[k] >= n) { delete ak; return inf->flag = SSIDS_ERROR_A_ALL_OOR; }
      }
      for (int i = 0; i < n; ++i) {
         int v = order[i] < 0 ? -order[i] : order[i];   // check_order: signs are reset (anal.f90:176-178)
         o0[i] = v - base;
      }
      flag = analyse(n, p0.data(), r0.data(), o0.data(), opt->nemin, ak->sym);
   } catch (const std::bad_alloc&) {
      delete ak;
      return inf->flag = SSIDS_ERROR_ALLOCATION;
   }
   if (flag < 0) {
      delete ak;
      return inf->flag = flag;
   }
   {
      const char* fz = std::getenv("B200LDLT_FUSE_ZERO_FRAC");
      double frac = fz ? std::atof(fz) : 0.05;
      if (opt->reserved[1] == 1) frac = -1.0;     // reserved[1] = 1: engine runs on the reference tree as is
      fuse_chains(ak->sym, ak->nsym, frac);
   }
   const Symbolic& s = ak->sym;
   for (int i = 0; i < n; ++i) order[i] = s.order[i] < 0 ? (base - 1) : s.order[i] + base;
   inf->flag = flag;
   inf->matrix_rank = s.matrix_rank;
   inf->maxdepth = s.maxdepth;
   inf->maxfront = s.maxfront;
   inf->num_factor = s.num_factor;
   inf->num_flops = s.num_flops;
   inf->num_sup = s.nnodes;
   inf->nparts = 1;
   inf->time_analyse = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
   ak->inf_analyse = *inf;
   if (*akeep) akeep_release(static_cast<Akeep*>(*akeep));   // an fkeep made from it keeps it alive until it is freed
   *akeep = ak;
   return flag;
}

int b200ldlt_order(int n, const int64_t* ptr, const int* row, int kind, int* order, const b200ldlt_options* opt) {
   if (n < 0) return SSIDS_ERROR_A_N_OOR;
   if (n == 0) return 0;
   if (!ptr || !row || !order || !opt) return SSIDS_ERROR_A_PTR;
   const int base = opt->array_base ? 1 : 0;
   try {
      const int64_t nz = ptr[n] - base;
      std::vector<int64_t> p0(n + 1);
      std::vector<int> r0((size_t)std::max<int64_t>(nz, 0));
      for (int i = 0; i <= n; ++i) p0[i] = ptr[i] - base;
      for (int64_t k = 0; k < nz; ++k) r0[k] = row[k] - base;
      int rc;
      if (kind == 1) rc = amd_order(n, p0.data(), r0.data(), order);
      else if (kind == 3) rc = nd_order(n, p0.data(), r0.data(), order);
      else return SSIDS_ERROR_UNIMPLEMENTED;
      if (rc < 0) return rc;
      if (base) for (int i = 0; i < n; ++i) order[i] += 1;
      return 0;
   } catch (const std::bad_alloc&) {
      return SSIDS_ERROR_ALLOCATION;
   }
}

int b200ldlt_analyse(void** akeep, int n, const int* ptr, const int* row, int* order,
                     const b200ldlt_options* opt, b200ldlt_inform* inf) {
   if (n < 0 || !ptr) {
      clear_inform(inf);
      if (inf) inf->flag = n < 0 ? SSIDS_ERROR_A_N_OOR : SSIDS_ERROR_A_PTR;
      return n < 0 ? SSIDS_ERROR_A_N_OOR : SSIDS_ERROR_A_PTR;
   }
   std::vector<int64_t> p64(ptr, ptr + n + 1);
   return b200ldlt_analyse_ptr64(akeep, n, p64.data(), row, order, opt, inf);
}

// plan and numeric handle for this akeep (created on first use)
static int ensure_fkeep(Akeep* ak, void** fkeep, const b200ldlt_options* opt, b200ldlt_inform* inf) {
   if (ensure_plan(ak, opt->device)) {
      inf->cuda_error = ak->plan_cuda_error;
      return inf->flag = SSIDS_ERROR_CUDA_UNKNOWN;
   }
   Fkeep* fk = static_cast<Fkeep*>(*fkeep);
   if (fk && fk->ak != ak) { b200ldlt_free_fkeep(fkeep); fk = nullptr; }
   if (!fk) {
      fk = new (std::nothrow) Fkeep();
      if (!fk) return inf->flag = SSIDS_ERROR_ALLOCATION;
      fk->ak = ak;
      ak->refs.fetch_add(1);
      fk->num = numeric_create(ak->plan);
      *fkeep = fk;
   }
   return 0;
}

// mapped_ne >= 0: `val` holds the caller's mapped_ne values in input order, gathered on the device through the
// value map (b200ldlt_set_value_map)
static int factor_common(void* akeep, void** fkeep, int posdef, const double* val, bool on_device,
                         const b200ldlt_options* opt, b200ldlt_inform* inf, int mapped_ne = -1, int nchg = -1,
                         const int* which = nullptr) {
   if (!akeep || !fkeep || !opt || !inf) return SSIDS_ERROR_CALL_SEQUENCE;
   Akeep* ak = static_cast<Akeep*>(akeep);
   *inf = ak->inf_analyse;
   inf->time_analyse = 0;
   if (ak->sym.n > 0 && !val && nchg != 0) return inf->flag = SSIDS_ERROR_VAL;
   if (int rc = ensure_fkeep(ak, fkeep, opt, inf)) return rc;
   Fkeep* fk = static_cast<Fkeep*>(*fkeep);
   FactorStats st;
   int flag = nchg >= 0      ? numeric_factor_changed(fk->num, nchg, which, val, to_engine(opt, posdef), st)
              : mapped_ne >= 0 ? numeric_factor_mapped(fk->num, mapped_ne, val, to_engine(opt, posdef), st)
                               : numeric_factor(fk->num, val, on_device, to_engine(opt, posdef), st);
   inf->cuda_error = st.cuda_error;
   inf->num_delay = st.num_delay;
   inf->num_neg = st.num_neg;
   inf->num_two = st.num_two;
   inf->not_first_pass = st.not_first_pass;
   inf->not_second_pass = st.not_second_pass;
   inf->maxfront = st.maxfront > 0 ? st.maxfront : inf->maxfront;
   inf->matrix_rank = ak->sym.matrix_rank - st.num_zero;   // cpu_iface.f90:71-88
   inf->time_factor = st.t_total;
   inf->time_factor_device = st.t_device;
   inf->factor_bytes = st.factor_bytes;
   inf->reserved[0] = (int)st.launches;
   inf->reserved[2] = (int)st.upd_launches;
   inf->upd_flops = st.upd_flops;
   inf->upd_seconds = st.upd_seconds;
   // singular: warning with action, error without (ssids.f90:1057-1066)
   if (flag >= 0 && st.num_zero > 0) flag = opt->action ? SSIDS_WARNING_FACT_SINGULAR : SSIDS_ERROR_SINGULAR;
   if (flag >= 0 && flag == 0 && ak->inf_analyse.flag > 0) flag = ak->inf_analyse.flag;
   return inf->flag = flag;
}

int b200ldlt_factor(void* akeep, void** fkeep, int posdef, const double* val,
                    const b200ldlt_options* opt, b200ldlt_inform* inf) {
   return factor_common(akeep, fkeep, posdef, val, false, opt, inf);
}
int b200ldlt_factor_device(void* akeep, void** fkeep, int posdef, const double* d_val,
                           const b200ldlt_options* opt, b200ldlt_inform* inf) {
   return factor_common(akeep, fkeep, posdef, d_val, true, opt, inf);
}

int b200ldlt_set_value_map(void* akeep, void** fkeep, int ne, const int* maps, const b200ldlt_options* opt,
                           b200ldlt_inform* inf) {
   if (!akeep || !fkeep || !opt || !inf || ne < 0 || (ne > 0 && !maps)) return SSIDS_ERROR_CALL_SEQUENCE;
   Akeep* ak = static_cast<Akeep*>(akeep);
   *inf = ak->inf_analyse;
   if (int rc = ensure_fkeep(ak, fkeep, opt, inf)) return rc;
   return inf->flag = numeric_set_value_map(static_cast<Fkeep*>(*fkeep)->num, ne, maps);
}
// ssids_factor with its optional SCALE argument and options%scaling (ssids.f90:900-1030)
// the Hungarian scaling alone (tests, callers that keep the scaling over several factorisations)
int b200ldlt_hungarian_scale(int n, const int64_t* ptr, const int* row, const double* val, double* scaling, int* match,
                             int* matched, int scale_if_singular) {
   if (n < 0 || !ptr || !scaling || (n > 0 && ptr[n] > 0 && (!row || !val))) return SSIDS_ERROR_CALL_SEQUENCE;
   return hungarian_scale_sym(n, ptr, row, val, scaling, scale_if_singular != 0, match, matched);
}
int b200ldlt_auction_scale(int n, const int64_t* ptr, const int* row, const double* val, double* scaling, int* match,
                           int* matched, int* iterations) {
   if (n < 0 || !ptr || !scaling || (n > 0 && ptr[n] > 0 && (!row || !val))) return SSIDS_ERROR_CALL_SEQUENCE;
   return auction_scale_sym(n, ptr, row, val, scaling, match, matched, iterations);
}
// 0-based entry points of the matching-based order alone (tests, callers with their own analyse)
int b200ldlt_match_order(int n, const int64_t* ptr, const int* row, const double* val, int kind, int* order, double* scale,
                         int* pairs, int* ncomp, int64_t* cptr, int* crow) {
   if (n < 0 || !ptr || !order || (n > 0 && ptr[n] > 0 && (!row || !val))) return SSIDS_ERROR_CALL_SEQUENCE;
   try {
      std::vector<int64_t> cp;
      std::vector<int> cr;
      const int flag = match_order(n, ptr, row, val, kind, order, scale, pairs, &cp, &cr);
      if (flag < 0) return flag;
      if (ncomp) *ncomp = cp.empty() ? 0 : (int)cp.size() - 1;
      if (cptr) std::copy(cp.begin(), cp.end(), cptr);
      if (crow) std::copy(cr.begin(), cr.end(), crow);
      return flag;
   } catch (const std::bad_alloc&) {
      return SSIDS_ERROR_ALLOCATION;
   }
}

// ssids_analyse with options%ordering = 2 (ssids.f90:283-286, 321-358): the order comes from the matching on the values,
// its scaling stays in akeep for options%scaling = 3.
int b200ldlt_analyse_matched(void** akeep, int n, const int64_t* ptr, const int* row, const double* val, int* order,
                             const b200ldlt_options* opt, b200ldlt_inform* inf) {
   clear_inform(inf);
   if (!akeep || !opt || !inf) return SSIDS_ERROR_CALL_SEQUENCE;
   if (n < 0) return inf->flag = SSIDS_ERROR_A_N_OOR;
   if (n > 0 && (!ptr || !row || !order)) return inf->flag = SSIDS_ERROR_A_PTR;
   if (n > 0 && !val) return inf->flag = SSIDS_ERROR_VAL;
   const int base = opt->array_base ? 1 : 0;
   if (n > 0 && ptr[0] != base) return inf->flag = SSIDS_ERROR_A_PTR;
   for (int i = 0; i < n; ++i)
      if (ptr[i + 1] < ptr[i]) return inf->flag = SSIDS_ERROR_A_PTR;
   std::vector<double> sc(std::max(n, 1));
   int mflag = 0;
   try {
      const int64_t nz = n > 0 ? ptr[n] - base : 0;
      std::vector<int64_t> p0(n + 1);
      std::vector<int> r0((size_t)nz), o0(std::max(n, 1));
      for (int i = 0; i <= n; ++i) p0[i] = ptr[i] - base;
      for (int64_t k = 0; k < nz; ++k) {
         r0[k] = row[k] - base;
         if (r0[k] < 0 || r0[k] >= n) return inf->flag = SSIDS_ERROR_A_ALL_OOR;
      }
      mflag = match_order(n, p0.data(), r0.data(), val, 3, o0.data(), sc.data());
      if (mflag < 0) return inf->flag = mflag;
      for (int i = 0; i < n; ++i) order[i] = o0[i] + base;
   } catch (const std::bad_alloc&) {
      return inf->flag = SSIDS_ERROR_ALLOCATION;
   }
   const int flag = b200ldlt_analyse_ptr64(akeep, n, ptr, row, order, opt, inf);
   if (flag < 0) return flag;
   Akeep* ak = static_cast<Akeep*>(*akeep);
   ak->match_scale.assign(sc.begin(), sc.begin() + n);
   if (mflag == 1) inf->flag = ak->inf_analyse.flag = SSIDS_WARNING_ANAL_SINGULAR;
   return inf->flag;
}

int b200ldlt_factor_scaled(void* akeep, void** fkeep, int posdef, const double* val, double* scale,
                           const b200ldlt_options* opt, b200ldlt_inform* inf) {
   if (!akeep || !fkeep || !opt || !inf) return SSIDS_ERROR_CALL_SEQUENCE;
   Akeep* ak = static_cast<Akeep*>(akeep);
   *inf = ak->inf_analyse;
   const int sc = opt->reserved[2];
   if (sc == 3 && (int)ak->match_scale.size() != ak->sym.n) return inf->flag = SSIDS_ERROR_NO_SAVED_SCALING;   // ssids.f90:990-995
   if (ak->sym.n > 0 && !val) return inf->flag = SSIDS_ERROR_VAL;
   if (int rc = ensure_fkeep(ak, fkeep, opt, inf)) return rc;
   Fkeep* fk = static_cast<Fkeep*>(*fkeep);
   const int n = ak->sym.n;
   int rc = 0;
   if (sc == 1) {                       // Hungarian algorithm (ssids.f90:928-960), returned in `scale` if given
      std::vector<double> s(std::max(n, 1));
      const int hf = hungarian_scale_sym(n, ak->sym.ptr.data(), ak->sym.row.data(), val, s.data(), opt->action != 0);
      if (hf == -2) return inf->flag = SSIDS_ERROR_SINGULAR;      // structurally singular and action = .false.
      if (scale) std::copy(s.begin(), s.begin() + n, scale);
      rc = numeric_set_scaling(fk->num, s.data());
   } else if (sc == 2) {                // auction algorithm (ssids.f90:962-990), returned in `scale` if given
      std::vector<double> s(std::max(n, 1));
      auction_scale_sym(n, ak->sym.ptr.data(), ak->sym.row.data(), val, s.data());
      if (scale) std::copy(s.begin(), s.begin() + n, scale);
      rc = numeric_set_scaling(fk->num, s.data());
   } else if (sc == 3) {                // saved by the matching-based order of the analyse phase (ssids.f90:990-998)
      if (scale) std::copy(ak->match_scale.begin(), ak->match_scale.end(), scale);
      rc = numeric_set_scaling(fk->num, n > 0 ? ak->match_scale.data() : nullptr);
   } else if (sc >= 4) {                // norm equilibration, computed here, returned in `scale` if given
      std::vector<double> s(std::max(n, 1));
      equilib_scale_sym(n, ak->sym.ptr.data(), ak->sym.row.data(), val, s.data());
      if (scale) std::copy(s.begin(), s.begin() + n, scale);
      rc = numeric_set_scaling(fk->num, s.data());
   } else {
      rc = numeric_set_scaling(fk->num, scale);   // user-supplied (original variable order) or none
   }
   if (rc) return inf->flag = rc;
   return factor_common(akeep, fkeep, posdef, val, false, opt, inf);
}
int b200ldlt_equilib_scale(int n, const int64_t* ptr, const int* row, const double* val, double* scaling, int* iterations) {
   if (n < 0 || !ptr || (n > 0 && (!row || !val || !scaling))) return SSIDS_ERROR_CALL_SEQUENCE;
   const int it = equilib_scale_sym(n, ptr, row, val, scaling);
   if (iterations) *iterations = it;
   return 0;
}
int b200ldlt_set_scaling(void* akeep, void** fkeep, const double* scale, const b200ldlt_options* opt,
                         b200ldlt_inform* inf) {
   if (!akeep || !fkeep || !opt || !inf) return SSIDS_ERROR_CALL_SEQUENCE;
   Akeep* ak = static_cast<Akeep*>(akeep);
   *inf = ak->inf_analyse;
   if (!scale && !*fkeep) return inf->flag = 0;          // nothing to clear
   if (int rc = ensure_fkeep(ak, fkeep, opt, inf)) return rc;
   return inf->flag = numeric_set_scaling(static_cast<Fkeep*>(*fkeep)->num, scale);
}
int b200ldlt_set_user_matrix(void* akeep, void** fkeep, int ne, const int* row, const int* col,
                             const b200ldlt_options* opt, b200ldlt_inform* inf) {
   if (!akeep || !fkeep || !opt || !inf || ne < 0 || (ne > 0 && (!row || !col))) return SSIDS_ERROR_CALL_SEQUENCE;
   Akeep* ak = static_cast<Akeep*>(akeep);
   *inf = ak->inf_analyse;
   if (int rc = ensure_fkeep(ak, fkeep, opt, inf)) return rc;
   return inf->flag = numeric_set_user_matrix(static_cast<Fkeep*>(*fkeep)->num, ne, row, col);
}
int b200ldlt_solve_refined(void* akeep, void* fkeep, double* x, int max_refinements, double stop_absolute,
                           double stop_relative, int* refinements, const b200ldlt_options*, b200ldlt_inform* inf) {
   if (!akeep || !fkeep || !inf || !x) return SSIDS_ERROR_CALL_SEQUENCE;
   Fkeep* fk = static_cast<Fkeep*>(fkeep);
   if (fk->ak != akeep || !fk->num || !numeric_has_factors(fk->num)) return inf->flag = SSIDS_ERROR_CALL_SEQUENCE;
   if (fk->ak->sym.n == 0) { if (refinements) *refinements = 0; return inf->flag = 0; }
   double secs = 0;
   int flag = numeric_solve_refined(fk->num, x, std::max(max_refinements, 0), stop_absolute, stop_relative, refinements, &secs);
   inf->time_solve = secs;
   return inf->flag = flag;
}
int b200ldlt_factor_mapped(void* akeep, void** fkeep, int posdef, int ne, const double* user_val,
                           const b200ldlt_options* opt, b200ldlt_inform* inf) {
   if (ne < 0) return SSIDS_ERROR_CALL_SEQUENCE;
   return factor_common(akeep, fkeep, posdef, user_val, false, opt, inf, ne);
}

// Refactorisation after a few of the caller's entries changed (the interior-point loop: SBLS rebuilds K with new_h = new_a = 0
// and only the barrier terms new, sbls.f90:143-153): the values of the latest b200ldlt_factor_mapped stay resident on the
// device, `which[i]` (positions in the caller's entry list, options%array_base) take `new_val[i]`.
int b200ldlt_factor_changed(void* akeep, void** fkeep, int posdef, int nchg, const int* which, const double* new_val,
                            const b200ldlt_options* opt, b200ldlt_inform* inf) {
   if (nchg < 0 || (nchg > 0 && (!which || !new_val)) || !opt) return SSIDS_ERROR_CALL_SEQUENCE;
   const int base = opt->array_base ? 1 : 0;
   if (base == 0) return factor_common(akeep, fkeep, posdef, new_val, false, opt, inf, -1, nchg, which);
   std::vector<int> w0(which, which + nchg);
   for (int& v : w0) v -= base;
   return factor_common(akeep, fkeep, posdef, new_val, false, opt, inf, -1, nchg, w0.data());
}

static int solve_common(void* akeep, void* fkeep, int job, int nrhs, double* x, int ldx, bool on_device,
                        b200ldlt_inform* inf) {
   if (!akeep || !fkeep || !inf) return SSIDS_ERROR_CALL_SEQUENCE;
   Fkeep* fk = static_cast<Fkeep*>(fkeep);
   if (fk->ak != akeep || !fk->num || !numeric_has_factors(fk->num)) return inf->flag = SSIDS_ERROR_CALL_SEQUENCE;
   if (job < 0 || job > 4) return inf->flag = SSIDS_ERROR_JOB_OOR;
   if (nrhs < 1 || ldx < fk->ak->sym.n) return inf->flag = SSIDS_ERROR_X_SIZE;
   if (fk->ak->sym.n == 0) return inf->flag = 0;
   double secs = 0;
   int64_t nl = 0;
   int flag = numeric_solve(fk->num, job, nrhs, x, ldx, on_device, &secs, &nl);
   inf->time_solve = secs;
   inf->reserved[1] = (int)nl;
   return inf->flag = flag;
}

int b200ldlt_solve(void* akeep, void* fkeep, int job, int nrhs, double* x, int ldx,
                   const b200ldlt_options*, b200ldlt_inform* inf) {
   return solve_common(akeep, fkeep, job, nrhs, x, ldx, false, inf);
}
int b200ldlt_solve_device(void* akeep, void* fkeep, int job, int nrhs, double* d_x, int ldx,
                          const b200ldlt_options*, b200ldlt_inform* inf) {
   return solve_common(akeep, fkeep, job, nrhs, d_x, ldx, true, inf);
}

int b200ldlt_enquire_indef(void* akeep, void* fkeep, int* piv_order, double* d,
                           const b200ldlt_options* opt, b200ldlt_inform* inf) {
   if (!akeep || !fkeep || !inf) return SSIDS_ERROR_CALL_SEQUENCE;
   Fkeep* fk = static_cast<Fkeep*>(fkeep);
   Akeep* ak = static_cast<Akeep*>(akeep);
   if (fk->ak != ak || !numeric_has_factors(fk->num)) return inf->flag = SSIDS_ERROR_CALL_SEQUENCE;
   if (numeric_is_posdef(fk->num)) return inf->flag = SSIDS_ERROR_NOT_LDLT;
   const int n = ak->sym.n;
   const int base = (opt && opt->array_base) ? 1 : 0;
   std::vector<int> po(n, 0);
   if (d) std::memset(d, 0, sizeof(double) * 2 * n);
   int flag = numeric_enquire(fk->num, piv_order ? po.data() : nullptr, d);
   if (flag < 0) return inf->flag = flag;
   if (piv_order) {
      // po is indexed by analysis pivot position and holds +-(position+1); map back to variables
      // (fkeep.F90:405-410) and shift to the caller's base
      for (int i = 0; i < n; ++i) {
         int p = ak->sym.order[i];
         if (p < 0) { piv_order[i] = 0; continue; }
         int v = po[p];
         if (v == 0) { piv_order[i] = 0; continue; }   // delayed beyond the root / never eliminated
         int pos = (v < 0 ? -v : v) - 1 + base;
         if (base == 0 && pos == 0 && v < 0) pos = 0;   // -0 is not representable; see INTEGRATION.md
         piv_order[i] = v < 0 ? -pos : pos;
      }
   }
   return inf->flag = 0;
}

int b200ldlt_enquire_posdef(void* akeep, void* fkeep, double* d, b200ldlt_inform* inf) {
   if (!akeep || !fkeep || !inf || !d) return SSIDS_ERROR_CALL_SEQUENCE;
   Fkeep* fk = static_cast<Fkeep*>(fkeep);
   Akeep* ak = static_cast<Akeep*>(akeep);
   if (fk->ak != ak || !numeric_has_factors(fk->num)) return inf->flag = SSIDS_ERROR_CALL_SEQUENCE;
   if (!numeric_is_posdef(fk->num)) return inf->flag = SSIDS_ERROR_NOT_LLT;
   const int n = ak->sym.n;
   std::vector<double> dd(2 * (size_t)n, 0.0);
   int flag = numeric_enquire(fk->num, nullptr, dd.data());
   if (flag < 0) return inf->flag = flag;
   // we hold L D L^T with D^{-1} stored; the Cholesky diagonal is sqrt(D)
   for (int i = 0; i < n; ++i) d[i] = dd[2 * i] > 0 ? 1.0 / std::sqrt(dd[2 * i]) : 0.0;
   return inf->flag = 0;
}

int b200ldlt_alter(void* akeep, void* fkeep, const double* d, b200ldlt_inform* inf) {
   if (!akeep || !fkeep || !inf || !d) return SSIDS_ERROR_CALL_SEQUENCE;
   Fkeep* fk = static_cast<Fkeep*>(fkeep);
   if (fk->ak != akeep || !numeric_has_factors(fk->num)) return inf->flag = SSIDS_ERROR_CALL_SEQUENCE;
   return inf->flag = numeric_alter(fk->num, d);
}

int b200ldlt_free_fkeep(void** fkeep) {
   if (!fkeep || !*fkeep) return 0;
   Fkeep* fk = static_cast<Fkeep*>(*fkeep);
   if (fk->num) numeric_destroy(fk->num);   // before the Plan it points into
   akeep_release(fk->ak);
   delete fk;
   *fkeep = nullptr;
   return 0;
}

int b200ldlt_free(void** akeep, void** fkeep) {
   if (fkeep) b200ldlt_free_fkeep(fkeep);
   if (akeep && *akeep) {
      akeep_release(static_cast<Akeep*>(*akeep));
      *akeep = nullptr;
   }
   return 0;
}

int b200ldlt_akeep_sizes(void* akeep, int* nnodes, int64_t* nrlist, int64_t* nnlist) {
   if (!akeep) return SSIDS_ERROR_CALL_SEQUENCE;
   const Symbolic& s = static_cast<Akeep*>(akeep)->sym;
   if (nnodes) *nnodes = s.nnodes;
   if (nrlist) *nrlist = (int64_t)s.rlist.size();
   if (nnlist) *nnlist = (int64_t)s.nlist.size() / 2;
   return 0;
}

int b200ldlt_akeep_get(void* akeep, int base, int* sptr, int* sparent, int64_t* rptr, int* rlist,
                       int64_t* nptr, int64_t* nlist, int* invp) {
   if (!akeep) return SSIDS_ERROR_CALL_SEQUENCE;
   const Symbolic& s = static_cast<Akeep*>(akeep)->sym;
   base = base ? 1 : 0;
   if (sptr) for (int i = 0; i <= s.nnodes; ++i) sptr[i] = s.sptr[i] + base;
   if (sparent) for (int i = 0; i < s.nnodes; ++i) sparent[i] = s.sparent[i] + base;
   if (rptr) for (int i = 0; i <= s.nnodes; ++i) rptr[i] = s.rptr[i] + base;
   if (rlist) for (size_t i = 0; i < s.rlist.size(); ++i) rlist[i] = s.rlist[i] + base;
   if (nptr) for (int i = 0; i <= s.nnodes; ++i) nptr[i] = s.nptr[i] + base;
   if (nlist) for (size_t i = 0; i < s.nlist.size(); ++i) nlist[i] = s.nlist[i] + base;
   if (invp) for (int i = 0; i < s.n; ++i) invp[i] = s.invp[i] + base;
   return 0;
}

int b200ldlt_ntree_sizes(void* akeep, int* nnodes, int64_t* nrlist, int64_t* nnlist) {
   if (!akeep) return SSIDS_ERROR_CALL_SEQUENCE;
   const Symbolic& s = static_cast<Akeep*>(akeep)->nsym;
   if (nnodes) *nnodes = s.nnodes;
   if (nrlist) *nrlist = (int64_t)s.rlist.size();
   if (nnlist) *nnlist = (int64_t)s.nlist.size() / 2;
   return 0;
}

int b200ldlt_ntree_get(void* akeep, int base, int* sptr, int* sparent, int64_t* rptr, int* rlist,
                       int64_t* nptr, int64_t* nlist) {
   if (!akeep) return SSIDS_ERROR_CALL_SEQUENCE;
   const Symbolic& s = static_cast<Akeep*>(akeep)->nsym;
   base = base ? 1 : 0;
   if (sptr) for (int i = 0; i <= s.nnodes; ++i) sptr[i] = s.sptr[i] + base;
   if (sparent) for (int i = 0; i < s.nnodes; ++i) sparent[i] = s.sparent[i] + base;
   if (rptr) for (int i = 0; i <= s.nnodes; ++i) rptr[i] = s.rptr[i] + base;
   if (rlist) for (size_t i = 0; i < s.rlist.size(); ++i) rlist[i] = s.rlist[i] + base;
   if (nptr) for (int i = 0; i <= s.nnodes; ++i) nptr[i] = s.nptr[i] + base;
   if (nlist) for (size_t i = 0; i < s.nlist.size(); ++i) nlist[i] = s.nlist[i] + base;
   return 0;
}

int b200ldlt_partition(void* akeep, int nranks) {
   if (!akeep) return SSIDS_ERROR_CALL_SEQUENCE;
   Akeep* ak = static_cast<Akeep*>(akeep);
   find_subtree_partition(ak->nsym, nranks, 1.2, 0);   // on the engine's (fused) tree
   proportional_mapping(ak->nsym, nranks);
   return ak->nsym.nparts;
}

int b200ldlt_map_info(void* akeep, int node, int* owner, int* stage) {
   if (!akeep) return SSIDS_ERROR_CALL_SEQUENCE;
   const Symbolic& s = static_cast<Akeep*>(akeep)->nsym;
   if (node < 0 || node >= s.nnodes || (int)s.map_owner.size() < s.nnodes) return SSIDS_ERROR_CALL_SEQUENCE;
   if (owner) *owner = s.map_owner[node];
   if (stage) *stage = s.map_stage[node];
   return 0;
}

int b200ldlt_part_info(void* akeep, int part, int* first_node, int* last_node, int* owner,
                       int* dest_node, int64_t* flops) {
   if (!akeep) return SSIDS_ERROR_CALL_SEQUENCE;
   const Symbolic& s = static_cast<Akeep*>(akeep)->nsym;
   if (part < 0 || part >= s.nparts) return SSIDS_ERROR_CALL_SEQUENCE;
   if (first_node) *first_node = s.part[part];
   if (last_node) *last_node = s.part[part + 1] - 1;
   if (owner) *owner = s.exec_loc[part];
   if (dest_node) {
      int last = s.part[part + 1] - 1;
      *dest_node = (last >= 0 && last < s.nnodes && s.sparent[last] < s.nnodes) ? s.sparent[last] : -1;
   }
   if (flops) {
      int64_t fl = 0;
      for (int node = s.part[part]; node < s.part[part + 1]; ++node) {
         int64_t m = s.rptr[node + 1] - s.rptr[node], nc = s.sptr[node + 1] - s.sptr[node];
         for (int64_t j = m - nc + 1; j <= m; ++j) fl += j * j;
      }
      *flops = fl;
   }
   return 0;
}

// ---- multi-GPU -------------------------------------------------------------------------------
int b200ldlt_dist_init(void* akeep, void** fkeep, int rank, int nranks, const b200ldlt_options* opt,
                       b200ldlt_inform* inf) {
   if (!akeep || !fkeep || !opt || !inf) return SSIDS_ERROR_CALL_SEQUENCE;
   Akeep* ak = static_cast<Akeep*>(akeep);
   *inf = ak->inf_analyse;
   find_subtree_partition(ak->nsym, nranks, 1.2, 0);   // reported through part_info (SSIDS's single-level view)
   proportional_mapping(ak->nsym, nranks);             // what the engine executes
   inf->nparts = ak->nsym.nparts;
   if (ensure_plan(ak, opt->device)) { inf->cuda_error = ak->plan_cuda_error; return inf->flag = SSIDS_ERROR_CUDA_UNKNOWN; }
   if (*fkeep) b200ldlt_free_fkeep(fkeep);
   Fkeep* fk = new (std::nothrow) Fkeep();
   if (!fk) return inf->flag = SSIDS_ERROR_ALLOCATION;
   fk->ak = ak;
   ak->refs.fetch_add(1);
   fk->num = numeric_create(ak->plan);
   *fkeep = fk;
   return inf->flag = numeric_set_partition(fk->num, rank, nranks);
}

int b200ldlt_dist_factor(void* akeep, void* fkeep, int phase, int posdef, const double* val,
                         const b200ldlt_options* opt, b200ldlt_inform* inf) {
   if (!akeep || !fkeep || !opt || !inf) return SSIDS_ERROR_CALL_SEQUENCE;
   Akeep* ak = static_cast<Akeep*>(akeep);
   Fkeep* fk = static_cast<Fkeep*>(fkeep);
   *inf = ak->inf_analyse;
   FactorStats st;
   int flag = numeric_factor_phase(fk->num, val, to_engine(opt, posdef), st, phase);
   inf->cuda_error = st.cuda_error;
   inf->num_delay = st.num_delay;
   inf->num_neg = st.num_neg;
   inf->num_two = st.num_two;
   inf->not_first_pass = st.not_first_pass;
   inf->not_second_pass = st.not_second_pass;
   inf->maxfront = st.maxfront;
   inf->matrix_rank = ak->sym.matrix_rank - st.num_zero;
   inf->time_factor = st.t_total;
   inf->time_factor_device = st.t_device;
   inf->factor_bytes = st.factor_bytes;
   inf->reserved[0] = (int)st.launches;
   return inf->flag = flag;
}

// stages of the proportional mapping, and its boundary fronts: nodes whose parent is factorised in a later
// stage; (node, owner of the node, owner of its parent, stage of the node)
int b200ldlt_dist_nstages(void* akeep) {
   if (!akeep) return SSIDS_ERROR_CALL_SEQUENCE;
   return static_cast<Akeep*>(akeep)->nsym.nstages;
}
int b200ldlt_dist_nboundary(void* akeep) {
   if (!akeep) return SSIDS_ERROR_CALL_SEQUENCE;
   const Symbolic& s = static_cast<Akeep*>(akeep)->nsym;
   if ((int)s.map_stage.size() < s.nnodes) return 0;
   int c = 0;
   for (int node = 0; node < s.nnodes; ++node) {
      const int par = s.sparent[node];
      if (par < s.nnodes && s.map_stage[par] > s.map_stage[node]) ++c;
   }
   return c;
}
int b200ldlt_dist_boundary(void* akeep, int i, int* node_out, int* src, int* dst, int* stage) {
   if (!akeep) return SSIDS_ERROR_CALL_SEQUENCE;
   const Symbolic& s = static_cast<Akeep*>(akeep)->nsym;
   if ((int)s.map_stage.size() < s.nnodes) return SSIDS_ERROR_CALL_SEQUENCE;
   int c = 0;
   for (int node = 0; node < s.nnodes; ++node) {
      const int par = s.sparent[node];
      if (par < s.nnodes && s.map_stage[par] > s.map_stage[node]) {
         if (c == i) {
            if (node_out) *node_out = node;
            if (src) *src = s.map_owner[node];
            if (dst) *dst = s.map_owner[par];
            if (stage) *stage = s.map_stage[node];
            return 0;
         }
         ++c;
      }
   }
   return SSIDS_ERROR_CALL_SEQUENCE;
}
int64_t b200ldlt_dist_contrib_bytes(void* fkeep, int node) { return numeric_contrib_bytes(static_cast<Fkeep*>(fkeep)->num, node); }
int b200ldlt_dist_contrib_pack(void* fkeep, int node, void* d_dst) { return numeric_contrib_pack(static_cast<Fkeep*>(fkeep)->num, node, d_dst); }
int b200ldlt_dist_contrib_unpack(void* fkeep, int node, const void* d_src, int64_t bytes) {
   return numeric_contrib_unpack(static_cast<Fkeep*>(fkeep)->num, node, d_src, bytes);
}
int64_t b200ldlt_dist_xtail_bytes(void* fkeep, int node) { return numeric_xtail_bytes(static_cast<Fkeep*>(fkeep)->num, node); }
int b200ldlt_dist_xtail_pack(void* fkeep, int node, void* d_dst) { return numeric_xtail_pack(static_cast<Fkeep*>(fkeep)->num, node, d_dst); }
int b200ldlt_dist_xtail_unpack(void* fkeep, int node, const void* d_src) { return numeric_xtail_unpack(static_cast<Fkeep*>(fkeep)->num, node, d_src); }
int b200ldlt_dist_solve(void* fkeep, int step, double* d_x, const double* d_src) {
   return numeric_dist_solve(static_cast<Fkeep*>(fkeep)->num, step, d_x, d_src);
}
void* b200ldlt_dist_xp(void* fkeep) { return numeric_xp(static_cast<Fkeep*>(fkeep)->num); }
int b200ldlt_dist_xp_get(void* fkeep, void* d_dst, int n) { return numeric_xp_get(static_cast<Fkeep*>(fkeep)->num, d_dst, n); }
int b200ldlt_dist_xp_get_stage(void* fkeep, int stage, void* d_dst, int n) { return numeric_xp_get_stage(static_cast<Fkeep*>(fkeep)->num, stage, d_dst, n); }
int b200ldlt_dist_xp_merge_stage(void* fkeep, int stage, const void* d_src, int n) { return numeric_xp_merge_stage(static_cast<Fkeep*>(fkeep)->num, stage, d_src, n); }

// per-node results (parity tests compare these with the oracle's)
int b200ldlt_node_info(void* fkeep, int node, int* m, int* n, int* nelim) {
   if (!fkeep) return SSIDS_ERROR_CALL_SEQUENCE;
   return numeric_node_info(static_cast<Fkeep*>(fkeep)->num, node, m, n, nelim);
}
int b200ldlt_node_copy(void* fkeep, int node, double* L, double* d, int* perm) {
   if (!fkeep) return SSIDS_ERROR_CALL_SEQUENCE;
   return numeric_node_copy(static_cast<Fkeep*>(fkeep)->num, node, L, d, perm);
}

}  // extern "C"
