// TEST INFRASTRUCTURE ONLY (oracle).  Thin driver around the reference's C++ SSIDS CPU engine,
// compiled by oracle/build_ref.sh together with the unmodified reference sources into
// oracle/_ref/libssids_ref.so.  It does what the (uncompilable here) Fortran layers do around the
// C ABI of src/ssids/cpu/subtree.f90:42-184:
//   * supplies the two Fortran-side symbols the engine links against
//     (src/ssids/contrib.h:16-21; only reached when a subtree has child contributions),
//   * calls the numeric constructor from inside `omp parallel` + `omp single`, as
//     src/ssids/fkeep.F90:125-130 arranges, so that the engine's task DAG uses all threads,
//   * sequences solve_fwd / solve_diag / solve_diag_bwd / solve_bwd like fkeep.F90:264-292.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load it.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <omp.h>

#include "ssids/cpu/cpu_iface.hxx"
#include "ssids/cpu/ThreadStats.hxx"

using namespace spral::ssids::cpu;

extern "C" {
// reference entry points (src/ssids/cpu/SymbolicSubtree.cxx:10, NumericSubtree.cxx:29-262)
void* spral_ssids_cpu_create_symbolic_subtree(int n, int sa, int en, int const* sptr,
      int const* sparent, long const* rptr, int const* rlist, long const* nptr,
      long const* nlist, int ncontrib, int const* contrib_idx,
      struct cpu_factor_options const* options);
void spral_ssids_cpu_destroy_symbolic_subtree(void*);
void* spral_ssids_cpu_create_num_subtree_dbl(bool posdef, void const* symb, const double* aval,
      const double* scaling, void** child_contrib, struct cpu_factor_options const* options,
      ThreadStats* stats);
void spral_ssids_cpu_destroy_num_subtree_dbl(bool posdef, void*);
int spral_ssids_cpu_subtree_solve_fwd_dbl(bool, void const*, int, double*, int);
int spral_ssids_cpu_subtree_solve_diag_dbl(bool, void const*, int, double*, int);
int spral_ssids_cpu_subtree_solve_diag_bwd_dbl(bool, void const*, int, double*, int);
int spral_ssids_cpu_subtree_solve_bwd_dbl(bool, void const*, int, double*, int);
void spral_ssids_cpu_subtree_enquire_dbl(bool, void const*, int*, double*);
void spral_ssids_cpu_subtree_alter_dbl(bool, void*, double const*);

// Fortran-side helpers (spral_ssids_contrib / _contrib_free modules); single-part runs never
// reach them, so they abort loudly rather than pretend.
void spral_ssids_contrib_get_data(const void*, int*, const double**, int*, const int**, int*,
      const int**, const double**, int*) {
   fprintf(stderr, "ref_driver: spral_ssids_contrib_get_data reached (multi-part not driven)\n");
   abort();
}
void spral_ssids_contrib_free_dbl(void*) {
   fprintf(stderr, "ref_driver: spral_ssids_contrib_free_dbl reached\n");
   abort();
}

struct refssids_opts {   // plain-C mirror of cpu_factor_options (cpu_iface.hxx)
   int print_level;
   int action;
   double small_;
   double u;
   double multiplier;
   long small_subtree_threshold;
   int cpu_block_size;
   int pivot_method;         // 1 app_aggressive, 2 app_block, 3 tpp
   int failed_pivot_method;  // 1 tpp, 2 pass
};

static cpu_factor_options to_ref(const refssids_opts* o) {
   cpu_factor_options r;
   r.print_level = o->print_level;
   r.action = o->action != 0;
   r.small = o->small_;
   r.u = o->u;
   r.multiplier = o->multiplier;
   r.small_subtree_threshold = o->small_subtree_threshold;
   r.cpu_block_size = o->cpu_block_size;
   r.pivot_method = static_cast<PivotMethod>(o->pivot_method);
   r.failed_pivot_method = static_cast<FailedPivotMethod>(o->failed_pivot_method);
   return r;
}

// All index arrays are Fortran (1-based) exactly as akeep holds them, and must outlive the handle.
void* refssids_symbolic(int n, int nnodes, const int* sptr, const int* sparent, const long* rptr,
      const int* rlist, const long* nptr, const long* nlist, const refssids_opts* o) {
   cpu_factor_options ro = to_ref(o);
   return spral_ssids_cpu_create_symbolic_subtree(n, 1, nnodes + 1, sptr, sparent, rptr, rlist,
         nptr, nlist, 0, nullptr, &ro);
}

// stats_out: flag, num_delay, num_neg, num_two, num_zero, maxfront, not_first_pass, not_second_pass
void* refssids_factor(void* symb, int posdef, const double* aval, const double* scaling,
      const refssids_opts* o, int nthreads, int* stats_out, double* seconds) {
   cpu_factor_options ro = to_ref(o);
   ThreadStats st;
   void* num = nullptr;
   if (nthreads < 1) nthreads = omp_get_max_threads();
   omp_set_max_active_levels(4);
   double t0 = omp_get_wtime();
#pragma omp parallel num_threads(nthreads) default(shared)
   {
#pragma omp single
      {
         num = spral_ssids_cpu_create_num_subtree_dbl(posdef != 0, symb, aval, scaling, nullptr,
               &ro, &st);
      }
   }
   double t1 = omp_get_wtime();
   if (seconds) *seconds = t1 - t0;
   stats_out[0] = (int)st.flag;
   stats_out[1] = st.num_delay;
   stats_out[2] = st.num_neg;
   stats_out[3] = st.num_two;
   stats_out[4] = st.num_zero;
   stats_out[5] = st.maxfront;
   stats_out[6] = st.not_first_pass;
   stats_out[7] = st.not_second_pass;
   return num;
}

// job: 0 = full solve, 1 = fwd, 2 = diag, 3 = bwd, 4 = diag+bwd (src/ssids/datatypes.f90:62-66).
// x is in *pivot order* (the caller applies invp like fkeep.F90:248-261).
int refssids_solve(void* num, int posdef, int job, int nrhs, double* x, int ldx) {
   bool pd = posdef != 0;
   int f = 0;
   if (job == 0 || job == 1) f = spral_ssids_cpu_subtree_solve_fwd_dbl(pd, num, nrhs, x, ldx);
   if (f < 0) return f;
   if (job == 0 || job == 4) {
      if (pd) f = spral_ssids_cpu_subtree_solve_bwd_dbl(pd, num, nrhs, x, ldx);
      else f = spral_ssids_cpu_subtree_solve_diag_bwd_dbl(pd, num, nrhs, x, ldx);
   } else if (job == 2) {
      f = spral_ssids_cpu_subtree_solve_diag_dbl(pd, num, nrhs, x, ldx);
   } else if (job == 3) {
      f = spral_ssids_cpu_subtree_solve_bwd_dbl(pd, num, nrhs, x, ldx);
   }
   return f;
}

void refssids_enquire(void* num, int posdef, int* piv_order, double* d) {
   spral_ssids_cpu_subtree_enquire_dbl(posdef != 0, num, piv_order, d);
}
void refssids_alter(void* num, const double* d) {
   spral_ssids_cpu_subtree_alter_dbl(false, num, d);
}
void refssids_free_numeric(void* num, int posdef) {
   spral_ssids_cpu_destroy_num_subtree_dbl(posdef != 0, num);
}
void refssids_free_symbolic(void* symb) { spral_ssids_cpu_destroy_symbolic_subtree(symb); }
int refssids_max_threads(void) { return omp_get_max_threads(); }
}  // extern "C"
