/* SLS C interface for the 'b200ldlt' solver -- host side of the drop-in boundary.
 *
 * Mirrors GALAHAD's own C interface to SLS (include/sls.h:915-1246, implemented by
 * src/sls/C/sls_ciface.f90:721-1108 on top of src/sls/sls.f90): same entry-point names, argument
 * order and meaning, status codes and call sequence
 *     sls_initialize -> sls_analyse_matrix -> sls_factorize_matrix -> sls_solve_system /
 *     sls_partial_solve_system -> sls_information -> sls_terminate
 * so that the reference's C test (src/sls/C/slst.c) reads the same against this library.  Inside
 * GALAHAD the same behaviour is obtained by the CASE('b200ldlt') blocks of INTEGRATION.md; this
 * C++ restatement of the SLS host logic (sls.f90:817, 2178, 3521, 4692, 6551, 8409) exists because
 * no Fortran compiler is available in the build image.
 *
 * control / inform carry the members of sls_control_type / sls_inform_type (include/sls.h:315-905)
 * that act on this solver; members that only concern other back ends (MA57, MA77 out-of-core files,
 * PARDISO, ...) are not reproduced, so the struct LAYOUT is this header's, not sls.h's.
 */
#ifndef B200_SLS_H
#define B200_SLS_H

#include <stdbool.h>
#include <stdint.h>

#include "b200ldlt.h"

#ifdef __cplusplus
extern "C" {
#endif

/* GALAHAD status values used by SLS (src/symbols/symbols.f90:50-137) */
#define GALAHAD_OK 0
#define GALAHAD_ERROR_ALLOCATE (-1)
#define GALAHAD_ERROR_DEALLOCATE (-2)
#define GALAHAD_ERROR_RESTRICTIONS (-3)
#define GALAHAD_ERROR_INERTIA (-20)
#define GALAHAD_ERROR_UNKNOWN_SOLVER (-26)
#define GALAHAD_UNAVAILABLE_OPTION (-29)
#define GALAHAD_ERROR_PERMUTATION (-39)
#define GALAHAD_ERROR_TECHNICAL (-50)

struct sls_control_type {
   bool f_indexing;                 /* Fortran (1-based) or C (0-based) sparse-matrix indexing */
   int error, warning, out, statistics;
   int print_level, print_level_solver;
   int node_amalgamation;           /* -> nemin (32) */
   int pivot_control;               /* 1 pivoted LDL^T (default), 2 Cholesky, 3 Cholesky + singular is an error, 4 no threshold */
   int ordering;                    /* <0: solver's own ordering; >=0: SLS-side (identity here unless a PERM is passed) */
   int scaling;                     /* only 0 is implemented for this solver */
   double relative_pivot_tolerance; /* -> u (0.01) */
   double absolute_pivot_tolerance; /* -> small (DBL_EPSILON) */
   double zero_tolerance;
   int max_iterative_refinements;
   double acceptable_residual_relative, acceptable_residual_absolute;
   bool multiple_rhs;
   int device;                      /* CUDA device ordinal, -1 = current (engine knob, not in sls.h) */
   char prefix[31];
};

struct sls_time_type {
   double total, analyse, factorize, solve;
   double clock_total, clock_analyse, clock_factorize, clock_solve;
   double analyse_external, factorize_external, solve_external;
};

struct sls_inform_type {
   int status, alloc_status;
   char bad_alloc[81];
   int more_info;
   int entries, out_of_range, duplicates, upper, missing_diagonals;
   int max_depth_assembly_tree, nodes_assembly_tree;
   int64_t real_size_factors, integer_size_factors, entries_in_factors;
   int max_front_size;
   int two_by_two_pivots, rank, negative_eigenvalues, delayed_pivots, num_zero;
   int iterative_refinements;
   int64_t flops_assembly, flops_elimination;
   double backward_error_1, backward_error_2;
   struct sls_time_type time;
   struct b200ldlt_inform b200ldlt_inform;   /* the solver's own inform (sls.h keeps ssids_inform here) */
};

void sls_initialize(const char solver[], void **data, struct sls_control_type *control, int *status);
void sls_analyse_matrix(struct sls_control_type *control, void **data, int *status, int n, const char type[],
                        int ne, const int row[], const int col[], const int ptr[]);
/* as sls_analyse_matrix with the optional PERM argument of the Fortran SLS_analyse (sls.f90:2178):
 * perm[i] = position of variable i in the pivot sequence (f_indexing base) */
void sls_analyse_matrix_perm(struct sls_control_type *control, void **data, int *status, int n,
                             const char type[], int ne, const int row[], const int col[], const int ptr[],
                             const int perm[]);
void sls_reset_control(struct sls_control_type *control, void **data, int *status);
void sls_factorize_matrix(void **data, int *status, int ne, const double val[]);
void sls_solve_system(void **data, int *status, int n, double sol[]);
void sls_partial_solve_system(const char part[], void **data, int *status, int n, double sol[]);
void sls_information(void **data, struct sls_inform_type *inform, int *status);
void sls_terminate(void **data, struct sls_control_type *control, struct sls_inform_type *inform);

/* The sorted-CSR mapping SLS builds for its back ends (SLS_coord_to_sorted_csr, sls.f90:8409-8578),
 * exposed for parity tests.  1-based, exactly as the Fortran: map[ne], ptr[n+1]. */
void sls_coord_to_sorted_csr(int n, int ne, const int row[], const int col[], int map[], int ptr[],
                             int *dup, int *oor, int *upper, int *missing_diagonals);

#ifdef __cplusplus
}
#endif
#endif
