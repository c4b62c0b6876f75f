/* b200ldlt -- C ABI of the B200-native sparse symmetric-indefinite LDL^T engine.
 *
 * This is the drop-in boundary for GALAHAD's SLS package (seam B1): a new
 * solver name 'b200ldlt' in src/sls/sls.f90 binds these entry points through ISO_C_BINDING, one per
 * call the 'ssids' CASE blocks make today:
 *
 *   b200ldlt_analyse         <- SSIDS_analyse      (sls.f90:3115-3150  -> ssids.f90:93/148)
 *   b200ldlt_factor          <- SSIDS_factor       (sls.f90:4273-4297  -> ssids.f90:736/770)
 *   b200ldlt_solve           <- SSIDS_solve        (sls.f90:5392-5397, 5693 -> ssids.f90:1114/1139)
 *   b200ldlt_enquire_indef   <- SSIDS_enquire_indef (sls.f90:6312      -> ssids.f90:1299)
 *   b200ldlt_enquire_posdef  <- SSIDS_enquire_posdef                   (ssids.f90:1255)
 *   b200ldlt_alter           <- SSIDS_alter        (sls.f90:6488       -> ssids.f90:1347)
 *   b200ldlt_free            <- SSIDS_free         (sls.f90:5972       -> ssids.f90:1388-1419)
 *
 * Plain pointers and sizes only.  Handles are opaque, owned by the library.  The caller owns every
 * array it passes and may release it on return (the pattern is copied at analyse).  Nothing throws
 * or aborts across this boundary: every call returns an SSIDS-style flag (src/ssids/datatypes.f90:
 * 25-59; 0 ok, <0 error, >0 warning) and records it in inform->flag.  There is no CPU fallback:
 * without a usable CUDA device factor/solve return B200LDLT_ERROR_CUDA_UNKNOWN (-51).
 */
#ifndef B200LDLT_H
#define B200LDLT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* flags: the SSIDS values (src/ssids/datatypes.f90:25-59) */
#define B200LDLT_SUCCESS 0
#define B200LDLT_ERROR_CALL_SEQUENCE (-1)
#define B200LDLT_ERROR_A_N_OOR (-2)
#define B200LDLT_ERROR_A_PTR (-3)
#define B200LDLT_ERROR_SINGULAR (-5)
#define B200LDLT_ERROR_NOT_POS_DEF (-6)
#define B200LDLT_ERROR_ORDER (-8)
#define B200LDLT_ERROR_X_SIZE (-10)
#define B200LDLT_ERROR_JOB_OOR (-11)
#define B200LDLT_ERROR_NOT_LLT (-13)
#define B200LDLT_ERROR_NOT_LDLT (-14)
#define B200LDLT_ERROR_ALLOCATION (-50)
#define B200LDLT_ERROR_CUDA_UNKNOWN (-51)
#define B200LDLT_ERROR_UNIMPLEMENTED (-98)
#define B200LDLT_WARNING_ANAL_SINGULAR 6
#define B200LDLT_WARNING_FACT_SINGULAR 7

/* solve jobs (src/ssids/datatypes.f90:62-66) */
#define B200LDLT_JOB_ALL 0
#define B200LDLT_JOB_FWD 1
#define B200LDLT_JOB_DIAG 2
#define B200LDLT_JOB_BWD 3
#define B200LDLT_JOB_DIAG_BWD 4

/* Options: the subset of ssids_options (src/ssids/datatypes.f90:187-283) that SLS sets through
 * SLS_copy_control_to_ssids (sls.f90:1385-1439), plus engine knobs. */
typedef struct b200ldlt_options {
   int array_base;      /* 0: C indexing of ptr/row/order; 1: Fortran indexing (the sls.f90 shim) */
   int print_level;     /* <0 silent */
   int nemin;           /* node amalgamation (SLS node_amalgamation), <1 -> 32 */
   int action;          /* 1: continue on singular matrix with a warning (7); 0: error -5 */
   double u;            /* relative pivot threshold (SLS relative_pivot_tolerance), clipped to [0,0.5] */
   double small_;       /* |pivot| below this is zero (SLS absolute_pivot_tolerance) */
   double multiplier;   /* factor-storage slack for delayed pivots (>= 1.0) */
   int device;          /* CUDA device ordinal, -1: current device */
   int panel_width;     /* reserved; 0 = engine default */
   int reserved[7];     /* reserved[0] = 1: profile mode (CUDA events around every trailing update), 2: around the
                           launches of the large-front update kernel only;
                           reserved[1] = 1: do not fuse chains (engine tree == reference tree);
                           reserved[2]: ssids options%scaling for b200ldlt_factor_scaled */
} b200ldlt_options;

/* Inform: shaped like ssids_inform (src/ssids/inform.f90:17-49). */
typedef struct b200ldlt_inform {
   int flag;
   int matrix_dup;
   int matrix_missing_diag;
   int matrix_outrange;
   int matrix_rank;
   int maxdepth;
   int maxfront;
   int num_delay;
   int64_t num_factor;
   int64_t num_flops;
   int num_neg;
   int num_sup;
   int num_two;
   int stat;
   int cuda_error;
   int nparts;
   int not_first_pass;   /* columns that failed the a-posteriori pass (ThreadStats.hxx) */
   int not_second_pass;  /* columns still failing after the TPP pass */
   double time_analyse;  /* seconds, host wall clock */
   double time_factor;   /* seconds, CUDA events around the numeric phase incl. H2D of val */
   double time_factor_device; /* seconds, CUDA events, val already resident */
   double time_solve;
   int64_t factor_bytes; /* device bytes held by L, D and pivot lists */
   double upd_flops;     /* profile mode: algorithmic flops of the trailing-update launches */
   double upd_seconds;   /* profile mode: summed CUDA-event time of those launches */
   int reserved[8];      /* [0] kernels launched by factor, [1] by solve, [2] update launches profiled */
} b200ldlt_inform;

void b200ldlt_default_options(b200ldlt_options *opt);

/* Symbolic phase (CPU).  ptr[n+1], row[ptr[n]-base]: LOWER triangle by columns == upper triangle by
 * rows as SLS_coord_to_sorted_csr (sls.f90:8409) leaves it: sorted, diagonal first and present.
 * order[n]: in = position of each variable in the pivot sequence; out = the order actually used. */
int b200ldlt_analyse(void **akeep, int n, const int *ptr, const int *row, int *order,
                     const b200ldlt_options *opt, b200ldlt_inform *inf);

/* Fill-reducing orderings computed on the host, for callers that pass no order of their own (what SLS_analyse gets from
 * MC68 / METIS, sls.f90:2395-2470 and ssids.f90 options%ordering = 1; stubs in a stock build of the reference).
 * kind 1: approximate minimum degree; kind 3: nested dissection (level-structure separators, minimum-degree leaves).
 * ptr / row as for b200ldlt_analyse_ptr64 (base from opt->array_base); order[n] out, same base. */
int b200ldlt_order(int n, const int64_t *ptr, const int *row, int kind, int *order, const b200ldlt_options *opt);

/* Matching-based ordering: ssids_analyse with options%ordering = 2 and VAL (ssids.f90:230-238, 321-358;
 * match_order_metis, src/spral/match_order.f90:135-208).  A maximum-product matching of |a_ij| (the Hungarian algorithm
 * of the scaling) is split into singletons and pairs, each pair becomes one vertex of a compressed graph, the compressed
 * graph is ordered by nested dissection (this library's, where the reference calls METIS) and each pair takes two
 * consecutive places of the order handed to the symbolic analysis (whose postorder of the elimination tree, as in
 * ssids_analyse, is what `order` returns); the matching's scaling is kept in akeep for b200ldlt_factor_scaled with
 * scaling 3.  val[] aligned with ptr / row (host).  order[n]: out.  Flag 6 (SSIDS_WARNING_ANAL_SINGULAR) when the
 * matrix is structurally singular, -9 (SSIDS_ERROR_VAL) without val.
 * b200ldlt_match_order: the order alone on 0-based arrays; kind 1 orders the compressed graph by minimum degree, else
 * nested dissection; every output but `order` may be NULL: scale[n]; pairs[n] = the split matching (1-based partner, -1
 * singleton, -2 unmatched); *ncomp, cptr[ncomp+1], crow[] = lower triangle of the compressed pattern (at most n + 1 and
 * 2 ptr[n] entries).  Returns 0, 1 (structurally singular) or an error flag. */
int b200ldlt_analyse_matched(void **akeep, int n, const int64_t *ptr, const int *row, const double *val, int *order,
                             const b200ldlt_options *opt, b200ldlt_inform *inf);
int b200ldlt_match_order(int n, const int64_t *ptr, const int *row, const double *val, int kind, int *order, double *scale,
                         int *pairs, int *ncomp, int64_t *cptr, int *crow);

/* 64-bit ptr variant (ssids.f90:148 analyse_double takes int64 ptr). */
int b200ldlt_analyse_ptr64(void **akeep, int n, const int64_t *ptr, const int *row, int *order,
                           const b200ldlt_options *opt, b200ldlt_inform *inf);

/* Numeric phase (GPU).  val[] aligned with ptr/row, HOST memory.  *fkeep may hold a previous
 * factorization of the same akeep: its storage is reused. */
int b200ldlt_factor(void *akeep, void **fkeep, int posdef, const double *val,
                    const b200ldlt_options *opt, b200ldlt_inform *inf);

/* As b200ldlt_factor, with val already resident in DEVICE memory (refactorization loops). */
int b200ldlt_factor_device(void *akeep, void **fkeep, int posdef, const double *d_val,
                           const b200ldlt_options *opt, b200ldlt_inform *inf);

/* SLS's value gather on the device (replaces the loop of SLS_factorize, sls.f90:4106-4147, that fills
 * data%matrix%VAL from the caller's values through data%MAPS).  set_value_map: once per analyse, `maps[l]` as
 * SLS_coord_to_sorted_csr leaves it (1-based: > 0 position of entry l, < 0 add to position -maps[l], 0 out of
 * range; sls.f90:8409-8578).  factor_mapped: `user_val[ne]` in the caller's order (host); positions without an
 * entry (missing diagonals) are zero, duplicates are summed in input order. */
int b200ldlt_set_value_map(void *akeep, void **fkeep, int ne, const int *maps, const b200ldlt_options *opt,
                           b200ldlt_inform *inf);
/* b200ldlt_factor_changed: refactorisation after a FEW of the caller's entries changed -- the interior-point loop, where
 * SBLS rebuilds K every iteration although only the barrier terms are new (new_h = new_a = 0, sbls.f90:143-153).  The
 * values of the latest b200ldlt_factor_mapped stay resident on the device; entry `which[i]` of the caller's list
 * (options%array_base) takes `new_val[i]`, then gather + factorisation run as in factor_mapped: H2D traffic is 12 bytes
 * per changed entry.  -1 (call sequence) without a previous complete factor_mapped on this fkeep. */
int b200ldlt_factor_changed(void *akeep, void **fkeep, int posdef, int nchg, const int *which, const double *new_val,
                            const b200ldlt_options *opt, b200ldlt_inform *inf);
int b200ldlt_factor_mapped(void *akeep, void **fkeep, int posdef, int ne, const double *user_val,
                           const b200ldlt_options *opt, b200ldlt_inform *inf);

/* Factorisation with scaling: ssids_factor's optional SCALE argument and options%scaling (ssids.f90:900-1030),
 * the option in opt->reserved[2]:  <= 0: `scale` (n factors, original variable order) is used if not NULL, else no
 * scaling;  >= 4: symmetric infinity-norm equilibration (inf_norm_equilib_sym, src/spral/scaling.f90:480-521)
 * computed from `val` and returned in `scale` if not NULL;  1: matching-based scaling by the Hungarian algorithm
 * (hungarian_scale_sym, src/spral/scaling.f90:134-170, MC64's algorithm; ssids.f90:928-960), computed on the host from
 * `val`, returned in `scale` if not NULL; a structurally singular matrix is scaled all the same when options%action is
 * set and is SSIDS_ERROR_SINGULAR (-5) otherwise;  2: matching-based scaling by SPRAL's auction algorithm
 * (auction_scale_sym, scaling.f90:269-309, default auction_options; ssids.f90:962-990);  3: the scaling saved by
 * b200ldlt_analyse_matched (ssids.f90:990-998), SSIDS_ERROR_NO_SAVED_SCALING (-15) after any other analyse.  The scaled matrix S A S is factorised; solves scale on the way in (jobs 0, 1) and on the way out
 * (jobs 0, 3, 4) exactly as fkeep.F90:247-310.  A later b200ldlt_factor / _factor_mapped keeps the scaling that was
 * set last; pass opt->reserved[2] = 0 and scale = NULL here to clear it.  `val` is a host pointer.
 * b200ldlt_equilib_scale: the equilibration alone (0-based lower triangle by columns). */
int b200ldlt_factor_scaled(void *akeep, void **fkeep, int posdef, const double *val, double *scale,
                           const b200ldlt_options *opt, b200ldlt_inform *inf);
int b200ldlt_equilib_scale(int n, const int64_t *ptr, const int *row, const double *val, double *scaling, int *iterations);
/* b200ldlt_hungarian_scale: the Hungarian scaling alone (0-based lower triangle by columns).  match (n, may be NULL):
 * row i is matched to column match[i] (1-based; negative: completion of a structurally singular matrix); *matched: the
 * cardinality of the matching.  Returns 0, 1 (singular, scaled because scale_if_singular) or -2 (singular, scaling = 1). */
int b200ldlt_hungarian_scale(int n, const int64_t *ptr, const int *row, const double *val, double *scaling, int *match,
                             int *matched, int scale_if_singular);
/* b200ldlt_auction_scale: the auction scaling alone; match[i] = column matched to row i (1-based, 0 = unmatched). */
int b200ldlt_auction_scale(int n, const int64_t *ptr, const int *row, const double *val, double *scaling, int *match,
                           int *matched, int *iterations);
/* The scaling alone: `scale` (n factors, original variable order, host) applies to every later factorisation of this
 * fkeep and to its solves, NULL clears it.  What the SLS layer uses when it computes the factors itself
 * (data%SCALE with explicit_scaling, sls.f90:3558) and the values arrive through b200ldlt_factor_mapped. */
int b200ldlt_set_scaling(void *akeep, void **fkeep, const double *scale, const b200ldlt_options *opt,
                         b200ldlt_inform *inf);

/* Iterative refinement on the device (the loop of SLS_solve_ir, sls.f90:4751-4947): x_0 = 0, r_0 = b; repeat
 * { solve A dx = r; x += dx; r = b - A x } until max|r_i| < max(stop_absolute, stop_relative max|b_i|) or
 * max_refinements.  The residual uses the caller's own entries: set_user_matrix (once per analyse; 1-based
 * coordinates exactly as given to SLS, out-of-range entries ignored, duplicates kept) and the values of the
 * latest b200ldlt_factor_mapped.  x: b on entry, the solution on return (host). */
int b200ldlt_set_user_matrix(void *akeep, void **fkeep, int ne, const int *row, const int *col,
                             const b200ldlt_options *opt, b200ldlt_inform *inf);
int b200ldlt_solve_refined(void *akeep, void *fkeep, double *x, int max_refinements, double stop_absolute,
                           double stop_relative, int *refinements, const b200ldlt_options *opt, b200ldlt_inform *inf);

/* Solve with nrhs right-hand sides held by columns in HOST x(ldx,nrhs), overwritten by the solution.
 * job as B200LDLT_JOB_*. */
int b200ldlt_solve(void *akeep, void *fkeep, int job, int nrhs, double *x, int ldx,
                   const b200ldlt_options *opt, b200ldlt_inform *inf);

/* As b200ldlt_solve with x in DEVICE memory. */
int b200ldlt_solve_device(void *akeep, void *fkeep, int job, int nrhs, double *d_x, int ldx,
                          const b200ldlt_options *opt, b200ldlt_inform *inf);

/* piv_order[n] (nullable): |piv_order[i]| = pivot position (array_base) of variable i, negative if
 * part of a 2x2 pivot.  d[2*n] (nullable): D^{-1}, d[2k] diagonal, d[2k+1] off-diagonal of pivot k. */
int b200ldlt_enquire_indef(void *akeep, void *fkeep, int *piv_order, double *d,
                           const b200ldlt_options *opt, b200ldlt_inform *inf);
/* d[n]: diagonal of the Cholesky factor L. */
int b200ldlt_enquire_posdef(void *akeep, void *fkeep, double *d, b200ldlt_inform *inf);
/* Overwrite D^{-1} with d[2*n] (layout as enquire_indef). */
int b200ldlt_alter(void *akeep, void *fkeep, const double *d, b200ldlt_inform *inf);

int b200ldlt_free_fkeep(void **fkeep);
int b200ldlt_free(void **akeep, void **fkeep);

/* ---- symbolic results, for parity checks against SSIDS's akeep (all array_base-shifted) -------- */
int b200ldlt_akeep_sizes(void *akeep, int *nnodes, int64_t *nrlist, int64_t *nnlist);
int b200ldlt_akeep_get(void *akeep, int array_base, int *sptr, int *sparent, int64_t *rptr,
                       int *rlist, int64_t *nptr, int64_t *nlist, int *invp);

/* The engine's own (numeric) tree: the reference tree with chain links fused into single fronts
 * (explicit zeros bounded; analyse.cpp fuse_chains).  Same array conventions as akeep_get.  It is a
 * valid assembly tree for the same matrix and order, so the oracles can be driven with it too. */
int b200ldlt_ntree_sizes(void *akeep, int *nnodes, int64_t *nrlist, int64_t *nnlist);
int b200ldlt_ntree_get(void *akeep, int array_base, int *sptr, int *sparent, int64_t *rptr,
                       int *rlist, int64_t *nptr, int64_t *nlist);

/* ---- multi-GPU (one process per GPU; the host side moves contribution blocks with NCCL) -------- */
/* Partition the assembly tree into leaf subtrees for `nranks` devices + a root part
 * (after find_subtree_partition, anal.f90:284-459).  Returns the number of parts. */
int b200ldlt_partition(void *akeep, int nranks);
/* proportional mapping computed by the same call: owner rank and stage of a node of the engine tree */
int b200ldlt_map_info(void *akeep, int node, int *owner, int *stage);
int b200ldlt_part_info(void *akeep, int part, int *first_node, int *last_node, int *owner,
                       int *dest_node, int64_t *flops);

/* Tree-parallel execution, one process per GPU.  dist_init maps the assembly tree
 * onto the ranks (proportional mapping: a node's ranks are split among its children by subtree flops; a
 * subtree left with one rank is a leaf part; a node belongs to the first rank of its range) and numbers
 * stages so that stage k only needs fronts of earlier stages from other ranks.
 *   factor: for k = 0 .. nstages-1: dist_factor(phase k+1: my fronts of stage k), then every boundary
 *           front of stage k whose parent lives elsewhere is packed by its owner (contrib_bytes /
 *           contrib_pack), moved by the host (NCCL send/recv) and imported by the parent's owner
 *           (contrib_unpack);
 *   solve:  dist_solve(1) (all ranks: b in pivot order); for k ascending: dist_solve(100+k) (forward
 *           sweep of stage k), x tails of its boundary fronts to the parents' owners (xtail_*); for k
 *           descending: dist_solve(200+k) (diagonal + backward sweep), then the solution entries of stage k
 *           are summed over the ranks (xp_get_stage -> all-reduce -> xp_merge_stage); dist_solve(4) returns
 *           the whole solution in the caller's order on every rank. */
int b200ldlt_dist_init(void *akeep, void **fkeep, int rank, int nranks, const b200ldlt_options *opt,
                       b200ldlt_inform *inf);
int b200ldlt_dist_factor(void *akeep, void *fkeep, int phase, int posdef, const double *val,
                         const b200ldlt_options *opt, b200ldlt_inform *inf);
int b200ldlt_dist_nstages(void *akeep);
int b200ldlt_dist_nboundary(void *akeep);
int b200ldlt_dist_boundary(void *akeep, int i, int *node, int *src_rank, int *dst_rank, int *stage);
int64_t b200ldlt_dist_contrib_bytes(void *fkeep, int node);
int b200ldlt_dist_contrib_pack(void *fkeep, int node, void *d_dst);
int b200ldlt_dist_contrib_unpack(void *fkeep, int node, const void *d_src, int64_t bytes);
int64_t b200ldlt_dist_xtail_bytes(void *fkeep, int node);
int b200ldlt_dist_xtail_pack(void *fkeep, int node, void *d_dst);
int b200ldlt_dist_xtail_unpack(void *fkeep, int node, const void *d_src);
int b200ldlt_dist_solve(void *fkeep, int step, double *d_x, const double *d_src);
void *b200ldlt_dist_xp(void *fkeep);
int b200ldlt_dist_xp_get(void *fkeep, void *d_dst, int n);   /* copy the pivot-order vector (device to device) */
/* entries of the pivot-order vector this rank produced in `stage` (zero elsewhere) / take the entries of `stage` from d_src */
int b200ldlt_dist_xp_get_stage(void *fkeep, int stage, void *d_dst, int n);
int b200ldlt_dist_xp_merge_stage(void *fkeep, int stage, const void *d_src, int n);

/* ---- per-front results, for parity checks against the oracle ---------------------------------- */
int b200ldlt_node_info(void *fkeep, int node, int *m, int *n, int *nelim);
int b200ldlt_node_copy(void *fkeep, int node, double *L /* m*nelim */, double *d /* 2n */,
                       int *perm /* n */);

const char *b200ldlt_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200LDLT_H */
