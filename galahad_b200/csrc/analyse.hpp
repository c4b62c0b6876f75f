// Symbolic analysis for the b200ldlt engine (CPU side; the north-star keeps ordering and symbolic
// analysis on the host).  All arrays here are 0-based.  The integer results are required to be
// bit-identical (after the +1 shift) to what the reference computes in
//   src/ssids/anal.f90 (check_order :147, expand_pattern :37, build_map :1129, analyse_phase :940,
//   find_subtree_partition :284) and src/spral/core_analyse.f90 (basic_analyse :38 and children).
#pragma once
#include <cstdint>
#include <vector>

namespace b200 {

struct Symbolic {
   int n = 0;
   int64_t nz = 0;                 // entries of the lower triangle (CSC) handed to analyse
   std::vector<int64_t> ptr;       // lower-CSC column pointers (copied at analyse)
   std::vector<int> row;           // lower-CSC row indices
   std::vector<int> order;         // order[i] = pivot position of variable i (-1: structurally absent)
   std::vector<int> invp;          // invp[p] = variable at pivot position p
   int nnodes = 0;
   std::vector<int> sptr;          // nnodes+1, first pivot of each supernode
   std::vector<int> sparent;       // nnodes, parent supernode (nnodes == virtual root)
   std::vector<int64_t> rptr;      // nnodes+1
   std::vector<int> rlist;         // row lists (pivot-order indices), ascending per node
   std::vector<int64_t> nptr;      // nnodes+1
   std::vector<int64_t> nlist;     // 2*nz: (src position in val, dest = col*blkm + row), per node
   int64_t num_factor = 0;
   int64_t num_flops = 0;
   int maxfront = 0;               // max #columns of a supernode (anal.f90:1100)
   int maxdepth = 0;
   int matrix_rank = 0;            // structural rank = sptr[nnodes]
   int realn = 0;
   int flag = 0;                   // 0 ok, 6 = SSIDS_WARNING_ANAL_SINGULAR, <0 error
   // subtree partition (find_subtree_partition); parts are contiguous node ranges
   int nparts = 1;
   std::vector<int> part;          // nparts+1
   std::vector<int> exec_loc;      // nparts; -1 = root part, else owning device
   std::vector<int> contrib_ptr;   // nparts+1
   std::vector<int> contrib_idx;   // contributing parts
   std::vector<int> contrib_dest;  // destination node of each contribution
   // proportional mapping of the whole tree onto `ndev` devices (proportional_mapping): owner and stage of
   // every node.  A node's stage is larger than the stage of every child owned by another device, so stage k
   // can be factorised everywhere once the boundary fronts of the stages < k have been exchanged.
   std::vector<int> map_owner, map_stage;
   int nstages = 1;
};

enum : int {
   SSIDS_SUCCESS = 0,
   SSIDS_ERROR_CALL_SEQUENCE = -1,
   SSIDS_ERROR_A_N_OOR = -2,
   SSIDS_ERROR_A_PTR = -3,
   SSIDS_ERROR_A_ALL_OOR = -4,
   SSIDS_ERROR_SINGULAR = -5,
   SSIDS_ERROR_NOT_POS_DEF = -6,
   SSIDS_ERROR_PTR_ROW = -7,
   SSIDS_ERROR_ORDER = -8,
   SSIDS_ERROR_VAL = -9,
   SSIDS_ERROR_X_SIZE = -10,
   SSIDS_ERROR_JOB_OOR = -11,
   SSIDS_ERROR_NOT_LLT = -13,
   SSIDS_ERROR_NOT_LDLT = -14,
   SSIDS_ERROR_NO_SAVED_SCALING = -15,
   SSIDS_ERROR_ALLOCATION = -50,
   SSIDS_ERROR_CUDA_UNKNOWN = -51,
   SSIDS_ERROR_CUBLAS_UNKNOWN = -52,
   SSIDS_ERROR_UNIMPLEMENTED = -98,
   SSIDS_ERROR_UNKNOWN = -99,
   SSIDS_WARNING_ANAL_SINGULAR = 6,
   SSIDS_WARNING_FACT_SINGULAR = 7,
};

// Lower-triangular CSC pattern (ptr/row, 0-based) + user order (order[i] = position, 0-based).
// Returns sym.flag.  nemin < 1 selects the reference default (32).
int analyse(int n, const int64_t* ptr, const int* row, const int* order, int nemin, Symbolic& sym);

// Symmetric infinity-norm equilibration (options%scaling = 4 of SSIDS): restatement of inf_norm_equilib_sym,
// src/spral/scaling.f90:480-521 (Algorithm 1 of Knight, Ruiz & Ucar), with the reference's defaults
// max_iterations = 10, tol = 1e-8 (a default REAL there, hence the float).  ptr/row: lower triangle by columns,
// 0-based.  Same operations in the same order; max is order-independent, so the result is the reference's bits.
int equilib_scale_sym(int n, const int64_t* ptr, const int* row, const double* val, double* scaling,
                      int max_iterations = 10, float tol = 1e-8f);

// Matching-based scaling by the Hungarian algorithm (options%scaling = 1 of SSIDS; scaling.cpp): restatement of
// hungarian_scale_sym, src/spral/scaling.f90:134-170 with hungarian_wrapper / hungarian_match (MC64's algorithm).  Lower
// triangle by columns, 0-based.  Returns 0, 1 (structurally singular, scaled by Duff and Pralet's correction because
// scale_if_singular) or -2 (structurally singular and not scale_if_singular: scaling = 1).  match_out (optional, n): the
// matching, 1-based, negative for the completion of the singular part; matched_out: its cardinality.
int hungarian_scale_sym(int n, const int64_t* ptr, const int* row, const double* val, double* scaling, bool scale_if_singular,
                        int* match_out = nullptr, int* matched_out = nullptr);

// Matching-based scaling by SPRAL's auction algorithm (options%scaling = 2): restatement of auction_scale_sym,
// src/spral/scaling.f90:269-309 with auction_match / auction_match_core and the reference's default auction_options.
// match_out (optional, n): row i is matched to column match[i] (1-based, 0 = unmatched).  Returns 0.
int auction_scale_sym(int n, const int64_t* ptr, const int* row, const double* val, double* scaling, int* match_out = nullptr,
                      int* matched_out = nullptr, int* iterations_out = nullptr);

// Matching-based ordering (options%ordering = 2 of SSIDS; scaling.cpp): restatement of match_order_metis,
// src/spral/match_order.f90:135-208, with this library's orderings on the compressed graph (kind 1: minimum degree, else
// nested dissection) where the reference calls METIS.  order[i] = pivot position of variable i (0-based), scale (n, may be
// NULL) = the matching's scaling, pairs (n, may be NULL) = the matching split into singletons and pairs (1-based partner,
// -1 singleton, -2 unmatched), cptr / crow (may be NULL) = lower triangle of the compressed pattern, 0-based.  Returns 0,
// 1 (structurally singular) or an SSIDS error flag.
int match_order(int n, const int64_t* ptr, const int* row, const double* val, int kind, int* order, double* scale,
                int* pairs = nullptr, std::vector<int64_t>* cptr = nullptr, std::vector<int>* crow = nullptr);

// Fill-reducing orderings (ordering.cpp): approximate minimum degree, and nested dissection from level structures with
// minimum-degree leaves.  Lower triangle by columns, 0-based; order[i] = pivot position of variable i.  0 or an SSIDS flag.
int amd_order(int n, const int64_t* ptr, const int* row, int* order);
int nd_order(int n, const int64_t* ptr, const int* row, int* order, int leaf_size = 200);

// Proportional mapping (the multi-level generalisation of the subtree partition: SURVEY section 8e): the root
// gets all devices, the devices of a node are split among its children in proportion to their subtree flops,
// a subtree that is left with one device is a leaf part of that device; a node is owned by the first device of
// its range.  Fills map_owner / map_stage / nstages.
void proportional_mapping(Symbolic& sym, int ndev);

// Partition the assembly tree into `ndev` leaf parts + one root part following
// anal.f90:284-459 with a GPU-only topology of `ndev` equal devices.
void find_subtree_partition(Symbolic& sym, int ndev, double max_load_inbalance, int64_t min_gpu_work);

// Engine-internal tree: absorb chain links (last child into parent) while the explicit zeros stay
// below max_zero_frac of the fused L.  Reported symbolic results remain those of `in`.
void fuse_chains(const Symbolic& in, Symbolic& out, double max_zero_frac);

}  // namespace b200
