// Numeric (GPU) side of b200ldlt: interface between the C ABI (capi.cpp) and the CUDA engine
// (engine.cu).  No CPU numeric path exists behind this interface.
#pragma once
#include <cstdint>

#include "analyse.hpp"

namespace b200 {

struct FactorOptions {
   int posdef = 0;
   int action = 1;
   double u = 0.01;
   double small_ = 1e-20;
   double multiplier = 1.1;
   int print_level = -1;
   int profile = 0;     // bench roofline: CUDA events around every trailing-update launch (1), or around the launches of
                        // the large-front kernel k_upd3 only (2); single stream, no look-ahead
};

struct FactorStats {
   int flag = 0;
   int num_delay = 0, num_neg = 0, num_two = 0, num_zero = 0, maxfront = 0;
   int not_first_pass = 0, not_second_pass = 0;
   int cuda_error = 0;
   double t_total = 0, t_device = 0;   // seconds (CUDA events)
   int64_t factor_bytes = 0;
   int64_t launches = 0;               // kernels launched by the last factor call
   // profile mode: the dominant kernel (trailing update), measured with CUDA events on its stream
   double upd_flops = 0, upd_seconds = 0;
   int64_t upd_launches = 0;
};

class Plan;      // static per-akeep structures (levels, maps) on host + device
class Numeric;   // per-fkeep factors on device

// Restrict a plan to the node range [first,last] (a part of the subtree partition); -1,-1 = all.
Plan* plan_create(const Symbolic& sym, int device, int* cuda_error);
void plan_destroy(Plan*);
int plan_device(const Plan*);

Numeric* numeric_create(Plan*);
void numeric_destroy(Numeric*);
int numeric_set_value_map(Numeric*, int ne, const int* maps);
int numeric_set_scaling(Numeric*, const double* scale);
int numeric_set_user_matrix(Numeric*, int ne, const int* row, const int* col);
int numeric_solve_refined(Numeric*, double* x, int max_ref, double stop_abs, double stop_rel, int* nref, double* seconds);
int numeric_factor_mapped(Numeric*, int ne, const double* uval, const FactorOptions&, FactorStats&);
int numeric_factor_changed(Numeric*, int nchg, const int* which, const double* newval, const FactorOptions&, FactorStats&);
int numeric_factor(Numeric*, const double* val, bool val_on_device, const FactorOptions&,
                   FactorStats&);
// x(ldx,nrhs) indexed by ORIGINAL variable; job 0..4
int numeric_solve(Numeric*, int job, int nrhs, double* x, int ldx, bool x_on_device,
                  double* seconds, int64_t* launches);
// indexed by pivot position within the analysis order (perm values): piv_order[n], d[2n]
int numeric_enquire(Numeric*, int* piv_order, double* d);
int numeric_alter(Numeric*, const double* d);
bool numeric_is_posdef(const Numeric*);
bool numeric_has_factors(const Numeric*);

// multi-GPU (one process per GPU): the partition lives in Symbolic (part / exec_loc)
int numeric_set_partition(Numeric*, int rank, int nranks);
int numeric_factor_phase(Numeric*, const double* val, const FactorOptions&, FactorStats&, int phase);
int64_t numeric_contrib_bytes(Numeric*, int node);
int numeric_contrib_pack(Numeric*, int node, void* d_dst);
int numeric_contrib_unpack(Numeric*, int node, const void* d_src, int64_t bytes);
int64_t numeric_xtail_bytes(Numeric*, int node);
int numeric_xtail_pack(Numeric*, int node, void* d_dst);
int numeric_xtail_unpack(Numeric*, int node, const void* d_src);
int numeric_dist_solve(Numeric*, int step, double* d_x, const double* d_src);
double* numeric_xp(Numeric*);
int numeric_xp_get(Numeric*, void* d_dst, int n);
int numeric_xp_get_stage(Numeric*, int stage, void* d_dst, int n);
int numeric_xp_merge_stage(Numeric*, int stage, const void* d_src, int n);

// debugging / parity: per-node results copied to the host
int numeric_node_info(Numeric*, int node, int* m, int* n, int* nelim);
int numeric_node_copy(Numeric*, int node, double* L /*m*nelim, ld m*/, double* d /*2n*/,
                      int* perm /*n*/);

}  // namespace b200
