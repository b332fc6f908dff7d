// internal.h -- structures shared by the translation units of librfslam_b200.so (not part of the ABI).
#pragma once
#include <string.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

#include "../../include/rfslam_b200.h"
#include "common.cuh"

namespace rfslam {

struct __align__(16) RowRec { double rt; uint32_t es; uint32_t pad; };   // es = event | (stratum - 1) << 8

// Device view of the resident CPIU table.  Rows are held in INTERNAL order = sorted by
// (stratum, original row): every stratum is one contiguous row range, so a node's in-bag list
// (kept sorted by internal row) is stratum-major and per-stratum histograms need only D bins.
struct DevData {
  int n, p, K;                    // rows, covariates, largest stratum id
  const uint32_t* perm;           // [n] internal -> original row
  const uint32_t* inv;            // [n] original -> internal row
  const uint32_t* sb;             // [K+2] stratum k owns internal rows [sb[k], sb[k+1]); sb[0] = sb[1] = 0
  const uint8_t*  ev;             // [n] 1 = event (response code 2)
  const double*   rt;             // [n] risk time
  const uint16_t* strat;          // [n] stratum id (redundant with sb; used by the sorted-walk path)
  const void* const* codes;       // [p] rank-coded column, internal order, `width[c]` bytes per row
  const uint8_t*  width;          // [p] 1 / 2 / 4
  const int*      D;              // [p] number of distinct values
  const double* const* uniq;      // [p] the distinct values, ascending: cut value = uniq[c][code]
  const uint8_t*  ftype;          // [p] 1 = unordered factor (xType 'C'): cuts are level subsets (MWCP, rf.c:14452-14518), or nullptr
  double rt_mode;                 // the table's dominant risk time (exposure = rt_mode * count + exceptions)
  // Row-major mirror of the table (present when every column is byte-coded): one row's codes are `rowPitch`
  // contiguous bytes and its (risk time, event, stratum) one 16-byte record.  A node deep in a tree holds a small
  // scattered fraction of the rows: gathering c candidate codes of a row from the column-major arrays costs c
  // 32-byte DRAM sectors for c bytes, from here at most rowPitch / 32 sectors -- and a whole row fits the
  // shared-memory tile that serves the subtree below a small node.
  const uint8_t*  rowcodes;       // [n][rowPitch] or nullptr (built only for tables that do not fit L2 anyway)
  int             rowPitch;       // p rounded up to a multiple of 16
  const RowRec*   rowrec;         // [n] always present
};

}  // namespace rfslam

namespace rfslam {
// recycles device scratch across grow calls on a resident dataset (cudaMalloc / cudaFree of
// multi-GB buffers would otherwise be paid on every call)
struct WsPool {
  struct Blk { void* p; size_t bytes; bool used; };
  std::vector<Blk> blks;
  void* get(size_t bytes) {
    int best = -1;
    for (size_t i = 0; i < blks.size(); i++)
      if (!blks[i].used && blks[i].bytes >= bytes && blks[i].bytes <= 2 * bytes + (1u << 20) && (best < 0 || blks[i].bytes < blks[best].bytes)) best = (int)i;
    if (best >= 0) { blks[best].used = true; return blks[best].p; }
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
      cudaGetLastError();
      trim();                                   // drop idle blocks and retry once
      if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    }
    blks.push_back({p, bytes, true});
    return p;
  }
  void put(void* p) { for (auto& b : blks) if (b.p == p) { b.used = false; return; } }
  size_t idle_bytes() const { size_t s = 0; for (auto& b : blks) if (!b.used) s += b.bytes; return s; }
  void trim() {
    std::vector<Blk> keep;
    for (auto& b : blks) { if (b.used) keep.push_back(b); else cudaFree(b.p); }
    blks.swap(keep);
  }
  void release_all() { for (auto& b : blks) cudaFree(b.p); blks.clear(); }
};
}  // namespace rfslam

// The opaque handle of the C-ABI.
struct rfslam_dataset {
  int device = 0;
  int n = 0, p = 0, K = 0;
  rfslam::DevData dev{};
  std::vector<int> D;             // host mirrors
  std::vector<uint8_t> width;
  std::vector<void*> allocs;      // everything to cudaFree
  std::vector<void*> codes_h;     // device pointers per column
  std::vector<double*> uniq_h;
  int64_t bytes = 0;
  int max_D = 0;
  double max_rt = 0.0;            // largest risk time (scale of the fixed-point exposure sums)
  int hist_bins = 32;             // pow2 >= the widest column with <= 256 distinct values (and >= the cuts of the widest factor)
  bool has_factor = false;        // some covariate is an unordered factor
  std::vector<uint8_t> ftype;
  rfslam::WsPool pool;            // scratch recycled across grow calls
};

namespace rfslam {

struct GrowConfig {
  int rule;            // 1..6
  int mtry, nodesize, nodedepth;
  double alpha;        // 1 / k_for_alpha^2
  int nsplit;          // 0 = every cut; > 0 = at most nsplit random cuts per covariate
  int ntree;           // forest size (seeds depend on it, rf.c:6677-6692)
  int seed;
  bool by_user;
  bool want_membership;
  bool want_membr_lists;
  bool finalize;
};

int  dataset_build(const rfslam_grow_args* a, rfslam_dataset** out, WsPool* seed = nullptr);
void dataset_free(rfslam_dataset* ds);
int  grow_forest(rfslam_dataset* ds, const rfslam_grow_args* a, rfslam_forest_out* out);
int  predict_forest(const rfslam_predict_args* a, rfslam_predict_out* out);
// seedC != nullptr: chain-C seeds of the grow call this restore belongs to (importance inside rfslam_b200_grow)
int  predict_forest_ex(const rfslam_predict_args* a, rfslam_predict_out* out, const int* seedC);
// importance after growth (api.cu): restores the finished forest on its training rows with the grow call's chain C
int  grow_importance(const rfslam_grow_args* a, rfslam_forest_out* out);
int  score_node_impl(int rule, double k_for_alpha, int m, const double* x, const double* response,
                     const int32_t* stratum, const double* risk_time, double* cut_value, double* delta,
                     int32_t* n_cuts, int device);
int  bootstrap_counts_impl(int seed, int ntree, int n, int tree, int32_t* counts, int device);
int  validate_grow_args(const rfslam_grow_args* a, GrowConfig* cfg);
void chain_seeds(int seed, int ntree, std::vector<int>& seedA, std::vector<int>& seedB);
// chain-A replay of `ntrees` trees (grow.cu bootstrap_draw_kernel): draws[t][0..bootSize[t]) = 0-based original rows
void launch_bootstrap_draw(const int* d_seedA, const int* d_bootSize, uint32_t* d_draws, size_t stride, int n, int ntrees);

// plugin.cu: family x slot -> device rule (the reference's customFunctionArray[4][16], rf.c:486)
void ensure_rule_defaults();
void reset_rule_defaults();
int  rule_table_get(int family, int slot);            // 1..6 device rule, 0 empty, -1 foreign (user-supplied CPU function)
void rule_table_set(int family, int slot, int rule);
bool rule_slot_is_foreign(int family, int slot);

// comm.cu: the exchange step of a tree-sharded forest (NCCL over NVLink), all on device buffers
void comm_allreduce_ensembles(rfslam_comm* c, double* allNum, double* oobNum, uint32_t* allDen, uint32_t* oobDen, int n);
void comm_allreduce_sum_u64(rfslam_comm* c, unsigned long long* dev, int count);
void comm_allgather_leaf_counts(rfslam_comm* c, const std::vector<int32_t>& mine, int ntree, std::vector<int32_t>& all);
// gather-v to rank 0 + merge into ascending tree order (comm.cu); unitsOfTree[b] = units (of `width` elements) of tree b + 1.
// The destination of rank 0 is device-accessible memory: a device array or a page-locked host block.
struct GatherLayout {
  int W = 0, ntree = 0; std::vector<long long> rankUnits; long long total = 0;
  int* d_rank = nullptr; long long *d_so = nullptr, *d_do = nullptr, *d_len = nullptr;
  std::vector<void*> devKeep; std::vector<std::vector<void*>> hostKeep;
};
void comm_gather_layout(rfslam_comm* c, const std::vector<long long>& unitsOfTree, GatherLayout& L);
void comm_gather_i32_into(rfslam_comm* c, GatherLayout& L, const int32_t* local, int width, int32_t* dst);
void comm_gather_f64_into(rfslam_comm* c, GatherLayout& L, const double* local, int width, double* dst);
void comm_gather_finish(rfslam_comm* c, GatherLayout& L);     // synchronises the communicator's stream, releases the scratch
int comm_rank(const rfslam_comm* c);
int comm_world(const rfslam_comm* c);

// R's NA_real_: the NaN with payload 1954 (arithmetic NaN is a different bit pattern)
inline double rfslam_na_real() { const unsigned long long b = 0x7FF00000000007A2ull; double d; memcpy(&d, &b, 8); return d; }

// hostpool.cu: recycled page-locked memory for the arrays handed to the caller
void* host_alloc(size_t bytes);
void* host_calloc(size_t bytes);
void  host_free(void* p);      // pool blocks return to the pool, anything else goes to free()
void  host_pool_trim();

}  // namespace rfslam
