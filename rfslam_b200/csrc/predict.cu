// predict.cu -- rfsrcPredict on the device (rf.c:22219-22402): route rows down every tree of a forest in the
// reference's wire format and average the leaf class proportions.
//   forest restore .. rf.c:10807-10881 (pre-order records -> links), done on the device, one warp per tree
//   routing ......... rf.c:3687-3747   (LEFT iff contPT - x >= 0, rf.c:10610)
//   leaf estimate ... rf.c:755-858     (class counts of the leaf's in-bag members); when the caller does not pass
//                                      tnCLAS they are rebuilt as restoreNodeMembership does (rf.c:3446-3900): the
//                                      bootstrap is replayed from forest$seed (rf.c:7046-7050) and the training rows
//                                      are routed down every tree
//   ensemble ........ rf.c:928-962, rf.c:8113-8125; restore mode adds the OOB ensemble and perfClas
//
// Data movement: rows arrive in chunks (p strided segments of the caller's [p][rows] matrix -> one [p][chunk] tile),
// uploaded on a copy stream into one of two tiles while the routing kernel works on the other.  Rows shard
// across GPUs with the forest replicated (frow_first / frow_count): no collective.
// Algorithmic bytes per (row, tree) of depth d: d * (16-byte record + 8-byte covariate) + 8 (SURVEY 8d B_pred).
#include <cub/cub.cuh>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "internal.h"
#include "perf.cuh"

namespace rfslam {

namespace {

// one visited node = one 16-byte load: cut value, covariate + 1 (0 = terminal), right child (internal) or leaf id (terminal)
struct __align__(16) PredRec { double cont; uint32_t parm; uint32_t link; };

// a factor split (mwcpSZ > 0): parm carries kFacBit, `cont` holds the MWCP word (bit level-1 set = LEFT, rf.c:1607-1616)
constexpr uint32_t kFacBit = 0x40000000u;
__device__ __forceinline__ uint32_t rec_cov(uint32_t parm) { return (parm & ~kFacBit) - 1u; }
__device__ __forceinline__ bool rec_goes_left(const struct PredRec& r, double xv);

constexpr int kLinkStack = 4096;             // pending right children while restoring a tree (= the grow path's DFS stack)
constexpr int kWalk = 4;                     // trees a thread walks at once (independent chains of dependent loads)

__device__ __forceinline__ bool rec_goes_left(const PredRec& r, double xv) {
  if (r.parm & kFacBit) return (((uint32_t)__double_as_longlong(r.cont) >> ((uint32_t)xv - 1u)) & 1u) != 0u;   // splitOnFactor
  return (r.cont - xv) >= 0.0;                                       // LEFT iff cut - x >= 0 (rf.c:10610)
}

__global__ void boundary_flag_kernel(const int32_t* treeID, uint8_t* flag, int64_t NN) {
  int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < NN) flag[q] = (q == 0 || treeID[q] != treeID[q - 1]) ? 1 : 0;
}

// One warp per tree (in order of appearance): right-child links of the pre-order records, leaf count, checks.
// A leaf closes the left subtree of the innermost pending internal node: that node's right child is the next record.
__global__ void restore_links_kernel(int ntree, const int64_t* nodeOff, const int32_t* treeID, const int32_t* nodeID, const int32_t* parmID,
                                     const double* contPT, int p, uint32_t* stack, PredRec* rec, int32_t* slotTree, int32_t* slotLeaves, int* bad) {
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= ntree) return;
  const int64_t off = nodeOff[t], cnt = nodeOff[t + 1] - off;
  uint32_t* st = stack + (size_t)t * kLinkStack;
  int sp = 0, leaves = 0, prevLeaf = 0, fail = 0;
  for (int64_t base = 0; base < cnt; base += 32) {
    const int64_t k = base + lane;
    const int32_t pm = (k < cnt) ? parmID[off + k] : 1;
    if (k < cnt && (pm < 0 || (int32_t)((uint32_t)pm & ~kFacBit) > p)) fail = 3;
    const unsigned internal = __ballot_sync(0xffffffffu, pm != 0);
    const int m = (int)min((int64_t)32, cnt - base);
    // lane 0 replays the chunk; the links it finds are handed to the lanes that own the records
    if (lane == 0) {
      for (int j = 0; j < m; j++) {
        const uint32_t kk = (uint32_t)(base + j);
        if (prevLeaf) {
          if (sp == 0) { fail = 1; break; }
          const uint32_t par = st[--sp];
          rec[off + par].link = kk;                              // right child, as an offset inside the tree
        }
        if ((internal >> j) & 1u) { if (sp >= kLinkStack) { fail = 2; break; } st[sp++] = kk; prevLeaf = 0; }
        else { leaves++; prevLeaf = 1; }
      }
    }
    if (k < cnt) {
      // (the link of an internal record is written by lane 0 above: different words of the record)
      if (pm == 0) rec[off + k].link = (uint32_t)nodeID[off + k];
      rec[off + k].cont = pm ? contPT[off + k] : 0.0; rec[off + k].parm = (uint32_t)pm;
    }
    __syncwarp();
  }
  fail = __shfl_sync(0xffffffffu, fail, 0) | fail;
  sp = __shfl_sync(0xffffffffu, sp, 0);
  if (lane == 0) {
    if (sp != 0 || !prevLeaf) fail = fail ? fail : 1;
    slotTree[t] = treeID[off]; slotLeaves[t] = leaves;
  }
  if (fail) atomicExch(bad, fail);
}

// leaves must carry ids 1..leafCount of their tree
__global__ void check_leaf_ids_kernel(int ntree, const int64_t* nodeOff, const PredRec* rec, const int32_t* slotLeaves, int* bad) {
  const int t = blockIdx.y;
  const int64_t off = nodeOff[t], cnt = nodeOff[t + 1] - off;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < cnt; k += (int64_t)gridDim.x * blockDim.x) {
    const PredRec r = rec[off + k];
    if (r.parm == 0u && (r.link < 1u || r.link > (uint32_t)slotLeaves[t])) atomicExch(bad, 4);
  }
}

// chain-A replay lands on ORIGINAL rows here (the grow path counts per internal row): u16 packed two per word
__global__ void count_draws_kernel(const uint32_t* draws, size_t dstride, const int* bootSize, uint32_t* cnt32, size_t cstride32) {
  const int t = blockIdx.y;
  const int bs = bootSize[t];
  const uint32_t* dr = draws + (size_t)t * dstride;
  uint32_t* c = cnt32 + (size_t)t * cstride32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < bs; i += gridDim.x * blockDim.x) {
    const uint32_t r = dr[i];
    atomicAdd(&c[r >> 1], 1u << (16u * (r & 1u)));
  }
}
__global__ void user_counts_kernel(const int32_t* samp, uint16_t* cnt, int n, int* bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { int32_t v = samp[i]; if (v < 0 || v > 65535) { atomicExch(bad, 5); v = 0; } cnt[i] = (uint16_t)v; }
}
__global__ void widen_counts_kernel(const uint16_t* cnt, size_t cstride, int32_t* out, int n, int ntree) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y;
  if (i < n && t < ntree) out[(size_t)t * n + i] = (int32_t)cnt[(size_t)t * cstride + i];
}

struct RouteArgs {
  int ntree, p;
  const int64_t* nodeOff;        // by slot (order of appearance)
  const int32_t* slotOfTree;     // tree id - 1 -> slot
  const int64_t* leafOff;        // by tree id - 1
  const PredRec* rec;
  const double* x;               // [p][pitch] tile
  size_t pitch;
  int rows;                      // rows in the tile
  int64_t row0;                  // first row of the tile in the output arrays
};

__device__ __forceinline__ uint32_t walk_tree(const RouteArgs& a, int tree, int i, unsigned long long& steps) {
  const int64_t off = a.nodeOff[a.slotOfTree[tree]];
  int64_t q = off;
  PredRec r = a.rec[q];
  while (r.parm != 0u) {
    const double xv = a.x[(size_t)rec_cov(r.parm) * a.pitch + i];
    q = rec_goes_left(r, xv) ? q + 1 : off + r.link;
    r = a.rec[q];
    steps++;
  }
  return r.link;
}

// kWalk trees at a time per row: the walks are independent chains of dependent loads, interleaving them multiplies
// the loads in flight; leaf ids come back in tree order so every sum below is taken in tree order
__device__ __forceinline__ void walk_group(const RouteArgs& a, int t0, int i, uint32_t (&leaf)[kWalk], unsigned long long& steps) {
  int64_t off[kWalk], q[kWalk]; PredRec r[kWalk]; bool live[kWalk];
#pragma unroll
  for (int u = 0; u < kWalk; u++) {
    live[u] = t0 + u < a.ntree;
    off[u] = live[u] ? a.nodeOff[a.slotOfTree[t0 + u]] : 0;
    q[u] = off[u];
    r[u] = live[u] ? a.rec[q[u]] : PredRec{0.0, 0u, 0u};
  }
  bool any = true;
  while (any) {
    any = false;
    double xv[kWalk];
#pragma unroll
    for (int u = 0; u < kWalk; u++) xv[u] = (r[u].parm != 0u) ? a.x[(size_t)rec_cov(r[u].parm) * a.pitch + i] : 0.0;
#pragma unroll
    for (int u = 0; u < kWalk; u++) {
      if (r[u].parm != 0u) {
        q[u] = rec_goes_left(r[u], xv[u]) ? q[u] + 1 : off[u] + r[u].link;
        r[u] = a.rec[q[u]];
        steps++;
        any |= r[u].parm != 0u;
      }
    }
  }
#pragma unroll
  for (int u = 0; u < kWalk; u++) leaf[u] = r[u].link;
}

// pass A (only when tnCLAS is not given): in-bag class counts of every leaf from the routed TRAINING rows
__global__ void leaf_class_kernel(RouteArgs a, const uint16_t* cnt, size_t cstride, const double* y, int32_t* tnCLAS, unsigned long long* visits) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long steps = 0;
  if (i < a.rows) {
    const int64_t row = a.row0 + i;
    const int cls = (y[row] == 2.0) ? 1 : 0;
    for (int t0 = 0; t0 < a.ntree; t0 += kWalk) {
      uint32_t leaf[kWalk];
      walk_group(a, t0, i, leaf, steps);
#pragma unroll
      for (int u = 0; u < kWalk; u++) {
        if (t0 + u < a.ntree) {
          const uint16_t c = cnt[(size_t)(t0 + u) * cstride + row];
          if (c) atomicAdd(&tnCLAS[2 * (a.leafOff[t0 + u] + (int64_t)leaf[u] - 1) + cls], (int32_t)c);
        }
      }
    }
  }
  steps = warp_sum(steps);
  if ((threadIdx.x & 31) == 0 && steps) atomicAdd(visits, steps);
}

// pass B: ensembles (and membership) of the target rows; cnt != nullptr = restore mode (OOB ensemble of the training rows)
__global__ void route_kernel(RouteArgs a, const int32_t* __restrict__ tnCLAS, const uint16_t* cnt, size_t cstride,
                             double* allNum, uint32_t* allDen, double* oobNum, uint32_t* oobDen, size_t M,
                             int32_t* membership, unsigned long long* visits) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long steps = 0;
  if (i < a.rows) {
    const int64_t row = a.row0 + i;
    double a1 = 0.0, a2 = 0.0, o1 = 0.0, o2 = 0.0; uint32_t od = 0;
    for (int t0 = 0; t0 < a.ntree; t0 += kWalk) {
      uint32_t leaf[kWalk];
      walk_group(a, t0, i, leaf, steps);
#pragma unroll
      for (int u = 0; u < kWalk; u++) {
        if (t0 + u < a.ntree) {
          const int64_t li = a.leafOff[t0 + u] + (int64_t)leaf[u] - 1;
          const double c1 = (double)tnCLAS[2 * li], c2 = (double)tnCLAS[2 * li + 1];
          const double rc = c1 + c2;
          const double p1 = c1 / rc, p2 = c2 / rc;                 // rf.c:948
          a1 += p1; a2 += p2;
          if (cnt && cnt[(size_t)(t0 + u) * cstride + row] == 0) { o1 += p1; o2 += p2; od++; }
          if (membership) membership[(size_t)(t0 + u) * M + row] = (int32_t)leaf[u];
        }
      }
    }
    allNum[row] = a1; allNum[M + row] = a2; allDen[row] = (uint32_t)a.ntree;
    if (cnt) { oobNum[row] = o1; oobNum[M + row] = o2; oobDen[row] = od; }
  }
  steps = warp_sum(steps);
  if ((threadIdx.x & 31) == 0 && steps) atomicAdd(visits, steps);
}

__global__ void finalize_pred_kernel(double* ens, const uint32_t* den, size_t m) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m && den[i]) { ens[i] /= (double)den[i]; ens[m + i] /= (double)den[i]; }              // rf.c:8115-8121
}

struct Buf {
  void* p = nullptr;
  ~Buf() { if (p) cudaFree(p); }
  template <typename T> T* alloc(size_t count) {
    const size_t bytes = std::max<size_t>(count * sizeof(T), 16);
    if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); set_error("out of device memory (%zu bytes)", bytes); throw CudaFail{RFSLAM_ENOMEM}; }
    return (T*)p;
  }
};

struct Streams {
  cudaStream_t copy = nullptr; cudaEvent_t up[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr}, e0 = nullptr, e1 = nullptr, k0 = nullptr, k1 = nullptr;
  ~Streams() {
    if (copy) { cudaStreamSynchronize(copy); cudaStreamDestroy(copy); }
    for (auto e : {up[0], up[1], done[0], done[1], e0, e1, k0, k1}) if (e) cudaEventDestroy(e);
  }
};

}  // namespace
}  // namespace rfslam
#include "vimp.cuh"
namespace rfslam {
namespace {

// chain-C seeds (rf.c:6685-6692): the third block of the seed LCG; in predict / restore mode the LCG starts from 0 and the
// chain-A block is skipped (rf.c:6665)
void chain_seed_c_restore(int ntree, std::vector<int>& seedC) {
  unsigned s = 0;
  seedC.resize(ntree);
  for (int pass = 0; pass < 2; pass++)
    for (int b = 0; b < ntree; b++) {
      s = lcg_next(s); s = lcg_next(s);
      while (s == 0) s = lcg_next(s);
      if (pass == 1) seedC[b] = -(int)s;
    }
}

int validate_predict(const rfslam_predict_args* a) {
  if (!a) { set_error("null arguments"); return RFSLAM_EINVAL; }
  if (a->ntree < 1 || a->total_node_count < 1 || a->x_size < 1 || a->observation_size < 1) { set_error("bad predict arguments (ntree, totalNodeCount, xSize, observationSize)"); return RFSLAM_EINVAL; }
  if (!a->treeID || !a->nodeID || !a->parmID || !a->contPT) { set_error("forest arrays missing (treeID, nodeID, hc_zero)"); return RFSLAM_EINVAL; }
  if (a->htry != 0) { set_error("htry must be 0 (rf.c:6490-6494)"); return RFSLAM_EINVAL; }
  if (a->ptn_count != 0) { set_error("pruned forests (ptnCount > 0) are outside the accelerated path"); return RFSLAM_ENOSUP; }
  if (a->y_size > 1) { set_error("the Poisson class path takes exactly one response (ySize = %d)", a->y_size); return RFSLAM_ENOSUP; }
  if ((a->opt_low & RFSLAM_OPT_VIMP) && (a->opt_low & RFSLAM_OPT_PERF)) {       // without OPT_PERF the reference drops the request (rf.c:1409-1413)
    if (!(a->opt_low & RFSLAM_OPT_VIMP_TYP1) || (a->opt_low & RFSLAM_OPT_VIMP_TYP2)) { set_error("importance: only the permutation type (\"permute\", \"permute.ensemble\") is on the accelerated path"); return RFSLAM_ENOSUP; }
    if (a->opt_low & RFSLAM_OPT_VIMP_JOIN) { set_error("joint importance is outside the accelerated path"); return RFSLAM_ENOSUP; }
    if (a->fobservation_size > 0 && !(a->fr_size > 0 && a->fr_data)) { set_error("importance on held-out rows needs their responses (frSize, frData)"); return RFSLAM_EINVAL; }
    if (a->fobservation_size > 0 && a->frow_count > 0 && a->frow_count < a->fobservation_size) { set_error("importance on held-out rows needs the whole row set in one process (no row shard)"); return RFSLAM_ENOSUP; }
    if (a->intr_predictor_size < 0 || (a->intr_predictor_size > 0 && !a->intr_predictor)) { set_error("intrPredictor missing"); return RFSLAM_EINVAL; }
    for (int k = 0; k < a->intr_predictor_size; k++)
      if (a->intr_predictor[k] < 1 || a->intr_predictor[k] > a->x_size) { set_error("intrPredictor[%d] = %d is outside 1..xSize (rf.c:19248)", k, a->intr_predictor[k]); return RFSLAM_EINVAL; }
  }
  if (a->partial_type != 0 || a->partial_length != 0) { set_error("partial plots are outside the accelerated path"); return RFSLAM_ENOSUP; }
  if (a->sobservation_size > 0) { set_error("subsetted restore (sobservationSize > 0) is outside the accelerated path"); return RFSLAM_ENOSUP; }
  if (a->x_type) for (int c = 0; c < a->x_size; c++) {
    if (!(a->x_type[c] == 'R' || a->x_type[c] == 'I' || a->x_type[c] == 'B' || a->x_type[c] == 'C')) { set_error("covariate %d has type '%c': only real / integer / boolean / unordered-factor covariates are on the accelerated predict path", c + 1, a->x_type[c]); return RFSLAM_ENOSUP; }
    if (a->x_type[c] == 'C' && (!a->x_levels || a->x_levels[c] < 1 || a->x_levels[c] > 32)) { set_error("factor covariate %d: xLevels must be in [1, 32] (one MWCP word)", c + 1); return RFSLAM_ENOSUP; }
  }
  if (a->mwcp_total > 0 && (!a->mwcpSZ || !a->mwcpPT)) { set_error("the forest holds factor (MWCP) splits: mwcpSZ and mwcpPT are needed"); return RFSLAM_EINVAL; }
  if (a->fobservation_size > 0 && !a->fx_data) { set_error("fxData missing"); return RFSLAM_EINVAL; }
  if (a->frow_first < 0 || a->frow_count < 0 || (a->fobservation_size > 0 && (int64_t)a->frow_first + a->frow_count > a->fobservation_size)) { set_error("row shard outside [0, fobservationSize)"); return RFSLAM_EINVAL; }
  return RFSLAM_OK;
}

}  // namespace

int predict_forest(const rfslam_predict_args* a, rfslam_predict_out* out) { return predict_forest_ex(a, out, nullptr); }

int predict_forest_ex(const rfslam_predict_args* a, rfslam_predict_out* out, const int* seedC_grow) {
  memset(out, 0, sizeof(*out));
  int rc = validate_predict(a);
  if (rc) return rc;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: librfslam_b200 has no CPU fallback"); return RFSLAM_ENODEV; }
  const int ntree = a->ntree, p = a->x_size, n = a->observation_size;
  const int64_t NN = a->total_node_count;
  const bool restore = a->fobservation_size == 0;
  const bool byUser = (a->opt_low & RFSLAM_OPT_BOOT_TYP1) && (a->opt_low & RFSLAM_OPT_BOOT_TYP2);
  const bool wantMemb = (a->opt_high & RFSLAM_OPTH_MEMB_USER) != 0;
  const bool haveClas = a->tnCLAS && a->tnCLAS_len > 0;
  const bool canReplay = byUser ? a->bootstrap != nullptr : a->forest_seed != nullptr;
  const bool needReplay = restore || !haveClas || (wantMemb && canReplay);  // bootMembership is part of the membership output
  if ((restore || !haveClas) && (!a->x_data || !a->r_data)) { set_error("training xData / rData are needed to restore node membership"); return RFSLAM_EINVAL; }
  if (needReplay && !byUser && !a->forest_seed) { set_error("forest$seed is needed to replay the bootstrap (rf.c:7046-7050)"); return RFSLAM_EINVAL; }
  if (needReplay && byUser && !a->bootstrap) { set_error("bootstrap = by.user needs the samp matrix"); return RFSLAM_EINVAL; }
  const int64_t mAll = restore ? n : a->fobservation_size;
  const int64_t r0 = restore ? 0 : a->frow_first;
  const int64_t M = restore ? n : (a->frow_count > 0 ? a->frow_count : mAll - r0);
  const bool timing = getenv("RFSLAM_TIMING") != nullptr;
  auto wall = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double tw0 = wall();
  double twForest = 0, twReplay = 0, twSweep = 0;
  try {
    if (a->device >= 0) RF_CUDA(cudaSetDevice(a->device));
    Streams st;
    RF_CUDA(cudaStreamCreateWithFlags(&st.copy, cudaStreamNonBlocking));
    for (int k = 0; k < 2; k++) { RF_CUDA(cudaEventCreateWithFlags(&st.up[k], cudaEventDisableTiming)); RF_CUDA(cudaEventCreateWithFlags(&st.done[k], cudaEventDisableTiming)); }
    RF_CUDA(cudaEventCreate(&st.e0)); RF_CUDA(cudaEventCreate(&st.e1)); RF_CUDA(cudaEventCreate(&st.k0)); RF_CUDA(cudaEventCreate(&st.k1));
    RF_CUDA(cudaEventRecord(st.e0));
    long long launches = 0;

    // ---- forest: records + links on the device ----
    Buf b_tid, b_nid, b_pid, b_cpt, b_flag, b_pos, b_iota, b_noff, b_rec, b_stack, b_slotTree, b_slotLeaves, b_bad, b_tmp, b_nsel;
    int32_t* d_tid = b_tid.alloc<int32_t>(NN); int32_t* d_nid = b_nid.alloc<int32_t>(NN); int32_t* d_pid = b_pid.alloc<int32_t>(NN);
    double* d_cpt = b_cpt.alloc<double>(NN);
    RF_CUDA(cudaMemcpyAsync(d_tid, a->treeID, NN * 4, cudaMemcpyHostToDevice));
    RF_CUDA(cudaMemcpyAsync(d_nid, a->nodeID, NN * 4, cudaMemcpyHostToDevice));
    std::vector<int32_t> facParm; std::vector<double> facCont;     // (alive until the copies below are done: synchronous for pageable memory)
    if (a->mwcp_total > 0) {
      // factor splits (restoreTree rf.c:10842-10848): the node's MWCP word travels in the record's cut slot, a flag in parmID.
      // Held-out factor values are levels 1..32 (checked per tile below would cost a pass: levels outside a word's range
      // simply test a clear bit and go RIGHT, as splitOnFactor does for a level the word does not hold)
      facParm.assign(a->parmID, a->parmID + NN); facCont.assign(a->contPT, a->contPT + NN);
      int64_t at = 0;
      for (int64_t q = 0; q < NN; q++) {
        const int sz = a->mwcpSZ[q];
        if (sz == 0) continue;
        if (sz != 1 || at >= a->mwcp_total) { set_error("factor split at record %lld: mwcpSZ = %d (only single-word factors, <= 32 levels) or mwcpPT too short", (long long)q, sz); throw CudaFail{sz != 1 ? RFSLAM_ENOSUP : RFSLAM_EINVAL}; }
        const unsigned long long word = (uint32_t)a->mwcpPT[at++];
        memcpy(&facCont[q], &word, 8);
        facParm[q] = (int32_t)((uint32_t)facParm[q] | kFacBit);
      }
    }
    RF_CUDA(cudaMemcpyAsync(d_pid, a->mwcp_total > 0 ? facParm.data() : a->parmID, NN * 4, cudaMemcpyHostToDevice));
    RF_CUDA(cudaMemcpyAsync(d_cpt, a->mwcp_total > 0 ? facCont.data() : a->contPT, NN * 8, cudaMemcpyHostToDevice));
    uint8_t* d_flag = b_flag.alloc<uint8_t>(NN);
    int64_t* d_noff = b_noff.alloc<int64_t>((size_t)ntree + 2);
    int* d_nsel = b_nsel.alloc<int>(1);
    int* d_bad = b_bad.alloc<int>(1);
    RF_CUDA(cudaMemsetAsync(d_bad, 0, 4));
    boundary_flag_kernel<<<(unsigned)((NN + 255) / 256), 256>>>(d_tid, d_flag, NN);
    launches++;
    {   // positions of the first record of every tree, in order of appearance (one-off preparation: CUB select)
      cub::CountingInputIterator<int64_t> iota(0);
      size_t tb = 0;
      // more boundaries than ntree would overrun d_noff: select into a buffer of NN only when that can happen
      cub::DeviceSelect::Flagged(nullptr, tb, iota, d_flag, d_noff, d_nsel, NN);
      void* d_tmp = b_tmp.alloc<unsigned char>(tb);
      // guard: count first
      Buf b_cnt; int* d_cnt = b_cnt.alloc<int>(1);
      size_t tb2 = 0; Buf b_tmp2;
      cub::DeviceReduce::Sum(nullptr, tb2, d_flag, d_cnt, NN);
      void* d_tmp2 = b_tmp2.alloc<unsigned char>(tb2);
      RF_CUDA(cub::DeviceReduce::Sum(d_tmp2, tb2, d_flag, d_cnt, NN));
      int nb = 0;
      RF_CUDA(cudaMemcpy(&nb, d_cnt, 4, cudaMemcpyDeviceToHost));
      if (nb != ntree) { set_error("the forest arrays hold %d trees, ntree = %d (records of one tree must be contiguous)", nb, ntree); throw CudaFail{RFSLAM_EINVAL}; }
      RF_CUDA(cub::DeviceSelect::Flagged(d_tmp, tb, iota, d_flag, d_noff, d_nsel, NN));
      RF_CUDA(cudaMemcpyAsync(d_noff + ntree, &NN, 8, cudaMemcpyHostToDevice));
      launches += 2;
    }
    PredRec* d_rec = b_rec.alloc<PredRec>(NN);
    uint32_t* d_stack = b_stack.alloc<uint32_t>((size_t)ntree * kLinkStack);
    int32_t* d_slotTree = b_slotTree.alloc<int32_t>(ntree); int32_t* d_slotLeaves = b_slotLeaves.alloc<int32_t>(ntree);
    restore_links_kernel<<<(ntree + 3) / 4, 128>>>(ntree, d_noff, d_tid, d_nid, d_pid, d_cpt, p, d_stack, d_rec, d_slotTree, d_slotLeaves, d_bad);
    {
      dim3 g(64, (unsigned)ntree);
      check_leaf_ids_kernel<<<g, 256>>>(ntree, d_noff, d_rec, d_slotLeaves, d_bad);
    }
    launches += 2;
    RF_CUDA(cudaGetLastError());
    std::vector<int32_t> slotTree(ntree), slotLeaves(ntree);
    int bad = 0;
    RF_CUDA(cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost));
    if (bad) {
      const char* why = bad == 1 ? "malformed pre-order forest" : bad == 2 ? "tree deeper than the restore stack" : bad == 3 ? "parmID out of range" : "leaf nodeID outside 1..leafCount";
      set_error("%s", why); throw CudaFail{bad == 2 ? RFSLAM_ENOSUP : RFSLAM_EINVAL};
    }
    RF_CUDA(cudaMemcpy(slotTree.data(), d_slotTree, (size_t)ntree * 4, cudaMemcpyDeviceToHost));
    RF_CUDA(cudaMemcpy(slotLeaves.data(), d_slotLeaves, (size_t)ntree * 4, cudaMemcpyDeviceToHost));
    // per-leaf arrays are stored by tree id (R/rfsrc.R:1846-1861), records by appearance (treeID column)
    std::vector<int32_t> slotOfTree(ntree, -1), leafCount(ntree, 0);
    for (int s = 0; s < ntree; s++) {
      const int id = slotTree[s];
      if (id < 1 || id > ntree || slotOfTree[id - 1] >= 0) { set_error("treeID values must be a permutation of 1..ntree"); throw CudaFail{RFSLAM_EINVAL}; }
      slotOfTree[id - 1] = s; leafCount[id - 1] = slotLeaves[s];
    }
    std::vector<int64_t> leafOff(ntree + 1, 0);
    for (int b = 0; b < ntree; b++) leafOff[b + 1] = leafOff[b] + leafCount[b];
    const int64_t NL = leafOff[ntree];
    if (haveClas && a->tnCLAS_len != 2 * NL) { set_error("tnCLAS holds %lld entries, the forest has %lld leaves x 2 classes", (long long)a->tnCLAS_len, (long long)NL); throw CudaFail{RFSLAM_EINVAL}; }
    Buf b_sot, b_loff, b_clas;
    int32_t* d_sot = b_sot.alloc<int32_t>(ntree); int64_t* d_loff = b_loff.alloc<int64_t>((size_t)ntree + 1);
    int32_t* d_clas = b_clas.alloc<int32_t>(2 * (size_t)NL);
    RF_CUDA(cudaMemcpy(d_sot, slotOfTree.data(), (size_t)ntree * 4, cudaMemcpyHostToDevice));
    RF_CUDA(cudaMemcpy(d_loff, leafOff.data(), ((size_t)ntree + 1) * 8, cudaMemcpyHostToDevice));
    if (haveClas) RF_CUDA(cudaMemcpyAsync(d_clas, a->tnCLAS, 2 * (size_t)NL * 4, cudaMemcpyHostToDevice));
    else RF_CUDA(cudaMemsetAsync(d_clas, 0, 2 * (size_t)NL * 4));

    if (timing) { RF_CUDA(cudaDeviceSynchronize()); twForest = wall(); }
    // ---- bootstrap replay (restore mode, missing tnCLAS, or bootMembership wanted) ----
    Buf b_cnt16, b_y, b_seed, b_bs, b_draws, b_samp;
    uint16_t* d_cnt = nullptr; double* d_y = nullptr;
    const size_t cstride = ((size_t)n + 1) & ~(size_t)1;
    if (needReplay) {
      d_cnt = b_cnt16.alloc<uint16_t>((size_t)ntree * cstride);
      RF_CUDA(cudaMemsetAsync(d_cnt, 0, (size_t)ntree * cstride * 2));
      if (a->r_data) { d_y = b_y.alloc<double>(n); RF_CUDA(cudaMemcpyAsync(d_y, a->r_data, (size_t)n * 8, cudaMemcpyHostToDevice)); }
      if (byUser) {
        int32_t* d_samp = b_samp.alloc<int32_t>(n);
        for (int b = 0; b < ntree; b++) {
          RF_CUDA(cudaMemcpy(d_samp, a->bootstrap + (size_t)b * n, (size_t)n * 4, cudaMemcpyHostToDevice));
          user_counts_kernel<<<(n + 255) / 256, 256>>>(d_samp, d_cnt + (size_t)b * cstride, n, d_bad);
          launches++;
        }
      } else {
        std::vector<int> bs(ntree);
        int maxBoot = 1;
        for (int b = 0; b < ntree; b++) { bs[b] = a->bootstrap_size ? a->bootstrap_size[b] : n; if (bs[b] < 1) { set_error("empty bootstrap sample for tree %d", b + 1); throw CudaFail{RFSLAM_EINVAL}; } maxBoot = std::max(maxBoot, bs[b]); }
        if (maxBoot > 65535) {
          const double mean = (double)maxBoot / (double)n;
          if (mean + 12.0 * sqrt(mean) + 50.0 > 65535.0) { set_error("sampsize %d over %d rows would overflow the 16-bit in-bag counts", maxBoot, n); throw CudaFail{RFSLAM_ENOSUP}; }
        }
        int* d_seed = b_seed.alloc<int>(ntree); int* d_bs = b_bs.alloc<int>(ntree);
        RF_CUDA(cudaMemcpy(d_seed, a->forest_seed, (size_t)ntree * 4, cudaMemcpyHostToDevice));      // chain A is re-seeded from forest$seed
        RF_CUDA(cudaMemcpy(d_bs, bs.data(), (size_t)ntree * 4, cudaMemcpyHostToDevice));
        // replay in batches of trees to bound the draw buffer
        const int tb = (int)std::max<size_t>(1, std::min<size_t>((size_t)ntree, ((size_t)2 << 30) / ((size_t)maxBoot * 4)));
        uint32_t* d_draws = b_draws.alloc<uint32_t>((size_t)tb * maxBoot);
        for (int b0 = 0; b0 < ntree; b0 += tb) {
          const int nb = std::min(tb, ntree - b0);
          launch_bootstrap_draw(d_seed + b0, d_bs + b0, d_draws, (size_t)maxBoot, n, nb);
          dim3 gc((unsigned)std::min<size_t>(((size_t)maxBoot + 255) / 256, 1024), (unsigned)nb);
          count_draws_kernel<<<gc, 256>>>(d_draws, (size_t)maxBoot, d_bs + b0, (uint32_t*)(d_cnt + (size_t)b0 * cstride), cstride / 2);
          launches += 2;
        }
      }
      RF_CUDA(cudaGetLastError());
    }

    if (timing) { RF_CUDA(cudaDeviceSynchronize()); twReplay = wall(); }
    // ---- row tiles ----
    size_t freeB = 0, totalB = 0;
    RF_CUDA(cudaMemGetInfo(&freeB, &totalB));
    size_t tileRows = std::min<size_t>((size_t)std::max<int64_t>(M, n), std::max<size_t>(4096, ((size_t)512 << 20) / ((size_t)p * 8)));
    if (const char* e = getenv("RFSLAM_PREDICT_TILE_ROWS")) { long v = atol(e); if (v > 0) tileRows = (size_t)v; }
    tileRows = (tileRows + 127) & ~(size_t)127;
    Buf b_tile[2];
    double* d_tile[2] = {b_tile[0].alloc<double>((size_t)p * tileRows), b_tile[1].alloc<double>((size_t)p * tileRows)};
    Buf b_visits;
    unsigned long long* d_visits = b_visits.alloc<unsigned long long>(1);
    RF_CUDA(cudaMemsetAsync(d_visits, 0, 8));
    float kernMs = 0.f;
    int tileUse = 0;
    // sweep rows [first, first + count) of a host matrix [p][pitch]: upload tile k + 1 while the kernel works on tile k
    auto sweep = [&](const double* host, size_t pitch, int64_t first, int64_t count, auto&& launch) {
      const int64_t ntiles = (count + (int64_t)tileRows - 1) / (int64_t)tileRows;
      auto upload = [&](int64_t ti) {
        const int k = (int)((tileUse + ti) & 1);
        const int64_t lo = first + ti * (int64_t)tileRows, rows = std::min<int64_t>((int64_t)tileRows, first + count - lo);
        RF_CUDA(cudaStreamWaitEvent(st.copy, st.done[k], 0));                 // the kernel that last read this tile
        RF_CUDA(cudaMemcpy2DAsync(d_tile[k], tileRows * 8, host + lo, pitch * 8, (size_t)rows * 8, (size_t)p, cudaMemcpyHostToDevice, st.copy));
        RF_CUDA(cudaEventRecord(st.up[k], st.copy));
      };
      if (ntiles > 0) upload(0);
      for (int64_t ti = 0; ti < ntiles; ti++) {
        const int k = (int)((tileUse + ti) & 1);
        if (ti + 1 < ntiles) upload(ti + 1);
        const int64_t lo = first + ti * (int64_t)tileRows, rows = std::min<int64_t>((int64_t)tileRows, first + count - lo);
        RF_CUDA(cudaStreamWaitEvent(0, st.up[k], 0));
        RouteArgs ra{ntree, p, d_noff, d_sot, d_loff, d_rec, d_tile[k], tileRows, (int)rows, lo - first};
        RF_CUDA(cudaEventRecord(st.k0));
        launch(ra);
        RF_CUDA(cudaEventRecord(st.k1));
        RF_CUDA(cudaEventRecord(st.done[k]));
        launches++;
        RF_CUDA(cudaGetLastError());
        RF_CUDA(cudaEventSynchronize(st.k1));
        float ms = 0.f; RF_CUDA(cudaEventElapsedTime(&ms, st.k0, st.k1)); kernMs += ms;
      }
      tileUse += (int)(ntiles & 1);
    };

    long long rowTrees = 0;
    if (!haveClas) {                                                          // pass A: leaf class counts from the training rows
      sweep(a->x_data, (size_t)n, 0, n, [&](const RouteArgs& ra) {
        leaf_class_kernel<<<(ra.rows + 127) / 128, 128>>>(ra, d_cnt, cstride, d_y, d_clas, d_visits);
      });
      rowTrees += (long long)n * ntree;
    }
    // pass B: the target rows
    Buf b_all, b_allDen, b_oob, b_oobDen, b_memb;
    double* d_all = b_all.alloc<double>(2 * (size_t)M); uint32_t* d_allDen = b_allDen.alloc<uint32_t>((size_t)M);
    double* d_oob = restore ? b_oob.alloc<double>(2 * (size_t)M) : nullptr;
    uint32_t* d_oobDen = restore ? b_oobDen.alloc<uint32_t>((size_t)M) : nullptr;
    int32_t* d_memb = wantMemb ? b_memb.alloc<int32_t>((size_t)ntree * (size_t)M) : nullptr;
    sweep(restore ? a->x_data : a->fx_data, (size_t)mAll, r0, M, [&](const RouteArgs& ra) {
      route_kernel<<<(ra.rows + 127) / 128, 128>>>(ra, d_clas, restore ? d_cnt : nullptr, cstride, d_all, d_allDen, d_oob, d_oobDen, (size_t)M, d_memb, d_visits);
    });
    rowTrees += (long long)M * ntree;
    if (a->finalize_ensembles) {
      finalize_pred_kernel<<<(unsigned)((M + 255) / 256), 256>>>(d_all, d_allDen, (size_t)M);
      if (restore) finalize_pred_kernel<<<(unsigned)((M + 255) / 256), 256>>>(d_oob, d_oobDen, (size_t)M);
      launches += restore ? 2 : 1;
    }
    double* h_perf = nullptr;
    if (restore && a->finalize_ensembles && (a->opt_low & RFSLAM_OPT_PERF)) {
      h_perf = perf_clas_table(a->opt_low, ntree, n, d_oob, d_oobDen, nullptr, nullptr, d_y, &launches);
      out->perfClas = h_perf;                  // owned by `out` from here on: the catch below releases it with the rest
    }
    // predict mode with held-out responses (frSize > 0): performance of the full ensemble on the held-out rows (rf.c:1323-1326,
    // rf.c:7770-7800) -- for the whole set of rows only (a row shard holds a part of the counts)
    Buf b_fy;
    double* d_fy = nullptr;
    unsigned long long trainCls[2] = {0, 0};
    if (!restore && a->finalize_ensembles && (a->opt_low & RFSLAM_OPT_PERF) && a->fr_size > 0 && a->fr_data && r0 == 0 && M == mAll) {
      d_fy = b_fy.alloc<double>((size_t)M);
      RF_CUDA(cudaMemcpy(d_fy, a->fr_data, (size_t)M * 8, cudaMemcpyHostToDevice));
      if (a->opt_low & RFSLAM_OPT_CLAS_RFQ) {
        if (!a->r_data) { set_error("rfq performance on held-out rows needs the training response (rData)"); throw CudaFail{RFSLAM_EINVAL}; }
        for (int i = 0; i < n; i++) trainCls[a->r_data[i] == 2.0 ? 1 : 0]++;
      }
      h_perf = perf_clas_table(a->opt_low, ntree, (int)M, d_all, d_allDen, nullptr, nullptr, d_fy, &launches,
                               (a->opt_low & RFSLAM_OPT_CLAS_RFQ) ? trainCls : nullptr);
      out->perfClas = h_perf;
    }
    // ---- permutation importance (restore mode, or held-out rows with responses; the grow entry comes here with its own chain-C seeds) ----
    double* h_vimp = nullptr;
    const int nIntr = a->intr_predictor_size;
    if (h_perf && (a->opt_low & RFSLAM_OPT_VIMP) && nIntr > 0) {
      // rows the importance is taken over: the OOB rows of every tree (restore / grow), or ALL held-out rows of every tree
      // (predict mode with responses: membershipIndex = identity, rf.c:2125-2131)
      const bool leob = (a->opt_low & RFSLAM_OPT_VIMP_LEOB) != 0;
      const int vn = restore ? n : (int)M;
      const double* hx = restore ? a->x_data : a->fx_data;
      const double* hy = restore ? a->r_data : a->fr_data;
      const double* d_vy = restore ? d_y : d_fy;
      const uint16_t* d_vcnt = restore ? d_cnt : nullptr;
      RF_CUDA(cudaEventRecord(st.k0));
      Buf b_x, b_oobc, b_seedC, b_intr, b_scr, b_acc, b_vens, b_vden;
      double* d_x = b_x.alloc<double>((size_t)p * vn);
      RF_CUDA(cudaMemcpy(d_x, hx, (size_t)p * vn * 8, cudaMemcpyHostToDevice));
      int maxOob = vn;
      if (restore) {
        int* d_oobc = b_oobc.alloc<int>(ntree);
        vimp_oob_count_kernel<<<ntree, kVimpThreads>>>(d_cnt, cstride, n, ntree, d_oobc);
        std::vector<int> oobc(ntree);
        RF_CUDA(cudaMemcpy(oobc.data(), d_oobc, (size_t)ntree * 4, cudaMemcpyDeviceToHost));
        maxOob = std::max(1, *std::max_element(oobc.begin(), oobc.end()));
      }
      const size_t smem = SlotIndex::bytes(maxOob);
      if (smem > 220 * 1024) { set_error("importance: %d OOB rows in one tree exceed the ~1.7 M the permutation's shared-memory slot index holds", maxOob); throw CudaFail{RFSLAM_ENOSUP}; }
      RF_CUDA(cudaFuncSetAttribute(vimp_tree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int perSm = 0, nsm = 0, dev = 0;
      RF_CUDA(cudaGetDevice(&dev));
      RF_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
      RF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, vimp_tree_kernel, kVimpThreads, smem));
      const int grid = std::max(1, std::min(ntree, nsm * std::max(1, perSm)));
      std::vector<int> seedC;
      if (seedC_grow) seedC.assign(seedC_grow, seedC_grow + ntree); else chain_seed_c_restore(ntree, seedC);
      std::vector<int> intr0(nIntr);
      for (int k = 0; k < nIntr; k++) intr0[k] = a->intr_predictor[k] - 1;
      int* d_seedC = b_seedC.alloc<int>(ntree); int* d_intr = b_intr.alloc<int>(nIntr);
      RF_CUDA(cudaMemcpy(d_seedC, seedC.data(), (size_t)ntree * 4, cudaMemcpyHostToDevice));
      RF_CUDA(cudaMemcpy(d_intr, intr0.data(), (size_t)nIntr * 4, cudaMemcpyHostToDevice));
      unsigned long long cls[2] = {0, 0};                             // of the rows scored: the per-class denominators (rf.c:1037-1040)
      for (int i = 0; i < vn; i++) cls[hy[i] == 2.0 ? 1 : 0]++;
      const unsigned long long* tcls = restore ? cls : trainCls;      // of the training rows: the RFQ vote (rf.c:16859-16895)
      VimpArgs va{};
      va.ntree = ntree; va.p = p; va.n = vn; va.nIntr = nIntr;
      va.nodeOff = d_noff; va.slotOfTree = d_sot; va.leafOff = d_loff; va.rec = d_rec; va.tnCLAS = d_clas;
      va.cnt = d_vcnt; va.cstride = cstride; va.x = d_x; va.y = d_vy; va.seedC = d_seedC; va.intr = d_intr;
      va.maxOob = maxOob; va.scratch = b_scr.alloc<uint32_t>((size_t)grid * 3 * maxOob);
      va.leob = leob ? 1 : 0; va.rfqMinority = -1; va.rfqThreshold = 0.0;
      if (a->opt_low & RFSLAM_OPT_CLAS_RFQ) { va.rfqMinority = tcls[1] < tcls[0] ? 1 : 0; va.rfqThreshold = (double)tcls[va.rfqMinority] / (double)(tcls[0] + tcls[1]); }
      va.acc = b_acc.alloc<VimpAcc>((size_t)ntree * (nIntr + 1));
      if (!leob) {
        va.vens = b_vens.alloc<double>((size_t)nIntr * 2 * vn); va.vden = b_vden.alloc<uint32_t>((size_t)nIntr * vn);
        RF_CUDA(cudaMemsetAsync(va.vens, 0, (size_t)nIntr * 2 * vn * 8)); RF_CUDA(cudaMemsetAsync(va.vden, 0, (size_t)nIntr * vn * 4));
      }
      vimp_tree_kernel<<<grid, kVimpThreads, smem>>>(va);
      launches += 2;
      RF_CUDA(cudaGetLastError());
      h_vimp = (double*)host_alloc((size_t)nIntr * 3 * 8);
      if (!h_vimp) { set_error("out of host memory"); throw CudaFail{RFSLAM_ENOMEM}; }
      out->vimpClas = h_vimp; out->vimp_count = nIntr;
      const double na = rfslam_na_real();
      if (leob) {
        // per tree: performance of the tree's own OOB estimate with covariate v permuted minus without, averaged over the
        // trees where both exist; the per-class slots are scaled by e (rf.c:2599-2633)
        std::vector<VimpAcc> acc((size_t)ntree * (nIntr + 1));
        RF_CUDA(cudaMemcpy(acc.data(), va.acc, acc.size() * sizeof(VimpAcc), cudaMemcpyDeviceToHost));
        auto row = [&](const VimpAcc& c, double* o) {
          o[0] = o[1] = o[2] = na;
          const unsigned long long cum[2] = {c.cum[0], c.cum[1]}, ok[2] = {c.ok[0], c.ok[1]};
          perf_from_counts(a->opt_low, cum, ok, cls, c.brier, o);
        };
        std::vector<double> sum((size_t)nIntr * 3, 0.0); std::vector<int> cntv((size_t)nIntr * 3, 0);
        for (int b = 0; b < ntree; b++) {
          double p0[3]; row(acc[(size_t)b * (nIntr + 1)], p0);
          for (int v = 0; v < nIntr; v++) {
            double pv[3]; row(acc[(size_t)b * (nIntr + 1) + v + 1], pv);
            for (int k = 0; k < 3; k++) if (pv[k] == pv[k] && p0[k] == p0[k]) { sum[v * 3 + k] += pv[k] - p0[k]; cntv[v * 3 + k]++; }
          }
        }
        for (int v = 0; v < nIntr; v++) for (int k = 0; k < 3; k++) {
          const int c = cntv[v * 3 + k];
          h_vimp[v * 3 + k] = c ? (k > 0 ? M_E * sum[v * 3 + k] / (double)c : sum[v * 3 + k] / (double)c) : na;
        }
      } else {
        // ensemble: the perturbed OOB ensemble of covariate v against the forest's (rf.c:2416-2470, 2685-2694)
        const double* f0 = h_perf + (size_t)(ntree - 1) * 3;
        for (int v = 0; v < nIntr; v++) {
          double* ev = va.vens + (size_t)v * 2 * vn; uint32_t* dv = va.vden + (size_t)v * vn;
          finalize_pred_kernel<<<(unsigned)((vn + 255) / 256), 256>>>(ev, dv, (size_t)vn);
          launches++;
          double* r = perf_clas_table(a->opt_low, 1, vn, ev, dv, nullptr, nullptr, d_vy, &launches,
                                      (!restore && (a->opt_low & RFSLAM_OPT_CLAS_RFQ)) ? trainCls : nullptr);
          for (int k = 0; k < 3; k++) h_vimp[v * 3 + k] = (r[k] == r[k] && f0[k] == f0[k]) ? r[k] - f0[k] : r[k];
          host_free(r);
        }
      }
      RF_CUDA(cudaEventRecord(st.k1));
      RF_CUDA(cudaEventSynchronize(st.k1));
      float ms = 0.f; RF_CUDA(cudaEventElapsedTime(&ms, st.k0, st.k1));
      out->vimp_ms = ms;
    }
    RF_CUDA(cudaEventRecord(st.e1));
    RF_CUDA(cudaGetLastError());
    RF_CUDA(cudaEventSynchronize(st.e1));
    if (timing) twSweep = wall();
    RF_CUDA(cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost));
    if (bad == 5) { set_error("by.user in-bag counts must be in [0, 65535]"); throw CudaFail{RFSLAM_EINVAL}; }
    float totMs = 0.f;
    RF_CUDA(cudaEventElapsedTime(&totMs, st.e0, st.e1));
    unsigned long long visits = 0;
    RF_CUDA(cudaMemcpy(&visits, d_visits, 8, cudaMemcpyDeviceToHost));

    out->m = (int32_t)M; out->ntree = ntree; out->kernel_ms = kernMs; out->total_device_ms = totMs;
    out->kernel_launches = launches; out->row_tree_visits = rowTrees; out->node_visits = (int64_t)visits;
    auto fetch = [&](const void* dev, size_t bytes) -> void* {
      void* h = host_alloc(std::max<size_t>(bytes, 8));
      if (!h) { set_error("out of host memory"); throw CudaFail{RFSLAM_ENOMEM}; }
      if (bytes) RF_CUDA(cudaMemcpy(h, dev, bytes, cudaMemcpyDeviceToHost));
      return h;
    };
    out->allEnsbCLS = (double*)fetch(d_all, 2 * (size_t)M * 8);
    out->allEnsbDen = (uint32_t*)fetch(d_allDen, (size_t)M * 4);
    if (restore) { out->oobEnsbCLS = (double*)fetch(d_oob, 2 * (size_t)M * 8); out->oobEnsbDen = (uint32_t*)fetch(d_oobDen, (size_t)M * 4); }
    out->leafCount = (int32_t*)host_alloc((size_t)ntree * 4);
    memcpy(out->leafCount, leafCount.data(), (size_t)ntree * 4);
    if (wantMemb) out->nodeMembership = (int32_t*)fetch(d_memb, (size_t)ntree * (size_t)M * 4);
    if (wantMemb && d_cnt) {
      Buf b_wide; int32_t* d_wide = b_wide.alloc<int32_t>((size_t)ntree * n);
      dim3 g((unsigned)((n + 255) / 256), (unsigned)ntree);
      widen_counts_kernel<<<g, 256>>>(d_cnt, cstride, d_wide, n, ntree);
      RF_CUDA(cudaGetLastError());
      out->bootMembership = (int32_t*)fetch(d_wide, (size_t)ntree * n * 4);
    }
    if (timing) fprintf(stderr, "[rfslam_b200_predict] forest upload + restore %.1f ms, bootstrap replay %.1f ms, row sweeps %.1f ms (routing kernels %.1f ms), download %.1f ms\n",
                        twForest - tw0, twReplay - twForest, twSweep - twReplay, (double)out->kernel_ms, wall() - twSweep);
  } catch (CudaFail f) {
    rfslam_b200_free_predict(out);
    return f.code;
  }
  if (timing) fprintf(stderr, "[rfslam_b200_predict] call %.1f ms (incl. releasing device buffers)\n", wall() - tw0);
  return RFSLAM_OK;
}

}  // namespace rfslam
