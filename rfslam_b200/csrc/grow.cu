// grow.cu -- host orchestration of forest growth plus the supporting kernels either side of the
// tree kernel: chain-A bootstrap replay (K1), record compaction, and the row-parallel traversal
// that yields node membership, all-row leaf counts and the OOB / full ensembles (K4/K5).
#include <cub/cub.cuh>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#define RF_PASS_CAND 8            // the lean 256-thread build of this translation unit serves mtry <= 8 (make_plan)
#include "grow_kernel.cuh"
#include "perf.cuh"
#include "internal.h"

namespace rfslam {

void chain_seeds(int seed, int ntree, std::vector<int>& seedA, std::vector<int>& seedB) {
  // rf.c:6666-6692: one LCG (modulus 714025) stepped twice per tree, first for all chain-A seeds,
  // then for all chain-B seeds (chain C, the vimp chain, is not on this path)
  unsigned s = (unsigned)std::abs(seed);
  if (s >= 714025u) s %= 714025u;
  seedA.resize(ntree); seedB.resize(ntree);
  for (int pass = 0; pass < 2; pass++)
    for (int b = 0; b < ntree; b++) {
      s = lcg_next(s); s = lcg_next(s);
      while (s == 0) s = lcg_next(s);
      (pass == 0 ? seedA : seedB)[b] = -(int)s;
    }
}

std::vector<unsigned long long> g_last_profile;

namespace {

// K1a: chain-A replay (rf.c:492-627), one warp per tree, 32 draws per round.  The Park-Miller states of
// a round come from jump-ahead (lane k: k + 1 steps); the Bays-Durham table lives in registers (lane j
// holds iv[j]) so the inherently sequential shuffle walk is two warp shuffles per draw; the float ->
// row index conversion and the stores are lane-parallel and coalesced.  draws are 0-based original rows.
__global__ void bootstrap_draw_kernel(const int* seedA, const int* bootSize, uint32_t* draws, size_t stride, int n) {
  __shared__ Ran1 rng;
  const int t = blockIdx.x, lane = threadIdx.x;
  if (lane == 0) { ran1_seed(rng, seedA[t]); ran1_init(rng); }
  __syncwarp();
  int ivReg = rng.iv[lane], iy = rng.iy, idum = rng.idum;
  uint32_t* out = draws + (size_t)t * stride;
  const int bs = bootSize[t];
  for (int base = 0; base < bs; base += 32) {
    const int cnt = min(32, bs - base);
    const int myIdum = ran1_mulmod(kJump[lane], (uint32_t)idum);
    int myIy = 0;
    for (int k = 0; k < cnt; k++) {
      const int j = iy / 67108864;
      const int idk = __shfl_sync(0xffffffffu, myIdum, k);
      const int niy = __shfl_sync(0xffffffffu, ivReg, j);
      if (lane == j) ivReg = idk;
      iy = niy;
      if (lane == k) myIy = niy;
    }
    idum = __shfl_sync(0xffffffffu, myIdum, cnt - 1);
    if (lane < cnt) out[base + lane] = (uint32_t)ceil((double)ran1_float(myIy) * ((double)n * 1.0)) - 1u;     // rf.c:531
  }
}

// K1b: in-bag multiplicity per INTERNAL row (u16 packed two per word, native 32-bit atomics)
__global__ void bag_count_kernel(const uint32_t* draws, size_t dstride, const int* bootSize, const uint32_t* inv,
                                 uint32_t* cnt32, size_t cstride32) {
  const int t = blockIdx.y;
  const int bs = bootSize[t];
  const uint32_t* dr = draws + (size_t)t * dstride;
  uint32_t* c = cnt32 + (size_t)t * cstride32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < bs; i += gridDim.x * blockDim.x) {
    uint32_t ii = inv[dr[i]];
    atomicAdd(&c[ii >> 1], 1u << (16u * (ii & 1u)));
  }
}

// by.user bootstrap (rf.c:520-526): the caller supplies the in-bag counts per original row
// (blockIdx.y = tree of the uploaded chunk: samp [chunk][n], cnt [chunk][cntStride])
__global__ void user_count_kernel(const int32_t* samp, const uint32_t* perm, uint16_t* cnt, size_t cntStride, int n, int* bad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    int32_t v = samp[(size_t)blockIdx.y * n + perm[i]];
    if (v < 0 || v > 65535) { atomicExch(bad, 1); v = 0; }
    cnt[(size_t)blockIdx.y * cntStride + i] = (uint16_t)v;
  }
}

// compaction: pool chunks -> flat pre-order arrays at the tree's offset
__global__ void compact_kernel(GrowWork w, DevData d, const int* treeIds, const unsigned long long* nodeOff,
                               const unsigned long long* leafOff,
                               int32_t* treeID, int32_t* nodeID, int32_t* parmID, double* contPT, double* splitStat,
                               uint4* rec4, int32_t* tnRCNT, int32_t* tnCLAS, int32_t* mwcpWord) {
  const int t = blockIdx.y;
  const uint32_t cntN = w.nodeCount[t];
  const unsigned long long off = nodeOff[t], loff = leafOff[t];
  const double na = __longlong_as_double(0x7FF00000000007A2LL);     // R's NA_real_ (payload 1954)
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < cntN; q += gridDim.x * blockDim.x) {
    uint32_t c = w.ctab[(size_t)t * w.maxChunksPerTree + q / kRecChunk];
    size_t at = (size_t)c * kRecChunk + (q % kRecChunk);
    Rec r = w.pool[at];
    treeID[off + q] = treeIds[t];
    nodeID[off + q] = (int32_t)r.nodeID;
    const uint32_t cov1 = r.parm & ~kFactorFlag;                   // covariate + 1
    parmID[off + q] = (int32_t)cov1;
    splitStat[off + q] = w.statPool[at];
    // the traversal kernels read ONE 16-byte record per visited node: {covariate + 1 (| factor flag), cut code or
    // LEFT mask over codes, right child, node id}
    rec4[off + q] = r.parm ? make_uint4(r.parm, r.a, r.b, r.nodeID) : make_uint4(0u, 0u, 0u, r.nodeID);
    if (mwcpWord) mwcpWord[off + q] = 0;
    if (r.parm & kFactorFlag) {
      // wire format of a factor split (saveTree rf.c:10773-10779): contPT = NA, one MWCP word whose bit (level - 1) is
      // set for the levels that go LEFT (convertRelToAbsBinaryPair rf.c:15303-15330); codes -> absolute levels here
      contPT[off + q] = na;
      uint32_t word = 0u, mk = r.a;
      while (mk) { const int cd = __ffs(mk) - 1; mk &= mk - 1u; word |= 1u << ((uint32_t)d.uniq[cov1 - 1][cd] - 1u); }
      mwcpWord[off + q] = (int32_t)word;
    } else if (r.parm) {
      contPT[off + q] = d.uniq[r.parm - 1][r.a];
    } else {
      contPT[off + q] = na;
      tnRCNT[loff + r.nodeID - 1] = (int32_t)r.b;
      tnCLAS[2 * (loff + r.nodeID - 1)] = (int32_t)(r.b - r.a);
      tnCLAS[2 * (loff + r.nodeID - 1) + 1] = (int32_t)r.a;
    }
  }
}

// K4/K5: one thread per internal row walks every tree of the batch (rf.c:10610 routing on codes:
// LEFT iff code <= cut code  <=>  cut - x >= 0), accumulating the class proportions of the leaf
// (rf.c:928-962) into the full ensemble and, when the row is out of bag, the OOB ensemble.
__global__ void membership_kernel(DevData d, int ntrees, const unsigned long long* nodeOff, const unsigned long long* leafOff,
                                  const uint4* __restrict__ rec4,
                                  const int32_t* tnRCNT, const int32_t* tnCLAS, const uint16_t* cnt,
                                  int32_t* tnACNT, double* allNum, double* oobNum, uint32_t* allDen, uint32_t* oobDen,
                                  int32_t* nodeMembership, int32_t* bootMembership, int batchFirst, size_t cntStride,
                                  unsigned long long* pathSum) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long steps = 0;
  if (i < d.n) {
  const uint32_t orig = d.perm[i];
  double a1 = 0.0, a2 = 0.0, o1 = 0.0, o2 = 0.0; uint32_t od = 0;
  for (int t = 0; t < ntrees; t++) {
    const unsigned long long off = nodeOff[t];
    unsigned long long q = off;
    uint4 r = rec4[q];
    while (r.x != 0u) {
      const uint32_t cv = (r.x & ~kFactorFlag) - 1u;
      const uint32_t cd = load_code(d.codes[cv], d.width[cv], (uint32_t)i);
      const bool left = (r.x & kFactorFlag) ? (((r.y >> cd) & 1u) != 0u) : (cd <= r.y);     // splitOnFactor rf.c:1607-1616
      q = left ? q + 1 : off + r.z;
      r = rec4[q];
      steps++;
    }
    const int leaf = (int)r.w;
    const unsigned long long li = leafOff[t] + (unsigned long long)leaf - 1ull;
    const double rc = (double)tnRCNT[li];
    const double p1 = (double)tnCLAS[2 * li] / rc, p2 = (double)tnCLAS[2 * li + 1] / rc;
    a1 += p1; a2 += p2;
    const uint16_t c = cnt[(size_t)t * cntStride + i];
    if (c == 0) { o1 += p1; o2 += p2; od++; }
    atomicAdd(&tnACNT[li], 1);
    if (nodeMembership) nodeMembership[(size_t)(batchFirst + t) * d.n + orig] = leaf;
    if (bootMembership) bootMembership[(size_t)(batchFirst + t) * d.n + orig] = (int32_t)c;
  }
  allNum[orig] += a1; allNum[(size_t)d.n + orig] += a2; allDen[orig] += (uint32_t)ntrees;
  oobNum[orig] += o1; oobNum[(size_t)d.n + orig] += o2; oobDen[orig] += od;
  }
  // every internal node on a row's path routes that row once: sum of path lengths = sum over split nodes of their all-row size
  steps = warp_sum(steps);
  if ((threadIdx.x & 31) == 0) atomicAdd(pathSum, steps);
}

// ---- rmbrMembership / ambrMembership (rf.c:797-803, rf.c:7602-7607): row ids grouped by leaf in the
// order leaves are reached by the DFS, draw order (in-bag) / ascending row order (all rows) inside a leaf.
// Built after the fact from node membership: leaf rank = position of the terminal record in pre-order,
// then a stable sort of the draws / of all rows by the rank of their leaf.
__global__ void leaf_flag_kernel(const int32_t* parmID, uint32_t* flag, size_t NN) {
  size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < NN) flag[q] = parmID[q] == 0 ? 1u : 0u;
}
__global__ void leaf_rank_kernel(const int32_t* parmID, const int32_t* nodeID, const uint32_t* exscan, const unsigned long long* nodeOff,
                                 const unsigned long long* leafOff, int ntrees, uint32_t* leafRank) {
  const int t = blockIdx.y;
  const unsigned long long off = nodeOff[t], end = nodeOff[t + 1];
  for (unsigned long long q = off + blockIdx.x * blockDim.x + threadIdx.x; q < end; q += (unsigned long long)gridDim.x * blockDim.x)
    if (parmID[q] == 0) leafRank[leafOff[t] + (unsigned long long)nodeID[q] - 1ull] = exscan[q] - exscan[off];
  (void)ntrees;
}
// keys for the stable sort: rank of the leaf that original row r fell into (vals = r + 1, the 1-based id R sees)
__global__ void member_keys_kernel(const uint32_t* rows, int count, const int32_t* nodeMembT, const uint32_t* leafRankT,
                                   uint32_t* keys, uint32_t* vals) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < count) {
    const uint32_t r = rows ? rows[j] : (uint32_t)j;
    keys[j] = leafRankT[nodeMembT[r] - 1];
    vals[j] = r + 1u;
  }
}
// by.user bootstrap: the in-bag list is every row repeated count times, ascending (rf.c:520-526)
__global__ void expand_counts_kernel(const int32_t* cnt, const uint32_t* exscan, int n, uint32_t* rows) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const uint32_t o = exscan[i]; for (int k = 0; k < cnt[i]; k++) rows[o + k] = (uint32_t)i; }
}

__global__ void finalize_kernel(double* num, const uint32_t* den, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    uint32_t dd = den[i];
    if (dd) { num[i] /= (double)dd; num[(size_t)n + i] /= (double)dd; }     // rf.c:8115-8121
  }
}

// device scratch; with a pool (the resident dataset's) buffers are recycled across grow calls
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t count = 0;
  WsPool* pool = nullptr;
  DevBuf() {}
  explicit DevBuf(WsPool* pl) : pool(pl) {}
  void alloc(size_t c) {
    count = c;
    size_t bytes = std::max<size_t>(c, 1) * sizeof(T);
    if (pool) { p = (T*)pool->get(bytes); if (!p) { set_error("out of device memory (%zu bytes)", bytes); throw CudaFail{RFSLAM_ENOMEM}; } }
    else RF_CUDA(cudaMalloc((void**)&p, bytes));
  }
  void zero() { RF_CUDA(cudaMemsetAsync(p, 0, std::max<size_t>(count, 1) * sizeof(T))); }
  ~DevBuf() { if (p) { if (pool) pool->put(p); else cudaFree(p); } }
};

template <typename T>
T* host_copy(const T* dev, size_t count) {
  T* h = (T*)host_alloc(std::max<size_t>(count, 1) * sizeof(T));
  if (!h) { set_error("out of host memory"); throw CudaFail{RFSLAM_ENOMEM}; }
  if (count) RF_CUDA(cudaMemcpy(h, dev, count * sizeof(T), cudaMemcpyDeviceToHost));
  return h;
}

}  // namespace

void launch_bootstrap_draw(const int* d_seedA, const int* d_bootSize, uint32_t* d_draws, size_t stride, int n, int ntrees) {
  bootstrap_draw_kernel<<<ntrees, 32>>>(d_seedA, d_bootSize, d_draws, stride, n);
}

}  // namespace rfslam
#include "grow_variant.h"
RFSLAM_DEFINE_VARIANT(t256, rfslam_grow_kernel)
namespace rfslam {

int validate_grow_args(const rfslam_grow_args* a, GrowConfig* cfg) {
  if (!a) { set_error("null arguments"); return RFSLAM_EINVAL; }
  if (a->seed >= 0) { set_error("Random seed must be less than zero (rf.c:6454)"); return RFSLAM_EINVAL; }
  if (a->ntree < 1) { set_error("ntree must be >= 1"); return RFSLAM_EINVAL; }
  if (a->y_size != 1) { set_error("the Poisson class path takes exactly one response (ySize = %d)", a->y_size); return RFSLAM_ENOSUP; }
  if (a->r_type && !(a->r_type[0] == 'C' || a->r_type[0] == 'I' || a->r_type[0] == 'B')) {
    set_error("response must be a factor (rType '%c'); the regression slots of the custom rules abort in the reference too (sc.c:785-805)", a->r_type[0]);
    return RFSLAM_ENOSUP;
  }
  if (a->split_rule != 16) { set_error("splitRule %d is not the custom rule (16): no device implementation", a->split_rule); return RFSLAM_ENOSUP; }
  int slot = (int)((a->opt_high & RFSLAM_OPTH_SPLT_CUST) >> 8) + 1;
  int rule = rfslam_b200_rule_of_slot(RFSLAM_CLAS_FAM, slot);
  if (rule < 1) {
    if (rule_slot_is_foreign(RFSLAM_CLAS_FAM, slot)) set_error("custom split slot %d holds a user-supplied CPU function (registerThis): it has no device implementation and there is no CPU fallback", slot);
    else set_error("custom split slot %d has no device rule (only the six Poisson rules are accelerated; no CPU fallback)", slot);
    return RFSLAM_ENOSUP;
  }
  if (a->nsplit < 0) { set_error("nsplit = %d must be >= 0 (R/data.utilities.R:483-488)", a->nsplit); return RFSLAM_EINVAL; }
  cfg->nsplit = a->nsplit;
  if (a->htry != 0) { set_error("htry must be 0 (rf.c:6490-6494)"); return RFSLAM_EINVAL; }
  if (a->mtry < 1 || a->mtry > a->x_size) { set_error("mtry must be in [1, xSize]"); return RFSLAM_EINVAL; }
  if (a->mtry > kMaxCand) { set_error("mtry %d exceeds the %d candidates scored jointly per node", a->mtry, kMaxCand); return RFSLAM_ENOSUP; }
  if (a->x_size > 1024) { set_error("xSize %d exceeds 1024 covariates", a->x_size); return RFSLAM_ENOSUP; }
  if (a->node_size < 1) { set_error("nodeSize must be >= 1"); return RFSLAM_EINVAL; }
  if (a->n_impute > 1) { set_error("multiple imputation is outside the accelerated path"); return RFSLAM_ENOSUP; }
  // Variable categories (rf.c:14887-15014) only bound how many FAILED draws one call of
  // selectRandomCovariates may make: ceil(mtry * weight) per category, categories tried in turn.
  // Which covariate is drawn never depends on them: the sampler is chosen once from the incoming
  // xWeight (uniform -> RF_WGHT_UNIFORM, rf.c:19110-19114) and the uniform sampler ignores the
  // per-category RF_xWeight rewrites (rf.c:8378-8381).  As long as the per-call budget cannot bind
  // before mtry does, the draws are those of the single-category path.
  if (a->n_categories > 1 || (a->n_categories == 1 && a->category_weights && a->category_weights[0] != 1.0)) {
    if (!a->category_weights) { set_error("categoryWeights missing"); return RFSLAM_EINVAL; }
    long budget = 0;
    for (int k = 0; k < a->n_categories; k++) budget += (long)ceil((double)a->mtry * a->category_weights[k]);
    if (budget < a->mtry) {
      set_error("category weights give a per-call draw budget of %ld < mtry = %d: that truncation is outside the accelerated path", budget, a->mtry);
      return RFSLAM_ENOSUP;
    }
  }
  if (a->n_to_fix > 0) { set_error("fixed_during_bootstrap covariates are outside the accelerated path"); return RFSLAM_ENOSUP; }
  if (a->opt_high & 0x1000u) { set_error("sampling without replacement is outside the accelerated path"); return RFSLAM_ENOSUP; }
  bool t1 = (a->opt_low & RFSLAM_OPT_BOOT_TYP1) != 0, t2 = (a->opt_low & RFSLAM_OPT_BOOT_TYP2) != 0;
  if (t1 != t2) { set_error("bootstrap must be by.root or by.user"); return RFSLAM_ENOSUP; }
  if (a->case_weight) {
    for (int i = 1; i < a->observation_size; i++)
      if (a->case_weight[i] != a->case_weight[0]) { set_error("non-uniform case weights are outside the accelerated path"); return RFSLAM_ENOSUP; }
  }
  if (a->x_weight) {
    for (int i = 1; i < a->x_size; i++)
      if (a->x_weight[i] != a->x_weight[0]) { set_error("non-uniform covariate weights are outside the accelerated path"); return RFSLAM_ENOSUP; }
  }
  if (!(a->k_for_alpha > 0.0)) { set_error("k_for_alpha must be > 0"); return RFSLAM_EINVAL; }
  if (!a->r_data || !a->x_data || !a->risk_time || !a->stratifier) { set_error("null data pointer"); return RFSLAM_EINVAL; }
  cfg->rule = rule; cfg->mtry = a->mtry; cfg->nodesize = a->node_size; cfg->nodedepth = a->node_depth;
  cfg->alpha = 1.0 / (a->k_for_alpha * a->k_for_alpha);
  cfg->ntree = a->ntree; cfg->seed = a->seed; cfg->by_user = t1 && t2;
  cfg->want_membership = (a->opt_high & RFSLAM_OPTH_MEMB_USER) != 0;
  cfg->want_membr_lists = (a->opt_high & RFSLAM_OPTH_MEMB_OUTG) != 0 && a->want_member_lists != 0;
  cfg->finalize = a->finalize_ensembles != 0;
  if (cfg->by_user && !a->bootstrap) { set_error("bootstrap = by.user needs the samp matrix"); return RFSLAM_EINVAL; }
  if (a->comm && a->gather_forest && (cfg->want_membership || cfg->want_membr_lists)) {
    set_error("per-row membership outputs stay with the shard that grew the tree: ask for them without gather_forest");
    return RFSLAM_ENOSUP;
  }
  return RFSLAM_OK;
}

namespace {
// growable host array handed to the caller (pool block, see hostpool.cu; released by rfslam_b200_free_forest)
template <typename T>
struct HostOut {
  T* p = nullptr; size_t n = 0, cap = 0;
  void reserve(size_t want) {
    if (want <= cap) return;
    const size_t ncap = std::max<size_t>(want, cap * 2);
    T* q = (T*)host_alloc(std::max<size_t>(ncap, 1) * sizeof(T));
    if (!q) { set_error("out of host memory"); throw CudaFail{RFSLAM_ENOMEM}; }
    if (n) memcpy(q, p, n * sizeof(T));
    host_free(p);
    p = q; cap = ncap;
  }
  T* grow(size_t add) { reserve(n + add); T* at = p + n; n += add; return at; }
  T* release() { if (!p) p = (T*)host_alloc(sizeof(T)); T* q = p; p = nullptr; return q; }
  ~HostOut() { host_free(p); }
};
}  // namespace

int grow_forest(rfslam_dataset* ds, const rfslam_grow_args* a, rfslam_forest_out* out) {
  memset(out, 0, sizeof(*out));
  const bool tmgAll = getenv("RFSLAM_TIMING") != nullptr;
  auto wallAll = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double tw0 = wallAll(); double twBatches = 0, twHand = 0;
  GrowConfig cfg;
  int rc = validate_grow_args(a, &cfg);
  if (rc) return rc;
  if (a->observation_size != ds->n || a->x_size != ds->p) { set_error("arguments do not match the resident dataset"); return RFSLAM_EINVAL; }
  const int n = ds->n, p = ds->p;
  WsPool* pool = &ds->pool;
  try {
    RF_CUDA(cudaSetDevice(ds->device));
    // ---- which trees does this process grow? (tree b -> rank (b-1) mod G) ----
    std::vector<int> ids;
    int first = a->tree_first > 0 ? a->tree_first : 1, step = a->tree_step > 0 ? a->tree_step : 1;
    for (int b = first; b <= cfg.ntree; b += step) ids.push_back(b);
    const int T = (int)ids.size();
    if (T == 0) { set_error("this shard holds no trees (tree_first %d > ntree %d)", first, cfg.ntree); return RFSLAM_EINVAL; }
    std::vector<int> seedA, seedB;
    chain_seeds(cfg.seed, cfg.ntree, seedA, seedB);
    std::vector<int> bootSize(T);
    int maxBoot = 1;
    for (int t = 0; t < T; t++) {
      if (cfg.by_user) {
        long sum = 0;
        const int32_t* sp = a->bootstrap + (size_t)(ids[t] - 1) * n;
        for (int i = 0; i < n; i++) sum += sp[i];
        bootSize[t] = (int)sum;
      } else {
        bootSize[t] = a->bootstrap_size ? a->bootstrap_size[ids[t] - 1] : n;
      }
      if (bootSize[t] < 1) { set_error("empty bootstrap sample for tree %d", ids[t]); return RFSLAM_EINVAL; }
      maxBoot = std::max(maxBoot, bootSize[t]);
    }

    // ---- kernel parameters ----
    GrowParams g{};
    g.rule = cfg.rule; g.mtry = cfg.mtry; g.nodesize = cfg.nodesize; g.nodedepth = cfg.nodedepth; g.alpha = cfg.alpha;
    g.pw = (p + 31) / 32; g.ncmax = std::min(cfg.mtry, kMaxCand);
    g.has_sort = ds->max_D > kHistBins;
    g.hbins = ds->hist_bins;
    g.nsplit = cfg.nsplit;
    g.seqDraw = (cfg.nsplit > 0 || ds->has_factor) ? 1 : 0;
    g.facCuts = g.seqDraw ? 256 : 0;                              // factor cut masks / drawn cut ordinals of wide covariates
    if (cfg.nsplit > 256 && ds->max_D > kHistBins) {
      set_error("nsplit = %d with a covariate of %d distinct values: at most 256 drawn cuts per covariate are kept for covariates of more than %d distinct values", cfg.nsplit, ds->max_D, kHistBins);
      return RFSLAM_ENOSUP;
    }
    {   // fixed-point scale of the exposure accumulators: the largest power of two that keeps
        // (bootstrap size) * (max risk time) below 2^62
      int ex = 0;
      frexp(std::max((double)maxBoot * std::max(ds->max_rt, 1e-300), 1.0), &ex);
      g.tscale = ldexp(1.0, 62 - ex); g.tinv = ldexp(1.0, ex - 62);
    }
    DevBuf<double> d_sw(pool);
    if (a->x_split_stat_weight) {
      bool allOne = true;
      for (int i = 0; i < p; i++) if (a->x_split_stat_weight[i] != 1.0) allOne = false;
      if (!allOne) {
        d_sw.alloc(p);
        RF_CUDA(cudaMemcpy(d_sw.p, a->x_split_stat_weight, (size_t)p * 8, cudaMemcpyHostToDevice));
        g.split_weight = d_sw.p;
      }
    }
    // Kernel plan per launch: CTA width (grow_variant.h) and the shared-memory carve-up that goes with it.
    //   * 256 threads, two CTAs per SM: the default;
    //   * 320 threads, two CTAs per SM: mtry of 9 or 10 -- one warp per candidate instead of two scoring rounds;
    //   * 512 threads, one CTA per SM: a launch with no more trees than SMs (a forest sharded over many GPUs), where half
    //     of every SM would idle and the passes over large nodes are the per-tree critical path.
    // Shared memory: the staged-subtree tile (subMax whole rows) is worth more than wide histogram passes, so the strata
    // slots per pass go 4 -> 2 (same entries, same flushes, two more barriers per large node) before the tile shrinks.
    const int tilePitch = (ds->dev.rowPitch + 15) & ~15;
    int sms = 0; long long quantumCycles = 0;
    {
      int dev = 0, khz = 0;
      RF_CUDA(cudaGetDevice(&dev));
      RF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      RF_CUDA(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));                  // kHz = cycles per ms
      const char* q = getenv("RFSLAM_QUANTUM_MS");
      quantumCycles = (long long)((q ? atof(q) : 10.0) * (double)khz);
    }
    struct Plan { const GrowVariant* kv; GrowParams g; size_t smem; int residentCtas; };
    auto make_plan = [&](int trees) -> Plan {
      Plan pl; pl.g = g;
      pl.kv = grow_variant_t256();
      if (!g.has_sort) {                                   // (the sorted-walk path's tile arrays are laid out for 256 threads)
        if (trees <= sms) pl.kv = grow_variant_t512();
        else if (g.ncmax > 8) pl.kv = grow_variant_t320();
      }
      if (const char* e = getenv("RFSLAM_CTA_THREADS")) {
        const int v = atoi(e);
        if (!g.has_sort) pl.kv = v == 512 ? grow_variant_t512() : v == 320 ? grow_variant_t320() : v == 128 ? grow_variant_t128() : grow_variant_t256();
      }
      if (pl.kv == grow_variant_t256() && g.ncmax > 8) pl.kv = grow_variant_t320();   // (the lean 256-thread build's histogram pass holds 8 candidates)
      // tables with unordered factors or covariates of more than 256 distinct values take the FULL build
      const bool wantSubtree = ds->dev.rowcodes && getenv("RFSLAM_SUBTREE_MAX") && !getenv("RFSLAM_NO_SUBTREE");
      const bool full = ds->has_factor || g.has_sort || wantSubtree || getenv("RFSLAM_PROFILE") || getenv("RFSLAM_FULL_KERNEL");
      if (full) pl.kv = grow_variant_t256f();
      const int ctasWanted = pl.kv->ctasPerSm;
      // shared memory available to one CTA's dynamic part (227 KB per SM, 1 KB reserved per CTA, GrowShared is static)
      const size_t limit = (size_t)(227u << 10) / (size_t)ctasWanted - 1024 - sizeof(GrowShared);
      pl.g.hslots = hist_slots(g.hbins);
      pl.g.tabSlots = kTabSlots;
      if (const char* e = getenv("RFSLAM_TAB_SLOTS")) pl.g.tabSlots = std::max(1, std::min(kTabSlots, atoi(e)));
      if (const char* e = getenv("RFSLAM_HIST_SLOTS")) pl.g.hslots = std::max(1, std::min(pl.g.hslots, atoi(e)));
      while (grow_smem_bytes(g.ncmax, ds->K, g.hbins, pl.g.hslots, pl.g.tabSlots, 0, tilePitch, g.facCuts) > limit && (pl.g.hslots > 1 || pl.g.tabSlots > 1)) {
        if (pl.g.hslots * 2 > pl.g.tabSlots) pl.g.hslots >>= 1; else pl.g.tabSlots >>= 1;      // (histogram 32 B / bin, tables 16 B / bin)
      }
      pl.g.subMax = 0;
      if (wantSubtree && !g.seqDraw) {                       // (staged subtrees bought nothing at 10 M rows: opt-in, FULL build)
        const int wantSub = getenv("RFSLAM_SUBTREE_MAX") ? atoi(getenv("RFSLAM_SUBTREE_MAX")) : kSmallMax;
        const int hsFull = pl.g.hslots, hsMin = std::min(2, hsFull);
        for (int sm = std::min(wantSub, kSmallMax); sm >= 128 && pl.g.subMax == 0; sm >>= 1)
          for (int hs = hsFull; hs >= hsMin; hs >>= 1)
            if (grow_smem_bytes(g.ncmax, ds->K, g.hbins, hs, pl.g.tabSlots, sm, tilePitch, g.facCuts) <= limit) { pl.g.hslots = hs; pl.g.subMax = sm; break; }
      }
      pl.smem = grow_smem_bytes(g.ncmax, ds->K, g.hbins, pl.g.hslots, pl.g.tabSlots, pl.g.subMax, tilePitch, g.facCuts);
      int perSm = 0;
      RF_CUDA(pl.kv->prepare(pl.smem, &perSm));
      pl.residentCtas = std::max(1, perSm) * sms;
      if (getenv("RFSLAM_TIMING")) fprintf(stderr, "[rfslam_b200_grow] plan for %d trees: %d threads/CTA, %d CTAs/SM resident, %zu B dynamic smem, hist slots %d, table slots %d, subtree tile %d\n",
                                          trees, pl.kv->threads, perSm, pl.smem, pl.g.hslots, pl.g.tabSlots, pl.g.subMax);
      return pl;
    };

    // ---- batch size from free memory ----
    size_t freeB = 0, totalB = 0;
    RF_CUDA(cudaMemGetInfo(&freeB, &totalB));
    freeB += pool->idle_bytes();
    // entries of a tree: a row drawn c times folds into ceil(c / kMaxMult) <= 1 + c / kMaxMult entries
    const size_t listStride = (size_t)std::min(n, maxBoot) + (size_t)maxBoot / kMaxMult + 1;
    if (!cfg.by_user && maxBoot > 65535) {                         // u16 in-bag counts: refuse samples under which a row could plausibly exceed them
      const double mean = (double)maxBoot / (double)n;
      if (mean + 12.0 * sqrt(mean) + 50.0 > 65535.0) { set_error("sampsize %d over %d rows would overflow the 16-bit in-bag counts", maxBoot, n); return RFSLAM_ENOSUP; }
    }
    const size_t perTree = listStride * (8 + (g.has_sort ? 16 : 0)) + (size_t)n * 2 + (cfg.by_user ? 0 : (size_t)maxBoot * 4)
                         + (size_t)kStackCap * (sizeof(NodeEnt) + (size_t)g.pw * 4) + 4096;
    const size_t worstNodes = 2ull * (size_t)std::min(n, maxBoot) + 1;
    size_t estNodes = std::min<size_t>(worstNodes, 4ull * (size_t)maxBoot / (size_t)std::max(1, cfg.nodesize) + 256);
    size_t budget = (size_t)(0.80 * (double)freeB);
    size_t fixed = (size_t)n * (16 + 16 + 8) + (64u << 20);
    int B = (int)std::max<size_t>(1, std::min<size_t>((size_t)T, (budget > fixed ? budget - fixed : 0) / (perTree + estNodes * 80)));
    if (const char* e = getenv("RFSLAM_TREES_PER_BATCH")) { int v = atoi(e); if (v > 0) B = std::min(T, v); }
    B = (T + (T + B - 1) / B - 1) / ((T + B - 1) / B);             // equal batches: 1000 trees with room for 900 run as 500 + 500, not 900 + 100

    // ---- persistent accumulators ----
    DevBuf<double> d_allNum(pool), d_oobNum(pool); DevBuf<uint32_t> d_allDen(pool), d_oobDen(pool);
    d_allNum.alloc(2 * (size_t)n); d_oobNum.alloc(2 * (size_t)n); d_allDen.alloc(n); d_oobDen.alloc(n);
    d_allNum.zero(); d_oobNum.zero(); d_allDen.zero(); d_oobDen.zero();
    DevBuf<int32_t> d_nodeMemb(pool), d_bootMemb(pool);
    if (cfg.want_membership || cfg.want_membr_lists) { d_nodeMemb.alloc((size_t)T * n); d_bootMemb.alloc((size_t)T * n); }
    DevBuf<int32_t> d_rmbr(pool), d_ambr(pool);
    const size_t rmbrStride = (size_t)std::max(a->max_bootstrap_size, maxBoot);
    if (cfg.want_membr_lists) { d_rmbr.alloc((size_t)T * rmbrStride); d_ambr.alloc((size_t)T * n); d_rmbr.zero(); d_ambr.zero(); }
    DevBuf<unsigned long long> d_pathSum(pool);
    d_pathSum.alloc(1); d_pathSum.zero();

    HostOut<int32_t> h_treeID, h_nodeID, h_parmID, h_tnRCNT, h_tnACNT, h_tnCLAS, h_mwcpW;
    HostOut<double> h_contPT, h_stat;
    std::vector<int32_t> h_leafCount(T), h_seed(T);
    long long totalCand = 0, totalAlgo = 0, launches = 0;
    float growMs = 0.f, totalMs = 0.f, membMs = 0.f, bootMs = 0.f;
    long long bootDraws = 0;
    rfslam_comm* comm = (a->comm && comm_world(a->comm) > 1) ? a->comm : nullptr;
    const bool gather = comm && a->gather_forest != 0;
    // gathering: the shard's records stay on the device (appended batch by batch) instead of going home
    struct Acc {
      WsPool* pool; void* p = nullptr; size_t used = 0, cap = 0;
      explicit Acc(WsPool* pl) : pool(pl) {}
      ~Acc() { if (p) pool->put(p); }
      void append(const void* src, size_t bytes, size_t hint) {
        if (used + bytes > cap) {
          const size_t ncap = std::max<size_t>({used + bytes, 2 * cap, hint});
          void* q = pool->get(ncap);
          if (!q) { set_error("out of device memory (%zu bytes)", ncap); throw CudaFail{RFSLAM_ENOMEM}; }
          if (used) RF_CUDA(cudaMemcpyAsync(q, p, used, cudaMemcpyDeviceToDevice));
          RF_CUDA(cudaDeviceSynchronize());
          if (p) pool->put(p);
          p = q; cap = ncap;
        }
        if (bytes) RF_CUDA(cudaMemcpyAsync((char*)p + used, src, bytes, cudaMemcpyDeviceToDevice));
        used += bytes;
      }
    };
    Acc g_tree(pool), g_node(pool), g_parm(pool), g_cont(pool), g_stat(pool), g_rcnt(pool), g_acnt(pool), g_clas(pool), g_mwcp(pool);
    // events and the copy stream are released on every exit path; the stream is drained first, so no copy can
    // still target a page-locked output block when the HostOut destructors (declared above) hand it back
    struct Sync {
      cudaEvent_t e[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}; cudaStream_t st = nullptr;
      ~Sync() { if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); } for (auto x : e) if (x) cudaEventDestroy(x); }
    } sync;
    for (int k = 0; k < 4; k++) RF_CUDA(cudaEventCreate(&sync.e[k]));
    RF_CUDA(cudaEventCreateWithFlags(&sync.e[4], cudaEventDisableTiming));
    RF_CUDA(cudaStreamCreateWithFlags(&sync.st, cudaStreamNonBlocking));   // forest records go home while the traversal kernel runs
    cudaEvent_t ev0 = sync.e[0], ev1 = sync.e[1], ev2 = sync.e[2], ev3 = sync.e[3], evC = sync.e[4];
    struct Ev2 { cudaEvent_t a = nullptr, b = nullptr; ~Ev2() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); } } evM;
    RF_CUDA(cudaEventCreate(&evM.a)); RF_CUDA(cudaEventCreate(&evM.b));
    cudaStream_t copyStream = sync.st;
    bool mwcpCleared = false;
    const bool wantStat = a->want_split_stat != 0;

    // ---- per-batch buffers ----
    DevBuf<uint32_t> d_lists(pool), d_sort(pool), d_draws(pool), d_stackPerm(pool), d_ctab(pool), d_nodeCount(pool), d_leafCount(pool), d_cursor(pool);
    DevBuf<uint16_t> d_cnt(pool); DevBuf<NodeEnt> d_stack(pool); DevBuf<int> d_seedA(pool), d_seedB(pool), d_boot(pool), d_ids(pool), d_status(pool);
    DevBuf<unsigned long long> d_cand(pool), d_algo(pool), d_nodeOff(pool), d_leafOff(pool);
    DevBuf<int32_t> d_samp(pool);
    const size_t cntStride = ((size_t)n + 7) & ~(size_t)7;         // u16 count rows stay 16-byte aligned (the root build reads eight at a time)
    d_lists.alloc((size_t)B * 2 * listStride);
    if (g.has_sort) d_sort.alloc((size_t)B * 4 * listStride);
    // by.user: the caller's in-bag counts go up in chunks of whole trees (one copy + one launch per chunk, not per tree)
    const int sampChunk = cfg.by_user ? (int)std::max<size_t>(1, std::min<size_t>((size_t)B, ((size_t)256 << 20) / ((size_t)n * 4))) : 0;
    if (!cfg.by_user) d_draws.alloc((size_t)B * maxBoot); else d_samp.alloc((size_t)sampChunk * n);
    d_cnt.alloc((size_t)B * cntStride);
    d_stack.alloc((size_t)B * kStackCap); d_stackPerm.alloc((size_t)B * kStackCap * g.pw);
    d_nodeCount.alloc(B); d_leafCount.alloc(B); d_cursor.alloc(1); d_status.alloc(1);
    d_seedA.alloc(B); d_seedB.alloc(B); d_boot.alloc(B); d_ids.alloc(B);
    d_cand.alloc(B); d_algo.alloc(B); d_nodeOff.alloc(B + 1); d_leafOff.alloc(B + 1);
    const uint32_t maxChunksPerTree = (uint32_t)((worstNodes + kRecChunk - 1) / kRecChunk) + 1;
    d_ctab.alloc((size_t)B * maxChunksPerTree);
    const bool profile = getenv("RFSLAM_PROFILE") != nullptr;
    DevBuf<unsigned long long> d_prof(pool);
    if (profile) { d_prof.alloc((size_t)B * kProfSlots); g_last_profile.assign(kProfSlots, 0ull); }

    for (int b0 = 0; b0 < T; b0 += B) {
      const int nb = std::min(B, T - b0);
      std::vector<int> hsA(nb), hsB(nb), hbs(nb), hid(nb);
      for (int t = 0; t < nb; t++) { hsA[t] = seedA[ids[b0 + t] - 1]; hsB[t] = seedB[ids[b0 + t] - 1]; hbs[t] = bootSize[b0 + t]; hid[t] = ids[b0 + t]; h_seed[b0 + t] = hsA[t]; }
      RF_CUDA(cudaMemcpy(d_seedA.p, hsA.data(), nb * 4, cudaMemcpyHostToDevice));
      RF_CUDA(cudaMemcpy(d_seedB.p, hsB.data(), nb * 4, cudaMemcpyHostToDevice));
      RF_CUDA(cudaMemcpy(d_boot.p, hbs.data(), nb * 4, cudaMemcpyHostToDevice));
      RF_CUDA(cudaMemcpy(d_ids.p, hid.data(), nb * 4, cudaMemcpyHostToDevice));

      size_t poolChunks = std::max<size_t>((size_t)nb * 2, (estNodes * nb + kRecChunk - 1) / kRecChunk + nb);
      for (int attempt = 0;; attempt++) {
        DevBuf<Rec> d_pool(pool); DevBuf<double> d_statPool(pool);
        d_pool.alloc(poolChunks * kRecChunk); d_statPool.alloc(poolChunks * kRecChunk);
        RF_CUDA(cudaEventRecord(ev0));
        // K1: bootstrap
        d_cnt.zero();
        if (cfg.by_user) {
          int* d_bad = d_status.p;
          RF_CUDA(cudaMemset(d_bad, 0, 4));
          for (int t0 = 0; t0 < nb; ) {
            // trees of a shard are tree_step apart in the caller's matrix: consecutive ones (tree_step = 1) go up in one copy
            int cnt = 1;
            while (t0 + cnt < nb && cnt < sampChunk && hid[t0 + cnt] == hid[t0 + cnt - 1] + 1) cnt++;
            RF_CUDA(cudaMemcpy(d_samp.p, a->bootstrap + (size_t)(hid[t0] - 1) * n, (size_t)cnt * n * 4, cudaMemcpyHostToDevice));
            dim3 gu((unsigned)((n + 255) / 256), (unsigned)cnt);
            user_count_kernel<<<gu, 256>>>(d_samp.p, ds->dev.perm, d_cnt.p + (size_t)t0 * cntStride, cntStride, n, d_bad);
            launches++;
            t0 += cnt;
          }
          int bad = 0;
          RF_CUDA(cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost));
          if (bad) { set_error("by.user in-bag counts must be in [0, 65535]"); throw CudaFail{RFSLAM_EINVAL}; }
        } else {
          bootstrap_draw_kernel<<<nb, 32>>>(d_seedA.p, d_boot.p, d_draws.p, (size_t)maxBoot, n);
          dim3 gc((unsigned)std::min<size_t>(((size_t)maxBoot + 255) / 256, 1024), (unsigned)nb);
          bag_count_kernel<<<gc, 256>>>(d_draws.p, (size_t)maxBoot, d_boot.p, ds->dev.inv, (uint32_t*)d_cnt.p, cntStride / 2);
          launches += 2;
        }
        RF_CUDA(cudaGetLastError());
        // K2: grow
        RF_CUDA(cudaMemsetAsync(d_cursor.p, 0, 4)); RF_CUDA(cudaMemsetAsync(d_status.p, 0, 4));
        GrowWork w{};
        w.ntrees = nb; w.seedB = d_seedB.p; w.bootSize = d_boot.p; w.lists = d_lists.p; w.listStride = listStride; w.cnt = d_cnt.p; w.cntStride = cntStride;
        w.sortbuf = g.has_sort ? d_sort.p : nullptr; w.stack = d_stack.p; w.stackPerm = d_stackPerm.p;
        w.pool = d_pool.p; w.statPool = d_statPool.p; w.poolChunks = (uint32_t)poolChunks; w.poolCursor = d_cursor.p;
        w.ctab = d_ctab.p; w.maxChunksPerTree = maxChunksPerTree; w.nodeCount = d_nodeCount.p; w.leafCount = d_leafCount.p;
        w.candCount = d_cand.p; w.algoBytes = d_algo.p; w.status = d_status.p;
        w.prof = profile ? d_prof.p : nullptr;
        // more trees than resident CTAs: the grid is the resident CTAs and trees are time-sliced
        const Plan pl = make_plan(nb);
        int gridTrees = nb;
        DevBuf<int> d_tstate(pool); DevBuf<unsigned char> d_saved(pool);
        if (nb > pl.residentCtas && !getenv("RFSLAM_NO_SLICING")) {
          d_tstate.alloc((size_t)nb + 2); d_tstate.zero();
          d_saved.alloc((size_t)nb * sizeof(GrowShared));
          RF_CUDA(cudaMemcpyAsync(d_tstate.p + nb, &nb, 4, cudaMemcpyHostToDevice));     // waiting = nb; ticket = 0
          w.treeState = d_tstate.p; w.waiting = d_tstate.p + nb; w.ticket = (unsigned*)(d_tstate.p + nb + 1);
          w.saved = d_saved.p; w.quantum = quantumCycles;
          gridTrees = pl.residentCtas;
        }
        RF_CUDA(cudaEventRecord(ev1));
        pl.kv->launch(ds->dev, pl.g, w, gridTrees, pl.smem);
        launches++;
        RF_CUDA(cudaEventRecord(ev2));
        RF_CUDA(cudaGetLastError());
        int status = 0;
        RF_CUDA(cudaMemcpy(&status, d_status.p, 4, cudaMemcpyDeviceToHost));
        if (status == 1) {
          if (attempt > 8) { set_error("record pool kept overflowing"); throw CudaFail{RFSLAM_ENOMEM}; }
          poolChunks *= 2;
          continue;                                            // deterministic: simply regrow the batch
        }
        if (status == 2) { set_error("tree deeper than %d pending nodes", kStackCap); throw CudaFail{RFSLAM_ENOSUP}; }
        if (profile) {
          std::vector<unsigned long long> hp((size_t)nb * kProfSlots);
          RF_CUDA(cudaMemcpy(hp.data(), d_prof.p, hp.size() * 8, cudaMemcpyDeviceToHost));
          for (int t = 0; t < nb; t++) for (int k = 0; k < kProfSlots; k++) if (k != 12 && k != 13) g_last_profile[k] += hp[(size_t)t * kProfSlots + k];
        }
        // ---- offsets, compaction, traversal ----
        std::vector<uint32_t> hNodes(nb), hLeaves(nb);
        std::vector<unsigned long long> hCand(nb), hAlgo(nb), nodeOff(nb + 1, 0), leafOff(nb + 1, 0);
        RF_CUDA(cudaMemcpy(hNodes.data(), d_nodeCount.p, nb * 4, cudaMemcpyDeviceToHost));
        RF_CUDA(cudaMemcpy(hLeaves.data(), d_leafCount.p, nb * 4, cudaMemcpyDeviceToHost));
        RF_CUDA(cudaMemcpy(hCand.data(), d_cand.p, nb * 8, cudaMemcpyDeviceToHost));
        RF_CUDA(cudaMemcpy(hAlgo.data(), d_algo.p, nb * 8, cudaMemcpyDeviceToHost));
        for (int t = 0; t < nb; t++) {
          nodeOff[t + 1] = nodeOff[t] + hNodes[t]; leafOff[t + 1] = leafOff[t] + hLeaves[t];
          totalCand += (long long)hCand[t]; totalAlgo += (long long)hAlgo[t]; bootDraws += (long long)hbs[t];
          h_leafCount[b0 + t] = (int32_t)hLeaves[t];
          if (hNodes[t] != 2 * hLeaves[t] - 1) { set_error("internal: tree %d has %u records for %u leaves", hid[t], hNodes[t], hLeaves[t]); throw CudaFail{RFSLAM_EINVAL}; }
        }
        RF_CUDA(cudaMemcpy(d_nodeOff.p, nodeOff.data(), (nb + 1) * 8, cudaMemcpyHostToDevice));
        RF_CUDA(cudaMemcpy(d_leafOff.p, leafOff.data(), (nb + 1) * 8, cudaMemcpyHostToDevice));
        const size_t NN = nodeOff[nb], NL = leafOff[nb];
        DevBuf<int32_t> f_tree(pool), f_node(pool), f_parm(pool), f_rcnt(pool), f_acnt(pool), f_clas(pool);
        DevBuf<double> f_cont(pool), f_stat(pool); DevBuf<uint4> f_rec4(pool);
        f_tree.alloc(NN); f_node.alloc(NN); f_parm.alloc(NN); f_cont.alloc(NN); f_stat.alloc(NN); f_rec4.alloc(NN);
        DevBuf<int32_t> f_mwcp(pool);
        if (ds->has_factor) f_mwcp.alloc(NN);
        f_rcnt.alloc(NL); f_acnt.alloc(NL); f_clas.alloc(2 * NL); f_acnt.zero();
        uint32_t maxNodes = *std::max_element(hNodes.begin(), hNodes.end());
        dim3 gk((unsigned)std::min<uint32_t>((maxNodes + 255) / 256, 2048), (unsigned)nb);
        compact_kernel<<<gk, 256>>>(w, ds->dev, d_ids.p, d_nodeOff.p, d_leafOff.p, f_tree.p, f_node.p, f_parm.p, f_cont.p, f_stat.p,
                                    f_rec4.p, f_rcnt.p, f_clas.p, ds->has_factor ? f_mwcp.p : nullptr);
        RF_CUDA(cudaEventRecord(evC));
        RF_CUDA(cudaEventRecord(evM.a));
        membership_kernel<<<(n + 127) / 128, 128>>>(ds->dev, nb, d_nodeOff.p, d_leafOff.p, f_rec4.p,
                                                    f_rcnt.p, f_clas.p, d_cnt.p, f_acnt.p, d_allNum.p, d_oobNum.p, d_allDen.p, d_oobDen.p,
                                                    d_nodeMemb.p, d_bootMemb.p, b0, cntStride, d_pathSum.p);
        launches += 2;
        RF_CUDA(cudaEventRecord(evM.b));
        RF_CUDA(cudaEventRecord(ev3));
        RF_CUDA(cudaGetLastError());
        if (gather) {
          const size_t hintN = (size_t)(estNodes * T), hintL = hintN / 2 + T;
          g_tree.append(f_tree.p, NN * 4, hintN * 4); g_node.append(f_node.p, NN * 4, hintN * 4); g_parm.append(f_parm.p, NN * 4, hintN * 4);
          g_cont.append(f_cont.p, NN * 8, hintN * 8);
          if (ds->has_factor) g_mwcp.append(f_mwcp.p, NN * 4, hintN * 4);
          if (wantStat) g_stat.append(f_stat.p, NN * 8, hintN * 8);
          g_rcnt.append(f_rcnt.p, NL * 4, hintL * 4); g_acnt.append(f_acnt.p, NL * 4, hintL * 4); g_clas.append(f_clas.p, 2 * NL * 4, hintL * 8);
          RF_CUDA(cudaEventSynchronize(ev3));
          RF_CUDA(cudaDeviceSynchronize());
        } else {
        // ---- straight into the caller's arrays: the forest records are final after compaction, so
        // their copies overlap the traversal kernel on a second stream ----
        RF_CUDA(cudaStreamWaitEvent(copyStream, evC, 0));
        RF_CUDA(cudaMemcpyAsync(h_treeID.grow(NN), f_tree.p, NN * 4, cudaMemcpyDeviceToHost, copyStream));
        RF_CUDA(cudaMemcpyAsync(h_nodeID.grow(NN), f_node.p, NN * 4, cudaMemcpyDeviceToHost, copyStream));
        RF_CUDA(cudaMemcpyAsync(h_parmID.grow(NN), f_parm.p, NN * 4, cudaMemcpyDeviceToHost, copyStream));
        RF_CUDA(cudaMemcpyAsync(h_contPT.grow(NN), f_cont.p, NN * 8, cudaMemcpyDeviceToHost, copyStream));
        if (ds->has_factor) RF_CUDA(cudaMemcpyAsync(h_mwcpW.grow(NN), f_mwcp.p, NN * 4, cudaMemcpyDeviceToHost, copyStream));
        if (wantStat) RF_CUDA(cudaMemcpyAsync(h_stat.grow(NN), f_stat.p, NN * 8, cudaMemcpyDeviceToHost, copyStream));
        RF_CUDA(cudaMemcpyAsync(h_tnRCNT.grow(NL), f_rcnt.p, NL * 4, cudaMemcpyDeviceToHost, copyStream));
        RF_CUDA(cudaMemcpyAsync(h_tnCLAS.grow(2 * NL), f_clas.p, 2 * NL * 4, cudaMemcpyDeviceToHost, copyStream));
        RF_CUDA(cudaStreamSynchronize(copyStream));
        RF_CUDA(cudaEventSynchronize(ev3));
        RF_CUDA(cudaMemcpy(h_tnACNT.grow(NL), f_acnt.p, NL * 4, cudaMemcpyDeviceToHost));
        }
        float ms = 0.f;
        RF_CUDA(cudaEventElapsedTime(&ms, ev1, ev2)); growMs += ms;
        RF_CUDA(cudaEventElapsedTime(&ms, ev0, ev3)); totalMs += ms;
        RF_CUDA(cudaEventElapsedTime(&ms, ev0, ev1)); bootMs += ms;
        RF_CUDA(cudaEventElapsedTime(&ms, evM.a, evM.b)); membMs += ms;
        if (cfg.want_membr_lists) {
          DevBuf<uint32_t> d_flag(pool), d_scan(pool), d_lrank(pool), d_k1(pool), d_k2(pool), d_v1(pool), d_v2(pool), d_rows(pool);
          d_flag.alloc(NN); d_scan.alloc(std::max<size_t>(NN, (size_t)n)); d_lrank.alloc(NL);
          const size_t mcap = std::max<size_t>((size_t)n, (size_t)maxBoot);
          d_k1.alloc(mcap); d_k2.alloc(mcap); d_v1.alloc(mcap); d_v2.alloc(mcap);
          if (cfg.by_user) d_rows.alloc(mcap);
          size_t tb = 0, tb2 = 0;
          cub::DeviceScan::ExclusiveSum(nullptr, tb, d_flag.p, d_scan.p, (int)std::max<size_t>(NN, (size_t)n));
          cub::DeviceRadixSort::SortPairs(nullptr, tb2, d_k1.p, d_k2.p, d_v1.p, d_v2.p, (int)mcap);
          DevBuf<unsigned char> d_tmp(pool);
          d_tmp.alloc(std::max(tb, tb2));
          size_t tbytes = std::max(tb, tb2);
          leaf_flag_kernel<<<(unsigned)((NN + 255) / 256), 256>>>(f_parm.p, d_flag.p, NN);
          RF_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp.p, tbytes, d_flag.p, d_scan.p, (int)NN));
          dim3 gl((unsigned)std::min<uint32_t>((maxNodes + 255) / 256, 2048), (unsigned)nb);
          leaf_rank_kernel<<<gl, 256>>>(f_parm.p, f_node.p, d_scan.p, d_nodeOff.p, d_leafOff.p, nb, d_lrank.p);
          for (int t = 0; t < nb; t++) {
            const int32_t* memb = d_nodeMemb.p + (size_t)(b0 + t) * n;
            const uint32_t* lr = d_lrank.p + leafOff[t];
            int bits = 1; while ((1u << bits) < hLeaves[t]) bits++;
            // all rows, ascending row order inside a leaf
            member_keys_kernel<<<(n + 255) / 256, 256>>>(nullptr, n, memb, lr, d_k1.p, d_v1.p);
            tbytes = std::max(tb, tb2);
            RF_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp.p, tbytes, d_k1.p, d_k2.p, d_v1.p, (uint32_t*)(d_ambr.p + (size_t)(b0 + t) * n), n, 0, bits));
            // in-bag rows, draw order inside a leaf
            const uint32_t* rows; const int bs = hbs[t];
            if (cfg.by_user) {
              RF_CUDA(cudaMemcpy(d_samp.p, a->bootstrap + (size_t)(hid[t] - 1) * n, (size_t)n * 4, cudaMemcpyHostToDevice));
              tbytes = std::max(tb, tb2);
              RF_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp.p, tbytes, (const uint32_t*)d_samp.p, d_scan.p, n));
              expand_counts_kernel<<<(n + 255) / 256, 256>>>(d_samp.p, d_scan.p, n, d_rows.p);
              rows = d_rows.p;
            } else {
              rows = d_draws.p + (size_t)t * maxBoot;
            }
            member_keys_kernel<<<(bs + 255) / 256, 256>>>(rows, bs, memb, lr, d_k1.p, d_v1.p);
            tbytes = std::max(tb, tb2);
            RF_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp.p, tbytes, d_k1.p, d_k2.p, d_v1.p, (uint32_t*)(d_rmbr.p + (size_t)(b0 + t) * rmbrStride), bs, 0, bits));
          }
          RF_CUDA(cudaDeviceSynchronize());
        }
        break;
      }
    }

    twBatches = wallAll();
    // ---- the one exchange of a tree-sharded forest: sum the ensemble accumulators over ranks, on the device ----
    bool wholeForest = T == cfg.ntree;
    double collMs = 0.0;
    auto wall = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    if (comm) {
      const double t0 = wall();
      comm_allreduce_ensembles(comm, d_allNum.p, d_oobNum.p, d_allDen.p, d_oobDen.p, n);
      DevBuf<unsigned long long> d_cntT(pool);
      d_cntT.alloc(1);
      const unsigned long long mine = (unsigned long long)T;
      RF_CUDA(cudaMemcpy(d_cntT.p, &mine, 8, cudaMemcpyHostToDevice));
      comm_allreduce_sum_u64(comm, d_cntT.p, 1);
      unsigned long long all = 0;
      RF_CUDA(cudaMemcpy(&all, d_cntT.p, 8, cudaMemcpyDeviceToHost));
      wholeForest = all == (unsigned long long)cfg.ntree;      // the ranks of this communicator hold every tree exactly once
      collMs += wall() - t0;
    }
    if (cfg.finalize) {
      finalize_kernel<<<(n + 255) / 256, 256>>>(d_allNum.p, d_allDen.p, n);
      finalize_kernel<<<(n + 255) / 256, 256>>>(d_oobNum.p, d_oobDen.p, n);
      launches += 2;
      RF_CUDA(cudaGetLastError());
    }
    // perfClas: only for a whole, finalized forest (a shard's ensembles are partial sums)
    double* h_perf = nullptr;
    if (cfg.finalize && (a->opt_low & RFSLAM_OPT_PERF) && wholeForest) {
      h_perf = perf_clas_table(a->opt_low, cfg.ntree, n, d_oobNum.p, d_oobDen.p, ds->dev.ev, ds->dev.inv, nullptr, &launches);   // [ntree][3]; a rank of a sharded forest returns the whole forest's table too
    }
    unsigned long long pathSum = 0;
    RF_CUDA(cudaMemcpy(&pathSum, d_pathSum.p, 8, cudaMemcpyDeviceToHost));
    const long long growAlgo = totalAlgo;                        // B_score + the in-bag half of B_part: the tree kernel's own
    const long long membAlgo = 16ll * (long long)pathSum;        // B_part's all-row term: the traversal kernel routes those rows
    const long long bootAlgo = 4ll * bootDraws;                  // 4 B per bootstrap draw written
    totalAlgo = growAlgo + membAlgo + bootAlgo;

    // ---- gather the shards' records to rank 0 (NCCL send / recv over NVLink, then a device merge into tree order) ----
    std::vector<int32_t> allLeaf;
    if (gather) {
      const bool tmg = getenv("RFSLAM_TIMING") != nullptr;
      const double tA = wall();
      RF_CUDA(cudaDeviceSynchronize());
      const double t0 = wall();
      comm_allgather_leaf_counts(comm, h_leafCount, cfg.ntree, allLeaf);
      const double tB = wall();
      std::vector<long long> nodeUnits(cfg.ntree), leafUnits(cfg.ntree);
      for (int b = 0; b < cfg.ntree; b++) { leafUnits[b] = allLeaf[b]; nodeUnits[b] = 2ll * allLeaf[b] - 1; }
      long long NNall = 0, NLall = 0;
      for (int b = 0; b < cfg.ntree; b++) { NNall += nodeUnits[b]; NLall += leafUnits[b]; }
      // rank 0's destinations are the page-locked host blocks the forest is returned in: the merge kernels write them
      // directly (device-accessible under unified addressing), so the gather, the merge and the download are one pass.
      // A destination that is not device-accessible (a block the pool had to malloc) goes through a device array.
      const bool root = comm_rank(comm) == 0;
      struct Stage { void* dev = nullptr; void* host = nullptr; size_t bytes = 0; };
      std::vector<Stage> staged;
      auto target = [&](void* hostDst, size_t bytes) -> void* {
        if (!root) return nullptr;
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, hostDst) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) return at.devicePointer;
        cudaGetLastError();
        Stage st; st.host = hostDst; st.bytes = bytes;
        RF_CUDA(cudaMalloc(&st.dev, std::max<size_t>(bytes, 16)));
        staged.push_back(st);
        return st.dev;
      };
      GatherLayout LN, LL;
      try {
        comm_gather_layout(comm, nodeUnits, LN); comm_gather_layout(comm, leafUnits, LL);
        const size_t nn = root ? (size_t)NNall : 0, nl = root ? (size_t)NLall : 0;
        comm_gather_i32_into(comm, LN, (const int32_t*)g_tree.p, 1, (int32_t*)target(h_treeID.grow(nn), nn * 4));
        comm_gather_i32_into(comm, LN, (const int32_t*)g_node.p, 1, (int32_t*)target(h_nodeID.grow(nn), nn * 4));
        comm_gather_i32_into(comm, LN, (const int32_t*)g_parm.p, 1, (int32_t*)target(h_parmID.grow(nn), nn * 4));
        comm_gather_f64_into(comm, LN, (const double*)g_cont.p, 1, (double*)target(h_contPT.grow(nn), nn * 8));
        if (ds->has_factor) comm_gather_i32_into(comm, LN, (const int32_t*)g_mwcp.p, 1, (int32_t*)target(h_mwcpW.grow(nn), nn * 4));
        if (wantStat) comm_gather_f64_into(comm, LN, (const double*)g_stat.p, 1, (double*)target(h_stat.grow(nn), nn * 8));
        comm_gather_i32_into(comm, LL, (const int32_t*)g_rcnt.p, 1, (int32_t*)target(h_tnRCNT.grow(nl), nl * 4));
        comm_gather_i32_into(comm, LL, (const int32_t*)g_acnt.p, 1, (int32_t*)target(h_tnACNT.grow(nl), nl * 4));
        comm_gather_i32_into(comm, LL, (const int32_t*)g_clas.p, 2, (int32_t*)target(h_tnCLAS.grow(2 * nl), 2 * nl * 4));
        const double tC = wall();
        comm_gather_finish(comm, LN); comm_gather_finish(comm, LL);
        for (auto& st : staged) RF_CUDA(cudaMemcpy(st.host, st.dev, st.bytes, cudaMemcpyDeviceToHost));
        if (tmg) fprintf(stderr, "[rfslam_b200_grow] rank %d exchange: device drain %.1f ms, leaf counts %.1f ms, issue (layouts, host blocks, recv buffers, NCCL, merges) %.1f ms, wait + release %.1f ms, %zu staged arrays\n",
                         comm_rank(comm), t0 - tA, tB - t0, tC - tB, wall() - tC, staged.size());
      } catch (...) {
        for (auto& st : staged) cudaFree(st.dev);
        try { comm_gather_finish(comm, LN); } catch (...) {}
        try { comm_gather_finish(comm, LL); } catch (...) {}
        throw;
      }
      for (auto& st : staged) cudaFree(st.dev);
      collMs += wall() - t0;
      if (root) {
        // rank 0 now returns the WHOLE forest: every tree id, leaf count and seed
        ids.resize(cfg.ntree); h_leafCount.resize(cfg.ntree); h_seed.resize(cfg.ntree);
        for (int b = 0; b < cfg.ntree; b++) { ids[b] = b + 1; h_leafCount[b] = allLeaf[b]; h_seed[b] = seedA[b]; }
      }
    }
    const int Tout = (int)ids.size();

    twHand = wallAll();
    // ---- hand over ----
    out->ntree = Tout; out->n = n;
    out->total_nodes = (int64_t)h_treeID.n; out->total_leaves = (int64_t)h_tnRCNT.n;
    out->tree_ids = (int32_t*)host_alloc((size_t)Tout * 4); memcpy(out->tree_ids, ids.data(), (size_t)Tout * 4);
    {
      // mwcpSZ is all zeros unless the table has factors: a page-locked block is cleared by the device's copy engine while
      // the host copies the ensembles below (a 128 MB host memset was 1 % of a configs[1] step, 1 GB of it on rank 0 at N = 8)
      const size_t bytes = std::max<size_t>(h_treeID.n, 1) * 4;
      out->mwcpSZ = (int32_t*)host_alloc(bytes);
      if (!out->mwcpSZ) { set_error("out of host memory"); throw CudaFail{RFSLAM_ENOMEM}; }
      cudaPointerAttributes at{};
      if (bytes >= (1u << 20) && cudaPointerGetAttributes(&at, out->mwcpSZ) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
        RF_CUDA(cudaMemsetAsync(at.devicePointer, 0, bytes, copyStream));
        mwcpCleared = true;
      } else { cudaGetLastError(); memset(out->mwcpSZ, 0, bytes); }
    }
    out->mwcpCT = (int32_t*)host_calloc((size_t)Tout * 4); out->mwcpPT = nullptr; out->mwcp_total = 0;
    if (ds->has_factor && h_mwcpW.n == h_treeID.n) {
      if (mwcpCleared) { RF_CUDA(cudaStreamSynchronize(copyStream)); mwcpCleared = false; }
      // mwcpSZ / mwcpPT / mwcpCT (saveTree rf.c:10773-10779): one word per factor split in record order, words per tree
      size_t words = 0;
      for (size_t q = 0; q < h_mwcpW.n; q++) if (h_mwcpW.p[q]) words++;
      out->mwcpPT = (int32_t*)host_alloc(std::max<size_t>(words, 1) * 4);
      size_t at = 0; int slot = 0;
      for (size_t q = 0; q < h_mwcpW.n; q++) {
        if (q > 0 && h_treeID.p[q] != h_treeID.p[q - 1]) slot++;
        if (h_mwcpW.p[q]) { out->mwcpSZ[q] = 1; out->mwcpPT[at++] = h_mwcpW.p[q]; if (slot < Tout) out->mwcpCT[slot]++; }
      }
      out->mwcp_total = (int64_t)words;
    }
    out->treeID = h_treeID.release(); out->nodeID = h_nodeID.release(); out->parmID = h_parmID.release();
    out->contPT = h_contPT.release(); out->splitStat = wantStat ? h_stat.release() : nullptr;
    out->leafCount = (int32_t*)host_alloc((size_t)Tout * 4); memcpy(out->leafCount, h_leafCount.data(), (size_t)Tout * 4);
    out->seed = (int32_t*)host_alloc((size_t)Tout * 4); memcpy(out->seed, h_seed.data(), (size_t)Tout * 4);
    out->tnRCNT = h_tnRCNT.release(); out->tnACNT = h_tnACNT.release(); out->tnCLAS = h_tnCLAS.release();
    out->oobEnsbCLS = host_copy(d_oobNum.p, 2 * (size_t)n); out->allEnsbCLS = host_copy(d_allNum.p, 2 * (size_t)n);
    out->oobEnsbDen = host_copy(d_oobDen.p, (size_t)n); out->allEnsbDen = host_copy(d_allDen.p, (size_t)n);
    if (cfg.want_membership) {
      out->nodeMembership = host_copy(d_nodeMemb.p, (size_t)T * n);
      out->bootMembership = host_copy(d_bootMemb.p, (size_t)T * n);
    }
    if (cfg.want_membr_lists) {
      out->rmbrMembership = host_copy(d_rmbr.p, (size_t)T * rmbrStride);
      out->ambrMembership = host_copy(d_ambr.p, (size_t)T * n);
      out->rmbr_stride = (int64_t)rmbrStride;
    }
    out->perfClas = h_perf;
    if (mwcpCleared) RF_CUDA(cudaStreamSynchronize(copyStream));
    out->split_candidates = totalCand; out->algorithmic_bytes = totalAlgo;
    out->grow_kernel_ms = growMs; out->total_device_ms = totalMs; out->kernel_launches = launches;
    out->grow_kernel_algorithmic_bytes = growAlgo; out->membership_algorithmic_bytes = membAlgo; out->bootstrap_algorithmic_bytes = bootAlgo;
    out->membership_kernel_ms = membMs; out->bootstrap_ms = bootMs; out->collective_ms = collMs;
    if (tmgAll) fprintf(stderr, "[rfslam_b200_grow] call %.1f ms: setup + batches (bootstrap, tree kernel, compaction, traversal, record copies) %.1f ms [device %.1f ms], exchange + ensembles + perf %.1f ms, hand-over %.1f ms\n",
                        wallAll() - tw0, twBatches - tw0, (double)totalMs, twHand - twBatches, wallAll() - twHand);
  } catch (CudaFail f) {
    rfslam_b200_free_forest(out);
    return f.code;
  }
  return RFSLAM_OK;
}

// ---------------------------------------------------------------------------------------------
// rfslam_b200_score_node: every cut of one covariate in one node, scored by the same device code
// the tree kernel runs (emit-only mode), for direct comparison with the reference plugin.
// ---------------------------------------------------------------------------------------------
int score_node_impl(int rule, double k_for_alpha, int m, const double* x, const double* response,
                    const int32_t* stratum, const double* risk_time, double* cut_value, double* delta,
                    int32_t* n_cuts, int device) {
  if (rule < 1 || rule > 6) { set_error("rule must be 1..6"); return RFSLAM_EINVAL; }
  if (!(k_for_alpha > 0.0)) { set_error("k_for_alpha must be > 0"); return RFSLAM_EINVAL; }
  rfslam_grow_args a;
  memset(&a, 0, sizeof(a));
  a.observation_size = m; a.x_size = 1; a.x_data = x; a.r_data = response; a.risk_time = risk_time;
  a.stratifier = stratum; a.device = device; a.x_type = "R";
  rfslam_dataset* ds = nullptr;
  int rc = dataset_build(&a, &ds);
  if (rc) return rc;
  try {
    const int n = m;
    GrowParams g{};
    g.rule = rule; g.mtry = 1; g.nodesize = 1; g.nodedepth = -1; g.alpha = 1.0 / (k_for_alpha * k_for_alpha);
    g.pw = 1; g.ncmax = 1; g.has_sort = ds->max_D > kHistBins; g.hbins = ds->hist_bins;
    { int ex = 0; frexp(std::max((double)m * std::max(ds->max_rt, 1e-300), 1.0), &ex); g.tscale = ldexp(1.0, 62 - ex); g.tinv = ldexp(1.0, ex - 62); }
    g.hslots = hist_slots(g.hbins); g.tabSlots = kTabSlots; g.subMax = 0; g.seqDraw = 0; g.facCuts = 0;
    const size_t smem = grow_smem_bytes(g.ncmax, ds->K, g.hbins, g.hslots, g.tabSlots, 0, 0);
    const GrowVariant* kv = grow_variant_t256f();                // the FULL build: emit mode, wide covariates
    { int perSm = 0; RF_CUDA(kv->prepare(smem, &perSm)); }
    DevBuf<uint32_t> d_lists, d_sort, d_stackPerm, d_ctab, d_nodeCount, d_leafCount, d_cursor;
    DevBuf<uint16_t> d_cnt; DevBuf<NodeEnt> d_stack; DevBuf<int> d_seedB, d_boot, d_status, d_sinkCount;
    DevBuf<unsigned long long> d_cand, d_algo; DevBuf<Rec> d_pool; DevBuf<double> d_statPool, d_sinkCut, d_sinkDelta;
    const size_t cntStride = ((size_t)n + 7) & ~(size_t)7;
    d_lists.alloc(2 * (size_t)n); if (g.has_sort) d_sort.alloc(4 * (size_t)n);
    d_cnt.alloc(cntStride); d_stack.alloc(kStackCap); d_stackPerm.alloc(kStackCap);
    d_ctab.alloc(4); d_nodeCount.alloc(1); d_leafCount.alloc(1); d_cursor.alloc(1); d_status.alloc(1);
    d_seedB.alloc(1); d_boot.alloc(1); d_cand.alloc(1); d_algo.alloc(1); d_pool.alloc(2 * kRecChunk); d_statPool.alloc(2 * kRecChunk);
    d_sinkCut.alloc(n); d_sinkDelta.alloc(n); d_sinkCount.alloc(1);
    std::vector<uint16_t> ones(cntStride, 1);
    RF_CUDA(cudaMemcpy(d_cnt.p, ones.data(), cntStride * 2, cudaMemcpyHostToDevice));
    int sb = -1, bs = n;
    RF_CUDA(cudaMemcpy(d_seedB.p, &sb, 4, cudaMemcpyHostToDevice));
    RF_CUDA(cudaMemcpy(d_boot.p, &bs, 4, cudaMemcpyHostToDevice));
    d_cursor.zero(); d_status.zero(); d_sinkCount.zero();
    GrowWork w{};
    w.ntrees = 1; w.seedB = d_seedB.p; w.bootSize = d_boot.p; w.lists = d_lists.p; w.listStride = (size_t)n; w.cnt = d_cnt.p; w.cntStride = cntStride;
    w.sortbuf = g.has_sort ? d_sort.p : nullptr; w.stack = d_stack.p; w.stackPerm = d_stackPerm.p;
    w.pool = d_pool.p; w.statPool = d_statPool.p; w.poolChunks = 2; w.poolCursor = d_cursor.p; w.ctab = d_ctab.p; w.maxChunksPerTree = 4;
    w.nodeCount = d_nodeCount.p; w.leafCount = d_leafCount.p; w.candCount = d_cand.p; w.algoBytes = d_algo.p; w.status = d_status.p;
    w.sinkCut = d_sinkCut.p; w.sinkDelta = d_sinkDelta.p; w.sinkCount = d_sinkCount.p;
    kv->launch(ds->dev, g, w, 1, smem);
    RF_CUDA(cudaGetLastError());
    RF_CUDA(cudaDeviceSynchronize());
    int nc = 0;
    RF_CUDA(cudaMemcpy(&nc, d_sinkCount.p, 4, cudaMemcpyDeviceToHost));
    *n_cuts = nc;
    if (nc > 0) {
      RF_CUDA(cudaMemcpy(cut_value, d_sinkCut.p, (size_t)nc * 8, cudaMemcpyDeviceToHost));
      RF_CUDA(cudaMemcpy(delta, d_sinkDelta.p, (size_t)nc * 8, cudaMemcpyDeviceToHost));
    }
  } catch (CudaFail f) {
    dataset_free(ds);
    return f.code;
  }
  dataset_free(ds);
  return RFSLAM_OK;
}

int bootstrap_counts_impl(int seed, int ntree, int n, int tree, int32_t* counts, int device) {
  if (seed >= 0 || ntree < 1 || tree < 1 || tree > ntree || n < 1) { set_error("bad bootstrap arguments"); return RFSLAM_EINVAL; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: librfslam_b200 has no CPU fallback"); return RFSLAM_ENODEV; }
  try {
    if (device >= 0) RF_CUDA(cudaSetDevice(device));
    std::vector<int> sA, sB;
    chain_seeds(seed, ntree, sA, sB);
    DevBuf<int> d_seed, d_bs; DevBuf<uint32_t> d_draws;
    d_seed.alloc(1); d_bs.alloc(1); d_draws.alloc(n);
    RF_CUDA(cudaMemcpy(d_seed.p, &sA[tree - 1], 4, cudaMemcpyHostToDevice));
    RF_CUDA(cudaMemcpy(d_bs.p, &n, 4, cudaMemcpyHostToDevice));
    bootstrap_draw_kernel<<<1, 32>>>(d_seed.p, d_bs.p, d_draws.p, (size_t)n, n);
    RF_CUDA(cudaGetLastError());
    std::vector<uint32_t> h(n);
    RF_CUDA(cudaMemcpy(h.data(), d_draws.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++) counts[i] = 0;
    for (int i = 0; i < n; i++) counts[h[i]]++;
  } catch (CudaFail f) {
    return f.code;
  }
  return RFSLAM_OK;
}

}  // namespace rfslam
