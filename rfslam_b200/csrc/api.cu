// api.cu -- the extern "C" surface declared in include/rfslam_b200.h.
#include <chrono>
#include <map>
#include <mutex>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "internal.h"

namespace rfslam {

extern std::vector<unsigned long long> g_last_profile;

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

}  // namespace rfslam

namespace rfslam {
// Importance inside a grow call (rf.c:20889-20910 computes it tree by tree while the forest grows; here the finished
// forest is restored on its own training rows -- same trees, same bootstrap, the grow call's chain-C seeds).
int grow_importance(const rfslam_grow_args* a, rfslam_forest_out* out) {
  if (!(a->opt_low & RFSLAM_OPT_VIMP) || !(a->opt_low & RFSLAM_OPT_PERF)) return RFSLAM_OK;      // rf.c:1282-1286
  const bool byUser = (a->opt_low & RFSLAM_OPT_BOOT_TYP1) && (a->opt_low & RFSLAM_OPT_BOOT_TYP2);
  if (a->comm || a->tree_step > 1 || out->ntree != a->ntree || !a->finalize_ensembles) {
    set_error("importance needs the whole forest in one process (no tree sharding) with finalized ensembles"); return RFSLAM_ENOSUP;
  }
  if (!a->x_data || !a->r_data) { set_error("importance needs xData / rData"); return RFSLAM_EINVAL; }
  if (!out->treeID) { set_error("importance needs the forest records (RFSLAM_OPT_TREE)"); return RFSLAM_EINVAL; }
  std::vector<int32_t> intr(a->x_size);
  for (int k = 0; k < a->x_size; k++) intr[k] = k + 1;                                           // processDefaultGrow rf.c:1220
  // chain C: third block of the seed LCG (rf.c:6685-6692)
  std::vector<int> seedC(a->ntree);
  {
    unsigned s = (unsigned)abs(a->seed);
    if (s >= 714025u) s %= 714025u;
    for (int pass = 0; pass < 3; pass++)
      for (int b = 0; b < a->ntree; b++) {
        s = lcg_next(s); s = lcg_next(s);
        while (s == 0) s = lcg_next(s);
        if (pass == 2) seedC[b] = -(int)s;
      }
  }
  rfslam_predict_args pa;
  memset(&pa, 0, sizeof pa);
  pa.seed = a->seed; pa.opt_low = a->opt_low; pa.opt_high = 0;
  pa.ntree = a->ntree; pa.observation_size = a->observation_size; pa.y_size = a->y_size; pa.r_type = a->r_type; pa.r_levels = a->r_levels;
  pa.r_data = a->r_data; pa.x_size = a->x_size; pa.x_type = a->x_type; pa.x_levels = a->x_levels; pa.x_data = a->x_data;
  pa.k_for_alpha = a->k_for_alpha; pa.max_bootstrap_size = a->max_bootstrap_size; pa.bootstrap_size = a->bootstrap_size;
  pa.bootstrap = byUser ? a->bootstrap : nullptr;
  pa.total_node_count = (int32_t)out->total_nodes; pa.forest_seed = out->seed;
  pa.treeID = out->treeID; pa.nodeID = out->nodeID; pa.parmID = out->parmID; pa.contPT = out->contPT; pa.mwcpSZ = out->mwcpSZ;
  pa.mwcpPT = out->mwcpPT; pa.mwcp_total = out->mwcp_total;
  pa.tnCLAS = out->tnCLAS; pa.tnCLAS_len = out->tnCLAS ? 2 * out->total_leaves : 0;
  pa.intr_predictor_size = a->x_size; pa.intr_predictor = intr.data();
  pa.device = a->device; pa.finalize_ensembles = 1;
  rfslam_predict_out po;
  const int rc = predict_forest_ex(&pa, &po, seedC.data());
  if (rc) return rc;
  out->vimpClas = po.vimpClas; out->vimp_count = po.vimp_count; out->vimp_ms = po.vimp_ms;
  po.vimpClas = nullptr;
  rfslam_b200_free_predict(&po);
  return RFSLAM_OK;
}
}  // namespace rfslam

using namespace rfslam;

static std::mutex g_scratchMu;
static std::map<int, WsPool> g_scratch;      // per device: idle scratch blocks kept between rfslam_b200_grow calls

extern "C" {

const char* rfslam_b200_last_error(void) { return g_err; }
const char* rfslam_b200_version(void) { return "rfslam_b200 0.1 (sm_100a)"; }

int rfslam_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

void rfslam_b200_register_defaults(void) { reset_rule_defaults(); }

int rfslam_b200_register_this(int rule, int family, int slot) {
  if (family < 0 || family > 3 || slot < 1 || slot > 16) { set_error("family must be 0..3 and slot 1..16 (rf.c:8856-8866)"); return RFSLAM_EINVAL; }
  if (rule < 0 || rule > 6) { set_error("only the six Poisson rules have device implementations"); return RFSLAM_ENOSUP; }
  if (rule != 0 && family != RFSLAM_CLAS_FAM) { set_error("the Poisson rules are classification-family rules (sc.c:47-62)"); return RFSLAM_ENOSUP; }
  rule_table_set(family, slot, rule);
  return RFSLAM_OK;
}

int rfslam_b200_rule_of_slot(int family, int slot) {
  if (family < 0 || family > 3 || slot < 1 || slot > 16) return 0;
  const int r = rule_table_get(family, slot);
  return r > 0 ? r : 0;                       // empty and foreign (user-supplied CPU rule) slots have no device rule
}

int rfslam_b200_dataset_create(const rfslam_grow_args* args, rfslam_dataset** ds) {
  if (!args || !ds) { set_error("null arguments"); return RFSLAM_EINVAL; }
  return dataset_build(args, ds);
}

void rfslam_b200_dataset_destroy(rfslam_dataset* ds) { dataset_free(ds); }

int64_t rfslam_b200_dataset_bytes(const rfslam_dataset* ds) { return ds ? ds->bytes : 0; }


int rfslam_b200_grow_resident(rfslam_dataset* ds, const rfslam_grow_args* args, rfslam_forest_out* out) {
  if (!ds || !args || !out) { set_error("null arguments"); return RFSLAM_EINVAL; }
  int rc = grow_forest(ds, args, out);
  if (rc == RFSLAM_OK && (rc = grow_importance(args, out)) != RFSLAM_OK) rfslam_b200_free_forest(out);
  return rc;
}

int rfslam_b200_grow(const rfslam_grow_args* args, rfslam_forest_out* out) {
  if (!args || !out) { set_error("null arguments"); return RFSLAM_EINVAL; }
  memset(out, 0, sizeof(*out));
  GrowConfig cfg;
  int rc = validate_grow_args(args, &cfg);
  if (rc) return rc;
  const bool timing = getenv("RFSLAM_TIMING") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = now();
  // device memory of the previous call on this device (several GB of entry lists and stacks, the rank-coded
  // table): cudaMalloc / cudaFree of it cost up to 0.6 s per call, and contend between the ranks of a node
  WsPool seed;
  int dev = 0;
  {
    if (args->device >= 0) dev = args->device; else cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_scratchMu);
    auto& cache = g_scratch[dev];
    seed.blks.swap(cache.blks);
  }
  rfslam_dataset* ds = nullptr;
  rc = dataset_build(args, &ds, &seed);
  if (rc) {                                   // dataset_build released what it had adopted; anything left goes back
    std::lock_guard<std::mutex> lk(g_scratchMu);
    for (auto& b : seed.blks) g_scratch[dev].blks.push_back(b);      // back to the device they were taken from
    return rc;
  }
  const double t1 = now();
  rc = grow_forest(ds, args, out);
  if (rc == RFSLAM_OK && (rc = grow_importance(args, out)) != RFSLAM_OK) rfslam_b200_free_forest(out);
  const double t2 = now();
  {
    std::lock_guard<std::mutex> lk(g_scratchMu);
    auto& cache = g_scratch[ds->device];
    for (auto& b : ds->pool.blks) { b.used = false; cache.blks.push_back(b); }
    ds->pool.blks.clear();
  }
  dataset_free(ds);
  if (timing) fprintf(stderr, "[rfslam_b200_grow] dataset %.1f ms, grow %.1f ms, release %.1f ms\n", t1 - t0, t2 - t1, now() - t2);
  return rc;
}

void rfslam_b200_free_forest(rfslam_forest_out* o) {
  if (!o) return;
  void* arrays[] = {o->tree_ids, o->treeID, o->nodeID, o->parmID, o->contPT, o->mwcpSZ, o->splitStat, o->leafCount, o->seed,
                    o->tnRCNT, o->tnACNT, o->tnCLAS, o->nodeMembership, o->bootMembership, o->rmbrMembership, o->ambrMembership,
                    o->oobEnsbCLS, o->allEnsbCLS, o->oobEnsbDen, o->allEnsbDen, o->perfClas, o->mwcpPT, o->mwcpCT, o->vimpClas};
  for (void* p : arrays) host_free(p);
  memset(o, 0, sizeof(*o));
}

void rfslam_b200_free_array(void* p) { host_free(p); }
void rfslam_b200_host_pool_trim(void) {
  host_pool_trim();
  std::lock_guard<std::mutex> lk(g_scratchMu);
  for (auto& kv : g_scratch) { cudaSetDevice(kv.first); kv.second.release_all(); }
}

int rfslam_b200_predict(const rfslam_predict_args* args, rfslam_predict_out* out) {
  if (!args || !out) { set_error("null arguments"); return RFSLAM_EINVAL; }
  return predict_forest(args, out);
}

void rfslam_b200_free_predict(rfslam_predict_out* o) {
  if (!o) return;
  void* arrays[] = {o->allEnsbCLS, o->allEnsbDen, o->oobEnsbCLS, o->oobEnsbDen, o->perfClas, o->leafCount, o->nodeMembership, o->bootMembership, o->vimpClas};
  for (void* p : arrays) host_free(p);
  memset(o, 0, sizeof(*o));
}

int rfslam_b200_score_node(int rule, double k_for_alpha, int m, const double* x, const double* response,
                           const int32_t* stratum, const double* risk_time, double* cut_value, double* delta,
                           int32_t* n_cuts, int device) {
  if (!x || !response || !stratum || !risk_time || !cut_value || !delta || !n_cuts) { set_error("null arguments"); return RFSLAM_EINVAL; }
  return score_node_impl(rule, k_for_alpha, m, x, response, stratum, risk_time, cut_value, delta, n_cuts, device);
}

int rfslam_b200_last_profile(uint64_t* slots, int capacity) {
  int k = 0;
  for (; k < capacity && k < (int)g_last_profile.size(); k++) slots[k] = g_last_profile[k];
  return k;
}

int rfslam_b200_bootstrap_counts(int seed, int ntree, int n, int tree, int32_t* counts, int device) {
  if (!counts) { set_error("null arguments"); return RFSLAM_EINVAL; }
  return bootstrap_counts_impl(seed, ntree, n, tree, counts, device);
}

}  // extern "C"
