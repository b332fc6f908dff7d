// plugin.cu -- the reference's custom split-rule ABI on the device.
//
// The reference selects a split rule through a table of function pointers,
// customFunctionArray[4 families][16 slots] (rf.c:486), filled by registerCustomFunctions() (sc.c:23-67)
// through registerThis(func, family, slot) (rf.c:8856-8866) and called once per candidate cut with the
// customFunction signature (splitCustom.h:24-44, rf.c:13469-13482).  This file keeps those symbols and that
// table so that code written against the reference's plugin header links and behaves the same way:
//
//   * poissonSplit1..6 have the reference's exact signature and evaluate ONE cut on the device
//     (per-(side, stratum) sums, then the rule's statistic) -- the symbol a known-answer test or a
//     maintainer's tool calls; the tree kernel never goes through them, it scores all cuts of all drawn
//     covariates of a node at once (grow_kernel.cuh), with the same per-cell term (common.cuh side_term);
//   * registerThis() recognises those six pointers and maps the slot to the device rule; any other
//     function pointer is a user-supplied CPU rule, which has no device implementation: the slot is
//     marked foreign and a grow call that selects it fails with RFSLAM_ENOSUP (no CPU fallback);
//   * registerCustomFunctions() makes the registrations of sc.c:47-62.
#include <mutex>
#include <vector>

#include "internal.h"

namespace rfslam {

namespace {

constexpr int kForeign = -1;
int g_rule_table[4][16];
bool g_rule_init = false;
std::mutex g_ruleMu;                          // the registry may be read by concurrent grow calls

// one CTA: per-(side, stratum) sufficient statistics of the node, then the statistic of the cut.
// Arrays are the plugin's 1-based ones shifted to 0-based by the caller.
__global__ void plugin_stat_kernel(int rule, int n, double alpha, const char* __restrict__ memb, const double* __restrict__ resp,
                                   const double* __restrict__ strat, const double* __restrict__ rt, int K, double* out, int* bad) {
  extern __shared__ unsigned char raw[];
  double* T = (double*)raw;                   // [2][K+1]
  double* W = T + 2 * (K + 1);
  unsigned* N = (unsigned*)(W + 2 * (K + 1));
  unsigned* E = N + 2 * (K + 1);
  __shared__ double ysumS, rtsumS;
  const bool stratified = rule_is_stratified(rule), unit = rule_unit_time(rule), bayes = rule_is_bayes(rule);
  for (int i = threadIdx.x; i < 2 * (K + 1); i += blockDim.x) { T[i] = 0.0; W[i] = 0.0; N[i] = 0u; E[i] = 0u; }
  if (threadIdx.x == 0) { ysumS = 0.0; rtsumS = 0.0; }
  __syncthreads();
  double ysum = 0.0, rtsum = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double y = resp[i];
    if (!(y >= 1.0 && y <= 2.0)) atomicExch(bad, 1);             // sc.c:676-687: the plugin exits on other codes
    const int side = memb[i] == 1 ? 0 : 1;                       // LEFT = 1 (splitCustom.h:5-6)
    const int j = stratified ? (int)strat[i] : 1;
    const double t = unit ? 1.0 : rt[i];
    const double e = y - 1.0;
    const int cell = side * (K + 1) + j;
    atomicAdd(&N[cell], 1u);
    if (y == 2.0) atomicAdd(&E[cell], 1u);
    atomicAdd(&T[cell], t);
    atomicAdd(&W[cell], e * t);
    ysum += y; rtsum += rt[i];                                   // beta uses the real risk time even under the unit-time rule (sc.c:999-1004)
  }
  atomicAdd(&ysumS, ysum); atomicAdd(&rtsumS, rtsum);
  __syncthreads();
  if (threadIdx.x == 0) {
    const double beta = bayes ? alpha / (ysumS / rtsumS) : 0.0;
    double sL = 0.0, sR = 0.0, sP = 0.0;
    for (int j = 1; j <= K; j++) {
      const double EL = E[j], NL = N[j], TL = T[j], WL = W[j];
      const double ER = E[K + 1 + j], NR = N[K + 1 + j], TR = T[K + 1 + j], WR = W[K + 1 + j];
      sL += side_term(bayes, EL, NL, TL, WL, alpha, beta);
      sR += side_term(bayes, ER, NR, TR, WR, alpha, beta);
      sP += side_term(bayes, EL + ER, NL + NR, TL + TR, WL + WR, alpha, beta);
    }
    *out = sL + sR - sP;                                         // sc.c:780, sc.c:1311
  }
}

struct Tmp {
  void* p = nullptr;
  ~Tmp() { if (p) cudaFree(p); }
  template <typename T> T* alloc(size_t count) { RF_CUDA(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T))); return (T*)p; }
};

double plugin_call(int rule, unsigned n, double k_for_alpha, char* membership, double* response, unsigned maxLevel,
                   double** feature, unsigned featureCount) {
  const double na = rfslam_na_real();
  if (!membership || !response || !feature || featureCount < 2 || n < 1) { set_error("poissonSplit%d: membership, response and feature[1..2] (stratum, risk time) are required", rule); return na; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: librfslam_b200 has no CPU fallback"); return na; }
  (void)maxLevel;
  try {
    // the plugin's arrays are 1-based (rf.c:13455-13482)
    const char* memb = membership + 1; const double* resp = response + 1;
    const double* strat = feature[1] + 1; const double* rt = feature[2] + 1;
    int K = 1;
    if (rule_is_stratified(rule)) for (unsigned i = 0; i < n; i++) { if (!(strat[i] >= 1.0)) { set_error("stratum must be >= 1"); return na; } K = std::max(K, (int)strat[i]); }   // sc.c:673-675
    if (K > kMaxStrata) { set_error("largest stratum %d exceeds the %d strata the device path holds in shared memory", K, kMaxStrata); return na; }
    Tmp tm, tr, ts, tt, to, tb;
    char* d_m = tm.alloc<char>(n); double* d_r = tr.alloc<double>(n); double* d_s = ts.alloc<double>(n); double* d_t = tt.alloc<double>(n);
    double* d_o = to.alloc<double>(1); int* d_b = tb.alloc<int>(1);
    RF_CUDA(cudaMemcpy(d_m, memb, n, cudaMemcpyHostToDevice));
    RF_CUDA(cudaMemcpy(d_r, resp, (size_t)n * 8, cudaMemcpyHostToDevice));
    RF_CUDA(cudaMemcpy(d_s, strat, (size_t)n * 8, cudaMemcpyHostToDevice));
    RF_CUDA(cudaMemcpy(d_t, rt, (size_t)n * 8, cudaMemcpyHostToDevice));
    RF_CUDA(cudaMemset(d_b, 0, 4));
    const size_t smem = (size_t)2 * (K + 1) * (8 + 8 + 4 + 4);
    plugin_stat_kernel<<<1, 256, smem>>>(rule, (int)n, 1.0 / (k_for_alpha * k_for_alpha), d_m, d_r, d_s, d_t, K, d_o, d_b);
    RF_CUDA(cudaGetLastError());
    double out = na; int bad = 0;
    RF_CUDA(cudaMemcpy(&out, d_o, 8, cudaMemcpyDeviceToHost));
    RF_CUDA(cudaMemcpy(&bad, d_b, 4, cudaMemcpyDeviceToHost));
    if (bad && (rule == 1 || rule == 3 || rule == 4 || rule == 6)) { set_error("response codes must be in [1, 2] (sc.c:676-687)"); return na; }
    return out;
  } catch (CudaFail) {
    return na;
  }
}

}  // namespace

void ensure_rule_defaults() {
  std::lock_guard<std::mutex> lk(g_ruleMu);
  if (!g_rule_init) {
    memset(g_rule_table, 0, sizeof g_rule_table);
    for (int r = 1; r <= 6; r++) g_rule_table[RFSLAM_CLAS_FAM][r] = r;      // poissonSplitN in CLAS_FAM slot N + 1 (sc.c:47-62)
    g_rule_init = true;
  }
}

void reset_rule_defaults() {
  { std::lock_guard<std::mutex> lk(g_ruleMu); g_rule_init = false; }
  ensure_rule_defaults();
}

int rule_table_get(int family, int slot) {
  ensure_rule_defaults();
  std::lock_guard<std::mutex> lk(g_ruleMu);
  return g_rule_table[family][slot - 1];
}

void rule_table_set(int family, int slot, int rule) {
  ensure_rule_defaults();
  std::lock_guard<std::mutex> lk(g_ruleMu);
  g_rule_table[family][slot - 1] = rule;
}

bool rule_slot_is_foreign(int family, int slot) { return rule_table_get(family, slot) == kForeign; }

}  // namespace rfslam

using namespace rfslam;

extern "C" {

#define RFSLAM_PLUGIN(N)                                                                                              \
  double poissonSplit##N(unsigned int n, double k_for_alpha, char* membership, double* time, double* event,           \
                         unsigned int eventTypeSize, unsigned int eventTimeSize, double* eventTime, double* response, \
                         double mean, double variance, unsigned int maxLevel, double** feature, unsigned int featureCount) { \
    (void)time; (void)event; (void)eventTypeSize; (void)eventTimeSize; (void)eventTime; (void)mean; (void)variance;   \
    return plugin_call(N, n, k_for_alpha, membership, response, maxLevel, feature, featureCount);                     \
  }
RFSLAM_PLUGIN(1) RFSLAM_PLUGIN(2) RFSLAM_PLUGIN(3) RFSLAM_PLUGIN(4) RFSLAM_PLUGIN(5) RFSLAM_PLUGIN(6)
#undef RFSLAM_PLUGIN

void registerThis(rfslam_custom_function func, unsigned int family, unsigned int slot) {
  if (slot < 1 || slot > 16 || family > 3) {                     // rf.c:8860-8865 prints and exits; here the registration is refused
    set_error("Invalid slot for custom split rule: family %u slot %u (family 0..3, slot 1..16)", family, slot);
    fprintf(stderr, "RF-SRC (b200): %s\n", rfslam_b200_last_error());
    return;
  }
  static const rfslam_custom_function known[6] = {poissonSplit1, poissonSplit2, poissonSplit3, poissonSplit4, poissonSplit5, poissonSplit6};
  int rule = func ? kForeign : 0;
  for (int r = 0; r < 6; r++) if (func == known[r]) rule = r + 1;
  if (rule > 0 && family != RFSLAM_CLAS_FAM) rule = kForeign;    // the Poisson rules are classification-family rules (sc.c:47-62)
  rule_table_set((int)family, (int)slot, rule);
}

void registerCustomFunctions(void) {
  // sc.c:47-62; the other registrations of the reference (getCustomSplitStatistic* in slot 1 and the REGR_FAM
  // slots) are CPU rules with no device implementation and stay empty
  registerThis(&poissonSplit1, RFSLAM_CLAS_FAM, 2);
  registerThis(&poissonSplit2, RFSLAM_CLAS_FAM, 3);
  registerThis(&poissonSplit3, RFSLAM_CLAS_FAM, 4);
  registerThis(&poissonSplit4, RFSLAM_CLAS_FAM, 5);
  registerThis(&poissonSplit5, RFSLAM_CLAS_FAM, 6);
  registerThis(&poissonSplit6, RFSLAM_CLAS_FAM, 7);
}

}  // extern "C"
