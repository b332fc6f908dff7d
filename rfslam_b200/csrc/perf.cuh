// perf.cuh -- performance of the finalized OOB ensemble of a two-class CPIU forest (the callers of the ensemble,
// SURVEY 8 f3): getPerformance's class-family branch (rf.c:7954-8010) for all three perf.type values and the
// RFQ vote, from one pass over the rows.
//   * default  (OPT_PERF):       misclassification rate (getClassificationIndex rf.c:1067-1095) + per class over ALL
//                                rows of the class (getConditionalClassificationIndex rf.c:1025-1066)
//   * "brier"  (OPT_PERF_CALB):  normalised Brier score + per class one-against-all (getBrierScore rf.c:974-1024)
//   * "g.mean" (OPT_PERF_GMN2):  1 - sqrt(prod of per-class true rates) (getGMeanIndex rf.c:1096-1140); the per-class
//                                slots stay NA
//   * rfq      (OPT_CLAS_RFQ):   the vote is "minority class iff its probability >= its prevalence" instead of the
//                                last class holding the maximum (getMaxVote rf.c:1175-1218, rf.c:16859-16903)
#pragma once
#include "common.cuh"
#include "internal.h"

namespace rfslam {
namespace {

constexpr int kPerfBlocks = 296;

struct PerfAcc {                         // integer counts (exact) and per-block Brier partial sums (summed in block order)
  unsigned long long cls[2], cum[2], ok[2];
};

__global__ void class_count_kernel(const uint8_t* ev, const double* y, int n, unsigned long long* cnt) {
  unsigned long long c1 = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) c1 += y ? (y[i] == 2.0 ? 1u : 0u) : (ev[i] ? 1u : 0u);
  c1 = warp_sum(c1);
  if ((threadIdx.x & 31) == 0 && c1) atomicAdd(cnt, c1);
}

// rows are ORIGINAL rows; class of row i = ev[inv[i]] (grow: the resident table) or y[i] (predict / restore).
// rfqMinority < 0: plain max vote, else the 0-based minority class with its threshold.
__global__ void perf_kernel(const double* oob, const uint32_t* den, const uint8_t* ev, const uint32_t* inv, const double* y, int n,
                            int rfqMinority, double rfqThreshold, PerfAcc* acc, double* brier /* [gridDim.x][2] */) {
  __shared__ double sb[2][8];
  unsigned long long cum0 = 0, cum1 = 0, ok0 = 0, ok1 = 0;
  double b0 = 0.0, b1 = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int cls = y ? (y[i] == 2.0 ? 1 : 0) : (ev[inv[i]] ? 1 : 0);
    if (den[i] != 0u) {
      const double p1 = oob[i], p2 = oob[(size_t)n + i];
      int vote;
      if (rfqMinority < 0) vote = (p1 <= p2) ? 1 : 0;              // the LAST class whose probability is >= the running maximum
      else vote = ((rfqMinority ? p2 : p1) >= rfqThreshold) ? rfqMinority : 1 - rfqMinority;
      if (cls) { cum1++; if (vote == cls) ok1++; } else { cum0++; if (vote == cls) ok0++; }
      const double d1 = (cls == 0 ? 1.0 : 0.0) - p1, d2 = (cls == 1 ? 1.0 : 0.0) - p2;
      b0 += d1 * d1; b1 += d2 * d2;                                // pow(x, 2.0) is x * x exactly
    }
  }
  cum0 = warp_sum(cum0); cum1 = warp_sum(cum1); ok0 = warp_sum(ok0); ok1 = warp_sum(ok1);
  b0 = warp_sum(b0); b1 = warp_sum(b1);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    if (cum0) atomicAdd(&acc->cum[0], cum0);
    if (cum1) atomicAdd(&acc->cum[1], cum1);
    if (ok0) atomicAdd(&acc->ok[0], ok0);
    if (ok1) atomicAdd(&acc->ok[1], ok1);
    sb[0][w] = b0; sb[1][w] = b1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s0 = 0.0, s1 = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); k++) { s0 += sb[0][k]; s1 += sb[1][k]; }
    brier[2 * blockIdx.x] = s0; brier[2 * blockIdx.x + 1] = s1;
  }
}

// One performance row [overall, class 1, class 2] from the counts of a pass over the rows (getPerformance's class
// branch, rf.c:7954-8010): cum / ok = rows with an estimate per class and the correctly voted ones, cls = ALL rows of
// each class, brier = sums of squared one-against-all errors.  `out` must hold NA_REAL on entry.
inline void perf_from_counts(uint32_t opt_low, const unsigned long long* cumK, const unsigned long long* okK,
                             const unsigned long long* cls, const double* brier, double* out) {
  const unsigned long long cum = cumK[0] + cumK[1];
  if (!cum) return;
  if (opt_low & RFSLAM_OPT_PERF_CALB) {
    const double cpv[2] = {brier[0] / (double)cum, brier[1] / (double)cum};
    out[0] = (cpv[0] + cpv[1]) * 2.0 / 1.0;                         // result * levels / (levels - 1)
    out[1] = cpv[0]; out[2] = cpv[1];
  } else if (opt_low & RFSLAM_OPT_PERF_GMN2) {                      // two classes: the minority flag is always set (rf.c:16899-16903)
    double result = 1.0;
    for (int k = 0; k < 2; k++) {
      const double t = (double)okK[k], denom = (double)cumK[k];     // true + false votes of the class
      result = denom > 0 ? result * t / denom : result * (1 + t) / (1 + denom);
    }
    out[0] = 1.0 - sqrt(result);
  } else {
    out[0] = 1.0 - (double)(okK[0] + okK[1]) / (double)cum;
    if (cls[0]) out[1] = 1.0 - (double)okK[0] / (double)cls[0];
    if (cls[1]) out[2] = 1.0 - (double)okK[1] / (double)cls[1];
  }
}

// perfClas [ntree][3]: NA_REAL except the last row (perfBlock = ntree, rf.c:8155).  `ev`/`inv` or `y` as in perf_kernel.
// `trainCls` (predict mode): class counts of the TRAINING response -- the RFQ vote's minority class and threshold come from
// it in every mode (rf.c:16859-16895), the per-class denominators from the rows scored (rf.c:1037-1040).
inline double* perf_clas_table(uint32_t opt_low, int ntree, int n, const double* d_oob, const uint32_t* d_den,
                               const uint8_t* d_ev, const uint32_t* d_inv, const double* d_y, long long* launches,
                               const unsigned long long* trainCls = nullptr) {
  struct Dev { void* p = nullptr; ~Dev() { if (p) cudaFree(p); } } b_acc, b_brier, b_cnt;
  RF_CUDA(cudaMalloc(&b_acc.p, sizeof(PerfAcc))); RF_CUDA(cudaMalloc(&b_brier.p, kPerfBlocks * 2 * sizeof(double))); RF_CUDA(cudaMalloc(&b_cnt.p, 8));
  RF_CUDA(cudaMemset(b_acc.p, 0, sizeof(PerfAcc))); RF_CUDA(cudaMemset(b_cnt.p, 0, 8));
  class_count_kernel<<<kPerfBlocks, 256>>>(d_ev, d_y, n, (unsigned long long*)b_cnt.p);
  unsigned long long c1 = 0;
  RF_CUDA(cudaMemcpy(&c1, b_cnt.p, 8, cudaMemcpyDeviceToHost));
  const unsigned long long cls[2] = {(unsigned long long)n - c1, c1};
  int minority = -1; double threshold = 0.0;
  if (opt_low & RFSLAM_OPT_CLAS_RFQ) {                              // rf.c:16876-16895: first class with the smallest count
    const unsigned long long* tc = trainCls ? trainCls : cls;
    minority = tc[1] < tc[0] ? 1 : 0;
    threshold = (double)tc[minority] / (double)(tc[0] + tc[1]);
  }
  perf_kernel<<<kPerfBlocks, 256>>>(d_oob, d_den, d_ev, d_inv, d_y, n, minority, threshold, (PerfAcc*)b_acc.p, (double*)b_brier.p);
  if (launches) *launches += 2;
  RF_CUDA(cudaGetLastError());
  PerfAcc acc; double brier[kPerfBlocks * 2];
  RF_CUDA(cudaMemcpy(&acc, b_acc.p, sizeof acc, cudaMemcpyDeviceToHost));
  RF_CUDA(cudaMemcpy(brier, b_brier.p, sizeof brier, cudaMemcpyDeviceToHost));
  double* perf = (double*)host_alloc((size_t)ntree * 3 * sizeof(double));
  if (!perf) { set_error("out of host memory"); throw CudaFail{RFSLAM_ENOMEM}; }
  const double na = rfslam_na_real();
  for (size_t k = 0; k < (size_t)ntree * 3; k++) perf[k] = na;
  double bsum[2] = {0.0, 0.0};
  for (int b = 0; b < kPerfBlocks; b++) { bsum[0] += brier[2 * b]; bsum[1] += brier[2 * b + 1]; }
  perf_from_counts(opt_low, acc.cum, acc.ok, cls, bsum, perf + (size_t)(ntree - 1) * 3);
  return perf;
}

}  // namespace
}  // namespace rfslam
