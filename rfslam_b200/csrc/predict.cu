// predict.cu -- held-out prediction: route test rows down every tree of a forest in the
// reference's wire format and average the leaf class proportions.
//   routing ........ rf.c:3687-3747 (LEFT iff contPT - x >= 0, rf.c:10610)
//   leaf estimate .. rf.c:755-858   (class counts of the leaf's in-bag members)
//   ensemble ....... rf.c:928-962, rf.c:8113-8125
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "internal.h"

namespace rfslam {

namespace {

// one visited node = one 16-byte load: cut value, covariate + 1 (0 = terminal), right child (internal) or leaf id (terminal)
struct __align__(16) PredRec { double cont; uint32_t parm; uint32_t link; };

// one thread per test row; trees are walked in order so the sum is reproducible
__global__ void predict_kernel(int ntree, int m, const unsigned long long* nodeOff, const unsigned long long* leafOff,
                               const PredRec* __restrict__ rec, const int32_t* tnCLAS, const double* __restrict__ fx,
                               double* ens, uint32_t* den, int32_t* membership) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  double a1 = 0.0, a2 = 0.0;
  for (int t = 0; t < ntree; t++) {
    const unsigned long long off = nodeOff[t];
    unsigned long long q = off;
    PredRec r = rec[q];
    while (r.parm != 0u) {
      const double xv = fx[(size_t)(r.parm - 1u) * m + i];
      q = ((r.cont - xv) >= 0.0) ? q + 1 : off + r.link;          // LEFT iff cut - x >= 0 (rf.c:10610)
      r = rec[q];
    }
    const int leaf = (int)r.link;
    const unsigned long long li = leafOff[t] + (unsigned long long)leaf - 1ull;
    const double c1 = (double)tnCLAS[2 * li], c2 = (double)tnCLAS[2 * li + 1];
    const double rc = c1 + c2;
    a1 += c1 / rc; a2 += c2 / rc;
    if (membership) membership[(size_t)t * m + i] = leaf;
  }
  ens[i] = a1; ens[(size_t)m + i] = a2; den[i] = (uint32_t)ntree;
}

__global__ void finalize_pred_kernel(double* ens, const uint32_t* den, int m) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m && den[i]) { ens[i] /= (double)den[i]; ens[(size_t)m + i] /= (double)den[i]; }
}

struct Buf {
  void* p = nullptr;
  ~Buf() { if (p) cudaFree(p); }
  void* alloc(size_t bytes) { RF_CUDA(cudaMalloc(&p, std::max<size_t>(bytes, 16))); return p; }
};

}  // namespace

int predict_forest(const rfslam_predict_args* a, rfslam_predict_out* out) {
  memset(out, 0, sizeof(*out));
  if (!a || a->ntree < 1 || a->total_nodes < 1 || a->fobservation_size < 1 || !a->fx_data) { set_error("bad predict arguments"); return RFSLAM_EINVAL; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: librfslam_b200 has no CPU fallback"); return RFSLAM_ENODEV; }
  const int ntree = a->ntree, m = a->fobservation_size, p = a->x_size;
  const int64_t NN = a->total_nodes;
  // rebuild right-child links from the pre-order records (what restoreTree does by recursion, rf.c:10807-10881)
  std::vector<uint32_t> right((size_t)NN, 0);
  std::vector<unsigned long long> nodeOff(ntree + 1, 0), leafOff(ntree + 1, 0);
  {
    int64_t q = 0;
    for (int t = 0; t < ntree; t++) {
      const int64_t cnt = 2 * (int64_t)a->leafCount[t] - 1;
      if (q + cnt > NN) { set_error("forest arrays shorter than sum(2*leafCount-1)"); return RFSLAM_EINVAL; }
      nodeOff[t] = (unsigned long long)q; leafOff[t + 1] = leafOff[t] + (unsigned long long)a->leafCount[t];
      std::vector<int64_t> stack;
      for (int64_t k = 0; k < cnt; k++) {
        const int64_t at = q + k;
        if (a->treeID[at] != a->treeID[q]) { set_error("tree records are not contiguous"); return RFSLAM_EINVAL; }
        if (a->parmID[at] < 0 || a->parmID[at] > p) { set_error("parmID out of range"); return RFSLAM_EINVAL; }
        if (k > 0 && a->parmID[at - 1] == 0) {          // previous record closed a left subtree
          if (stack.empty()) { set_error("malformed pre-order forest"); return RFSLAM_EINVAL; }
          right[(size_t)stack.back()] = (uint32_t)k; stack.pop_back();
        }
        if (a->parmID[at] != 0) stack.push_back(at);
      }
      if (!stack.empty()) { set_error("malformed pre-order forest"); return RFSLAM_EINVAL; }
      q += cnt;
    }
    nodeOff[ntree] = (unsigned long long)q;
  }
  try {
    if (a->device >= 0) RF_CUDA(cudaSetDevice(a->device));
    const size_t NL = (size_t)leafOff[ntree];
    Buf b_noff, b_loff, b_rec, b_clas, b_fx, b_ens, b_den, b_memb;
    std::vector<PredRec> recs((size_t)NN);
    for (int64_t q = 0; q < NN; q++) {
      const uint32_t pm = (uint32_t)a->parmID[q];
      recs[(size_t)q] = PredRec{pm ? a->contPT[q] : 0.0, pm, pm ? right[(size_t)q] : (uint32_t)a->nodeID[q]};
    }
    RF_CUDA(cudaMemcpy(b_noff.alloc((ntree + 1) * 8), nodeOff.data(), (ntree + 1) * 8, cudaMemcpyHostToDevice));
    RF_CUDA(cudaMemcpy(b_loff.alloc((ntree + 1) * 8), leafOff.data(), (ntree + 1) * 8, cudaMemcpyHostToDevice));
    RF_CUDA(cudaMemcpy(b_rec.alloc((size_t)NN * sizeof(PredRec)), recs.data(), (size_t)NN * sizeof(PredRec), cudaMemcpyHostToDevice));
    RF_CUDA(cudaMemcpy(b_clas.alloc(NL * 8), a->tnCLAS, NL * 8, cudaMemcpyHostToDevice));
    RF_CUDA(cudaMemcpy(b_fx.alloc((size_t)p * m * 8), a->fx_data, (size_t)p * m * 8, cudaMemcpyHostToDevice));
    b_ens.alloc((size_t)2 * m * 8); b_den.alloc((size_t)m * 4);
    if (a->want_membership) b_memb.alloc((size_t)ntree * m * 4);
    cudaEvent_t e0, e1;
    RF_CUDA(cudaEventCreate(&e0)); RF_CUDA(cudaEventCreate(&e1));
    RF_CUDA(cudaEventRecord(e0));
    predict_kernel<<<(m + 127) / 128, 128>>>(ntree, m, (const unsigned long long*)b_noff.p, (const unsigned long long*)b_loff.p,
                                             (const PredRec*)b_rec.p, (const int32_t*)b_clas.p, (const double*)b_fx.p,
                                             (double*)b_ens.p, (uint32_t*)b_den.p, (int32_t*)b_memb.p);
    if (a->finalize_ensembles) finalize_pred_kernel<<<(m + 255) / 256, 256>>>((double*)b_ens.p, (const uint32_t*)b_den.p, m);
    RF_CUDA(cudaEventRecord(e1));
    RF_CUDA(cudaGetLastError());
    RF_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    RF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    out->m = m; out->ntree = ntree; out->kernel_ms = ms;
    out->allEnsbCLS = (double*)host_alloc((size_t)2 * m * 8);
    out->allEnsbDen = (uint32_t*)host_alloc((size_t)m * 4);
    RF_CUDA(cudaMemcpy(out->allEnsbCLS, b_ens.p, (size_t)2 * m * 8, cudaMemcpyDeviceToHost));
    RF_CUDA(cudaMemcpy(out->allEnsbDen, b_den.p, (size_t)m * 4, cudaMemcpyDeviceToHost));
    if (a->want_membership) {
      out->nodeMembership = (int32_t*)host_alloc((size_t)ntree * m * 4);
      RF_CUDA(cudaMemcpy(out->nodeMembership, b_memb.p, (size_t)ntree * m * 4, cudaMemcpyDeviceToHost));
    }
  } catch (CudaFail f) {
    rfslam_b200_free_predict(out);
    return f.code;
  }
  return RFSLAM_OK;
}

}  // namespace rfslam
