// cpiu.cu -- CPIU construction on the device: the step immediately before the forest path.
//
// Replaces `cpiu.fc` / `event` of utilities/cpiu.R:361-668 (one dplyr group per person, a data frame replicated per
// interval) and the four Rcpp helpers of utilities/cpiu.cpp:27-249 it calls.  A person's follow-up [0, t_outcome] is
// cut into ceil(t_outcome / (interval * unit_norm)) counting-process information units; every unit carries its interval
// number, bounds, risk time (the `rt` and `int.n` columns the Poisson rules stratify and offset by) and the death
// indicator, plus derived time-varying covariates: event indicators (i.*), counts in this / in previous intervals
// (ni.* / n.*), time since the last event (t.*) and carried-forward measurements (c.*).
//
// Layout: persons are independent segments.  K1 counts the units per person, a CUB exclusive scan turns the counts
// into row offsets, K2 (one thread per person) writes the interval table of its rows, K3 (one thread per person and
// derived variable) replays the helper's own loop over that person's intervals -- the loops carry state from one
// interval to the next (running counts, last event, age of a measurement), so a person's intervals are walked in
// order; at <= a few dozen intervals and <= a few dozen event slots per person the walk is a few hundred operations,
// and there are persons x variables of them.  Output columns are contiguous per variable ([n_vars][rows]): a thread
// writes its person's consecutive rows of one column.  HBM-bound: ~ (6 + n_vars) * 8 B written per unit.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstring>
#include <vector>

#include "internal.h"

namespace rfslam {
namespace {

constexpr int kMaxSlots = 64;                 // event / measurement slots per person and variable (sorted in registers / local memory)

__global__ void cpiu_count_kernel(const double* tOutcome, int persons, double imod, uint32_t* count, int* bad) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= persons) return;
  const double t = tOutcome[q];
  if (!(t > 0.0) || !(t / imod < 1.0e6)) { atomicExch(bad, 1); count[q] = 0; return; }
  count[q] = (uint32_t)ceil(t / imod);                                       // cpiu.R:476
}

__global__ void cpiu_base_kernel(const double* tOutcome, const double* tDeath, const int32_t* death, int persons,
                                 double interval, double unitNorm, const uint32_t* rowOff,
                                 int32_t* person, int32_t* intN, double* tStart, double* tEnd, double* rt, int32_t* iDeath) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= persons) return;
  const double imod = interval * unitNorm, to = tOutcome[q], td = tDeath[q];
  const uint32_t r0 = rowOff[q], ni = rowOff[q + 1] - r0;
  for (uint32_t i = 0; i < ni; i++) {
    const double ts = imod * (double)i;                                      // interval.mod * (int.n - 1), cpiu.R:568
    const bool inside = ts + imod < to;
    person[r0 + i] = q; intN[r0 + i] = (int32_t)i + 1;
    tStart[r0 + i] = ts;
    tEnd[r0 + i] = inside ? ts + imod : to;                                  // cpiu.R:569-571
    rt[r0 + i] = inside ? interval : (to - ts) / unitNorm;                   // cpiu.R:572-574
    iDeath[r0 + i] = (ts + imod < td) ? 0 : death[q];                        // cpiu.R:575-577
  }
}

struct VarDesc { int kind, width; const double* times; const double* values; };

// one thread per (person, variable): the helper's loop over the person's intervals (cpiu.cpp:27-249)
__global__ void cpiu_vars_kernel(const VarDesc* vars, int nVars, int persons, const double* tOutcome, double imod,
                                 const uint32_t* rowOff, size_t rows, int kToPersist, double na, double* out) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)persons * nVars) return;
  const int v = (int)(g / persons), q = (int)(g % persons);                 // adjacent threads = adjacent persons of one variable
  const VarDesc d = vars[v];
  const double to = tOutcome[q];
  const uint32_t r0 = rowOff[q], ni = rowOff[q + 1] - r0;
  double* col = out + (size_t)v * rows + r0;
  const double* tm = d.times + (size_t)q * d.width;
  const int w = d.width;
  auto ts_of = [&](uint32_t i) { return imod * (double)i; };
  auto te_of = [&](uint32_t i) { const double ts = imod * (double)i; return (ts + imod < to) ? ts + imod : to; };
  if (d.kind == RFSLAM_CPIU_INDICATOR) {                                     // cpiu.cpp:27-50
    for (uint32_t i = 0; i < ni; i++) {
      const double ts = ts_of(i), te = te_of(i);
      int hit = 0;
      for (int j = 0; j < w; j++) { const double e = tm[j]; if (e > ts && e <= te) hit = 1; }   // (NaN slots never satisfy it, cpiu.R:583-584)
      col[i] = (double)hit;
    }
  } else if (d.kind == RFSLAM_CPIU_COUNT_PREV || d.kind == RFSLAM_CPIU_COUNT_CUR) {   // cpiu.cpp:75-113, flag 1 / flag 0
    int count = 0;
    for (uint32_t i = 0; i < ni; i++) {
      const double ts = ts_of(i), te = te_of(i);
      if (d.kind == RFSLAM_CPIU_COUNT_CUR) count = 0; else col[i] = (double)count;
      for (int j = 0; j < w; j++) { const double e = tm[j]; if (e > ts && e <= te) count++; }
      if (d.kind == RFSLAM_CPIU_COUNT_CUR) col[i] = (double)count;
    }
  } else if (d.kind == RFSLAM_CPIU_ELAPSED) {                                // cpiu.cpp:129-165
    double last = 0.0;
    for (uint32_t i = 0; i < ni; i++) {
      const double ts = ts_of(i), te = te_of(i);
      bool toggle = false;
      for (int j = 0; j < w; j++) { const double e = tm[j]; if (e > ts && e <= te) { last = e; toggle = true; } }
      const double r = toggle ? 0.0 : te - last;
      col[i] = (r == -1.0) ? na : r;                                         // mod.df[mod.df == -1] <- NA, cpiu.R:642
    }
  } else {                                                                   // cpiu.cpp:186-249 after cvars.sort (cpiu.R:488-494)
    const double* vl = d.values + (size_t)q * d.width;
    double st[kMaxSlots], sv[kMaxSlots];
    int m = 0;
    for (int j = 0; j < w; j++) {                                            // stable insertion sort by time, missing slots dropped
      const double t = tm[j];
      if (t != t) continue;
      int k = m++;
      while (k > 0 && st[k - 1] > t) { st[k] = st[k - 1]; sv[k] = sv[k - 1]; k--; }
      st[k] = t; sv[k] = vl[j];
    }
    const int keep = (kToPersist == -1) ? (int)ni : kToPersist;
    int jStart = 0, age = 0;
    double cur = m > 0 ? sv[0] : na;
    if (ni > 0) col[0] = (cur == -1.0) ? na : cur;
    for (uint32_t i = 0; i + 1 < ni; i++) {
      const double ts = ts_of(i), te = te_of(i);
      bool changed = false; double next = cur;
      for (int j = jStart; j < m; j++)
        if (st[j] > ts && st[j] <= te) { next = sv[j]; changed = true; jStart = j; age = 0; }
      if (!changed) { next = (age + 1 >= keep) ? -1.0 : cur; age++; }
      cur = next;
      col[i + 1] = (cur == -1.0) ? na : cur;
    }
  }
}

template <typename T> struct DBuf {
  T* p = nullptr;
  ~DBuf() { if (p) cudaFree(p); }
  void alloc(size_t n) { RF_CUDA(cudaMalloc((void**)&p, std::max<size_t>(n, 1) * sizeof(T))); }
};

template <typename T> T* fetch(const T* dev, size_t n) {
  T* h = (T*)host_alloc(std::max<size_t>(n, 1) * sizeof(T));
  if (!h) { set_error("out of host memory"); throw CudaFail{RFSLAM_ENOMEM}; }
  if (n) RF_CUDA(cudaMemcpy(h, dev, n * sizeof(T), cudaMemcpyDeviceToHost));
  return h;
}

}  // namespace

int cpiu_build(const rfslam_cpiu_args* a, rfslam_cpiu_out* out) {
  memset(out, 0, sizeof(*out));
  if (!a || a->persons < 1 || !a->t_outcome || !a->t_death || !a->death) { set_error("cpiu: persons, t_outcome, t_death and death are needed"); return RFSLAM_EINVAL; }
  if (!(a->interval > 0.0) || !(a->unit_norm > 0.0)) { set_error("cpiu: interval and unit_norm must be > 0"); return RFSLAM_EINVAL; }
  if (a->n_vars < 0 || (a->n_vars > 0 && (!a->var_kind || !a->var_width || !a->var_times))) { set_error("cpiu: variable descriptors missing"); return RFSLAM_EINVAL; }
  for (int v = 0; v < a->n_vars; v++) {
    if (a->var_kind[v] < RFSLAM_CPIU_INDICATOR || a->var_kind[v] > RFSLAM_CPIU_CONTINUOUS) { set_error("cpiu: variable %d has kind %d", v, a->var_kind[v]); return RFSLAM_EINVAL; }
    if (a->var_width[v] < 1 || a->var_width[v] > kMaxSlots) { set_error("cpiu: variable %d has %d slots per person (1..%d)", v, a->var_width[v], kMaxSlots); return RFSLAM_ENOSUP; }
    if (!a->var_times[v] || (a->var_kind[v] == RFSLAM_CPIU_CONTINUOUS && (!a->var_values || !a->var_values[v]))) { set_error("cpiu: variable %d: times / values missing", v); return RFSLAM_EINVAL; }
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: librfslam_b200 has no CPU fallback"); return RFSLAM_ENODEV; }
  try {
    if (a->device >= 0) RF_CUDA(cudaSetDevice(a->device));
    const int P = a->persons, V = a->n_vars;
    const double imod = a->interval * a->unit_norm;
    DBuf<double> d_to, d_td; DBuf<int32_t> d_death; DBuf<uint32_t> d_cnt, d_off; DBuf<int> d_bad; DBuf<unsigned char> d_tmp;
    d_to.alloc(P); d_td.alloc(P); d_death.alloc(P); d_cnt.alloc((size_t)P + 1); d_off.alloc((size_t)P + 1); d_bad.alloc(1);
    RF_CUDA(cudaMemcpy(d_to.p, a->t_outcome, (size_t)P * 8, cudaMemcpyHostToDevice));
    RF_CUDA(cudaMemcpy(d_td.p, a->t_death, (size_t)P * 8, cudaMemcpyHostToDevice));
    RF_CUDA(cudaMemcpy(d_death.p, a->death, (size_t)P * 4, cudaMemcpyHostToDevice));
    std::vector<DBuf<double>> d_times(V), d_values(V);
    std::vector<VarDesc> desc(V);
    for (int v = 0; v < V; v++) {
      const size_t cells = (size_t)P * a->var_width[v];
      d_times[v].alloc(cells);
      RF_CUDA(cudaMemcpy(d_times[v].p, a->var_times[v], cells * 8, cudaMemcpyHostToDevice));
      desc[v] = VarDesc{a->var_kind[v], a->var_width[v], d_times[v].p, nullptr};
      if (a->var_kind[v] == RFSLAM_CPIU_CONTINUOUS) {
        d_values[v].alloc(cells);
        RF_CUDA(cudaMemcpy(d_values[v].p, a->var_values[v], cells * 8, cudaMemcpyHostToDevice));
        desc[v].values = d_values[v].p;
      }
    }
    DBuf<VarDesc> d_desc; d_desc.alloc(V);
    if (V) RF_CUDA(cudaMemcpy(d_desc.p, desc.data(), (size_t)V * sizeof(VarDesc), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    RF_CUDA(cudaEventCreate(&e0)); RF_CUDA(cudaEventCreate(&e1));
    RF_CUDA(cudaMemset(d_bad.p, 0, 4));
    RF_CUDA(cudaMemset(d_cnt.p, 0, ((size_t)P + 1) * 4));
    RF_CUDA(cudaEventRecord(e0));
    cpiu_count_kernel<<<(P + 255) / 256, 256>>>(d_to.p, P, imod, d_cnt.p, d_bad.p);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, d_cnt.p, d_off.p, P + 1);
    d_tmp.alloc(tb);
    RF_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp.p, tb, d_cnt.p, d_off.p, P + 1));
    uint32_t rows32 = 0; int bad = 0;
    RF_CUDA(cudaMemcpy(&rows32, d_off.p + P, 4, cudaMemcpyDeviceToHost));
    RF_CUDA(cudaMemcpy(&bad, d_bad.p, 4, cudaMemcpyDeviceToHost));
    if (bad) { cudaEventDestroy(e0); cudaEventDestroy(e1); set_error("cpiu: t_outcome must be > 0 and span fewer than 10^6 intervals"); return RFSLAM_EINVAL; }
    const size_t rows = rows32;
    DBuf<int32_t> d_person, d_intn, d_idth; DBuf<double> d_ts, d_te, d_rt, d_vars;
    d_person.alloc(rows); d_intn.alloc(rows); d_idth.alloc(rows); d_ts.alloc(rows); d_te.alloc(rows); d_rt.alloc(rows); d_vars.alloc(rows * (size_t)std::max(V, 1));
    cpiu_base_kernel<<<(P + 127) / 128, 128>>>(d_to.p, d_td.p, d_death.p, P, a->interval, a->unit_norm, d_off.p,
                                               d_person.p, d_intn.p, d_ts.p, d_te.p, d_rt.p, d_idth.p);
    if (V) {
      const long long threads = (long long)P * V;
      cpiu_vars_kernel<<<(unsigned)((threads + 127) / 128), 128>>>(d_desc.p, V, P, d_to.p, imod, d_off.p, rows, a->k_to_persist, rfslam_na_real(), d_vars.p);
    }
    RF_CUDA(cudaEventRecord(e1));
    RF_CUDA(cudaGetLastError());
    RF_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    RF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    out->rows = (int64_t)rows; out->n_vars = V; out->kernel_ms = ms; out->kernel_launches = V ? 4 : 3;
    out->person = fetch(d_person.p, rows); out->int_n = fetch(d_intn.p, rows); out->i_death = fetch(d_idth.p, rows);
    out->t_start = fetch(d_ts.p, rows); out->t_end = fetch(d_te.p, rows); out->rt = fetch(d_rt.p, rows);
    out->vars = fetch(d_vars.p, rows * (size_t)V);
  } catch (CudaFail f) {
    rfslam_b200_free_cpiu(out);
    return f.code;
  }
  return RFSLAM_OK;
}

}  // namespace rfslam

extern "C" int rfslam_b200_cpiu_build(const rfslam_cpiu_args* args, rfslam_cpiu_out* out) {
  if (!out) { rfslam::set_error("null output"); return RFSLAM_EINVAL; }
  return rfslam::cpiu_build(args, out);
}

extern "C" void rfslam_b200_free_cpiu(rfslam_cpiu_out* o) {
  if (!o) return;
  void* arrays[] = {o->person, o->int_n, o->i_death, o->t_start, o->t_end, o->rt, o->vars};
  for (void* p : arrays) rfslam::host_free(p);
  memset(o, 0, sizeof(*o));
}
