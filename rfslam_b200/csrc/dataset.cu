// dataset.cu -- upload the CPIU table once and lay it out in HBM for the tree kernels.
//
// What the reference does per NODE and per CANDIDATE (gather the column, indexx() argsort, dedupe
// into splitVector: rf.c:14946-14996) is hoisted here to once per DATASET: every covariate column
// is rank-coded (code = index of the value among the column's sorted distinct values).  Cut values
// returned by the path are looked up in uniq[], so contPT is still the exact data value
// (rf.c:15539).  CUB's radix sort is used for this one-off preparation step only; nothing in the
// per-tree hot path calls a library.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstring>

#include "internal.h"

namespace rfslam {

constexpr int kMaxFactorLevelsHost = 9;   // = kMaxFactorLevels of the tree kernel (grow_kernel.cuh)

namespace {

__global__ void iota_kernel(uint32_t* v, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (uint32_t)i;
}

__global__ void strat_keys_kernel(const int32_t* strat, uint32_t* keys, int n, int* bad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    int s = strat[i];
    if (s < 1) atomicExch(bad, 1);
    keys[i] = (uint32_t)s;
  }
}

__global__ void invert_kernel(const uint32_t* perm, uint32_t* inv, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) inv[perm[i]] = (uint32_t)i;
}

// sb[k] = first internal row whose stratum is >= k  (k = 0..K+1)
__global__ void strat_bounds_kernel(const uint32_t* sorted_strat, int n, int K, uint32_t* sb) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > K + 1) return;
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (sorted_strat[mid] < (uint32_t)k) lo = mid + 1; else hi = mid;
  }
  sb[k] = (uint32_t)lo;
}

__global__ void row_arrays_kernel(const uint32_t* perm, const double* y, const double* rt_in, const uint32_t* sorted_strat,
                                  uint8_t* ev, double* rt, uint16_t* strat, int n, int* bad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    uint32_t r = perm[i];
    double yy = y[r];
    if (!(yy == 1.0 || yy == 2.0)) atomicExch(bad, 2);       // sc.c:676-687: the plugin exits on other codes
    ev[i] = (yy == 2.0) ? 1 : 0;
    double t = rt_in[r];
    if (!(t == t)) atomicExch(bad, 3);
    rt[i] = t;
    strat[i] = (uint16_t)sorted_strat[i];
  }
}

// keys = x + 0.0 (folds -0.0 onto +0.0: the reference compares doubles, rf.c:14992), NaN flagged
__global__ void column_keys_kernel(const double* x, double* keys, int n, int* bad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    double v = x[i];
    if (!(v == v)) atomicExch(bad, 4);
    keys[i] = v + 0.0;
  }
}

__global__ void diff_flags_kernel(const double* sorted, uint32_t* flag, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = (i > 0 && sorted[i] != sorted[i - 1]) ? 1u : 0u;
}

template <typename CodeT>
__global__ void scatter_codes_kernel(const double* sorted, const uint32_t* sorted_row, const uint32_t* rank,
                                     const uint32_t* inv, CodeT* codes, double* uniq, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    uint32_t r = rank[i];
    codes[inv[sorted_row[i]]] = (CodeT)r;
    if (i == 0 || sorted[i] != sorted[i - 1]) uniq[r] = sorted[i];
  }
}

// row-major mirror: rowcodes[row][c] = codes[c][row]; one thread block transposes a 32-row x 32-column tile
__global__ void row_major_kernel(const void* const* codes, int p, int n, int pitch, uint8_t* rowcodes) {
  __shared__ uint8_t tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;               // 32 x 8
  for (int k = ty; k < 32; k += 8) {
    const int c = c0 + k, r = r0 + tx;
    tile[k][tx] = (c < p && r < n) ? ((const uint8_t*)codes[c])[r] : 0;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int r = r0 + k, c = c0 + tx;
    if (r < n && c < pitch) rowcodes[(size_t)r * pitch + c] = tile[tx][k];
  }
}
__global__ void row_rec_kernel(const double* rt, const uint8_t* ev, const uint16_t* strat, RowRec* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { RowRec r; r.rt = rt[i]; r.es = (ev[i] ? 1u : 0u) | (((uint32_t)strat[i] - 1u) << 8); r.pad = 0u; out[i] = r; }
}

// every device array of a dataset comes from its pool: rfslam_b200_grow seeds the pool with the blocks of
// the previous call (api.cu), so a second call of the same shape makes no cudaMalloc / cudaFree at all
template <typename T>
T* dalloc(rfslam_dataset* ds, size_t count, bool resident = true) {
  const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
  void* p = ds->pool.get(bytes);
  if (!p) { set_error("out of device memory (%zu bytes)", bytes); throw CudaFail{RFSLAM_ENOMEM}; }
  if (resident) ds->bytes += (int64_t)(count * sizeof(T));
  return (T*)p;
}

}  // namespace

void dataset_free(rfslam_dataset* ds) {
  if (!ds) return;
  cudaSetDevice(ds->device);
  ds->pool.release_all();
  for (void* p : ds->allocs) cudaFree(p);
  delete ds;
}

int dataset_build(const rfslam_grow_args* a, rfslam_dataset** out, WsPool* seed) {
  *out = nullptr;
  const int n = a->observation_size, p = a->x_size;
  if (n < 2 || p < 1) { set_error("observationSize must be >= 2 and xSize >= 1"); return RFSLAM_EINVAL; }
  if ((uint64_t)n > (uint64_t)kRowMask) {
    set_error("observationSize %d exceeds the %u rows an in-bag entry can address", n, kRowMask);
    return RFSLAM_ENOSUP;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device: librfslam_b200 has no CPU fallback");
    return RFSLAM_ENODEV;
  }
  rfslam_dataset* ds = new rfslam_dataset();
  try {
    if (a->device >= 0) RF_CUDA(cudaSetDevice(a->device));
    RF_CUDA(cudaGetDevice(&ds->device));
    if (seed) { for (auto& b : seed->blks) { b.used = false; ds->pool.blks.push_back(b); } seed->blks.clear(); }
    ds->n = n; ds->p = p;
    const int T = 256, G = (n + T - 1) / T;

    int* d_bad = dalloc<int>(ds, 1, false);
    RF_CUDA(cudaMemset(d_bad, 0, sizeof(int)));

    // ---- transient buffers -------------------------------------------------------------
    int32_t* d_strat_in = dalloc<int32_t>(ds, n, false);
    double* d_col = dalloc<double>(ds, n, false); double* d_keys = dalloc<double>(ds, n, false);
    double* d_keys_sorted = dalloc<double>(ds, n, false); double* d_y = dalloc<double>(ds, n, false);
    double* d_rt_in = dalloc<double>(ds, n, false);
    uint32_t* d_iota = dalloc<uint32_t>(ds, n, false); uint32_t* d_rows_sorted = dalloc<uint32_t>(ds, n, false);
    uint32_t* d_u32a = dalloc<uint32_t>(ds, n, false); uint32_t* d_u32b = dalloc<uint32_t>(ds, n, false);
    std::vector<void*> transient = {d_strat_in, d_col, d_keys, d_keys_sorted, d_y, d_rt_in, d_iota, d_rows_sorted, d_u32a, d_u32b, d_bad};
    auto free_transient = [&]() { RF_CUDA(cudaDeviceSynchronize()); for (void* q : transient) ds->pool.put(q); transient.clear(); };

    size_t tmp_bytes = 0, tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, n);
    tmp_bytes = std::max(tmp_bytes, tb);
    cub::DeviceRadixSort::SortPairs(nullptr, tb, (const double*)nullptr, (double*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, n);
    tmp_bytes = std::max(tmp_bytes, tb);
    cub::DeviceScan::InclusiveSum(nullptr, tb, (const uint32_t*)nullptr, (uint32_t*)nullptr, n);
    tmp_bytes = std::max(tmp_bytes, tb);
    void* d_tmp = dalloc<unsigned char>(ds, tmp_bytes, false);
    transient.push_back(d_tmp);

    try {
      // ---- internal row order: stable sort by stratum ------------------------------------
      RF_CUDA(cudaMemcpy(d_strat_in, a->stratifier, (size_t)n * 4, cudaMemcpyHostToDevice));
      RF_CUDA(cudaMemcpy(d_y, a->r_data, (size_t)n * 8, cudaMemcpyHostToDevice));
      RF_CUDA(cudaMemcpy(d_rt_in, a->risk_time, (size_t)n * 8, cudaMemcpyHostToDevice));
      iota_kernel<<<G, T>>>(d_iota, n);
      strat_keys_kernel<<<G, T>>>(d_strat_in, d_u32a, n, d_bad);
      uint32_t* perm = dalloc<uint32_t>(ds, n);
      uint32_t* inv = dalloc<uint32_t>(ds, n);
      tb = tmp_bytes;
      RF_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tb, d_u32a, d_u32b, d_iota, perm, n));   // stable
      invert_kernel<<<G, T>>>(perm, inv, n);
      uint32_t maxs = 0;
      RF_CUDA(cudaMemcpy(&maxs, d_u32b + (n - 1), 4, cudaMemcpyDeviceToHost));
      int bad = 0;
      RF_CUDA(cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost));
      if (bad == 1) { set_error("stratifier must be >= 1"); throw CudaFail{RFSLAM_EINVAL}; }
      if (maxs > (uint32_t)kMaxStrata) {
        set_error("largest stratum %u exceeds the %d strata the device path holds in shared memory", maxs, kMaxStrata);
        throw CudaFail{RFSLAM_ENOSUP};
      }
      ds->K = (int)maxs;
      uint32_t* sb = dalloc<uint32_t>(ds, ds->K + 2);
      strat_bounds_kernel<<<(ds->K + 2 + 63) / 64, 64>>>(d_u32b, n, ds->K, sb);
      uint8_t* ev = dalloc<uint8_t>(ds, n);
      double* rt = dalloc<double>(ds, n);
      uint16_t* strat = dalloc<uint16_t>(ds, n);
      row_arrays_kernel<<<G, T>>>(perm, d_y, d_rt_in, d_u32b, ev, rt, strat, n, d_bad);

      // ---- rank-code every covariate column ----------------------------------------------
      ds->D.resize(p); ds->width.resize(p); ds->codes_h.resize(p); ds->uniq_h.resize(p); ds->ftype.assign(p, 0);
      for (int c = 0; c < p; c++) {
        const bool isFactor = a->x_type && a->x_type[c] == 'C';
        if (a->x_type && !(a->x_type[c] == 'R' || a->x_type[c] == 'I' || a->x_type[c] == 'B' || isFactor)) {
          set_error("covariate %d has type '%c': only real / integer / boolean / unordered-factor covariates are on the accelerated path", c + 1, a->x_type[c]);
          throw CudaFail{RFSLAM_ENOSUP};
        }
        if (isFactor) {                                  // levels are the integers 1 .. xLevels (one MWCP word: rf.c:1421)
          if (!a->x_levels || a->x_levels[c] < 1 || a->x_levels[c] > 32) {
            set_error("factor covariate %d: xLevels must be in [1, 32] (one MWCP word)", c + 1);
            throw CudaFail{a->x_levels && a->x_levels[c] > 32 ? RFSLAM_ENOSUP : RFSLAM_EINVAL};
          }
          const double* col = a->x_data + (size_t)c * n;
          for (int i = 0; i < n; i++) {
            const double v = col[i];
            if (!(v >= 1.0 && v <= (double)a->x_levels[c] && v == (double)(int)v)) {
              set_error("factor covariate %d, row %d: level %g is not an integer in [1, %d]", c + 1, i + 1, v, a->x_levels[c]);
              throw CudaFail{RFSLAM_EINVAL};
            }
          }
        }
        RF_CUDA(cudaMemcpy(d_col, a->x_data + (size_t)c * n, (size_t)n * 8, cudaMemcpyHostToDevice));
        column_keys_kernel<<<G, T>>>(d_col, d_keys, n, d_bad);
        tb = tmp_bytes;
        RF_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tb, d_keys, d_keys_sorted, d_iota, d_rows_sorted, n));
        diff_flags_kernel<<<G, T>>>(d_keys_sorted, d_u32a, n);
        tb = tmp_bytes;
        RF_CUDA(cub::DeviceScan::InclusiveSum(d_tmp, tb, d_u32a, d_u32b, n));
        uint32_t last = 0;
        RF_CUDA(cudaMemcpy(&last, d_u32b + (n - 1), 4, cudaMemcpyDeviceToHost));
        int D = (int)last + 1;
        ds->D[c] = D;
        ds->ftype[c] = isFactor ? 1 : 0;
        if (isFactor) {
          // a factor's cuts are level SUBSETS: 2^(D-1) - 1 of them share one candidate's per-cut accumulators (hist_bins)
          if (D > kMaxFactorLevelsHost) {
            set_error("factor covariate %d takes %d distinct levels: unordered-factor splits are accelerated for at most %d levels present", c + 1, D, kMaxFactorLevelsHost);
            throw CudaFail{RFSLAM_ENOSUP};
          }
          ds->has_factor = true;
          while (ds->hist_bins < (1 << (D > 1 ? D - 1 : 0))) ds->hist_bins *= 2;
        }
        ds->max_D = std::max(ds->max_D, D);
        if (D <= kHistBins) while (ds->hist_bins < D) ds->hist_bins *= 2;
        else ds->hist_bins = kHistBins;            // small nodes re-rank such a column locally: up to 256 ranks per node
        double* uniq = dalloc<double>(ds, D);
        ds->uniq_h[c] = uniq;
        if (D <= 256) {
          ds->width[c] = 1;
          uint8_t* codes = dalloc<uint8_t>(ds, n);
          ds->codes_h[c] = codes;
          scatter_codes_kernel<uint8_t><<<G, T>>>(d_keys_sorted, d_rows_sorted, d_u32b, inv, codes, uniq, n);
        } else if (D <= 65536) {
          ds->width[c] = 2;
          uint16_t* codes = dalloc<uint16_t>(ds, n);
          ds->codes_h[c] = codes;
          scatter_codes_kernel<uint16_t><<<G, T>>>(d_keys_sorted, d_rows_sorted, d_u32b, inv, codes, uniq, n);
        } else {
          ds->width[c] = 4;
          uint32_t* codes = dalloc<uint32_t>(ds, n);
          ds->codes_h[c] = codes;
          scatter_codes_kernel<uint32_t><<<G, T>>>(d_keys_sorted, d_rows_sorted, d_u32b, inv, codes, uniq, n);
        }
      }
      RF_CUDA(cudaDeviceSynchronize());
      RF_CUDA(cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost));
      if (bad == 2) { set_error("response codes must be 1.0 or 2.0 (two-class CPIU outcome, splitCustom.c:676-687)"); throw CudaFail{RFSLAM_EINVAL}; }
      if (bad == 3 || bad == 4) { set_error("missing values (NA/NaN) are outside the accelerated path"); throw CudaFail{RFSLAM_ENOSUP}; }

      // ---- device-side tables of per-column pointers --------------------------------------
      const void** d_codes = (const void**)dalloc<void*>(ds, p);
      const double** d_uniq = (const double**)dalloc<double*>(ds, p);
      uint8_t* d_width = dalloc<uint8_t>(ds, p);
      int* d_D = dalloc<int>(ds, p);
      RF_CUDA(cudaMemcpy(d_codes, ds->codes_h.data(), (size_t)p * sizeof(void*), cudaMemcpyHostToDevice));
      RF_CUDA(cudaMemcpy(d_uniq, ds->uniq_h.data(), (size_t)p * sizeof(double*), cudaMemcpyHostToDevice));
      RF_CUDA(cudaMemcpy(d_width, ds->width.data(), (size_t)p, cudaMemcpyHostToDevice));
      RF_CUDA(cudaMemcpy(d_D, ds->D.data(), (size_t)p * sizeof(int), cudaMemcpyHostToDevice));
      uint8_t* d_ftype = nullptr;
      if (ds->has_factor) {
        d_ftype = dalloc<uint8_t>(ds, p);
        RF_CUDA(cudaMemcpy(d_ftype, ds->ftype.data(), (size_t)p, cudaMemcpyHostToDevice));
      }

      // Row-major mirror (see DevData): only when every column is byte-coded, and only when the column-major table
      // does not fit the 126 MB L2 anyway.  A table that does (configs[1]: 69 MB) is served from L2 whatever the
      // access pattern, and a mirror would push the working set out of it (measured: 1.42 s -> 1.53 s per 500 trees).
      uint8_t* rowcodes = nullptr; RowRec* rowrec = nullptr; int rowPitch = 0;
      bool wantRow = (size_t)n * (size_t)p > ((size_t)96 << 20);
      if (const char* e = getenv("RFSLAM_ROWMAJOR")) wantRow = atoi(e) != 0;
      rowrec = dalloc<RowRec>(ds, n);                              // the 16-byte row records are always built (one load instead of three)
      row_rec_kernel<<<G, T>>>(rt, ev, strat, rowrec, n);
      if (ds->max_D <= 256 && wantRow && !getenv("RFSLAM_NO_ROWMAJOR")) {
        rowPitch = (p + 15) & ~15;
        rowcodes = dalloc<uint8_t>(ds, (size_t)n * rowPitch);
        dim3 g((unsigned)((n + 31) / 32), (unsigned)((rowPitch + 31) / 32)), b(32, 8);
        row_major_kernel<<<g, b>>>((const void* const*)d_codes, p, n, rowPitch, rowcodes);
      }
      RF_CUDA(cudaGetLastError());

      // dominant risk time: the modal value of a sample (any value is correct; the mode is fastest)
      double rt_mode = a->risk_time[0];
      for (int i = 0; i < n; i++) {
        const double t = a->risk_time[i];
        if (t < 0.0) { set_error("negative risk time at row %d", i + 1); throw CudaFail{RFSLAM_EINVAL}; }
        if (t > ds->max_rt) ds->max_rt = t;
      }
      {
        std::vector<double> smp;
        const int step = std::max(1, n / 65536);
        for (int i = 0; i < n; i += step) smp.push_back(a->risk_time[i]);
        std::sort(smp.begin(), smp.end());
        size_t bestLen = 0;
        for (size_t i = 0; i < smp.size();) {
          size_t j = i;
          while (j < smp.size() && smp[j] == smp[i]) j++;
          if (j - i > bestLen) { bestLen = j - i; rt_mode = smp[i]; }
          i = j;
        }
      }
      DevData& d = ds->dev;
      d.rt_mode = rt_mode;
      d.n = n; d.p = p; d.K = ds->K;
      d.perm = perm; d.inv = inv; d.sb = sb; d.ev = ev; d.rt = rt; d.strat = strat;
      d.codes = (const void* const*)d_codes; d.width = d_width; d.D = d_D; d.uniq = (const double* const*)d_uniq;
      d.rowcodes = rowcodes; d.rowPitch = rowPitch; d.rowrec = rowrec; d.ftype = d_ftype;
    } catch (...) {
      free_transient();
      throw;
    }
    free_transient();
  } catch (CudaFail f) {
    dataset_free(ds);
    return f.code;
  }
  *out = ds;
  return RFSLAM_OK;
}

}  // namespace rfslam
