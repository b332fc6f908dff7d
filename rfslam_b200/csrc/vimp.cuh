// vimp.cuh -- permutation importance on the device (getPermuteMembership rf.c:2106-2218, permute() rf.c:1974-2007,
// updateGenericVimpEnsemble rf.c:2317-2415, summarize / finalizeVimpPerformance rf.c:2416-2697).  Included by predict.cu
// (needs PredRec / rec_goes_left).
//
// One CTA per tree, trees time-sliced over the resident CTAs.  Chain C of a tree is one sequential stream over the
// covariates of interest, so warp 0 is the PRODUCER: it replays the chain (32 Park-Miller states per jump-ahead, the
// Bays-Durham table spread over the lanes) and builds the permutation of the tree's OOB rows for covariate v --
// permute() gives slot "k-th still-empty" the value i for i = m..1, which the reference finds by a linear walk
// (O(m^2)); here an occupancy bitmap with two levels of empty-slot counts in shared memory answers "k-th empty" in three
// warp-wide prefix steps.  Warps 1..7 are the CONSUMERS: while the producer works on covariate v they drop the OOB rows
// down the tree with covariate v - 1 replaced by its permuted value and either count the votes of the tree's own
// estimate (per-tree importance) or add the leaf proportions to the covariate's perturbed ensemble.
#pragma once

namespace rfslam {
namespace {

constexpr int kVimpThreads = 256;
constexpr int kVimpConsumers = kVimpThreads - 32;

struct __align__(16) VimpAcc { uint32_t cum[2], ok[2]; double brier[2]; };     // one (tree, covariate) pass over the OOB rows

struct VimpArgs {
  int ntree, p, n, nIntr;
  const int64_t* nodeOff;        // by slot
  const int32_t* slotOfTree;     // tree id - 1 -> slot
  const int64_t* leafOff;        // by tree id - 1
  const PredRec* rec;
  const int32_t* tnCLAS;         // [leaf][2]
  const uint16_t* cnt; size_t cstride;   // in-bag counts [tree][row]
  const double* x;               // [p][n] the training table, original rows
  const double* y;               // [n]
  const int* seedC;              // [ntree]
  const int* intr;               // [nIntr] 0-based covariates
  int maxOob;                    // capacity of the per-CTA lists
  uint32_t* scratch;             // [gridDim.x][3][maxOob]: OOB rows, two permutation buffers
  int leob;                      // 1 = per-tree performance (acc), 0 = perturbed ensembles (vens / vden)
  int rfqMinority; double rfqThreshold;
  VimpAcc* acc;                  // [ntree][nIntr + 1]: slot 0 = the unperturbed tree
  double* vens; uint32_t* vden;  // [nIntr][2][n], [nIntr][n]
};

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += t; }
  return v;
}

// chain replay by one warp: 32 draws per refill (as bootstrap_draw_kernel, grow.cu)
struct WarpRan1 {
  int ivReg, iy, idum; float u; int pos;
  __device__ void seed(int seedValue, int lane, Ran1* tmp) {
    if (lane == 0) { ran1_seed(*tmp, seedValue); ran1_init(*tmp); }
    __syncwarp();
    ivReg = tmp->iv[lane]; iy = tmp->iy; idum = tmp->idum; pos = 32; u = 0.f;
    __syncwarp();
  }
  __device__ __forceinline__ float next(int lane) {
    if (pos == 32) {
      const int myIdum = ran1_mulmod(kJump[lane], (uint32_t)idum);
      int myIy = 0;
      for (int k = 0; k < 32; k++) {
        const int j = iy / 67108864;
        const int idk = __shfl_sync(0xffffffffu, myIdum, k);
        const int niy = __shfl_sync(0xffffffffu, ivReg, j);
        if (lane == j) ivReg = idk;
        iy = niy;
        if (lane == k) myIy = niy;
      }
      idum = __shfl_sync(0xffffffffu, myIdum, 31);
      u = ran1_float(myIy);
      pos = 0;
    }
    return __shfl_sync(0xffffffffu, u, pos++);
  }
};

__device__ __forceinline__ uint32_t vimp_walk(const VimpArgs& a, int64_t off, int row, int target, double shadow) {
  int64_t q = off;
  PredRec r = a.rec[q];
  while (r.parm != 0u) {
    const int c = (int)rec_cov(r.parm);
    const double xv = (c == target) ? shadow : a.x[(size_t)c * a.n + row];
    q = rec_goes_left(r, xv) ? q + 1 : off + r.link;
    r = a.rec[q];
  }
  return r.link;
}

__global__ void __launch_bounds__(kVimpThreads) vimp_oob_count_kernel(const uint16_t* cnt, size_t cstride, int n, int ntree, int* oobCount) {
  const int t = blockIdx.x;
  __shared__ int sh;
  if (threadIdx.x == 0) sh = 0;
  __syncthreads();
  int c = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) c += cnt[(size_t)t * cstride + i] == 0 ? 1 : 0;
  c = (int)warp_sum((unsigned long long)c);
  if ((threadIdx.x & 31) == 0) atomicAdd(&sh, c);
  __syncthreads();
  if (threadIdx.x == 0) oobCount[t] = sh;
}

// dynamic shared memory: occupancy bitmap [W0 words] | L1 empty counts per 32 words (u32) [G1] | L2 per 32 L1 groups [G2]
__global__ void __launch_bounds__(kVimpThreads) vimp_tree_kernel(VimpArgs a) {
  extern __shared__ uint32_t dsm[];
  __shared__ Ran1 rngTmp;
  __shared__ int warpBase[kVimpThreads / 32 + 1];
  __shared__ VimpAcc red[kVimpThreads / 32];
  __shared__ int sM;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int W0cap = (a.maxOob + 31) >> 5, G1cap = (W0cap + 31) >> 5;
  uint32_t* L0 = dsm; uint32_t* L1 = L0 + W0cap; uint32_t* L2 = L1 + G1cap;
  uint32_t* idx = a.scratch + (size_t)blockIdx.x * 3 * a.maxOob;
  uint32_t* permBuf[2] = {idx + a.maxOob, idx + 2 * (size_t)a.maxOob};

  for (int tree = blockIdx.x; tree < a.ntree; tree += gridDim.x) {
    const uint16_t* cnt = a.cnt + (size_t)tree * a.cstride;
    // ---- the tree's OOB rows, ascending (RF_oobMembershipIndex) ----
    int base = 0;
    for (int i0 = 0; i0 < a.n; i0 += kVimpThreads) {
      const int i = i0 + tid;
      const bool oob = i < a.n && cnt[i] == 0;
      const unsigned bal = __ballot_sync(0xffffffffu, oob);
      if (lane == 0) warpBase[w] = __popc(bal);
      __syncthreads();
      int before = base, total = base;
      for (int k = 0; k < kVimpThreads / 32; k++) { const int c = warpBase[k]; if (k < w) before += c; total += c; }
      if (oob) idx[before + __popc(bal & ((1u << lane) - 1u))] = (uint32_t)i;
      base = total;
      __syncthreads();
    }
    const int m = base;
    const int64_t off = a.nodeOff[a.slotOfTree[tree]];
    const int64_t leaf0 = a.leafOff[tree];
    WarpRan1 rng;
    if (w == 0) rng.seed(a.seedC[tree], lane, &rngTmp);
    __syncthreads();           // idx visible to all
    if (m == 0) {              // getVimpMembership does nothing for a tree without OOB rows: its slots stay NA
      for (int s = tid; s <= a.nIntr; s += kVimpThreads) { VimpAcc z{}; a.acc[(size_t)tree * (a.nIntr + 1) + s] = z; }
      continue;
    }
    const int W0 = (m + 31) >> 5, G1 = (W0 + 31) >> 5, G2 = (G1 + 31) >> 5;

    for (int s = 0; s <= a.nIntr; s++) {
      if (w == 0) {
        // ---------------- producer: the permutation of covariate s ----------------
        if (s < a.nIntr) {
          uint32_t* perm = permBuf[s & 1];
          for (int k = lane; k < W0; k += 32) L0[k] = (k == W0 - 1 && (m & 31)) ? ~((1u << (m & 31)) - 1u) : 0u;       // slots beyond m count as taken
          for (int g = lane; g < G1; g += 32) L1[g] = (uint32_t)(min(m, (g + 1) * 1024) - g * 1024);
          for (int h = lane; h < G2; h += 32) L2[h] = (uint32_t)(min(m, (h + 1) * 32768) - h * 32768);
          __syncwarp();
          for (int i = m; i > 0; i--) {
            uint32_t k = (uint32_t)ceil((double)rng.next(lane) * ((double)i * 1.0));      // rf.c:1998
            // level 2
            int h = 0;
            for (int b2 = 0; b2 < G2; b2 += 32) {
              const uint32_t c = (b2 + lane < G2) ? L2[b2 + lane] : 0u;
              const uint32_t inc = warp_incl_scan(c, lane);
              const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
              if (k <= tot) {
                const int f = __ffs(__ballot_sync(0xffffffffu, inc >= k)) - 1;
                k -= __shfl_sync(0xffffffffu, inc - c, f);
                h = b2 + f;
                break;
              }
              k -= tot;
            }
            // level 1
            int g;
            {
              const int gi = h * 32 + lane;
              const uint32_t c = (gi < G1) ? L1[gi] : 0u;
              const uint32_t inc = warp_incl_scan(c, lane);
              const int f = __ffs(__ballot_sync(0xffffffffu, inc >= k)) - 1;
              k -= __shfl_sync(0xffffffffu, inc - c, f);
              g = h * 32 + f;
            }
            // level 0
            int wi; uint32_t ww;
            {
              const int wj = g * 32 + lane;
              const uint32_t word = (wj < W0) ? L0[wj] : 0xffffffffu;
              const uint32_t c = 32u - (uint32_t)__popc(word);
              const uint32_t inc = warp_incl_scan(c, lane);
              const int f = __ffs(__ballot_sync(0xffffffffu, inc >= k)) - 1;
              k -= __shfl_sync(0xffffffffu, inc - c, f);
              wi = g * 32 + f;
              ww = __shfl_sync(0xffffffffu, word, f);
            }
            const uint32_t freeBits = ~ww;
            const bool mine = ((freeBits >> lane) & 1u) && (uint32_t)__popc(freeBits & ((1u << lane) - 1u)) == k - 1u;
            const int bit = __ffs(__ballot_sync(0xffffffffu, mine)) - 1;
            if (lane == 0) {
              L0[wi] = ww | (1u << bit); L1[g] -= 1u; L2[h] -= 1u;
              perm[wi * 32 + bit] = (uint32_t)i;                 // slot (1-based wi*32+bit+1) receives i
            }
            __syncwarp();
          }
          __threadfence_block();
        }
      } else {
        // ---------------- consumers: drop the OOB rows with covariate s - 1 permuted (s = 0: the tree as grown) ----------------
        const int v = s - 1;
        if (v >= 0 || a.leob) {
          const int target = v >= 0 ? a.intr[v] : -1;
          const uint32_t* perm = permBuf[(s + 1) & 1];
          uint32_t cum0 = 0, cum1 = 0, ok0 = 0, ok1 = 0; double b0 = 0.0, b1 = 0.0;
          for (int k = tid - 32; k < m; k += kVimpConsumers) {
            const int row = (int)idx[k];
            const double shadow = v >= 0 ? a.x[(size_t)target * a.n + idx[perm[k] - 1u]] : 0.0;     // rf.c:2185
            const uint32_t leaf = vimp_walk(a, off, row, target, shadow);
            const int64_t li = leaf0 + (int64_t)leaf - 1;
            const double c1 = (double)a.tnCLAS[2 * li], c2 = (double)a.tnCLAS[2 * li + 1];
            const double rc = c1 + c2;
            const double p1 = c1 / rc, p2 = c2 / rc;
            if (a.leob) {
              const int cls = (a.y[row] == 2.0) ? 1 : 0;
              int vote;
              if (a.rfqMinority < 0) vote = (p1 <= p2) ? 1 : 0;
              else vote = ((a.rfqMinority ? p2 : p1) >= a.rfqThreshold) ? a.rfqMinority : 1 - a.rfqMinority;
              if (cls) { cum1++; if (vote == cls) ok1++; } else { cum0++; if (vote == cls) ok0++; }
              const double d1 = (cls == 0 ? 1.0 : 0.0) - p1, d2 = (cls == 1 ? 1.0 : 0.0) - p2;
              b0 += d1 * d1; b1 += d2 * d2;
            } else {
              atomicAdd(&a.vens[((size_t)v * 2) * a.n + row], p1);
              atomicAdd(&a.vens[((size_t)v * 2 + 1) * a.n + row], p2);
              atomicAdd(&a.vden[(size_t)v * a.n + row], 1u);
            }
          }
          if (a.leob) {
            cum0 = (uint32_t)warp_sum((unsigned long long)cum0); cum1 = (uint32_t)warp_sum((unsigned long long)cum1);
            ok0 = (uint32_t)warp_sum((unsigned long long)ok0); ok1 = (uint32_t)warp_sum((unsigned long long)ok1);
            b0 = warp_sum(b0); b1 = warp_sum(b1);
            if (lane == 0) { red[w].cum[0] = cum0; red[w].cum[1] = cum1; red[w].ok[0] = ok0; red[w].ok[1] = ok1; red[w].brier[0] = b0; red[w].brier[1] = b1; }
            asm volatile("bar.sync 1, %0;" ::"n"(kVimpConsumers) : "memory");
            if (tid == 32) {
              VimpAcc t{};
              for (int k = 1; k < kVimpThreads / 32; k++) {
                t.cum[0] += red[k].cum[0]; t.cum[1] += red[k].cum[1]; t.ok[0] += red[k].ok[0]; t.ok[1] += red[k].ok[1];
                t.brier[0] += red[k].brier[0]; t.brier[1] += red[k].brier[1];
              }
              a.acc[(size_t)tree * (a.nIntr + 1) + s] = t;
            }
          }
        }
      }
      __syncthreads();
    }
  }
}

}  // namespace
}  // namespace rfslam
