// vimp.cuh -- permutation importance on the device (getPermuteMembership rf.c:2106-2218, permute() rf.c:1974-2007,
// updateGenericVimpEnsemble rf.c:2317-2415, summarize / finalizeVimpPerformance rf.c:2416-2697).  Included by predict.cu
// (needs PredRec / rec_goes_left).
//
// One CTA per tree, trees time-sliced over the resident CTAs.  Chain C of a tree is one sequential stream over the
// covariates of interest, so warp 0 is the PRODUCER: it replays the chain (32 Park-Miller states per jump-ahead, the
// Bays-Durham table spread over the lanes) and builds the permutation of the tree's OOB rows for covariate v --
// permute() gives slot "k-th still-empty" the value i for i = m..1, which the reference finds by a linear walk
// (O(m^2)).  Here 32 draws are resolved per step: their ranks among the slots empty at the START of the step follow from
// the earlier draws of the step alone (rank = k + #earlier ranks below it: a ballot fixed point, exact), then every lane
// looks its rank up in an occupancy bitmap with 8-ary empty-slot counts in shared memory (one 16-byte load per level)
// and the 32 slots are taken at once.  Warps 1..7 are the CONSUMERS: while the producer works on covariate v they drop the
// OOB rows down the tree with covariate v - 1 replaced by its permuted value and either count the votes of the tree's own
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
  const uint16_t* cnt; size_t cstride;   // in-bag counts [tree][row]; nullptr = the importance is taken over all n rows
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

// chain replay by one warp: 32 draws per refill (as bootstrap_draw_kernel, grow.cu)
struct WarpRan1 {
  int ivReg, iy, idum; float u; int pos;
  __device__ void seed(int seedValue, int lane, Ran1* tmp) {
    if (lane == 0) { ran1_seed(*tmp, seedValue); ran1_init(*tmp); }
    __syncwarp();
    ivReg = tmp->iv[lane]; iy = tmp->iy; idum = tmp->idum; pos = 32; u = 0.f;
    __syncwarp();
  }
  __device__ __forceinline__ void refill(int lane) {
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
  }
  // the next `cnt` (uniform, 1..32) draws of the chain: lane j < cnt receives the j-th of them
  __device__ __forceinline__ float take(int lane, int cnt) {
    if (pos == 32) { refill(lane); pos = 0; }
    float v = __shfl_sync(0xffffffffu, u, (pos + lane) & 31);
    const bool fromNext = pos + lane >= 32;
    if (pos + cnt > 32) {
      refill(lane);
      const float v2 = __shfl_sync(0xffffffffu, u, (pos + lane) & 31);
      if (fromNext) v = v2;
      pos = pos + cnt - 32;
    } else {
      pos += cnt;
    }
    return v;
  }
};

// Empty-slot index of one permutation: occupancy bitmap (bit set = slot taken) and 8-ary counts of EMPTY slots --
// c1 per 8 words (256 slots), c2 per 2048, c3 per 16384 (u16, 8 to a 16-byte load), top per 131072 (u32).
struct SlotIndex {
  uint32_t* bits; uint16_t* c1; uint16_t* c2; uint16_t* c3; uint32_t* top;
  int W0, n1, n2, n3, nt;
  __host__ __device__ static size_t bytes(int m) {
    const int W0 = ((m + 31) >> 5), n1 = (W0 + 7) >> 3, n2 = (n1 + 7) >> 3, n3 = (n2 + 7) >> 3, nt = (n3 + 7) >> 3;
    return (size_t)(((W0 + 7) & ~7) * 4 + (((n1 + 7) & ~7) + ((n2 + 7) & ~7) + ((n3 + 7) & ~7)) * 2 + nt * 4 + 64);
  }
  __device__ void carve(uint32_t* base, int mcap) {
    const int W0c = ((mcap + 31) >> 5), n1c = (W0c + 7) >> 3, n2c = (n1c + 7) >> 3, n3c = (n2c + 7) >> 3;
    bits = base;
    c1 = (uint16_t*)(bits + ((W0c + 7) & ~7));
    c2 = c1 + ((n1c + 7) & ~7);
    c3 = c2 + ((n2c + 7) & ~7);
    top = (uint32_t*)(c3 + ((n3c + 7) & ~7));
  }
  // all slots of [0, m) empty (one warp)
  __device__ void reset(int m, int lane) {
    W0 = (m + 31) >> 5; n1 = (W0 + 7) >> 3; n2 = (n1 + 7) >> 3; n3 = (n2 + 7) >> 3; nt = (n3 + 7) >> 3;
    for (int k = lane; k < ((W0 + 7) & ~7); k += 32) bits[k] = k < W0 - 1 ? 0u : (k == W0 - 1 ? ((m & 31) ? ~((1u << (m & 31)) - 1u) : 0u) : 0xffffffffu);
    for (int k = lane; k < ((n1 + 7) & ~7); k += 32) c1[k] = (uint16_t)max(0, min(m, (k + 1) * 256) - k * 256);
    for (int k = lane; k < ((n2 + 7) & ~7); k += 32) c2[k] = (uint16_t)max(0, min(m, (k + 1) * 2048) - k * 2048);
    for (int k = lane; k < ((n3 + 7) & ~7); k += 32) c3[k] = (uint16_t)max(0, min(m, (k + 1) * 16384) - k * 16384);
    for (int k = lane; k < nt; k += 32) top[k] = (uint32_t)max(0, min(m, (k + 1) * 131072) - k * 131072);
  }
  // which of 8 u16 counts (one 16-byte load) holds the k-th empty slot; k becomes the rank inside it
  __device__ __forceinline__ static int pick8(const uint16_t* c, uint32_t& k) {
    const uint4 v = *reinterpret_cast<const uint4*>(c);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    int at = 7;
#pragma unroll
    for (int q = 7; q >= 0; q--) {                  // prefix sums in registers
      uint32_t before = 0;
#pragma unroll
      for (int t = 0; t < 8; t++) if (t < q) before += (w[t >> 1] >> (16 * (t & 1))) & 0xffffu;
      const uint32_t mine = (w[q >> 1] >> (16 * (q & 1))) & 0xffffu;
      if (k > before && k <= before + mine) at = q;
    }
    uint32_t before = 0;
#pragma unroll
    for (int t = 0; t < 8; t++) if (t < at) before += (w[t >> 1] >> (16 * (t & 1))) & 0xffffu;
    k -= before;
    return at;
  }
  // slot of the r-th (1-based) empty slot
  __device__ __forceinline__ int select(uint32_t r) const {
    int t = 0;
    for (; t < nt - 1; t++) { const uint32_t c = top[t]; if (r <= c) break; r -= c; }
    const int i3 = t * 8 + pick8(c3 + t * 8, r);
    const int i2 = i3 * 8 + pick8(c2 + i3 * 8, r);
    const int i1 = i2 * 8 + pick8(c1 + i2 * 8, r);
    const uint4 a = *reinterpret_cast<const uint4*>(bits + i1 * 8), b = *reinterpret_cast<const uint4*>(bits + i1 * 8 + 4);
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    int wi = 7; uint32_t before = 0, acc = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const uint32_t e = 32u - (uint32_t)__popc(w[q]);
      if (r > acc && r <= acc + e) { wi = q; before = acc; }
      acc += e;
    }
    uint32_t word = w[0];
#pragma unroll
    for (int q = 1; q < 8; q++) if (wi == q) word = w[q];
    const int bit = (int)__fns(~word, 0, (int)(r - before));
    return (i1 * 8 + wi) * 32 + bit;
  }
  // several lanes take their slots at once
  __device__ __forceinline__ void take(int slot) {
    const int wi = slot >> 5, i1 = wi >> 3, i2 = i1 >> 3, i3 = i2 >> 3, t = i3 >> 3;
    atomicOr(&bits[wi], 1u << (slot & 31));
    atomicSub(reinterpret_cast<uint32_t*>(c1) + (i1 >> 1), 1u << (16 * (i1 & 1)));
    atomicSub(reinterpret_cast<uint32_t*>(c2) + (i2 >> 1), 1u << (16 * (i2 & 1)));
    atomicSub(reinterpret_cast<uint32_t*>(c3) + (i3 >> 1), 1u << (16 * (i3 & 1)));
    atomicSub(&top[t], 1u);
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

// dynamic shared memory: the producer's SlotIndex (SlotIndex::bytes(maxOob))
__global__ void __launch_bounds__(kVimpThreads) vimp_tree_kernel(VimpArgs a) {
  extern __shared__ __align__(16) uint32_t dsm[];
  __shared__ Ran1 rngTmp;
  __shared__ int warpBase[kVimpThreads / 32 + 1];
  __shared__ VimpAcc red[kVimpThreads / 32];
  __shared__ int sM;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  SlotIndex si; si.carve(dsm, a.maxOob);
  uint32_t* idx = a.scratch + (size_t)blockIdx.x * 3 * a.maxOob;
  uint32_t* permBuf[2] = {idx + a.maxOob, idx + 2 * (size_t)a.maxOob};

  for (int tree = blockIdx.x; tree < a.ntree; tree += gridDim.x) {
    const uint16_t* cnt = a.cnt ? a.cnt + (size_t)tree * a.cstride : nullptr;       // nullptr: every row (held-out rows)
    // ---- the tree's OOB rows, ascending (RF_oobMembershipIndex) ----
    int base = 0;
    for (int i0 = 0; i0 < a.n; i0 += kVimpThreads) {
      const int i = i0 + tid;
      const bool oob = i < a.n && (!cnt || cnt[i] == 0);
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

    for (int s = 0; s <= a.nIntr; s++) {
      if (w == 0) {
        // ---------------- producer: the permutation of covariate s ----------------
        if (s < a.nIntr) {
          uint32_t* perm = permBuf[s & 1];
          si.reset(m, lane);
          __syncwarp();
          for (int i0 = m; i0 > 0; i0 -= 32) {
            const int cntB = min(32, i0);
            const float uu = rng.take(lane, cntB);
            const int mine = i0 - lane;                                                     // lane j draws for i = i0 - j
            const uint32_t k = lane < cntB ? (uint32_t)ceil((double)uu * ((double)mine * 1.0)) : 0u;     // rf.c:1998
            // ranks among the slots empty at the start of the step: r_j = k_j + #{t < j : r_t <= r_j} (least fixed point)
            uint32_t r = 0;
            for (int j = 0; j < cntB; j++) {
              const uint32_t kj = __shfl_sync(0xffffffffu, k, j);
              uint32_t c = 0;
              while (true) {
                const uint32_t cn = (uint32_t)__popc(__ballot_sync(0xffffffffu, lane < j && r <= kj + c));
                if (cn == c) break;
                c = cn;
              }
              if (lane == j) r = kj + c;
            }
            int slot = -1;
            if (lane < cntB) { slot = si.select(r); perm[slot] = (uint32_t)mine; }          // slot (1-based slot + 1) receives i
            __syncwarp();
            if (lane < cntB) si.take(slot);
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
