// common.cuh -- shared device/host helpers for the B200-native rfSLAM hot path (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include <string>

namespace rfslam {

// ------------------------------------------------------------------------------------------
// error plumbing: every CUDA failure becomes a non-zero C-ABI status with a message
// ------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
struct CudaFail { int code; };

#define RF_CUDA(call)                                                                              \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      rfslam::set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e__), __FILE__, __LINE__, #call); \
      throw rfslam::CudaFail{2};                                                                   \
    }                                                                                              \
  } while (0)

// ------------------------------------------------------------------------------------------
// constants of the path
// ------------------------------------------------------------------------------------------
#ifndef RF_THREADS
#define RF_THREADS 256
#endif
constexpr int      kThreads     = RF_THREADS;   // CTA size of the tree kernel (128 or 256)
constexpr int      kWarps       = kThreads / 32;
constexpr int      kHistBins    = 256;          // covariates with <= 256 distinct values score by histogram
constexpr int      kMaxCand     = 16;           // mtry limit of the joint histogram pass
constexpr int      kMaxStrata   = 256;          // largest stratum id held in shared memory
constexpr uint32_t kRowBits     = 26;           // in-bag entry = (multiplicity-1) << 26 | internal row
constexpr uint32_t kRowMask     = (1u << kRowBits) - 1u;
constexpr uint32_t kMaxMult     = 64;
constexpr int      kStackCap    = 4096;         // pending right children per tree
constexpr int      kRecChunk    = 1024;         // node records per pool chunk
constexpr double   kEpsilon     = 1.0e-9;       // rf.h:169 EPSILON (absolute hysteresis of the tracker)

// ------------------------------------------------------------------------------------------
// ran1 chain replay (rf.c:5976-6117); state kept in one small struct so it can live in smem
// ------------------------------------------------------------------------------------------
struct Ran1 {
  int iy;
  int idum;
  int iv[32];
};

__host__ __device__ inline int ran1_step(int v) {
  // Schrage's factorisation of v * 16807 mod (2^31 - 1)
  int k = v / 127773;
  v = 16807 * (v - k * 127773) - 2836 * k;
  if (v < 0) v += 2147483647;
  return v;
}

__host__ __device__ inline void ran1_seed(Ran1& s, int seed) {
  s.iy = 0;
  s.idum = seed;
}

__host__ __device__ inline float ran1_next(Ran1& s) {
  if (s.idum <= 0 || s.iy == 0) {
    s.idum = (-(s.idum) < 1) ? 1 : -(s.idum);
    for (int j = 39; j >= 0; j--) {
      s.idum = ran1_step(s.idum);
      if (j < 32) s.iv[j] = s.idum;
    }
    s.iy = s.iv[1];                       // the reference's generic variant reads slot 1 (rf.c:6095)
  }
  s.idum = ran1_step(s.idum);
  int j = s.iy / 67108864;                // NDIV = 1 + (IM - 1) / 32 = 2^26
  s.iy = s.iv[j];
  s.iv[j] = s.idum;
  float t = (float)((1.0 / 2147483647.0) * (double)s.iy);
  const double rnmx = 1.0 - 1.2e-7;
  if ((double)t > rnmx) return (float)rnmx;
  return t;
}

// The initialisation branch of ran1_generic taken on its own (rf.c:6082-6096): running it right
// after seeding changes nothing, and lets later draws skip the test.
__host__ __device__ inline void ran1_init(Ran1& s) {
  if (s.idum <= 0 || s.iy == 0) {
    s.idum = (-(s.idum) < 1) ? 1 : -(s.idum);
    for (int j = 39; j >= 0; j--) {
      s.idum = ran1_step(s.idum);
      if (j < 32) s.iv[j] = s.idum;
    }
    s.iy = s.iv[1];
  }
}

// 16807^k mod (2^31 - 1), k = 1..32: the Park-Miller state k steps ahead is (kJump[k-1] * idum) mod M,
// bit-identical to k Schrage steps (both are exact modular products)
__device__ __constant__ const uint32_t kJump[32] = {16807u, 282475249u, 1622650073u, 984943658u, 1144108930u, 470211272u,
  101027544u, 1457850878u, 1458777923u, 2007237709u, 823564440u, 1115438165u, 1784484492u, 74243042u, 114807987u, 1137522503u,
  1441282327u, 16531729u, 823378840u, 143542612u, 896544303u, 1474833169u, 1264817709u, 1998097157u, 1817129560u, 1131570933u,
  197493099u, 1404280278u, 893351816u, 1505795335u, 1954899097u, 1636807826u};

__host__ __device__ inline int ran1_mulmod(uint32_t a, uint32_t v) {
  unsigned long long p = (unsigned long long)a * (unsigned long long)v;
  unsigned long long r = (p & 0x7FFFFFFFull) + (p >> 31);
  r = (r & 0x7FFFFFFFull) + (r >> 31);
  if (r >= 0x7FFFFFFFull) r -= 0x7FFFFFFFull;
  return (int)r;
}

__host__ __device__ inline float ran1_float(int iy) {
  float t = (float)((1.0 / 2147483647.0) * (double)iy);
  const double rnmx = 1.0 - 1.2e-7;
  if ((double)t > rnmx) return (float)rnmx;
  return t;
}

// draw k in [1, size]: (uint) ceil(ran1 * (size * 1.0))  (rf.c:531, rf.c:8380)
__host__ __device__ inline uint32_t ran1_index(Ran1& s, uint32_t size) {
  return (uint32_t)ceil((double)ran1_next(s) * ((double)size * 1.0));
}

inline unsigned lcg_next(unsigned s) { return (1366u * s + 150889u) % 714025u; }

// ------------------------------------------------------------------------------------------
// the six Poisson rules (sc.c:639-1687) as one per-(side, stratum) term
// ------------------------------------------------------------------------------------------
__host__ __device__ inline bool rule_is_bayes(int rule)      { return rule <= 3; }
__host__ __device__ inline bool rule_is_stratified(int rule) { return rule == 1 || rule == 3 || rule == 4 || rule == 6; }
__host__ __device__ inline bool rule_unit_time(int rule)     { return rule == 3 || rule == 6; }

// E, N: event / row counts; T: exposure; W: sum of event * exposure.  Zero-event cells contribute
// exactly 0 under every rule (Bayes: 0 * log(.), sc.c:761; raw: mu = 0 is skipped, sc.c:1282).
// Branch-free form of side_term.  On the device this is ONE out-of-line function: the tree kernel is
// instruction-fetch bound on its once-per-node code, and every inlined copy of the FP64 divide + log
// sequence is ~120 instructions.
#ifdef __CUDACC__
static __device__ __noinline__
#else
inline
#endif
double side_term_nb(bool bayes, double E, double N, double T, double W, double alpha, double beta) {
  if (bayes) {
    const double lam = (E + alpha) / (beta + T);
    const double r = E * log(lam);
    return (E == 0.0) ? 0.0 : r;
  }
  const double mu0 = (T != 0.0) ? W / T : 0.0;
  const bool ok = (E != 0.0) && (mu0 != 0.0);
  const double mu = ok ? mu0 : 1.0;
  const double r = (E * log(mu)) - (N * mu);
  return ok ? r : 0.0;
}


__host__ __device__ inline double side_term(bool bayes, double E, double N, double T, double W, double alpha, double beta) {
#ifdef __CUDA_ARCH__
  return side_term_nb(bayes, E, N, T, W, alpha, beta);
#else
  if (E == 0.0) return 0.0;
  if (bayes) {
    double lam = (E + alpha) / (beta + T);
    return E * log(lam);
  }
  if (T == 0.0) return 0.0;
  double mu = W / T;
  if (mu == 0.0) return 0.0;
  return (E * log(mu)) - (N * mu);
#endif
}

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------
// warp / block primitives
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
// Virtual warp / thread ids.  The tree kernel gives warp 0 / thread 0 all the serial work of a node
// (draws, tracker combine, records).  Hardware warp w issues on scheduler w % 4, so with several CTAs
// per SM every CTA's serial warp would pile onto scheduler 0; rotating the numbering per CTA spreads
// them (CTAs b and b + 148 share an SM under the round-robin placement and get different rotations).
__device__ __forceinline__ unsigned warp_wrap(unsigned v) { return (kWarps & (kWarps - 1)) == 0 ? (v & (unsigned)(kWarps - 1)) : (v % (unsigned)kWarps); }
__device__ __forceinline__ int warp_rot() { return (int)warp_wrap(blockIdx.x * 5u + (blockIdx.x / 148u) * 3u); }
__device__ __forceinline__ int warp_id() { return (int)warp_wrap((threadIdx.x >> 5) + (unsigned)warp_rot()); }
__device__ __forceinline__ int vtid() { return (warp_id() << 5) | (int)(threadIdx.x & 31); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__device__ __forceinline__ T warp_incl_scan(T v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += u;
  }
  return v;
}

// block-wide sum; every thread gets the result.  `scratch` holds kWarps elements of T in smem.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
  v = warp_sum(v);
  __syncthreads();
  if (lane_id() == 0) scratch[warp_id()] = v;
  __syncthreads();
  T r = scratch[0];
#pragma unroll
  for (int w = 1; w < kWarps; w++) r += scratch[w];
  return r;
}

// block-wide inclusive scan (kThreads elements, one per thread); returns the inclusive prefix and
// the block total through `total`.  `scratch` holds kWarps elements.
template <typename T>
__device__ __forceinline__ T block_incl_scan(T v, T* scratch, T& total) {
  T inc = warp_incl_scan(v);
  __syncthreads();
  if (lane_id() == 31) scratch[warp_id()] = inc;
  __syncthreads();
  T off = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kWarps; w++) {
    T s = scratch[w];
    if (w < warp_id()) off += s;
    tot += s;
  }
  total = tot;
  return inc + off;
}

__device__ __forceinline__ uint32_t load_code(const void* p, int width, uint32_t row) {
  if (width == 1) return ((const uint8_t*)p)[row];
  if (width == 2) return ((const uint16_t*)p)[row];
  return ((const uint32_t*)p)[row];
}
#endif  // __CUDACC__

}  // namespace rfslam
