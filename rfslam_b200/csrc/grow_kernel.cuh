// grow_kernel.cuh -- the persistent tree-growing kernel: one CTA grows one whole tree, depth
// first, with no host round trips.
//
// Why one CTA per tree: chain B of the reference's RNG is consumed in
// left-first DFS node order (rf.c:21666-21691, rf.c:14915), so a node's draws are unknown until
// every node before it in DFS order has been grown; trees, however, are independent
// (rf.c:7082-7087).  148+ trees in flight (one or two per SM) keep every SM streaming its share
// of HBM while each CTA follows the sequential DFS the results are defined by.
//
// Per attempted node the CTA does (reference functions replaced in brackets):
//   1. draw <= mtry covariates without replacement          [selectRandomCovariates rf.c:14793]
//      (nsplit > 0: candidate by candidate, each followed by its cut draws, rf.c:14533-14564)
//   2. sufficient statistics of every (candidate, cut):
//      - node of <= 512 entries (97 % of nodes): the node is staged in shared memory and warp ci scores
//        candidate ci alone -- rule terms once per row of an event stratum (lane = row) or per cut
//        (lane = cut) for long strata; covariates with > 256 distinct values are re-ranked locally
//      - larger node: passes over the entries, up to four strata per pass, per-(candidate, code)
//        histograms in shared memory with native 32-bit atomics, then per stratum prefix sums over
//        codes, the rule's term at every code where the stratum has rows, fill-forward
//      [the per-cut regather rf.c:13455-13465 + the O(m) passes of poissonSplitN sc.c:710-747,
//       sc.c:751-764 / sc.c:1262-1290]; ascending stratum order as in the reference
//   3. sequential epsilon-hysteresis tracker over candidates in draw order and cuts in ascending
//      order, exact and parallel via per-candidate summaries    [updateMaximumSplit rf.c:15471]
//   4. stable partition of the entry list into the two children [forkAndUpdate rf.c:10568,
//      rf.c:21653-21663]
// Covariates with more than 256 distinct values in nodes above 256 entries take the sorted-walk path:
// the CTA radix-sorts the node's codes in global scratch and walks the sorted order with block scans.
// Batches with more trees than resident CTAs are time-sliced (see rfslam_grow_kernel).
#pragma once
#include "common.cuh"
#include "internal.h"

#ifndef RF_ASSUME
#define RF_ASSUME 3    // bits 4, 8 (the staging / tile region) and 16 (GrowShared: faults when trees are parked and resumed) miscompute on
                       // nvcc 12.9 when assumed shared: left generic.  GrowShared is a static __shared__ object instead.
#endif

namespace rfslam {

// RF_CHECK builds (RFSLAM_NVCC_EXTRA=-DRF_CHECK): device-side sanity checks of the entry lists and the DFS state
// (compute-sanitizer is not available on the test boxes); a violation prints once and traps.
#ifdef RF_CHECK
#define RF_ASSERT(cond, ...) do { if (!(cond)) { printf("RF_CHECK " __VA_ARGS__); __trap(); } } while (0)
#else
#define RF_ASSERT(cond, ...) do { } while (0)
#endif

struct Rec {            // one forest record (pre-order); 16 bytes
  uint32_t nodeID;
  uint32_t parm;        // covariate + 1, 0 = terminal
  uint32_t a;           // internal: cut code      terminal: in-bag events (with multiplicity)
  uint32_t b;           // internal: record index of the right child   terminal: in-bag count
};

struct NodeEnt {        // a pending node of the DFS
  uint32_t start, len;  // entry range in the tree's list buffer
  uint32_t m, E;        // in-bag rows / events with multiplicity
  double   rtsum;       // sum of multiplicity * risk time (beta of the Bayes rules, sc.c:662-667)
  uint32_t nodeID;
  uint32_t parentRec;   // record whose right-child link is this node, or kNone
  uint32_t depth;
  uint32_t buf;         // which of the two list buffers holds the entries
};

constexpr uint32_t kNone = 0xFFFFFFFFu;
#ifndef RF_FULL
#define RF_FULL 0
#endif
// The kernel exists in a LEAN build (tables whose covariates all have <= 256 distinct values and no unordered factor: every
// benchmarked configuration) and a FULL one (grow_t256f.cu).  It is instruction-fetch sensitive on its once-per-node code:
// merely compiling the level-subset paths of factors in slowed the all-numeric case from 1.46 s to 1.56 s per 500 trees
// (16.3 k -> 18.8 k SASS instructions), the rebuilt sorted walk another 2 %.  So the paths a lean table never executes are
// compile-time switches (RF_K_*): factors, the sorted walk of wide covariates, the single-node emit mode and the phase
// profile are FULL-only (each switch measured: -3 % for the sorted walk + emit mode, -3 % for the profile stamps);
// sequential draws stay in both (staged subtrees, opt-in and unhelpful at 10 M rows, are FULL-only too).  Compiling the sequential-draw path OUT (RF_K_SEQ = 0) makes
// time-sliced launches FAULT (the second tree on a CTA reads garbage in-bag counts; bisected to this one switch, although
// a table that never draws sequentially executes the same statements either way) -- not understood; every
// shipped build keeps it in and passes the time-slicing stress tests.
constexpr bool kFull = RF_FULL != 0;
#ifndef RF_K_SORT
#define RF_K_SORT RF_FULL
#endif
#ifndef RF_K_SINK
#define RF_K_SINK RF_FULL
#endif
#ifndef RF_K_SUB
#define RF_K_SUB RF_FULL
#endif
#ifndef RF_K_PROF
#define RF_K_PROF RF_FULL
#endif
#ifndef RF_K_SEQ
#define RF_K_SEQ 1
#endif
constexpr bool kFactors = kFull, kSeq = RF_K_SEQ != 0, kSort = RF_K_SORT != 0, kSub = RF_K_SUB != 0, kSink = RF_K_SINK != 0, kProf = RF_K_PROF != 0;
#define RF_CANDHIST(ci) (!kSort || s->candHist[ci])
constexpr uint32_t kFactorFlag = 0x80000000u;   // Rec::parm of a factor split: `a` is then the LEFT set as a mask over codes
#ifndef RF_SMALL_MAX
#define RF_SMALL_MAX 512
#endif
constexpr int kSmallMax = RF_SMALL_MAX;      // nodes with <= this many entries are scored entirely out of shared memory
constexpr int kLocalMax = 256;      // small nodes up to this many entries also take covariates with > 256 distinct values (local ranks)      // nodes with <= this many entries are scored entirely out of shared memory
#ifndef RF_PASS_CAND
#define RF_PASS_CAND 16
#endif
constexpr int kPassCand = RF_PASS_CAND;   // candidates the unrolled histogram pass of a large node is compiled for (8: the lean 256-thread build,
                                          // launched only when mtry <= 8; the layout of GrowShared keeps kMaxCand slots in every build)
constexpr int kTileThreads = 256;   // lanes of a sorted-walk tile (the wide-covariate path runs in the 256-thread builds only)
constexpr int kSparseRatio = 12;    // a node with fewer than n / kSparseRatio entries reads the row-major mirror of the table
constexpr int kProfSlots = 96;   // 64..79: the phase slots again, restricted to nodes with < 64 entries   // 0..15 phases, 16..47 node time by log2(len), 48..63 node count by log2(len)/2

struct GrowParams {
  int rule, mtry, nodesize, nodedepth;
  double alpha;
  const double* split_weight;    // [p] or nullptr
  int pw;                        // words of a permissible-covariate mask
  int ncmax;                     // min(mtry, kMaxCand)
  int has_sort;                  // any covariate with D > kHistBins
  int hbins;                     // histogram stride (see GrowSmem::hb)
  double tscale, tinv;           // 2^S and 2^-S of the fixed-point exposure accumulators (see hist_add)
  int nsplit;                    // > 0: at most nsplit randomly drawn cuts per covariate (rf.c:14533-14564)
  int hslots;                    // stratum slots per candidate histogram (GrowSmem::hs), chosen by the host to fit shared memory
  int seqDraw;                   // candidates are drawn one at a time by the whole CTA (nsplit > 0 or a factor covariate: cut draws
                                 // interleave with the covariate draws in chain B); the tracker then reads the cut masks
  int facCuts;                   // > 0: room for this many cut masks per factor candidate (GrowSmem::facMask)
  int tabSlots;                  // strata per group of the small-node path (<= RF_MAX_SLOTS): sizes the term tables (GrowSmem::ts)
  int subMax;                    // > 0: a node of at most this many entries stages ITS WHOLE ROWS in shared memory and the subtree
                                 //      below it is grown out of that tile (see stage_subtree); 0 = off
};

struct GrowWork {
  int ntrees;
  const int* seedB;              // [ntrees]
  const int* bootSize;           // [ntrees] sum of in-bag counts
  uint32_t* lists;               // [ntrees][2][listStride]
  size_t listStride;             // entries a tree can hold: min(n, bootstrap size) + bootstrap size / kMaxMult + 1
  const uint16_t* cnt;           // [ntrees][cntStride] in-bag count per internal row
  size_t cntStride;
  uint32_t* sortbuf;             // [ntrees][4][listStride] or nullptr
  NodeEnt* stack;                // [ntrees][kStackCap]
  uint32_t* stackPerm;           // [ntrees][kStackCap][pw]
  Rec* pool; double* statPool;   // [poolChunks * kRecChunk]
  uint32_t poolChunks;
  uint32_t* poolCursor;
  uint32_t* ctab;                // [ntrees][maxChunksPerTree]
  uint32_t maxChunksPerTree;
  uint32_t* nodeCount;           // [ntrees]
  uint32_t* leafCount;           // [ntrees]
  unsigned long long* candCount; // [ntrees]
  unsigned long long* algoBytes; // [ntrees]
  int* status;                   // 0 ok, 1 record pool exhausted, 2 DFS stack overflow
  unsigned long long* prof;      // optional [ntrees][kProfSlots] clock64 sums per phase / node-size class
  // single-node scoring mode (rfslam_b200_score_node): emit every cut instead of tracking
  double* sinkCut; double* sinkDelta; int* sinkCount;
  // time slicing (batches with more trees than resident CTAs): see rfslam_grow_kernel
  int* treeState;                // [ntrees] 0 fresh, 1 suspended, 2 running, 3 done; nullptr = one CTA per tree, no slicing
  int* waiting;                  // trees that are fresh or suspended
  unsigned* ticket;              // round-robin cursor
  unsigned char* saved;          // [ntrees][sizeof(GrowShared)] state of suspended trees
  long long quantum;             // cycles a CTA works on a tree before it offers the slot to a waiting tree
};

// fixed-size part of the CTA's shared memory
struct GrowShared {
  Ran1 rng;
  NodeEnt cur;
  uint32_t perm[32];                 // permissible covariates of the current node (p <= 1024)
  int nc;
  int candCov[kMaxCand];
  int candD[kMaxCand];
  int candW[kMaxCand];
  int candHist[kMaxCand];
  const void* candPtr[kMaxCand];
  // tracker
  int have; double bestDelta; int bestCand; uint32_t bestCode;
  int cutsSeen;
  // scratch for block reductions / scans
  double scrD[32]; unsigned long long scrU[32]; int scrI[32];   // (sized for the widest CTA: the struct is the same in every build of the kernel)
  uint32_t nodeCount, leafCount, sp;
  unsigned long long candCount, algoBytes;
  int abort;
  int needSlow; int isSmall; int poolSize; int histDirty;
  // sparse evaluation of small nodes: the rows of the strata that contain events (<= 32), stratum-ascending
  uint32_t evMask[8];                // small nodes: strata that contain events
  int trkMaxBin[kMaxCand]; double trkMax[kMaxCand]; int trkClean[kMaxCand]; int trkNp[kMaxCand]; int trkLast[kMaxCand];
  unsigned long long prof[kProfSlots];
  long long tmark;
  long long sliceStart; int curTree; int fresh;
  // nsplit > 0: the cuts drawn for each candidate (bit = code), scratch of the draw, cuts per candidate
  uint32_t cutMask[kMaxCand][8]; uint32_t rankSel[8]; uint8_t swor[256]; int trkCuts[kMaxCand];
  int candStride;                    // byte-coded candidates: code of row r = candPtr[k][r * candStride] (1 = column-major array, rowPitch = row-major mirror)
  int inSub; uint32_t subSp, subStart; // growing inside a staged subtree: DFS depth at its root, list offset of its first entry
  uint8_t candMask[kMaxCand];        // sequential draws: the tracker takes this candidate's cuts from cutMask (drawn subset, factor cuts)
  int seqNext;
  uint8_t candFac[kMaxCand];         // candidate is an unordered factor: its cuts are the level subsets facMask[k][0 .. facN[k])
  int facN[kMaxCand];
  uint8_t candLocal[kMaxCand];       // candidate re-ranked locally by this node (its codes in shared memory are node-local ranks)
  uint32_t bestWide;                 // the winning cut of a locally re-ranked candidate as a table-wide code
};

#define RF_PROF(slot)                                                                  \
  do {                                                                                 \
    if (kProf && w.prof && vtid() == 0) {                                         \
      long long now__ = clock64();                                                     \
      s->prof[slot] += (unsigned long long)(now__ - s->tmark);                         \
      if (s->cur.len < 64u) s->prof[64 + (slot)] += (unsigned long long)(now__ - s->tmark); \
      else if (s->cur.len <= (uint32_t)kSmallMax) s->prof[80 + (slot)] += (unsigned long long)(now__ - s->tmark); \
      s->tmark = now__;                                                                \
    }                                                                                  \
  } while (0)

// the same inside a candidate's warp (small_candidate): thread 0 of the CTA stamps the sub-phases of ITS candidate
#define RF_PROFC(slot)                                                                 \
  do {                                                                                 \
    if (kProf && c.w.prof && vtid() == 0) {                                            \
      long long now__ = clock64();                                                     \
      s->prof[slot] += (unsigned long long)(now__ - s->tmark);                         \
      if (s->cur.len < 64u) s->prof[64 + (slot)] += (unsigned long long)(now__ - s->tmark); \
      else if (s->cur.len <= (uint32_t)kSmallMax) s->prof[80 + (slot)] += (unsigned long long)(now__ - s->tmark); \
      s->tmark = now__;                                                                \
    }                                                                                  \
  } while (0)

__device__ __forceinline__ Rec* rec_ptr(const GrowWork& w, int tree, uint32_t q) {
  uint32_t c = w.ctab[(size_t)tree * w.maxChunksPerTree + q / kRecChunk];
  return w.pool + (size_t)c * kRecChunk + (q % kRecChunk);
}

__device__ __forceinline__ int block_incl_scan_max(int v, int* scratch) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v = max(v, u);
  }
  __syncthreads();
  if (lane == 31) scratch[warp_id()] = v;
  __syncthreads();
  int off = -1;
  for (int w2 = 0; w2 < warp_id(); w2++) off = max(off, scratch[w2]);
  return max(v, off);
}

// one staged entry of a small node (score_small): multiplicity * risk time, multiplicity, event | (stratum - 1) << 8
struct __align__(16) SmallRec { double mrt; uint32_t mult; uint32_t es; };

// dynamic shared memory carve-up
struct GrowSmem {
  double *sL, *sR, *hT, *hW, *statP;
  uint32_t *hN, *hE, *hNx, *hEx, *pres;
  uint32_t *smk;                     // small nodes: [ncmax][kTabSlots][8] codes that have rows in a slot's stratum
  double *tabL, *tabR;               // small nodes: per-candidate term tables [ncmax][kTabSlots][hb], laid over the whole histogram region
  // sorted-walk tile arrays (alias the small-node staging: the two paths never overlap in time)
  double *tileTL, *tileTR, *tileDelta; uint32_t *tileKey, *radixBase; uint8_t* tileCut;
  uint16_t *tileS, *tileSeg; uint8_t *tileO, *tileWcnt;   // sorted walk: stratum of a position, segment starts, origin lane, per-warp stratum counts
  uint16_t* pool;                    // permissible-covariate pool of the draw step (same aliasing: used between nodes only)
  uint32_t *segpos, *PN, *PE, *cN, *cE;
  double *PT, *PW, *cT, *cW, *cTL, *cTR;
  // small-node staging: the node's entries and everything scoring needs about them
  uint32_t* smEnt; SmallRec* smRec; uint8_t* smCodes;
  // staged subtree (subMax entries): entry words, records, the entries' whole code rows, and the node order (slot per position)
  uint32_t* stRow; SmallRec* stRec; uint8_t* stTile; uint16_t* stIdx; int subMax, tilePitch;
  uint32_t* facMask;                 // [ncmax][facCuts] LEFT sets of a factor candidate's cuts, as masks over CODES (bit = rank code)
  int facCuts;
  GrowShared* s;
  int hb;                            // bins per candidate histogram: pow2 >= the widest byte-coded column, 32..256
  int hs;                            // stratum slots per candidate histogram (large nodes accumulate hs strata per pass)
  int ts;                            // table slots per candidate of the small-node path (strata per group)
};

#ifndef RF_MAX_SLOTS
#define RF_MAX_SLOTS 4
#endif
#ifndef RF_ROWS_MAXSEG
#define RF_ROWS_MAXSEG 64   // small nodes: longest event stratum for which the lane = row evaluation is used
#endif
constexpr int kTabSlots = RF_MAX_SLOTS;    // strata per group of the small-node path (independent of the large-node passes' hs)
__host__ __device__ inline int hist_slots(int hb) { int g = 256 / hb; return g < 1 ? 1 : (g > RF_MAX_SLOTS ? RF_MAX_SLOTS : g); }

__host__ __device__ inline size_t grow_smem_bytes(int ncmax, int K, int hb, int hslots, int tabSlots, int subMax, int tilePitch, int facCuts = 0) {
  const size_t hs = (size_t)hslots;
  size_t KS = (size_t)K + 2;
  size_t b = 0;
  b += (size_t)ncmax * hb * 8 * 2;             // sL sR
  {                                            // hT hW hN hE hNx hEx = one region, at least the small nodes' two term tables
    const size_t hist = (size_t)ncmax * hb * hs * 32, tabs = (size_t)ncmax * hb * tabSlots * 16;
    b += hist > tabs ? hist : tabs;
  }
  b += (size_t)ncmax * 8;                      // statP
  b += (size_t)ncmax * 8 * 4;                  // pres
  b += (size_t)ncmax * tabSlots * 8 * 4;       // smk
  b += KS * 8 * 6;                             // PT PW cT cW cTL cTR
  b += KS * 4 * 5;                             // segpos PN PE cN cE
  {                                            // small-node staging; the sorted-walk tile arrays are aliased onto it
    const size_t stage = (size_t)kSmallMax * (16 + 4 + (size_t)ncmax) + 32;
    const size_t tiles = (size_t)kTileThreads * (8 * 3 + 4 + 1 + 2 + 1) + 256 * 4 + (KS + 2) * 2 + (size_t)(kTileThreads / 32) * KS + 64;
    b += stage > tiles ? stage : tiles;
  }
  b = (b + 15) & ~(size_t)15;
  b += (size_t)subMax * (4 + 16 + (size_t)tilePitch + 2) + 32;   // staged subtree
  b = (b + 15) & ~(size_t)15;
  b += (size_t)ncmax * (size_t)facCuts * 4;                      // factor cut masks
  return b + 16;                               // (GrowShared is a static __shared__ object of the kernel, not part of this)
}

__device__ inline GrowSmem carve(unsigned char* raw, int ncmax, int K, int hb, int hslots, int tabSlots, int subMax, int tilePitch, int facCuts) {
  GrowSmem m;
  m.hb = hb;
  m.hs = hslots;
  m.ts = tabSlots;
  m.subMax = subMax; m.tilePitch = tilePitch;
  const size_t hsz = (size_t)ncmax * hb * m.hs;
  size_t KS = (size_t)K + 2;
  uintptr_t tileEnd = 0;
  double* d = (double*)raw;
  m.sL = d; d += (size_t)ncmax * hb;
  m.sR = d; d += (size_t)ncmax * hb;
  m.hT = d; d += hsz;
  m.hW = d; d += hsz;
  {                                            // the u32 histograms follow at once: [hT hW hN hE hNx hEx] is one region of hsz * 32 bytes
    uint32_t* uh = (uint32_t*)d;
    m.hN = uh; uh += hsz; m.hE = uh; uh += hsz; m.hNx = uh; uh += hsz; m.hEx = uh; uh += hsz;
    const size_t hist = hsz * 32, tabs = (size_t)ncmax * hb * tabSlots * 16;
    m.tabL = m.hT; m.tabR = m.hT + (size_t)ncmax * hb * tabSlots;
    d = (double*)((unsigned char*)m.hT + (hist > tabs ? hist : tabs));
  }
  m.statP = d; d += ncmax;
  m.PT = d; d += KS; m.PW = d; d += KS; m.cT = d; d += KS; m.cW = d; d += KS; m.cTL = d; d += KS; m.cTR = d; d += KS;
  uint32_t* u = (uint32_t*)d;
  m.pres = u; u += (size_t)ncmax * 8;
  m.smk = u; u += (size_t)ncmax * tabSlots * 8;
  m.segpos = u; u += KS; m.PN = u; u += KS; m.PE = u; u += KS; m.cN = u; u += KS; m.cE = u; u += KS;
  {
    uintptr_t al = ((uintptr_t)u + 7) & ~(uintptr_t)7;
    u = (uint32_t*)al;
    // (extents for kTileThreads = 256 lanes in every build: the sorted walk only ever runs in the 256-thread builds, and the
    // host sizes the launch with this same layout)
    m.tileTL = (double*)u; m.tileTR = m.tileTL + kTileThreads; m.tileDelta = m.tileTR + kTileThreads;
    m.tileKey = (uint32_t*)(m.tileDelta + kTileThreads); m.radixBase = m.tileKey + kTileThreads; m.tileCut = (uint8_t*)(m.radixBase + 256);
    m.tileS = (uint16_t*)(m.tileCut + kTileThreads); m.tileSeg = m.tileS + kTileThreads;
    m.tileO = (uint8_t*)(m.tileSeg + KS + 2); m.tileWcnt = m.tileO + kTileThreads;
    tileEnd = (uintptr_t)(m.tileWcnt + (size_t)(kTileThreads / 32) * KS);
    m.pool = (uint16_t*)u;
  }
  {
    uintptr_t al = ((uintptr_t)u + 15) & ~(uintptr_t)15;
    m.smRec = (SmallRec*)al;
    u = (uint32_t*)(al + (size_t)kSmallMax * 16);
  }
  m.smEnt = u; u += kSmallMax;
  m.smCodes = (uint8_t*)u;
  u = (uint32_t*)(m.smCodes + (size_t)kSmallMax * ncmax);
  if ((uintptr_t)u < tileEnd) u = (uint32_t*)tileEnd;             // (one candidate: the sorted-walk tile arrays are the larger alias)
  {
    uintptr_t al = ((uintptr_t)u + 15) & ~(uintptr_t)15;
    m.stRec = (SmallRec*)al;
    m.stTile = (uint8_t*)(al + (size_t)subMax * 16);
    m.stRow = (uint32_t*)(m.stTile + (size_t)subMax * tilePitch);
    m.stIdx = (uint16_t*)(m.stRow + subMax);
    u = (uint32_t*)(m.stIdx + subMax + (subMax & 1));
  }
  {
    uintptr_t al = ((uintptr_t)u + 15) & ~(uintptr_t)15;
    m.facMask = (uint32_t*)al; m.facCuts = facCuts;
  }
  m.s = nullptr;                               // set by the kernel: GrowShared is a static __shared__ object, so every s->field
                                               // access is an LDS / STS by construction (no address-space assumption needed)
  // tell the compiler these all live in shared memory (LDS/STS/ATOMS instead of generic LD/ST)
#define RF_SH(x) __builtin_assume(__isShared(x))
#if RF_ASSUME & 1
  RF_SH(m.sL); RF_SH(m.sR); RF_SH(m.hT); RF_SH(m.hW); RF_SH(m.statP); RF_SH(m.hN); RF_SH(m.hE); RF_SH(m.hNx); RF_SH(m.hEx); RF_SH(m.pres);
#endif
#if RF_ASSUME & 2
  RF_SH(m.segpos); RF_SH(m.PN); RF_SH(m.PE); RF_SH(m.cN); RF_SH(m.cE); RF_SH(m.PT); RF_SH(m.PW); RF_SH(m.cT); RF_SH(m.cW); RF_SH(m.cTL); RF_SH(m.cTR);
#endif
#if RF_ASSUME & 4
  RF_SH(m.smEnt); RF_SH(m.smRec); RF_SH(m.smCodes);
#endif
#if RF_ASSUME & 8
  RF_SH(m.tileTL); RF_SH(m.tileTR); RF_SH(m.tileDelta); RF_SH(m.tileKey); RF_SH(m.radixBase); RF_SH(m.tileCut);
#endif
#undef RF_SH
  return m;
}

// ------------------------------------------------------------------------------------------
// tracker: the reference's updateMaximumSplit applied to one 32-wide chunk of candidate cuts in
// ascending order.  `val` is this lane's delta, `present` whether the lane holds a real cut.
// have/best/bestLane are warp-uniform; returns the number of selections (bestLane = last one).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void track_chunk(bool present, double val, int& have, double& best, int& selLane) {
  unsigned pend = __ballot_sync(0xffffffffu, present);
  const int lane = threadIdx.x & 31;
  selLane = -1;
  while (pend) {
    int j;
    if (!have) {
      j = __ffs(pend) - 1;
    } else {
      bool beat = ((pend >> lane) & 1u) && ((val - best) > kEpsilon);
      unsigned mask = __ballot_sync(0xffffffffu, beat);
      if (!mask) break;
      j = __ffs(mask) - 1;
    }
    best = __shfl_sync(0xffffffffu, val, j);
    have = 1;
    selLane = j;
    pend &= ~((2u << j) - 1u);
  }
}

struct NodeCtx {
  const DevData& d;
  const GrowParams& g;
  const GrowWork& w;
  GrowSmem& m;
  const uint32_t* list;   // the node's entries: list[0..len)
  int tree;
  bool bayes, unit, strat;
  double beta;
  int Kloop;
};

__device__ inline double factor_accumulate(const NodeCtx& c, int ci, uint32_t vN, uint32_t vE, double vT, double vW);
__device__ __forceinline__ void factor_set_presence(const NodeCtx& c, int ci);

// one in-bag entry into one candidate's histogram: native 32-bit shared atomics only.
// f64 (and 64-bit) shared atomics are CAS loops on sm_100a (ATOMS.CAST.SPIN) and were the top stall
// of the kernel, so (a) exposure is kept as rt_mode * (N - Nx) + Tx: only rows whose risk time
// differs from the table's dominant value t0 (in CPIU data: the last interval of a person) touch
// the exposure accumulators at all, and (b) those accumulate in 64-bit fixed point (resolution
// 2^-S, S = 62 - ceil(log2(bootstrap size * max risk time)), 2^-43 at 1M rows) as two native 32-bit
// adds: the low word's returned old value tells this add whether IT caused the carry.
__device__ __forceinline__ void fixed_add(double* acc, unsigned long long x) {
  uint32_t* w32 = (uint32_t*)acc;
  const uint32_t xl = (uint32_t)x, xh = (uint32_t)(x >> 32);
  const uint32_t old = atomicAdd(&w32[0], xl);
  const uint32_t hi = xh + (((uint32_t)(old + xl) < old) ? 1u : 0u);
  if (hi) atomicAdd(&w32[1], hi);
}
__device__ __forceinline__ double fixed_get(const double* acc, double tinv) {
  return (double)(*(const unsigned long long*)acc) * tinv;
}

__device__ __forceinline__ void hist_add(const GrowSmem& m, int idx, uint32_t mult, uint32_t e, bool unit, bool isT0, unsigned long long fx) {
  atomicAdd(&m.hN[idx], mult);
  if (e) atomicAdd(&m.hE[idx], mult);
  if (!unit && !isT0) {
    atomicAdd(&m.hNx[idx], mult);
    fixed_add(&m.hT[idx], fx);
    if (e) { atomicAdd(&m.hEx[idx], mult); fixed_add(&m.hW[idx], fx); }
  }
}

// ---- per-stratum flush of one candidate's histogram (step 3), executed by one warp ----------
// tN/tE/tT/tW are the stratum's totals in this node; the histogram is left clean.
__device__ __forceinline__ void flush_candidate(const NodeCtx& c, int ci, int slot, uint32_t tN, uint32_t tE, double tT, double tW,
                                                double termP, bool updatePres) {
  GrowSmem& m = c.m;
  GrowShared* s = m.s;
  const int lane = lane_id();
  const int Dc = s->candD[ci];
  const int nch = (Dc + 31) >> 5;
  const int ho = (ci * m.hs + slot) * m.hb;
  uint32_t* hN = m.hN + ho; uint32_t* hE = m.hE + ho;
  uint32_t* hNx = m.hNx + ho; uint32_t* hEx = m.hEx + ho;
  double* hT = m.hT + ho;   double* hW = m.hW + ho;
  const double t0 = c.d.rt_mode;
  const bool active = tE > 0;
  uint32_t cN = 0, cE = 0; double cT = 0.0, cW = 0.0;     // carries (warp-uniform)
  double cTL = 0.0, cTR = termP;
  for (int ch = 0; ch < nch; ch++) {
    int bin = (ch << 5) + lane;
    uint32_t vN = 0, vE = 0; double vT = 0.0, vW = 0.0;
    if (bin < Dc) {
      vN = hN[bin]; vE = hE[bin];
      if (!c.unit) {                                               // exposure = t0 * (rows at the dominant risk time) + the rest
        vT = t0 * (double)(vN - hNx[bin]) + fixed_get(&hT[bin], c.g.tinv);
        vW = t0 * (double)(vE - hEx[bin]) + fixed_get(&hW[bin], c.g.tinv);
      }
      hN[bin] = 0; hE[bin] = 0; hNx[bin] = 0; hEx[bin] = 0; hT[bin] = 0.0; hW[bin] = 0.0;   // leave the histogram clean
    }
    unsigned own = __ballot_sync(0xffffffffu, vN > 0);
    if (updatePres && lane == 0 && own) m.pres[ci * 8 + ch] |= own;
    if (!active) continue;
    if (c.unit) { vT = (double)vN; vW = (double)vE; }
    uint32_t iN = warp_incl_scan(vN) + cN, iE = warp_incl_scan(vE) + cE;
    double iT = warp_incl_scan(vT) + cT, iW = warp_incl_scan(vW) + cW;
    double tL = 0.0, tR = 0.0;
    if (vN > 0) {
      tL = side_term(c.bayes, (double)iE, (double)iN, iT, iW, c.g.alpha, c.beta);
      tR = side_term(c.bayes, (double)(tE - iE), (double)(tN - iN), tT - iT, tW - iW, c.g.alpha, c.beta);
    }
    // fill-forward from the nearest lane <= me that holds rows of this stratum
    unsigned le = own & (0xffffffffu >> (31 - lane));
    int src = le ? (31 - __clz(le)) : -1;
    double fL = __shfl_sync(0xffffffffu, tL, src < 0 ? 0 : src);
    double fR = __shfl_sync(0xffffffffu, tR, src < 0 ? 0 : src);
    if (src < 0) { fL = cTL; fR = cTR; }
    if (bin < Dc) { m.sL[ci * m.hb + bin] += fL; m.sR[ci * m.hb + bin] += fR; }
    cN = __shfl_sync(0xffffffffu, iN, 31); cE = __shfl_sync(0xffffffffu, iE, 31);
    cT = __shfl_sync(0xffffffffu, iT, 31); cW = __shfl_sync(0xffffffffu, iW, 31);
    cTL = __shfl_sync(0xffffffffu, fL, 31); cTR = __shfl_sync(0xffffffffu, fR, 31);
  }
}

// large-node variant: totals come from the histogram itself; `nslots` strata were accumulated in
// this pass (slot q = stratum k0 + q) and are flushed in ascending order
__device__ inline void flush_stratum(const NodeCtx& c, int nc, int nslots) {
  GrowSmem& m = c.m;
  GrowShared* s = m.s;
  const int lane = lane_id();
  for (int ci = warp_id(); ci < nc; ci += kWarps) {
    if (!RF_CANDHIST(ci)) continue;
    const int Dc = s->candD[ci];
    const int nch = (Dc + 31) >> 5;
    const double t0 = c.d.rt_mode;
    if (kFactors && s->candFac[ci]) {                                // unordered factor: per-level bins -> every level-subset cut
      for (int slot = 0; slot < nslots; slot++) {
        const int ho = (ci * m.hs + slot) * m.hb;
        uint32_t vN = 0, vE = 0; double vT = 0.0, vW = 0.0;
        if (lane < Dc) {
          vN = m.hN[ho + lane]; vE = m.hE[ho + lane];
          if (!c.unit) {
            vT = t0 * (double)(vN - m.hNx[ho + lane]) + fixed_get(&m.hT[ho + lane], c.g.tinv);
            vW = t0 * (double)(vE - m.hEx[ho + lane]) + fixed_get(&m.hW[ho + lane], c.g.tinv);
          }
          m.hN[ho + lane] = 0; m.hE[ho + lane] = 0; m.hNx[ho + lane] = 0; m.hEx[ho + lane] = 0; m.hT[ho + lane] = 0.0; m.hW[ho + lane] = 0.0;
        }
        const double termP = factor_accumulate(c, ci, vN, vE, vT, vW);
        if (lane == 0) m.statP[ci] += termP;
      }
      factor_set_presence(c, ci);
      continue;
    }
    for (int slot = 0; slot < nslots; slot++) {
      const int ho = (ci * m.hs + slot) * m.hb;
      const uint32_t* hN = m.hN + ho; const uint32_t* hE = m.hE + ho;
      const uint32_t* hNx = m.hNx + ho; const uint32_t* hEx = m.hEx + ho;
      const double* hT = m.hT + ho;   const double* hW = m.hW + ho;
      uint32_t tN = 0, tE = 0; double tT = 0.0, tW = 0.0;
      for (int ch = 0; ch < nch; ch++) {
        int bin = (ch << 5) + lane;
        if (bin < Dc) {
          tN += hN[bin]; tE += hE[bin];
          if (!c.unit) { tT += t0 * (double)(hN[bin] - hNx[bin]) + fixed_get(&hT[bin], c.g.tinv); tW += t0 * (double)(hE[bin] - hEx[bin]) + fixed_get(&hW[bin], c.g.tinv); }
        }
      }
      tN = warp_sum(tN); tE = warp_sum(tE);
      if (c.unit) { tT = (double)tN; tW = (double)tE; } else { tT = warp_sum(tT); tW = warp_sum(tW); }
      if (tN == 0) continue;
      double termP = 0.0;
      if (tE > 0) {
        termP = side_term(c.bayes, (double)tE, (double)tN, tT, tW, c.g.alpha, c.beta);
        if (lane == 0) m.statP[ci] += termP;
      }
      flush_candidate(c, ci, slot, tN, tE, tT, tW, termP, true);
    }
  }
}

// ---- tracker (step 4): exact replay of updateMaximumSplit (rf.c:15471-15551) ------------------
// The tracker can only ever select strict prefix-max records, so when no cut BEFORE the first
// occurrence of a candidate's maximum M lies within EPSILON of M, the candidate's final selection
// is (M, first position) if M beats the incoming best by more than EPSILON and nothing otherwise.
// track_summary (run by the warp that scored the candidate, all candidates in parallel) reduces a
// candidate to {#cuts, M, position, "needs the exact walk"}; track_combine (warp 0) then replays
// the candidates in draw order, walking the exact sequential chain only for near-ties and NaN.
__device__ inline void track_summary(const NodeCtx& c, int ci) {
  GrowSmem& m = c.m;
  GrowShared* s = m.s;
  const int lane = lane_id();
  const int Dc = (kFactors && s->candFac[ci]) ? s->facN[ci] + 1 : s->candD[ci];      // bins the tracker walks: codes, or a factor's cuts + the closing bin
  const int nch = (Dc + 31) >> 5;
  uint32_t word = (lane < 8) ? m.pres[ci * 8 + lane] : 0u;
  int np = warp_sum((int)__popc(word));
  int last = word ? (lane * 32 + 31 - __clz(word)) : -1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
  if (np < 2) { if (lane == 0) { s->trkNp[ci] = np; s->trkClean[ci] = 1; s->trkCuts[ci] = 0; } return; }
  const int cov = s->candCov[ci];
  const double wgt = c.g.split_weight ? c.g.split_weight[cov] : 1.0;
  const double statP = m.statP[ci];
  const uint32_t* cm = (kSeq && c.g.seqDraw && s->candMask[ci]) ? s->cutMask[ci] : nullptr;   // the candidate's cuts as a mask (drawn subset, factor cuts)
  int ncuts = np - 1;
  if (cm) ncuts = warp_sum((lane < 8) ? (int)__popc(cm[lane]) : 0);
  double M = -INFINITY; int pos = 0x7fffffff; bool anyNan = false;
#pragma unroll 1
  for (int ch = 0; ch < nch; ch++) {
    int bin = (ch << 5) + lane;
    uint32_t wd = cm ? cm[ch] : m.pres[ci * 8 + ch];
    bool present = ((wd >> lane) & 1u) && bin != last;
    if (present) {
      double val = (m.sL[ci * m.hb + bin] + m.sR[ci * m.hb + bin] - statP) * wgt;
      m.sL[ci * m.hb + bin] = val;                     // keep delta for the combine / exact walk
      if (val != val) anyNan = true;
      if (val > M) { M = val; pos = bin; }
    }
  }
  anyNan = __any_sync(0xffffffffu, anyNan);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double oM = __shfl_xor_sync(0xffffffffu, M, o); int oP = __shfl_xor_sync(0xffffffffu, pos, o);
    if (oM > M || (oM == M && oP < pos)) { M = oM; pos = oP; }
  }
  bool near = false;                                        // an earlier cut within EPSILON of the maximum?
  if (!anyNan) {
#pragma unroll 1
    for (int ch = 0; ch <= (pos >> 5); ch++) {
      int bin = (ch << 5) + lane;
      uint32_t wd = cm ? cm[ch] : m.pres[ci * 8 + ch];
      bool present = ((wd >> lane) & 1u) && bin != last && bin < pos;
      if (present && !((M - m.sL[ci * m.hb + bin]) > kEpsilon)) near = true;
    }
    near = __any_sync(0xffffffffu, near);
  }
  if (lane == 0) {
    s->trkNp[ci] = np; s->trkMax[ci] = M; s->trkMaxBin[ci] = pos; s->trkClean[ci] = (!anyNan && !near) ? 1 : 0;
    s->trkCuts[ci] = ncuts;
    s->trkLast[ci] = last;
  }
}

// exact chain walk of one candidate (warp 0); sL holds the deltas (track_summary)
__device__ inline void track_exact(const NodeCtx& c, int ci) {
  GrowSmem& m = c.m;
  GrowShared* s = m.s;
  const int lane = lane_id();
  const int Dc = (kFactors && s->candFac[ci]) ? s->facN[ci] + 1 : s->candD[ci];      // bins the tracker walks: codes, or a factor's cuts + the closing bin
  const int nch = (Dc + 31) >> 5;
  const int last = s->trkLast[ci];
  const int cov = s->candCov[ci];
  int have = s->have; double best = s->bestDelta; int bestBin = -1;
  int emitted = s->cutsSeen;
  const uint32_t* cm = (kSeq && c.g.seqDraw && s->candMask[ci]) ? s->cutMask[ci] : nullptr;
#pragma unroll 1
  for (int ch = 0; ch < nch; ch++) {
    int bin = (ch << 5) + lane;
    uint32_t wd = cm ? cm[ch] : m.pres[ci * 8 + ch];
    bool present = ((wd >> lane) & 1u) && bin != last;
    double val = present ? m.sL[ci * m.hb + bin] : 0.0;
    if (kSink && c.w.sinkDelta) {
      unsigned pm = __ballot_sync(0xffffffffu, present);
      if (present) {
        int k = emitted + __popc(pm & ((1u << lane) - 1u));
        c.w.sinkDelta[k] = val;
        c.w.sinkCut[k] = c.d.uniq[cov][bin];
      }
      emitted += __popc(pm);
    }
    int sel;
    track_chunk(present, val, have, best, sel);
    if (sel >= 0) bestBin = (ch << 5) + sel;
  }
  if (lane == 0) {
    s->cutsSeen = emitted;
    if (bestBin >= 0) { s->have = 1; s->bestDelta = best; s->bestCand = ci; s->bestCode = (uint32_t)bestBin; }
  }
  __syncwarp();
}

// warp 0: one histogram candidate, in draw order
__device__ inline void track_combine_one(const NodeCtx& c, int ci) {
  GrowShared* s = c.m.s;
  const int lane = lane_id();
  const int np = s->trkNp[ci];
  const int cov = s->candCov[ci];
  if (np < 2) {                       // constant in this node: not permissible below (rf.c:15004-15007)
    if (lane == 0) s->perm[cov >> 5] &= ~(1u << (cov & 31));
    __syncwarp();
    return;
  }
  if (lane == 0) s->candCount += (unsigned long long)s->trkCuts[ci];
  if (s->trkClean[ci] && !(kSink && c.w.sinkDelta)) {
    if (lane == 0) {
      const double M = s->trkMax[ci];
      if (!s->have || (M - s->bestDelta) > kEpsilon) { s->have = 1; s->bestDelta = M; s->bestCand = ci; s->bestCode = (uint32_t)s->trkMaxBin[ci]; }
    }
    __syncwarp();
    return;
  }
  track_exact(c, ci);
}

// warp 0, nodes whose candidates are all histogram candidates: lane = candidate; the draw-order
// replay of the tracker is a short register loop, the exact walk only runs for near-ties / NaN
__device__ inline void track_combine_all(const NodeCtx& c, int nc) {
  GrowShared* s = c.m.s;
  const int lane = lane_id();
  const bool mine = lane < nc;
  const int np = mine ? s->trkNp[lane] : 0;
  const bool clean = mine ? (s->trkClean[lane] != 0) : true;
  const bool slow = (kSink && c.w.sinkDelta) || __any_sync(0xffffffffu, mine && np >= 2 && !clean);
  if (lane == 0) s->needSlow = slow ? 1 : 0;              // near-ties / NaN / emit mode: the caller walks the candidates one by one
  if (slow) return;
  if (mine && np < 2) {                                   // constant in this node (rf.c:15004-15007)
    const int cov = s->candCov[lane];
    atomicAnd(&s->perm[cov >> 5], ~(1u << (cov & 31)));
  }
  const int cuts = warp_sum((mine && np >= 2) ? s->trkCuts[lane] : 0);
  const double M = mine ? s->trkMax[lane] : 0.0;
  int have = s->have; double best = s->bestDelta; int bc = -1;
#pragma unroll 1
  for (int ci = 0; ci < nc; ci++) {
    const double Mi = __shfl_sync(0xffffffffu, M, ci);
    const int npi = __shfl_sync(0xffffffffu, np, ci);
    if (npi >= 2 && (!have || (Mi - best) > kEpsilon)) { have = 1; best = Mi; bc = ci; }
  }
  if (lane == 0) {
    s->candCount += (unsigned long long)cuts;
    if (bc >= 0) { s->have = 1; s->bestDelta = best; s->bestCand = bc; s->bestCode = (uint32_t)s->trkMaxBin[bc]; }
  }
  __syncwarp();
}

// ---- sorted-walk path for covariates with more than kHistBins distinct values ----------------
__device__ inline void radix_pass(const uint32_t* kin, const uint32_t* vin, uint32_t* kout, uint32_t* vout,
                                  uint32_t len, int shift, GrowShared* s, uint32_t* rbase, uint32_t* dOld, void* wdRaw) {
  const int tid = vtid(), lane = lane_id(), wid = warp_id();
  __syncthreads();
  constexpr int kDigPer = (256 + kThreads - 1) / kThreads;   // digits owned by one thread (consecutive); threads past digit 255 own none
  const bool owns = tid * kDigPer < 256;
  for (int q = 0; q < kDigPer; q++) if (owns) rbase[tid * kDigPer + q] = 0;
  __syncthreads();
  for (uint32_t i = tid; i - (uint32_t)lane < len; i += kThreads) {  // digit histogram, one atomic per distinct digit of a warp's 32 entries
    const bool v = i < len;
    const uint32_t dg = v ? ((kin[i] >> shift) & 255u) : (0x100u + (uint32_t)lane);
    const unsigned pr = __match_any_sync(0xffffffffu, dg);
    if (v && (pr & ((1u << lane) - 1u)) == 0u) atomicAdd(&rbase[dg], (uint32_t)__popc(pr));
  }
  __syncthreads();
  uint32_t cnt[kDigPer]; uint32_t mine = 0;
  for (int q = 0; q < kDigPer; q++) { cnt[q] = owns ? rbase[tid * kDigPer + q] : 0u; mine += cnt[q]; }
  unsigned long long tot;
  unsigned long long inc = block_incl_scan<unsigned long long>((unsigned long long)mine, s->scrU, tot);
  __syncthreads();
  uint32_t run = (uint32_t)(inc - mine);
  for (int q = 0; q < kDigPer; q++) { if (owns) rbase[tid * kDigPer + q] = run; run += cnt[q]; }
  __syncthreads();
  // stable scatter, one tile of kThreads entries at a time: the rank of an entry among its tile's entries of the same digit is
  // its rank inside the warp (__match_any) plus the counts of the warps before it (a [warp][digit] byte table, prefixed by
  // the digit's own thread) -- three barriers per tile (the warps used to take turns: one barrier per warp and tile)
  uint8_t* wd = (uint8_t*)wdRaw;                                     // [kWarps][256]
  const uint32_t ntiles = (len + kThreads - 1) / kThreads;
  for (uint32_t t = 0; t < ntiles; t++) {
    uint32_t i = t * kThreads + tid;
    bool valid = i < len;
    uint32_t key = valid ? kin[i] : 0u;
    uint32_t val = valid ? vin[i] : 0u;
    uint32_t dgt = valid ? ((key >> shift) & 255u) : (0x100u + (uint32_t)lane);
    for (int q = tid; q < kWarps * 64; q += kThreads) ((uint32_t*)wd)[q] = 0u;
    __syncthreads();
    const unsigned peers = __match_any_sync(0xffffffffu, dgt);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank == 0) wd[wid * 256 + dgt] = (uint8_t)__popc(peers);              // 1..32
    __syncthreads();
    for (int q = 0; q < kDigPer; q++) {
      if (owns) {
        const int dg = tid * kDigPer + q;
        const uint32_t old = rbase[dg];
        uint32_t run = 0;
        for (int wv = 0; wv < kWarps; wv++) {                          // counts -> exclusive prefix over the warps (<= 224: a byte)
          const uint32_t cell = wd[wv * 256 + dg];
          wd[wv * 256 + dg] = (uint8_t)run;
          run += cell;
        }
        dOld[dg] = old;
        rbase[dg] = old + run;
      }
    }
    __syncthreads();
    if (valid) { const uint32_t pos = dOld[dgt] + (uint32_t)wd[wid * 256 + dgt] + rank; kout[pos] = key; vout[pos] = val; }
    __syncthreads();
  }
  __syncthreads();
}

// per-stratum totals of the node (needed up front by the sorted walk)
__device__ inline void stratum_totals(const NodeCtx& c, uint32_t len) {
  GrowSmem& m = c.m;
  GrowShared* s = m.s;
  for (int k = 1; k <= c.Kloop; k++) {
    uint32_t s0 = m.segpos[k], s1 = m.segpos[k + 1];
    unsigned long long ne = 0; double vT = 0.0, vW = 0.0;
    for (uint32_t i = s0 + vtid(); i < s1; i += kThreads) {
      uint32_t ent = c.list[i];
      uint32_t row = ent & kRowMask, mult = (ent >> kRowBits) + 1u;
      uint32_t e = c.d.ev[row];
      double t = c.unit ? 1.0 : c.d.rt[row];
      ne += ((unsigned long long)mult << 32) | (unsigned long long)(e ? mult : 0u);
      vT += (double)mult * t;
      if (e) vW += (double)mult * t;
    }
    unsigned long long tne = block_sum<unsigned long long>(ne, s->scrU);
    double tT = block_sum<double>(vT, s->scrD);
    double tW = block_sum<double>(vW, s->scrD);
    if (vtid() == 0) {
      m.PN[k] = (uint32_t)(tne >> 32); m.PE[k] = (uint32_t)(tne & 0xffffffffu);
      m.PT[k] = tT; m.PW[k] = tW;
    }
  }
  (void)len;
  __syncthreads();
}

// stable LSD radix sort of (code, position) of the node's entries for candidate ci, in the tree's global scratch;
// ks / vs come back pointing at the sorted keys / positions
__device__ inline void sort_node_codes(const GrowWork& w, int tree, GrowShared* s, GrowSmem& m, const uint32_t* list, uint32_t len, int ci,
                                       uint32_t*& ks, uint32_t*& vs) {
  const int tid = vtid();
  const size_t n = w.listStride;
  uint32_t* kA = w.sortbuf + (size_t)tree * 4 * n;
  uint32_t* kB = kA + n; uint32_t* vA = kB + n; uint32_t* vB = vA + n;
  const void* ptr = s->candPtr[ci]; const int wd = s->candW[ci]; const int Dc = s->candD[ci];
  for (uint32_t i = tid; i < len; i += kThreads) {
    uint32_t row = list[i] & kRowMask;
    kA[i] = load_code(ptr, wd, row);
    vA[i] = i;
  }
  int bits = 32 - __clz(Dc - 1);
  int npass = (bits + 7) >> 3;
  ks = kA; vs = vA; uint32_t *kd = kB, *vd = vB;
  for (int ps = 0; ps < npass; ps++) {
    radix_pass(ks, vs, kd, vd, len, 8 * ps, s, m.radixBase, m.tileKey, m.tileTR);   // (tileTL is the draw step's covariate pool: the nsplit prologue sorts too)
    uint32_t* t1 = ks; ks = kd; kd = t1;
    uint32_t* t2 = vs; vs = vd; vd = t2;
  }
  __syncthreads();
}

__device__ inline void sort_candidate(const NodeCtx& c, int ci, uint32_t len) {
  GrowSmem& m = c.m;
  GrowShared* s = m.s;
  const int tid = vtid(), lane = lane_id(), wid = warp_id();
  const int cov = s->candCov[ci];
  uint32_t *ks, *vs;
  sort_node_codes(c.w, c.tree, s, m, c.list, len, ci, ks, vs);
  RF_PROFC(14);
  // nsplit > 0 on a covariate with more distinct values than nsplit: only the DRAWN cut ordinals (ascending, 1-based rank of
  // the distinct value the cut follows; node_prologue_nsplit) are candidates (rf.c:14533-14564)
  const bool drawn = kSeq && c.g.seqDraw && s->candMask[ci] && s->facN[ci] > 0;
  const uint32_t* sel = m.facMask + (size_t)ci * m.facCuts;
  const int nsel = drawn ? s->facN[ci] : 0;
  uint32_t cutBase = 0;                                              // cuts in the tiles before this one
  __syncthreads();
  // parent statistic and carries
  double statP = 0.0;
  for (int a = 1; a <= c.Kloop; a++)
    if (m.PE[a] > 0) statP += side_term(c.bayes, (double)m.PE[a], (double)m.PN[a], m.PT[a], m.PW[a], c.g.alpha, c.beta);
  for (int a = 1 + tid; a <= c.Kloop; a += kThreads) {
    m.cN[a] = 0; m.cE[a] = 0; m.cT[a] = 0.0; m.cW[a] = 0.0; m.cTL[a] = 0.0;
    m.cTR[a] = (m.PE[a] > 0) ? side_term(c.bayes, (double)m.PE[a], (double)m.PN[a], m.PT[a], m.PW[a], c.g.alpha, c.beta) : 0.0;
  }
  __syncthreads();
  const double wgt = c.g.split_weight ? c.g.split_weight[cov] : 1.0;
  const uint32_t ntiles = (len + kThreads - 1) / kThreads;
  int have = s->have; double best = s->bestDelta; uint32_t bestKey = 0; int gotBest = 0;   // warp 0's view
  int totalCuts = 0;
  int emitted = s->cutsSeen;
  // The walk, one tile of kThreads code-sorted entries at a time.  A cut's statistic is the sum over strata of the terms of
  // each stratum's LAST row at or before the cut, and a row's terms need its own stratum's running statistics -- sixteen
  // interleaved prefix sums along one order.  Per tile the entries are therefore regrouped by stratum (a stable rank:
  // __match_any inside the warp, per-warp counts across warps), ONE segmented scan gives every row its stratum's
  // statistics (carried across tiles per stratum), the rule terms are evaluated once per row, and a row hands the cut
  // order the CHANGE it makes to its stratum's terms; one plain scan over the tile on top of the exact sum of the strata's
  // carried terms gives every position's statistic.  (It used to be three block scans and eleven barriers per stratum and
  // tile, every entry taking part in every stratum's scans: 65 % of the kernel on a continuous 1 M x 50 table.)
  const int KB = c.Kloop;                                            // stratum bins of a tile; invalid lanes keep their own slot
  unsigned long long* qNE = (unsigned long long*)m.tileTL;
  double* qT = m.tileTR; double* qW = m.tileDelta; double* qInc = m.tileTL;
  uint16_t* qS = m.tileS; uint8_t* qO = m.tileO; uint16_t* segStart = m.tileSeg; uint8_t* wcnt = m.tileWcnt;
  for (uint32_t t = 0; t < ntiles; t++) {
    uint32_t j = t * kThreads + tid;
    bool valid = j < len;
    uint32_t key = 0; int a0 = KB;
    unsigned long long vne = 0ull; double vT = 0.0, vW = 0.0;
    bool cut = false;
    if (valid) {
      key = ks[j];
      const uint32_t ent = c.list[vs[j]];
      const uint32_t row = ent & kRowMask, mult = (ent >> kRowBits) + 1u;
      const RowRec rr = c.d.rowrec[row];
      const uint32_t e = rr.es & 1u;
      a0 = c.strat ? (int)(rr.es >> 8) : 0;
      const double tt = c.unit ? 1.0 : rr.rt;
      vne = ((unsigned long long)mult << 32) | (unsigned long long)(e ? mult : 0u);
      vT = (double)mult * tt; vW = e ? vT : 0.0;
      cut = (j + 1 < len) && (ks[j + 1] != key);
    }
    // (1) stable rank by stratum inside the tile
    for (int i = tid; i < kWarps * KB; i += kThreads) wcnt[i] = 0;
    double bsum = 0.0;                                               // exact sum of the strata's carried terms (strata with events)
    for (int a = 1 + tid; a <= KB; a += kThreads) if (m.PE[a] > 0) bsum += m.cTL[a] + m.cTR[a];
    const double base = block_sum<double>(bsum, s->scrD);            // (its barriers also publish the zeroed counts)
    const unsigned peers = __match_any_sync(0xffffffffu, a0);
    const int rankW = __popc(peers & ((1u << lane) - 1u));
    if (valid && rankW == 0) wcnt[wid * KB + a0] = (uint8_t)__popc(peers);
    __syncthreads();
    uint32_t tot = 0;
    if (tid < KB) {
      uint32_t run = 0;
      for (int w2 = 0; w2 < kWarps; w2++) { const uint32_t cw = wcnt[w2 * KB + tid]; wcnt[w2 * KB + tid] = (uint8_t)run; run += cw; }
      tot = run;
    }
    uint32_t totAll;
    const uint32_t incl = block_incl_scan<uint32_t>(tot, (uint32_t*)s->scrI, totAll);
    if (tid < KB) segStart[tid] = (uint16_t)(incl - tot);
    __syncthreads();
    const int ps = valid ? (int)segStart[a0] + (int)wcnt[wid * KB + a0] + rankW : tid;   // (valid entries fill [0, totAll), the tail keeps its slots)
    qNE[ps] = vne; qT[ps] = vT; qW[ps] = vW; qS[ps] = (uint16_t)a0; qO[ps] = (uint8_t)tid;
    __syncthreads();
    // (2) segmented inclusive scan in stratum order: thread = position
    const int sp = qS[tid];
    const bool pvalid = sp < KB;
    const bool head = tid == 0 || qS[tid - 1] != sp;
    unsigned long long ine = qNE[tid]; double iT = qT[tid], iW = qW[tid];
    const unsigned hbits = __ballot_sync(0xffffffffu, head);
    const unsigned hm = hbits & (0xFFFFFFFFu >> (31 - lane));
    const int startLane = hm ? 31 - __clz(hm) : -1;                   // -1: my segment began in an earlier warp
    const int lim = startLane < 0 ? 0 : startLane;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long une = __shfl_up_sync(0xffffffffu, ine, o);
      const double uT = __shfl_up_sync(0xffffffffu, iT, o), uW = __shfl_up_sync(0xffffffffu, iW, o);
      if (lane - o >= lim) { ine += une; iT += uT; iW += uW; }
    }
    if (lane == 31) { s->scrU[wid] = ine; s->scrD[wid] = iT; s->scrD[kWarps + wid] = iW; }
    if (lane == 0) s->scrI[wid] = hbits != 0u ? 1 : 0;
    __syncthreads();
    if (startLane < 0) {                                              // add the open segment's part in the warps before mine
      for (int w2 = wid - 1; w2 >= 0; w2--) {
        ine += s->scrU[w2]; iT += s->scrD[w2]; iW += s->scrD[kWarps + w2];
        if (s->scrI[w2]) break;
      }
    }
    // (3) the row's stratum statistics (carried across tiles), its terms, and the change it makes to its stratum's terms
    const int a = sp + 1;
    double tL = 0.0, tR = 0.0; uint32_t LN = 0, LE = 0; double LT = 0.0, LW = 0.0;
    const bool ev = pvalid && m.PE[a] > 0;
    if (pvalid) { LN = (uint32_t)(ine >> 32) + m.cN[a]; LE = (uint32_t)(ine & 0xffffffffu) + m.cE[a]; LT = iT + m.cT[a]; LW = iW + m.cW[a]; }
    if (__any_sync(0xffffffffu, ev)) {
      const double xL = side_term(c.bayes, (double)LE, (double)LN, LT, LW, c.g.alpha, c.beta);
      const double xR = ev ? side_term(c.bayes, (double)(m.PE[a] - LE), (double)(m.PN[a] - LN), m.PT[a] - LT, m.PW[a] - LW, c.g.alpha, c.beta) : 0.0;
      if (ev) { tL = xL; tR = xR; }
    }
    __syncthreads();                                                  // (every thread has read its qT / qW / scratch)
    qT[tid] = tL; qW[tid] = tR;
    __syncthreads();
    double inc = 0.0;
    if (ev) {
      const double pL = head ? m.cTL[a] : qT[tid - 1], pR = head ? m.cTR[a] : qW[tid - 1];
      inc = (tL - pL) + (tR - pR);
    }
    const bool tail = pvalid && (tid == kThreads - 1 || qS[tid + 1] != sp);
    __syncthreads();                                                  // (carries and neighbours read before they change)
    qInc[qO[tid]] = inc;                                              // back to the cut order
    if (tail) { m.cN[a] = LN; m.cE[a] = LE; m.cT[a] = LT; m.cW[a] = LW; if (ev) { m.cTL[a] = tL; m.cTR[a] = tR; } }
    __syncthreads();
    double tsum;
    const double sIncl = block_incl_scan<double>(qInc[tid], s->scrD, tsum);
    const double stat = base + sIncl;                                 // sum over strata of (left term + right term) at this position
    if (drawn) {                                                     // keep a cut only if its ordinal was drawn
      int tot;
      const int inc = block_incl_scan<int>(cut ? 1 : 0, s->scrI, tot);
      if (cut) {
        const uint32_t ord = cutBase + (uint32_t)inc;
        int lo = 0, hi = nsel;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (sel[mid] < ord) lo = mid + 1; else hi = mid; }
        cut = lo < nsel && sel[lo] == ord;
      }
      cutBase += (uint32_t)tot;
      __syncthreads();
    }
    __syncthreads();                                                  // (the scan has read every slot of qInc / the scratch)
    m.tileDelta[tid] = cut ? (stat - statP) * wgt : 0.0;
    m.tileCut[tid] = cut ? 1 : 0;
    m.tileKey[tid] = key;
    __syncthreads();
    if (wid == 0) {
      for (int ch = 0; ch < kWarps; ch++) {
        int q = (ch << 5) + lane;
        bool present = m.tileCut[q] != 0;
        double val = m.tileDelta[q];
        unsigned pm = __ballot_sync(0xffffffffu, present);
        if (kSink && c.w.sinkDelta && present) {
          int k = emitted + __popc(pm & ((1u << lane) - 1u));
          c.w.sinkDelta[k] = val;
          c.w.sinkCut[k] = c.d.uniq[cov][m.tileKey[q]];
        }
        emitted += __popc(pm);
        totalCuts += __popc(pm);
        int sel;
        track_chunk(present, val, have, best, sel);
        if (sel >= 0) { bestKey = m.tileKey[(ch << 5) + sel]; gotBest = 1; }
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    s->cutsSeen = emitted;
    if (totalCuts == 0) {
      s->perm[cov >> 5] &= ~(1u << (cov & 31));
    } else {
      s->candCount += (unsigned long long)totalCuts;
      if (gotBest) { s->have = 1; s->bestDelta = best; s->bestCand = ci; s->bestCode = bestKey; }
    }
  }
  __syncthreads();
}


// ---- unordered factors (MWCP splits, rf.c:14452-14518) ----------------------------------------------------------
// A factor candidate's cuts are LEFT sets of the r levels present in the node.  Deterministic mode: all 2^(r-1) - 1
// complementary pairs, by cardinality of the LEFT set (1 .. floor(r/2)), lexicographic inside a cardinality, only the
// first half of the middle cardinality when r is even -- the order bookPair (rf.c:1498-1530) books them in, pinned
// against the live reference's tables for r = 2..12 (tests/test_oracle.py).  Cut j is UNRANKED on the fly (no table):
// thread j finds its cardinality group, then the combination with that lexicographic rank.  Masks are over rank
// CODES (bit = code of the level in this table); the forest record converts them to absolute levels (compact_kernel).
static __device__ __constant__ const uint8_t kChoose[10][10] = {
  {1,0,0,0,0,0,0,0,0,0}, {1,1,0,0,0,0,0,0,0,0}, {1,2,1,0,0,0,0,0,0,0}, {1,3,3,1,0,0,0,0,0,0}, {1,4,6,4,1,0,0,0,0,0},
  {1,5,10,10,5,1,0,0,0,0}, {1,6,15,20,15,6,1,0,0,0}, {1,7,21,35,35,21,7,1,0,0}, {1,8,28,56,70,56,28,8,1,0}, {1,9,36,84,126,126,84,36,9,1}};
constexpr int kMaxFactorLevels = 9;          // 2^8 - 1 = 255 cuts fit one candidate's accumulators (hb <= 256)

// the code of the k-th (0-based) level present in the node
__device__ __forceinline__ uint32_t nth_present_code(uint32_t presW, int k) {
  for (int i = 0; i < k; i++) presW &= presW - 1u;
  return (uint32_t)__ffs(presW) - 1u;
}

__device__ inline uint32_t factor_cut_mask(int r, int j, uint32_t presW) {
  int g = 1, idx = j;
  while (true) {                                                     // cardinality group of cut j
    int gs = kChoose[r][g];
    if (!(r & 1) && g == (r >> 1)) gs >>= 1;
    if (idx < gs) break;
    idx -= gs; g++;
  }
  uint32_t mask = 0u;
  int x = 1;                                                         // relative levels 1..r
  for (int i = 0; i < g; i++) {                                      // lexicographic unranking of a g-combination
    while (true) {
      const int cnt = kChoose[r - x][g - i - 1];                     // combinations that start with level x at position i
      if (idx < cnt) break;
      idx -= cnt; x++;
    }
    mask |= 1u << nth_present_code(presW, x - 1);
    x++;
  }
  return mask;
}

// One stratum of a factor candidate: lane l holds the stratum's (rows, events, exposure, event exposure) at code l.
// Adds the stratum's left / right terms of every cut to sL / sR (cut j at index j) and returns its parent term.
// A stratum without events contributes exactly 0 to every cut and to the parent (see side_term).
__device__ inline double factor_accumulate(const NodeCtx& c, int ci, uint32_t vN, uint32_t vE, double vT, double vW) {
  GrowSmem& m = c.m;
  GrowShared* s = m.s;
  const int lane = lane_id();
  const uint32_t tE = warp_sum(vE);
  if (tE == 0u) return 0.0;
  const uint32_t tN = warp_sum(vN);
  const double tT = c.unit ? (double)tN : warp_sum(vT), tW = c.unit ? (double)tE : warp_sum(vW);
  if (c.unit) { vT = (double)vN; vW = (double)vE; }
  const double termP = side_term(c.bayes, (double)tE, (double)tN, tT, tW, c.g.alpha, c.beta);
  const int ncuts = s->facN[ci], Dc = s->candD[ci];
  const uint32_t* masks = m.facMask + (size_t)ci * m.facCuts;
#pragma unroll 1
  for (int base = 0; base < ncuts; base += 32) {
    const int j = base + lane;
    const uint32_t mk = j < ncuts ? masks[j] : 0u;
    uint32_t LN = 0, LE = 0; double LT = 0.0, LW = 0.0;
#pragma unroll 1
    for (int l = 0; l < Dc; l++) {                                   // ascending level: the order the plugin's sorted rows arrive in
      const uint32_t n_l = __shfl_sync(0xffffffffu, vN, l), e_l = __shfl_sync(0xffffffffu, vE, l);
      const double t_l = __shfl_sync(0xffffffffu, vT, l), w_l = __shfl_sync(0xffffffffu, vW, l);
      if ((mk >> l) & 1u) { LN += n_l; LE += e_l; LT += t_l; LW += w_l; }
    }
    const double tL = side_term(c.bayes, (double)LE, (double)LN, LT, LW, c.g.alpha, c.beta);
    const double tR = side_term(c.bayes, (double)(tE - LE), (double)(tN - LN), tT - LT, tW - LW, c.g.alpha, c.beta);
    if (j < ncuts) { m.sL[ci * m.hb + j] += tL; m.sR[ci * m.hb + j] += tR; }
  }
  return termP;
}

// the tracker sees a factor candidate as cuts 0 .. ncuts-1 "present" plus one closing bin it never selects
__device__ __forceinline__ void factor_set_presence(const NodeCtx& c, int ci) {
  const int lane = lane_id();
  const int ncuts = c.m.s->facN[ci];
  if (lane < 8) {
    const int lo = lane * 32;
    uint32_t w = 0u;
    if (ncuts > 0 && ncuts + 1 > lo) w = (ncuts + 1 - lo >= 32) ? 0xFFFFFFFFu : ((1u << (ncuts + 1 - lo)) - 1u);
    c.m.pres[ci * 8 + lane] = w;
  }
  __syncwarp();
}

// small node: the candidate's warp builds each event stratum's per-level statistics from the staged rows
__device__ inline void small_factor_candidate(const NodeCtx& c, int ci) {
  GrowSmem& m = c.m;
  GrowShared* s = m.s;
  const int lane = lane_id();
  const int ncuts = s->facN[ci];
  const SmallRec* __restrict__ rec = m.smRec;
  const uint8_t* __restrict__ cptr = m.smCodes + ci * kSmallMax;
  for (int b = lane; b <= ncuts; b += 32) { m.sL[ci * m.hb + b] = 0.0; m.sR[ci * m.hb + b] = 0.0; }
  __syncwarp();
  double statP = 0.0;
  if (ncuts > 0) {
    const int nw = (c.Kloop + 31) >> 5;
#pragma unroll 1
    for (int w8 = 0; w8 < nw; w8++) {
      uint32_t word = s->evMask[w8];
#pragma unroll 1
      while (word) {                                                 // event strata, ascending (sc.c:760-764 sums in this order)
        const int sidx = (w8 << 5) + __ffs(word) - 1;
        word &= word - 1u;
        const uint32_t a0 = m.segpos[sidx], a1 = m.cN[sidx];
        uint32_t vN = 0, vE = 0; double vT = 0.0, vW = 0.0;
#pragma unroll 1
        for (uint32_t k = a0; k < a1; k++) {
          const SmallRec r = rec[k];
          if ((uint32_t)cptr[k] == (uint32_t)lane) {
            vN += r.mult; vT += r.mrt;
            if (r.es & 1u) { vE += r.mult; vW += r.mrt; }
          }
        }
        statP += factor_accumulate(c, ci, vN, vE, vT, vW);
      }
    }
  }
  if (lane == 0) m.statP[ci] = statP;
  factor_set_presence(c, ci);
}

// ---- small nodes (<= kSmallMax entries, histogram candidates only) ---------------------------
// The node is staged in shared memory once -- per entry a 16-byte record (multiplicity * risk time,
// multiplicity, event, stratum) and one code byte per candidate -- with every gather of an entry in
// flight together.  Then warp ci scores candidate ci on its own, with warp-level synchronisation only
// and no atomics, visiting just the strata that contain events in this node (every other cell
// contributes exactly 0 to every rule, see side_term; their rows matter only for which cuts exist).
// The kernel was instruction-fetch bound on its once-per-node code (stall_no_instruction was the top
// stall at two CTAs per SM) and, inside this function, bound by the FP64 divide + log of the rule
// terms, so it is organised to be small and to evaluate as few terms as possible:
//
//  * short strata (<= RF_ROWS_MAXSEG rows, the usual case under the stratified rules): LANE = ROW.
//    A stratum's terms only change at the codes of its own rows, so they are evaluated once per row of
//    an event stratum (32 rows per batch, two term evaluations for the whole batch), written to a
//    per-stratum table, and each cut then adds, stratum by stratum, the entry of the largest code <=
//    the cut (fill-forward by bit search).  Rows holding their stratum's largest code have "left =
//    everything": their left term IS the stratum's parent term.  Tables live in this candidate's
//    slice of the (idle) large-node histograms, G = hs strata per group.
//  * long strata: LANE = CUT.  The warp streams the stratum's records, every lane keeps the left sums
//    of its own cut(s) of the current 64-code window, three to five terms per stratum and window.
//
// Either way a cut's left / right statistics are the sums over strata in ascending order, exactly
// what the histogram + prefix-scan + fill-forward of the large-node path produces.
template <bool TWO>
__device__ __forceinline__ void stream_rows(const SmallRec* __restrict__ rec, const uint8_t* __restrict__ cptr, uint32_t r0, uint32_t r1,
                                            bool unit, uint32_t b0, uint32_t b1, uint32_t& TN, uint32_t& TE, double& TT, double& TW,
                                            uint32_t& LN0, uint32_t& LE0, double& LT0, double& LW0,
                                            uint32_t& LN1, uint32_t& LE1, double& LT1, double& LW1) {
#pragma unroll 4
  for (uint32_t i = r0; i < r1; i++) {
    const SmallRec r = rec[i];
    const uint32_t code = cptr[i];
    const uint32_t mu = r.mult;
    const double ft = unit ? (double)mu : r.mrt;
    const bool ev = (r.es & 1u) != 0u;
    const uint32_t me = ev ? mu : 0u; const double fe = ev ? ft : 0.0;
    TN += mu; TT += ft; TE += me; TW += fe;
    const bool l0 = code <= b0;
    LN0 += l0 ? mu : 0u; LT0 += l0 ? ft : 0.0; LE0 += l0 ? me : 0u; LW0 += l0 ? fe : 0.0;
    if (TWO) {
      const bool l1 = code <= b1;
      LN1 += l1 ? mu : 0u; LT1 += l1 ? ft : 0.0; LE1 += l1 ? me : 0u; LW1 += l1 ? fe : 0.0;
    }
  }
}

// one long stratum, lane = cut; adds its terms to sL / sR and returns its parent term
__device__ inline double stream_stratum(const NodeCtx& c, int ci, int Dc, uint32_t r0, uint32_t r1) {
  GrowSmem& m = c.m;
  const int lane = lane_id();
  const SmallRec* __restrict__ rec = m.smRec;
  const uint8_t* __restrict__ cptr = m.smCodes + ci * kSmallMax;
  double tP = 0.0;
#pragma unroll 1
  for (int base = 0; base < Dc; base += 64) {
    const uint32_t b0 = (uint32_t)(base + lane), b1 = b0 + 32u;
    const bool two = Dc > base + 32;
    uint32_t TN = 0, TE = 0, LN0 = 0, LE0 = 0, LN1 = 0, LE1 = 0;
    double TT = 0.0, TW = 0.0, LT0 = 0.0, LW0 = 0.0, LT1 = 0.0, LW1 = 0.0;
    if (two) stream_rows<true>(rec, cptr, r0, r1, c.unit, b0, b1, TN, TE, TT, TW, LN0, LE0, LT0, LW0, LN1, LE1, LT1, LW1);
    else stream_rows<false>(rec, cptr, r0, r1, c.unit, b0, b1, TN, TE, TT, TW, LN0, LE0, LT0, LW0, LN1, LE1, LT1, LW1);
    if (base == 0) tP = side_term_nb(c.bayes, (double)TE, (double)TN, TT, TW, c.g.alpha, c.beta);
    const double tL0 = side_term_nb(c.bayes, (double)LE0, (double)LN0, LT0, LW0, c.g.alpha, c.beta);
    const double tR0 = side_term_nb(c.bayes, (double)(TE - LE0), (double)(TN - LN0), TT - LT0, TW - LW0, c.g.alpha, c.beta);
    if (b0 < (uint32_t)Dc) { m.sL[ci * m.hb + b0] += tL0; m.sR[ci * m.hb + b0] += tR0; }
    if (two) {
      const double tL1 = side_term_nb(c.bayes, (double)LE1, (double)LN1, LT1, LW1, c.g.alpha, c.beta);
      const double tR1 = side_term_nb(c.bayes, (double)(TE - LE1), (double)(TN - LN1), TT - LT1, TW - LW1, c.g.alpha, c.beta);
      if (b1 < (uint32_t)Dc) { m.sL[ci * m.hb + b1] += tL1; m.sR[ci * m.hb + b1] += tR1; }
    }
  }
  return tP;
}

// A covariate with more than 256 distinct values in the table has at most len <= kLocalMax of them in a
// small node: the candidate's warp gathers the node's wide codes (<= 8 per lane), ranks them among the
// node's distinct values by all-pairs comparison through warp shuffles (len * len / 32 compares), and
// writes the ranks as the node-local byte codes the rest of the small path works on.  The winning rank is
// turned back into a table-wide code by the split path (every entry with that rank holds the same code).
__device__ inline void local_rank_candidate(GrowShared* s, const uint32_t* smEnt, uint8_t* cptr, int ci, uint32_t len) {
  const int lane = lane_id();
  const void* ptr = s->candPtr[ci]; const int wd = s->candW[ci];
  uint32_t wv[kLocalMax / 32];
#pragma unroll
  for (int r = 0; r < kLocalMax / 32; r++) {
    const uint32_t j = (uint32_t)(r * 32 + lane);
    wv[r] = (j < len) ? load_code(ptr, wd, smEnt[j] & kRowMask) : 0xFFFFFFFFu;
  }
  // first[r]: no earlier entry holds the same value; rank[r]: distinct values below mine
  uint32_t firstMask = 0u;
  uint32_t rank[kLocalMax / 32];
#pragma unroll
  for (int r = 0; r < kLocalMax / 32; r++) rank[r] = 0u;
  const int nr = (int)((len + 31u) >> 5);
#pragma unroll
  for (int rk = 0; rk < kLocalMax / 32; rk++) {
    if (rk < nr) {
#pragma unroll 1
      for (int src = 0; src < 32; src++) {
        const uint32_t k = (uint32_t)(rk * 32 + src);
        if (k >= len) break;                                          // warp-uniform
        const uint32_t wk = __shfl_sync(0xffffffffu, wv[rk], src);
        // is entry k the first occurrence of its value?  (any earlier entry equal to it)
        bool dup = false;
#pragma unroll
        for (int r = 0; r < kLocalMax / 32; r++) dup |= (wv[r] == wk) && ((uint32_t)(r * 32 + lane) < k);
        const bool isFirst = !__any_sync(0xffffffffu, dup);
        if (isFirst) {
#pragma unroll
          for (int r = 0; r < kLocalMax / 32; r++) rank[r] += (wk < wv[r]) ? 1u : 0u;
          if (lane == src) firstMask |= 1u << rk;
        }
      }
    }
  }
  int np = __popc(firstMask);
  np = warp_sum(np);
#pragma unroll
  for (int r = 0; r < kLocalMax / 32; r++) {
    const uint32_t j = (uint32_t)(r * 32 + lane);
    if (j < len) cptr[j] = (uint8_t)rank[r];
  }
  if (lane == 0) { s->candD[ci] = np; s->candHist[ci] = 1; s->candLocal[ci] = 1; }   // node-local cardinality / byte codes from here on
  __syncwarp();
}

__device__ inline void small_candidate(const NodeCtx& c, int ci, uint32_t len) {
  GrowSmem& m = c.m;
  GrowShared* s = m.s;
  const int lane = lane_id();
  const int Dc = s->candD[ci];
  const int nwords = (Dc + 31) >> 5;
  const SmallRec* __restrict__ rec = m.smRec;
  const uint8_t* __restrict__ cptr = m.smCodes + ci * kSmallMax;
  const int nw = (c.Kloop + 31) >> 5;
  const int G = m.ts;                                              // table slots = strata per group
  double* tabL = m.tabL + (size_t)ci * G * m.hb;
  double* tabR = m.tabR + (size_t)ci * G * m.hb;
  uint32_t* mk = m.smk + (size_t)ci * G * 8;               // [slot][8] codes that have rows in the slot's stratum
  if (lane < 8) m.pres[ci * 8 + lane] = 0u;
  for (int b = lane; b < Dc; b += 32) { m.sL[ci * m.hb + b] = 0.0; m.sR[ci * m.hb + b] = 0.0; }
  __syncwarp();
#pragma unroll 1
  for (int base = 0; base < Dc; base += 64) {                      // which codes occur in the node (the cuts that exist)
    uint32_t plo = 0u, phi = 0u;
    for (uint32_t i = (uint32_t)lane; i < len; i += 32u) {
      const uint32_t rel = (uint32_t)cptr[i] - (uint32_t)base;
      if (rel < 32u) plo |= 1u << rel; else if (rel < 64u) phi |= 1u << (rel - 32u);
    }
    plo = __reduce_or_sync(0xffffffffu, plo); phi = __reduce_or_sync(0xffffffffu, phi);
    if (lane == 0) { m.pres[ci * 8 + (base >> 5)] = plo; if (Dc > base + 32) m.pres[ci * 8 + (base >> 5) + 1] = phi; }
  }
  RF_PROFC(5);
  double statP = 0.0;
  int w8 = 0; uint32_t word = s->evMask[0];
  uint32_t bigR0 = 0u, bigR1 = 0u;                                 // a long stratum met while collecting a group: handled after it
#pragma unroll 1
  while (true) {
    // next group: up to G short strata with events, ascending; a long stratum ends the group
    uint32_t r0[RF_MAX_SLOTS], r1[RF_MAX_SLOTS]; double tP[RF_MAX_SLOTS];
    int ng = 0;
#pragma unroll
    for (int q = 0; q < RF_MAX_SLOTS; q++) {
      r0[q] = 0u; r1[q] = 0u; tP[q] = 0.0;
      if (q < G && bigR1 == 0u) {
        while (!word && ++w8 < nw) word = s->evMask[w8];
        if (word) {
          const int sidx = (w8 << 5) + __ffs(word) - 1;
          word &= word - 1u;
          const uint32_t a0 = m.segpos[sidx], a1 = m.cN[sidx];
          if (a1 - a0 > (uint32_t)RF_ROWS_MAXSEG) { bigR0 = a0; bigR1 = a1; }
          else { r0[q] = a0; r1[q] = a1; ng = q + 1; }
        }
      }
    }
    if (ng == 0 && bigR1 == 0u) break;
    if (ng > 0) {
    if (lane < G * 8) mk[lane] = 0u;
    __syncwarp();
    uint32_t R = 0;
#pragma unroll
    for (int q = 0; q < RF_MAX_SLOTS; q++) R += r1[q] - r0[q];
#pragma unroll 1
    for (uint32_t bb = 0; bb < R; bb += 32u) {
      const uint32_t idx = bb + (uint32_t)lane;
      int myG = -1; uint32_t myJ = 0u, lo = 0u, hi = 0u, cum = 0u;
#pragma unroll
      for (int q = 0; q < RF_MAX_SLOTS; q++) {
        const uint32_t sz = r1[q] - r0[q];
        if (myG < 0 && idx < cum + sz) { myG = q; myJ = r0[q] + (idx - cum); lo = r0[q]; hi = r1[q]; }
        cum += sz;
      }
      const uint32_t myCode = (myG >= 0) ? (uint32_t)cptr[myJ] : 0u;
      uint32_t TN = 0, TE = 0, LN = 0, LE = 0;
      double TT = 0.0, TW = 0.0, LT = 0.0, LW = 0.0;
      for (uint32_t k = lo; k < hi; k++) {                          // my stratum's rows
        const SmallRec r = rec[k];
        const uint32_t mu = r.mult;
        const double ft = c.unit ? (double)mu : r.mrt;
        const bool ev = (r.es & 1u) != 0u;
        const bool le = (uint32_t)cptr[k] <= myCode;
        const uint32_t me = ev ? mu : 0u; const double fe = ev ? ft : 0.0;
        TN += mu; TT += ft; TE += me; TW += fe;
        LN += le ? mu : 0u; LT += le ? ft : 0.0; LE += le ? me : 0u; LW += le ? fe : 0.0;
      }
      const double tL = side_term_nb(c.bayes, (double)LE, (double)LN, LT, LW, c.g.alpha, c.beta);
      const double tR = side_term_nb(c.bayes, (double)(TE - LE), (double)(TN - LN), TT - LT, TW - LW, c.g.alpha, c.beta);
      if (myG >= 0) {
        tabL[myG * m.hb + (int)myCode] = tL; tabR[myG * m.hb + (int)myCode] = tR;
        atomicOr(&mk[myG * 8 + (int)(myCode >> 5)], 1u << (myCode & 31u));
      }
#pragma unroll
      for (int q = 0; q < RF_MAX_SLOTS; q++) {
        const unsigned bal = __ballot_sync(0xffffffffu, myG == q && LN == TN);
        if (bal) tP[q] = __shfl_sync(0xffffffffu, tL, __ffs(bal) - 1);
      }
    }
    __syncwarp();
    RF_PROFC(7);
#pragma unroll 1
    for (int wq = 0; wq < nwords; wq++) {                          // lane = cut: add this group's strata in ascending order
      const int q = (wq << 5) + lane;
      if (q < Dc) {
        double aL = m.sL[ci * m.hb + q], aR = m.sR[ci * m.hb + q];
#pragma unroll
        for (int g2 = 0; g2 < RF_MAX_SLOTS; g2++) {
          if (g2 < ng) {
            int src = -1;
            for (int w2 = wq; w2 >= 0 && src < 0; w2--) {
              uint32_t mw = mk[g2 * 8 + w2];
              if (w2 == wq) mw &= 0xFFFFFFFFu >> (31 - lane);
              if (mw) src = (w2 << 5) + 31 - __clz(mw);
            }
            if (src >= 0) { aL += tabL[g2 * m.hb + src]; aR += tabR[g2 * m.hb + src]; }
            else aR += tP[g2];
          }
        }
        m.sL[ci * m.hb + q] = aL; m.sR[ci * m.hb + q] = aR;
      }
    }
#pragma unroll
    for (int q = 0; q < RF_MAX_SLOTS; q++) if (q < ng) statP += tP[q];
    __syncwarp();
    RF_PROFC(15);
    }
    if (bigR1 != 0u) {
      statP += stream_stratum(c, ci, Dc, bigR0, bigR1);
      bigR0 = 0u; bigR1 = 0u;
      __syncwarp();
    }
  }
  if (lane == 0) m.statP[ci] = statP;
  __syncwarp();
}

__device__ inline void score_small(NodeCtx& c) {
  GrowSmem& m = c.m;
  GrowShared* s = m.s;
  const int tid = vtid(), lane = lane_id(), wid = warp_id();
  const int nc = s->nc;
  const uint32_t len = s->cur.len;
  const DevData& d = c.d;
  const GrowWork& w = c.w;
  // A node of at most subMax entries stages its entries' WHOLE ROWS once (entry word, 16-byte record, every byte code
  // of the row from the row-major mirror: tilePitch / 16 coalesced 16-byte loads per entry).  The subtree below it is
  // then grown out of shared memory: each of its nodes fetches the codes of its own drawn covariates from the tile,
  // partitions by permuting 16-bit slot numbers, and never touches the lists or the table in global memory again.
  // (Deep in a tree every per-node gather used to be one DRAM round trip per level for one byte per 32-byte sector.)
  if (kSub && m.subMax > 0 && !s->inSub && len <= (uint32_t)m.subMax && !(kSink && w.sinkDelta)) {
    const int chunks = m.tilePitch >> 4;
    for (uint32_t q = (uint32_t)tid; q < len * (uint32_t)chunks; q += kThreads) {
      const uint32_t j = q / (uint32_t)chunks, ch = q - j * (uint32_t)chunks;
      const uint32_t row = c.list[j] & kRowMask;
      ((uint4*)(m.stTile + (size_t)j * m.tilePitch))[ch] = ((const uint4*)(d.rowcodes + (size_t)row * (size_t)d.rowPitch))[ch];
    }
    for (uint32_t j = (uint32_t)tid; j < len; j += kThreads) {
      const uint32_t ent = c.list[j];
      const uint32_t row = ent & kRowMask, mult = (ent >> kRowBits) + 1u;
      const RowRec rr = d.rowrec[row];
      SmallRec r; r.mrt = (double)mult * rr.rt; r.mult = mult; r.es = (rr.es & 1u) | (c.strat ? (rr.es & 0xFFFFFF00u) : 0u);
      m.stRow[j] = ent; m.stRec[j] = r; m.stIdx[j] = (uint16_t)j;
    }
    __syncthreads();                       // every thread has read inSub == 0 (it is inside this block) before the flag flips
    if (tid == 0) { s->inSub = 1; s->subSp = s->sp; s->subStart = s->cur.start; }
    __syncthreads();
    RF_PROF(14);
  }
  const bool sub = kSub && s->inSub != 0;
  const uint32_t soff = sub ? s->cur.start - s->subStart : 0u;
  const int cstride = s->candStride;
  RF_ASSERT(!sub || (s->cur.start >= s->subStart && soff + len <= (uint32_t)m.subMax), "staged subtree: tree %d node %u start %u len %u outside the tile (subStart %u subMax %d sp %u subSp %u)\n", c.tree, s->cur.nodeID, s->cur.start, len, s->subStart, m.subMax, s->sp, s->subSp);
  // stage the node: one thread per entry, all of an entry's gathers issued before the first shared-memory store.
  // With the row-major mirror an entry costs one 16-byte record + the sectors of ONE code row instead of three
  // per-row arrays + one 32-byte sector per candidate column; inside a staged subtree it costs no global access.
  for (uint32_t j0 = (uint32_t)(tid - lane); j0 < len; j0 += kThreads) {
    const uint32_t j = j0 + lane;
    const bool valid = j < len;
    uint32_t ent, e, sidx, mult; double mrt;
    uint32_t edgeS = 0xFFFFFFFFu;        // the stratum of the entry before this warp's first / after its last one (segment boundaries)
    const uint8_t* __restrict__ rc = nullptr;
    uint32_t row = 0u;
    if (sub) {
      const uint32_t slot = valid ? (uint32_t)m.stIdx[soff + j] : 0u;
      RF_ASSERT(slot < (uint32_t)m.subMax, "staged subtree: tree %d slot %u at position %u + %u (subMax %d)\n", c.tree, slot, soff, j, m.subMax);
      const SmallRec r = m.stRec[slot];
      ent = m.stRow[slot]; e = r.es & 1u; sidx = r.es >> 8; mult = r.mult; mrt = r.mrt;
      rc = m.stTile + (size_t)slot * m.tilePitch;
      if (c.strat) {
        if (lane == 0 && j > 0u) edgeS = m.stRec[m.stIdx[soff + j - 1u]].es >> 8;
        if (lane == 31 && j + 1u < len) edgeS = m.stRec[m.stIdx[soff + j + 1u]].es >> 8;
      }
    } else {
      ent = valid ? c.list[j] : 0u;
      RF_ASSERT((ent & kRowMask) < (uint32_t)d.n, "small staging: tree %d entry %u/%u = %08x (start %u buf %u node %u)\n", c.tree, j, len, ent, s->cur.start, s->cur.buf, s->cur.nodeID);
      row = ent & kRowMask; mult = (ent >> kRowBits) + 1u;
      const RowRec rr = d.rowrec[row];
      e = rr.es & 1u; sidx = c.strat ? (rr.es >> 8) : 0u;
      mrt = (double)mult * rr.rt;
      if (c.strat) {
        if (lane == 0 && j > 0u) edgeS = d.rowrec[c.list[j - 1u] & kRowMask].es >> 8;
        if (lane == 31 && j + 1u < len) edgeS = d.rowrec[c.list[j + 1u] & kRowMask].es >> 8;
      }
    }
    const size_t roff = (size_t)row * (size_t)cstride;
    for (int half = 0; half < nc; half += 8) {
      uint32_t raw[8];
#pragma unroll
      for (int q = 0; q < 8; q++)
        raw[q] = (half + q < nc) ? (rc ? (uint32_t)rc[s->candCov[half + q]] : (uint32_t)((const uint8_t*)s->candPtr[half + q])[roff]) : 0u;
#pragma unroll
      for (int q = 0; q < 8; q++)
        if (valid && half + q < nc) m.smCodes[(half + q) * kSmallMax + j] = (uint8_t)raw[q];
    }
    uint32_t prevS = __shfl_up_sync(0xffffffffu, sidx, 1), nextS = __shfl_down_sync(0xffffffffu, sidx, 1);
    if (lane == 0) prevS = (j > 0u && !c.strat) ? 0u : edgeS;
    if (lane == 31) nextS = (j + 1u < len && !c.strat) ? 0u : edgeS;
    if (j + 1u >= len) nextS = 0xFFFFFFFFu;
    if (valid) {
      m.smEnt[j] = ent;
      SmallRec r; r.mrt = mrt; r.mult = mult; r.es = (e ? 1u : 0u) | (sidx << 8);
      m.smRec[j] = r;
      if (prevS != sidx) m.segpos[sidx] = j;
      if (nextS != sidx) m.cN[sidx] = j + 1u;
      if (e) atomicOr(&s->evMask[sidx >> 5], 1u << (sidx & 31u));
    }
  }
  __syncthreads();
  RF_PROF(3);
  if (tid == 0) s->histDirty = 1;                                    // the candidates keep tables in the large-node histograms
  if (kSort && m.hb >= kLocalMax) {                                 // only tables with a covariate of > 256 distinct values get here
    for (int ci = wid; ci < nc; ci += kWarps)
      if (!s->candHist[ci]) local_rank_candidate(s, m.smEnt, m.smCodes + ci * kSmallMax, ci, len);
  }
  for (int ci = wid; ci < nc; ci += kWarps) {                               // the same warp summarises ci next (score_node)
    if (kFactors && s->candFac[ci]) small_factor_candidate(c, ci); else small_candidate(c, ci, len);
  }
}

// ---- score all drawn candidates of the current node (steps 2-4) ------------------------------
__device__ inline void score_node(const DevData& d, const GrowParams& g, const GrowWork& w, GrowSmem& m, int tree) {
  GrowShared* s = m.s;
  const int tid = vtid();
  const int nc = s->nc;
  const NodeEnt cur = s->cur;
  const uint32_t* list = w.lists + ((size_t)tree * 2 + cur.buf) * w.listStride + cur.start;
  const bool bayes = rule_is_bayes(g.rule), unit = rule_unit_time(g.rule), strat = rule_is_stratified(g.rule);
  double beta = 0.0;
  if (bayes) beta = g.alpha / (((double)cur.m + (double)cur.E) / cur.rtsum);
  const int Kloop = strat ? d.K : 1;
  NodeCtx c{d, g, w, m, list, tree, bayes, unit, strat, beta, Kloop};

  bool allHist = true;
  for (int ci = 0; ci < nc; ci++) if (!RF_CANDHIST(ci)) allHist = false;
  const bool small = (allHist && cur.len <= (uint32_t)kSmallMax) ||
                     (kSort && cur.len <= (uint32_t)kLocalMax && m.hb >= kLocalMax && !(kSink && w.sinkDelta) && g.nsplit == 0);
  bool anyHist = false, anySort = false;
  for (int ci = 0; ci < nc; ci++) { if (RF_CANDHIST(ci)) anyHist = true; else anySort = true; }
  if (small) {
    if (tid == 0) s->isSmall = 1;
    score_small(c);
    anySort = false;                                                 // wide covariates were re-ranked locally (local_rank_candidate)
  } else {
  // segment boundaries of the strata inside this node's (row-sorted) entry list
  if (strat) {
    for (int k = tid; k <= d.K + 1; k += kThreads) {
      uint32_t key = d.sb[k];
      uint32_t lo = 0, hi = cur.len;
      while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if ((list[mid] & kRowMask) < key) lo = mid + 1; else hi = mid;
      }
      m.segpos[k] = lo;
    }
  } else if (tid == 0) {
    m.segpos[0] = 0; m.segpos[1] = 0; m.segpos[2] = cur.len;
  }
  if (s->histDirty) {                                                // small nodes kept tables in the histograms since the last large node
    for (int i = tid; i < g.ncmax * m.hb * m.hs; i += kThreads) { m.hN[i] = 0; m.hE[i] = 0; m.hNx[i] = 0; m.hEx[i] = 0; m.hT[i] = 0.0; m.hW[i] = 0.0; }   // (the term tables lie over all six)
  }
  for (int i = tid; i < nc * m.hb; i += kThreads) { m.sL[i] = 0.0; m.sR[i] = 0.0; }
  for (int i = tid; i < nc * 8; i += kThreads) m.pres[i] = 0u;
  if (tid < nc) m.statP[tid] = 0.0;
  if (tid == 0) { s->have = 0; s->bestDelta = 0.0; s->bestCand = -1; s->bestCode = 0; s->cutsSeen = 0; }
  __syncthreads();
  if (tid == 0) s->histDirty = 0;                                    // (every thread has read the flag above)
  RF_PROF(2);

  if (anyHist) {
    unsigned hmask = 0u;
    for (int ci = 0; ci < nc; ci++) if (RF_CANDHIST(ci)) hmask |= 1u << ci;
    const uint8_t* __restrict__ hptr[kPassCand];
#pragma unroll
    for (int ci = 0; ci < kPassCand; ci++) hptr[ci] = (ci < nc) ? (const uint8_t*)s->candPtr[ci] : nullptr;
    // (candPtr / candStride already point a sparse node at the row-major mirror of the table: see the draw step)
    const size_t cstride = (size_t)s->candStride;
    const int hstride = m.hs * m.hb;
    for (int k0 = 1; k0 <= Kloop; k0 += m.hs) {
      // one pass accumulates the next hs strata (a contiguous entry range: the list is stratum-major),
      // slot = stratum - k0, so the dependent gathers of a mid-size node are exposed once per hs strata
      const int ns = min(m.hs, Kloop - k0 + 1);
      const uint32_t s0 = m.segpos[k0], s1 = m.segpos[k0 + ns];
      if (s0 == s1) continue;
      uint32_t bnd[3];
#pragma unroll
      for (int q = 0; q < 3; q++) bnd[q] = (q + 1 < ns) ? m.segpos[k0 + q + 1] : 0xFFFFFFFFu;
      uint32_t entNext = (s0 + tid < s1) ? list[s0 + tid] : 0u;
      for (uint32_t i = s0 + tid; i < s1; i += kThreads) {
        const uint32_t ent = entNext;
        RF_ASSERT((ent & kRowMask) < (uint32_t)d.n, "large pass: tree %d entry %u = %08x (len %u start %u buf %u)\n", tree, i, ent, cur.len, cur.start, cur.buf);
        if (i + kThreads < s1) entNext = list[i + kThreads];       // software pipelining: next entry's load overlaps this entry's gathers
        const uint32_t row = ent & kRowMask, mult = (ent >> kRowBits) + 1u;
        // issue every gather of this entry before the first shared-memory atomic
        const RowRec rr = d.rowrec[row];
        const uint32_t e = rr.es & 1u; const double tv = unit ? 0.0 : rr.rt;
        const size_t roff = (size_t)row * cstride;
        const int slotOff = ((i >= bnd[0]) + (i >= bnd[1]) + (i >= bnd[2])) * m.hb;
        const bool isT0 = tv == d.rt_mode;
        const unsigned long long ft = (unsigned long long)__double2ll_rn(tv * g.tscale) * (unsigned long long)mult;
#pragma unroll
        for (int half = 0; half < kPassCand; half += 8) {
          if (half < nc) {
            uint32_t raw[8];                                       // histogram candidates are byte-coded
#pragma unroll
            for (int q = 0; q < 8; q++) {
              const int ci = half + q;
              raw[q] = (ci < nc && ((hmask >> ci) & 1u)) ? (uint32_t)hptr[ci][roff] : 0u;
            }
#pragma unroll
            for (int q = 0; q < 8; q++) {
              const int ci = half + q;
              if (ci < nc && ((hmask >> ci) & 1u)) hist_add(m, ci * hstride + slotOff + (int)raw[q], mult, e, unit, isT0, ft);
            }
          }
        }
      }
      __syncthreads();
      RF_PROF(3);
      flush_stratum(c, nc, ns);
      __syncthreads();
      RF_PROF(4);
    }
  }
  if (kSort && anySort) { stratum_totals(c, cur.len); RF_PROF(5); }
  }
  // shared tail: warp ci summarises candidate ci (the warp that scored it), then warp 0 replays the
  // tracker in draw order
  for (int ci = warp_id(); ci < nc; ci += kWarps)
    if (RF_CANDHIST(ci)) track_summary(c, ci);
  __syncthreads();
  RF_PROF(4);

  if (!kSort || !anySort) {
    if (warp_id() == 0) track_combine_all(c, nc);
    __syncthreads();
    RF_PROF(6);
    if (!s->needSlow) return;
  }
#pragma unroll 1
  for (int ci = 0; ci < nc; ci++) {
    if (RF_CANDHIST(ci)) {
      if (warp_id() == 0) track_combine_one(c, ci);
      __syncthreads();
      RF_PROF(6);
    } else {
      sort_candidate(c, ci, cur.len);
      RF_PROF(7);
    }
  }
  __syncthreads();
}

// ---- thread 0: draw the node's candidate covariates (step 1) ---------------------------------
// ---- warp 0: draw the node's candidate covariates (step 1) -----------------------------------
// min(mtry, pool) draws without replacement, slot = ceil(ran1B * size), swap-remove
// (rf.c:8378-8381, rf.c:8305-8307).  The Park-Miller states of all draws are produced at once by
// jump-ahead (lane k: k+1 steps), only the Bays-Durham table walk and the swap-remove are serial.
__device__ inline void draw_candidates(const DevData& d, const GrowParams& g, GrowShared* s, GrowSmem& m) {
  const int lane = lane_id();
  const int size0 = s->poolSize;                                                  // built just before, ascending (rf.c:8209-8213)
  const int nd = min(g.mtry, size0);
  if (nd > 0) {
    const int idum0 = s->rng.idum;
    const int myIdum = (lane < nd) ? ran1_mulmod(kJump[lane], (uint32_t)idum0) : 0;
    int iy = s->rng.iy, myIy = 0;
    int ivReg = s->rng.iv[lane];                                                  // the Bays-Durham table, one entry per lane
#pragma unroll 1
    for (int k = 0; k < nd; k++) {
      const int idk = __shfl_sync(0xffffffffu, myIdum, k);
      const int j = iy / 67108864;
      const int niy = __shfl_sync(0xffffffffu, ivReg, j);
      if (lane == j) ivReg = idk;
      iy = niy;
      if (lane == k) myIy = niy;
    }
    s->rng.iv[lane] = ivReg;
    if (lane == 0) s->rng.iy = iy;
    const int lastIdum = __shfl_sync(0xffffffffu, myIdum, nd - 1);
    if (lane == 0) s->rng.idum = lastIdum;
    uint32_t slot = 0;
    if (lane < nd) slot = (uint32_t)ceil((double)ran1_float(myIy) * ((double)(size0 - lane) * 1.0));   // in [1, size]
#pragma unroll 1
    for (int k = 0; k < nd; k++) {
      const uint32_t sl = __shfl_sync(0xffffffffu, slot, k);
      const int cov = m.pool[sl - 1];
      const int tail = m.pool[size0 - 1 - k];
      __syncwarp();
      if (lane == 0) { m.pool[sl - 1] = (uint16_t)tail; s->candCov[k] = cov; }            // swap-remove (rf.c:8306)
      __syncwarp();
    }
  }
  if (lane == 0) s->nc = nd;
}

// ---- nsplit > 0: the whole CTA draws the node's candidates one at a time ------------------------
// stackAndConstructSplitVector (rf.c:14533-14564) draws a covariate's cuts from chain B right after the
// covariate itself, and how many it draws depends on the number of distinct values the covariate has in
// THIS node (all D - 1 positions when D <= nsplit, else nsplit of them without replacement).  So the
// chain position of candidate k + 1 is only known once candidate k's values have been looked at: per
// candidate, one pass of the CTA over the node's entries collects which codes occur, then thread 0 draws
// the cut ranks (swap-remove over 1..D-1, slot = ceil(ran1B * size)) and turns them into a code mask the
// tracker uses in place of "every present code but the last".
__device__ inline void node_prologue_nsplit(const DevData& d, const GrowParams& g, const GrowWork& w, int tree, GrowShared* s, GrowSmem& m,
                                            const uint32_t* list, uint32_t len, uint32_t mInBag) {
  const int tid = vtid(), lane = lane_id();
  if (warp_id() == 0) {
    if (lane == 0) { s->nc = 0; s->have = 0; s->isSmall = 0; s->bestDelta = 0.0; s->bestCand = -1; s->bestCode = 0; s->cutsSeen = 0; }
    if (lane < 8) s->evMask[lane] = 0u;
    int size = 0;
#pragma unroll 1
    for (int base = 0; base < d.p; base += 32) {
      const int k = base + lane;
      const bool ok = k < d.p && ((s->perm[base >> 5] >> lane) & 1u);
      const unsigned bm = __ballot_sync(0xffffffffu, ok);
      if (ok) m.pool[size + __popc(bm & ((1u << lane) - 1u))] = (uint16_t)k;
      size += __popc(bm);
    }
    if (lane == 0) s->poolSize = size;
  }
  __syncthreads();
  const int size0 = s->poolSize;
  const int nd = min(g.mtry, size0);
  // Thread 0 draws covariates until one needs a look at the node before the chain can go on: a candidate whose cuts are
  // DRAWN (nsplit > 0) or a factor (random pairs when the node is small).  Plain numeric candidates (nsplit = 0) consume
  // no further draws and are scored exactly as in the all-at-once prologue, so a run of them costs one barrier.
  int kDone = 0;
#pragma unroll 1
  while (kDone < nd) {
    if (tid == 0) {
      int kk = kDone;
      while (kk < nd) {                                              // the covariate (rf.c:8378-8381, rf.c:8305-8307)
        const uint32_t sl = ran1_index(s->rng, (uint32_t)(size0 - kk));
        const int cov = m.pool[sl - 1];
        m.pool[sl - 1] = m.pool[size0 - 1 - kk];
        const int Dv = d.D[cov];
        s->candCov[kk] = cov; s->candD[kk] = Dv; s->candW[kk] = d.width[cov]; s->candPtr[kk] = d.codes[cov]; s->candHist[kk] = Dv <= kHistBins;
        const bool plain = g.nsplit == 0 && !(kFactors && d.ftype && d.ftype[cov]);
        s->candMask[kk] = plain ? 0 : 1; s->candFac[kk] = 0; s->facN[kk] = 0;
        kk++;
        if (!plain) break;
      }
      s->candStride = 1;
      s->seqNext = kk;
    }
    __syncthreads();
    const int k = s->seqNext - 1;                                    // the last candidate drawn
    kDone = k + 1;
    if (!s->candMask[k]) continue;                                   // (only possible at the end of the node's draws)
    if (kSort && !s->candHist[k]) {
      // nsplit > 0 on a covariate with more than 256 distinct values: how many cuts are drawn depends on the number of distinct
      // values IN THIS NODE (rf.c:14544-14564), so the node's codes are sorted here to count them; the drawn cut ORDINALS
      // (rank of the distinct value a cut follows, ascending) are kept for sort_candidate, which sorts again when it scores
      uint32_t *ks, *vs;
      sort_node_codes(w, tree, s, m, list, len, k, ks, vs);
      int cnt = 0;
      for (uint32_t i = (uint32_t)tid; i < len; i += kThreads) cnt += (i == 0u || ks[i] != ks[i - 1u]) ? 1 : 0;
      const int np = block_sum<int>(cnt, s->scrI);
      if (tid == 0) {
        int nsel = 0;
        if (np >= 2 && np > g.nsplit) {
          // swap-remove over the ranks 1 .. np-1 (sworVector, rf.c:14551-14561) without materialising the vector: only the
          // slots a draw has overwritten are remembered (at most nsplit of them; the accumulators are idle between nodes)
          uint32_t* modIdx = (uint32_t*)m.sL; uint32_t* modVal = modIdx + 256;
          uint32_t* out = m.facMask + (size_t)k * m.facCuts;
          int nm = 0; uint32_t sz = (uint32_t)np - 1u;
          for (int j = 0; j < g.nsplit; j++) {
            const uint32_t idx = ran1_index(s->rng, sz);
            uint32_t pick = idx, tail = sz; int at = -1;
            for (int t = 0; t < nm; t++) { if (modIdx[t] == idx) { pick = modVal[t]; at = t; } if (modIdx[t] == sz) tail = modVal[t]; }
            if (at >= 0) modVal[at] = tail; else { modIdx[nm] = idx; modVal[nm] = tail; nm++; }
            sz--;
            int q = nsel++;                                            // insertion into the ascending list (sort, rf.c:14563)
            while (q > 0 && out[q - 1] > pick) { out[q] = out[q - 1]; q--; }
            out[q] = pick;
          }
        }
        s->facN[k] = nsel;
      }
      __syncthreads();
      continue;
    }
    if (tid < 8) { m.pres[k * 8 + tid] = 0u; s->cutMask[k][tid] = 0u; s->rankSel[tid] = 0u; }
    __syncthreads();
    {                                                                // which codes occur in the node
      const uint8_t* __restrict__ ptr = (const uint8_t*)s->candPtr[k];
      uint32_t acc0 = 0u, acc1 = 0u;
      for (uint32_t i = (uint32_t)tid; i < len; i += kThreads) {
        const uint32_t code = ptr[list[i] & kRowMask];
        if (code < 32u) acc0 |= 1u << code;
        else if (code < 64u) acc1 |= 1u << (code - 32u);
        else atomicOr(&m.pres[k * 8 + (int)(code >> 5)], 1u << (code & 31u));
      }
      acc0 = __reduce_or_sync(0xffffffffu, acc0); acc1 = __reduce_or_sync(0xffffffffu, acc1);
      if (lane == 0) { if (acc0) atomicOr(&m.pres[k * 8], acc0); if (acc1) atomicOr(&m.pres[k * 8 + 1], acc1); }
    }
    __syncthreads();
    const int covk = s->candCov[k];
    if (kFactors && d.ftype && d.ftype[covk]) {
      // unordered factor (stackAndConstructSplitVector, rf.c:14452-14518): deterministic when the 2^(r-1) - 1 complementary
      // pairs of the r levels present are fewer than the node's in-bag members (nsplit > 0: no more than min(nsplit,
      // members)), every thread then unranks its own cuts; otherwise as many RANDOM pairs as members (nsplit > 0:
      // min(nsplit, members)), drawn from chain B by thread 0: a cardinality in proportion to the group sizes
      // (getRandomPair rf.c:15232-15262), then that many levels without replacement (createRandomBinaryPair rf.c:15263-15302)
      const uint32_t presW = m.pres[k * 8];                          // (a factor has at most kMaxFactorLevels codes)
      const int r = __popc(presW);
      int ncuts = 0; bool det = true;
      if (r >= 2) {
        const uint32_t pairs = (1u << (r - 1)) - 1u;
        if (g.nsplit == 0) { det = !(pairs >= mInBag); ncuts = det ? (int)pairs : (int)mInBag; }
        else { const uint32_t lim = min((uint32_t)g.nsplit, mInBag); det = pairs <= lim; ncuts = det ? (int)pairs : (int)lim; }
      }
      uint32_t* masks = m.facMask + (size_t)k * m.facCuts;
      if (det) {
        for (int j = tid; j < ncuts; j += kThreads) masks[j] = factor_cut_mask(r, j, presW);
      } else if (tid == 0) {
        double cdf[kMaxFactorLevels / 2 + 2];
        const int G = r >> 1;
        cdf[0] = 0.0;
        for (int q = 1; q <= G; q++) { int gs = kChoose[r][q]; if (!(r & 1) && q == G) gs >>= 1; cdf[q] = cdf[q - 1] + (double)gs; }
        for (int j = 0; j < ncuts; j++) {
          const double rv = ceil(ran1_next(s->rng) * cdf[G]);         // float * double, as in the reference
          int gi = 1; while (rv > cdf[gi]) gi++;
          uint8_t lv[kMaxFactorLevels + 1];
          for (int q = 1; q <= r; q++) lv[q] = (uint8_t)q;
          int size = r; uint32_t mk = 0u;
          for (int q = 1; q <= gi; q++) {
            const uint32_t slot = ran1_index(s->rng, (uint32_t)size);
            const int rel = lv[slot];
            lv[slot] = lv[size]; size--;
            mk |= 1u << nth_present_code(presW, rel - 1);
          }
          masks[j] = mk;
        }
      }
      if (tid < 8) {                                                 // the tracker's cut mask: cuts 0 .. ncuts-1
        const int lo = tid * 32;
        s->cutMask[k][tid] = ncuts > lo ? ((ncuts - lo >= 32) ? 0xFFFFFFFFu : ((1u << (ncuts - lo)) - 1u)) : 0u;
      }
      if (tid == 0) { s->candFac[k] = 1; s->facN[k] = ncuts; }
      __syncthreads();
      continue;
    }
    if (tid == 0) {
      int np = 0, last = -1;
      for (int wq = 0; wq < 8; wq++) { const uint32_t wd = m.pres[k * 8 + wq]; np += __popc(wd); if (wd) last = (wq << 5) + 31 - __clz(wd); }
      if (np >= 2) {
        if (g.nsplit > 0 && np > g.nsplit) {
          int sz = np - 1;
          for (int j = 0; j < sz; j++) s->swor[j] = (uint8_t)j;      // ranks 0 .. D-2 of the distinct values (the last is never a cut)
          for (int j = 0; j < g.nsplit; j++) {
            const uint32_t idx = ran1_index(s->rng, (uint32_t)sz);
            const int r = s->swor[idx - 1];
            s->rankSel[r >> 5] |= 1u << (r & 31);
            s->swor[idx - 1] = s->swor[sz - 1]; sz--;
          }
          int rank = 0;
          for (int wq = 0; wq < 8; wq++) {
            uint32_t wd = m.pres[k * 8 + wq], out = 0u;
            while (wd) {
              const int b = __ffs(wd) - 1; wd &= wd - 1u;
              if ((s->rankSel[rank >> 5] >> (rank & 31)) & 1u) out |= 1u << b;
              rank++;
            }
            s->cutMask[k][wq] = out;
          }
        } else {
          for (int wq = 0; wq < 8; wq++) s->cutMask[k][wq] = m.pres[k * 8 + wq];
          s->cutMask[k][last >> 5] &= ~(1u << (last & 31));
        }
      }
    }
    __syncthreads();
  }
  if (tid == 0) { s->nc = nd; s->algoBytes += (unsigned long long)mInBag * (12ull * (unsigned long long)nd + 20ull); }
}

__device__ inline bool append_record(const GrowWork& w, GrowShared* s, int tree, uint32_t nodeID, uint32_t parm,
                                     uint32_t a, uint32_t b, double stat, uint32_t parentRec, uint32_t* outIdx) {
  uint32_t q = s->nodeCount;
  if ((q % kRecChunk) == 0) {
    uint32_t ch = q / kRecChunk;
    uint32_t id = atomicAdd(w.poolCursor, 1u);
    if (id >= w.poolChunks || ch >= w.maxChunksPerTree) { atomicExch(w.status, 1); s->abort = 1; return false; }
    w.ctab[(size_t)tree * w.maxChunksPerTree + ch] = id;
  }
  uint32_t cidx = w.ctab[(size_t)tree * w.maxChunksPerTree + q / kRecChunk];
  size_t at = (size_t)cidx * kRecChunk + (q % kRecChunk);
  w.pool[at] = Rec{nodeID, parm, a, b};
  w.statPool[at] = stat;
  if (parentRec != kNone) rec_ptr(w, tree, parentRec)->b = q;
  s->nodeCount = q + 1;
  *outIdx = q;
  return true;
}


// A node the DFS will not try to split (node gate rf.c:15354-15371 + purity rf.c:6391-6416, see the kernel's loop): its
// record is all there is to it, so it is written as soon as the node is met -- by the split that creates it or by the
// pop that reaches it -- and never becomes an iteration of the CTA's node loop (half of all nodes are such leaves).
__device__ __forceinline__ bool node_is_leaf(const GrowParams& g, const NodeEnt& e) {
  const bool attempt = (e.m >= 2u * (uint32_t)g.nodesize) && (g.nodedepth < 0 || e.depth < (uint32_t)g.nodedepth);
  return !(attempt && e.E > 0u && e.E < e.m);
}

// lane 0 of warp 0: pop pending right children until one needs an attempt (it becomes s->cur; returns true: the caller
// restores its permissible mask from stackPerm[s->sp]) or the stack is empty (s->abort = 2, tree finished)
__device__ inline bool pop_next(const GrowParams& g, const GrowWork& w, GrowShared* s, int tree) {
  while (true) {
    if (s->sp == 0) { s->abort = 2; return false; }
    s->sp--;
    const NodeEnt e = w.stack[(size_t)tree * kStackCap + s->sp];
    if (kSub && s->inSub && s->sp < s->subSp) s->inSub = 0;         // the DFS is back above the staged subtree's root
    if (node_is_leaf(g, e)) {
      uint32_t q;
      append_record(w, s, tree, e.nodeID, 0u, e.E, e.m, nan(""), e.parentRec, &q);
      if (s->abort) return false;
      continue;
    }
    s->cur = e;
    return true;
  }
}

#ifndef RF_MIN_CTAS
#define RF_MIN_CTAS 2
#endif
#ifndef RF_KERNEL_NAME
#define RF_KERNEL_NAME rfslam_grow_kernel      // the 256-thread build; grow_variants.cu instantiates wider CTAs under other names
#endif
extern "C" __global__ void __launch_bounds__(kThreads, RF_MIN_CTAS)
RF_KERNEL_NAME(DevData d, GrowParams g, GrowWork w) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ GrowShared g_state;
  GrowSmem m = carve(smem_raw, g.ncmax, d.K, g.hbins, g.hslots, g.tabSlots, g.subMax, (d.rowPitch + 15) & ~15, g.facCuts);
  m.s = &g_state;
  GrowShared* s = &g_state;
#ifdef RF_CHECK
  { unsigned dyn; asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    RF_ASSERT((size_t)((unsigned char*)(m.stIdx + m.subMax) - smem_raw) <= (size_t)dyn, "shared memory: carve-up ends at %llu of %u bytes\n", (unsigned long long)((unsigned char*)(m.stIdx + m.subMax) - smem_raw), dyn); }
#endif
  const int tid = vtid();
  const int n = d.n;
  const bool sliced = w.treeState != nullptr;

  // ---- clean histograms (flushes keep them clean afterwards) ----
  for (int i = tid; i < g.ncmax * m.hb * m.hs; i += kThreads) { m.hN[i] = 0; m.hE[i] = 0; m.hNx[i] = 0; m.hEx[i] = 0; m.hT[i] = 0.0; m.hW[i] = 0.0; }

  // Time slicing.  A batch with more trees than resident CTAs (500 trees, 296 slots) used to run as a
  // full wave and a partial one: 1.61 s against 1.43 s for a balanced load, and a tree cannot be split
  // (chain B).  But a tree's whole state between two nodes is GrowShared (2 KB: RNG, current node,
  // permissible mask, counters) plus what already lives in global memory, so the grid is the resident
  // CTAs and every CTA works on a tree for a quantum, parks it and takes the next waiting one: all trees
  // advance at the same rate and finish together.  Results do not depend on the schedule.
  while (true) {
  if (tid == 0) {
    int t = -1, fresh = 0;
    if (!sliced) { t = (int)blockIdx.x; fresh = 1; }
    else {
      while (atomicAdd(w.waiting, 0) > 0) {
        const unsigned k = atomicAdd(w.ticket, 1u) % (unsigned)w.ntrees;
        const int old = atomicCAS(&w.treeState[k], 0, 2);
        if (old == 0) { t = (int)k; fresh = 1; atomicSub(w.waiting, 1); break; }
        if (old == 1 && atomicCAS(&w.treeState[k], 1, 2) == 1) { t = (int)k; atomicSub(w.waiting, 1); break; }
      }
    }
    s->curTree = t; s->fresh = fresh;
  }
  __syncthreads();
  const int tree = s->curTree;
  const bool isFresh = s->fresh != 0;
  if (tree < 0) break;
  __syncthreads();                                                   // every thread holds the pick in registers: the restore below overwrites GrowShared (curTree, fresh included)
  __threadfence();                                                   // acquire: what the previous owner of this tree wrote
  uint32_t* listA = w.lists + (size_t)tree * 2 * w.listStride;
  const uint16_t* cnt = w.cnt + (size_t)tree * w.cntStride;
  if (isFresh) {
  if (tid == 0) {
    for (int k = 0; k < kProfSlots; k++) s->prof[k] = 0;
    s->tmark = clock64();
    ran1_seed(s->rng, w.seedB[tree]);
    ran1_init(s->rng);
    s->nodeCount = 0; s->leafCount = 1; s->sp = 0; s->candCount = 0; s->algoBytes = 0; s->abort = 0; s->inSub = 0;
    s->histDirty = sliced ? 1 : 0;                                  // another tree's small nodes may have used the histograms
  }
  for (int k = tid; k < 32; k += kThreads) {
    uint32_t msk = 0;
    for (int b = 0; b < 32; b++) if (k * 32 + b < d.p) msk |= (1u << b);
    s->perm[k] = msk;
  }
  __syncthreads();

  // ---- root: in-bag entry list in internal row order from the bootstrap counts (rf.c:492-627) ----
  // Eight consecutive rows per thread and tile: one 16-byte load of their u16 counts, one block scan per 8 * kThreads
  // rows (the counts are streamed once; a tile per kThreads rows was one dependent load + two barriers per row).
  {
    constexpr int kRpt = 8;
    uint32_t base = 0;
    unsigned long long ne = 0; double rts = 0.0;
    const uint32_t ntiles = ((uint32_t)n + kRpt * kThreads - 1) / (kRpt * kThreads);
    const uint4* cnt8 = (const uint4*)cnt;                                   // cntStride is a multiple of 8 counts
    for (uint32_t t = 0; t < ntiles; t++) {
      const uint32_t i0 = (t * kThreads + tid) * kRpt;
      uint4 raw = make_uint4(0u, 0u, 0u, 0u);
      if (i0 < (uint32_t)n) raw = cnt8[i0 / kRpt];                           // rows past n inside the last vector hold zero counts
      uint32_t cv[kRpt] = {raw.x & 0xFFFFu, raw.x >> 16, raw.y & 0xFFFFu, raw.y >> 16, raw.z & 0xFFFFu, raw.z >> 16, raw.w & 0xFFFFu, raw.w >> 16};
      uint32_t mine = 0;
#pragma unroll
      for (int k = 0; k < kRpt; k++) { if (i0 + k >= (uint32_t)n) cv[k] = 0u; mine += (cv[k] + kMaxMult - 1) / kMaxMult; }
      unsigned long long tot;
      unsigned long long inc = block_incl_scan<unsigned long long>((unsigned long long)mine, s->scrU, tot);
      uint32_t pos = base + (uint32_t)(inc - mine);
#pragma unroll
      for (int k = 0; k < kRpt; k++) {
        uint32_t left = cv[k];
        if (left) {
          const RowRec rr = d.rowrec[i0 + k];
          ne += ((unsigned long long)left << 32) | (unsigned long long)((rr.es & 1u) ? left : 0u);
          rts += (double)left * rr.rt;
          while (left) {
            const uint32_t mu = left > kMaxMult ? kMaxMult : left;
            listA[pos++] = ((mu - 1u) << kRowBits) | (i0 + k);
            left -= mu;
          }
        }
      }
      base += (uint32_t)tot;
      __syncthreads();
    }
    unsigned long long tne = block_sum<unsigned long long>(ne, s->scrU);
    double trt = block_sum<double>(rts, s->scrD);
    if (tid == 0) {
      s->cur.start = 0; s->cur.len = base; s->cur.m = (uint32_t)(tne >> 32); s->cur.E = (uint32_t)(tne & 0xffffffffu);
      s->cur.rtsum = trt; s->cur.nodeID = 1; s->cur.parentRec = kNone; s->cur.depth = 0; s->cur.buf = 0;
    }
    __syncthreads();
  }
  RF_PROF(0);
  } else {
    // resume a parked tree
    const uint32_t* src32 = (const uint32_t*)(w.saved + (size_t)tree * sizeof(GrowShared));
    uint32_t* dst32 = (uint32_t*)s;
    for (int k = tid; k < (int)(sizeof(GrowShared) / 4); k += kThreads) dst32[k] = __ldcg(src32 + k);
    __syncthreads();
    if (tid == 0) { s->curTree = tree; s->fresh = 0; s->abort = 0; s->histDirty = 1; s->tmark = clock64(); s->prof[13] = 0; }
    __syncthreads();
  }
  if (tid == 0) s->sliceStart = clock64();

  const bool bayes = rule_is_bayes(g.rule);
  // ---- depth-first growth (rf.c:21460-21792) ----
  while (true) {
    // park -- never inside a staged subtree (its state lives in this CTA's shared memory; it is over in well under a quantum)
    if (sliced && tid == 0 && !s->abort && !(kSub && s->inSub) && clock64() - s->sliceStart > w.quantum && atomicAdd(w.waiting, 0) > 0) s->abort = 3;
    __syncthreads();
    if (s->abort) break;
    const NodeEnt cur = s->cur;
    RF_ASSERT((size_t)cur.start + cur.len <= w.listStride && cur.len >= 1u && cur.buf <= 1u && s->sp <= (uint32_t)kStackCap, "node state: tree %d start %u len %u buf %u sp %u nodeID %u\n", tree, cur.start, cur.len, cur.buf, s->sp, cur.nodeID);
    if (kProf && w.prof && tid == 0) {
      long long now = clock64();
      if (s->prof[13]) s->prof[16 + s->prof[12]] += (unsigned long long)now - s->prof[13];
      int cls = 31 - __clz(cur.len | 1u);
      s->prof[48 + (cls >> 1)] += 1ull;
      s->prof[12] = (unsigned long long)cls;
      s->prof[13] = (unsigned long long)now;
    }
    // node gate (rf.c:15354-15371) and purity (rf.c:6391-6416): both classes present
    bool attempt = (cur.m >= 2u * (uint32_t)g.nodesize) && (g.nodedepth < 0 || cur.depth < (uint32_t)g.nodedepth);
    // purity (rf.c:6391-6416): the variance of the response codes is (E/m)(1 - E/m) >= (1/m)(1 - 1/m),
    // i.e. > 1.4e-8 for any m < 2^26 whenever both classes are present, and exactly 0 otherwise --
    // so "variance > EPSILON" is the integer test 0 < E < m
    if (attempt) attempt = cur.E > 0u && cur.E < cur.m;
    if (kSink && w.sinkDelta) attempt = true;              // single-node scoring mode: no gate
    // One serial section per node, run by warp 0 between two block barriers: tracker reset, the draws
    // and the candidate metadata (a barrier + phase boundary costs ~1k cycles on a node that has a few
    // dozen rows; there used to be five of them before the first gather).
    if (kSeq && g.seqDraw && attempt) {
      node_prologue_nsplit(d, g, w, tree, s, m, w.lists + ((size_t)tree * 2 + cur.buf) * w.listStride + cur.start, cur.len, cur.m);
    } else if (warp_id() == 0) {
      const int lane = lane_id();
      if (lane == 0) { s->nc = 0; s->have = 0; s->isSmall = 0; s->bestDelta = 0.0; s->bestCand = -1; s->bestCode = 0; s->cutsSeen = 0; }
      if (lane < 8) s->evMask[lane] = 0u;
      __syncwarp();
      if (attempt) {                                       // permissible covariates, ascending, by ballot compaction
        int size = 0;
#pragma unroll 1
        for (int base = 0; base < d.p; base += 32) {
          const int k = base + lane;
          const bool ok = k < d.p && ((s->perm[base >> 5] >> lane) & 1u);
          const unsigned bm = __ballot_sync(0xffffffffu, ok);
          if (ok) m.pool[size + __popc(bm & ((1u << lane) - 1u))] = (uint16_t)k;
          size += __popc(bm);
        }
        if (lane == 0) s->poolSize = size;
        __syncwarp();
        draw_candidates(d, g, s, m);
        __syncwarp();
        const int ncd = s->nc;
        if (lane == 0) s->algoBytes += (unsigned long long)cur.m * (12ull * (unsigned long long)ncd + 20ull);
        // A node that holds few of the table's rows gathers from the row-major mirror (when the table has one): the c
        // code bytes of a row sit in at most rowPitch / 32 sectors there, against one 32-byte sector per column from
        // the column-major arrays (whose sectors are shared between entries only while the node still holds a good
        // fraction of consecutive rows).  One stride per node: code = candPtr[k][row * candStride].
        const bool sparse = d.rowcodes != nullptr && (unsigned long long)cur.len * (unsigned long long)kSparseRatio < (unsigned long long)d.n;
        if (lane == 0) s->candStride = sparse ? d.rowPitch : 1;
        if (lane < ncd) {                                  // candidate metadata: one parallel round of loads
          const int cov = s->candCov[lane];
          const int Dv = d.D[cov];
          s->candD[lane] = Dv; s->candW[lane] = d.width[cov]; s->candHist[lane] = Dv <= kHistBins; s->candLocal[lane] = 0; s->candFac[lane] = 0;
          s->candPtr[lane] = sparse ? (const void*)(d.rowcodes + cov) : d.codes[cov];
        }
      }
    }
    __syncthreads();
    RF_PROF(1);
    if (attempt && s->nc > 0) score_node(d, g, w, m, tree);   // ends with a block barrier
    if (kSink && w.sinkDelta) {                            // emit-only mode stops after the root
      if (tid == 0) *w.sinkCount = s->cutsSeen;
      break;
    }
    const bool split = attempt && s->have;
    if (!split) {
      if (warp_id() == 0) {
        int restore = 0;
        if (lane_id() == 0) {
          uint32_t q;
          append_record(w, s, tree, cur.nodeID, 0u, cur.E, cur.m, nan(""), cur.parentRec, &q);
          if (!s->abort) restore = pop_next(g, w, s, tree) ? 1 : 0;
        }
        restore = __shfl_sync(0xffffffffu, restore, 0);
        if (restore) {
          const uint32_t* sp = w.stackPerm + ((size_t)tree * kStackCap + s->sp) * g.pw;
#pragma unroll 1
          for (int k = lane_id(); k < g.pw; k += 32) s->perm[k] = sp[k];
        }
      }
      RF_PROF(10);
      continue;                                             // the loop-top barrier publishes cur / perm / abort
    }

    // ---- split: record, partition, push right, descend left ----
    const int bc = s->bestCand;
    const uint32_t bcode = s->bestCode;
    // an unordered factor routes by level subset (splitOnFactor, rf.c:1607-1616): bestCode is the cut's index, the LEFT set
    // is its mask over codes.  goes_left(code) covers both kinds of split with one shift: a numeric cut at code b is the
    // mask of codes 0..b when the column has at most 32 codes, the compare otherwise.
    const bool isFac = kFactors && s->candFac[bc] != 0;
    const uint32_t fmask = isFac ? m.facMask[(size_t)bc * m.facCuts + bcode] : 0u;
#define RF_GOES_LEFT(code) (isFac ? (((fmask >> (code)) & 1u) != 0u) : ((code) <= bcode))
    const void* ptr = s->candPtr[bc]; const int wd = s->candW[bc]; const int cov = s->candCov[bc];
    const uint32_t* src = w.lists + ((size_t)tree * 2 + cur.buf) * w.listStride + cur.start;
    uint32_t* dst = w.lists + ((size_t)tree * 2 + (cur.buf ^ 1u)) * w.listStride + cur.start;
    const bool small = s->isSmall != 0;
    const bool localCut = kSort && small && s->candLocal[bc] != 0;
    const uint8_t* cptr = m.smCodes + bc * kSmallMax;
    uint32_t nL; unsigned long long tneL; double trtL = 0.0, trtR = 0.0;
    if (small) {
      // the node is staged in shared memory: thread t owns entries 2t and 2t+1, one block scan of a
      // packed (left entries | left rows << 16 | left events << 40) word gives sizes and positions
      constexpr int kEpt = (kSmallMax + kThreads - 1) / kThreads;   // staged entries per thread (consecutive)
      static_assert(kEpt <= 8, "staged entries per thread");
      unsigned long long my = 0ull; bool lf[kEpt]; uint32_t en[kEpt];
      double rtL = 0.0, rtR = 0.0;
      const bool sub = kSub && s->inSub != 0;                    // staged subtree: the "list" is the slot order in shared memory
      const uint32_t soff = sub ? cur.start - s->subStart : 0u;
#pragma unroll
      for (int u = 0; u < kEpt; u++) {
        const uint32_t i = (uint32_t)kEpt * tid + u;
        lf[u] = false; en[u] = 0u;
        if (i < cur.len) {
          en[u] = sub ? (uint32_t)m.stIdx[soff + i] : m.smEnt[i];
          lf[u] = RF_GOES_LEFT((uint32_t)cptr[i]);                // LEFT iff cut - x >= 0 (rf.c:15194)
          if (localCut && cptr[i] == bcode) s->bestWide = load_code(ptr, wd, en[u] & kRowMask);   // same value from every such entry
          const SmallRec r = m.smRec[i];
          const unsigned long long mult = r.mult;
          if (lf[u]) my += 1ull | (mult << 16) | (((r.es & 1u) ? mult : 0ull) << 40);
          if (bayes) { if (lf[u]) rtL += r.mrt; else rtR += r.mrt; }
        }
      }
      unsigned long long tot;
      const unsigned long long inc = block_incl_scan<unsigned long long>(my, s->scrU, tot);
      nL = (uint32_t)(tot & 0xFFFFull);
      tneL = (((tot >> 16) & 0xFFFFFFull) << 32) | (tot >> 40);
      if (bayes) { trtL = block_sum<double>(rtL, s->scrD); trtR = block_sum<double>(rtR, s->scrD); }
      RF_PROF(8);
      uint32_t before = (uint32_t)((inc - my) & 0xFFFFull);
#pragma unroll
      for (int u = 0; u < kEpt; u++) {
        const uint32_t i = (uint32_t)kEpt * tid + u;
        if (i < cur.len) {
          if (sub) {                                              // (every thread read its slots before the scan's barriers)
            if (lf[u]) { m.stIdx[soff + before] = (uint16_t)en[u]; before++; }
            else m.stIdx[soff + nL + (i - before)] = (uint16_t)en[u];
          } else {
            if (lf[u]) { dst[before] = en[u]; before++; }
            else dst[nL + (i - before)] = en[u];
          }
        }
      }
    } else {
    // Four consecutive entries per thread in both passes: the four list loads, then the four
    // dependent code gathers, are in flight together (the kernel is latency-bound, not HBM-bound).
    // pass 1: sizes of the left child
    unsigned long long cL = 0, neL = 0; double rtL = 0.0, rtR = 0.0;
    const size_t pstride = (wd == 1) ? (size_t)s->candStride : 1;   // (the winning candidate's pointer follows the node's layout)
    for (uint32_t base = 0; base < cur.len; base += 4u * kThreads) {
      uint32_t ent[4]; uint32_t cd[4]; uint32_t ev4[4]; double rt4[4];
#pragma unroll
      for (int u = 0; u < 4; u++) { const uint32_t i = base + 4u * tid + u; ent[u] = (i < cur.len) ? src[i] : 0xFFFFFFFFu; }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const bool valid = ent[u] != 0xFFFFFFFFu;
        const uint32_t row = ent[u] & kRowMask;
        cd[u] = valid ? (wd == 1 ? (uint32_t)((const uint8_t*)ptr)[(size_t)row * pstride] : load_code(ptr, wd, row)) : 0xFFFFFFFFu;
        const RowRec rr = valid ? d.rowrec[row] : RowRec{0.0, 0u, 0u};
        ev4[u] = rr.es & 1u; rt4[u] = bayes ? rr.rt : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        if (ent[u] != 0xFFFFFFFFu) {
          const bool left = RF_GOES_LEFT(cd[u]);                   // LEFT iff cut - x >= 0 (rf.c:15194)
          const uint32_t mult = (ent[u] >> kRowBits) + 1u;
          if (left) { cL++; neL += ((unsigned long long)mult << 32) | (unsigned long long)(ev4[u] ? mult : 0u); }
          if (bayes) { const double v = (double)mult * rt4[u]; if (left) rtL += v; else rtR += v; }
        }
      }
    }
    nL = (uint32_t)block_sum<unsigned long long>(cL, s->scrU);
    tneL = block_sum<unsigned long long>(neL, s->scrU);
    if (bayes) { trtL = block_sum<double>(rtL, s->scrD); trtR = block_sum<double>(rtR, s->scrD); }
    RF_PROF(8);
    // pass 2: stable scatter
    {
      uint32_t baseL = 0, baseR = 0;
      for (uint32_t base = 0; base < cur.len; base += 4u * kThreads) {
        uint32_t ent[4]; uint32_t cd[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { const uint32_t i = base + 4u * tid + u; ent[u] = (i < cur.len) ? src[i] : 0xFFFFFFFFu; }
#pragma unroll
        for (int u = 0; u < 4; u++)
          cd[u] = (ent[u] != 0xFFFFFFFFu) ? (wd == 1 ? (uint32_t)((const uint8_t*)ptr)[(size_t)(ent[u] & kRowMask) * pstride] : load_code(ptr, wd, ent[u] & kRowMask)) : 0xFFFFFFFFu;
        uint32_t mine = 0;
#pragma unroll
        for (int u = 0; u < 4; u++) mine += (ent[u] != 0xFFFFFFFFu && RF_GOES_LEFT(cd[u])) ? 1u : 0u;
        unsigned long long tot;
        const unsigned long long inc = block_incl_scan<unsigned long long>((unsigned long long)mine, s->scrU, tot);
        uint32_t before = (uint32_t)inc - mine;                     // left entries of this tile before mine
#pragma unroll
        for (int u = 0; u < 4; u++) {
          if (ent[u] != 0xFFFFFFFFu) {
            if (RF_GOES_LEFT(cd[u])) { dst[baseL + before] = ent[u]; before++; }
            else dst[nL + baseR + (4u * tid + u - before)] = ent[u];
          }
        }
        const uint32_t nvalid = min(4u * kThreads, cur.len - base);
        baseL += (uint32_t)tot;
        baseR += nvalid - (uint32_t)tot;
        __syncthreads();
      }
    }
    }
    RF_PROF(9);
    if (warp_id() == 0) {
      int act = 0;                                                   // 1: right child pushed (save the mask), 2: a pending node popped (restore its mask)
      if (lane_id() == 0) {
        uint32_t q;
        append_record(w, s, tree, cur.nodeID, ((uint32_t)cov + 1u) | (isFac ? kFactorFlag : 0u), isFac ? fmask : (localCut ? s->bestWide : bcode), kNone, s->bestDelta, cur.parentRec, &q);
        s->algoBytes += 16ull * (unsigned long long)cur.m;
        if (!s->abort) {
          s->leafCount++;
          NodeEnt r;
          r.start = cur.start + nL; r.len = cur.len - nL;
          r.m = cur.m - (uint32_t)(tneL >> 32); r.E = cur.E - (uint32_t)(tneL & 0xffffffffu);
          r.rtsum = trtR; r.nodeID = s->leafCount; r.parentRec = q; r.depth = cur.depth + 1; r.buf = cur.buf ^ 1u;
          NodeEnt l;
          l.start = cur.start; l.len = nL; l.m = (uint32_t)(tneL >> 32); l.E = (uint32_t)(tneL & 0xffffffffu);
          l.rtsum = trtL; l.nodeID = cur.nodeID; l.parentRec = kNone; l.depth = cur.depth + 1; l.buf = cur.buf ^ 1u;
          if (!node_is_leaf(g, l)) {                                 // descend left, the right child waits on the stack (a leaf among them is written by the pop)
            if (s->sp >= (uint32_t)kStackCap) { atomicExch(w.status, 2); s->abort = 1; }
            else { w.stack[(size_t)tree * kStackCap + s->sp] = r; s->cur = l; act = 1; }
          } else {                                                   // the left child is a leaf: its record follows its parent's at once (pre-order)
            uint32_t ql;
            append_record(w, s, tree, l.nodeID, 0u, l.E, l.m, nan(""), kNone, &ql);
            if (!s->abort) {
              if (!node_is_leaf(g, r)) s->cur = r;                   // straight on to the right child: same permissible mask, nothing to push
              else {
                append_record(w, s, tree, r.nodeID, 0u, r.E, r.m, nan(""), r.parentRec, &ql);
                if (!s->abort) act = pop_next(g, w, s, tree) ? 2 : 0;
              }
            }
          }
        }
      }
      act = __shfl_sync(0xffffffffu, act, 0);
      if (act == 1) {
        // children inherit the parent's (possibly updated) permissible mask (rf.c:10725-10727)
        uint32_t* sp = w.stackPerm + ((size_t)tree * kStackCap + s->sp) * g.pw;
#pragma unroll 1
        for (int k = lane_id(); k < g.pw; k += 32) sp[k] = s->perm[k];
        __syncwarp();
        if (lane_id() == 0) s->sp++;
      } else if (act == 2) {
        const uint32_t* sp = w.stackPerm + ((size_t)tree * kStackCap + s->sp) * g.pw;
#pragma unroll 1
        for (int k = lane_id(); k < g.pw; k += 32) s->perm[k] = sp[k];
      }
    }
    RF_PROF(11);
  }
  __syncthreads();
  if (s->abort == 3) {
    // park: state to global, then publish (every thread's writes of this slice are fenced first)
    uint32_t* dst32 = (uint32_t*)(w.saved + (size_t)tree * sizeof(GrowShared));
    const uint32_t* src32 = (const uint32_t*)s;
    for (int k = tid; k < (int)(sizeof(GrowShared) / 4); k += kThreads) dst32[k] = src32[k];
    __threadfence();
    __syncthreads();
    if (tid == 0) { atomicExch(&w.treeState[tree], 1); atomicAdd(w.waiting, 1); }
    __syncthreads();
    continue;
  }
  if (tid == 0) {
    if (kProf && w.prof) for (int k = 0; k < kProfSlots; k++) w.prof[(size_t)tree * kProfSlots + k] = s->prof[k];
    w.nodeCount[tree] = s->nodeCount;
    w.leafCount[tree] = s->leafCount;
    w.candCount[tree] = s->candCount;
    w.algoBytes[tree] = s->algoBytes;
    if (sliced) atomicExch(&w.treeState[tree], 3);
  }
  __syncthreads();
  if (!sliced) break;
  }
}

}  // namespace rfslam
