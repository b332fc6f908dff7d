// hostpool.cu -- page-locked, recycled host memory for the arrays the C-ABI hands to the caller.
//
// A 500-tree forest on 1M rows is ~1 GB of records (randomForestSRC.c:9-79 lists them).  Fresh
// malloc'd pages cost a page fault + kernel zeroing per 4 KB and pageable device-to-host copies are
// staged through the driver's bounce buffers; together that was ~0.4 s per rfslam_b200_grow call,
// a fifth of the step.  Blocks come from cudaHostAlloc instead and go back to a free list when the
// caller releases the forest (rfslam_b200_free_forest / rfslam_b200_free_array), so a second call
// of similar size touches no new pages and its copies run at PCIe speed, asynchronously.
#include <sys/mman.h>
#include <unistd.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <unordered_map>
#include "common.cuh"
#include "internal.h"

namespace rfslam {
namespace {
std::mutex g_mu;
std::unordered_map<void*, size_t> g_owned;          // every live pool block (handed out or idle) -> capacity
std::unordered_map<void*, int> g_kind;              // 1 = huge-page aligned malloc + cudaHostRegister, 0 = cudaHostAlloc
std::multimap<size_t, void*> g_idle;                // capacity -> idle block
size_t g_idleBytes = 0;

constexpr size_t kSmall = 1u << 16;                 // below this plain malloc is used (free() releases it)

size_t idle_cap() {
  static size_t cap = [] {
    // idle page-locked blocks kept for reuse: a quarter of the machine's memory, at least 8 GB.  The forest gathered on
    // rank 0 of an 8-GPU job is 8.5 GB per call: with a fixed 8 GB cap every call freed and re-pinned it (seconds per step)
    const char* e = getenv("RFSLAM_HOST_POOL_MB");
    if (e) return (size_t)atoll(e) << 20;
    const long pages = sysconf(_SC_PHYS_PAGES), psz = sysconf(_SC_PAGE_SIZE);
    const size_t ram = (pages > 0 && psz > 0) ? (size_t)pages * (size_t)psz : ((size_t)32 << 30);
    return std::max<size_t>((size_t)8 << 30, ram / 4);
  }();
  return cap;
}

size_t round_cap(size_t bytes) {                    // eighth-of-an-octave size classes, 2 MB granularity
  size_t p2 = 1; while ((p2 << 1) <= bytes) p2 <<= 1;
  size_t step = std::max<size_t>(p2 >> 3, (size_t)2 << 20);
  return (bytes + step - 1) / step * step;
}
void release_block_locked(void* p) {                 // (g_mu held)
  auto it = g_kind.find(p);
  const int kind = it == g_kind.end() ? 0 : it->second;
  if (it != g_kind.end()) g_kind.erase(it);
  if (kind == 1) { cudaHostUnregister(p); free(p); } else cudaFreeHost(p);
}
void release_block(void* p) { std::lock_guard<std::mutex> lk(g_mu); release_block_locked(p); }
}  // namespace

void* host_alloc(size_t bytes) {
  if (bytes < kSmall) return malloc(std::max<size_t>(bytes, 1));
  const size_t cap = round_cap(bytes);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_idle.lower_bound(cap);
    if (it != g_idle.end() && it->first < 2 * cap) {
      void* p = it->second; g_idleBytes -= it->first; g_idle.erase(it);
      return p;
    }
  }
  // A fresh block: 2 MB-aligned memory advised to transparent huge pages, then registered.  Pinning 4 KB pages is what makes
  // cudaHostAlloc slow (measured on the B200 host: 4 GB in 2.14 s; this way 0.89 s; the device-to-host rate is the same
  // 49 GB/s) -- and a configs[2] forest is 20 GB of records coming home into blocks that do not exist yet on the first call.
  void* p = nullptr;
  int kind = 0;
  if (cap >= ((size_t)8 << 20) && posix_memalign(&p, (size_t)2 << 20, cap) == 0 && p) {
    madvise(p, cap, MADV_HUGEPAGE);
    if (cudaHostRegister(p, cap, cudaHostRegisterPortable | cudaHostRegisterMapped) == cudaSuccess) kind = 1;
    else { cudaGetLastError(); free(p); p = nullptr; }
  }
  if (!p) {
    if (cudaHostAlloc(&p, cap, cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();                            // not page-lockable right now: pageable memory still works
      return malloc(bytes);
    }
  }
  std::lock_guard<std::mutex> lk(g_mu);
  g_owned[p] = cap; g_kind[p] = kind;
  return p;
}

void* host_calloc(size_t bytes) {
  void* p = host_alloc(bytes);
  if (p) memset(p, 0, bytes);
  return p;
}

void host_free(void* p) {
  if (!p) return;
  size_t cap = 0;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_owned.find(p);
    if (it == g_owned.end()) { cap = 0; }
    else {
      cap = it->second;
      if (g_idleBytes + cap <= idle_cap()) { g_idle.emplace(cap, p); g_idleBytes += cap; return; }
      g_owned.erase(it);
    }
  }
  if (cap) release_block(p); else free(p);
}

void host_pool_trim() {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& kv : g_idle) { g_owned.erase(kv.second); release_block_locked(kv.second); }
  g_idle.clear(); g_idleBytes = 0;
}

}  // namespace rfslam
