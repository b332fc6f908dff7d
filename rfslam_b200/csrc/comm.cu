// comm.cu -- the one exchange step of a tree-sharded forest, on the device.
//
// Trees shard across GPUs with no traffic during growth (rf.c:7082-7087: trees share nothing).  After growth
//   (1) the ensemble numerators / denominators are summed over ranks (what rf.c:944-949 does under per-row
//       locks between OpenMP threads): one NCCL all-reduce group over the DEVICE accumulators, before they are
//       finalized and sent home -- no host bounce;
//   (2) the shards' forest records are gathered to rank 0 over NVLink (the reference's saveTree under
//       `omp critical`, rf.c:20938-20946): sizes first, then one group of ncclSend / ncclRecv per array (a
//       gather-v), then a device kernel that interleaves the trees back into ascending treeID order
//       (tree b lives on rank (b - 1) mod world at local position (b - 1) / world).
// NCCL is bound at run time (dlopen "libnccl.so.2"): inside a torch process that is the library torch already
// loaded, elsewhere the system one; librfslam_b200.so itself links against nothing but the CUDA runtime.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

#include "internal.h"

namespace rfslam {

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};
NcclApi g_nccl;
std::mutex g_ncclMu;

bool nccl_load() {
  std::lock_guard<std::mutex> lk(g_ncclMu);
  if (g_nccl.handle) return true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!h) { set_error("NCCL not found (dlopen libnccl.so.2: %s)", dlerror()); return false; }
#define RF_SYM(field, name) \
  *(void**)(&g_nccl.field) = dlsym(h, name); \
  if (!g_nccl.field) { set_error("NCCL symbol %s missing", name); dlclose(h); return false; }
  RF_SYM(GetUniqueId, "ncclGetUniqueId") RF_SYM(CommInitRank, "ncclCommInitRank") RF_SYM(CommDestroy, "ncclCommDestroy")
  RF_SYM(AllReduce, "ncclAllReduce") RF_SYM(AllGather, "ncclAllGather") RF_SYM(Send, "ncclSend") RF_SYM(Recv, "ncclRecv")
  RF_SYM(GroupStart, "ncclGroupStart") RF_SYM(GroupEnd, "ncclGroupEnd") RF_SYM(GetErrorString, "ncclGetErrorString")
  RF_SYM(GetVersion, "ncclGetVersion")
#undef RF_SYM
  g_nccl.handle = h;
  return true;
}

#define RF_NCCL(call)                                                                                        \
  do {                                                                                                       \
    ncclResult_t r__ = (call);                                                                               \
    if (r__ != ncclSuccess) {                                                                                \
      rfslam::set_error("NCCL error %s at %s:%d (%s)", g_nccl.GetErrorString(r__), __FILE__, __LINE__, #call); \
      throw rfslam::CudaFail{RFSLAM_ENODEV};                                                                 \
    }                                                                                                        \
  } while (0)

// interleave the shards' per-tree segments back into ascending tree order: one block per tree
template <typename T>
__global__ void merge_segments_kernel(T* dst, const T* const* src, const int* srcRank, const long long* srcOff, const long long* dstOff,
                                      const long long* len, int width) {
  const int b = blockIdx.x;
  const T* s = src[srcRank[b]] + srcOff[b] * width;
  T* d = dst + dstOff[b] * width;
  const long long cnt = len[b] * width;
  for (long long k = threadIdx.x; k < cnt; k += blockDim.x) d[k] = s[k];
}

}  // namespace

}  // namespace rfslam

using namespace rfslam;

struct rfslam_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
  cudaStream_t stream = nullptr;
  long long allreduces = 0, gathers = 0;       // collectives issued (for the caller's log)
};

namespace rfslam {

// (1) sum the ensemble accumulators over ranks, in place, on the device
void comm_allreduce_ensembles(rfslam_comm* c, double* allNum, double* oobNum, uint32_t* allDen, uint32_t* oobDen, int n) {
  if (!c || c->world < 2) return;
  RF_CUDA(cudaDeviceSynchronize());                                       // the accumulators are final
  RF_NCCL(g_nccl.GroupStart());
  RF_NCCL(g_nccl.AllReduce(allNum, allNum, 2 * (size_t)n, ncclDouble, ncclSum, c->comm, c->stream));
  RF_NCCL(g_nccl.AllReduce(oobNum, oobNum, 2 * (size_t)n, ncclDouble, ncclSum, c->comm, c->stream));
  RF_NCCL(g_nccl.AllReduce(allDen, allDen, (size_t)n, ncclUint32, ncclSum, c->comm, c->stream));
  RF_NCCL(g_nccl.AllReduce(oobDen, oobDen, (size_t)n, ncclUint32, ncclSum, c->comm, c->stream));
  RF_NCCL(g_nccl.GroupEnd());
  RF_CUDA(cudaStreamSynchronize(c->stream));
  c->allreduces += 4;
}

void comm_allreduce_sum_u64(rfslam_comm* c, unsigned long long* dev, int count) {
  if (!c || c->world < 2) return;
  RF_NCCL(g_nccl.AllReduce(dev, dev, (size_t)count, ncclUint64, ncclSum, c->comm, c->stream));
  RF_CUDA(cudaStreamSynchronize(c->stream));
  c->allreduces += 1;
}

// (2) gather-v of the shard arrays to rank 0, merged into ascending tree order.  All of a call's arrays share one layout
// per unit kind (records, leaves): the offset tables go to the device once, every array is then three asynchronous steps
// on the communicator's stream -- NCCL recv of the other ranks' segments over NVLink, one merge kernel, nothing else.  The
// merge writes the caller's destination directly: a device array, or the PAGE-LOCKED HOST block the forest is returned in
// (host memory from cudaHostAlloc is device-accessible under unified addressing), so rank 0's download IS the merge --
// no merged device copy, no second pass over PCIe.  comm_gather_finish synchronises and releases the scratch.
void comm_gather_layout(rfslam_comm* c, const std::vector<long long>& unitsOfTree, GatherLayout& L) {
  const int W = c->world, ntree = (int)unitsOfTree.size();
  L.W = W; L.ntree = ntree; L.rankUnits.assign(W, 0); L.total = 0;
  std::vector<long long> srcOff(ntree), dstOff(ntree), len(ntree);
  std::vector<int> srcRank(ntree);
  for (int b = 0; b < ntree; b++) {
    const int r = b % W;
    srcRank[b] = r; srcOff[b] = L.rankUnits[r]; dstOff[b] = L.total; len[b] = unitsOfTree[b];
    L.rankUnits[r] += unitsOfTree[b]; L.total += unitsOfTree[b];
  }
  if (c->rank != 0) return;
  RF_CUDA(cudaMalloc((void**)&L.d_rank, (size_t)ntree * 4));
  RF_CUDA(cudaMalloc((void**)&L.d_so, (size_t)ntree * 8)); RF_CUDA(cudaMalloc((void**)&L.d_do, (size_t)ntree * 8)); RF_CUDA(cudaMalloc((void**)&L.d_len, (size_t)ntree * 8));
  RF_CUDA(cudaMemcpy(L.d_rank, srcRank.data(), (size_t)ntree * 4, cudaMemcpyHostToDevice));
  RF_CUDA(cudaMemcpy(L.d_so, srcOff.data(), (size_t)ntree * 8, cudaMemcpyHostToDevice));
  RF_CUDA(cudaMemcpy(L.d_do, dstOff.data(), (size_t)ntree * 8, cudaMemcpyHostToDevice));
  RF_CUDA(cudaMemcpy(L.d_len, len.data(), (size_t)ntree * 8, cudaMemcpyHostToDevice));
}

template <typename T>
void comm_gather_into(rfslam_comm* c, GatherLayout& L, const T* local, int width, ncclDataType_t dt, T* dst) {
  const int W = c->world;
  std::vector<void*> recv(W, nullptr);
  if (c->rank == 0) {
    for (int r = 1; r < W; r++) {
      void* q = nullptr;
      RF_CUDA(cudaMalloc(&q, std::max<size_t>((size_t)L.rankUnits[r] * width * sizeof(T), 16)));
      recv[r] = q; L.devKeep.push_back(q);
    }
    recv[0] = const_cast<T*>(local);
  }
  RF_NCCL(g_nccl.GroupStart());
  if (c->rank == 0) { for (int r = 1; r < W; r++) RF_NCCL(g_nccl.Recv(recv[r], (size_t)L.rankUnits[r] * width, dt, r, c->comm, c->stream)); }
  else RF_NCCL(g_nccl.Send(local, (size_t)L.rankUnits[c->rank] * width, dt, 0, c->comm, c->stream));
  RF_NCCL(g_nccl.GroupEnd());
  c->gathers += 1;
  if (c->rank != 0) return;
  void* d_src = nullptr;
  RF_CUDA(cudaMalloc(&d_src, W * sizeof(void*)));
  L.devKeep.push_back(d_src);
  L.hostKeep.push_back(recv);                                        // (alive until the copy below has run)
  RF_CUDA(cudaMemcpyAsync(d_src, L.hostKeep.back().data(), W * sizeof(void*), cudaMemcpyHostToDevice, c->stream));
  merge_segments_kernel<T><<<L.ntree, 256, 0, c->stream>>>(dst, (const T* const*)d_src, L.d_rank, L.d_so, L.d_do, L.d_len, width);
  RF_CUDA(cudaGetLastError());
}
void comm_gather_i32_into(rfslam_comm* c, GatherLayout& L, const int32_t* local, int width, int32_t* dst) { comm_gather_into<int32_t>(c, L, local, width, ncclInt32, dst); }
void comm_gather_f64_into(rfslam_comm* c, GatherLayout& L, const double* local, int width, double* dst) { comm_gather_into<double>(c, L, local, width, ncclDouble, dst); }

void comm_gather_finish(rfslam_comm* c, GatherLayout& L) {
  cudaError_t e = cudaStreamSynchronize(c->stream);
  for (void* q : L.devKeep) cudaFree(q);
  L.devKeep.clear(); L.hostKeep.clear();
  for (void* q : {(void*)L.d_rank, (void*)L.d_so, (void*)L.d_do, (void*)L.d_len}) if (q) cudaFree(q);
  L.d_rank = nullptr; L.d_so = L.d_do = L.d_len = nullptr;
  if (e != cudaSuccess) { set_error("CUDA error %s while gathering the forest", cudaGetErrorString(e)); throw CudaFail{RFSLAM_ENODEV}; }
}

// leafCount of every tree of the forest on every rank: rank r contributes its trees at b = r, r + W, ...
void comm_allgather_leaf_counts(rfslam_comm* c, const std::vector<int32_t>& mine, int ntree, std::vector<int32_t>& all) {
  const int W = c->world;
  const int per = (ntree + W - 1) / W;
  int32_t *d_in = nullptr, *d_out = nullptr;
  std::vector<int32_t> in(per, 0), out((size_t)per * W, 0);
  for (size_t k = 0; k < mine.size() && k < (size_t)per; k++) in[k] = mine[k];
  try {
    RF_CUDA(cudaMalloc((void**)&d_in, (size_t)per * 4)); RF_CUDA(cudaMalloc((void**)&d_out, (size_t)per * W * 4));
    RF_CUDA(cudaMemcpyAsync(d_in, in.data(), (size_t)per * 4, cudaMemcpyHostToDevice, c->stream));
    RF_NCCL(g_nccl.AllGather(d_in, d_out, (size_t)per, ncclInt32, c->comm, c->stream));
    RF_CUDA(cudaMemcpyAsync(out.data(), d_out, (size_t)per * W * 4, cudaMemcpyDeviceToHost, c->stream));
    RF_CUDA(cudaStreamSynchronize(c->stream));
  } catch (...) { if (d_in) cudaFree(d_in); if (d_out) cudaFree(d_out); throw; }
  cudaFree(d_in); cudaFree(d_out);
  c->gathers += 1;
  all.assign(ntree, 0);
  for (int b = 0; b < ntree; b++) all[b] = out[(size_t)(b % W) * per + b / W];
}

int comm_rank(const rfslam_comm* c) { return c ? c->rank : 0; }
int comm_world(const rfslam_comm* c) { return c ? c->world : 1; }

}  // namespace rfslam

extern "C" {

int rfslam_b200_comm_unique_id(void* id, int bytes) {
  if (!id || bytes < (int)sizeof(ncclUniqueId)) { set_error("unique id buffer must hold %d bytes", (int)sizeof(ncclUniqueId)); return RFSLAM_EINVAL; }
  if (!nccl_load()) return RFSLAM_ENODEV;
  ncclUniqueId u;
  ncclResult_t r = g_nccl.GetUniqueId(&u);
  if (r != ncclSuccess) { set_error("ncclGetUniqueId: %s", g_nccl.GetErrorString(r)); return RFSLAM_ENODEV; }
  memcpy(id, &u, sizeof u);
  return RFSLAM_OK;
}

int rfslam_b200_comm_init(rfslam_comm** out, int rank, int world, const void* id, int device) {
  if (!out || !id || world < 1 || rank < 0 || rank >= world) { set_error("bad communicator arguments"); return RFSLAM_EINVAL; }
  *out = nullptr;
  if (!nccl_load()) return RFSLAM_ENODEV;
  rfslam_comm* c = new rfslam_comm();
  try {
    if (device >= 0) RF_CUDA(cudaSetDevice(device));
    RF_CUDA(cudaGetDevice(&c->device));
    c->rank = rank; c->world = world;
    ncclUniqueId u; memcpy(&u, id, sizeof u);
    RF_NCCL(g_nccl.CommInitRank(&c->comm, world, u, rank));
    RF_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  } catch (CudaFail f) {
    if (c->comm) g_nccl.CommDestroy(c->comm);
    delete c;
    return f.code;
  }
  *out = c;
  return RFSLAM_OK;
}

void rfslam_b200_comm_destroy(rfslam_comm* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
  if (c->comm) g_nccl.CommDestroy(c->comm);
  delete c;
}

int rfslam_b200_comm_rank(const rfslam_comm* c) { return comm_rank(c); }
int rfslam_b200_comm_world(const rfslam_comm* c) { return comm_world(c); }
int64_t rfslam_b200_comm_collectives(const rfslam_comm* c, int64_t* allreduces, int64_t* gathers) {
  if (!c) return 0;
  if (allreduces) *allreduces = c->allreduces;
  if (gathers) *gathers = c->gathers;
  return c->allreduces + c->gathers;
}
int rfslam_b200_nccl_version(void) {
  if (!nccl_load()) return 0;
  int v = 0; g_nccl.GetVersion(&v); return v;
}

}  // extern "C"
