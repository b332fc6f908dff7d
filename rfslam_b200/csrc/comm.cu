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

// (2) gather-v of one shard array to rank 0 and merge into tree order.  `local` holds this rank's trees
// back to back (`width` elements per unit); unitsOfTree[b] = units of tree b + 1 (all ranks know all trees).
// Rank 0 receives into scratch and returns the merged device array (caller frees with cudaFree); others return nullptr.
template <typename T>
T* comm_gather_merge(rfslam_comm* c, const T* local, const std::vector<long long>& unitsOfTree, int width, ncclDataType_t dt) {
  const int W = c->world, ntree = (int)unitsOfTree.size();
  std::vector<long long> rankUnits(W, 0), srcOff(ntree), dstOff(ntree), len(ntree);
  std::vector<int> srcRank(ntree);
  long long total = 0;
  for (int b = 0; b < ntree; b++) {
    const int r = b % W;
    srcRank[b] = r; srcOff[b] = rankUnits[r]; dstOff[b] = total; len[b] = unitsOfTree[b];
    rankUnits[r] += unitsOfTree[b]; total += unitsOfTree[b];
  }
  std::vector<T*> recv(W, nullptr);
  T* merged = nullptr;
  auto cleanup = [&]() { for (int r = 1; r < W; r++) if (recv[r]) cudaFree(recv[r]); };
  try {
    if (c->rank == 0) {
      for (int r = 1; r < W; r++) RF_CUDA(cudaMalloc((void**)&recv[r], std::max<size_t>((size_t)rankUnits[r] * width * sizeof(T), 16)));
      recv[0] = const_cast<T*>(local);
    }
    RF_NCCL(g_nccl.GroupStart());
    if (c->rank == 0) { for (int r = 1; r < W; r++) RF_NCCL(g_nccl.Recv(recv[r], (size_t)rankUnits[r] * width, dt, r, c->comm, c->stream)); }
    else RF_NCCL(g_nccl.Send(local, (size_t)rankUnits[c->rank] * width, dt, 0, c->comm, c->stream));
    RF_NCCL(g_nccl.GroupEnd());
    c->gathers += 1;
    if (c->rank == 0) {
      RF_CUDA(cudaMalloc((void**)&merged, std::max<size_t>((size_t)total * width * sizeof(T), 16)));
      T** d_src = nullptr; int* d_rank = nullptr; long long *d_so = nullptr, *d_do = nullptr, *d_len = nullptr;
      RF_CUDA(cudaMalloc((void**)&d_src, W * sizeof(T*))); RF_CUDA(cudaMalloc((void**)&d_rank, (size_t)ntree * 4));
      RF_CUDA(cudaMalloc((void**)&d_so, (size_t)ntree * 8)); RF_CUDA(cudaMalloc((void**)&d_do, (size_t)ntree * 8)); RF_CUDA(cudaMalloc((void**)&d_len, (size_t)ntree * 8));
      RF_CUDA(cudaMemcpyAsync(d_src, recv.data(), W * sizeof(T*), cudaMemcpyHostToDevice, c->stream));
      RF_CUDA(cudaMemcpyAsync(d_rank, srcRank.data(), (size_t)ntree * 4, cudaMemcpyHostToDevice, c->stream));
      RF_CUDA(cudaMemcpyAsync(d_so, srcOff.data(), (size_t)ntree * 8, cudaMemcpyHostToDevice, c->stream));
      RF_CUDA(cudaMemcpyAsync(d_do, dstOff.data(), (size_t)ntree * 8, cudaMemcpyHostToDevice, c->stream));
      RF_CUDA(cudaMemcpyAsync(d_len, len.data(), (size_t)ntree * 8, cudaMemcpyHostToDevice, c->stream));
      merge_segments_kernel<T><<<ntree, 256, 0, c->stream>>>(merged, (const T* const*)d_src, d_rank, d_so, d_do, d_len, width);
      RF_CUDA(cudaGetLastError());
      RF_CUDA(cudaStreamSynchronize(c->stream));
      cudaFree(d_src); cudaFree(d_rank); cudaFree(d_so); cudaFree(d_do); cudaFree(d_len);
    } else {
      RF_CUDA(cudaStreamSynchronize(c->stream));
    }
  } catch (...) { cleanup(); if (merged) cudaFree(merged); throw; }
  cleanup();
  return merged;
}
template int32_t* comm_gather_merge<int32_t>(rfslam_comm*, const int32_t*, const std::vector<long long>&, int, ncclDataType_t);
template double* comm_gather_merge<double>(rfslam_comm*, const double*, const std::vector<long long>&, int, ncclDataType_t);

int32_t* comm_gather_i32(rfslam_comm* c, const int32_t* local, const std::vector<long long>& units, int width) { return comm_gather_merge<int32_t>(c, local, units, width, ncclInt32); }
double*  comm_gather_f64(rfslam_comm* c, const double* local, const std::vector<long long>& units, int width) { return comm_gather_merge<double>(c, local, units, width, ncclDouble); }

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
