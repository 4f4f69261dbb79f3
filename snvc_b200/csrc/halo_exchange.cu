// Depth-slab halo exchange for the single-volume stress configuration (SURVEY.md 8(b) `snvc_halo_exchange`, 8(e) cfg-5).
//
// The reference has no counterpart: its only parallelism is nn.DataParallel over the batch
// (tools/inference_agnostic.py:472).  Here ONE volume is split along depth over the ranks; every 3x3x3 convolution
// needs one plane from each neighbour (a stride-2 conv needs only the lower one, a transposed conv only the upper
// one), so after each layer a rank sends its first / last REAL plane to rank r-1 / r+1 and receives their last /
// first real plane into its inner halo planes -- one ncclGroup of up to two sends and two receives on the caller's
// stream, NVLink 5 / NVSwitch point to point, no host synchronisation.  At the global boundary the inner halo plane is
// zero-filled instead (it stands for the convolution's zero padding).
//
// NCCL is resolved at run time (dlopen of the libnccl the process already has -- torch's bundled one -- so the library
// still loads, and exports every symbol, on a CPU-only host).
#include <dlfcn.h>

#include <mutex>

#include "common.cuh"

namespace snvc {
namespace {

typedef struct { char internal[128]; } NcclUniqueId;
typedef void* NcclComm;
typedef int (*FnGetUniqueId)(NcclUniqueId*);
typedef int (*FnCommInitRank)(NcclComm*, int, NcclUniqueId, int);
typedef int (*FnCommDestroy)(NcclComm);
typedef int (*FnGroup)(void);
typedef int (*FnSendRecv)(void*, size_t, int, int, NcclComm, cudaStream_t);   // ncclSend(const void*...) / ncclRecv
typedef const char* (*FnErr)(int);

struct Nccl {
  FnGetUniqueId GetUniqueId = nullptr;
  FnCommInitRank CommInitRank = nullptr;
  FnCommDestroy CommDestroy = nullptr;
  FnGroup GroupStart = nullptr, GroupEnd = nullptr;
  FnSendRecv Send = nullptr, Recv = nullptr;
  FnErr GetErrorString = nullptr;
  bool ok = false;
};

const Nccl& nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = nullptr;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      h = dlopen(name, RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD);     // the copy torch has already loaded, if any
      if (!h) h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) return;
    n.GetUniqueId = (FnGetUniqueId)dlsym(h, "ncclGetUniqueId");
    n.CommInitRank = (FnCommInitRank)dlsym(h, "ncclCommInitRank");
    n.CommDestroy = (FnCommDestroy)dlsym(h, "ncclCommDestroy");
    n.GroupStart = (FnGroup)dlsym(h, "ncclGroupStart");
    n.GroupEnd = (FnGroup)dlsym(h, "ncclGroupEnd");
    n.Send = (FnSendRecv)dlsym(h, "ncclSend");
    n.Recv = (FnSendRecv)dlsym(h, "ncclRecv");
    n.GetErrorString = (FnErr)dlsym(h, "ncclGetErrorString");
    n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.GroupStart && n.GroupEnd && n.Send && n.Recv;
  });
  return n;
}

int nccl_fail(int r, const char* what) {
  const Nccl& n = nccl();
  return fail(1000 + r, "%s failed: %s", what, n.GetErrorString ? n.GetErrorString(r) : "NCCL error");
}

#define SNVC_NCCL_OK(expr)                         \
  do {                                             \
    int r__ = (expr);                              \
    if (r__ != 0) return nccl_fail(r__, #expr);    \
  } while (0)

}  // namespace
}  // namespace snvc

using namespace snvc;

extern "C" int snvc_halo_unique_id(void* id128) {
  SNVC_CHECK_ARG(id128 != nullptr, "null pointer");
  if (!nccl().ok) return fail(SNVC_E_DRIVER, "NCCL (libnccl.so.2) is not available in this process");
  SNVC_NCCL_OK(nccl().GetUniqueId(static_cast<NcclUniqueId*>(id128)));
  return 0;
}

extern "C" int snvc_halo_comm_create(const void* id128, int32_t world, int32_t rank, void** comm) {
  SNVC_CHECK_ARG(id128 && comm && world >= 1 && rank >= 0 && rank < world, "bad arguments");
  if (!nccl().ok) return fail(SNVC_E_DRIVER, "NCCL (libnccl.so.2) is not available in this process");
  NcclUniqueId id = *static_cast<const NcclUniqueId*>(id128);
  NcclComm c = nullptr;
  SNVC_NCCL_OK(nccl().CommInitRank(&c, world, id, rank));
  *comm = c;
  return 0;
}

extern "C" int snvc_halo_comm_destroy(void* comm) {
  if (comm == nullptr) return 0;
  if (!nccl().ok) return fail(SNVC_E_DRIVER, "NCCL is not available");
  SNVC_NCCL_OK(nccl().CommDestroy(comm));
  return 0;
}

extern "C" int snvc_halo_exchange(void* comm, void* x, int64_t planes_ext, int64_t plane_bytes, int32_t halo, int32_t rank,
                                  int32_t world, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SNVC_CHECK_ARG(x != nullptr && plane_bytes > 0 && halo >= 1 && planes_ext >= 2 * halo + 1, "bad slab geometry");
  SNVC_CHECK_ARG(world >= 1 && rank >= 0 && rank < world, "bad rank / world");
  SNVC_CHECK_ARG(world == 1 || comm != nullptr, "a communicator is required for world > 1");
  char* base = static_cast<char*>(x);
  char* first_real = base + (int64_t)halo * plane_bytes;
  char* last_real = base + (planes_ext - halo - 1) * plane_bytes;
  char* halo_lo = base + (int64_t)(halo - 1) * plane_bytes;            // inner halo plane below the slab
  char* halo_hi = base + (planes_ext - halo) * plane_bytes;            // inner halo plane above the slab
  const bool has_lo = rank > 0, has_hi = rank < world - 1;
  if (!has_lo) SNVC_CUDA_OK(cudaMemsetAsync(halo_lo, 0, (size_t)plane_bytes, stream));
  if (!has_hi) SNVC_CUDA_OK(cudaMemsetAsync(halo_hi, 0, (size_t)plane_bytes, stream));
  if (!has_lo && !has_hi) return 0;
  const Nccl& n = nccl();
  if (!n.ok) return fail(SNVC_E_DRIVER, "NCCL is not available");
  SNVC_NCCL_OK(n.GroupStart());
  int r = 0;
  // ncclInt8 == 0: plane_bytes elements of one byte
  if (has_lo) {
    if (!r) r = n.Send(first_real, (size_t)plane_bytes, 0, rank - 1, comm, stream);
    if (!r) r = n.Recv(halo_lo, (size_t)plane_bytes, 0, rank - 1, comm, stream);
  }
  if (has_hi) {
    if (!r) r = n.Send(last_real, (size_t)plane_bytes, 0, rank + 1, comm, stream);
    if (!r) r = n.Recv(halo_hi, (size_t)plane_bytes, 0, rank + 1, comm, stream);
  }
  const int e = n.GroupEnd();
  if (r) return nccl_fail(r, "ncclSend / ncclRecv");
  if (e) return nccl_fail(e, "ncclGroupEnd");
  return 0;
}
