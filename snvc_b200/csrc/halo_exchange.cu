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

#include <algorithm>
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

// ---- peer-memory halo push ------------------------------------------------------------------------------------
// The NCCL path above moves a 30.7 MB plane at 146 GB/s per direction and neighbour on the 8-GPU box (ncclSend / ncclRecv
// run on a couple of channels, scripts/halo_bw_probe.py) -- at 8 ranks the ten exchanges of a volume cost more than half
// of the slab's kernels.  NVLink 5 / NVSwitch give a GPU 900 GB/s out, so the product path pushes the planes itself: the
// slab buffers of every rank live in a cudaMalloc'ed arena that the neighbours map through CUDA IPC, one kernel per layer
// (all SMs, 16-byte loads from the local slab, 16-byte peer stores straight into the neighbours' inner halo planes),
// and the same kernel is the neighbour barrier: the last block to finish publishes a monotonically increasing epoch in
// the neighbours' control words (release at system scope, after every block's stores were fenced) and waits until both
// neighbours have published theirs, i.e. until their planes have landed here.  No NCCL, no host, capturable in a graph
// (the epoch lives in device memory, so a replay needs no new arguments).
struct PeerCtl {
  // words written by the neighbours through their mapping of this arena
  unsigned long long ready_from_lo, ready_from_hi;   // "my kernels that write the slab of push #e have finished"
  unsigned long long done_from_lo, done_from_hi;     // "my planes of push #e have landed in your slab"
  // local words
  unsigned long long epoch;                          // pushes completed by this rank
  unsigned int blocks_done;                          // arrival counter of the running push
  unsigned int pad;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// bounded wait (a lost neighbour must not hang the GPU): ~4 s at 2 GHz, then trap
__device__ __forceinline__ void wait_flags(const unsigned long long* a, const unsigned long long* b, unsigned long long e) {
  const long long t0 = clock64();
  while ((a && ld_acquire_sys(a) < e) || (b && ld_acquire_sys(b) < e)) {
    if (clock64() - t0 > 8000000000ll) __trap();
    __nanosleep(100);
  }
}

// grid <= the number of co-resident blocks (every block waits for the neighbours' "ready" before it stores)
__global__ void __launch_bounds__(512)
halo_push_kernel(const uint4* __restrict__ first_real, const uint4* __restrict__ last_real, uint4* __restrict__ halo_lo,
                 uint4* __restrict__ halo_hi, uint4* __restrict__ peer_lo_halo_hi, uint4* __restrict__ peer_hi_halo_lo,
                 int64_t n16, PeerCtl* ctl, PeerCtl* ctl_lo, PeerCtl* ctl_hi) {
  // phase 0 -- the neighbours' slabs must be ready to receive: the convolution that produced THEIR slab also wrote
  // (meaningless) values into its inner halo planes, so a plane pushed before that kernel has finished would be
  // overwritten.  "ready" = this rank's stream has reached its push, i.e. its producer kernel is complete.
  // (`epoch` is advanced by the last block only after every block has arrived, so all blocks read the same value.)
  const unsigned long long e = *reinterpret_cast<volatile unsigned long long*>(&ctl->epoch) + 1ull;
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) {
      __threadfence_system();
      if (ctl_lo) st_release_sys(&ctl_lo->ready_from_hi, e);   // we are rank-1's upper neighbour
      if (ctl_hi) st_release_sys(&ctl_hi->ready_from_lo, e);
    }
    wait_flags(ctl_lo ? &ctl->ready_from_lo : nullptr, ctl_hi ? &ctl->ready_from_hi : nullptr, e);
  }
  __syncthreads();
  // phase 1 -- the first real plane goes up into rank-1's upper inner halo, the last one down into rank+1's lower inner
  // halo; at the ends of the volume the inner halo plane is the convolution's zero padding
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
    if (peer_lo_halo_hi) peer_lo_halo_hi[i] = first_real[i];
    else halo_lo[i] = make_uint4(0u, 0u, 0u, 0u);
    if (peer_hi_halo_lo) peer_hi_halo_lo[i] = last_real[i];
    else halo_hi[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __threadfence_system();                                  // this thread's peer stores are performed system-wide ...
  __syncthreads();
  if (threadIdx.x != 0) return;
  if (atomicAdd(&ctl->blocks_done, 1u) != gridDim.x - 1) return;   // ... before its block counts as done
  // phase 2 (last block) -- publish "landed", then wait for the neighbours' planes
  __threadfence();
  ctl->blocks_done = 0u;
  ctl->epoch = e;
  __threadfence_system();
  if (ctl_lo) st_release_sys(&ctl_lo->done_from_hi, e);
  if (ctl_hi) st_release_sys(&ctl_hi->done_from_lo, e);
  wait_flags(ctl_lo ? &ctl->done_from_lo : nullptr, ctl_hi ? &ctl->done_from_hi : nullptr, e);
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

// ---- peer-memory path: arena + IPC mapping + push ------------------------------------------------------------------
extern "C" int snvc_peer_alloc(int64_t bytes, void** ptr) {
  SNVC_CHECK_ARG(bytes > 0 && ptr != nullptr, "bad arguments");
  void* p = nullptr;
  SNVC_CUDA_OK(cudaMalloc(&p, (size_t)bytes));            // a plain cudaMalloc block: exportable with cudaIpcGetMemHandle
  SNVC_CUDA_OK(cudaMemset(p, 0, (size_t)bytes));
  *ptr = p;
  return 0;
}

extern "C" int snvc_peer_free(void* ptr) {
  if (ptr) SNVC_CUDA_OK(cudaFree(ptr));
  return 0;
}

extern "C" int snvc_peer_export(void* ptr, void* handle64) {
  SNVC_CHECK_ARG(ptr && handle64, "null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  SNVC_CUDA_OK(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle64), ptr));
  return 0;
}

extern "C" int snvc_peer_open(const void* handle64, void** peer_ptr) {
  SNVC_CHECK_ARG(handle64 && peer_ptr, "null pointer");
  void* p = nullptr;
  SNVC_CUDA_OK(cudaIpcOpenMemHandle(&p, *static_cast<const cudaIpcMemHandle_t*>(handle64), cudaIpcMemLazyEnablePeerAccess));
  *peer_ptr = p;
  return 0;
}

extern "C" int snvc_peer_close(void* peer_ptr) {
  if (peer_ptr) SNVC_CUDA_OK(cudaIpcCloseMemHandle(peer_ptr));
  return 0;
}

extern "C" int64_t snvc_peer_ctl_bytes(void) { return 256; }

extern "C" int snvc_halo_push(void* x, void* x_in_lo_peer, void* x_in_hi_peer, int64_t planes_ext, int64_t plane_bytes,
                              int32_t halo, void* ctl, void* ctl_lo_peer, void* ctl_hi_peer, int32_t max_blocks,
                              void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SNVC_CHECK_ARG(x != nullptr && ctl != nullptr && plane_bytes > 0 && plane_bytes % 16 == 0 && halo >= 1 &&
                     planes_ext >= 2 * halo + 1, "bad slab geometry");
  SNVC_CHECK_ARG((x_in_lo_peer == nullptr) == (ctl_lo_peer == nullptr) && (x_in_hi_peer == nullptr) == (ctl_hi_peer == nullptr),
                 "a neighbour needs both its slab and its control block");
  SNVC_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(x_in_lo_peer) |
                   reinterpret_cast<uintptr_t>(x_in_hi_peer)) & 15) == 0, "slabs must be 16-byte aligned");
  char* base = static_cast<char*>(x);
  const int64_t o_first = (int64_t)halo * plane_bytes, o_last = (planes_ext - halo - 1) * plane_bytes;
  const int64_t o_halo_lo = (int64_t)(halo - 1) * plane_bytes, o_halo_hi = (planes_ext - halo) * plane_bytes;
  const int64_t n16 = plane_bytes / 16;
  // every block waits inside the kernel, so the grid must be co-resident: at most one 512-thread block per SM
  int blocks = (int)std::min<int64_t>(ceil_div(n16, 512 * 4), std::min<int64_t>(max_blocks > 0 ? max_blocks : sm_count(), sm_count()));
  if (blocks < 1) blocks = 1;
  halo_push_kernel<<<blocks, 512, 0, stream>>>(
      reinterpret_cast<const uint4*>(base + o_first), reinterpret_cast<const uint4*>(base + o_last),
      reinterpret_cast<uint4*>(base + o_halo_lo), reinterpret_cast<uint4*>(base + o_halo_hi),
      x_in_lo_peer ? reinterpret_cast<uint4*>(static_cast<char*>(x_in_lo_peer) + o_halo_hi) : nullptr,
      x_in_hi_peer ? reinterpret_cast<uint4*>(static_cast<char*>(x_in_hi_peer) + o_halo_lo) : nullptr, n16,
      static_cast<PeerCtl*>(ctl), static_cast<PeerCtl*>(ctl_lo_peer), static_cast<PeerCtl*>(ctl_hi_peer));
  return launch_status("halo_push_kernel");
}
