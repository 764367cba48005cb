// comm.cu -- multi-GPU plumbing: one process per GPU, NCCL over NVLink / NVSwitch.
//
// The reference has no distributed layer at all; this file adds only what
// BASELINE.json's north_star partitions: reductions combine per-GPU partials with an
// allreduce, and the heat grid is slab-decomposed along axis 0 with one-plane halos
// exchanged by grouped ncclSend/ncclRecv on a side stream while the interior planes are
// being updated on the main stream.
#include "ph_common.cuh"
#include "comm.cuh"
#include "ops.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <dlfcn.h>
#include <vector>
#include <algorithm>
#include <nccl.h>   // types only: the library is resolved at run time (see nccl_api)

namespace ph {

// NCCL is bound lazily with dlopen: a process that already loaded an NCCL (e.g. the one
// bundled with PyTorch) keeps using that copy, and a process that never calls ph_comm_*
// never loads NCCL at all.
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
};
static NcclApi& nccl() {
  static NcclApi api;
  return api;
}
static int32_t nccl_load() {
  NcclApi& n = nccl();
  if (n.handle) return PH_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);       // already in the process?
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
  if (!h) return set_error(PH_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
#define PH_SYM(field, name)                                                             \
  *(void**)(&n.field) = dlsym(h, name);                                                 \
  if (!n.field) return set_error(PH_ERR_NCCL, "libnccl is missing symbol %s", name);
  PH_SYM(GetUniqueId, "ncclGetUniqueId") PH_SYM(CommInitRank, "ncclCommInitRank")
  PH_SYM(CommDestroy, "ncclCommDestroy") PH_SYM(AllReduce, "ncclAllReduce")
  PH_SYM(AllGather, "ncclAllGather") PH_SYM(Send, "ncclSend") PH_SYM(Recv, "ncclRecv")
  PH_SYM(GroupStart, "ncclGroupStart") PH_SYM(GroupEnd, "ncclGroupEnd")
  PH_SYM(GetErrorString, "ncclGetErrorString")
#undef PH_SYM
  n.handle = h;
  return PH_OK;
}
#define PH_NCCL_LOAD()                      \
  do {                                      \
    int32_t _s = nccl_load();               \
    if (_s != PH_OK) return _s;             \
  } while (0)

struct Comm {
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  bool inited = false;
};
static Comm& cm() {
  static Comm c;
  return c;
}

// data-path NCCL calls (all-reduce, all-gather, send, recv) issued since ph_init: the peer-memory forms must leave
// this number unchanged over a timed region (bench.py prints the difference beside every multi-GPU leg)
static long long g_nccl_calls = 0;

static int32_t check_nccl(ncclResult_t r, const char* what) {
  if (r == ncclSuccess) return PH_OK;
  return set_error(PH_ERR_NCCL, "NCCL error %d (%s) at %s", (int)r, nccl().GetErrorString(r), what);
}
#define PH_NCCL(call)                                   \
  do {                                                  \
    int32_t _s = check_nccl((call), #call);             \
    if (_s != PH_OK) return _s;                         \
  } while (0)
#define PH_NCCL_DATA(call) do { g_nccl_calls++; PH_NCCL(call); } while (0)

int32_t heat_slab_dispatch(int32_t dtype, int rank, const int64_t* ext, const void* coeff_host, int ghost, int has_lo,
                           int has_hi, int64_t p_begin, int64_t p_end, const void* in, void* out,
                           cudaStream_t stream, bool two_step, const HeatMirror* mir, bool* mirrored);
bool heat_two_step_usable(int32_t dtype, int rank, const int64_t* ext);

// ---------------------------------------------------------------- peer-mapped memory (CUDA IPC over NVLink)
PeerInfo& peers() {
  static PeerInfo p;
  return p;
}
bool comm_is_multi() { return cm().inited && cm().nranks > 1; }

struct SymmAlloc {
  char* local = nullptr;
  size_t nbytes = 0;
  bool mapped = false;
  char* peer[PH_MAX_PEERS] = {nullptr};
};
static std::vector<SymmAlloc>& symm_table() {
  static std::vector<SymmAlloc> t;
  return t;
}
static const SymmAlloc* symm_find(const void* p) {
  const char* q = reinterpret_cast<const char*>(p);
  for (const SymmAlloc& a : symm_table())
    if (q >= a.local && q < a.local + a.nbytes) return &a;
  return nullptr;
}

static int32_t ensure_exchange_buffers() {
  PeerInfo& p = peers();
  if (!p.gather_send) {
    PH_CUDA(cudaMalloc((void**)&p.gather_send, sizeof(ReduceSlot) * PH_MAX_PEERS));
    PH_CUDA(cudaMalloc((void**)&p.gather_recv, sizeof(ReduceSlot) * PH_MAX_PEERS));
    PH_CUDA(cudaMemset(p.gather_send, 0, sizeof(ReduceSlot) * PH_MAX_PEERS));
    PH_CUDA(cudaMemset(p.gather_recv, 0, sizeof(ReduceSlot) * PH_MAX_PEERS));
  }
  if (!p.host_result) {
    PH_CUDA(cudaHostAlloc((void**)&p.host_result, 64, cudaHostAllocMapped));
    memset(p.host_result, 0, 64);
    PH_CUDA(cudaHostGetDevicePointer((void**)&p.host_result_dev, p.host_result, 0));
  }
  return PH_OK;
}

int32_t comm_allgather_records(cudaStream_t s) {
  Comm& c = cm();
  PeerInfo& p = peers();
  if (c.nranks == 1) {
    PH_CUDA(cudaMemcpyAsync(p.gather_recv, p.gather_send, sizeof(ReduceSlot), cudaMemcpyDeviceToDevice, s));
    return PH_OK;
  }
  PH_NCCL_DATA(nccl().AllGather(p.gather_send + c.rank, p.gather_recv, sizeof(ReduceSlot), ncclUint8, c.comm, s));
  return PH_OK;
}

// 64 bytes from every rank to every rank through the host (set-up paths only)
static int32_t exchange64(const void* mine, void* all /* nranks x 64 */) {
  Comm& c = cm();
  PeerInfo& p = peers();
  cudaStream_t s = rt().stream;
  PH_CUDA(cudaMemcpyAsync(p.gather_send + c.rank, mine, 64, cudaMemcpyHostToDevice, s));
  int32_t st = comm_allgather_records(s);
  if (st != PH_OK) return st;
  PH_CUDA(cudaMemcpyAsync(all, p.gather_recv, (size_t)64 * c.nranks, cudaMemcpyDeviceToHost, s));
  PH_CUDA(cudaStreamSynchronize(s));
  return PH_OK;
}

// every rank reports `ok`; true only when all did (so that all ranks take the same path afterwards)
static int32_t all_agree(bool ok, bool* all_ok) {
  uint64_t mine[8] = {ok ? 1ull : 0ull, 0, 0, 0, 0, 0, 0, 0};
  uint64_t all[8 * PH_MAX_PEERS];
  int32_t st = exchange64(mine, all);
  if (st != PH_OK) return st;
  *all_ok = true;
  for (int r = 0; r < cm().nranks; r++) if (all[8 * r] != 1ull) *all_ok = false;
  return PH_OK;
}

// export `local`, import everybody else's: peer[r] = rank r's allocation mapped into this process
static int32_t map_peers(void* local, char** peer, bool* mapped) {
  Comm& c = cm();
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "an IPC handle travels as one 64-byte record");
  cudaIpcMemHandle_t mine, all[PH_MAX_PEERS];
  memset(&mine, 0, sizeof(mine));
  bool ok = cudaIpcGetMemHandle(&mine, local) == cudaSuccess;
  if (!ok) cudaGetLastError();
  int32_t st = exchange64(&mine, all);
  if (st != PH_OK) return st;
  bool all_exported = false;
  if ((st = all_agree(ok, &all_exported)) != PH_OK) return st;
  for (int r = 0; r < c.nranks; r++) peer[r] = nullptr;
  peer[c.rank] = reinterpret_cast<char*>(local);
  if (all_exported) {
    for (int r = 0; r < c.nranks && ok; r++) {
      if (r == c.rank) continue;
      void* q = nullptr;
      if (cudaIpcOpenMemHandle(&q, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; }
      else peer[r] = reinterpret_cast<char*>(q);
    }
  }
  bool all_mapped = false;
  if ((st = all_agree(ok && all_exported, &all_mapped)) != PH_OK) return st;
  if (!all_mapped)
    for (int r = 0; r < c.nranks; r++)
      if (r != c.rank && peer[r]) { cudaIpcCloseMemHandle(peer[r]); peer[r] = nullptr; }
  *mapped = all_mapped;
  return PH_OK;
}

static void p2p_teardown() {
  PeerInfo& p = peers();
  for (SymmAlloc& a : symm_table()) {
    for (int r = 0; r < PH_MAX_PEERS; r++)
      if (a.mapped && a.peer[r] && a.peer[r] != a.local) cudaIpcCloseMemHandle(a.peer[r]);
    if (a.local) cudaFree(a.local);
  }
  symm_table().clear();
  for (int r = 0; r < PH_MAX_PEERS; r++)
    if (p.ctrl[r] && r != p.rank) cudaIpcCloseMemHandle(p.ctrl[r]);
  if (p.ctrl[p.rank]) cudaFree(p.ctrl[p.rank]);
  if (p.gather_send) cudaFree(p.gather_send);
  if (p.gather_recv) cudaFree(p.gather_recv);
  if (p.host_result) cudaFreeHost(p.host_result);
  cudaGetLastError();
  p = PeerInfo();
}

// Map every peer's control block.  Failure is not an error: the NCCL paths remain (PH_NO_P2P=1 forces that).
static int32_t p2p_setup() {
  Comm& c = cm();
  PeerInfo& p = peers();
  p.nranks = c.nranks;
  p.rank = c.rank;
  int32_t st = ensure_exchange_buffers();
  if (st != PH_OK) return st;
  if (c.nranks == 1 || c.nranks > PH_MAX_PEERS || getenv("PH_NO_P2P")) return PH_OK;
  CtrlBlock* mine = nullptr;
  const size_t bytes = std::max<size_t>(sizeof(CtrlBlock), (size_t)2 << 20);     // its own 2 MiB granule
  PH_CUDA(cudaMalloc((void**)&mine, bytes));
  PH_CUDA(cudaMemset(mine, 0, bytes));
  PH_CUDA(cudaDeviceSynchronize());
  char* peer[PH_MAX_PEERS];
  bool mapped = false;
  if ((st = map_peers(mine, peer, &mapped)) != PH_OK) return st;
  if (!mapped) { cudaFree(mine); return PH_OK; }
  for (int r = 0; r < c.nranks; r++) p.ctrl[r] = reinterpret_cast<CtrlBlock*>(peer[r]);
  p.ready = true;
  return PH_OK;
}

int32_t comm_combine_args(CombineArgs* out, int64_t elems_before) {
  Comm& c = cm();
  PeerInfo& p = peers();
  // no communicator: a single process with one GPU -- the record mode of the reduction kernels (result and
  // flags in the pinned host record, no copy, no stream synchronisation) works on its own
  int32_t st = ensure_exchange_buffers();
  if (st != PH_OK) return st;
  CombineArgs a;
  a.nranks = c.inited ? c.nranks : 1;
  a.rank = c.inited ? c.rank : 0;
  a.seq = ++p.reduce_seq;
  a.elems_before = elems_before;
  a.host_out = p.host_result_dev;
  static const bool force_nccl = getenv("PH_REDUCE_NCCL") != nullptr;       // A/B knob: NCCL transport of the records
  if (a.nranks > 1) {
    if (p.ready && !force_nccl) {
      const int parity = (int)(a.seq & 1u);
      a.my_slots = &p.ctrl[c.rank]->slot[parity][0];
      for (int r = 0; r < c.nranks; r++) a.peer_slots[r] = &p.ctrl[r]->slot[parity][0];
    } else {
      a.my_slots = p.gather_send;
      a.peer_slots[0] = nullptr;
    }
  }
  *out = a;
  return PH_OK;
}

// ---- stream-ordered signal / wait on flag words in the control blocks
typedef CUresult (*StreamWaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static StreamWaitValue32Fn stream_wait_fn() {
  static StreamWaitValue32Fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    if (!getenv("PH_P2P_SPIN_WAIT")) {
      void* q = nullptr;
      cudaDriverEntryPointQueryResult res;
      if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &q, cudaEnableDefault, &res) == cudaSuccess &&
          res == cudaDriverEntryPointSuccess)
        fn = (StreamWaitValue32Fn)q;
      else
        cudaGetLastError();
    }
  }
  return fn;
}
static __global__ void wait_flag_kernel(const volatile uint32_t* flag, uint32_t value, uint32_t* err_flags) {
  const long long t0 = clock64();
  while ((int32_t)(*flag - value) < 0) {
    if (clock64() - t0 > 20000000000LL) { atomicOr(err_flags + 1, 1u); break; }     // ~10 s: a neighbour died
  }
  __threadfence_system();
}
int32_t stream_wait_geq(cudaStream_t s, uint32_t* flag_dev, uint32_t value) {
  if (StreamWaitValue32Fn fn = stream_wait_fn()) {
    CUresult r = fn((CUstream)s, (CUdeviceptr)flag_dev, value, CU_STREAM_WAIT_VALUE_GEQ);
    if (r != CUDA_SUCCESS) return set_error(PH_ERR_CUDA, "cuStreamWaitValue32 failed (%d)", (int)r);
    return PH_OK;
  }
  wait_flag_kernel<<<1, 1, 0, s>>>(flag_dev, value, rt().d_flags);
  PH_LAUNCH_CHECK("wait_flag_kernel");
  return PH_OK;
}

static __device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// "I have reached the start of run `event`, my slab has `my_planes` planes": everything this rank launched
// before is complete (stream order), so the neighbours may write into its ghost planes again
static __global__ void halo_hello_kernel(uint32_t* flag_lo, uint32_t* flag_hi, int64_t* planes_lo, int64_t* planes_hi,
                                         int64_t my_planes, uint32_t event) {
  if (planes_lo) *reinterpret_cast<volatile int64_t*>(planes_lo) = my_planes;
  if (planes_hi) *reinterpret_cast<volatile int64_t*>(planes_hi) = my_planes;
  __threadfence_system();
  if (flag_lo) st_release_sys(flag_lo, event);
  if (flag_hi) st_release_sys(flag_hi, event);
}
// copy my edge planes into the neighbours' ghost planes (peer stores), then release their flags
static __global__ void __launch_bounds__(256) halo_push_kernel(const uint4* __restrict__ src_lo, uint4* __restrict__ dst_lo,
                                                               const uint4* __restrict__ src_hi, uint4* __restrict__ dst_hi,
                                                               int64_t n16, uint32_t* ticket, uint32_t* flag_lo,
                                                               uint32_t* flag_hi, uint32_t event) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
    if (dst_lo) dst_lo[i] = src_lo[i];
    if (dst_hi) dst_hi[i] = src_hi[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
      *ticket = 0;
      __threadfence_system();
      if (flag_lo) st_release_sys(flag_lo, event);
      if (flag_hi) st_release_sys(flag_hi, event);
    }
  }
}

static int32_t nccl_type(int32_t dtype, ncclDataType_t* t) {
  switch (dtype) {
    case PH_F32: *t = ncclFloat32; return PH_OK;
    case PH_F64: *t = ncclFloat64; return PH_OK;
    case PH_I32: *t = ncclInt32; return PH_OK;
    case PH_I64: *t = ncclInt64; return PH_OK;
    case PH_U8: *t = ncclUint8; return PH_OK;
    case PH_I8: *t = ncclInt8; return PH_OK;
    case PH_U32: *t = ncclUint32; return PH_OK;
    case PH_U64: *t = ncclUint64; return PH_OK;
    default: return set_error(PH_ERR_UNSUPPORTED, "dtype %d has no NCCL type", dtype);
  }
}

static int32_t halo_exchange_impl(const void* send_lo, void* recv_lo, int lo_rank, const void* send_hi,
                                  void* recv_hi, int hi_rank, int64_t nbytes, cudaStream_t s) {
  Comm& c = cm();
  if (lo_rank < 0 && hi_rank < 0) return PH_OK;
  if (!c.inited) return set_error(PH_ERR_NOT_INIT, "ph_comm_init was not called");
  PH_NCCL(nccl().GroupStart());
  if (lo_rank >= 0) {
    PH_NCCL_DATA(nccl().Send(send_lo, (size_t)nbytes, ncclUint8, lo_rank, c.comm, s));
    PH_NCCL_DATA(nccl().Recv(recv_lo, (size_t)nbytes, ncclUint8, lo_rank, c.comm, s));
  }
  if (hi_rank >= 0) {
    PH_NCCL_DATA(nccl().Send(send_hi, (size_t)nbytes, ncclUint8, hi_rank, c.comm, s));
    PH_NCCL_DATA(nccl().Recv(recv_hi, (size_t)nbytes, ncclUint8, hi_rank, c.comm, s));
  }
  PH_NCCL(nccl().GroupEnd());
  return PH_OK;
}

}  // namespace ph

using namespace ph;

// ---- flag rounds over the peer-mapped control blocks (strided all-to-all, ordered all-reduce)
struct PeerFlagPtrs {
  uint32_t* p[PH_MAX_PEERS];
};
static __global__ void xchg_signal_kernel(PeerFlagPtrs f, int n, uint32_t event) {
  __threadfence_system();                       // the copies launched before this kernel have completed (stream order)
  if ((int)threadIdx.x < n && f.p[threadIdx.x]) st_release_sys(f.p[threadIdx.x], event);
}
static __global__ void xchg_wait_kernel(const volatile uint32_t* flags, int n, uint32_t event, uint32_t* err_flags) {
  if ((int)threadIdx.x < n) {
    const long long t0 = clock64();
    while ((int32_t)(flags[threadIdx.x] - event) < 0) {
      if (clock64() - t0 > 20000000000LL) { atomicOr(err_flags + 1, 1u); break; }   // ~10 s: a peer died
    }
  }
  __threadfence_system();
}
static int32_t xchg_round(int which, uint32_t event) {
  Comm& c = cm();
  PeerInfo& p = peers();
  Runtime& r = rt();
  PeerFlagPtrs f;
  for (int q = 0; q < PH_MAX_PEERS; q++) f.p[q] = q < c.nranks ? &p.ctrl[q]->xchg_flag[which][c.rank] : nullptr;
  xchg_signal_kernel<<<1, 32, 0, r.stream>>>(f, c.nranks, event);
  PH_LAUNCH_CHECK("xchg_signal_kernel");
  xchg_wait_kernel<<<1, 32, 0, r.stream>>>(&p.ctrl[c.rank]->xchg_flag[which][0], c.nranks, event, r.d_flags);
  PH_LAUNCH_CHECK("xchg_wait_kernel");
  return PH_OK;
}

// ---------------------------------------------------------------- ordered all-reduce over peer memory
// ph_allreduce when every peer is mapped: reduce-scatter + all-gather as THREE launches over NVLink, folded IN RANK
// ORDER (deterministic; for an axis-0 sharded array rank order is row order, so integer sums are overflow-checked
// exactly like the single-GPU fold) with every rank's arithmetic flags delivered to every rank:
//   1. push : chunk q of my buffer -> rank q's staging slot [my rank] (peer stores); the last block to finish
//             releases "my partials have landed" in every peer's control block
//   2. fold : every block waits for that word from all N ranks, then my chunk = slot[0] (op) slot[1] (op) ... in rank
//             order -> my result chunk (double-buffered by call parity); the last block sends my flag word to every
//             peer and releases "my result is ready"
//   3. pull : every block waits for that word from all N ranks, then every rank's result chunk -> my buffer (peer
//             loads); the peers' flag words are OR-ed into mine.
// The launches carry their own flag rounds (no separate signal / wait kernels): lanes 0..n-1 of each block spin on the
// n words; the grids are small enough to be resident at once, and the signals a block waits for come from OTHER GPUs.
__device__ __forceinline__ void ar_wait_round(const uint32_t* words, int n, uint32_t event, uint32_t* err_flags) {
  if ((int)threadIdx.x < n) {
    const volatile uint32_t* w = words;
    const long long t0 = clock64();
    while ((int32_t)(w[threadIdx.x] - event) < 0) {
      if (clock64() - t0 > 20000000000LL) { atomicOr(err_flags + 1, 1u); break; }      // ~10 s: a peer died
    }
    __threadfence_system();
  }
  __syncthreads();
}
__device__ __forceinline__ void ar_copy(char* dst, const char* src, size_t n, size_t tid, size_t stride) {
  if ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
    const size_t n16 = n / 16;
    for (size_t i = tid; i < n16; i += stride) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
    for (size_t i = n16 * 16 + tid; i < n; i += stride) dst[i] = src[i];
  } else {
    for (size_t i = tid; i < n; i += stride) dst[i] = src[i];
  }
}
// true in exactly one block of the launch: the last one to get here (every block fences before it takes its
// ticket, so all the launch's earlier stores -- also those into peer memory -- are visible system-wide by then)
__device__ __forceinline__ bool ar_last_block(uint32_t* ticket) {
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    last = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1;
    if (last) { *ticket = 0; __threadfence_system(); }
  }
  __syncthreads();
  return last;
}
static __global__ void __launch_bounds__(256) ar_push_kernel(const char* __restrict__ buf, size_t total, size_t chunk,
                                                             PeerFlagPtrs stage_of /* base of each peer's staging */, size_t my_slot_offset,
                                                             uint32_t* ticket, PeerFlagPtrs landed, int nranks, uint32_t event) {
  const int q = blockIdx.y;
  const size_t lo = (size_t)q * chunk;
  if (lo < total) {
    const size_t n = (total - lo < chunk) ? total - lo : chunk;
    ar_copy(reinterpret_cast<char*>(stage_of.p[q]) + my_slot_offset, buf + lo, n,
            (size_t)blockIdx.x * blockDim.x + threadIdx.x, (size_t)gridDim.x * blockDim.x);
  }
  if (ar_last_block(ticket) && (int)threadIdx.x < nranks) st_release_sys(landed.p[threadIdx.x], event);
}
template <typename T, int RED>
static __global__ void __launch_bounds__(256) ar_fold_kernel(const char* __restrict__ staging, size_t slot_stride, int64_t n_mine,
                                                             int nranks, T* __restrict__ result, uint32_t* __restrict__ flags,
                                                             const uint32_t* landed_words, uint32_t* ticket, PeerFlagPtrs ready,
                                                             PeerFlagPtrs flag_slot, uint32_t event) {
  ar_wait_round(landed_words, nranks, event, flags);        // every rank's partials are in my staging
  uint32_t err = 0;
  bool nan = false;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_mine; i += stride) {
    T acc = reinterpret_cast<const T*>(staging)[i];
    if constexpr (is_float_t<T>::value && RED != PH_SUM) nan |= (acc != acc);
    for (int r = 1; r < nranks; r++) {
      const T v = reinterpret_cast<const T*>(staging + (size_t)r * slot_stride)[i];
      if constexpr (RED == PH_SUM) {
        if constexpr (is_float_t<T>::value) acc = f_add(acc, v);
        else acc = i_add<T>(acc, v, true, err);
      } else {
        if constexpr (is_float_t<T>::value) nan |= (v != v);
        const bool take = RED == PH_MAX ? (v > acc) : (v < acc);
        if (take) acc = v;
      }
    }
    result[i] = acc;
  }
  if (err) atomicOr(flags, err);
  if (nan) atomicOr(flags, (uint32_t)PH_FLAG_NAN);
  if (ar_last_block(ticket) && (int)threadIdx.x < nranks) {           // my result chunk is ready; my flag word goes with it
    *reinterpret_cast<volatile uint32_t*>(flag_slot.p[threadIdx.x]) = *reinterpret_cast<const volatile uint32_t*>(flags);
    __threadfence_system();
    st_release_sys(ready.p[threadIdx.x], event);
  }
}
static __global__ void __launch_bounds__(256) ar_pull_kernel(char* __restrict__ buf, size_t total, size_t chunk,
                                                             PeerFlagPtrs result_of, const uint32_t* __restrict__ peer_flags, int nranks,
                                                             uint32_t* __restrict__ flags, const uint32_t* ready_words, uint32_t event) {
  ar_wait_round(ready_words, nranks, event, flags);         // every rank's result chunk is ready
  const int q = blockIdx.y;
  const size_t lo = (size_t)q * chunk;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    uint32_t f = 0;
    for (int r = 0; r < nranks; r++) f |= reinterpret_cast<const volatile uint32_t*>(peer_flags)[r];
    if (f) atomicOr(flags, f);
  }
  if (lo >= total) return;
  const size_t n = (total - lo < chunk) ? total - lo : chunk;
  ar_copy(buf + lo, reinterpret_cast<const char*>(result_of.p[q]), n, (size_t)blockIdx.x * blockDim.x + threadIdx.x,
          (size_t)gridDim.x * blockDim.x);
}

struct ArFoldArgs {
  const char* staging; size_t slot_stride; int64_t n_mine; int nranks; void* result;
  const uint32_t* landed_words; uint32_t* ticket; PeerFlagPtrs ready, flag_slot; uint32_t event; unsigned blocks;
};
template <typename T>
static int32_t ar_fold_launch(int32_t red, const ArFoldArgs& a) {
  Runtime& r = rt();
  T* res = reinterpret_cast<T*>(a.result);
#define PH_FOLD(RED) ar_fold_kernel<T, RED><<<a.blocks, 256, 0, r.stream>>>(a.staging, a.slot_stride, a.n_mine, a.nranks, res, r.d_flags, \
                                                                             a.landed_words, a.ticket, a.ready, a.flag_slot, a.event)
  switch (red) {
    case PH_SUM: PH_FOLD(PH_SUM); break;
    case PH_MIN: PH_FOLD(PH_MIN); break;
    case PH_MAX: PH_FOLD(PH_MAX); break;
    default: return set_error(PH_ERR_UNSUPPORTED, "allreduce supports SUM / MIN / MAX (arg* use ph_allgather)");
  }
#undef PH_FOLD
  PH_LAUNCH_CHECK("ar_fold_kernel");
  return PH_OK;
}

// returns PH_OK with *done = false when the P2P form cannot take this call (the caller then uses NCCL)
static int32_t allreduce_p2p(int32_t red, int32_t dtype, void* buf_dev, int64_t count, bool* done) {
  *done = false;
  Comm& c = cm();
  PeerInfo& p = peers();
  static const bool off = getenv("PH_ALLREDUCE_NCCL") != nullptr;
  const size_t esz = (size_t)dtype_size(dtype);
  const size_t total = (size_t)count * esz;
  if (off || !p.ready || esz == 0 || total > ((size_t)256 << 20)) return PH_OK;
  if (red != PH_SUM && red != PH_MIN && red != PH_MAX) return PH_OK;
  Runtime& r = rt();
  const int n = c.nranks;
  // the user's buffer is cut into n chunks of `chunk` bytes (a multiple of 256, so vector copies and folds stay aligned)
  const int64_t per = ceil_div(count, (int64_t)n);
  const size_t chunk = (((size_t)per * esz + 255) / 256) * 256;
  if (chunk > p.ar_chunk_cap) {                       // collective growth: every rank sees the same `count`
    if (p.ar_scratch) { int32_t st = ph_symm_free(p.ar_scratch); if (st != PH_OK) return st; p.ar_scratch = nullptr; p.ar_chunk_cap = 0; }
    const size_t want = std::max<size_t>(chunk, (size_t)1 << 20);
    void* block = nullptr;
    int32_t st = ph_symm_alloc(want * (size_t)(n + 2), &block);
    if (st != PH_OK) return st;
    p.ar_scratch = reinterpret_cast<char*>(block);
    p.ar_chunk_cap = want;
  }
  const SymmAlloc* sa = symm_find(p.ar_scratch);
  if (!sa || !sa->mapped) return PH_OK;
  const size_t cap = p.ar_chunk_cap;                  // staging slots and result chunks are `cap` bytes apart in the block
  const uint32_t event = ++p.xchg_event;
  const int parity = (int)(event & 1u);
  CtrlBlock* mine = p.ctrl[c.rank];
  PeerFlagPtrs stage_of, result_of, landed, ready, flag_slot;
  for (int q = 0; q < PH_MAX_PEERS; q++) {
    const bool live = q < n;
    char* base = live ? (q == c.rank ? p.ar_scratch : sa->peer[q]) : nullptr;
    stage_of.p[q] = reinterpret_cast<uint32_t*>(base);
    result_of.p[q] = reinterpret_cast<uint32_t*>(live ? base + cap * (size_t)(n + parity) : nullptr);
    landed.p[q] = live ? &p.ctrl[q]->xchg_flag[0][c.rank] : nullptr;
    ready.p[q] = live ? &p.ctrl[q]->xchg_flag[1][c.rank] : nullptr;
    flag_slot.p[q] = live ? &p.ctrl[q]->ar_flags[c.rank] : nullptr;
  }
  const unsigned bx = (unsigned)std::max<size_t>(1, std::min<size_t>(32, chunk / ((size_t)256 * 64) + 1));
  ar_push_kernel<<<dim3(bx, (unsigned)n), 256, 0, r.stream>>>(reinterpret_cast<const char*>(buf_dev), total, chunk, stage_of,
                                                              cap * (size_t)c.rank, &mine->ar_ticket[0], landed, n, event);
  PH_LAUNCH_CHECK("ar_push_kernel");
  const size_t my_lo = (size_t)c.rank * chunk;
  ArFoldArgs fa;
  fa.staging = p.ar_scratch; fa.slot_stride = cap; fa.nranks = n;
  fa.n_mine = my_lo >= total ? 0 : (int64_t)(std::min(chunk, total - my_lo) / esz);
  fa.result = p.ar_scratch + cap * (size_t)(n + parity);
  fa.landed_words = &mine->xchg_flag[0][0]; fa.ticket = &mine->ar_ticket[1]; fa.ready = ready; fa.flag_slot = flag_slot; fa.event = event;
  fa.blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(4 * r.sm_count, ceil_div(std::max<int64_t>(fa.n_mine, 1), 256 * 4)));
  int32_t st;
  switch (dtype) {
#define PH_AR(code, T) case code: st = ar_fold_launch<T>(red, fa); break;
    PH_AR(PH_F32, float) PH_AR(PH_F64, double) PH_AR(PH_I32, int32_t) PH_AR(PH_I64, int64_t) PH_AR(PH_U8, uint8_t)
    PH_AR(PH_I8, int8_t) PH_AR(PH_I16, int16_t) PH_AR(PH_U16, uint16_t) PH_AR(PH_U32, uint32_t) PH_AR(PH_U64, uint64_t)
#undef PH_AR
    default: return set_error(PH_ERR_UNSUPPORTED, "dtype %d has no all-reduce", dtype);
  }
  if (st != PH_OK) return st;
  ar_pull_kernel<<<dim3(bx, (unsigned)n), 256, 0, r.stream>>>(reinterpret_cast<char*>(buf_dev), total, chunk, result_of,
                                                              &mine->ar_flags[0], n, r.d_flags, &mine->xchg_flag[1][0], event);
  PH_LAUNCH_CHECK("ar_pull_kernel");
  *done = true;
  return PH_OK;
}

extern "C" {

int32_t ph_comm_unique_id(uint8_t* out128) {
  if (!out128) return set_error(PH_ERR_INVALID, "null out128");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  PH_NCCL_LOAD();
  ncclUniqueId id;
  PH_NCCL(nccl().GetUniqueId(&id));
  memcpy(out128, &id, 128);
  return PH_OK;
}

int32_t ph_comm_init(int32_t nranks, int32_t rank, const uint8_t* id128) {
  PH_REQUIRE_INIT();
  Comm& c = cm();
  if (c.inited) ph_comm_destroy();
  else {                        // single-process record-mode reductions may have run: call numbers restart at 0 on every rank
    cudaDeviceSynchronize();
    PeerInfo& p = peers();
    p.reduce_seq = 0;
    if (p.host_result) memset(p.host_result, 0, 64);
  }
  if (nranks < 1 || rank < 0 || rank >= nranks) return set_error(PH_ERR_INVALID, "bad rank %d of %d", rank, nranks);
  c.nranks = nranks; c.rank = rank;
  if (nranks == 1) { c.inited = true; c.comm = nullptr; return p2p_setup(); }
  if (!id128) return set_error(PH_ERR_INVALID, "null unique id");
  PH_NCCL_LOAD();
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  PH_NCCL(nccl().CommInitRank(&c.comm, nranks, id, rank));
  c.inited = true;
  return p2p_setup();     // map the peers' control blocks (CUDA IPC); the NCCL paths remain when that is impossible
}

int64_t ph_nccl_call_count(void) { return g_nccl_calls; }

int32_t ph_comm_p2p_ready(int32_t* out) {
  if (!out) return set_error(PH_ERR_INVALID, "null out");
  *out = peers().ready ? 1 : 0;
  return PH_OK;
}

int32_t ph_symm_alloc(size_t nbytes, void** out_dev) {
  PH_REQUIRE_INIT();
  Comm& c = cm();
  if (!c.inited) return set_error(PH_ERR_NOT_INIT, "ph_comm_init was not called");
  if (!out_dev) return set_error(PH_ERR_INVALID, "null out_dev");
  *out_dev = nullptr;
  SymmAlloc a;
  a.nbytes = nbytes ? nbytes : 1;
  PH_CUDA(cudaMalloc((void**)&a.local, a.nbytes));
  if (peers().ready) {
    int32_t st = map_peers(a.local, a.peer, &a.mapped);
    if (st != PH_OK) { cudaFree(a.local); return st; }
  }
  symm_table().push_back(a);
  *out_dev = a.local;
  return PH_OK;
}

int32_t ph_symm_free(void* dev) {
  PH_REQUIRE_INIT();
  if (!dev) return PH_OK;
  std::vector<SymmAlloc>& t = symm_table();
  for (size_t i = 0; i < t.size(); i++) {
    if (t[i].local != dev) continue;
    PH_CUDA(cudaDeviceSynchronize());
    if (t[i].mapped) {
      bool dummy;
      int32_t st = all_agree(true, &dummy);        // nobody unmaps while a peer may still be writing
      if (st != PH_OK) return st;
      for (int r = 0; r < cm().nranks; r++)
        if (t[i].peer[r] && t[i].peer[r] != t[i].local) cudaIpcCloseMemHandle(t[i].peer[r]);
      if ((st = all_agree(true, &dummy)) != PH_OK) return st;     // every importer has let go before the owner frees
    }
    PH_CUDA(cudaFree(t[i].local));
    t.erase(t.begin() + (long)i);
    return PH_OK;
  }
  return set_error(PH_ERR_INVALID, "ph_symm_free: not a ph_symm_alloc pointer");
}

int32_t ph_symm_peer(const void* local_dev, int32_t peer_rank, void** out_peer_dev) {
  if (!out_peer_dev) return set_error(PH_ERR_INVALID, "null out_peer_dev");
  *out_peer_dev = nullptr;
  const SymmAlloc* a = symm_find(local_dev);
  if (!a || !a->mapped || peer_rank < 0 || peer_rank >= cm().nranks) return PH_OK;
  *out_peer_dev = a->peer[peer_rank] + (reinterpret_cast<const char*>(local_dev) - a->local);
  return PH_OK;
}

int32_t ph_comm_destroy(void) {
  Comm& c = cm();
  if (c.inited) {
    cudaDeviceSynchronize();
    p2p_teardown();
    if (c.comm) nccl().CommDestroy(c.comm);
  }
  c = Comm();
  return PH_OK;
}

int32_t ph_allreduce(int32_t red, int32_t dtype, void* buf_dev, int64_t count) {
  PH_REQUIRE_INIT();
  Comm& c = cm();
  if (!c.inited) return set_error(PH_ERR_NOT_INIT, "ph_comm_init was not called");
  if (c.nranks == 1 || count == 0) return PH_OK;
  {
    bool done = false;                                  // peers mapped: reduce-scatter + all-gather over peer memory, rank order
    int32_t st = allreduce_p2p(red, dtype, buf_dev, count, &done);
    if (st != PH_OK || done) return st;
  }
  ncclDataType_t t;
  int32_t st = nccl_type(dtype, &t);
  if (st != PH_OK) return st;
  ncclRedOp_t op;
  switch (red) {
    case PH_SUM: op = ncclSum; break;
    case PH_MIN: op = ncclMin; break;
    case PH_MAX: op = ncclMax; break;
    default: return set_error(PH_ERR_UNSUPPORTED, "allreduce supports SUM / MIN / MAX (arg* use ph_allgather)");
  }
  PH_NCCL_DATA(nccl().AllReduce(buf_dev, buf_dev, (size_t)count, t, op, c.comm, rt().stream));
  return PH_OK;
}

int32_t ph_allgather(const void* send_dev, void* recv_dev, int64_t nbytes_per_rank) {
  PH_REQUIRE_INIT();
  Comm& c = cm();
  if (!c.inited) return set_error(PH_ERR_NOT_INIT, "ph_comm_init was not called");
  if (c.nranks == 1) {
    if (send_dev != recv_dev)
      PH_CUDA(cudaMemcpyAsync(recv_dev, send_dev, (size_t)nbytes_per_rank, cudaMemcpyDeviceToDevice, rt().stream));
    return PH_OK;
  }
  PH_NCCL_DATA(nccl().AllGather(send_dev, recv_dev, (size_t)nbytes_per_rank, ncclUint8, c.comm, rt().stream));
  return PH_OK;
}

// Personalised all-to-all: block p of the send list goes to rank p, block p of the receive list
// comes from rank p (byte counts per peer; 0 = nothing).  One grouped ncclSend/ncclRecv per peer
// pair, the rank's own block is a device copy.  This is the exchange step of a transpose across
// axis-0 shards (ShardedNArray.permute), the one place on the path where data must change owner.
int32_t ph_alltoallv(const void* const* send_dev, const int64_t* send_bytes, void* const* recv_dev,
                     const int64_t* recv_bytes) {
  PH_REQUIRE_INIT();
  Comm& c = cm();
  if (!c.inited) return set_error(PH_ERR_NOT_INIT, "ph_comm_init was not called");
  if (!send_dev || !send_bytes || !recv_dev || !recv_bytes) return set_error(PH_ERR_INVALID, "null argument to ph_alltoallv");
  cudaStream_t s = rt().stream;
  const int me = c.rank;
  if (send_bytes[me] != recv_bytes[me]) return set_error(PH_ERR_INVALID, "ph_alltoallv: own block sizes differ");
  if (send_bytes[me] > 0 && send_dev[me] != recv_dev[me])
    PH_CUDA(cudaMemcpyAsync(recv_dev[me], send_dev[me], (size_t)send_bytes[me], cudaMemcpyDeviceToDevice, s));
  if (c.nranks == 1) return PH_OK;
  PH_NCCL(nccl().GroupStart());
  for (int p = 0; p < c.nranks; p++) {
    if (p == me) continue;
    if (send_bytes[p] > 0) PH_NCCL_DATA(nccl().Send(send_dev[p], (size_t)send_bytes[p], ncclUint8, p, c.comm, s));
    if (recv_bytes[p] > 0) PH_NCCL_DATA(nccl().Recv(recv_dev[p], (size_t)recv_bytes[p], ncclUint8, p, c.comm, s));
  }
  PH_NCCL(nccl().GroupEnd());
  return PH_OK;
}

// Personalised all-to-all of STRIDED blocks as peer stores: the exchange step of a transpose across axis-0
// shards in ONE pass -- for every peer q the block this rank owes it (src_descs[q]: a view of the rank's rows,
// already in the destination's axis order) is copied by the ordinary gather / transpose kernels straight into
// rank q's result through the peer mapping (dst_descs[q], relative to q's copy of the symmetric block): the
// permuting copy IS the transfer.  No staging buffer, no ncclSend/ncclRecv, no scatter afterwards.  Peers are
// visited in the order rank+1, rank+2, ... so every phase is a permutation (no receiver takes N streams at once).
// Two flag rounds bracket the copies: "my destination may be overwritten" before, "my stores have landed" after
// (stream-ordered; the host does not block).
int32_t ph_alltoall_strided(int32_t elem_size, const void* src_dev, const ph_desc* src_descs, void* dst_symm,
                            const ph_desc* dst_descs) {
  PH_REQUIRE_INIT();
  Comm& c = cm();
  PeerInfo& p = peers();
  if (!c.inited) return set_error(PH_ERR_NOT_INIT, "ph_comm_init was not called");
  if (!src_dev || !src_descs || !dst_symm || !dst_descs) return set_error(PH_ERR_INVALID, "null argument to ph_alltoall_strided");
  auto elems = [](const ph_desc& d) { int64_t n = 1; for (int i = 0; i < d.rank; i++) n *= d.extent[i]; return n; };
  if (c.nranks == 1) return elems(src_descs[0]) ? ph_copy_strided(elem_size, src_dev, &src_descs[0], dst_symm, &dst_descs[0]) : PH_OK;
  if (!p.ready) return set_error(PH_ERR_UNSUPPORTED, "peers are not mapped (no P2P): use ph_alltoallv");
  const SymmAlloc* sa = symm_find(dst_symm);
  if (!sa || !sa->mapped) return set_error(PH_ERR_INVALID, "ph_alltoall_strided: the destination must come from ph_symm_alloc");
  const ptrdiff_t rel = reinterpret_cast<const char*>(dst_symm) - sa->local;
  const uint32_t event = ++p.xchg_event;
  int32_t st = xchg_round(0, event);            // every rank has finished whatever read its destination before
  if (st != PH_OK) return st;
  // The peer stores go out on the (high-priority) auxiliary stream: they are bound by the link, need few SMs
  // and must start first; my own block, which never leaves the GPU and is HBM-bound, is copied on the main
  // stream at the same time.
  Runtime& r = rt();
  cudaStream_t main_stream = r.stream;
  PH_CUDA(cudaEventRecord(r.ev_a, main_stream));
  PH_CUDA(cudaStreamWaitEvent(r.aux_stream, r.ev_a, 0));
  r.stream = r.aux_stream;
  for (int i = 1; i < c.nranks && st == PH_OK; i++) {
    const int q = (c.rank + i) % c.nranks;
    if (elems(src_descs[q]) == 0) continue;
    st = ph_copy_strided(elem_size, src_dev, &src_descs[q], (void*)(sa->peer[q] + rel), &dst_descs[q]);
  }
  r.stream = main_stream;
  if (st != PH_OK) return st;
  PH_CUDA(cudaEventRecord(r.ev_b, r.aux_stream));
  if (elems(src_descs[c.rank]) != 0 &&
      (st = ph_copy_strided(elem_size, src_dev, &src_descs[c.rank], dst_symm, &dst_descs[c.rank])) != PH_OK)
    return st;
  PH_CUDA(cudaStreamWaitEvent(main_stream, r.ev_b, 0));
  return xchg_round(1, event);                  // every block destined for me has landed
}

int32_t ph_halo_exchange(const void* send_lo, void* recv_lo, int32_t lo_rank, const void* send_hi,
                         void* recv_hi, int32_t hi_rank, int64_t nbytes, void* cuda_stream) {
  PH_REQUIRE_INIT();
  cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : rt().stream;
  return halo_exchange_impl(send_lo, recv_lo, lo_rank, send_hi, recv_hi, hi_rank, nbytes, s);
}

// local_extents[0] = n0_local + 2 g, g = ghost_planes per side (1 or 2).  Per pass:
//   side stream (high priority): update the g edge planes on either side (what the neighbours
//                need), then exchange them with the neighbours' ghost planes;
//   main stream: update the interior planes at the same time;
//   each stream waits for the other's previous pass (events), see the loop below.
// g = 1: every pass is one time step.  g = 2 (rank 3, shape permitting): a pass advances TWO time
// steps with the temporally blocked kernel (heat_tma.cu), so both the HBM traffic and the number
// of exchanges per time step halve; an odd last step is a single step.
int32_t ph_heat_run_sharded(int32_t dtype, int32_t rank, const int64_t* local_extents, const void* coeff_host,
                            int32_t ghost_planes, void* buf_a, void* buf_b, int64_t steps, int32_t* final_is_b) {
  PH_REQUIRE_INIT();
  Runtime& r = rt();
  Comm& c = cm();
  if (!c.inited) return set_error(PH_ERR_NOT_INIT, "ph_comm_init was not called");
  if (!local_extents || !coeff_host || !buf_a || !buf_b) return set_error(PH_ERR_INVALID, "null argument");
  if (rank < 2 || rank > 3) return set_error(PH_ERR_UNSUPPORTED, "sharded stencil needs rank 2 or 3");
  const int esz = dtype_size(dtype);
  if (dtype != PH_F32 && dtype != PH_F64) return set_error(PH_ERR_UNSUPPORTED, "heat stencil is F32 / F64");
  const int g = ghost_planes;
  if (g != 1 && g != 2) return set_error(PH_ERR_INVALID, "ghost_planes must be 1 or 2 (got %d)", g);
  const int64_t n0 = local_extents[0];
  if (n0 - 2 * g < g) return set_error(PH_ERR_INVALID, "a slab needs at least %d owned planes", g);
  int64_t plane = 1;
  for (int i = 1; i < rank; i++) plane *= local_extents[i];
  const int64_t pbytes = plane * esz;
  const int lo = c.rank > 0 ? c.rank - 1 : -1;
  const int hi = c.rank < c.nranks - 1 ? c.rank + 1 : -1;
  const int has_lo = lo >= 0, has_hi = hi >= 0;
  char* bufs[2] = {reinterpret_cast<char*>(buf_a), reinterpret_cast<char*>(buf_b)};
  int32_t dummy = 0;
  if (!final_is_b) final_is_b = &dummy;
  *final_is_b = 0;
  const bool can_two = g == 2 && heat_two_step_usable(dtype, rank, local_extents);

  const int64_t own_b = g, own_e = n0 - g;

  // ---- P2P path: both slabs live in peer-mapped memory (ph_symm_alloc) -> no NCCL, no side stream.
  // A pass is ONE launch over all owned planes: the marches at either end of the slab are dispatched first
  // and store the planes the neighbours need straight into their ghost planes (compute + halo in one
  // kernel); the last of those blocks releases the neighbour's flag word, and the neighbour's next pass is
  // stream-ordered behind a wait on it.  Events are numbered identically on every rank (same calls, same
  // step counts): hello, initial ghosts, then one per pass.
  {
    PeerInfo& P = peers();
    static const bool force_nccl = getenv("PH_HEAT_NCCL") != nullptr;           // A/B knob
    void *pa_lo = nullptr, *pb_lo = nullptr, *pa_hi = nullptr, *pb_hi = nullptr;
    bool p2p = P.ready && !force_nccl && c.nranks > 1 && (g * pbytes) % 16 == 0 && ((uintptr_t)buf_a % 16) == 0 &&
               ((uintptr_t)buf_b % 16) == 0;
    if (p2p) {
      if (has_lo) { ph_symm_peer(buf_a, lo, &pa_lo); ph_symm_peer(buf_b, lo, &pb_lo); p2p = p2p && pa_lo && pb_lo; }
      if (has_hi) { ph_symm_peer(buf_a, hi, &pa_hi); ph_symm_peer(buf_b, hi, &pb_hi); p2p = p2p && pa_hi && pb_hi; }
      if (!symm_find(buf_a) || !symm_find(buf_b)) p2p = false;
    }
    if (p2p) {
      cudaStream_t s = r.stream;
      CtrlBlock* me = P.ctrl[c.rank];
      uint32_t* nbr_flag_lo = has_lo ? &P.ctrl[lo]->halo_flag[1] : nullptr;    // I am lo's HI neighbour
      uint32_t* nbr_flag_hi = has_hi ? &P.ctrl[hi]->halo_flag[0] : nullptr;
      auto wait_nbrs = [&](uint32_t ev) -> int32_t {
        int32_t w = PH_OK;
        if (has_lo && (w = stream_wait_geq(s, &me->halo_flag[0], ev)) != PH_OK) return w;
        if (has_hi && (w = stream_wait_geq(s, &me->halo_flag[1], ev)) != PH_OK) return w;
        return PH_OK;
      };
      // hello: my earlier work is complete (stream order) and this is my plane count
      uint32_t ev = ++P.halo_event;
      halo_hello_kernel<<<1, 1, 0, s>>>(nbr_flag_lo, nbr_flag_hi, has_lo ? &P.ctrl[lo]->nbr_planes[1] : nullptr,
                                        has_hi ? &P.ctrl[hi]->nbr_planes[0] : nullptr, n0, ev);
      PH_LAUNCH_CHECK("halo_hello_kernel");
      int32_t st = wait_nbrs(ev);
      if (st != PH_OK) return st;
      int64_t nbr_planes[2] = {0, 0};
      PH_CUDA(cudaMemcpyAsync(nbr_planes, me->nbr_planes, sizeof(nbr_planes), cudaMemcpyDeviceToHost, s));
      PH_CUDA(cudaStreamSynchronize(s));
      const int64_t n0_lo = nbr_planes[0];
      char* peer_bufs_lo[2] = {reinterpret_cast<char*>(pa_lo), reinterpret_cast<char*>(pb_lo)};
      char* peer_bufs_hi[2] = {reinterpret_cast<char*>(pa_hi), reinterpret_cast<char*>(pb_hi)};
      // my g lowest owned planes -> lo's upper ghosts [n0_lo - g, n0_lo); my g highest -> hi's lower ghosts [0, g)
      auto push = [&](int which, uint32_t event) -> int32_t {
        const int64_t n16 = g * pbytes / 16;
        const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)r.sm_count * 4, ceil_div(n16, (int64_t)256 * 4)));
        halo_push_kernel<<<blocks, 256, 0, s>>>(
            reinterpret_cast<const uint4*>(bufs[which] + own_b * pbytes),
            has_lo ? reinterpret_cast<uint4*>(peer_bufs_lo[which] + (n0_lo - g) * pbytes) : nullptr,
            reinterpret_cast<const uint4*>(bufs[which] + (own_e - g) * pbytes),
            has_hi ? reinterpret_cast<uint4*>(peer_bufs_hi[which]) : nullptr, n16, &me->halo_ticket[2], nbr_flag_lo,
            nbr_flag_hi, event);
        PH_LAUNCH_CHECK("halo_push_kernel");
        return PH_OK;
      };
      ev = ++P.halo_event;
      if ((st = push(0, ev)) != PH_OK) return st;
      if (!has_lo) PH_CUDA(cudaMemcpyAsync(bufs[1] + g * pbytes, bufs[0] + g * pbytes, pbytes, cudaMemcpyDeviceToDevice, s));
      if (!has_hi) PH_CUDA(cudaMemcpyAsync(bufs[1] + (n0 - g - 1) * pbytes, bufs[0] + (n0 - g - 1) * pbytes, pbytes,
                                           cudaMemcpyDeviceToDevice, s));
      int cur = 0;
      int64_t left = steps;
      while (left > 0) {
        const bool two = can_two && left >= 2;
        if ((st = wait_nbrs(ev)) != PH_OK) return st;        // my ghosts of `in` are complete; theirs of `out` are free
        ev = ++P.halo_event;
        const char* in = bufs[cur];
        char* out = bufs[cur ^ 1];
        HeatMirror mir;
        if (has_lo) {
          mir.lo_end = own_b + g;
          mir.delta_lo = (peer_bufs_lo[cur ^ 1] - out) + (n0_lo - g - own_b) * pbytes;
          mir.flag_lo = nbr_flag_lo;
        }
        if (has_hi) {
          mir.hi_begin = own_e - g;
          mir.delta_hi = (peer_bufs_hi[cur ^ 1] - out) - (own_e - g) * pbytes;
          mir.flag_hi = nbr_flag_hi;
        }
        mir.ticket = me->halo_ticket;
        mir.event = ev;
        bool mirrored = false;
        st = heat_slab_dispatch(dtype, rank, local_extents, coeff_host, g, has_lo, has_hi, own_b, own_e, in, out, s, two,
                                two ? &mir : nullptr, &mirrored);
        if (st != PH_OK) return st;
        if (!mirrored && (st = push(cur ^ 1, ev)) != PH_OK) return st;
        cur ^= 1;
        left -= two ? 2 : 1;
      }
      if ((st = wait_nbrs(ev)) != PH_OK) return st;          // the ghosts of the final state have arrived
      *final_is_b = cur;
      return PH_OK;
    }
  }

  // ---- NCCL path
  // g lowest owned planes -> lo neighbour's upper ghosts, g highest owned planes -> hi neighbour's lower ghosts
  auto exchange = [&](char* buf, cudaStream_t s) {
    return halo_exchange_impl(buf + g * pbytes, buf, lo, buf + (n0 - 2 * g) * pbytes, buf + (n0 - g) * pbytes, hi,
                              g * pbytes, s);
  };
  // initial ghosts of buf_a, and the (constant) global-boundary planes of buf_b
  int32_t st = exchange(bufs[0], r.stream);
  if (st != PH_OK) return st;
  if (!has_lo) PH_CUDA(cudaMemcpyAsync(bufs[1] + g * pbytes, bufs[0] + g * pbytes, pbytes, cudaMemcpyDeviceToDevice, r.stream));
  if (!has_hi) PH_CUDA(cudaMemcpyAsync(bufs[1] + (n0 - g - 1) * pbytes, bufs[0] + (n0 - g - 1) * pbytes, pbytes,
                                       cudaMemcpyDeviceToDevice, r.stream));

  // Two streams per pass.  side (high priority): the g edge planes on either side, then the grouped
  // send/recv of exactly those planes.  main: the interior planes, CONCURRENTLY with the edges (both
  // read `in` and write disjoint planes of `out`), so the small, latency-bound edge launches never
  // hold the interior back and the exchange runs under it.  Dependencies of pass k:
  //   edges(k)    <- exchange(k-1) [same stream], interior(k-1) [ev_a: main -> side]
  //   interior(k) <- edges(k-1)                  [ev_b: side -> main]
  static const bool no_overlap = getenv("PH_HEAT_NO_OVERLAP") != nullptr;    // measurement knob
  const bool split = own_e - own_b > 2 * g && c.nranks > 1 && !no_overlap;
  cudaStream_t side = r.aux_stream;
  if (split) {
    PH_CUDA(cudaEventRecord(r.ev_a, r.stream));
    PH_CUDA(cudaStreamWaitEvent(side, r.ev_a, 0));             // initial ghosts / boundary copies are on main
    PH_CUDA(cudaEventRecord(r.ev_b, side));
  }
  int cur = 0;
  int64_t left = steps;
  while (left > 0) {
    const bool two = can_two && left >= 2;
    const char* in = bufs[cur];
    char* out = bufs[cur ^ 1];
    auto update = [&](int64_t b, int64_t e, cudaStream_t s) {
      return heat_slab_dispatch(dtype, rank, local_extents, coeff_host, g, has_lo, has_hi, b, e, in, out, s, two, nullptr, nullptr);
    };
    if (split) {
      PH_CUDA(cudaStreamWaitEvent(side, r.ev_a, 0));           // previous interior
      PH_CUDA(cudaStreamWaitEvent(r.stream, r.ev_b, 0));       // previous edges + exchange
      if ((st = update(own_b, own_b + g, side)) != PH_OK) return st;
      if ((st = update(own_e - g, own_e, side)) != PH_OK) return st;
      if ((st = exchange(out, side)) != PH_OK) return st;
      PH_CUDA(cudaEventRecord(r.ev_b, side));
      if ((st = update(own_b + g, own_e - g, r.stream)) != PH_OK) return st;
      PH_CUDA(cudaEventRecord(r.ev_a, r.stream));
    } else {
      if ((st = update(own_b, own_e, r.stream)) != PH_OK) return st;
      if (c.nranks > 1 && (st = exchange(out, r.stream)) != PH_OK) return st;
    }
    cur ^= 1;
    left -= two ? 2 : 1;
  }
  if (split) PH_CUDA(cudaStreamWaitEvent(r.stream, r.ev_b, 0));   // the caller's stream sees the whole run
  *final_is_b = cur;
  return PH_OK;
}

}  // extern "C"
