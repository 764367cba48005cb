// comm.cu -- multi-GPU plumbing: one process per GPU, NCCL over NVLink / NVSwitch.
//
// The reference has no distributed layer at all; this file adds only what
// BASELINE.json's north_star partitions: reductions combine per-GPU partials with an
// allreduce, and the heat grid is slab-decomposed along axis 0 with one-plane halos
// exchanged by grouped ncclSend/ncclRecv on a side stream while the interior planes are
// being updated on the main stream.
#include "ph_common.cuh"
#include <stdlib.h>
#include <dlfcn.h>
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

static int32_t check_nccl(ncclResult_t r, const char* what) {
  if (r == ncclSuccess) return PH_OK;
  return set_error(PH_ERR_NCCL, "NCCL error %d (%s) at %s", (int)r, nccl().GetErrorString(r), what);
}
#define PH_NCCL(call)                                   \
  do {                                                  \
    int32_t _s = check_nccl((call), #call);             \
    if (_s != PH_OK) return _s;                         \
  } while (0)

int32_t heat_slab_dispatch(int32_t dtype, int rank, const int64_t* ext, const void* coeff_host, int ghost, int has_lo,
                           int has_hi, int64_t p_begin, int64_t p_end, const void* in, void* out,
                           cudaStream_t stream, bool two_step);
bool heat_two_step_usable(int32_t dtype, int rank, const int64_t* ext);

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
    PH_NCCL(nccl().Send(send_lo, (size_t)nbytes, ncclUint8, lo_rank, c.comm, s));
    PH_NCCL(nccl().Recv(recv_lo, (size_t)nbytes, ncclUint8, lo_rank, c.comm, s));
  }
  if (hi_rank >= 0) {
    PH_NCCL(nccl().Send(send_hi, (size_t)nbytes, ncclUint8, hi_rank, c.comm, s));
    PH_NCCL(nccl().Recv(recv_hi, (size_t)nbytes, ncclUint8, hi_rank, c.comm, s));
  }
  PH_NCCL(nccl().GroupEnd());
  return PH_OK;
}

}  // namespace ph

using namespace ph;

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
  if (nranks < 1 || rank < 0 || rank >= nranks) return set_error(PH_ERR_INVALID, "bad rank %d of %d", rank, nranks);
  c.nranks = nranks; c.rank = rank;
  if (nranks == 1) { c.inited = true; c.comm = nullptr; return PH_OK; }
  if (!id128) return set_error(PH_ERR_INVALID, "null unique id");
  PH_NCCL_LOAD();
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  PH_NCCL(nccl().CommInitRank(&c.comm, nranks, id, rank));
  c.inited = true;
  return PH_OK;
}

int32_t ph_comm_destroy(void) {
  Comm& c = cm();
  if (c.inited && c.comm) {
    cudaDeviceSynchronize();
    nccl().CommDestroy(c.comm);
  }
  c = Comm();
  return PH_OK;
}

int32_t ph_allreduce(int32_t red, int32_t dtype, void* buf_dev, int64_t count) {
  PH_REQUIRE_INIT();
  Comm& c = cm();
  if (!c.inited) return set_error(PH_ERR_NOT_INIT, "ph_comm_init was not called");
  if (c.nranks == 1 || count == 0) return PH_OK;
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
  PH_NCCL(nccl().AllReduce(buf_dev, buf_dev, (size_t)count, t, op, c.comm, rt().stream));
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
  PH_NCCL(nccl().AllGather(send_dev, recv_dev, (size_t)nbytes_per_rank, ncclUint8, c.comm, rt().stream));
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
    if (send_bytes[p] > 0) PH_NCCL(nccl().Send(send_dev[p], (size_t)send_bytes[p], ncclUint8, p, c.comm, s));
    if (recv_bytes[p] > 0) PH_NCCL(nccl().Recv(recv_dev[p], (size_t)recv_bytes[p], ncclUint8, p, c.comm, s));
  }
  PH_NCCL(nccl().GroupEnd());
  return PH_OK;
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
  const int64_t own_b = g, own_e = n0 - g;
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
      return heat_slab_dispatch(dtype, rank, local_extents, coeff_host, g, has_lo, has_hi, b, e, in, out, s, two);
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
