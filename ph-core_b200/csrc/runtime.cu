// runtime.cu -- device selection, stream, pool allocation, transfers, flags, timers,
// and the descriptor planner shared by every kernel family.
// Replaces the storage half of NArray (src/n_array.cr:20-79, 230-232, 372-395).
#include "ph_common.cuh"
#include <atomic>
#include <stdarg.h>
#include <stdio.h>
#include <algorithm>
#include <vector>

namespace ph {

Runtime& rt() {
  static Runtime r;
  return r;
}

int32_t set_error(int32_t code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(rt().err, sizeof(rt().err), fmt, ap);
  va_end(ap);
  return code;
}

int32_t check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return PH_OK;
  return set_error(PH_ERR_CUDA, "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
}

int32_t ensure_scratch(size_t bytes) {
  Runtime& r = rt();
  if (r.scratch_bytes >= bytes) return PH_OK;
  if (r.d_scratch) {
    PH_CUDA(cudaStreamSynchronize(r.stream));
    PH_CUDA(cudaFree(r.d_scratch));
    r.d_scratch = nullptr;
    r.scratch_bytes = 0;
  }
  size_t want = std::max(bytes, (size_t)1 << 20);
  PH_CUDA(cudaMalloc(&r.d_scratch, want));
  r.scratch_bytes = want;
  return PH_OK;
}

// Normalise N descriptors over shared extents: drop size-1 axes,
// merge axis j into j+1 when stride[j] == stride[j+1] * extent[j+1] for EVERY operand.
int32_t make_plan(Plan& p, int nops, const ph_desc* const* descs) {
  if (nops < 1 || nops > MAX_OPERANDS) return set_error(PH_ERR_INVALID, "bad operand count %d", nops);
  for (int k = 0; k < nops; k++) {
    if (!descs[k]) return set_error(PH_ERR_INVALID, "null descriptor (operand %d)", k);
    if (descs[k]->rank < 0 || descs[k]->rank > PH_MAX_RANK)
      return set_error(PH_ERR_INVALID, "descriptor rank %d out of range (operand %d)", descs[k]->rank, k);
    if (descs[k]->rank != descs[0]->rank)
      return set_error(PH_ERR_INVALID, "descriptor ranks differ (%d vs %d)", descs[k]->rank, descs[0]->rank);
    for (int i = 0; i < descs[0]->rank; i++)
      if (descs[k]->extent[i] != descs[0]->extent[i])
        return set_error(PH_ERR_INVALID, "descriptor extents differ on axis %d", i);
  }
  p.nops = nops;
  p.rank = 0;
  p.total = 1;
  for (int k = 0; k < nops; k++) p.offset[k] = descs[k]->offset;
  const int rank = descs[0]->rank;
  for (int i = 0; i < rank; i++) {
    int64_t e = descs[0]->extent[i];
    if (e < 0) return set_error(PH_ERR_INVALID, "negative extent on axis %d", i);
    p.total *= e;
  }
  if (rank == 0) p.total = 1;
  if (p.total == 0) return PH_OK;
  for (int i = 0; i < rank; i++) {
    int64_t e = descs[0]->extent[i];
    if (e == 1) continue;
    bool merged = false;
    if (p.rank > 0) {
      int j = p.rank - 1;
      bool ok = true;
      for (int k = 0; k < nops; k++)
        if (p.stride[k][j] != descs[k]->stride[i] * e) { ok = false; break; }
      if (ok) {
        p.extent[j] *= e;
        for (int k = 0; k < nops; k++) p.stride[k][j] = descs[k]->stride[i];
        merged = true;
      }
    }
    if (!merged) {
      p.extent[p.rank] = e;
      for (int k = 0; k < nops; k++) p.stride[k][p.rank] = descs[k]->stride[i];
      p.rank++;
    }
  }
  return PH_OK;
}

// ---- position-weighted 64-bit checksum of a device buffer (verification aid of bench.py / tests)
// sum over 8-byte words w_i of w_i * (2 * (i + word_offset) + 1)  (mod 2^64): wrapping adds commute, so
// the value does not depend on the grid or on how an array is sharded over ranks (a rank passes the
// GLOBAL word index of its first word and the per-rank values are added), but it does depend on WHERE
// every word sits -- a misplaced plane or a stale halo changes it.
static __global__ void __launch_bounds__(256) checksum64_kernel(const uint64_t* __restrict__ x, int64_t nwords,
                                                                uint64_t word_offset, unsigned long long* __restrict__ out) {
  uint64_t acc = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < nwords; i += stride) {
    if (i + 4 <= nwords) {
      const RawVec<32> v = ld_stream<32>(x + i);
#pragma unroll
      for (int k = 0; k < 4; k++) acc += v.q[k] * (2 * ((uint64_t)(i + k) + word_offset) + 1);
    } else {
      for (int64_t k = i; k < nwords; k++) acc += x[k] * (2 * ((uint64_t)k + word_offset) + 1);
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
  __shared__ uint64_t sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t t = 0;
    for (int w = 0; w < 8; w++) t += sh[w];
    atomicAdd(out, (unsigned long long)t);
  }
}

}  // namespace ph

using namespace ph;

extern "C" {

int32_t ph_checksum64(const void* dev, size_t nbytes, uint64_t word_offset, uint64_t* out_host) {
  PH_REQUIRE_INIT();
  if (!out_host) return set_error(PH_ERR_INVALID, "null out_host");
  if (nbytes % 8 || ((uintptr_t)dev % 32)) return set_error(PH_ERR_INVALID, "ph_checksum64 needs a 32-byte aligned buffer of whole 8-byte words");
  Runtime& r = rt();
  int32_t st = ensure_scratch(1 << 20);
  if (st != PH_OK) return st;
  unsigned long long* acc = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(r.d_scratch) + (1 << 20) - 128);
  PH_CUDA(cudaMemsetAsync(acc, 0, 8, r.stream));
  const int64_t nwords = (int64_t)(nbytes / 8);
  if (nwords) {
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)r.sm_count * 8, ceil_div(nwords, (int64_t)256 * 4)));
    checksum64_kernel<<<blocks, 256, 0, r.stream>>>(reinterpret_cast<const uint64_t*>(dev), nwords, word_offset, acc);
    PH_LAUNCH_CHECK("checksum64_kernel");
  }
  PH_CUDA(cudaMemcpyAsync(r.h_scratch, acc, 8, cudaMemcpyDeviceToHost, r.stream));
  PH_CUDA(cudaStreamSynchronize(r.stream));
  memcpy(out_host, r.h_scratch, 8);
  return PH_OK;
}

// ---- caller-visible streams: chunked host <-> device pipelines through the array API
static std::vector<cudaStream_t>& live_streams() {
  static std::vector<cudaStream_t> v;
  return v;
}
static std::vector<cudaStream_t>& adopted_streams() {      // caller-owned streams seen by ph_set_stream
  static std::vector<cudaStream_t> v;
  return v;
}
static bool stream_is_live(cudaStream_t s) {
  if (s == rt().own_stream || s == rt().aux_stream) return true;
  for (cudaStream_t t : live_streams()) if (t == s) return true;
  for (cudaStream_t t : adopted_streams()) if (t == s) return true;      // the caller keeps those alive
  return false;
}

int32_t ph_stream_create(void** out_stream) {
  PH_REQUIRE_INIT();
  if (!out_stream) return set_error(PH_ERR_INVALID, "null out_stream");
  cudaStream_t s;
  PH_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  live_streams().push_back(s);
  *out_stream = (void*)s;
  return PH_OK;
}

int32_t ph_stream_destroy(void* stream) {
  PH_REQUIRE_INIT();
  if (!stream) return PH_OK;
  cudaStream_t s = (cudaStream_t)stream;
  std::vector<cudaStream_t>& v = live_streams();
  auto it = std::find(v.begin(), v.end(), s);
  if (it == v.end()) return set_error(PH_ERR_INVALID, "ph_stream_destroy: not a live ph_stream_create stream");
  if (s == rt().stream) rt().stream = rt().own_stream;
  PH_CUDA(cudaStreamSynchronize(s));
  v.erase(it);
  PH_CUDA(cudaStreamDestroy(s));
  return PH_OK;
}

// everything queued on `waiter` AFTER this call runs after everything queued on `signaler` BEFORE it
// (NULL = the library's own stream)
int32_t ph_stream_wait(void* waiter, void* signaler) {
  PH_REQUIRE_INIT();
  Runtime& r = rt();
  cudaStream_t w = waiter ? (cudaStream_t)waiter : r.own_stream, s = signaler ? (cudaStream_t)signaler : r.own_stream;
  if (w == s) return PH_OK;
  cudaEvent_t ev;
  PH_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  PH_CUDA(cudaEventRecord(ev, s));
  PH_CUDA(cudaStreamWaitEvent(w, ev, 0));
  PH_CUDA(cudaEventDestroy(ev));          // released once the recorded work completes
  return PH_OK;
}

int32_t ph_stream_sync(void* stream) {
  PH_REQUIRE_INIT();
  PH_CUDA(cudaStreamSynchronize(stream ? (cudaStream_t)stream : rt().own_stream));
  return PH_OK;
}

// stream-ordered release on the stream the block was last used on (ph_free releases on the CURRENT stream)
int32_t ph_free_on(void* dev, void* stream) {
  PH_REQUIRE_INIT();
  if (!dev) return PH_OK;
  // a block may outlive the stream it was used on: that stream was synchronised when it was destroyed (or
  // was a caller-owned stream adopted with ph_set_stream and is unknown here), so the library's own stream is
  // as good an order as any
  cudaStream_t s = stream ? (cudaStream_t)stream : rt().own_stream;
  if (!stream_is_live(s)) s = rt().own_stream;
  PH_CUDA(cudaFreeAsync(dev, s));
  return PH_OK;
}

int32_t ph_device_count(int32_t* out) {
  if (!out) return set_error(PH_ERR_INVALID, "null out");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { *out = 0; return check_cuda(e, "cudaGetDeviceCount"); }
  *out = n;
  return PH_OK;
}

int32_t ph_init(int32_t device) {
  Runtime& r = rt();
  if (r.inited && r.device == device) return PH_OK;
  if (r.inited) ph_shutdown();
  int n = 0;
  PH_CUDA(cudaGetDeviceCount(&n));
  if (device < 0 || device >= n) return set_error(PH_ERR_INVALID, "device %d not in [0,%d)", device, n);
  PH_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  PH_CUDA(cudaGetDeviceProperties(&prop, device));
  r.sm_count = prop.multiProcessorCount;
  // (An L2 set-aside for persisting lines -- cudaLimitPersistingL2CacheSize -- makes the chained elementwise launches
  // SLOWER, not faster: a*b+c 6831 GB/s with the driver's default, 5959 / 5190 / 4859 with 32 / 64 / 128 MB set aside,
  // profiles/r02_l2_persist_setaside.txt.  The evict_first / evict_last qualifiers of ph_common.cuh work without it.)
  PH_CUDA(cudaStreamCreateWithFlags(&r.own_stream, cudaStreamNonBlocking));
  {
    // the side stream carries halo edges + exchange of sharded stencil runs: highest priority, so its
    // few blocks are placed ahead of the interior update that runs concurrently on the main stream
    int prio_lo = 0, prio_hi = 0;
    PH_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    PH_CUDA(cudaStreamCreateWithPriority(&r.aux_stream, cudaStreamNonBlocking, prio_hi));
  }
  r.stream = r.own_stream;
  PH_CUDA(cudaEventCreateWithFlags(&r.ev_a, cudaEventDisableTiming));
  PH_CUDA(cudaEventCreateWithFlags(&r.ev_b, cudaEventDisableTiming));
  PH_CUDA(cudaEventCreate(&r.ev_t0));
  PH_CUDA(cudaEventCreate(&r.ev_t1));
  PH_CUDA(cudaMalloc((void**)&r.d_flags, 64));
  PH_CUDA(cudaMemset(r.d_flags, 0, 64));
  PH_CUDA(cudaHostAlloc((void**)&r.h_flags, 64, cudaHostAllocMapped));   // [0] flag word, [1] call number (see take_flags)
  memset(r.h_flags, 0, 64);
  PH_CUDA(cudaHostGetDevicePointer((void**)&r.h_flags_dev, r.h_flags, 0));
  r.flag_seq = 0;
  PH_CUDA(cudaMallocHost(&r.h_scratch, 256));
  // keep freed blocks in the pool: fluent ops allocate one result array per operator
  cudaMemPool_t pool;
  PH_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t threshold = UINT64_MAX;
  PH_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
  r.device = device;
  r.launches = 0;
  r.inited = true;
  return PH_OK;
}

int32_t ph_shutdown(void) {
  Runtime& r = rt();
  if (!r.inited) return PH_OK;
  cudaSetDevice(r.device);
  cudaDeviceSynchronize();
  if (r.d_scratch) cudaFree(r.d_scratch);
  if (r.d_flags) cudaFree(r.d_flags);
  if (r.h_flags) cudaFreeHost(r.h_flags);
  if (r.h_scratch) cudaFreeHost(r.h_scratch);
  if (r.ev_a) cudaEventDestroy(r.ev_a);
  if (r.ev_b) cudaEventDestroy(r.ev_b);
  if (r.ev_t0) cudaEventDestroy(r.ev_t0);
  if (r.ev_t1) cudaEventDestroy(r.ev_t1);
  if (r.own_stream) cudaStreamDestroy(r.own_stream);
  if (r.aux_stream) cudaStreamDestroy(r.aux_stream);
  r = Runtime();
  return PH_OK;
}

int32_t ph_sm_count(int32_t* out) {
  PH_REQUIRE_INIT();
  if (!out) return set_error(PH_ERR_INVALID, "null out");
  *out = rt().sm_count;
  return PH_OK;
}

int32_t ph_alloc(size_t nbytes, void** out_dev) {
  PH_REQUIRE_INIT();
  if (!out_dev) return set_error(PH_ERR_INVALID, "null out_dev");
  *out_dev = nullptr;
  if (nbytes == 0) nbytes = 1;   // empty NArrays still own a (tiny) buffer
  PH_CUDA(cudaMallocAsync(out_dev, nbytes, rt().stream));
  return PH_OK;
}

int32_t ph_free(void* dev) {
  PH_REQUIRE_INIT();
  if (!dev) return PH_OK;
  PH_CUDA(cudaFreeAsync(dev, rt().stream));
  return PH_OK;
}

int32_t ph_h2d(void* dst_dev, const void* src_host, size_t nbytes) {
  PH_REQUIRE_INIT();
  if (nbytes == 0) return PH_OK;
  if (!dst_dev || !src_host) return set_error(PH_ERR_INVALID, "null pointer in ph_h2d");
  // synchronous w.r.t. the host pointer unless it is pinned (SURVEY.md 8(b) ownership)
  PH_CUDA(cudaMemcpyAsync(dst_dev, src_host, nbytes, cudaMemcpyHostToDevice, rt().stream));
  return PH_OK;
}

int32_t ph_d2h(void* dst_host, const void* src_dev, size_t nbytes) {
  PH_REQUIRE_INIT();
  if (nbytes == 0) return PH_OK;
  if (!dst_host || !src_dev) return set_error(PH_ERR_INVALID, "null pointer in ph_d2h");
  PH_CUDA(cudaMemcpyAsync(dst_host, src_dev, nbytes, cudaMemcpyDeviceToHost, rt().stream));
  PH_CUDA(cudaStreamSynchronize(rt().stream));
  return PH_OK;
}

// ph_d2h plus the arithmetic flag word in the SAME synchronisation: the read a host layer uses for
// `to_host` / `get`, so that a data-dependent error of any earlier launch surfaces at the read that
// would hand its result to the caller (the reference raises at the offending operator).
// The flag word travels like the result record of a full reduction: one tiny launch takes it (read and clear in
// ONE atomic), stores it and then the call number into a pinned, device-mapped record, and the host polls that
// number.  Stream order makes the poll a synchronisation point for everything enqueued before it (the data copy
// of ph_d2h_flags included).  A 4-byte cudaMemcpyAsync + cudaMemsetAsync + cudaStreamSynchronize cost ~25 us per
// raise point (a per-axis fold of [16384,16384] f32 runs for 150 us); PH_FLAGS_MEMCPY=1 restores that form (A/B).
static __global__ void take_flags_kernel(uint32_t* __restrict__ d_flags, volatile uint32_t* __restrict__ host_rec, uint32_t seq) {
  const uint32_t f = atomicExch(d_flags, 0u);
  host_rec[0] = f;
  __threadfence_system();
  host_rec[1] = seq;
}
static int32_t take_flags(uint32_t* out_flags) {
  Runtime& r = rt();
  static const bool memcpy_form = getenv("PH_FLAGS_MEMCPY") != nullptr;
  if (memcpy_form || !r.h_flags_dev) {
    PH_CUDA(cudaMemcpyAsync(r.h_flags, r.d_flags, 4, cudaMemcpyDeviceToHost, r.stream));
    PH_CUDA(cudaMemsetAsync(r.d_flags, 0, 4, r.stream));
    PH_CUDA(cudaStreamSynchronize(r.stream));
    *out_flags = r.h_flags[0];
    return PH_OK;
  }
  if (++r.flag_seq == 0) ++r.flag_seq;                // 0 is the record's initial state: never a call number
  const uint32_t seq = r.flag_seq;
  take_flags_kernel<<<1, 1, 0, r.stream>>>(r.d_flags, r.h_flags_dev, seq);
  PH_CUDA(cudaGetLastError());                       // not counted by ph_launch_count: it stands in for a 4-byte copy
  const volatile uint32_t* done = r.h_flags + 1;
  uint32_t polls = 0;
  while (*done != seq) {
    if ((++polls & 0xfff) == 0) {                    // a faulted launch never writes the record: ask the stream now and then
      const cudaError_t q = cudaStreamQuery(r.stream);
      if (q != cudaErrorNotReady) {
        PH_CUDA(cudaStreamSynchronize(r.stream));
        if (*done != seq) return set_error(PH_ERR_CUDA, "the flag record was not written");
        break;
      }
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  *out_flags = r.h_flags[0];
  return PH_OK;
}

int32_t ph_d2h_flags(void* dst_host, const void* src_dev, size_t nbytes, uint32_t* out_flags) {
  PH_REQUIRE_INIT();
  if (!out_flags) return set_error(PH_ERR_INVALID, "null out_flags");
  Runtime& r = rt();
  if (nbytes) {
    if (!dst_host || !src_dev) return set_error(PH_ERR_INVALID, "null pointer in ph_d2h_flags");
    PH_CUDA(cudaMemcpyAsync(dst_host, src_dev, nbytes, cudaMemcpyDeviceToHost, r.stream));
  }
  return take_flags(out_flags);
}

int32_t ph_d2h_async(void* dst_host, const void* src_dev, size_t nbytes) {
  PH_REQUIRE_INIT();
  if (nbytes == 0) return PH_OK;
  if (!dst_host || !src_dev) return set_error(PH_ERR_INVALID, "null pointer in ph_d2h_async");
  PH_CUDA(cudaMemcpyAsync(dst_host, src_dev, nbytes, cudaMemcpyDeviceToHost, rt().stream));
  return PH_OK;
}

int32_t ph_d2d(void* dst_dev, const void* src_dev, size_t nbytes) {
  PH_REQUIRE_INIT();
  if (nbytes == 0) return PH_OK;
  if (!dst_dev || !src_dev) return set_error(PH_ERR_INVALID, "null pointer in ph_d2d");
  PH_CUDA(cudaMemcpyAsync(dst_dev, src_dev, nbytes, cudaMemcpyDeviceToDevice, rt().stream));
  return PH_OK;
}

int32_t ph_host_alloc(size_t nbytes, void** out_host) {
  if (!out_host) return set_error(PH_ERR_INVALID, "null out_host");
  PH_CUDA(cudaMallocHost(out_host, nbytes ? nbytes : 1));
  return PH_OK;
}

int32_t ph_host_free(void* host) {
  if (!host) return PH_OK;
  PH_CUDA(cudaFreeHost(host));
  return PH_OK;
}

int32_t ph_sync(void) {
  PH_REQUIRE_INIT();
  PH_CUDA(cudaStreamSynchronize(rt().stream));
  return PH_OK;
}

void* ph_stream(void) { return (void*)rt().stream; }

int32_t ph_set_stream(void* cuda_stream) {
  PH_REQUIRE_INIT();
  cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : rt().own_stream;
  if (!stream_is_live(s)) adopted_streams().push_back(s);
  rt().stream = s;
  return PH_OK;
}

const char* ph_last_error_string(void) { return rt().err; }

int32_t ph_take_arith_flags(uint32_t* out_flags) {
  PH_REQUIRE_INIT();
  if (!out_flags) return set_error(PH_ERR_INVALID, "null out_flags");
  return take_flags(out_flags);
}

int32_t ph_timer_start(void) {
  PH_REQUIRE_INIT();
  PH_CUDA(cudaEventRecord(rt().ev_t0, rt().stream));
  return PH_OK;
}

int32_t ph_timer_stop(float* out_ms) {
  PH_REQUIRE_INIT();
  if (!out_ms) return set_error(PH_ERR_INVALID, "null out_ms");
  PH_CUDA(cudaEventRecord(rt().ev_t1, rt().stream));
  PH_CUDA(cudaEventSynchronize(rt().ev_t1));
  PH_CUDA(cudaEventElapsedTime(out_ms, rt().ev_t0, rt().ev_t1));
  return PH_OK;
}

int64_t ph_launch_count(void) { return rt().launches; }

}  // extern "C"
