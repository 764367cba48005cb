// ph_common.cuh -- shared device/host helpers for libphgpu (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <type_traits>
#include "../../include/ph_gpu.h"

namespace ph {

// ---------------------------------------------------------------- runtime state
struct Runtime {
  int device = -1;
  bool inited = false;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;       // the stream launches go to
  cudaStream_t aux_stream = nullptr;   // halo exchange / overlap
  cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
  uint32_t* d_flags = nullptr;         // device arithmetic flag word (d_flags[8] = reduction ticket)
  uint32_t* h_flags = nullptr;         // pinned, device-mapped record: [0] flag word, [1] call number
  uint32_t* h_flags_dev = nullptr;     // its device address
  uint32_t flag_seq = 0;
  void* d_scratch = nullptr;           // reduction partials
  size_t scratch_bytes = 0;
  void* h_scratch = nullptr;           // pinned result staging (64 B)
  int sm_count = 148;
  long long launches = 0;
  unsigned long long flat_launches = 0; // parity = traversal direction of the next flat elementwise launch
  char err[512] = {0};
};
Runtime& rt();
int32_t set_error(int32_t code, const char* fmt, ...);
int32_t check_cuda(cudaError_t e, const char* what);
int32_t ensure_scratch(size_t bytes);

#define PH_CUDA(call)                                                 \
  do {                                                                \
    int32_t _s = ph::check_cuda((call), #call);                       \
    if (_s != PH_OK) return _s;                                       \
  } while (0)
#define PH_REQUIRE_INIT()                                             \
  do {                                                                \
    if (!ph::rt().inited)                                             \
      return ph::set_error(PH_ERR_NOT_INIT, "ph_init was not called"); \
  } while (0)
#define PH_LAUNCH_CHECK(name)                                         \
  do {                                                                \
    ph::rt().launches++;                                              \
    int32_t _s = ph::check_cuda(cudaGetLastError(), name);            \
    if (_s != PH_OK) return _s;                                       \
  } while (0)

inline int dtype_size(int32_t dt) {
  switch (dt) {
    case PH_F32: case PH_I32: case PH_U32: return 4;
    case PH_F64: case PH_I64: case PH_U64: return 8;
    case PH_U8: case PH_I8: return 1;
    case PH_I16: case PH_U16: return 2;
    default: return 0;
  }
}

// ---------------------------------------------------------------- descriptors
// A "plan" is N operand descriptors over shared extents, normalised: size-1 axes
// removed, adjacent axes merged whenever every operand allows it.
constexpr int MAX_OPERANDS = 4;
struct Plan {
  int rank = 0;                        // after coalescing (0 => a single element)
  int nops = 0;
  int64_t extent[PH_MAX_RANK];
  int64_t stride[MAX_OPERANDS][PH_MAX_RANK];
  int64_t offset[MAX_OPERANDS];
  int64_t total = 0;                   // number of elements (0 => nothing to do)
};
int32_t make_plan(Plan& p, int nops, const ph_desc* const* descs);

// Device-side copy of the outer (non-innermost) axes of a plan.
struct OuterAxes {
  int n;                               // number of outer axes
  int64_t extent[PH_MAX_RANK - 1];
  int64_t stride[MAX_OPERANDS][PH_MAX_RANK - 1];
};

// ---------------------------------------------------------------- vector access
// 32-byte (256-bit) global accesses are an sm_100 feature (LDG.E.256 / STG.E.256).
template <int BYTES> struct RawVec;
template <> struct alignas(32) RawVec<32> { uint64_t q[4]; };
template <> struct alignas(16) RawVec<16> { uint64_t q[2]; };
template <> struct alignas(8)  RawVec<8>  { uint64_t q[1]; };
template <> struct alignas(4)  RawVec<4>  { uint32_t q[1]; };
template <> struct alignas(2)  RawVec<2>  { uint16_t q[1]; };
template <> struct alignas(1)  RawVec<1>  { uint8_t q[1]; };

template <int BYTES>
__device__ __forceinline__ RawVec<BYTES> ld_stream(const void* p) {
  RawVec<BYTES> r;
  if constexpr (BYTES == 32) {
    asm("ld.global.L1::no_allocate.v4.b64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(r.q[0]), "=l"(r.q[1]), "=l"(r.q[2]), "=l"(r.q[3]) : "l"(p));
  } else if constexpr (BYTES == 16) {
    asm("ld.global.L1::no_allocate.v2.b64 {%0,%1}, [%2];"
                 : "=l"(r.q[0]), "=l"(r.q[1]) : "l"(p));
  } else if constexpr (BYTES == 8) {
    asm("ld.global.L1::no_allocate.b64 %0, [%1];" : "=l"(r.q[0]) : "l"(p));
  } else if constexpr (BYTES == 4) {
    asm("ld.global.L1::no_allocate.b32 %0, [%1];" : "=r"(r.q[0]) : "l"(p));
  } else if constexpr (BYTES == 2) {
    asm("ld.global.L1::no_allocate.b16 %0, [%1];" : "=h"(r.q[0]) : "l"(p));
  } else {
    r.q[0] = *reinterpret_cast<const uint8_t*>(p);
  }
  return r;
}

// identical today (all loads are coherent so exact in-place aliasing out == in is legal)
template <int BYTES>
__device__ __forceinline__ RawVec<BYTES> ld_plain(const void* p) {
  RawVec<BYTES> r;
  if constexpr (BYTES == 32) {
    asm("ld.global.v4.b64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(r.q[0]), "=l"(r.q[1]), "=l"(r.q[2]), "=l"(r.q[3]) : "l"(p));
  } else {
    r = *reinterpret_cast<const RawVec<BYTES>*>(p);
  }
  return r;
}

template <int BYTES>
__device__ __forceinline__ void st_stream(void* p, const RawVec<BYTES>& r) {
  if constexpr (BYTES == 32) {
    asm volatile("st.global.L1::no_allocate.v4.b64 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "l"(r.q[0]), "l"(r.q[1]), "l"(r.q[2]), "l"(r.q[3]) : "memory");
  } else if constexpr (BYTES == 16) {
    asm volatile("st.global.L1::no_allocate.v2.b64 [%0], {%1,%2};"
                 :: "l"(p), "l"(r.q[0]), "l"(r.q[1]) : "memory");
  } else if constexpr (BYTES == 8) {
    asm volatile("st.global.L1::no_allocate.b64 [%0], %1;" :: "l"(p), "l"(r.q[0]) : "memory");
  } else if constexpr (BYTES == 4) {
    asm volatile("st.global.L1::no_allocate.b32 [%0], %1;" :: "l"(p), "r"(r.q[0]) : "memory");
  } else if constexpr (BYTES == 2) {
    asm volatile("st.global.L1::no_allocate.b16 [%0], %1;" :: "l"(p), "h"(r.q[0]) : "memory");
  } else {
    *reinterpret_cast<uint8_t*>(p) = r.q[0];
  }
}

// A typed group of N elements backed by one RawVec.
template <typename T, int N>
struct Group {
  static constexpr int BYTES = int(sizeof(T)) * N;
  union { RawVec<BYTES> raw; T v[N]; };
  __device__ __forceinline__ Group() {}
};

template <typename T, int N>
__device__ __forceinline__ Group<T, N> load_group(const T* p) {
  Group<T, N> g;
  g.raw = ld_stream<Group<T, N>::BYTES>(p);
  return g;
}
template <typename T, int N>
__device__ __forceinline__ Group<T, N> load_group_plain(const T* p) {
  Group<T, N> g;
  g.raw = ld_plain<Group<T, N>::BYTES>(p);
  return g;
}
template <typename T, int N>
__device__ __forceinline__ void store_group(T* p, const Group<T, N>& g) {
  st_stream<Group<T, N>::BYTES>(p, g.raw);
}
// L2 eviction priorities for CHAINED elementwise launches (map_flat_kernel): inputs are read once -> evict_first
// (SASS LDG.E.NA.EFL2), results are what the next operator of a fluent expression reads -> evict_last (STG ... ELL2).
// A consumed line that was hit with evict_first becomes the first victim, so the part of a temporary still waiting in
// L2 is not pushed out by the consumer's own output.  `hint` is launch-uniform.
template <int BYTES>
__device__ __forceinline__ RawVec<BYTES> ld_stream_evict_first(const void* p) {
  RawVec<BYTES> r;
  if constexpr (BYTES == 32) {
    asm("ld.global.L1::no_allocate.L2::evict_first.v4.b64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(r.q[0]), "=l"(r.q[1]), "=l"(r.q[2]), "=l"(r.q[3]) : "l"(p));
  } else {                                   // the .L2::evict_* qualifiers exist for 256-bit accesses only
    r = ld_stream<BYTES>(p);
  }
  return r;
}
template <int BYTES>
__device__ __forceinline__ void st_stream_evict_last(void* p, const RawVec<BYTES>& r) {
  if constexpr (BYTES == 32) {
    asm volatile("st.global.L1::no_allocate.L2::evict_last.v4.b64 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "l"(r.q[0]), "l"(r.q[1]), "l"(r.q[2]), "l"(r.q[3]) : "memory");
  } else {
    st_stream<BYTES>(p, r);
  }
}
template <typename T, int N>
__device__ __forceinline__ Group<T, N> load_group_hint(const T* p, bool hint) {
  Group<T, N> g;
  if (hint) g.raw = ld_stream_evict_first<Group<T, N>::BYTES>(p);
  else g.raw = ld_stream<Group<T, N>::BYTES>(p);
  return g;
}
template <typename T, int N>
__device__ __forceinline__ void store_group_hint(T* p, const Group<T, N>& g, bool hint) {
  if (hint) st_stream_evict_last<Group<T, N>::BYTES>(p, g.raw);
  else st_stream<Group<T, N>::BYTES>(p, g.raw);
}
template <typename T, int N>
__device__ __forceinline__ Group<T, N> splat_group(T x) {
  Group<T, N> g;
#pragma unroll
  for (int i = 0; i < N; i++) g.v[i] = x;
  return g;
}

// largest power-of-two byte width (<= cap) dividing the byte address / byte length
inline int align_of(uintptr_t byte_value, int cap = 32) {
  int a = cap;
  while (a > 1 && (byte_value % (uintptr_t)a) != 0) a >>= 1;
  return a;
}

template <typename T>
__host__ __device__ __forceinline__ T bits_to(uint64_t bits) {
  T x;
  if constexpr (sizeof(T) == 8) { memcpy(&x, &bits, 8); }
  else if constexpr (sizeof(T) == 4) { uint32_t b = (uint32_t)bits; memcpy(&x, &b, 4); }
  else if constexpr (sizeof(T) == 2) { uint16_t b = (uint16_t)bits; memcpy(&x, &b, 2); }
  else { uint8_t b = (uint8_t)bits; memcpy(&x, &b, 1); }
  return x;
}
inline uint64_t host_scalar_bits(const void* p, int size) {
  uint64_t b = 0;
  memcpy(&b, p, size);
  return b;
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace ph
