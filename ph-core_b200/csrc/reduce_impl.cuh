// reduce_impl.cuh -- full and per-axis reductions (K7, K8 of SURVEY.md 2.3): kernels and per-dtype launchers.
// Included by reduce_<dtype>.cu, each of which instantiates reduce_full_t / reduce_axis_t for ONE element type
// (ten translation units compile side by side; as a single file this took three minutes); reduce.cu holds the
// C-ABI entry points and only declares the two templates.
//
// Replaces Crystal's Enumerable#sum/min/max folded over NArray#each
// (src/n_array.cr:556-564, block form src/multi_indexable.cr:788-793) and the README's
// argmax idiom (README.md:56-61).  Semantics kept from the reference:
//   * max/min/argmax/argmin: strict > / < from the left, so the FIRST extremum wins (this
//     also decides between -0.0 and +0.0); NaN -> PH_FLAG_NAN (host raises ArgumentError).
//   * integer sum: OverflowError when ANY prefix of the lexicographic fold leaves T.  A
//     commutative pre-filter (sum of positives / of negatives fit in T => no prefix can
//     overflow) decides almost always; otherwise an ordered (sum, max-prefix, min-prefix)
//     monoid pass decides exactly.
//   * float sum: per-block tree in T, then the last block to finish folds the partials in a
//     fixed order inside the same launch (deterministic for a given shape), tolerance-checked.
// Per-axis reductions are defined as the fold of each_slice(axis) in increasing index
// (src/multi_indexable.cr:742-748); the column-strip kernel keeps exactly that order.
#pragma once
#include "map_kernels.cuh"
#include "ops.cuh"
#include "comm.cuh"
#include <limits>
#include <atomic>

namespace ph {

constexpr int RED_THREADS = 256;

// Full reductions over a ROW-STRIDED view (`narr[0..2.., ..]`, a column block of a matrix, ...): logical
// element i = (row, col) lives at x[row * row_stride + col].  When the row length is a whole number of
// tiles every tile lies inside one row, so the kernels only re-base a tile: one 32-bit division per tile
// (tiles_per_row == 0: the plain contiguous array, no division).  The view is read in place -- no gather,
// no temporary: algorithmic bytes = the bytes of the view.
struct RowsArgs {
  uint32_t tiles_per_row = 0;
  int64_t row_stride = 0;
};
__device__ __forceinline__ int64_t tile_base(const RowsArgs& r, int64_t t, int64_t tile) {
  if (r.tiles_per_row == 0) return t * tile;
  const uint32_t row = (uint32_t)t / r.tiles_per_row;
  return (int64_t)row * r.row_stride + (int64_t)((uint32_t)t - row * r.tiles_per_row) * tile;
}

// ---------------------------------------------------------------- helpers
template <typename T> struct Acc { using type = T; };             // sum accumulator
template <> struct Acc<int32_t> { using type = int64_t; };
template <> struct Acc<int64_t> { using type = __int128; };
template <> struct Acc<uint8_t> { using type = int64_t; };
template <> struct Acc<int8_t> { using type = int64_t; };
template <> struct Acc<int16_t> { using type = int64_t; };
template <> struct Acc<uint16_t> { using type = int64_t; };
template <> struct Acc<uint32_t> { using type = int64_t; };
template <> struct Acc<uint64_t> { using type = __int128; };

template <typename T> __device__ __forceinline__ T shfl_down_t(T v, int off) {
  if constexpr (sizeof(T) == 16) {
    uint64_t lo = (uint64_t)v, hi = (uint64_t)((unsigned __int128)v >> 64);
    lo = __shfl_down_sync(0xffffffffu, lo, off);
    hi = __shfl_down_sync(0xffffffffu, hi, off);
    return (T)(((unsigned __int128)hi << 64) | lo);
  } else {
    return __shfl_down_sync(0xffffffffu, v, off);
  }
}

template <typename T> __device__ __forceinline__ T shfl_xor_t(T v, int off) {
  if constexpr (sizeof(T) == 16) {
    uint64_t lo = (uint64_t)v, hi = (uint64_t)((unsigned __int128)v >> 64);
    lo = __shfl_xor_sync(0xffffffffu, lo, off);
    hi = __shfl_xor_sync(0xffffffffu, hi, off);
    return (T)(((unsigned __int128)hi << 64) | lo);
  } else {
    return __shfl_xor_sync(0xffffffffu, v, off);
  }
}

// read a value another block of this launch wrote (L2, never a stale L1 line)
template <typename T> __device__ __forceinline__ T ld_cg(const T* p) {
  if constexpr (sizeof(T) == 16) {
    const unsigned long long* q = reinterpret_cast<const unsigned long long*>(p);
    const unsigned long long lo = __ldcg(q), hi = __ldcg(q + 1);
    return (T)(((unsigned __int128)hi << 64) | lo);
  } else if constexpr (sizeof(T) == 8) {
    const unsigned long long v = __ldcg(reinterpret_cast<const unsigned long long*>(p));
    T r; memcpy(&r, &v, 8); return r;
  } else if constexpr (sizeof(T) == 4) {
    const unsigned int v = __ldcg(reinterpret_cast<const unsigned int*>(p));
    T r; memcpy(&r, &v, 4); return r;
  } else {
    return *reinterpret_cast<const volatile T*>(p);
  }
}

template <typename T> __device__ __forceinline__ T lowest_of() {
  if constexpr (std::is_same<T, float>::value) return -__int_as_float(0x7f800000);
  else if constexpr (std::is_same<T, double>::value) return -__longlong_as_double(0x7ff0000000000000LL);
  else return std::numeric_limits<T>::lowest();
}
template <typename T> __device__ __forceinline__ T highest_of() {
  if constexpr (std::is_same<T, float>::value) return __int_as_float(0x7f800000);
  else if constexpr (std::is_same<T, double>::value) return __longlong_as_double(0x7ff0000000000000LL);
  else return std::numeric_limits<T>::max();
}

// (value, index) candidate for first-extremum selection
template <typename T> struct Cand { T v; int64_t i; };
template <typename T, bool IS_MAX>
__device__ __forceinline__ Cand<T> better(const Cand<T>& a, const Cand<T>& b) {
  // strictly better value wins; equal values -> lower index (the first one met in lex order)
  const bool b_wins = IS_MAX ? (b.v > a.v || (b.v == a.v && b.i < a.i))
                             : (b.v < a.v || (b.v == a.v && b.i < a.i));
  return b_wins ? b : a;
}

static __global__ void set_flag_kernel(uint32_t* flags, uint32_t bits) { atomicOr(flags, bits); }

// ---------------------------------------------------------------- cross-rank combine (one process per GPU)
// A sharded reduction is ONE launch per rank: the block that finishes last stores this rank's partial into
// slot[rank] of EVERY peer's control block over NVLink (payload, system fence, then the word carrying the
// call number), waits until all N slots of its own block carry this call's number, and folds them in rank
// order -- every rank computes the same result from the same records in the same order (deterministic), and
// writes it to a pinned host record.  No NCCL kernel, no second launch, no D2H copy.  Slots are double
// buffered by call parity: a peer can only be one call ahead (it needs MY record to finish a call).
template <typename A> __device__ __forceinline__ void pack_wide(uint64_t* w, A v) {
  if constexpr (sizeof(A) == 16) {
    w[0] = (uint64_t)(unsigned __int128)v;
    w[1] = (uint64_t)((unsigned __int128)v >> 64);
  } else {
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(A));
    w[0] = b;
    w[1] = 0;
  }
}
template <typename A> __device__ __forceinline__ A unpack_wide(const uint64_t* w) {
  if constexpr (sizeof(A) == 16) {
    return (A)(((unsigned __int128)w[1] << 64) | (unsigned __int128)w[0]);
  } else {
    A v;
    const uint64_t b = w[0];
    memcpy(&v, &b, sizeof(A));
    return v;
  }
}

// Called by every thread of the finishing block; `mine` is thread 0's.  Afterwards sh[r] holds rank r's
// record for r < nranks (visible to all threads).  false = a peer never arrived (time-out, ~4 s).
__device__ __forceinline__ bool exchange_records(const CombineArgs& c, const ReduceSlot& mine, ReduceSlot* sh) {
  __shared__ int timed_out;
  if (threadIdx.x == 0) { sh[PH_MAX_PEERS] = mine; timed_out = 0; }
  __syncthreads();
  if ((int)threadIdx.x < c.nranks) {
    const int r = threadIdx.x;
    volatile uint64_t* dst = reinterpret_cast<volatile uint64_t*>(c.peer_slots[r] + c.rank);
    const uint64_t* src = sh[PH_MAX_PEERS].w;
#pragma unroll
    for (int i = 0; i < 7; i++) dst[i] = src[i];
    __threadfence_system();
    dst[7] = (src[7] & 0xffffffffull) | ((uint64_t)c.seq << 32);
    const volatile uint64_t* mine_r = reinterpret_cast<const volatile uint64_t*>(c.my_slots + r);
    const long long t0 = clock64();
    uint64_t tail;
    while ((uint32_t)((tail = mine_r[7]) >> 32) != c.seq) {
      if (clock64() - t0 > 8000000000LL) { timed_out = 1; break; }
    }
    __threadfence_system();
#pragma unroll
    for (int i = 0; i < 7; i++) sh[r].w[i] = mine_r[i];
    sh[r].w[7] = tail;
  }
  __syncthreads();
  return timed_out == 0;
}

template <typename T> __device__ __forceinline__ void sum_finish(typename Acc<T>::type fs, typename Acc<T>::type fp,
                                                                  typename Acc<T>::type fn, uint32_t fl, bool ok,
                                                                  T* out_value, int* status, ReduceResult* host_out,
                                                                  uint32_t seq) {
  using A = typename Acc<T>::type;
  int st = PH_RED_OK;
  if constexpr (!is_float_t<T>::value) {
    const A hi = (A)std::numeric_limits<T>::max(), lo = (A)std::numeric_limits<T>::lowest();
    if (fs > hi || fs < lo) st = PH_RED_OVERFLOW;               // the final prefix itself overflows
    else if (fp > hi || fn < lo) st = PH_RED_NEED_EXACT;        // some prefix MIGHT overflow: decide exactly
  }
  if (!ok) st = PH_RED_TIMEOUT;
  const T v = (T)fs;
  if (status) *status = st;
  if (out_value) *out_value = v;
  if (host_out) {
    uint64_t w[2];
    pack_wide<T>(w, v);
    host_out->value[0] = w[0]; host_out->value[1] = w[1];
    host_out->index = 0;
    host_out->flags = fl;
    host_out->status = st;
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t*>(&host_out->seq) = seq;      // the host is polling this word
  }
}

// fold of the N sum records in rank order (thread 0)
template <typename T> __device__ __forceinline__ void sum_fold_records(const ReduceSlot* rec, int nranks, bool ok, T* out_value,
                                                                        int* status, ReduceResult* host_out, uint32_t seq) {
  using A = typename Acc<T>::type;
  A fs = 0, fp = 0, fn = 0;
  uint32_t fl = 0;
  for (int r = 0; r < nranks; r++) {
    if constexpr (is_float_t<T>::value) fs = f_add(fs, unpack_wide<A>(rec[r].w));
    else { fs += unpack_wide<A>(rec[r].w); fp += unpack_wide<A>(rec[r].w + 2); fn += unpack_wide<A>(rec[r].w + 4); }
    fl |= (uint32_t)rec[r].w[7];
  }
  sum_finish<T>(fs, fp, fn, fl, ok, out_value, status, host_out, seq);
}

template <typename T> __device__ __forceinline__ void ext_finish(Cand<T> fin, uint32_t fl, bool ok, T* out_value,
                                                                  int64_t* out_index, ReduceResult* host_out, uint32_t seq) {
  if (out_value) *out_value = fin.v;
  if (out_index) *out_index = fin.i;
  if (host_out) {
    uint64_t w[2];
    pack_wide<T>(w, fin.v);
    host_out->value[0] = w[0]; host_out->value[1] = w[1];
    host_out->index = fin.i;
    host_out->flags = fl;
    host_out->status = !ok ? PH_RED_TIMEOUT : (fin.i == INT64_MAX ? PH_RED_EMPTY : PH_RED_OK);
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t*>(&host_out->seq) = seq;
  }
}

template <typename T, bool IS_MAX> __device__ __forceinline__ void ext_fold_records(const ReduceSlot* rec, int nranks, bool ok,
                                                                                    T* out_value, int64_t* out_index,
                                                                                    ReduceResult* host_out, uint32_t seq) {
  Cand<T> fin;
  fin.v = IS_MAX ? lowest_of<T>() : highest_of<T>();
  fin.i = INT64_MAX;
  uint32_t fl = 0;
  for (int r = 0; r < nranks; r++) {
    Cand<T> c;
    c.v = unpack_wide<T>(rec[r].w);
    c.i = (int64_t)rec[r].w[6];
    if (c.i != INT64_MAX) fin = better<T, IS_MAX>(fin, c);      // an empty shard contributes nothing
    fl |= (uint32_t)rec[r].w[7];
  }
  ext_finish<T>(fin, fl, ok, out_value, out_index, host_out, seq);
}

// NCCL transport of the same combine: the records were allgathered between the two launches
template <typename T> __global__ void sum_combine_kernel(const ReduceSlot* rec, int nranks, T* out_value, int* status,
                                                         ReduceResult* host_out, uint32_t seq) {
  if (threadIdx.x == 0) sum_fold_records<T>(rec, nranks, true, out_value, status, host_out, seq);
}
template <typename T, bool IS_MAX> __global__ void ext_combine_kernel(const ReduceSlot* rec, int nranks, T* out_value,
                                                                       int64_t* out_index, ReduceResult* host_out, uint32_t seq) {
  if (threadIdx.x == 0) ext_fold_records<T, IS_MAX>(rec, nranks, true, out_value, out_index, host_out, seq);
}

// ---------------------------------------------------------------- full sum (floats; ints: S, P, N)
template <typename T> struct SumState {
  using A = typename Acc<T>::type;
  A s, pos, neg;
};

// The block that finishes LAST (atomic ticket) folds the per-block partials: no second launch, and
// the fold order is fixed (thread j takes partials j, j+256, ... in order, then a fixed tree), so
// the result is deterministic for a given grid.  out_value = the sum in T; status = 0 ok /
// 1 definitely-overflow / 2 need the exact ordered pass (integers only).
template <typename T, int E>
__global__ void __launch_bounds__(RED_THREADS) sum_partial_kernel(const T* __restrict__ x, int64_t n,
                                                                  SumState<T>* __restrict__ partials,
                                                                  T* __restrict__ out_value, int* __restrict__ status,
                                                                  unsigned int* __restrict__ ticket,
                                                                  uint32_t* __restrict__ flags, const CombineArgs cmb,
                                                                  const RowsArgs rows) {
  using A = typename Acc<T>::type;
  constexpr bool IS_INT = !is_float_t<T>::value;
  constexpr int UNROLL = 4;
  A acc[E];
  A pos = 0, neg = 0;
#pragma unroll
  for (int i = 0; i < E; i++) acc[i] = 0;
  const int64_t tile = (int64_t)RED_THREADS * E * UNROLL;
  const int64_t ntiles = n / tile;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t base = tile_base(rows, t, tile) + (int64_t)threadIdx.x * E;
    Group<T, E> g[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) g[u] = load_group<T, E>(x + base + (int64_t)u * RED_THREADS * E);
#pragma unroll
    for (int u = 0; u < UNROLL; u++)
#pragma unroll
      for (int i = 0; i < E; i++) {
        if constexpr (IS_INT) {
          const A v = (A)g[u].v[i];
          acc[i] += v;
          if (v > 0) pos += v; else neg += v;
        } else {
          acc[i] = f_add(acc[i], g[u].v[i]);
        }
      }
  }
  // tail (< one tile): spread over the first block
  if (blockIdx.x == 0) {
    for (int64_t i = ntiles * tile + threadIdx.x; i < n; i += RED_THREADS) {
      if constexpr (IS_INT) {
        const A v = (A)x[i];
        acc[0] += v;
        if (v > 0) pos += v; else neg += v;
      } else {
        acc[0] = f_add(acc[0], x[i]);
      }
    }
  }
  A s = acc[0];
#pragma unroll
  for (int i = 1; i < E; i++) {
    if constexpr (IS_INT) s += acc[i]; else s = f_add(s, acc[i]);
  }
  __shared__ A sh_s[RED_THREADS / 32], sh_p[RED_THREADS / 32], sh_n[RED_THREADS / 32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // warp then block tree (fixed order => deterministic); thread 0 ends up with the block totals
  auto block_fold = [&](A& ts, A& tp, A& tn) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      if constexpr (IS_INT) {
        ts += shfl_down_t<A>(ts, off); tp += shfl_down_t<A>(tp, off); tn += shfl_down_t<A>(tn, off);
      } else {
        ts = f_add(ts, shfl_down_t<A>(ts, off));
      }
    }
    if (lane == 0) { sh_s[warp] = ts; sh_p[warp] = tp; sh_n[warp] = tn; }
    __syncthreads();
    if (threadIdx.x == 0) {
      ts = sh_s[0]; tp = sh_p[0]; tn = sh_n[0];
      for (int w = 1; w < RED_THREADS / 32; w++) {
        if constexpr (IS_INT) { ts += sh_s[w]; tp += sh_p[w]; tn += sh_n[w]; }
        else ts = f_add(ts, sh_s[w]);
      }
    }
  };
  block_fold(s, pos, neg);
  if (threadIdx.x == 0) {
    partials[blockIdx.x].s = s; partials[blockIdx.x].pos = pos; partials[blockIdx.x].neg = neg;
    __threadfence();
    is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  A fs = 0, fp = 0, fn = 0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += RED_THREADS) {
    const SumState<T>* q = partials + i;
    if constexpr (IS_INT) { fs += ld_cg(&q->s); fp += ld_cg(&q->pos); fn += ld_cg(&q->neg); }
    else fs = f_add(fs, ld_cg(&q->s));
  }
  __syncthreads();                                   // sh_* are reused
  block_fold(fs, fp, fn);
  if (threadIdx.x == 0) *ticket = 0;                 // ready for the next launch on this stream
  if (cmb.nranks > 1) {                              // sharded: combine the per-GPU partials across ranks
    __shared__ ReduceSlot sh_rec[PH_MAX_PEERS + 1];
    ReduceSlot mine;
    if (threadIdx.x == 0) {
      pack_wide<A>(mine.w, fs); pack_wide<A>(mine.w + 2, fp); pack_wide<A>(mine.w + 4, fn);
      mine.w[6] = 0;
      mine.w[7] = atomicExch(flags, 0u);             // pending arithmetic flags travel with the record
    }
    if (cmb.peer_slots[0] == nullptr) {              // NCCL transport: the allgather follows this launch
      if (threadIdx.x == 0) { mine.w[7] |= (uint64_t)cmb.seq << 32; cmb.my_slots[cmb.rank] = mine; }
      return;
    }
    const bool ok = exchange_records(cmb, mine, sh_rec);
    if (threadIdx.x == 0) sum_fold_records<T>(sh_rec, cmb.nranks, ok, out_value, status, cmb.host_out, cmb.seq);
    return;
  }
  if (threadIdx.x == 0)
    sum_finish<T>(fs, fp, fn, cmb.host_out ? atomicExch(flags, 0u) : 0u, true, out_value, status, cmb.host_out, cmb.seq);
}

// Exact ordered pass for integer sums: monoid (sum, max prefix, min prefix) combined in lex order.
template <typename T> struct Prefix {
  using A = typename Acc<T>::type;
  A s, mx, mn;
};
template <typename T>
__device__ __forceinline__ Prefix<T> pcombine(const Prefix<T>& l, const Prefix<T>& r) {
  Prefix<T> o;
  o.s = l.s + r.s;
  const typename Prefix<T>::A a = l.s + r.mx, b = l.s + r.mn;
  o.mx = l.mx > a ? l.mx : a;
  o.mn = l.mn < b ? l.mn : b;
  return o;
}
// each block owns a CONTIGUOUS range; each thread a contiguous sub-range (uncoalesced but exact;
// this kernel only runs when the commutative filter could not decide).
template <typename T>
__global__ void __launch_bounds__(RED_THREADS) sum_exact_kernel(const T* __restrict__ x, int64_t n,
                                                                Prefix<T>* __restrict__ partials) {
  using A = typename Acc<T>::type;
  const int64_t nthreads = (int64_t)gridDim.x * RED_THREADS;
  const int64_t per = (n + nthreads - 1) / nthreads;
  const int64_t gid = (int64_t)blockIdx.x * RED_THREADS + threadIdx.x;
  int64_t b = gid * per, e = b + per;
  if (b > n) b = n;
  if (e > n) e = n;
  Prefix<T> st; st.s = 0; st.mx = 0; st.mn = 0;     // prefixes include the empty prefix (acc = 0)
  for (int64_t i = b; i < e; i++) {
    st.s += (A)x[i];
    if (st.s > st.mx) st.mx = st.s;
    if (st.s < st.mn) st.mn = st.s;
  }
  __shared__ Prefix<T> sh[RED_THREADS];
  sh[threadIdx.x] = st;
  __syncthreads();
  if (threadIdx.x == 0) {
    Prefix<T> acc = sh[0];
    for (int t = 1; t < RED_THREADS; t++) acc = pcombine<T>(acc, sh[t]);
    partials[blockIdx.x] = acc;
  }
}
// thread 0 folds this rank's per-block prefixes in order; sharded: the per-rank (sum, max prefix, min prefix)
// triples are exchanged like any other record and folded in RANK order -- the lexicographic order of the
// global array -- so the overflow decision is exact across ranks too.
template <typename T>
__global__ void sum_exact_final_kernel(const Prefix<T>* __restrict__ partials, int nparts, int* __restrict__ status,
                                       const CombineArgs cmb) {
  using A = typename Acc<T>::type;
  Prefix<T> acc;
  acc.s = 0; acc.mx = 0; acc.mn = 0;
  if (threadIdx.x == 0 && nparts > 0) {
    acc = partials[0];
    for (int i = 1; i < nparts; i++) acc = pcombine<T>(acc, partials[i]);
  }
  bool ok = true;
  if (cmb.nranks > 1) {
    __shared__ ReduceSlot sh_rec[PH_MAX_PEERS + 1];
    ReduceSlot mine;
    if (threadIdx.x == 0) {
      pack_wide<A>(mine.w, acc.s); pack_wide<A>(mine.w + 2, acc.mx); pack_wide<A>(mine.w + 4, acc.mn);
      mine.w[6] = 0; mine.w[7] = 0;
    }
    const ReduceSlot* rec = sh_rec;
    if (cmb.peer_slots[0] == nullptr) rec = cmb.my_slots;          // NCCL transport: already gathered (second launch)
    else ok = exchange_records(cmb, mine, sh_rec);
    if (threadIdx.x == 0) {
      acc.s = unpack_wide<A>(rec[0].w); acc.mx = unpack_wide<A>(rec[0].w + 2); acc.mn = unpack_wide<A>(rec[0].w + 4);
      for (int r = 1; r < cmb.nranks; r++) {
        Prefix<T> o;
        o.s = unpack_wide<A>(rec[r].w); o.mx = unpack_wide<A>(rec[r].w + 2); o.mn = unpack_wide<A>(rec[r].w + 4);
        acc = pcombine<T>(acc, o);
      }
    }
  }
  if (threadIdx.x != 0) return;
  const A hi = (A)std::numeric_limits<T>::max(), lo = (A)std::numeric_limits<T>::lowest();
  const int st = !ok ? PH_RED_TIMEOUT : ((acc.mx > hi || acc.mn < lo) ? PH_RED_OVERFLOW : PH_RED_OK);
  if (status) *status = st;
  if (cmb.host_out) { cmb.host_out->status = st; __threadfence_system(); }
}
// NCCL transport, first launch: leave this rank's prefix triple in the gather send buffer
template <typename T>
__global__ void sum_exact_record_kernel(const Prefix<T>* __restrict__ partials, int nparts, const CombineArgs cmb) {
  using A = typename Acc<T>::type;
  if (threadIdx.x != 0) return;
  Prefix<T> acc;
  acc.s = 0; acc.mx = 0; acc.mn = 0;
  if (nparts > 0) {
    acc = partials[0];
    for (int i = 1; i < nparts; i++) acc = pcombine<T>(acc, partials[i]);
  }
  ReduceSlot mine;
  pack_wide<A>(mine.w, acc.s); pack_wide<A>(mine.w + 2, acc.mx); pack_wide<A>(mine.w + 4, acc.mn);
  mine.w[6] = 0; mine.w[7] = (uint64_t)cmb.seq << 32;
  cmb.my_slots[cmb.rank] = mine;
}

// ---------------------------------------------------------------- full min / max / argmin / argmax
// Running extremum of a register tile.  f32 uses max.NaN / min.NaN (one instruction that also
// propagates NaN, so NaN detection is free); other types compare + select and test NaN apart.
template <typename T, bool IS_MAX>
__device__ __forceinline__ T ext2(T a, T b, bool& nan) {
  if constexpr (std::is_same<T, float>::value) {
    float r;
    if (IS_MAX) asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    else asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
  } else {
    if constexpr (is_float_t<T>::value) nan |= (b != b);
    return IS_MAX ? (b > a ? b : a) : (b < a ? b : a);
  }
}

// Hot loop: 4 x 32-byte loads, the extremum of the 32 register values, and -- only when a tile
// strictly improves on the running extremum -- the tile NUMBER.  The first index inside the
// winning tile is searched once per thread after the loop (the tile is re-read, 128 bytes), so
// the loop carries no index arithmetic and few registers.  Tiles are visited in increasing
// index order, so "strictly improves" keeps the FIRST extremum.  The last block to finish
// (atomic ticket) folds the per-block candidates: no second launch.
template <typename T, int E, bool IS_MAX>
__global__ void __launch_bounds__(RED_THREADS) ext_partial_kernel(const T* __restrict__ x, int64_t n,
                                                                  Cand<T>* __restrict__ partials,
                                                                  uint32_t* __restrict__ flags,
                                                                  T* __restrict__ out_value,
                                                                  int64_t* __restrict__ out_index,
                                                                  unsigned int* __restrict__ ticket, const CombineArgs cmb,
                                                                  const RowsArgs rows) {
  constexpr int UNROLL = 4;
  T best_v = IS_MAX ? lowest_of<T>() : highest_of<T>();
  int64_t best_t = -1;
  bool nan = false;
  const int64_t tile = (int64_t)RED_THREADS * E * UNROLL;
  const int64_t ntiles = n / tile;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t base = tile_base(rows, t, tile) + (int64_t)threadIdx.x * E;
    Group<T, E> g[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) g[u] = load_group<T, E>(x + base + (int64_t)u * RED_THREADS * E);
    T m = g[0].v[0];
    if constexpr (is_float_t<T>::value && !std::is_same<T, float>::value) nan |= (m != m);
#pragma unroll
    for (int u = 0; u < UNROLL; u++)
#pragma unroll
      for (int i = 0; i < E; i++)
        if (u || i) m = ext2<T, IS_MAX>(m, g[u].v[i], nan);
    if constexpr (std::is_same<T, float>::value) nan |= (m != m);
    const bool improves = IS_MAX ? (m > best_v) : (m < best_v);
    if (improves || (best_t < 0 && m == best_v)) { best_v = m; best_t = t; }
  }
  Cand<T> best;
  best.v = best_v;
  best.i = INT64_MAX;
  // Only threads whose own extremum EQUALS the block's can hold the block's first extremum: the others skip the
  // re-read of their winning tile (4 scattered 32-byte groups per thread, which DRAM serves as 128-byte lines:
  // ~140 MB of extra reads per launch when every thread did it -- 13 % on a 1 GB array).
  __shared__ T sh_bv[RED_THREADS / 32];
  T block_v = best_v;
  {
    bool ignore = false;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) block_v = ext2<T, IS_MAX>(block_v, shfl_xor_t<T>(block_v, off), ignore);
    if ((threadIdx.x & 31) == 0) sh_bv[threadIdx.x >> 5] = block_v;
    __syncthreads();
    block_v = sh_bv[0];
#pragma unroll
    for (int w = 1; w < RED_THREADS / 32; w++) block_v = ext2<T, IS_MAX>(block_v, sh_bv[w], ignore);
  }
  if (best_t >= 0 && best_v == block_v) {            // first register of the winning tile holding the extremum
    const int64_t base = tile_base(rows, best_t, tile) + (int64_t)threadIdx.x * E;      // where the tile lives
    const int64_t lbase = best_t * tile + (int64_t)threadIdx.x * E;                     // its LOGICAL (lex) index
    bool found = false;
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const Group<T, E> g = load_group_plain<T, E>(x + base + (int64_t)u * RED_THREADS * E);
#pragma unroll
      for (int i = 0; i < E; i++)
        if (!found && g.v[i] == best_v) {
          found = true;
          best.v = g.v[i];                           // the element itself (keeps the sign of a zero)
          best.i = lbase + (int64_t)u * RED_THREADS * E + i;
        }
    }
  }
  if (blockIdx.x == 0) {
    for (int64_t i = ntiles * tile + threadIdx.x; i < n; i += RED_THREADS) {
      const T v = x[i];
      if constexpr (is_float_t<T>::value) nan |= (v != v);
      Cand<T> c; c.v = v; c.i = i;
      best = better<T, IS_MAX>(best, c);
    }
  }
  __shared__ Cand<T> sh[RED_THREADS / 32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  auto block_fold = [&](Cand<T>& b) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      Cand<T> o;
      o.v = __shfl_down_sync(0xffffffffu, b.v, off);
      o.i = __shfl_down_sync(0xffffffffu, b.i, off);
      b = better<T, IS_MAX>(b, o);
    }
    if (lane == 0) sh[warp] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
      b = sh[0];
      for (int w = 1; w < RED_THREADS / 32; w++) b = better<T, IS_MAX>(b, sh[w]);
    }
  };
  if (nan) atomicOr(flags, (uint32_t)PH_FLAG_NAN);   // before the fold's barrier: ordered ahead of thread 0's ticket
  block_fold(best);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = best;
    __threadfence();
    is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  Cand<T> fin;
  fin.v = IS_MAX ? lowest_of<T>() : highest_of<T>();
  fin.i = INT64_MAX;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += RED_THREADS) {
    Cand<T> c;
    c.v = ld_cg(&partials[i].v);
    c.i = ld_cg(&partials[i].i);
    fin = better<T, IS_MAX>(fin, c);
  }
  __syncthreads();                                   // sh is reused
  block_fold(fin);
  if (threadIdx.x == 0) *ticket = 0;
  if (cmb.nranks > 1) {                              // sharded: first extremum across ranks (value, then GLOBAL index)
    __shared__ ReduceSlot sh_rec[PH_MAX_PEERS + 1];
    ReduceSlot mine;
    if (threadIdx.x == 0) {
      pack_wide<T>(mine.w, fin.v);
      mine.w[2] = mine.w[3] = mine.w[4] = mine.w[5] = 0;
      mine.w[6] = (uint64_t)(fin.i == INT64_MAX ? INT64_MAX : fin.i + cmb.elems_before);
      mine.w[7] = atomicExch(flags, 0u);
    }
    if (cmb.peer_slots[0] == nullptr) {
      if (threadIdx.x == 0) { mine.w[7] |= (uint64_t)cmb.seq << 32; cmb.my_slots[cmb.rank] = mine; }
      return;
    }
    const bool ok = exchange_records(cmb, mine, sh_rec);
    if (threadIdx.x == 0) ext_fold_records<T, IS_MAX>(sh_rec, cmb.nranks, ok, out_value, out_index, cmb.host_out, cmb.seq);
    return;
  }
  if (threadIdx.x == 0) ext_finish<T>(fin, cmb.host_out ? atomicExch(flags, 0u) : 0u, true, out_value, out_index, cmb.host_out, cmb.seq);
}

// ---------------------------------------------------------------- per-axis: [outer, K, inner]
// inner > 1: threads run along `inner` (coalesced), each folds its column over k = 0..K-1 in order.
// UNROLL groups are in flight per thread; the block size is the launch's (64..256).
template <typename T, int E, int RED, int UNROLL>
__global__ void __launch_bounds__(RED_THREADS) axis_strip_kernel(const T* __restrict__ x, void* __restrict__ out,
                                                                 int64_t outer, int64_t K, int64_t inner,
                                                                 int64_t ostride, int64_t kstride, int irev,
                                                                 uint32_t* __restrict__ flags) {
  // element (o, k, c) lives at x[o * ostride + k * kstride + c]: a contiguous array has ostride = K * inner and
  // kstride = inner; a row-strided or reversed-outer view only changes the two strides (read in place).
  // irev: the inner axis runs BACKWARDS in memory (x[... - c]): a thread's group of logical columns c .. c+E-1
  // is the memory group ending at -c, loaded whole and reversed in registers (free: the loop is unrolled)
  const int64_t groups = inner / E;                         // inner % E == 0 by dispatch
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= outer * groups) return;
  const int64_t o = gid / groups;
  const int64_t c = (gid - o * groups) * E;
  const T* p = x + o * ostride + (irev ? -(c + E - 1) : c);
  auto load = [&](const T* q) {
    Group<T, E> g = load_group<T, E>(q);
    if (irev) {
#pragma unroll
      for (int i = 0; i < E / 2; i++) { const T t = g.v[i]; g.v[i] = g.v[E - 1 - i]; g.v[E - 1 - i] = t; }
    }
    return g;
  };
  uint32_t err = 0;
  bool nan = false;
  T acc[E];
  int32_t arg[E];                                           // an axis extent always fits Int32
#pragma unroll
  for (int i = 0; i < E; i++) { acc[i] = (RED == PH_SUM) ? (T)0 : p[irev ? E - 1 - i : i]; arg[i] = 0; }
  int64_t k = (RED == PH_SUM) ? 0 : 1;
  if constexpr (RED != PH_SUM && is_float_t<T>::value) {
#pragma unroll
    for (int i = 0; i < E; i++) nan |= (acc[i] != acc[i]);
  }
  auto fold = [&](const Group<T, E>& g, int32_t kk) {
#pragma unroll
    for (int i = 0; i < E; i++) {
      const T v = g.v[i];
      if constexpr (RED == PH_SUM) {
        if constexpr (is_float_t<T>::value) acc[i] = f_add(acc[i], v);
        else acc[i] = i_add<T>(acc[i], v, true, err);
      } else {
        if constexpr (is_float_t<T>::value) nan |= (v != v);
        const bool take = (RED == PH_MAX || RED == PH_ARGMAX) ? (v > acc[i]) : (v < acc[i]);
        if (take) { acc[i] = v; arg[i] = kk; }
      }
    }
  };
  for (; k + UNROLL <= K; k += UNROLL) {
    Group<T, E> g[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) g[u] = load(p + (k + u) * kstride);
#pragma unroll
    for (int u = 0; u < UNROLL; u++) fold(g[u], (int32_t)(k + u));
  }
  for (; k < K; k++) fold(load(p + k * kstride), (int32_t)k);
  if constexpr (RED == PH_ARGMAX || RED == PH_ARGMIN) {
    int64_t* q = reinterpret_cast<int64_t*>(out) + o * inner + c;
#pragma unroll
    for (int i = 0; i < E; i++) q[i] = (int64_t)arg[i];
  } else {
    Group<T, E> r;
#pragma unroll
    for (int i = 0; i < E; i++) r.v[i] = acc[i];
    store_group<T, E>(reinterpret_cast<T*>(out) + o * inner + c, r);
  }
  if (err) atomicOr(flags, err);
  if (nan) atomicOr(flags, (uint32_t)PH_FLAG_NAN);
}

// ---- the same ordered fold for FEW columns (a [16384, 16384] matrix folded down axis 0 has 16 384 columns: one
// thread per column leaves ~110 threads per SM, 14 KB in flight, 1.3 TB/s however deep the unrolling; and K
// cannot be split across threads without changing the association order of a float sum).  Bytes in flight are
// decoupled from the thread count instead: a block is ONE warp that owns a strip of 32 columns for the whole
// of K and streams it through a ring of STAGES shared-memory tiles of KT rows with 16-byte `cp.async` copies
// (LDGSTS, L1 bypassed), STAGES-1 tiles ahead of the fold; lane c then folds column c of each tile in k order
// from shared memory (conflict-free).  No __syncthreads: cp.async.wait_group + __syncwarp.  Strides may be
// negative (reversed views); rows of a strip must be 16-byte aligned (dispatch).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T, int RED, int KT, int STAGES>
__global__ void __launch_bounds__(32) axis_strip_staged_kernel(const T* __restrict__ x, void* __restrict__ out,
                                                               int64_t K, int64_t inner, int64_t ostride, int64_t kstride,
                                                               int irev, int strips_per_outer, uint32_t* __restrict__ flags) {
  constexpr int CW = 32;                                    // columns per strip: one lane each
  constexpr int EPC = 16 / (int)sizeof(T);                  // elements per 16-byte chunk
  constexpr int CPR = CW / EPC;                             // chunks per tile row
  constexpr int CPL = KT * CPR / 32;                        // chunks per lane per tile
  static_assert((KT * CPR) % 32 == 0, "a tile is a whole number of chunks per lane");
  __shared__ __align__(16) T ring[STAGES][KT][CW];
  const int lane = threadIdx.x;
  const int64_t o = blockIdx.x / strips_per_outer;
  const int64_t c0 = (int64_t)(blockIdx.x - o * strips_per_outer) * CW;
  const int ncols = (int)((inner - c0) < CW ? (inner - c0) : CW);      // a multiple of EPC (dispatch)
  // irev: the inner axis runs backwards in memory; the strip's LOGICAL columns c0 .. c0+ncols-1 are the memory
  // columns -(c0+ncols-1) .. -c0: the tile is copied in memory order and lane c reads tile column ncols-1-c
  const T* base = x + o * ostride + (irev ? -(c0 + ncols - 1) : c0);
  const int mycol = irev ? ncols - 1 - lane : lane;
  const int ntiles = (int)((K + KT - 1) / KT);
  // this lane's chunks of a tile: rows r0 + j * (32 / CPR), columns cc .. cc + EPC
  const int r0 = lane / CPR, cc = (lane % CPR) * EPC;
  const bool col_ok = cc < ncols;
  // Issue path kept to ~3 instructions per 16-byte chunk: the lane's source pointer walks the strip (one 64-bit add
  // per chunk, one per tile), whole tiles carry no bounds test.  (Recomputing `(k0 + r) * kstride` and testing
  // `k0 + r < K` per chunk cost ~12 instructions each -- more than the fold itself -- and made f32 strips
  // instruction-bound per warp.)
  const int Ki = (int)K;                                    // K < 2^31 (dispatch)
  const int64_t row_step = (int64_t)(32 / CPR) * kstride;   // between a lane's consecutive chunks of one tile
  const int64_t tile_step = (int64_t)KT * kstride;
  const T* next_src = base + (int64_t)r0 * kstride + cc;    // this lane's first chunk of the next tile to issue
  int next_tile = 0;
  auto issue = [&]() {
    const int slot = next_tile % STAGES;
    const int k0 = next_tile * KT;
    if (col_ok) {
      T* dst = &ring[slot][r0][cc];
      const T* p = next_src;
      if (k0 + KT <= Ki) {
#pragma unroll
        for (int j = 0; j < CPL; j++) { cp_async16(dst + j * (32 / CPR) * CW, p); p += row_step; }
      } else {
#pragma unroll
        for (int j = 0; j < CPL; j++) {
          if (k0 + r0 + j * (32 / CPR) < Ki) cp_async16(dst + j * (32 / CPR) * CW, p);
          p += row_step;
        }
      }
    }
    next_src += tile_step;
    next_tile++;
  };
#pragma unroll
  for (int t = 0; t < STAGES - 1; t++) {
    if (t < ntiles) issue();
    cp_async_commit();
  }
  // A sum is ONE chain (the k order is the result).  Extrema are order-free, so they run as NA independent
  // chains over rows k % NA and merge at the end (better value, then lower k): with ~3.5 warps per SM a single
  // dependent compare-select chain per lane -- ~10 cycles per f64 element -- was the limit (0.82 of the peak).
  constexpr bool IS_MAXLIKE = (RED == PH_MAX || RED == PH_ARGMAX);
  constexpr int NA = (RED == PH_SUM) ? 1 : 4;
  static_assert(KT % NA == 0, "a tile is a whole number of accumulator rounds");
  T acc[NA];
  int32_t arg[NA];
#pragma unroll
  for (int j = 0; j < NA; j++) { acc[j] = (RED == PH_SUM) ? (T)0 : (IS_MAXLIKE ? lowest_of<T>() : highest_of<T>()); arg[j] = INT32_MAX; }
  uint32_t err = 0;
  bool nan = false;
  const bool act = lane < ncols;
  auto fold = [&](T v, int32_t kk, int j) {
    if constexpr (RED == PH_SUM) {
      if constexpr (is_float_t<T>::value) acc[0] = f_add(acc[0], v);
      else acc[0] = i_add<T>(acc[0], v, true, err);
    } else {
      if constexpr (is_float_t<T>::value) nan |= (v != v);
      const bool take = IS_MAXLIKE ? (v > acc[j]) : (v < acc[j]);
      if (take || arg[j] == INT32_MAX) { acc[j] = v; arg[j] = kk; }       // the chain's first element always enters
    }
  };
  for (int t = 0; t < ntiles; t++) {
    if (t + STAGES - 1 < ntiles) issue();                   // tile t + STAGES - 1, into the slot folded in the previous iteration
    cp_async_commit();
    cp_async_wait<STAGES - 1>();                            // tile t has landed (this lane's chunks) ...
    __syncwarp();                                           // ... and every other lane's
    const int slot = t % STAGES;
    const int k0 = t * KT;
    if (act) {
      if (k0 + KT <= Ki) {
#pragma unroll
        for (int r = 0; r < KT; r++) fold(ring[slot][r][mycol], (int32_t)(k0 + r), r % NA);
      } else {
        for (int r = 0; k0 + r < Ki; r++) fold(ring[slot][r][mycol], (int32_t)(k0 + r), 0);
      }
    }
    __syncwarp();                                           // the slot may be refilled
  }
  if (act) {
    T best = acc[0];
    int32_t barg = arg[0];
#pragma unroll
    for (int j = 1; j < NA; j++) {                    // first extremum: better value, then lower k (a zero keeps ITS sign)
      const bool better = arg[j] != INT32_MAX &&
                          (barg == INT32_MAX || (IS_MAXLIKE ? acc[j] > best : acc[j] < best) || (acc[j] == best && arg[j] < barg));
      if (better) { best = acc[j]; barg = arg[j]; }
    }
    if constexpr (RED == PH_ARGMAX || RED == PH_ARGMIN) reinterpret_cast<int64_t*>(out)[o * inner + c0 + lane] = (int64_t)(barg == INT32_MAX ? 0 : barg);
    else reinterpret_cast<T*>(out)[o * inner + c0 + lane] = best;
  }
  if (err) atomicOr(flags, err);
  if (nan) atomicOr(flags, (uint32_t)PH_FLAG_NAN);
}

// inner == 1: each row of K contiguous elements is reduced by TX cooperating threads (TX a
// power of two; <= 32: lanes of one warp, several rows per warp; larger: TX / 32 whole warps).  Lanes read
// consecutive 32-byte groups (E elements), UNROLL groups in flight per lane; lanes combine by a shuffle tree,
// warps through one shared-memory word each (a serial fold of TX shared-memory words by one lane cost a
// 128 KB row ~4 us of single-thread work: 5.0 TB/s; the tree form streams at the copy peak).
// min/max/arg*: ONE pass.  The loop keeps the running extremum and the BATCH (first group + group count)
// that gave it -- the first such batch in the view's order (a reversed view takes the later physical batch
// on ties) -- exactly like the full reduction keeps a tile number; after the row extremum M is known, only
// lanes holding M re-read their own winning batch (<= 128 bytes) for the first logical index.  That index is
// needed for arg*, and for max/min only when M is a zero (its sign depends on which zero came first).
template <typename T, int E, int RED>
__global__ void __launch_bounds__(RED_THREADS) axis_row_kernel(const T* __restrict__ x, void* __restrict__ out,
                                                               int64_t rows, int64_t K, int64_t rstride, int rev,
                                                               int tx, int tx_log2, uint32_t* __restrict__ flags) {
  // row r occupies x[r * rstride .. + K) in memory; rev != 0: its LOGICAL element j is the physical K-1-j (a
  // reversed view): values are order-free, only "first occurrence" is taken on the logical index
  constexpr bool IS_MAXLIKE = (RED == PH_MAX || RED == PH_ARGMAX);
  constexpr bool IS_ARG = (RED == PH_ARGMAX || RED == PH_ARGMIN);
  constexpr int UNROLL = 4;
  constexpr int NWARPS = RED_THREADS / 32;
  using A = typename Acc<T>::type;
  __shared__ A sh_s[NWARPS], sh_p[NWARPS], sh_n[NWARPS];
  __shared__ T sh_m[NWARPS];
  __shared__ int32_t sh_i[NWARPS];
  const int lane = threadIdx.x & (tx - 1);
  const int ty = threadIdx.x >> tx_log2;
  const int TY = RED_THREADS >> tx_log2;
  const int warp = threadIdx.x >> 5;
  const int w0 = (ty << tx_log2) >> 5;           // first warp of this row (tx > 32)
  const int wn = tx >> 5;                         // warps per row (tx > 32)
  const int tree = tx < 32 ? tx : 32;             // shuffle-tree width
  const int64_t row = (int64_t)blockIdx.x * TY + ty;
  const bool live = row < rows;
  const T* p = x + (live ? row : 0) * rstride;
  const int64_t groups = K / E;                   // K % E == 0 by dispatch (E == 1 otherwise)
  bool nan = false;
  uint32_t err = 0;

  if constexpr (RED == PH_SUM) {
    A s = 0, pos = 0, neg = 0;
    auto fold = [&](const Group<T, E>& g) {
#pragma unroll
      for (int i = 0; i < E; i++) {
        const T v = g.v[i];
        if constexpr (is_float_t<T>::value) s = f_add(s, v);
        else { s += (A)v; if (v > 0) pos += (A)v; else neg += (A)v; }
      }
    };
    if (live) {
      int64_t g0 = lane;
      for (; g0 + (int64_t)(UNROLL - 1) * tx < groups; g0 += (int64_t)UNROLL * tx) {
        Group<T, E> g[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) g[u] = load_group<T, E>(p + (g0 + (int64_t)u * tx) * E);
#pragma unroll
        for (int u = 0; u < UNROLL; u++) fold(g[u]);
      }
      for (; g0 < groups; g0 += tx) fold(load_group<T, E>(p + g0 * E));
    }
    for (int off = tree >> 1; off > 0; off >>= 1) {
      if constexpr (is_float_t<T>::value) s = f_add(s, shfl_xor_t<A>(s, off));
      else { s += shfl_xor_t<A>(s, off); pos += shfl_xor_t<A>(pos, off); neg += shfl_xor_t<A>(neg, off); }
    }
    if (tx > 32) {
      if ((threadIdx.x & 31) == 0) { sh_s[warp] = s; sh_p[warp] = pos; sh_n[warp] = neg; }
      __syncthreads();
      if (lane == 0)
        for (int j = 1; j < wn; j++) {
          if constexpr (is_float_t<T>::value) s = f_add(s, sh_s[w0 + j]);
          else { s += sh_s[w0 + j]; pos += sh_p[w0 + j]; neg += sh_n[w0 + j]; }
        }
    }
    if (lane == 0 && live) {
      if constexpr (!is_float_t<T>::value) {
        // checked fold: overflow at ANY prefix raises.  sum(positives) / sum(negatives) inside T
        // proves no prefix can leave T; otherwise lane 0 replays the row in order (rare).
        const A hi = (A)std::numeric_limits<T>::max(), lo = (A)std::numeric_limits<T>::lowest();
        if (pos > hi || neg < lo) {
          A run = 0;
          for (int64_t k = 0; k < K; k++) { run += (A)p[rev ? K - 1 - k : k]; if (run > hi || run < lo) { err |= PH_FLAG_OVERFLOW; break; } }
        }
      }
      reinterpret_cast<T*>(out)[row] = (T)s;
    }
  } else {
    // ---- the lane's extremum and the batch that gave it
    T m = IS_MAXLIKE ? lowest_of<T>() : highest_of<T>();
    int64_t bg = -1;                              // first group of that batch (-1: nothing seen yet)
    int bn = 0;                                   // groups in it (UNROLL, or 1 in the remainder loop)
    auto consider = [&](T bm, int64_t g0, int n) {
      if constexpr (std::is_same<T, float>::value) nan |= (bm != bm);
      const bool improves = IS_MAXLIKE ? (bm > m) : (bm < m);
      if (improves || (bm == m && (rev || bg < 0))) { m = bm; bg = g0; bn = n; }
    };
    if (live) {
      int64_t g0 = lane;
      for (; g0 + (int64_t)(UNROLL - 1) * tx < groups; g0 += (int64_t)UNROLL * tx) {
        Group<T, E> g[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) g[u] = load_group_plain<T, E>(p + (g0 + (int64_t)u * tx) * E);
        T bm = g[0].v[0];
        if constexpr (is_float_t<T>::value && !std::is_same<T, float>::value) nan |= (bm != bm);
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
#pragma unroll
          for (int i = 0; i < E; i++)
            if (u || i) bm = ext2<T, IS_MAXLIKE>(bm, g[u].v[i], nan);
        consider(bm, g0, UNROLL);
      }
      for (; g0 < groups; g0 += tx) {
        const Group<T, E> g = load_group_plain<T, E>(p + g0 * E);
        T bm = g.v[0];
        if constexpr (is_float_t<T>::value && !std::is_same<T, float>::value) nan |= (bm != bm);
#pragma unroll
        for (int i = 1; i < E; i++) bm = ext2<T, IS_MAXLIKE>(bm, g.v[i], nan);
        consider(bm, g0, 1);
      }
    }
    // ---- the row's extremum M
    T M = m;
    bool dummy = false;
    for (int off = tree >> 1; off > 0; off >>= 1) M = ext2<T, IS_MAXLIKE>(M, __shfl_xor_sync(0xffffffffu, M, off), dummy);
    if (tx > 32) {
      if ((threadIdx.x & 31) == 0) sh_m[warp] = M;
      __syncthreads();
      M = sh_m[w0];
      for (int j = 1; j < wn; j++) M = ext2<T, IS_MAXLIKE>(M, sh_m[w0 + j], dummy);
    }
    // ---- first logical index of M: only lanes whose own extremum IS M look, and only at their winning batch
    int32_t bi = INT32_MAX;
    const bool need_index = IS_ARG || (is_float_t<T>::value && M == (T)0);
    if (live && need_index && bg >= 0 && m == M) {
      for (int u = 0; u < bn; u++) {
        const int64_t gq = bg + (int64_t)u * tx;
        const Group<T, E> g = load_group_plain<T, E>(p + gq * E);
#pragma unroll
        for (int i = 0; i < E; i++) {
          const int32_t phys = (int32_t)(gq * E) + i;
          const int32_t logical = rev ? (int32_t)K - 1 - phys : phys;
          if (g.v[i] == M && logical < bi) bi = logical;
        }
      }
    }
    for (int off = tree >> 1; off > 0; off >>= 1) {
      const int32_t o = __shfl_xor_sync(0xffffffffu, bi, off);
      bi = o < bi ? o : bi;
    }
    if (tx > 32) {
      if ((threadIdx.x & 31) == 0) sh_i[warp] = bi;
      __syncthreads();
      if (lane == 0)
        for (int j = 1; j < wn; j++) bi = sh_i[w0 + j] < bi ? sh_i[w0 + j] : bi;
    }
    if (lane == 0 && live) {
      if constexpr (IS_ARG) reinterpret_cast<int64_t*>(out)[row] = (bi == INT32_MAX) ? 0 : (int64_t)bi;
      else reinterpret_cast<T*>(out)[row] = (need_index && bi != INT32_MAX) ? p[rev ? K - 1 - bi : bi] : M;   // keeps the first zero's sign
    }
  }
  if (err) atomicOr(flags, err);
  if (nan) atomicOr(flags, (uint32_t)PH_FLAG_NAN);
}

// inner == 1, short rows (K <= 32 lanes x G groups x E elements): the whole row lives in the
// registers of TX <= 32 lanes of ONE warp.  Every lane issues all of its G 32-byte loads before
// the first use (no loop, no trip-count arithmetic: the looped kernel above spent ~7 instructions
// per element and was latency-bound at 0.87 of the copy peak), folds them, and the lanes combine
// with shuffles.  arg* / a zero extremum: only lanes whose own extremum equals the row's search
// their registers for the first match (no second pass over memory).
template <typename T, int E, int G, int RED>
__global__ void __launch_bounds__(RED_THREADS) axis_rowreg_kernel(const T* __restrict__ x, void* __restrict__ out,
                                                                  int64_t rows, int64_t K, int64_t rstride, int rev,
                                                                  int tx_log2, uint32_t* __restrict__ flags) {
  constexpr bool IS_MAXLIKE = (RED == PH_MAX || RED == PH_ARGMAX);
  constexpr bool IS_ARG = (RED == PH_ARGMAX || RED == PH_ARGMIN);
  using A = typename Acc<T>::type;
  const int tx = 1 << tx_log2;
  const int lane = threadIdx.x & (tx - 1);
  const int64_t row = (int64_t)blockIdx.x * (RED_THREADS >> tx_log2) + (threadIdx.x >> tx_log2);
  const bool live = row < rows;
  const T* p = x + (live ? row : 0) * rstride;
  const int groups = (int)(K / E);                // K % E == 0 and groups <= 32 * G by dispatch
  Group<T, E> g[G];
  bool has[G];
#pragma unroll
  for (int u = 0; u < G; u++) {
    const int gi = lane + u * tx;
    has[u] = live && gi < groups;
    if (has[u]) g[u] = load_group<T, E>(p + (int64_t)gi * E);
  }
  bool nan = false;
  uint32_t err = 0;
  if constexpr (RED == PH_SUM) {
    A s = 0, pos = 0, neg = 0;
    if constexpr (is_float_t<T>::value) {
      A acc[E];
#pragma unroll
      for (int i = 0; i < E; i++) acc[i] = has[0] ? g[0].v[i] : (T)0;
#pragma unroll
      for (int u = 1; u < G; u++)
        if (has[u]) {
#pragma unroll
          for (int i = 0; i < E; i++) acc[i] = f_add(acc[i], g[u].v[i]);
        }
#pragma unroll
      for (int w = E / 2; w > 0; w >>= 1)
#pragma unroll
        for (int i = 0; i < w; i++) acc[i] = f_add(acc[i], acc[i + w]);
      s = acc[0];
    } else {
#pragma unroll
      for (int u = 0; u < G; u++)
        if (has[u]) {
#pragma unroll
          for (int i = 0; i < E; i++) { const A v = (A)g[u].v[i]; s += v; if (v > 0) pos += v; else neg += v; }
        }
    }
    for (int off = tx >> 1; off > 0; off >>= 1) {
      if constexpr (is_float_t<T>::value) s = f_add(s, shfl_xor_t<A>(s, off));
      else { s += shfl_xor_t<A>(s, off); pos += shfl_xor_t<A>(pos, off); neg += shfl_xor_t<A>(neg, off); }
    }
    if (lane == 0 && live) {
      if constexpr (!is_float_t<T>::value) {
        const A hi = (A)std::numeric_limits<T>::max(), lo = (A)std::numeric_limits<T>::lowest();
        if (pos > hi || neg < lo) {               // some prefix might leave T: replay the row in order (rare)
          A run = 0;
          for (int64_t k = 0; k < K; k++) { run += (A)p[rev ? K - 1 - k : k]; if (run > hi || run < lo) { err |= PH_FLAG_OVERFLOW; break; } }
        }
      }
      reinterpret_cast<T*>(out)[row] = (T)s;
    }
  } else {
    T m = IS_MAXLIKE ? lowest_of<T>() : highest_of<T>();
#pragma unroll
    for (int u = 0; u < G; u++)
      if (has[u]) {
#pragma unroll
        for (int i = 0; i < E; i++) m = ext2<T, IS_MAXLIKE>(m, g[u].v[i], nan);
      }
    T M = m;
    for (int off = tx >> 1; off > 0; off >>= 1) M = ext2<T, IS_MAXLIKE>(M, __shfl_xor_sync(0xffffffffu, M, off), nan);
    if constexpr (std::is_same<T, float>::value) nan |= (M != M);
    int32_t bi = INT32_MAX;
    T zv = M;
    const bool need_index = IS_ARG || (is_float_t<T>::value && M == (T)0);
    // rows sharing a warp (TX < 32) may disagree on need_index: branch on a warp-uniform vote so
    // the shuffles below are executed by all 32 lanes
    const bool warp_need = IS_ARG ? true : (__any_sync(0xffffffffu, need_index) != 0);
    if (warp_need) {
      if (need_index && m == M) {                 // NaN rows never match
#pragma unroll
        for (int u = 0; u < G; u++)
#pragma unroll
          for (int i = 0; i < E; i++) {
            const int32_t phys = (lane + u * tx) * E + i;
            const int32_t logical = rev ? (int32_t)K - 1 - phys : phys;
            if (has[u] && g[u].v[i] == M && logical < bi) { bi = logical; zv = g[u].v[i]; }
          }
      }
      for (int off = tx >> 1; off > 0; off >>= 1) {
        const int32_t ob = __shfl_xor_sync(0xffffffffu, bi, off);
        const T oz = __shfl_xor_sync(0xffffffffu, zv, off);
        if (ob < bi) { bi = ob; zv = oz; }
      }
    }
    if (lane == 0 && live) {
      if constexpr (IS_ARG) reinterpret_cast<int64_t*>(out)[row] = (bi == INT32_MAX) ? 0 : (int64_t)bi;
      else reinterpret_cast<T*>(out)[row] = (need_index && bi != INT32_MAX) ? zv : M;   // keeps the first zero's sign
    }
  }
  if (err) atomicOr(flags, err);
  if (nan) atomicOr(flags, (uint32_t)PH_FLAG_NAN);
}

// ---------------------------------------------------------------- host side
static bool desc_is_contiguous(const ph_desc* d, int64_t& total) {
  total = 1;
  int64_t expect = 1;
  bool ok = true;
  for (int i = d->rank - 1; i >= 0; i--) {
    if (d->extent[i] != 1 && d->stride[i] != expect) ok = false;
    expect *= d->extent[i];
    total *= d->extent[i];
  }
  if (d->rank == 0) total = 0;
  return ok;
}

// Can the full-reduction kernels read this view IN PLACE?  Yes when it coalesces to rows of C contiguous
// elements at a constant (any sign) row stride and C is a whole number of tiles; `vec` = the 32-byte form
// is usable (aligned base and stride).  Contiguous arrays never come here (desc_is_contiguous first).
template <typename T>
static bool rows_layout(const void* a, const ph_desc* d, const T** x, int64_t& n, RowsArgs& rows, bool& vec) {
  static const bool off = getenv("PH_REDUCE_GATHER") != nullptr;           // A/B knob: always gather first
  if (off) return false;
  int64_t total;
  if (desc_is_contiguous(d, total) || total == 0) return false;
  // coalesce: drop extent-1 axes, merge axis j into j+1 when stride[j] == stride[j+1] * extent[j+1]
  int64_t ext[PH_MAX_RANK], str[PH_MAX_RANK];
  int rk = 0;
  for (int i = 0; i < d->rank; i++) {
    if (d->extent[i] == 1) continue;
    if (rk > 0 && str[rk - 1] == d->stride[i] * d->extent[i]) { ext[rk - 1] *= d->extent[i]; str[rk - 1] = d->stride[i]; }
    else { ext[rk] = d->extent[i]; str[rk] = d->stride[i]; rk++; }
  }
  if (rk != 2 || str[1] != 1) return false;
  constexpr int E32 = 32 / (int)sizeof(T);
  const int64_t R = ext[0], C = ext[1], S = str[0];
  const T* base = reinterpret_cast<const T*>(a) + d->offset;
  const int64_t tile32 = (int64_t)RED_THREADS * E32 * 4, tile1 = (int64_t)RED_THREADS * 4;
  if (E32 > 1 && C % tile32 == 0 && S % E32 == 0 && ((uintptr_t)base % 32) == 0) { vec = true; rows.tiles_per_row = (uint32_t)(C / tile32); }
  else if (C % tile1 == 0) { vec = false; rows.tiles_per_row = (uint32_t)(C / tile1); }
  else return false;
  if (C / tile1 > 0x7fffffffLL || R * (C / tile1) > 0x7fffffffLL) { rows.tiles_per_row = 0; return false; }
  rows.row_stride = S;
  *x = base;
  n = R * C;
  return true;
}

// returns a contiguous device pointer holding the described region (gathers if needed)
template <typename T>
static int32_t contiguous_input(const void* a, const ph_desc* d, const T** out_ptr, void** temp, int64_t& total) {
  *temp = nullptr;
  if (desc_is_contiguous(d, total)) {
    *out_ptr = reinterpret_cast<const T*>(a) + d->offset;
    return PH_OK;
  }
  if (total == 0) { *out_ptr = reinterpret_cast<const T*>(a); return PH_OK; }
  PH_CUDA(cudaMallocAsync(temp, (size_t)total * sizeof(T), rt().stream));
  ph_desc cd;
  memset(&cd, 0, sizeof(cd));
  cd.rank = d->rank;
  int64_t acc = 1;
  for (int i = d->rank - 1; i >= 0; i--) { cd.extent[i] = d->extent[i]; cd.stride[i] = acc; acc *= d->extent[i]; }
  int32_t st = ph_copy_strided((int32_t)sizeof(T), a, d, *temp, &cd);
  if (st != PH_OK) return st;
  *out_ptr = reinterpret_cast<const T*>(*temp);
  return PH_OK;
}

// one zero-initialised device word (allocated with the flag word by ph_init, so it follows the
// device); the last block of every fused reduction resets it
static unsigned int* reduce_ticket() { return rt().d_flags ? rt().d_flags + 8 : nullptr; }

// `cmb` == nullptr: single-GPU reduction, result left on the device (out_value_dev / out_index_dev).
// `cmb` with nranks > 1: this rank's part of a sharded reduction -- the kernel also runs for an EMPTY shard
// (one block contributing the identity), so every rank enters the combine; the result goes to cmb->host_out.
template <typename T>
int32_t reduce_full_t(int32_t red, const void* a, const ph_desc* d, void* out_value_dev, int64_t* out_index_dev,
                      const CombineArgs* cmb_in) {
  Runtime& r = rt();
  const CombineArgs cmb = cmb_in ? *cmb_in : CombineArgs();
  const bool sharded = cmb.host_out != nullptr;       // record mode: always launch, result to the pinned host record
  const T* x;
  void* temp = nullptr;
  int64_t n;
  int32_t st = PH_OK;
  RowsArgs rows;
  bool rows_vec = false;
  if (rows_layout<T>(a, d, &x, n, rows, rows_vec)) {
    // a row-strided view whose rows are whole tiles: read in place (no gather, no temporary)
  } else {
    st = contiguous_input<T>(a, d, &x, &temp, n);
    if (st != PH_OK) return st;
  }
  // persistent grid: exactly the blocks that are resident at once (a partial second wave would
  // leave the GPU mostly idle while it runs), fewer for small inputs
  static int resident_sum = 0, resident_ext = 0, resident_dev = -1;
  if (resident_dev != r.device) {
    resident_dev = r.device;
    int a = 0, b = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, sum_partial_kernel<T, 32 / (int)sizeof(T)>, RED_THREADS, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, ext_partial_kernel<T, 32 / (int)sizeof(T), true>, RED_THREADS, 0);
    resident_sum = std::max(1, a);
    resident_ext = std::max(1, b);
  }
  const int per_sm = (red == PH_SUM) ? resident_sum : resident_ext;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)r.sm_count * per_sm,
                                                               ceil_div(n, (int64_t)RED_THREADS * 32)));
  st = ensure_scratch((size_t)grid * 64 + 256);
  if (st != PH_OK) return st;
  int* status = reinterpret_cast<int*>(reinterpret_cast<char*>(r.d_scratch) + (size_t)grid * 64);
  unsigned int* ticket = reduce_ticket();
  if (!ticket) return set_error(PH_ERR_CUDA, "cannot allocate the reduction ticket");
  constexpr int E32 = 32 / (int)sizeof(T);
  const bool al32 = rows.tiles_per_row ? rows_vec : ((uintptr_t)x % 32) == 0;
  if (red == PH_SUM) {
    if (n == 0 && !sharded) {                             // Enumerable#sum of nothing is T.zero
      PH_CUDA(cudaMemsetAsync(out_value_dev, 0, sizeof(T), r.stream));
    } else {
      SumState<T>* parts = reinterpret_cast<SumState<T>*>(r.d_scratch);
      static_assert(sizeof(SumState<T>) <= 64, "partial too large");
      T* ov = reinterpret_cast<T*>(out_value_dev);
      if (al32) sum_partial_kernel<T, E32><<<grid, RED_THREADS, 0, r.stream>>>(x, n, parts, ov, status, ticket, r.d_flags, cmb, rows);
      else sum_partial_kernel<T, 1><<<grid, RED_THREADS, 0, r.stream>>>(x, n, parts, ov, status, ticket, r.d_flags, cmb, rows);
      PH_LAUNCH_CHECK("sum_partial_kernel");
      if constexpr (!is_float_t<T>::value) {
        if (!sharded) {                                   // (sharded: the caller decides from the combined record)
          int* h = reinterpret_cast<int*>(r.h_scratch);
          PH_CUDA(cudaMemcpyAsync(h, status, sizeof(int), cudaMemcpyDeviceToHost, r.stream));
          PH_CUDA(cudaStreamSynchronize(r.stream));
          int code = *h;
          if (code == PH_RED_NEED_EXACT) {                // the filter could not decide: exact ordered pass
            if (rows.tiles_per_row) {                       // (the ordered pass walks a contiguous buffer)
              if ((st = contiguous_input<T>(a, d, &x, &temp, n)) != PH_OK) return st;
            }
            Prefix<T>* pp = reinterpret_cast<Prefix<T>*>(r.d_scratch);
            static_assert(sizeof(Prefix<T>) <= 64, "partial too large");
            sum_exact_kernel<T><<<grid, RED_THREADS, 0, r.stream>>>(x, n, pp);
            PH_LAUNCH_CHECK("sum_exact_kernel");
            sum_exact_final_kernel<T><<<1, 32, 0, r.stream>>>(pp, grid, status, CombineArgs());
            PH_LAUNCH_CHECK("sum_exact_final_kernel");
            PH_CUDA(cudaMemcpyAsync(h, status, sizeof(int), cudaMemcpyDeviceToHost, r.stream));
            PH_CUDA(cudaStreamSynchronize(r.stream));
            code = *h;
          }
          if (code == PH_RED_OVERFLOW) {
            set_flag_kernel<<<1, 1, 0, r.stream>>>(r.d_flags, (uint32_t)PH_FLAG_OVERFLOW);
            PH_LAUNCH_CHECK("set_flag_kernel");
          }
        }
      }
    }
    if (out_index_dev) PH_CUDA(cudaMemsetAsync(out_index_dev, 0, 8, r.stream));
  } else if (red == PH_MIN || red == PH_MAX || red == PH_ARGMIN || red == PH_ARGMAX) {
    if (n == 0 && !sharded) {                             // host raises Enumerable::EmptyError
      PH_CUDA(cudaMemsetAsync(out_value_dev, 0, sizeof(T), r.stream));
      if (out_index_dev) PH_CUDA(cudaMemsetAsync(out_index_dev, 0xff, 8, r.stream));
    } else {
      Cand<T>* parts = reinterpret_cast<Cand<T>*>(r.d_scratch);
      const bool is_max = (red == PH_MAX || red == PH_ARGMAX);
      T* ov = reinterpret_cast<T*>(out_value_dev);
      if (is_max) {
        if (al32) ext_partial_kernel<T, E32, true><<<grid, RED_THREADS, 0, r.stream>>>(x, n, parts, r.d_flags, ov, out_index_dev, ticket, cmb, rows);
        else ext_partial_kernel<T, 1, true><<<grid, RED_THREADS, 0, r.stream>>>(x, n, parts, r.d_flags, ov, out_index_dev, ticket, cmb, rows);
      } else {
        if (al32) ext_partial_kernel<T, E32, false><<<grid, RED_THREADS, 0, r.stream>>>(x, n, parts, r.d_flags, ov, out_index_dev, ticket, cmb, rows);
        else ext_partial_kernel<T, 1, false><<<grid, RED_THREADS, 0, r.stream>>>(x, n, parts, r.d_flags, ov, out_index_dev, ticket, cmb, rows);
      }
      PH_LAUNCH_CHECK("ext_partial_kernel");
    }
  } else {
    if (temp) cudaFreeAsync(temp, r.stream);
    return set_error(PH_ERR_INVALID, "unknown reduction %d", red);
  }
  if (temp) PH_CUDA(cudaFreeAsync(temp, r.stream));
  return PH_OK;
}

// Full reduction of an array sharded along axis 0 over the ranks of the communicator (collective: every
// rank calls it with its shard).  One launch per rank with the in-kernel combine when the peers' control
// blocks are mapped; otherwise the records are allgathered with NCCL and folded by a second tiny launch.
// The result (identical on every rank) is read from the pinned host record after ONE synchronisation.
template <typename T>
int32_t reduce_full_sharded_t(int32_t red, const void* a, const ph_desc* d, int64_t elems_before, void* out_value_host,
                              int64_t* out_index_host, uint32_t* out_flags) {
  Runtime& r = rt();
  PeerInfo& pi = peers();
  CombineArgs cmb;
  int32_t st = comm_combine_args(&cmb, elems_before);
  if (st != PH_OK) return st;
  const bool nccl = cmb.nranks > 1 && cmb.peer_slots[0] == nullptr;      // records travel by NCCL allgather
  const bool is_max = (red == PH_MAX || red == PH_ARGMAX);
  st = reduce_full_t<T>(red, a, d, nullptr, nullptr, &cmb);
  if (st != PH_OK) return st;
  if (nccl) {
    if ((st = comm_allgather_records(r.stream)) != PH_OK) return st;
    if (red == PH_SUM) sum_combine_kernel<T><<<1, 32, 0, r.stream>>>(pi.gather_recv, cmb.nranks, nullptr, nullptr, cmb.host_out, cmb.seq);
    else if (is_max) ext_combine_kernel<T, true><<<1, 32, 0, r.stream>>>(pi.gather_recv, cmb.nranks, nullptr, nullptr, cmb.host_out, cmb.seq);
    else ext_combine_kernel<T, false><<<1, 32, 0, r.stream>>>(pi.gather_recv, cmb.nranks, nullptr, nullptr, cmb.host_out, cmb.seq);
    PH_LAUNCH_CHECK("combine_kernel");
  }
  // The finishing block writes the pinned record and then its call number: the host polls that word (the
  // result is visible ~1 us after the store) instead of paying a stream synchronisation's wake-up; a stream
  // query every few thousand polls surfaces a faulted launch
  {
    const volatile uint32_t* done = &pi.host_result->seq;
    uint32_t polls = 0;
    while (*done != cmb.seq) {
      if ((++polls & 0xfff) == 0) {
        cudaError_t q = cudaStreamQuery(r.stream);
        if (q != cudaErrorNotReady) {                      // finished (or failed) without our word: settle it the slow way
          PH_CUDA(cudaStreamSynchronize(r.stream));
          if (*done != cmb.seq) return set_error(PH_ERR_CUDA, "sharded reduction finished without writing its result record");
          break;
        }
      }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
  }
  ReduceResult res = *pi.host_result;
  if constexpr (!is_float_t<T>::value) {
    if (red == PH_SUM && res.status == PH_RED_NEED_EXACT) {
      // rare: some prefix of the global fold MIGHT leave T.  Every rank sees the same status, so all of them
      // take this branch: ordered (sum, max prefix, min prefix) pass per shard, triples folded in rank order.
      const T* x;
      void* temp;
      int64_t n;
      if ((st = contiguous_input<T>(a, d, &x, &temp, n)) != PH_OK) return st;
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)r.sm_count * 4, ceil_div(n, (int64_t)RED_THREADS * 32)));
      if ((st = ensure_scratch((size_t)grid * 64 + 256)) != PH_OK) return st;
      Prefix<T>* pp = reinterpret_cast<Prefix<T>*>(r.d_scratch);
      sum_exact_kernel<T><<<grid, RED_THREADS, 0, r.stream>>>(x, n, pp);
      PH_LAUNCH_CHECK("sum_exact_kernel");
      CombineArgs c2;
      if ((st = comm_combine_args(&c2, elems_before)) != PH_OK) return st;
      if (c2.nranks > 1 && c2.peer_slots[0] == nullptr) {
        sum_exact_record_kernel<T><<<1, 32, 0, r.stream>>>(pp, grid, c2);
        PH_LAUNCH_CHECK("sum_exact_record_kernel");
        if ((st = comm_allgather_records(r.stream)) != PH_OK) return st;
        c2.my_slots = pi.gather_recv;
      }
      sum_exact_final_kernel<T><<<1, 32, 0, r.stream>>>(pp, grid, nullptr, c2);
      PH_LAUNCH_CHECK("sum_exact_final_kernel");
      if (temp) PH_CUDA(cudaFreeAsync(temp, r.stream));
      PH_CUDA(cudaStreamSynchronize(r.stream));
      res.status = pi.host_result->status;
    }
  }
  if (res.status == PH_RED_TIMEOUT)
    return set_error(PH_ERR_CUDA, "sharded reduction: a peer rank never delivered its partial (is every rank calling?)");
  uint32_t fl = res.flags;
  if (res.status == PH_RED_OVERFLOW) fl |= PH_FLAG_OVERFLOW;
  if (out_flags) *out_flags = fl;
  memcpy(out_value_host, res.value, sizeof(T));
  if (out_index_host) *out_index_host = (red == PH_SUM) ? 0 : (res.index == INT64_MAX ? -1 : res.index);
  return PH_OK;
}

// The view's layout as the kernels see it: element (o, k, c) at x[o * ostride + k * kstride + c * istride].
// inner > 1 needs istride == 1 (threads run along the contiguous inner axis); inner == 1 (the reduced axis is
// the innermost one) needs kstride == +1 or -1 (rev): rows of K elements, row o at x + o * ostride.
struct AxisLayout {
  int64_t ostride, kstride;
  int rev;                // inner == 1: the reduced (innermost) axis runs backwards
  int irev = 0;           // inner > 1: the inner axis runs backwards (x is the address of LOGICAL column 0)
};

template <typename T, int RED>
static int32_t reduce_axis_launch(const T* x, void* out, int64_t outer, int64_t K, int64_t inner, const AxisLayout& lay) {
  Runtime& r = rt();
  if (inner > 1) {
    constexpr int E32 = 32 / (int)sizeof(T), E16 = 16 / (int)sizeof(T);
    constexpr bool ARG = (RED == PH_ARGMAX || RED == PH_ARGMIN);
    const uintptr_t xa = (uintptr_t)(lay.irev ? x - (inner - 1) : x), oa = (uintptr_t)out;      // lowest address of a row
    const bool can32 = inner % E32 == 0 && xa % 32 == 0 && (ARG || oa % 32 == 0) && lay.ostride % E32 == 0 && lay.kstride % E32 == 0;
    const bool can16 = inner % E16 == 0 && xa % 16 == 0 && (ARG || oa % 16 == 0) && lay.ostride % E16 == 0 && lay.kstride % E16 == 0;
    // Plan = (group width E, groups in flight per thread U, block size), measured on the axis-0
    // shards of the 1e9-element f32 array ([1000/N, 1000, 1000], benchmarks/probe_axis.py):
    //   * outer == 1 (the reduced axis is the leading one; consecutive k are `inner` apart): the widest
    //     group with 4 in flight streams at 6.4-7.1 TB/s down to [125, 1000, 1000];
    //   * outer > 1 (a middle axis: every `outer` index is its own K x inner panel): with many columns the
    //     widest group wins, 8 in flight in 128-thread blocks (6.8 TB/s, argmax 6.5 -> 6.75); with few
    //     columns only MORE THREADS keep the memory system busy -- 32-byte groups leave 15 625 threads on
    //     [125, 1000, 1000] and reach 2.1 TB/s however deep the unrolling, scalar columns 5.4 TB/s.
    const int64_t cols = outer * inner;
    const int64_t col_bytes = cols * (int64_t)sizeof(T);
    // few columns: the cp.async-staged kernel (one warp per 32-column strip); needs 16-byte aligned strip rows
    // and enough strips to occupy the machine.  PH_AXIS_STAGED=0 / 1 forces the choice (A/B runs).
    if constexpr (sizeof(T) >= 4) {
      constexpr int EPC = 16 / (int)sizeof(T);
      static const int force_staged = getenv("PH_AXIS_STAGED") ? atoi(getenv("PH_AXIS_STAGED")) : -1;
      const int64_t spo = ceil_div(inner, (int64_t)32);
      const int64_t strips = outer * spo;
      const bool aligned = xa % 16 == 0 && inner % EPC == 0 && (lay.ostride * (int64_t)sizeof(T)) % 16 == 0 &&
                           (lay.kstride * (int64_t)sizeof(T)) % 16 == 0;
      const bool wanted = force_staged >= 0 ? force_staged != 0
                                            : (col_bytes < (3LL << 19) && strips >= (int64_t)r.sm_count && K >= 64);
      if (aligned && wanted && strips <= 0x7fffffffLL && spo <= 0x7fffffffLL && K <= 0x7fffffffLL) {
        // 4 KB tiles, 7 in flight per warp: bytes in flight, not instructions, bound this kernel (f32 with 16-row
        // = 2 KB tiles: 7 MB in flight over the GPU, 4.1 TB/s on [16384,16384]; Little's law wants ~5 MB at 6.5 TB/s
        // and ~0.8 us, with little margin)
        constexpr int KT = sizeof(T) >= 8 ? 16 : 32, STAGES = 8;
        axis_strip_staged_kernel<T, RED, KT, STAGES><<<(unsigned)strips, 32, 0, r.stream>>>(x, out, K, inner, lay.ostride,
                                                                                             lay.kstride, lay.irev, (int)spo, r.d_flags);
        PH_LAUNCH_CHECK("axis_strip_staged_kernel");
        return PH_OK;
      }
    }
    int e = 1, u = 4, block = RED_THREADS;
    if (outer == 1) {
      if (can32 && E32 > 1 && cols / E32 >= (int64_t)r.sm_count * 64) e = E32;
      else if (can16 && E16 > 1 && cols / E16 >= (int64_t)r.sm_count * 64) e = E16;
      u = col_bytes * 4 >= (6LL << 20) ? 4 : 16;
      // 8 groups in flight in 128-thread blocks: argmax 6.26 -> 6.70 TB/s, sum / max 6.64 / 6.59 -> 6.81 / 6.73 on
      // [1000,1000,1000] f32 (256-thread blocks x 8 or x 16 are slower: 5.2 / 5.1 for argmax)
      if (e == E32 && u == 4) { u = 8; block = 128; }
    } else if (col_bytes >= (3LL << 20) && can32 && E32 > 1) {
      e = E32; u = 8; block = 128;
    } else if (col_bytes >= (3LL << 19) && can32 && E32 > 1) {
      e = E32; u = 4;
    } else if (ARG && col_bytes >= (3LL << 18) && can16 && E16 > 1) {
      e = E16; u = 8;
    } else {
      e = 1; u = (!ARG && col_bytes >= (3LL << 18)) ? 4 : 16;
    }
    static const int force_e = getenv("PH_AXIS_E") ? atoi(getenv("PH_AXIS_E")) : 0;       // tuning knobs
    static const int force_u = getenv("PH_AXIS_U") ? atoi(getenv("PH_AXIS_U")) : 0;
    static const int force_b = getenv("PH_AXIS_BLOCK") ? atoi(getenv("PH_AXIS_BLOCK")) : 0;
    if (force_e == 32 && can32 && E32 > 1) e = E32;
    else if (force_e == 16 && can16 && E16 > 1) e = E16;
    else if (force_e == 1) e = 1;
    if (force_u) u = force_u;
    if (sizeof(T) < 4 && u > 8) u = 8;                       // 32 one-byte accumulators + 16 groups would spill
    const int64_t threads = cols / e;
    if (force_b) block = force_b;
    const int64_t blocks = ceil_div(threads, (int64_t)block);
    if (blocks > 0x7fffffffLL) return set_error(PH_ERR_INVALID, "array too large for one launch");
#define PH_STRIP(EE, UU) axis_strip_kernel<T, EE, RED, UU><<<(unsigned)blocks, block, 0, r.stream>>>(x, out, outer, K, inner, lay.ostride, lay.kstride, lay.irev, r.d_flags)
    if (e == E32 && E32 > 1) { if (u >= 16) PH_STRIP(E32, 16); else if (u >= 8) PH_STRIP(E32, 8); else PH_STRIP(E32, 4); }
    else if (e == E16 && E16 > 1) { if (u >= 8) PH_STRIP(E16, 8); else PH_STRIP(E16, 4); }
    else { if (u >= 16) PH_STRIP(1, 16); else PH_STRIP(1, 4); }
#undef PH_STRIP
    PH_LAUNCH_CHECK("axis_strip_kernel");
    return PH_OK;
  }
  // last axis: rows of K contiguous elements
  constexpr int E32 = 32 / (int)sizeof(T);
  const bool vec = E32 > 1 && K % E32 == 0 && ((uintptr_t)x % 32) == 0 && lay.ostride % E32 == 0;
  const int e = vec ? E32 : 1;
  const int64_t groups = K / e;
  if (vec && groups <= 32 * 8) {                  // the row fits the registers of one warp
    const int G = groups <= 32 * 4 ? 4 : 8;
    int tx = 1, lg = 0;
    while (tx < 32 && (int64_t)tx * G < groups) { tx <<= 1; lg++; }
    const int64_t blocks = ceil_div(outer, (int64_t)(RED_THREADS / tx));
    if (blocks > 0x7fffffffLL) return set_error(PH_ERR_INVALID, "array too large for one launch");
    if (G == 4) axis_rowreg_kernel<T, E32, 4, RED><<<(unsigned)blocks, RED_THREADS, 0, r.stream>>>(x, out, outer, K, lay.ostride, lay.rev, lg, r.d_flags);
    else axis_rowreg_kernel<T, E32, 8, RED><<<(unsigned)blocks, RED_THREADS, 0, r.stream>>>(x, out, outer, K, lay.ostride, lay.rev, lg, r.d_flags);
    PH_LAUNCH_CHECK("axis_rowreg_kernel");
    return PH_OK;
  }
  int tx = 1, lg = 0;
  while (tx < RED_THREADS && (int64_t)tx * 8 < groups) { tx <<= 1; lg++; }   // <= 8 groups per lane
  const int ty = RED_THREADS / tx;
  const int64_t blocks = ceil_div(outer, ty);
  if (blocks > 0x7fffffffLL) return set_error(PH_ERR_INVALID, "array too large for one launch");
  if (vec) axis_row_kernel<T, E32, RED><<<(unsigned)blocks, RED_THREADS, 0, r.stream>>>(x, out, outer, K, lay.ostride, lay.rev, tx, lg, r.d_flags);
  else axis_row_kernel<T, 1, RED><<<(unsigned)blocks, RED_THREADS, 0, r.stream>>>(x, out, outer, K, lay.ostride, lay.rev, tx, lg, r.d_flags);
  PH_LAUNCH_CHECK("axis_row_kernel");
  return PH_OK;
}

template <typename T>
int32_t reduce_axis_t(int32_t red, const void* a, const ph_desc* d, int32_t axis, void* out,
                             const ph_desc* od) {
  Runtime& r = rt();
  if (axis < 0 || axis >= d->rank) return set_error(PH_ERR_INVALID, "axis %d out of range for rank %d", axis, d->rank);
  int64_t ototal;
  if (!od || !desc_is_contiguous(od, ototal) )
    return set_error(PH_ERR_UNSUPPORTED, "ph_reduce_axis writes a contiguous output");
  const T* x = nullptr;
  void* temp = nullptr;
  int64_t n = 0;
  int32_t st = PH_OK;
  int64_t outer = 1, inner = 1;
  const int64_t K = d->extent[axis];
  for (int i = 0; i < axis; i++) outer *= d->extent[i];
  for (int i = axis + 1; i < d->rank; i++) inner *= d->extent[i];
  if (outer * inner == 0 || K == 0) return PH_OK;
  // Read the view IN PLACE when its outer axes coalesce to one stride, its inner axes to one unit-stride
  // axis (or, for the innermost axis, the reduced axis itself has stride +1 / -1): row-strided slices,
  // column blocks, reversed views (`view.reverse.max(axis: 1)`).  Anything else is gathered first.
  AxisLayout lay{K * inner, inner, 0};
  bool in_place = false;
  {
    static const bool off = getenv("PH_REDUCE_GATHER") != nullptr;         // A/B knob: always gather first
    auto coalesce = [&](int lo, int hi, int64_t& stride) -> bool {        // axes [lo, hi) as ONE axis of `stride`
      bool have = false;
      int64_t s_prev = 0, e_prev = 0;
      for (int i = hi - 1; i >= lo; i--) {
        if (d->extent[i] == 1) continue;
        if (!have) { stride = d->stride[i]; have = true; }
        else if (d->stride[i] != s_prev * e_prev) return false;
        s_prev = d->stride[i]; e_prev = d->extent[i];
      }
      if (!have) stride = 0;
      return true;
    };
    int64_t total;
    int64_t os = 0, is = 0;
    if (!off && !desc_is_contiguous(d, total) && coalesce(0, axis, os) && coalesce(axis + 1, d->rank, is)) {
      const int64_t ks = d->stride[axis];
      if (inner > 1 && (is == 1 || is == -1)) { lay = AxisLayout{os, ks, 0, is == -1}; in_place = true; }
      else if (inner == 1 && (ks == 1 || ks == -1 || K == 1)) { lay = AxisLayout{os, 1, ks == -1 && K > 1}; in_place = true; }
    }
  }
  if (in_place) {
    x = reinterpret_cast<const T*>(a) + d->offset;
    if (lay.rev) x -= (K - 1);                    // rows are addressed by their lowest element
  } else {
    st = contiguous_input<T>(a, d, &x, &temp, n);
    if (st != PH_OK) return st;
  }
  const bool arg = (red == PH_ARGMAX || red == PH_ARGMIN);
  void* o = arg ? (void*)(reinterpret_cast<int64_t*>(out) + od->offset) : (void*)(reinterpret_cast<T*>(out) + od->offset);
  switch (red) {
    case PH_SUM: st = reduce_axis_launch<T, PH_SUM>(x, o, outer, K, inner, lay); break;
    case PH_MIN: st = reduce_axis_launch<T, PH_MIN>(x, o, outer, K, inner, lay); break;
    case PH_MAX: st = reduce_axis_launch<T, PH_MAX>(x, o, outer, K, inner, lay); break;
    case PH_ARGMAX: st = reduce_axis_launch<T, PH_ARGMAX>(x, o, outer, K, inner, lay); break;
    case PH_ARGMIN: st = reduce_axis_launch<T, PH_ARGMIN>(x, o, outer, K, inner, lay); break;
    default: st = set_error(PH_ERR_INVALID, "unknown reduction %d", red);
  }
  if (temp) cudaFreeAsync(temp, r.stream);
  return st;
}

}  // namespace ph

