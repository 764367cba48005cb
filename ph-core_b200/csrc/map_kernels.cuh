// map_kernels.cuh -- the elementwise "map" kernel family (K1/K2/K3 of SURVEY.md 2.3).
//
// One functor F describes an N-input, 1-output element operation:
//     using In = ...; using Out = ...; static constexpr int NIN = ...;
//     static __device__ Out apply(const In (&x)[NIN], uint32_t& err);
// and launch_map<F>() picks one of three kernels from the coalesced plan:
//   flat : every operand is contiguous (or a broadcast scalar)      -> 256-bit vectors
//   rows : innermost axis contiguous / broadcast, outer axes strided -> 256-bit vectors,
//          row-vector operands are loaded once per block and kept in registers
//   any  : arbitrary strides -> one element per thread per step
// This replaces MultiIndexable.each_with's per-element coordinate loop
// (src/multi_indexable.cr:1080-1102) and NArray#map (src/n_array.cr:589-595).
#pragma once
#include "ph_common.cuh"
#include <algorithm>
#include <stdlib.h>

namespace ph {

constexpr int MAP_THREADS = 256;

enum OperandMode : int { OPND_ARRAY = 0, OPND_BCAST = 1, OPND_PARAM = 2, OPND_ROWVEC = 3, OPND_PERIODIC = 4 };

template <int NIN>
struct MapArgs {
  const void* in[NIN];      // element 0 of each operand (offset already applied)
  uint64_t param[NIN];      // scalar bits for OPND_PARAM
  int mode[NIN];
  void* out;
  int64_t n;                // flat: elements; rows: inner extent
  int64_t rows;             // rows: number of rows
  int rows_per_block;
  int all_array;            // every input is OPND_ARRAY
  uint32_t period;          // flat: OPND_PERIODIC operands repeat every `period` elements (0 = none)
  int reverse;              // flat: visit the tiles from the last to the first (see launch_flat)
  int l2_hint;              // flat: 1 = inputs evict_first; 2 = + results evict_last in L2 (see launch_flat)
  int64_t gx;               // rows: number of column tiles (1-D grid = gx * slabs * chunks)
  uint32_t chunks;          // rows: row chunks of the last outer axis per slab
  uint32_t slabs;           // rows: product of the leading outer extents
  int slab_fastest;         // rows: consecutive blocks visit consecutive slabs (see launch_rows)
  int tx, tx_log2;          // rows: threads along the inner axis (power of two <= 256)
  uint32_t* flags;
  OuterAxes outer;          // rows/any: operand k uses stride[k], the output uses stride[NIN]
  // rows kernel with E == 1: innermost strides (any value, incl. negative / zero)
  int64_t inner_stride[NIN + 1];
};

template <typename F, int E>
__device__ __forceinline__ void apply_group(const Group<typename F::In, E> (&x)[F::NIN],
                                            Group<typename F::Out, E>& y, uint32_t& err) {
#pragma unroll
  for (int i = 0; i < E; i++) {
    typename F::In v[F::NIN];
#pragma unroll
    for (int k = 0; k < F::NIN; k++) v[k] = x[k].v[i];
    y.v[i] = F::apply(v, err);
  }
}

// ------------------------------------------------------------------ flat
template <typename F, int E, int UNROLL>
__global__ void __launch_bounds__(MAP_THREADS) map_flat_kernel(const MapArgs<F::NIN> a) {
  using In = typename F::In;
  using Out = typename F::Out;
  constexpr int NIN = F::NIN;
  const int64_t tile = (int64_t)MAP_THREADS * E * UNROLL;
  const int64_t base = (int64_t)(a.reverse ? gridDim.x - 1 - blockIdx.x : blockIdx.x) * tile;
  uint32_t err = 0;
  Out* __restrict__ out = reinterpret_cast<Out*>(a.out);

  In scalar[NIN];
#pragma unroll
  for (int k = 0; k < NIN; k++) {
    if (a.mode[k] == OPND_PARAM) scalar[k] = bits_to<In>(a.param[k]);
    else if (a.mode[k] == OPND_BCAST) scalar[k] = *reinterpret_cast<const In*>(a.in[k]);
    else scalar[k] = In();
  }

  const bool hint_in = a.l2_hint >= 1, hint_out = a.l2_hint >= 2;
  if (base + tile <= a.n) {
    Group<In, E> x[UNROLL][NIN];
    if (a.all_array) {                       // hot path: no per-operand predicates
#pragma unroll
      for (int u = 0; u < UNROLL; u++) {
        const int64_t idx = base + ((int64_t)u * MAP_THREADS + threadIdx.x) * E;
#pragma unroll
        for (int k = 0; k < NIN; k++) x[u][k] = load_group_hint<In, E>(reinterpret_cast<const In*>(a.in[k]) + idx, hint_in);
      }
    } else {
#pragma unroll
      for (int u = 0; u < UNROLL; u++) {
        const int64_t idx = base + ((int64_t)u * MAP_THREADS + threadIdx.x) * E;
        // a row vector broadcast over contiguous rows is a flat array with a periodic operand
        // (period % E == 0, so a group never wraps); it stays in L1 (cached load, no streaming hint)
        uint32_t pcol = 0;
        if (a.period) pcol = (a.n >> 32) ? (uint32_t)(idx % (int64_t)a.period) : (uint32_t)idx % a.period;
#pragma unroll
        for (int k = 0; k < NIN; k++) {
          if (a.mode[k] == OPND_ARRAY) x[u][k] = load_group_hint<In, E>(reinterpret_cast<const In*>(a.in[k]) + idx, hint_in);
          else if (a.mode[k] == OPND_PERIODIC) x[u][k] = load_group_plain<In, E>(reinterpret_cast<const In*>(a.in[k]) + pcol);
          else x[u][k] = splat_group<In, E>(scalar[k]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const int64_t idx = base + ((int64_t)u * MAP_THREADS + threadIdx.x) * E;
      Group<Out, E> y;
      apply_group<F, E>(x[u], y, err);
      store_group_hint<Out, E>(out + idx, y, hint_out);
    }
  } else {
    for (int64_t i = base + threadIdx.x; i < a.n; i += MAP_THREADS) {
      In v[NIN];
#pragma unroll
      for (int k = 0; k < NIN; k++) {
        if (a.mode[k] == OPND_ARRAY) v[k] = reinterpret_cast<const In*>(a.in[k])[i];
        else if (a.mode[k] == OPND_PERIODIC) v[k] = reinterpret_cast<const In*>(a.in[k])[i % (int64_t)a.period];
        else v[k] = scalar[k];
      }
      out[i] = F::apply(v, err);
    }
  }
  if (err) atomicOr(a.flags, err);
}

// ------------------------------------------------------------------ rows
// The general kernel: the innermost axis of the plan is spread over TX threads (E elements
// each; E > 1 needs unit/zero inner strides, E == 1 takes any inner stride), TY = 256/TX
// rows are processed side by side and UNROLL row-groups are in flight per thread.
// 1-D grid: the fastest block index tiles the inner axis, the rest enumerates (slab over the
// leading outer axes) x (chunk of rows_per_block rows of the LAST outer axis), so inside a
// block consecutive rows differ by one constant stride per operand: no div/mod per element.
template <typename F, int E, int UNROLL>
__global__ void __launch_bounds__(MAP_THREADS) map_rows_kernel(const MapArgs<F::NIN> a) {
  using In = typename F::In;
  using Out = typename F::Out;
  constexpr int NIN = F::NIN;
  const int tx = threadIdx.x & (a.tx - 1);
  const int ty = threadIdx.x >> a.tx_log2;
  const int TY = MAP_THREADS >> a.tx_log2;
  // blockIdx.x < 2^31, so the whole decomposition runs in 32-bit unsigned arithmetic (a 64-bit
  // division is ~100 instructions, and every thread of the block repeats this set-up)
  const uint32_t gx = (uint32_t)a.gx;
  const uint32_t by = blockIdx.x / gx;
  const uint32_t ctile = blockIdx.x - by * gx;
  const int64_t col = ((int64_t)ctile * a.tx + tx) * E;
  if (col >= a.n) return;                          // whole groups only: n % E == 0 by dispatch
  const int last = a.outer.n - 1;
  const int64_t rows_last = a.outer.extent[last];
  uint32_t slab, chunk;
  if (a.slab_fastest) { chunk = by / a.slabs; slab = by - chunk * a.slabs; }
  else { slab = by / a.chunks; chunk = by - slab * a.chunks; }
  uint32_t err = 0;

  int64_t base[NIN + 1];
#pragma unroll
  for (int k = 0; k <= NIN; k++) base[k] = 0;
  for (int ax = last - 1; ax >= 0; ax--) {         // leading outer axes: once per block
    const int64_t e64 = a.outer.extent[ax];
    const uint32_t e = e64 > 0x7fffffffLL ? 0x80000000u : (uint32_t)e64;   // slab < 2^31 <= e: q = 0
    const uint32_t q = slab / e;
    const int64_t c = slab - q * e;
    slab = q;
#pragma unroll
    for (int k = 0; k <= NIN; k++) base[k] += c * a.outer.stride[k][ax];
  }
  int64_t step[NIN + 1];
#pragma unroll
  for (int k = 0; k <= NIN; k++) {
    step[k] = a.outer.stride[k][last];
    if constexpr (E == 1) base[k] += col * a.inner_stride[k];
    else base[k] += (k == NIN || a.mode[k] == OPND_ARRAY || a.mode[k] == OPND_ROWVEC) ? col : 0;
  }

  const In* in[NIN];
  Group<In, E> fixed[NIN];                         // row-vector / scalar operands: loaded once
#pragma unroll
  for (int k = 0; k < NIN; k++) {
    in[k] = reinterpret_cast<const In*>(a.in[k]) + base[k];
    if (a.mode[k] == OPND_PARAM) fixed[k] = splat_group<In, E>(bits_to<In>(a.param[k]));
    else if (a.mode[k] == OPND_ROWVEC) fixed[k] = load_group<In, E>(in[k]);
    else fixed[k] = splat_group<In, E>(In());
  }
  Out* out = reinterpret_cast<Out*>(a.out) + base[NIN];

  const int64_t r0 = (int64_t)chunk * a.rows_per_block;
  const int64_t r1 = (r0 + a.rows_per_block < rows_last) ? r0 + a.rows_per_block : rows_last;
  int64_t r = r0 + ty;
  // x[u][k] of a fixed operand is written once here and never again: inside the loop
  // only array operands are (re)loaded, so the hot loop has no selects.
  Group<In, E> x[UNROLL][NIN];
#pragma unroll
  for (int u = 0; u < UNROLL; u++)
#pragma unroll
    for (int k = 0; k < NIN; k++) x[u][k] = fixed[k];
  for (; r + (int64_t)(UNROLL - 1) * TY < r1; r += (int64_t)UNROLL * TY) {
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
#pragma unroll
      for (int k = 0; k < NIN; k++) {
        if (a.mode[k] == OPND_ARRAY) x[u][k] = load_group<In, E>(in[k] + (r + (int64_t)u * TY) * step[k]);
        else if (a.mode[k] == OPND_BCAST) x[u][k] = splat_group<In, E>(in[k][(r + (int64_t)u * TY) * step[k]]);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      Group<Out, E> y;
      apply_group<F, E>(x[u], y, err);
      store_group<Out, E>(out + (r + (int64_t)u * TY) * step[NIN], y);
    }
  }
  for (; r < r1; r += TY) {
#pragma unroll
    for (int k = 0; k < NIN; k++) {
      if (a.mode[k] == OPND_ARRAY) x[0][k] = load_group<In, E>(in[k] + r * step[k]);
      else if (a.mode[k] == OPND_BCAST) x[0][k] = splat_group<In, E>(in[k][r * step[k]]);
    }
    Group<Out, E> y;
    apply_group<F, E>(x[0], y, err);
    store_group<Out, E>(out + r * step[NIN], y);
  }
  if (err) atomicOr(a.flags, err);
}

// ------------------------------------------------------------------ dispatch
struct MapOperand {
  const void* base = nullptr;    // device pointer (element 0 of the buffer), or null for a param
  const ph_desc* desc = nullptr;
  uint64_t param = 0;
  bool is_param = false;
};

// Consecutive flat launches traverse their arrays in OPPOSITE directions.  The fluent API runs an
// expression as a chain of launches, each consuming the array the previous one produced
// (`t = a * b` then `t + c`); when a launch ends, the last ~100 MB it wrote are still in the
// 126 MB L2.  A consumer that starts from the far end reads them from L2 instead of HBM (and a
// temporary that is overwritten while still resident never reaches HBM at all).  Elementwise
// results do not depend on the visiting order.  PH_FLAT_NO_ALTERNATE=1 disables it (A/B runs).
template <typename F, int E, int UNROLL>
inline int32_t launch_flat(MapArgs<F::NIN>& a) {
  const int64_t tile = (int64_t)MAP_THREADS * E * UNROLL;
  const int64_t blocks = ceil_div(a.n, tile);
  if (blocks > 0x7fffffffLL) return set_error(PH_ERR_INVALID, "array too large for one launch");
  static const bool alternate = getenv("PH_FLAT_NO_ALTERNATE") == nullptr;
  a.reverse = alternate ? (int)(rt().flat_launches++ & 1) : 0;
  // L2 eviction priorities (ld_stream_evict_first / st_stream_evict_last): inputs evict_first, results evict_last.
  // Bench headline 6786 -> 6843 GB/s (`t + c` 117.5 -> 114.5 us: more of the temporary survives in L2 until its
  // consumer arrives; `t = a * b` 84.5 -> 85.7 us).  PH_FLAT_L2_HINT=0 restores plain accesses, 1 = loads only.
  static const int l2_hint = getenv("PH_FLAT_L2_HINT") ? atoi(getenv("PH_FLAT_L2_HINT")) : 2;
  a.l2_hint = l2_hint;
  map_flat_kernel<F, E, UNROLL><<<(unsigned)blocks, MAP_THREADS, 0, rt().stream>>>(a);
  PH_LAUNCH_CHECK("map_flat_kernel");
  return PH_OK;
}

template <typename F, int E, int UNROLL>
inline int32_t launch_rows(MapArgs<F::NIN>& a) {
  const int64_t groups = ceil_div(a.n, (int64_t)E);
  int tx = 1, lg = 0;
  while (tx < MAP_THREADS && tx < groups) { tx <<= 1; lg++; }
  const int ty = MAP_THREADS / tx;
  a.tx = tx; a.tx_log2 = lg;
  const int64_t gx = ceil_div(groups, (int64_t)tx);
  const int64_t rows_last = a.outer.extent[a.outer.n - 1];
  const int64_t slabs = a.rows / rows_last;
  // enough blocks for ~8 waves, but several row-groups per thread so that the per-block set-up and
  // row-vector loads amortise: 4 unrolled batches per thread, or one when that would leave fewer
  // than half the wanted blocks (a 256 MB output tiled [2,2]: 2048 blocks = 2.3 waves with a tail)
  const int64_t target_blocks = (int64_t)rt().sm_count * 8 * 8;
  int64_t chunks = std::max<int64_t>(1, target_blocks / std::max<int64_t>(1, gx * slabs));
  int64_t rows_min = (int64_t)UNROLL * 4 * ty;
  if (gx * slabs * ceil_div(rows_last, rows_min) < target_blocks / 2) rows_min = (int64_t)UNROLL * ty;
  int64_t rpb = std::max<int64_t>(ceil_div(rows_last, chunks), std::min<int64_t>(rows_last, rows_min));
  chunks = ceil_div(rows_last, rpb);
  const int64_t blocks = gx * slabs * chunks;
  if (blocks > 0x7fffffffLL) return set_error(PH_ERR_INVALID, "array too large for one launch");
  a.rows_per_block = (int)std::min<int64_t>(rpb, 0x7fffffff);
  a.gx = gx;
  a.chunks = (uint32_t)chunks;
  a.slabs = (uint32_t)slabs;
  // An input with stride 0 on a leading outer axis (a tile count, a broadcast plane) is re-read by
  // every slab along that axis: let consecutive blocks walk the slabs first, so the re-reads run
  // together and hit L2 instead of coming back to HBM once per slab.
  a.slab_fastest = 0;
  for (int k = 0; k < F::NIN && slabs > 1; k++)
    if (a.mode[k] == OPND_ARRAY || a.mode[k] == OPND_BCAST)
      for (int ax = 0; ax + 1 < a.outer.n; ax++)
        if (a.outer.extent[ax] > 1 && a.outer.stride[k][ax] == 0) a.slab_fastest = 1;
  map_rows_kernel<F, E, UNROLL><<<(unsigned)blocks, MAP_THREADS, 0, rt().stream>>>(a);
  PH_LAUNCH_CHECK("map_rows_kernel");
  return PH_OK;
}

// ops[0..NIN-1] are the inputs, `out`/`out_desc` the output.
template <typename F>
int32_t launch_map(const MapOperand* ops, void* out, const ph_desc* out_desc) {
  using In = typename F::In;
  using Out = typename F::Out;
  constexpr int NIN = F::NIN;
  constexpr int WIDE = sizeof(In) > sizeof(Out) ? sizeof(In) : sizeof(Out);
  PH_REQUIRE_INIT();
  if (!out || !out_desc) return set_error(PH_ERR_INVALID, "null output");

  // plan over the array operands + the output (output last)
  const ph_desc* descs[MAX_OPERANDS];
  int slot[NIN];
  int nd = 0;
  for (int k = 0; k < NIN; k++) {
    if (ops[k].is_param) { slot[k] = -1; continue; }
    if (!ops[k].base || !ops[k].desc) return set_error(PH_ERR_INVALID, "null input operand %d", k);
    slot[k] = nd;
    descs[nd++] = ops[k].desc;
  }
  const int out_slot = nd;
  descs[nd++] = out_desc;
  Plan p;
  int32_t st = make_plan(p, nd, descs);
  if (st != PH_OK) return st;
  if (p.total == 0) return PH_OK;

  MapArgs<NIN> a;
  memset(&a, 0, sizeof(a));
  a.flags = rt().d_flags;
  a.out = reinterpret_cast<Out*>(out) + p.offset[out_slot];
  for (int k = 0; k < NIN; k++) {
    if (ops[k].is_param) { a.mode[k] = OPND_PARAM; a.param[k] = ops[k].param; a.in[k] = nullptr; }
    else a.in[k] = reinterpret_cast<const In*>(ops[k].base) + p.offset[slot[k]];
  }

  const int inner = p.rank - 1;   // -1 when rank == 0 (single element)
  auto in_stride = [&](int k, int ax) { return p.stride[slot[k]][ax]; };

  // ---- vector width: every array operand and the output must be aligned to its group
  auto pick_width = [&](bool rows) {
    int vb = 32;
    for (; vb > WIDE; vb >>= 1) {
      const int e = vb / WIDE;
      bool ok = (p.rank == 0 ? 1 : p.extent[inner]) % e == 0 || !rows;
      auto aligned = [&](const void* ptr, int esz, int slot_idx) {
        if (((uintptr_t)ptr) % (uintptr_t)(e * esz)) return false;
        if (rows)
          for (int ax = 0; ax < inner; ax++)
            if ((p.stride[slot_idx][ax] * esz) % (e * esz)) return false;
        return true;
      };
      if (!aligned(a.out, sizeof(Out), out_slot)) ok = false;
      for (int k = 0; k < NIN && ok; k++) {
        if (a.mode[k] == OPND_PARAM) continue;
        if (p.rank > 0 && in_stride(k, inner) == 0) {   // broadcast along the inner axis: scalar loads
          continue;
        }
        if (!aligned(a.in[k], sizeof(In), slot[k])) ok = false;
      }
      if (ok) break;
    }
    return vb;
  };

  // ---- flat?
  bool flat = p.rank <= 1;
  if (flat && p.rank == 1) {
    if (p.stride[out_slot][0] != 1) flat = false;
    for (int k = 0; k < NIN && flat; k++)
      if (a.mode[k] != OPND_PARAM && in_stride(k, 0) != 1 && in_stride(k, 0) != 0) flat = false;
  }
  if (flat) {
    a.n = p.total;
    for (int k = 0; k < NIN; k++)
      if (a.mode[k] != OPND_PARAM)
        a.mode[k] = (p.rank == 0 || in_stride(k, 0) == 0) ? OPND_BCAST : OPND_ARRAY;
    a.all_array = 1;
    for (int k = 0; k < NIN; k++) if (a.mode[k] != OPND_ARRAY) a.all_array = 0;
    const int vb = pick_width(false);
    if (vb == 32 && WIDE <= 16) return launch_flat<F, 32 / WIDE, 2>(a);
    if constexpr (!F::kCompact) {
      if (vb >= 16 && WIDE <= 8) return launch_flat<F, 16 / WIDE, 4>(a);
    }
    return launch_flat<F, 1, 4>(a);
  }

  // ---- periodic flat: rank 2, the output and every full operand contiguous over both axes, the
  //      others a contiguous row vector repeated down axis 0 (stride 0).  One tile per block like
  //      the flat kernel: measured 6 % faster than the rows kernel on the BASELINE broadcast
  //      (benchmarks/micro_stream.cu), because tiles spread over HBM better than per-block row chunks.
  if (p.rank == 2 && p.extent[1] < 0xffffffffLL && p.stride[out_slot][1] == 1 && p.stride[out_slot][0] == p.extent[1]) {
    bool ok = true, any_periodic = false;
    for (int k = 0; k < NIN && ok; k++) {
      if (a.mode[k] == OPND_PARAM) continue;
      const int64_t s0 = in_stride(k, 0), s1 = in_stride(k, 1);
      if (s1 == 1 && s0 == p.extent[1]) continue;
      if (s1 == 1 && s0 == 0) { any_periodic = true; continue; }
      ok = false;
    }
    if (ok && any_periodic) {
      // vector width: every pointer aligned to the group and the period a multiple of it
      int vb = 32;
      for (; vb > WIDE; vb >>= 1) {
        const int e = vb / WIDE;
        bool fits = p.extent[1] % e == 0 && ((uintptr_t)a.out) % (uintptr_t)(e * sizeof(Out)) == 0;
        for (int k = 0; k < NIN && fits; k++)
          if (a.mode[k] != OPND_PARAM && ((uintptr_t)a.in[k]) % (uintptr_t)(e * sizeof(In))) fits = false;
        if (fits) break;
      }
      if (vb == 32 && WIDE <= 16) {
        a.n = p.total;
        a.period = (uint32_t)p.extent[1];
        a.all_array = 0;
        for (int k = 0; k < NIN; k++)
          if (a.mode[k] != OPND_PARAM) a.mode[k] = in_stride(k, 0) == 0 ? OPND_PERIODIC : OPND_ARRAY;
        return launch_flat<F, 32 / WIDE, 2>(a);
      }
    }
  }

  // ---- outer axes (a strided 1-D plan gets a dummy outer axis of extent 1)
  a.outer.n = inner > 0 ? inner : 1;
  if (inner == 0) {
    a.outer.extent[0] = 1;
    for (int k = 0; k <= NIN; k++) a.outer.stride[k][0] = 0;
  }
  // A block walks rows of the LAST outer axis only (the leading ones are decomposed once per
  // block).  The visiting order of an elementwise map is free, so when that axis is short --
  // `tile` makes axes of extent 2, a [N, 3, C] slice has one of 3 -- the longest outer axis
  // takes its place: otherwise every thread would pay the per-block set-up (two 64-bit
  // divisions per leading axis) for one or two 32-byte groups (tile [4096,4096] x [2,2]:
  // 197 instructions per thread, 0.52 of the copy peak).
  int order[PH_MAX_RANK];
  for (int ax = 0; ax < inner; ax++) order[ax] = ax;
  if (inner >= 2 && p.extent[inner - 1] < 64) {
    int longest = inner - 1;
    for (int ax = 0; ax < inner - 1; ax++) if (p.extent[ax] > p.extent[longest]) longest = ax;
    order[longest] = inner - 1;
    order[inner - 1] = longest;
  }
  for (int ax = 0; ax < inner; ax++) {
    const int from = order[ax];
    a.outer.extent[ax] = p.extent[from];
    for (int k = 0; k < NIN; k++) a.outer.stride[k][ax] = (a.mode[k] == OPND_PARAM) ? 0 : in_stride(k, from);
    a.outer.stride[NIN][ax] = p.stride[out_slot][from];
  }
  a.n = p.extent[inner];
  a.rows = p.total / a.n;
  for (int k = 0; k < NIN; k++) a.inner_stride[k] = (a.mode[k] == OPND_PARAM) ? 0 : in_stride(k, inner);
  a.inner_stride[NIN] = p.stride[out_slot][inner];

  // ---- vectorisable rows: unit inner stride on the output, unit / zero on the inputs
  bool vec = p.stride[out_slot][inner] == 1;
  for (int k = 0; k < NIN && vec; k++)
    if (a.mode[k] != OPND_PARAM && in_stride(k, inner) != 1 && in_stride(k, inner) != 0) vec = false;
  if (vec) {
    for (int k = 0; k < NIN; k++) {
      if (a.mode[k] == OPND_PARAM) continue;
      if (in_stride(k, inner) == 0) { a.mode[k] = OPND_BCAST; continue; }
      bool rowvec = true;
      for (int ax = 0; ax < inner; ax++) if (in_stride(k, ax) != 0) rowvec = false;
      a.mode[k] = (rowvec && inner > 0) ? OPND_ROWVEC : OPND_ARRAY;
    }
    const int vb = pick_width(true);
    if (vb == 32 && WIDE <= 16) return launch_rows<F, 32 / WIDE, 4>(a);
    if constexpr (!F::kCompact) {
      if (vb >= 16 && WIDE <= 8) return launch_rows<F, 16 / WIDE, 4>(a);
    }
    return launch_rows<F, 1, 4>(a);
  }
  // ---- arbitrary inner strides: one element per thread per row, still no div/mod per element
  for (int k = 0; k < NIN; k++)
    if (a.mode[k] != OPND_PARAM) a.mode[k] = OPND_ARRAY;
  return launch_rows<F, 1, 4>(a);
}

}  // namespace ph
