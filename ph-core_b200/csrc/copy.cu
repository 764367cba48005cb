// copy.cu -- strided gather / scatter / fill / masked store (K4, K5, K6, K10).
//
// ph_copy_strided moves every element of a descriptor-described source region onto a
// descriptor-described destination region.  It replaces the per-element loops of
//   NArray#unsafe_fetch_chunk  src/n_array.cr:450-453     (gather, lex order)
//   NArray#unsafe_set_chunk    src/n_array.cr:484-500     (scatter / fill)
//   View#to_narr               src/view.cr:123-126        (transform chain -> copy)
//   MutableView writes         src/mutable_view.cr:16-18
// Element order is irrelevant on the device: both descriptors enumerate the same logical
// coordinates, so element i of the region lands where the reference's lex iteration puts it.
//
// Kernel choice: contiguous -> 256-bit flat copy; inner axis contiguous on both sides ->
// vectorised rows; inner stride != 1 (step slicing, reversal) -> scalar rows (coalesced at
// sector granularity); unit-stride axis differs between source and destination (permuted
// views, transposes) -> shared-memory tile transpose so both sides stay coalesced.
#include "map_kernels.cuh"

namespace ph {

template <typename T>
struct CopyOp {
  using In = T;
  using Out = T;
  static constexpr int NIN = 1;
  static constexpr bool kCompact = false;
  static __device__ __forceinline__ Out apply(const In (&x)[1], uint32_t&) { return x[0]; }
};

// ------------------------------------------------------------------ tile transpose
constexpr int TT = 32;      // tile edge (elements)
constexpr int TROWS = 8;    // block = TT x TROWS threads

struct TransposeArgs {
  const void* src;
  void* dst;
  int64_t extA, extB;                 // A: unit-stride axis of the source, B: of the destination
  int64_t sA_src, sB_src, sA_dst, sB_dst;
  int64_t tilesA, tilesB;
  int n_outer;
  int grid3d;
  int64_t outer_extent[PH_MAX_RANK];
  int64_t outer_src[PH_MAX_RANK], outer_dst[PH_MAX_RANK];
};

template <typename T>
__global__ void __launch_bounds__(TT * TROWS) transpose_kernel(const TransposeArgs a) {
  __shared__ T tile[TT][TT + 1];
  // 3-D grid (tiles along A, tiles along B, outer index): no 64-bit division per thread for the tile
  // coordinates (two of them cost more than moving the thread's four elements); a 1-D grid only when
  // a dimension exceeds the grid limits
  int64_t ta, tb, bid;
  if (a.grid3d) { ta = blockIdx.x; tb = blockIdx.y; bid = blockIdx.z; }
  else { bid = blockIdx.x; ta = bid % a.tilesA; bid /= a.tilesA; tb = bid % a.tilesB; bid /= a.tilesB; }
  int64_t off_src = 0, off_dst = 0;
  for (int ax = a.n_outer - 1; ax >= 0; ax--) {
    const int64_t c = bid % a.outer_extent[ax];
    bid /= a.outer_extent[ax];
    off_src += c * a.outer_src[ax];
    off_dst += c * a.outer_dst[ax];
  }
  const T* __restrict__ src = reinterpret_cast<const T*>(a.src) + off_src;
  T* __restrict__ dst = reinterpret_cast<T*>(a.dst) + off_dst;
  const int tx = threadIdx.x, ty = threadIdx.y;
  // Whole tiles (all but the last row / column of tiles) take a path with no range tests and one
  // pointer + constant step per thread: the checked form costs ~60 instructions per element and was
  // issue-bound (ncu: issue 70 %, DRAM 63 %).
  const bool whole = (ta + 1) * TT <= a.extA && (tb + 1) * TT <= a.extB;      // block-uniform
  if (whole) {
    const T* ps = src + (ta * TT + tx) * a.sA_src + (tb * TT + ty) * a.sB_src;
    const int64_t step_s = (int64_t)TROWS * a.sB_src;
    T v[TT / TROWS];
#pragma unroll
    for (int j = 0; j < TT / TROWS; j++) v[j] = ps[j * step_s];               // all loads in flight first
#pragma unroll
    for (int j = 0; j < TT / TROWS; j++) tile[ty + j * TROWS][tx] = v[j];
    __syncthreads();
    T* pd = dst + (ta * TT + ty) * a.sA_dst + (tb * TT + tx) * a.sB_dst;
    const int64_t step_d = (int64_t)TROWS * a.sA_dst;
#pragma unroll
    for (int j = 0; j < TT / TROWS; j++) pd[j * step_d] = tile[tx][ty + j * TROWS];
    return;
  }
  {
    const int64_t ia = ta * TT + tx;               // threads run along A: coalesced source reads
#pragma unroll
    for (int j = 0; j < TT; j += TROWS) {
      const int64_t ib = tb * TT + ty + j;
      if (ia < a.extA && ib < a.extB) tile[ty + j][tx] = src[ia * a.sA_src + ib * a.sB_src];
    }
  }
  __syncthreads();
  {
    const int64_t ib = tb * TT + tx;               // threads run along B: coalesced destination writes
#pragma unroll
    for (int j = 0; j < TT; j += TROWS) {
      const int64_t ia = ta * TT + ty + j;
      if (ia < a.extA && ib < a.extB) dst[ia * a.sA_dst + ib * a.sB_dst] = tile[tx][ty + j];
    }
  }
}

template <typename T>
static int32_t try_transpose(const Plan& p, const void* src, void* dst, bool& done) {
  done = false;
  if (p.rank < 2) return PH_OK;
  int A = -1, B = -1;
  for (int ax = 0; ax < p.rank; ax++) {
    if (A < 0 && (p.stride[0][ax] == 1 || p.stride[0][ax] == -1)) A = ax;
    if (B < 0 && (p.stride[1][ax] == 1 || p.stride[1][ax] == -1)) B = ax;
  }
  if (A < 0 || B < 0 || A == B) return PH_OK;
  if (p.extent[A] < 8 || p.extent[B] < 8) return PH_OK;      // too thin to tile: rows kernel
  TransposeArgs a;
  memset(&a, 0, sizeof(a));
  a.src = reinterpret_cast<const T*>(src) + p.offset[0];
  a.dst = reinterpret_cast<T*>(dst) + p.offset[1];
  a.extA = p.extent[A]; a.extB = p.extent[B];
  a.sA_src = p.stride[0][A]; a.sB_src = p.stride[0][B];
  a.sA_dst = p.stride[1][A]; a.sB_dst = p.stride[1][B];
  a.tilesA = ceil_div(a.extA, TT); a.tilesB = ceil_div(a.extB, TT);
  int64_t blocks = a.tilesA * a.tilesB;
  for (int ax = 0; ax < p.rank; ax++) {
    if (ax == A || ax == B) continue;
    a.outer_extent[a.n_outer] = p.extent[ax];
    a.outer_src[a.n_outer] = p.stride[0][ax];
    a.outer_dst[a.n_outer] = p.stride[1][ax];
    a.n_outer++;
    blocks *= p.extent[ax];
  }
  if (blocks > 0x7fffffffLL) return PH_OK;
  dim3 block(TT, TROWS);
  const int64_t outer_count = blocks / (a.tilesA * a.tilesB);
  a.grid3d = a.tilesA <= 0x7fffffffLL && a.tilesB <= 65535 && outer_count <= 65535;
  if (a.grid3d) {
    dim3 grid((unsigned)a.tilesA, (unsigned)a.tilesB, (unsigned)outer_count);
    transpose_kernel<T><<<grid, block, 0, rt().stream>>>(a);
  } else
  transpose_kernel<T><<<(unsigned)blocks, block, 0, rt().stream>>>(a);
  PH_LAUNCH_CHECK("transpose_kernel");
  done = true;
  return PH_OK;
}

// ------------------------------------------------------------------ inner-strided rows
// Gather / scatter whose innermost axis is not unit-stride on one side (step slicing
// `0..2..`, reversal `..-1..`, a scatter into every k-th column).  One element per thread per
// row is instruction-bound on B200 (~20 instructions per 8 bytes), so a thread moves a group
// of E elements (32 bytes): each side is read / written as
//   VEC    stride +1 : one 256-bit access
//   REV    stride -1 : one 256-bit access + in-register reversal
//   STEP2  stride +2 : two 256-bit loads, keep the even lanes (source side only)
//   STRIDED any other stride: E scalar accesses
// Same block mapping as map_rows_kernel: TX threads along the inner axis, 256/TX rows side by
// side, outer coordinates decomposed once per block.
enum SideMode : int { SIDE_VEC = 0, SIDE_REV = 1, SIDE_STEP2 = 2, SIDE_STRIDED = 3 };

struct CopyRowsArgs {
  const void* src;
  void* dst;
  int64_t n;                  // inner extent
  int64_t s_src, s_dst;       // inner strides (elements)
  int64_t rows;
  int rows_per_block;
  int64_t gx;
  int tx, tx_log2;
  OuterAxes outer;            // stride[0] = source, stride[1] = destination
};

template <typename T, int E, int SM>
__device__ __forceinline__ Group<T, E> load_side(const T* row, int64_t col, int64_t stride) {
  Group<T, E> g;
  if constexpr (SM == SIDE_VEC) {
    g = load_group<T, E>(row + col);
  } else if constexpr (SM == SIDE_REV) {
    const Group<T, E> t = load_group<T, E>(row - col - (E - 1));
#pragma unroll
    for (int i = 0; i < E; i++) g.v[i] = t.v[E - 1 - i];
  } else if constexpr (SM == SIDE_STEP2) {
    const Group<T, E> a = load_group<T, E>(row + 2 * col);
    const Group<T, E> b = load_group<T, E>(row + 2 * col + E);
#pragma unroll
    for (int i = 0; i < E; i++) g.v[i] = (2 * i < E) ? a.v[(2 * i) % E] : b.v[(2 * i) % E];
  } else {
#pragma unroll
    for (int i = 0; i < E; i++) g.v[i] = row[(col + i) * stride];
  }
  return g;
}

template <typename T, int E, int DM>
__device__ __forceinline__ void store_side(T* row, int64_t col, int64_t stride, const Group<T, E>& g) {
  if constexpr (DM == SIDE_VEC) {
    store_group<T, E>(row + col, g);
  } else if constexpr (DM == SIDE_REV) {
    Group<T, E> t;
#pragma unroll
    for (int i = 0; i < E; i++) t.v[i] = g.v[E - 1 - i];
    store_group<T, E>(row - col - (E - 1), t);
  } else {
#pragma unroll
    for (int i = 0; i < E; i++) row[(col + i) * stride] = g.v[i];
  }
}

template <typename T, int E, int SM, int DM>
__global__ void __launch_bounds__(MAP_THREADS) copy_rows_kernel(const CopyRowsArgs a) {
  constexpr int UNROLL = 2;
  const int tx = threadIdx.x & (a.tx - 1);
  const int ty = threadIdx.x >> a.tx_log2;
  const int TY = MAP_THREADS >> a.tx_log2;
  const int64_t by = (int64_t)blockIdx.x / a.gx;
  const int64_t ctile = (int64_t)blockIdx.x - by * a.gx;
  const int64_t col = (ctile * a.tx + tx) * E;
  if (col >= a.n) return;
  // STEP2 loads touch one element past the last kept one: keep the last group of a row scalar
  const bool full = (col + E <= a.n) && !(SM == SIDE_STEP2 && col + E >= a.n);
  const int last = a.outer.n - 1;
  const int64_t rows_last = a.outer.extent[last];
  const int64_t chunks = (rows_last + a.rows_per_block - 1) / a.rows_per_block;
  int64_t slab = by / chunks;
  const int64_t chunk = by - slab * chunks;
  int64_t bs = 0, bd = 0;
  for (int ax = last - 1; ax >= 0; ax--) {
    const int64_t e = a.outer.extent[ax];
    const int64_t q = slab / e;
    const int64_t c = slab - q * e;
    slab = q;
    bs += c * a.outer.stride[0][ax];
    bd += c * a.outer.stride[1][ax];
  }
  const int64_t ss = a.outer.stride[0][last], sd = a.outer.stride[1][last];
  const T* src = reinterpret_cast<const T*>(a.src) + bs;
  T* dst = reinterpret_cast<T*>(a.dst) + bd;
  const int64_t r0 = chunk * a.rows_per_block;
  const int64_t r1 = (r0 + a.rows_per_block < rows_last) ? r0 + a.rows_per_block : rows_last;
  if (full) {
    int64_t r = r0 + ty;
    for (; r + (int64_t)(UNROLL - 1) * TY < r1; r += (int64_t)UNROLL * TY) {
      Group<T, E> g[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; u++) g[u] = load_side<T, E, SM>(src + (r + (int64_t)u * TY) * ss, col, a.s_src);
#pragma unroll
      for (int u = 0; u < UNROLL; u++) store_side<T, E, DM>(dst + (r + (int64_t)u * TY) * sd, col, a.s_dst, g[u]);
    }
    for (; r < r1; r += TY) {
      const Group<T, E> g = load_side<T, E, SM>(src + r * ss, col, a.s_src);
      store_side<T, E, DM>(dst + r * sd, col, a.s_dst, g);
    }
  } else {                                       // ragged last group of a row: element-wise
    for (int64_t r = r0 + ty; r < r1; r += TY)
      for (int64_t c = col; c < a.n; c++) dst[r * sd + c * a.s_dst] = src[r * ss + c * a.s_src];
  }
}

template <typename T, int E, int SM, int DM>
static int32_t launch_copy_rows(CopyRowsArgs& a) {
  const int64_t groups = ceil_div(a.n, (int64_t)E);
  int tx = 1, lg = 0;
  while (tx < MAP_THREADS && tx < groups) { tx <<= 1; lg++; }
  const int ty = MAP_THREADS / tx;
  a.tx = tx; a.tx_log2 = lg;
  a.gx = ceil_div(groups, (int64_t)tx);
  const int64_t rows_last = a.outer.extent[a.outer.n - 1];
  const int64_t slabs = a.rows / rows_last;
  const int64_t target_blocks = (int64_t)rt().sm_count * 8 * 8;
  int64_t chunks = std::max<int64_t>(1, target_blocks / std::max<int64_t>(1, a.gx * slabs));
  int64_t rpb = std::max<int64_t>(ceil_div(rows_last, chunks), std::min<int64_t>(rows_last, (int64_t)4 * ty));
  chunks = ceil_div(rows_last, rpb);
  const int64_t blocks = a.gx * slabs * chunks;
  if (blocks > 0x7fffffffLL) return set_error(PH_ERR_INVALID, "array too large for one launch");
  a.rows_per_block = (int)std::min<int64_t>(rpb, 0x7fffffff);
  copy_rows_kernel<T, E, SM, DM><<<(unsigned)blocks, MAP_THREADS, 0, rt().stream>>>(a);
  PH_LAUNCH_CHECK("copy_rows_kernel");
  return PH_OK;
}

// returns done = false when both inner strides are +1 (the vectorised map path is better)
template <typename T>
static int32_t try_copy_rows(const Plan& p, const void* src, void* dst, bool& done) {
  done = false;
  if (p.rank < 1) return PH_OK;
  const int inner = p.rank - 1;
  const int64_t s_src = p.stride[0][inner], s_dst = p.stride[1][inner];
  if (s_src == 1 && s_dst == 1) return PH_OK;
  if (s_dst == 0) return PH_OK;
  constexpr int E = 32 / (int)sizeof(T);
  if (inner == 0) {
    // One long strided row (e.g. a fully reversed contiguous array, whose axes coalesce into a
    // single axis of stride -1): as a single row every thread would move ONE group and pay the
    // whole per-thread setup for it (measured 55 instructions per element, 0.78 of peak).  Fold
    // it back into rows of L elements so the row loop below amortises the setup; the remainder
    // (< L elements) goes through the same path as a short row.
    constexpr int64_t L = (int64_t)MAP_THREADS * E * 4;
    if (p.extent[0] >= 2 * L) {
      const int64_t rows = p.extent[0] / L;
      Plan q = p;
      q.rank = 2;
      q.extent[0] = rows; q.extent[1] = L;
      for (int k = 0; k < 2; k++) { q.stride[k][0] = L * p.stride[k][0]; q.stride[k][1] = p.stride[k][0]; }
      q.total = rows * L;
      int32_t st = try_copy_rows<T>(q, src, dst, done);
      if (st != PH_OK || !done) return st;
      const int64_t rem = p.extent[0] - rows * L;
      if (rem > 0) {
        Plan t = p;
        t.extent[0] = rem; t.total = rem;
        for (int k = 0; k < 2; k++) t.offset[k] = p.offset[k] + rows * L * p.stride[k][0];
        st = try_copy_rows<T>(t, src, dst, done);
      }
      return st;
    }
  }
  CopyRowsArgs a;
  memset(&a, 0, sizeof(a));
  const T* sp = reinterpret_cast<const T*>(src) + p.offset[0];
  T* dp = reinterpret_cast<T*>(dst) + p.offset[1];
  a.src = sp; a.dst = dp;
  a.n = p.extent[inner];
  a.s_src = s_src; a.s_dst = s_dst;
  a.rows = p.total / a.n;
  a.outer.n = inner > 0 ? inner : 1;
  if (inner == 0) { a.outer.extent[0] = 1; a.outer.stride[0][0] = 0; a.outer.stride[1][0] = 0; }
  bool outer_src_ok = true, outer_dst_ok = true;
  for (int ax = 0; ax < inner; ax++) {
    a.outer.extent[ax] = p.extent[ax];
    a.outer.stride[0][ax] = p.stride[0][ax];
    a.outer.stride[1][ax] = p.stride[1][ax];
    if (p.stride[0][ax] % E) outer_src_ok = false;
    if (p.stride[1][ax] % E) outer_dst_ok = false;
  }
  auto elem_index = [](const void* ptr) { return (int64_t)((uintptr_t)ptr / sizeof(T)); };
  const bool byte_ok_s = ((uintptr_t)sp % sizeof(T)) == 0, byte_ok_d = ((uintptr_t)dp % sizeof(T)) == 0;
  int sm = SIDE_STRIDED, dm = SIDE_STRIDED;
  if (byte_ok_s && outer_src_ok) {
    if (s_src == 1 && elem_index(sp) % E == 0) sm = SIDE_VEC;
    else if (s_src == -1 && (elem_index(sp) + 1) % E == 0) sm = SIDE_REV;
    else if (s_src == 2 && elem_index(sp) % E == 0) sm = SIDE_STEP2;
  }
  if (byte_ok_d && outer_dst_ok) {
    if (s_dst == 1 && elem_index(dp) % E == 0) dm = SIDE_VEC;
    else if (s_dst == -1 && (elem_index(dp) + 1) % E == 0) dm = SIDE_REV;
  }
  int32_t st;
#define PH_COPY_CASE(SMV, DMV) if (sm == SMV && dm == DMV) { st = launch_copy_rows<T, E, SMV, DMV>(a); done = (st == PH_OK); return st; }
  PH_COPY_CASE(SIDE_VEC, SIDE_REV) PH_COPY_CASE(SIDE_VEC, SIDE_STRIDED)
  PH_COPY_CASE(SIDE_REV, SIDE_VEC) PH_COPY_CASE(SIDE_REV, SIDE_REV) PH_COPY_CASE(SIDE_REV, SIDE_STRIDED)
  PH_COPY_CASE(SIDE_STEP2, SIDE_VEC) PH_COPY_CASE(SIDE_STEP2, SIDE_REV) PH_COPY_CASE(SIDE_STEP2, SIDE_STRIDED)
  PH_COPY_CASE(SIDE_STRIDED, SIDE_VEC) PH_COPY_CASE(SIDE_STRIDED, SIDE_REV) PH_COPY_CASE(SIDE_STRIDED, SIDE_STRIDED)
#undef PH_COPY_CASE
  return PH_OK;   // (VEC, VEC): both sides contiguous but stride flags said otherwise -> map path
}

template <typename T>
static int32_t copy_strided_t(const void* src, const ph_desc* sd, void* dst, const ph_desc* dd) {
  const ph_desc* descs[2] = {sd, dd};
  Plan p;
  int32_t st = make_plan(p, 2, descs);
  if (st != PH_OK) return st;
  if (p.total == 0) return PH_OK;
  bool done = false;
  st = try_transpose<T>(p, src, dst, done);
  if (st != PH_OK || done) return st;
  st = try_copy_rows<T>(p, src, dst, done);
  if (st != PH_OK || done) return st;
  MapOperand ops[1];
  ops[0].base = src; ops[0].desc = sd;
  return launch_map<CopyOp<T>>(ops, dst, dd);
}

template <typename T>
static int32_t fill_t(void* dst, const ph_desc* dd, uint64_t bits) {
  MapOperand ops[1];
  ops[0].is_param = true; ops[0].param = bits;
  return launch_map<CopyOp<T>>(ops, dst, dd);
}

// ------------------------------------------------------------------ masked store
// dst[i] = mask[i] ? value : dst[i]   (NArray#[]=(mask, value), src/n_array.cr:510-551)
template <typename T, int E>
__global__ void __launch_bounds__(MAP_THREADS) mask_set_flat_kernel(T* __restrict__ dst,
                                                                    const uint8_t* __restrict__ mask,
                                                                    const T* __restrict__ src, uint64_t param,
                                                                    int has_src, int64_t n) {
  const int64_t tile = (int64_t)MAP_THREADS * E;
  const int64_t base = (int64_t)blockIdx.x * tile;
  const T scalar = bits_to<T>(param);
  if (base + tile <= n) {
    const int64_t idx = base + (int64_t)threadIdx.x * E;
    Group<uint8_t, E> m = load_group<uint8_t, E>(mask + idx);
    bool any = false;
#pragma unroll
    for (int i = 0; i < E; i++) any |= (m.v[i] != 0);
    if (!any) return;                               // untouched groups cost only their mask bytes
    Group<T, E> d = load_group<T, E>(dst + idx);
    if (has_src) {
      Group<T, E> s = load_group<T, E>(src + idx);
#pragma unroll
      for (int i = 0; i < E; i++) if (m.v[i]) d.v[i] = s.v[i];
    } else {
#pragma unroll
      for (int i = 0; i < E; i++) if (m.v[i]) d.v[i] = scalar;
    }
    store_group<T, E>(dst + idx, d);
  } else {
    for (int64_t i = base + threadIdx.x; i < n; i += MAP_THREADS)
      if (mask[i]) dst[i] = has_src ? src[i] : scalar;
  }
}

struct MaskAnyArgs {
  int rank;
  int64_t extent[PH_MAX_RANK];
  int64_t s_dst[PH_MAX_RANK], s_mask[PH_MAX_RANK], s_src[PH_MAX_RANK];
};

template <typename T>
__global__ void __launch_bounds__(MAP_THREADS) mask_set_any_kernel(T* dst, const uint8_t* mask, const T* src,
                                                                   uint64_t param, int has_src,
                                                                   const MaskAnyArgs a, int64_t total) {
  const T scalar = bits_to<T>(param);
  const int64_t stride = (int64_t)gridDim.x * MAP_THREADS;
  for (int64_t i = (int64_t)blockIdx.x * MAP_THREADS + threadIdx.x; i < total; i += stride) {
    int64_t r = i, od = 0, om = 0, os = 0;
    for (int ax = a.rank - 1; ax >= 0; ax--) {
      const int64_t q = r / a.extent[ax];
      const int64_t c = r - q * a.extent[ax];
      r = q;
      od += c * a.s_dst[ax]; om += c * a.s_mask[ax]; os += c * a.s_src[ax];
    }
    if (mask[om]) dst[od] = has_src ? src[os] : scalar;
  }
}

template <typename T>
static int32_t mask_set_t(void* dst, const ph_desc* dd, const uint8_t* mask, const ph_desc* md, const void* src,
                          const ph_desc* sd, uint64_t param) {
  const int nops = src ? 3 : 2;
  const ph_desc* descs[3] = {dd, md, sd};
  Plan p;
  int32_t st = make_plan(p, nops, descs);
  if (st != PH_OK) return st;
  if (p.total == 0) return PH_OK;
  T* d = reinterpret_cast<T*>(dst) + p.offset[0];
  const uint8_t* m = mask + p.offset[1];
  const T* s = src ? reinterpret_cast<const T*>(src) + p.offset[2] : nullptr;
  bool flat = p.rank <= 1;
  if (flat && p.rank == 1)
    for (int k = 0; k < nops; k++) if (p.stride[k][0] != 1) flat = false;
  if (flat) {
    constexpr int E = 32 / (int)sizeof(T) > 16 ? 16 : 32 / (int)sizeof(T);   // mask group <= 16 bytes
    const bool aligned = ((uintptr_t)d % (E * sizeof(T)) == 0) && ((uintptr_t)m % E == 0) &&
                         (!s || (uintptr_t)s % (E * sizeof(T)) == 0);
    if (aligned) {
      const int64_t blocks = ceil_div(p.total, (int64_t)MAP_THREADS * E);
      mask_set_flat_kernel<T, E><<<(unsigned)blocks, MAP_THREADS, 0, rt().stream>>>(d, m, s, param, s != nullptr, p.total);
    } else {
      const int64_t blocks = ceil_div(p.total, (int64_t)MAP_THREADS);
      mask_set_flat_kernel<T, 1><<<(unsigned)blocks, MAP_THREADS, 0, rt().stream>>>(d, m, s, param, s != nullptr, p.total);
    }
    PH_LAUNCH_CHECK("mask_set_flat_kernel");
    return PH_OK;
  }
  MaskAnyArgs a;
  memset(&a, 0, sizeof(a));
  a.rank = p.rank;
  for (int ax = 0; ax < p.rank; ax++) {
    a.extent[ax] = p.extent[ax];
    a.s_dst[ax] = p.stride[0][ax]; a.s_mask[ax] = p.stride[1][ax];
    a.s_src[ax] = src ? p.stride[2][ax] : 0;
  }
  const int64_t blocks = std::min<int64_t>(ceil_div(p.total, MAP_THREADS), (int64_t)rt().sm_count * 32);
  mask_set_any_kernel<T><<<(unsigned)blocks, MAP_THREADS, 0, rt().stream>>>(d, m, s, param, s != nullptr, a, p.total);
  PH_LAUNCH_CHECK("mask_set_any_kernel");
  return PH_OK;
}

}  // namespace ph

using namespace ph;

#define PH_SIZE_SWITCH(elem_size, CALL)                                                       \
  switch (elem_size) {                                                                        \
    case 1: return CALL(uint8_t);                                                             \
    case 2: return CALL(uint16_t);                                                            \
    case 4: return CALL(uint32_t);                                                            \
    case 8: return CALL(uint64_t);                                                            \
    default: return set_error(PH_ERR_UNSUPPORTED, "element size %d is not 1, 2, 4 or 8", elem_size); \
  }

extern "C" {

int32_t ph_copy_strided(int32_t elem_size, const void* src, const ph_desc* src_desc, void* dst,
                        const ph_desc* dst_desc) {
  PH_REQUIRE_INIT();
  if (!src || !dst || !src_desc || !dst_desc) return set_error(PH_ERR_INVALID, "null argument to ph_copy_strided");
#define CALL(T) copy_strided_t<T>(src, src_desc, dst, dst_desc)
  PH_SIZE_SWITCH(elem_size, CALL)
#undef CALL
}

int32_t ph_fill_region(int32_t elem_size, void* dst, const ph_desc* dst_desc, const void* scalar_host) {
  PH_REQUIRE_INIT();
  if (!dst || !dst_desc || !scalar_host) return set_error(PH_ERR_INVALID, "null argument to ph_fill_region");
  if (elem_size != 1 && elem_size != 2 && elem_size != 4 && elem_size != 8)
    return set_error(PH_ERR_UNSUPPORTED, "element size %d is not 1, 2, 4 or 8", elem_size);
  const uint64_t bits = host_scalar_bits(scalar_host, elem_size);
#define CALL(T) fill_t<T>(dst, dst_desc, bits)
  PH_SIZE_SWITCH(elem_size, CALL)
#undef CALL
}

int32_t ph_mask_set_scalar(int32_t elem_size, void* dst, const ph_desc* dst_desc, const uint8_t* mask,
                           const ph_desc* mask_desc, const void* scalar_host) {
  PH_REQUIRE_INIT();
  if (!dst || !dst_desc || !mask || !mask_desc || !scalar_host)
    return set_error(PH_ERR_INVALID, "null argument to ph_mask_set_scalar");
  if (elem_size != 1 && elem_size != 2 && elem_size != 4 && elem_size != 8)
    return set_error(PH_ERR_UNSUPPORTED, "element size %d is not 1, 2, 4 or 8", elem_size);
  const uint64_t bits = host_scalar_bits(scalar_host, elem_size);
#define CALL(T) mask_set_t<T>(dst, dst_desc, mask, mask_desc, nullptr, nullptr, bits)
  PH_SIZE_SWITCH(elem_size, CALL)
#undef CALL
}

int32_t ph_mask_set_array(int32_t elem_size, void* dst, const ph_desc* dst_desc, const uint8_t* mask,
                          const ph_desc* mask_desc, const void* src, const ph_desc* src_desc) {
  PH_REQUIRE_INIT();
  if (!dst || !dst_desc || !mask || !mask_desc || !src || !src_desc)
    return set_error(PH_ERR_INVALID, "null argument to ph_mask_set_array");
#define CALL(T) mask_set_t<T>(dst, dst_desc, mask, mask_desc, src, src_desc, 0)
  PH_SIZE_SWITCH(elem_size, CALL)
#undef CALL
}

}  // extern "C"
