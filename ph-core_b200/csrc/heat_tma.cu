// heat_tma.cu -- 3-D heat stencil with a TMA-fed shared-memory plane pipeline (sm_100a).
//
// The register-streaming kernel in heat.cu is latency-bound on B200: each thread has one
// 16-byte load in flight per plane, i.e. ~20 KB per SM, where ~30 KB are needed to cover
// HBM latency at full bandwidth.  Here the bytes in flight are decoupled from registers and
// occupancy: one elected thread per block issues `cp.async.bulk.tensor.3d` (TMA) loads of
// whole halo'd tile planes into a ring of STAGES shared-memory buffers, several planes
// ahead; mbarriers signal arrival; all threads then read their cells and the 6 neighbours
// from shared memory.  TMA zero-fills out-of-range halo coordinates, so there is no edge
// clamping (boundary cells are held fixed, their neighbours are never used).
//
// Same arithmetic, same association order, same bit-exact results as heat.cu
// (examples/heat_equation.cr:38-51 generalised by SURVEY.md 8(a) a-9).
#include "ph_common.cuh"
#include "ops.cuh"
#include <cuda.h>
#include <algorithm>
#include <stdlib.h>

namespace ph {

constexpr int TMA_STAGES = 6;       // planes resident in shared memory (3 in use + 3 in flight)

template <typename T, int TMA_TY> struct TmaTile {
  static constexpr int E = 16 / (int)sizeof(T);       // cells per thread (one 16-byte group)
  static constexpr int W = 32 * E;                     // tile width in cells
  static constexpr int PITCH = W + 2 * E;              // + one group of halo each side (keeps 16-B alignment)
  static constexpr int ROWS = TMA_TY + 2;
  static constexpr int PLANE_BYTES = ROWS * PITCH * (int)sizeof(T);
  static constexpr int STAGE_BYTES = (PLANE_BYTES + 127) / 128 * 128;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

template <typename T>
struct HeatTmaArgs {
  T* out;
  int64_t n0, n1, n2;
  int64_t z_begin, z_end, z_chunk;
  T coeff;
};

template <typename T, int TMA_TY>
__global__ void __launch_bounds__(32 * TMA_TY) heat_tma_kernel(const __grid_constant__ CUtensorMap in_map,
                                                               const HeatTmaArgs<T> a) {
  using Tile = TmaTile<T, TMA_TY>;
  constexpr int E = Tile::E;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t full[TMA_STAGES];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t tile_x = (int64_t)blockIdx.x * Tile::W;
  const int64_t tile_y = (int64_t)blockIdx.y * TMA_TY;
  const int64_t zb = a.z_begin + (int64_t)blockIdx.z * a.z_chunk;
  const int64_t ze = (zb + a.z_chunk < a.z_end) ? zb + a.z_chunk : a.z_end;
  if (zb >= ze) return;

  // plane p (absolute index) lives in stage (p - (zb - 1)) % STAGES; the loop below tracks the
  // three stages it reads and their barrier parity incrementally (no div/mod per plane)
  auto issue = [&](int s, int64_t p) {
    mbar_expect_tx(&full[s], (uint32_t)Tile::PLANE_BYTES);
    tma_load_3d(smem_raw + (size_t)s * Tile::STAGE_BYTES, &in_map, (int)(tile_x - E), (int)(tile_y - 1), (int)p, &full[s]);
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < TMA_STAGES; s++) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const int nplanes = (int)(ze - zb);              // planes to update; planes zb-1 .. ze are read
  if (threadIdx.x == 0) {
    for (int k = 0; k < TMA_STAGES; k++)
      if (k <= nplanes + 1) issue(k, zb - 1 + k);
  }

  const int64_t x0 = tile_x + (int64_t)lane * E;
  const int64_t y = tile_y + warp;
  const bool active = x0 < a.n2 && y < a.n1;
  const bool fix_all = (y == 0 || y >= a.n1 - 1);
  const bool fix_first = (x0 == 0), fix_last = (x0 + E >= a.n2);
  const int64_t plane = a.n1 * a.n2;
  T* p_out = a.out + zb * plane + y * a.n2 + x0;
  // element offset of this thread's first cell inside a stage
  const int cell = (warp + 1) * Tile::PITCH + E + lane * E;
  const T* base = reinterpret_cast<const T*>(smem_raw);
  constexpr int STAGE_ELEMS = Tile::STAGE_BYTES / (int)sizeof(T);

  int sp = 0, sc = 1, sn = 2;                      // stages of planes z-1, z, z+1
  uint32_t par_n = 0;                              // parity of the current use of stage sn
  mbar_wait(&full[0], 0);
  mbar_wait(&full[1], 0);
  for (int it = 0; it < nplanes; it++) {
    // planes are issued in order: once z+1 has landed, z and z-1 have too
    mbar_wait(&full[sn], par_n);
    const T* P = base + sp * STAGE_ELEMS + cell;
    const T* C = base + sc * STAGE_ELEMS + cell;
    const T* N = base + sn * STAGE_ELEMS + cell;
    const Group<T, E> c = *reinterpret_cast<const Group<T, E>*>(C);
    Group<T, E> res = c;
    if (!fix_all) {
      const Group<T, E> zl = *reinterpret_cast<const Group<T, E>*>(P);
      const Group<T, E> zh = *reinterpret_cast<const Group<T, E>*>(N);
      const Group<T, E> up = *reinterpret_cast<const Group<T, E>*>(C - Tile::PITCH);
      const Group<T, E> dn = *reinterpret_cast<const Group<T, E>*>(C + Tile::PITCH);
      const T xl = C[-1];
      const T xr = C[E];
#pragma unroll
      for (int i = 0; i < E; i++) {
        const T l = (i > 0) ? c.v[i - 1] : xl;
        const T r = (i < E - 1) ? c.v[i + 1] : xr;
        const T two_c = f_mul((T)2, c.v[i]);
        const T d0 = f_add(f_sub(zl.v[i], two_c), zh.v[i]);
        const T d1 = f_add(f_sub(up.v[i], two_c), dn.v[i]);
        const T d2 = f_add(f_sub(l, two_c), r);
        res.v[i] = f_add(c.v[i], f_mul(f_add(f_add(d0, d1), d2), a.coeff));
      }
      if (fix_first) res.v[0] = c.v[0];
      if (fix_last) res.v[E - 1] = c.v[E - 1];
    }
    if (active) store_group<T, E>(p_out, res);
    p_out += plane;
    __syncthreads();                               // every thread is done with plane z-1: recycle its stage
    if (threadIdx.x == 0 && it + TMA_STAGES <= nplanes + 1) issue(sp, zb - 1 + it + TMA_STAGES);
    sp = sc; sc = sn;
    if (++sn == TMA_STAGES) { sn = 0; par_n ^= 1; }
  }
}

// ---------------------------------------------------------------------------- two steps per pass
// Temporal blocking: one pass over HBM advances the grid by TWO time steps, so the traffic
// per cell-update drops from 8 to 4 bytes (f32).  Ring A holds TMA-fed planes of time t with a
// 2-cell halo; every plane p the block first computes time t+1 on a (TY+2) x (W+2) region
// into ring B (shared memory only), then time t+2 on its TY x W tile for plane p-1 from ring
// B and writes that to HBM.  Every cell value is produced by exactly the same operations in
// the same order as two single-step passes (boundary cells are copied forward at t+1 just as
// they are held by a single step), so the result is bit-identical.
constexpr int TMA2_STAGES_A = 6;
constexpr int TMA2_STAGES_B = 4;

template <typename T, int TY> struct Tma2Tile {
  static constexpr int E = 16 / (int)sizeof(T);
  static constexpr int W = 32 * E;
  static constexpr int PITCH = W + 2 * E;
  static constexpr int ROWS_A = TY + 4, ROWS_B = TY + 2;
  static constexpr int A_BYTES = ROWS_A * PITCH * (int)sizeof(T);
  static constexpr int A_STAGE = (A_BYTES + 127) / 128 * 128;
  static constexpr int B_STAGE = (ROWS_B * PITCH * (int)sizeof(T) + 127) / 128 * 128;
  static constexpr int SMEM = TMA2_STAGES_A * A_STAGE + TMA2_STAGES_B * B_STAGE;
};

template <typename T>
__device__ __forceinline__ T heat7(T c, T zl, T zh, T yl, T yh, T xl, T xh, T coeff) {
  const T two_c = f_mul((T)2, c);
  const T d0 = f_add(f_sub(zl, two_c), zh);
  const T d1 = f_add(f_sub(yl, two_c), yh);
  const T d2 = f_add(f_sub(xl, two_c), xh);
  return f_add(c, f_mul(f_add(f_add(d0, d1), d2), coeff));
}

template <typename T, int TY>
__global__ void __launch_bounds__(32 * TY) heat_tma2_kernel(const __grid_constant__ CUtensorMap in_map,
                                                            const HeatTmaArgs<T> a) {
  using Tile = Tma2Tile<T, TY>;
  constexpr int E = Tile::E, PITCH = Tile::PITCH;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t full[TMA2_STAGES_A];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t tile_x = (int64_t)blockIdx.x * Tile::W;
  const int64_t tile_y = (int64_t)blockIdx.y * TY;
  const int64_t zb = a.z_begin + (int64_t)blockIdx.z * a.z_chunk;
  const int64_t ze = (zb + a.z_chunk < a.z_end) ? zb + a.z_chunk : a.z_end;
  if (zb >= ze) return;
  T* const ringA = reinterpret_cast<T*>(smem_raw);
  T* const ringB = reinterpret_cast<T*>(smem_raw + (size_t)TMA2_STAGES_A * Tile::A_STAGE);
  constexpr int A_ELEMS = Tile::A_STAGE / (int)sizeof(T), B_ELEMS = Tile::B_STAGE / (int)sizeof(T);

  // time-t planes zb-2 .. ze+1 stream through ring A in order; plane index k = plane - (zb-2)
  const int nA = (int)(ze - zb) + 4;
  auto issue = [&](int stage, int k) {
    mbar_expect_tx(&full[stage], (uint32_t)Tile::A_BYTES);
    tma_load_3d(smem_raw + (size_t)stage * Tile::A_STAGE, &in_map, (int)(tile_x - E), (int)(tile_y - 2),
                (int)(zb - 2 + k), &full[stage]);
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < TMA2_STAGES_A; s++) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int k = 0; k < TMA2_STAGES_A && k < nA; k++) issue(k, k);

  // ---- per-thread geometry (tile-local row r in [-2, TY+1], col c in [-E, W+E-1])
  const int c0 = lane * E;
  auto offA = [&](int r, int c) { return (r + 2) * PITCH + (c + E); };
  auto offB = [&](int r, int c) { return (r + 1) * PITCH + (c + E); };
  // Boundary cells are carried forward unchanged.  Positions OUTSIDE the grid only ever feed
  // boundary cells, so their values are don't-cares: only the grid's first / last row and
  // column need a test, and those are per-thread constants.
  const int64_t plane = a.n1 * a.n2;
  const int64_t gy_own = tile_y + warp, gx_own = tile_x + c0;
  const bool active = gx_own < a.n2 && gy_own < a.n1;
  const bool fix_first = gx_own == 0, fix_last = gx_own + E == a.n2;
  const bool row_fixed_own = gy_own == 0 || gy_own == a.n1 - 1;
  const int r_extra = warp == 0 ? -1 : TY;
  const int64_t gy_extra = tile_y + r_extra;
  const bool row_fixed_extra = gy_extra == 0 || gy_extra == a.n1 - 1;
  T* p_out = a.out + zb * plane + gy_own * a.n2 + gx_own;

  // time t+1 of the E cells at (r, c0) / of the single cell (r, c) of plane p into ring B
  auto t1_group = [&](const T* Ap, const T* Ac, const T* An, T* Bc, int r, bool fixed) {
    const int o = offA(r, c0);
    const Group<T, E> cc = *reinterpret_cast<const Group<T, E>*>(Ac + o);
    Group<T, E> res = cc;
    if (!fixed) {
      const Group<T, E> zl = *reinterpret_cast<const Group<T, E>*>(Ap + o);
      const Group<T, E> zh = *reinterpret_cast<const Group<T, E>*>(An + o);
      const Group<T, E> up = *reinterpret_cast<const Group<T, E>*>(Ac + o - PITCH);
      const Group<T, E> dn = *reinterpret_cast<const Group<T, E>*>(Ac + o + PITCH);
      const T xl = Ac[o - 1], xr = Ac[o + E];
#pragma unroll
      for (int i = 0; i < E; i++) {
        const T l = (i > 0) ? cc.v[i - 1] : xl;
        const T rr = (i < E - 1) ? cc.v[i + 1] : xr;
        res.v[i] = heat7<T>(cc.v[i], zl.v[i], zh.v[i], up.v[i], dn.v[i], l, rr, a.coeff);
      }
      if (fix_first) res.v[0] = cc.v[0];
      if (fix_last) res.v[E - 1] = cc.v[E - 1];
    }
    *reinterpret_cast<Group<T, E>*>(Bc + offB(r, c0)) = res;
  };
  // the extra columns -1 and W are never grid-boundary columns (tiles start on multiples of W)
  auto t1_cell = [&](const T* Ap, const T* Ac, const T* An, T* Bc, int r, int c, bool fixed) {
    const int o = offA(r, c);
    const T cc = Ac[o];
    Bc[offB(r, c)] = fixed ? cc : heat7<T>(cc, Ap[o], An[o], Ac[o - PITCH], Ac[o + PITCH], Ac[o - 1], Ac[o + 1], a.coeff);
  };

  int sa_p = 0, sa_c = 1, sa_n = 2;              // ring-A stages of time-t planes p-1, p, p+1
  uint32_t par_n = 0;
  int sb = 0;                                    // ring-B stage that receives time-(t+1) plane p
  mbar_wait(&full[0], 0);
  mbar_wait(&full[1], 0);
  // p runs over the time-(t+1) planes zb-1 .. ze; output plane q = p-1 is written once p >= zb+1
  const int iters = (int)(ze - zb) + 2;
  for (int it = 0; it < iters; it++) {
    const int64_t p = zb - 1 + it;
    mbar_wait(&full[sa_n], par_n);
    const T* Ap = ringA + sa_p * A_ELEMS;
    const T* Ac = ringA + sa_c * A_ELEMS;
    const T* An = ringA + sa_n * A_ELEMS;
    T* Bc = ringB + sb * B_ELEMS;
    const bool plane_free = p >= 1 && p <= a.n0 - 2;
    // ---- time t+1 on rows -1 .. TY, cols -1 .. W of plane p
    {
      const bool fx = !plane_free || row_fixed_own;
      t1_group(Ap, Ac, An, Bc, warp, fx);
      if (lane == 0) t1_cell(Ap, Ac, An, Bc, warp, -1, fx);
      if (lane == 31) t1_cell(Ap, Ac, An, Bc, warp, Tile::W, fx);
    }
    if (warp < 2) {                              // the two extra rows
      const bool fx = !plane_free || row_fixed_extra;
      t1_group(Ap, Ac, An, Bc, r_extra, fx);
      if (lane == 0) t1_cell(Ap, Ac, An, Bc, r_extra, -1, fx);
      if (lane == 31) t1_cell(Ap, Ac, An, Bc, r_extra, Tile::W, fx);
    }
    __syncthreads();                             // ring-B plane p complete; ring-A plane p-1 no longer needed
    if (threadIdx.x == 0 && it + TMA2_STAGES_A < nA) issue(sa_p, it + TMA2_STAGES_A);
    // ---- time t+2 on the TY x W tile of plane q = p-1 from ring-B planes q-1, q, q+1
    if (it >= 2) {
      const T* Bq_l = ringB + ((sb + 2) & 3) * B_ELEMS;   // plane p-2
      const T* Bq = ringB + ((sb + 3) & 3) * B_ELEMS;     // plane p-1
      const T* Bq_h = Bc;                                  // plane p
      const int o = offB(warp, c0);
      const Group<T, E> cc = *reinterpret_cast<const Group<T, E>*>(Bq + o);
      Group<T, E> res = cc;
      if (!row_fixed_own) {
        const Group<T, E> zl = *reinterpret_cast<const Group<T, E>*>(Bq_l + o);
        const Group<T, E> zh = *reinterpret_cast<const Group<T, E>*>(Bq_h + o);
        const Group<T, E> up = *reinterpret_cast<const Group<T, E>*>(Bq + o - PITCH);
        const Group<T, E> dn = *reinterpret_cast<const Group<T, E>*>(Bq + o + PITCH);
        const T xl = Bq[o - 1], xr = Bq[o + E];
#pragma unroll
        for (int i = 0; i < E; i++) {
          const T l = (i > 0) ? cc.v[i - 1] : xl;
          const T rr = (i < E - 1) ? cc.v[i + 1] : xr;
          res.v[i] = heat7<T>(cc.v[i], zl.v[i], zh.v[i], up.v[i], dn.v[i], l, rr, a.coeff);
        }
        if (fix_first) res.v[0] = cc.v[0];
        if (fix_last) res.v[E - 1] = cc.v[E - 1];
      }
      if (active) store_group<T, E>(p_out, res);
      p_out += plane;
    }
    sa_p = sa_c; sa_c = sa_n;
    if (++sa_n == TMA2_STAGES_A) { sa_n = 0; par_n ^= 1; }
    sb = (sb + 1) & 3;
  }
}

// ---------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// returns PH_OK and *used = true when the TMA kernel ran; *used = false => caller falls back
template <typename T, int TMA_TY>
static int32_t heat_tma_launch(const T* in, T* out, int64_t n0, int64_t n1, int64_t n2, T coeff, int64_t z_begin,
                               int64_t z_end, cudaStream_t stream, bool* used) {
  using Tile = TmaTile<T, TMA_TY>;
  *used = false;
  if (z_begin >= z_end) { *used = true; return PH_OK; }
  static const bool disabled = getenv("PH_HEAT_NO_TMA") != nullptr;
  if (disabled) return PH_OK;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return PH_OK;
  // TMA constraints: 16-byte aligned base and row pitch, extents within 32-bit coordinates
  if ((uintptr_t)in % 16 || (uintptr_t)out % 16 || (n2 * sizeof(T)) % 16 || n2 % Tile::E) return PH_OK;
  if (n0 > 0x7fffffff || n1 > 0x7fffffff || n2 > 0x7fffffff) return PH_OK;
  CUtensorMap map;
  const cuuint64_t dims[3] = {(cuuint64_t)n2, (cuuint64_t)n1, (cuuint64_t)n0};
  const cuuint64_t strides[2] = {(cuuint64_t)n2 * sizeof(T), (cuuint64_t)n1 * n2 * sizeof(T)};
  const cuuint32_t box[3] = {(cuuint32_t)Tile::PITCH, (cuuint32_t)Tile::ROWS, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapDataType dt = sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
  CUresult r = enc(&map, dt, 3, const_cast<T*>(in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return PH_OK;             // shape not encodable: register kernel
  HeatTmaArgs<T> a;
  a.out = out; a.n0 = n0; a.n1 = n1; a.n2 = n2; a.coeff = coeff;
  a.z_begin = z_begin; a.z_end = z_end;
  const int64_t gx = ceil_div(n2, (int64_t)Tile::W), gy = ceil_div(n1, (int64_t)TMA_TY);
  const int64_t planes = z_end - z_begin;
  // ~6 resident blocks per SM (shared memory bound) x ~6 waves, but chunks of >= 48 planes so the
  // ring prologue (STAGES planes) stays a few percent
  const int64_t want = (int64_t)rt().sm_count * std::max(1, 48 / TMA_TY) * 6;
  int64_t gz = std::max<int64_t>(1, std::min<int64_t>(ceil_div(want, gx * gy), ceil_div(planes, 48)));
  a.z_chunk = ceil_div(planes, gz);
  gz = ceil_div(planes, a.z_chunk);
  if (gy > 65535 || gz > 65535) return PH_OK;
  const size_t smem = (size_t)TMA_STAGES * Tile::STAGE_BYTES;
  static bool attr_set = false;
  if (!attr_set) {
    PH_CUDA(cudaFuncSetAttribute(heat_tma_kernel<T, TMA_TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)gz), block(32 * TMA_TY);
  heat_tma_kernel<T, TMA_TY><<<grid, block, smem, stream>>>(map, a);
  PH_LAUNCH_CHECK("heat_tma_kernel");
  *used = true;
  return PH_OK;
}

// two steps per pass; *used = false => caller runs two single steps instead
template <typename T>
int32_t heat_tma2_planes(const T* in, T* out, int64_t n0, int64_t n1, int64_t n2, T coeff, int64_t z_begin,
                         int64_t z_end, cudaStream_t stream, bool* used) {
  constexpr int TY = 16;
  using Tile = Tma2Tile<T, TY>;
  *used = false;
  if (z_begin >= z_end) { *used = true; return PH_OK; }
  static const bool disabled = getenv("PH_HEAT_NO_TMA") != nullptr || getenv("PH_HEAT_NO_FUSE2") != nullptr;
  if (disabled) return PH_OK;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return PH_OK;
  if ((uintptr_t)in % 16 || (uintptr_t)out % 16 || (n2 * sizeof(T)) % 16 || n2 % Tile::E) return PH_OK;
  if (n0 > 0x7fffffff || n1 > 0x7fffffff || n2 > 0x7fffffff) return PH_OK;
  CUtensorMap map;
  const cuuint64_t dims[3] = {(cuuint64_t)n2, (cuuint64_t)n1, (cuuint64_t)n0};
  const cuuint64_t strides[2] = {(cuuint64_t)n2 * sizeof(T), (cuuint64_t)n1 * n2 * sizeof(T)};
  const cuuint32_t box[3] = {(cuuint32_t)Tile::PITCH, (cuuint32_t)Tile::ROWS_A, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapDataType dt = sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
  CUresult r = enc(&map, dt, 3, const_cast<T*>(in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return PH_OK;
  HeatTmaArgs<T> a;
  a.out = out; a.n0 = n0; a.n1 = n1; a.n2 = n2; a.coeff = coeff;
  a.z_begin = z_begin; a.z_end = z_end;
  const int64_t gx = ceil_div(n2, (int64_t)Tile::W), gy = ceil_div(n1, (int64_t)TY);
  const int64_t planes = z_end - z_begin;
  // 2 resident blocks per SM x ~6 waves; chunks of >= 64 planes (each chunk recomputes 2 planes of t+1)
  const int64_t want = (int64_t)rt().sm_count * 2 * 6;
  int64_t gz = std::max<int64_t>(1, std::min<int64_t>(ceil_div(want, gx * gy), ceil_div(planes, 64)));
  a.z_chunk = ceil_div(planes, gz);
  gz = ceil_div(planes, a.z_chunk);
  if (gy > 65535 || gz > 65535) return PH_OK;
  static bool attr_set = false;
  if (!attr_set) {
    PH_CUDA(cudaFuncSetAttribute(heat_tma2_kernel<T, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tile::SMEM));
    attr_set = true;
  }
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)gz), block(32 * TY);
  heat_tma2_kernel<T, TY><<<grid, block, Tile::SMEM, stream>>>(map, a);
  PH_LAUNCH_CHECK("heat_tma2_kernel");
  *used = true;
  return PH_OK;
}

template int32_t heat_tma2_planes<float>(const float*, float*, int64_t, int64_t, int64_t, float, int64_t, int64_t,
                                         cudaStream_t, bool*);
template int32_t heat_tma2_planes<double>(const double*, double*, int64_t, int64_t, int64_t, double, int64_t, int64_t,
                                          cudaStream_t, bool*);

template <typename T>
int32_t heat_tma_planes(const T* in, T* out, int64_t n0, int64_t n1, int64_t n2, T coeff, int64_t z_begin,
                        int64_t z_end, cudaStream_t stream, bool* used) {
  static const int ty = getenv("PH_HEAT_TMA_ROWS") ? atoi(getenv("PH_HEAT_TMA_ROWS")) : 16;   // tuning knob (8 / 16 / 32 rows per tile)
  if (ty == 16) return heat_tma_launch<T, 16>(in, out, n0, n1, n2, coeff, z_begin, z_end, stream, used);
  if (ty == 32) return heat_tma_launch<T, 32>(in, out, n0, n1, n2, coeff, z_begin, z_end, stream, used);
  if (ty == 8) return heat_tma_launch<T, 8>(in, out, n0, n1, n2, coeff, z_begin, z_end, stream, used);
  return heat_tma_launch<T, 16>(in, out, n0, n1, n2, coeff, z_begin, z_end, stream, used);
}

template int32_t heat_tma_planes<float>(const float*, float*, int64_t, int64_t, int64_t, float, int64_t, int64_t,
                                        cudaStream_t, bool*);
template int32_t heat_tma_planes<double>(const double*, double*, int64_t, int64_t, int64_t, double, int64_t, int64_t,
                                         cudaStream_t, bool*);

}  // namespace ph
