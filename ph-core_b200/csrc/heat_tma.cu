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
#include "comm.cuh"
#include "ops.cuh"
#include <cuda.h>
#include <algorithm>
#include <type_traits>
#include <stdlib.h>

namespace ph {

// Packed f32x2 adds (FADD2) in the two-step stencil: bit-identical and 20 % fewer instructions
// (ncu: 1175 M -> 941 M warp instructions, issue 68 % -> 53 %), but SLOWER on B200 (2048^3: 1452 ->
// 1381 Gcell/s; barrier + short-scoreboard stalls up): FADD2 does not run at twice the FADD rate, so
// the FP pipe, not the issue slot, becomes the limit.  Kept for A/B runs (PH_HEAT_TB_CFG=5 / 6), off by default.
constexpr int TMA_STAGES = 6;       // planes resident in shared memory (3 in use + 3 in flight)

__device__ __forceinline__ void st_flag_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

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
// per cell-update drops from 8 to ~4.5 bytes (f32) and the kernel stops being HBM-bound.
// It is then bound by instruction issue, so the loop is built for a low instruction count:
//   * ring A: TMA-fed planes of time t with a 2-cell halo (as above);
//   * a thread owns TWO adjacent rows x one 16-byte group and marches along z keeping its own
//     cells of planes p-1, p, p+1 (time t) AND of planes p-2, p-1, p (time t+1) in registers:
//     per plane it reads only the new plane's own cells, the two rows above / below its pair
//     and the x-neighbour cells from shared memory (1 LDS per cell-update instead of 1.75);
//   * time t+1 of plane p goes to ring B (2 stages) only so that NEIGHBOUR threads can read
//     it; time t+2 of plane p-1 is computed in the same iteration and streamed to HBM;
//   * the (TY+2) x (W+2) halo of time t+1 is produced by two dedicated warps (one for the
//     rows -1 / TY, one for the columns -1 / W), so the TY/2 main warps never diverge;
//   * one __syncthreads per plane; role rotation by a 3x unrolled loop (no register copies).
// Every cell value is produced by exactly the same operations in the same order as two
// single-step passes (fixed cells are copied forward at t+1 just as a single step holds
// them), so the result is bit-identical.
template <typename T, int TY, int STAGES> struct Tma2Tile {
  static constexpr int E = 16 / (int)sizeof(T);
  static constexpr int W = 32 * E;
  static constexpr int PITCH = W + 2 * E;
  static constexpr int ROWS_A = TY + 4, ROWS_B = TY + 2;
  static constexpr int A_BYTES = ROWS_A * PITCH * (int)sizeof(T);
  static constexpr int A_STAGE = (A_BYTES + 127) / 128 * 128;
  static constexpr int B_STAGE = (ROWS_B * PITCH * (int)sizeof(T) + 127) / 128 * 128;
  static constexpr int SMEM = STAGES * A_STAGE + 2 * B_STAGE;
  static constexpr int OWN_WARPS = TY / 2;
  static constexpr int THREADS = 32 * (OWN_WARPS + 2);
};

template <typename T>
struct HeatTma2Args {
  T* out;
  int64_t n0, n1, n2;
  int64_t z_begin, z_end, z_chunk;
  int64_t fixed_lo, fixed_hi;        // planes <= fixed_lo or >= fixed_hi are held (global boundary planes)
  T coeff;
  HeatMirror mir;                    // sharded runs over peer-mapped slabs: compute + halo in one kernel
};

template <typename T>
__device__ __forceinline__ T heat7(T c, T zl, T zh, T yl, T yh, T xl, T xh, T coeff) {
  const T two_c = f_mul((T)2, c);
  const T d0 = f_add(f_sub(zl, two_c), zh);
  const T d1 = f_add(f_sub(yl, two_c), yh);
  const T d2 = f_add(f_sub(xl, two_c), xh);
  return f_add(c, f_mul(f_add(f_add(d0, d1), d2), coeff));
}

// ---- packed f32x2 arithmetic (sm_100: FADD2).  add / sub with an explicit .rn are single IEEE
// operations per lane exactly like __fadd_rn; there is NO packed multiply in heat_row_f32x2 (ptxas
// fuses mul.f32x2 + add.f32x2 into FFMA2 even under --fmad=false), the two multiplies per cell pair
// stay scalar __fmul_rn, which is never contracted.  Halves the FP issue slots of the stencil.
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2_sub(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// heat7 on one 4-cell group of floats, two cells per instruction where the operands are aligned
// register pairs (z and y neighbours, the centre); the x-neighbour differences are shifted by one
// cell and stay scalar.  2*c is computed as c + c (the same value, exactly).  Same operations in
// the same order per cell as heat7 => bit-identical.
template <bool EDGE>
__device__ __forceinline__ Group<float, 4> heat_row_f32x2(const Group<float, 4>& c, const Group<float, 4>& zl,
                                                          const Group<float, 4>& zh, const Group<float, 4>& yl,
                                                          const Group<float, 4>& yh, float xl, float xr, float coeff,
                                                          bool fix_first, bool fix_last) {
  const uint64_t c0 = c.raw.q[0], c1 = c.raw.q[1];
  const uint64_t tc0 = f2_add(c0, c0), tc1 = f2_add(c1, c1);
  const uint64_t d0a = f2_add(f2_sub(zl.raw.q[0], tc0), zh.raw.q[0]), d0b = f2_add(f2_sub(zl.raw.q[1], tc1), zh.raw.q[1]);
  const uint64_t d1a = f2_add(f2_sub(yl.raw.q[0], tc0), yh.raw.q[0]), d1b = f2_add(f2_sub(yl.raw.q[1], tc1), yh.raw.q[1]);
  float t0, t1, t2, t3;
  f2_unpack(tc0, t0, t1);
  f2_unpack(tc1, t2, t3);
  const float e0 = f_add(f_sub(xl, t0), c.v[1]);
  const float e1 = f_add(f_sub(c.v[0], t1), c.v[2]);
  const float e2 = f_add(f_sub(c.v[1], t2), c.v[3]);
  const float e3 = f_add(f_sub(c.v[2], t3), xr);
  const uint64_t lapa = f2_add(f2_add(d0a, d1a), f2_pack(e0, e1));
  const uint64_t lapb = f2_add(f2_add(d0b, d1b), f2_pack(e2, e3));
  float l0, l1, l2, l3;
  f2_unpack(lapa, l0, l1);
  f2_unpack(lapb, l2, l3);
  Group<float, 4> res;
  res.raw.q[0] = f2_add(c0, f2_pack(f_mul(l0, coeff), f_mul(l1, coeff)));
  res.raw.q[1] = f2_add(c1, f2_pack(f_mul(l2, coeff), f_mul(l3, coeff)));
  if (EDGE) {
    if (fix_first) res.v[0] = c.v[0];
    if (fix_last) res.v[3] = c.v[3];
  }
  return res;
}

template <typename T, int E, bool EDGE, bool X2 = false>
__device__ __forceinline__ Group<T, E> heat_row(const Group<T, E>& c, const Group<T, E>& zl, const Group<T, E>& zh,
                                                const Group<T, E>& yl, const Group<T, E>& yh, T xl, T xr, T coeff,
                                                bool fix_first, bool fix_last) {
  if constexpr (std::is_same<T, float>::value && E == 4 && X2) {
    return heat_row_f32x2<EDGE>(c, zl, zh, yl, yh, xl, xr, coeff, fix_first, fix_last);
  }
  Group<T, E> res;
#pragma unroll
  for (int i = 0; i < E; i++) {
    const T l = (i > 0) ? c.v[i - 1] : xl;
    const T r = (i < E - 1) ? c.v[i + 1] : xr;
    res.v[i] = heat7<T>(c.v[i], zl.v[i], zh.v[i], yl.v[i], yh.v[i], l, r, coeff);
  }
  if (EDGE) {                                   // first / last column of the grid is held
    if (fix_first) res.v[0] = c.v[0];
    if (fix_last) res.v[E - 1] = c.v[E - 1];
  }
  return res;
}

template <typename T, int TY, int STAGES, bool X2 = false, bool LDSX = true>
__global__ void __launch_bounds__(Tma2Tile<T, TY, STAGES>::THREADS, (TY <= 16 ? 2 : 1))
heat_tma2_kernel(const __grid_constant__ CUtensorMap in_map, const HeatTma2Args<T> a) {
  using Tile = Tma2Tile<T, TY, STAGES>;
  using G = Group<T, Tile::E>;
  constexpr int E = Tile::E, PITCH = Tile::PITCH, W = Tile::W;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t full[STAGES];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t tile_x = (int64_t)blockIdx.x * W;
  const int64_t tile_y = (int64_t)blockIdx.y * TY;
  // sharded runs: the chunks at either end of the slab produce the planes the neighbours need, so they are
  // numbered first (blocks are dispatched in blockIdx order) and the halo leaves in the first wave
  int64_t zc = blockIdx.z;
  if (a.mir.edges_first && gridDim.z > 2) zc = blockIdx.z == 0 ? 0 : (blockIdx.z == 1 ? (int64_t)gridDim.z - 1 : (int64_t)blockIdx.z - 1);
  const int64_t zb = a.z_begin + zc * a.z_chunk;
  const int64_t ze = (zb + a.z_chunk < a.z_end) ? zb + a.z_chunk : a.z_end;
  if (zb >= ze) return;
  const T* const ringA = reinterpret_cast<const T*>(smem_raw);
  T* const ringB = reinterpret_cast<T*>(smem_raw + (size_t)STAGES * Tile::A_STAGE);
  constexpr int A_ELEMS = Tile::A_STAGE / (int)sizeof(T), B_ELEMS = Tile::B_STAGE / (int)sizeof(T);

  // time-t planes zb-2 .. ze+1 stream through ring A in order; k = plane - (zb-2) lives in stage k % STAGES
  const int nA = (int)(ze - zb) + 4;
  auto issue = [&](int stage, int k) {
    mbar_expect_tx(&full[stage], (uint32_t)Tile::A_BYTES);
    tma_load_3d(smem_raw + (size_t)stage * Tile::A_STAGE, &in_map, (int)(tile_x - E), (int)(tile_y - 2),
                (int)(zb - 2 + k), &full[stage]);
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int k = 0; k < STAGES && k < nA; k++) issue(k, k);

  // tile-local row r in [-2, TY+1], column c in [-E, W+E-1]
  auto offA = [](int r, int c) { return (r + 2) * PITCH + (c + E); };
  auto offB = [](int r, int c) { return (r + 1) * PITCH + (c + E); };
  const int64_t plane = a.n1 * a.n2;
  const int iters = (int)(ze - zb) + 2;          // t+1 planes p = zb-1 .. ze; output plane q = p-1 once it >= 2

  // ring bookkeeping shared by the three roles (identical in every thread)
  int sc = 1, sn = 2;                            // stages of time-t planes p, p+1
  uint32_t par_n = 0;
  int bsel = 0;                                  // ring-B stage receiving time-(t+1) plane p
  int it = 0;
  int64_t p = zb - 1;
  auto advance = [&]() {                         // end of a plane: recycle A[p]'s stage, rotate
    __syncthreads();
    if (threadIdx.x == 0 && it + STAGES + 1 < nA) issue(sc, it + STAGES + 1);
    sc = sn;
    if (++sn == STAGES) { sn = 0; par_n ^= 1; }
    bsel ^= 1;
    it++; p++;
  };
  // prologue common to all roles: planes zb-2 and zb-1 have landed; own cells are read by the
  // caller, then stage 0 is recycled
  mbar_wait(&full[0], 0);
  mbar_wait(&full[1], 0);

  if (warp < Tile::OWN_WARPS + 1) {
    // ------------------------------------------------------------ group rows: a pair of rows x one group per thread
    const bool own = warp < Tile::OWN_WARPS;     // own rows 2w, 2w+1 (t+1 and t+2)  /  halo rows -1, TY (t+1 only)
    const int r0 = own ? 2 * warp : -1, r1 = own ? 2 * warp + 1 : TY;
    const int c0 = lane * E;
    const int oa0 = offA(r0, c0), oa1 = offA(r1, c0), ob0 = offB(r0, c0), ob1 = offB(r1, c0);
    const int64_t gx = tile_x + c0, gy0 = tile_y + r0, gy1 = tile_y + r1;
    const bool fix_first = gx == 0, fix_last = gx + E == a.n2;
    const bool rowfix0 = gy0 == 0 || gy0 == a.n1 - 1, rowfix1 = gy1 == 0 || gy1 == a.n1 - 1;
    const bool act0 = own && gx < a.n2 && gy0 < a.n1, act1 = own && gx < a.n2 && gy1 < a.n1;
    T* p_out0 = a.out + zb * plane + gy0 * a.n2 + gx;
    T* p_out1 = p_out0 + a.n2;
    const T coeff = a.coeff;
    // warp-uniform: both rows lie strictly inside the grid and the tile is not the first / last
    // in x, so no cell of this warp is ever held and every store is in range
    const bool fast = own && gy0 > 0 && gy1 < a.n1 - 1 && tile_x > 0 && tile_x + W < a.n2;

    // x-neighbours of a thread's group: the last cell of the lane to the left, the first of the lane to the right.
    // Default (LDSX): one scalar LDS per neighbour -- a 4-way bank conflict (lanes are 4 words apart), which makes
    // the shared-memory pipe the busiest unit (ncu: 81 % of its wavefront peak, l1tex 93 %).  The shuffle form
    // (PH_HEAT_TB_CFG=8: neighbour lanes' registers, only lanes 0 / 31 read the adjacent tile's halo column)
    // cuts the wavefronts by 26 % and the conflicts by 69 %, but costs 12 % more instructions (selects, the
    // predicated edge loads) and the kernel is then issue-bound: 1.637 ms vs 1.568 ms on a (256, 2048, 2048)
    // slab, 1182 vs 1220 Gcell/s sustained (profiles/r02_ncu_heat_shuffle_vs_lds.csv).  Measured, kept off.
    const bool edge_lane = lane == 0 || lane == 31;
    const int eoff = lane == 0 ? -1 : E;
    auto x_neighbours = [&](const T* S, int o0, int o1, const G& c0v, const G& c1v, T& xl0, T& xr0, T& xl1, T& xr1) {
      if constexpr (sizeof(T) == 4 && !LDSX) {
        xl0 = __shfl_up_sync(0xffffffffu, c0v.v[E - 1], 1);
        xr0 = __shfl_down_sync(0xffffffffu, c0v.v[0], 1);
        xl1 = __shfl_up_sync(0xffffffffu, c1v.v[E - 1], 1);
        xr1 = __shfl_down_sync(0xffffffffu, c1v.v[0], 1);
        if (edge_lane) {
          const T e0 = S[o0 + eoff], e1 = S[o1 + eoff];
          if (lane == 0) { xl0 = e0; xl1 = e1; } else { xr0 = e0; xr1 = e1; }
        }
      } else {
        xl0 = S[o0 - 1]; xr0 = S[o0 + E]; xl1 = S[o1 - 1]; xr1 = S[o1 + E];
      }
    };

    G a0[3], a1[3], b0[3], b1[3];                // rotating roles: [P, C, N] of time t (a) and time t+1 (b)
    a0[0] = *reinterpret_cast<const G*>(ringA + oa0);
    a1[0] = *reinterpret_cast<const G*>(ringA + oa1);
    a0[1] = *reinterpret_cast<const G*>(ringA + A_ELEMS + oa0);
    a1[1] = *reinterpret_cast<const G*>(ringA + A_ELEMS + oa1);
    b0[0] = a0[1]; b1[0] = a1[1]; b0[1] = a0[1]; b1[1] = a1[1];      // placeholders until two t+1 planes exist
    __syncthreads();
    if (threadIdx.x == 0 && STAGES < nA) issue(0, STAGES);

    // OWN: adjacent rows (the pair shares its middle neighbours through registers) and the t+2
    // phase; !OWN: the two halo rows.  FAST: no held cell, no range test (see `fast`); the first
    // and the last plane of a march always take the general form (they may be held planes).
    auto step = [&](auto own_t, auto fast_t, auto mir_t, G& aP0, G& aP1, G& aC0, G& aC1, G& aN0, G& aN1, G& bP0, G& bP1,
                    G& bC0, G& bC1, G& bN0, G& bN1) {
      constexpr bool OWN = decltype(own_t)::value, FAST = decltype(fast_t)::value, MIR = decltype(mir_t)::value;
      mbar_wait(&full[sn], par_n);               // planes land in order: p+1 here => p here
      const T* An = ringA + sn * A_ELEMS;
      const T* Ac = ringA + sc * A_ELEMS;
      aN0 = *reinterpret_cast<const G*>(An + oa0);
      aN1 = *reinterpret_cast<const G*>(An + oa1);
      // ---- time t+1 of plane p, rows r0 / r1
      {
        const G up0 = *reinterpret_cast<const G*>(Ac + oa0 - PITCH);
        const G dn1 = *reinterpret_cast<const G*>(Ac + oa1 + PITCH);
        T xl0, xr0, xl1, xr1;
        x_neighbours(Ac, oa0, oa1, aC0, aC1, xl0, xr0, xl1, xr1);
        if constexpr (OWN && FAST) {
          bN0 = heat_row<T, E, false, X2>(aC0, aP0, aN0, up0, aC1, xl0, xr0, coeff, false, false);
          bN1 = heat_row<T, E, false, X2>(aC1, aP1, aN1, aC0, dn1, xl1, xr1, coeff, false, false);
        } else {
          const bool plane_held = p <= a.fixed_lo || p >= a.fixed_hi;
          G dn0, up1;
          if constexpr (OWN) { dn0 = aC1; up1 = aC0; }
          else { dn0 = *reinterpret_cast<const G*>(Ac + oa0 + PITCH); up1 = *reinterpret_cast<const G*>(Ac + oa1 - PITCH); }
          const G t0 = heat_row<T, E, true, X2>(aC0, aP0, aN0, up0, dn0, xl0, xr0, coeff, fix_first, fix_last);
          const G t1 = heat_row<T, E, true, X2>(aC1, aP1, aN1, up1, dn1, xl1, xr1, coeff, fix_first, fix_last);
          bN0 = (plane_held || rowfix0) ? aC0 : t0;
          bN1 = (plane_held || rowfix1) ? aC1 : t1;
        }
      }
      T* Bw = ringB + bsel * B_ELEMS;
      *reinterpret_cast<G*>(Bw + ob0) = bN0;
      *reinterpret_cast<G*>(Bw + ob1) = bN1;
      // ---- time t+2 of plane q = p-1 (own rows only): z-neighbours from registers, the rest from ring B
      if (OWN && (FAST || it >= 2)) {            // a FAST step is never one of the first two planes
        const T* Bq = ringB + (bsel ^ 1) * B_ELEMS;
        const G up0 = *reinterpret_cast<const G*>(Bq + ob0 - PITCH);
        const G dn1 = *reinterpret_cast<const G*>(Bq + ob1 + PITCH);
        T xl0, xr0, xl1, xr1;
        x_neighbours(Bq, ob0, ob1, bC0, bC1, xl0, xr0, xl1, xr1);
        // MIR: output plane q = p-1 is one of the planes a neighbour slab keeps as ghosts -> the same group
        // also goes to that neighbour's memory over NVLink (a plain store through the peer mapping)
        auto mirror_delta = [&]() -> int64_t {
          return (p - 1 < a.mir.lo_end) ? a.mir.delta_lo : ((p - 1 >= a.mir.hi_begin) ? a.mir.delta_hi : 0);
        };
        if constexpr (FAST) {
          const G res0 = heat_row<T, E, false, X2>(bC0, bP0, bN0, up0, bC1, xl0, xr0, coeff, false, false);
          const G res1 = heat_row<T, E, false, X2>(bC1, bP1, bN1, bC0, dn1, xl1, xr1, coeff, false, false);
          store_group<T, E>(p_out0, res0);
          store_group<T, E>(p_out1, res1);
          if constexpr (MIR) {
            if (const int64_t mdelta = mirror_delta()) {
              store_group<T, E>(reinterpret_cast<T*>(reinterpret_cast<char*>(p_out0) + mdelta), res0);
              store_group<T, E>(reinterpret_cast<T*>(reinterpret_cast<char*>(p_out1) + mdelta), res1);
            }
          }
        } else {
          const G t0 = heat_row<T, E, true, X2>(bC0, bP0, bN0, up0, bC1, xl0, xr0, coeff, fix_first, fix_last);
          const G t1 = heat_row<T, E, true, X2>(bC1, bP1, bN1, bC0, dn1, xl1, xr1, coeff, fix_first, fix_last);
          if (act0) store_group<T, E>(p_out0, rowfix0 ? bC0 : t0);
          if (act1) store_group<T, E>(p_out1, rowfix1 ? bC1 : t1);
          if constexpr (MIR) {
            if (const int64_t mdelta = mirror_delta()) {
              if (act0) store_group<T, E>(reinterpret_cast<T*>(reinterpret_cast<char*>(p_out0) + mdelta), rowfix0 ? bC0 : t0);
              if (act1) store_group<T, E>(reinterpret_cast<T*>(reinterpret_cast<char*>(p_out1) + mdelta), rowfix1 ? bC1 : t1);
            }
          }
        }
        p_out0 += plane; p_out1 += plane;
      }
      advance();
    };
#define PH_ROT0 a0[0], a1[0], a0[1], a1[1], a0[2], a1[2], b0[0], b1[0], b0[1], b1[1], b0[2], b1[2]
#define PH_ROT1 a0[1], a1[1], a0[2], a1[2], a0[0], a1[0], b0[1], b1[1], b0[2], b1[2], b0[0], b1[0]
#define PH_ROT2 a0[2], a1[2], a0[0], a1[0], a0[1], a1[1], b0[2], b1[2], b0[0], b1[0], b0[1], b1[1]
    auto run = [&](auto own_t, auto fast_t, auto mir_t) {
      const std::false_type general;
      step(own_t, general, mir_t, PH_ROT0);      // it = 0: plane zb-1 may be a held plane, no output yet
      step(own_t, general, mir_t, PH_ROT1);      // it = 1: no output yet
      while (iters - it >= 4) {
        step(own_t, fast_t, mir_t, PH_ROT2);
        step(own_t, fast_t, mir_t, PH_ROT0);
        step(own_t, fast_t, mir_t, PH_ROT1);
      }
      const int rem = iters - it;                // 1..3 planes left (iters >= 3); the last one may be a held plane
      if (rem == 1) {
        step(own_t, general, mir_t, PH_ROT2);
      } else if (rem == 2) {
        step(own_t, fast_t, mir_t, PH_ROT2);
        step(own_t, general, mir_t, PH_ROT0);
      } else {
        step(own_t, fast_t, mir_t, PH_ROT2);
        step(own_t, fast_t, mir_t, PH_ROT0);
        step(own_t, general, mir_t, PH_ROT1);
      }
    };
    // block-uniform: does this march produce planes that are ALSO stored into a neighbour's ghost planes?
    const bool mirror = own && (zb < a.mir.lo_end || ze > a.mir.hi_begin);
    if (fast) { if (mirror) run(std::true_type{}, std::true_type{}, std::true_type{}); else run(std::true_type{}, std::true_type{}, std::false_type{}); }
    else if (own) { if (mirror) run(std::true_type{}, std::false_type{}, std::true_type{}); else run(std::true_type{}, std::false_type{}, std::false_type{}); }
    else run(std::false_type{}, std::false_type{}, std::false_type{});
#undef PH_ROT0
#undef PH_ROT1
#undef PH_ROT2
  } else {
    // ------------------------------------------------------------ edge columns -1 and W of time t+1, rows -1 .. TY
    // (never grid-boundary columns: tiles start on multiples of W and n2 % E == 0)
    const bool work = lane < TY + 2;
    const int r = work ? lane - 1 : 0;
    const int oa0 = offA(r, -1), oa1 = offA(r, W), ob0 = offB(r, -1), ob1 = offB(r, W);
    const int64_t gy = tile_y + r;
    const bool rowfix = gy == 0 || gy == a.n1 - 1;
    const T coeff = a.coeff;
    T e0[3], e1[3];
    e0[0] = ringA[oa0]; e1[0] = ringA[oa1];
    e0[1] = ringA[A_ELEMS + oa0]; e1[1] = ringA[A_ELEMS + oa1];
    __syncthreads();
    auto step = [&](T& eP0, T& eP1, T& eC0, T& eC1, T& eN0, T& eN1) {
      mbar_wait(&full[sn], par_n);
      const T* An = ringA + sn * A_ELEMS;
      const T* Ac = ringA + sc * A_ELEMS;
      eN0 = An[oa0]; eN1 = An[oa1];
      T v0 = eC0, v1 = eC1;
      if (!(p <= a.fixed_lo || p >= a.fixed_hi) && !rowfix) {
        v0 = heat7<T>(eC0, eP0, eN0, Ac[oa0 - PITCH], Ac[oa0 + PITCH], Ac[oa0 - 1], Ac[oa0 + 1], coeff);
        v1 = heat7<T>(eC1, eP1, eN1, Ac[oa1 - PITCH], Ac[oa1 + PITCH], Ac[oa1 - 1], Ac[oa1 + 1], coeff);
      }
      if (work) {
        T* Bw = ringB + bsel * B_ELEMS;
        Bw[ob0] = v0; Bw[ob1] = v1;
      }
      advance();
    };
    while (it + 3 <= iters) {
      step(e0[0], e1[0], e0[1], e1[1], e0[2], e1[2]);
      step(e0[1], e1[1], e0[2], e1[2], e0[0], e1[0]);
      step(e0[2], e1[2], e0[0], e1[0], e0[1], e1[1]);
    }
    if (it < iters) {
      step(e0[0], e1[0], e0[1], e1[1], e0[2], e1[2]);
      if (it < iters) step(e0[1], e1[1], e0[2], e1[2], e0[0], e1[0]);
    }
  }
  // ---- halo signal: when the LAST block that stored into a neighbour's ghost planes is done, that
  // neighbour's flag word receives this pass's event number (release at system scope after every block's
  // own system fence), and its next pass -- stream-ordered behind a wait on that word -- may start
  // (the march bounds are recomputed from blockIdx here rather than kept in registers across the loops)
  int64_t zc2 = blockIdx.z;
  if (a.mir.edges_first && gridDim.z > 2) zc2 = blockIdx.z == 0 ? 0 : (blockIdx.z == 1 ? (int64_t)gridDim.z - 1 : (int64_t)blockIdx.z - 1);
  const int64_t zb2 = a.z_begin + zc2 * a.z_chunk;
  const int64_t ze2 = (zb2 + a.z_chunk < a.z_end) ? zb2 + a.z_chunk : a.z_end;
  const bool covers_lo = zb2 < a.mir.lo_end, covers_hi = ze2 > a.mir.hi_begin;
  if (covers_lo || covers_hi) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      if (covers_lo && atomicAdd(&a.mir.ticket[0], 1u) == a.mir.lo_blocks - 1) {
        a.mir.ticket[0] = 0;
        __threadfence_system();
        st_flag_sys(a.mir.flag_lo, a.mir.event);
      }
      if (covers_hi && atomicAdd(&a.mir.ticket[1], 1u) == a.mir.hi_blocks - 1) {
        a.mir.ticket[1] = 0;
        __threadfence_system();
        st_flag_sys(a.mir.flag_hi, a.mir.event);
      }
    }
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
  static int attr_dev = -1;                         // function attributes are per device
  if (attr_dev != rt().device) {
    PH_CUDA(cudaFuncSetAttribute(heat_tma_kernel<T, TMA_TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_dev = rt().device;
  }
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)gz), block(32 * TMA_TY);
  heat_tma_kernel<T, TMA_TY><<<grid, block, smem, stream>>>(map, a);
  PH_LAUNCH_CHECK("heat_tma_kernel");
  *used = true;
  return PH_OK;
}

// two steps per pass; *used = false => caller runs two single steps instead
template <typename T, int TY, int STAGES, bool X2 = false, bool LDSX = true>
static int32_t heat_tma2_launch(const T* in, T* out, int64_t n0, int64_t n1, int64_t n2, T coeff, int64_t z_begin,
                                int64_t z_end, int64_t fixed_lo, int64_t fixed_hi, cudaStream_t stream, bool* used,
                                const HeatMirror* mir) {
  using Tile = Tma2Tile<T, TY, STAGES>;
  static_assert(TY % 2 == 0 && TY + 2 <= 32, "the edge-column warp covers rows -1 .. TY with one lane each");
  EncodeTiledFn enc = encode_fn();
  if (!enc) return PH_OK;
  CUtensorMap map;
  const cuuint64_t dims[3] = {(cuuint64_t)n2, (cuuint64_t)n1, (cuuint64_t)n0};
  const cuuint64_t strides[2] = {(cuuint64_t)n2 * sizeof(T), (cuuint64_t)n1 * n2 * sizeof(T)};
  const cuuint32_t box[3] = {(cuuint32_t)Tile::PITCH, (cuuint32_t)Tile::ROWS_A, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapDataType dt = sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
  CUresult r = enc(&map, dt, 3, const_cast<T*>(in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return PH_OK;
  HeatTma2Args<T> a;
  a.out = out; a.n0 = n0; a.n1 = n1; a.n2 = n2; a.coeff = coeff;
  a.z_begin = z_begin; a.z_end = z_end; a.fixed_lo = fixed_lo; a.fixed_hi = fixed_hi;
  const int64_t gx = ceil_div(n2, (int64_t)Tile::W), gy = ceil_div(n1, (int64_t)TY);
  const int64_t planes = z_end - z_begin;
  // Marches of ~128 planes: blocks of one wave start together and share their tile halos through
  // L2, but over a long march they drift apart and the halo rows come from DRAM again (measured on
  // 2048^3: 1451 Gcell/s with 128-plane marches, 1389 / 1296 / 1210 with 256 / 512 / 2046).  Each
  // chunk re-reads 4 planes and recomputes 2 planes of t+1, so never shorter than 32 planes; more
  // chunks when the tile count alone cannot fill the machine for ~6 waves.
  const int per_sm = std::max(1, std::min(2, (int)(220 * 1024 / Tile::SMEM)));
  const int64_t want = (int64_t)rt().sm_count * per_sm * 6;
  static const int64_t march = getenv("PH_HEAT_TB_MARCH") ? atoll(getenv("PH_HEAT_TB_MARCH")) : 128;   // tuning knob
  int64_t gz = std::max<int64_t>(ceil_div(planes, march), ceil_div(want, gx * gy));
  gz = std::max<int64_t>(1, std::min<int64_t>(gz, ceil_div(planes, 32)));
  a.z_chunk = ceil_div(planes, gz);
  gz = ceil_div(planes, a.z_chunk);
  if (gy > 65535 || gz > 65535) return PH_OK;
  if (mir) {                                        // how many blocks store into each neighbour's ghost planes
    a.mir = *mir;
    a.mir.edges_first = 1;
    a.mir.lo_blocks = a.mir.hi_blocks = 0;
    for (int64_t c = 0; c < gz; c++) {
      const int64_t zb = z_begin + c * a.z_chunk, ze = std::min(zb + a.z_chunk, z_end);
      if (zb < a.mir.lo_end) a.mir.lo_blocks += (uint32_t)(gx * gy);
      if (ze > a.mir.hi_begin) a.mir.hi_blocks += (uint32_t)(gx * gy);
    }
  }
  static int attr_dev = -1;                         // function attributes are per device
  if (attr_dev != rt().device) {
    PH_CUDA(cudaFuncSetAttribute(heat_tma2_kernel<T, TY, STAGES, X2, LDSX>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tile::SMEM));
    attr_dev = rt().device;
  }
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)gz), block(Tile::THREADS);
  heat_tma2_kernel<T, TY, STAGES, X2, LDSX><<<grid, block, Tile::SMEM, stream>>>(map, a);
  PH_LAUNCH_CHECK("heat_tma2_kernel");
  *used = true;
  return PH_OK;
}

// Planes [z_begin, z_end) of `out` receive time t+2; `in` holds time t on planes
// [z_begin-2, z_end+2) (missing planes outside the array are never used: planes <= fixed_lo and
// >= fixed_hi are held fixed, and a fixed plane ignores its neighbours).
template <typename T>
int32_t heat_tma2_planes(const T* in, T* out, int64_t n0, int64_t n1, int64_t n2, T coeff, int64_t z_begin,
                         int64_t z_end, int64_t fixed_lo, int64_t fixed_hi, cudaStream_t stream, bool* used,
                         const HeatMirror* mir) {
  *used = false;
  if (z_begin >= z_end) { *used = true; return PH_OK; }
  static const bool disabled = getenv("PH_HEAT_NO_TMA") != nullptr || getenv("PH_HEAT_NO_FUSE2") != nullptr;
  if (disabled) return PH_OK;
  constexpr int E = 16 / (int)sizeof(T);
  if ((uintptr_t)in % 16 || (uintptr_t)out % 16 || (n2 * sizeof(T)) % 16 || n2 % E) return PH_OK;
  if (n0 > 0x7fffffff || n1 > 0x7fffffff || n2 > 0x7fffffff) return PH_OK;
  static const int cfg = getenv("PH_HEAT_TB_CFG") ? atoi(getenv("PH_HEAT_TB_CFG")) : 0;     // tuning knob
  switch (cfg) {
    case 1: return heat_tma2_launch<T, 16, 4>(in, out, n0, n1, n2, coeff, z_begin, z_end, fixed_lo, fixed_hi, stream, used, mir);
    case 2: return heat_tma2_launch<T, 16, 8>(in, out, n0, n1, n2, coeff, z_begin, z_end, fixed_lo, fixed_hi, stream, used, mir);
    case 3: return heat_tma2_launch<T, 24, 5>(in, out, n0, n1, n2, coeff, z_begin, z_end, fixed_lo, fixed_hi, stream, used, mir);
    case 4: return heat_tma2_launch<T, 8, 6>(in, out, n0, n1, n2, coeff, z_begin, z_end, fixed_lo, fixed_hi, stream, used, mir);
    case 5: return heat_tma2_launch<T, 16, 6, true>(in, out, n0, n1, n2, coeff, z_begin, z_end, fixed_lo, fixed_hi, stream, used, mir);   // packed f32x2 adds
    case 6: return heat_tma2_launch<T, 16, 4, true>(in, out, n0, n1, n2, coeff, z_begin, z_end, fixed_lo, fixed_hi, stream, used, mir);
    case 8: return heat_tma2_launch<T, 16, 6, false, false>(in, out, n0, n1, n2, coeff, z_begin, z_end, fixed_lo, fixed_hi, stream, used, mir);   // x-neighbours by warp shuffle (see x_neighbours)
    case 7: return heat_tma2_launch<T, 24, 5, true>(in, out, n0, n1, n2, coeff, z_begin, z_end, fixed_lo, fixed_hi, stream, used, mir);   // one block per SM: no register cap
    default: return heat_tma2_launch<T, 16, 6>(in, out, n0, n1, n2, coeff, z_begin, z_end, fixed_lo, fixed_hi, stream, used, mir);
  }
}

// shape-only test (identical on every rank of a sharded run): can the two-step kernel take this grid?
template <typename T>
bool heat_tma2_usable(int64_t n1, int64_t n2) {
  constexpr int E = 16 / (int)sizeof(T);
  if (getenv("PH_HEAT_NO_TMA") != nullptr || getenv("PH_HEAT_NO_FUSE2") != nullptr) return false;
  if (!encode_fn()) return false;
  return n2 % E == 0 && n1 >= 3 && n2 >= 3 && n1 <= 0x7fffffff && n2 <= 0x7fffffff && ceil_div(n1, (int64_t)8) <= 65535;
}
template bool heat_tma2_usable<float>(int64_t, int64_t);
template bool heat_tma2_usable<double>(int64_t, int64_t);

template int32_t heat_tma2_planes<float>(const float*, float*, int64_t, int64_t, int64_t, float, int64_t, int64_t,
                                         int64_t, int64_t, cudaStream_t, bool*, const HeatMirror*);
template int32_t heat_tma2_planes<double>(const double*, double*, int64_t, int64_t, int64_t, double, int64_t, int64_t,
                                          int64_t, int64_t, cudaStream_t, bool*, const HeatMirror*);

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
