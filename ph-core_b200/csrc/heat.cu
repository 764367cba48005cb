// heat.cu -- explicit heat-diffusion stencil (K9 of SURVEY.md 2.3).
//
// Reference: examples/heat_equation.cr:38-51 (update_temp) and :26-36 (simulate).
//   PH_HEAT_EXAMPLE1D  the example verbatim: one-sided (zero-flux) ends,
//        d[0] = (s[1]-s[0])*C; d[n-1] = (s[n-2]-s[n-1])*C;
//        d[i] = ((s[i-1] - 2*s[i]) + s[i+1])*C;  s' = s + d
//   PH_HEAT_FIXED      the N-D rule SURVEY.md 8(a) a-9 defines in reference operators:
//        c = s[interior]; d_k = (s[lo_k] - 2*c) + s[hi_k]; lap = (d_0 + d_1) + d_2;
//        s'[interior] = c + lap*C; boundary cells are held.
// Every operator is one rounding in T (the __f*_rn intrinsics cannot be contracted), in
// exactly the association order above, so ONE step is bit-identical to the reference's
// operator-by-operator evaluation with materialised temporaries.
//
// Kernels.  `simulate` (ph_heat_run, ph_heat_run_sharded with two ghost planes) advances TWO time
// steps per pass over HBM: heat2d_tb_kernel below for rank 2 (warp-shuffle temporal blocking) and
// heat_tma2_kernel in heat_tma.cu for rank 3 (TMA ring + register marching).  A single step
// (`update_temp`, an odd last step, shapes the two-step kernels cannot take) uses heat_tma_kernel
// (rank 3) or the 2.5-D streaming kernel here: a block owns a (rows x columns) tile and marches along axis 0;
// each thread keeps its cells of planes z-1, z, z+1 in registers, so every cell is read from
// HBM once (8 B per cell-update for f32: one read + one write).  In-plane neighbours come
// from warp shuffles (x) and a double-buffered shared-memory row exchange (y); only the tile
// halo is re-read (L2 hits).  Slabs for multi-GPU runs are split along axis 0, so halo
// planes are contiguous and need no packing.
#include "ph_common.cuh"
#include "comm.cuh"
#include "ops.cuh"
#include <algorithm>
#include <type_traits>
#include <stdlib.h>

namespace ph {

constexpr int HEAT_TY = 8;            // warps per block (rows of the tile for rank 3)

template <typename T>
struct HeatArgs {
  const T* in;
  T* out;
  int64_t n0, n1, n2;                 // extents: axis 0 (march), axis 1 (rows; 1 for rank 2), axis 2 (x)
  int64_t z_begin, z_end;             // planes to update (interior only: 1 <= z < n0-1)
  int64_t z_chunk;                    // planes per block along the march axis
  T coeff;
};

template <typename T>
__device__ __forceinline__ T heat_cell(T c, T zl, T zh, T yl, T yh, T xl, T xh, T coeff, int rank) {
  const T two_c = f_mul((T)2, c);
  const T d0 = f_add(f_sub(zl, two_c), zh);
  T lap = d0;
  if (rank == 3) {
    const T d1 = f_add(f_sub(yl, two_c), yh);
    const T d2 = f_add(f_sub(xl, two_c), xh);
    lap = f_add(f_add(d0, d1), d2);
  } else if (rank == 2) {
    const T d1 = f_add(f_sub(xl, two_c), xh);
    lap = f_add(d0, d1);
  }
  return f_add(c, f_mul(lap, coeff));
}

// RANK 3: block = 32 lanes (x, E cells each) x HEAT_TY warps (consecutive y rows).
// RANK 2: block = 32 lanes x HEAT_TY warps, every warp owns its own x tile (no y axis).
//
// The loop is written for a low instruction count (the first version was issue-bound at
// ~48 instructions per cell): pointers advance by one plane per iteration, the y-halo rows
// travel through the same shared-memory exchange as the tile rows (rows 0 and TY+1), all
// halo loads are issued one plane ahead, inactive threads read clamped addresses instead of
// branching, and the fixed-boundary test is three per-thread flags.
template <typename T, int E, int RANK>
__global__ void __launch_bounds__(32 * HEAT_TY) heat_march_kernel(const HeatArgs<T> a) {
  __shared__ Group<T, E> rows[2][HEAT_TY + 2][32];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  int64_t x0, y;
  if (RANK == 3) {
    x0 = ((int64_t)blockIdx.x * 32 + lane) * E;
    y = (int64_t)blockIdx.y * HEAT_TY + warp;
  } else {
    x0 = (((int64_t)blockIdx.x * HEAT_TY + warp) * 32 + lane) * E;
    y = 0;
  }
  const int64_t zb = a.z_begin + (int64_t)blockIdx.z * a.z_chunk;
  const int64_t ze = (zb + a.z_chunk < a.z_end) ? zb + a.z_chunk : a.z_end;
  if (zb >= ze) return;
  const bool active = x0 < a.n2 && y < a.n1;       // n2 % E == 0 by dispatch: whole groups
  const int64_t xc = active ? x0 : a.n2 - E;       // clamped: inactive threads load valid data
  const int64_t yc = (y < a.n1) ? y : a.n1 - 1;
  const int64_t plane = a.n1 * a.n2;
  const bool fix_all = (RANK == 3) && (yc == 0 || yc == a.n1 - 1);
  const bool fix_first = (xc == 0), fix_last = (xc + E == a.n2);
  // y halo: warp 0 fetches the row above the tile, the last warp the row below it
  const bool halo_up = (RANK == 3) && warp == 0;
  const bool halo_dn = (RANK == 3) && warp == HEAT_TY - 1;
  const int64_t yh = halo_up ? (yc > 0 ? yc - 1 : yc) : (yc + 1 < a.n1 ? yc + 1 : yc);
  const int halo_row = halo_up ? 0 : HEAT_TY + 1;
  // x halo: lane 0 fetches the cell left of the warp's span, lane 31 the cell right of it
  const bool xh_left = lane == 0, xh_right = lane == 31;
  const int64_t xh = xh_left ? (xc > 0 ? xc - 1 : xc) : (xc + E < a.n2 ? xc + E : xc);

  const T* p_next = a.in + zb * plane + yc * a.n2 + xc;       // -> plane z of this thread's group
  const T* p_halo = a.in + zb * plane + yh * a.n2 + xc;       // -> plane z of the halo row
  const T* p_xh = a.in + zb * plane + yc * a.n2 + xh;         // -> plane z of the x-halo cell
  T* p_out = a.out + zb * plane + yc * a.n2 + xc;

  // Register pipeline over planes: P = z-1, C = z, N = z+1 are needed by the stencil at z; F = z+2
  // is requested now and first used one step later, so one full step of work hides its latency.
  const T* p_last = a.in + (a.n0 - 1) * plane + yc * a.n2 + xc;     // clamp for the final look-ahead
  Group<T, E> g0 = load_group<T, E>(p_next - plane);     // plane z-1
  Group<T, E> g1 = load_group<T, E>(p_next);             // plane z
  p_next += plane;
  Group<T, E> g2 = load_group<T, E>(p_next);             // plane z+1
  Group<T, E> g3;                                        // plane z+2
  Group<T, E> halo = g1;
  if (halo_up || halo_dn) halo = load_group_plain<T, E>(p_halo);
  T xhv = (xh_left || xh_right) ? *p_xh : (T)0;
  int buf = 0;
  int64_t z = zb;
  // one plane; the caller rotates which register group plays which role (four steps per trip),
  // so there are no register-to-register copies
  auto step = [&](const Group<T, E>& P, const Group<T, E>& C, const Group<T, E>& N, Group<T, E>& F) {
    p_next += plane; p_halo += plane; p_xh += plane;
    F = load_group<T, E>(p_next <= p_last ? p_next : p_last);
    Group<T, E> halo_n = C;
    T xhv_n = (T)0;
    if (z + 1 < ze) {
      if (halo_up || halo_dn) halo_n = load_group_plain<T, E>(p_halo);
      if (xh_left || xh_right) xhv_n = *p_xh;
    }
    Group<T, E> up = C, dn = C;
    if (RANK == 3) {
      rows[buf][warp + 1][lane] = C;
      if (halo_up || halo_dn) rows[buf][halo_row][lane] = halo;
      __syncthreads();
      up = rows[buf][warp][lane];
      dn = rows[buf][warp + 2][lane];
    }
    const T from_left = __shfl_up_sync(0xffffffffu, C.v[E - 1], 1);
    const T from_right = __shfl_down_sync(0xffffffffu, C.v[0], 1);
    const T xl = xh_left ? xhv : from_left;
    const T xr = xh_right ? xhv : from_right;
    Group<T, E> res = C;
    if (!fix_all) {                                      // warp-uniform
#pragma unroll
      for (int i = 0; i < E; i++) {
        const T l = (i > 0) ? C.v[i - 1] : xl;
        const T r = (i < E - 1) ? C.v[i + 1] : xr;
        res.v[i] = heat_cell<T>(C.v[i], P.v[i], N.v[i], up.v[i], dn.v[i], l, r, a.coeff, RANK);
      }
      if (fix_first) res.v[0] = C.v[0];
      if (fix_last) res.v[E - 1] = C.v[E - 1];
    }
    if (active) store_group<T, E>(p_out, res);
    p_out += plane;
    halo = halo_n;
    xhv = xhv_n;
    buf ^= 1;
    z++;
  };
  while (z + 4 <= ze) {
    step(g0, g1, g2, g3);
    step(g1, g2, g3, g0);
    step(g2, g3, g0, g1);
    step(g3, g0, g1, g2);
  }
  if (z < ze) {
    step(g0, g1, g2, g3);
    if (z < ze) {
      step(g1, g2, g3, g0);
      if (z < ze) step(g2, g3, g0, g1);
    }
  }
}

// ------------------------------------------------------------------ rank 2, two steps per pass
// Temporal blocking for the 2-D grid: one pass over HBM advances TWO time steps (4.3 instead of
// 8 bytes of DRAM traffic per cell-update for f32).  No shared memory at all: a WARP owns an
// x-segment of 32 groups and marches down axis 0; a lane keeps its group of rows p-1 .. p+4 at
// time t (three rows are loads in flight: 96 bytes per lane) and of rows p-2, p-1, p at time
// t+1 in registers, rotated by a 6x unrolled loop (no register copies); x-neighbours come from
// warp shuffles at both time levels (the default keeps 9 time-t rows, i.e. 6 loads in flight per
// lane).  Lanes 0 and 31 are halo lanes: they compute time t+1
// only (of which just the cell next to lane 1 / lane 30 is ever used, so THEIR outer neighbour
// is a don't-care and no halo loads exist); lanes 1..30 write time t+2.  Same operations in
// the same order as two single steps => bit-identical.
template <typename T>
struct Heat2dTbArgs {
  const T* in;
  T* out;
  int64_t n0, n2;                     // rows (march axis), columns
  int64_t z_begin, z_end, z_chunk;    // rows to update
  int64_t fixed_lo, fixed_hi;         // rows <= fixed_lo or >= fixed_hi are held
  int64_t tiles;                      // x tiles of 30 * E output cells
  T coeff;
};

template <typename T>
__device__ __forceinline__ T heat5(T c, T up, T dn, T xl, T xr, T coeff) {
  const T two_c = f_mul((T)2, c);
  const T d0 = f_add(f_sub(up, two_c), dn);
  const T d1 = f_add(f_sub(xl, two_c), xr);
  return f_add(c, f_mul(f_add(d0, d1), coeff));
}

template <typename T, int E, bool EDGE>
__device__ __forceinline__ Group<T, E> heat5_row(const Group<T, E>& c, const Group<T, E>& up, const Group<T, E>& dn,
                                                 T coeff, bool fix_first, bool fix_last) {
  const T xl = __shfl_up_sync(0xffffffffu, c.v[E - 1], 1);
  const T xr = __shfl_down_sync(0xffffffffu, c.v[0], 1);
  Group<T, E> res;
#pragma unroll
  for (int i = 0; i < E; i++) {
    const T l = (i > 0) ? c.v[i - 1] : xl;
    const T r = (i < E - 1) ? c.v[i + 1] : xr;
    res.v[i] = heat5<T>(c.v[i], up.v[i], dn.v[i], l, r, coeff);
  }
  if (EDGE) {
    if (fix_first) res.v[0] = c.v[0];
    if (fix_last) res.v[E - 1] = c.v[E - 1];
  }
  return res;
}

template <int N, typename F, int I = 0>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<N, F, I + 1>(static_cast<F&&>(f));
  }
}

// NA = time-t row slots per lane (a multiple of 3): rows p-1 .. p+NA-2, of which NA-3 are loads in flight.
template <typename T, int E, int MINB, int NA>
__global__ void __launch_bounds__(32 * HEAT_TY, MINB) heat2d_tb_kernel(const Heat2dTbArgs<T> a) {
  static_assert(NA % 3 == 0 && NA >= 6, "the t and t+1 register rings rotate together");
  using G = Group<T, E>;
  const int lane = threadIdx.x & 31;
  const int64_t tile = (int64_t)blockIdx.x * HEAT_TY + (threadIdx.x >> 5);
  if (tile >= a.tiles) return;                                   // whole warps only: shuffles stay full
  const int64_t zb = a.z_begin + (int64_t)blockIdx.y * a.z_chunk;
  const int64_t ze = (zb + a.z_chunk < a.z_end) ? zb + a.z_chunk : a.z_end;
  if (zb >= ze) return;
  const int64_t gx = tile * (30 * E) - E + (int64_t)lane * E;    // first column of this lane's group
  const bool in_range = gx >= 0 && gx + E <= a.n2;               // n2 % E == 0: a group is all in or all out
  const bool writes = in_range && lane >= 1 && lane <= 30;
  const bool fix_first = gx == 0, fix_last = gx + E == a.n2;
  const int64_t x_first = tile * (30 * E) - E;
  const bool fast = x_first > 0 && x_first + 32 * E < a.n2;      // warp-uniform: no held column, all lanes in range
  const T coeff = a.coeff;
  const T* col = a.in + (in_range ? gx : 0);
  T* p_out = a.out + zb * a.n2 + (in_range ? gx : 0);
  auto load_row = [&](int64_t r) {                               // rows outside the grid are never used: clamp
    r = r < 0 ? 0 : (r > a.n0 - 1 ? a.n0 - 1 : r);
    return load_group<T, E>(col + r * a.n2);
  };
  auto zero = [&]() { return splat_group<T, E>((T)0); };

  G ra[NA], rb[3];                                               // time t rows p-1 .. p+NA-2 ; time t+1 rows p-2 .. p
  int64_t p = zb - 1;
  int it = 0;
  const int iters = (int)(ze - zb) + 2;
#pragma unroll
  for (int k = 0; k < NA - 1; k++) ra[k] = in_range ? load_row(p - 1 + k) : zero();
  rb[0] = ra[1]; rb[1] = ra[1];                                  // placeholders until two t+1 rows exist

  // aP, aC, aN: time-t rows p-1, p, p+1; aF: receives row p+NA-2.  FAST: no held cell, every lane in
  // range; the first and last row of a march always take the general form (possibly held rows).
  auto step = [&](auto fast_t, G& aP, G& aC, G& aN, G& aF, G& bP, G& bC, G& bN) {
    constexpr bool FAST = decltype(fast_t)::value;
    if constexpr (FAST) aF = load_row(p + NA - 2);
    else aF = in_range ? load_row(p + NA - 2) : zero();
    if constexpr (FAST) {
      bN = heat5_row<T, E, false>(aC, aP, aN, coeff, false, false);
    } else {
      const bool held = p <= a.fixed_lo || p >= a.fixed_hi;
      const G t = heat5_row<T, E, true>(aC, aP, aN, coeff, fix_first, fix_last);
      bN = held ? aC : t;
    }
    if (FAST || it >= 2) {                                       // time t+2 of row q = p-1
      if constexpr (FAST) {
        const G res = heat5_row<T, E, false>(bC, bP, bN, coeff, false, false);
        if (lane >= 1 && lane <= 30) store_group<T, E>(p_out, res);
      } else {
        const G res = heat5_row<T, E, true>(bC, bP, bN, coeff, fix_first, fix_last);
        if (writes) store_group<T, E>(p_out, res);
      }
      p_out += a.n2;
    }
    it++; p++;
  };
  // rotation I = it mod NA (NA % 3 == 0, so the t+1 ring's rotation is I mod 3)
  auto step_rot = [&](auto fast_t, auto rot) {
    constexpr int I = decltype(rot)::value;
    step(fast_t, ra[I % NA], ra[(I + 1) % NA], ra[(I + 2) % NA], ra[(I + NA - 1) % NA], rb[I % 3], rb[(I + 1) % 3],
         rb[(I + 2) % 3]);
  };
  auto step_k = [&](auto fast_t, int k) {
    static_for<NA>([&](auto rot) { if (k == decltype(rot)::value) step_rot(fast_t, rot); });
  };
  auto run = [&](auto fast_t) {
    const std::false_type general;
    step_rot(general, std::integral_constant<int, 0>{});         // it = 0: row zb-1 may be held, no output yet
    step_rot(general, std::integral_constant<int, 1>{});         // it = 1: no output yet
    while (iters - it >= NA + 1)
      static_for<NA>([&](auto r) { step_rot(fast_t, std::integral_constant<int, (decltype(r)::value + 2) % NA>{}); });
    int k = 2;                                                   // 1..NA rows left (iters >= 3); the last may be held
    while (iters - it > 1) { step_k(fast_t, k); k = k + 1 == NA ? 0 : k + 1; }
    step_k(general, k);
  };
  if (fast) run(std::true_type{});
  else run(std::false_type{});
}

// rank 1, either boundary mode: one cell per thread
template <typename T>
__global__ void heat_1d_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t n, T coeff, int mode,
                               int64_t begin, int64_t end) {
  const int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= end) return;
  const T c = in[i];
  if (mode == PH_HEAT_EXAMPLE1D) {
    T d;
    if (i == 0) d = f_mul(f_sub(in[1], c), coeff);
    else if (i == n - 1) d = f_mul(f_sub(in[n - 2], c), coeff);
    else d = f_mul(f_add(f_sub(in[i - 1], f_mul((T)2, c)), in[i + 1]), coeff);
    out[i] = f_add(c, d);
  } else {
    if (i == 0 || i == n - 1) out[i] = c;
    else out[i] = f_add(c, f_mul(f_add(f_sub(in[i - 1], f_mul((T)2, c)), in[i + 1]), coeff));
  }
}

// rank 1, n <= 1024: the whole rod lives in shared memory for ALL steps (one launch instead
// of one per step: the reference example is 21 points x 10 001 steps).
template <typename T>
__global__ void __launch_bounds__(1024) heat_1d_resident_kernel(T* __restrict__ a_buf, T* __restrict__ b_buf,
                                                                 int n, T coeff, int mode, int64_t steps) {
  __shared__ T s[2][1024];
  const int i = threadIdx.x;
  if (i < n) s[0][i] = a_buf[i];
  __syncthreads();
  int cur = 0;
  for (int64_t t = 0; t < steps; t++) {
    if (i < n) {
      const T c = s[cur][i];
      T r;
      if (mode == PH_HEAT_EXAMPLE1D) {
        T d;
        if (i == 0) d = f_mul(f_sub(s[cur][1], c), coeff);
        else if (i == n - 1) d = f_mul(f_sub(s[cur][n - 2], c), coeff);
        else d = f_mul(f_add(f_sub(s[cur][i - 1], f_mul((T)2, c)), s[cur][i + 1]), coeff);
        r = f_add(c, d);
      } else {
        if (i == 0 || i == n - 1) r = c;
        else r = f_add(c, f_mul(f_add(f_sub(s[cur][i - 1], f_mul((T)2, c)), s[cur][i + 1]), coeff));
      }
      s[cur ^ 1][i] = r;
    }
    __syncthreads();
    cur ^= 1;
  }
  T* dst = (steps & 1) ? b_buf : a_buf;            // same convention as the ping-pong path
  if (i < n) dst[i] = s[cur][i];
}

// ------------------------------------------------------------------ host side
template <typename T, int RANK>
static int32_t launch_march(const T* in, T* out, int64_t n0, int64_t n1, int64_t n2, T coeff, int64_t z_begin,
                            int64_t z_end, cudaStream_t stream) {
  if (z_begin >= z_end) return PH_OK;
  HeatArgs<T> a;
  a.in = in; a.out = out; a.n0 = n0; a.n1 = n1; a.n2 = n2; a.coeff = coeff;
  a.z_begin = z_begin; a.z_end = z_end;
  constexpr int EW = 32 / (int)sizeof(T), EV = 16 / (int)sizeof(T);
  // widest group (256-bit, then 128-bit) that divides the row, keeps rows aligned and still
  // fills a warp; scalar otherwise
  const bool wide = (n2 % EW == 0) && n2 >= 32 * EW && ((uintptr_t)in % 32 == 0) && ((uintptr_t)out % 32 == 0);
  const bool vec = (n2 % EV == 0) && ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0);
  int e = wide ? EW : (vec ? EV : 1);
  if (const char* knob = getenv("PH_HEAT_GROUP_BYTES")) {      // tuning knob: 32 / 16 / 0 (scalar)
    const int b = atoi(knob);
    if (b == 16 && vec) e = EV;
    else if (b == 32 && !wide && vec) e = EV;
    else if (b == 0) e = 1;
  }
  int64_t gx, gy;
  if (RANK == 3) { gx = ceil_div(n2, (int64_t)32 * e); gy = ceil_div(n1, HEAT_TY); }
  else { gx = ceil_div(n2, (int64_t)32 * e * HEAT_TY); gy = 1; }
  // enough blocks for a few waves, but long marches so the 2-plane prologue stays negligible
  const int64_t planes = z_end - z_begin;
  const int64_t want = (int64_t)rt().sm_count * 8 * 4;
  int64_t gz = std::max<int64_t>(1, std::min<int64_t>(ceil_div(want, gx * gy), ceil_div(planes, 32)));
  a.z_chunk = ceil_div(planes, gz);
  gz = ceil_div(planes, a.z_chunk);
  if (gy > 65535 || gz > 65535) return set_error(PH_ERR_INVALID, "heat grid too large for one launch");
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)gz), block(32 * HEAT_TY);
  if (e == EW && EW != EV) heat_march_kernel<T, EW, RANK><<<grid, block, 0, stream>>>(a);
  else if (e == EV && EV > 1) heat_march_kernel<T, EV, RANK><<<grid, block, 0, stream>>>(a);
  else heat_march_kernel<T, 1, RANK><<<grid, block, 0, stream>>>(a);
  PH_LAUNCH_CHECK("heat_march_kernel");
  return PH_OK;
}

// rank 2, two steps per pass; *used = false => the caller runs two single steps instead
template <typename T>
static bool heat2d_tb_shape_ok(int64_t n2) {
  constexpr int E = 32 / (int)sizeof(T);
  return getenv("PH_HEAT_NO_FUSE2") == nullptr && n2 % E == 0 && n2 >= 3;
}

template <typename T, int E, int MINB, int NA>
static int32_t heat2d_tb_launch(const T* in, T* out, int64_t n0, int64_t n2, T coeff, int64_t z_begin, int64_t z_end,
                                int64_t fixed_lo, int64_t fixed_hi, cudaStream_t stream) {
  Heat2dTbArgs<T> a;
  a.in = in; a.out = out; a.n0 = n0; a.n2 = n2; a.coeff = coeff;
  a.z_begin = z_begin; a.z_end = z_end; a.fixed_lo = fixed_lo; a.fixed_hi = fixed_hi;
  a.tiles = ceil_div(n2 + E, (int64_t)30 * E);          // lane 1 of tile 0 holds column 0; cover column n2-1
  const int64_t gx = ceil_div(a.tiles, HEAT_TY);
  // resident blocks per SM x ~4 waves of warps; chunks of >= 128 rows (each chunk re-reads 4 rows)
  const int64_t rows = z_end - z_begin;
  const int64_t want = (int64_t)rt().sm_count * MINB * 4;
  int64_t gy = std::max<int64_t>(1, std::min<int64_t>(ceil_div(want, gx), ceil_div(rows, 128)));
  a.z_chunk = ceil_div(rows, gy);
  gy = ceil_div(rows, a.z_chunk);
  if (gy > 65535) return set_error(PH_ERR_INVALID, "heat grid too large for one launch");
  dim3 grid((unsigned)gx, (unsigned)gy), block(32 * HEAT_TY);
  heat2d_tb_kernel<T, E, MINB, NA><<<grid, block, 0, stream>>>(a);
  PH_LAUNCH_CHECK("heat2d_tb_kernel");
  return PH_OK;
}

template <typename T>
static int32_t heat2d_tb_planes(const T* in, T* out, int64_t n0, int64_t n2, T coeff, int64_t z_begin, int64_t z_end,
                                int64_t fixed_lo, int64_t fixed_hi, cudaStream_t stream, bool* used) {
  *used = false;
  if (z_begin >= z_end) { *used = true; return PH_OK; }
  if (!heat2d_tb_shape_ok<T>(n2) || (uintptr_t)in % 32 || (uintptr_t)out % 32) return PH_OK;
  static const int cfg = getenv("PH_HEAT2D_CFG") ? atoi(getenv("PH_HEAT2D_CFG")) : 0;     // tuning knob
  *used = true;
  constexpr int EW = 32 / (int)sizeof(T), EV = 16 / (int)sizeof(T);
  // measured on 16384^2 f32 (Gcell/s): 16-byte groups, 3 blocks/SM, 9 row slots 1389; 6 slots 1352;
  // 2 blocks/SM with 9 / 12 slots 1171 / 1148; 32-byte groups, 2 blocks/SM 1167 (occupancy beats depth)
  switch (cfg) {
    case 1: return heat2d_tb_launch<T, EV, 3, 6>(in, out, n0, n2, coeff, z_begin, z_end, fixed_lo, fixed_hi, stream);
    case 2: return heat2d_tb_launch<T, EW, 2, 6>(in, out, n0, n2, coeff, z_begin, z_end, fixed_lo, fixed_hi, stream);
    default: return heat2d_tb_launch<T, EV, 3, 9>(in, out, n0, n2, coeff, z_begin, z_end, fixed_lo, fixed_hi, stream);
  }
}

// heat_tma.cu: TMA-fed shared-memory pipeline (preferred for rank 3 when the shape allows)
template <typename T>
int32_t heat_tma_planes(const T* in, T* out, int64_t n0, int64_t n1, int64_t n2, T coeff, int64_t z_begin,
                        int64_t z_end, cudaStream_t stream, bool* used);

template <typename T>
int32_t heat_tma2_planes(const T* in, T* out, int64_t n0, int64_t n1, int64_t n2, T coeff, int64_t z_begin,
                         int64_t z_end, int64_t fixed_lo, int64_t fixed_hi, cudaStream_t stream, bool* used,
                         const HeatMirror* mir);

template <typename T>
bool heat_tma2_usable(int64_t n1, int64_t n2);

// one step on planes [z_begin, z_end) of a rank-2/3 grid; boundary planes are NOT touched
template <typename T>
static int32_t heat_planes(int rank, const int64_t* ext, T coeff, const T* in, T* out, int64_t z_begin,
                           int64_t z_end, cudaStream_t stream) {
  if (rank == 3) {
    bool used = false;
    int32_t st = heat_tma_planes<T>(in, out, ext[0], ext[1], ext[2], coeff, z_begin, z_end, stream, &used);
    if (st != PH_OK || used) return st;
    return launch_march<T, 3>(in, out, ext[0], ext[1], ext[2], coeff, z_begin, z_end, stream);
  }
  return launch_march<T, 2>(in, out, ext[0], 1, ext[1], coeff, z_begin, z_end, stream);
}

template <typename T>
static int32_t heat_step_t(int rank, const int64_t* ext, const void* coeff_host, int mode, const void* in_v,
                           void* out_v) {
  Runtime& r = rt();
  T coeff;
  memcpy(&coeff, coeff_host, sizeof(T));
  const T* in = reinterpret_cast<const T*>(in_v);
  T* out = reinterpret_cast<T*>(out_v);
  int64_t total = 1;
  for (int i = 0; i < rank; i++) total *= ext[i];
  if (total == 0) return PH_OK;
  if (rank == 1) {
    const int64_t n = ext[0];
    if (n < 3 && mode == PH_HEAT_FIXED) { PH_CUDA(cudaMemcpyAsync(out, in, n * sizeof(T), cudaMemcpyDeviceToDevice, r.stream)); return PH_OK; }
    if (n < 2) { PH_CUDA(cudaMemcpyAsync(out, in, n * sizeof(T), cudaMemcpyDeviceToDevice, r.stream)); return PH_OK; }
    heat_1d_kernel<T><<<(unsigned)ceil_div(n, 256), 256, 0, r.stream>>>(in, out, n, coeff, mode, 0, n);
    PH_LAUNCH_CHECK("heat_1d_kernel");
    return PH_OK;
  }
  if (mode != PH_HEAT_FIXED) return set_error(PH_ERR_UNSUPPORTED, "PH_HEAT_EXAMPLE1D is rank-1 only");
  bool thin = false;
  for (int i = 0; i < rank; i++) if (ext[i] < 3) thin = true;
  const int64_t plane = total / ext[0];
  if (thin) { PH_CUDA(cudaMemcpyAsync(out, in, total * sizeof(T), cudaMemcpyDeviceToDevice, r.stream)); return PH_OK; }
  // nxt = s.clone on the two boundary planes, stencil on the interior planes
  PH_CUDA(cudaMemcpyAsync(out, in, plane * sizeof(T), cudaMemcpyDeviceToDevice, r.stream));
  PH_CUDA(cudaMemcpyAsync(out + (ext[0] - 1) * plane, in + (ext[0] - 1) * plane, plane * sizeof(T),
                          cudaMemcpyDeviceToDevice, r.stream));
  return heat_planes<T>(rank, ext, coeff, in, out, 1, ext[0] - 1, r.stream);
}

template <typename T>
static int32_t heat_run_t(int rank, const int64_t* ext, const void* coeff_host, int mode, void* a_v, void* b_v,
                          int64_t steps, int32_t* final_is_b) {
  Runtime& r = rt();
  T coeff;
  memcpy(&coeff, coeff_host, sizeof(T));
  T* bufs[2] = {reinterpret_cast<T*>(a_v), reinterpret_cast<T*>(b_v)};
  *final_is_b = (int32_t)(steps > 0 ? (steps & 1) : 0);     // the plain ping-pong convention
  if (steps <= 0) return PH_OK;
  if (rank == 1 && ext[0] <= 1024 && ext[0] >= 3) {
    heat_1d_resident_kernel<T><<<1, 1024, 0, r.stream>>>(bufs[0], bufs[1], (int)ext[0], coeff, mode, steps);
    PH_LAUNCH_CHECK("heat_1d_resident_kernel");
    return PH_OK;
  }
  if (rank >= 2 && mode == PH_HEAT_FIXED) {
    bool thin = false;
    int64_t total = 1;
    for (int i = 0; i < rank; i++) { total *= ext[i]; if (ext[i] < 3) thin = true; }
    if (total == 0) return PH_OK;
    if (!thin) {
      // the boundary planes never change: copy them into the second buffer once, then
      // every step is a single stencil launch over the interior planes
      const int64_t plane = total / ext[0];
      PH_CUDA(cudaMemcpyAsync(bufs[1], bufs[0], plane * sizeof(T), cudaMemcpyDeviceToDevice, r.stream));
      PH_CUDA(cudaMemcpyAsync(bufs[1] + (ext[0] - 1) * plane, bufs[0] + (ext[0] - 1) * plane, plane * sizeof(T),
                              cudaMemcpyDeviceToDevice, r.stream));
      int cur = 0;
      int64_t left = steps;
      while (left > 0) {
        if (left >= 2) {                         // two time steps per pass over HBM when the shape allows
          bool used = false;
          int32_t st = rank == 3
              ? heat_tma2_planes<T>(bufs[cur], bufs[cur ^ 1], ext[0], ext[1], ext[2], coeff, 1, ext[0] - 1, 0,
                                    ext[0] - 1, r.stream, &used, nullptr)
              : heat2d_tb_planes<T>(bufs[cur], bufs[cur ^ 1], ext[0], ext[1], coeff, 1, ext[0] - 1, 0, ext[0] - 1,
                                    r.stream, &used);
          if (st != PH_OK) return st;
          if (used) { cur ^= 1; left -= 2; continue; }
        }
        int32_t st = heat_planes<T>(rank, ext, coeff, bufs[cur], bufs[cur ^ 1], 1, ext[0] - 1, r.stream);
        if (st != PH_OK) return st;
        cur ^= 1; left -= 1;
      }
      *final_is_b = cur;
      return PH_OK;
    }
  }
  for (int64_t t = 0; t < steps; t++) {
    int32_t st = heat_step_t<T>(rank, ext, coeff_host, mode, bufs[t & 1], bufs[(t & 1) ^ 1]);
    if (st != PH_OK) return st;
  }
  return PH_OK;
}

// Slab with `g` ghost planes per side (g = 1: one step per exchange; g = 2: two).  Owned planes are
// [g, n0-g); without a neighbour the first / last owned plane is the fixed global boundary.
// two_step: planes [p_begin, p_end) of `out` receive time t+2 (rank 3, g = 2, shape accepted by
// heat_tma2_usable); otherwise time t+1.
template <typename T>
static int32_t heat_slab_t(int rank, const int64_t* ext, const void* coeff_host, int g, int has_lo, int has_hi,
                           int64_t p_begin, int64_t p_end, const void* in_v, void* out_v, cudaStream_t stream,
                           bool two_step, const HeatMirror* mir, bool* mirrored) {
  if (mirrored) *mirrored = false;
  T coeff;
  memcpy(&coeff, coeff_host, sizeof(T));
  const int64_t n0 = ext[0];
  const int64_t lo = has_lo ? g : g + 1;
  const int64_t hi = has_hi ? n0 - g : n0 - g - 1;
  const int64_t b = std::max<int64_t>(p_begin, lo), e = std::min<int64_t>(p_end, hi);
  const T* in = reinterpret_cast<const T*>(in_v);
  T* out = reinterpret_cast<T*>(out_v);
  for (int i = 1; i < rank; i++)
    if (ext[i] < 3) {
      // a thin grid has no interior: every cell is a held boundary cell, a step is `nxt = s.clone`
      // (heat_step_t does the same for the undivided grid) -- the caller flips buffers after every pass
      int64_t plane = 1;
      for (int j = 1; j < rank; j++) plane *= ext[j];
      if (p_end > p_begin && plane > 0)
        PH_CUDA(cudaMemcpyAsync(out + p_begin * plane, in + p_begin * plane, (size_t)((p_end - p_begin) * plane) * sizeof(T),
                                cudaMemcpyDeviceToDevice, stream));
      return PH_OK;
    }
  if (two_step) {
    if (g < 2) return set_error(PH_ERR_INVALID, "two-step slab update needs 2 ghost planes");
    bool used = false;
    const int64_t flo = has_lo ? -1 : g, fhi = has_hi ? n0 : n0 - g - 1;
    int32_t st = rank == 3 ? heat_tma2_planes<T>(in, out, n0, ext[1], ext[2], coeff, b, e, flo, fhi, stream, &used, mir)
                           : heat2d_tb_planes<T>(in, out, n0, ext[1], coeff, b, e, flo, fhi, stream, &used);
    if (st != PH_OK) return st;
    if (mirrored) *mirrored = used && rank == 3 && mir != nullptr;     // the kernel delivered the halo itself
    if (!used) return set_error(PH_ERR_INVALID, "two-step slab update: buffers must be 32-byte aligned");
    return PH_OK;
  }
  return heat_planes<T>(rank, ext, coeff, in, out, b, e, stream);
}

// exported to comm.cu
int32_t heat_slab_dispatch(int32_t dtype, int rank, const int64_t* ext, const void* coeff_host, int ghost, int has_lo,
                           int has_hi, int64_t p_begin, int64_t p_end, const void* in, void* out,
                           cudaStream_t stream, bool two_step, const HeatMirror* mir, bool* mirrored) {
  if (rank < 2 || rank > 3) return set_error(PH_ERR_UNSUPPORTED, "slab stencil needs rank 2 or 3 (got %d)", rank);
  if (dtype == PH_F32) return heat_slab_t<float>(rank, ext, coeff_host, ghost, has_lo, has_hi, p_begin, p_end, in, out, stream, two_step, mir, mirrored);
  if (dtype == PH_F64) return heat_slab_t<double>(rank, ext, coeff_host, ghost, has_lo, has_hi, p_begin, p_end, in, out, stream, two_step, mir, mirrored);
  return set_error(PH_ERR_UNSUPPORTED, "the heat stencil is defined for F32 / F64");
}

bool heat_two_step_usable(int32_t dtype, int rank, const int64_t* ext) {
  if (rank == 2) return dtype == PH_F32 ? heat2d_tb_shape_ok<float>(ext[1]) : (dtype == PH_F64 && heat2d_tb_shape_ok<double>(ext[1]));
  if (rank != 3) return false;
  if (dtype == PH_F32) return heat_tma2_usable<float>(ext[1], ext[2]);
  if (dtype == PH_F64) return heat_tma2_usable<double>(ext[1], ext[2]);
  return false;
}

}  // namespace ph

using namespace ph;

extern "C" {

int32_t ph_heat_step(int32_t dtype, int32_t rank, const int64_t* extents, const void* coeff_host,
                     int32_t boundary_mode, const void* in, void* out) {
  PH_REQUIRE_INIT();
  if (!extents || !coeff_host || !in || !out) return set_error(PH_ERR_INVALID, "null argument to ph_heat_step");
  if (rank < 1 || rank > 3) return set_error(PH_ERR_UNSUPPORTED, "heat stencil rank must be 1..3 (got %d)", rank);
  if (in == out) return set_error(PH_ERR_INVALID, "ph_heat_step needs distinct in / out buffers");
  if (dtype == PH_F32) return heat_step_t<float>(rank, extents, coeff_host, boundary_mode, in, out);
  if (dtype == PH_F64) return heat_step_t<double>(rank, extents, coeff_host, boundary_mode, in, out);
  return set_error(PH_ERR_UNSUPPORTED, "the heat stencil is defined for F32 / F64");
}

int32_t ph_heat_run(int32_t dtype, int32_t rank, const int64_t* extents, const void* coeff_host,
                    int32_t boundary_mode, void* buf_a, void* buf_b, int64_t steps, int32_t* final_is_b) {
  PH_REQUIRE_INIT();
  if (!extents || !coeff_host || !buf_a || !buf_b) return set_error(PH_ERR_INVALID, "null argument to ph_heat_run");
  if (rank < 1 || rank > 3) return set_error(PH_ERR_UNSUPPORTED, "heat stencil rank must be 1..3 (got %d)", rank);
  int32_t dummy = 0;
  if (!final_is_b) final_is_b = &dummy;
  if (dtype == PH_F32) return heat_run_t<float>(rank, extents, coeff_host, boundary_mode, buf_a, buf_b, steps, final_is_b);
  if (dtype == PH_F64) return heat_run_t<double>(rank, extents, coeff_host, boundary_mode, buf_a, buf_b, steps, final_is_b);
  return set_error(PH_ERR_UNSUPPORTED, "the heat stencil is defined for F32 / F64");
}

int32_t ph_heat_step_slab(int32_t dtype, int32_t rank, const int64_t* extents, const void* coeff_host,
                          int32_t has_lo, int32_t has_hi, int64_t p_begin, int64_t p_end, const void* in,
                          void* out, void* cuda_stream) {
  PH_REQUIRE_INIT();
  if (!extents || !coeff_host || !in || !out) return set_error(PH_ERR_INVALID, "null argument to ph_heat_step_slab");
  cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : rt().stream;
  return heat_slab_dispatch(dtype, rank, extents, coeff_host, 1, has_lo, has_hi, p_begin, p_end, in, out, s, false, nullptr, nullptr);
}

int32_t ph_heat_pass_slab(int32_t dtype, int32_t rank, const int64_t* extents, const void* coeff_host,
                          int32_t ghost_planes, int32_t two_steps, int32_t has_lo, int32_t has_hi, int64_t p_begin,
                          int64_t p_end, const void* in, void* out, void* cuda_stream) {
  PH_REQUIRE_INIT();
  if (!extents || !coeff_host || !in || !out) return set_error(PH_ERR_INVALID, "null argument to ph_heat_pass_slab");
  if (ghost_planes < 1 || ghost_planes > 2) return set_error(PH_ERR_INVALID, "ghost_planes must be 1 or 2");
  if (two_steps && !heat_two_step_usable(dtype, rank, extents))
    return set_error(PH_ERR_UNSUPPORTED, "the two-steps-per-pass kernels cannot take this grid (rank 2 / 3, x extent a multiple of 32 / 16 bytes)");
  cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : rt().stream;
  return heat_slab_dispatch(dtype, rank, extents, coeff_host, ghost_planes, has_lo, has_hi, p_begin, p_end, in, out, s,
                            two_steps != 0, nullptr, nullptr);
}

}  // extern "C"
