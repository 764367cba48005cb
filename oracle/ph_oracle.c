/* ph_oracle.c -- CPU restatement (plain C) of ph-core's hot loops.  TEST
 * INFRASTRUCTURE: the parity checker for large cases and the timed CPU baseline of
 * bench.py.  Never linked into, loaded by, or called from the product library.
 *
 * Two flavours per operation:
 *   ref_*  keeps the reference's STRUCTURE: one thread, a lexicographic coordinate
 *          iterator advanced per element (src/buffered/indexed/lex_iterator.cr:7-21,
 *          src/iterators/stride_iterator.cr:110-122) and a coordinate -> index dot
 *          product per operand (src/buffered/buffered.cr:44-52), one materialised
 *          temporary per operator (src/multi_indexable.cr:1026-1102).
 *   flat_* flat loops + OpenMP on every host core: the most generous CPU baseline.
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile).
 * Parity status: see the header of oracle/ph_oracle.py (same pins, same caveats).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXR 8

/* torchrun exports OMP_NUM_THREADS=1; the CPU baseline wants every host core it can use */
void oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---- Indexed::LexIterator over IndexRegion.cover(shape) --------------------------- */
typedef struct {
  int rank;
  int64_t first[MAXR], last[MAXR], step[MAXR], coord[MAXR], bstep[MAXR];
  int64_t buffer_index;
  int hold;
} lex_iter;

static void axis_strides(const int64_t* shape, int rank, int64_t* out) { /* buffered.cr:15-24 */
  out[rank - 1] = 1;
  for (int i = rank - 2; i >= 0; i--) out[i] = out[i + 1] * shape[i + 1];
}

static void lex_init(lex_iter* it, int rank, const int64_t* first, const int64_t* step,
                     const int64_t* last, const int64_t* shape) {
  it->rank = rank;
  axis_strides(shape, rank, it->bstep);
  it->buffer_index = 0;
  for (int i = 0; i < rank; i++) {
    it->first[i] = first[i]; it->step[i] = step[i]; it->last[i] = last[i];
    it->coord[i] = first[i];
    it->buffer_index += it->bstep[i] * first[i];       /* indexed/stride_iterator.cr:10-13 */
  }
  it->hold = 1;
}

/* StrideIterator#next + Indexed::LexIterator#advance! ; returns 0 on Stop */
static int lex_next(lex_iter* it) {
  if (it->hold) {
    for (int i = 0; i < it->rank; i++) if (it->step[i] == 0) return 0;
    it->hold = 0;
    return 1;
  }
  for (int i = it->rank - 1; i >= 0; i--) {
    if (it->coord[i] == it->last[i]) {
      it->buffer_index -= (it->coord[i] - it->first[i]) * it->bstep[i];
      it->coord[i] = it->first[i];
      if (i == 0) return 0;
    } else {
      it->coord[i] += it->step[i];
      it->buffer_index += it->bstep[i] * it->step[i];
      break;
    }
  }
  return 1;
}

static inline int64_t coord_to_index_fast(const int64_t* coord, const int64_t* strides, int rank) {
  int64_t index = 0;                                   /* buffered.cr:44-52 */
  for (int i = 0; i < rank; i++) index += coord[i] * strides[i];
  return index;
}

static void cover(const int64_t* shape, int rank, int64_t* first, int64_t* step, int64_t* last) {
  for (int i = 0; i < rank; i++) {                     /* index_region.cr:232-238 */
    first[i] = 0; step[i] = shape[i] == 0 ? 0 : 1; last[i] = shape[i] > 0 ? shape[i] - 1 : 0;
  }
}

/* ---- elementwise, reference structure ------------------------------------------------
 * op: 0 + , 1 - , 2 * , 3 /   (single IEEE operations; -ffp-contract=off)             */
#define DEF_REF_MAP_WITH(NAME, T)                                                          \
  void NAME(int op, const T* a, const T* b, T* out, const int64_t* shape, int rank) {       \
    int64_t first[MAXR], step[MAXR], last[MAXR], strides[MAXR];                             \
    cover(shape, rank, first, step, last);                                                  \
    axis_strides(shape, rank, strides);                                                     \
    lex_iter it;                                                                            \
    lex_init(&it, rank, first, step, last, shape);                                          \
    int64_t idx = 0;                                                                        \
    while (lex_next(&it)) {                          /* each_with: first.each_coord :1094 */ \
      const T x = a[coord_to_index_fast(it.coord, strides, rank)];   /* :1097 */            \
      const T y = b[coord_to_index_fast(it.coord, strides, rank)];                          \
      T r;                                                                                  \
      switch (op) { case 0: r = x + y; break; case 1: r = x - y; break;                     \
                    case 2: r = x * y; break; default: r = x / y; break; }                  \
      out[idx++] = r;                                /* buffer[idx] = yield :1047-1049 */   \
    }                                                                                       \
  }
DEF_REF_MAP_WITH(ref_map_with_f32, float)
DEF_REF_MAP_WITH(ref_map_with_f64, double)

/* MultiIndexable#tile (multi_indexable.cr:818-827) through TilingLexIterator
 * (tiling_lex_iterator.cr:27-41) + get(smaller_coord): the broadcast oracle. */
#define DEF_REF_TILE(NAME, T)                                                               \
  void NAME(const T* src, const int64_t* src_shape, const int64_t* counts, int rank, T* out) { \
    int64_t new_shape[MAXR], first[MAXR], step[MAXR], last[MAXR], sstr[MAXR], small[MAXR];  \
    for (int i = 0; i < rank; i++) new_shape[i] = src_shape[i] * counts[i];                 \
    cover(new_shape, rank, first, step, last);                                              \
    axis_strides(src_shape, rank, sstr);                                                    \
    lex_iter it;                                                                            \
    lex_init(&it, rank, first, step, last, new_shape);                                      \
    int64_t idx = 0;                                                                        \
    while (lex_next(&it)) {                                                                 \
      for (int i = 0; i < rank; i++) small[i] = it.coord[i] % src_shape[i];                 \
      out[idx++] = src[coord_to_index_fast(small, sstr, rank)];                             \
    }                                                                                       \
  }
DEF_REF_TILE(ref_tile_f32, float)
DEF_REF_TILE(ref_tile_f64, double)

/* a*b+c with b a [1, C] row vector, exactly as the fluent API would run it:
 * tb = b.tile([R,1]); t = a * tb; out = t + c  (three materialised arrays). */
void ref_mul_rowvec_add_f32(const float* a, const float* b, const float* c, float* out,
                            int64_t rows, int64_t cols) {
  const int64_t shape[2] = {rows, cols}, bshape[2] = {1, cols}, counts[2] = {rows, 1};
  float* tb = (float*)malloc(sizeof(float) * rows * cols);
  float* t = (float*)malloc(sizeof(float) * rows * cols);
  ref_tile_f32(b, bshape, counts, 2, tb);
  ref_map_with_f32(2, a, tb, t, shape, 2);
  ref_map_with_f32(0, t, c, out, shape, 2);
  free(tb); free(t);
}

/* ---- elementwise, flat + OpenMP -------------------------------------------------------- */
void flat_mul_rowvec_add_f32(const float* a, const float* b, const float* c, float* out,
                             float* tmp, int64_t rows, int64_t cols) {
  /* two passes with a materialised temporary, like the reference's two operators */
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < rows; r++)
    for (int64_t j = 0; j < cols; j++) tmp[r * cols + j] = a[r * cols + j] * b[j];
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < rows; r++)
    for (int64_t j = 0; j < cols; j++) out[r * cols + j] = tmp[r * cols + j] + c[r * cols + j];
}

#define DEF_FLAT_BINARY(NAME, T)                                                            \
  void NAME(int op, const T* a, const T* b, T* out, int64_t n) {                            \
    switch (op) {                                                                           \
      case 0: _Pragma("omp parallel for schedule(static)")                                  \
        for (int64_t i = 0; i < n; i++) out[i] = a[i] + b[i]; break;                        \
      case 1: _Pragma("omp parallel for schedule(static)")                                  \
        for (int64_t i = 0; i < n; i++) out[i] = a[i] - b[i]; break;                        \
      case 2: _Pragma("omp parallel for schedule(static)")                                  \
        for (int64_t i = 0; i < n; i++) out[i] = a[i] * b[i]; break;                        \
      default: _Pragma("omp parallel for schedule(static)")                                 \
        for (int64_t i = 0; i < n; i++) out[i] = a[i] / b[i]; break;                        \
    }                                                                                       \
  }
DEF_FLAT_BINARY(flat_binary_f32, float)
DEF_FLAT_BINARY(flat_binary_f64, double)

/* ---- gather / scatter, reference structure ---------------------------------------------
 * NArray#unsafe_fetch_chunk (src/n_array.cr:450-453): a lexicographic iterator over the region
 * yields buffer indices, the chunk is filled in lex order.  unsafe_set_chunk (:484-492) is the
 * mirror image.  first/step/last are per axis, in elements of `shape`. */
static void ref_fetch_chunk_f32(const float* src, const int64_t* shape, int rank, const int64_t* first,
                                const int64_t* step, const int64_t* last, float* out) {
  lex_iter it;
  lex_init(&it, rank, first, step, last, shape);
  int64_t idx = 0;
  while (lex_next(&it)) out[idx++] = src[it.buffer_index];
}

static void ref_set_chunk_f32(float* dst, const int64_t* shape, int rank, const int64_t* first,
                              const int64_t* step, const int64_t* last, const float* src) {
  lex_iter it;
  lex_init(&it, rank, first, step, last, shape);
  int64_t idx = 0;
  while (lex_next(&it)) dst[it.buffer_index] = src[idx++];
}

/* NArray#map with a scalar (src/n_array.cr:589-595 through multi_indexable.cr:947-951) */
static void ref_map_scalar_f32(int op, const float* a, float s, float* out, int64_t n) {
  for (int64_t i = 0; i < n; i++) out[i] = op == 2 ? a[i] * s : (op == 0 ? a[i] + s : a[i] - s);
}

/* One 3-D heat step exactly as the slice arithmetic of SURVEY.md 8(a) a-9 would run through
 * ph-core's operators (examples/heat_equation.cr:38-51 generalised): every slice is a fetched
 * chunk, every operator a materialised array, one thread.
 *   c = s[1...-1, 1...-1, 1...-1]; two_c = c * 2
 *   d_k = (s[lo_k] - two_c) + s[hi_k];  lap = (d_0 + d_1) + d_2
 *   nxt = s.clone; nxt[1...-1, ...] = c + lap * C                                            */
void ref_heat_step_3d_f32(const float* s, float* nxt, const int64_t* shape, float coeff) {
  const int64_t n0 = shape[0], n1 = shape[1], n2 = shape[2];
  memcpy(nxt, s, sizeof(float) * n0 * n1 * n2);                       /* s.clone */
  if (n0 < 3 || n1 < 3 || n2 < 3) return;
  const int64_t ishape[3] = {n0 - 2, n1 - 2, n2 - 2};
  const int64_t m = ishape[0] * ishape[1] * ishape[2];
  float* c = (float*)malloc(sizeof(float) * m);
  float* two_c = (float*)malloc(sizeof(float) * m);
  float* lo = (float*)malloc(sizeof(float) * m);
  float* hi = (float*)malloc(sizeof(float) * m);
  float* t = (float*)malloc(sizeof(float) * m);
  float* d = (float*)malloc(sizeof(float) * m);
  float* lap = (float*)malloc(sizeof(float) * m);
  const int64_t one[3] = {1, 1, 1};
  int64_t first[3] = {1, 1, 1}, last[3] = {n0 - 2, n1 - 2, n2 - 2};
  ref_fetch_chunk_f32(s, shape, 3, first, one, last, c);
  ref_map_scalar_f32(2, c, 2.0f, two_c, m);
  for (int k = 0; k < 3; k++) {
    int64_t f[3] = {1, 1, 1}, l[3] = {n0 - 2, n1 - 2, n2 - 2};
    f[k] = 0; l[k] = shape[k] - 3;
    ref_fetch_chunk_f32(s, shape, 3, f, one, l, lo);                  /* s[lo_k] */
    f[k] = 2; l[k] = shape[k] - 1;
    ref_fetch_chunk_f32(s, shape, 3, f, one, l, hi);                  /* s[hi_k] */
    ref_map_with_f32(1, lo, two_c, t, ishape, 3);                     /* lo - two_c */
    ref_map_with_f32(0, t, hi, d, ishape, 3);                         /* ... + hi   */
    if (k == 0) memcpy(lap, d, sizeof(float) * m);
    else { ref_map_with_f32(0, lap, d, t, ishape, 3); memcpy(lap, t, sizeof(float) * m); }
  }
  ref_map_scalar_f32(2, lap, coeff, t, m);                            /* lap * C */
  ref_map_with_f32(0, c, t, d, ishape, 3);                            /* c + ... */
  ref_set_chunk_f32(nxt, shape, 3, first, one, last, d);              /* nxt[interior] = ... */
  free(c); free(two_c); free(lo); free(hi); free(t); free(d); free(lap);
}

/* the same step as one flat loop nest + OpenMP: the most generous CPU baseline */
void flat_heat_step_3d_f32(const float* s, float* nxt, const int64_t* shape, float coeff) {
  const int64_t n0 = shape[0], n1 = shape[1], n2 = shape[2], p = n1 * n2;
#pragma omp parallel for schedule(static)
  for (int64_t z = 0; z < n0; z++)
    for (int64_t y = 0; y < n1; y++)
      for (int64_t x = 0; x < n2; x++) {
        const int64_t i = z * p + y * n2 + x;
        if (z == 0 || z == n0 - 1 || y == 0 || y == n1 - 1 || x == 0 || x == n2 - 1) { nxt[i] = s[i]; continue; }
        const float c = s[i], two_c = c * 2.0f;
        const float d0 = (s[i - p] - two_c) + s[i + p];
        const float d1 = (s[i - n2] - two_c) + s[i + n2];
        const float d2 = (s[i - 1] - two_c) + s[i + 1];
        nxt[i] = c + ((d0 + d1) + d2) * coeff;
      }
}

/* Enumerable#sum over NArray#each (src/n_array.cr:556-564): a left fold in T, one thread */
float ref_sum_f32(const float* x, int64_t n) {
  float acc = 0.0f;
  for (int64_t i = 0; i < n; i++) acc = acc + x[i];
  return acc;
}

/* per-thread partial sums in double + OpenMP reduction: the most generous CPU baseline */
double flat_sum_f32(const float* x, int64_t n) {
  double acc = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : acc)
  for (int64_t i = 0; i < n; i++) acc += (double)x[i];
  return acc;
}
