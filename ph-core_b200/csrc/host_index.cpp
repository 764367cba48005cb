// host_index.cpp -- host-side index math for the device path (include/ph_host.h).
// Region literals -> canonical regions -> stride/offset descriptors; view transforms are
// folded straight onto descriptors, so no per-element coordinate iterator survives on this
// path (BASELINE.json north_star).  Pure C++, no CUDA.
#include "../../include/ph_host.h"
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <vector>

namespace {

thread_local char g_err[384] = {0};

int32_t fail(int32_t code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

inline int sgn(int64_t x) { return (x > 0) - (x < 0); }

// floor division for possibly negative operands (Crystal Int#//)
inline int64_t floordiv(int64_t a, int64_t b) {
  int64_t q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) q -= 1;
  return q;
}

// negative index -> bound + index, unchecked (CoordUtil.canonicalize_index_unsafe)
inline int64_t wrap_index(int64_t i, int64_t bound) { return i < 0 ? bound + i : i; }

struct Axis { int64_t first, step, last, size; };

// The per-axis rule of RangeSyntax.infer_range (range_syntax.cr:84-136), without the final
// in-bounds check of canonicalize_range.
int32_t infer_axis(const ph_range_lit& lit, int64_t bound, Axis& ax) {
  if (lit.is_index) {
    const int64_t c = wrap_index(lit.first, bound);
    ax = {c, 1, c, 1};
    return PH_HOST_OK;
  }
  int64_t first, tail, step;
  if (!lit.has_step) {
    first = lit.has_first ? wrap_index(lit.first, bound) : 0;
    tail = lit.has_last ? wrap_index(lit.last, bound) : bound - 1;
    step = tail >= first ? 1 : -1;          // direction is inferred AFTER negatives are resolved
  } else {
    step = lit.step;
    first = lit.has_first ? wrap_index(lit.first, bound) : (step > 0 ? 0 : bound - 1);
    tail = lit.has_last ? wrap_index(lit.last, bound) : (step > 0 ? bound - 1 : 0);
  }
  if (lit.has_last && lit.exclusive) {
    if (tail == first) { ax = {0, 0, 0, 0}; return PH_HOST_OK; }   // spans no integer
    tail -= sgn(step);
  }
  if (first < 0 || tail < 0)
    return fail(PH_HOST_INDEX_ERROR, "Invalid index: at least one endpoint is negative after canonicalization");
  if (tail != first && sgn(step) != sgn(tail - first))
    return fail(PH_HOST_INDEX_ERROR, "Could not canonicalize range: conflict between implicit direction and step %lld",
                (long long)step);
  if (step == 0) return fail(PH_HOST_DIV0_ERROR, "Division by 0 (explicit step 0)");
  const int64_t count = (tail >= first ? floordiv(tail - first, step) : floordiv(first - tail, -step)) + 1;
  ax = {first, step, first + step * (count - 1), count};
  return PH_HOST_OK;
}

void finish_region(ph_region& r) {
  // IndexRegion.compute_reduced_shape (index_region.cr:323-335)
  r.reduced_rank = 0;
  if (!r.drop) {
    for (int i = 0; i < r.rank; i++) r.reduced_shape[r.reduced_rank++] = r.proper_shape[i];
    return;
  }
  for (int i = 0; i < r.rank; i++)
    if (!r.degeneracy[i]) r.reduced_shape[r.reduced_rank++] = r.proper_shape[i];
  if (r.reduced_rank == 0) {                 // every axis dropped: 1-D with 0 or 1 element
    int64_t size = r.rank == 0 ? 0 : 1;
    for (int i = 0; i < r.rank; i++) size *= r.proper_shape[i];
    r.reduced_shape[0] = size;
    r.reduced_rank = 1;
  }
}

}  // namespace

namespace {
// extremum records are compared in the value's own type (exact for int64 / uint64 too)
template <typename T>
bool better(const uint8_t* a, const uint8_t* b, bool is_max) {
  T x, y;
  memcpy(&x, a, sizeof(T));
  memcpy(&y, b, sizeof(T));
  return is_max ? x > y : x < y;
}
template <typename T>
bool same(const uint8_t* a, const uint8_t* b) {
  T x, y;
  memcpy(&x, a, sizeof(T));
  memcpy(&y, b, sizeof(T));
  return x == y;
}
}  // namespace

extern "C" {

const char* ph_host_last_error(void) { return g_err; }

int32_t ph_canonicalize_range(const ph_range_lit* lit, int64_t bound, int64_t* first, int64_t* step,
                              int64_t* last, int64_t* size) {
  if (!lit) return fail(PH_HOST_INVALID, "null literal");
  Axis ax = {0, 0, 0, 0};
  int32_t st = infer_axis(*lit, bound, ax);
  if (st != PH_HOST_OK) return st;
  // range_valid? (range_syntax.cr:138-140)
  if (!(ax.last >= 0 && ax.last < bound && ax.first >= 0 && ax.first < bound))
    return fail(PH_HOST_INDEX_ERROR, "Could not canonicalize range: not a sensible index range for axis of length %lld",
                (long long)bound);
  if (first) *first = ax.first;
  if (step) *step = ax.step;
  if (last) *last = ax.last;
  if (size) *size = ax.size;
  return PH_HOST_OK;
}

int32_t ph_canonicalize_coord(const int64_t* coord, int32_t ncoord, const int64_t* shape, int32_t rank,
                              int64_t* out) {
  if (ncoord != rank)
    return fail(PH_HOST_DIMENSION_ERROR, "Could not canonicalize coordinate: shape has %d dimensions, coord has %d", rank, ncoord);
  for (int i = 0; i < rank; i++) {
    if (!(coord[i] < shape[i] && coord[i] >= -shape[i]))
      return fail(PH_HOST_INDEX_ERROR, "Could not canonicalize index: %lld is not a valid index for an axis of length %lld",
                  (long long)coord[i], (long long)shape[i]);
    out[i] = wrap_index(coord[i], shape[i]);
  }
  return PH_HOST_OK;
}

int32_t ph_region_new(const ph_range_lit* lits, int32_t nlits, const int64_t* bound_shape, int32_t rank,
                      int32_t drop, ph_region* out) {
  if (!out || (nlits > 0 && !lits) || (rank > 0 && !bound_shape)) return fail(PH_HOST_INVALID, "null argument");
  if (rank > PH_MAX_RANK) return fail(PH_HOST_INVALID, "rank %d exceeds PH_MAX_RANK", rank);
  if (nlits > rank)
    return fail(PH_HOST_DIMENSION_ERROR, "The region literal had more dimensions (%d) than its bound shape (%d)", nlits, rank);
  memset(out, 0, sizeof(*out));
  out->rank = rank;
  out->drop = drop ? 1 : 0;
  for (int i = 0; i < nlits; i++) {
    int64_t f, s, l, n;
    int32_t st = ph_canonicalize_range(&lits[i], bound_shape[i], &f, &s, &l, &n);
    if (st != PH_HOST_OK) return st;
    out->first[i] = f; out->step[i] = s; out->last[i] = l; out->proper_shape[i] = n;
    out->degeneracy[i] = (lits[i].is_index && drop) ? 1 : 0;
  }
  for (int i = nlits; i < rank; i++) {        // implicit trailing `..` (index_region.cr:215-221)
    out->first[i] = 0; out->step[i] = 1;
    out->last[i] = bound_shape[i] - 1;
    out->proper_shape[i] = bound_shape[i];
  }
  finish_region(*out);
  return PH_HOST_OK;
}

int32_t ph_region_new_trimmed(const ph_range_lit* lits, int32_t nlits, const int64_t* bound_shape,
                              const int64_t* trim_to, int32_t rank, int32_t drop, ph_region* out) {
  if (!out || (nlits > 0 && !lits) || (rank > 0 && !trim_to)) return fail(PH_HOST_INVALID, "null argument");
  if (rank > PH_MAX_RANK || nlits > rank) return fail(PH_HOST_DIMENSION_ERROR, "region literal has more dimensions than the shape");
  const bool allow_relative = bound_shape != nullptr;
  const int64_t* bound = allow_relative ? bound_shape : trim_to;
  memset(out, 0, sizeof(*out));
  out->rank = rank;
  out->drop = drop ? 1 : 0;
  for (int i = 0; i < nlits; i++) {
    const ph_range_lit& l = lits[i];
    if (!allow_relative) {                        // RangeSyntax.ensure_nonnegative (range_syntax.cr:71-82)
      const bool neg = l.is_index ? l.first < 0 : ((l.has_first && l.first < 0) || (l.has_last && l.last < 0));
      if (neg) return fail(PH_HOST_INDEX_ERROR, "Negative indices have no meaning when a bounding shape is not provided.");
    }
    Axis ax = {0, 0, 0, 0};
    int32_t st = infer_axis(l, bound[i], ax);
    if (st != PH_HOST_OK) return st;
    out->first[i] = ax.first; out->step[i] = ax.step; out->last[i] = ax.last; out->proper_shape[i] = ax.size;
    out->degeneracy[i] = (l.is_index && drop) ? 1 : 0;
  }
  for (int i = nlits; i < rank; i++) {
    out->first[i] = 0; out->step[i] = 1;
    out->last[i] = bound[i] - 1;
    out->proper_shape[i] = bound[i];
  }
  finish_region(*out);
  return ph_region_trim(out, trim_to, rank);
}

int32_t ph_region_cover(const int64_t* bound_shape, int32_t rank, int32_t drop, ph_region* out) {
  if (!out || (rank > 0 && !bound_shape) || rank > PH_MAX_RANK) return fail(PH_HOST_INVALID, "bad argument");
  memset(out, 0, sizeof(*out));
  out->rank = rank;
  out->drop = drop ? 1 : 0;
  for (int i = 0; i < rank; i++) {
    out->first[i] = 0;
    out->step[i] = bound_shape[i] == 0 ? 0 : 1;
    out->last[i] = std::max<int64_t>(0, bound_shape[i] - 1);
    out->proper_shape[i] = bound_shape[i];
  }
  finish_region(*out);
  return PH_HOST_OK;
}

int32_t ph_region_fits_in(const ph_region* r, const int64_t* bound_shape, int32_t rank, int32_t* fits) {
  if (!r || !fits) return fail(PH_HOST_INVALID, "null argument");
  if (rank != r->rank)
    return fail(PH_HOST_DIMENSION_ERROR, "The bound shape had a different number of dimensions than this IndexRegion");
  *fits = 1;
  for (int i = 0; i < rank; i++)
    if (bound_shape[i] <= std::max(r->first[i], r->last[i])) { *fits = 0; break; }
  return PH_HOST_OK;
}

int32_t ph_region_trim(ph_region* r, const int64_t* bound_shape, int32_t rank) {
  if (!r) return fail(PH_HOST_INVALID, "null region");
  if (rank != r->rank) return fail(PH_HOST_DIMENSION_ERROR, "trim!: dimension mismatch");
  for (int i = 0; i < rank; i++) {           // trim_axis (index_region.cr:684-703)
    const int64_t nb = bound_shape[i];
    int64_t &f = r->first[i], &s = r->step[i], &l = r->last[i], &n = r->proper_shape[i];
    const int64_t mag = s < 0 ? -s : s;
    if (f >= nb) {
      if (l >= nb) { f = 0; s = 0; l = 0; n = 0; }
      else if (s < 0) {                      // started too high, ends inside
        int64_t span = (nb - 1) - l;
        n = span / mag + 1;
        span -= span % mag;
        f = l + span;
      }
    } else if (s > 0 && l >= nb) {           // starts inside, runs past the bound
      int64_t span = (nb - 1) - f;
      span -= span % mag;
      n = span / mag + 1;
      l = f + span;
    }
  }
  finish_region(*r);
  return PH_HOST_OK;
}

int32_t ph_region_reverse(ph_region* r) {
  if (!r) return fail(PH_HOST_INVALID, "null region");
  for (int i = 0; i < r->rank; i++) {
    std::swap(r->first[i], r->last[i]);
    r->step[i] = -r->step[i];
  }
  return PH_HOST_OK;
}

int32_t ph_region_translate(ph_region* r, const int64_t* offset, int32_t noffset) {
  if (!r || !offset) return fail(PH_HOST_INVALID, "null argument");
  const int n = std::min<int>(noffset, r->rank);
  for (int i = 0; i < n; i++)
    if (offset[i] < 0 && (r->first[i] < -offset[i] || r->last[i] < -offset[i]))
      return fail(PH_HOST_INDEX_ERROR, "Can't translate to negative indices");
  for (int i = 0; i < n; i++) { r->first[i] += offset[i]; r->last[i] += offset[i]; }
  return PH_HOST_OK;
}

int32_t ph_shapes_compatible(const int64_t* a, int32_t na, const int64_t* b, int32_t nb, int32_t* ok) {
  if (!ok) return fail(PH_HOST_INVALID, "null ok");
  *ok = 0;
  if (na == 0 || nb == 0) { *ok = (na == nb); return PH_HOST_OK; }
  const int shared = std::min(na, nb);
  for (int i = 0; i < shared; i++) if (a[i] != b[i]) return PH_HOST_OK;
  const int64_t* longer = na > nb ? a : b;
  for (int i = shared; i < std::max(na, nb); i++) if (longer[i] != 1) return PH_HOST_OK;
  *ok = 1;
  return PH_HOST_OK;
}

int32_t ph_broadcast_shapes(const int64_t* a, const int64_t* b, int32_t rank, int64_t* out) {
  for (int i = 0; i < rank; i++) {
    if (a[i] == b[i] || b[i] == 1) out[i] = a[i];
    else if (a[i] == 1) out[i] = b[i];
    else return fail(PH_HOST_SHAPE_ERROR, "shapes cannot be broadcast on axis %d (%lld vs %lld)", i, (long long)a[i], (long long)b[i]);
  }
  return PH_HOST_OK;
}

int32_t ph_row_chunks(int64_t n, int64_t chunks, int32_t taper, int32_t ramp, int64_t* bounds, int32_t cap, int32_t* count) {
  if (!bounds || !count || cap < 0 || n < 0) return fail(PH_HOST_INVALID, "bad argument to ph_row_chunks");
  std::vector<std::pair<int64_t, int64_t>> b;
  const int64_t k = std::max<int64_t>(1, chunks);
  const int64_t per = (n + k - 1) / k;
  for (int64_t r = 0; r < n; r += per) b.push_back({r, std::min(n, r + per)});
  if (taper > 0 && !b.empty()) {
    int64_t r0 = b.back().first, r1 = b.back().second;
    b.pop_back();
    for (int t = 0; t < taper; t++) {
      const int64_t mid = r0 + (r1 - r0 + 1) / 2;
      if (mid >= r1) break;
      b.push_back({r0, mid});
      r0 = mid;
    }
    b.push_back({r0, r1});
  }
  if (ramp > 0 && !b.empty()) {
    int64_t r0 = b.front().first, r1 = b.front().second;
    b.erase(b.begin());
    std::vector<std::pair<int64_t, int64_t>> head;
    for (int t = 0; t < ramp; t++) {
      const int64_t mid = r1 - (r1 - r0 + 1) / 2;
      if (mid <= r0) break;
      head.insert(head.begin(), {mid, r1});
      r1 = mid;
    }
    head.insert(head.begin(), {r0, r1});
    b.insert(b.begin(), head.begin(), head.end());
  }
  if ((int64_t)b.size() > (int64_t)cap) return fail(PH_HOST_INVALID, "ph_row_chunks: %lld chunks do not fit the caller's %d", (long long)b.size(), cap);
  for (size_t i = 0; i < b.size(); i++) { bounds[2 * i] = b[i].first; bounds[2 * i + 1] = b[i].second; }
  *count = (int32_t)b.size();
  return PH_HOST_OK;
}

int32_t ph_concat_shape(const int64_t* shapes, const int32_t* ranks, int32_t n, int32_t axis, int64_t* out_shape, int32_t* out_axis) {
  if (!shapes || !ranks || !out_shape || !out_axis || n < 1) return fail(PH_HOST_INVALID, "bad argument to ph_concat_shape");
  const int rank = ranks[0];
  if (rank < 0 || rank > PH_MAX_RANK) return fail(PH_HOST_INVALID, "rank out of range");
  for (int idx = 0; idx < rank; idx++)                       // shape.each_with_index { |dim, idx| others.each { ... } }
    for (int k = 0; k < n; k++) {
      if (ranks[k] > PH_MAX_RANK) return fail(PH_HOST_INVALID, "rank out of range");
      if (idx >= ranks[k]) return fail(PH_HOST_INDEX_ERROR, "Index out of bounds");
      if (shapes[(size_t)k * PH_MAX_RANK + idx] != shapes[idx] && idx != axis)
        return fail(PH_HOST_DIMENSION_ERROR, "Cannot concatenate these arrays along axis %d: shapes do not match", axis);
    }
  for (int k = 1; k < n; k++)                                // (an input of HIGHER rank would desynchronise the reference's iterators)
    if (ranks[k] != rank) return fail(PH_HOST_DIMENSION_ERROR, "Cannot concatenate these arrays along axis %d: shapes do not match", axis);
  if (axis >= rank || axis < -rank) return fail(PH_HOST_INDEX_ERROR, "Index out of bounds");
  const int ax = axis < 0 ? axis + rank : axis;
  int64_t total = 0;
  for (int k = 0; k < n; k++) total += shapes[(size_t)k * PH_MAX_RANK + ax];
  for (int i = 0; i < rank; i++) out_shape[i] = shapes[i];
  out_shape[ax] = total;
  *out_axis = ax;
  return PH_HOST_OK;
}

int32_t ph_desc_contiguous(const int64_t* shape, int32_t rank, ph_desc* out) {
  if (!out || rank < 0 || rank > PH_MAX_RANK) return fail(PH_HOST_INVALID, "bad argument");
  memset(out, 0, sizeof(*out));
  out->rank = rank;
  int64_t acc = 1;
  for (int i = rank - 1; i >= 0; i--) {
    out->extent[i] = shape[i];
    out->stride[i] = acc;
    acc *= shape[i];
  }
  return PH_HOST_OK;
}

int32_t ph_desc_region(const ph_desc* src, const ph_region* region, ph_desc* out) {
  if (!src || !region || !out) return fail(PH_HOST_INVALID, "null argument");
  if (region->rank != src->rank)
    return fail(PH_HOST_DIMENSION_ERROR, "region has %d dimensions, the array has %d", region->rank, src->rank);
  ph_desc d;
  memset(&d, 0, sizeof(d));
  d.offset = src->offset;
  bool empty = false;
  for (int i = 0; i < src->rank; i++) {
    if (region->step[i] == 0) { empty = true; continue; }     // first/last are meaningless (index_region.cr:50-53)
    d.offset += region->first[i] * src->stride[i];
  }
  if (empty) d.offset = src->offset;
  bool any_kept = false;
  for (int i = 0; i < src->rank; i++) {
    if (region->drop && region->degeneracy[i]) continue;
    any_kept = true;
    d.extent[d.rank] = region->proper_shape[i];
    d.stride[d.rank] = region->step[i] * src->stride[i];
    d.rank++;
  }
  if (!any_kept) {                            // all axes dropped: [size] (index_region.cr:323-335)
    d.rank = 1;
    d.extent[0] = region->reduced_shape[0];
    d.stride[0] = 1;
  }
  *out = d;
  return PH_HOST_OK;
}

int32_t ph_desc_permute(const ph_desc* src, const int32_t* pattern, int32_t npattern, ph_desc* out) {
  if (!src || !out) return fail(PH_HOST_INVALID, "null argument");
  ph_desc d = *src;
  const int n = pattern ? npattern : src->rank;
  if (pattern)
    for (int i = 0; i < n; i++)
      if (pattern[i] < 0 || pattern[i] >= src->rank)
        return fail(PH_HOST_INDEX_ERROR, "Could not use pattern to permute: axis %d is not present in a %d-dimensional MultiIndexable",
                    pattern[i], src->rank);
  d.rank = n;
  for (int i = 0; i < n; i++) {
    const int from = pattern ? pattern[i] : (src->rank - 1 - i);
    d.extent[i] = src->extent[from];
    d.stride[i] = src->stride[from];
  }
  *out = d;
  return PH_HOST_OK;
}

int32_t ph_desc_reverse(const ph_desc* src, ph_desc* out) {
  if (!src || !out) return fail(PH_HOST_INVALID, "null argument");
  ph_desc d = *src;
  for (int i = 0; i < d.rank; i++) {
    if (d.extent[i] > 0) d.offset += (d.extent[i] - 1) * d.stride[i];
    d.stride[i] = -d.stride[i];
  }
  *out = d;
  return PH_HOST_OK;
}

int32_t ph_desc_reshape(const ph_desc* src, const int64_t* new_shape, int32_t new_rank, ph_desc* out) {
  if (!src || !out || new_rank < 0 || new_rank > PH_MAX_RANK) return fail(PH_HOST_INVALID, "bad argument");
  int64_t old_n = src->rank == 0 ? 0 : 1, new_n = new_rank == 0 ? 0 : 1;
  for (int i = 0; i < src->rank; i++) old_n *= src->extent[i];
  for (int i = 0; i < new_rank; i++) new_n *= new_shape[i];
  if (old_n != new_n)
    return fail(PH_HOST_SHAPE_ERROR, "Cannot change shape (%lld elements) to one with %lld elements because reshape cannot add or remove elements.",
                (long long)old_n, (long long)new_n);
  ph_desc d;
  memset(&d, 0, sizeof(d));
  d.rank = new_rank;
  d.offset = src->offset;
  for (int i = 0; i < new_rank; i++) d.extent[i] = new_shape[i];
  if (old_n == 0) { for (int i = 0; i < new_rank; i++) d.stride[i] = 0; *out = d; return PH_HOST_OK; }
  // Split the source into maximal runs that are contiguous in lex order; every new axis must
  // subdivide one run (numpy's no-copy reshape rule).  Size-1 axes are free.
  int oi = 0, ni = 0;
  int64_t os[PH_MAX_RANK], oe[PH_MAX_RANK];
  int on = 0;
  for (int i = 0; i < src->rank; i++) if (src->extent[i] != 1) { oe[on] = src->extent[i]; os[on] = src->stride[i]; on++; }
  int nn_idx[PH_MAX_RANK], nn = 0;
  for (int i = 0; i < new_rank; i++) { if (new_shape[i] != 1) nn_idx[nn++] = i; else d.stride[i] = 0; }
  while (oi < on && ni < nn) {
    int oj = oi + 1, nj = ni + 1;
    int64_t op = oe[oi], np = new_shape[nn_idx[ni]];
    while (op != np) {
      if (op < np) op *= oe[oj++]; else np *= new_shape[nn_idx[nj++]];
    }
    for (int k = oi; k + 1 < oj; k++)
      if (os[k] != os[k + 1] * oe[k + 1]) return fail(PH_HOST_NEEDS_COPY, "reshape of a non-contiguous view needs a copy");
    int64_t stride = os[oj - 1];
    for (int k = nj - 1; k >= ni; k--) {
      d.stride[nn_idx[k]] = stride;
      stride *= new_shape[nn_idx[k]];
    }
    oi = oj; ni = nj;
  }
  *out = d;
  return PH_HOST_OK;
}

int32_t ph_desc_broadcast(const ph_desc* src, const int64_t* shape, int32_t rank, ph_desc* out) {
  if (!src || !out) return fail(PH_HOST_INVALID, "null argument");
  if (rank != src->rank) return fail(PH_HOST_SHAPE_ERROR, "broadcast requires equal rank (%d vs %d)", src->rank, rank);
  ph_desc d = *src;
  for (int i = 0; i < rank; i++) {
    if (src->extent[i] == shape[i]) continue;
    if (src->extent[i] != 1) return fail(PH_HOST_SHAPE_ERROR, "axis %d of length %lld cannot stretch to %lld", i,
                                         (long long)src->extent[i], (long long)shape[i]);
    d.extent[i] = shape[i];
    d.stride[i] = 0;
  }
  *out = d;
  return PH_HOST_OK;
}

int32_t ph_desc_offset_of(const ph_desc* d, const int64_t* coord, int32_t ncoord, int64_t* out) {
  if (!d || !out) return fail(PH_HOST_INVALID, "null argument");
  if (ncoord != d->rank) return fail(PH_HOST_DIMENSION_ERROR, "coordinate has %d dimensions, the array has %d", ncoord, d->rank);
  int64_t off = d->offset;
  for (int i = 0; i < ncoord; i++) off += coord[i] * d->stride[i];
  *out = off;
  return PH_HOST_OK;
}

// ------------------------------------------------------------------ partitioning plans (one process per GPU)
int32_t ph_shard_range(int64_t n, int32_t world, int32_t rank, int64_t* start, int64_t* stop) {
  if (!start || !stop || world <= 0 || rank < 0 || rank >= world || n < 0) return fail(PH_HOST_INVALID, "bad argument to ph_shard_range");
  const int64_t base = n / world, extra = n % world;
  *start = rank * base + std::min<int64_t>(rank, extra);
  *stop = *start + base + (rank < extra ? 1 : 0);
  return PH_HOST_OK;
}

int32_t ph_slab_layout(int64_t n0, int32_t world, int32_t rank, int32_t ghost, ph_slab* out) {
  if (!out || ghost < 0) return fail(PH_HOST_INVALID, "bad argument to ph_slab_layout");
  memset(out, 0, sizeof(*out));
  int32_t st = ph_shard_range(n0, world, rank, &out->start, &out->stop);
  if (st != PH_HOST_OK) return st;
  out->count = out->stop - out->start;
  out->ghost = ghost;
  out->local_planes = out->count + 2 * (int64_t)ghost;
  out->lo_rank = rank > 0 ? rank - 1 : -1;
  out->hi_rank = rank < world - 1 ? rank + 1 : -1;
  return PH_HOST_OK;
}

int32_t ph_transpose_plan_of(const int64_t* shape, int32_t dims, const int32_t* pattern, int32_t world,
                             int32_t rank, ph_transpose_plan* plan, ph_transpose_peer* peers) {
  if (!shape || !pattern || !plan || dims <= 0 || dims > PH_MAX_RANK || world <= 0 || rank < 0 || rank >= world)
    return fail(PH_HOST_INVALID, "bad argument to ph_transpose_plan_of");
  bool seen[PH_MAX_RANK] = {false};
  for (int i = 0; i < dims; i++) {
    if (pattern[i] < 0 || pattern[i] >= dims || seen[pattern[i]])
      return fail(PH_HOST_INDEX_ERROR, "permute pattern is not a permutation of the axes of a rank-%d array", dims);
    seen[pattern[i]] = true;
  }
  memset(plan, 0, sizeof(*plan));
  plan->dims = dims;
  for (int i = 0; i < dims; i++) plan->new_shape[i] = shape[pattern[i]];
  if (pattern[0] == 0) { plan->local = 1; return PH_HOST_OK; }
  if (!peers) return fail(PH_HOST_INVALID, "ph_transpose_plan_of needs `world` peer entries for an exchanging plan");
  const int k = pattern[0];
  int j = 0;
  for (int i = 0; i < dims; i++) if (pattern[i] == 0) j = i;
  plan->k = k;
  plan->j = j;
  ph_shard_range(shape[0], world, rank, &plan->my_rows[0], &plan->my_rows[1]);
  ph_shard_range(shape[k], world, rank, &plan->my_new_rows[0], &plan->my_new_rows[1]);
  for (int q = 0; q < world; q++) {
    ph_transpose_peer& p = peers[q];
    memset(&p, 0, sizeof(p));
    ph_shard_range(shape[k], world, q, &p.send0, &p.send1);
    ph_shard_range(shape[0], world, q, &p.recv0, &p.recv1);
    for (int i = 0; i < dims; i++) { p.send_shape[i] = plan->new_shape[i]; p.recv_shape[i] = plan->new_shape[i]; }
    p.send_shape[0] = p.send1 - p.send0;
    p.send_shape[j] = plan->my_rows[1] - plan->my_rows[0];
    p.recv_shape[0] = plan->my_new_rows[1] - plan->my_new_rows[0];
    p.recv_shape[j] = p.recv1 - p.recv0;
  }
  return PH_HOST_OK;
}

int32_t ph_slice_plan_of(const int64_t* shape, int32_t dims, const ph_region* reg, int32_t world, int32_t rank,
                         ph_slice_plan* plan, ph_slice_peer* peers) {
  if (!shape || !reg || !plan || dims <= 0 || dims > PH_MAX_RANK || reg->rank != dims || world <= 0 || rank < 0 || rank >= world)
    return fail(PH_HOST_INVALID, "bad argument to ph_slice_plan_of");
  memset(plan, 0, sizeof(*plan));
  plan->dims = reg->reduced_rank;
  for (int i = 0; i < reg->reduced_rank; i++) plan->new_shape[i] = reg->reduced_shape[i];
  bool dropped[PH_MAX_RANK];
  for (int i = 0; i < dims; i++) dropped[i] = reg->drop && reg->degeneracy[i];
  if (!dropped[0] && reg->first[0] == 0 && reg->step[0] == 1 && reg->proper_shape[0] == shape[0]) { plan->local = 1; return PH_HOST_OK; }
  if (!peers) return fail(PH_HOST_INVALID, "ph_slice_plan_of needs `world` peer entries for an exchanging plan");
  int64_t gstride[PH_MAX_RANK];
  int64_t acc = 1;
  for (int i = dims - 1; i >= 0; i--) { gstride[i] = acc; acc *= shape[i]; }
  int kept[PH_MAX_RANK], nkept = 0;
  for (int i = 0; i < dims; i++) if (!dropped[i]) kept[nkept++] = i;
  const int a = nkept ? kept[0] : -1;                 // the result's leading axis (-1: every axis indexed, shape [1])
  const int rnk = reg->reduced_rank;
  const int64_t n_new = plan->new_shape[0];
  int64_t inner_off = 0;
  for (int i = 1; i < dims; i++) inner_off += reg->first[i] * gstride[i];
  const int64_t f0 = reg->first[0], s0 = reg->step[0];
  // rows [lo, hi) of the result's leading axis that `dst` owns and whose data `src` holds
  auto owned = [&](int src, int dst, int64_t* lo, int64_t* hi) {
    int64_t r0, r1, j0, j1;
    ph_shard_range(shape[0], world, src, &r0, &r1);
    ph_shard_range(n_new, world, dst, &j0, &j1);
    *lo = *hi = 0;
    if (r1 <= r0 || j1 <= j0) return;
    if (a != 0) { if (r0 <= f0 && f0 < r1) { *lo = j0; *hi = j1; } return; }
    auto ceil_div64 = [](int64_t x, int64_t y) { return x >= 0 ? (x + y - 1) / y : -((-x) / y); };
    int64_t l, h;
    if (s0 > 0) {
      l = std::max<int64_t>(0, ceil_div64(r0 - f0, s0));
      h = (r1 - 1 >= f0) ? (r1 - 1 - f0) / s0 + 1 : 0;
    } else {
      const int64_t t = -s0;
      l = std::max<int64_t>(0, ceil_div64(f0 - (r1 - 1), t));
      h = (f0 >= r0) ? (f0 - r0) / t + 1 : 0;
    }
    l = std::max(l, j0);
    h = std::min(h, j1);
    if (h > l) { *lo = l; *hi = h; }
  };
  int64_t my0, my1;
  ph_shard_range(shape[0], world, rank, &my0, &my1);
  ph_shard_range(n_new, world, rank, &plan->my_new_rows[0], &plan->my_new_rows[1]);
  for (int q = 0; q < world; q++) {
    ph_slice_peer& p = peers[q];
    memset(&p, 0, sizeof(p));
    p.send.rank = p.land.rank = rnk;
    owned(q, rank, &p.recv0, &p.recv1);
    int64_t lo, hi;
    owned(rank, q, &lo, &hi);
    if (hi <= lo) continue;
    int64_t off = (a == 0 ? f0 + s0 * lo - my0 : f0 - my0) * gstride[0] + inner_off;
    if (a > 0) off += reg->step[a] * lo * gstride[a];
    if (nkept == 0) { p.send.extent[0] = 1; p.send.stride[0] = 1; }
    for (int d = 0; d < nkept; d++) {
      const int i = kept[d];
      p.send.extent[d] = (i == a) ? hi - lo : reg->proper_shape[i];
      p.send.stride[d] = reg->step[i] * gstride[i];
    }
    p.send.offset = off;
    int64_t j0, j1;
    ph_shard_range(n_new, world, q, &j0, &j1);
    int64_t dacc = 1;
    for (int d = rnk - 1; d >= 0; d--) {
      p.land.extent[d] = d == 0 ? hi - lo : plan->new_shape[d];
      p.land.stride[d] = dacc;
      dacc *= d == 0 ? j1 - j0 : plan->new_shape[d];
    }
    p.land.offset = (lo - j0) * p.land.stride[0];
  }
  return PH_HOST_OK;
}

int32_t ph_combine_extremum_records(const uint8_t* records, int32_t world, int32_t dtype, int32_t is_max,
                                    int32_t* winner_rank, int64_t* global_index) {
  if (!records || !winner_rank || !global_index || world <= 0) return fail(PH_HOST_INVALID, "bad argument to ph_combine_extremum_records");
  int best = -1;
  int64_t best_index = -1;
  for (int r = 0; r < world; r++) {
    const uint8_t* rec = records + (size_t)r * PH_EXTREMUM_RECORD_BYTES;
    int64_t local, offset;
    memcpy(&local, rec + 16, 8);
    memcpy(&offset, rec + 24, 8);
    if (local < 0) continue;                       // empty shard
    const int64_t gidx = local + offset;
    bool take = best < 0;
    if (!take) {
      const uint8_t* cur = records + (size_t)best * PH_EXTREMUM_RECORD_BYTES;
      bool gt = false, eq = false;
#define PH_CASE(code, T) case code: gt = better<T>(rec, cur, is_max != 0); eq = same<T>(rec, cur); break;
      switch (dtype) {
        PH_CASE(PH_F32, float) PH_CASE(PH_F64, double) PH_CASE(PH_I32, int32_t) PH_CASE(PH_I64, int64_t)
        PH_CASE(PH_U8, uint8_t) PH_CASE(PH_I8, int8_t) PH_CASE(PH_I16, int16_t) PH_CASE(PH_U16, uint16_t)
        PH_CASE(PH_U32, uint32_t) PH_CASE(PH_U64, uint64_t)
        default: return fail(PH_HOST_INVALID, "unknown dtype %d", dtype);
      }
#undef PH_CASE
      take = gt || (eq && gidx < best_index);
    }
    if (take) { best = r; best_index = gidx; }
  }
  *winner_rank = best;
  *global_index = best_index;
  return PH_HOST_OK;
}

}  // extern "C"
