/* ph_host.h -- host-side index math of the device path: region literals -> canonical
 * regions -> stride/offset descriptors, and view transforms folded onto descriptors.
 *
 * In the real drop-in this logic is Crystal (BASELINE.json north_star: "src/index_region.cr,
 * src/coord_util.cr and src/shape_util.cr ... compile an IndexRegion into a compact
 * stride/offset descriptor the kernels consume").  No Crystal compiler
 * exists in this image, so the same rules are written here in C++ (the reference is
 * compiled code) behind a C-ABI, mirroring the reference's names, argument meaning and
 * error behaviour; INTEGRATION.md shows the Crystal methods they correspond to.
 * Pure host code: no CUDA call, usable without a GPU.
 */
#ifndef PH_HOST_H
#define PH_HOST_H

#include <stdint.h>
#include "ph_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* status codes: which exception class the reference raises */
enum {
  PH_HOST_OK = 0,
  PH_HOST_INDEX_ERROR = 101,      /* IndexError      (range_syntax.cr:120-122,148-150; coord_util.cr:44-46) */
  PH_HOST_DIMENSION_ERROR = 102,  /* DimensionError  (index_region.cr:199-201; coord_util.cr:77-79) */
  PH_HOST_SHAPE_ERROR = 103,      /* ShapeError      (multi_indexable.cr:935-940; multi_writable.cr:58-60) */
  PH_HOST_DIV0_ERROR = 104,       /* DivisionByZeroError (explicit step 0 with first == last) */
  PH_HOST_NEEDS_COPY = 110,       /* reshape of a non-contiguous view: materialise first */
  PH_HOST_INVALID = 111
};

/* One entry of a region literal after RangeSyntax.parse_range (range_syntax.cr:41-69):
 * an Int (`is_index`), or first/step/last with nil-ness flags and exclusivity. */
typedef struct ph_range_lit {
  int32_t is_index;   /* literal was an Int: axis is dropped when `drop` */
  int32_t has_first, has_last, has_step, exclusive;
  int64_t first, last, step;   /* `first` holds the index when is_index */
} ph_range_lit;

/* IndexRegion(T) fields (index_region.cr:54-94) */
typedef struct ph_region {
  int32_t rank;                        /* proper dimensions */
  int32_t drop;
  int64_t first[PH_MAX_RANK], step[PH_MAX_RANK], last[PH_MAX_RANK], proper_shape[PH_MAX_RANK];
  int32_t degeneracy[PH_MAX_RANK];
  int32_t reduced_rank;
  int64_t reduced_shape[PH_MAX_RANK];
} ph_region;

/* RangeSyntax.infer_range / canonicalize_range (range_syntax.cr:84-153) on one axis */
int32_t ph_canonicalize_range(const ph_range_lit* lit, int64_t bound, int64_t* first, int64_t* step,
                              int64_t* last, int64_t* size);
/* CoordUtil.canonicalize_coord (coord_util.cr:76-82) */
int32_t ph_canonicalize_coord(const int64_t* coord, int32_t ncoord, const int64_t* shape, int32_t rank,
                              int64_t* out);
/* IndexRegion.new(region_literal, bound_shape, drop) (index_region.cr:192-224) */
int32_t ph_region_new(const ph_range_lit* lits, int32_t nlits, const int64_t* bound_shape, int32_t rank,
                      int32_t drop, ph_region* out);
/* IndexRegion.new(region_literal, bound_shape, drop, trim_to:) (index_region.cr:133-168):
 * infer (no bounds check) against `bound_shape`, then trim! to `trim_to`.  `bound_shape` may
 * be NULL (absolute literals only: a negative index raises IndexError).  Used by
 * MultiIndexable#get_available (multi_indexable.cr:411-413). */
int32_t ph_region_new_trimmed(const ph_range_lit* lits, int32_t nlits, const int64_t* bound_shape,
                              const int64_t* trim_to, int32_t rank, int32_t drop, ph_region* out);
/* IndexRegion.cover (index_region.cr:232-238) */
int32_t ph_region_cover(const int64_t* bound_shape, int32_t rank, int32_t drop, ph_region* out);
/* IndexRegion#fits_in? (:468-478): *fits = 0/1 */
int32_t ph_region_fits_in(const ph_region* r, const int64_t* bound_shape, int32_t rank, int32_t* fits);
/* IndexRegion#trim! (:502-515), #reverse! (:533-537), #translate! (:577-585) -- in place */
int32_t ph_region_trim(ph_region* r, const int64_t* bound_shape, int32_t rank);
int32_t ph_region_reverse(ph_region* r);
int32_t ph_region_translate(ph_region* r, const int64_t* offset, int32_t noffset);

/* ShapeUtil.compatible_shapes? (shape_util.cr:6-32): *ok = 0/1 */
int32_t ph_shapes_compatible(const int64_t* a, int32_t na, const int64_t* b, int32_t nb, int32_t* ok);
/* NEW ShapeUtil.broadcast_shapes: equal rank, each axis equal or 1 */
int32_t ph_broadcast_shapes(const int64_t* a, const int64_t* b, int32_t rank, int64_t* out);
/* Row schedule of a chunked host -> device -> host pipeline over `n` leading-axis rows (pipeline.RowPipeline /
 * Phase::RowPipeline; no reference counterpart -- NArray lives on the host): `chunks` equal chunks; the LAST one is cut
 * into halves `taper` times (per/2, per/4, ..., per/2^t, per/2^t: the tail nothing overlaps is one small chunk), the
 * FIRST one is the mirror image `ramp` times (per/2^r, per/2^r, ..., per/2: the first download starts early).
 * bounds[2i], bounds[2i+1] = [r0, r1) of chunk i, in order, a partition of [0, n); *count <= cap chunks. */
int32_t ph_row_chunks(int64_t n, int64_t chunks, int32_t taper, int32_t ramp, int64_t* bounds, int32_t cap, int32_t* count);
/* NArray.concatenate's shape rule (n_array.cr:666-673 compatible?, :722-731): `shapes` holds n shapes of `ranks[i]`
 * entries each, packed at a pitch of PH_MAX_RANK.  Every dimension of the FIRST shape must equal the same dimension of
 * every other shape except at index `axis` -- compared on the raw argument, so a negative axis excludes nothing
 * (PH_HOST_DIMENSION_ERROR; a shorter shape, or an axis outside the first shape, is PH_HOST_INDEX_ERROR).
 * out_shape = the first shape with extent[axis] summed over the inputs; *out_axis = the canonical (non-negative) axis. */
int32_t ph_concat_shape(const int64_t* shapes, const int32_t* ranks, int32_t n, int32_t axis, int64_t* out_shape, int32_t* out_axis);

/* Buffered.axis_strides (buffered.cr:15-24) as a descriptor of a whole row-major array */
int32_t ph_desc_contiguous(const int64_t* shape, int32_t rank, ph_desc* out);
/* NEW IndexRegion#to_descriptor: restrict the array / view `src` to `region`:
 * offset += sum first*stride; kept axes get extent = proper_shape, stride = step*stride;
 * all-dropped => rank-1 [size].  Replaces RegionTransform#apply (transforms.cr:215-221). */
int32_t ph_desc_region(const ph_desc* src, const ph_region* region, ph_desc* out);
/* PermuteTransform (transforms.cr:224-271): out axis i = src axis pattern[i];
 * pattern == NULL => reversed axes (:236-238).  Bad axis -> IndexError (view.cr:73-75). */
int32_t ph_desc_permute(const ph_desc* src, const int32_t* pattern, int32_t npattern, ph_desc* out);
/* ReverseTransform (transforms.cr:273-305): every axis flipped */
int32_t ph_desc_reverse(const ph_desc* src, ph_desc* out);
/* ReshapeTransform (transforms.cr:118-189) when expressible as strides; element-count
 * mismatch -> ShapeError (view.cr:59-61); not expressible -> PH_HOST_NEEDS_COPY */
int32_t ph_desc_reshape(const ph_desc* src, const int64_t* new_shape, int32_t new_rank, ph_desc* out);
/* stretch size-1 axes of `src` to `shape` with stride 0 (broadcast operand) */
int32_t ph_desc_broadcast(const ph_desc* src, const int64_t* shape, int32_t rank, ph_desc* out);
/* buffer offset of one coordinate (Buffered.coord_to_index_fast, buffered.cr:44-52) */
int32_t ph_desc_offset_of(const ph_desc* d, const int64_t* coord, int32_t ncoord, int64_t* out);

/* ---- partitioning across one process per GPU.  ph-core is single-process, so
 * these have no reference counterpart; they are the host plans the sharded operations of the path run on
 * (ph_allreduce / ph_allgather / ph_alltoallv / ph_heat_run_sharded in ph_gpu.h do the data movement). */

/* Contiguous split of `n` leading-axis indices over `world` ranks: the first n % world ranks get one extra. */
int32_t ph_shard_range(int64_t n, int32_t world, int32_t rank, int64_t* start, int64_t* stop);

/* Slab of a grid split along axis 0: the owned planes [start, stop) live at local planes
 * [ghost, ghost + count); `ghost` planes on either side are ghosts (1: one time step per halo exchange,
 * 2: two, for the temporally blocked stencil).  lo_rank / hi_rank = neighbour ranks or -1 at the ends of the
 * grid, where the first / last owned plane is the fixed global boundary. */
typedef struct ph_slab {
  int64_t start, stop, count, local_planes;
  int32_t ghost, lo_rank, hi_rank, _pad;
} ph_slab;
int32_t ph_slab_layout(int64_t n0, int32_t world, int32_t rank, int32_t ghost, ph_slab* out);

/* Plan of `permute(pattern)` (PermuteTransform, transforms.cr:224-271) on an array sharded along axis 0
 * whose result is sharded along ITS axis 0 (= old axis k = pattern[0]).  For every peer q (peers[q]):
 *   send0..send1: the slice of old axis k this rank cuts out of its rows for q; the block travels already
 *                 permuted, shape send_shape;
 *   recv0..recv1: the old-axis-0 rows q owns; its block lands at those positions of new axis j (where old
 *                 axis 0 ends up), shape recv_shape.
 * local != 0 (pattern[0] == 0): no exchange, every rank permutes its own shard.  A pattern that is not a
 * permutation of the axes -> IndexError. */
typedef struct ph_transpose_peer {
  int64_t send0, send1, recv0, recv1;
  int64_t send_shape[PH_MAX_RANK], recv_shape[PH_MAX_RANK];
} ph_transpose_peer;
typedef struct ph_transpose_plan {
  int32_t local, dims, k, j;
  int64_t new_shape[PH_MAX_RANK];
  int64_t my_rows[2], my_new_rows[2];
} ph_transpose_plan;
int32_t ph_transpose_plan_of(const int64_t* shape, int32_t dims, const int32_t* pattern, int32_t world,
                             int32_t rank, ph_transpose_plan* plan, ph_transpose_peer* peers /* world entries */);

/* Plan of `narr[region]` (gather, multi_indexable.cr:338-356) on an array sharded along axis 0 whose result is
 * sharded along ITS axis 0.  local != 0: the region leaves axis 0 whole, every rank slices its own shard.  Otherwise
 * the result's leading axis is the first axis the region does not drop -- axis 0 itself (its selected rows are
 * re-split over the ranks) or, when axis 0 is ONE row, a later axis (the row's owner deals it out) -- and for every
 * peer q (peers[q]):
 *   send: the block this rank owes q as a strided view of ITS local shard (an arithmetic progression of its rows);
 *   land: where that block lands in q's shard of the result (a contiguous range of q's rows);
 *         both have the result's rank; extent 0 on every axis = nothing to send;
 *   recv0..recv1: the rows of the result (global numbering) q holds for this rank.
 * send / land are exactly the descriptor lists of ph_alltoall_strided. */
typedef struct ph_slice_peer {
  ph_desc send, land;
  int64_t recv0, recv1;
} ph_slice_peer;
typedef struct ph_slice_plan {
  int32_t local, dims;                 /* dims = rank of the result */
  int64_t new_shape[PH_MAX_RANK];
  int64_t my_new_rows[2];
} ph_slice_plan;
int32_t ph_slice_plan_of(const int64_t* shape, int32_t dims, const ph_region* region, int32_t world, int32_t rank,
                         ph_slice_plan* plan, ph_slice_peer* peers /* world entries */);

/* The sharded argmax / argmin: every rank contributes one 32-byte record to ph_allgather --
 * value @0 (one element of dtype, <= 8 bytes), LOCAL flat index of its first extremum @16 (int64, -1 = empty
 * shard), elements owned by lower ranks @24 (int64).  Picks the best value, then the lowest GLOBAL index
 * (README.md:56-61 across shards).  *winner_rank = -1 when every shard is empty. */
#define PH_EXTREMUM_RECORD_BYTES 32
int32_t ph_combine_extremum_records(const uint8_t* records, int32_t world, int32_t dtype, int32_t is_max,
                                    int32_t* winner_rank, int64_t* global_index);

const char* ph_host_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* PH_HOST_H */
