/* ph_gpu.h -- C-ABI of libphgpu.so: the device-resident NArray backing for
 * ph-core's data-parallel hot path on NVIDIA B200 (sm_100a).
 *
 * ph-core (Crystal) has no FFI today; its seam is the mixin contract of
 * MultiIndexable / MultiWritable (src/multi_indexable.cr:30-65,
 * src/multi_writable.cr:5-44).  A Crystal `lib LibPhGpu` binding (see
 * INTEGRATION.md) declares exactly the functions below, and the device-backed
 * NArray overrides the reference methods each entry cites.
 *
 * Conventions
 *   - every function returns int32 status: 0 = PH_OK; no exception / longjmp
 *     crosses the boundary; ph_last_error_string() describes the last failure.
 *   - plain pointers and sizes only.  `const void* dev` arguments are DEVICE
 *     pointers (from ph_alloc or any CUDA allocator); `host` arguments are host.
 *   - all SEMANTIC validation (ShapeError, DimensionError, IndexError) stays in
 *     the host language and happens before a call; data-dependent errors
 *     (OverflowError, DivisionByZeroError, ArgumentError) are accumulated in a
 *     device flag word read by ph_take_arith_flags().
 *   - launches are asynchronous on the library's per-device stream (ph_stream);
 *     ph_d2h / ph_reduce_full / ph_take_arith_flags / ph_sync synchronise.
 *   - not thread-safe by contract: the reference is single-threaded
 *     (no spawn/Channel/Mutex anywhere in src/).
 *   - there is no CPU fallback: without a CUDA device every compute entry
 *     returns PH_ERR_CUDA.
 */
#ifndef PH_GPU_H
#define PH_GPU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PH_MAX_RANK 8

/* ---- status codes ------------------------------------------------------ */
enum {
  PH_OK = 0,
  PH_ERR_CUDA = 1,        /* a CUDA runtime call failed (incl. "no device") */
  PH_ERR_INVALID = 2,     /* bad argument: null pointer, rank > 8, unknown enum */
  PH_ERR_UNSUPPORTED = 3, /* op not defined for dtype (e.g. `&` on floats) */
  PH_ERR_NCCL = 4,
  PH_ERR_NOT_INIT = 5
};

/* ---- element types (Crystal primitive numerics + Bool) ------------------ */
enum {
  PH_F32 = 0, PH_F64 = 1, PH_I32 = 2, PH_I64 = 3,
  PH_U8 = 4,  /* UInt8 and Bool (Slice(Bool) is one byte per element) */
  PH_I8 = 5, PH_I16 = 6, PH_U16 = 7, PH_U32 = 8, PH_U64 = 9
};

/* ---- binary operators: the op list of src/multi_indexable.cr:960-975 ---- */
enum {
  PH_ADD = 0,      /* +   ints: overflow-checked                           */
  PH_SUB = 1,      /* -                                                     */
  PH_MUL = 2,      /* *                                                     */
  PH_DIV = 3,      /* /   ints: result dtype is F64 (Int / Int -> Float64)  */
  PH_FLOORDIV = 4, /* //  ints: floored, /0 -> DIV0, MIN//-1 -> ARGUMENT;
                          floats: (a / b).floor                             */
  PH_MOD = 5,      /* %   ints: floored modulo; floats: a - b*(a/b).floor,
                          b == 0 -> DIV0                                    */
  PH_POW = 6,      /* **  ints: checked square-and-multiply, exp<0 -> ARGUMENT;
                          floats array**array: pow() (tolerance only)       */
  PH_WADD = 7, PH_WSUB = 8, PH_WMUL = 9, PH_WPOW = 10, /* &+ &- &* &** wrap */
  PH_AND = 11, PH_OR = 12, PH_XOR = 13,                 /* & | ^ (ints)     */
  PH_POWI = 14     /* Float ** Int32 scalar: llvm.powi (compiler-rt loop);
                      ph_ewise_scalar only, scalar holds an int32           */
};

/* comparisons: src/multi_indexable.cr:977-980, eq :899-913 */
enum { PH_GT = 0, PH_LT = 1, PH_GE = 2, PH_LE = 3, PH_EQ = 4, PH_NE = 5 };

/* unary: src/multi_indexable.cr:983-985 */
enum { PH_POS = 0, PH_NEG = 1, PH_NOT = 2 };

/* reductions: Enumerable#sum/min/max over NArray#each (src/n_array.cr:556-564),
 * argmax idiom README.md:56-61 */
enum { PH_SUM = 0, PH_MIN = 1, PH_MAX = 2, PH_ARGMAX = 3, PH_ARGMIN = 4 };

/* arithmetic flag bits (ph_take_arith_flags) -> exception the host raises */
enum {
  PH_FLAG_OVERFLOW = 1, /* OverflowError                                   */
  PH_FLAG_DIV0 = 2,     /* DivisionByZeroError                             */
  PH_FLAG_NAN = 4,      /* ArgumentError: NaN met by max/min/argmax        */
  PH_FLAG_ARGUMENT = 8  /* ArgumentError: MIN // -1, negative int exponent */
};

/* heat boundary modes */
enum {
  PH_HEAT_FIXED = 0,   /* N-D rule of SURVEY.md 8(a) a-9: boundary cells held */
  PH_HEAT_EXAMPLE1D = 1 /* examples/heat_equation.cr:43-44 one-sided ends     */
};

/* ---- strided descriptor --------------------------------------------------
 * What an IndexRegion (src/index_region.cr:54-94) or a View transform chain
 * (src/view_util/transforms.cr) compiles to: element (not byte) units, signed
 * strides, stride 0 = broadcast axis.  Element i = (i_0..i_{rank-1}) lives at
 * base[offset + sum_k i_k * stride[k]].  Replaces per-element
 * IndexRegion#local_to_absolute_unsafe (:621-637) + Buffered.coord_to_index_fast
 * (src/buffered/buffered.cr:44-52). */
typedef struct ph_desc {
  int32_t rank;
  int32_t _pad;
  int64_t offset;
  int64_t extent[PH_MAX_RANK];
  int64_t stride[PH_MAX_RANK];
} ph_desc;

/* ---- runtime / storage: NArray buffer ownership, src/n_array.cr:20-79,
 *      clone :372-374, fill :230-232 ----------------------------------------- */
int32_t ph_init(int32_t device);              /* select device, create stream + pool */
int32_t ph_shutdown(void);
int32_t ph_device_count(int32_t* out);
int32_t ph_sm_count(int32_t* out);
int32_t ph_alloc(size_t nbytes, void** out_dev);     /* stream-ordered pool allocation */
int32_t ph_free(void* dev);
int32_t ph_h2d(void* dst_dev, const void* src_host, size_t nbytes);
int32_t ph_d2h(void* dst_host, const void* src_dev, size_t nbytes);   /* synchronises */
/* ph_d2h + ph_take_arith_flags in one synchronisation: what `to_host` / `get` call, so an
 * OverflowError / DivisionByZeroError of any earlier operator is raised by the read that would
 * hand its result to the caller (the reference raises at the operator itself). */
int32_t ph_d2h_flags(void* dst_host, const void* src_dev, size_t nbytes, uint32_t* out_flags);
int32_t ph_d2h_async(void* dst_host, const void* src_dev, size_t nbytes); /* pinned dst; ph_sync before reading */
int32_t ph_d2d(void* dst_dev, const void* src_dev, size_t nbytes);    /* NArray#clone */
int32_t ph_host_alloc(size_t nbytes, void** out_host);  /* pinned staging memory */
int32_t ph_host_free(void* host);
int32_t ph_sync(void);
void*   ph_stream(void);                      /* the cudaStream_t launches go to */
int32_t ph_set_stream(void* cuda_stream);     /* adopt a caller-owned stream (0 = own) */
/* Caller-visible streams, for chunked host <-> device pipelines written with the array API (upload of
 * chunk i+1, the operators of chunk i and the download of chunk i-1 overlap: PCIe is full duplex).
 * ph_set_stream(s) makes `s` the stream every later call launches on; ph_stream_wait orders two streams;
 * ph_free_on releases a block on the stream it was used on.  ph_h2d / ph_d2h_async are asynchronous when
 * the host side is pinned (ph_host_alloc). */
int32_t ph_stream_create(void** out_stream);
int32_t ph_stream_destroy(void* stream);
int32_t ph_stream_wait(void* waiter_stream, void* signaler_stream);   /* NULL = the library's own stream */
int32_t ph_stream_sync(void* stream);
int32_t ph_free_on(void* dev, void* stream);
/* Position-weighted 64-bit checksum of a device buffer: sum of word_i * (2 * (i + word_offset) + 1) mod 2^64
 * over its 8-byte words.  Independent of grid and sharding (ranks pass the GLOBAL index of their first word
 * and add their values), sensitive to where every word sits.  Verification aid (bench.py field_hash). */
int32_t ph_checksum64(const void* dev, size_t nbytes, uint64_t word_offset, uint64_t* out_host);
const char* ph_last_error_string(void);
int32_t ph_take_arith_flags(uint32_t* out_flags);    /* read + clear; synchronises */
int32_t ph_timer_start(void);                 /* CUDA event on ph_stream */
int32_t ph_timer_stop(float* out_ms);         /* records, synchronises, elapsed ms */
int64_t ph_launch_count(void);                /* kernels launched since ph_init */

/* ---- elementwise (K1/K2): def_elementwise_binary, src/multi_indexable.cr:931-952
 * -> map_with/each_with :1026-1102; all descriptors share rank and extents,
 * broadcast operands carry stride 0.  Output dtype = dtype, except PH_DIV on an
 * integer dtype which writes F64. */
int32_t ph_ewise_binary(int32_t op, int32_t dtype,
                        const void* a, const ph_desc* a_desc,
                        const void* b, const ph_desc* b_desc,
                        void* out, const ph_desc* out_desc);
/* array (op) scalar: src/multi_indexable.cr:947-951 -> NArray#map src/n_array.cr:589-595;
 * scalar (op) array (scalar_on_left != 0): src/patches/number.cr:6-15.
 * `scalar_host` points at one element of `dtype` (an int32 for PH_POWI). */
int32_t ph_ewise_scalar(int32_t op, int32_t dtype,
                        const void* a, const ph_desc* a_desc,
                        const void* scalar_host, int32_t scalar_on_left,
                        void* out, const ph_desc* out_desc);
/* unary + - ~ : src/multi_indexable.cr:954-958 */
int32_t ph_ewise_unary(int32_t op, int32_t dtype,
                       const void* a, const ph_desc* a_desc,
                       void* out, const ph_desc* out_desc);
/* fused out = (a * b) + c with TWO roundings (never an FMA): SURVEY.md 8(f) f-1;
 * bit-identical to ph_ewise_binary(MUL) then ph_ewise_binary(ADD). */
int32_t ph_ewise_mul_add(int32_t dtype,
                         const void* a, const ph_desc* a_desc,
                         const void* b, const ph_desc* b_desc,
                         const void* c, const ph_desc* c_desc,
                         void* out, const ph_desc* out_desc);

/* ---- comparisons -> NArray(Bool) (K3): src/multi_indexable.cr:977-980, eq :899-920 */
int32_t ph_compare(int32_t cmp, int32_t dtype,
                   const void* a, const ph_desc* a_desc,
                   const void* b, const ph_desc* b_desc,
                   uint8_t* out, const ph_desc* out_desc);
int32_t ph_compare_scalar(int32_t cmp, int32_t dtype,
                          const void* a, const ph_desc* a_desc,
                          const void* scalar_host, int32_t scalar_on_left,
                          uint8_t* out, const ph_desc* out_desc);

/* `<=>` of the operator list (src/multi_indexable.cr:960-981): -1 / 0 / 1 as Int32, INTEGER element
 * types only (Float#<=> is Int32?, nil against NaN: PH_ERR_UNSUPPORTED). */
int32_t ph_compare3(int32_t dtype,
                    const void* a, const ph_desc* a_desc,
                    const void* b, const ph_desc* b_desc,
                    int32_t* out, const ph_desc* out_desc);
int32_t ph_compare3_scalar(int32_t dtype,
                           const void* a, const ph_desc* a_desc,
                           const void* scalar_host, int32_t scalar_on_left,
                           int32_t* out, const ph_desc* out_desc);

/* ---- masked store (K4): NArray#[]=(mask, value) src/n_array.cr:510-551,
 *      MultiWritable#set_mask src/multi_writable.cr:154-172 ------------------- */
int32_t ph_mask_set_scalar(int32_t elem_size, void* dst, const ph_desc* dst_desc,
                           const uint8_t* mask, const ph_desc* mask_desc,
                           const void* scalar_host);
int32_t ph_mask_set_array(int32_t elem_size, void* dst, const ph_desc* dst_desc,
                          const uint8_t* mask, const ph_desc* mask_desc,
                          const void* src, const ph_desc* src_desc);

/* ---- strided gather / scatter / fill (K5/K6/K10) ----------------------------
 * ph_copy_strided: dst[dst_desc(i)] = src[src_desc(i)] for every i of the shared
 * extents, in any order (the regions never alias).  Covers
 *   gather  NArray#unsafe_fetch_chunk src/n_array.cr:450-453, View#to_narr
 *           src/view.cr:123-126, MultiIndexable#permute/reverse/reshape
 *           src/multi_indexable.cr:795-803   (dst_desc contiguous)
 *   scatter NArray#unsafe_set_chunk(region, src) src/n_array.cr:484-492,
 *           MutableView writes src/mutable_view.cr:16-18 (dst_desc strided) */
int32_t ph_copy_strided(int32_t elem_size,
                        const void* src, const ph_desc* src_desc,
                        void* dst, const ph_desc* dst_desc);
/* NArray#unsafe_set_chunk(region, value) src/n_array.cr:495-500; NArray.fill :230-232 */
int32_t ph_fill_region(int32_t elem_size, void* dst, const ph_desc* dst_desc,
                       const void* scalar_host);

/* ---- reductions (K7/K8) --------------------------------------------------------
 * Full: Enumerable#sum/min/max over NArray#each (src/n_array.cr:556-564); result
 * written to HOST memory (`out_value_host`: one element of dtype; for ARGMAX /
 * ARGMIN also the flat lex index of the FIRST extremum in `out_index_host`).
 * `count_out` (optional) receives the element count (0 -> host raises EmptyError
 * for min/max).  NaN under min/max/arg* sets PH_FLAG_NAN; integer SUM that
 * leaves T's range at ANY prefix of the lex-order fold sets PH_FLAG_OVERFLOW. */
int32_t ph_reduce_full(int32_t red, int32_t dtype,
                       const void* a, const ph_desc* a_desc,
                       void* out_value_host, int64_t* out_index_host);
/* Same, result left on the device (for NCCL allreduce of per-GPU partials):
 * out_value_dev: one element of dtype, out_index_dev: one int64 (may be null). */
int32_t ph_reduce_full_dev(int32_t red, int32_t dtype,
                           const void* a, const ph_desc* a_desc,
                           void* out_value_dev, int64_t* out_index_dev);
/* Per-axis (defined by each_slice(axis) + fold, src/multi_indexable.cr:742-748):
 * out has a_desc's extents without `axis`; dtype for SUM/MIN/MAX, int64 for ARG*. */
int32_t ph_reduce_axis(int32_t red, int32_t dtype,
                       const void* a, const ph_desc* a_desc, int32_t axis,
                       void* out, const ph_desc* out_desc);

/* ---- heat stencil (K9): examples/heat_equation.cr:38-51 -------------------------
 * One explicit step on a contiguous row-major grid of `rank` (1..3) axes;
 * `coeff_host` points at one element of dtype (F32/F64).  in != out. */
int32_t ph_heat_step(int32_t dtype, int32_t rank, const int64_t* extents,
                     const void* coeff_host, int32_t boundary_mode,
                     const void* in, void* out);
/* `steps` steps between the two buffers.  *final_is_b (may be NULL) receives 1 when the final
 * state is in buf_b, 0 when it is in buf_a: rank-3 grids advance TWO time steps per pass over
 * HBM when the shape allows (temporal blocking, bit-identical to single steps), so the parity
 * of `steps` does not tell.  With a NULL pointer the caller must not rely on either buffer. */
int32_t ph_heat_run(int32_t dtype, int32_t rank, const int64_t* extents, const void* coeff_host,
                    int32_t boundary_mode, void* buf_a, void* buf_b, int64_t steps, int32_t* final_is_b);
/* Slab-decomposed step (axis 0 sharded): the local slab holds planes
 * [1, n0_local] plus ghost planes 0 and n0_local+1 (extents[0] = n0_local + 2);
 * has_lo / has_hi say whether a neighbour exists on that side (else the edge
 * plane is a fixed global boundary).  Computes planes [p_begin, p_end) only, so
 * the caller can do the two edge planes first, start the halo exchange, then the
 * interior (examples/heat_equation.cr generalised). */
int32_t ph_heat_step_slab(int32_t dtype, int32_t rank, const int64_t* extents,
                          const void* coeff_host, int32_t has_lo, int32_t has_hi,
                          int64_t p_begin, int64_t p_end,
                          const void* in, void* out, void* cuda_stream);

/* The same with `ghost_planes` (1 or 2) ghost planes per side (extents[0] = owned + 2 * ghost_planes;
 * owned planes are [ghost_planes, extents[0] - ghost_planes)).  two_steps != 0 (needs 2 ghost planes,
 * rank 3): planes [p_begin, p_end) of `out` receive time t+2 from time t in `in` in ONE pass over HBM
 * (temporal blocking; bit-identical to two ph_heat_step_slab passes). */
int32_t ph_heat_pass_slab(int32_t dtype, int32_t rank, const int64_t* extents,
                          const void* coeff_host, int32_t ghost_planes, int32_t two_steps,
                          int32_t has_lo, int32_t has_hi, int64_t p_begin, int64_t p_end,
                          const void* in, void* out, void* cuda_stream);

/* ---- multi-GPU (one process per GPU, NCCL over NVLink) ---------------------------- */
int32_t ph_comm_unique_id(uint8_t* out128);                 /* rank 0; broadcast by the host */
int32_t ph_comm_init(int32_t nranks, int32_t rank, const uint8_t* id128);
int32_t ph_comm_destroy(void);
/* 1 when every rank has mapped its peers' memory (CUDA IPC over NVLink / NVSwitch): sharded reductions then
 * combine inside the reduction kernel and the stencil delivers its own halos (no NCCL on either path).
 * 0 (IPC refused, PH_NO_P2P=1, a single rank): the NCCL forms of the same entry points run instead. */
int32_t ph_comm_p2p_ready(int32_t* out);
/* Data-path NCCL calls (all-reduce, all-gather, send, recv) this process has issued so far: a measurement aid --
 * the peer-memory forms leave it unchanged over a timed region (bench.py prints the difference per leg). */
int64_t ph_nccl_call_count(void);
/* Peer-mapped device memory for arrays whose kernels write into a neighbour rank (heat slabs).
 * COLLECTIVE: every rank calls ph_symm_alloc / ph_symm_free in the same order.  Works (as a plain
 * allocation) without P2P too.  ph_symm_peer: the address of `local_dev` (any address inside a
 * ph_symm_alloc block) in rank `peer_rank`'s copy of that block, NULL when not mapped. */
int32_t ph_symm_alloc(size_t nbytes, void** out_dev);
int32_t ph_symm_free(void* dev);
int32_t ph_symm_peer(const void* local_dev, int32_t peer_rank, void** out_peer_dev);
/* Full reduction of an array sharded along axis 0 (Enumerable#sum/min/max and the argmax idiom over the
 * whole distributed array; src/n_array.cr:556-564, README.md:56-61).  COLLECTIVE: every rank passes its
 * shard (possibly EMPTY: it contributes the identity) and `elems_before` = the number of elements owned by
 * lower ranks.  Every rank receives the same result: value in `out_value_host`, for ARG* the GLOBAL flat
 * index of the first extremum in `out_index_host` (-1: every shard is empty -> the host raises EmptyError
 * for min/max).  Integer SUM is overflow-checked over the lexicographic fold of the GLOBAL array.
 * `out_flags` receives the arithmetic flags of every rank, read and cleared (PH_FLAG_*).
 * One kernel launch per rank and one synchronisation: the last block stores the rank's partial into every
 * peer's slot over NVLink and folds the N slots in rank order (deterministic); NCCL allgather of the
 * records + a second tiny launch when peers are not mapped (or PH_REDUCE_NCCL=1).
 * WITHOUT a communicator (ph_comm_init never called) this is the single-process full reduction in "record
 * mode": one launch, the finishing block writes value / index / flags into a pinned host record the call
 * polls -- no device-to-host copy, no stream synchronisation, no second read for the flags.  The host layers
 * use it for every `sum` / `min` / `max` / argmax (ph_reduce_full + ph_d2h_flags is the two-read form). */
int32_t ph_reduce_full_sharded(int32_t red, int32_t dtype, const void* a, const ph_desc* a_desc,
                               int64_t elems_before, void* out_value_host, int64_t* out_index_host,
                               uint32_t* out_flags);
/* SUM / MIN / MAX in place over the ranks.  With peer-mapped memory (ph_comm_p2p_ready): reduce-scatter + all-gather
 * as three small launches of peer stores / loads over NVLink, folded IN RANK ORDER -- deterministic, and for an axis-0
 * sharded array rank order is row order, so integer SUMs are overflow-checked (PH_FLAG_OVERFLOW) like the single-GPU
 * fold; every rank's arithmetic flags reach every rank.  Otherwise (or PH_ALLREDUCE_NCCL=1) ncclAllReduce: unordered,
 * integer sums wrap -- callers that need the check gather and fold (ph_allgather + ph_reduce_axis). */
int32_t ph_allreduce(int32_t red, int32_t dtype, void* buf_dev, int64_t count);
int32_t ph_allgather(const void* send_dev, void* recv_dev, int64_t nbytes_per_rank);
/* personalised all-to-all (the exchange step of a transpose across axis-0 shards): arrays of
 * nranks entries; block p of the send list goes to rank p, block p of the receive list comes from
 * rank p; byte counts may be 0.  No reference counterpart (ph-core is single-process): SURVEY.md 8(f) f-3. */
int32_t ph_alltoallv(const void* const* send_dev, const int64_t* send_bytes,
                     void* const* recv_dev, const int64_t* recv_bytes);
/* The same exchange as ONE pass of peer stores (needs ph_comm_p2p_ready): block q of `src_dev` -- src_descs[q],
 * a strided view already in the DESTINATION's axis order -- is copied by the gather / transpose kernels straight
 * into rank q's copy of the symmetric block `dst_symm` (from ph_symm_alloc, same call order on every rank) at
 * dst_descs[q] (element offsets relative to dst_symm).  The permuting copy is the transfer: no staging buffers,
 * no ncclSend/ncclRecv, no scatter.  COLLECTIVE; stream-ordered (returns without blocking the host): two flag
 * rounds over the peer-mapped control blocks bracket the copies.  Blocks with no elements are skipped.
 * PH_ERR_UNSUPPORTED when peers are not mapped -- the caller falls back to ph_alltoallv. */
int32_t ph_alltoall_strided(int32_t elem_size, const void* src_dev, const ph_desc* src_descs,
                            void* dst_symm, const ph_desc* dst_descs);
/* exchange one plane with each neighbour rank (lo = rank-1, hi = rank+1; -1 = none):
 * sends send_lo -> lo, send_hi -> hi; receives recv_lo <- lo, recv_hi <- hi. */
int32_t ph_halo_exchange(const void* send_lo, void* recv_lo, int32_t lo_rank,
                         const void* send_hi, void* recv_hi, int32_t hi_rank,
                         int64_t nbytes, void* cuda_stream);
/* `steps` slab steps with the exchange overlapped with the interior update.  The local slab
 * holds `ghost_planes` (1 or 2) ghost planes on either side of its owned planes
 * (local_extents[0] = owned + 2 * ghost_planes).  With 2 ghost planes a rank-3 grid advances
 * two time steps per pass over HBM and per exchange (bit-identical to single steps).
 * *final_is_b (may be NULL) = 1 when the final state is in buf_b.
 * When both slabs come from ph_symm_alloc and peers are mapped, a pass is ONE launch: the stencil kernel
 * stores the planes its neighbours need straight into their ghost planes and releases their flag words
 * (compute + halo in one kernel, no ncclSend/ncclRecv; PH_HEAT_NCCL=1 forces the NCCL form for A/B). */
int32_t ph_heat_run_sharded(int32_t dtype, int32_t rank, const int64_t* local_extents,
                            const void* coeff_host, int32_t ghost_planes, void* buf_a, void* buf_b,
                            int64_t steps, int32_t* final_is_b);

#ifdef __cplusplus
}
#endif
#endif /* PH_GPU_H */
