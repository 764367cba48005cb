// ph_narray.hpp -- header-only C++17 host layer above the C-ABI (ph_gpu.h + ph_host.h):
// Phase::DeviceNArray<T> / Phase::DeviceView<T>, the device-backed twins of ph-core's
// NArray(T) / View / MutableView with the reference's method names, argument meaning and
// exception classes, so a host program (or a spec) reads like the Crystal original.
//
// Why C++: the reference is compiled code (Crystal) and no Crystal compiler exists in this
// image; INTEGRATION.md gives the Crystal twin of every method below.  This header holds NO
// array arithmetic and NO index math: every data operation is one libphgpu launch, every
// region / transform rule is a ph_host.h call.  There is no CPU fallback -- without a CUDA
// device Device::init throws DeviceError and nothing else can be constructed.
//
// Reference map (file:line under the reference's src/):
//   class DeviceNArray<T>        n_array.cr:15-79 (storage), :230-232 fill, :372-395 clone/dup,
//                                :424-437 reshape/flatten (ALIAS the buffer), :440-447 ==
//   get_chunk / operator[]       multi_indexable.cr:338-356, 523-531 -> n_array.cr:450-453
//   get_available / has_region   multi_indexable.cr:313-318, 397-413
//   get / set_element            multi_indexable.cr:567-575, multi_writable.cr:48-50
//   set_chunk                    multi_writable.cr:55-84 -> n_array.cr:484-500
//   set_mask                     n_array.cr:510-551, multi_writable.cr:154-172
//   + - * / % floordiv pow ...   def_elementwise_binary multi_indexable.cr:931-985,
//                                scalar on the left: patches/number.cr:6-15
//   > < >= <= eq match           multi_indexable.cr:899-920, 977-980
//   sum min max argmax           Enumerable over NArray#each n_array.cr:556-564, README.md:56-61
//   each_slice / slices / tile   multi_indexable.cr:742-786, 818-843
//   view / DeviceView            view.cr:10-126, mutable_view.cr:16-18, view_util/transforms.cr
//   map / apply / process ...    arbitrary blocks: throw DeviceBlockError (BASELINE north_star)
//   Heat::step / Heat::run       examples/heat_equation.cr:26-51
#ifndef PH_NARRAY_HPP
#define PH_NARRAY_HPP

#include <cstdint>
#include <initializer_list>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <cstring>
#include <vector>

#include "ph_gpu.h"
#include "ph_host.h"

namespace Phase {

// ---- the reference's exception classes (exceptions/exceptions.cr:4-22 + Crystal stdlib) ----
struct ShapeError : std::runtime_error { using std::runtime_error::runtime_error; };
struct DimensionError : ShapeError { using ShapeError::ShapeError; };
struct IndexError : std::runtime_error { using std::runtime_error::runtime_error; };
struct OverflowError : std::runtime_error { using std::runtime_error::runtime_error; };
struct DivisionByZeroError : std::runtime_error { using std::runtime_error::runtime_error; };
struct ArgumentError : std::runtime_error { using std::runtime_error::runtime_error; };
struct EmptyError : std::runtime_error { using std::runtime_error::runtime_error; };   // Enumerable::EmptyError
// arbitrary Crystal blocks cannot run on the device path: raise, never run on the CPU
struct DeviceBlockError : std::runtime_error { using std::runtime_error::runtime_error; };
// CUDA / NCCL failure or a missing device (RuntimeError in INTEGRATION.md)
struct DeviceError : std::runtime_error { using std::runtime_error::runtime_error; };

using Bool = uint8_t;                 // Slice(Bool): one byte per element
using Shape = std::vector<int64_t>;   // type_aliases.cr:16 (Int32 in the reference, Int64 here)
using Coord = std::vector<int64_t>;

template <class U> struct type_identity { using type = U; };
template <class U> using type_identity_t = typename type_identity<U>::type;

template <class T> struct DType;
#define PH_DTYPE_(T, code) template <> struct DType<T> { static constexpr int32_t value = code; }
PH_DTYPE_(float, PH_F32);   PH_DTYPE_(double, PH_F64);  PH_DTYPE_(int32_t, PH_I32); PH_DTYPE_(int64_t, PH_I64);
PH_DTYPE_(uint8_t, PH_U8);  PH_DTYPE_(int8_t, PH_I8);   PH_DTYPE_(int16_t, PH_I16); PH_DTYPE_(uint16_t, PH_U16);
PH_DTYPE_(uint32_t, PH_U32); PH_DTYPE_(uint64_t, PH_U64);
#undef PH_DTYPE_

namespace Device {
inline void check(int32_t status) {
  if (status != PH_OK) throw DeviceError(std::string("libphgpu status ") + std::to_string(status) + ": " + ph_last_error_string());
}
// ph_host.h status -> the exception class the reference raises at that point
inline void host_check(int32_t status) {
  if (status == PH_HOST_OK) return;
  const char* m = ph_host_last_error();
  std::string msg = m ? m : "";
  switch (status) {
    case PH_HOST_INDEX_ERROR: throw IndexError(msg);
    case PH_HOST_DIMENSION_ERROR: throw DimensionError(msg);
    case PH_HOST_SHAPE_ERROR: throw ShapeError(msg);
    case PH_HOST_DIV0_ERROR: throw DivisionByZeroError(msg);
    default: throw DeviceError("ph_host status " + std::to_string(status) + ": " + msg);
  }
}
inline bool& initialised() { static bool v = false; return v; }
inline void init(int32_t device = 0) { check(ph_init(device)); initialised() = true; }
inline void ensure_init() { if (!initialised()) init(0); }
inline void shutdown() { if (initialised()) { ph_shutdown(); initialised() = false; } }
inline void wait() { check(ph_sync()); }    // internal: keep a host temporary alive until it is copied; never raises
inline uint32_t take_flags() { uint32_t f = 0; check(ph_take_arith_flags(&f)); return f; }
// Data-dependent errors come back as a flag word after the stream sync and are re-raised as
// the classes the CPU path raises.
inline void raise_for(uint32_t f) {
  if (f & PH_FLAG_DIV0) throw DivisionByZeroError("Division by 0");
  if (f & PH_FLAG_OVERFLOW) throw OverflowError("Arithmetic overflow");
  if (f & PH_FLAG_ARGUMENT) throw ArgumentError("invalid integer argument (MIN // -1 or negative exponent)");
  if (f & PH_FLAG_NAN) throw ArgumentError("Comparison of NaN failed");
}
inline void raise_pending() { raise_for(take_flags()); }
// Every synchronising READ is a raise point: the copy and the flag word come back in one
// synchronisation (ph_d2h_flags), so `(a + b).to_host()` raises OverflowError like `a + b` does on the CPU.
inline void read_checked(void* dst_host, const void* src_dev, size_t nbytes) {
  uint32_t f = 0;
  check(ph_d2h_flags(dst_host, src_dev, nbytes, &f));
  raise_for(f);
}
inline void sync() { read_checked(nullptr, nullptr, 0); }
}  // namespace Device

// ---- region literals (range_syntax.cr:41-69).  Crystal writes a..b, a...b, a..s..b, `..`; here:
//   range(a, b)  range_ex(a, b)  range(a, s, b)  range_ex(a, s, b)  all      with `nil` for an open end
struct Nil {};
constexpr Nil nil{};
struct Bound {
  bool has = false;
  int64_t v = 0;
  Bound(Nil) {}
  template <class I, class = std::enable_if_t<std::is_integral<I>::value>> Bound(I x) : has(true), v((int64_t)x) {}
};
struct Lit {
  ph_range_lit l{};
  template <class I, class = std::enable_if_t<std::is_integral<I>::value>> Lit(I index) { l.is_index = 1; l.first = (int64_t)index; }
  Lit(Bound first, Bound last, bool exclusive, bool has_step = false, int64_t step = 0) {
    l.has_first = first.has; l.first = first.v; l.has_last = last.has; l.last = last.v;
    l.has_step = has_step; l.step = step; l.exclusive = exclusive;
  }
};
inline Lit range(Bound first, Bound last) { return Lit(first, last, false); }
inline Lit range_ex(Bound first, Bound last) { return Lit(first, last, true); }
inline Lit range(Bound first, int64_t step, Bound last) { return Lit(first, last, false, true, step); }
inline Lit range_ex(Bound first, int64_t step, Bound last) { return Lit(first, last, true, true, step); }
static const Lit all = range(nil, nil);
using RegionLiteral = std::vector<Lit>;

// ---- IndexRegion (index_region.cr:40-705): canonical {first, step, last, shape, degeneracy} ----
class IndexRegion {
 public:
  ph_region r{};
  IndexRegion() = default;
  // IndexRegion.new(region_literal, bound_shape, drop) :192-224
  IndexRegion(const RegionLiteral& lits, const Shape& bound_shape, bool drop = true) {
    std::vector<ph_range_lit> raw;
    for (const Lit& x : lits) raw.push_back(x.l);
    raw.push_back(ph_range_lit{});   // keep .data() valid for the empty literal
    Device::host_check(ph_region_new(raw.data(), (int32_t)lits.size(), bound_shape.data(), (int32_t)bound_shape.size(),
                                     drop ? 1 : 0, &r));
  }
  // IndexRegion.new(region_literal, bound_shape, drop, trim_to:) :133-168
  static IndexRegion trimmed(const RegionLiteral& lits, const Shape& bound_shape, const Shape& trim_to, bool drop = true) {
    IndexRegion out;
    std::vector<ph_range_lit> raw;
    for (const Lit& x : lits) raw.push_back(x.l);
    raw.push_back(ph_range_lit{});
    Device::host_check(ph_region_new_trimmed(raw.data(), (int32_t)lits.size(), bound_shape.data(), trim_to.data(),
                                             (int32_t)bound_shape.size(), drop ? 1 : 0, &out.r));
    return out;
  }
  static IndexRegion cover(const Shape& bound_shape, bool drop = true) {           // :232-238
    IndexRegion out;
    Device::host_check(ph_region_cover(bound_shape.data(), (int32_t)bound_shape.size(), drop ? 1 : 0, &out.r));
    return out;
  }
  Shape shape() const { return Shape(r.reduced_shape, r.reduced_shape + r.reduced_rank); }   // :323-335
  Shape proper_shape() const { return Shape(r.proper_shape, r.proper_shape + r.rank); }
  int64_t size() const { int64_t n = 1; for (int i = 0; i < r.rank; i++) n *= r.proper_shape[i]; return n; }
  bool fits_in(const Shape& bound) const {                                         // :468-478
    int32_t fits = 0;
    Device::host_check(ph_region_fits_in(&r, bound.data(), (int32_t)bound.size(), &fits));
    return fits != 0;
  }
  IndexRegion& trim(const Shape& bound) { Device::host_check(ph_region_trim(&r, bound.data(), (int32_t)bound.size())); return *this; }   // trim! :502-515
  IndexRegion& reverse() { Device::host_check(ph_region_reverse(&r)); return *this; }                                                 // reverse! :533-537
  IndexRegion& translate(const Coord& by) { Device::host_check(ph_region_translate(&r, by.data(), (int32_t)by.size())); return *this; } // translate! :577-585
};

// Ref-counted owner of one device allocation: reshape aliases the buffer (n_array.cr:429-433)
// and views keep their source alive (view.cr:7).
class DeviceBuffer {
 public:
  void* ptr = nullptr;
  size_t nbytes = 0;
  explicit DeviceBuffer(size_t n) : nbytes(n ? n : 1) {
    Device::ensure_init();
    Device::check(ph_alloc(nbytes, &ptr));
    home_ = ph_stream();              // the pool block is released on the stream it was handed out on
  }
  // A byte range of another buffer (one slice of a batched `slices` copy): keeps the parent alive and
  // never frees; the parent releases the whole allocation when the last range dies.
  DeviceBuffer(std::shared_ptr<DeviceBuffer> parent, size_t byte_offset, size_t n)
      : ptr(static_cast<char*>(parent->ptr) + byte_offset), nbytes(n), parent_(std::move(parent)) {}
  // Storage this layer does not release (peer-mapped blocks from ph_symm_alloc, whose release is collective:
  // ph_symm_free on every rank, or ph_comm_destroy)
  static std::shared_ptr<DeviceBuffer> adopt(void* external, size_t n) {
    std::shared_ptr<DeviceBuffer> b(new DeviceBuffer());
    b->ptr = external;
    b->nbytes = n;
    b->adopted_ = true;
    return b;
  }
  ~DeviceBuffer() { if (ptr && !parent_ && !adopted_) ph_free_on(ptr, home_); }
  void* home_stream() const { return parent_ ? parent_->home_stream() : home_; }
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;

 private:
  DeviceBuffer() = default;
  std::shared_ptr<DeviceBuffer> parent_;
  bool adopted_ = false;
  void* home_ = nullptr;
};

template <class T> class DeviceNArray;
template <class T> class DeviceView;

inline int64_t shape_to_size(const Shape& s) {   // ShapeUtil.shape_to_size (shape_util.cr:41-50): [] has size 0
  if (s.empty()) return 0;
  int64_t n = 1;
  for (int64_t e : s) n *= e;
  return n;
}
inline std::string shape_str(const Shape& s) {
  std::string out = "[";
  for (size_t i = 0; i < s.size(); i++) out += (i ? ", " : "") + std::to_string(s[i]);
  return out + "]";
}
inline ph_desc contiguous_desc(const Shape& s) {
  ph_desc d{};
  Device::host_check(ph_desc_contiguous(s.data(), (int32_t)s.size(), &d));
  return d;
}

// ---- the device-side MultiIndexable::Mutable(T) (multi_indexable.cr:30-65, mutable.cr) --------
// What DeviceNArray and DeviceView share: a buffer, ONE descriptor, a shape.
template <class T>
class MultiIndexable {
 public:
  using Elem = T;
  using DivResult = std::conditional_t<std::is_integral<T>::value, double, T>;   // Int / Int -> Float64

  const Shape& shape() const { return shape_; }
  int64_t size() const { return shape_to_size(shape_); }
  int32_t dimensions() const { return (int32_t)shape_.size(); }
  const ph_desc& desc() const { return desc_; }
  T* data() const { return static_cast<T*>(buf_->ptr); }
  const std::shared_ptr<DeviceBuffer>& buffer_owner() const { return buf_; }

  // ---- small host-side queries (multi_indexable.cr:100-237) ------------------------------------
  bool empty() const { return size() == 0; }                  // empty? :107-109
  bool scalar() const { return size() == 1; }                 // scalar? :119-121
  T first() const {                                           // :176-182
    if (size() == 0) throw ShapeError("This MultiIndexable has zero elements (shape: " + shape_str(shape_) + ").");
    return get(Coord(shape_.size(), 0));
  }
  T last() const {                                            // :197-203
    if (size() == 0) throw ShapeError("This MultiIndexable has zero elements (shape: " + shape_str(shape_) + ").");
    Coord c(shape_);
    for (int64_t& x : c) x -= 1;
    return get(c);
  }
  T to_scalar() const {                                       // :131-137
    if (!scalar())
      throw ShapeError("Only single-element MultiIndexables can be converted to scalars, but this one has " + std::to_string(size()) +
                       " elements (shape: " + shape_str(shape_) + ").");
    return first();
  }
  double to_f() const { return (double)to_scalar(); }         // :160-162

  // ---- blocks: out of scope on the device path, and they say so ---------------------------
  template <class... A> [[noreturn]] void map(A&&...) const { no_blocks("map"); }
  template <class... A> [[noreturn]] void map_with(A&&...) const { no_blocks("map_with"); }
  template <class... A> [[noreturn]] void map_with_coord(A&&...) const { no_blocks("map_with_coord"); }
  template <class... A> [[noreturn]] void each(A&&...) const { no_blocks("each"); }
  template <class... A> [[noreturn]] void each_with(A&&...) const { no_blocks("each_with"); }
  template <class... A> [[noreturn]] void each_coord(A&&...) const { no_blocks("each_coord"); }
  template <class... A> [[noreturn]] void fast_each(A&&...) const { no_blocks("fast_each"); }
  template <class... A> [[noreturn]] void apply(A&&...) const { no_blocks("apply"); }
  template <class... A> [[noreturn]] void process(A&&...) const { no_blocks("process"); }

  // ---- gather (multi_indexable.cr:338-356, 523-531 -> n_array.cr:450-453): ONE launch -----
  DeviceNArray<T> unsafe_fetch_chunk(const IndexRegion& region) const;
  DeviceNArray<T> get_chunk(const RegionLiteral& lits, bool drop = true) const { return unsafe_fetch_chunk(IndexRegion(lits, shape_, drop)); }
  DeviceNArray<T> get_chunk(const IndexRegion& region) const { return unsafe_fetch_chunk(region); }
  // get_chunk(coord, region_shape) :369-395: the block of `region_shape` whose lowermost corner is `coord`
  DeviceNArray<T> get_chunk(const Coord& coord, const Shape& region_shape) const {
    if (coord.size() != region_shape.size())
      throw DimensionError("'coord' and 'region_shape' had a different number of dimensions. Note that you must fully specify your coordinate and region shape for this overload of get_chunk.");
    if (coord.size() != shape_.size())
      throw DimensionError("'coord' had a different number of dimensions than this MultiIndexable (must have " + std::to_string(shape_.size()) + ", but has " + std::to_string(coord.size()) + ").");
    for (size_t i = 0; i < coord.size(); i++) {
      if (coord[i] < 0) throw ArgumentError("'coord' was negative on axis " + std::to_string(i) + ", but must be strictly nonnegative.");
      if (region_shape[i] < 0) throw ArgumentError("'region_shape' was negative on axis " + std::to_string(i) + ", but must be strictly nonnegative.");
      if (coord[i] + region_shape[i] > shape_[i])
        throw ShapeError("The region defined by shape " + shape_str(region_shape) + " and lowermost coordinate " + shape_str(coord) +
                         " is not contained within this MultiIndexable on axis " + std::to_string(i) + ".");
    }
    return unsafe_fetch_chunk(IndexRegion::cover(region_shape).translate(coord));
  }
  DeviceNArray<T> operator[](const RegionLiteral& lits) const { return get_chunk(lits); }
  // narr[mask] returns self (multi_indexable.cr:479-481); the spec-visible use is `narr[mask] = v`
  const MultiIndexable<T>& operator[](const MultiIndexable<Bool>&) const { return *this; }
  DeviceNArray<T> get_available(const RegionLiteral& lits, bool drop = true) const {   // :397-413
    return unsafe_fetch_chunk(IndexRegion::trimmed(lits, shape_, shape_, drop));
  }
  bool has_region(const RegionLiteral& lits, bool drop = true) const {                 // has_region? :313-318
    try { IndexRegion probe(lits, shape_, drop); (void)probe; return true; }
    catch (const IndexError&) { return false; }
    catch (const DimensionError&) { return false; }
  }
  // one element; legal but slow (one tiny D2H) -- the conformance tester enumerates coordinates
  T get(const Coord& coord) const {                                                     // :567-575
    T out;
    Device::read_checked(&out, data() + offset_of(coord), sizeof(T));
    return out;
  }
  T get_element(const Coord& coord) const { return get(coord); }
  bool has_coord(const Coord& coord) const {                                            // has_coord? coord_util.cr:9-21
    Coord canon(coord.size() + 1);
    return ph_canonicalize_coord(coord.data(), (int32_t)coord.size(), shape_.data(), (int32_t)shape_.size(), canon.data()) == PH_HOST_OK;
  }

  // ---- scatter / fill (multi_writable.cr:55-84 -> n_array.cr:484-500) ------------------------
  void set_element(const Coord& coord, T value) {
    Device::check(ph_h2d(data() + offset_of(coord), &value, sizeof(T)));
    Device::wait();   // `value` is a stack temporary
  }
  void set_chunk(const RegionLiteral& lits, const MultiIndexable<T>& src) { set_chunk(IndexRegion(lits, shape_), src); }
  void set_chunk(const RegionLiteral& lits, type_identity_t<T> value) { unsafe_set_chunk(IndexRegion(lits, shape_), value); }
  void set_chunk(const IndexRegion& region, const MultiIndexable<T>& src) {
    Shape rs = region.shape();
    int32_t ok = 0;
    Device::host_check(ph_shapes_compatible(src.shape().data(), (int32_t)src.shape().size(), rs.data(), (int32_t)rs.size(), &ok));
    if (!ok)   // multi_writable.cr:58-60
      throw ShapeError("Cannot substitute: the given array has shape " + shape_str(src.shape()) + ", but the region has shape " + shape_str(rs) + ".");
    unsafe_set_chunk(region, src);
  }
  void unsafe_set_chunk(const IndexRegion& region, const MultiIndexable<T>& src);
  void unsafe_set_chunk(const IndexRegion& region, T value) {
    ph_desc dst{};
    Device::host_check(ph_desc_region(&desc_, &region.r, &dst));
    if (desc_count(dst) == 0) return;
    Device::check(ph_fill_region((int32_t)sizeof(T), buf_->ptr, &dst, &value));
  }

  // ---- masked store (n_array.cr:510-551) -------------------------------------------------
  void set_mask(const MultiIndexable<Bool>& mask, type_identity_t<T> value) {
    if (mask.shape() != shape_)   // :511-513
      throw DimensionError("Cannot perform masking: mask shape " + shape_str(mask.shape()) + " does not match array shape " + shape_str(shape_) + ".");
    if (size() == 0) return;
    Device::check(ph_mask_set_scalar((int32_t)sizeof(T), buf_->ptr, &desc_, mask.data(), &mask.desc(), &value));
  }
  void set_mask(const MultiIndexable<Bool>& mask, const MultiIndexable<T>& value) {
    if (mask.shape() != shape_)
      throw DimensionError("Cannot perform masking: mask shape " + shape_str(mask.shape()) + " does not match array shape " + shape_str(shape_) + ".");
    if (value.shape() != shape_)   // :525-527
      throw DimensionError("Cannot perform masking: value shape " + shape_str(value.shape()) + " does not match array shape " + shape_str(shape_) + ".");
    if (size() == 0) return;
    Device::check(ph_mask_set_array((int32_t)sizeof(T), buf_->ptr, &desc_, mask.data(), &mask.desc(), value.buffer_owner()->ptr, &value.desc()));
  }

  // ---- views turned into copies (view.cr:123-126, multi_indexable.cr:795-803, 852-856) ------
  DeviceNArray<T> to_narr() const;
  std::vector<T> to_host() const;      // explicit device -> host transfer of the (materialised) contents
  DeviceView<T> view() const;
  DeviceView<T> view(const RegionLiteral& lits, bool drop = true) const;
  DeviceView<T> mutable_view() const;
  DeviceView<T> mutable_view(const RegionLiteral& lits, bool drop = true) const;

  // ---- slices / tile (multi_indexable.cr:742-786, 818-843) ------------------------------------
  std::vector<DeviceNArray<T>> slices(int32_t axis = 0) const;
  // each_slice (:742-748): the slices as VIEWS over this array's buffer -- descriptors only, no copy and no
  // launch -- so the reference's per-axis idiom (a fold over each_slice) costs just the consumer's kernels,
  // which read the strided slices directly; `slices` makes independent arrays with one batched copy.
  std::vector<DeviceView<T>> each_slice(int32_t axis = 0) const;
  DeviceNArray<T> tile(const Shape& counts) const;

  // ---- elementwise, named forms (operators are free functions below) --------------------------
  DeviceNArray<T> binary(int32_t op, const MultiIndexable<T>& other, const char* name) const;
  DeviceNArray<T> scalar(int32_t op, T s, bool scalar_on_left) const;
  DeviceNArray<DivResult> divide(const MultiIndexable<T>& other) const;
  DeviceNArray<DivResult> divide(T s, bool scalar_on_left) const;
  DeviceNArray<T> floordiv(const MultiIndexable<T>& o) const { return binary(PH_FLOORDIV, o, "//"); }
  DeviceNArray<T> floordiv(type_identity_t<T> s) const { return scalar(PH_FLOORDIV, s, false); }
  DeviceNArray<T> pow(const MultiIndexable<T>& o) const { return binary(PH_POW, o, "**"); }
  // Float ** Int32 is llvm.powi (bit-exact); everything else is T ** T
  template <class E, class = std::enable_if_t<std::is_arithmetic<E>::value>> DeviceNArray<T> pow(E e) const;
  DeviceNArray<T> wrapping_add(const MultiIndexable<T>& o) const { return binary(PH_WADD, o, "&+"); }
  DeviceNArray<T> wrapping_sub(const MultiIndexable<T>& o) const { return binary(PH_WSUB, o, "&-"); }
  DeviceNArray<T> wrapping_mul(const MultiIndexable<T>& o) const { return binary(PH_WMUL, o, "&*"); }
  DeviceNArray<T> wrapping_pow(const MultiIndexable<T>& o) const { return binary(PH_WPOW, o, "&**"); }
  DeviceNArray<T> wrapping_add(type_identity_t<T> s) const { return scalar(PH_WADD, s, false); }
  DeviceNArray<T> wrapping_sub(type_identity_t<T> s) const { return scalar(PH_WSUB, s, false); }
  DeviceNArray<T> wrapping_mul(type_identity_t<T> s) const { return scalar(PH_WMUL, s, false); }
  DeviceNArray<T> wrapping_pow(type_identity_t<T> s) const { return scalar(PH_WPOW, s, false); }
  DeviceNArray<T> unary(int32_t op) const;
  // NEW (no broadcasting exists in the reference; = tile + op): equal rank, size-1 axes stretch
  DeviceNArray<T> broadcast_op(int32_t op, const MultiIndexable<T>& other) const;
  // NEW fused (self * b) + c, two roundings; b and c may broadcast
  DeviceNArray<T> mul_add(const MultiIndexable<T>& b, const MultiIndexable<T>& c) const;

  DeviceNArray<Bool> compare(int32_t cmp, const MultiIndexable<T>& other, bool eq_style = false) const;
  DeviceNArray<Bool> compare(int32_t cmp, T s, bool scalar_on_left = false) const;
  DeviceNArray<Bool> eq(const MultiIndexable<T>& o) const;     // :899-913
  DeviceNArray<Bool> eq(type_identity_t<T> s) const;
  DeviceNArray<Bool> match(type_identity_t<T> s) const;        // =~ :916-920
  // <=> of the operator list (:960-981): -1 / 0 / 1 as Int32; integer element types only (Float#<=> is Int32?)
  DeviceNArray<int32_t> cmp(const MultiIndexable<T>& other) const;
  DeviceNArray<int32_t> cmp(type_identity_t<T> s) const;
  bool equals(const MultiIndexable<T>& other) const;                                            // NArray#== n_array.cr:440-447

  // ---- reductions (Enumerable over NArray#each, n_array.cr:556-564) ---------------------------
  T sum() const { return size() == 0 ? T(0) : reduce_full(PH_SUM).first; }
  T min() const { return reduce_full(PH_MIN).first; }
  T max() const { return reduce_full(PH_MAX).first; }
  // README.md:56-61 idiom: {max, coord of the FIRST maximum}
  std::pair<T, Coord> argmax() const { auto r = reduce_full(PH_ARGMAX); return {r.first, index_to_coord(r.second)}; }
  std::pair<T, Coord> argmin() const { auto r = reduce_full(PH_ARGMIN); return {r.first, index_to_coord(r.second)}; }
  // per axis (= each_slice(axis) + fold, multi_indexable.cr:742-748): the shape loses `axis`
  DeviceNArray<T> sum(int32_t axis) const { return reduce_axis<T>(PH_SUM, axis); }
  DeviceNArray<T> min(int32_t axis) const { return reduce_axis<T>(PH_MIN, axis); }
  DeviceNArray<T> max(int32_t axis) const { return reduce_axis<T>(PH_MAX, axis); }
  DeviceNArray<int64_t> argmax(int32_t axis) const;
  DeviceNArray<int64_t> argmin(int32_t axis) const;
  Coord index_to_coord(int64_t index) const {   // Buffered.index_to_coord (buffered.cr:58-77)
    Coord c(shape_.size());
    for (size_t i = shape_.size(); i-- > 0;) { c[i] = index % shape_[i]; index /= shape_[i]; }
    return c;
  }

 protected:
  std::shared_ptr<DeviceBuffer> buf_;
  ph_desc desc_{};
  Shape shape_;

  MultiIndexable() = default;
  MultiIndexable(std::shared_ptr<DeviceBuffer> b, const ph_desc& d, Shape s) : buf_(std::move(b)), desc_(d), shape_(std::move(s)) {}

  [[noreturn]] static void no_blocks(const char* what) {
    throw DeviceBlockError(std::string(what) + ": arbitrary blocks cannot run on the device path");
  }
  static int64_t desc_count(const ph_desc& d) { int64_t n = 1; for (int i = 0; i < d.rank; i++) n *= d.extent[i]; return n; }
  int64_t offset_of(const Coord& coord) const {   // canonicalize_coord (coord_util.cr:76-82) + coord_to_index (buffered.cr:44-52)
    Coord canon(coord.size() + 1);
    Device::host_check(ph_canonicalize_coord(coord.data(), (int32_t)coord.size(), shape_.data(), (int32_t)shape_.size(), canon.data()));
    int64_t off = 0;
    Device::host_check(ph_desc_offset_of(&desc_, canon.data(), (int32_t)coord.size(), &off));
    return off;
  }
  ph_desc bcast_desc(const Shape& to) const {
    ph_desc d{};
    Device::host_check(ph_desc_broadcast(&desc_, to.data(), (int32_t)to.size(), &d));
    return d;
  }
  std::pair<T, int64_t> reduce_full(int32_t red) const {
    if (size() == 0) throw EmptyError("Empty enumerable");
    T val{};
    int64_t idx = -1;
    // record mode (the sharded entry on one process exchanges nothing): one launch, the finishing block writes
    // value, index and the pending flags into a pinned host record -- no copy, no second read for the flags
    alignas(16) unsigned char cell[16] = {0};
    uint32_t flags = 0;
    Device::check(ph_reduce_full_sharded(red, DType<T>::value, buf_->ptr, &desc_, 0, cell, &idx, &flags));
    Device::raise_for(flags);
    std::memcpy(&val, cell, sizeof(T));
    return {val, idx};
  }
  template <class R> DeviceNArray<R> reduce_axis(int32_t red, int32_t axis) const;
  template <class U> friend class MultiIndexable;
};

// ---- DeviceNArray<T>: row-major N-D array resident in HBM (n_array.cr:15) -------------------
template <class T>
class DeviceNArray : public MultiIndexable<T> {
  using Base = MultiIndexable<T>;

 public:
  // uninitialised storage of `shape` (NArray.new(shape) { block } would need a block: use fill / from_host)
  explicit DeviceNArray(const Shape& shape)
      : Base(std::make_shared<DeviceBuffer>((size_t)shape_to_size(shape) * sizeof(T)), contiguous_desc(shape), shape) {}
  DeviceNArray(const Shape& shape, std::shared_ptr<DeviceBuffer> alias) : Base(std::move(alias), contiguous_desc(shape), shape) {}

  // NArray#to_device: explicit host -> device transfer
  static DeviceNArray from_host(const Shape& shape, const T* host) {
    DeviceNArray out(shape);
    if (out.size()) {
      Device::check(ph_h2d(out.data(), host, (size_t)out.size() * sizeof(T)));
      Device::wait();   // a pageable source must stay alive until copied
    }
    return out;
  }
  static DeviceNArray from_host(const Shape& shape, const std::vector<T>& host) {
    if ((int64_t)host.size() != shape_to_size(shape)) throw ShapeError("from_host: " + std::to_string(host.size()) + " elements for shape " + shape_str(shape));
    return from_host(shape, host.data());
  }
  static DeviceNArray fill(const Shape& shape, T value) {   // NArray.fill n_array.cr:230-232
    DeviceNArray out(shape);
    if (out.size()) Device::check(ph_fill_region((int32_t)sizeof(T), out.data(), &out.desc(), &value));
    return out;
  }
  template <class... A> [[noreturn]] static DeviceNArray build(A&&...) { Base::no_blocks("build"); }

  DeviceNArray clone() const {   // deep copy n_array.cr:372-374
    DeviceNArray out(this->shape_);
    if (this->size()) Device::check(ph_d2d(out.data(), this->data(), (size_t)this->size() * sizeof(T)));
    return out;
  }
  DeviceNArray dup() const { return clone(); }
  DeviceNArray reshape(const Shape& new_shape) const {   // ALIASES the buffer n_array.cr:429-433
    if (shape_to_size(new_shape) != this->size())
      throw ShapeError("Cannot change shape from " + shape_str(this->shape_) + " to " + shape_str(new_shape) + ": reshape cannot add or remove elements.");
    return DeviceNArray(new_shape, this->buf_);
  }
  DeviceNArray flatten() const { return reshape({this->size()}); }
  // ---- joins (n_array.cr:321-344, 666-750): one strided copy per input ----------------------
  static DeviceNArray concatenate(const std::vector<const MultiIndexable<T>*>& narrs, int32_t axis = 0);   // NArray.concatenate :722-724
  DeviceNArray concatenate(const MultiIndexable<T>& other, int32_t axis = 0) const { return concatenate({this, &other}, axis); }   // :712-714
  DeviceNArray& push(const std::vector<const MultiIndexable<T>*>& others, int32_t axis = 0);               // push :688-710, in place
  DeviceNArray& operator<<(const MultiIndexable<T>& other) { return push({&other}); }                      // << :682-684
  static DeviceNArray wrap(const std::vector<const MultiIndexable<T>*>& narrs);                            // NArray.wrap :321-340
  // copying permute / reverse = view + to_narr (multi_indexable.cr:795-803)
  DeviceNArray permute(const std::vector<int32_t>& order = {}) const;
  DeviceNArray reverse() const;
};

// ---- DeviceView<T>: source buffer + ONE descriptor (view.cr, mutable_view.cr) ---------------
// Region / Permute / Reverse transforms (view_util/transforms.cr) are affine in the coordinate,
// so a chain folds into (offset, extent[], stride[]) as it is built; Reshape folds when it is
// expressible in strides and otherwise materialises first.  Reads gather, writes scatter.
template <class T>
class DeviceView : public MultiIndexable<T> {
  using Base = MultiIndexable<T>;

 public:
  DeviceView(std::shared_ptr<DeviceBuffer> b, const ph_desc& d, Shape s) : Base(std::move(b), d, std::move(s)) {}
  DeviceView view() const { return *this; }
  DeviceView view(const RegionLiteral& lits, bool drop = true) const { return restrict_to(IndexRegion(lits, this->shape_, drop)); }
  DeviceView restrict_to(const IndexRegion& region) const {   // view.cr:43-56
    ph_desc d{};
    Device::host_check(ph_desc_region(&this->desc_, &region.r, &d));
    return DeviceView(this->buf_, d, region.shape());
  }
  DeviceView unsafe_fetch_chunk(const IndexRegion& region) const { return restrict_to(region); }   // view.cr:105-107
  DeviceView get_chunk(const RegionLiteral& lits, bool drop = true) const { return view(lits, drop); }
  DeviceView operator[](const RegionLiteral& lits) const { return view(lits); }
  DeviceView permute(const std::vector<int32_t>& order = {}) const {   // permute! view.cr:72-81; {} = reversed axes
    ph_desc d{};
    Device::host_check(ph_desc_permute(&this->desc_, order.empty() ? nullptr : order.data(), (int32_t)order.size(), &d));
    return DeviceView(this->buf_, d, Shape(d.extent, d.extent + d.rank));
  }
  DeviceView reverse() const {   // reverse! view.cr:96-99
    ph_desc d{};
    Device::host_check(ph_desc_reverse(&this->desc_, &d));
    return DeviceView(this->buf_, d, this->shape_);
  }
  DeviceView reshape(const Shape& new_shape) const {   // reshape! view.cr:58-66
    ph_desc d{};
    int32_t st = ph_desc_reshape(&this->desc_, new_shape.data(), (int32_t)new_shape.size(), &d);
    if (st == PH_HOST_NEEDS_COPY) return this->to_narr().view().reshape(new_shape);
    Device::host_check(st);
    return DeviceView(this->buf_, d, new_shape);
  }
};

// =============================== implementation ===============================================
template <class T>
DeviceNArray<T> MultiIndexable<T>::unsafe_fetch_chunk(const IndexRegion& region) const {
  ph_desc src{};
  Device::host_check(ph_desc_region(&desc_, &region.r, &src));
  DeviceNArray<T> out(region.shape());
  if (out.size()) Device::check(ph_copy_strided((int32_t)sizeof(T), buf_->ptr, &src, out.data(), &out.desc()));
  return out;
}

// An array source is streamed in ITS OWN lex order onto the region's lex order (n_array.cr:484-492).
template <class T>
void MultiIndexable<T>::unsafe_set_chunk(const IndexRegion& region, const MultiIndexable<T>& src) {
  ph_desc dst{};
  Device::host_check(ph_desc_region(&desc_, &region.r, &dst));
  if (desc_count(dst) == 0) return;
  Shape rs(dst.extent, dst.extent + dst.rank);
  // compatible_shapes? allows trailing ones: view the source with the region's extents
  ph_desc sd{};
  int32_t st = ph_desc_reshape(&src.desc(), rs.data(), (int32_t)rs.size(), &sd);
  if (st == PH_HOST_NEEDS_COPY) {
    DeviceNArray<T> tmp = src.to_narr();
    Device::host_check(ph_desc_reshape(&tmp.desc(), rs.data(), (int32_t)rs.size(), &sd));
    Device::check(ph_copy_strided((int32_t)sizeof(T), tmp.data(), &sd, buf_->ptr, &dst));
    return;
  }
  Device::host_check(st);
  Device::check(ph_copy_strided((int32_t)sizeof(T), src.buffer_owner()->ptr, &sd, buf_->ptr, &dst));
}

template <class T>
DeviceNArray<T> MultiIndexable<T>::to_narr() const {
  DeviceNArray<T> out(shape_);
  if (out.size()) Device::check(ph_copy_strided((int32_t)sizeof(T), buf_->ptr, &desc_, out.data(), &out.desc()));
  return out;
}

template <class T>
std::vector<T> MultiIndexable<T>::to_host() const {
  DeviceNArray<T> flat = to_narr();   // one gather; a whole contiguous array takes the flat copy kernel
  std::vector<T> out((size_t)flat.size());
  Device::read_checked(out.empty() ? nullptr : out.data(), flat.data(), out.size() * sizeof(T));   // a raise point
  return out;
}

template <class T> DeviceView<T> MultiIndexable<T>::view() const { return DeviceView<T>(buf_, desc_, shape_); }
template <class T> DeviceView<T> MultiIndexable<T>::view(const RegionLiteral& lits, bool drop) const { return view().view(lits, drop); }
template <class T> DeviceView<T> MultiIndexable<T>::mutable_view() const { return view(); }
template <class T> DeviceView<T> MultiIndexable<T>::mutable_view(const RegionLiteral& lits, bool drop) const { return view(lits, drop); }

// The reference gathers the slices one by one (ChunkIterator -> unsafe_fetch_chunk per index): on the
// device one launch per slice, launch-bound for every axis but the leading one.  All slices along `axis`
// together ARE the array with `axis` moved to the front, so ONE permuting copy produces them and they are
// handed out as consecutive ranges of its buffer (disjoint, so still independent arrays).
template <class T>
std::vector<DeviceNArray<T>> MultiIndexable<T>::slices(int32_t axis) const {
  if (axis < 0 || axis >= dimensions()) throw IndexError("axis " + std::to_string(axis) + " is out of range for shape " + shape_str(shape_));
  const int64_t n = shape_[axis];
  Shape rest;
  for (int32_t i = 0; i < dimensions(); i++) if (i != axis) rest.push_back(shape_[i]);
  if (rest.empty()) rest.push_back(1);
  std::vector<DeviceNArray<T>> out;
  out.reserve((size_t)n);
  if (n == 0 || size() == 0) {
    for (int64_t i = 0; i < n; i++) out.emplace_back(rest);
    return out;
  }
  std::vector<int32_t> order{axis};
  for (int32_t i = 0; i < dimensions(); i++) if (i != axis) order.push_back(i);
  DeviceNArray<T> moved = view().permute(order).to_narr();   // one launch
  const size_t step = (size_t)(size() / n) * sizeof(T);
  for (int64_t i = 0; i < n; i++)
    out.emplace_back(rest, std::make_shared<DeviceBuffer>(moved.buffer_owner(), (size_t)i * step, step));
  return out;
}

template <class T>
std::vector<DeviceView<T>> MultiIndexable<T>::each_slice(int32_t axis) const {
  if (axis < 0 || axis >= dimensions()) throw IndexError("axis " + std::to_string(axis) + " is out of range for shape " + shape_str(shape_));
  ph_desc d{};
  Shape rest;
  for (int32_t i = 0; i < dimensions(); i++) {
    if (i == axis) continue;
    d.extent[d.rank] = desc_.extent[i];
    d.stride[d.rank] = desc_.stride[i];
    d.rank++;
    rest.push_back(shape_[i]);
  }
  if (rest.empty()) { d.rank = 1; d.extent[0] = 1; d.stride[0] = 1; rest.push_back(1); }   // slices of a vector have shape [1]
  std::vector<DeviceView<T>> out;
  out.reserve((size_t)shape_[axis]);
  for (int64_t i = 0; i < shape_[axis]; i++) {
    d.offset = desc_.offset + i * desc_.stride[axis];
    out.emplace_back(buf_, d, rest);
  }
  return out;
}

// out[c] = self[c % shape] (multi_indexable.cr:818-827); as a descriptor every axis becomes
// (count, extent) with strides (0, stride): no modulo on the device.
template <class T>
DeviceNArray<T> MultiIndexable<T>::tile(const Shape& counts) const {
  if (counts.size() != shape_.size()) throw DimensionError("tile counts have the wrong number of dimensions");
  if (2 * shape_.size() > PH_MAX_RANK) throw ShapeError("tile supports rank <= 4 on the device path");
  ph_desc src{};
  src.rank = (int32_t)(2 * shape_.size());
  src.offset = desc_.offset;
  Shape ext, out_shape;
  for (size_t i = 0; i < shape_.size(); i++) {
    src.extent[2 * i] = counts[i];         src.stride[2 * i] = 0;
    src.extent[2 * i + 1] = desc_.extent[i]; src.stride[2 * i + 1] = desc_.stride[i];
    ext.push_back(counts[i]); ext.push_back(desc_.extent[i]);
    out_shape.push_back(counts[i] * shape_[i]);
  }
  DeviceNArray<T> out(out_shape);
  ph_desc od = contiguous_desc(ext);
  if (out.size()) Device::check(ph_copy_strided((int32_t)sizeof(T), buf_->ptr, &src, out.data(), &od));
  return out;
}

template <class T>
DeviceNArray<T> MultiIndexable<T>::binary(int32_t op, const MultiIndexable<T>& other, const char* name) const {
  if (shape_ != other.shape_)   // multi_indexable.cr:935-940, word for word
    throw ShapeError("The shape of this MultiIndexable (" + shape_str(shape_) + ") does not match the shape of the one provided (" +
                     shape_str(other.shape_) + "), so '" + name + "' cannot be applied element-wise.");
  DeviceNArray<T> out(shape_);
  if (out.size()) Device::check(ph_ewise_binary(op, DType<T>::value, buf_->ptr, &desc_, other.buf_->ptr, &other.desc_, out.data(), &out.desc()));
  return out;
}

template <class T>
DeviceNArray<T> MultiIndexable<T>::scalar(int32_t op, T s, bool scalar_on_left) const {
  DeviceNArray<T> out(shape_);
  if (out.size()) Device::check(ph_ewise_scalar(op, DType<T>::value, buf_->ptr, &desc_, &s, scalar_on_left ? 1 : 0, out.data(), &out.desc()));
  return out;
}

template <class T>
DeviceNArray<typename MultiIndexable<T>::DivResult> MultiIndexable<T>::divide(const MultiIndexable<T>& other) const {
  if (shape_ != other.shape_)
    throw ShapeError("The shape of this MultiIndexable (" + shape_str(shape_) + ") does not match the shape of the one provided (" +
                     shape_str(other.shape_) + "), so '/' cannot be applied element-wise.");
  DeviceNArray<DivResult> out(shape_);
  if (out.size()) Device::check(ph_ewise_binary(PH_DIV, DType<T>::value, buf_->ptr, &desc_, other.buf_->ptr, &other.desc_, out.data(), &out.desc()));
  return out;
}

template <class T>
DeviceNArray<typename MultiIndexable<T>::DivResult> MultiIndexable<T>::divide(T s, bool scalar_on_left) const {
  DeviceNArray<DivResult> out(shape_);
  if (out.size()) Device::check(ph_ewise_scalar(PH_DIV, DType<T>::value, buf_->ptr, &desc_, &s, scalar_on_left ? 1 : 0, out.data(), &out.desc()));
  return out;
}

template <class T>
template <class E, class>
DeviceNArray<T> MultiIndexable<T>::pow(E e) const {
  if (std::is_floating_point<T>::value && std::is_integral<E>::value) {
    int32_t k = (int32_t)e;
    DeviceNArray<T> out(shape_);
    if (out.size()) Device::check(ph_ewise_scalar(PH_POWI, DType<T>::value, buf_->ptr, &desc_, &k, 0, out.data(), &out.desc()));
    return out;
  }
  return scalar(PH_POW, (T)e, false);
}

template <class T>
DeviceNArray<T> MultiIndexable<T>::unary(int32_t op) const {
  DeviceNArray<T> out(shape_);
  if (out.size()) Device::check(ph_ewise_unary(op, DType<T>::value, buf_->ptr, &desc_, out.data(), &out.desc()));
  return out;
}

template <class T>
DeviceNArray<T> MultiIndexable<T>::broadcast_op(int32_t op, const MultiIndexable<T>& other) const {
  if (shape_.size() != other.shape_.size()) throw ShapeError("broadcast requires equal rank");
  Shape to(shape_.size() + 1);
  Device::host_check(ph_broadcast_shapes(shape_.data(), other.shape_.data(), (int32_t)shape_.size(), to.data()));
  to.resize(shape_.size());
  DeviceNArray<T> out(to);
  ph_desc da = bcast_desc(to), db = other.bcast_desc(to);
  if (out.size()) Device::check(ph_ewise_binary(op, DType<T>::value, buf_->ptr, &da, other.buf_->ptr, &db, out.data(), &out.desc()));
  return out;
}

template <class T>
DeviceNArray<T> MultiIndexable<T>::mul_add(const MultiIndexable<T>& b, const MultiIndexable<T>& c) const {
  DeviceNArray<T> out(shape_);
  ph_desc db = b.bcast_desc(shape_), dc = c.bcast_desc(shape_);
  if (out.size())
    Device::check(ph_ewise_mul_add(DType<T>::value, buf_->ptr, &desc_, b.buf_->ptr, &db, c.buf_->ptr, &dc, out.data(), &out.desc()));
  return out;
}

template <class T>
DeviceNArray<Bool> MultiIndexable<T>::compare(int32_t cmp, const MultiIndexable<T>& other, bool eq_style) const {
  if (shape_ != other.shape_) {
    if (eq_style) throw DimensionError("Cannot compute the element-wise equality between shapes " + shape_str(shape_) + " and " + shape_str(other.shape_));   // :900-902
    throw ShapeError("The shape of this MultiIndexable (" + shape_str(shape_) + ") does not match the shape of the one provided (" +
                     shape_str(other.shape_) + "), so the comparison cannot be applied element-wise.");
  }
  DeviceNArray<Bool> out(shape_);
  if (out.size()) Device::check(ph_compare(cmp, DType<T>::value, buf_->ptr, &desc_, other.buf_->ptr, &other.desc_, out.data(), &out.desc()));
  return out;
}

template <class T>
DeviceNArray<Bool> MultiIndexable<T>::compare(int32_t cmp, T s, bool scalar_on_left) const {
  DeviceNArray<Bool> out(shape_);
  if (out.size()) Device::check(ph_compare_scalar(cmp, DType<T>::value, buf_->ptr, &desc_, &s, scalar_on_left ? 1 : 0, out.data(), &out.desc()));
  return out;
}

template <class T>
DeviceNArray<int32_t> MultiIndexable<T>::cmp(const MultiIndexable<T>& other) const {
  static_assert(std::is_integral<T>::value, "<=> on the device path is defined for integer element types (Float#<=> is nilable)");
  if (shape_ != other.shape_)
    throw ShapeError("The shape of this MultiIndexable (" + shape_str(shape_) + ") does not match the shape of the one provided (" +
                     shape_str(other.shape_) + "), so '<=>' cannot be applied element-wise.");
  DeviceNArray<int32_t> out(shape_);
  if (out.size()) Device::check(ph_compare3(DType<T>::value, buf_->ptr, &desc_, other.buf_->ptr, &other.desc_, out.data(), &out.desc()));
  return out;
}
template <class T>
DeviceNArray<int32_t> MultiIndexable<T>::cmp(type_identity_t<T> s) const {
  static_assert(std::is_integral<T>::value, "<=> on the device path is defined for integer element types (Float#<=> is nilable)");
  DeviceNArray<int32_t> out(shape_);
  if (out.size()) Device::check(ph_compare3_scalar(DType<T>::value, buf_->ptr, &desc_, &s, 0, out.data(), &out.desc()));
  return out;
}

template <class T> DeviceNArray<Bool> MultiIndexable<T>::eq(const MultiIndexable<T>& o) const { return compare(PH_EQ, o, true); }
template <class T> DeviceNArray<Bool> MultiIndexable<T>::eq(type_identity_t<T> s) const { return compare(PH_EQ, s); }
template <class T> DeviceNArray<Bool> MultiIndexable<T>::match(type_identity_t<T> s) const { return compare(PH_EQ, s); }
template <class T> DeviceNArray<int64_t> MultiIndexable<T>::argmax(int32_t axis) const { return reduce_axis<int64_t>(PH_ARGMAX, axis); }
template <class T> DeviceNArray<int64_t> MultiIndexable<T>::argmin(int32_t axis) const { return reduce_axis<int64_t>(PH_ARGMIN, axis); }

template <class T>
bool MultiIndexable<T>::equals(const MultiIndexable<T>& other) const {
  if (shape_ != other.shape_) return false;
  if (size() == 0) return true;
  return eq(other).min() != 0;
}

template <class T>
template <class R>
DeviceNArray<R> MultiIndexable<T>::reduce_axis(int32_t red, int32_t axis) const {
  if (axis < 0 || axis >= dimensions()) throw IndexError("axis " + std::to_string(axis) + " is not present in a " + std::to_string(dimensions()) + "-dimensional MultiIndexable");
  if (shape_[axis] == 0 && red != PH_SUM) throw EmptyError("Empty enumerable");
  Shape os;
  for (int32_t i = 0; i < dimensions(); i++) if (i != axis) os.push_back(shape_[i]);
  if (os.empty()) os.push_back(1);
  DeviceNArray<R> out(os);
  if (out.size()) {
    if (shape_[axis] == 0) {
      R zero{};
      Device::check(ph_fill_region((int32_t)sizeof(R), out.data(), &out.desc(), &zero));
    } else {
      Device::check(ph_reduce_axis(red, DType<T>::value, buf_->ptr, &desc_, axis, out.data(), &out.desc()));
    }
  }
  Device::raise_pending();
  return out;
}

template <class T> DeviceNArray<T> DeviceNArray<T>::permute(const std::vector<int32_t>& order) const { return this->view().permute(order).to_narr(); }
template <class T> DeviceNArray<T> DeviceNArray<T>::reverse() const { return this->view().reverse().to_narr(); }

// ---- infix operators: the op list of multi_indexable.cr:960-985 and patches/number.cr:6-15 ----
#define PH_INFIX_(sym, code)                                                                                       \
  template <class T> DeviceNArray<T> operator sym(const MultiIndexable<T>& a, const MultiIndexable<T>& b) { return a.binary(code, b, #sym); } \
  template <class T> DeviceNArray<T> operator sym(const MultiIndexable<T>& a, type_identity_t<T> s) { return a.scalar(code, s, false); }      \
  template <class T> DeviceNArray<T> operator sym(type_identity_t<T> s, const MultiIndexable<T>& a) { return a.scalar(code, s, true); }
PH_INFIX_(+, PH_ADD)
PH_INFIX_(-, PH_SUB)
PH_INFIX_(*, PH_MUL)
PH_INFIX_(%, PH_MOD)
PH_INFIX_(&, PH_AND)
PH_INFIX_(|, PH_OR)
PH_INFIX_(^, PH_XOR)
#undef PH_INFIX_
template <class T> auto operator/(const MultiIndexable<T>& a, const MultiIndexable<T>& b) { return a.divide(b); }
template <class T> auto operator/(const MultiIndexable<T>& a, type_identity_t<T> s) { return a.divide(s, false); }
template <class T> auto operator/(type_identity_t<T> s, const MultiIndexable<T>& a) { return a.divide(s, true); }
template <class T> DeviceNArray<T> operator+(const MultiIndexable<T>& a) { return a.unary(PH_POS); }
template <class T> DeviceNArray<T> operator-(const MultiIndexable<T>& a) { return a.unary(PH_NEG); }
template <class T> DeviceNArray<T> operator~(const MultiIndexable<T>& a) { return a.unary(PH_NOT); }
#define PH_CMP_(sym, code, mirrored)                                                                               \
  template <class T> DeviceNArray<Bool> operator sym(const MultiIndexable<T>& a, const MultiIndexable<T>& b) { return a.compare(code, b); } \
  template <class T> DeviceNArray<Bool> operator sym(const MultiIndexable<T>& a, type_identity_t<T> s) { return a.compare(code, s, false); } \
  template <class T> DeviceNArray<Bool> operator sym(type_identity_t<T> s, const MultiIndexable<T>& a) { return a.compare(code, s, true); }
PH_CMP_(>, PH_GT, PH_LT)
PH_CMP_(<, PH_LT, PH_GT)
PH_CMP_(>=, PH_GE, PH_LE)
PH_CMP_(<=, PH_LE, PH_GE)
#undef PH_CMP_
template <class T> bool operator==(const MultiIndexable<T>& a, const MultiIndexable<T>& b) { return a.equals(b); }
template <class T> bool operator!=(const MultiIndexable<T>& a, const MultiIndexable<T>& b) { return !a.equals(b); }

// ---- joins: NArray.concatenate / push / wrap (n_array.cr:321-344, 666-750) ----------------------------------
// The shape rule is the reference's `compatible?` (ph_concat_shape: `idx != axis` on the raw argument, so a negative
// axis excludes nothing and every dimension must then match).  Each input is ONE scatter into its range of the result.
inline Shape concat_shape_of(const std::vector<Shape>& shapes, int32_t axis, int32_t* canonical_axis) {
  std::vector<int64_t> packed(shapes.size() * PH_MAX_RANK, 0);
  std::vector<int32_t> ranks(shapes.size());
  for (size_t k = 0; k < shapes.size(); k++) {
    if (shapes[k].size() > (size_t)PH_MAX_RANK) throw ShapeError("the device path supports rank <= " + std::to_string(PH_MAX_RANK));
    ranks[k] = (int32_t)shapes[k].size();
    for (size_t i = 0; i < shapes[k].size(); i++) packed[k * PH_MAX_RANK + i] = shapes[k][i];
  }
  int64_t out[PH_MAX_RANK] = {0};
  Device::host_check(ph_concat_shape(packed.data(), ranks.data(), (int32_t)shapes.size(), axis, out, canonical_axis));
  return Shape(out, out + shapes[0].size());
}
template <class T>
DeviceNArray<T> DeviceNArray<T>::concatenate(const std::vector<const MultiIndexable<T>*>& narrs, int32_t axis) {
  if (narrs.empty()) throw DimensionError("Cannot concatenate: nothing to concatenate");
  std::vector<Shape> shapes;
  for (const MultiIndexable<T>* a : narrs) shapes.push_back(a->shape());
  int32_t ax = 0;
  const Shape shape = concat_shape_of(shapes, axis, &ax);
  DeviceNArray<T> out(shape);
  int64_t at = 0;
  for (const MultiIndexable<T>* a : narrs) {
    const int64_t n = a->shape()[(size_t)ax];
    if (n && out.size()) {
      RegionLiteral lit(shape.size(), all);
      lit[(size_t)ax] = range(at, at + n - 1);
      out.unsafe_set_chunk(IndexRegion(lit, shape, false), *a);
    }
    at += n;
  }
  return out;
}
// push: the buffers are appended as they lie and ONLY shape[0] grows, whatever `axis` says (`axis` merely relaxes the
// compatibility test -- the reference's own TODO).  Arrays made by reshape before the push keep the old buffer.
template <class T>
DeviceNArray<T>& DeviceNArray<T>::push(const std::vector<const MultiIndexable<T>*>& others, int32_t axis) {
  if (others.empty()) return *this;
  std::vector<Shape> shapes{this->shape_};
  for (const MultiIndexable<T>* o : others) shapes.push_back(o->shape());
  int32_t ax = 0;
  concat_shape_of(shapes, axis, &ax);                                  // compatible? -> DimensionError
  int64_t total = this->size(), rows = this->shape_[0];
  for (const MultiIndexable<T>* o : others) { total += o->size(); rows += o->shape()[0]; }
  Shape new_shape = this->shape_;
  new_shape[0] = rows;
  if (shape_to_size(new_shape) != total)
    throw ShapeError("Cannot change shape from [" + std::to_string(total) + "] to " + shape_str(new_shape) + ": reshape cannot add or remove elements.");
  auto buf = std::make_shared<DeviceBuffer>((size_t)total * sizeof(T));
  int64_t at = 0;
  auto append = [&](const MultiIndexable<T>& a) {
    const DeviceNArray<T> flat = a.to_narr();
    if (flat.size()) Device::check(ph_d2d(static_cast<T*>(buf->ptr) + at, flat.data(), (size_t)flat.size() * sizeof(T)));
    at += flat.size();
  };
  append(*this);
  for (const MultiIndexable<T>* o : others) append(*o);
  this->buf_ = std::move(buf);
  this->shape_ = new_shape;
  this->desc_ = contiguous_desc(new_shape);
  return *this;
}
template <class T>
DeviceNArray<T> DeviceNArray<T>::wrap(const std::vector<const MultiIndexable<T>*>& narrs) {
  if (narrs.empty()) throw DimensionError("Cannot wrap these arrays: nothing to wrap");
  const Shape container = narrs[0]->shape();
  for (const MultiIndexable<T>* a : narrs)
    if (a->shape() != container)
      throw DimensionError("Cannot wrap these arrays: shapes do not match. Pass argument pad:true if you want to reshape arrays as necessary.");
  Shape row = container;
  row.insert(row.begin(), 1);
  std::vector<DeviceNArray<T>> rows;
  for (const MultiIndexable<T>* a : narrs) rows.push_back(a->to_narr().reshape(row));
  std::vector<const MultiIndexable<T>*> ptrs;
  for (const DeviceNArray<T>& r : rows) ptrs.push_back(&r);
  return concatenate(ptrs, 0);
}

// ---- the stencil of examples/heat_equation.cr:26-51 ---------------------------------------------
namespace Heat {
// update_temp: one explicit step; mode PH_HEAT_EXAMPLE1D reproduces the example's one-sided ends,
// PH_HEAT_FIXED is the N-D rule (boundary cells held).
template <class T>
DeviceNArray<T> update_temp(const DeviceNArray<T>& state, T coeff, int32_t mode = PH_HEAT_FIXED) {
  static_assert(std::is_floating_point<T>::value, "the stencil is defined for Float32 / Float64");
  DeviceNArray<T> out(state.shape());
  if (state.size())
    Device::check(ph_heat_step(DType<T>::value, state.dimensions(), state.shape().data(), &coeff, mode, state.data(), out.data()));
  return out;
}
// simulate: `steps` steps; the input array is left untouched
template <class T>
DeviceNArray<T> simulate(const DeviceNArray<T>& initial, T coeff, int64_t steps, int32_t mode = PH_HEAT_FIXED) {
  static_assert(std::is_floating_point<T>::value, "the stencil is defined for Float32 / Float64");
  DeviceNArray<T> a = initial.clone(), b(initial.shape());
  if (!initial.size() || steps <= 0) return a;
  int32_t final_is_b = 0;
  Device::check(ph_heat_run(DType<T>::value, a.dimensions(), a.shape().data(), &coeff, mode, a.data(), b.data(), steps, &final_is_b));
  return final_is_b ? b : a;
}
}  // namespace Heat

}  // namespace Phase
#endif  // PH_NARRAY_HPP
