// ph_pipeline.hpp -- streams, pinned host arrays, asynchronous transfers and the chunked
// host -> device -> host row pipeline for the C++ host layer (twin of ph-core_b200/pipeline.py and of
// crystal/src/device/pipeline.cr).
//
// The reference keeps every NArray in host memory (n_array.cr:20, 58); a device-backed array adds the two
// explicit transfers.  ONE elementwise expression over host-resident operands is bound
// by the host link, not by any kernel: RowPipeline cuts the leading axis into chunks and runs them through an
// upload stream, a compute stream (the caller's expression, written with the ordinary operators) and a
// download stream, so the upload of chunk i+1, the kernels of chunk i and the download of chunk i-1 overlap.
// Everything goes through the stream entry points of ph_gpu.h (ph_stream_create / ph_set_stream /
// ph_stream_wait / ph_free_on); there is no arithmetic and no descriptor construction here.
#ifndef PH_PIPELINE_HPP
#define PH_PIPELINE_HPP

#include <functional>

#include "ph_narray.hpp"

namespace Phase {

// A CUDA stream of the library.  `StreamScope on(s);` makes every array operation of the enclosing block
// launch on it; `s.wait(other)` orders it behind what `other` (nullptr = the library's own stream) has queued.
class Stream {
 public:
  Stream() { Device::ensure_init(); Device::check(ph_stream_create(&h_)); }
  ~Stream() { if (h_) ph_stream_destroy(h_); }
  Stream(const Stream&) = delete;
  Stream& operator=(const Stream&) = delete;
  void* handle() const { return h_; }
  void wait(const Stream* other = nullptr) const { Device::check(ph_stream_wait(h_, other ? other->h_ : nullptr)); }
  void synchronize() const { Device::check(ph_stream_sync(h_)); }

 private:
  void* h_ = nullptr;
};
inline void main_stream_wait(const Stream& s) { Device::check(ph_stream_wait(nullptr, s.handle())); }

class StreamScope {
 public:
  explicit StreamScope(const Stream& s) : saved_(ph_stream()) { Device::check(ph_set_stream(s.handle())); }
  ~StreamScope() { ph_set_stream(saved_); }
  StreamScope(const StreamScope&) = delete;
  StreamScope& operator=(const StreamScope&) = delete;

 private:
  void* saved_;
};

// A row-major host array in PINNED memory (ph_host_alloc): the only kind of host memory an asynchronous
// transfer may use.  rows(r0, r1) is the address of a block of leading-axis rows (no copy).
template <class T>
class PinnedArray {
 public:
  explicit PinnedArray(const Shape& shape) : shape_(shape) {
    void* p = nullptr;
    Device::check(ph_host_alloc((size_t)std::max<int64_t>(1, shape_to_size(shape)) * sizeof(T), &p));
    ptr_ = static_cast<T*>(p);
  }
  PinnedArray(const Shape& shape, const std::vector<T>& values) : PinnedArray(shape) {
    if ((int64_t)values.size() != shape_to_size(shape)) throw ShapeError("PinnedArray: " + std::to_string(values.size()) + " elements for shape " + shape_str(shape));
    std::memcpy(ptr_, values.data(), values.size() * sizeof(T));
  }
  ~PinnedArray() { if (ptr_) ph_host_free(ptr_); }
  PinnedArray(const PinnedArray&) = delete;
  PinnedArray& operator=(const PinnedArray&) = delete;
  const Shape& shape() const { return shape_; }
  int64_t size() const { return shape_to_size(shape_); }
  int64_t row_elems() const { return shape_.empty() || shape_[0] == 0 ? 0 : size() / shape_[0]; }
  T* data() { return ptr_; }
  const T* data() const { return ptr_; }
  T& operator[](int64_t i) { return ptr_[i]; }
  const T& operator[](int64_t i) const { return ptr_[i]; }
  T* rows(int64_t r0) { return ptr_ + r0 * row_elems(); }
  const T* rows(int64_t r0) const { return ptr_ + r0 * row_elems(); }
  Shape rows_shape(int64_t r0, int64_t r1) const { Shape s = shape_; s[0] = r1 - r0; return s; }
  std::vector<T> to_vector() const { return std::vector<T>(ptr_, ptr_ + size()); }

 private:
  Shape shape_;
  T* ptr_ = nullptr;
};

// Host -> device from pinned memory: returns at once, the copy is ordered on the current stream like an operator.
template <class T>
DeviceNArray<T> from_host_async(const Shape& shape, const T* pinned) {
  DeviceNArray<T> out(shape);
  if (out.size()) Device::check(ph_h2d(out.data(), pinned, (size_t)out.size() * sizeof(T)));
  return out;
}
// Device -> host into pinned memory, asynchronous: valid after the next Device::sync() (the raise point for
// data-dependent errors) or Stream::synchronize().  `src` may be a temporary: its block is released on the stream
// it was allocated on, which is made to wait for this copy when that is not the current stream.
template <class T>
void to_host_async(const DeviceNArray<T>& a, T* pinned_dst) {
  if (a.size()) Device::check(ph_d2h_async(pinned_dst, a.data(), (size_t)a.size() * sizeof(T)));
  void* home = a.buffer_owner()->home_stream();
  void* cur = ph_stream();
  if (home != cur) Device::check(ph_stream_wait(home, cur));
}
template <class T>
void to_host_async(const DeviceView<T>& v, T* pinned_dst) { to_host_async<T>(v.to_narr(), pinned_dst); }   // one gather first

// [r0, r1) row ranges: `chunks` equal chunks; with taper = t the LAST one is cut again into halves t times
// (per/2, per/4, ..., per/2^t, per/2^t).  What is left when the last upload ends is one chunk's kernels and
// download -- nothing overlaps that tail -- so the final chunks are small while the early ones stay large
// (every copy pays a fixed set-up: 16 equal chunks 120.4 GB/s on the BASELINE config, 4 + 7 halvings 125.1).
inline std::vector<std::pair<int64_t, int64_t>> row_chunks(int64_t n, int64_t chunks, int taper = 0, int ramp = 0) {
  // ramp = t: the mirror image at the FRONT (per/2^t, per/2^t, ..., per/2): the first download can only start after
  // the first chunk, so that one is small too (no cost on one GPU -- 10.73 vs 10.77 ms -- and the downloads start
  // 32 times earlier when the host side is the slower half).  The schedule itself is ph_row_chunks (ph_host.h).
  const int32_t cap = (int32_t)std::max<int64_t>(1, chunks) + std::max(0, taper) + std::max(0, ramp) + 2;
  std::vector<int64_t> bounds((size_t)2 * cap);
  int32_t count = 0;
  Device::host_check(ph_row_chunks(n, chunks, taper, ramp, bounds.data(), cap, &count));
  std::vector<std::pair<int64_t, int64_t>> b;
  for (int32_t i = 0; i < count; i++) b.push_back({bounds[(size_t)2 * i], bounds[(size_t)2 * i + 1]});
  return b;
}

// Three streams by ROLE, reused across calls: one uploads, one computes, one downloads.  Uploads of all chunks
// are queued back to back; chunk k's operators wait for its upload, its download for its operators.
class RowPipeline {
 public:
  explicit RowPipeline(int64_t chunks = 4, int taper = 7, int ramp = 5) : chunks_(chunks), taper_(taper), ramp_(ramp) {}

  // out[r0:r1] = fn(rows[k][r0:r1] as device arrays, shared as device arrays) for every row chunk.  `rows` and
  // `out` are pinned host arrays with the same leading extent; `shared` operands (a broadcast row vector) are
  // uploaded once.  wait = true: returns when `out` is complete and raises pending data-dependent errors.
  template <class T>
  void map_rows(const std::function<DeviceNArray<T>(const std::vector<DeviceNArray<T>>&, const std::vector<DeviceNArray<T>>&)>& fn,
                const std::vector<const PinnedArray<T>*>& rows, PinnedArray<T>& out,
                const std::vector<const PinnedArray<T>*>& shared = {}, bool wait = true) {
    const int64_t n = out.shape().empty() ? 0 : out.shape()[0];
    for (const PinnedArray<T>* r : rows)
      if (r->shape().empty() || r->shape()[0] != n) throw ShapeError("map_rows: every row operand needs the leading extent of the output");
    up_.wait(); comp_.wait(); down_.wait();                   // behind whatever the main stream has queued
    std::vector<DeviceNArray<T>> shared_dev;
    {
      StreamScope on(up_);
      for (const PinnedArray<T>* x : shared) shared_dev.push_back(from_host_async<T>(x->shape(), x->data()));
    }
    std::vector<std::vector<DeviceNArray<T>>> keep_in;
    std::vector<DeviceNArray<T>> keep_out;
    try {
      for (const auto& ch : row_chunks(n, chunks_, taper_, ramp_)) {
        const int64_t r0 = ch.first, r1 = ch.second;
        std::vector<DeviceNArray<T>> ins;
        {
          StreamScope on(up_);
          for (const PinnedArray<T>* r : rows) ins.push_back(from_host_async<T>(r->rows_shape(r0, r1), r->rows(r0)));
        }
        comp_.wait(&up_);                                       // chunk k's operands (and the shared ones) have landed
        DeviceNArray<T> res = [&] { StreamScope on(comp_); return fn(ins, shared_dev); }();
        if (res.shape() != out.rows_shape(r0, r1)) throw ShapeError("map_rows: the expression returned shape " + shape_str(res.shape()) + " for a chunk of shape " + shape_str(out.rows_shape(r0, r1)));
        down_.wait(&comp_);
        {
          StreamScope on(down_);
          to_host_async<T>(res, out.rows(r0));
        }
        keep_in.push_back(std::move(ins));
        keep_out.push_back(std::move(res));
      }
    } catch (...) {
      // the caller's expression threw half way: chunks already queued still read and write the temporaries, so
      // nothing is released before the three streams have drained
      ph_stream_sync(up_.handle()); ph_stream_sync(comp_.handle()); ph_stream_sync(down_.handle());
      throw;
    }
    // device temporaries are released on the streams they were allocated on: order each of those behind every
    // consumer before letting go, then join the main stream
    up_.wait(&comp_); up_.wait(&down_); comp_.wait(&down_);
    main_stream_wait(up_); main_stream_wait(comp_); main_stream_wait(down_);
    keep_in.clear(); keep_out.clear(); shared_dev.clear();
    if (wait) Device::sync();
  }

 private:
  int64_t chunks_;
  int taper_, ramp_;
  Stream up_, comp_, down_;
};

}  // namespace Phase
#endif  // PH_PIPELINE_HPP
