// ph_sharded.hpp -- Phase::ShardedNArray<T>: an NArray distributed along axis 0 over the ranks of a
// one-process-per-GPU job, for a compiled host (the Python mirror is
// ph-core_b200/sharding.py, the Crystal one crystal/src/device/sharded_n_array.cr).
//
// The reference is single-process (no distributed layer); what BASELINE.json's north_star
// partitions is the seam this class covers (src/multi_indexable.cr:30-65 for the array it shards):
//   * every rank holds the contiguous row range ph_shard_range(shape[0], world, rank) as an ordinary
//     DeviceNArray<T>: elementwise operators, comparisons and masked stores are local launches;
//   * full reductions (Enumerable#sum/min/max, the argmax idiom README.md:56-61) are ONE launch per rank with
//     the cross-rank combine inside the kernel over peer-mapped slots (ph_reduce_full_sharded), the same
//     result -- and the same exceptions -- on every rank;
//   * per-axis reductions: local unless the reduced axis is the sharded one (then the [inner] partials are
//     combined: ph_allreduce, or for checked integer sums ph_allgather + the checked axis-0 fold);
//   * permute (a transpose across shards) is one pass of peer stores: the transpose kernel writes every block
//     straight into its owner's shard over NVLink (ph_alltoall_strided); without P2P the NCCL form
//     (gather, ph_alltoallv, scatter) runs.
// Header-only over include/ph_gpu.h + include/ph_host.h; host logic (who owns what) comes from ph_host.h's
// plans, so the three host layers cannot disagree.
#pragma once
#include "ph_narray.hpp"
#include <algorithm>
#include <limits>
#include <tuple>

namespace Phase {

namespace Comm {
inline int32_t& world_ref() { static int32_t w = 1; return w; }
inline int32_t& rank_ref() { static int32_t r = 0; return r; }
inline int32_t world() { return world_ref(); }
inline int32_t rank() { return rank_ref(); }
// rank 0 creates the id and hands it to the other ranks by whatever means the launcher has (file, MPI, env)
inline void unique_id(uint8_t out128[128]) { Device::ensure_init(); Device::check(ph_comm_unique_id(out128)); }
inline void init(int32_t nranks, int32_t rank, const uint8_t* id128) {
  Device::ensure_init();
  Device::check(ph_comm_init(nranks, rank, id128));
  world_ref() = nranks;
  rank_ref() = rank;
}
inline void destroy() { ph_comm_destroy(); world_ref() = 1; rank_ref() = 0; }
inline bool p2p_ready() { int32_t v = 0; Device::check(ph_comm_p2p_ready(&v)); return v != 0; }
inline std::pair<int64_t, int64_t> shard_range(int64_t n, int32_t w = world(), int32_t r = rank()) {
  int64_t a = 0, b = 0;
  Device::host_check(ph_shard_range(n, w, r, &a, &b));
  return {a, b};
}
// Peer-mapped device memory (ph_symm_alloc): allocation is COLLECTIVE (every rank, same order).  The block is
// released by Comm::destroy (ph_comm_destroy) -- never by a destructor, which would make the collective free
// depend on object lifetimes -- or explicitly with ph_symm_free on every rank.
inline std::shared_ptr<DeviceBuffer> symm_buffer(size_t nbytes) {
  void* p = nullptr;
  Device::check(ph_symm_alloc(nbytes ? nbytes : 1, &p));
  return DeviceBuffer::adopt(p, nbytes ? nbytes : 1);
}
}  // namespace Comm

template <class T>
class ShardedNArray {
 public:
  // `local` = this rank's rows [row0, row1) of the global array
  ShardedNArray(Shape global_shape, DeviceNArray<T> local) : shape_(std::move(global_shape)), local_(std::move(local)) {
    if (shape_.empty()) throw ShapeError("a sharded array needs at least one axis");
    std::tie(row0_, row1_) = Comm::shard_range(shape_[0]);
    Shape want = shape_;
    want[0] = row1_ - row0_;
    if (local_.shape() != want)
      throw ShapeError("local shard has shape " + shape_str(local_.shape()) + ", expected " + shape_str(want));
  }
  // every rank passes the same host array (lexicographic order) and keeps its own rows
  static ShardedNArray from_global(const Shape& shape, const std::vector<T>& host) {
    if ((int64_t)host.size() != shape_to_size(shape)) throw ShapeError("from_global: " + std::to_string(host.size()) + " elements for shape " + shape_str(shape));
    auto rr = Comm::shard_range(shape[0]);
    Shape mine = shape;
    mine[0] = rr.second - rr.first;
    const int64_t row = row_elems(shape);
    return ShardedNArray(shape, DeviceNArray<T>::from_host(mine, host.data() + rr.first * row));
  }

  const Shape& shape() const { return shape_; }
  int64_t size() const { return shape_to_size(shape_); }
  const DeviceNArray<T>& local() const { return local_; }
  DeviceNArray<T>& local() { return local_; }
  int64_t row0() const { return row0_; }
  int64_t row1() const { return row1_; }

  // the whole array on every rank's host (allgather of the shards, padded to the largest shard)
  std::vector<T> to_global() const {
    const int32_t w = Comm::world();
    std::vector<T> mine = local_.to_host();
    if (w == 1) return mine;
    const int64_t row = row_elems(shape_);
    const int64_t rows_max = (shape_[0] + w - 1) / w;
    const size_t slot = (size_t)(rows_max * row) * sizeof(T);
    DeviceBuffer send(slot), recv(slot * (size_t)w);
    if (!mine.empty()) Device::check(ph_d2d(send.ptr, local_.data(), mine.size() * sizeof(T)));
    Device::check(ph_allgather(send.ptr, recv.ptr, (int64_t)slot));
    std::vector<unsigned char> raw(slot * (size_t)w);
    Device::read_checked(raw.data(), recv.ptr, raw.size());
    std::vector<T> out((size_t)size());
    for (int32_t r = 0; r < w; r++) {
      auto rr = Comm::shard_range(shape_[0], w, r);
      const size_t n = (size_t)((rr.second - rr.first) * row);
      if (n) std::memcpy(out.data() + rr.first * row, raw.data() + (size_t)r * slot, n * sizeof(T));
    }
    return out;
  }

  // ---- elementwise / compare / masks: local, shapes checked on the GLOBAL shape ---------------
  ShardedNArray binary(int32_t op, const ShardedNArray& o, const char* name) const { same_shape(o); return wrap(local_.binary(op, o.local_, name)); }
  ShardedNArray scalar(int32_t op, T s, bool scalar_on_left = false) const { return wrap(local_.scalar(op, s, scalar_on_left)); }
  ShardedNArray<Bool> compare(int32_t cmp, const ShardedNArray& o) const { same_shape(o); return ShardedNArray<Bool>(shape_, local_.compare(cmp, o.local_)); }
  ShardedNArray<Bool> compare(int32_t cmp, T s) const { return ShardedNArray<Bool>(shape_, local_.compare(cmp, s)); }
  ShardedNArray<Bool> eq(const ShardedNArray& o) const { same_shape(o); return ShardedNArray<Bool>(shape_, local_.eq(o.local_)); }
  void set_mask(const ShardedNArray<Bool>& mask, type_identity_t<T> value) { local_.set_mask(mask.local(), value); }
  void set_mask(const ShardedNArray<Bool>& mask, const ShardedNArray& value) { same_shape(value); local_.set_mask(mask.local(), value.local_); }

  // ---- full reductions: one launch per rank, combined inside the kernel (collective) ------------
  T sum() const { return reduce_full(PH_SUM).first; }
  T min() const { return need(reduce_full(PH_MIN)).first; }
  T max() const { return need(reduce_full(PH_MAX)).first; }
  std::pair<T, Coord> argmax() const { auto r = need(reduce_full(PH_ARGMAX)); return {r.first, index_to_coord(r.second)}; }
  std::pair<T, Coord> argmin() const { auto r = need(reduce_full(PH_ARGMIN)); return {r.first, index_to_coord(r.second)}; }

  // ---- per-axis reductions ----------------------------------------------------------------------
  // axis >= 1: axis 0 survives, the result is still sharded
  ShardedNArray sum(int32_t axis) const { return kept(axis, local_.sum(check_axis(axis))); }
  ShardedNArray min(int32_t axis) const { return kept(axis, local_.min(check_axis(axis))); }
  ShardedNArray max(int32_t axis) const { return kept(axis, local_.max(check_axis(axis))); }
  // axis 0 is the sharded one: partial over my rows, combined across ranks; REPLICATED result of shape[1:]
  DeviceNArray<T> sum0() const { return over_shards(PH_SUM); }
  DeviceNArray<T> min0() const { return over_shards(PH_MIN); }
  DeviceNArray<T> max0() const { return over_shards(PH_MAX); }

  // ---- narr[region_literal] (gather, multi_indexable.cr:338-356) across shards -----------------------
  // A literal that leaves axis 0 whole is local.  Anything else re-splits the result over the ranks along ITS
  // axis 0 (ph_slice_plan_of): every (source, destination) block is ONE strided descriptor over the source's
  // rows and a contiguous row range of the destination -- stored straight into its owner over NVLink with P2P
  // (ph_alltoall_strided), gathered + ph_alltoallv (received in place) without.
  ShardedNArray get_chunk(const RegionLiteral& lits) const {
    const int32_t nd = (int32_t)shape_.size(), w = Comm::world(), me = Comm::rank();
    IndexRegion reg(lits, shape_, true);
    ph_slice_plan plan{};
    std::vector<ph_slice_peer> peers((size_t)std::max(1, w));
    Device::host_check(ph_slice_plan_of(shape_.data(), nd, &reg.r, w, me, &plan, peers.data()));
    Shape new_shape(plan.new_shape, plan.new_shape + plan.dims);
    if (plan.local) {
      if (row1_ <= row0_) { Shape e = new_shape; e[0] = 0; return ShardedNArray(new_shape, DeviceNArray<T>(e)); }
      RegionLiteral mine = lits;
      mine[0] = all;
      return ShardedNArray(new_shape, local_.get_chunk(mine));
    }
    Shape my_shape = new_shape;
    my_shape[0] = plan.my_new_rows[1] - plan.my_new_rows[0];
    if (w > 1 && Comm::p2p_ready()) {
      DeviceNArray<T> res(my_shape, Comm::symm_buffer((size_t)shape_to_size(my_shape) * sizeof(T)));   // collective
      std::vector<ph_desc> src((size_t)w), dst((size_t)w);
      for (int32_t q = 0; q < w; q++) { src[(size_t)q] = peers[(size_t)q].send; dst[(size_t)q] = peers[(size_t)q].land; }
      Device::check(ph_alltoall_strided((int32_t)sizeof(T), local_.data(), src.data(), res.data(), dst.data()));
      return ShardedNArray(new_shape, res);
    }
    DeviceNArray<T> res(my_shape);
    int64_t row = 1;
    for (size_t i = 1; i < new_shape.size(); i++) row *= new_shape[i];
    std::vector<DeviceNArray<T>> sends;
    std::vector<const void*> sp((size_t)w, nullptr);
    std::vector<void*> rp((size_t)w, nullptr);
    std::vector<int64_t> sb((size_t)w, 0), rb((size_t)w, 0);
    sends.reserve((size_t)w);
    for (int32_t q = 0; q < w; q++) {
      const ph_slice_peer& pq = peers[(size_t)q];
      int64_t n = 1;
      for (int i = 0; i < pq.send.rank; i++) n *= pq.send.extent[i];
      if (n > 0) {
        sends.push_back(DeviceView<T>(local_.buffer_owner(), pq.send, Shape(pq.send.extent, pq.send.extent + pq.send.rank)).to_narr());
        sp[(size_t)q] = sends.back().data();
        sb[(size_t)q] = n * (int64_t)sizeof(T);
      }
      if (pq.recv1 > pq.recv0 && row > 0) {
        rp[(size_t)q] = res.data() + (pq.recv0 - plan.my_new_rows[0]) * row;
        rb[(size_t)q] = (pq.recv1 - pq.recv0) * row * (int64_t)sizeof(T);
      }
    }
    Device::check(ph_alltoallv(sp.data(), sb.data(), rp.data(), rb.data()));
    Device::wait();                               // the gathered blocks are released when `sends` goes out of scope
    return ShardedNArray(new_shape, res);
  }
  ShardedNArray operator[](const RegionLiteral& lits) const { return get_chunk(lits); }

  // ---- narr[region_literal] = value (scatter / fill, multi_writable.cr:55-84) across shards -------------
  // A scalar fills this rank's cells of the region (no exchange).  An array source of the region's shape is the
  // gather run backwards with the same plan: the rows of `src` this rank holds leave as contiguous blocks
  // (ph_alltoallv), every block received is scattered into the arithmetic progression of local rows it belongs to.
  void set_chunk(const RegionLiteral& lits, type_identity_t<T> value) {
    ph_slice_plan plan{};
    std::vector<ph_slice_peer> peers = slice_plan(lits, &plan);
    if (plan.local) {
      if (row1_ > row0_) { RegionLiteral mine = lits; mine[0] = all; local_.set_chunk(mine, value); }
      return;
    }
    for (const ph_slice_peer& pq : peers)
      if (desc_elems(pq.send) > 0) Device::check(ph_fill_region((int32_t)sizeof(T), local_.data(), &pq.send, &value));
  }
  void set_chunk(const RegionLiteral& lits, const ShardedNArray& src) {
    ph_slice_plan plan{};
    std::vector<ph_slice_peer> peers = slice_plan(lits, &plan);
    Shape new_shape(plan.new_shape, plan.new_shape + plan.dims);
    if (src.shape_ != new_shape)   // multi_writable.cr:58-60
      throw ShapeError("Cannot substitute: the given array has shape " + shape_str(src.shape_) + ", but the region has shape " + shape_str(new_shape) + ".");
    if (plan.local) {
      if (row1_ > row0_) { RegionLiteral mine = lits; mine[0] = all; local_.set_chunk(mine, src.local_); }
      return;
    }
    const int32_t w = Comm::world();
    int64_t row = 1;
    for (size_t i = 1; i < new_shape.size(); i++) row *= new_shape[i];
    std::vector<DeviceNArray<T>> temps;
    std::vector<int32_t> from;
    std::vector<const void*> sp((size_t)w, nullptr);
    std::vector<void*> rp((size_t)w, nullptr);
    std::vector<int64_t> sb((size_t)w, 0), rb((size_t)w, 0);
    temps.reserve((size_t)w);
    for (int32_t q = 0; q < w; q++) {
      const ph_slice_peer& pq = peers[(size_t)q];
      if (pq.recv1 > pq.recv0 && row > 0) {                      // rows of `src` I hold that q's shard receives
        sp[(size_t)q] = src.local_.data() + (pq.recv0 - plan.my_new_rows[0]) * row;
        sb[(size_t)q] = (pq.recv1 - pq.recv0) * row * (int64_t)sizeof(T);
      }
      if (desc_elems(pq.send) > 0) {                              // where q's rows land in MY shard
        temps.push_back(DeviceNArray<T>(Shape(pq.send.extent, pq.send.extent + pq.send.rank)));
        from.push_back(q);
        rp[(size_t)q] = temps.back().data();
        rb[(size_t)q] = desc_elems(pq.send) * (int64_t)sizeof(T);
      }
    }
    Device::check(ph_alltoallv(sp.data(), sb.data(), rp.data(), rb.data()));
    for (size_t i = 0; i < temps.size(); i++)
      Device::check(ph_copy_strided((int32_t)sizeof(T), temps[i].data(), &temps[i].desc(), local_.data(), &peers[(size_t)from[i]].send));
    Device::wait();                               // the temporaries die with this scope
  }

  // ---- permute (MultiIndexable#permute multi_indexable.cr:795-803; default = reversed axes) ------
  // `reuse`: an earlier P2P result of the same shape whose (peer-mapped) storage receives the new result
  ShardedNArray permute(std::vector<int32_t> pattern = {}, const ShardedNArray* reuse = nullptr) const {
    const int32_t nd = (int32_t)shape_.size(), w = Comm::world(), me = Comm::rank();
    if (pattern.empty()) for (int32_t i = nd - 1; i >= 0; i--) pattern.push_back(i);
    ph_transpose_plan plan{};
    std::vector<ph_transpose_peer> peers((size_t)std::max(1, w));
    Device::host_check(ph_transpose_plan_of(shape_.data(), nd, pattern.data(), w, me, &plan, peers.data()));
    Shape new_shape(plan.new_shape, plan.new_shape + nd);
    if (plan.local) return ShardedNArray(new_shape, local_.permute(pattern));       // axis 0 stays put: no exchange
    Shape my_shape = new_shape;
    my_shape[0] = plan.my_new_rows[1] - plan.my_new_rows[0];
    auto block_view = [&](int32_t q) {                                              // my rows x q's slice of old axis k, in q's axis order
      RegionLiteral lit((size_t)nd, all);
      lit[(size_t)plan.k] = range_ex(peers[(size_t)q].send0, peers[(size_t)q].send1);
      return local_.view(lit, false).permute(pattern);
    };
    if (w > 1 && Comm::p2p_ready()) {
      DeviceNArray<T> res = reuse ? reuse->local_
                                  : DeviceNArray<T>(my_shape, Comm::symm_buffer((size_t)shape_to_size(my_shape) * sizeof(T)));   // collective
      if (reuse && reuse->shape_ != new_shape) throw ShapeError("permute(reuse): shape " + shape_str(reuse->shape_) + " is not " + shape_str(new_shape));
      std::vector<ph_desc> src((size_t)w), dst((size_t)w);
      for (int32_t q = 0; q < w; q++) {
        const ph_transpose_peer& pq = peers[(size_t)q];
        ph_desc zero{};
        zero.rank = nd;
        src[(size_t)q] = dst[(size_t)q] = zero;
        if (pq.send1 - pq.send0 <= 0 || row1_ - row0_ <= 0) continue;
        src[(size_t)q] = block_view(q).desc();
        auto qr = Comm::shard_range(new_shape[0], w, q);                            // q's rows of the result
        Shape q_shape = new_shape;
        q_shape[0] = qr.second - qr.first;
        ph_desc d = contiguous_desc(q_shape);
        d.extent[plan.j] = row1_ - row0_;                                           // old axis 0 (my rows) lands on new axis j
        d.offset = row0_ * d.stride[plan.j];
        dst[(size_t)q] = d;
      }
      Device::check(ph_alltoall_strided((int32_t)sizeof(T), local_.data(), src.data(), res.data(), dst.data()));
      return ShardedNArray(new_shape, res);
    }
    // NCCL form: permuting gathers, personalised all-to-all, scatters
    DeviceNArray<T> out(my_shape);
    std::vector<DeviceNArray<T>> sends, recvs;
    std::vector<const void*> sp((size_t)w, nullptr);
    std::vector<void*> rp((size_t)w, nullptr);
    std::vector<int64_t> sb((size_t)w, 0), rb((size_t)w, 0);
    sends.reserve((size_t)w); recvs.reserve((size_t)w);
    auto landing = [&](int32_t q) {
      RegionLiteral lit((size_t)nd, all);
      lit[(size_t)plan.j] = range_ex(peers[(size_t)q].recv0, peers[(size_t)q].recv1);
      return lit;
    };
    for (int32_t q = 0; q < w; q++) {
      const ph_transpose_peer& pq = peers[(size_t)q];
      const int64_t n_send = count(pq.send_shape, nd), n_recv = count(pq.recv_shape, nd);
      if (q == me) {                                // my own block never leaves the GPU: one permuting copy into the result
        if (n_send) out.set_chunk(landing(q), block_view(q));
        continue;
      }
      if (n_send) {
        sends.push_back(block_view(q).to_narr());
        sp[(size_t)q] = sends.back().data();
        sb[(size_t)q] = n_send * (int64_t)sizeof(T);
      }
      if (n_recv) {
        recvs.push_back(DeviceNArray<T>(Shape(pq.recv_shape, pq.recv_shape + nd)));
        rp[(size_t)q] = recvs.back().data();
        rb[(size_t)q] = n_recv * (int64_t)sizeof(T);
      }
    }
    Device::check(ph_alltoallv(sp.data(), sb.data(), rp.data(), rb.data()));
    size_t ri = 0;
    for (int32_t q = 0; q < w; q++) {
      if (q == me || rb[(size_t)q] == 0) continue;
      const DeviceNArray<T>& blk = recvs[ri++];
      out.set_chunk(landing(q), blk);
    }
    return ShardedNArray(new_shape, out);
  }

  Coord index_to_coord(int64_t index) const {       // Buffered.index_to_coord (buffered.cr:58-77) on the GLOBAL shape
    Coord c(shape_.size());
    for (size_t i = shape_.size(); i-- > 0;) { c[i] = index % shape_[i]; index /= shape_[i]; }
    return c;
  }

 private:
  Shape shape_;
  DeviceNArray<T> local_;
  int64_t row0_ = 0, row1_ = 0;
  template <class U> friend class ShardedNArray;

  static int64_t desc_elems(const ph_desc& d) { int64_t n = 1; for (int i = 0; i < d.rank; i++) n *= d.extent[i]; return d.rank ? n : 0; }
  std::vector<ph_slice_peer> slice_plan(const RegionLiteral& lits, ph_slice_plan* plan) const {
    IndexRegion reg(lits, shape_, true);
    std::vector<ph_slice_peer> peers((size_t)std::max(1, Comm::world()));
    Device::host_check(ph_slice_plan_of(shape_.data(), (int32_t)shape_.size(), &reg.r, Comm::world(), Comm::rank(), plan, peers.data()));
    return peers;
  }
  static int64_t row_elems(const Shape& s) { int64_t n = 1; for (size_t i = 1; i < s.size(); i++) n *= s[i]; return n; }
  static int64_t count(const int64_t* ext, int32_t nd) { int64_t n = 1; for (int32_t i = 0; i < nd; i++) n *= ext[i]; return n; }
  ShardedNArray wrap(DeviceNArray<T> l) const { return ShardedNArray(shape_, std::move(l)); }
  void same_shape(const ShardedNArray& o) const {
    if (o.shape_ != shape_)
      throw ShapeError("The shape of this MultiIndexable (" + shape_str(shape_) + ") does not match the shape of the one provided (" + shape_str(o.shape_) + ").");
  }
  int32_t check_axis(int32_t axis) const {
    if (axis < 0 || axis >= (int32_t)shape_.size())
      throw IndexError("axis " + std::to_string(axis) + " is not present in a " + std::to_string(shape_.size()) + "-dimensional MultiIndexable");
    if (axis == 0) throw IndexError("axis 0 is the sharded axis: use sum0 / min0 / max0 (replicated result)");
    return axis;
  }
  ShardedNArray kept(int32_t axis, DeviceNArray<T> part) const {
    Shape s;
    for (int32_t i = 0; i < (int32_t)shape_.size(); i++) if (i != axis) s.push_back(shape_[(size_t)i]);
    return ShardedNArray(s, std::move(part));
  }
  std::pair<T, int64_t> reduce_full(int32_t red) const {
    alignas(16) unsigned char cell[16] = {0};
    int64_t idx = -1;
    uint32_t flags = 0;
    Device::check(ph_reduce_full_sharded(red, DType<T>::value, local_.data(), &local_.desc(), row0_ * row_elems(shape_), cell, &idx, &flags));
    Device::raise_for(flags);
    T val{};
    std::memcpy(&val, cell, sizeof(T));
    return {val, idx};
  }
  static std::pair<T, int64_t> need(std::pair<T, int64_t> r) {
    if (r.second < 0) throw EmptyError("Empty enumerable");
    return r;
  }
  DeviceNArray<T> over_shards(int32_t red) const {
    // every decision that can raise is taken on the GLOBAL shape, so all ranks reach the collective (or none does)
    if (shape_[0] == 0 && red != PH_SUM) throw EmptyError("Empty enumerable");
    Shape out_shape(shape_.begin() + 1, shape_.end());
    if (out_shape.empty()) out_shape.push_back(1);
    DeviceNArray<T> part = row1_ > row0_ ? (red == PH_SUM ? local_.sum(0) : red == PH_MAX ? local_.max(0) : local_.min(0))
                                         : DeviceNArray<T>::fill(out_shape, identity(red));     // an empty shard contributes the identity
    const int32_t w = Comm::world();
    if (w == 1) return part;
    if (red == PH_SUM && std::is_integral<T>::value && !Comm::p2p_ready()) {   // (P2P: ph_allreduce folds in rank order, checked)
      // integer sums are overflow-CHECKED: an ncclSum would wrap silently.  The per-rank partials ([world, inner],
      // rank order = row order) are gathered and folded by the checked axis-0 sum.
      Shape gs = out_shape;
      gs.insert(gs.begin(), (int64_t)w);
      DeviceNArray<T> gathered(gs);
      Device::check(ph_allgather(part.data(), gathered.data(), part.size() * (int64_t)sizeof(T)));
      return gathered.sum(0);
    }
    Device::check(ph_allreduce(red, DType<T>::value, part.data(), part.size()));
    Device::raise_pending();                      // per-axis folds are raise points (reduce_axis): the cross-rank fold too
    return part;
  }
  static T identity(int32_t red) {
    if (red == PH_SUM) return T(0);
    if (std::is_floating_point<T>::value) return red == PH_MAX ? -std::numeric_limits<T>::infinity() : std::numeric_limits<T>::infinity();
    return red == PH_MAX ? std::numeric_limits<T>::lowest() : std::numeric_limits<T>::max();
  }
};

#define PH_SHARDED_ARITH(sym, code)                                                                                              \
  template <class T> ShardedNArray<T> operator sym(const ShardedNArray<T>& a, const ShardedNArray<T>& b) { return a.binary(code, b, #sym); } \
  template <class T> ShardedNArray<T> operator sym(const ShardedNArray<T>& a, type_identity_t<T> s) { return a.scalar(code, s, false); }       \
  template <class T> ShardedNArray<T> operator sym(type_identity_t<T> s, const ShardedNArray<T>& a) { return a.scalar(code, s, true); }
PH_SHARDED_ARITH(+, PH_ADD)
PH_SHARDED_ARITH(-, PH_SUB)
PH_SHARDED_ARITH(*, PH_MUL)
#undef PH_SHARDED_ARITH
#define PH_SHARDED_CMP(sym, code)                                                                                               \
  template <class T> ShardedNArray<Bool> operator sym(const ShardedNArray<T>& a, const ShardedNArray<T>& b) { return a.compare(code, b); } \
  template <class T> ShardedNArray<Bool> operator sym(const ShardedNArray<T>& a, type_identity_t<T> s) { return a.compare(code, s); }
PH_SHARDED_CMP(>, PH_GT)
PH_SHARDED_CMP(<, PH_LT)
PH_SHARDED_CMP(>=, PH_GE)
PH_SHARDED_CMP(<=, PH_LE)
#undef PH_SHARDED_CMP

}  // namespace Phase
