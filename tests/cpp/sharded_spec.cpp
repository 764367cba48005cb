// sharded_spec.cpp -- Phase::ShardedNArray (include/ph_sharded.hpp) checked against the undivided array,
// one process per GPU.  Product-side host code + libphgpu only: expectations are computed here on the host
// with plain loops over the GLOBAL array every rank builds from the same seed.
//
// Launch (N ranks, one GPU each; rank 0 writes the NCCL id to PH_ID_FILE, the others wait for it):
//   for r in 0 1; do RANK=$r WORLD_SIZE=2 PH_ID_FILE=/tmp/ph_id tests/cpp/sharded_spec & done; wait
// A single process (no env) runs the same checks with world = 1.  Exit 0 = all passed.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ph_sharded.hpp"

using namespace Phase;
template <class T> using V = std::vector<T>;

static int g_failed = 0, g_passed = 0, g_rank = 0;
static std::string g_current;
#define EXPECT(...)                                                                              \
  do {                                                                                           \
    if (!(__VA_ARGS__)) { std::printf("  [rank %d] FAIL %s:%d  %s   [%s]\n", g_rank, __FILE__, __LINE__, #__VA_ARGS__, g_current.c_str()); g_failed++; } \
    else g_passed++;                                                                             \
  } while (0)
#define EXPECT_RAISES(ExcType, ...)                                                              \
  do {                                                                                           \
    bool raised_ = false;                                                                        \
    try { __VA_ARGS__; } catch (const ExcType&) { raised_ = true; }                              \
    if (!raised_) { std::printf("  [rank %d] FAIL %s:%d  expected %s from: %s   [%s]\n", g_rank, __FILE__, __LINE__, #ExcType, #__VA_ARGS__, g_current.c_str()); g_failed++; } \
    else g_passed++;                                                                             \
  } while (0)
static void it(const char* name, const std::function<void()>& body) {
  g_current = name;
  const int before = g_failed;
  try { body(); }
  catch (const std::exception& e) { std::printf("  [rank %d] FAIL unexpected exception in '%s': %s\n", g_rank, name, e.what()); g_failed++; }
  if (g_rank == 0) std::printf("%s %s\n", g_failed == before ? "ok  " : "FAIL", name);
}

// small LCG: the same global array on every rank
static uint32_t g_state = 12345;
static int32_t next_int(int32_t lo, int32_t hi) {
  g_state = g_state * 1664525u + 1013904223u;
  return lo + (int32_t)((g_state >> 8) % (uint32_t)(hi - lo + 1));
}
template <class T> static V<T> ints(int64_t n, int32_t lo = -8, int32_t hi = 8) {
  V<T> v((size_t)n);
  for (T& x : v) x = (T)next_int(lo, hi);
  return v;
}
// host permute: out[pattern-permuted coord] = in[coord]
template <class T> static V<T> host_permute(const V<T>& in, const Shape& shape, const std::vector<int32_t>& pat, Shape* out_shape) {
  const size_t nd = shape.size();
  Shape ns(nd);
  for (size_t i = 0; i < nd; i++) ns[i] = shape[(size_t)pat[i]];
  V<T> out(in.size());
  Coord c(nd, 0);
  for (size_t flat = 0; flat < in.size(); flat++) {
    int64_t o = 0;
    for (size_t i = 0; i < nd; i++) o = o * ns[i] + c[(size_t)pat[i]];
    out[(size_t)o] = in[flat];
    for (size_t i = nd; i-- > 0;) { if (++c[i] < shape[i]) break; c[i] = 0; }
  }
  *out_shape = ns;
  return out;
}

int main() {
  const char* wr = std::getenv("WORLD_SIZE");
  const int world = wr ? std::atoi(wr) : 1;
  g_rank = std::getenv("RANK") ? std::atoi(std::getenv("RANK")) : 0;
  const char* lr = std::getenv("LOCAL_RANK");
  Device::init(lr ? std::atoi(lr) : g_rank);
  uint8_t id[128] = {0};
  if (world > 1) {
    const char* path = std::getenv("PH_ID_FILE");
    if (!path) { std::printf("PH_ID_FILE is not set\n"); return 2; }
    const std::string tmp = std::string(path) + ".tmp";
    if (g_rank == 0) {
      Comm::unique_id(id);
      FILE* f = std::fopen(tmp.c_str(), "wb");
      std::fwrite(id, 1, 128, f);
      std::fclose(f);
      std::rename(tmp.c_str(), path);
    } else {
      FILE* f = nullptr;
      for (int tries = 0; tries < 600 && !(f = std::fopen(path, "rb")); tries++) std::this_thread::sleep_for(std::chrono::milliseconds(100));
      if (!f || std::fread(id, 1, 128, f) != 128) { std::printf("rank %d: no id file\n", g_rank); return 2; }
      std::fclose(f);
    }
  }
  Comm::init(world, g_rank, id);
  if (g_rank == 0) std::printf("sharded_spec: world=%d p2p=%d\n", world, (int)Comm::p2p_ready());

  const Shape gs = {6 * world + 2, 12, 10};
  const int64_t n = shape_to_size(gs);
  const V<float> g = ints<float>(n), h = ints<float>(n);
  auto sg = ShardedNArray<float>::from_global(gs, g);
  auto sh = ShardedNArray<float>::from_global(gs, h);

  it("elementwise and comparisons are local and equal the undivided array", [&] {
    V<float> want((size_t)n);
    for (size_t i = 0; i < (size_t)n; i++) want[i] = (g[i] * h[i] + g[i]) - 2.0f;
    EXPECT(((sg * sh + sg) - 2.0f).to_global() == want);
    V<Bool> wm((size_t)n);
    for (size_t i = 0; i < (size_t)n; i++) wm[i] = g[i] > h[i];
    EXPECT((sg > sh).to_global() == wm);
    EXPECT_RAISES(ShapeError, sg + ShardedNArray<float>::from_global({2 * world, 3}, ints<float>(6 * world)));
  });
  it("full reductions: one launch per rank, the same result on every rank (README.md:56-61 across shards)", [&] {
    double s = 0;
    float mx = g[0], mn = g[0];
    int64_t amax = 0, amin = 0;
    for (int64_t i = 0; i < n; i++) {
      s += g[(size_t)i];
      if (g[(size_t)i] > mx) { mx = g[(size_t)i]; amax = i; }
      if (g[(size_t)i] < mn) { mn = g[(size_t)i]; amin = i; }
    }
    EXPECT(sg.sum() == (float)s);
    EXPECT(sg.max() == mx && sg.min() == mn);
    auto am = sg.argmax();
    EXPECT(am.first == mx && am.second == sg.index_to_coord(amax));
    auto an = sg.argmin();
    EXPECT(an.first == mn && an.second == sg.index_to_coord(amin));
  });
  it("integer sums are overflow-checked over the GLOBAL lexicographic fold; NaN under max raises everywhere", [&] {
    V<int32_t> big((size_t)(4 * world), 0);
    big[0] = std::numeric_limits<int32_t>::max();
    big.back() = 1;                                   // the overflowing prefix ends on the LAST rank
    big[1] = -5;
    auto sb = ShardedNArray<int32_t>::from_global({4 * world}, big);
    EXPECT(sb.sum() == std::numeric_limits<int32_t>::max() - 4);
    big[1] = 0; big[2] = 1;
    EXPECT_RAISES(OverflowError, ShardedNArray<int32_t>::from_global({4 * world}, big).sum());
    V<double> nn((size_t)(3 * world), 1.0);
    nn.back() = std::numeric_limits<double>::quiet_NaN();
    EXPECT_RAISES(ArgumentError, ShardedNArray<double>::from_global({3 * world}, nn).max());
    EXPECT(ShardedNArray<double>::from_global({3 * world}, V<double>((size_t)(3 * world), 2.0)).sum() == 6.0 * world);
  });
  it("per-axis reductions: sharded axis combined across ranks, other axes local", [&] {
    const int64_t inner = gs[1] * gs[2];
    V<float> s0((size_t)inner, 0.0f), m2((size_t)(gs[0] * gs[1]));
    for (int64_t r = 0; r < gs[0]; r++)
      for (int64_t c = 0; c < inner; c++) s0[(size_t)c] = s0[(size_t)c] + g[(size_t)(r * inner + c)];
    EXPECT(sg.sum0().to_host() == s0);
    for (int64_t r = 0; r < gs[0] * gs[1]; r++) {
      float m = g[(size_t)(r * gs[2])];
      for (int64_t c = 1; c < gs[2]; c++) m = g[(size_t)(r * gs[2] + c)] > m ? g[(size_t)(r * gs[2] + c)] : m;
      m2[(size_t)r] = m;
    }
    auto mx2 = sg.max(2);
    EXPECT(mx2.shape() == Shape({gs[0], gs[1]}) && mx2.to_global() == m2);
    const V<int32_t> gi = ints<int32_t>(n);
    V<int32_t> si((size_t)inner, 0);
    for (int64_t r = 0; r < gs[0]; r++)
      for (int64_t c = 0; c < inner; c++) si[(size_t)c] += gi[(size_t)(r * inner + c)];
    EXPECT(ShardedNArray<int32_t>::from_global(gs, gi).sum0().to_host() == si);
    EXPECT_RAISES(IndexError, sg.sum(0));
    EXPECT_RAISES(IndexError, sg.max(3));
  });
  it("permute across shards (multi_indexable.cr:795-803): default, [1,0,2], local [0,2,1]; twice = identity", [&] {
    Shape ns;
    EXPECT(sg.permute().to_global() == host_permute(g, gs, {2, 1, 0}, &ns));
    EXPECT(sg.permute().shape() == ns);
    EXPECT(sg.permute({1, 0, 2}).to_global() == host_permute(g, gs, {1, 0, 2}, &ns));
    EXPECT(sg.permute({0, 2, 1}).to_global() == host_permute(g, gs, {0, 2, 1}, &ns));
    EXPECT_RAISES(IndexError, sg.permute({0, 0, 1}));
    const Shape ms = {1000 * world + 3, 517};
    const V<double> m2 = ints<double>(shape_to_size(ms), -1000, 1000);
    auto t2 = ShardedNArray<double>::from_global(ms, m2).permute();
    EXPECT(t2.shape() == Shape({517, 1000 * world + 3}) && t2.to_global() == host_permute(m2, ms, {1, 0}, &ns));
    EXPECT(t2.permute().to_global() == m2);
    if (Comm::world() > 1 && Comm::p2p_ready()) {      // the P2P form writes into an earlier result on request
      V<double> m3 = m2;
      for (double& x : m3) x += 1.0;
      auto again = ShardedNArray<double>::from_global(ms, m3).permute({}, &t2);
      EXPECT(again.local().data() == t2.local().data() && t2.to_global() == host_permute(m3, ms, {1, 0}, &ns));
    }
    const Shape os = {3 * world, 1};                   // the result has ONE row: every rank but 0 owns nothing
    const V<double> m4 = ints<double>(3 * world);
    auto t4 = ShardedNArray<double>::from_global(os, m4).permute();
    EXPECT(t4.shape() == Shape({1, 3 * world}) && t4.to_global() == m4);
    EXPECT(t4.permute().to_global() == m4);
  });
  it("slicing across shards (multi_indexable.cr:338-356): ranges, steps, reversal, an Int on the sharded axis", [&] {
    // host gather: out[j, y, x] = g[f0 + s0 * j, y0 + sy * y, x]
    auto host_slice = [&](int64_t f0, int64_t s0, int64_t n0, int64_t y0, int64_t sy, int64_t ny) {
      V<float> out;
      for (int64_t j = 0; j < n0; j++)
        for (int64_t y = 0; y < ny; y++)
          for (int64_t x = 0; x < gs[2]; x++) out.push_back(g[(size_t)(((f0 + s0 * j) * gs[1] + (y0 + sy * y)) * gs[2] + x)]);
      return out;
    };
    auto a = sg[{range(2, 4 * world)}];
    EXPECT(a.shape() == Shape({4 * world - 1, 12, 10}) && a.to_global() == host_slice(2, 1, 4 * world - 1, 0, 1, 12));
    auto b = sg[{range(nil, -1, nil), range(1, 2, 9)}];                      // rows reversed, every 2nd column-row
    EXPECT(b.shape() == Shape({gs[0], 5, 10}) && b.to_global() == host_slice(gs[0] - 1, -1, gs[0], 1, 2, 5));
    auto c = sg[{range(1, 3, nil)}];
    EXPECT(c.to_global() == host_slice(1, 3, (gs[0] - 1 - 1) / 3 + 1, 0, 1, 12));
    auto d = sg[{5, range(nil, -1, nil)}];                                   // ONE row of the sharded axis: its owner deals it out
    EXPECT(d.shape() == Shape({12, 10}) && d.to_global() == host_slice(5, 1, 1, 11, -1, 12));
    auto e = sg[{6 * world + 1, 3, 2}];                                      // every axis indexed: shape [1]
    EXPECT(e.shape() == Shape({1}) && e.to_global() == V<float>{g[(size_t)(((6 * world + 1) * 12 + 3) * 10 + 2)]});
    auto f = sg[{all, range(1, 3)}];                                         // axis 0 whole: local
    EXPECT(f.shape() == Shape({gs[0], 3, 10}) && f.to_global() == host_slice(0, 1, gs[0], 1, 1, 3));
    EXPECT_RAISES(IndexError, sg[{range(0, gs[0])}]);
  });
  it("scatter / fill across shards (multi_writable.cr:55-84): the gather plan run backwards", [&] {
    // rows reversed, every 2nd column-row <- a sharded source of that shape; then a scalar into a stepped range
    const Shape ss = {gs[0], 5, 10};
    const V<float> src = ints<float>(shape_to_size(ss));
    auto dst = ShardedNArray<float>::from_global(gs, g);
    dst.set_chunk({range(nil, -1, nil), range(1, 2, 9)}, ShardedNArray<float>::from_global(ss, src));
    V<float> want = g;
    for (int64_t j = 0; j < gs[0]; j++)
      for (int64_t y = 0; y < 5; y++)
        for (int64_t x = 0; x < 10; x++) want[(size_t)(((gs[0] - 1 - j) * 12 + (1 + 2 * y)) * 10 + x)] = src[(size_t)((j * 5 + y) * 10 + x)];
    EXPECT(dst.to_global() == want);
    dst.set_chunk({range(1, 3, nil), all, 4}, 2.5f);
    for (int64_t r = 1; r < gs[0]; r += 3)
      for (int64_t y = 0; y < 12; y++) want[(size_t)((r * 12 + y) * 10 + 4)] = 2.5f;
    EXPECT(dst.to_global() == want);
    const V<float> rowsrc = ints<float>(12 * 10);
    dst.set_chunk({5}, ShardedNArray<float>::from_global({12, 10}, rowsrc));      // ONE row of the sharded axis
    for (int64_t i = 0; i < 120; i++) want[(size_t)(5 * 120 + i)] = rowsrc[(size_t)i];
    EXPECT(dst.to_global() == want);
    EXPECT_RAISES(ShapeError, dst.set_chunk({range(0, 1)}, ShardedNArray<float>::from_global(gs, g)));
  });
  it("masked store on the distributed array (n_array.cr:510-551)", [&] {
    auto a = ShardedNArray<float>::from_global(gs, g);
    a.set_mask(a > sh, 0.0f);
    V<float> want = g;
    for (size_t i = 0; i < (size_t)n; i++) if (g[i] > h[i]) want[i] = 0.0f;
    EXPECT(a.to_global() == want);
  });

  Device::sync();
  std::printf("[rank %d] sharded_spec: %d expectations passed, %d failed\n", g_rank, g_passed, g_failed);
  Comm::destroy();
  return g_failed ? 1 : 0;
}
