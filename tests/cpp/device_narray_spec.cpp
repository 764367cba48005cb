// device_narray_spec.cpp -- the reference's own specs for the hot path, replayed on the
// device-backed array through the C++ host layer (include/ph_narray.hpp).  Each block cites
// the spec it replays; expectations are the reference's golden literals, typed in here (no
// oracle is linked: this program is product-side host code + libphgpu only).
//
//   spec/n_array_spec.cr:211-333, 446-447   fetch / set chunk, mask store, elementwise
//   spec/multi_writable_spec.cr:14-93       [1.., 1..] scalar / array / out of bounds, set_element
//   spec/view_util/*_transform_spec.cr      permute / reshape / reverse known answers
//   README.md:22-64                         narr + narr2, narr * narr2, get, [.., 1], argmax, slices
//   examples/heat_equation.cr               21-point rod, 10 001 steps
//
// Build: make -C tests/cpp      Run: tests/cpp/device_narray_spec   (exit 0 = all specs passed)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <string>
#include <vector>

#include "../../include/ph_narray.hpp"
#include "../../include/ph_narray_io.hpp"
#include "../../include/ph_pipeline.hpp"

using namespace Phase;
template <class T> using V = std::vector<T>;

static int g_failed = 0, g_passed = 0;
static std::string g_current;

// variadic: region literals `{1, all}` carry commas
#define EXPECT(...)                                                                              \
  do {                                                                                           \
    if (!(__VA_ARGS__)) { std::printf("  FAIL %s:%d  %s   [%s]\n", __FILE__, __LINE__, #__VA_ARGS__, g_current.c_str()); g_failed++; } \
    else g_passed++;                                                                             \
  } while (0)
#define EXPECT_RAISES(ExcType, ...)                                                              \
  do {                                                                                           \
    bool raised_ = false;                                                                        \
    try { __VA_ARGS__; } catch (const ExcType&) { raised_ = true; }                              \
    if (!raised_) { std::printf("  FAIL %s:%d  expected %s from: %s   [%s]\n", __FILE__, __LINE__, #ExcType, #__VA_ARGS__, g_current.c_str()); g_failed++; } \
    else g_passed++;                                                                             \
  } while (0)

static void it(const char* name, const std::function<void()>& body) {
  g_current = name;
  int before = g_failed;
  try { body(); }
  catch (const std::exception& e) { std::printf("  FAIL unexpected exception in '%s': %s\n", name, e.what()); g_failed++; }
  std::printf("%s %s\n", g_failed == before ? "ok  " : "FAIL", name);
}

static DeviceNArray<int32_t> stock_narr() { return DeviceNArray<int32_t>::from_host({2, 3}, V<int32_t>{0, 1, 2, 3, 4, 5}); }   // spec/n_array_spec.cr:7
template <class T> static DeviceNArray<T> narr(const Shape& shape, const V<T>& v) { return DeviceNArray<T>::from_host(shape, v); }

// Region literals -> {first, step, last} at bound 10: spec/spec_helper.cr:64-158 through
// spec/index_region_spec.cr.  Pure host code (ph_host.h): runs with --host-only on a box without a GPU.
static void host_specs() {
  it("region literals at bound 10 (spec_helper.cr:64-131)", [] {
    struct Case { Lit lit; int64_t first, step, last; };
    const int64_t bound = 10, mid = 5;
    const Case cases[] = {
        {range(mid, mid), mid, 1, mid},   {range(nil, mid), 0, 1, mid},      {range_ex(nil, mid), 0, 1, mid - 1},
        {Lit(0), 0, 1, 0},                {Lit(bound - 1), 9, 1, 9},         {range(mid, 0), mid, -1, 0},
        {range(mid, -1, nil), mid, -1, 0}, {range_ex(0, 2, bound), 0, 2, 8},  {range_ex(0, 2, bound - 1), 0, 2, 8},
        {range(5, -3, 1), 5, -3, 2},
        // implicit bounds
        {range(nil, nil), 0, 1, 9},       {range_ex(nil, nil), 0, 1, 9},     {range(mid, nil), mid, 1, 9},
        {range(nil, -1, nil), 9, -1, 0},  {range_ex(nil, -1, 2), 9, -1, 3},  {range(nil, -4, nil), 9, -4, 1},
        // negative indices
        {range(-bound, nil), 0, 1, 9},    {range(nil, -bound), 0, 1, 0},     {range(nil, -1), 0, 1, 9},
        {range(-mid, -mid + 2), 5, 1, 7}, {Lit(-mid), 5, 1, 5},
    };
    for (const Case& c : cases) {
      IndexRegion reg({c.lit}, {bound}, false);
      EXPECT(reg.r.first[0] == c.first && reg.r.step[0] == c.step && reg.r.last[0] == c.last);
      EXPECT(reg.r.proper_shape[0] == (c.last - c.first) / c.step + 1);
    }
  });
  it("out-of-bounds and malformed literals raise (spec_helper.cr:133-158; index_region_spec.cr:64-77)", [] {
    EXPECT_RAISES(IndexError, IndexRegion({range(nil, 10)}, {10}));
    EXPECT_RAISES(IndexError, IndexRegion({range(-11, nil)}, {10}));
    EXPECT_RAISES(IndexError, IndexRegion({Lit(10)}, {10}));
    EXPECT_RAISES(IndexError, IndexRegion({Lit(-11)}, {10}));
    EXPECT_RAISES(DimensionError, IndexRegion({all, all, all}, {10, 10}));
    IndexRegion padded({Lit(1)}, {2, 3});                       // short literals are padded with `..`
    EXPECT(padded.shape() == Shape({3}) && padded.proper_shape() == Shape({1, 3}));
    IndexRegion scalar({Lit(1), Lit(2)}, {2, 3});               // every axis dropped -> [1] (index_region.cr:323-335)
    EXPECT(scalar.shape() == Shape({1}));
    IndexRegion kept({Lit(1), Lit(2)}, {2, 3}, false);
    EXPECT(kept.shape() == Shape({1, 1}));
    IndexRegion empty({range_ex(0, 0), range_ex(0, 0)}, {2, 3});
    EXPECT(empty.shape() == Shape({0, 0}) && empty.size() == 0);
  });
  it("trim! / translate! / reverse! / fits_in? / cover (index_region_spec.cr:218-371)", [] {
    IndexRegion reg({range(0, 2, 8)}, {10});
    EXPECT(reg.shape() == Shape({5}) && reg.fits_in({9}) && !reg.fits_in({8}));
    reg.trim({5});
    EXPECT(reg.shape() == Shape({3}) && reg.r.last[0] == 4);
    IndexRegion two({range(1, 2), range(0, 2, 4)}, {10, 10});
    IndexRegion moved = two;
    moved.translate({3, 1});
    EXPECT(moved.r.first[0] == 4 && moved.r.last[0] == 5 && moved.r.first[1] == 1 && moved.r.last[1] == 5 && moved.shape() == two.shape());
    EXPECT_RAISES(IndexError, IndexRegion(two).translate({-4, 0}));
    IndexRegion rev = two;
    rev.reverse();
    EXPECT(rev.r.first[1] == 4 && rev.r.last[1] == 0 && rev.r.step[1] == -2 && rev.shape() == two.shape());
    // multidimensional goldens (index_region_spec.cr:64-77, 300-305)
    IndexRegion md({range(0, 2, 8), range(nil, -1, nil)}, {10, 4});
    EXPECT(md.shape() == Shape({5, 4}) && md.r.first[0] == 0 && md.r.first[1] == 3 && md.r.last[0] == 8 && md.r.last[1] == 0 &&
           md.r.step[0] == 2 && md.r.step[1] == -1);
    EXPECT_RAISES(IndexError, IndexRegion({range(0, 3), range(0, 2, 8)}, {10, 4}));
    IndexRegion big = IndexRegion::cover({20, 20});
    big.trim({10, 4});
    EXPECT(big.fits_in({10, 4}) && !big.fits_in({4, 10}));
    IndexRegion cov = IndexRegion::cover({4, 0, 2});
    EXPECT(cov.shape() == Shape({4, 0, 2}) && cov.size() == 0);
    EXPECT(shape_to_size({}) == 0 && shape_to_size({3, 4}) == 12);
  });
}

// The partition plans of the multi-GPU path, called from a compiled host (ph_host.h); no reference counterpart.
static void partition_host_specs() {
  it("shard ranges, slab layout, cross-shard transpose plan, extremum records (SURVEY.md 8(e), f-3)", [] {
    int64_t a = 0, b = 0, covered = 0;
    for (int r = 0; r < 3; r++) {                       // 10 rows over 3 ranks: 4 + 3 + 3, contiguous
      EXPECT(ph_shard_range(10, 3, r, &a, &b) == PH_HOST_OK && a == covered && b - a == (r == 0 ? 4 : 3));
      covered = b;
    }
    EXPECT(covered == 10);
    ph_slab slab;
    EXPECT(ph_slab_layout(2048, 8, 0, 2, &slab) == PH_HOST_OK && slab.count == 256 && slab.local_planes == 260 &&
           slab.lo_rank == -1 && slab.hi_rank == 1);
    EXPECT(ph_slab_layout(2048, 8, 7, 2, &slab) == PH_HOST_OK && slab.start == 1792 && slab.lo_rank == 6 && slab.hi_rank == -1);
    // [6, 4] transposed over 2 ranks: rank 0 owns rows 0..2, keeps new rows 0..1
    const int64_t shape[2] = {6, 4};
    const int32_t pattern[2] = {1, 0};
    ph_transpose_plan plan;
    ph_transpose_peer peers[2];
    EXPECT(ph_transpose_plan_of(shape, 2, pattern, 2, 0, &plan, peers) == PH_HOST_OK && !plan.local && plan.k == 1 && plan.j == 1);
    EXPECT(plan.new_shape[0] == 4 && plan.new_shape[1] == 6 && plan.my_rows[1] == 3 && plan.my_new_rows[1] == 2);
    EXPECT(peers[1].send0 == 2 && peers[1].send1 == 4 && peers[1].send_shape[0] == 2 && peers[1].send_shape[1] == 3);
    EXPECT(peers[1].recv0 == 3 && peers[1].recv1 == 6 && peers[1].recv_shape[0] == 2 && peers[1].recv_shape[1] == 3);
    const int32_t same[2] = {0, 1}, bad[2] = {1, 1};
    EXPECT(ph_transpose_plan_of(shape, 2, same, 2, 0, &plan, peers) == PH_HOST_OK && plan.local);
    EXPECT(ph_transpose_plan_of(shape, 2, bad, 2, 0, &plan, peers) == PH_HOST_INDEX_ERROR);
    // three shards: a tie between ranks 0 and 2 goes to the lower GLOBAL index, an empty shard is skipped
    uint8_t recs[3 * PH_EXTREMUM_RECORD_BYTES] = {0};
    auto put = [&](int r, float v, int64_t local, int64_t offset) {
      std::memcpy(recs + r * PH_EXTREMUM_RECORD_BYTES, &v, 4);
      std::memcpy(recs + r * PH_EXTREMUM_RECORD_BYTES + 16, &local, 8);
      std::memcpy(recs + r * PH_EXTREMUM_RECORD_BYTES + 24, &offset, 8);
    };
    put(0, 9.0f, 7, 0); put(1, 99.0f, -1, 10); put(2, 9.0f, 1, 10);
    int32_t winner = -5; int64_t gidx = -5;
    EXPECT(ph_combine_extremum_records(recs, 3, PH_F32, 1, &winner, &gidx) == PH_HOST_OK && winner == 0 && gidx == 7);
    put(0, 9.0f, 12, 0);
    EXPECT(ph_combine_extremum_records(recs, 3, PH_F32, 1, &winner, &gidx) == PH_HOST_OK && winner == 2 && gidx == 11);
  });
}

// JSON / YAML goldens of spec/n_array_spec.cr:520-558 and the binary dump, host side only
static void pipeline_host_specs() {
  it("row_chunks: every schedule is an ordered partition of the rows; the taper halves the last chunk", [] {
    for (int64_t n : {0, 1, 7, 10, 1000, 8192})
      for (int64_t chunks : {1, 2, 3, 4, 8, 16, 50})
        for (int taper : {0, 1, 3, 7, 12}) {
          int64_t at = 0;
          bool ok = true;
          for (const auto& c : row_chunks(n, chunks, taper)) { ok = ok && c.first == at && c.second > c.first; at = c.second; }
          EXPECT(ok && at == n);
        }
    std::vector<int64_t> sizes;
    for (const auto& c : row_chunks(8192, 4, 7)) sizes.push_back(c.second - c.first);
    EXPECT(sizes == std::vector<int64_t>({2048, 2048, 2048, 1024, 512, 256, 128, 64, 32, 16, 16}));
    sizes.clear();
    for (const auto& c : row_chunks(8192, 4, 7, 5)) sizes.push_back(c.second - c.first);
    EXPECT(sizes == std::vector<int64_t>({64, 64, 128, 256, 512, 1024, 2048, 2048, 1024, 512, 256, 128, 64, 32, 16, 16}));
    for (int64_t n : {0, 1, 7, 10, 1000})
      for (int ramp : {1, 2, 5}) {
        int64_t at = 0;
        bool ok = true;
        for (const auto& c : row_chunks(n, 3, 2, ramp)) { ok = ok && c.first == at && c.second > c.first; at = c.second; }
        EXPECT(ok && at == n);
      }
    EXPECT(row_chunks(8192, 16).size() == 16 && row_chunks(8192, 16)[15] == std::make_pair<int64_t, int64_t>(7680, 8192));
  });
}

static void join_host_specs() {
  it("ph_concat_shape: NArray#compatible? (n_array.cr:666-673) and the concatenated shape", [] {
    int32_t ax = -9;
    EXPECT(concat_shape_of({{2, 3}, {4, 3}}, 0, &ax) == Shape({6, 3}) && ax == 0);
    EXPECT(concat_shape_of({{2, 3, 4}, {2, 1, 4}, {2, 5, 4}}, 1, &ax) == Shape({2, 9, 4}) && ax == 1);
    EXPECT(concat_shape_of({{2, 3}, {2, 3}}, -1, &ax) == Shape({2, 6}) && ax == 1);
    EXPECT_RAISES(DimensionError, concat_shape_of({{2, 3}, {2, 4}}, 0, &ax));
    EXPECT_RAISES(DimensionError, concat_shape_of({{2, 3}, {2, 4}}, -1, &ax));      // `idx != axis` on the raw argument
    EXPECT_RAISES(IndexError, concat_shape_of({{2, 3}, {2}}, 0, &ax));
    EXPECT_RAISES(IndexError, concat_shape_of({{2, 3}, {2, 3}}, 2, &ax));
    EXPECT_RAISES(DimensionError, concat_shape_of({{2}, {2, 3}}, 0, &ax));
  });
}

static void io_host_specs() {
  it("to_json / from_json / to_yaml / from_yaml goldens (n_array_spec.cr:520-558)", [] {
    namespace H = IO::host;
    const Shape shape{2, 3};
    const V<int32_t> stock{0, 1, 2, 3, 4, 5};
    EXPECT(H::to_json(shape, stock) == "{\"shape\":[2,3],\"elements\":[0,1,2,3,4,5]}");
    EXPECT(H::to_json(Shape{0}, V<int32_t>{}) == "{\"shape\":[0],\"elements\":[]}");
    EXPECT(H::to_yaml(shape, stock) == "---\nshape: [2, 3]\nelements: [0, 1, 2, 3, 4, 5]\n");
    EXPECT(H::to_yaml(Shape{0}, V<int32_t>{}) == "---\nshape: [0]\nelements: []\n");
    Shape s; V<int32_t> e;
    H::from_json("{\"shape\":[2,3],\"elements\":[0,1,2,3,4,5]}", s, e);
    EXPECT(s == shape && e == stock);
    H::from_json(" { \"elements\" : [ ] , \"shape\" : [0] } ", s, e);          // any key order, any whitespace
    EXPECT(s == Shape{0} && e.empty());
    H::from_yaml("---\nshape: [2, 3]\nelements: [0, 1, 2, 3, 4, 5]\n", s, e);
    EXPECT(s == shape && e == stock);
    H::from_yaml("---\nshape: [0]\nelements: []\n", s, e);
    EXPECT(s == Shape{0} && e.empty());
    EXPECT_RAISES(ShapeError, H::from_json("{\"shape\":[2,2],\"elements\":[1,2,3]}", s, e));
    EXPECT_RAISES(IO::ParseError, H::from_json("{\"shape\":[2,2]}", s, e));
    EXPECT_RAISES(IO::ParseError, H::from_json("{\"shape\":[1],\"elements\":[1.5]}", s, e));    // not an Int32
    EXPECT_RAISES(IO::ParseError, H::from_json("{\"shape\":[1],\"elements\":[1],\"extra\":[2]}", s, e));
    // floats as Crystal's Float#to_s writes them: shortest text that round-trips, always with a fraction,
    // positional up to 1e15, d.de+X beyond (unpadded exponent); bools as true / false on the way in
    const V<double> fl{0.1, 2.0, -1.5e-7, 1e22, 5e-324, 0.30000000000000004};
    EXPECT(H::to_json(Shape{6}, fl) == "{\"shape\":[6],\"elements\":[0.1,2.0,-1.5e-7,1.0e+22,5.0e-324,0.30000000000000004]}");
    EXPECT(H::to_json(Shape{5}, V<double>{1e14, 1e15, 0.0001, 0.00001, -0.0}) == "{\"shape\":[5],\"elements\":[100000000000000.0,1.0e+15,0.0001,1.0e-5,-0.0]}");
    EXPECT(H::to_json(Shape{2}, V<float>{0.1f, 16777216.0f}) == "{\"shape\":[2],\"elements\":[0.1,16777216.0]}");   // Float32 digits, not the widened double's
    Shape fs; V<double> fe;
    H::from_json(H::to_json(Shape{6}, fl), fs, fe);
    EXPECT(fe == fl);
    H::from_yaml(H::to_yaml(Shape{2, 3}, fl), fs, fe);
    EXPECT(fs == Shape({2, 3}) && fe == fl);
    V<Bool> be;
    H::from_json("{\"shape\":[3],\"elements\":[true,false,true]}", fs, be);
    EXPECT(be == V<Bool>({1, 0, 1}));
    EXPECT_RAISES(std::invalid_argument, H::to_json(Shape{1}, V<float>{std::nanf("")}));
  });
}

// --write-dump PATH: a known [3,4] f32 array in the binary format; --read-dump PATH DTYPE: print it back as JSON.
// tests/test_cpp_host_layer.py exchanges files with the Python mirror's io module through these.
static int dump_cli(int argc, char** argv) {
  namespace H = IO::host;
  if (std::strcmp(argv[1], "--write-dump") == 0 && argc > 2) {
    V<float> v(12);
    for (int i = 0; i < 12; i++) v[i] = 0.25f * i - 1.0f;
    H::dump(argv[2], Shape{3, 4}, v);
    return 0;
  }
  if (std::strcmp(argv[1], "--read-dump") == 0 && argc > 3) {
    Shape s;
    std::string dt = argv[3];
    if (dt == "f64") { V<double> e; H::load(argv[2], s, e); std::printf("%s\n", H::to_json(s, e).c_str()); }
    else if (dt == "i16") { V<int16_t> e; H::load(argv[2], s, e); std::printf("%s\n", H::to_json(s, e).c_str()); }
    else if (dt == "u8") { V<uint8_t> e; H::load(argv[2], s, e); std::printf("%s\n", H::to_json(s, e).c_str()); }
    else { V<float> e; H::load(argv[2], s, e); std::printf("%s\n", H::to_json(s, e).c_str()); }
    return 0;
  }
  return 2;
}

int main(int argc, char** argv) {
  if (argc > 1 && (std::strcmp(argv[1], "--write-dump") == 0 || std::strcmp(argv[1], "--read-dump") == 0)) {
    try { return dump_cli(argc, argv); }
    catch (const std::exception& e) { std::fprintf(stderr, "%s\n", e.what()); return 4; }
  }
  if (argc > 1 && std::strcmp(argv[1], "--host-only") == 0) {
    host_specs();
    io_host_specs();
    partition_host_specs();
    pipeline_host_specs();
    join_host_specs();
    std::printf("%d expectations passed, %d failed (host-only)\n", g_passed, g_failed);
    return g_failed ? 1 : 0;
  }
  try { Device::init(0); }
  catch (const DeviceError& e) {
    std::fprintf(stderr, "device_narray_spec: %s\nno CUDA device -- the device path has no CPU fallback\n", e.what());
    return 3;
  }

  host_specs();
  io_host_specs();
  partition_host_specs();

  // (ran on a B200 in round 2: profiles/r02_cpp_host_spec.log; PH_SPEC_NO_DEVICE_IO=1 skips it, e.g. on a read-only /tmp)
  if (!std::getenv("PH_SPEC_NO_DEVICE_IO"))
  it("device arrays through the I/O formats (n_array.cr:807-912; binary checkpoint)", [] {
    auto stock = stock_narr();
    EXPECT(IO::to_json(stock) == "{\"shape\":[2,3],\"elements\":[0,1,2,3,4,5]}");
    EXPECT(IO::from_json<int32_t>(IO::to_json(stock)) == stock);
    EXPECT(IO::from_yaml<int32_t>(IO::to_yaml(stock)) == stock);
    EXPECT(IO::to_yaml(stock.view().permute()) == "---\nshape: [3, 2]\nelements: [0, 3, 1, 4, 2, 5]\n");   // a view is materialised first
    V<float> v(300 * 20);
    for (size_t i = 0; i < v.size(); i++) v[i] = (float)((i * 2654435761u) % 1000) / 7.0f;
    auto big = narr<float>({300, 20}, v);
    const char* path = "/tmp/ph_cpp_spec_checkpoint.phbin";
    IO::dump(big, path);
    auto back = IO::load<float>(path);
    EXPECT(back == big);
    // checkpoint / resume of a heat run: 3 + 4 steps == 7 steps, bit for bit
    auto grid = narr<float>({30, 200}, v);
    auto whole = Heat::simulate(grid, 0.1f, 7);
    IO::dump(Heat::simulate(grid, 0.1f, 3), path);
    EXPECT(Heat::simulate(IO::load<float>(path), 0.1f, 4) == whole);
    EXPECT_RAISES(IO::ParseError, IO::load<double>(path));                     // wrong element type
    std::remove(path);
  });

  it("#unsafe_fetch_chunk goldens (n_array_spec.cr:211-229)", [] {
    auto s = stock_narr();
    auto a = s.unsafe_fetch_chunk(IndexRegion({1, range(0, 2, 2)}, {2, 3}));
    EXPECT(a.shape() == Shape({2}) && a.to_host() == V<int32_t>({3, 5}));
    auto b = s.unsafe_fetch_chunk(IndexRegion({-2, range(-1, 0)}, {2, 3}));
    EXPECT(b.to_host() == V<int32_t>({2, 1, 0}));
    auto e = s.unsafe_fetch_chunk(IndexRegion({range_ex(0, 0), range_ex(0, 0)}, {2, 3}));
    EXPECT(e.shape() == Shape({0, 0}) && e.size() == 0 && e.to_host().empty());
    EXPECT(s.get({1, 1}) == 4);                       // #unsafe_fetch_element :231-236
    EXPECT(s.get({-1, -1}) == 5);
  });

  it("#unsafe_set_chunk goldens (n_array_spec.cr:238-295)", [] {
    auto n = stock_narr().clone();
    n.unsafe_set_chunk(IndexRegion({1, range(0, 2, 2)}, {2, 3}), narr<int32_t>({2}, {6, 7}));
    EXPECT(n.to_host() == V<int32_t>({0, 1, 2, 6, 4, 7}));
    n = stock_narr().clone();
    n.unsafe_set_chunk(IndexRegion({-2, range(-1, 0)}, {2, 3}), narr<int32_t>({3}, {6, 7, 8}));
    EXPECT(n.to_host() == V<int32_t>({8, 7, 6, 3, 4, 5}));
    n = stock_narr().clone();
    IndexRegion empty({range_ex(0, 0), range_ex(0, 0)}, {2, 3});
    n.unsafe_set_chunk(empty, narr<int32_t>({1}, {0})[{range_ex(nil, -1)}]);       // NArray[0][...-1]
    EXPECT(n == stock_narr());
    n.unsafe_set_chunk(IndexRegion({1, range(0, 2, 2)}, {2, 3}), 6);
    EXPECT(n.to_host() == V<int32_t>({0, 1, 2, 6, 4, 6}));
    n = stock_narr().clone();
    n.unsafe_set_chunk(IndexRegion({-2, range(-1, 0)}, {2, 3}), 6);
    EXPECT(n.to_host() == V<int32_t>({6, 6, 6, 3, 4, 5}));
    n = stock_narr().clone();
    n.unsafe_set_chunk(empty, 6);
    EXPECT(n == stock_narr());
  });

  it("[]=(mask, value) goldens (n_array_spec.cr:297-333)", [] {
    auto mask = narr<Bool>({2, 3}, {1, 0, 1, 0, 1, 0});
    auto n = stock_narr().clone();
    n.set_mask(mask, 6);
    EXPECT(n.to_host() == V<int32_t>({6, 1, 6, 3, 6, 5}));
    n = stock_narr().clone();
    n.set_mask(mask, stock_narr() + 10);
    EXPECT(n.to_host() == V<int32_t>({10, 1, 12, 3, 14, 5}));
    EXPECT_RAISES(DimensionError, n.set_mask(narr<Bool>({3, 2}, {1, 0, 1, 0, 1, 0}), 6));
    EXPECT(&n[mask] == static_cast<const MultiIndexable<int32_t>*>(&n));   // narr[mask] is self (multi_indexable.cr:479-481)
  });

  it("elementwise goldens (n_array_spec.cr:317-319, 446-447, 462-466; README.md:22-41)", [] {
    auto s = stock_narr();
    EXPECT((s + 10).to_host() == V<int32_t>({10, 11, 12, 13, 14, 15}));
    EXPECT((s * 2).to_host() == V<int32_t>({0, 2, 4, 6, 8, 10}));
    EXPECT(s.pow(2).to_host() == V<int32_t>({0, 1, 4, 9, 16, 25}));
    EXPECT((s * 2 + s) == s * 3);
    auto a = narr<int32_t>({2, 3}, {1, 0, 0, 0, 1, 0});
    auto b = narr<int32_t>({2, 3}, {0, 1, 2, 10, 11, 12});
    EXPECT((a + b).to_host() == V<int32_t>({1, 1, 2, 10, 12, 12}));
    EXPECT((a * b).to_host() == V<int32_t>({0, 0, 0, 0, 11, 0}));
    EXPECT(a.get({0, 0}) == 1);
    EXPECT(a[{all, 1}].to_host() == V<int32_t>({0, 1}));
    EXPECT(a.view({all, 1}).to_narr().to_host() == V<int32_t>({0, 1}));
    // scalar on the LEFT keeps operand order (patches/number.cr:6-15)
    EXPECT((10 - s).to_host() == V<int32_t>({10, 9, 8, 7, 6, 5}));
    // Int / Int -> Float64
    auto q = (s / 2);
    static_assert(std::is_same<decltype(q), DeviceNArray<double>>::value, "Int / Int is Float64");
    EXPECT(q.to_host() == V<double>({0.0, 0.5, 1.0, 1.5, 2.0, 2.5}));
    EXPECT(s.floordiv(-2).to_host() == V<int32_t>({0, -1, -1, -2, -2, -3}));   // floored
    EXPECT((s % -4).to_host() == V<int32_t>({0, -3, -2, -1, 0, -3}));          // sign of the divisor
    EXPECT((-s).to_host() == V<int32_t>({0, -1, -2, -3, -4, -5}));
    EXPECT((~s).to_host() == V<int32_t>({-1, -2, -3, -4, -5, -6}));
    EXPECT(((s & 6) | 1).to_host() == V<int32_t>({1, 1, 3, 3, 5, 5}));
    // comparisons -> NArray(Bool)
    EXPECT((s > 2).to_host() == V<Bool>({0, 0, 0, 1, 1, 1}));
    EXPECT((s <= b).to_host() == V<Bool>({1, 1, 1, 1, 1, 1}));
    EXPECT(s.eq(b).to_host() == V<Bool>({1, 1, 1, 0, 0, 0}));
    EXPECT(s.match(4).to_host() == V<Bool>({0, 0, 0, 0, 1, 0}));
    EXPECT_RAISES(ShapeError, s + narr<int32_t>({3, 2}, {0, 1, 2, 3, 4, 5}));        // multi_indexable.cr:935-940
    EXPECT_RAISES(DimensionError, s.eq(narr<int32_t>({3, 2}, {0, 1, 2, 3, 4, 5})));   // :900-902
  });

  it("float ops are single IEEE operations; a*b+c is two roundings", [] {
    // 1 + 2^-24 is not representable: (1 * (1 + 2^-23)) + 2^-24 ... choose operands where an FMA would differ
    float a = 1.0f + std::ldexp(1.0f, -12), b = 1.0f + std::ldexp(1.0f, -12), c = -1.0f;
    float two = (a * b);            // rounded product
    volatile float want = two + c;  // second rounding
    auto A = DeviceNArray<float>::fill({1000}, a), B = DeviceNArray<float>::fill({1000}, b), Cc = DeviceNArray<float>::fill({1000}, c);
    auto got = (A * B + Cc).to_host();
    auto fused = A.mul_add(B, Cc).to_host();
    float fma_result = std::fma(a, b, c);
    EXPECT(got[0] == want && got[999] == want);
    EXPECT(fused[0] == want && fused[500] == want);
    EXPECT(fma_result != want);     // the test would not notice an FMA otherwise
    auto p = DeviceNArray<double>::fill({4}, 0.05).pow(2).to_host();               // Float ** Int = powi (heat_equation.cr:20)
    EXPECT(p[0] == 0.05 * 0.05);
  });

  it("data-dependent errors come back as the reference's classes", [] {
    auto big = DeviceNArray<int32_t>::fill({100}, std::numeric_limits<int32_t>::max());
    auto r = big + 1;
    EXPECT_RAISES(OverflowError, Device::raise_pending());
    EXPECT((big.wrapping_add(1)).get({0}) == std::numeric_limits<int32_t>::min());   // &+ wraps
    Device::raise_pending();
    auto z = stock_narr().floordiv(0);
    EXPECT_RAISES(DivisionByZeroError, Device::raise_pending());
    auto nanarr = narr<float>({3}, {1.0f, std::nanf(""), 2.0f});
    EXPECT_RAISES(ArgumentError, nanarr.max());
    EXPECT_RAISES(EmptyError, DeviceNArray<float>::fill({3, 0, 2}, 0.0f).max());
    EXPECT(DeviceNArray<float>::fill({3, 0, 2}, 0.0f).sum() == 0.0f);
  });

  it("MultiWritable goldens (multi_writable_spec.cr:14-93)", [] {
    V<int64_t> base(12);
    for (int i = 0; i < 12; i++) base[i] = i;
    auto d = narr<int64_t>({3, 4}, base);
    d.set_chunk({range(1, nil), range(1, nil)}, 10);
    EXPECT(d.to_host() == V<int64_t>({0, 1, 2, 3, 4, 10, 10, 10, 8, 10, 10, 10}));
    d = narr<int64_t>({3, 4}, base);
    d.set_chunk({range(1, nil), range(1, nil)}, narr<int64_t>({2, 3}, {10, 11, 12, 13, 14, 15}));
    EXPECT(d.to_host() == V<int64_t>({0, 1, 2, 3, 4, 10, 11, 12, 8, 13, 14, 15}));
    EXPECT_RAISES(ShapeError, d.set_chunk({range(2, nil), range(2, nil)}, narr<int64_t>({2, 3}, {10, 11, 12, 13, 14, 15})));
    EXPECT_RAISES(IndexError, d.set_chunk({range(1, 7), range(1, nil)}, 10));
    d.set_element({-1, -2}, 77);
    EXPECT(d.get({2, 2}) == 77);
    EXPECT_RAISES(IndexError, d.set_element({10, 10}, 1));
    EXPECT_RAISES(IndexError, d.set_element({-10, -10}, 1));
    // trailing ones are compatible (shape_util.cr:6-32)
    d.set_chunk({range(1, nil), range(1, nil)}, narr<int64_t>({2, 3, 1}, {20, 21, 22, 23, 24, 25}));
    EXPECT(d[{range(1, nil), range(1, nil)}].to_host() == V<int64_t>({20, 21, 22, 23, 24, 25}));
  });

  it("view transforms (permute / reshape / reverse specs; view.cr)", [] {
    V<int32_t> v(24);
    for (int i = 0; i < 24; i++) v[i] = i;
    auto n = narr<int32_t>({2, 3, 4}, v);
    auto t = n.permute();                                   // reversed axes -> [4,3,2]
    EXPECT(t.shape() == Shape({4, 3, 2}) && t.get({3, 2, 1}) == n.get({1, 2, 3}) && t.get({1, 0, 1}) == 13);
    auto p = n.view().permute({1, 2, 0});                   // out axis i = source axis pattern[i]
    EXPECT(p.shape() == Shape({3, 4, 2}) && p.get({2, 3, 1}) == n.get({1, 2, 3}));
    auto r = n.reverse();                                   // every axis flipped
    EXPECT(r.get({0, 0, 0}) == 23 && r.to_host()[1] == 22);
    auto rs = n.view().reshape({6, 4});                      // [3,4]->[6,2]-style reshape keeps lex order
    EXPECT(rs.get({5, 3}) == 23 && rs.get({2, 1}) == 9);
    auto chain = n.view({all, range(nil, -1, nil), range(0, 2, nil)}).permute().reverse();
    auto host = chain.to_narr();
    EXPECT(host.shape() == Shape({2, 3, 2}));
    EXPECT(host.get({0, 0, 0}) == n.get({1, 0, 2}) && host.get({1, 2, 1}) == n.get({0, 2, 0}));
    EXPECT_RAISES(ShapeError, n.view().reshape({5, 5}));
    EXPECT_RAISES(IndexError, n.view().permute({0, 1, 5}));
    // MutableView write-through: scatter through the chain (mutable_view.cr:16-18)
    auto m = narr<int32_t>({2, 3}, {0, 0, 0, 0, 0, 0});
    m.mutable_view().permute().set_chunk({all, all}, narr<int32_t>({3, 2}, {1, 2, 3, 4, 5, 6}));
    EXPECT(m.to_host() == V<int32_t>({1, 3, 5, 2, 4, 6}));
    // reshape ALIASES the buffer (n_array.cr:429-433); clone does not
    auto alias = m.reshape({3, 2});
    alias.set_element({0, 0}, 99);
    EXPECT(m.get({0, 0}) == 99 && m.clone().data() != m.data() && alias.data() == m.data());
  });

  it("reductions, argmax idiom and slices (README.md:56-64)", [] {
    auto b = narr<int32_t>({2, 3}, {0, 1, 2, 10, 11, 12});
    auto am = b.argmax();
    EXPECT(am.first == 12 && am.second == Coord({1, 2}));
    EXPECT(b.sum() == 36 && b.min() == 0 && b.max() == 12);
    auto sl = b.slices(1);
    EXPECT(sl.size() == 3 && sl[0].to_host() == V<int32_t>({0, 10}) && sl[2].to_host() == V<int32_t>({2, 12}));
    {   // all slices along an axis come from ONE permuting copy and stay independent arrays
      V<int32_t> v(2 * 3 * 4);
      for (int i = 0; i < 24; i++) v[i] = i;
      auto cube = narr<int32_t>({2, 3, 4}, v);
      long long before = ph_launch_count();
      auto mid = cube.slices(1);
      EXPECT(ph_launch_count() - before == 1 && mid.size() == 3 && mid[1].shape() == Shape({2, 4}));
      EXPECT(mid[1].to_host() == V<int32_t>({4, 5, 6, 7, 16, 17, 18, 19}) && mid[2].get({1, 3}) == 23);
      mid[1].set_chunk({all, all}, -1);
      EXPECT(mid[0].to_host() == V<int32_t>({0, 1, 2, 3, 12, 13, 14, 15}) && cube.get({0, 1, 0}) == 4);
      EXPECT_RAISES(IndexError, cube.slices(3));
    }
    EXPECT(b.sum(0).to_host() == V<int32_t>({10, 12, 14}) && b.sum(1).to_host() == V<int32_t>({3, 33}));
    EXPECT(b.argmax(1).to_host() == V<int64_t>({2, 2}) && b.max(0).to_host() == V<int32_t>({10, 11, 12}));
    auto ties = narr<float>({2, 4}, {1, 7, 7, 0, 7, 7, 7, 7});
    EXPECT(ties.argmax().second == Coord({0, 1}));         // the FIRST maximum wins
    EXPECT(ties.argmax(1).to_host() == V<int64_t>({1, 0}));
    auto tiled = narr<int32_t>({1, 3}, {1, 2, 3}).tile({2, 2});   // multi_indexable.cr:818-827
    EXPECT(tiled.shape() == Shape({2, 6}) && tiled.to_host() == V<int32_t>({1, 2, 3, 1, 2, 3, 1, 2, 3, 1, 2, 3}));
    // broadcasting is defined as tile + op
    auto col = narr<int32_t>({2, 1}, {100, 200});
    EXPECT(b.broadcast_op(PH_ADD, col) == b + col.tile({1, 3}));
  });

  it("arbitrary blocks raise instead of running on the CPU", [] {
    auto s = stock_narr();
    EXPECT_RAISES(DeviceBlockError, s.map([](int32_t x) { return x * x; }));
    EXPECT_RAISES(DeviceBlockError, s.each_with(s, [](int32_t, int32_t) {}));
    EXPECT_RAISES(DeviceBlockError, s.view().process([](int32_t x) { return x; }));
    EXPECT_RAISES(DeviceBlockError, DeviceNArray<int32_t>::build(Shape({2, 3}), [](const Coord&) { return 0; }));
  });

  it("get_available / has_region? / region edge cases", [] {
    auto s = stock_narr();
    EXPECT(s.get_available({range(0, 5), range(1, 9)}).to_host() == V<int32_t>({1, 2, 4, 5}));
    EXPECT(s.has_region({range(0, 1), 2}) && !s.has_region({range(0, 2), 2}) && !s.has_region({0, 0, 0}));
    EXPECT_RAISES(IndexError, s[{5, 0}]);
    EXPECT_RAISES(DimensionError, s[{0, 0, 0}]);
    EXPECT(s.get_chunk({1, all}, false).shape() == Shape({1, 3}));     // drop: false keeps the axis
    IndexRegion reg({range(0, 2, 8)}, {10});
    EXPECT(reg.shape() == Shape({5}) && reg.fits_in({9}) && !reg.fits_in({8}));
    reg.trim({5});
    EXPECT(reg.shape() == Shape({3}));
  });

  it("examples/heat_equation.cr: 21 points, 10 001 steps", [] {
    const double COEFF = (237 * 0.01) / ((double)(2700 * 900) * (0.05 * 0.05));
    auto state = DeviceNArray<double>::fill({21}, 20.0);
    state.set_element({0}, 0.0);
    state.set_element({-1}, 100.0);
    auto fin = Heat::simulate(state, COEFF, 10001, PH_HEAT_EXAMPLE1D).to_host();
    double sum = 0;
    for (double x : fin) sum += x;
    EXPECT(std::fabs(sum - 480.0) < 1e-9);                                   // zero-flux ends conserve heat
    EXPECT(std::fabs(fin[0] - 14.381532) < 1e-6 && std::fabs(fin[20] - 42.473872) < 1e-6);
    // update_temp written literally with the reference's operators (heat_equation.cr:38-51): one-element
    // chunks, array-valued []=, scalar on the left -- bit-identical to the fused example kernel
    {
      auto st = state.clone();
      for (int step = 0; step < 2; step++) {
        auto temp_diff = DeviceNArray<double>::fill(st.shape(), 0.0);
        temp_diff.set_chunk({0}, (st[{1}] - st[{0}]) * COEFF);
        temp_diff.set_chunk({-1}, (st[{-2}] - st[{-1}]) * COEFF);
        auto centre = st[{range_ex(1, -1)}];
        for (int64_t idx = 0; idx < centre.shape()[0]; idx++) {
          double center_temp = centre[{idx}].to_scalar();
          temp_diff.set_chunk({idx + 1}, (st[{idx}] - 2 * center_temp + st[{idx + 2}]) * COEFF);
        }
        st = st + temp_diff;
      }
      EXPECT(st == Heat::simulate(state, COEFF, 2, PH_HEAT_EXAMPLE1D));
      EXPECT(st[{3}].scalar() && !st.scalar() && !st.empty() && st.last() == 100.0 + (st.last() - 100.0));
      EXPECT_RAISES(ShapeError, st.to_scalar());
    }
    // one fused step == the slice-arithmetic form written with the reference's operators (N-D rule, 2-D)
    const int64_t H = 37, W = 53;
    V<float> init((size_t)(H * W));
    for (size_t i = 0; i < init.size(); i++) init[i] = (float)((i * 2654435761u) % 1000) / 7.0f;
    auto s = narr<float>({H, W}, init);
    const float C = 0.1f;
    auto c = s[{range_ex(1, -1), range_ex(1, -1)}];
    auto d0 = (s[{range_ex(0, -2), range_ex(1, -1)}] - 2.0f * c) + s[{range(2, nil), range_ex(1, -1)}];
    auto d1 = (s[{range_ex(1, -1), range_ex(0, -2)}] - 2.0f * c) + s[{range_ex(1, -1), range(2, nil)}];
    auto nxt = s.clone();
    nxt.set_chunk({range_ex(1, -1), range_ex(1, -1)}, c + (d0 + d1) * C);
    EXPECT(Heat::update_temp(s, C) == nxt);                                  // bit-identical
  });

  it("get_chunk(coord, region_shape) (multi_indexable.cr:369-395): the source's own examples", [] {
    auto n = narr<int32_t>({3, 3}, {1, 2, 3, 4, 5, 6, 7, 8, 9});
    EXPECT(n.get_chunk(Coord{1, 0}, Shape{1, 3}).to_host() == V<int32_t>({4, 5, 6}));
    EXPECT(n.get_chunk(Coord{1, 1}, Shape{2, 2}).to_host() == V<int32_t>({5, 6, 8, 9}));
    EXPECT(n.get_chunk(Coord{3, 3}, Shape{0, 0}).shape() == Shape({0, 0}));
    EXPECT_RAISES(ShapeError, n.get_chunk(Coord{1, 0}, Shape{10, 10}));
    EXPECT_RAISES(DimensionError, n.get_chunk(Coord{0}, Shape{1}));
    EXPECT_RAISES(ArgumentError, n.get_chunk(Coord{-1, 0}, Shape{1, 1}));
  });

  it("NArray.concatenate / push / << / wrap (n_array.cr:321-344, 666-750): one strided copy per input", [] {
    auto a = narr<int32_t>({2, 3}, {0, 1, 2, 3, 4, 5});
    auto b = narr<int32_t>({1, 3}, {10, 11, 12});
    auto c = narr<int32_t>({2, 2}, {20, 21, 22, 23});
    EXPECT(DeviceNArray<int32_t>::concatenate({&a, &b}, 0).to_host() == V<int32_t>({0, 1, 2, 3, 4, 5, 10, 11, 12}));
    auto side = a.concatenate(c, 1);
    EXPECT(side.shape() == Shape({2, 5}) && side.to_host() == V<int32_t>({0, 1, 2, 20, 21, 3, 4, 5, 22, 23}));
    auto t = a.view().permute();                                            // views join like arrays: [3,2] ++ [3,2] along axis 1
    EXPECT(DeviceNArray<int32_t>::concatenate({&t, &t}, 1).to_host() == V<int32_t>({0, 3, 0, 3, 1, 4, 1, 4, 2, 5, 2, 5}));
    EXPECT(DeviceNArray<int32_t>::concatenate({&a, &a}, -1).shape() == Shape({2, 6}));
    EXPECT_RAISES(DimensionError, DeviceNArray<int32_t>::concatenate({&a, &c}, 0));
    EXPECT_RAISES(DimensionError, DeviceNArray<int32_t>::concatenate({&a, &c}, -1));   // a negative axis excludes nothing (compatible?)
    EXPECT_RAISES(IndexError, DeviceNArray<int32_t>::concatenate({&a, &a}, 2));
    auto p = a.clone();
    auto alias = p.reshape({3, 2});
    p << b;
    EXPECT(p.shape() == Shape({3, 3}) && p.to_host() == V<int32_t>({0, 1, 2, 3, 4, 5, 10, 11, 12}));
    EXPECT(alias.to_host() == V<int32_t>({0, 1, 2, 3, 4, 5}));              // an alias made before the push keeps the old buffer
    p.push({&b, &a});
    EXPECT(p.shape() == Shape({6, 3}) && p.get({5, 2}) == 5 && p.get({3, 0}) == 10);
    EXPECT_RAISES(DimensionError, p.push({&c}));
    auto w = DeviceNArray<int32_t>::wrap({&a, &a});
    EXPECT(w.shape() == Shape({2, 2, 3}) && w.get({1, 1, 2}) == 5 && w.get({0, 0, 1}) == 1);
    EXPECT_RAISES(DimensionError, DeviceNArray<int32_t>::wrap({&a, &b}));
  });

  // ---- streams, pinned arrays, asynchronous transfers, the chunked host -> device -> host pipeline (ph_pipeline.hpp)
  it("RowPipeline.map_rows: a*b+c over pinned host operands equals the resident computation, bit for bit", [] {
    const int64_t R = 1000, Cc = 768;                                        // ragged chunks
    V<float> av((size_t)(R * Cc)), cv((size_t)(R * Cc)), bv((size_t)Cc);
    for (size_t i = 0; i < av.size(); i++) { av[i] = (float)((i * 2654435761u) % 2001) / 1000.0f - 1.0f; cv[i] = (float)((i * 40503u) % 1999) / 999.0f - 1.0f; }
    for (size_t i = 0; i < bv.size(); i++) bv[i] = (float)((i * 7919u) % 2003) / 1001.0f - 1.0f;
    PinnedArray<float> a({R, Cc}, av), c({R, Cc}, cv), b({1, Cc}, bv), out({R, Cc});
    auto want = (narr<float>({R, Cc}, av).broadcast_op(PH_MUL, narr<float>({1, Cc}, bv)) + narr<float>({R, Cc}, cv)).to_host();
    auto expr = [](const std::vector<DeviceNArray<float>>& in, const std::vector<DeviceNArray<float>>& shared) {
      return in[0].broadcast_op(PH_MUL, shared[0]) + in[1];
    };
    RowPipeline equal(8, 0, 0), tapered(4, 7, 5);
    for (RowPipeline* pipe : {&equal, &tapered, &equal}) {                   // streams and pool blocks are reused
      std::memset(out.data(), 0, (size_t)out.size() * sizeof(float));
      pipe->map_rows<float>(expr, {&a, &c}, out, {&b});
      EXPECT(std::memcmp(out.data(), want.data(), want.size() * sizeof(float)) == 0);
    }
    // data-dependent errors of a pipelined step surface at its synchronising end
    PinnedArray<int32_t> ia({64, 8}, V<int32_t>(512, std::numeric_limits<int32_t>::max())), io({64, 8});
    RowPipeline ip(4, 2);
    EXPECT_RAISES(OverflowError, ip.map_rows<int32_t>([](const std::vector<DeviceNArray<int32_t>>& in, const std::vector<DeviceNArray<int32_t>>&) { return in[0] + 1; }, {&ia}, io));
    // a row operand with another leading extent is a ShapeError before anything is queued
    PinnedArray<float> shorter({R - 1, Cc});
    EXPECT_RAISES(ShapeError, equal.map_rows<float>(expr, {&a, &shorter}, out, {&b}));
    // explicit streams: two chains ordered by wait(), joined by the main stream
    Stream s1, s2;
    DeviceNArray<float> x = [&] { StreamScope on(s1); return from_host_async<float>({R, Cc}, a.data()) * 2.0f; }();
    s2.wait(&s1);
    { StreamScope on(s2); to_host_async<float>(x + 1.0f, out.data()); }
    main_stream_wait(s2);
    Device::sync();
    bool same = true;
    for (size_t i = 0; i < av.size() && same; i++) same = out[(int64_t)i] == av[i] * 2.0f + 1.0f;
    EXPECT(same);
  });

  std::printf("%d expectations passed, %d failed, %lld kernel launches\n", g_passed, g_failed, (long long)ph_launch_count());
  Device::shutdown();
  return g_failed ? 1 : 0;
}
