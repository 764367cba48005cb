// device_narray_bench.cpp -- what the fluent API costs from a COMPILED host (the Crystal drop-in's
// situation): the BASELINE expression `a * b_rowvec + c` written with operators on
// Phase::DeviceNArray (two kernels, two result arrays from the stream-ordered pool per step), the same
// through the fused entry point, and each_slice over a rank-3 array.  CUDA-event timing on the library's
// stream, one JSON line per row.   Build: make -C tests/cpp   Run: tests/cpp/device_narray_bench
#include <cstdio>
#include <vector>

#include "../../include/ph_narray.hpp"
#include "../../include/ph_pipeline.hpp"

using namespace Phase;

template <class F>
static double time_ms(F&& fn, int warm, int reps, int inner) {
  for (int i = 0; i < warm; i++) fn();
  double best = 1e30;
  for (int r = 0; r < reps; r++) {
    float ms = 0;
    Device::check(ph_timer_start());
    for (int i = 0; i < inner; i++) fn();
    Device::check(ph_timer_stop(&ms));
    if (ms / inner < best) best = ms / inner;
  }
  return best;
}

int main() {
  try { Device::init(0); }
  catch (const DeviceError& e) { std::fprintf(stderr, "device_narray_bench: %s (no CPU fallback)\n", e.what()); return 3; }
  const int64_t R = 8192, Cc = 8192, N = R * Cc;
  auto a = DeviceNArray<float>::fill({R, Cc}, 1.5f), c = DeviceNArray<float>::fill({R, Cc}, -0.25f);
  auto b = DeviceNArray<float>::fill({1, Cc}, 0.75f);
  const double bytes_two = 2.0 * N * 4 + Cc * 4 + 3.0 * N * 4, bytes_fused = 3.0 * N * 4 + Cc * 4;

  double ms = time_ms([&] { auto out = a.broadcast_op(PH_MUL, b) + c; (void)out; }, 3, 20, 8);
  std::printf("{\"row\": \"a.broadcast_op(*, b) + c through operators (2 kernels, 2 pool allocations)\", \"ms\": %.5f, \"gbs\": %.1f}\n",
              ms, bytes_two / ms / 1e6);
  ms = time_ms([&] { auto out = a.mul_add(b, c); (void)out; }, 3, 20, 8);
  std::printf("{\"row\": \"a.mul_add(b, c) fused entry\", \"ms\": %.5f, \"gbs\": %.1f}\n", ms, bytes_fused / ms / 1e6);
  ms = time_ms([&] { auto m = a > c; (void)m; }, 3, 20, 8);
  std::printf("{\"row\": \"a > c -> DeviceNArray<Bool>\", \"ms\": %.5f, \"gbs\": %.1f}\n", ms, (2.0 * N * 4 + N) / ms / 1e6);
  ms = time_ms([&] { auto s = a[{range(0, 2, nil), range(nil, -1)}]; (void)s; }, 3, 20, 8);
  std::printf("{\"row\": \"a[0..2.., ..-1] gather incl. region parse + descriptor compile\", \"ms\": %.5f, \"gbs\": %.1f}\n", ms,
              (double)N * 4 / ms / 1e6);
  {
    auto cube = DeviceNArray<double>::fill({64, 4096, 1024}, 2.0);
    const double cb = 2.0 * 64 * 4096 * 1024 * 8;
    ms = time_ms([&] { auto s = cube.slices(0); (void)s; }, 1, 5, 1);
    std::printf("{\"row\": \"slices(0) of [64,4096,1024] f64: 64 arrays from one copy\", \"ms\": %.5f, \"gbs\": %.1f}\n",
                ms, cb / ms / 1e6);
    ms = time_ms([&] { auto s = cube.slices(1); (void)s; }, 1, 5, 1);
    std::printf("{\"row\": \"slices(1) of [64,4096,1024] f64: 4096 arrays of [64,1024] from one permuting copy\", \"ms\": %.5f, \"gbs\": %.1f}\n",
                ms, cb / ms / 1e6);
  }
  float total = 0;
  ms = time_ms([&] { total = a.sum(); }, 2, 10, 1);
  std::printf("{\"row\": \"a.sum() incl. result read-back and flag check\", \"ms\": %.5f, \"gbs\": %.1f, \"ok\": %s}\n", ms,
              (double)N * 4 / ms / 1e6, (total > 1.5 * N * (1 - 1e-4) && total < 1.5 * N * (1 + 1e-4)) ? "true" : "false");
  {
    // per-axis fold incl. its raise point (the flag word read through the pinned record)
    auto m32 = DeviceNArray<float>::fill({16384, 16384}, 0.5f);
    ms = time_ms([&] { auto s = m32.sum(0); (void)s; }, 2, 10, 4);
    std::printf("{\"row\": \"sum(axis 0) of [16384,16384] f32 incl. the flag check\", \"ms\": %.5f, \"gbs\": %.1f}\n", ms,
                (16384.0 * 16384 * 4 + 16384 * 4) / ms / 1e6);
  }
  {
    // END TO END with HOST operands through the compiled host layer: pinned a, b, c -> a * b + c -> pinned result,
    // uploads, kernels and download inside every timed step (RowPipeline, ph_pipeline.hpp)
    PinnedArray<float> ha({R, Cc}), hc({R, Cc}), hb({1, Cc}), hout({R, Cc});
    for (int64_t i = 0; i < N; i++) { ha[i] = (float)((i * 2654435761u) % 2001) / 1000.0f - 1.0f; hc[i] = (float)((i * 40503u) % 1999) / 999.0f - 1.0f; }
    for (int64_t i = 0; i < Cc; i++) hb[i] = (float)((i * 7919u) % 2003) / 1001.0f - 1.0f;
    RowPipeline pipe;                                            // 4 chunks, the last halved 7x, the first 5x
    auto expr = [](const std::vector<DeviceNArray<float>>& in, const std::vector<DeviceNArray<float>>& shared) {
      return in[0].broadcast_op(PH_MUL, shared[0]) + in[1];
    };
    ms = time_ms([&] { pipe.map_rows<float>(expr, {&ha, &hc}, hout, {&hb}, false); }, 1, 3, 4);
    Device::sync();
    bool ok = true;
    for (int64_t i = 0; i < N && ok; i += 4099) ok = hout[i] == ha[i] * hb[i % Cc] + hc[i];       // two roundings, no FMA (-ffp-contract=off)
    std::printf("{\"row\": \"e2e: RowPipeline.map_rows(a * b + c), pinned host in / out, 16 ramped / tapered chunks\", \"ms\": %.5f, \"gbs\": %.1f, "
                "\"h2d_bytes\": %.0f, \"d2h_bytes\": %.0f, \"ok\": %s}\n", ms, bytes_two / ms / 1e6, 2.0 * N * 4 + Cc * 4, 1.0 * N * 4, ok ? "true" : "false");
  }
  std::printf("{\"kernel_launches\": %lld}\n", (long long)ph_launch_count());
  Device::shutdown();
  return 0;
}
