// heat_equation.cpp -- the reference's examples/heat_equation.cr on the device path, as a C++
// host program over the C-ABI (the reference is Crystal; no Crystal compiler exists in this
// image, so the host side of the example is written in C++; INTEGRATION.md section 6 shows the
// Crystal edit).  Same constants, same update rule, same number of steps:
//   examples/heat_equation.cr:5-20  constants          -> main()
//   examples/heat_equation.cr:22-24 initial state      -> NArray.fill + two element writes
//   examples/heat_equation.cr:26-36 simulate           -> ph_heat_run (ONE launch for 10 001 steps)
//   examples/heat_equation.cr:38-51 update_temp        -> PH_HEAT_EXAMPLE1D
// Build: make -C examples        Run: ./examples/heat_equation [steps]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../include/ph_gpu.h"

#define CHECK(call)                                                                   \
  do {                                                                                \
    int32_t st_ = (call);                                                             \
    if (st_ != PH_OK) {                                                               \
      std::fprintf(stderr, "%s failed (%d): %s\n", #call, st_, ph_last_error_string()); \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

int main(int argc, char** argv) {
  const double T_LEFT = 0.0, T_RIGHT = 100.0, T_INITIAL = 20.0;           // :5-7
  const int LENGTH = 1;                                                    // :9
  const double SPACING = 0.05, TIMESTEP = 0.01;                            // :10-11
  const int NUM_POINTS = (int)std::floor(LENGTH / SPACING) + 1;            // :12  (= 21)
  const int CONDUCTIVITY = 237, DENSITY = 2700, SPECIFIC_HEAT = 900;       // :14-16
  // :20  (CONDUCTIVITY * TIMESTEP) / (DENSITY * SPECIFIC_HEAT * (SPACING ** 2)); Float ** Int = powi
  const double COEFF = (CONDUCTIVITY * TIMESTEP) / ((double)(DENSITY * SPECIFIC_HEAT) * (SPACING * SPACING));
  const double duration = 100.0;
  long steps = (long)(duration / TIMESTEP) + 1;                            // :27  (= 10 001)
  if (argc > 1) steps = std::atol(argv[1]);

  CHECK(ph_init(0));
  const int64_t n = NUM_POINTS;
  ph_desc whole = {};
  whole.rank = 1; whole.extent[0] = n; whole.stride[0] = 1;
  void *state = nullptr, *other = nullptr;
  CHECK(ph_alloc(n * sizeof(double), &state));
  CHECK(ph_alloc(n * sizeof(double), &other));
  CHECK(ph_fill_region(8, state, &whole, &T_INITIAL));                     // NArray.fill([NUM_POINTS], T_INITIAL)
  ph_desc first = whole, last = whole;                                      // state[0] = T_LEFT; state[-1] = T_RIGHT
  first.extent[0] = 1;
  last.extent[0] = 1; last.offset = n - 1;
  CHECK(ph_fill_region(8, state, &first, &T_LEFT));
  CHECK(ph_fill_region(8, state, &last, &T_RIGHT));

  const int64_t ext[1] = {n};
  int32_t final_is_b = 0;
  CHECK(ph_heat_run(PH_F64, 1, ext, &COEFF, PH_HEAT_EXAMPLE1D, state, other, steps, &final_is_b));
  std::vector<double> host(n);
  CHECK(ph_d2h(host.data(), final_is_b ? other : state, n * sizeof(double)));
  uint32_t flags = 0;
  CHECK(ph_take_arith_flags(&flags));

  std::printf("%dx1 NArray(Float64) after %ld steps, COEFF = %a\n[", NUM_POINTS, steps, COEFF);
  double sum = 0;
  for (int64_t i = 0; i < n; i++) { std::printf("%s%.17g", i ? ", " : "", host[i]); sum += host[i]; }
  std::printf("]\nsum = %.12f  launches = %lld  flags = %u\n", sum, (long long)ph_launch_count(), flags);
  CHECK(ph_free(state));
  CHECK(ph_free(other));
  CHECK(ph_shutdown());
  return 0;
}
