// micro_stream.cu -- throw-away microbenchmark behind DESIGN.md's choices for the flat elementwise
// kernels: 1R1W (a*s), 2R1W (a+c) and row-vector broadcast (a*b[col]) on 8192x8192 f32, over
// (bytes per thread in flight) x (persistent or one tile per block) x (cache hints).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o micro_stream micro_stream.cu && ./micro_stream
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

struct alignas(32) V32 { float v[8]; };

__device__ __forceinline__ V32 ld_na(const float* p) {
  V32 r; uint64_t* q = reinterpret_cast<uint64_t*>(&r);
  asm("ld.global.L1::no_allocate.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(q[0]), "=l"(q[1]), "=l"(q[2]), "=l"(q[3]) : "l"(p));
  return r;
}
__device__ __forceinline__ V32 ld_ca(const float* p) {
  V32 r; uint64_t* q = reinterpret_cast<uint64_t*>(&r);
  asm("ld.global.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(q[0]), "=l"(q[1]), "=l"(q[2]), "=l"(q[3]) : "l"(p));
  return r;
}
template <int HINT>
__device__ __forceinline__ void st(float* p, const V32& r) {
  const uint64_t* q = reinterpret_cast<const uint64_t*>(&r);
  if (HINT == 0) asm volatile("st.global.L1::no_allocate.v4.b64 [%0], {%1,%2,%3,%4};" :: "l"(p), "l"(q[0]), "l"(q[1]), "l"(q[2]), "l"(q[3]) : "memory");
  else if (HINT == 1) asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" :: "l"(p), "l"(q[0]), "l"(q[1]), "l"(q[2]), "l"(q[3]) : "memory");
  else asm volatile("st.global.cs.v4.b64 [%0], {%1,%2,%3,%4};" :: "l"(p), "l"(q[0]), "l"(q[1]), "l"(q[2]), "l"(q[3]) : "memory");
}

// MODE 0: out = a*s   1: out = a + c   2: out = a * b[col] (b cached loads, period = cols)
template <int MODE, int U, int THREADS, bool PERSIST, int HINT>
__global__ void __launch_bounds__(THREADS) k(const float* __restrict__ a, const float* __restrict__ c,
                                             const float* __restrict__ b, float* __restrict__ out, int64_t n,
                                             uint32_t cols, float s) {
  const int64_t tile = (int64_t)THREADS * 8 * U;
  const int64_t ntiles = n / tile;
  for (int64_t t = blockIdx.x; t < ntiles; t += PERSIST ? gridDim.x : ntiles) {
    const int64_t base = t * tile + (int64_t)threadIdx.x * 8;
    V32 x[U], y[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int64_t i = base + (int64_t)u * THREADS * 8;
      x[u] = ld_na(a + i);
      if (MODE == 1) y[u] = ld_na(c + i);
      if (MODE == 2) y[u] = ld_ca(b + (uint32_t)((uint64_t)i % cols));
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      V32 r;
#pragma unroll
      for (int j = 0; j < 8; j++) r.v[j] = MODE == 0 ? __fmul_rn(x[u].v[j], s) : (MODE == 1 ? __fadd_rn(x[u].v[j], y[u].v[j]) : __fmul_rn(x[u].v[j], y[u].v[j]));
      st<HINT>(out + base + (int64_t)u * THREADS * 8, r);
    }
  }
}

template <int MODE, int U, int THREADS, bool PERSIST, int HINT>
static void run(const char* name, const float* a, const float* c, const float* b, float* out, int64_t n, int sms, int per_sm) {
  const int64_t tile = (int64_t)THREADS * 8 * U;
  const int64_t ntiles = n / tile;
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<MODE, U, THREADS, PERSIST, HINT>, THREADS, 0);
  const int grid = PERSIST ? sms * (per_sm > 0 ? per_sm : occ) : (int)ntiles;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 5; i++) k<MODE, U, THREADS, PERSIST, HINT><<<grid, THREADS>>>(a, c, b, out, n, 8192, 1.5f);
  const int reps = 50;
  cudaEventRecord(e0);
  for (int i = 0; i < reps; i++) k<MODE, U, THREADS, PERSIST, HINT><<<grid, THREADS>>>(a, c, b, out, n, 8192, 1.5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
  const double bytes = (MODE == 1 ? 3.0 : 2.0) * n * 4;
  printf("%-46s mode %d U %d thr %d %s hint %d occ %d grid %7d : %.4f ms  %.0f GB/s\n", name, MODE, U, THREADS,
         PERSIST ? "persist" : "tiled  ", HINT, occ, grid, ms, bytes / ms / 1e6);
}

int main() {
  const int64_t n = 8192LL * 8192;
  float *a, *c, *b, *out;
  cudaMalloc(&a, n * 4); cudaMalloc(&c, n * 4); cudaMalloc(&out, n * 4); cudaMalloc(&b, 8192 * 4);
  cudaMemset(a, 0, n * 4); cudaMemset(c, 0, n * 4); cudaMemset(b, 0, 8192 * 4);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
#define ALLMODES(U, T, P, H, PS) \
  run<0, U, T, P, H>("scale 1R1W", a, c, b, out, n, sms, PS); \
  run<1, U, T, P, H>("add 2R1W", a, c, b, out, n, sms, PS); \
  run<2, U, T, P, H>("mul rowvec (periodic flat)", a, c, b, out, n, sms, PS);
  ALLMODES(2, 256, false, 0, 0)
  ALLMODES(4, 256, false, 0, 0)
  ALLMODES(1, 256, false, 0, 0)
  ALLMODES(2, 512, false, 0, 0)
  ALLMODES(2, 256, false, 1, 0)
  ALLMODES(2, 256, false, 2, 0)
  ALLMODES(2, 256, true, 0, 0)
  ALLMODES(4, 256, true, 0, 0)
  ALLMODES(2, 256, true, 0, 4)
  ALLMODES(4, 256, true, 0, 2)
  ALLMODES(2, 512, true, 0, 0)
  ALLMODES(2, 1024, true, 0, 1)
  cudaMemcpy(out, a, n * 4, cudaMemcpyDeviceToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int i = 0; i < 50; i++) cudaMemcpyAsync(out, a, n * 4, cudaMemcpyDeviceToDevice);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("cudaMemcpy D2D 256 MiB: %.4f ms  %.0f GB/s\n", ms / 50, 2.0 * n * 4 / (ms / 50) / 1e6);
  return 0;
}
