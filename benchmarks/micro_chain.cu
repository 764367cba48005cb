// micro_chain.cu -- throw-away microbenchmark: the two-kernel chain t = a*b[col]; out = t + c on
// 8192x8192 f32 with (a) same / opposite traversal direction and (b) L2 eviction-priority hints
// (inputs evict_first, outputs evict_last), to see how much of the 256 MiB temporary the 126 MB L2
// can carry from the producer to the consumer.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o micro_chain micro_chain.cu && ./micro_chain
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

struct alignas(32) V32 { float v[8]; };

template <int HINT>   // 0 none, 1 evict_first
__device__ __forceinline__ V32 ld(const float* p, uint64_t pol) {
  V32 r; uint64_t* q = reinterpret_cast<uint64_t*>(&r);
  if (HINT == 0) asm("ld.global.L1::no_allocate.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(q[0]), "=l"(q[1]), "=l"(q[2]), "=l"(q[3]) : "l"(p));
  else asm("ld.global.L1::no_allocate.L2::cache_hint.v4.b64 {%0,%1,%2,%3}, [%4], %5;" : "=l"(q[0]), "=l"(q[1]), "=l"(q[2]), "=l"(q[3]) : "l"(p), "l"(pol));
  return r;
}
template <int HINT>   // 0 none, 1 evict_last
__device__ __forceinline__ void st(float* p, const V32& r, uint64_t pol) {
  const uint64_t* q = reinterpret_cast<const uint64_t*>(&r);
  if (HINT == 0) asm volatile("st.global.L1::no_allocate.v4.b64 [%0], {%1,%2,%3,%4};" :: "l"(p), "l"(q[0]), "l"(q[1]), "l"(q[2]), "l"(q[3]) : "memory");
  else asm volatile("st.global.L2::cache_hint.v4.b64 [%0], {%1,%2,%3,%4}, %5;" :: "l"(p), "l"(q[0]), "l"(q[1]), "l"(q[2]), "l"(q[3]), "l"(pol) : "memory");
}

// MODE 0: out = x * b[col]   MODE 1: out = x + y
template <int MODE, int HINT>
__global__ void __launch_bounds__(256) k(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out,
                                         int64_t n, uint32_t cols, int reverse) {
  uint64_t pf = 0, pl = 0;
  if (HINT) {
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pf));
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pl));
  }
  const int64_t tile = 256 * 8 * 2;
  const int64_t t = reverse ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const int64_t base = t * tile + (int64_t)threadIdx.x * 8;
  V32 a[2], b[2];
#pragma unroll
  for (int u = 0; u < 2; u++) {
    const int64_t i = base + (int64_t)u * 256 * 8;
    a[u] = ld<HINT>(x + i, pf);
    if (MODE == 1) b[u] = ld<HINT>(y + i, pf);
    else b[u] = *reinterpret_cast<const V32*>(y + (uint32_t)((uint64_t)i % cols));
  }
#pragma unroll
  for (int u = 0; u < 2; u++) {
    V32 r;
#pragma unroll
    for (int j = 0; j < 8; j++) r.v[j] = MODE == 0 ? __fmul_rn(a[u].v[j], b[u].v[j]) : __fadd_rn(a[u].v[j], b[u].v[j]);
    st<HINT>(out + base + (int64_t)u * 256 * 8, r, pl);
  }
}

template <int HINT>
static void run(const char* name, float* a, float* b, float* c, float* t, float* out, int64_t n, bool alternate) {
  const int grid = (int)(n / (256 * 8 * 2));
  cudaEvent_t e0, e1, m0, m1; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&m0); cudaEventCreate(&m1);
  for (int i = 0; i < 5; i++) { k<0, HINT><<<grid, 256>>>(a, b, t, n, 8192, 0); k<1, HINT><<<grid, 256>>>(t, c, out, n, 8192, alternate); }
  const int reps = 200;
  cudaEventRecord(e0);
  for (int i = 0; i < reps; i++) { k<0, HINT><<<grid, 256>>>(a, b, t, n, 8192, 0); k<1, HINT><<<grid, 256>>>(t, c, out, n, 8192, alternate); }
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
  const double bytes = 5.0 * n * 4 + 8192 * 4;
  printf("%-34s hints %d alternate %d : %.4f ms per step  %.0f GB/s algorithmic\n", name, HINT, (int)alternate, ms, bytes / ms / 1e6);
}

int main() {
  const int64_t n = 8192LL * 8192;
  float *a, *b, *c, *t, *out;
  cudaMalloc(&a, n * 4); cudaMalloc(&c, n * 4); cudaMalloc(&t, n * 4); cudaMalloc(&out, n * 4); cudaMalloc(&b, 8192 * 4);
  cudaMemset(a, 0, n * 4); cudaMemset(c, 0, n * 4); cudaMemset(b, 0, 8192 * 4);
  run<0>("t=a*b ; out=t+c", a, b, c, t, out, n, false);
  run<0>("t=a*b ; out=t+c", a, b, c, t, out, n, true);
  run<1>("t=a*b ; out=t+c", a, b, c, t, out, n, false);
  run<1>("t=a*b ; out=t+c", a, b, c, t, out, n, true);
  return 0;
}
