// lat_probe.cu -- dependent-issue latencies that bound the serial covariance recursion (one lane per launch).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/lat_probe tools/lat_probe.cu && ./tools/lat_probe
// Prints cycles per operation of: a dependent DFMA chain, a dependent FFMA chain, the float64 reciprocal
// used by the filters (rcp.approx.ftz.f64 + cubic step), a shared-memory store->load round trip and
// mbarrier.try_wait on an already completed phase, each on ONE thread of an otherwise idle SM.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double rcp3(double a) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
  const double e = fma(-a, x, 1.0);
  return fma(x, fma(e, e, e), x);
}

__global__ void probe(double* out, long long* cyc, int iters) {
  __shared__ double sm[32];
  __shared__ unsigned long long bar;
  double a = 1.000001, b = 0.999999, c = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) a = fma(a, b, c);
  }
  long long t1 = clock64();
  float fa = 1.0001f, fb = 0.9999f, fc = 1e-6f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) fa = fmaf(fa, fb, fc);
  }
  long long t2 = clock64();
  double r = 3.0 + a * 1e-30;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) r = rcp3(r) + 2.5;
  }
  long long t3 = clock64();
  double s = a;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      *reinterpret_cast<volatile double*>(&sm[j]) = s;
      s = *reinterpret_cast<volatile double*>(&sm[j]) + 1.0;
    }
  }
  long long t4 = clock64();
  unsigned bar_a = static_cast<unsigned>(__cvta_generic_to_shared(&bar));
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_a) : "memory");  // phase 0 complete
  unsigned okc = 0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      unsigned ok;
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(ok) : "r"(bar_a), "r"(okc & 0u) : "memory");
      okc += ok;  // the next wait's operand depends on this one
    }
  }
  long long t5 = clock64();
  out[0] = a + fa + r + s + okc;
  cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8); cudaMalloc(&cyc, 5 * 8);
  const int iters = 20000;
  probe<<<1, 1>>>(out, cyc, iters);
  probe<<<1, 1>>>(out, cyc, iters);
  long long h[5];
  if (cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) { printf("CUDA error\n"); return 1; }
  printf("dependent DFMA            %.1f cycles\n", double(h[0]) / (16.0 * iters));
  printf("dependent FFMA            %.1f cycles\n", double(h[1]) / (16.0 * iters));
  printf("rcp64 (approx + cubic) +1 DADD  %.1f cycles\n", double(h[2]) / (4.0 * iters));
  printf("STS.64 -> LDS.64 -> DADD  %.1f cycles\n", double(h[3]) / (4.0 * iters));
  printf("mbarrier.try_wait (complete phase, dependent) %.1f cycles\n", double(h[4]) / (4.0 * iters));
  return 0;
}
