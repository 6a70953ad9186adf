// stream_probe.cu -- read-streaming ceilings for the filter loop's access pattern on one GPU.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/stream_probe tools/stream_probe.cu
//   ./tools/stream_probe [N=131072] [T=2000]
//
// The payload of b200kf_run_batch is real[T, D, N] (time-major, filter fastest) and a CTA owns a
// contiguous slice of F filters, so per step it reads D rows of F*8 bytes, 8*N bytes apart.  This tool
// measures how fast that pattern can be pulled through (a) a cp.async.bulk + mbarrier ring with a
// dedicated producer warp (what kf_run_tma_kernel / kf_run_shared_kernel do), for several CTA sizes,
// ring depths and CTAs per SM; (b) plain per-thread LDG with U steps of register prefetch; and (c) a
// flat LDG.128 grid-stride read of the same bytes (the read-only HBM ceiling).  Every consumer does
// one DADD per value so the loads cannot be elided.  Prints GB/s per configuration.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) {                                                         \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                       \
    }                                                                                \
  } while (0)


__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c));
}
__device__ __forceinline__ void mbar_expect_tx_arrive(unsigned long long* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(
          smem_u32(b)),
      "r"(parity), "r"(0x100000u)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// (a) TMA ring: F consumer threads + one producer warp; STAGES x CHUNK steps x D rows of F doubles
template <int F, int STAGES, int CHUNK, int D>
__global__ void __launch_bounds__(F + 32) ring_kernel(const double* __restrict__ pay, long long N, int T, double* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  double* buf = reinterpret_cast<double*>(smem);
  constexpr size_t stage_reals = static_cast<size_t>(CHUNK) * D * F;
  unsigned long long* full = reinterpret_cast<unsigned long long*>(buf + STAGES * stage_reals);
  unsigned long long* empty = full + STAGES;
  const int tid = threadIdx.x;
  const long long base = static_cast<long long>(blockIdx.x) * F;
  const int nfb = static_cast<int>((N - base) < F ? (N - base) : F);
  const int nchunks = (T + CHUNK - 1) / CHUNK;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], F / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid >= F) {
    if (tid == F) {
      const unsigned row_bytes = nfb * 8;
      for (int c = 0; c < nchunks; ++c) {
        const int s = c % STAGES;
        if (c >= STAGES) mbar_wait(&empty[s], (c / STAGES - 1) & 1);
        const int k0 = c * CHUNK;
        const int steps = (T - k0) < CHUNK ? (T - k0) : CHUNK;
        mbar_expect_tx_arrive(&full[s], steps * D * row_bytes);
        for (int cc = 0; cc < steps; ++cc)
          for (int j = 0; j < D; ++j)
            bulk_g2s(buf + s * stage_reals + (cc * D + j) * F, pay + (static_cast<long long>(k0 + cc) * D + j) * N + base,
                     row_bytes, &full[s]);
      }
    }
    return;
  }
  double acc = 0.0;
  for (int c = 0; c < nchunks; ++c) {
    const int s = c % STAGES;
    mbar_wait(&full[s], (c / STAGES) & 1);
    const int k0 = c * CHUNK;
    const int steps = (T - k0) < CHUNK ? (T - k0) : CHUNK;
    if (tid < nfb) {
      const double* sb = buf + s * stage_reals + tid;
#pragma unroll
      for (int cc = 0; cc < CHUNK; ++cc)
        if (cc < steps)
#pragma unroll
          for (int j = 0; j < D; ++j) acc += sb[(cc * D + j) * F];
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&empty[s]);
  }
  if (tid < nfb) out[base + tid] = acc;
}

// (b) per-thread LDG, U steps in flight
template <int U, int D>
__global__ void ldg_kernel(const double* __restrict__ pay, long long N, int T, double* out) {
  const long long f = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (f >= N) return;
  const double* p = pay + f;
  double acc = 0.0;
  double buf[U][D];
#pragma unroll
  for (int u = 0; u < U; ++u)
#pragma unroll
    for (int j = 0; j < D; ++j) buf[u][j] = (u < T) ? __ldcs(p + (static_cast<long long>(u) * D + j) * N) : 0.0;
  for (int k0 = 0; k0 < T; k0 += U) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k = k0 + u;
      if (k < T) {
#pragma unroll
        for (int j = 0; j < D; ++j) acc += buf[u][j];
        if (k + U < T) {
#pragma unroll
          for (int j = 0; j < D; ++j) buf[u][j] = __ldcs(p + (static_cast<long long>(k + U) * D + j) * N);
        }
      }
    }
  }
  out[f] = acc;
}

// (c) flat read
__global__ void flat_kernel(const double2* __restrict__ p, long long n2, double* out) {
  double acc = 0.0;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  for (; i + 3 * stride < n2; i += 4 * stride) {
    const double2 a = __ldcs(p + i), b = __ldcs(p + i + stride), c = __ldcs(p + i + 2 * stride), d = __ldcs(p + i + 3 * stride);
    acc += (a.x + a.y) + (b.x + b.y) + (c.x + c.y) + (d.x + d.y);
  }
  for (; i < n2; i += stride) {
    const double2 a = __ldcs(p + i);
    acc += a.x + a.y;
  }
  if (acc == 123.456) out[0] = acc;
}

__global__ void fill_kernel(double* p, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += stride)
    p[i] = 1e-3 * static_cast<double>(i & 1023);
}

template <class L>
static double time_ms(L&& launch, int reps = 5) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; ++r) launch();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  CK(cudaGetLastError());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}

template <int F, int STAGES, int CHUNK, int D = 3>
static void run_ring(const double* pay, long long N, int T, double* out, double bytes) {
  constexpr size_t smem = static_cast<size_t>(STAGES) * CHUNK * D * F * 8 + 2 * STAGES * 8;
  if (smem > 227 * 1024) return;
  CK(cudaFuncSetAttribute(ring_kernel<F, STAGES, CHUNK, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const unsigned blocks = static_cast<unsigned>((N + F - 1) / F);
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ring_kernel<F, STAGES, CHUNK, D>, F + 32, smem));
  const double ms = time_ms([&] { ring_kernel<F, STAGES, CHUNK, D><<<blocks, F + 32, smem>>>(pay, N, T, out); });
  printf("ring  D=%d F=%4d stages=%2d chunk=%d  smem=%6zu B  ctas=%4u (%d/SM)  %8.3f ms  %7.1f GB/s\n", D, F, STAGES, CHUNK, smem,
         blocks, occ, ms, bytes / ms / 1e6);
}

template <int U, int D = 3>
static void run_ldg(const double* pay, long long N, int T, double* out, double bytes, int tpb) {
  const unsigned blocks = static_cast<unsigned>((N + tpb - 1) / tpb);
  const double ms = time_ms([&] { ldg_kernel<U, D><<<blocks, tpb>>>(pay, N, T, out); });
  printf("ldg   U=%2d tpb=%4d  ctas=%5u  %8.3f ms  %7.1f GB/s\n", U, tpb, blocks, ms, bytes / ms / 1e6);
}

template <int D>
static void sweep(long long N, int T) {
  const long long total = static_cast<long long>(T) * D * N;
  const double bytes = static_cast<double>(total) * 8;
  double *pay, *out;
  CK(cudaMalloc(&pay, total * 8));
  CK(cudaMalloc(&out, N * 8));
  fill_kernel<<<148 * 8, 256>>>(pay, total);
  CK(cudaDeviceSynchronize());
  printf("payload [T=%d, D=%d, N=%lld] f64 = %.2f GB\n", T, D, N, bytes / 1e9);
  {
    const double ms = time_ms([&] { flat_kernel<<<148 * 16, 512>>>(reinterpret_cast<const double2*>(pay), total / 2, out); });
    printf("flat  LDG.128 grid-stride                      %8.3f ms  %7.1f GB/s\n", ms, bytes / ms / 1e6);
  }
  if constexpr (D == 3) {
    run_ring<896, 5, 2, D>(pay, N, T, out, bytes);
    run_ring<896, 3, 3, D>(pay, N, T, out, bytes);
    run_ring<896, 10, 1, D>(pay, N, T, out, bytes);
    run_ring<896, 2, 5, D>(pay, N, T, out, bytes);
    run_ring<448, 8, 2, D>(pay, N, T, out, bytes);
    run_ring<448, 4, 2, D>(pay, N, T, out, bytes);
    run_ring<448, 4, 4, D>(pay, N, T, out, bytes);
    run_ring<224, 4, 4, D>(pay, N, T, out, bytes);
    run_ring<128, 4, 4, D>(pay, N, T, out, bytes);
    run_ring<1024 - 32, 4, 2, D>(pay, N, T, out, bytes);
    run_ring<512, 6, 2, D>(pay, N, T, out, bytes);
    run_ring<256, 6, 2, D>(pay, N, T, out, bytes);
  } else {
    run_ring<896, 3, 1, D>(pay, N, T, out, bytes);
    run_ring<896, 2, 1, D>(pay, N, T, out, bytes);
    run_ring<448, 6, 1, D>(pay, N, T, out, bytes);
    run_ring<448, 3, 1, D>(pay, N, T, out, bytes);
    run_ring<448, 3, 2, D>(pay, N, T, out, bytes);
    run_ring<224, 6, 1, D>(pay, N, T, out, bytes);
    run_ring<224, 12, 1, D>(pay, N, T, out, bytes);
    run_ring<224, 6, 2, D>(pay, N, T, out, bytes);
    run_ring<128, 5, 1, D>(pay, N, T, out, bytes);
    run_ring<128, 10, 2, D>(pay, N, T, out, bytes);
    run_ring<640, 4, 1, D>(pay, N, T, out, bytes);
  }
  for (int tpb : {128, 256, 896}) {
    run_ldg<1, D>(pay, N, T, out, bytes, tpb);
    run_ldg<2, D>(pay, N, T, out, bytes, tpb);
    run_ldg<4, D>(pay, N, T, out, bytes, tpb);
  }
  CK(cudaFree(pay));
  CK(cudaFree(out));
}

int main(int argc, char** argv) {
  const long long N = argc > 1 ? atoll(argv[1]) : 131072;
  const int T = argc > 2 ? atoi(argv[2]) : 2000;
  const int d = argc > 3 ? atoi(argv[3]) : 3;
  if (d == 9) sweep<9>(N, T);
  else if (d == 4) sweep<4>(N, T);
  else sweep<3>(N, T);
  return 0;
}
