// kf_api.cu -- the extern "C" surface of libb200kf.so (include/b200kf.h): argument checking,
// model objects, dispatch to the per-model kernels, the host-buffer streaming path, and the
// small model-independent kernels (schedule-search finaliser, batched cov_matrix, synthetic
// payload generator, FMA-peak probe).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

#include <mutex>
#include <thread>
#include <vector>

#include "kf_internal.h"

namespace b200kf {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) in %s", static_cast<int>(e),
           cudaGetErrorString(e), what);
  return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? B200KF_ERR_NO_DEVICE
                                                                      : B200KF_ERR_CUDA;
}

struct ModelInfo {
  int n, d, m_gps;
};
static const ModelInfo kInfo[MODEL_COUNT] = {{6, 3, 3}, {8, 4, 2}, {9, 3, 3}, {15, 9, 3}, {6, 3, 3}};

static int dispatch_run(const b200kf_model* m, const RunParams& p, unsigned flags, cudaStream_t st) {
  switch (m->desc.model_id) {
    case MODEL_CV6: return launch_run_m<MODEL_CV6>(m, p, flags, st);
    case MODEL_CA8: return launch_run_m<MODEL_CA8>(m, p, flags, st);
    case MODEL_BIAS9: return launch_run_m<MODEL_BIAS9>(m, p, flags, st);
    case MODEL_CA15: return launch_run_m<MODEL_CA15>(m, p, flags, st);
    case MODEL_CVBU6: return launch_run_m<MODEL_CVBU6>(m, p, flags, st);
  }
  return set_error(B200KF_ERR_UNSUPPORTED, "unknown model id %d", m->desc.model_id);
}

static int dispatch_score(const b200kf_model* m, const ScoreParams& p, unsigned flags,
                          b200kf_best* best, cudaStream_t st) {
  switch (m->desc.model_id) {
    case MODEL_CV6: return launch_score_m<MODEL_CV6>(m, p, flags, best, st);
    case MODEL_CA8: return launch_score_m<MODEL_CA8>(m, p, flags, best, st);
    case MODEL_BIAS9: return launch_score_m<MODEL_BIAS9>(m, p, flags, best, st);
    case MODEL_CA15: return launch_score_m<MODEL_CA15>(m, p, flags, best, st);
    case MODEL_CVBU6: return launch_score_m<MODEL_CVBU6>(m, p, flags, best, st);
  }
  return set_error(B200KF_ERR_UNSUPPORTED, "unknown model id %d", m->desc.model_id);
}

static int dispatch_masked(const b200kf_model* m, const MaskedParams& p, unsigned flags, cudaStream_t st) {
  switch (m->desc.model_id) {
    case MODEL_CV6: return launch_masked_m<MODEL_CV6>(m, p, flags, st);
    case MODEL_CA8: return launch_masked_m<MODEL_CA8>(m, p, flags, st);
    case MODEL_BIAS9: return launch_masked_m<MODEL_BIAS9>(m, p, flags, st);
    case MODEL_CA15: return launch_masked_m<MODEL_CA15>(m, p, flags, st);
    case MODEL_CVBU6: return launch_masked_m<MODEL_CVBU6>(m, p, flags, st);
  }
  return set_error(B200KF_ERR_UNSUPPORTED, "unknown model id %d", m->desc.model_id);
}

static int dispatch_product(const b200kf_model* m, const ProductParams& p, unsigned flags, cudaStream_t st) {
  switch (m->desc.model_id) {
    case MODEL_CV6: return launch_product_m<MODEL_CV6>(m, p, flags, st);
    case MODEL_CA8: return launch_product_m<MODEL_CA8>(m, p, flags, st);
    case MODEL_BIAS9: return launch_product_m<MODEL_BIAS9>(m, p, flags, st);
    case MODEL_CA15: return launch_product_m<MODEL_CA15>(m, p, flags, st);
    case MODEL_CVBU6: return launch_product_m<MODEL_CVBU6>(m, p, flags, st);
  }
  return set_error(B200KF_ERR_UNSUPPORTED, "unknown model id %d", m->desc.model_id);
}

static int dispatch_gains(const b200kf_model* m, const RunParams& p, cudaStream_t st) {
  switch (m->desc.model_id) {
    case MODEL_CV6: return launch_gains_m<MODEL_CV6>(m, p, st);
    case MODEL_CA8: return launch_gains_m<MODEL_CA8>(m, p, st);
    case MODEL_BIAS9: return launch_gains_m<MODEL_BIAS9>(m, p, st);
    case MODEL_CA15: return launch_gains_m<MODEL_CA15>(m, p, st);
    case MODEL_CVBU6: return launch_gains_m<MODEL_CVBU6>(m, p, st);
  }
  return set_error(B200KF_ERR_UNSUPPORTED, "unknown model id %d", m->desc.model_id);
}

static int dispatch_select(const b200kf_model* m, const SelectParams& p, unsigned flags, cudaStream_t st) {
  switch (m->desc.model_id) {
    case MODEL_CV6: return launch_select_m<MODEL_CV6>(m, p, flags, st);
    case MODEL_CA8: return launch_select_m<MODEL_CA8>(m, p, flags, st);
    case MODEL_BIAS9: return launch_select_m<MODEL_BIAS9>(m, p, flags, st);
    case MODEL_CA15: return launch_select_m<MODEL_CA15>(m, p, flags, st);
    case MODEL_CVBU6: return launch_select_m<MODEL_CVBU6>(m, p, flags, st);
  }
  return set_error(B200KF_ERR_UNSUPPORTED, "unknown model id %d", m->desc.model_id);
}

// ---- schedule-search finaliser: reduce the per-block partials to one b200kf_best -------------
__device__ __forceinline__ void score_partial_merge(ScorePartial& a, const ScorePartial& b) {
  if (b.score < a.score || (b.score == a.score && b.index < a.index)) {
    a.score = b.score;
    a.index = b.index;
  }
  a.key = b.key < a.key ? b.key : a.key;
  a.n_feasible += b.n_feasible;
  a.n_near += b.n_near;
}

// Launched with programmatic stream serialisation behind the last scoring kernel: the block is resident and waiting
// at the grid dependency when that kernel drains, so the search pays the reduce (warp shuffles, one shared-memory
// hand-off) but not a second launch latency -- a 2^20 search is 40 us in all.
__global__ void __launch_bounds__(256) score_final_kernel(const ScorePartial* parts, int n_parts, int L, b200kf_best* out) {
  cudaGridDependencySynchronize();
  __shared__ ScorePartial sh[8];
  ScorePartial r{INFINITY, ~0ULL, ~0ULL, 0ULL, 0ULL};
  for (int i = threadIdx.x; i < n_parts; i += blockDim.x) score_partial_merge(r, parts[i]);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    ScorePartial o;
    o.score = __shfl_down_sync(0xffffffffu, r.score, off);
    o.index = __shfl_down_sync(0xffffffffu, r.index, off);
    o.key = __shfl_down_sync(0xffffffffu, r.key, off);
    o.n_feasible = __shfl_down_sync(0xffffffffu, r.n_feasible, off);
    o.n_near = __shfl_down_sync(0xffffffffu, r.n_near, off);
    score_partial_merge(r, o);
  }
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = r;
  __syncthreads();
  if (threadIdx.x == 0) {
    ScorePartial a = sh[0];
    for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w) score_partial_merge(a, sh[w]);
    out->argmin_sum_score = a.score;
    out->argmin_sum_index = a.index;
    out->first_feasible_key = a.key;
    out->n_feasible = a.n_feasible;
    out->n_near_threshold = a.n_near;
    if (a.key == ~0ULL) {
      out->first_feasible_index = ~0ULL;
    } else {
      const unsigned long long full = (1ULL << L) - 1ULL;
      const unsigned long long rev = full - (a.key & 0xffffffffULL);
      out->first_feasible_index = __brevll(rev) >> (64 - L);
    }
  }
}

// ---- product-search finaliser ------------------------------------------------------------------
__global__ void product_final_kernel(const ProductPartial* parts, int n_parts, unsigned long long count,
                                     b200kf_best_product* out) {
  __shared__ ProductPartial sh[256];
  ProductPartial r{INFINITY, ~0ULL, INFINITY};
  auto merge = [](ProductPartial& a, const ProductPartial& b) {
    const bool take = b.score < a.score || (b.score == a.score && b.index < a.index);
    const double loser = take ? a.score : b.score;
    a.runner_up = fmin(fmin(a.runner_up, b.runner_up), loser);
    if (take) {
      a.score = b.score;
      a.index = b.index;
    }
  };
  for (int i = threadIdx.x; i < n_parts; i += blockDim.x) merge(r, parts[i]);
  sh[threadIdx.x] = r;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) merge(sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out->score = sh[0].score;
    out->index = sh[0].index;
    out->runner_up_score = sh[0].runner_up;
    out->n_scored = count;
  }
}

// ---- accuracy metric (kf_workers.py:1162-1216, KF_SensorFusion.ipynb:766-802) ------------------
// interp1d(kind='linear', fill_value='extrapolate') on a sorted track, evaluated as scipy does
// (_call_linear): hi = clip(searchsorted(t, x, side='left'), 1, n-1), lo = hi - 1,
// y = ((x - t_lo) / (t_hi - t_lo)) * y_hi + ((t_hi - x) / (t_hi - t_lo)) * y_lo, every operation
// rounded separately (no FMA contraction) so the result is bit-identical to numpy's.
__global__ void track_lerp_kernel(long long n_track, const double* __restrict__ tt, const double* __restrict__ tp,
                                  int k, long long n_query, const double* __restrict__ qt, double* __restrict__ out) {
  const long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (q >= n_query) return;
  const double x = qt[q];
  if (n_track == 1) {
    for (int c = 0; c < k; ++c) out[q * k + c] = tp[c];
    return;
  }
  long long lo = 0, hi = n_track;  // first index with tt[i] >= x
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (tt[mid] < x) lo = mid + 1;
    else hi = mid;
  }
  long long ih = lo < 1 ? 1 : (lo > n_track - 1 ? n_track - 1 : lo);
  const long long il = ih - 1;
  const double w = __dsub_rn(tt[ih], tt[il]);
  const double w1 = __ddiv_rn(__dsub_rn(x, tt[il]), w), w0 = __ddiv_rn(__dsub_rn(tt[ih], x), w);
  for (int c = 0; c < k; ++c)
    out[q * k + c] = __dadd_rn(__dmul_rn(w1, tp[ih * k + c]), __dmul_rn(w0, tp[il * k + c]));
}

// Partial sums of |x[0:3] - ref|^2 over one time segment for every filter of a time-major
// trajectory [T, ncomp, N] (coalesced over filters); segments are summed in order by the finaliser
// so the result does not depend on scheduling.
template <typename real>
__global__ void rmse_partial_kernel(const real* __restrict__ traj, int ncomp, long long N, long long T,
                                    const double* __restrict__ ref, long long seg_len, double* __restrict__ part) {
  const long long f = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (f >= N) return;
  const long long k0 = blockIdx.y * seg_len;
  const long long k1 = (k0 + seg_len < T) ? k0 + seg_len : T;
  double sse = 0.0;
  for (long long k = k0; k < k1; ++k) {
    const real* row = traj + (k * ncomp) * N + f;
    const double dx = double(row[0]) - ref[k * 3 + 0];
    const double dy = double(row[N]) - ref[k * 3 + 1];
    const double dz = double(row[2 * N]) - ref[k * 3 + 2];
    const double nrm = sqrt(dx * dx + dy * dy + dz * dz);  // np.linalg.norm, then **2 (:1207-1208)
    sse += nrm * nrm;
  }
  part[blockIdx.y * N + f] = sse;
}
__global__ void rmse_final_kernel(const double* part, int n_seg, long long N, long long T, double* rmse) {
  const long long f = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (f >= N) return;
  double s = 0.0;
  for (int g = 0; g < n_seg; ++g) s += part[g * N + f];
  rmse[f] = sqrt(s / static_cast<double>(T));
}

// ---- batched Scheduler.cov_matrix (kf_workers.py:112-147) ------------------------------------
struct CovParams {
  int n, k;
  double H[kMaxN * kMaxN];  // row-selected H^  [k, n]
  double R[kMaxN * kMaxN];  // row/col-selected R^ [k, k]
};

__global__ void cov_update_kernel(const CovParams p, long long batch, const double* __restrict__ Sigma,
                                  double* __restrict__ post, double* __restrict__ trace) {
  const long long b = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (b >= batch) return;
  const int n = p.n, k = p.k;
  const double* S0 = Sigma + b * n * n;
  double A[kMaxN * kMaxN];   // H^ Sigma            [k, n]
  double Bm[kMaxN * kMaxN];  // Sigma H^^T          [n, k]
  double S[kMaxN * kMaxN];   // R^ + H^ Sigma H^^T  [k, k]
  double Si[kMaxN * kMaxN];  // inverse
  for (int r = 0; r < k; ++r)
    for (int c = 0; c < n; ++c) {
      double a = 0, bb = 0;
      for (int j = 0; j < n; ++j) {
        a += p.H[r * n + j] * S0[j * n + c];
        bb += S0[c * n + j] * p.H[r * n + j];
      }
      A[r * n + c] = a;
      Bm[c * k + r] = bb;
    }
  for (int r = 0; r < k; ++r)
    for (int c = 0; c < k; ++c) {
      double a = p.R[r * k + c];
      for (int j = 0; j < n; ++j) a += p.H[r * n + j] * Bm[j * k + c];
      S[r * k + c] = a;
      Si[r * k + c] = (r == c) ? 1.0 : 0.0;
    }
  // Gauss-Jordan with partial pivoting (numpy's inv is LU with partial pivoting)
  for (int c = 0; c < k; ++c) {
    int piv = c;
    double best = fabs(S[c * k + c]);
    for (int r = c + 1; r < k; ++r)
      if (fabs(S[r * k + c]) > best) {
        best = fabs(S[r * k + c]);
        piv = r;
      }
    if (piv != c)
      for (int j = 0; j < k; ++j) {
        double t = S[c * k + j];
        S[c * k + j] = S[piv * k + j];
        S[piv * k + j] = t;
        t = Si[c * k + j];
        Si[c * k + j] = Si[piv * k + j];
        Si[piv * k + j] = t;
      }
    const double inv = 1.0 / S[c * k + c];
    for (int j = 0; j < k; ++j) {
      S[c * k + j] *= inv;
      Si[c * k + j] *= inv;
    }
    for (int r = 0; r < k; ++r)
      if (r != c) {
        const double f = S[r * k + c];
        if (f != 0.0)
          for (int j = 0; j < k; ++j) {
            S[r * k + j] -= f * S[c * k + j];
            Si[r * k + j] -= f * Si[c * k + j];
          }
      }
  }
  // K = Bm Si [n, k];  post = Sigma - K A
  double tr = 0;
  double* out = post + b * n * n;
  for (int r = 0; r < n; ++r) {
    double Krow[kMaxN];
    for (int c = 0; c < k; ++c) {
      double a = 0;
      for (int j = 0; j < k; ++j) a += Bm[r * k + j] * Si[j * k + c];
      Krow[c] = a;
    }
    for (int c = 0; c < n; ++c) {
      double a = 0;
      for (int j = 0; j < k; ++j) a += Krow[j] * A[j * n + c];
      const double v = S0[r * n + c] - a;
      out[r * n + c] = v;
      if (r == c) tr += v;
    }
  }
  if (trace != nullptr) trace[b] = tr;
}

// ---- synthetic payload (BASELINE config 3/5 shape) --------------------------------------------
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

template <typename real>
__global__ void fill_payload_kernel(real* __restrict__ out, const uint8_t* __restrict__ types,
                                    long long N, long long T, int d, int m_gps,
                                    unsigned long long seed) {
  const long long total = T * d * N;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / N;  // k * d + c
    const int c = static_cast<int>(row % d);
    const long long k = row / d;
    const bool gps = types[k] == EV_GPS;
    const double sigma = gps ? (c < m_gps ? 1.7320508075688772 : 0.0) : 0.1;
    const unsigned long long h = mix64(seed ^ mix64(static_cast<unsigned long long>(i)));
    const double u1 = (static_cast<double>(h >> 40) + 0.5) * (1.0 / 16777216.0);
    const double u2 = (static_cast<double>((h >> 16) & 0xFFFFFF) + 0.5) * (1.0 / 16777216.0);
    const double g = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    out[i] = static_cast<real>(sigma * g);
  }
}

// ---- FMA peak probe ----------------------------------------------------------------------------
template <typename real>
__global__ void fma_probe_kernel(real* out, int iters, real a, real b) {
  real v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5,
       v6 = v0 + 6, v7 = v0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v0 = fma(v0, a, b);
      v1 = fma(v1, a, b);
      v2 = fma(v2, a, b);
      v3 = fma(v3, a, b);
      v4 = fma(v4, a, b);
      v5 = fma(v5, a, b);
      v6 = fma(v6, a, b);
      v7 = fma(v7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7;
}

// Same FMA count, but every FMA reads three DISTINCT, changing registers (no operand-reuse cache
// hits): the register-file-bandwidth-limited rate a real filter step sees.
template <typename real>
__global__ void fma_probe_distinct_kernel(real* out, int iters, real a) {
  real v0 = threadIdx.x * real(1e-3), v1 = v0 + a, v2 = v0 - a, v3 = v0 * a, v4 = v0 + 2, v5 = v0 + 3,
       v6 = v0 + 4, v7 = v0 + 5;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v0 = fma(v1, v2, v0);
      v1 = fma(v2, v3, v1);
      v2 = fma(v3, v4, v2);
      v3 = fma(v4, v5, v3);
      v4 = fma(v5, v6, v4);
      v5 = fma(v6, v7, v5);
      v6 = fma(v7, v0, v6);
      v7 = fma(v0, v1, v7);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7;
}

template <typename real>
static int probe_peak(double* tflops, bool distinct) {
  int dev = 0, sms = 0;
  B200KF_CUDA_TRY(cudaGetDevice(&dev));
  B200KF_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int tpb = 256, blocks = sms * 8, iters = 4096;
  real* out = nullptr;
  B200KF_CUDA_TRY(cudaMalloc(&out, sizeof(real) * tpb * blocks));
  cudaEvent_t e0, e1;
  B200KF_CUDA_TRY(cudaEventCreate(&e0));
  B200KF_CUDA_TRY(cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    B200KF_CUDA_TRY(cudaEventRecord(e0));
    if (distinct)
      fma_probe_distinct_kernel<real><<<blocks, tpb>>>(out, iters, real(1e-9));
    else
      fma_probe_kernel<real><<<blocks, tpb>>>(out, iters, real(0.999999), real(1e-6));
    B200KF_CUDA_TRY(cudaEventRecord(e1));
    B200KF_CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0;
    B200KF_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 64.0 * iters * tpb * static_cast<double>(blocks);
    const double tf = flops / (ms * 1e-3) * 1e-12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  return B200KF_OK;
}

}  // namespace b200kf

using namespace b200kf;

extern "C" {

int b200kf_version(int* major, int* minor) {
  if (major) *major = B200KF_VERSION_MAJOR;
  if (minor) *minor = B200KF_VERSION_MINOR;
  return B200KF_OK;
}

size_t b200kf_last_error(char* buf, size_t n) {
  const size_t len = strlen(g_err);
  if (buf && n) {
    strncpy(buf, g_err, n - 1);
    buf[n - 1] = 0;
  }
  return len;
}

int b200kf_device_count(int* count) {
  if (!count) return set_error(B200KF_ERR_INVALID, "count is null");
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) {
    cudaGetLastError();
    *count = 0;
    return B200KF_OK;
  }
  *count = c;
  return B200KF_OK;
}

int b200kf_model_default_desc(int32_t model_id, int32_t precision, b200kf_model_desc* desc) {
  if (!desc) return set_error(B200KF_ERR_INVALID, "desc is null");
  if (model_id < 0 || model_id >= MODEL_COUNT)
    return set_error(B200KF_ERR_UNSUPPORTED, "unknown model id %d", model_id);
  memset(desc, 0, sizeof(*desc));
  desc->model_id = model_id;
  desc->precision = precision;
  desc->accel_coef = 0.5;
  for (int i = 0; i < 4; ++i) desc->r_gps[i] = 3.0;  // kf_workers.py:583, hw5_2.py:282
  auto set3 = [](double* a, int at, double v) { a[at] = a[at + 1] = a[at + 2] = v; };
  switch (model_id) {
    case MODEL_CV6:
      set3(desc->q, 0, 5.0), set3(desc->q, 3, 1.0);
      set3(desc->r_imu, 0, 50.0), set3(desc->r_imu, 3, 10.0);
      break;
    case MODEL_CA8: {  // hw5_2.py:235-249, 288-303
      const double q[8] = {5, 5, 0.05, 1, 1, 0.1, 2, 2};
      const double r[8] = {50, 50, 0.05, 10, 10, 0.1, 100, 100};
      memcpy(desc->q, q, sizeof(q));
      memcpy(desc->r_imu, r, sizeof(r));
      break;
    }
    case MODEL_BIAS9:
      set3(desc->q, 0, 5.0), set3(desc->q, 3, 1.0), set3(desc->q, 6, 2.0);
      break;
    case MODEL_CVBU6:  // CV6's process noise; the IMU event is a control input: no IMU variances
      set3(desc->q, 0, 5.0), set3(desc->q, 3, 1.0);
      break;
    case MODEL_CA15:  // kf_workers.py:521-543, 589-613
      set3(desc->q, 0, 5.0), set3(desc->q, 3, 0.05), set3(desc->q, 6, 1.0), set3(desc->q, 9, 0.1),
          set3(desc->q, 12, 2.0);
      set3(desc->r_imu, 0, 50.0), set3(desc->r_imu, 3, 0.05), set3(desc->r_imu, 6, 10.0),
          set3(desc->r_imu, 9, 0.1), set3(desc->r_imu, 12, 100.0);
      break;
  }
  return B200KF_OK;
}

// One stream-ordered scratch pool per DEVICE, created on first use and kept for the life of the process
// (release threshold = max: a steady-state call sequence never reaches the driver's allocator).  It is shared by
// every model on the device -- stream-ordered allocations of different models / streams never alias -- and is
// never destroyed: destroying a per-model pool with frees still in flight (an Engine dropped right after an
// asynchronous launch) crashed the process on the GPU box.
static cudaError_t scratch_pool_for(int device, cudaMemPool_t* out) {
  static std::mutex mu;
  static cudaMemPool_t pools[64] = {};
  if (device < 0 || device >= 64) return cudaErrorInvalidDevice;
  std::lock_guard<std::mutex> lk(mu);
  if (pools[device] == nullptr) {
    cudaMemPoolProps props;
    memset(&props, 0, sizeof(props));
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    cudaMemPool_t pool = nullptr;
    cudaError_t e = cudaMemPoolCreate(&pool, &props);
    if (e != cudaSuccess) return e;
    unsigned long long keep = ~0ULL;
    e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    if (e != cudaSuccess) return e;
    pools[device] = pool;
  }
  *out = pools[device];
  return cudaSuccess;
}

int b200kf_model_create(const b200kf_model_desc* desc, b200kf_model** out) {
  if (!desc || !out) return set_error(B200KF_ERR_INVALID, "null argument");
  if (desc->model_id < 0 || desc->model_id >= MODEL_COUNT)
    return set_error(B200KF_ERR_UNSUPPORTED, "unknown model id %d", desc->model_id);
  if (desc->precision != B200KF_F64 && desc->precision != B200KF_F32)
    return set_error(B200KF_ERR_UNSUPPORTED, "unknown precision %d", desc->precision);
  const ModelInfo info = kInfo[desc->model_id];
  for (int i = 0; i < info.n; ++i) {
    if (!(desc->q[i] >= 0.0)) return set_error(B200KF_ERR_INVALID, "q[%d] must be >= 0", i);
    if (!(desc->r_imu[i] >= 0.0)) return set_error(B200KF_ERR_INVALID, "r_imu[%d] must be >= 0", i);
  }
  for (int i = 0; i < info.m_gps; ++i)
    if (!(desc->r_gps[i] > 0.0)) return set_error(B200KF_ERR_INVALID, "r_gps[%d] must be > 0", i);
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return set_error(B200KF_ERR_NO_DEVICE,
                     "no CUDA device: libb200kf has no CPU fallback (cudaGetDeviceCount: %s)",
                     e == cudaSuccess ? "0 devices" : cudaGetErrorString(e));
  }
  b200kf_model* m = new (std::nothrow) b200kf_model();
  if (!m) return set_error(B200KF_ERR_INVALID, "out of host memory");
  m->desc = *desc;
  memcpy(m->mc.q, desc->q, sizeof(m->mc.q));
  memcpy(m->mc.r_gps, desc->r_gps, sizeof(m->mc.r_gps));
  memcpy(m->mc.r_imu, desc->r_imu, sizeof(m->mc.r_imu));
  m->mc.c2 = desc->accel_coef;
  m->n = info.n, m->d = info.d, m->m_gps = info.m_gps;
  m->pool = nullptr;
  m->host_ws.ready = 0;
  m->host_ws.slab = nullptr;
  m->host_ws.cap = 0;
  m->host_ws.h_stage[0] = m->host_ws.h_stage[1] = nullptr;
  m->host_ws.h_cap = 0;
  m->tpb = 64;
  m->use_tma = 1;
  if (const char* s = getenv("B200KF_TMA")) m->use_tma = atoi(s) != 0;
  m->use_select_warp = 1;
  if (const char* s = getenv("B200KF_SELECT_WARP")) m->use_select_warp = atoi(s) != 0;
  m->use_axis_dense = 1;
  if (const char* s = getenv("B200KF_AXIS_DENSE")) m->use_axis_dense = atoi(s) != 0;
  m->use_ca_ring = 1;
  if (const char* s = getenv("B200KF_CA_RING")) m->use_ca_ring = atoi(s) != 0;
  m->use_bulk_out = 1;
  if (const char* s = getenv("B200KF_BULK_OUT")) m->use_bulk_out = atoi(s) != 0;
  m->use_sym = 1;
  if (const char* s = getenv("B200KF_SYM")) m->use_sym = atoi(s) != 0;
  m->ring_f_cap = 896;
  if (const char* s = getenv("B200KF_RING_F")) m->ring_f_cap = atoi(s);
  m->use_tree = 1;
  if (const char* s = getenv("B200KF_TREE")) m->use_tree = atoi(s) != 0;
  m->use_shared = 1;
  if (const char* s = getenv("B200KF_SHARED")) m->use_shared = atoi(s) != 0;
  if (const char* s = getenv("B200KF_TPB")) {
    const int v = atoi(s);
    if (v >= 32 && v <= 256 && v % 32 == 0) m->tpb = v;
  }
  e = cudaGetDevice(&m->device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, m->device);
  if (e == cudaSuccess) e = scratch_pool_for(m->device, &m->pool);
  if (e != cudaSuccess) {
    delete m;
    return cuda_fail(e, "b200kf_model_create");
  }
  *out = m;
  return B200KF_OK;
}

int b200kf_model_destroy(b200kf_model* m) {
  if (!m) return B200KF_OK;
  // (the scratch pool belongs to the device, not to the model: a model may be destroyed while its last
  // launches -- and their stream-ordered scratch frees -- are still in flight)
  if (m->host_ws.slab) cudaFree(m->host_ws.slab);
  for (int i = 0; i < 2; ++i)
    if (m->host_ws.h_stage[i]) cudaFreeHost(m->host_ws.h_stage[i]);
  if (m->host_ws.ready) {
    cudaStreamDestroy(m->host_ws.s_copy);
    cudaStreamDestroy(m->host_ws.s_comp);
    for (int i = 0; i < 2; ++i) {
      cudaEventDestroy(m->host_ws.ev_copied[i]);
      cudaEventDestroy(m->host_ws.ev_done[i]);
    }
  }
  delete m;
  return B200KF_OK;
}

int b200kf_model_dims(const b200kf_model* m, int32_t* n, int32_t* d, int32_t* m_gps) {
  if (!m) return set_error(B200KF_ERR_INVALID, "model is null");
  if (n) *n = m->n;
  if (d) *d = m->d;
  if (m_gps) *m_gps = m->m_gps;
  return B200KF_OK;
}

int b200kf_run_batch(const b200kf_model* m, int64_t n_filters, int64_t n_steps, const uint8_t* types,
                     const double* dts, const void* payload, void* x, void* P, double* agg,
                     void* traj, int32_t traj_ncomp, void* ptraj, void* logdet, uint8_t* used,
                     int32_t* status, double gate_threshold, uint32_t flags, void* stream) {
  return b200kf_run_batch_rows(m, n_filters, n_steps, types, dts, payload, nullptr, x, P, agg, traj, traj_ncomp, ptraj,
                               logdet, used, status, gate_threshold, flags, stream);
}

static int run_batch_impl(const b200kf_model* m, int64_t n_filters, int64_t n_steps, const uint8_t* types,
                          const double* dts, const void* payload, const int32_t* step_rows, void* x, void* P,
                          double* agg, void* traj, int32_t traj_ncomp, void* ptraj, void* logdet, uint8_t* used,
                          int32_t* status, double gate_threshold, const double* gate_thresholds, uint32_t flags,
                          void* stream);

int b200kf_run_batch_rows(const b200kf_model* m, int64_t n_filters, int64_t n_steps, const uint8_t* types,
                          const double* dts, const void* payload, const int32_t* step_rows, void* x, void* P,
                          double* agg, void* traj, int32_t traj_ncomp, void* ptraj, void* logdet, uint8_t* used,
                          int32_t* status, double gate_threshold, uint32_t flags, void* stream) {
  return run_batch_impl(m, n_filters, n_steps, types, dts, payload, step_rows, x, P, agg, traj, traj_ncomp, ptraj,
                        logdet, used, status, gate_threshold, nullptr, flags, stream);
}

int b200kf_run_batch_gated(const b200kf_model* m, int64_t n_filters, int64_t n_steps, const uint8_t* types,
                           const double* dts, const void* payload, const double* gate_thresholds, void* x, void* P,
                           double* agg, void* traj, int32_t traj_ncomp, void* ptraj, void* logdet, uint8_t* used,
                           int32_t* status, uint32_t flags, void* stream) {
  if (!gate_thresholds) return set_error(B200KF_ERR_INVALID, "gate_thresholds is required");
  return run_batch_impl(m, n_filters, n_steps, types, dts, payload, nullptr, x, P, agg, traj, traj_ncomp, ptraj, logdet,
                        used, status, 0.0, gate_thresholds, flags, stream);
}

static int run_batch_impl(const b200kf_model* m, int64_t n_filters, int64_t n_steps, const uint8_t* types,
                          const double* dts, const void* payload, const int32_t* step_rows, void* x, void* P,
                          double* agg, void* traj, int32_t traj_ncomp, void* ptraj, void* logdet, uint8_t* used,
                          int32_t* status, double gate_threshold, const double* gate_thresholds, uint32_t flags,
                          void* stream) {
  if (!m) return set_error(B200KF_ERR_INVALID, "model is null");
  if (n_filters < 0 || n_steps < 0) return set_error(B200KF_ERR_INVALID, "negative size");
  if (n_steps > 0x7fffffffLL) return set_error(B200KF_ERR_UNSUPPORTED, "n_steps > 2^31-1 per call");
  if (n_filters == 0) return B200KF_OK;
  if (!x || !P) return set_error(B200KF_ERR_INVALID, "x and P are required");
  if (n_steps > 0 && (!types || !dts || !payload))
    return set_error(B200KF_ERR_INVALID, "types, dts and payload are required when n_steps > 0");
  if (traj && (traj_ncomp < 1 || traj_ncomp > m->n))
    return set_error(B200KF_ERR_INVALID, "traj_ncomp must be in [1, %d]", m->n);
  if ((flags & B200KF_RUN_FORCE_DENSE) &&
      (flags & (B200KF_RUN_FORCE_DECOUPLED | B200KF_RUN_FORCE_SYMMETRIC | B200KF_RUN_FORCE_SHARED)))
    return set_error(B200KF_ERR_INVALID, "FORCE_DENSE excludes FORCE_DECOUPLED / FORCE_SYMMETRIC / FORCE_SHARED");
  RunParams p;
  p.mc = m->mc;
  p.n_filters = n_filters;
  p.n_steps = static_cast<int>(n_steps);
  p.types = types;
  p.dts = dts;
  p.payload = payload;
  p.x = x;
  p.P = P;
  p.agg = agg;
  p.traj = traj;
  p.traj_ncomp = traj ? traj_ncomp : 0;
  p.ptraj = ptraj;
  p.logdet = logdet;
  p.used = used;
  p.status = status;
  p.gate_enabled = (gate_thresholds != nullptr || !std::isnan(gate_threshold)) ? 1 : 0;
  p.gate_threshold = gate_threshold;
  p.gate_thresholds = gate_thresholds;
  p.agg_continue = (flags & B200KF_RUN_AGG_CONTINUE) ? 1 : 0;
  p.dense_flag = nullptr;
  p.gain_recs = nullptr;
  p.step_rows = step_rows;
  p.small_tiles = 0;
  return dispatch_run(m, p, flags, static_cast<cudaStream_t>(stream));
}

// Copy `bytes` from pageable memory into a pinned staging buffer on `n_threads` host threads (contiguous slices):
// one thread moves ~12-14 GB/s, less than a quarter of what the host->device link takes.
static void staged_memcpy(void* dst, const void* src, size_t bytes, int n_threads) {
  if (n_threads <= 1 || bytes < (size_t(8) << 20)) {
    memcpy(dst, src, bytes);
    return;
  }
  const size_t slice = ((bytes + n_threads - 1) / n_threads + 4095) & ~size_t(4095);
  std::vector<std::thread> th;
  th.reserve(n_threads);
  for (int t = 1; t < n_threads; ++t) {
    const size_t off = slice * t;
    if (off >= bytes) break;
    const size_t len = bytes - off < slice ? bytes - off : slice;
    try {
      th.emplace_back([=]() { memcpy(static_cast<char*>(dst) + off, static_cast<const char*>(src) + off, len); });
    } catch (...) {  // no thread to be had (std::system_error): this slice is copied here; nothing crosses the C ABI
      memcpy(static_cast<char*>(dst) + off, static_cast<const char*>(src) + off, len);
    }
  }
  memcpy(dst, src, slice < bytes ? slice : bytes);
  for (auto& t : th) t.join();
}

int b200kf_run_batch_host(const b200kf_model* m, int64_t n_filters, int64_t n_steps,
                          const uint8_t* types, const double* dts, const void* payload, void* x,
                          void* P, double* agg, int32_t* status, int64_t chunk_steps, uint32_t flags) {
  if (!m) return set_error(B200KF_ERR_INVALID, "model is null");
  if (n_filters <= 0 || n_steps < 0) return set_error(B200KF_ERR_INVALID, "bad size");
  if (!x || !P) return set_error(B200KF_ERR_INVALID, "x and P are required");
  if (n_steps > 0 && (!types || !dts || !payload))
    return set_error(B200KF_ERR_INVALID, "types, dts and payload are required");
  const size_t rs = m->desc.precision == B200KF_F32 ? 4 : 8;
  const size_t step_bytes = static_cast<size_t>(m->d) * n_filters * rs;
  if (chunk_steps <= 0) {
    static const size_t chunk_mb = []() {
      const char* e = getenv("B200KF_HOST_CHUNK_MB");
      const long v = e ? atol(e) : 0;
      return static_cast<size_t>(v > 0 ? v : 128);
    }();
    chunk_steps = static_cast<int64_t>((chunk_mb << 20) / (step_bytes ? step_bytes : 1));
    if (chunk_steps < 4) chunk_steps = 4;
  }
  if (chunk_steps > n_steps) chunk_steps = n_steps > 0 ? n_steps : 1;
  const size_t xb = static_cast<size_t>(n_filters) * m->n * rs;
  const size_t Pb = xb * m->n;

  // Persistent workspace owned by the model (streams, events, one growable device slab): the
  // per-call cudaMalloc/cudaFree of the first version cost more than the copies once NCCL had
  // peer-mapped the device.  Calls on one model serialise on the workspace mutex.
  HostWorkspace& ws = const_cast<b200kf_model*>(m)->host_ws;
  std::lock_guard<std::mutex> ws_lock(ws.mu);  // concurrent callers of one model take turns
  auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
  const size_t pay_b = up(step_bytes * chunk_steps);
  const size_t nst = n_steps > 0 ? static_cast<size_t>(n_steps) : 1;
  const size_t need = 2 * pay_b + up(xb) + up(Pb) + up(32 * n_filters) + up(4 * n_filters) + up(nst) + up(8 * nst);
  int rc = B200KF_OK;
  int stage_threads = 0;
#define HOST_TRY(expr)                              \
  do {                                              \
    cudaError_t _e = (expr);                        \
    if (_e != cudaSuccess) {                        \
      rc = cuda_fail(_e, #expr);                    \
      goto done;                                    \
    }                                               \
  } while (0)
  char* base = nullptr;
  void *d_pay[2], *d_x, *d_P;
  double *d_agg, *d_dts;
  int32_t* d_status;
  uint8_t* d_types;
  if (!ws.ready) {
    HOST_TRY(cudaStreamCreateWithFlags(&ws.s_copy, cudaStreamNonBlocking));
    HOST_TRY(cudaStreamCreateWithFlags(&ws.s_comp, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      HOST_TRY(cudaEventCreateWithFlags(&ws.ev_copied[i], cudaEventDisableTiming));
      HOST_TRY(cudaEventCreateWithFlags(&ws.ev_done[i], cudaEventDisableTiming));
    }
    ws.ready = 1;
  }
  if (ws.cap < need) {
    if (ws.slab) cudaFree(ws.slab);
    ws.slab = nullptr;
    ws.cap = 0;
    HOST_TRY(cudaMalloc(&ws.slab, need));
    ws.cap = need;
  }
  base = static_cast<char*>(ws.slab);
  d_pay[0] = base;
  d_pay[1] = base + pay_b;
  base += 2 * pay_b;
  d_x = base, base += up(xb);
  d_P = base, base += up(Pb);
  d_agg = reinterpret_cast<double*>(base), base += up(32 * n_filters);
  d_status = reinterpret_cast<int32_t*>(base), base += up(4 * n_filters);
  d_types = reinterpret_cast<uint8_t*>(base), base += up(nst);
  d_dts = reinterpret_cast<double*>(base);
  // A payload in pageable memory (what a drop-in caller's numpy array is) would be staged by the driver on one
  // thread (11 GB/s end to end); here a few host threads fill a pinned chunk buffer while the previous chunk's
  // copy is in flight.  B200KF_STAGE_THREADS=0 leaves it to the driver.
  if (n_steps > 0) {
    cudaPointerAttributes pa;
    const cudaError_t pe = cudaPointerGetAttributes(&pa, payload);
    if (pe != cudaSuccess) cudaGetLastError();
    if (pe == cudaSuccess && pa.type == cudaMemoryTypeUnregistered) {
      static const int env_threads = []() {
        const char* e = getenv("B200KF_STAGE_THREADS");
        if (e) return atoi(e);
        const unsigned hc = std::thread::hardware_concurrency();
        return static_cast<int>(hc >= 32 ? 16 : (hc >= 4 ? hc / 2 : 1));  // 4 threads already reach 3.4 x the driver's staging
      }();
      stage_threads = env_threads;
    }
  }
  if (stage_threads > 0 && ws.h_cap < pay_b) {
    for (int i = 0; i < 2; ++i) {
      if (ws.h_stage[i]) cudaFreeHost(ws.h_stage[i]);
      ws.h_stage[i] = nullptr;
    }
    ws.h_cap = 0;
    HOST_TRY(cudaHostAlloc(&ws.h_stage[0], pay_b, cudaHostAllocDefault));
    HOST_TRY(cudaHostAlloc(&ws.h_stage[1], pay_b, cudaHostAllocDefault));
    ws.h_cap = pay_b;
  }
  {
    cudaStream_t s_copy = ws.s_copy, s_comp = ws.s_comp;
    HOST_TRY(cudaMemcpyAsync(d_x, x, xb, cudaMemcpyHostToDevice, s_comp));
    HOST_TRY(cudaMemcpyAsync(d_P, P, Pb, cudaMemcpyHostToDevice, s_comp));
    if (n_steps > 0) {
      HOST_TRY(cudaMemcpyAsync(d_types, types, n_steps, cudaMemcpyHostToDevice, s_comp));
      HOST_TRY(cudaMemcpyAsync(d_dts, dts, sizeof(double) * n_steps, cudaMemcpyHostToDevice, s_comp));
    }
    int64_t c = 0;
    for (int64_t k0 = 0; k0 < n_steps || c == 0; k0 += chunk_steps, ++c) {
      const int b = static_cast<int>(c & 1);
      const int64_t steps = (n_steps - k0) < chunk_steps ? (n_steps - k0) : chunk_steps;
      if (steps > 0) {
        if (c >= 2) HOST_TRY(cudaStreamWaitEvent(s_copy, ws.ev_done[b], 0));
        const char* src = static_cast<const char*>(payload) + step_bytes * k0;
        if (stage_threads > 0) {
          if (c >= 2) HOST_TRY(cudaEventSynchronize(ws.ev_copied[b]));  // the copy out of this staging buffer is done
          staged_memcpy(ws.h_stage[b], src, step_bytes * steps, stage_threads);
          src = static_cast<const char*>(ws.h_stage[b]);
        }
        HOST_TRY(cudaMemcpyAsync(d_pay[b], src, step_bytes * steps, cudaMemcpyHostToDevice, s_copy));
        HOST_TRY(cudaEventRecord(ws.ev_copied[b], s_copy));
        HOST_TRY(cudaStreamWaitEvent(s_comp, ws.ev_copied[b], 0));
      }
      const uint32_t f = (flags & ~B200KF_RUN_AGG_CONTINUE) | (c > 0 ? B200KF_RUN_AGG_CONTINUE : 0u);
      rc = b200kf_run_batch(m, n_filters, steps, d_types + k0, d_dts + k0, d_pay[b], d_x, d_P, d_agg,
                            nullptr, 0, nullptr, nullptr, nullptr, d_status, NAN, f, s_comp);
      if (rc != B200KF_OK) goto done;
      HOST_TRY(cudaEventRecord(ws.ev_done[b], s_comp));
      if (n_steps == 0) break;
    }
    HOST_TRY(cudaMemcpyAsync(x, d_x, xb, cudaMemcpyDeviceToHost, s_comp));
    HOST_TRY(cudaMemcpyAsync(P, d_P, Pb, cudaMemcpyDeviceToHost, s_comp));
    if (agg)
      HOST_TRY(cudaMemcpyAsync(agg, d_agg, sizeof(double) * 4 * n_filters, cudaMemcpyDeviceToHost, s_comp));
    if (status)
      HOST_TRY(cudaMemcpyAsync(status, d_status, sizeof(int32_t) * n_filters, cudaMemcpyDeviceToHost, s_comp));
    HOST_TRY(cudaStreamSynchronize(s_comp));
    HOST_TRY(cudaStreamSynchronize(s_copy));
  }
done:
  if (rc != B200KF_OK) cudaDeviceSynchronize();
#undef HOST_TRY
  return rc;
}

int b200kf_score_schedules_strided(const b200kf_model* m, int32_t n_events, const uint8_t* types,
                                   const double* times, double t_prev, double t_end, const double* P0,
                                   uint64_t low_bits, int32_t n_low, uint64_t first, uint64_t count,
                                   double threshold, double initial_logdet, double* per_schedule,
                                   b200kf_best* best, uint32_t flags, void* stream) {
  if (!m) return set_error(B200KF_ERR_INVALID, "model is null");
  if (n_events < 1 || n_events > kMaxEvents)
    return set_error(B200KF_ERR_UNSUPPORTED, "n_events must be in [1, %d]", kMaxEvents);
  if (!types || !times || !P0 || !best) return set_error(B200KF_ERR_INVALID, "null argument");
  if (n_low < 0 || n_low > n_events || (low_bits >> n_low) != 0)
    return set_error(B200KF_ERR_INVALID, "n_low must be in [0, n_events] and low_bits < 2^n_low");
  const uint64_t total = 1ULL << (n_events - n_low);  // schedules that share these low bits
  if (first > total || count > total - first)
    return set_error(B200KF_ERR_INVALID, "schedule range exceeds 2^(n_events - n_low)");
  ScoreParams p;
  p.low_bits = low_bits;
  p.n_low = n_low;
  p.mc = m->mc;
  p.n_events = n_events;
  memset(p.types, 0, sizeof(p.types));
  memset(p.times, 0, sizeof(p.times));
  memset(p.P0, 0, sizeof(p.P0));
  for (int i = 0; i < n_events; ++i) {
    if (types[i] > B200KF_EV_PREDICT_ONLY) return set_error(B200KF_ERR_INVALID, "bad event type");
    p.types[i] = types[i];
    p.times[i] = times[i];
  }
  memcpy(p.P0, P0, sizeof(double) * m->n * m->n);
  p.t_prev = t_prev;
  p.t_end = t_end;
  p.first = first;
  p.count = count;
  p.threshold = threshold;
  p.initial_logdet = initial_logdet;
  p.per_schedule = per_schedule;
  {
    const double tol = 1e-9 * fmax(1.0, fabs(threshold));
    p.det_thr = exp(threshold);
    p.det_lo = exp(threshold - tol);  // NaN for an infinite threshold: nothing is "near" it
    p.det_hi = exp(threshold + tol);
    // sum = (exponent + log2 mantissa) ln 2 (+ term 0): a schedule whose exponent + mantissa exceeds the best's by
    // this much is larger by >= half of it in the sum, far above the rounding of either (1e-16 relative)
    p.q_slack = 1e-9 * (1.0 + (isnan(initial_logdet) ? 0.0 : fabs(initial_logdet)));
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ScratchBuf parts;  // per call: concurrent searches on one model never share their partial results
  B200KF_CUDA_TRY(parts.alloc(m, sizeof(ScorePartial) * kMaxSearchPartials, st));
  p.partials = parts.p;
  p.dense_flag = nullptr;
  const int blocks = dispatch_score(m, p, flags, best, st);
  if (blocks < 0) return blocks;
  {
    static const bool pdl = []() { const char* e = getenv("B200KF_PDL"); return !(e && e[0] == '0'); }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    B200KF_CUDA_TRY(cudaLaunchKernelEx(&cfg, score_final_kernel, static_cast<const ScorePartial*>(p.partials), blocks,
                                       static_cast<int>(n_events), best));
  }
  B200KF_CUDA_TRY(cudaGetLastError());
  return B200KF_OK;
}

int b200kf_score_schedules(const b200kf_model* m, int32_t n_events, const uint8_t* types,
                           const double* times, double t_prev, double t_end, const double* P0,
                           uint64_t first, uint64_t count, double threshold, double initial_logdet,
                           double* per_schedule, b200kf_best* best, uint32_t flags, void* stream) {
  return b200kf_score_schedules_strided(m, n_events, types, times, t_prev, t_end, P0, 0, 0, first, count, threshold,
                                        initial_logdet, per_schedule, best, flags, stream);
}

int b200kf_run_masked(const b200kf_model* m, int32_t n_events, const uint8_t* types, const double* times,
                      const double* payload, double t_prev, double t_end, const double* x0,
                      const double* P0, int64_t n_schedules, const uint64_t* masks, double* x_final,
                      double* P_final, double* traj, int32_t traj_ncomp, double* logdet,
                      int32_t* n_terms, uint32_t flags, void* stream) {
  if (!m) return set_error(B200KF_ERR_INVALID, "model is null");
  if (m->desc.precision != B200KF_F64)
    return set_error(B200KF_ERR_UNSUPPORTED, "b200kf_run_masked is float64 only");
  if (n_events < 1 || n_events > kMaxMaskedEvents)
    return set_error(B200KF_ERR_UNSUPPORTED, "n_events must be in [1, %d]", kMaxMaskedEvents);
  if (!types || !times || !payload || !x0 || !P0 || !masks || !x_final)
    return set_error(B200KF_ERR_INVALID, "null argument");
  if (traj && (traj_ncomp < 0 || traj_ncomp > m->n))
    return set_error(B200KF_ERR_INVALID, "traj_ncomp must be in [0, %d]", m->n);
  if (n_schedules <= 0) return B200KF_OK;
  MaskedParams p;
  memset(&p, 0, sizeof(p));
  p.mc = m->mc;
  p.n_events = n_events;
  for (int i = 0; i < n_events; ++i) {
    if (types[i] > B200KF_EV_PREDICT_ONLY) return set_error(B200KF_ERR_INVALID, "bad event type");
    p.types[i] = types[i];
    p.times[i] = times[i];
    for (int c = 0; c < m->d; ++c) p.payload[i * m->d + c] = payload[i * m->d + c];
  }
  p.t_prev = t_prev;
  p.t_end = t_end;
  memcpy(p.x0, x0, sizeof(double) * m->n);
  memcpy(p.P0, P0, sizeof(double) * m->n * m->n);
  p.n_schedules = n_schedules;
  p.masks = reinterpret_cast<const unsigned long long*>(masks);
  p.x_final = x_final;
  p.P_final = P_final;
  p.traj = traj;
  p.traj_ncomp = traj ? traj_ncomp : 0;
  p.logdet = logdet;
  p.n_terms = n_terms;
  return dispatch_masked(m, p, flags, static_cast<cudaStream_t>(stream));
}

int b200kf_score_products(const b200kf_model* m, int32_t n_windows, const int32_t* queue_len,
                          const uint8_t* types, const double* times, const double* payload,
                          const double* ref_pos, const double* x0, const double* P0, uint64_t first,
                          uint64_t count, double* per_combo, b200kf_best_product* best, uint32_t flags,
                          void* stream) {
  if (!m) return set_error(B200KF_ERR_INVALID, "model is null");
  if (m->desc.precision != B200KF_F64)
    return set_error(B200KF_ERR_UNSUPPORTED, "b200kf_score_products is float64 only");
  if (n_windows < 1 || n_windows > kMaxWindows)
    return set_error(B200KF_ERR_UNSUPPORTED, "n_windows must be in [1, %d]", kMaxWindows);
  if (!queue_len || !types || !times || !payload || !ref_pos || !x0 || !P0 || !best)
    return set_error(B200KF_ERR_INVALID, "null argument");
  ProductParams p;
  memset(&p, 0, sizeof(p));
  p.mc = m->mc;
  p.n_windows = n_windows;
  long long off = 0;
  for (int w = 0; w < n_windows; ++w) {
    if (queue_len[w] < 1) return set_error(B200KF_ERR_INVALID, "queue %d is empty: no sequence exists", w);
    p.q_len[w] = queue_len[w];
    p.q_off[w] = static_cast<int>(off);
    off += queue_len[w];
    if (off > 0x7fffffffLL) return set_error(B200KF_ERR_UNSUPPORTED, "more than 2^31-1 queued events");
  }
  unsigned long long total = 1;
  for (int w = n_windows - 1; w >= 0; --w) {  // itertools.product: the last queue varies fastest
    p.stride[w] = total;
    if (total > (~0ULL) / static_cast<unsigned long long>(queue_len[w]))
      return set_error(B200KF_ERR_UNSUPPORTED, "the product of the queue lengths exceeds 2^64");
    total *= static_cast<unsigned long long>(queue_len[w]);
  }
  if (first > total || count > total - first)
    return set_error(B200KF_ERR_INVALID, "sequence range exceeds the product of the queue lengths");
  p.types = types;
  p.times = times;
  p.payload = payload;
  p.ref = ref_pos;
  memcpy(p.x0, x0, sizeof(double) * m->n);
  memcpy(p.P0, P0, sizeof(double) * m->n * m->n);
  p.first = first;
  p.count = count;
  p.per_combo = per_combo;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ScratchBuf parts;
  B200KF_CUDA_TRY(parts.alloc(m, sizeof(ProductPartial) * kMaxSearchBlocks, st));
  p.partials = parts.p;
  const int blocks = dispatch_product(m, p, flags, st);
  if (blocks < 0) return blocks;
  product_final_kernel<<<1, 256, 0, st>>>(static_cast<const ProductPartial*>(p.partials), blocks, count, best);
  B200KF_CUDA_TRY(cudaGetLastError());
  return B200KF_OK;
}

int b200kf_track_lerp(int64_t n_track, const double* track_t, const double* track_p, int32_t k,
                      int64_t n_query, const double* query_t, double* out, void* stream) {
  if (n_track < 1 || k < 1 || n_query < 0) return set_error(B200KF_ERR_INVALID, "need n_track >= 1, k >= 1");
  if (n_query == 0) return B200KF_OK;
  if (!track_t || !track_p || !query_t || !out) return set_error(B200KF_ERR_INVALID, "null argument");
  const int tpb = 128;
  track_lerp_kernel<<<static_cast<unsigned>((n_query + tpb - 1) / tpb), tpb, 0, static_cast<cudaStream_t>(stream)>>>(
      n_track, track_t, track_p, k, n_query, query_t, out);
  B200KF_CUDA_TRY(cudaGetLastError());
  return B200KF_OK;
}

int b200kf_rmse_vs_track(const b200kf_model* m, int64_t n_filters, int64_t n_steps, const void* traj,
                         int32_t traj_ncomp, const double* ref_pos, double* rmse, void* stream) {
  if (!m) return set_error(B200KF_ERR_INVALID, "model is null");
  if (n_filters < 0 || n_steps < 1) return set_error(B200KF_ERR_INVALID, "need n_steps >= 1");
  if (traj_ncomp < 3) return set_error(B200KF_ERR_INVALID, "the trajectory must hold the 3 position components");
  if (n_filters == 0) return B200KF_OK;
  if (!traj || !ref_pos || !rmse) return set_error(B200KF_ERR_INVALID, "null argument");
  // time segments: enough threads to fill the GPU when there are few filters, bounded by 8 MB of partials
  const long long cap = static_cast<long long>((size_t(8) << 20) / sizeof(double));
  long long seg = (static_cast<long long>(m->sm_count) * 2048 + n_filters - 1) / n_filters;
  if (seg > n_steps) seg = n_steps;
  if (seg * n_filters > cap) seg = cap / n_filters;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ScratchBuf part;
  double* dst = nullptr;
  if (seg < 1) {  // more filters than partials: one segment, partial sums straight into rmse
    seg = 1;
    dst = rmse;
  }
  if (seg > 65535) seg = 65535;
  const long long seg_len = (n_steps + seg - 1) / seg;
  const int n_seg = static_cast<int>((n_steps + seg_len - 1) / seg_len);
  if (dst == nullptr) {
    B200KF_CUDA_TRY(part.alloc(m, sizeof(double) * static_cast<size_t>(n_seg) * n_filters, st));
    dst = static_cast<double*>(part.p);
  }
  const int tpb = 128;
  const dim3 grid(static_cast<unsigned>((n_filters + tpb - 1) / tpb), static_cast<unsigned>(n_seg));
  if (m->desc.precision == B200KF_F32)
    rmse_partial_kernel<float><<<grid, tpb, 0, st>>>(static_cast<const float*>(traj), traj_ncomp, n_filters, n_steps,
                                                    ref_pos, seg_len, dst);
  else
    rmse_partial_kernel<double><<<grid, tpb, 0, st>>>(static_cast<const double*>(traj), traj_ncomp, n_filters,
                                                     n_steps, ref_pos, seg_len, dst);
  B200KF_CUDA_TRY(cudaGetLastError());
  rmse_final_kernel<<<grid.x, tpb, 0, st>>>(dst, n_seg, n_filters, n_steps, rmse);
  B200KF_CUDA_TRY(cudaGetLastError());
  return B200KF_OK;
}

int b200kf_select_greedy_gains(const b200kf_model* m, int64_t n_events, const uint8_t* types,
                               const double* times, int32_t n_inst, const double* periods, const double* t_prev,
                               const double* P0, int64_t max_sel, int32_t* sel, int64_t* n_sel, double* P_final,
                               void* gain_records, uint32_t flags, void* stream) {
  if (!m) return set_error(B200KF_ERR_INVALID, "model is null");
  if (m->desc.model_id == MODEL_BIAS9 || m->desc.model_id == MODEL_CVBU6)
    return set_error(B200KF_ERR_UNSUPPORTED, "greedy selection needs an IMU measurement model");
  if (n_events < 0 || n_inst < 0 || max_sel < 0) return set_error(B200KF_ERR_INVALID, "negative size");
  if (n_events > 0x7fffffffLL) return set_error(B200KF_ERR_UNSUPPORTED, "n_events > 2^31-1");
  if (n_inst == 0) return B200KF_OK;
  if (!periods || !t_prev || !P0 || !sel || !n_sel || (n_events > 0 && (!types || !times)))
    return set_error(B200KF_ERR_INVALID, "null argument");
  SelectParams p;
  p.mc = m->mc;
  p.n_events = n_events;
  p.types = types;
  p.times = times;
  p.n_inst = n_inst;
  p.periods = periods;
  p.t_prev = t_prev;
  p.P0 = P0;
  p.max_sel = max_sel;
  p.sel = sel;
  p.n_sel = reinterpret_cast<long long*>(n_sel);
  p.P_final = P_final;
  p.gain_recs = gain_records;
  p.e_begin = 0;
  p.carry = nullptr;
  return dispatch_select(m, p, flags, static_cast<cudaStream_t>(stream));
}

int b200kf_select_greedy_chunk(const b200kf_model* m, int64_t e_begin, int64_t e_end, const uint8_t* types,
                               const double* times, int32_t n_inst, const double* periods, const double* t_prev,
                               const double* P0, double* carry, double* P_state, int64_t max_sel, int32_t* sel,
                               int64_t* n_sel, void* gain_records, uint32_t flags, void* stream) {
  if (!m) return set_error(B200KF_ERR_INVALID, "model is null");
  if (m->desc.model_id == MODEL_BIAS9 || m->desc.model_id == MODEL_CVBU6)
    return set_error(B200KF_ERR_UNSUPPORTED, "greedy selection needs an IMU measurement model");
  if (e_begin < 0 || e_end < e_begin || n_inst < 0 || max_sel < 0) return set_error(B200KF_ERR_INVALID, "bad range or size");
  if (e_end > 0x7fffffffLL) return set_error(B200KF_ERR_UNSUPPORTED, "n_events > 2^31-1");
  if (n_inst == 0) return B200KF_OK;
  if (!periods || !t_prev || !P0 || !carry || !P_state || !sel || !n_sel || (e_end > 0 && (!types || !times)))
    return set_error(B200KF_ERR_INVALID, "null argument");
  SelectParams p;
  p.mc = m->mc;
  p.n_events = e_end;
  p.types = types;
  p.times = times;
  p.n_inst = n_inst;
  p.periods = periods;
  p.t_prev = t_prev;
  p.P0 = P0;
  p.max_sel = max_sel;
  p.sel = sel;
  p.n_sel = reinterpret_cast<long long*>(n_sel);
  p.P_final = P_state;
  p.gain_recs = gain_records;
  p.e_begin = e_begin;
  p.carry = carry;
  return dispatch_select(m, p, flags, static_cast<cudaStream_t>(stream));
}

int b200kf_select_greedy(const b200kf_model* m, int64_t n_events, const uint8_t* types,
                         const double* times, int32_t n_inst, const double* periods, const double* t_prev,
                         const double* P0, int64_t max_sel, int32_t* sel, int64_t* n_sel, double* P_final,
                         uint32_t flags, void* stream) {
  return b200kf_select_greedy_gains(m, n_events, types, times, n_inst, periods, t_prev, P0, max_sel, sel, n_sel,
                                    P_final, nullptr, flags, stream);
}

int b200kf_gain_record_bytes(const b200kf_model* m, size_t* bytes) {
  if (!m || !bytes) return set_error(B200KF_ERR_INVALID, "null argument");
  switch (m->desc.model_id) {
    case MODEL_CV6: *bytes = gain_record_bytes_m<MODEL_CV6>(); break;
    case MODEL_CA8: *bytes = gain_record_bytes_m<MODEL_CA8>(); break;
    case MODEL_BIAS9: *bytes = gain_record_bytes_m<MODEL_BIAS9>(); break;
    case MODEL_CVBU6: *bytes = gain_record_bytes_m<MODEL_CVBU6>(); break;
    case MODEL_CA15: *bytes = gain_record_bytes_m<MODEL_CA15>(); break;
    default: return set_error(B200KF_ERR_UNSUPPORTED, "unknown model id %d", m->desc.model_id);
  }
  return B200KF_OK;
}

int b200kf_run_gains(const b200kf_model* m, int64_t n_filters, int64_t n_steps, const void* gain_records,
                     const void* payload, const int32_t* step_rows, void* x, double* agg, void* traj,
                     int32_t traj_ncomp, void* logdet, uint8_t* used, int32_t* status, uint32_t flags, void* stream) {
  if (!m) return set_error(B200KF_ERR_INVALID, "model is null");
  if (n_filters < 0 || n_steps < 0) return set_error(B200KF_ERR_INVALID, "negative size");
  if (n_steps > 0x7fffffffLL) return set_error(B200KF_ERR_UNSUPPORTED, "n_steps > 2^31-1 per call");
  if (n_filters == 0) return B200KF_OK;
  if (!x) return set_error(B200KF_ERR_INVALID, "x is required");
  if (n_steps > 0 && (!gain_records || !payload))
    return set_error(B200KF_ERR_INVALID, "gain_records and payload are required when n_steps > 0");
  if (traj && (traj_ncomp < 1 || traj_ncomp > m->n))
    return set_error(B200KF_ERR_INVALID, "traj_ncomp must be in [1, %d]", m->n);
  const size_t rb = m->desc.precision == B200KF_F32 ? 4 : 8;
  if ((n_filters * rb) % 16 != 0 || (reinterpret_cast<uintptr_t>(payload) & 15u) != 0 ||
      (reinterpret_cast<uintptr_t>(gain_records) & 15u) != 0)
    return set_error(B200KF_ERR_UNSUPPORTED,
                     "b200kf_run_gains streams with bulk copies: 16-byte aligned buffers and rows (even filter count "
                     "in float64, a multiple of 4 in float32)");
  RunParams p;
  p.mc = m->mc;
  p.n_filters = n_filters;
  p.n_steps = static_cast<int>(n_steps);
  p.types = nullptr;
  p.dts = nullptr;
  p.payload = payload;
  p.x = x;
  p.P = nullptr;
  p.agg = agg;
  p.traj = traj;
  p.traj_ncomp = traj ? traj_ncomp : 0;
  p.ptraj = nullptr;
  p.logdet = logdet;
  p.used = used;
  p.status = status;
  p.gate_enabled = 0;
  p.gate_threshold = 0.0;
  p.gate_thresholds = nullptr;
  p.agg_continue = (flags & B200KF_RUN_AGG_CONTINUE) ? 1 : 0;
  p.dense_flag = nullptr;
  p.gain_recs = gain_records;
  p.step_rows = step_rows;
  p.small_tiles = (flags & B200KF_RUN_SMALL_TILES) ? 1 : 0;
  return dispatch_gains(m, p, static_cast<cudaStream_t>(stream));
}

int b200kf_cov_update(int32_t n, int32_t m_rows, int32_t k, const int32_t* rows, const double* H,
                      const double* R, int64_t batch, const double* Sigma, double* Sigma_post,
                      double* trace, void* stream) {
  if (n < 1 || n > kMaxN || k < 1 || k > kMaxN || m_rows < k)
    return set_error(B200KF_ERR_UNSUPPORTED, "need 1 <= n <= 16 and 1 <= k <= min(m, 16)");
  if (!rows || !H || !R || !Sigma || !Sigma_post) return set_error(B200KF_ERR_INVALID, "null argument");
  if (batch <= 0) return B200KF_OK;
  CovParams p;
  memset(&p, 0, sizeof(p));
  p.n = n;
  p.k = k;
  for (int a = 0; a < k; ++a) {
    if (rows[a] < 0 || rows[a] >= m_rows) return set_error(B200KF_ERR_INVALID, "row index out of range");
    for (int j = 0; j < n; ++j) p.H[a * n + j] = H[rows[a] * n + j];
    for (int b = 0; b < k; ++b) p.R[a * k + b] = R[rows[a] * m_rows + rows[b]];
  }
  const int tpb = 64;
  cov_update_kernel<<<static_cast<unsigned>((batch + tpb - 1) / tpb), tpb, 0,
                      static_cast<cudaStream_t>(stream)>>>(p, batch, Sigma, Sigma_post, trace);
  B200KF_CUDA_TRY(cudaGetLastError());
  return B200KF_OK;
}

int b200kf_fill_payload(const b200kf_model* m, int64_t n_filters, int64_t n_steps,
                        const uint8_t* types, uint64_t seed, void* payload, void* stream) {
  if (!m || !types || !payload) return set_error(B200KF_ERR_INVALID, "null argument");
  if (n_filters <= 0 || n_steps <= 0) return B200KF_OK;
  const int blocks = m->sm_count * 16;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (m->desc.precision == B200KF_F32)
    fill_payload_kernel<float><<<blocks, 256, 0, st>>>(static_cast<float*>(payload), types, n_filters,
                                                       n_steps, m->d, m->m_gps, seed);
  else
    fill_payload_kernel<double><<<blocks, 256, 0, st>>>(static_cast<double*>(payload), types, n_filters,
                                                        n_steps, m->d, m->m_gps, seed);
  B200KF_CUDA_TRY(cudaGetLastError());
  return B200KF_OK;
}

int b200kf_probe_fma_peak(int32_t precision, double* tflops) {
  if (!tflops) return set_error(B200KF_ERR_INVALID, "tflops is null");
  const bool distinct = (precision & 2) != 0;  // bit 1: three distinct register operands per FMA
  if ((precision & 1) == B200KF_F32) return probe_peak<float>(tflops, distinct);
  return probe_peak<double>(tflops, distinct);
}

}  // extern "C"
