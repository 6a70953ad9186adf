// kf_run_model.cu -- the batched filter-loop kernels for ONE model id (compiled once per model
// with -DB200KF_INST_MODEL=<id> so the four models build in parallel).
//
// Thread <-> filter.  State and covariance stay in registers for the whole launch; the event
// schedule (type, dt) is shared by all filters so every branch is warp-uniform; the payload is
// read time-major ([T, d, N], filter fastest) so each warp load is one fully coalesced 256-byte
// (f64) request, software-prefetched PF steps ahead.
#include <tuple>

#include "kf_internal.h"
#include "kf_axis.cuh"

#ifndef B200KF_INST_MODEL
#error "compile with -DB200KF_INST_MODEL=<model id>"
#endif

namespace b200kf {

#ifndef B200KF_PF
#define B200KF_PF 1
#endif
constexpr int PF = B200KF_PF;  // payload prefetch depth (steps)

// Launch shape per kernel variant.  The closed-form (value, rate)-chain kernels are capped at 72
// registers so 7 blocks x 128 threads = 896 filters are resident per SM: the 131 072-filter shard
// of BASELINE config 3 (885.6 filters per SM over 148 SMs) then runs as ONE balanced wave
// instead of 1.15 waves with an almost empty second one.
template <int MODEL, int VAR, typename real>
struct LaunchCfg {
  static constexpr bool fast = VariantTraits<MODEL, VAR, real>::Filter::fast2;
  static constexpr int max_threads = fast ? 128 : 256;
  static constexpr int min_blocks = fast ? 7 : 1;
};

// AUTO dispatch: mode_flags[0] != 0 -> some P couples chains (dense kernel); else mode_flags[1] != 0
// -> the chains of some filter differ (per-chain kernel); else mode_flags[2] != 0 -> the filters'
// covariances differ from each other (axis-symmetric kernel, one covariance per filter); else all
// filters of the launch carry the SAME covariance (shared-covariance kernel, one per launch).
enum : int { MODE_DECOUPLED = 0, MODE_DENSE = 1, MODE_SYMMETRIC = 2, MODE_SHARED = 3 };
__device__ __forceinline__ bool mode_selected(const int* flags, int mine) {
  if (flags == nullptr) return true;
  const int actual = flags[0] ? MODE_DENSE
                              : (flags[1] ? MODE_DECOUPLED : (flags[2] ? MODE_SYMMETRIC : MODE_SHARED));
  return actual == mine;
}

// payload row of step k: k itself, or step_rows[k] when the launch walks a selection of a longer stream
// (the scheduled filter's selected events) without a gathered copy of the payload
__device__ __forceinline__ long long step_row(const int* __restrict__ rows, int k) {
  return rows != nullptr ? static_cast<long long>(__ldg(rows + k)) : static_cast<long long>(k);
}

template <int MODEL, int VAR, typename real, bool FULL>
__global__ void __launch_bounds__((LaunchCfg<MODEL, VAR, real>::max_threads),
                                  (LaunchCfg<MODEL, VAR, real>::min_blocks))
    kf_run_kernel(const RunParams p) {
  using MT = VariantTraits<MODEL, VAR, real>;
  constexpr bool DENSE = (VAR == VAR_DENSE);
  constexpr int n = MT::n;
  constexpr int D = MT::d;
  constexpr bool FAST = MT::Filter::fast2;
  if (!mode_selected(p.dense_flag, VAR)) return;
  const long long f = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (f >= p.n_filters) return;
  const long long N = p.n_filters;
  const int T = p.n_steps;

  typename MT::Filter flt;
  {
    const real* xg = static_cast<const real*>(p.x) + f * n;
    const real* Pg = static_cast<const real*>(p.P) + f * (n * n);
    flt.each([&](auto& b) { b.load(xg, Pg, n); });
  }

  const double gate = (FULL && p.gate_thresholds != nullptr) ? p.gate_thresholds[f] : p.gate_threshold;
  double rprod = 1.0;
  if constexpr (FAST) rprod = fast2_rprod(flt, p.mc);
  // lean running aggregate: mantissa product + exponent sum + max; "last" is recomputed from the
  // final P and the count is T, which keeps the hot loop inside the 72-register budget
  double acc_mant = 1.0;
  long long acc_max = 0;  // max det as raw bits: positive doubles order like their bit patterns
  int acc_expo = 0;
  bool ok = true;

  const real* pay = static_cast<const real*>(p.payload) + f;
  const long long step_stride = static_cast<long long>(D) * N;
  real ubuf[PF][D];
#pragma unroll
  for (int j = 0; j < PF; ++j) {
    if (j < T) {
#pragma unroll
      for (int c = 0; c < D; ++c) ubuf[j][c] = __ldcs(pay + step_row(p.step_rows, j) * step_stride + c * N);
    }
  }

  for (int k0 = 0; k0 < T; k0 += PF) {
#pragma unroll
    for (int j = 0; j < PF; ++j) {
      const int k = k0 + j;
      if (k < T) {
        real u[D];
#pragma unroll
        for (int c = 0; c < D; ++c) u[c] = ubuf[j][c];
        if (k + PF < T) {
          const real* src = pay + step_row(p.step_rows, k + PF) * step_stride;
#pragma unroll
          for (int c = 0; c < D; ++c) ubuf[j][c] = __ldcs(src + c * N);
        }
        const int ty = __ldg(p.types + k);
        const real dt = static_cast<real>(__ldg(p.dts + k));

        if constexpr (FAST) fast2_predict<true>(flt, dt, p.mc);
        else filter_predict<true>(flt, dt, p.mc, ty, u);
        bool do_update = (ty != EV_PREDICT_ONLY);
        if constexpr (FULL) {
          if (p.gate_enabled) {  // kf_workers.py:1023-1025
            const double dpre = filter_det(flt, ok);
            do_update = do_update && (fast_log_pos(dpre) > gate);
          }
        }
        double det;
        if constexpr (FAST) {
          det = do_update ? fast2_update<true>(flt, dt, p.mc, ty, u, rprod, ok) : filter_det(flt, ok);
        } else {
          if (do_update) filter_update<true>(flt, dt, p.mc, ty, u);
          det = filter_det(flt, ok);
        }
        {
          acc_max = max(acc_max, __double_as_longlong(det));
          const double m = acc_mant * det;
          const int hi = __double2hiint(m);
          const int e = ((hi >> 20) & 0x7ff) - 1023;
          acc_expo += e;
          acc_mant = __hiloint2double(hi - (e << 20), __double2loint(m));
        }

        if constexpr (FULL) {
          if (p.traj != nullptr) {
            real* tr = static_cast<real*>(p.traj) + (static_cast<long long>(k) * p.traj_ncomp) * N + f;
            const int nc = p.traj_ncomp;
            flt.each([&](const auto& b) {
              b.for_each_state([&](auto G, real v) {
                constexpr int g = G;
                if (g < nc) __stcs(tr + static_cast<long long>(g) * N, v);
              });
            });
          }
          if (p.ptraj != nullptr) {  // sf_KF_covariance.append(Pt.copy()), kf_workers.py:822
            real* pt = static_cast<real*>(p.ptraj) + (static_cast<long long>(k) * (n * n)) * N + f;
            if constexpr (!DENSE) {
              for (int e = 0; e < n * n; ++e) pt[static_cast<long long>(e) * N] = real(0);
            }
            flt.each([&](const auto& b) {
              b.for_each_cov([&](int gi, int gj, real v) { pt[static_cast<long long>(gi * n + gj) * N] = v; });
            });
          }
          if (p.logdet != nullptr)
            __stcs(static_cast<real*>(p.logdet) + static_cast<long long>(k) * N + f,
                   static_cast<real>(fast_log_pos(det)));
          if (p.used != nullptr)
            p.used[static_cast<long long>(k) * N + f] = do_update ? 1 : 0;
        }
      }
    }
  }

  {
    real* xg = static_cast<real*>(p.x) + f * n;
    real* Pg = static_cast<real*>(p.P) + f * (n * n);
    flt.each([&](const auto& b) { b.store(xg, Pg, n); });
  }
  if (p.agg != nullptr) {
    double* a = p.agg + f * 4;
    double mx = (T > 0) ? log(__longlong_as_double(acc_max)) : -INFINITY;
    double sm = log(acc_mant) + static_cast<double>(acc_expo) * 0.6931471805599453094;
    double cnt = static_cast<double>(T);
    if (p.agg_continue) {
      mx = fmax(mx, a[0]);
      sm += a[1];
      cnt += a[3];
    }
    a[0] = mx;
    a[1] = sm;
    a[2] = (T > 0) ? log(filter_det(flt, ok)) : (p.agg_continue ? a[2] : 0.0);
    a[3] = cnt;
  }
  if (p.status != nullptr) {
    const int bad = ok ? 0 : 1;
    p.status[f] = p.agg_continue ? (p.status[f] | bad) : bad;
  }
}


// ------------------------------------------------------------------------------------------
// Dense covariance of a three-axis model, three lanes per filter (kf_axis.cuh): the arbitrary-initial_pt
// path of the 15-state filter (kf_workers.py:642-645) without local memory.  Ten filters per warp; lanes 30
// and 31 run a clamped copy of the last filter and store nothing.  Same interface, outputs and aggregate
// conventions as kf_run_kernel<MODEL, VAR_DENSE>.
// ------------------------------------------------------------------------------------------
constexpr int kAxisThreads = 128;
constexpr int kAxisFiltersPerBlock = (kAxisThreads / 32) * 10;

template <int MODEL, typename real, bool FULL, bool GATED>
__global__ void __launch_bounds__(kAxisThreads, 2) kf_run_axis_kernel(const RunParams p) {
  using AT = typename AxisTraits<MODEL>::Table;
  using MT = ModelTraits<MODEL, true, real>;
  static_assert(!MT::imu_control, "axis kernel: measurement-type IMU events only");
  constexpr int n = MT::n;
  constexpr int D = MT::d;
  constexpr int NA = AT::n;
  constexpr int NC = AT::ncols();
  if (!mode_selected(p.dense_flag, VAR_DENSE)) return;
  const int lane = threadIdx.x & 31;
  const long long warp = blockIdx.x * static_cast<long long>(kAxisThreads / 32) + (threadIdx.x >> 5);
  const long long N = p.n_filters;
  const int T = p.n_steps;
  long long f = warp * 10 + lane / 3;
  const bool active = lane < 30 && f < N;
  if (warp * 10 >= N) return;  // whole warp beyond the batch
  if (f >= N) f = N - 1;

  AxisDense<AT, real> flt;
  flt.init_lane(lane, p.mc);
  const int a = flt.a;
  const bool lead = active && a == 0;
  flt.load(static_cast<const real*>(p.x) + f * n, static_cast<const real*>(p.P) + f * (n * n), n);

  // prod of the measured variances of an event (all three axes): det P+ = det P- * rprod / prod det S
  double rprod_imu = 1.0, rprod_gps = 1.0;
  static_for<0, NA>([&](auto I) {
    constexpr int g = AT::gidx0(I);
    if constexpr (AT::zk(I) != ZK_NONE) rprod_imu *= p.mc.r_imu[g] * p.mc.r_imu[g + 1] * p.mc.r_imu[g + 2];
    if constexpr (AT::gps0(I) >= 0) {
      constexpr int c = AT::gps0(I);
      rprod_gps *= p.mc.r_gps[c] * p.mc.r_gps[c + 1] * p.mc.r_gps[c + 2];
    }
  });

  const double gate = (GATED && p.gate_thresholds != nullptr) ? p.gate_thresholds[f] : p.gate_threshold;
  double acc_mant = 1.0;
  long long acc_max = 0;
  int acc_expo = 0;
  bool ok = true;
  double det = 1.0;

  const real* pay = static_cast<const real*>(p.payload) + f;
  const long long step_stride = static_cast<long long>(D) * N;
  real unext[NC];
  if (T > 0) {
    const real* src = pay + step_row(p.step_rows, 0) * step_stride;
    static_for<0, NC>([&](auto S) { unext[S] = __ldcs(src + static_cast<long long>(AT::col(S) + a) * N); });
  }

  for (int k = 0; k < T; ++k) {
    real u[NC];
    static_for<0, NC>([&](auto S) { u[S] = unext[S]; });
    if (k + 1 < T) {
      const real* src = pay + step_row(p.step_rows, k + 1) * step_stride;
      static_for<0, NC>([&](auto S) { unext[S] = __ldcs(src + static_cast<long long>(AT::col(S) + a) * N); });
    }
    const int ty = __ldg(p.types + k);
    const real dt = static_cast<real>(__ldg(p.dts + k));

    flt.predict(dt, real(p.mc.c2) * (dt * dt));
    const double dpre = flt.det(ok);
    bool do_update = (ty != EV_PREDICT_ONLY);
    bool mixed = false;
    if constexpr (GATED) {
      {  // kf_workers.py:1023-1025; the lanes of a filter follow its first lane's decision
        const int mine = (do_update && (fast_log_pos(dpre) > gate)) ? 1 : 0;
        do_update = __shfl_sync(0xffffffffu, mine, (lane / 3) * 3) != 0;
        const unsigned upd = __ballot_sync(0xffffffffu, do_update);
        mixed = upd != 0u && upd != 0xffffffffu;
      }
    }
    det = dpre;
    // The stages shuffle across the whole warp, so the warp updates as one: when the gate lets only some of its
    // filters update, all run the update and the others restore their predicted state afterwards.
    if (do_update || mixed) {
      AxisDense<AT, real> keep;
      if constexpr (GATED) {
        if (mixed) keep = flt;
      }
      if (ty == EV_GPS) {
        real z[NA];
        static_for<0, NA>([&](auto I) {
          if constexpr (AT::gps0(I) >= 0) z[I] = u[AT::slot(I, true)];
          else z[I] = real(0);
        });
        const double sprod = flt.template update<true>(z);
        det = dpre * rprod_gps / sprod;
      } else if (ty == EV_IMU) {
        real z[NA];
        flt.build_imu_z(dt, u, z);
        const double sprod = flt.template update<false>(z);
        det = dpre * rprod_imu / sprod;
      }
      if constexpr (GATED) {
        if (mixed && !do_update) {
          flt = keep;
          det = dpre;
        }
      }
    }
    ok = ok && (det > 0.0);
    {
      acc_max = max(acc_max, __double_as_longlong(det));
      const double m = acc_mant * det;
      const int hi = __double2hiint(m);
      const int e = ((hi >> 20) & 0x7ff) - 1023;
      acc_expo += e;
      acc_mant = __hiloint2double(hi - (e << 20), __double2loint(m));
    }

    if constexpr (FULL) {
      if (p.traj != nullptr && active) {
        real* tr = static_cast<real*>(p.traj) + (static_cast<long long>(k) * p.traj_ncomp) * N + f;
        const int nc = p.traj_ncomp;
        static_for<0, NA>([&](auto I) {
          const int g = AT::gidx0(I) + a;
          if (g < nc) __stcs(tr + static_cast<long long>(g) * N, flt.x[I]);
        });
      }
      if (p.ptraj != nullptr && active) {  // sf_KF_covariance.append(Pt.copy()), kf_workers.py:822
        real* pt = static_cast<real*>(p.ptraj) + (static_cast<long long>(k) * (n * n)) * N + f;
        flt.for_each_cov([&](int gi, int gj, real v) { pt[static_cast<long long>(gi * n + gj) * N] = v; });
      }
      if (p.logdet != nullptr && lead)
        __stcs(static_cast<real*>(p.logdet) + static_cast<long long>(k) * N + f, static_cast<real>(fast_log_pos(det)));
      if (p.used != nullptr && lead) p.used[static_cast<long long>(k) * N + f] = do_update ? 1 : 0;
    }
  }

  if (active) flt.store(static_cast<real*>(p.x) + f * n, static_cast<real*>(p.P) + f * (n * n), n);
  // status: a non-positive pivot seen by any lane of the filter
  const unsigned bad_lanes = __ballot_sync(0xffffffffu, !ok);
  if (p.agg != nullptr && lead) {
    double* ag = p.agg + f * 4;
    double mx = (T > 0) ? log(__longlong_as_double(acc_max)) : -INFINITY;
    double sm = log(acc_mant) + static_cast<double>(acc_expo) * 0.6931471805599453094;
    double cnt = static_cast<double>(T);
    if (p.agg_continue) {
      mx = fmax(mx, ag[0]);
      sm += ag[1];
      cnt += ag[3];
    }
    ag[0] = mx;
    ag[1] = sm;
    if (T > 0) ag[2] = log(det);
    else if (!p.agg_continue) ag[2] = 0.0;
    ag[3] = cnt;
  }
  if (p.status != nullptr && lead) {
    const int bad = (bad_lanes & flt.gmask) != 0 ? 1 : 0;
    p.status[f] = p.agg_continue ? (p.status[f] | bad) : bad;
  }
}

// ------------------------------------------------------------------------------------------
// TMA-staged variant for the closed-form filters (aggregate mode): one persistent 896-thread CTA
// per SM, each thread one filter.  The payload rows of a CTA's 896 filters are contiguous in the
// time-major layout (7 KB per (step, column)), so one elected thread streams them into a ring of
// shared-memory stages with cp.async.bulk (UBLKCP) completing on mbarriers, together with the
// packed (dt, type) records of the chunk.  The math warps never issue a global load in the loop:
// they wait on the stage's "full" barrier, read their 3 reals with conflict-free LDS, and one lane
// per warp arrives on the "empty" barrier.  Besides hiding HBM latency at 28 warps/SM, the ring
// bounds the skew between warps to kStages-1 chunks, which removes the long single-wave tail the
// LDG kernel shows (warps that win the issue arbitration finish early and leave the FP64 pipe
// half empty at the end).
// ------------------------------------------------------------------------------------------
constexpr int kTmaThreads = 896;            // math threads: 7 warps per SM sub-partition, one filter each
constexpr int kWsThreads = kTmaThreads + 32;  // + the producer warp (928 * 64 regs = 58 K registers)
#ifndef B200KF_CHUNK
#define B200KF_CHUNK 2
#endif
#ifndef B200KF_STAGES
#define B200KF_STAGES 5
#endif
constexpr int kChunk = B200KF_CHUNK;    // float32 launches: steps per stage
constexpr int kStages = B200KF_STAGES;  // float32 launches: ring depth (kStages * kChunk * d * 7 KB of shared memory)
// float64 launches take 3 stages of 3 steps (and three copy warps): at the power-capped clock the lean float64 loop is
// bound by its instruction issue, and a chunk costs every warp a try_wait + an arrive -- one per three steps instead
// of one per two is 3-4 % of the loop (headline 0.86 -> 0.89 at 1 470 MHz, 0.92 -> 0.95 at 1 580 MHz on the same boxes;
// CA15 / CA8 aggregate + 2-3 %).  The float32 rings keep 5 x 2 and four copy warps: their copy lanes are the bound
// (BIAS9 f32 loses 9 % with three).
#ifndef B200KF_STAGES_F64
#define B200KF_STAGES_F64 3
#endif
#ifndef B200KF_CHUNK_F64
#define B200KF_CHUNK_F64 3
#endif
template <typename real>
constexpr int base_stages() {
  return sizeof(real) == 8 ? B200KF_STAGES_F64 : B200KF_STAGES;
}
template <typename real>
constexpr int base_chunk() {
  return sizeof(real) == 8 ? B200KF_CHUNK_F64 : B200KF_CHUNK;
}
// Ring shape of the per-filter kernel.  The ring always holds the same BYTES (kStages * kChunk steps of
// 896 float64 filters, ~215 KB for d = 3): a CTA with fewer filters, or float32 payload, gets
// proportionally more steps.  Those go into LONGER chunks first (up to B200KF_MAX_CHUNK steps per stage):
// a chunk costs every warp one mbarrier.try_wait (~90 cycles even when the phase is complete) and the copy
// lanes one wait + expect_tx, which small CTAs (1.75 warps per scheduler) cannot hide; beyond that, into
// more stages.
#ifndef B200KF_MAX_CHUNK
#define B200KF_MAX_CHUNK 8
#endif
// (D = payload columns per step: 3 for the lean CV6 / BIAS9 rings the constants were tuned on; the constant-
// acceleration models carry 4 / 9 columns and get the same bytes: fewer steps, at least kCopyWarps stages)
template <typename real, int F, int D = 3>
constexpr int ring_steps() {
  return base_stages<real>() * base_chunk<real>() * (896 / F) * (sizeof(real) == 4 ? 2 : 1) * 3 / D;
}
template <typename real, int F, int D = 3>
constexpr int ring_chunk() {
  constexpr int want = ring_steps<real, F, D>() / base_stages<real>();
  return want < 1 ? 1 : (want < B200KF_MAX_CHUNK ? want : B200KF_MAX_CHUNK);
}
template <typename real, int F, int D = 3>
constexpr int ring_stages() {
  return ring_steps<real, F, D>() / ring_chunk<real, F, D>();
}

// step record of the ring: LeanRecT<real> (64 bytes in float64, 32 in float32; only (dt, type) are read
// outside the lean symmetric loop)
using SchedRec = LeanRecT<double>;

__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x100000u)  // suspend-time hint (ns): sleep in hardware instead of spinning
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// shared -> global bulk store (the per-step outputs of the trajectory mode leave the SM as whole rows)
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, unsigned bytes) {
#ifdef B200KF_BULK_EVICT_FIRST
  // streaming outputs: written once, never re-read by this launch -> do not let them displace the input rows in L2
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst),
               "r"(smem_u32(src)), "r"(bytes), "l"(pol)
               : "memory");
#else
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
               : "memory");
#endif
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {  // all but the N latest groups have READ their shared-memory source
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// make this thread's shared-memory writes visible to the async proxy (the bulk-copy engine)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Ring shape of the per-filter kernel WITH an output ring (trajectory mode through bulk stores): 4 input stages
// and 2 output stages of the same number of steps; at F = 896 float64 that is 4 x 21 KB in + 2 x 7 KB x rows out
// (rows <= kBulkMaxRows).  Smaller CTAs / float32 keep the bytes and take more steps per stage.
#ifndef B200KF_BULK_IN_STAGES
#define B200KF_BULK_IN_STAGES 4
#endif
#ifndef B200KF_BULK_OUT_STAGES
#define B200KF_BULK_OUT_STAGES 2
#endif
// Three-column payloads (CV6 / BIAS9: 24 B in, 56 B out per step) take 3 + 3 stages instead: the stores are 70 % of
// the traffic and a third output stage keeps the store lane fed (trajectory mode 5 540 -> 5 960 GB/s, 0.85 -> 0.91 of
// the HBM peak; the 72-byte-per-step CA15 ring measured 1 % slower with it and keeps 4 + 2).
template <int D>
constexpr int bulk_in_stages() {
  return D <= 3 ? 3 : B200KF_BULK_IN_STAGES;
}
template <int D>
constexpr int bulk_out_stages() {
  return D <= 3 ? 3 : B200KF_BULK_OUT_STAGES;
}
constexpr int kBulkMaxRows = 8;
#ifndef B200KF_BULK_CHUNK_D4
#define B200KF_BULK_CHUNK_D4 1
#endif
template <typename real, int F, int D = 3>
constexpr int bulk_chunk() {
  constexpr int c = (896 / F) * (sizeof(real) == 4 ? 2 : 1) * 3 / D;
  if (D == 4) return (c < 1 ? 1 : c) * B200KF_BULK_CHUNK_D4;  // four-column payload (CA8): room for longer chunks
  return c < 1 ? 1 : c;
}

// Per-step records of the launch: (dt, type) plus the warp-uniform constants of the lean symmetric
// step (Sym2Filter::LeanRec), which depend only on dt and the group's (q, r): evaluated once per
// step here instead of once per step and filter.
struct LeanConst {
  double q0, q1, r0, r1, rg;
};
template <typename real>
static __global__ void pack_schedule_kernel(const uint8_t* __restrict__ types, const double* __restrict__ dts,
                                            int T, LeanConst lc, LeanRecT<real>* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= T) return;
  const double dt = dts[k];
  const int ty = types[k];
  const double s10 = sqrt(lc.r1 / lc.r0), s01 = sqrt(lc.r0 / lc.r1);
  const double kap = dt * s10;
  LeanRecT<real> r{};
  r.dt = real(dt);
  r.dt2 = real(dt * dt);
  r.kap = real(kap);
  r.lam = real((ty == EV_GPS) ? s10 : dt * s01);
  r.kdt = real((ty == EV_GPS) ? lc.rg / lc.r0 : kap * dt);
  r.phi0 = real(lc.q0 * dt / lc.r0);
  r.phi1 = real(lc.q1 * dt / lc.r1);
  r.type = ty;
  out[k] = r;
}

// filter object of the per-filter-covariance ring kernel: per-chain / axis-symmetric closed form for the
// (value, rate) models (CV6), the lean 3-state control-chain form for BIAS9 (symmetric only)
template <int MODEL, typename real, bool SYM>
struct TmaFlt {
  using MT = ModelTraits<MODEL, false, real>;
  using type = std::conditional_t<
      !SYM, typename MT::Dec2,
      std::conditional_t<MT::Filter::fast2, typename MT::Sym2,
                         std::conditional_t<MT::Sym3::eligible, typename MT::Sym3, typename MT::SymCA>>>;
};

// step records of the constant-acceleration ring filter (SymCAFilter): everything that depends on (dt, q, r) only
struct CAConst {
  double r0, r1, r2, q0, q1, q2, b0, b1, qa0, qa1, rg, c2;
};
template <typename real>
static __global__ void pack_ca_schedule_kernel(const uint8_t* __restrict__ types, const double* __restrict__ dts, int T,
                                               CAConst c, CARecT<real>* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= T) return;
  const double dt = dts[k], e = c.c2 * dt * dt;
  CARecT<real> r{};
  r.dt = real(dt);
  r.e = real(e);
  r.k01 = real(dt * sqrt(c.r1 / c.r0));
  r.k02 = real(e * sqrt(c.r2 / c.r0));
  r.k12 = real(dt * sqrt(c.r2 / c.r1));
  r.f0 = real(c.q0 * dt / c.r0);
  r.f1 = real(c.q1 * dt / c.r1);
  r.f2 = real(c.q2 * dt / c.r2);
  r.ka = real(dt * sqrt(c.b1 / c.b0));
  r.fa0 = real(c.qa0 * dt / c.b0);
  r.fa1 = real(c.qa1 * dt / c.b1);
  r.gam = real(c.rg / c.r0);
  r.type = types[k];
  out[k] = r;
}

// The copy lane of every ring kernel: keeps the payload ring full, gated only by the ring's empty
// barriers.  Stage s holds CH steps x D rows of F reals; the rows of one (step, column) are contiguous in
// the time-major payload, so each is one cp.async.bulk completing on the stage's full barrier.  With
// `sched` the chunk's per-step records (packed schedule, or precomputed gain records) travel with it.  (An L2 prefetch of later rows with
// cp.async.bulk.prefetch.L2 was measured 8-25 % slower for the deep d = 3 ring and neutral for the
// shallow d = 9 one, and is not issued.)
template <typename real, int D, int F, int ST, int CH, class AuxRec = SchedRec>
__device__ __forceinline__ void ring_copy_loop(real* stage_buf, unsigned long long* full, unsigned long long* empty,
                                               const real* pay, long long N, int T, unsigned row_bytes,
                                               AuxRec* sched_buf = nullptr, const AuxRec* sched = nullptr,
                                               const int* __restrict__ step_rows = nullptr, int first_chunk = 0,
                                               int chunk_stride = 1) {
  // several copy lanes (one per warp, on different schedulers) may share the ring: lane w of W takes the
  // chunks congruent to w modulo W; the barriers do not care who fills a stage
  static_assert(sizeof(AuxRec) % 16 == 0, "bulk copies move multiples of 16 bytes");
  constexpr size_t kStageReals = static_cast<size_t>(CH) * D * F;
  const int nchunks = (T + CH - 1) / CH;
  // step_rows (a selection of a longer stream): the row indices of this lane's NEXT chunk are fetched before the
  // wait for the current one, so the global-load latency (~0.5 us, once per chunk on the lane's serial path: it
  // capped a launch at ~2.5 steps/us = 725 GB/s whatever the CTA count) hides behind the wait and the copies
  int rows_nx[CH] = {};
  auto fetch_rows = [&](int chunk) {
#pragma unroll
    for (int c = 0; c < CH; ++c) rows_nx[c] = (chunk * CH + c < T) ? __ldg(step_rows + chunk * CH + c) : 0;
  };
  if (step_rows != nullptr && first_chunk < nchunks) fetch_rows(first_chunk);
  for (int chunk = first_chunk; chunk < nchunks; chunk += chunk_stride) {
    const int s = chunk % ST;
    int rows_cur[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) rows_cur[c] = rows_nx[c];
    if (step_rows != nullptr && chunk + chunk_stride < nchunks) fetch_rows(chunk + chunk_stride);
    if (chunk >= ST) mbar_wait(&empty[s], static_cast<unsigned>(chunk / ST - 1) & 1u);
    const int k0 = chunk * CH;
    const int steps = (T - k0) < CH ? (T - k0) : CH;
    const unsigned rec_bytes = (sched != nullptr) ? static_cast<unsigned>(steps * sizeof(AuxRec)) : 0u;
    mbar_expect_tx(&full[s], static_cast<unsigned>(steps) * D * row_bytes + rec_bytes);
    if (sched != nullptr) bulk_g2s(sched_buf + s * CH, sched + k0, rec_bytes, &full[s]);
    real* dst = stage_buf + s * kStageReals;
    if (step_rows == nullptr) {  // consecutive rows: one running pointer for the steps * D copies of the chunk
      const real* src = pay + static_cast<long long>(k0) * D * N;
      for (int r = 0; r < steps * D; ++r, src += N, dst += F) bulk_g2s(dst, src, row_bytes, &full[s]);
    } else {
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        if (c < steps) {
          const long long row = rows_cur[c];
#pragma unroll
          for (int j = 0; j < D; ++j) bulk_g2s(dst + (c * D + j) * F, pay + (row * D + j) * N, row_bytes, &full[s]);
        }
      }
    }
  }
}

// copy warps of the per-filter ring kernel: one per scheduler, so the copy lanes' issue slots (about 15
// instructions per bulk copy, each waiting its turn behind seven math warps) come off four schedulers
// instead of one; with a single copy warp the float32 loop (two copies per 100 issue cycles and warp) ran
// the ring dry 18 % of the time (profiles/r01c_ncu_f32.txt)
#ifndef B200KF_COPY_WARPS
#define B200KF_COPY_WARPS 4
#endif
constexpr int kCopyWarps = B200KF_COPY_WARPS;  // float32 launches
#ifndef B200KF_COPY_WARPS_F64
#define B200KF_COPY_WARPS_F64 3
#endif
// service warps of a model's per-filter ring kernel.  The constant-acceleration filter (24 state / covariance values
// per thread) gets two: 448 math threads + 64 = 512 threads leave 128 registers per thread (576 threads round to 96),
// and its FP64-bound loop (~160 FP64 instructions against 9 bulk copies per step) needs no more copy lanes.
template <int MODEL, typename real>
constexpr int svc_warps() {
  using MT = ModelTraits<MODEL, false, real>;
  return (!MT::Filter::fast2 && !MT::Sym3::eligible && MT::SymCA::eligible)
             ? 2
             : (sizeof(real) == 8 ? B200KF_COPY_WARPS_F64 : kCopyWarps);
}

template <int MODEL, typename real, bool SYM, bool FULL, int F, bool BULK = false>
__global__ void __launch_bounds__(F + 32 * svc_warps<MODEL, real>(), 1)
    kf_run_tma_kernel(const RunParams p, const void* __restrict__ sched_v) {
  constexpr int kCopyWarps = svc_warps<MODEL, real>();  // (shadows the default: this model's service warps)
  // F math threads (filters) per CTA: 896 fills an SM with one CTA; 448 / 224 keep every SM busy when a
  // GPU's shard is small (N <= 148 * 448), with a proportionally deeper ring in the same shared memory
  // BULK (per-step outputs only): the math threads write a step's outputs (trajectory rows, log det) into a
  // shared-memory OUTPUT ring and the last service warp stores them with cp.async.bulk.global.shared -- whole
  // 7 KB rows, no per-thread global stores in the loop (the per-thread STGs of the plain FULL kernel back up
  // the store queue: long-scoreboard waits of 6.6 per issue, profiles/r01c_ncu_trajectory_kernel.txt).
  static_assert(!BULK || FULL, "the output ring belongs to the per-step-output kernel");
  constexpr int kTmaThreads = F;
  constexpr int kD = ModelTraits<MODEL, false, real>::d;
  constexpr int kOutStages = bulk_out_stages<kD>();
  constexpr int kStages = BULK ? bulk_in_stages<kD>() : ring_stages<real, F, kD>();
  constexpr int kChunk = BULK ? bulk_chunk<real, F, kD>() : ring_chunk<real, F, kD>();
  constexpr int kLoadWarps = BULK ? kCopyWarps - 1 : kCopyWarps;  // BULK: the last service warp stores
  // a copy lane waits on a stage's empty barrier by phase PARITY: that is only unambiguous while it can
  // never be two uses of a stage ahead, i.e. while its stride (kLoadWarps chunks) is within one ring turn
  static_assert(kStages >= kLoadWarps && kLoadWarps >= 1, "ring shallower than the copy lanes' stride");
  using MT = ModelTraits<MODEL, false, real>;
  using Flt = typename TmaFlt<MODEL, real, SYM>::type;
  constexpr int n = MT::n;
  constexpr int D = MT::d;
  static_assert(MT::Filter::fast2 || (SYM && (MT::Sym3::eligible || MT::SymCA::eligible)),
                "ring kernel is for the closed-form filters");
  if (!mode_selected(p.dense_flag, SYM ? MODE_SYMMETRIC : MODE_DECOUPLED)) return;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  real* stage_buf = reinterpret_cast<real*>(smem_raw);  // [kStages][kChunk][D][kTmaThreads]
  constexpr size_t kStageReals = static_cast<size_t>(kChunk) * D * kTmaThreads;
  using Rec = typename Flt::LeanRec;  // the filter's step record (LeanRecT, or CARecT for the constant-acceleration form)
  const Rec* __restrict__ sched = static_cast<const Rec*>(sched_v);
  // output ring (BULK): [kOutStages][kChunk][out_rows][kTmaThreads], rows = the trajectory components, then log det
  const int out_nc = (FULL && p.traj != nullptr) ? p.traj_ncomp : 0;
  const int out_rows = out_nc + ((FULL && p.logdet != nullptr) ? 1 : 0);
  real* out_buf = stage_buf + kStages * kStageReals;
  const size_t out_stage_reals = BULK ? static_cast<size_t>(kChunk) * out_rows * kTmaThreads : 0;
  Rec* sched_buf = reinterpret_cast<Rec*>(out_buf + (BULK ? kOutStages : 0) * out_stage_reals);  // [kStages][kChunk]
  unsigned long long* full = reinterpret_cast<unsigned long long*>(sched_buf + kStages * kChunk);
  unsigned long long* empty = full + kStages;
  unsigned long long* out_full = empty + kStages;     // BULK: a chunk's outputs are staged (all math warps arrived)
  unsigned long long* out_empty = out_full + kOutStages;  // BULK: the stage's bulk stores have read it

  const int tid = threadIdx.x;
  const long long N = p.n_filters;
  const long long base = static_cast<long long>(blockIdx.x) * kTmaThreads;
  const int nfb = static_cast<int>((N - base) < kTmaThreads ? (N - base) : kTmaThreads);
  const long long f = base + tid;
  const bool active = tid < nfb;
  const int T = p.n_steps;
  const int nchunks = (T + kChunk - 1) / kChunk;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kTmaThreads / 32);
    }
    if constexpr (BULK) {
      for (int s = 0; s < kOutStages; ++s) {
        mbar_init(&out_full[s], kTmaThreads / 32);
        mbar_init(&out_empty[s], 1);
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (tid >= kTmaThreads) {
    // ---- copy warps: one lane each keeps the ring full, decoupled from the math warps' progress ----
    const int sw = (tid - kTmaThreads) >> 5;
    if (((tid - kTmaThreads) & 31) != 0) return;
    if (sw < kLoadWarps) {
      ring_copy_loop<real, D, F, kStages, kChunk, Rec>(stage_buf, full, empty,
                                                       static_cast<const real*>(p.payload) + base, N, T,
                                                       static_cast<unsigned>(nfb * sizeof(real)), sched_buf, sched,
                                                       p.step_rows, sw, kLoadWarps);
    } else if constexpr (BULK) {
      // ---- store warp: one lane drains the output ring with bulk stores, a chunk per bulk group ----
      const unsigned row_bytes = static_cast<unsigned>(nfb * sizeof(real));
      real* trg = (p.traj != nullptr) ? static_cast<real*>(p.traj) + base : nullptr;
      real* ldg = (p.logdet != nullptr) ? static_cast<real*>(p.logdet) + base : nullptr;
      for (int chunk = 0; chunk < nchunks; ++chunk) {
        const int os = chunk % kOutStages;
        mbar_wait(&out_full[os], static_cast<unsigned>(chunk / kOutStages) & 1u);
        const int k0 = chunk * kChunk;
        const int steps = (T - k0) < kChunk ? (T - k0) : kChunk;
        const real* src = out_buf + os * out_stage_reals;
        for (int c = 0; c < steps; ++c) {
          const long long k = k0 + c;
          for (int g = 0; g < out_nc; ++g)
            bulk_s2g(trg + (k * out_nc + g) * N, src + (static_cast<size_t>(c) * out_rows + g) * kTmaThreads, row_bytes);
          if (ldg != nullptr)
            bulk_s2g(ldg + k * N, src + (static_cast<size_t>(c) * out_rows + out_nc) * kTmaThreads, row_bytes);
        }
        bulk_commit();
        if constexpr (kOutStages >= 3) {
          // one bulk group in flight behind the one just committed: the previous chunk's stage is released once ITS
          // stores have read it, while this chunk's are already issued (with only two output stages this makes the
          // math warps wait a chunk longer for a free stage: -8 % on the CA15 ring)
          if (chunk > 0) {
            bulk_wait_read<1>();
            mbar_arrive(&out_empty[(chunk - 1) % kOutStages]);
          }
        } else {
          bulk_wait_read<0>();  // the stores have READ the stage (not yet landed in HBM): the math warps may refill it
          mbar_arrive(&out_empty[os]);
        }
      }
      if constexpr (kOutStages >= 3) {
        if (nchunks > 0) {
          bulk_wait_read<0>();
          mbar_arrive(&out_empty[(nchunks - 1) % kOutStages]);
        }
      }
      bulk_wait_all();  // the launch's outputs are complete when the kernel ends
    }
    return;
  }

  Flt flt;
  if (active) {
    const real* xg = static_cast<const real*>(p.x) + f * n;
    const real* Pg = static_cast<const real*>(p.P) + f * (n * n);
    flt.load(xg, Pg, n, p.mc);
  }
  const double gate = (FULL && active && p.gate_thresholds != nullptr) ? p.gate_thresholds[f] : p.gate_threshold;
  double acc_mant = 1.0;
  long long acc_max = 0;  // max det as raw bits: positive doubles order like their bit patterns
  int acc_expo = 0;
  bool ok = true;
  // lean aggregate of the symmetric filter: IMU events contribute det / (r0 r1) of ONE chain, the
  // constant and the chain multiplicity are restored after the loop; validity (every det finite
  // and > 0) is tracked with integer min / max of the raw bits instead of an FP64 compare
  constexpr bool LEAN = SYM && !FULL && Flt::lean_ok;
  // float32 launches keep the lean aggregate in float32 too (no conversion, no FP64 in the loop): the
  // running product's mantissa picks up <= 2^-24 relative per step, i.e. <= T 2^-24 absolute on a sum of
  // T logarithms
  constexpr bool LEAN32 = LEAN && sizeof(real) == 4;
  constexpr int kRenorm = 4;  // float64 lean aggregate: steps between renormalisations of the running product
  // validity of the lean aggregate: integer min / max of the raw bits (a determinant <= 0 has a high word
  // <= 0, NaN / inf exceed the largest finite pattern).  The running product is renormalised every step by
  // masking its exponent field; the biased exponents are summed and the bias removed after the loop.
  int min_hi = 0x7ff00000;
  float mant32 = 1.0f;
  int max32 = 0;
  // filters that live in R-normalised coordinates: the lean aggregate loop, and the record-driven filter in every mode
  constexpr bool NORM = LEAN || Flt::rec_api;
  if constexpr (NORM) {
    if (active && T > 0) flt.to_normalized(p.mc);
  }

  real* tr_run = (FULL && !BULK && p.traj != nullptr) ? static_cast<real*>(p.traj) + f : nullptr;
  real* ld_run = (FULL && !BULK && p.logdet != nullptr) ? static_cast<real*>(p.logdet) + f : nullptr;
  unsigned char* used_run = (FULL && !BULK && p.used != nullptr) ? p.used + f : nullptr;
  for (int chunk = 0; chunk < nchunks; ++chunk) {
    const int s = chunk % kStages;
    const unsigned use = static_cast<unsigned>(chunk / kStages);
    mbar_wait(&full[s], use & 1u);
    const int k0 = chunk * kChunk;
    const int steps = (T - k0) < kChunk ? (T - k0) : kChunk;
    const int os = chunk % kOutStages;
    if constexpr (BULK) {
      if (chunk >= kOutStages) mbar_wait(&out_empty[os], static_cast<unsigned>(chunk / kOutStages - 1) & 1u);
    }
    if (active) {
      const real* sb = stage_buf + s * kStageReals + tid;
      real* ob = out_buf + os * out_stage_reals + tid;  // BULK: this thread's column of the output stage
      auto do_step = [&](int c) {
        const Rec& rec = sched_buf[s * kChunk + c];
        real u[D];
#pragma unroll
        for (int j = 0; j < D; ++j) u[j] = sb[(c * D + j) * kTmaThreads];
        if constexpr (LEAN) {
          real v;
          flt.step_lean(rec, p.mc, u, v);
          if constexpr (LEAN32) {
            const int vb = __float_as_int(v);
            max32 = max(max32, vb);
            min_hi = min(min_hi, vb);
            const unsigned mb = __float_as_uint(mant32 * v);
            acc_expo += static_cast<int>(mb >> 23);
            mant32 = __uint_as_float((mb & 0x007fffffu) | 0x3f800000u);
          } else {
            min_hi = min(min_hi, __double2hiint(double(v)));
            acc_max = max(acc_max, __double_as_longlong(double(v)));
#ifdef B200KF_AGG_RENORM_EVERY_STEP
            const double m = acc_mant * double(v);
            const unsigned hi = static_cast<unsigned>(__double2hiint(m));
            acc_expo += static_cast<int>(hi >> 20);
            acc_mant = __hiloint2double(static_cast<int>((hi & 0x000fffffu) | 0x3ff00000u), __double2loint(m));
#else
            // the running product is renormalised once per kRenorm steps (below), not per step: between two
            // renormalisations it holds a mantissa in [1, 2) times at most kRenorm determinants, each checked to
            // lie in [2^-240, 2^240] (ok flag) -- no overflow, the same sum of logarithms
            acc_mant *= double(v);
            if (((c + 1) % kRenorm) == 0 || c + 1 == steps) {
              const unsigned hi = static_cast<unsigned>(__double2hiint(acc_mant));
              acc_expo += static_cast<int>(hi >> 20) - 1023;
              acc_mant = __hiloint2double(static_cast<int>((hi & 0x000fffffu) | 0x3ff00000u), __double2loint(acc_mant));
            }
#endif
          }
          return;
        }
        const real dt = rec.dt;
        const int ty = rec.type;
        if constexpr (Flt::rec_api) flt.predict(rec);
        else flt.predict(dt, p.mc);
        bool do_update = (ty != EV_PREDICT_ONLY);
        if constexpr (FULL) {
          if (p.gate_enabled) do_update = do_update && (fast_log_pos(flt.det(ok)) > gate);  // :1023-1025
        }
        double det;
        if constexpr (Flt::rec_api) det = do_update ? flt.update(rec, ty, u, ok) : flt.det(ok);
        else det = do_update ? flt.update(dt, p.mc, ty, u, ok) : flt.det(ok);
        if constexpr (BULK) {
          real* o = ob + static_cast<size_t>(c) * out_rows * kTmaThreads;
          flt.for_each_state([&](auto G, real v) {
            constexpr int g = G;
            if (g < out_nc) o[g * kTmaThreads] = v;
          });
          if (out_rows > out_nc) o[out_nc * kTmaThreads] = static_cast<real>(fast_log_pos(det));
        } else if constexpr (FULL) {
          // running output pointers: no reload of the launch parameters and no 64-bit multiply per step
          if (tr_run != nullptr) {
            flt.for_each_state([&](auto G, real v) {
              constexpr int g = G;
              if (g < out_nc) __stcs(tr_run + static_cast<long long>(g) * N, v);
            });
            tr_run += static_cast<long long>(out_nc) * N;
          }
          if (ld_run != nullptr) {
            __stcs(ld_run, static_cast<real>(fast_log_pos(det)));
            ld_run += N;
          }
          if (used_run != nullptr) {
            *used_run = do_update ? 1 : 0;
            used_run += N;
          }
        }
        acc_max = max(acc_max, __double_as_longlong(det));
        const double m = acc_mant * det;
        const int hi = __double2hiint(m);
        const int e = ((hi >> 20) & 0x7ff) - 1023;
        acc_expo += e;
        acc_mant = __hiloint2double(hi - (e << 20), __double2loint(m));
      };
      if (steps == kChunk) {  // every chunk but the last: no per-step bound check
#pragma unroll
        for (int c = 0; c < kChunk; ++c) do_step(c);
      } else {
        for (int c = 0; c < steps; ++c) do_step(c);
      }
    }
    if constexpr (BULK) fence_async_smem();  // this thread's staged outputs -> visible to the bulk-copy engine
    __syncwarp();
    if ((tid & 31) == 0) {
      mbar_arrive(&empty[s]);
      if constexpr (BULK) mbar_arrive(&out_full[os]);
    }
  }

  if (!active) return;
  double last_det = 1.0;  // record-driven filters: det P of the final state, taken in their normalised coordinates
  if constexpr (Flt::rec_api) {
    if (T > 0) last_det = flt.det(ok);
  }
  if constexpr (LEAN) {
    if (T > 0) {
      if constexpr (LEAN32) {
        ok = ok && min_hi > 0 && max32 < 0x7f800000;
        acc_mant = double(mant32);
        acc_max = __double_as_longlong(double(__int_as_float(max32)));
        acc_expo -= 127 * T;
      } else {
#ifdef B200KF_AGG_RENORM_EVERY_STEP
        ok = ok && min_hi > 0 && acc_max < 0x7ff0000000000000LL;
        acc_expo -= 1023 * T;
#else
        // every (normalised) determinant within [2^-240, 2^240]: the renormalise-every-kRenorm-steps product cannot
        // have left the double range (a covariance beyond that is flagged like a non-positive one)
        ok = ok && min_hi >= ((1023 - 240) << 20) && acc_max < (static_cast<long long>(1023 + 240) << 52);
#endif
      }
      flt.from_normalized(p.mc);
    }
  } else if constexpr (NORM) {
    if (T > 0) flt.from_normalized(p.mc);
  }
  {
    real* xg = static_cast<real*>(p.x) + f * n;
    real* Pg = static_cast<real*>(p.P) + f * (n * n);
    flt.store(xg, Pg, n);
  }
  if (p.agg != nullptr) {
    double* a = p.agg + f * 4;
    double mx = (T > 0) ? log(__longlong_as_double(acc_max)) : -INFINITY;
    double sm = log(acc_mant) + static_cast<double>(acc_expo) * 0.6931471805599453094;
    if constexpr (LEAN) {  // every term was det / (r0 r1) of one chain: restore the constant, then the multiplicity
      const double lr = log(double(flt.r0r1));
      mx = (mx + lr) * static_cast<double>(Flt::M);
      sm = (sm + static_cast<double>(T) * lr) * static_cast<double>(Flt::M);
    }
    double cnt = static_cast<double>(T);
    if (p.agg_continue) {
      mx = fmax(mx, a[0]);
      sm += a[1];
      cnt += a[3];
    }
    a[0] = mx;
    a[1] = sm;
    if constexpr (Flt::rec_api) a[2] = (T > 0) ? log(last_det) : (p.agg_continue ? a[2] : 0.0);
    else a[2] = (T > 0) ? log(flt.det(ok)) : (p.agg_continue ? a[2] : 0.0);
    a[3] = cnt;
  }
  if (p.status != nullptr) {
    const int bad = ok ? 0 : 1;
    p.status[f] = p.agg_continue ? (p.status[f] | bad) : bad;
  }
}

// ------------------------------------------------------------------------------------------
// Shared-covariance variant.  The covariance / gain recursion never reads a measurement
// (kf_workers.py:546-549, 616-621, 711 touch no z) and the event schedule is shared by the launch,
// so when every filter ALSO starts from the same P -- Monte-Carlo seeds of one rate (BASELINE
// configs[1]), or the reference's fixed P0 (kf_workers.py:651) for a whole batch -- all filters
// have the same gains at every step.  One lane of a dedicated producer warp per CTA runs that
// recursion ONCE (in float64, whatever the model precision), writes a 64-byte gain record per step
// into the ring stage next to the TMA'd payload rows, and issues the bulk copies; the 28 math warps
// only apply the gains to their filter's state: 4 FP64 per axis and event.  Validity (all P
// bit-identical, axis-symmetric, chain-compatible; equal rate constants) is checked per launch on
// the device like the other variants.  The kernel is a pure HBM stream.
// ------------------------------------------------------------------------------------------
constexpr int kMathThreads = kTmaThreads;           // 896 consumer threads, one filter each
constexpr int kSharedThreads = kMathThreads + 64;   // + the gain warp + the copy warp

template <typename real>
struct GainRecT {  // 56 bytes in float64, 32 in float32: the gain warp rounds once per step, not every thread
  real dt;
  real g00, g01, g10, g11;  // IMU: x0 += g00 v + g01 u, x1 = g10 v + g11 u (predict folded in); GPS: (k0, -, k1, -)
  real logdet;              // log det P of the whole filter after the step (per-step output only)
  int type;                 // event type actually applied
  int pad;
};

// ring shape of the CV6 shared-covariance kernel: 5 stages; float32 and the smaller CTAs take longer chunks
// (same reasoning as SharedRing / ring_chunk), 224-filter CTAs half the bytes so that two share an SM
template <typename real, int F>
constexpr int shared_chunk() {
  return base_chunk<real>() * (sizeof(real) == 4 ? 2 : 1) * (F == 896 ? 1 : 2);
}

template <int MODEL, typename real, bool FULL, int F>
__global__ void __launch_bounds__(F + 64, F == 224 ? 2 : 1) kf_run_shared_kernel(const RunParams p) {
  constexpr int kMathThreads = F;
  constexpr int kStages = base_stages<real>();
  constexpr int kChunk = shared_chunk<real, F>();
  using GainRec = GainRecT<real>;
  using MT = ModelTraits<MODEL, false, real>;
  using S2 = typename MT::Sym2;
  using B0 = typename S2::Lead;
  constexpr int n = MT::n;
  constexpr int D = MT::d;
  constexpr int M = S2::M;
  static_assert(S2::lean_ok, "shared-covariance kernel: (position, velocity) pseudo-measurement chains only");
  if (!mode_selected(p.dense_flag, MODE_SHARED)) return;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  real* stage_buf = reinterpret_cast<real*>(smem_raw);  // [kStages][kChunk][D][kMathThreads]
  constexpr size_t kStageReals = static_cast<size_t>(kChunk) * D * kMathThreads;
  GainRec* gain_buf = reinterpret_cast<GainRec*>(stage_buf + kStages * kStageReals);  // [kStages][kChunk]
  static_assert((kStages * kChunk * sizeof(GainRec)) % 8 == 0, "aggregates behind the records must stay aligned");
  double* fin = reinterpret_cast<double*>(gain_buf + kStages * kChunk);  // a, b, c, max, sum, last, ok
  unsigned long long* full = reinterpret_cast<unsigned long long*>(fin + 8);
  unsigned long long* empty = full + kStages;

  const int tid = threadIdx.x;
  const long long N = p.n_filters;
  const long long base = static_cast<long long>(blockIdx.x) * kMathThreads;
  const int nfb = static_cast<int>((N - base) < kMathThreads ? (N - base) : kMathThreads);
  const int T = p.n_steps;
  const int nchunks = (T + kChunk - 1) / kChunk;
  const unsigned row_bytes = static_cast<unsigned>(nfb * sizeof(real));

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 2);  // the copy warp's expect_tx arrival + the gain warp's arrival
      mbar_init(&empty[s], kMathThreads / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (tid >= kMathThreads + 32) {
    // ---------------- copy warp: one lane keeps the payload ring full ------------------------------
    if (tid == kMathThreads + 32)
      ring_copy_loop<real, D, F, kStages, kChunk>(stage_buf, full, empty, static_cast<const real*>(p.payload) + base, N, T,
                                                  row_bytes, static_cast<SchedRec*>(nullptr), static_cast<const SchedRec*>(nullptr), p.step_rows);
  } else if (tid >= kMathThreads) {
    // ---------------- gain warp: the one covariance recursion of the launch ------------------------
    if (tid == kMathThreads) {
      // every filter carries the same P: read this CTA's first filter (only this CTA writes it, at the end)
      const real* Pg = static_cast<const real*>(p.P) + base * (n * n);
      double a = double(Pg[B0::gidx(0) * n + B0::gidx(0)]);
      double b = 0.5 * (double(Pg[B0::gidx(0) * n + B0::gidx(1)]) + double(Pg[B0::gidx(1) * n + B0::gidx(0)]));
      double c = double(Pg[B0::gidx(1) * n + B0::gidx(1)]);
      const double q0 = p.mc.q[B0::gidx(0)], q1 = p.mc.q[B0::gidx(1)];
      const double r0 = p.mc.r_imu[B0::gidx(0)], r1 = p.mc.r_imu[B0::gidx(1)];
      const double rg = p.mc.r_gps[B0::gps(0)];
      const double r0r1 = r0 * r1, ir0 = 1.0 / r0, ir1 = 1.0 / r1;
      LogDetAcc acc;
      bool ok = true;
      // the step's (dt, type) are fetched one step ahead: a global load per step would sit on the serial chain
      double dt_nx = (T > 0) ? __ldg(p.dts) : 0.0;
      int ty_nx = (T > 0) ? __ldg(p.types) : 0;
      for (int chunk = 0; chunk < nchunks; ++chunk) {
        const int s = chunk % kStages;
        if (chunk >= kStages) mbar_wait(&empty[s], static_cast<unsigned>(chunk / kStages - 1) & 1u);
        const int k0 = chunk * kChunk;
        const int steps = (T - k0) < kChunk ? (T - k0) : kChunk;
        for (int cc = 0; cc < steps; ++cc) {
          const double dt = dt_nx;
          const int ty = ty_nx;
          if (k0 + cc + 1 < T) {
            dt_nx = __ldg(p.dts + k0 + cc + 1);
            ty_nx = __ldg(p.types + k0 + cc + 1);
          }
          GainRecT<double> rec;
          rec.dt = dt;
          rec.type = ty;
          rec.pad = 0;
          rec.g00 = rec.g01 = rec.g10 = rec.g11 = 0.0;
          // same formulas as Sym2Filter::step_lean, evaluated once for the launch
          const double bn = fma(dt, c, b);
          const double ap = fma(dt, b + bn, fma(q0, dt, a));
          const double cp = fma(q1, dt, c);
          double det;
          if (ty == EV_IMU) {
            const double detp = fma(ap, cp, -(bn * bn));
            const double t0 = fma(ap, r1, detp), t1 = fma(cp, r0, detp);
            const double inv = fast_rcp3(fma(r0, cp + r1, t0));
            const double i0 = r0 * inv, i1 = r1 * inv;
            a = t0 * i0;
            c = t1 * i1;
            b = bn * (i0 * r1);
            const double w0 = dt * ir0, w1 = dt * ir1;
            const double g00 = a * w0, g10 = b * w0, g01 = b * w1, g11 = c * w1;
            rec.g00 = dt + g00;           // x0+ = x0 + (dt + g00) v + (g00 dt + g01) u
            rec.g01 = fma(g00, dt, g01);
            rec.g10 = 1.0 + g10;          // x1+ = (1 + g10) v + (g10 dt + g11) u
            rec.g11 = fma(g10, dt, g11);
            det = detp * inv * r0r1;
          } else if (ty == EV_GPS) {
            const double inv = fast_rcp3(ap + rg);
            rec.g00 = ap * inv;
            rec.g10 = bn * inv;
            const double g = rg * inv;
            c = fma(-rec.g10, bn, cp);
            a = ap * g;
            b = bn * g;
            det = fma(a, c, -(b * b));
          } else {
            a = ap;
            b = bn;
            c = cp;
            det = fma(a, c, -(b * b));
          }
          ok = ok && (det > 0.0) && (det < INFINITY);
          const double detM = S2::powM(det);
          acc.add(detM);
          rec.logdet = 0.0;
          if constexpr (FULL) {
            if (p.logdet != nullptr) rec.logdet = fast_log_pos(detM);
          }
          if constexpr (sizeof(real) == 8) {
            gain_buf[s * kChunk + cc] = rec;
          } else {
            GainRec r32;
            r32.dt = real(rec.dt), r32.g00 = real(rec.g00), r32.g01 = real(rec.g01), r32.g10 = real(rec.g10);
            r32.g11 = real(rec.g11), r32.logdet = real(rec.logdet), r32.type = rec.type, r32.pad = 0;
            gain_buf[s * kChunk + cc] = r32;
          }
        }
        mbar_arrive(&full[s]);  // release: the gain records are visible to whoever sees the phase flip
      }
      fin[0] = a;
      fin[1] = b;
      fin[2] = c;
      fin[3] = (T > 0) ? log(acc.maxdet) : -INFINITY;
      fin[4] = acc.sum_log();
      fin[5] = (T > 0) ? log(acc.last) : 0.0;
      fin[6] = ok ? 1.0 : 0.0;
    }
  } else {
    // ---------------- math warps: x += G [v + u dt, u] --------------------------------------------
    const bool active = tid < nfb;
    const long long f = base + tid;
    real x0[M], x1[M];
    if (active) {
      const real* xg = static_cast<const real*>(p.x) + f * n;
      S2::members([&](auto Bm, int m) {
        using BM = decltype(Bm);
        x0[m] = xg[BM::gidx(0)];
        x1[m] = xg[BM::gidx(1)];
      });
    }
    for (int chunk = 0; chunk < nchunks; ++chunk) {
      const int s = chunk % kStages;
      mbar_wait(&full[s], static_cast<unsigned>(chunk / kStages) & 1u);
      const int k0 = chunk * kChunk;
      const int steps = (T - k0) < kChunk ? (T - k0) : kChunk;
      if (active) {
        const real* sb = stage_buf + s * kStageReals + tid;
#pragma unroll
        for (int cc = 0; cc < kChunk; ++cc) {
          if (cc < steps) {
            const GainRec* rec = gain_buf + s * kChunk + cc;
            const real dt = static_cast<real>(rec->dt);
            const int ty = rec->type;
            real u[D];
#pragma unroll
            for (int j = 0; j < D; ++j) u[j] = sb[(cc * D + j) * kMathThreads];
            if (ty == EV_IMU) {
              const real g00 = static_cast<real>(rec->g00), g01 = static_cast<real>(rec->g01);
              const real g10 = static_cast<real>(rec->g10), g11 = static_cast<real>(rec->g11);
              S2::members([&](auto Bm, int m) {
                using BM = decltype(Bm);
                const real acc = u[BM::pz(1)];
                x0[m] = fma(g00, x1[m], fma(g01, acc, x0[m]));
                x1[m] = fma(g11, acc, g10 * x1[m]);
              });
            } else if (ty == EV_GPS) {
              const real k0g = static_cast<real>(rec->g00), k1g = static_cast<real>(rec->g10);
              S2::members([&](auto Bm, int m) {
                using BM = decltype(Bm);
                const real xp = fma(dt, x1[m], x0[m]);
                const real y = u[BM::gps(0)] - xp;
                x0[m] = fma(k0g, y, xp);
                x1[m] = fma(k1g, y, x1[m]);
              });
            } else {
#pragma unroll
              for (int m = 0; m < M; ++m) x0[m] = fma(dt, x1[m], x0[m]);
            }
            if constexpr (FULL) {
              const long long k = k0 + cc;
              if (p.traj != nullptr) {
                real* tr = static_cast<real*>(p.traj) + (k * p.traj_ncomp) * N + f;
                const int nc = p.traj_ncomp;
                S2::members([&](auto Bm, int m) {
                  using BM = decltype(Bm);
                  if (BM::gidx(0) < nc) __stcs(tr + static_cast<long long>(BM::gidx(0)) * N, x0[m]);
                  if (BM::gidx(1) < nc) __stcs(tr + static_cast<long long>(BM::gidx(1)) * N, x1[m]);
                });
              }
              if (p.logdet != nullptr) __stcs(static_cast<real*>(p.logdet) + k * N + f, static_cast<real>(rec->logdet));
              if (p.used != nullptr) p.used[k * N + f] = (ty != EV_PREDICT_ONLY) ? 1 : 0;
            }
          }
        }
      }
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&empty[s]);
    }
    if (active) {
      real* xg = static_cast<real*>(p.x) + f * n;
      S2::members([&](auto Bm, int m) {
        using BM = decltype(Bm);
        xg[BM::gidx(0)] = x0[m];
        xg[BM::gidx(1)] = x1[m];
      });
    }
  }

  __syncthreads();  // the producer's final covariance and aggregates
  if (tid >= nfb) return;
  const long long f = base + tid;
  if (T > 0) {
    real* Pg = static_cast<real*>(p.P) + f * (n * n);
    const real a = static_cast<real>(fin[0]), b = static_cast<real>(fin[1]), c = static_cast<real>(fin[2]);
    S2::members([&](auto Bm, int) {
      using BM = decltype(Bm);
      Pg[BM::gidx(0) * n + BM::gidx(0)] = a;
      Pg[BM::gidx(0) * n + BM::gidx(1)] = b;
      Pg[BM::gidx(1) * n + BM::gidx(0)] = b;
      Pg[BM::gidx(1) * n + BM::gidx(1)] = c;
    });
  }
  if (p.agg != nullptr) {
    double* ag = p.agg + f * 4;
    double mx = fin[3], sm = fin[4], cnt = static_cast<double>(T);
    if (p.agg_continue) {
      mx = fmax(mx, ag[0]);
      sm += ag[1];
      cnt += ag[3];
    }
    ag[0] = mx;
    ag[1] = sm;
    ag[2] = (T > 0) ? fin[5] : (p.agg_continue ? ag[2] : 0.0);
    ag[3] = cnt;
  }
  if (p.status != nullptr) {
    const int bad = (T > 0 && fin[6] == 0.0) ? 1 : 0;
    p.status[f] = p.agg_continue ? (p.status[f] | bad) : bad;
  }
}

// ------------------------------------------------------------------------------------------
// Shared-covariance kernel for every model (CA15 / CA8 / BIAS9; CV6 keeps the tuned kernel above).
// Same structure -- copy warp, covariance recursion on single lanes, 28 math warps, payload ring -- with the
// model's generic block engine: one CHAIN warp per block runs the covariance-only filter of the
// axis-symmetric form (filter_predict<false> / filter_update<false, joint>, float64) and stages P+ of
// every step; the GAIN warp, one chunk behind, publishes per step the gain matrix
// K = P+ H^T R^-1 of every block (columns of P+ over the measurement variances, exact for selector
// H and diagonal R) plus determinants / log-dets and the launch's aggregates; a math thread predicts its filter's state with the chain
// tables, builds the IMU pseudo-measurement from the predicted state (kf_workers.py:698-706) and
// applies x += K (z - H x) jointly -- the same update as the stage-by-stage scalar form, in exact
// arithmetic.  Monte-Carlo seeds of one configuration (BASELINE configs[1]: 4096 seeds per rate)
// and batches started from the reference's fixed P0 (configs[4]) take this path in AUTO mode.
// ------------------------------------------------------------------------------------------
template <int MODEL, typename real>
struct SharedRing {  // ring depth per model: stages * chunk * d rows of 896 float64 within ~200 KB
  static constexpr int d = ModelTraits<MODEL, false, real>::d;
  static constexpr int chunk = ((d >= 8) ? 1 : 2) * (sizeof(real) == 4 ? 2 : 1);  // float32: same bytes, twice the steps
  static constexpr int stages = (d >= 8) ? 3 : (d >= 4 ? 3 : 4);
  // Smaller CTAs take longer chunks, not more stages (a chunk costs every warp -- the gain warps' serial chain
  // included -- one mbarrier.try_wait of ~90 cycles).  F = 224 keeps HALF the bytes of the larger CTAs so that
  // two CTAs fit one SM: launches that small come in groups on concurrent streams (one per sampling rate in
  // the scheduled sweep, 19 CTAs each), and a second resident CTA per SM keeps them in one wave.
  // F = 128 (b200kf_run_gains with B200KF_RUN_SMALL_TILES): the gain-applying launches of a rate sweep share the GPU,
  // and a CTA streams only (ring bytes) / (memory latency) ~ 40 GB/s: the rate with the longest schedule sets the
  // sweep's time unless its filters are spread over more CTAs.  About the ring bytes of F = 224, two CTAs per SM.
  template <int F>
  static constexpr int chunk_for() {
    return chunk * (F == 896 ? 1 : (F == 128 ? 3 : 2));
  }
  template <int F>
  static constexpr int ctas_per_sm() {
    return F <= 224 ? 2 : 1;
  }
};

// the blocks (tags) of a filter type, for "one gain warp per block"
template <class Flt>
struct FilterTags;
template <typename real, bool IC, class... Bs>
struct FilterTags<FilterT<real, IC, Bs...>> {
  static constexpr int count = sizeof...(Bs);
  template <int K>
  using tag = std::tuple_element_t<K, std::tuple<Bs...>>;
  template <int K>
  using block = typename BlockObj<tag<K>, real>::type;
  static constexpr int slot_offset(int k) {  // first gain slot of block k
    constexpr int sl[] = {BlockObj<Bs, real>::type::kslots...};
    int o = 0;
    for (int i = 0; i < k; ++i) o += sl[i];
    return o;
  }
  static constexpr int cov_offset(int k) {
    constexpr int np[] = {BlockObj<Bs, real>::type::NP...};
    int o = 0;
    for (int i = 0; i < k; ++i) o += np[i];
    return o;
  }
};

// EXT: the gain records were computed beforehand (b200kf_select_greedy's gain output, or any earlier pass over
// the same schedule and P0) and stream in from global memory with the payload: no gain warps, no serial
// recursion in the launch at all -- the kernel is a pure stream whatever the number of filters.
template <int MODEL, typename real, bool FULL, int F, bool EXT = false>
__global__ void __launch_bounds__(F + 32 * ((EXT ? 0 : FilterTags<typename SymTraits<MODEL, double>::Filter>::count) + 2),
                                  SharedRing<MODEL, real>::template ctas_per_sm<F>())
    kf_run_shared_generic_kernel(const RunParams p) {
  constexpr int kMathThreads = F;
  using MT = ModelTraits<MODEL, false, real>;
  using FltX = typename SymTraits<MODEL, real>::Filter;    // math warps: state only
  using FltP = typename SymTraits<MODEL, double>::Filter;  // chain / gain warps: covariance only, float64
  using Tags = FilterTags<FltP>;
  constexpr int NB = Tags::count;
  static_assert(NB <= kMaxGainBlocks, "too many blocks");
  constexpr int n = MT::n;
  constexpr int D = MT::d;
  constexpr bool CTRL = MT::imu_control;
  constexpr int NG = GainSlots<FltP>::value;
  constexpr int NCOV = GainSlots<FltP>::cov;
  using Rec = GainRecG<NG>;
  constexpr int kSt = SharedRing<MODEL, real>::stages;
  constexpr int kCh = SharedRing<MODEL, real>::template chunk_for<F>();
  // warp roles after the F / 32 math warps: NB chain warps (one block's covariance recursion each), the copy
  // warp, and the aggregate warp, which is also the gain warp (gain records from the staged covariances)
  constexpr int NGW = EXT ? 0 : NB;  // chain warps in this launch
  constexpr int kGainBase = kMathThreads, kCopyTid = kMathThreads + 32 * NGW, kAggTid = kCopyTid + 32;
  if (!mode_selected(p.dense_flag, MODE_SHARED)) return;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  real* stage_buf = reinterpret_cast<real*>(smem_raw);  // [kSt][kCh][D][kMathThreads]
  constexpr size_t kStageReals = static_cast<size_t>(kCh) * D * kMathThreads;
  Rec* gain_buf = reinterpret_cast<Rec*>(stage_buf + kSt * kStageReals);  // [kSt][kCh]
  double* fin = reinterpret_cast<double*>(gain_buf + kSt * kCh);          // cov[NCOV], max, sum, last, ok
  double* pstage = fin + NCOV + 4;                                        // [kSt][kCh][NCOV]: P+ of every step, block after block
  double* irbuf = pstage + (EXT ? 0 : kSt * kCh * NCOV);                  // [2][NB][kMaxN]: 1 / r per measured column (GPS, IMU)
  unsigned long long* full = reinterpret_cast<unsigned long long*>(irbuf + 2 * NB * kMaxN);
  unsigned long long* empty = full + kSt;
  unsigned long long* pready = empty + kSt;  // the chain warps' covariances of a chunk are staged

  const int tid = threadIdx.x;
  const long long N = p.n_filters;
  const long long base = static_cast<long long>(blockIdx.x) * kMathThreads;
  const int nfb = static_cast<int>((N - base) < kMathThreads ? (N - base) : kMathThreads);
  const int T = p.n_steps;
  const int nchunks = (T + kCh - 1) / kCh;
  const unsigned row_bytes = static_cast<unsigned>(nfb * sizeof(real));

  if (tid == 0) {
    for (int s = 0; s < kSt; ++s) {
      // full: the copy warp's expect_tx arrival + the gain warp's (records written); with precomputed records
      // they arrive with the copy.  empty: the math warps (+ the aggregate warp when it reads the records).
      mbar_init(&full[s], EXT ? 1 : 2);
      mbar_init(&empty[s], kMathThreads / 32 + (EXT ? 1 : 0));
      mbar_init(&pready[s], NB > 0 ? NB : 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if constexpr (EXT) {  // covariance records (b200kf_select_greedy's warp walk): the gains are formed here, K = P+ H^T R^-1
      static_for<0, NB>([&](auto Kc) {
        constexpr int K = Kc;
        using Blk = typename Tags::template block<K>;
        Blk::template inv_noise<true>(p.mc, irbuf + K * kMaxN);
        if constexpr (!CTRL) Blk::template inv_noise<false>(p.mc, irbuf + (NB + K) * kMaxN);
      });
    }
  }
  __syncthreads();

  if (tid >= kAggTid) {
    // ---------------- aggregate warp ---------------------------------------------------------------------
    if (tid == kAggTid) {
      LogDetAcc acc;
      bool ok = true;
      if constexpr (EXT) {
        // precomputed records: log-det aggregates of the whole filter from the block determinants
        for (int chunk = 0; chunk < nchunks; ++chunk) {
          const int s = chunk % kSt;
          mbar_wait(&full[s], static_cast<unsigned>(chunk / kSt) & 1u);
          const int k0 = chunk * kCh;
          const int steps = (T - k0) < kCh ? (T - k0) : kCh;
          for (int cc = 0; cc < steps; ++cc) {
            const Rec* rec = gain_buf + s * kCh + cc;
            double det = 1.0;
            if (rec->pad == kCovRecord) {  // covariance record: g holds P+ block after block, det P from it
              static_for<0, NB>([&](auto Kc) {
                constexpr int K = Kc;
                using Blk = typename Tags::template block<K>;
                Blk blk;
#pragma unroll
                for (int e = 0; e < Blk::NP; ++e) blk.P[e] = rec->g[Tags::cov_offset(K) + e];
                const double dk = blk.det_total(ok);
                ok = ok && (dk < INFINITY);
                det *= dk;
              });
            } else {
#pragma unroll
              for (int k = 0; k < NB; ++k) {
                const double dk = rec->det[k];
                ok = ok && (dk > 0.0) && (dk < INFINITY);
                det *= dk;
              }
            }
            acc.add(det);
          }
          mbar_arrive(&empty[s]);
        }
      } else {
        // Gain warp: everything that is NOT on the covariance recursion's serial chain.  The chain warps stage
        // P+ of every step; this lane turns it into the step's gain record (K = P+ H^T R^-1 per block, block
        // determinants, their logs for the per-step output) and the launch's log-det aggregates, one chunk
        // behind the chain.  (With chain and gains on one lane a CA15 step was ~250 dependent instructions,
        // 0.45-0.5 us, and bound every launch below ~10^5 filters and the float32 launches.)
        static_for<0, NB>([&](auto Kc) {
          constexpr int K = Kc;
          using Blk = typename Tags::template block<K>;
          Blk::template inv_noise<true>(p.mc, irbuf + K * kMaxN);
          if constexpr (!CTRL) Blk::template inv_noise<false>(p.mc, irbuf + (NB + K) * kMaxN);
        });
        double dt_nx = (T > 0) ? __ldg(p.dts) : 0.0;
        int ty_nx = (T > 0) ? __ldg(p.types) : 0;
        for (int chunk = 0; chunk < nchunks; ++chunk) {
          const int s = chunk % kSt;
          mbar_wait(&pready[s], static_cast<unsigned>(chunk / kSt) & 1u);
          const int k0 = chunk * kCh;
          const int steps = (T - k0) < kCh ? (T - k0) : kCh;
          for (int cc = 0; cc < steps; ++cc) {
            const double dt = dt_nx;
            const int ty = ty_nx;
            if (k0 + cc + 1 < T) {
              dt_nx = __ldg(p.dts + k0 + cc + 1);
              ty_nx = __ldg(p.types + k0 + cc + 1);
            }
            Rec* rec = gain_buf + s * kCh + cc;
            const double* ps = pstage + (s * kCh + cc) * NCOV;
            double det = 1.0;
            static_for<0, NB>([&](auto Kc) {
              constexpr int K = Kc;
              using Blk = typename Tags::template block<K>;
              constexpr int so = Tags::slot_offset(K);
              constexpr int co = Tags::cov_offset(K);
              Blk blk;
#pragma unroll
              for (int e = 0; e < Blk::NP; ++e) blk.P[e] = ps[co + e];
              if (ty == EV_GPS) {
                blk.template extract_gain<true>(irbuf + K * kMaxN, rec->g + so);
              } else if (ty == EV_IMU) {
                if constexpr (!CTRL) blk.template extract_gain<false>(irbuf + (NB + K) * kMaxN, rec->g + so);
              }
              const double dk = blk.det_total(ok);
              ok = ok && (dk < INFINITY);
              rec->det[K] = dk;
              if constexpr (FULL) {
                if (p.logdet != nullptr) rec->ldet[K] = fast_log_pos(dk);
              }
              det *= dk;
            });
            rec->dt = dt;
            rec->type = ty;
            rec->pad = 0;
            acc.add(det);
          }
          mbar_arrive(&full[s]);  // release: the records are visible to whoever sees the phase flip
        }
      }
      fin[NCOV + 0] = (T > 0) ? log(acc.maxdet) : -INFINITY;
      fin[NCOV + 1] = acc.sum_log();
      fin[NCOV + 2] = (T > 0) ? log(acc.last) : 0.0;
      fin[NCOV + 3] = ok ? 1.0 : 0.0;
    }
  } else if (tid >= kCopyTid) {
    // ---------------- copy warp ----------------------------------------------------------------
    if (tid == kCopyTid) {
      if constexpr (EXT)
        ring_copy_loop<real, D, F, kSt, kCh, Rec>(stage_buf, full, empty, static_cast<const real*>(p.payload) + base, N,
                                                  T, row_bytes, gain_buf, static_cast<const Rec*>(p.gain_recs), p.step_rows);
      else
        ring_copy_loop<real, D, F, kSt, kCh>(stage_buf, full, empty, static_cast<const real*>(p.payload) + base, N, T,
                                             row_bytes, static_cast<SchedRec*>(nullptr), static_cast<const SchedRec*>(nullptr), p.step_rows);
    }
  } else if (!EXT && tid >= kGainBase) {
    // ---------------- chain warps: block k's covariance recursion, and nothing else, on one lane of warp k ----
    const int gw = (tid - kGainBase) >> 5;
    if (((tid - kGainBase) & 31) == 0) {
      static_for<0, NB>([&](auto Kc) {
        constexpr int K = Kc;
        if (gw != K) return;
        using Tag = typename Tags::template tag<K>;
        using Blk = typename Tags::template block<K>;
        FilterT<double, CTRL, Tag> flt;  // this block alone: blocks never couple
        const real* Pg = static_cast<const real*>(p.P) + base * (n * n);  // this CTA's first filter (all are equal)
        flt.head.load_cov_from(Pg, n);
        constexpr int co = Tags::cov_offset(K);
        double dt_nx = (T > 0) ? __ldg(p.dts) : 0.0;  // fetched one step ahead, off the serial chain
        int ty_nx = (T > 0) ? __ldg(p.types) : 0;
        for (int chunk = 0; chunk < nchunks; ++chunk) {
          const int s = chunk % kSt;
          // the stage's staging area is free once its records were consumed (the gain warp read it before that)
          if (chunk >= kSt) mbar_wait(&empty[s], static_cast<unsigned>(chunk / kSt - 1) & 1u);
          const int k0 = chunk * kCh;
          const int steps = (T - k0) < kCh ? (T - k0) : kCh;
          for (int cc = 0; cc < steps; ++cc) {
            const double dt = dt_nx;
            const int ty = ty_nx;
            if (k0 + cc + 1 < T) {
              dt_nx = __ldg(p.dts + k0 + cc + 1);
              ty_nx = __ldg(p.types + k0 + cc + 1);
            }
            filter_predict<false>(flt, dt, p.mc, ty, static_cast<const double*>(nullptr));
            filter_update<false, true>(flt, dt, p.mc, ty, static_cast<const double*>(nullptr));
            double* ps = pstage + (s * kCh + cc) * NCOV + co;
#pragma unroll
            for (int e = 0; e < Blk::NP; ++e) ps[e] = flt.head.P[e];
          }
          mbar_arrive(&pready[s]);  // release: this block's covariances of the chunk are staged
        }
#pragma unroll
        for (int e = 0; e < Blk::NP; ++e) fin[co + e] = flt.head.P[e];
      });
    }
  } else {
    // ---------------- math warps: predict the state, x += K (z - H x) -----------------------------
    const bool active = tid < nfb;
    const long long f = base + tid;
    FltX flt;
    if (active) {
      const real* xg = static_cast<const real*>(p.x) + f * n;
      flt.each([&](auto& b) { b.load_state_from(xg); });
    }
    for (int chunk = 0; chunk < nchunks; ++chunk) {
      const int s = chunk % kSt;
      mbar_wait(&full[s], static_cast<unsigned>(chunk / kSt) & 1u);
      const int k0 = chunk * kCh;
      const int steps = (T - k0) < kCh ? (T - k0) : kCh;
      if (active) {
        const real* sb = stage_buf + s * kStageReals + tid;
#pragma unroll
        for (int cc = 0; cc < kCh; ++cc) {
          if (cc < steps) {
            const Rec* rec = gain_buf + s * kCh + cc;
            const real dt = static_cast<real>(rec->dt);
            const int ty = rec->type;
            real u[D];
#pragma unroll
            for (int j = 0; j < D; ++j) u[j] = sb[(cc * D + j) * kMathThreads];
            flt.each([&](auto& b) { b.predict_state(dt, p.mc); });
            const bool cov_rec = EXT && rec->pad == kCovRecord;  // warp-uniform
            if (ty == EV_IMU) {
              if constexpr (CTRL) {
                flt.each([&](auto& b) { b.control(dt, p.mc, u); });
              } else if (cov_rec) {
                int oc = 0, kb = 0;
                flt.each([&](auto& b) {
                  using BT = std::remove_reference_t<decltype(b)>;
                  b.template correct_cov<false>(rec->g + oc, irbuf + (NB + kb) * kMaxN, dt, u);
                  oc += BT::NP;
                  kb += 1;
                });
              } else {
                int o = 0;
                flt.each([&](auto& b) {
                  using BT = std::remove_reference_t<decltype(b)>;
                  b.template correct<false>(rec->g + o, dt, u);
                  o += BT::kslots;
                });
              }
            } else if (ty == EV_GPS) {
              if (cov_rec) {
                int oc = 0, kb = 0;
                flt.each([&](auto& b) {
                  using BT = std::remove_reference_t<decltype(b)>;
                  b.template correct_cov<true>(rec->g + oc, irbuf + kb * kMaxN, dt, u);
                  oc += BT::NP;
                  kb += 1;
                });
              } else {
                int o = 0;
                flt.each([&](auto& b) {
                  using BT = std::remove_reference_t<decltype(b)>;
                  b.template correct<true>(rec->g + o, dt, u);
                  o += BT::kslots;
                });
              }
            }
            if constexpr (FULL) {
              const long long k = k0 + cc;
              if (p.traj != nullptr) {
                real* tr = static_cast<real*>(p.traj) + (k * p.traj_ncomp) * N + f;
                const int nc = p.traj_ncomp;
                flt.each([&](const auto& b) {
                  b.for_each_state([&](auto G, real v) {
                    constexpr int g = G;
                    if (g < nc) __stcs(tr + static_cast<long long>(g) * N, v);
                  });
                });
              }
              if (p.logdet != nullptr) {
                double ld = 0.0;  // log det P = sum of the blocks' log-dets
                if constexpr (EXT) {  // precomputed records carry determinants only (their producer takes no log)
                  double dp = 1.0;
                  if (cov_rec) {  // ... or the covariance itself
                    bool okd = true;
                    static_for<0, NB>([&](auto Kc) {
                      constexpr int K = Kc;
                      using Blk = typename Tags::template block<K>;
                      Blk blk;
#pragma unroll
                      for (int e = 0; e < Blk::NP; ++e) blk.P[e] = rec->g[Tags::cov_offset(K) + e];
                      dp *= blk.det_total(okd);
                    });
                  } else {
#pragma unroll
                    for (int k2 = 0; k2 < NB; ++k2) dp *= rec->det[k2];
                  }
                  ld = fast_log_pos(dp);
                } else {
#pragma unroll
                  for (int k2 = 0; k2 < NB; ++k2) ld += rec->ldet[k2];
                }
                __stcs(static_cast<real*>(p.logdet) + k * N + f, static_cast<real>(ld));
              }
              if (p.used != nullptr) p.used[k * N + f] = (ty != EV_PREDICT_ONLY) ? 1 : 0;
            }
          }
        }
      }
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&empty[s]);
    }
    if (active) {
      real* xg = static_cast<real*>(p.x) + f * n;
      flt.each([&](const auto& b) { b.store_state_to(xg); });
    }
  }

  __syncthreads();  // the chain warps' final covariance and the aggregate warp's results
  if (tid >= nfb) return;
  const long long f = base + tid;
  if (T > 0 && !EXT) {  // with external gains the covariance never enters the launch (its producer returns P)
    real* Pg = static_cast<real*>(p.P) + f * (n * n);
    FltP shape;  // only its static tables are used
    int o = 0;
    shape.each([&](const auto& b) {
      using BT = std::remove_cv_t<std::remove_reference_t<decltype(b)>>;
      b.store_cov_to(fin + o, Pg, n);
      o += BT::NP;
    });
  }
  if (p.agg != nullptr) {
    double* ag = p.agg + f * 4;
    double mx = fin[NCOV + 0], sm = fin[NCOV + 1], cnt = static_cast<double>(T);
    if (p.agg_continue) {
      mx = fmax(mx, ag[0]);
      sm += ag[1];
      cnt += ag[3];
    }
    ag[0] = mx;
    ag[1] = sm;
    ag[2] = (T > 0) ? fin[NCOV + 2] : (p.agg_continue ? ag[2] : 0.0);
    ag[3] = cnt;
  }
  if (p.status != nullptr) {
    const int bad = (T > 0 && fin[NCOV + 3] == 0.0) ? 1 : 0;
    p.status[f] = p.agg_continue ? (p.status[f] | bad) : bad;
  }
}

template <int MODEL, typename real, int F>
constexpr size_t shared_generic_smem_bytes() {
  using MT = ModelTraits<MODEL, false, real>;
  using FltP = typename SymTraits<MODEL, double>::Filter;
  constexpr int kSt = SharedRing<MODEL, real>::stages, kCh = SharedRing<MODEL, real>::template chunk_for<F>();
  return static_cast<size_t>(kSt) * kCh * MT::d * F * sizeof(real) +
         static_cast<size_t>(kSt) * kCh * sizeof(GainRecG<GainSlots<FltP>::value>) +
         (GainSlots<FltP>::cov + 4) * sizeof(double) +
         (static_cast<size_t>(kSt) * kCh * GainSlots<FltP>::cov + 2 * kMaxGainBlocks * kMaxN) * sizeof(double) +
         3 * kSt * sizeof(unsigned long long);
}

template <int MODEL, typename real, int F>
constexpr size_t shared_smem_bytes() {
  using MT = ModelTraits<MODEL, false, real>;
  constexpr int st = base_stages<real>(), ch = shared_chunk<real, F>();
  // (the gain records are padded to whole doubles so that the aggregates behind them stay 8-byte aligned)
  return static_cast<size_t>(st) * ch * MT::d * F * sizeof(real) +
         ((static_cast<size_t>(st) * ch * sizeof(GainRecT<real>) + 7) / 8) * 8 + 8 * sizeof(double) +
         2 * st * sizeof(unsigned long long);
}

template <int MODEL, typename real, int F>
constexpr size_t tma_smem_bytes() {
  using MT = ModelTraits<MODEL, false, real>;
  constexpr int st = ring_stages<real, F, MT::d>(), ch = ring_chunk<real, F, MT::d>();
  return static_cast<size_t>(st) * ch * MT::d * F * sizeof(real) +
         static_cast<size_t>(st) * ch * sizeof(CARecT<real>) + (2 * st + 2 * bulk_out_stages<MT::d>()) * sizeof(unsigned long long);
}
template <int MODEL, typename real, int F>
static size_t tma_bulk_smem_bytes(int out_rows) {
  using MT = ModelTraits<MODEL, false, real>;
  constexpr int st = bulk_in_stages<MT::d>(), kOut = bulk_out_stages<MT::d>(), ch = bulk_chunk<real, F, MT::d>();
  return static_cast<size_t>(st) * ch * MT::d * F * sizeof(real) +
         static_cast<size_t>(kOut) * ch * out_rows * F * sizeof(real) +
         static_cast<size_t>(st) * ch * sizeof(CARecT<real>) + (2 * st + 2 * kOut) * sizeof(unsigned long long);
}

// filters per CTA of the ring kernels: the largest of {896, 448, 224} that still gives every SM a CTA
static int pick_ring_f(long long n_filters, int sm_count, int cap = 896) {
  if (cap >= 896 && n_filters > static_cast<long long>(sm_count) * 448) return 896;
  if (cap <= 224) return 224;
  if (n_filters > static_cast<long long>(sm_count) * 224) return 448;
  return 224;
}

// AUTO/forced launch of the TMA kernels.  Returns true when they were enqueued (then the caller
// skips the LDG per-chain kernel); false -> caller uses the LDG kernel.  `flags_dev` != null: AUTO,
// both the per-chain and (when the model allows) the axis-symmetric variants are enqueued and
// pick themselves from the device flags.
template <int MODEL, typename real>
static bool tma_possible(const b200kf_model* m, const RunParams& p, bool full) {
  if constexpr (!ModelTraits<MODEL, false, real>::Filter::fast2) {
    return false;
  } else {
    if (p.ptraj != nullptr || m->use_tma == 0) return false;  // per-step P output: LDG kernel only
    (void)full;
    // bulk copies need 16-byte aligned rows: even filter count for f64, multiple of 4 for f32
    const long long align_elems = 16 / static_cast<long long>(sizeof(real));
    return p.n_steps >= 1 && (p.n_filters % align_elems) == 0 &&
           (reinterpret_cast<uintptr_t>(p.payload) & 15u) == 0;
  }
}

// BIAS9-like models (control-input chains): the symmetric ring kernel, with or without per-step outputs;
// the adaptive gate and per-chain covariances stay with the LDG kernels
template <int MODEL, typename real>
static bool tma_sym3_possible(const b200kf_model* m, const RunParams& p, bool full) {
  using MT = ModelTraits<MODEL, false, real>;
  if constexpr (MT::Filter::fast2 || !(MT::Sym3::eligible || MT::SymCA::eligible)) {
    return false;
  } else if constexpr (MT::SymCA::eligible) {
    // the constant-acceleration models (the reference's own CA15 / CA8): every mode incl. the adaptive gate runs on
    // the ring (the gate only skips update(), the state predict is in predict()); per-step P stays with the LDG kernel
    (void)full;
    if (p.ptraj != nullptr || m->use_tma == 0 || m->use_ca_ring == 0) return false;
    const long long align_elems = 16 / static_cast<long long>(sizeof(real));
    return p.n_steps >= 1 && (p.n_filters % align_elems) == 0 && (reinterpret_cast<uintptr_t>(p.payload) & 15u) == 0;
  } else {
    // per-step outputs run on the ring too; the adaptive gate does not (this filter applies the control input
    // inside update(), which a gated step skips) and stays with the LDG kernel
    (void)full;
    if (p.gate_enabled || p.ptraj != nullptr || m->use_tma == 0) return false;
    const long long align_elems = 16 / static_cast<long long>(sizeof(real));
    return p.n_steps >= 1 && (p.n_filters % align_elems) == 0 &&
           (reinterpret_cast<uintptr_t>(p.payload) & 15u) == 0;
  }
}

template <int MODEL, typename real>
static bool sym_possible(const b200kf_model* m) {
  return m->use_sym != 0 && SymTraits<MODEL, real>::Filter::constants_equal(m->mc);
}
// the TMA kernel has its own closed-form symmetric filter (Sym2Filter) for all-(value, rate) models
template <int MODEL, typename real>
constexpr bool tma_sym_eligible() {
  using MT = ModelTraits<MODEL, false, real>;
  if constexpr (!MT::Filter::fast2) return false;
  else return MT::Sym2::eligible;
}

template <int MODEL, typename real, bool SYM, bool FULL, int F>
static cudaError_t launch_tma_f(const RunParams& p, const void* sched, cudaStream_t st) {
  constexpr size_t smem = tma_smem_bytes<MODEL, real, F>();
  const unsigned blocks = static_cast<unsigned>((p.n_filters + F - 1) / F);
  cudaError_t err = cudaFuncSetAttribute(kf_run_tma_kernel<MODEL, real, SYM, FULL, F>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (err != cudaSuccess) return err;
  kf_run_tma_kernel<MODEL, real, SYM, FULL, F><<<blocks, F + 32 * svc_warps<MODEL, real>(), smem, st>>>(p, sched);
  return cudaGetLastError();
}
constexpr size_t kMaxDynSmem = 227 * 1024;  // opt-in dynamic shared memory per CTA on sm_100
template <int MODEL, typename real, bool SYM, int F>
static cudaError_t launch_tma_bulk_f(const RunParams& p, const void* sched, int out_rows, cudaStream_t st) {
  const size_t smem = tma_bulk_smem_bytes<MODEL, real, F>(out_rows);
  const unsigned blocks = static_cast<unsigned>((p.n_filters + F - 1) / F);
  const size_t smem_cap = tma_bulk_smem_bytes<MODEL, real, F>(kBulkMaxRows);
  cudaError_t err = cudaFuncSetAttribute(kf_run_tma_kernel<MODEL, real, SYM, true, F, true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem_cap < kMaxDynSmem ? smem_cap : kMaxDynSmem));
  if (err != cudaSuccess) return err;
  kf_run_tma_kernel<MODEL, real, SYM, true, F, true><<<blocks, F + 32 * svc_warps<MODEL, real>(), smem, st>>>(p, sched);
  return cudaGetLastError();
}
// per-step outputs through the output ring: trajectory components and / or log det (at most kBulkMaxRows rows per
// step), no `used` output (a byte per filter-step: per-thread stores)
static bool bulk_out_ok(const b200kf_model* m, const RunParams& p, int* rows) {
  const int nc = (p.traj != nullptr) ? p.traj_ncomp : 0;
  *rows = nc + (p.logdet != nullptr ? 1 : 0);
  const bool aligned = ((reinterpret_cast<uintptr_t>(p.traj) | reinterpret_cast<uintptr_t>(p.logdet)) & 15u) == 0;
  return m->use_bulk_out != 0 && p.used == nullptr && *rows >= 1 && *rows <= kBulkMaxRows && aligned;
}
template <int MODEL, typename real, bool SYM, bool FULL>
static cudaError_t launch_tma_one(const b200kf_model* m, const RunParams& p, const void* sched, int f,
                                  cudaStream_t st) {
  using MTo = ModelTraits<MODEL, false, real>;
  // (the constant-acceleration filter is never launched with 896 filters per CTA: not instantiated)
  constexpr bool k896 = MTo::Filter::fast2 || MTo::Sym3::eligible;
  if constexpr (FULL) {
    int rows = 0;
    if (bulk_out_ok(m, p, &rows)) {
      if constexpr (k896) {
        if (f == 896 && tma_bulk_smem_bytes<MODEL, real, 896>(rows) <= kMaxDynSmem)
          return launch_tma_bulk_f<MODEL, real, SYM, 896>(p, sched, rows, st);
      }
      if (f >= 448 && tma_bulk_smem_bytes<MODEL, real, 448>(rows) <= kMaxDynSmem)
        return launch_tma_bulk_f<MODEL, real, SYM, 448>(p, sched, rows, st);
      if (f < 448 && tma_bulk_smem_bytes<MODEL, real, 224>(rows) <= kMaxDynSmem)
        return launch_tma_bulk_f<MODEL, real, SYM, 224>(p, sched, rows, st);
    }
  }
  if constexpr (k896) {
    if (f == 896) return launch_tma_f<MODEL, real, SYM, FULL, 896>(p, sched, st);
  }
  if (f >= 448) return launch_tma_f<MODEL, real, SYM, FULL, 448>(p, sched, st);
  return launch_tma_f<MODEL, real, SYM, FULL, 224>(p, sched, st);
}

// IMU variances / process-noise rates of the angular chains' lead member
template <class A0, class... As>
static void ca_ang_const(ChainList<A0, As...>, const ModelConst& mc, CAConst& c) {
  c.b0 = mc.r_imu[A0::gidx(0)], c.b1 = mc.r_imu[A0::gidx(1)];
  c.qa0 = mc.q[A0::gidx(0)], c.qa1 = mc.q[A0::gidx(1)];
}
static void ca_ang_const(...) {}

template <int MODEL, typename real>
static cudaError_t launch_tma(const b200kf_model* m, const RunParams& p, bool with_sym, bool with_dec,
                              bool full, void* sched_mem, cudaStream_t st) {
  using MTl = ModelTraits<MODEL, false, real>;
  if constexpr (!MTl::Filter::fast2 && !MTl::Sym3::eligible && !MTl::SymCA::eligible) {
    return cudaErrorInvalidValue;
  } else {
    constexpr bool kFast2 = MTl::Filter::fast2;
    constexpr bool kCA = !kFast2 && !MTl::Sym3::eligible && MTl::SymCA::eligible;
    void* sched = sched_mem;  // [n_steps] step records of the filter's type, scratch of this call
    if constexpr (kCA) {
      using L0 = typename MTl::SymCA::Lead;
      using A0 = typename CALists<MODEL>::Ang;
      CAConst c{};
      c.r0 = m->mc.r_imu[L0::gidx(0)], c.r1 = m->mc.r_imu[L0::gidx(1)], c.r2 = m->mc.r_imu[L0::gidx(2)];
      c.q0 = m->mc.q[L0::gidx(0)], c.q1 = m->mc.q[L0::gidx(1)], c.q2 = m->mc.q[L0::gidx(2)];
      ca_ang_const(A0{}, m->mc, c);
      c.rg = m->mc.r_gps[L0::gps(0)];
      c.c2 = m->mc.c2;
      pack_ca_schedule_kernel<real><<<(p.n_steps + 255) / 256, 256, 0, st>>>(p.types, p.dts, p.n_steps, c,
                                                                             static_cast<CARecT<real>*>(sched));
    } else {
      LeanConst lc{1.0, 1.0, 1.0, 1.0, 1.0};  // rate constants of the symmetric group (only its lean step reads them)
      if constexpr (kFast2 && ModelTraits<MODEL, false, real>::Sym2::eligible) {
        using G0 = typename ModelTraits<MODEL, false, real>::Sym2::Lead;
        lc.q0 = m->mc.q[G0::gidx(0)];
        lc.q1 = m->mc.q[G0::gidx(1)];
        lc.r0 = m->mc.r_imu[G0::gidx(0)];
        lc.r1 = m->mc.r_imu[G0::gidx(1)];
        if constexpr (G0::gps(0) >= 0) lc.rg = m->mc.r_gps[G0::gps(0)];
      }
      pack_schedule_kernel<real><<<(p.n_steps + 255) / 256, 256, 0, st>>>(p.types, p.dts, p.n_steps, lc,
                                                                          static_cast<LeanRecT<real>*>(sched));
    }
    cudaError_t err = cudaSuccess;
    // the constant-acceleration filter holds 24 state / covariance values per thread: 448 filters per CTA at most
    // (576 threads x ~110 registers), which also keeps a step's rows (d = 9: 32 KB) small enough for a 6-deep ring
    const int fcap = (kCA && m->ring_f_cap > 448) ? 448 : m->ring_f_cap;
    const int f = pick_ring_f(p.n_filters, m->sm_count, fcap);
    if constexpr (kFast2) {
      if (with_dec) {
        err = full ? launch_tma_one<MODEL, real, false, true>(m, p, sched, f, st)
                   : launch_tma_one<MODEL, real, false, false>(m, p, sched, f, st);
        if (err != cudaSuccess) return err;
      }
    }
    if constexpr ((kFast2 && MTl::Sym2::eligible) || (!kFast2 && (MTl::Sym3::eligible || MTl::SymCA::eligible))) {
      if (with_sym) {
        err = full ? launch_tma_one<MODEL, real, true, true>(m, p, sched, f, st)
                   : launch_tma_one<MODEL, real, true, false>(m, p, sched, f, st);
        if (err != cudaSuccess) return err;
      }
    }
    return err;
  }
}

// shared-covariance kernel: eligible models (all chains one axis-symmetric (position, velocity) group),
// aggregate / trajectory / log-det outputs; the adaptive gate and per-step P stay with the other kernels
template <int MODEL, typename real>
constexpr bool shared_lean() {  // the tuned (position, velocity) kernel; every other model takes the generic one
  using MT = ModelTraits<MODEL, false, real>;
  if constexpr (!MT::Filter::fast2) return false;
  else if constexpr (!MT::Sym2::eligible) return false;
  else return MT::Sym2::lean_ok;
}
template <int MODEL, typename real>
constexpr bool shared_eligible() {
  return true;
}

template <int MODEL, typename real, bool FULL, int F>
static cudaError_t launch_shared_f(const RunParams& p, cudaStream_t st) {
  const unsigned blocks = static_cast<unsigned>((p.n_filters + F - 1) / F);
  cudaError_t err;
  if constexpr (!shared_lean<MODEL, real>()) {
    constexpr size_t smem = shared_generic_smem_bytes<MODEL, real, F>();
    err = cudaFuncSetAttribute(kf_run_shared_generic_kernel<MODEL, real, FULL, F>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (err != cudaSuccess) return err;
    if constexpr (F == 224) {  // two CTAs per SM need the whole shared-memory carve-out (a hint, not an error if refused)
      (void)cudaFuncSetAttribute(kf_run_shared_generic_kernel<MODEL, real, FULL, F>,
                                 cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    kf_run_shared_generic_kernel<MODEL, real, FULL, F>
        <<<blocks, F + 32 * (FilterTags<typename SymTraits<MODEL, double>::Filter>::count + 2), smem, st>>>(p);
  } else {
    constexpr size_t smem = shared_smem_bytes<MODEL, real, F>();
    err = cudaFuncSetAttribute(kf_run_shared_kernel<MODEL, real, FULL, F>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem));
    if (err != cudaSuccess) return err;
    if constexpr (F == 224) {
      (void)cudaFuncSetAttribute(kf_run_shared_kernel<MODEL, real, FULL, F>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared);
    }
    kf_run_shared_kernel<MODEL, real, FULL, F><<<blocks, F + 64, smem, st>>>(p);
  }
  return cudaGetLastError();
}
template <int MODEL, typename real>
static cudaError_t launch_shared(const b200kf_model* m, const RunParams& p, bool full, cudaStream_t st) {
  // a step's rows must leave room for a ring deep enough to cover the loaded DRAM latency: at most ~32 KB per
  // step and CTA (CA15 in float64 is 72 B per filter-step: 448 filters per CTA, 6 stages, +45 % over 896 / 3)
  constexpr int kStepBytes896 = ModelTraits<MODEL, false, real>::d * static_cast<int>(sizeof(real)) * 896;
  const int cap = (kStepBytes896 > 36 * 1024 && m->ring_f_cap > 448) ? 448 : m->ring_f_cap;
  const int f = pick_ring_f(p.n_filters, m->sm_count, cap);
  if (full) {
    if (f == 896) return launch_shared_f<MODEL, real, true, 896>(p, st);
    if (f == 448) return launch_shared_f<MODEL, real, true, 448>(p, st);
    return launch_shared_f<MODEL, real, true, 224>(p, st);
  }
  if (f == 896) return launch_shared_f<MODEL, real, false, 896>(p, st);
  if (f == 448) return launch_shared_f<MODEL, real, false, 448>(p, st);
  return launch_shared_f<MODEL, real, false, 224>(p, st);
}

// b200kf_run_gains: filters that only apply precomputed gain records (any model, axis-symmetric form)
template <int MODEL, typename real, bool FULL, int F>
static cudaError_t launch_gains_f(const RunParams& p, cudaStream_t st) {
  constexpr size_t smem = shared_generic_smem_bytes<MODEL, real, F>();
  const unsigned blocks = static_cast<unsigned>((p.n_filters + F - 1) / F);
  cudaError_t err = cudaFuncSetAttribute(kf_run_shared_generic_kernel<MODEL, real, FULL, F, true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (err != cudaSuccess) return err;
  if constexpr (F <= 224) {  // two CTAs per SM need the whole shared-memory carve-out (a hint, not an error if refused)
    (void)cudaFuncSetAttribute(kf_run_shared_generic_kernel<MODEL, real, FULL, F, true>,
                               cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }
  kf_run_shared_generic_kernel<MODEL, real, FULL, F, true><<<blocks, F + 64, smem, st>>>(p);
  return cudaGetLastError();
}
template <int MODEL, typename real>
static int launch_gains_precision(const b200kf_model* m, const RunParams& p, cudaStream_t st) {
  const bool full = p.traj != nullptr || p.logdet != nullptr || p.used != nullptr;
  constexpr int kStepBytes896 = ModelTraits<MODEL, false, real>::d * static_cast<int>(sizeof(real)) * 896;
  const int cap = (kStepBytes896 > 36 * 1024 && m->ring_f_cap > 448) ? 448 : m->ring_f_cap;
  const int f = pick_ring_f(p.n_filters, m->sm_count, cap);
  cudaError_t e;
  if (p.small_tiles && f == 224)
    e = full ? launch_gains_f<MODEL, real, true, 128>(p, st) : launch_gains_f<MODEL, real, false, 128>(p, st);
  else if (full)
    e = f == 896 ? launch_gains_f<MODEL, real, true, 896>(p, st)
                 : (f == 448 ? launch_gains_f<MODEL, real, true, 448>(p, st) : launch_gains_f<MODEL, real, true, 224>(p, st));
  else
    e = f == 896 ? launch_gains_f<MODEL, real, false, 896>(p, st)
                 : (f == 448 ? launch_gains_f<MODEL, real, false, 448>(p, st) : launch_gains_f<MODEL, real, false, 224>(p, st));
  B200KF_CUDA_TRY(e);
  return B200KF_OK;
}
template <>
int launch_gains_m<B200KF_INST_MODEL>(const b200kf_model* m, const RunParams& p, cudaStream_t st) {
  if (m->desc.precision == B200KF_F32) return launch_gains_precision<B200KF_INST_MODEL, float>(m, p, st);
  return launch_gains_precision<B200KF_INST_MODEL, double>(m, p, st);
}
template <>
size_t gain_record_bytes_m<B200KF_INST_MODEL>() {
  return sizeof(GainRecG<GainSlots<typename SymTraits<B200KF_INST_MODEL, double>::Filter>::value>);
}

template <int MODEL, typename real>
__global__ void chain_compat_kernel(const real* __restrict__ P, long long n_filters, int* flag) {
  using MT = ModelTraits<MODEL, false, real>;
  constexpr int n = MT::n;
  const long long f = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (f >= n_filters) return;
  const real* Pg = P + f * (n * n);
  if (flag[1] == 0 && !SymTraits<MODEL, real>::Filter::blocks_identical(Pg, n)) atomicOr(flag + 1, 1);
  if (flag[2] == 0) {  // shared-covariance mode needs every P bit-identical to filter 0's
    bool same = true;
    for (int e = 0; e < n * n; ++e) same = same && (Pg[e] == P[e]);
    if (!same) atomicOr(flag + 2, 1);
  }
  bool coupled = false;
  static_for<0, n>([&](auto I) {
    constexpr int i = I;
    static_for<0, n>([&](auto J) {
      constexpr int j = J;
      if constexpr (MT::chain_of(i) != MT::chain_of(j)) coupled = coupled || (Pg[i * n + j] != real(0));
    });
  });
  if (coupled) atomicOr(flag, 1);
}

template <int MODEL, int VAR, typename real>
static cudaError_t launch_variant(const b200kf_model* m, const RunParams& p, bool full,
                                  cudaStream_t st) {
  // float64 only: in float32 the 120 entries fit the thread-per-filter kernel's registers and it is the faster
  // one (6.0e9 vs 4.8e9 steps/s, profiles/r02i_dense_ca15.jsonl)
  if constexpr (VAR == VAR_DENSE && AxisTraits<MODEL>::eligible && sizeof(real) == 8) {
    if (m->use_axis_dense != 0) {
      const long long blocks = (p.n_filters + kAxisFiltersPerBlock - 1) / kAxisFiltersPerBlock;
      if (p.gate_enabled)
        kf_run_axis_kernel<MODEL, real, true, true><<<static_cast<unsigned>(blocks), kAxisThreads, 0, st>>>(p);
      else if (full)
        kf_run_axis_kernel<MODEL, real, true, false><<<static_cast<unsigned>(blocks), kAxisThreads, 0, st>>>(p);
      else
        kf_run_axis_kernel<MODEL, real, false, false><<<static_cast<unsigned>(blocks), kAxisThreads, 0, st>>>(p);
      return cudaGetLastError();
    }
  }
  using LC = LaunchCfg<MODEL, VAR, real>;
  const int tpb = LC::fast ? LC::max_threads : (m->tpb <= LC::max_threads ? m->tpb : LC::max_threads);
  const long long blocks = (p.n_filters + tpb - 1) / tpb;
  if (full)
    kf_run_kernel<MODEL, VAR, real, true><<<static_cast<unsigned>(blocks), tpb, 0, st>>>(p);
  else
    kf_run_kernel<MODEL, VAR, real, false><<<static_cast<unsigned>(blocks), tpb, 0, st>>>(p);
  return cudaGetLastError();
}

template <int MODEL, typename real>
static int launch_precision(const b200kf_model* m, RunParams p, unsigned flags, cudaStream_t st) {
  const bool full = p.traj != nullptr || p.ptraj != nullptr || p.logdet != nullptr || p.used != nullptr || p.gate_enabled;
  if (flags & B200KF_RUN_FORCE_DENSE) {
    p.dense_flag = nullptr;
    B200KF_CUDA_TRY((launch_variant<MODEL, VAR_DENSE, real>(m, p, full, st)));
    return B200KF_OK;
  }
  // the lean symmetric step and the shared-covariance kernels divide by the measurement variances (R-normalised
  // covariance, K = P+ H^T R^-1): a model with a zero variance (a perfect sensor) stays on the LDG kernels,
  // whose scalar updates only ever divide by P_ss + r
  bool r_pos = true;
  for (int i = 0; i < m->n; ++i) r_pos = r_pos && (m->mc.r_imu[i] > 0.0 || ModelTraits<MODEL, false, real>::imu_control);
  for (int i = 0; i < m->m_gps; ++i) r_pos = r_pos && m->mc.r_gps[i] > 0.0;
  const bool tma = tma_possible<MODEL, real>(m, p, full);
  const bool sym = sym_possible<MODEL, real>(m);
  // scratch of THIS call (stream-ordered, freed behind the last kernel below): the three AUTO dispatch flags
  // and the packed step records of the ring kernels
  ScratchBuf scratch;
  B200KF_CUDA_TRY(scratch.alloc(m, 256 + static_cast<size_t>(p.n_steps) * sizeof(CARecT<double>), st));
  static_assert(sizeof(CARecT<double>) >= sizeof(SchedRec), "the scratch holds either record type");
  void* sched_mem = static_cast<char*>(scratch.p) + 256;
  const bool tma_sym = tma && tma_sym_eligible<MODEL, real>() && r_pos;
  // the shared-covariance kernels stream the payload with bulk copies: 16-byte aligned rows
  const long long align_elems = 16 / static_cast<long long>(sizeof(real));
  const bool ring_ok = p.n_steps >= 1 && p.ptraj == nullptr && m->use_tma != 0 && (p.n_filters % align_elems) == 0 &&
                       (reinterpret_cast<uintptr_t>(p.payload) & 15u) == 0;
  const bool shared = ring_ok && r_pos && sym && shared_eligible<MODEL, real>() && m->use_shared != 0 && !p.gate_enabled &&
                      (!shared_lean<MODEL, real>() || tma_sym);
  if ((flags & B200KF_RUN_FORCE_SHARED) && shared) {
    p.dense_flag = nullptr;
    B200KF_CUDA_TRY((launch_shared<MODEL, real>(m, p, full, st)));
    return B200KF_OK;
  }
  if (flags & B200KF_RUN_FORCE_SHARED) flags |= B200KF_RUN_FORCE_SYMMETRIC;  // next-best guarantee
  // BIAS9 (control-input chains) and the constant-acceleration models CA15 / CA8: symmetric-form ring kernels; the
  // latter carry an R-normalised covariance and need positive measurement variances
  const bool tma3 = sym && tma_sym3_possible<MODEL, real>(m, p, full) &&
                    (!ModelTraits<MODEL, false, real>::SymCA::eligible || r_pos);
  if ((flags & B200KF_RUN_FORCE_SYMMETRIC) && sym) {
    p.dense_flag = nullptr;
    if (tma_sym || tma3) {
      B200KF_CUDA_TRY((launch_tma<MODEL, real>(m, p, true, false, full, sched_mem, st)));
    } else {
      B200KF_CUDA_TRY((launch_variant<MODEL, VAR_SYMMETRIC, real>(m, p, full, st)));
    }
    return B200KF_OK;
  }
  if (flags & (B200KF_RUN_FORCE_DECOUPLED | B200KF_RUN_FORCE_SYMMETRIC)) {
    p.dense_flag = nullptr;
    if (tma) {
      B200KF_CUDA_TRY((launch_tma<MODEL, real>(m, p, false, true, full, sched_mem, st)));
      return B200KF_OK;
    }
    B200KF_CUDA_TRY((launch_variant<MODEL, VAR_DECOUPLED, real>(m, p, full, st)));
    return B200KF_OK;
  }
  // AUTO: decide on the device, no host round trip.  Every candidate kernel is enqueued; the ones
  // whose mode does not match the flags exit at once.
  // A shared-covariance launch can never be faster than n_steps x (one step of the gain warp's serial recursion,
  // ~0.23 us); the lean per-filter ring kernels run the same recursion in every thread at ~0.17 us per step
  // when an SM holds few warps, so shards below the 896-filter CTA regime (measured cross-over ~1e5 filters,
  // profiles/r01c_small_shards.jsonl) stay per-filter in aggregate mode.
  const bool small_shard = pick_ring_f(p.n_filters, m->sm_count, m->ring_f_cap) < 896;
  const bool shared_auto = shared && !(small_shard && !full && (tma_sym || tma3));
  int* flag = static_cast<int*>(scratch.p);
  B200KF_CUDA_TRY(cudaMemsetAsync(flag, 0, sizeof(int), st));
  B200KF_CUDA_TRY(cudaMemsetAsync(flag + 1, sym ? 0 : 1, sizeof(int), st));
  B200KF_CUDA_TRY(cudaMemsetAsync(flag + 2, shared_auto ? 0 : 1, sizeof(int), st));
  const int ctpb = 128;
  chain_compat_kernel<MODEL, real>
      <<<static_cast<unsigned>((p.n_filters + ctpb - 1) / ctpb), ctpb, 0, st>>>(
          static_cast<const real*>(p.P), p.n_filters, flag);
  B200KF_CUDA_TRY(cudaGetLastError());
  p.dense_flag = flag;
  if (shared_auto) B200KF_CUDA_TRY((launch_shared<MODEL, real>(m, p, full, st)));
  if (tma) {
    B200KF_CUDA_TRY((launch_tma<MODEL, real>(m, p, sym && tma_sym, true, full, sched_mem, st)));
  } else {
    if (tma3) B200KF_CUDA_TRY((launch_tma<MODEL, real>(m, p, true, false, full, sched_mem, st)));
    B200KF_CUDA_TRY((launch_variant<MODEL, VAR_DECOUPLED, real>(m, p, full, st)));
  }
  if (sym && !tma_sym && !tma3) B200KF_CUDA_TRY((launch_variant<MODEL, VAR_SYMMETRIC, real>(m, p, full, st)));
  B200KF_CUDA_TRY((launch_variant<MODEL, VAR_DENSE, real>(m, p, full, st)));
  return B200KF_OK;
}

template <>
int launch_run_m<B200KF_INST_MODEL>(const b200kf_model* m, const RunParams& p, unsigned flags,
                                    cudaStream_t st) {
  if (m->desc.precision == B200KF_F32) return launch_precision<B200KF_INST_MODEL, float>(m, p, flags, st);
  return launch_precision<B200KF_INST_MODEL, double>(m, p, flags, st);
}

}  // namespace b200kf
