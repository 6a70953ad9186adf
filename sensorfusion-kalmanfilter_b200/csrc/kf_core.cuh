// kf_core.cuh -- register-resident Kalman-filter block engine and the model tables.
//
// A model (CV6 / CA8 / BIAS9 / CA15) is a list of independent *chains* (blocks): with the
// reference's diagonal Q, R and row-selector H (kf_workers.py:519-614, hw5_2.py:233-304) the
// filter never couples axes, so a block-compatible covariance can be filtered
// chain by chain ("decoupled"), and an arbitrary dense covariance as ONE block holding every
// state ("dense", the Union<> of the chains).  Both run the same code below.
//
// Per block the state x[n] and the upper triangle of the symmetric covariance P live in
// registers: every loop is unrolled at compile time (static_for) so all indices are constants.
//
//   predict   x = F x (+ B u),  P = F P F^T + Q     kf_workers.py:546-549, 688-691
//             F = I + dt*N1 + c2*dt^2*N2 is applied as elementary row/column operations
//   update    K = P H^T (H P H^T + R)^-1, x += K(z - Hx), P = (I - KH) P   :616-621, :709-711
//             H selects states and R is diagonal, so the joint update equals a sequence of
//             scalar updates (exact in exact arithmetic); each costs one reciprocal
//   metric    det P (log taken by the caller)        :716
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

namespace b200kf {

constexpr int kMaxN = 16;
constexpr int kMaxMGps = 4;

// Rate constants of a model instance (host getters -> desc -> here); passed by value as a kernel
// parameter so every access is a constant-offset ld.param.
struct ModelConst {
  double q[kMaxN];         // Q(dt) = dt * diag(q)              kf_workers.py:519-544
  double r_gps[kMaxMGps];  // R_GPS = diag(r_gps)               kf_workers.py:581-585
  double r_imu[kMaxN];     // R_IMU = diag(r_imu)               kf_workers.py:587-614
  double c2;               // coefficient of dt^2 in F (0.5)    kf_workers.py:501-503
};

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

// ---- reciprocals: every 1/x in the filters has x = P_ss + r > 0 and normal --------------------
__device__ __forceinline__ double fast_rcp(double a) {  // a normal and > 0; ~1 ulp
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
  double e = fma(-a, x, 1.0);
  x = fma(x, e, x);
  e = fma(-a, x, 1.0);
  x = fma(x, e, x);
  return x;
}
__device__ __forceinline__ float fast_rcp(float a) {  // a normal and > 0; one Newton step on the hardware seed
  float x;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(x) : "f"(a));
  return fmaf(x, fmaf(-a, x, 1.0f), x);
}
// one cubically convergent step x (1 + e + e^2), e = 1 - a x: error e0^3 from the hardware seed
// (e0 <= 2^-20 -> below 2^-59) with 3 FMAs instead of the 4 of two Newton steps
__device__ __forceinline__ double fast_rcp3(double a) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
  const double e = fma(-a, x, 1.0);
  return fma(x, fma(e, e, e), x);
}
__device__ __forceinline__ float fast_rcp3(float a) { return fast_rcp(a); }

// log of a POSITIVE NORMAL FINITE double (every determinant of a valid filter): exponent + the atanh series of the
// mantissa folded into [sqrt(1/2), sqrt(2)) -- log m = 2 s (1 + z/3 + z^2/5 + ...), s = (m - 1)/(m + 1), z = s^2 <= 0.0295,
// ten terms (truncation < 3e-17) -- ~30 instructions against the ~170 of log() with its special cases; anything
// else (zero, negative, subnormal, inf, NaN: an invalid filter, flagged in its status) takes log().  The per-step
// log det output of the trajectory mode is the only per-step transcendental of the filter loop.
__device__ __forceinline__ double fast_log_pos(double x) {
  int hi = __double2hiint(x);
  if (static_cast<unsigned>(hi - 0x00100000) >= 0x7fe00000u) return log(x);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;
  if (hi >= 0x3ff6a09f) {  // m > ~sqrt(2): halve the mantissa (any threshold near sqrt(2) keeps |s| <= 0.172)
    hi -= 0x00100000;
    e += 1;
  }
  const double m = __hiloint2double(hi, __double2loint(x));
  const double f = m - 1.0;
  const double s = f * fast_rcp(2.0 + f);
  const double z = s * s;
  double p = 1.0 / 21.0;
  p = fma(p, z, 1.0 / 19.0);
  p = fma(p, z, 1.0 / 17.0);
  p = fma(p, z, 1.0 / 15.0);
  p = fma(p, z, 1.0 / 13.0);
  p = fma(p, z, 1.0 / 11.0);
  p = fma(p, z, 1.0 / 9.0);
  p = fma(p, z, 1.0 / 7.0);
  p = fma(p, z, 1.0 / 5.0);
  p = fma(p, z, 1.0 / 3.0);
  const double s2 = s + s;
  const double lm = fma(s2, p * z, s2);
  return fma(static_cast<double>(e), 0.6931471805599453094, lm);
}

// Per-step record of the ring kernels' packed schedule, in the arithmetic type of the launch (the
// constants are evaluated in float64 and rounded once, by pack_schedule_kernel): (dt, type) for every
// kernel, plus the warp-uniform constants of Sym2Filter::step_lean.  Sizes are multiples of 16 bytes
// (bulk copies).
template <typename real>
struct LeanRecT;
template <>
struct alignas(16) LeanRecT<double> {  // 64 bytes
  double dt, dt2;
  double kap;   // dt sqrt(r1 / r0)
  double lam;   // IMU: dt sqrt(r0 / r1)       GPS: sqrt(r1 / r0)
  double kdt;   // IMU: kap dt                 GPS: r_gps / r0
  double phi0;  // q0 dt / r0
  double phi1;  // q1 dt / r1
  int type;
  int pad;
};
template <>
struct alignas(16) LeanRecT<float> {  // 32 bytes, same fields
  float dt, dt2, kap, lam, kdt, phi0, phi1;
  int type;
};

// inv[i] = 1 / d[i] from ONE reciprocal; returns 1 / prod(d)
template <int K>
__device__ __forceinline__ double rcp_batch(const double (&d)[K], double (&inv)[K]) {
  double pre[K];
  pre[0] = d[0];
#pragma unroll
  for (int i = 1; i < K; ++i) pre[i] = pre[i - 1] * d[i];
  const double rall = fast_rcp(pre[K - 1]);
  double r = rall;
#pragma unroll
  for (int i = K - 1; i > 0; --i) {
    inv[i] = r * pre[i - 1];
    r = r * d[i];
  }
  inv[0] = r;
  return rall;
}
template <int K>
__device__ __forceinline__ double rcp_batch(const float (&d)[K], float (&inv)[K]) {
  double rall = 1.0;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    inv[i] = 1.0f / d[i];
    rall *= double(inv[i]);
  }
  return rall;
}

// IMU pseudo-measurement kinds (kf_workers.py:698-704)
enum : int { ZK_POS = 0, ZK_VEL = 1, ZK_DIRECT = 2, ZK_NONE = 3 };

// --------------------------------------------------------------------------------------
// Chain tables.  Local state order inside a chain is (value, rate, rate-of-rate).
//   gidx(i)   global state index           l1/l2(i)  local index multiplied by dt / c2*dt^2
//   s1/s2(i)  sign of that coefficient      gps(i)    GPS payload column measuring state i
//   zk(i)     IMU pseudo-measurement kind   pz(i)     payload column it reads
//   cu(i)     control kind for imu_control models: 1 -> +c2*dt^2*u, 2 -> +dt*u, 0 none
// --------------------------------------------------------------------------------------
template <int AXIS>
struct CV6Chain {  // [p, v] of one axis; global layout [p(3), v(3)]
  static constexpr int n = 2;
  static constexpr int gidx(int i) { return i == 0 ? AXIS : 3 + AXIS; }
  static constexpr int l1(int i) { return i == 0 ? 1 : -1; }
  static constexpr int l2(int) { return -1; }
  static constexpr int s1(int) { return 1; }
  static constexpr int s2(int) { return 1; }
  static constexpr int gps(int i) { return i == 0 ? AXIS : -1; }
  static constexpr int zk(int i) { return i == 0 ? ZK_POS : ZK_VEL; }
  static constexpr int pz(int) { return AXIS; }
  static constexpr int cu(int) { return 0; }
};

template <int AXIS>
struct CA8LinChain {  // [x|y, v, a]; global [x, y, th, vx, vy, thdot, ax, ay]  hw5_2.py:219-231
  static constexpr int n = 3;
  static constexpr int gidx(int i) { return i == 0 ? AXIS : (i == 1 ? 3 + AXIS : 6 + AXIS); }
  static constexpr int l1(int i) { return i == 0 ? 1 : (i == 1 ? 2 : -1); }
  static constexpr int l2(int i) { return i == 0 ? 2 : -1; }
  static constexpr int s1(int) { return 1; }
  static constexpr int s2(int) { return 1; }
  static constexpr int gps(int i) { return i == 0 ? AXIS : -1; }
  static constexpr int zk(int i) { return i == 0 ? ZK_POS : (i == 1 ? ZK_VEL : ZK_DIRECT); }
  static constexpr int pz(int) { return 2 + AXIS; }  // payload [yaw, wz, ax, ay]
  static constexpr int cu(int) { return 0; }
};

struct CA8YawChain {  // [theta, theta_dot]
  static constexpr int n = 2;
  static constexpr int gidx(int i) { return i == 0 ? 2 : 5; }
  static constexpr int l1(int i) { return i == 0 ? 1 : -1; }
  static constexpr int l2(int) { return -1; }
  static constexpr int s1(int) { return 1; }
  static constexpr int s2(int) { return 1; }
  static constexpr int gps(int) { return -1; }
  static constexpr int zk(int) { return ZK_DIRECT; }
  static constexpr int pz(int i) { return i; }
  static constexpr int cu(int) { return 0; }
};

template <int AXIS>
struct CA15LinChain {  // [p, v, a]; global [p(3), rpy(3), v(3), w(3), a(3)]  kf_workers.py:493-517
  static constexpr int n = 3;
  static constexpr int gidx(int i) { return i == 0 ? AXIS : (i == 1 ? 6 + AXIS : 12 + AXIS); }
  static constexpr int l1(int i) { return i == 0 ? 1 : (i == 1 ? 2 : -1); }
  static constexpr int l2(int i) { return i == 0 ? 2 : -1; }
  static constexpr int s1(int) { return 1; }
  static constexpr int s2(int) { return 1; }
  static constexpr int gps(int i) { return i == 0 ? AXIS : -1; }
  static constexpr int zk(int i) { return i == 0 ? ZK_POS : (i == 1 ? ZK_VEL : ZK_DIRECT); }
  static constexpr int pz(int) { return 6 + AXIS; }  // payload [r,p,y, wx,wy,wz, ax,ay,az]
  static constexpr int cu(int) { return 0; }
};

template <int AXIS>
struct CA15AngChain {  // [angle, rate]
  static constexpr int n = 2;
  static constexpr int gidx(int i) { return i == 0 ? 3 + AXIS : 9 + AXIS; }
  static constexpr int l1(int i) { return i == 0 ? 1 : -1; }
  static constexpr int l2(int) { return -1; }
  static constexpr int s1(int) { return 1; }
  static constexpr int s2(int) { return 1; }
  static constexpr int gps(int) { return -1; }
  static constexpr int zk(int) { return ZK_DIRECT; }
  static constexpr int pz(int i) { return i == 0 ? AXIS : 3 + AXIS; }
  static constexpr int cu(int) { return 0; }
};

template <int AXIS>
struct BIAS9Chain {  // [p, v, b]; p += dt v - c2 dt^2 b, v -= dt b; u = accel (SURVEY.md 8(d) cfg 5)
  static constexpr int n = 3;
  static constexpr int gidx(int i) { return AXIS + 3 * i; }
  static constexpr int l1(int i) { return i == 0 ? 1 : (i == 1 ? 2 : -1); }
  static constexpr int l2(int i) { return i == 0 ? 2 : -1; }
  static constexpr int s1(int i) { return i == 0 ? 1 : -1; }
  static constexpr int s2(int) { return -1; }
  static constexpr int gps(int i) { return i == 0 ? AXIS : -1; }
  static constexpr int zk(int) { return ZK_NONE; }
  static constexpr int pz(int) { return AXIS; }
  static constexpr int cu(int i) { return i == 0 ? 1 : (i == 1 ? 2 : 0); }
};

template <int AXIS>
struct CVBU6Chain {  // [p, v]; the north star's x = F x + B u for six states: p += dt v + c2 dt^2 u, v += dt u (u = accel),
  static constexpr int n = 2;  // IMU events update nothing, GPS measures p (the explicit-Bu variant)
  static constexpr int gidx(int i) { return i == 0 ? AXIS : 3 + AXIS; }
  static constexpr int l1(int i) { return i == 0 ? 1 : -1; }
  static constexpr int l2(int) { return -1; }
  static constexpr int s1(int) { return 1; }
  static constexpr int s2(int) { return 1; }
  static constexpr int gps(int i) { return i == 0 ? AXIS : -1; }
  static constexpr int zk(int) { return ZK_NONE; }
  static constexpr int pz(int) { return AXIS; }
  static constexpr int cu(int i) { return i == 0 ? 1 : 2; }
};

// The dense form of a model: one block holding every state of the listed chains.
template <class... Bs>
struct Union {
  static constexpr int n = (Bs::n + ...);
  template <class B>
  static constexpr int local_of(int g) {
    for (int i = 0; i < B::n; ++i)
      if (B::gidx(i) == g) return i;
    return -1;
  }
  static constexpr int gidx(int i) { return i; }
#define B200KF_UNION_LOOKUP(name, expr, dflt)                                   \
  template <class B>                                                            \
  static constexpr bool name##_one(int g, int& r) {                             \
    const int i = local_of<B>(g);                                               \
    if (i >= 0) { r = (expr); return true; }                                    \
    return false;                                                               \
  }                                                                             \
  static constexpr int name(int g) {                                            \
    int r = (dflt);                                                             \
    (void)(name##_one<Bs>(g, r) || ...);                                        \
    return r;                                                                   \
  }
  B200KF_UNION_LOOKUP(l1, (B::l1(i) >= 0 ? B::gidx(B::l1(i)) : -1), -1)
  B200KF_UNION_LOOKUP(l2, (B::l2(i) >= 0 ? B::gidx(B::l2(i)) : -1), -1)
  B200KF_UNION_LOOKUP(s1, B::s1(i), 1)
  B200KF_UNION_LOOKUP(s2, B::s2(i), 1)
  B200KF_UNION_LOOKUP(gps, B::gps(i), -1)
  B200KF_UNION_LOOKUP(zk, B::zk(i), ZK_NONE)
  B200KF_UNION_LOOKUP(pz, B::pz(i), 0)
  B200KF_UNION_LOOKUP(cu, B::cu(i), 0)
#undef B200KF_UNION_LOOKUP
};

// --------------------------------------------------------------------------------------
// One block in registers
// --------------------------------------------------------------------------------------
template <class B, typename real>
struct BlockKF {
  static constexpr int n = B::n;
  static constexpr int M = 1;  // chains that share this covariance (SymBlock: the members of the group)
  static constexpr int NP = n * (n + 1) / 2;
  using Traits = B;
  real x[n];
  real P[NP];
  static constexpr int gidx(int i) { return B::gidx(i); }

  static constexpr int pidx(int i, int j) {
    return i <= j ? (i * n - i * (i - 1) / 2 + (j - i)) : (j * n - j * (j - 1) / 2 + (i - j));
  }

  // xg: [n_total] of this filter, Pg: [n_total, n_total] row-major of this filter.
  __device__ __forceinline__ void load(const real* __restrict__ xg, const real* __restrict__ Pg,
                                       int n_total) {
    static_for<0, n>([&](auto I) {
      constexpr int i = I;
      x[i] = xg[B::gidx(i)];
      static_for<i, n>([&](auto J) {
        constexpr int j = J;
        const real a = Pg[B::gidx(i) * n_total + B::gidx(j)];
        const real b = Pg[B::gidx(j) * n_total + B::gidx(i)];
        P[pidx(i, j)] = real(0.5) * (a + b);
      });
    });
  }

  __device__ __forceinline__ void store(real* __restrict__ xg, real* __restrict__ Pg,
                                        int n_total) const {
    static_for<0, n>([&](auto I) {
      constexpr int i = I;
      xg[B::gidx(i)] = x[i];
      static_for<i, n>([&](auto J) {
        constexpr int j = J;
        Pg[B::gidx(i) * n_total + B::gidx(j)] = P[pidx(i, j)];
        Pg[B::gidx(j) * n_total + B::gidx(i)] = P[pidx(i, j)];
      });
    });
  }

  // covariance-only initialisation from a dense [n_total, n_total] matrix in param/constant space
  __device__ __forceinline__ void load_cov(const double* __restrict__ Pg, int n_total) {
    static_for<0, n>([&](auto I) {
      constexpr int i = I;
      x[i] = real(0);
      static_for<i, n>([&](auto J) {
        constexpr int j = J;
        P[pidx(i, j)] = real(0.5 * (Pg[B::gidx(i) * n_total + B::gidx(j)] +
                                    Pg[B::gidx(j) * n_total + B::gidx(i)]));
      });
    });
  }

  // x = F x, P = F P F^T + Q.  F = prod of elementary ops "row i += c1*row l1 + c2*row l2";
  // links always point to a higher local index, so ascending i reads not-yet-modified rows.
  template <bool WITH_STATE>
  __device__ __forceinline__ void predict(real dt, const ModelConst& mc) {
    const real a = dt;
    const real b = real(mc.c2) * (dt * dt);
    static_for<0, n>([&](auto I) {
      constexpr int i = I;
      constexpr int L1 = B::l1(i);
      constexpr int L2 = B::l2(i);
      static_assert(L1 < 0 || L1 > i, "links must point forward");
      static_assert(L2 < 0 || L2 > i, "links must point forward");
      if constexpr (L1 >= 0 && L2 >= 0) {
        const real c1 = real(B::s1(i)) * a;
        const real c2 = real(B::s2(i)) * b;
        if constexpr (WITH_STATE) x[i] += c1 * x[L1] + c2 * x[L2];
        // diagonal first: it needs the old P(l, i)
        P[pidx(i, i)] += real(2) * (c1 * P[pidx(L1, i)] + c2 * P[pidx(L2, i)]) +
                         c1 * c1 * P[pidx(L1, L1)] + real(2) * (c1 * c2) * P[pidx(L1, L2)] +
                         c2 * c2 * P[pidx(L2, L2)];
        static_for<0, n>([&](auto J) {
          constexpr int j = J;
          if constexpr (j != i) P[pidx(i, j)] += c1 * P[pidx(L1, j)] + c2 * P[pidx(L2, j)];
        });
      } else if constexpr (L1 >= 0) {
        const real c1 = real(B::s1(i)) * a;
        if constexpr (WITH_STATE) x[i] += c1 * x[L1];
        P[pidx(i, i)] += real(2) * (c1 * P[pidx(L1, i)]) + c1 * c1 * P[pidx(L1, L1)];
        static_for<0, n>([&](auto J) {
          constexpr int j = J;
          if constexpr (j != i) P[pidx(i, j)] += c1 * P[pidx(L1, j)];
        });
      }
    });
    static_for<0, n>([&](auto I) {
      constexpr int i = I;
      P[pidx(i, i)] += real(mc.q[B::gidx(i)]) * dt;
    });
  }

  // x += B u for control-input models (u = payload columns pz)
  __device__ __forceinline__ void control(real dt, const ModelConst& mc, const real* u) {
    static_for<0, n>([&](auto I) {
      constexpr int i = I;
      if constexpr (B::cu(i) == 1) x[i] += (real(mc.c2) * (dt * dt)) * u[B::pz(i)];
      if constexpr (B::cu(i) == 2) x[i] += dt * u[B::pz(i)];
    });
  }

  // Scalar measurement of local state S with value z and variance r; inv = 1 / (P_ss + r) is
  // supplied by the caller so the blocks of a filter can share one reciprocal per stage.
  template <int S, bool WITH_STATE>
  __device__ __forceinline__ void scalar_update(real z, real r, real inv) {
    real row[n];
    static_for<0, n>([&](auto J) { row[J] = P[pidx(S, J)]; });
    if constexpr (WITH_STATE) {
      const real w = (z - x[S]) * inv;
      static_for<0, n>([&](auto I) { x[I] += row[I] * w; });
    }
    static_for<0, n>([&](auto I) {
      constexpr int i = I;
      if constexpr (i != S) {
        const real k = row[i] * inv;
        static_for<i, n>([&](auto J) {
          constexpr int j = J;
          if constexpr (j != S) P[pidx(i, j)] -= k * row[j];
        });
      }
    });
    const real shrink = r * inv;  // row S of (I - K H) P is row * r / (P_ss + r)
    static_for<0, n>([&](auto J) { P[pidx(S, J)] = row[J] * shrink; });
  }

  // IMU pseudo-measurement z built from the PREDICTED state (kf_workers.py:698-706); all of z is
  // formed before the first scalar update touches x.
  real zbuf[n];
  __device__ __forceinline__ void build_imu_z(real dt, const real* u) {
    static_for<0, n>([&](auto I) {
      constexpr int i = I;
      if constexpr (B::zk(i) == ZK_POS) {
        const real V = x[B::l1(i)] + u[B::pz(i)] * dt;
        zbuf[i] = x[i] + V * dt;
      } else if constexpr (B::zk(i) == ZK_VEL) {
        zbuf[i] = x[i] + u[B::pz(i)] * dt;
      } else if constexpr (B::zk(i) == ZK_DIRECT) {
        zbuf[i] = u[B::pz(i)];
      } else {
        zbuf[i] = real(0);
      }
    });
  }

  // ---- the block interface the filter-level code uses (BlockKF and SymBlock both provide it) ----
  template <int S, bool GPS_EVENT>
  __device__ __forceinline__ real stage_den(const ModelConst& mc) const {
    const real r = GPS_EVENT ? real(mc.r_gps[B::gps(S) >= 0 ? B::gps(S) : 0]) : real(mc.r_imu[B::gidx(S)]);
    return P[pidx(S, S)] + r;
  }
  template <int S, bool GPS_EVENT, bool WITH_STATE>
  __device__ __forceinline__ void stage_apply(const ModelConst& mc, const real* u, real inv) {
    const real r = GPS_EVENT ? real(mc.r_gps[B::gps(S) >= 0 ? B::gps(S) : 0]) : real(mc.r_imu[B::gidx(S)]);
    real z = real(0);
    if constexpr (WITH_STATE) z = GPS_EVENT ? u[B::gps(S) >= 0 ? B::gps(S) : 0] : zbuf[S];
    scalar_update<S, WITH_STATE>(z, r, inv);
  }
  __device__ __forceinline__ double det_total(bool& ok) const { return double(det(ok)); }
  template <typename T>
  __device__ __forceinline__ static bool blocks_identical(const T*, int) { return true; }
  template <typename T>
  static bool blocks_identical_host(const T*, int) { return true; }
  static bool constants_equal(const ModelConst&) { return true; }
  template <class F>
  __device__ __forceinline__ void for_each_state(F&& fn) const {  // fn(integral_constant global index, value)
    static_for<0, n>([&](auto I) { fn(std::integral_constant<int, B::gidx(I)>{}, x[I]); });
  }
  template <class F>
  __device__ __forceinline__ void for_each_cov(F&& fn) const {  // fn(global row, global col, value)
    static_for<0, n>([&](auto I) {
      static_for<0, n>([&](auto J) { fn(B::gidx(I), B::gidx(J), P[pidx(I, J)]); });
    });
  }
  __device__ __forceinline__ void load_state(const double* x0) {
    static_for<0, n>([&](auto I) { x[I] = real(x0[B::gidx(I)]); });
  }

  // ---- shared-covariance interface: the gain warp extracts K = P+ H^T R^-1 once per step, the math
  // warps predict and correct the state only (BlockKF and SymBlock both provide it) ----
  template <bool GPS_EVENT>
  static constexpr bool measured(int s) {
    return GPS_EVENT ? (B::gps(s) >= 0) : (B::zk(s) != ZK_NONE);
  }
  template <bool GPS_EVENT>
  static constexpr int n_meas() {
    int c = 0;
    for (int s = 0; s < n; ++s) c += measured<GPS_EVENT>(s) ? 1 : 0;
    return c;
  }
  template <bool GPS_EVENT>
  static constexpr int meas_state(int c) {  // local state of measured column c
    int k = 0;
    for (int s = 0; s < n; ++s)
      if (measured<GPS_EVENT>(s)) {
        if (k == c) return s;
        ++k;
      }
    return 0;
  }
  static constexpr int kslots = n * (n_meas<true>() > n_meas<false>() ? n_meas<true>() : n_meas<false>());
  template <typename T>
  __device__ __forceinline__ void load_cov_from(const T* __restrict__ Pg, int n_total) {
    static_for<0, n>([&](auto I) {
      constexpr int i = I;
      static_for<i, n>([&](auto J) {
        constexpr int j = J;
        P[pidx(i, j)] = real(0.5 * (double(Pg[B::gidx(i) * n_total + B::gidx(j)]) +
                                    double(Pg[B::gidx(j) * n_total + B::gidx(i)])));
      });
    });
  }
  template <typename T>
  __device__ __forceinline__ void load_state_from(const T* __restrict__ xg) {
    static_for<0, n>([&](auto I) { x[I] = real(xg[B::gidx(I)]); });
  }
  template <typename T>
  __device__ __forceinline__ void store_state_to(T* __restrict__ xg) const {
    static_for<0, n>([&](auto I) { xg[B::gidx(I)] = T(x[I]); });
  }
  // K[i][c] = P+[i][s_c] / r_{s_c} for the event's measured local states s_c (row-major [n][n_meas])
  // ir[c] = 1 / r_{s_c}, evaluated once per launch by the gain warp
  template <bool GPS_EVENT>
  __device__ __forceinline__ static void inv_noise(const ModelConst& mc, double* ir) {
    static_for<0, n_meas<GPS_EVENT>()>([&](auto C) {
      constexpr int sc = meas_state<GPS_EVENT>(C);
      ir[C] = 1.0 / (GPS_EVENT ? mc.r_gps[B::gps(sc) >= 0 ? B::gps(sc) : 0] : mc.r_imu[B::gidx(sc)]);
    });
  }
  template <bool GPS_EVENT>
  __device__ __forceinline__ void extract_gain(const double* ir, double* K) const {
    constexpr int NM = n_meas<GPS_EVENT>();
    static_for<0, NM>([&](auto C) {
      constexpr int c = C;
      constexpr int sc = meas_state<GPS_EVENT>(c);
      static_for<0, n>([&](auto I) { K[I * NM + c] = double(P[pidx(I, sc)]) * ir[c]; });
    });
  }
  __device__ __forceinline__ void predict_state(real dt, const ModelConst& mc) {
    const real a = dt;
    const real b = real(mc.c2) * (dt * dt);
    static_for<0, n>([&](auto I) {
      constexpr int i = I;
      constexpr int L1 = B::l1(i);
      constexpr int L2 = B::l2(i);
      if constexpr (L1 >= 0 && L2 >= 0) x[i] += (real(B::s1(i)) * a) * x[L1] + (real(B::s2(i)) * b) * x[L2];
      else if constexpr (L1 >= 0) x[i] += (real(B::s1(i)) * a) * x[L1];
    });
  }
  // x += K (z - H x) with the innovation formed from the predicted state before any component moves
  template <bool GPS_EVENT>
  __device__ __forceinline__ void apply_gain(const double* __restrict__ K, const real* u) {
    constexpr int NM = n_meas<GPS_EVENT>();
    if constexpr (NM > 0) {
      real y[NM];
      static_for<0, NM>([&](auto C) {
        constexpr int sc = meas_state<GPS_EVENT>(C);
        const real z = GPS_EVENT ? u[B::gps(sc) >= 0 ? B::gps(sc) : 0] : zbuf[sc];
        y[C] = z - x[sc];
      });
      static_for<0, n>([&](auto I) {
        real acc = x[I];
        static_for<0, NM>([&](auto C) { acc = fma(real(K[I * NM + C]), y[C], acc); });
        x[I] = acc;
      });
    }
  }
  // build the innovation from the predicted state and apply the gains in one pass (no z buffer):
  // y = z - x is (v + u dt) dt for the position pseudo-measurement, u dt for the velocity one
  // (kf_workers.py:698-709) and u - x for directly measured states
  template <bool GPS_EVENT>
  __device__ __forceinline__ void correct(const double* __restrict__ K, real dt, const real* u) {
    constexpr int NM = n_meas<GPS_EVENT>();
    if constexpr (NM > 0) {
      real y[NM];
      static_for<0, NM>([&](auto C) {
        constexpr int sc = meas_state<GPS_EVENT>(C);
        if constexpr (GPS_EVENT) y[C] = u[B::gps(sc) >= 0 ? B::gps(sc) : 0] - x[sc];
        else if constexpr (B::zk(sc) == ZK_POS) y[C] = (x[B::l1(sc)] + u[B::pz(sc)] * dt) * dt;
        else if constexpr (B::zk(sc) == ZK_VEL) y[C] = u[B::pz(sc)] * dt;
        else y[C] = u[B::pz(sc)] - x[sc];
      });
      static_for<0, n>([&](auto I) {
        real acc = x[I];
        static_for<0, NM>([&](auto C) { acc = fma(real(K[I * NM + C]), y[C], acc); });
        x[I] = acc;
      });
    }
  }
  // correct() from a COVARIANCE record: Pp = this block's P+ (upper triangle, pidx order), ir = 1 / r of the
  // event's measured columns; the gain K = P+ H^T R^-1 is formed where it is used (the same product, the same FMA)
  template <bool GPS_EVENT>
  __device__ __forceinline__ void correct_cov(const double* __restrict__ Pp, const double* __restrict__ ir, real dt,
                                              const real* u) {
    constexpr int NM = n_meas<GPS_EVENT>();
    if constexpr (NM > 0) {
      real y[NM];
      static_for<0, NM>([&](auto C) {
        constexpr int sc = meas_state<GPS_EVENT>(C);
        if constexpr (GPS_EVENT) y[C] = u[B::gps(sc) >= 0 ? B::gps(sc) : 0] - x[sc];
        else if constexpr (B::zk(sc) == ZK_POS) y[C] = (x[B::l1(sc)] + u[B::pz(sc)] * dt) * dt;
        else if constexpr (B::zk(sc) == ZK_VEL) y[C] = u[B::pz(sc)] * dt;
        else y[C] = u[B::pz(sc)] - x[sc];
      });
      static_for<0, n>([&](auto I) {
        real acc = x[I];
        static_for<0, NM>([&](auto C) {
          constexpr int sc = meas_state<GPS_EVENT>(C);
          acc = fma(real(Pp[pidx(I, sc)] * ir[C]), y[C], acc);
        });
        x[I] = acc;
      });
    }
  }
  template <typename T>
  __device__ __forceinline__ void store_cov_to(const double* cov, T* __restrict__ Pg, int n_total) const {
    static_for<0, n>([&](auto I) {
      static_for<0, n>([&](auto J) { Pg[B::gidx(I) * n_total + B::gidx(J)] = T(cov[pidx(I, J)]); });
    });
  }

  // det P: closed form for n <= 3, LDL^T pivots otherwise.  ok is cleared on a non-positive pivot.
  __device__ __forceinline__ real det(bool& ok) const {
    if constexpr (n == 1) {
      ok = ok && (P[0] > real(0));
      return P[0];
    } else if constexpr (n == 2) {
      const real d = P[pidx(0, 0)] * P[pidx(1, 1)] - P[pidx(0, 1)] * P[pidx(0, 1)];
      ok = ok && (d > real(0));
      return d;
    } else if constexpr (n == 3) {
      const real a = P[pidx(0, 0)], b = P[pidx(0, 1)], c = P[pidx(0, 2)];
      const real d = P[pidx(1, 1)], e = P[pidx(1, 2)], f = P[pidx(2, 2)];
      const real r = a * (d * f - e * e) - b * (b * f - c * e) + c * (b * e - c * d);
      ok = ok && (r > real(0));
      return r;
    } else {
      real A[NP];
      static_for<0, NP>([&](auto I) { A[I] = P[I]; });
      real prod = real(1);
      static_for<0, n>([&](auto K) {
        constexpr int k = K;
        const real d = A[pidx(k, k)];
        ok = ok && (d > real(0));
        prod *= d;
        if constexpr (k + 1 < n) {
          const real inv = fast_rcp(d);
          static_for<k + 1, n>([&](auto I) {
            constexpr int i = I;
            const real l = A[pidx(k, i)] * inv;
            static_for<i, n>([&](auto J) {
              constexpr int j = J;
              A[pidx(i, j)] -= l * A[pidx(k, j)];
            });
          });
        }
      });
      return prod;
    }
  }
};

// --------------------------------------------------------------------------------------
// SymBlock: M chains that are the SAME model on different axes (x/y/z position chains, or
// roll/pitch/yaw) and whose covariance blocks are identical share ONE covariance.  The covariance
// and gain recursion never reads a measurement, so it is evaluated once and applied to the M
// member states.  Validity (identical P blocks, equal rate constants) is checked per launch.
// SymGroup<B0, B1, ...> is the tag that selects it inside a FilterT.
// --------------------------------------------------------------------------------------
template <class B0, class... Bs>
struct SymGroup : B0 {
  static constexpr bool is_group = true;
};

template <typename real, class B0, class... Bs>
struct SymBlock {
  static constexpr int n = B0::n;
  static constexpr int NP = n * (n + 1) / 2;
  static constexpr int M = 1 + sizeof...(Bs);
  using Traits = B0;
  using B = B0;
  real P[NP];
  real x[M][n];
  real zbuf[M][n];
  static constexpr int pidx(int i, int j) {
    return i <= j ? (i * n - i * (i - 1) / 2 + (j - i)) : (j * n - j * (j - 1) / 2 + (i - j));
  }
  template <class F>
  __device__ __forceinline__ static void members(F&& fn) {  // fn(member tables, member index)
    fn(B0{}, std::integral_constant<int, 0>{});
    members_tail<1, Bs...>(fn);
  }
  template <int K, class F>
  __device__ __forceinline__ static void members_tail(F&&) {}
  template <int K, class C0, class... Cs, class F>
  __device__ __forceinline__ static void members_tail(F&& fn) {
    fn(C0{}, std::integral_constant<int, K>{});
    members_tail<K + 1, Cs...>(fn);
  }

  // every member's block of P equals member 0's (device, one filter)
  template <typename T>
  __device__ __forceinline__ static bool blocks_identical(const T* Pg, int n_total) {
    bool same = true;
    members([&](auto Bm, auto) {
      using BM = decltype(Bm);
      static_for<0, n>([&](auto I) {
        static_for<0, n>([&](auto J) {
          same = same && (Pg[BM::gidx(I) * n_total + BM::gidx(J)] == Pg[B0::gidx(I) * n_total + B0::gidx(J)]);
        });
      });
    });
    return same;
  }
  template <typename T>
  static bool blocks_identical_host(const T* Pg, int n_total) {
    bool same = true;
    auto chk = [&](auto Bm) {
      using BM = decltype(Bm);
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
          same = same && (Pg[BM::gidx(i) * n_total + BM::gidx(j)] == Pg[B0::gidx(i) * n_total + B0::gidx(j)]);
    };
    (chk(Bs{}), ...);
    return same;
  }
  static bool constants_equal(const ModelConst& mc) {
    bool same = true;
    auto chk = [&](auto Bm) {
      using BM = decltype(Bm);
      for (int i = 0; i < n; ++i) {
        same = same && mc.q[BM::gidx(i)] == mc.q[B0::gidx(i)] && mc.r_imu[BM::gidx(i)] == mc.r_imu[B0::gidx(i)];
        if (B0::gps(i) >= 0) same = same && mc.r_gps[BM::gps(i)] == mc.r_gps[B0::gps(i)];
      }
    };
    (chk(Bs{}), ...);
    return same;
  }

  __device__ __forceinline__ void load(const real* __restrict__ xg, const real* __restrict__ Pg, int n_total) {
    static_for<0, n>([&](auto I) {
      constexpr int i = I;
      static_for<i, n>([&](auto J) {
        constexpr int j = J;
        P[pidx(i, j)] = real(0.5) * (Pg[B0::gidx(i) * n_total + B0::gidx(j)] + Pg[B0::gidx(j) * n_total + B0::gidx(i)]);
      });
    });
    members([&](auto Bm, auto Mi) {
      using BM = decltype(Bm);
      static_for<0, n>([&](auto I) { x[Mi][I] = xg[BM::gidx(I)]; });
    });
  }
  __device__ __forceinline__ void store(real* __restrict__ xg, real* __restrict__ Pg, int n_total) const {
    members([&](auto Bm, auto Mi) {
      using BM = decltype(Bm);
      static_for<0, n>([&](auto I) {
        xg[BM::gidx(I)] = x[Mi][I];
        static_for<0, n>([&](auto J) { Pg[BM::gidx(I) * n_total + BM::gidx(J)] = P[pidx(I, J)]; });
      });
    });
  }
  __device__ __forceinline__ void load_cov(const double* __restrict__ Pg, int n_total) {
    static_for<0, n>([&](auto I) {
      constexpr int i = I;
      static_for<i, n>([&](auto J) {
        constexpr int j = J;
        P[pidx(i, j)] = real(0.5 * (Pg[B0::gidx(i) * n_total + B0::gidx(j)] + Pg[B0::gidx(j) * n_total + B0::gidx(i)]));
      });
    });
#pragma unroll
    for (int m = 0; m < M; ++m) static_for<0, n>([&](auto I) { x[m][I] = real(0); });
  }
  __device__ __forceinline__ void load_state(const double* x0) {
    members([&](auto Bm, auto Mi) {
      using BM = decltype(Bm);
      static_for<0, n>([&](auto I) { x[Mi][I] = real(x0[BM::gidx(I)]); });
    });
  }

  // ---- shared-covariance interface (see BlockKF) ----
  template <bool GPS_EVENT>
  static constexpr bool measured(int s) {
    return GPS_EVENT ? (B0::gps(s) >= 0) : (B0::zk(s) != ZK_NONE);
  }
  template <bool GPS_EVENT>
  static constexpr int n_meas() {
    int c = 0;
    for (int s = 0; s < n; ++s) c += measured<GPS_EVENT>(s) ? 1 : 0;
    return c;
  }
  template <bool GPS_EVENT>
  static constexpr int meas_state(int c) {
    int k = 0;
    for (int s = 0; s < n; ++s)
      if (measured<GPS_EVENT>(s)) {
        if (k == c) return s;
        ++k;
      }
    return 0;
  }
  static constexpr int kslots = n * (n_meas<true>() > n_meas<false>() ? n_meas<true>() : n_meas<false>());
  template <typename T>
  __device__ __forceinline__ void load_cov_from(const T* __restrict__ Pg, int n_total) {
    static_for<0, n>([&](auto I) {
      constexpr int i = I;
      static_for<i, n>([&](auto J) {
        constexpr int j = J;
        P[pidx(i, j)] = real(0.5 * (double(Pg[B0::gidx(i) * n_total + B0::gidx(j)]) +
                                    double(Pg[B0::gidx(j) * n_total + B0::gidx(i)])));
      });
    });
  }
  template <typename T>
  __device__ __forceinline__ void load_state_from(const T* __restrict__ xg) {
    members([&](auto Bm, auto Mi) {
      using BM = decltype(Bm);
      static_for<0, n>([&](auto I) { x[Mi][I] = real(xg[BM::gidx(I)]); });
    });
  }
  template <typename T>
  __device__ __forceinline__ void store_state_to(T* __restrict__ xg) const {
    members([&](auto Bm, auto Mi) {
      using BM = decltype(Bm);
      static_for<0, n>([&](auto I) { xg[BM::gidx(I)] = T(x[Mi][I]); });
    });
  }
  template <bool GPS_EVENT>
  __device__ __forceinline__ static void inv_noise(const ModelConst& mc, double* ir) {
    static_for<0, n_meas<GPS_EVENT>()>([&](auto C) {
      constexpr int sc = meas_state<GPS_EVENT>(C);
      ir[C] = 1.0 / (GPS_EVENT ? mc.r_gps[B0::gps(sc) >= 0 ? B0::gps(sc) : 0] : mc.r_imu[B0::gidx(sc)]);
    });
  }
  template <bool GPS_EVENT>
  __device__ __forceinline__ void extract_gain(const double* ir, double* K) const {
    constexpr int NM = n_meas<GPS_EVENT>();
    static_for<0, NM>([&](auto C) {
      constexpr int c = C;
      constexpr int sc = meas_state<GPS_EVENT>(c);
      static_for<0, n>([&](auto I) { K[I * NM + c] = double(P[pidx(I, sc)]) * ir[c]; });
    });
  }
  __device__ __forceinline__ void predict_state(real dt, const ModelConst& mc) {
    const real a = dt;
    const real b = real(mc.c2) * (dt * dt);
    static_for<0, n>([&](auto I) {
      constexpr int i = I;
      constexpr int L1 = B0::l1(i);
      constexpr int L2 = B0::l2(i);
      if constexpr (L1 >= 0 && L2 >= 0) {
        const real c1 = real(B0::s1(i)) * a, c2 = real(B0::s2(i)) * b;
#pragma unroll
        for (int m = 0; m < M; ++m) x[m][i] += c1 * x[m][L1] + c2 * x[m][L2];
      } else if constexpr (L1 >= 0) {
        const real c1 = real(B0::s1(i)) * a;
#pragma unroll
        for (int m = 0; m < M; ++m) x[m][i] += c1 * x[m][L1];
      }
    });
  }
  template <bool GPS_EVENT>
  __device__ __forceinline__ void apply_gain(const double* __restrict__ K, const real* u) {
    constexpr int NM = n_meas<GPS_EVENT>();
    if constexpr (NM > 0) {
      real k[n * NM];
      static_for<0, n * NM>([&](auto E) { k[E] = real(K[E]); });  // one broadcast read per gain, all members use it
      members([&](auto Bm, auto Mi) {
        using BM = decltype(Bm);
        real y[NM];
        static_for<0, NM>([&](auto C) {
          constexpr int sc = meas_state<GPS_EVENT>(C);
          const real z = GPS_EVENT ? u[BM::gps(sc) >= 0 ? BM::gps(sc) : 0] : zbuf[Mi][sc];
          y[C] = z - x[Mi][sc];
        });
        static_for<0, n>([&](auto I) {
          real acc = x[Mi][I];
          static_for<0, NM>([&](auto C) { acc = fma(k[I * NM + C], y[C], acc); });
          x[Mi][I] = acc;
        });
      });
    }
  }
  template <bool GPS_EVENT>
  __device__ __forceinline__ void correct(const double* __restrict__ K, real dt, const real* u) {
    constexpr int NM = n_meas<GPS_EVENT>();
    if constexpr (NM > 0) {
      members([&](auto Bm, auto Mi) {
        using BM = decltype(Bm);
        real y[NM];
        static_for<0, NM>([&](auto C) {
          constexpr int sc = meas_state<GPS_EVENT>(C);
          if constexpr (GPS_EVENT) y[C] = u[BM::gps(sc) >= 0 ? BM::gps(sc) : 0] - x[Mi][sc];
          else if constexpr (BM::zk(sc) == ZK_POS) y[C] = (x[Mi][BM::l1(sc)] + u[BM::pz(sc)] * dt) * dt;
          else if constexpr (BM::zk(sc) == ZK_VEL) y[C] = u[BM::pz(sc)] * dt;
          else y[C] = u[BM::pz(sc)] - x[Mi][sc];
        });
        static_for<0, n>([&](auto I) {
          real acc = x[Mi][I];
          static_for<0, NM>([&](auto C) { acc = fma(real(K[I * NM + C]), y[C], acc); });
          x[Mi][I] = acc;
        });
      });
    }
  }
  // correct() from a COVARIANCE record (see BlockKF::correct_cov)
  template <bool GPS_EVENT>
  __device__ __forceinline__ void correct_cov(const double* __restrict__ Pp, const double* __restrict__ ir, real dt,
                                              const real* u) {
    constexpr int NM = n_meas<GPS_EVENT>();
    if constexpr (NM > 0) {
      members([&](auto Bm, auto Mi) {
        using BM = decltype(Bm);
        real y[NM];
        static_for<0, NM>([&](auto C) {
          constexpr int sc = meas_state<GPS_EVENT>(C);
          if constexpr (GPS_EVENT) y[C] = u[BM::gps(sc) >= 0 ? BM::gps(sc) : 0] - x[Mi][sc];
          else if constexpr (BM::zk(sc) == ZK_POS) y[C] = (x[Mi][BM::l1(sc)] + u[BM::pz(sc)] * dt) * dt;
          else if constexpr (BM::zk(sc) == ZK_VEL) y[C] = u[BM::pz(sc)] * dt;
          else y[C] = u[BM::pz(sc)] - x[Mi][sc];
        });
        static_for<0, n>([&](auto I) {
          real acc = x[Mi][I];
          static_for<0, NM>([&](auto C) {
            constexpr int sc = meas_state<GPS_EVENT>(C);
            acc = fma(real(Pp[pidx(I, sc)] * ir[C]), y[C], acc);
          });
          x[Mi][I] = acc;
        });
      });
    }
  }
  template <typename T>
  __device__ __forceinline__ void store_cov_to(const double* cov, T* __restrict__ Pg, int n_total) const {
    members([&](auto Bm, auto) {
      using BM = decltype(Bm);
      static_for<0, n>([&](auto I) {
        static_for<0, n>([&](auto J) { Pg[BM::gidx(I) * n_total + BM::gidx(J)] = T(cov[pidx(I, J)]); });
      });
    });
  }

  template <bool WITH_STATE>
  __device__ __forceinline__ void predict(real dt, const ModelConst& mc) {
    const real a = dt;
    const real b = real(mc.c2) * (dt * dt);
    static_for<0, n>([&](auto I) {
      constexpr int i = I;
      constexpr int L1 = B0::l1(i);
      constexpr int L2 = B0::l2(i);
      if constexpr (L1 >= 0 && L2 >= 0) {
        const real c1 = real(B0::s1(i)) * a;
        const real c2 = real(B0::s2(i)) * b;
        if constexpr (WITH_STATE) {
#pragma unroll
          for (int m = 0; m < M; ++m) x[m][i] += c1 * x[m][L1] + c2 * x[m][L2];
        }
        P[pidx(i, i)] += real(2) * (c1 * P[pidx(L1, i)] + c2 * P[pidx(L2, i)]) + c1 * c1 * P[pidx(L1, L1)] +
                         real(2) * (c1 * c2) * P[pidx(L1, L2)] + c2 * c2 * P[pidx(L2, L2)];
        static_for<0, n>([&](auto J) {
          constexpr int j = J;
          if constexpr (j != i) P[pidx(i, j)] += c1 * P[pidx(L1, j)] + c2 * P[pidx(L2, j)];
        });
      } else if constexpr (L1 >= 0) {
        const real c1 = real(B0::s1(i)) * a;
        if constexpr (WITH_STATE) {
#pragma unroll
          for (int m = 0; m < M; ++m) x[m][i] += c1 * x[m][L1];
        }
        P[pidx(i, i)] += real(2) * (c1 * P[pidx(L1, i)]) + c1 * c1 * P[pidx(L1, L1)];
        static_for<0, n>([&](auto J) {
          constexpr int j = J;
          if constexpr (j != i) P[pidx(i, j)] += c1 * P[pidx(L1, j)];
        });
      }
    });
    static_for<0, n>([&](auto I) { P[pidx(I, I)] += real(mc.q[B0::gidx(I)]) * dt; });
  }
  __device__ __forceinline__ void control(real dt, const ModelConst& mc, const real* u) {
    members([&](auto Bm, auto Mi) {
      using BM = decltype(Bm);
      static_for<0, n>([&](auto I) {
        constexpr int i = I;
        if constexpr (BM::cu(i) == 1) x[Mi][i] += (real(mc.c2) * (dt * dt)) * u[BM::pz(i)];
        if constexpr (BM::cu(i) == 2) x[Mi][i] += dt * u[BM::pz(i)];
      });
    });
  }
  __device__ __forceinline__ void build_imu_z(real dt, const real* u) {
    members([&](auto Bm, auto Mi) {
      using BM = decltype(Bm);
      static_for<0, n>([&](auto I) {
        constexpr int i = I;
        if constexpr (BM::zk(i) == ZK_POS) {
          const real V = x[Mi][BM::l1(i)] + u[BM::pz(i)] * dt;
          zbuf[Mi][i] = x[Mi][i] + V * dt;
        } else if constexpr (BM::zk(i) == ZK_VEL) {
          zbuf[Mi][i] = x[Mi][i] + u[BM::pz(i)] * dt;
        } else if constexpr (BM::zk(i) == ZK_DIRECT) {
          zbuf[Mi][i] = u[BM::pz(i)];
        } else {
          zbuf[Mi][i] = real(0);
        }
      });
    });
  }
  template <int S, bool GPS_EVENT>
  __device__ __forceinline__ real stage_den(const ModelConst& mc) const {
    const real r = GPS_EVENT ? real(mc.r_gps[B0::gps(S) >= 0 ? B0::gps(S) : 0]) : real(mc.r_imu[B0::gidx(S)]);
    return P[pidx(S, S)] + r;
  }
  template <int S, bool GPS_EVENT, bool WITH_STATE>
  __device__ __forceinline__ void stage_apply(const ModelConst& mc, const real* u, real inv) {
    const real r = GPS_EVENT ? real(mc.r_gps[B0::gps(S) >= 0 ? B0::gps(S) : 0]) : real(mc.r_imu[B0::gidx(S)]);
    real row[n];
    static_for<0, n>([&](auto J) { row[J] = P[pidx(S, J)]; });
    if constexpr (WITH_STATE) {
      members([&](auto Bm, auto Mi) {
        using BM = decltype(Bm);
        const real z = GPS_EVENT ? u[BM::gps(S) >= 0 ? BM::gps(S) : 0] : zbuf[Mi][S];
        const real w = (z - x[Mi][S]) * inv;
        static_for<0, n>([&](auto I) { x[Mi][I] += row[I] * w; });
      });
    }
    static_for<0, n>([&](auto I) {
      constexpr int i = I;
      if constexpr (i != S) {
        const real k = row[i] * inv;
        static_for<i, n>([&](auto J) {
          constexpr int j = J;
          if constexpr (j != S) P[pidx(i, j)] -= k * row[j];
        });
      }
    });
    const real shrink = r * inv;
    static_for<0, n>([&](auto J) { P[pidx(S, J)] = row[J] * shrink; });
  }
  __device__ __forceinline__ real det(bool& ok) const {
    static_assert(n <= 3, "SymBlock chains have at most 3 states");
    real d;
    if constexpr (n == 1) {
      d = P[0];
    } else if constexpr (n == 2) {
      d = P[pidx(0, 0)] * P[pidx(1, 1)] - P[pidx(0, 1)] * P[pidx(0, 1)];
    } else {
      const real a = P[pidx(0, 0)], b = P[pidx(0, 1)], c = P[pidx(0, 2)];
      const real dd = P[pidx(1, 1)], e = P[pidx(1, 2)], f = P[pidx(2, 2)];
      d = a * (dd * f - e * e) - b * (b * f - c * e) + c * (b * e - c * dd);
    }
    ok = ok && (d > real(0));
    return d;
  }
  __device__ __forceinline__ double det_total(bool& ok) const {
    const double d = double(det(ok));
    double r = d;
#pragma unroll
    for (int m = 1; m < M; ++m) r *= d;
    return r;
  }
  template <class F>
  __device__ __forceinline__ void for_each_state(F&& fn) const {
    members([&](auto Bm, auto Mi) {
      using BM = decltype(Bm);
      static_for<0, n>([&](auto I) { fn(std::integral_constant<int, BM::gidx(I)>{}, x[Mi][I]); });
    });
  }
  template <class F>
  __device__ __forceinline__ void for_each_cov(F&& fn) const {
    members([&](auto Bm, auto) {
      using BM = decltype(Bm);
      static_for<0, n>([&](auto I) {
        static_for<0, n>([&](auto J) { fn(BM::gidx(I), BM::gidx(J), P[pidx(I, J)]); });
      });
    });
  }
};

// block object for a table type: plain chain -> BlockKF, SymGroup<...> -> SymBlock
template <class B, typename real>
struct BlockObj {
  using type = BlockKF<B, real>;
};
template <typename real, class B0, class... Bs>
struct BlockObj<SymGroup<B0, Bs...>, real> {
  using type = SymBlock<real, B0, Bs...>;
};

// --------------------------------------------------------------------------------------
// A filter = a tuple of blocks
// --------------------------------------------------------------------------------------
template <typename real, bool IMU_CONTROL, class... Bs>
struct FilterT;

// A chain the closed-form 2x2 path below can handle: (value, rate) with value += dt * rate and
// both states measured by the IMU event.
template <class B, class = void>
struct is_group_tag : std::false_type {};
template <class B>
struct is_group_tag<B, std::void_t<decltype(B::is_group)>> : std::true_type {};

template <class B>
constexpr bool is_pv2() {
  if (is_group_tag<B>::value) return false;
  if (B::n != 2) return false;
  return B::l1(0) == 1 && B::l2(0) < 0 && B::l1(1) < 0 && B::l2(1) < 0 && B::s1(0) == 1 &&
         (B::zk(0) == ZK_POS || B::zk(0) == ZK_DIRECT) && (B::zk(1) == ZK_VEL || B::zk(1) == ZK_DIRECT) &&
         B::cu(0) == 0 && B::cu(1) == 0 && B::gps(1) < 0;
}

template <typename real, bool IMU_CONTROL>
struct FilterT<real, IMU_CONTROL> {
  static constexpr bool fast2 = !IMU_CONTROL;
  static constexpr int nblocks = 0;
  template <typename T>
  __device__ __forceinline__ static bool blocks_identical(const T*, int) { return true; }
  template <typename T>
  static bool blocks_identical_host(const T*, int) { return true; }
  static bool constants_equal(const ModelConst&) { return true; }
  template <class F>
  __device__ __forceinline__ void each(F&&) {}
  template <class F>
  __device__ __forceinline__ void each(F&&) const {}
};

template <typename real, bool IMU_CONTROL, class B0, class... Bs>
struct FilterT<real, IMU_CONTROL, B0, Bs...> {
  typename BlockObj<B0, real>::type head;
  FilterT<real, IMU_CONTROL, Bs...> tail;
  static constexpr bool fast2 = is_pv2<B0>() && FilterT<real, IMU_CONTROL, Bs...>::fast2;
  static constexpr int nblocks = 1 + FilterT<real, IMU_CONTROL, Bs...>::nblocks;
  using Head = typename BlockObj<B0, real>::type;
  using Tail = FilterT<real, IMU_CONTROL, Bs...>;
  // symmetric-group validity: every group's member blocks of P identical / rate constants equal
  template <typename T>
  __device__ __forceinline__ static bool blocks_identical(const T* Pg, int n) {
    return Head::blocks_identical(Pg, n) && Tail::blocks_identical(Pg, n);
  }
  template <typename T>
  static bool blocks_identical_host(const T* Pg, int n) {
    return Head::blocks_identical_host(Pg, n) && Tail::blocks_identical_host(Pg, n);
  }
  static bool constants_equal(const ModelConst& mc) { return Head::constants_equal(mc) && Tail::constants_equal(mc); }
  template <class F>
  __device__ __forceinline__ void each(F&& f) {
    f(head);
    tail.each(f);
  }
  template <class F>
  __device__ __forceinline__ void each(F&& f) const {
    f(head);
    tail.each(f);
  }
};

enum : int { EV_GPS = 0, EV_IMU = 1, EV_PREDICT_ONLY = 2 };

// One filter step on every block.  `u` is the event payload (d reals).
template <bool WITH_STATE, typename real, bool IMU_CONTROL, class... Bs>
__device__ __forceinline__ void filter_predict(FilterT<real, IMU_CONTROL, Bs...>& f, real dt,
                                               const ModelConst& mc, int etype, const real* u) {
  f.each([&](auto& blk) {
    blk.template predict<WITH_STATE>(dt, mc);
    if constexpr (IMU_CONTROL && WITH_STATE) {
      if (etype == EV_IMU) blk.control(dt, mc, u);
    }
  });
}

// Update = a sequence of scalar updates (exact for selector H and diagonal R).  Stage S handles
// local state S of every block that measures it; the blocks of a stage are independent, so their
// reciprocals 1 / (P_ss + r) come from ONE hardware reciprocal (rcp_batch).
template <class... Bs>
constexpr int max_block_n() {
  int m = 0;
  ((m = Bs::n > m ? Bs::n : m), ...);
  return m;
}

template <bool GPS_EVENT, class B>
constexpr bool measures(int s) {
  if (s >= B::n) return false;
  return GPS_EVENT ? (B::gps(s) >= 0) : (B::zk(s) != ZK_NONE);
}

// --------------------------------------------------------------------------------------
// Joint covariance update for the single-lane recursions (the shared-covariance kernels' chain warps, the
// greedy selection): a block whose n <= 3 states are ALL measured by the IMU event (H = I on the block,
// R = diag r) in closed form, P+ = R S^-1 P with S = P + R inverted through its cofactors:
//   diagonal      P+_ii = r_i (sum_k C_ik p_ki) / det S
//   off-diagonal  P+_ij = -r_i r_j C_ij / det S            (from P+ = R - R S^-1 R)
// -- the reference's own formulation (K = P H^T inv(S), kf_workers.py:616-621), no cancelling subtraction
// beyond the cofactors of an SPD matrix with R on its diagonal.  Arithmetic is about that of the n
// sequential scalar updates (44 vs 42 FP64 for n = 3), but ONE reciprocal and a dependency depth of
// ~12 operations instead of three reciprocal-bound stages: a lane that runs the recursion alone is bound
// by latency (8 cycles per dependent DFMA, ~40 per reciprocal on B200; tools/lat_probe.cu), not by flops.
// --------------------------------------------------------------------------------------
template <class BT>
constexpr bool joint_imu_ok() {
  using TB = typename BT::Traits;
  if (BT::n != 2 && BT::n != 3) return false;
  for (int i = 0; i < BT::n; ++i) {
    if (!measures<false, TB>(i)) return false;
  }
  return true;
}

template <class BT>
__device__ __forceinline__ void joint_cov_update_imu(BT& b, const ModelConst& mc) {
  using TB = typename BT::Traits;
  using real = std::remove_cv_t<std::remove_reference_t<decltype(b.P[0])>>;
  if constexpr (BT::n == 2) {
    const real r0 = real(mc.r_imu[TB::gidx(0)]), r1 = real(mc.r_imu[TB::gidx(1)]);
    const real a = b.P[BT::pidx(0, 0)], o = b.P[BT::pidx(0, 1)], c = b.P[BT::pidx(1, 1)];
    const real detp = fma(a, c, -(o * o));
    const real t0 = fma(a, r1, detp), t1 = fma(c, r0, detp);
    const real inv = fast_rcp(fma(r0, c + r1, t0));
    const real i0 = r0 * inv, i1 = r1 * inv;
    b.P[BT::pidx(0, 0)] = t0 * i0;
    b.P[BT::pidx(1, 1)] = t1 * i1;
    b.P[BT::pidx(0, 1)] = o * (i0 * r1);
  } else {
    const real r0 = real(mc.r_imu[TB::gidx(0)]), r1 = real(mc.r_imu[TB::gidx(1)]), r2 = real(mc.r_imu[TB::gidx(2)]);
    const real p00 = b.P[BT::pidx(0, 0)], p01 = b.P[BT::pidx(0, 1)], p02 = b.P[BT::pidx(0, 2)];
    const real p11 = b.P[BT::pidx(1, 1)], p12 = b.P[BT::pidx(1, 2)], p22 = b.P[BT::pidx(2, 2)];
    const real s00 = p00 + r0, s11 = p11 + r1, s22 = p22 + r2;
    const real c00 = fma(s11, s22, -(p12 * p12));
    const real c01 = fma(p02, p12, -(p01 * s22));
    const real c02 = fma(p01, p12, -(p02 * s11));
    const real c11 = fma(s00, s22, -(p02 * p02));
    const real c12 = fma(p01, p02, -(s00 * p12));
    const real c22 = fma(s00, s11, -(p01 * p01));
    const real det = fma(s00, c00, fma(p01, c01, p02 * c02));
    const real n0 = fma(p00, c00, fma(p01, c01, p02 * c02));
    const real n1 = fma(p01, c01, fma(p11, c11, p12 * c12));
    const real n2 = fma(p02, c02, fma(p12, c12, p22 * c22));
    const real inv = fast_rcp(det);
    const real i0 = r0 * inv, i1 = r1 * inv, i2 = r2 * inv;
    b.P[BT::pidx(0, 0)] = n0 * i0;
    b.P[BT::pidx(1, 1)] = n1 * i1;
    b.P[BT::pidx(2, 2)] = n2 * i2;
    b.P[BT::pidx(0, 1)] = -((c01 * i0) * r1);
    b.P[BT::pidx(0, 2)] = -((c02 * i0) * r2);
    b.P[BT::pidx(1, 2)] = -((c12 * i1) * r2);
  }
}

template <typename real, bool IMU_CONTROL, class... Bs>
constexpr bool all_blocks_joint_ok(const FilterT<real, IMU_CONTROL, Bs...>*) {
  return (joint_imu_ok<typename BlockObj<Bs, real>::type>() && ...);
}

template <bool GPS_EVENT, bool WITH_STATE, typename real, bool IMU_CONTROL, class... Bs>
__device__ __forceinline__ void staged_update(FilterT<real, IMU_CONTROL, Bs...>& f, const ModelConst& mc,
                                              const real* u) {
  constexpr int MAXN = max_block_n<Bs...>();
  static_for<0, MAXN>([&](auto Sc) {
    constexpr int S = Sc;
    constexpr int NPART = (0 + ... + (measures<GPS_EVENT, Bs>(S) ? 1 : 0));
    if constexpr (NPART > 0) {
      real den[NPART], inv[NPART];
      int i = 0;
      f.each([&](auto& blk) {
        using BT = std::remove_reference_t<decltype(blk)>;
        using TB = typename BT::Traits;
        if constexpr (measures<GPS_EVENT, TB>(S)) {
          den[i] = blk.template stage_den<S, GPS_EVENT>(mc);
          ++i;
        }
      });
      rcp_batch<NPART>(den, inv);
      i = 0;
      f.each([&](auto& blk) {
        using BT = std::remove_reference_t<decltype(blk)>;
        using TB = typename BT::Traits;
        if constexpr (measures<GPS_EVENT, TB>(S)) {
          blk.template stage_apply<S, GPS_EVENT, WITH_STATE>(mc, u, inv[i]);
          ++i;
        }
      });
    }
  });
}

// JOINT (covariance-only callers that run the recursion on a single lane): the IMU event of a filter whose
// blocks are all jointly updatable takes joint_cov_update_imu instead of the scalar stages.
template <bool WITH_STATE, bool JOINT = false, typename real, bool IMU_CONTROL, class... Bs>
__device__ __forceinline__ void filter_update(FilterT<real, IMU_CONTROL, Bs...>& f, real dt,
                                              const ModelConst& mc, int etype, const real* u) {
  if (etype == EV_GPS) {
    staged_update<true, WITH_STATE>(f, mc, u);
  } else if (etype == EV_IMU) {
    if constexpr (!IMU_CONTROL) {
      if constexpr (JOINT && !WITH_STATE &&
                    all_blocks_joint_ok(static_cast<const FilterT<real, IMU_CONTROL, Bs...>*>(nullptr))) {
        f.each([&](auto& blk) { joint_cov_update_imu(blk, mc); });
      } else {
        if constexpr (WITH_STATE) f.each([&](auto& blk) { blk.build_imu_z(dt, u); });
        staged_update<false, WITH_STATE>(f, mc, u);
      }
    }
  }
}

template <typename real, bool IMU_CONTROL, class... Bs>
__device__ __forceinline__ double filter_det(const FilterT<real, IMU_CONTROL, Bs...>& f, bool& ok) {
  double d = 1.0;
  f.each([&](const auto& blk) { d *= blk.det_total(ok); });
  return d;
}

// --------------------------------------------------------------------------------------
// Closed-form path for filters made only of (value, rate) chains (CV6 decoupled).
//
// IMU event on one chain, P = [[a, b], [b, c]], H = I2, R = diag(r0, r1), S = P + R:
//   D    = det S = detP + a r1 + c r0 + r0 r1
//   P+   = R S^-1 P = (1/D) [[r0 (detP + a r1), r0 r1 b], [r0 r1 b, r1 (detP + c r0)]]
//   K    = P S^-1   = (1/D) [[detP + a r1, r0 b], [r1 b, detP + c r0]]
//   det P+ = r0 r1 detP / D
// which is (I - KH)P of kf_workers.py:711 written without a single cancelling subtraction.  The
// one reciprocal per chain is shared across the chains of a filter (rcp_batch): one MUFU per step.
// --------------------------------------------------------------------------------------
template <typename real>
struct Upd2 {
  real t0, t1, detp;
};

template <bool WITH_STATE, class B, typename real>
__device__ __forceinline__ void pv2_predict(BlockKF<B, real>& k, real dt, const ModelConst& mc) {
  real& a = k.P[0];
  real& b = k.P[1];
  real& c = k.P[2];
  if constexpr (WITH_STATE) k.x[0] = fma(dt, k.x[1], k.x[0]);
  const real bn = fma(dt, c, b);
  a = fma(dt, b + bn, fma(real(mc.q[B::gidx(0)]), dt, a));  // a + 2 dt b + dt^2 c + q0 dt
  b = bn;
  c = fma(real(mc.q[B::gidx(1)]), dt, c);
}

template <class B, typename real>
__device__ __forceinline__ real pv2_imu_begin(const BlockKF<B, real>& k, const ModelConst& mc,
                                              Upd2<real>& w) {
  const real a = k.P[0], b = k.P[1], c = k.P[2];
  const real r0 = real(mc.r_imu[B::gidx(0)]), r1 = real(mc.r_imu[B::gidx(1)]);
  w.detp = fma(a, c, -(b * b));
  w.t0 = fma(a, r1, w.detp);
  w.t1 = fma(c, r0, w.detp);
  return fma(r0, c + r1, w.t0);
}

template <bool WITH_STATE, class B, typename real>
__device__ __forceinline__ void pv2_imu_finish(BlockKF<B, real>& k, real dt, const ModelConst& mc,
                                               const real* u, real inv, const Upd2<real>& w) {
  const real r0 = real(mc.r_imu[B::gidx(0)]), r1 = real(mc.r_imu[B::gidx(1)]);
  const real bi = k.P[1] * inv;
  const real k00 = w.t0 * inv, k11 = w.t1 * inv;  // K = P S^-1 = P+ R^-1
  const real k01 = bi * r0, k10 = bi * r1;
  if constexpr (WITH_STATE) {
    // innovation y = z - x with z built from the predicted state (kf_workers.py:698-709)
    real y0, y1;
    if constexpr (B::zk(1) == ZK_VEL) y1 = u[B::pz(1)] * dt;
    else y1 = u[B::pz(1)] - k.x[1];
    if constexpr (B::zk(0) == ZK_POS) y0 = (k.x[1] + u[B::pz(0)] * dt) * dt;
    else y0 = u[B::pz(0)] - k.x[0];
    k.x[0] = fma(k00, y0, fma(k01, y1, k.x[0]));
    k.x[1] = fma(k10, y0, fma(k11, y1, k.x[1]));
  }
  k.P[0] = k00 * r0;
  k.P[1] = k01 * r1;
  k.P[2] = k11 * r1;
}

// GPS event on a chain whose state 0 is measured (variance r): scalar update, reciprocal shared.
template <bool WITH_STATE, class B, typename real>
__device__ __forceinline__ void pv2_gps_finish(BlockKF<B, real>& k, const ModelConst& mc, real inv,
                                               const real* u) {
  const real a = k.P[0], b = k.P[1];
  const real r = real(mc.r_gps[B::gps(0)]);
  const real k0 = a * inv, k1 = b * inv, g = r * inv;
  if constexpr (WITH_STATE) {
    const real y = u[B::gps(0)] - k.x[0];
    k.x[0] = fma(k0, y, k.x[0]);
    k.x[1] = fma(k1, y, k.x[1]);
  }
  k.P[2] = fma(-k1, b, k.P[2]);
  k.P[0] = a * g;
  k.P[1] = b * g;
}

template <bool WITH_STATE, typename real, bool IMU_CONTROL, class... Bs>
__device__ __forceinline__ void fast2_predict(FilterT<real, IMU_CONTROL, Bs...>& f, real dt,
                                              const ModelConst& mc) {
  f.each([&](auto& blk) { pv2_predict<WITH_STATE>(blk, dt, mc); });
}

// prod over chains of r0 * r1: det P+ = rprod * prod(det P-) / prod(det S) on an IMU event
template <typename real, bool IMU_CONTROL, class... Bs>
__device__ __forceinline__ double fast2_rprod(const FilterT<real, IMU_CONTROL, Bs...>& f, const ModelConst& mc) {
  double r = 1.0;
  f.each([&](const auto& blk) {
    using BT = std::remove_cv_t<std::remove_reference_t<decltype(blk)>>;
    r *= mc.r_imu[BT::gidx(0)] * mc.r_imu[BT::gidx(1)];
  });
  return r;
}

// Update every chain for an event of type etype; returns det P+ (product over chains).
template <bool WITH_STATE, typename real, bool IMU_CONTROL, class... Bs>
__device__ __forceinline__ double fast2_update(FilterT<real, IMU_CONTROL, Bs...>& f, real dt,
                                               const ModelConst& mc, int etype, const real* u, double rprod,
                                               bool& ok) {
  constexpr int NB = FilterT<real, IMU_CONTROL, Bs...>::nblocks;
  real den[NB], inv[NB];
  if (etype == EV_IMU) {
    Upd2<real> w[NB];
    double detp = 1.0;
    int i = 0;
    f.each([&](auto& blk) {
      den[i] = pv2_imu_begin(blk, mc, w[i]);
      detp *= double(w[i].detp);
      ++i;
    });
    const double rall = rcp_batch<NB>(den, inv);
    i = 0;
    f.each([&](auto& blk) {
      pv2_imu_finish<WITH_STATE>(blk, dt, mc, u, inv[i], w[i]);
      ++i;
    });
    const double det = detp * (rall * rprod);
    ok = ok && (det > 0.0);
    return det;
  }
  if (etype == EV_GPS) {
    int i = 0;
    f.each([&](auto& blk) {
      using BT = std::remove_reference_t<decltype(blk)>;
      if constexpr (BT::Traits::gps(0) >= 0) den[i] = blk.P[0] + real(mc.r_gps[BT::Traits::gps(0)]);
      else den[i] = real(1);
      ++i;
    });
    rcp_batch<NB>(den, inv);
    i = 0;
    f.each([&](auto& blk) {
      using BT = std::remove_reference_t<decltype(blk)>;
      if constexpr (BT::Traits::gps(0) >= 0) pv2_gps_finish<WITH_STATE>(blk, mc, inv[i], u);
      ++i;
    });
  }
  return filter_det(f, ok);
}

// --------------------------------------------------------------------------------------
// Filter objects for the TMA kernel.  Dec2Filter: every (value, rate) chain carries its own
// covariance.  Sym2Filter: the chains of a filter are the SAME model on different axes (CV6: x, y,
// z share q, r and -- for the reference's axis-symmetric P0, kf_workers.py:651 -- P), so the
// covariance / gain recursion, which never reads a measurement, is evaluated ONCE per filter and
// applied to the M axis states.  Valid only while the M chain blocks of P are bit-identical and
// the rate constants agree (checked per launch: chain_compat_kernel + sym2_constants_equal);
// results are the same numbers as Dec2Filter, det P = (det chain)^M.
// --------------------------------------------------------------------------------------
template <class A, class B>
constexpr bool same_pv2_shape() {
  return is_pv2<A>() && is_pv2<B>() && A::zk(0) == B::zk(0) && A::zk(1) == B::zk(1) &&
         ((A::gps(0) >= 0) == (B::gps(0) >= 0));
}

template <typename real, class... Bs>
struct Dec2Filter {
  static constexpr bool rec_api = false;  // predict(dt, mc) / update(dt, mc, ...) interface
  using LeanRec = LeanRecT<real>;  // only (dt, type) are read
  static constexpr bool lean_ok = false;
  FilterT<real, false, Bs...> f;
  double rprod;
  __device__ __forceinline__ void load(const real* xg, const real* Pg, int n, const ModelConst& mc) {
    f.each([&](auto& b) { b.load(xg, Pg, n); });
    rprod = fast2_rprod(f, mc);
  }
  __device__ __forceinline__ void store(real* xg, real* Pg, int n) const {
    f.each([&](const auto& b) { b.store(xg, Pg, n); });
  }
  __device__ __forceinline__ void predict(real dt, const ModelConst& mc) { fast2_predict<true>(f, dt, mc); }
  __device__ __forceinline__ double update(real dt, const ModelConst& mc, int ty, const real* u, bool& ok) {
    return fast2_update<true>(f, dt, mc, ty, u, rprod, ok);
  }
  template <class F>
  __device__ __forceinline__ void for_each_state(F&& fn) const {  // fn(global index, value)
    f.each([&](const auto& b) {
      using BT = std::remove_cv_t<std::remove_reference_t<decltype(b)>>;
      static_for<0, BT::n>([&](auto I) { fn(std::integral_constant<int, BT::gidx(I)>{}, b.x[I]); });
    });
  }
  __device__ __forceinline__ double det(bool& ok) const { return filter_det(f, ok); }
};

template <typename real, class B0, class... Bs>
struct Sym2Filter {
  static constexpr bool rec_api = false;  // predict(dt, mc) / update(dt, mc, ...) interface
  static constexpr int M = 1 + sizeof...(Bs);
  using Lead = B0;
  static constexpr bool eligible = (same_pv2_shape<B0, Bs>() && ... && is_pv2<B0>());
  real a, b, c;  // the shared chain covariance [[a, b], [b, c]]
  real x0[M], x1[M];
  real r0r1;

  template <class F>
  __device__ __forceinline__ static void members(F&& fn) {
    int m = 0;
    fn(B0{}, m++);
    (fn(Bs{}, m++), ...);
  }
  __device__ __forceinline__ static double powM(double d) {
    double r = d;
#pragma unroll
    for (int i = 1; i < M; ++i) r *= d;
    return r;
  }
  // device-side validity check of one filter's P [n, n]
  __device__ __forceinline__ static bool is_symmetric(const real* Pg, int n) {
    bool same = true;
    members([&](auto Bm, int) {
      using BM = decltype(Bm);
      static_for<0, 2>([&](auto I) {
        static_for<0, 2>([&](auto J) {
          same = same && (Pg[BM::gidx(I) * n + BM::gidx(J)] == Pg[B0::gidx(I) * n + B0::gidx(J)]);
        });
      });
    });
    return same;
  }
  // host-side: the rate constants of all members agree
  static bool constants_equal(const ModelConst& mc) {
    bool same = true;
    auto chk = [&](auto Bm) {
      using BM = decltype(Bm);
      for (int i = 0; i < 2; ++i) {
        same = same && mc.q[BM::gidx(i)] == mc.q[B0::gidx(i)] && mc.r_imu[BM::gidx(i)] == mc.r_imu[B0::gidx(i)];
      }
      if (B0::gps(0) >= 0) same = same && mc.r_gps[BM::gps(0)] == mc.r_gps[B0::gps(0)];
    };
    (chk(Bs{}), ...);
    return same;
  }

  __device__ __forceinline__ void load(const real* xg, const real* Pg, int n, const ModelConst& mc) {
    a = Pg[B0::gidx(0) * n + B0::gidx(0)];
    b = real(0.5) * (Pg[B0::gidx(0) * n + B0::gidx(1)] + Pg[B0::gidx(1) * n + B0::gidx(0)]);
    c = Pg[B0::gidx(1) * n + B0::gidx(1)];
    members([&](auto Bm, int m) {
      using BM = decltype(Bm);
      x0[m] = xg[BM::gidx(0)];
      x1[m] = xg[BM::gidx(1)];
    });
    r0r1 = real(mc.r_imu[B0::gidx(0)] * mc.r_imu[B0::gidx(1)]);
  }
  __device__ __forceinline__ void store(real* xg, real* Pg, int n) const {
    members([&](auto Bm, int m) {
      using BM = decltype(Bm);
      xg[BM::gidx(0)] = x0[m];
      xg[BM::gidx(1)] = x1[m];
      Pg[BM::gidx(0) * n + BM::gidx(0)] = a;
      Pg[BM::gidx(0) * n + BM::gidx(1)] = b;
      Pg[BM::gidx(1) * n + BM::gidx(0)] = b;
      Pg[BM::gidx(1) * n + BM::gidx(1)] = c;
    });
  }
  template <class F>
  __device__ __forceinline__ void for_each_state(F&& fn) const {  // fn(global index, value)
    members([&](auto Bm, int m) {
      using BM = decltype(Bm);
      fn(std::integral_constant<int, BM::gidx(0)>{}, x0[m]);
      fn(std::integral_constant<int, BM::gidx(1)>{}, x1[m]);
    });
  }
  __device__ __forceinline__ void predict(real dt, const ModelConst& mc) {
#pragma unroll
    for (int m = 0; m < M; ++m) x0[m] = fma(dt, x1[m], x0[m]);
    const real bn = fma(dt, c, b);
    a = fma(dt, b + bn, fma(real(mc.q[B0::gidx(0)]), dt, a));
    b = bn;
    c = fma(real(mc.q[B0::gidx(1)]), dt, c);
  }
  __device__ __forceinline__ double det(bool& ok) const {
    const real d = fma(a, c, -(b * b));
    ok = ok && (d > real(0));
    return powM(double(d));
  }
  // Lean fused predict + update for the aggregate-only loop of (position, velocity) chains whose
  // IMU event is the reference's pseudo-measurement z = [p + (v + u dt) dt, v + u dt]
  // (kf_workers.py:698-704).  Two rewrites, both exact in exact arithmetic:
  //  * the covariance is carried in R-normalised coordinates a~ = a / r0, b~ = b / sqrt(r0 r1),
  //    c~ = c / r1, where the H = I update is P~+ = P~ (P~ + I)^-1: D~ = det P~ + a~ + c~ + 1,
  //    a~+ = (det P~ + a~) / D~, b~+ = b~ / D~, c~+ = (det P~ + c~) / D~ -- no multiplication by r;
  //  * the innovation is y = [(v + u dt) dt, u dt], so predict and update of the state collapse to
  //    x0+ = x0 + cA v + cB u, x1+ = cC v + cE u with four per-filter coefficients (4 FP64 per axis).
  // The per-step constants that depend only on (dt, q, r) arrive with the step record (LeanRec).
  // v returns det P+ / (r0 r1) of ONE chain for every event type; the constant factor and the chain
  // multiplicity are restored once at the end of the launch.
  static constexpr bool lean_ok =
      (B0::zk(0) == ZK_POS && B0::zk(1) == ZK_VEL && B0::pz(0) == B0::pz(1) && B0::gps(0) >= 0);
  using LeanRec = LeanRecT<real>;  // warp-uniform
  __device__ __forceinline__ void to_normalized(const ModelConst& mc) {
    const double r0 = mc.r_imu[B0::gidx(0)], r1 = mc.r_imu[B0::gidx(1)];
    a = real(double(a) / r0);
    b = real(double(b) / sqrt(r0 * r1));
    c = real(double(c) / r1);
  }
  __device__ __forceinline__ void from_normalized(const ModelConst& mc) {
    const double r0 = mc.r_imu[B0::gidx(0)], r1 = mc.r_imu[B0::gidx(1)];
    a = real(double(a) * r0);
    b = real(double(b) * sqrt(r0 * r1));
    c = real(double(c) * r1);
  }
  __device__ __forceinline__ void step_lean(const LeanRec& r, const ModelConst&, const real* u, real& v) {
    const real dt = real(r.dt), kap = real(r.kap);
    const real bn = fma(kap, c, b);
    const real ap = fma(kap, b + bn, a + real(r.phi0));
    const real cp = c + real(r.phi1);
    if (r.type == EV_IMU) {
      const real detp = fma(ap, cp, -(bn * bn));
      const real t0 = detp + ap, t1 = detp + cp;
      const real inv = fast_rcp3((t0 + real(1)) + cp);
      a = t0 * inv;
      c = t1 * inv;
      b = bn * inv;
      const real cA = fma(a, dt, dt);                        // dt + k00 dt
      const real cB = fma(a, real(r.dt2), real(r.lam) * b);  // k00 dt^2 + k01 dt
      const real cC = fma(kap, b, real(1));                  // 1 + k10 dt
      const real cE = fma(real(r.kdt), b, dt * c);           // k10 dt^2 + k11 dt
      members([&](auto Bm, int m) {
        using BM = decltype(Bm);
        const real acc = u[BM::pz(1)];
        x0[m] = fma(cA, x1[m], fma(cB, acc, x0[m]));
        x1[m] = fma(cE, acc, cC * x1[m]);
      });
      v = detp * inv;
      return;
    }
    if (r.type == EV_GPS) {
      const real gam = real(r.kdt);
      const real inv = fast_rcp3(ap + gam);
      const real k0 = ap * inv, bi = bn * inv, g = gam * inv;
      const real k1 = real(r.lam) * bi;
      members([&](auto Bm, int m) {
        using BM = decltype(Bm);
        const real xp = fma(dt, x1[m], x0[m]);
        const real y = u[BM::gps(0)] - xp;
        x0[m] = fma(k0, y, xp);
        x1[m] = fma(k1, y, x1[m]);
      });
      c = fma(-bi, bn, cp);
      a = ap * g;
      b = bn * g;
    } else {
#pragma unroll
      for (int m = 0; m < M; ++m) x0[m] = fma(dt, x1[m], x0[m]);
      a = ap;
      b = bn;
      c = cp;
    }
    v = fma(a, c, -(b * b));
  }

  __device__ __forceinline__ double update(real dt, const ModelConst& mc, int ty, const real* u, bool& ok) {
    if (ty == EV_IMU) {
      const real r0 = real(mc.r_imu[B0::gidx(0)]), r1 = real(mc.r_imu[B0::gidx(1)]);
      const real detp = fma(a, c, -(b * b));
      const real t0 = fma(a, r1, detp);
      const real t1 = fma(c, r0, detp);
      const real inv = fast_rcp(fma(r0, c + r1, t0));
      const real bi = b * inv;
      const real k00 = t0 * inv, k11 = t1 * inv, k01 = bi * r0, k10 = bi * r1;
      members([&](auto Bm, int m) {
        using BM = decltype(Bm);
        real y0, y1;
        if constexpr (BM::zk(1) == ZK_VEL) y1 = u[BM::pz(1)] * dt;
        else y1 = u[BM::pz(1)] - x1[m];
        if constexpr (BM::zk(0) == ZK_POS) y0 = (x1[m] + u[BM::pz(0)] * dt) * dt;
        else y0 = u[BM::pz(0)] - x0[m];
        x0[m] = fma(k00, y0, fma(k01, y1, x0[m]));
        x1[m] = fma(k10, y0, fma(k11, y1, x1[m]));
      });
      a = k00 * r0;
      b = k01 * r1;
      c = k11 * r1;
      const real dc = detp * (inv * r0r1);
      ok = ok && (dc > real(0));
      return powM(double(dc));
    }
    if (ty == EV_GPS) {
      if constexpr (B0::gps(0) >= 0) {
        const real r = real(mc.r_gps[B0::gps(0)]);
        const real inv = fast_rcp(a + r);
        const real k0 = a * inv, k1 = b * inv, g = r * inv;
        members([&](auto Bm, int m) {
          using BM = decltype(Bm);
          const real y = u[BM::gps(0)] - x0[m];
          x0[m] = fma(k0, y, x0[m]);
          x1[m] = fma(k1, y, x1[m]);
        });
        c = fma(-k1, b, c);
        a = a * g;
        b = b * g;
      }
    }
    return det(ok);
  }
};

// --------------------------------------------------------------------------------------
// Sym3CtrlFilter: the lean form of a filter made of M identical (value, rate, bias) chains whose IMU
// event is a control input (BIAS9: x = F x + B u, no update) and whose GPS
// event measures the value.  One 3 x 3 covariance per filter (axis-symmetric P, checked per launch
// like Sym2Filter), the M axis states in registers.  With F = [[1, dt, -e], [0, 1, -dt], [0, 0, 1]],
// e = c2 dt^2, the predict is written out entry by entry (12 FMA + 3 for Q); the IMU step costs
// 15 (covariance) + 10 (det) + 5 per axis.
// --------------------------------------------------------------------------------------
template <class B>
constexpr bool is_vrb3() {  // [value, rate, bias]: value += dt rate - c2 dt^2 bias, rate -= dt bias; control on value and rate
  if (is_group_tag<B>::value) return false;
  if (B::n != 3) return false;
  return B::l1(0) == 1 && B::l2(0) == 2 && B::s1(0) == 1 && B::s2(0) == -1 && B::l1(1) == 2 && B::l2(1) < 0 &&
         B::s1(1) == -1 && B::l1(2) < 0 && B::l2(2) < 0 && B::cu(0) == 1 && B::cu(1) == 2 && B::cu(2) == 0 &&
         B::gps(0) >= 0 && B::gps(1) < 0 && B::gps(2) < 0 && B::zk(0) == ZK_NONE;
}

template <typename real, class B0, class... Bs>
struct Sym3CtrlFilter {
  static constexpr bool rec_api = false;  // predict(dt, mc) / update(dt, mc, ...) interface
  static constexpr int M = 1 + sizeof...(Bs);
  static constexpr bool eligible = (is_vrb3<B0>() && ... && is_vrb3<Bs>());
  static constexpr bool lean_ok = eligible;
  using Lead = B0;
  real p00, p01, p02, p11, p12, p22;  // the shared chain covariance (upper triangle)
  real x0[M], x1[M], x2[M];
  real r0r1;  // unused scale of the lean aggregate (v is the chain determinant itself)

  using LeanRec = LeanRecT<real>;  // only (dt, dt^2, type) are read
  template <class F>
  __device__ __forceinline__ static void members(F&& fn) {
    int m = 0;
    fn(B0{}, m++);
    (fn(Bs{}, m++), ...);
  }
  __device__ __forceinline__ void load(const real* xg, const real* Pg, int n, const ModelConst&) {
    auto sym = [&](int i, int j) {
      return real(0.5) * (Pg[B0::gidx(i) * n + B0::gidx(j)] + Pg[B0::gidx(j) * n + B0::gidx(i)]);
    };
    p00 = sym(0, 0), p01 = sym(0, 1), p02 = sym(0, 2), p11 = sym(1, 1), p12 = sym(1, 2), p22 = sym(2, 2);
    members([&](auto Bm, int m) {
      using BM = decltype(Bm);
      x0[m] = xg[BM::gidx(0)];
      x1[m] = xg[BM::gidx(1)];
      x2[m] = xg[BM::gidx(2)];
    });
    r0r1 = real(1);
  }
  __device__ __forceinline__ void store(real* xg, real* Pg, int n) const {
    const real P[3][3] = {{p00, p01, p02}, {p01, p11, p12}, {p02, p12, p22}};
    members([&](auto Bm, int m) {
      using BM = decltype(Bm);
      xg[BM::gidx(0)] = x0[m];
      xg[BM::gidx(1)] = x1[m];
      xg[BM::gidx(2)] = x2[m];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) Pg[BM::gidx(i) * n + BM::gidx(j)] = P[i][j];
    });
  }
  template <class F>
  __device__ __forceinline__ void for_each_state(F&& fn) const {
    members([&](auto Bm, int m) {
      using BM = decltype(Bm);
      fn(std::integral_constant<int, BM::gidx(0)>{}, x0[m]);
      fn(std::integral_constant<int, BM::gidx(1)>{}, x1[m]);
      fn(std::integral_constant<int, BM::gidx(2)>{}, x2[m]);
    });
  }
  __device__ __forceinline__ void to_normalized(const ModelConst&) {}
  __device__ __forceinline__ void from_normalized(const ModelConst&) {}
  __device__ __forceinline__ real chain_det() const {
    return p00 * fma(p11, p22, -(p12 * p12)) - p01 * fma(p01, p22, -(p02 * p12)) + p02 * fma(p01, p12, -(p02 * p11));
  }
  __device__ __forceinline__ double det(bool& ok) const {
    const real d = chain_det();
    ok = ok && (d > real(0));
    double r = double(d);
#pragma unroll
    for (int i = 1; i < M; ++i) r *= double(d);
    return r;
  }
  // P = F P F^T + Q, entry by entry
  __device__ __forceinline__ void predict_cov(real dt, real e, const ModelConst& mc) {
    const real a00 = fma(-e, p02, fma(dt, p01, p00));
    const real a01 = fma(-e, p12, fma(dt, p11, p01));
    const real a02 = fma(-e, p22, fma(dt, p12, p02));
    const real a11 = fma(-dt, p12, p11);
    const real a12 = fma(-dt, p22, p12);
    p00 = fma(real(mc.q[B0::gidx(0)]), dt, fma(-e, a02, fma(dt, a01, a00)));
    p01 = fma(-dt, a02, a01);
    p02 = a02;
    p11 = fma(real(mc.q[B0::gidx(1)]), dt, fma(-dt, a12, a11));
    p12 = a12;
    p22 = fma(real(mc.q[B0::gidx(2)]), dt, p22);
  }
  __device__ __forceinline__ void predict_state(real dt, real e) {
#pragma unroll
    for (int m = 0; m < M; ++m) {
      x0[m] = fma(-e, x2[m], fma(dt, x1[m], x0[m]));
      x1[m] = fma(-dt, x2[m], x1[m]);
    }
  }
  __device__ __forceinline__ void gps_update(const ModelConst& mc, const real* u) {
    const real r = real(mc.r_gps[B0::gps(0)]);
    const real inv = fast_rcp3(p00 + r);
    const real k0 = p00 * inv, k1 = p01 * inv, k2 = p02 * inv, g = r * inv;
    members([&](auto Bm, int m) {
      using BM = decltype(Bm);
      const real y = u[BM::gps(0)] - x0[m];
      x0[m] = fma(k0, y, x0[m]);
      x1[m] = fma(k1, y, x1[m]);
      x2[m] = fma(k2, y, x2[m]);
    });
    p11 = fma(-k1, p01, p11);
    p12 = fma(-k1, p02, p12);
    p22 = fma(-k2, p02, p22);
    p00 *= g;  // row 0 of (I - K H) P is row * r / (P00 + r)
    p01 *= g;
    p02 *= g;
  }
  // fused predict (+ control or GPS update); v = det of the chain covariance after the step
  __device__ __forceinline__ void step_lean(const LeanRec& rec, const ModelConst& mc, const real* u, real& v) {
    const real dt = real(rec.dt);
    const real e = real(mc.c2) * real(rec.dt2);
    predict_cov(dt, e, mc);
    if (rec.type == EV_IMU) {  // x = F x + B u: the control enters value and rate
      members([&](auto Bm, int m) {
        using BM = decltype(Bm);
        const real acc = u[BM::pz(0)];
        x0[m] = fma(e, acc - x2[m], fma(dt, x1[m], x0[m]));
        x1[m] = fma(dt, acc - x2[m], x1[m]);
      });
    } else {
      predict_state(dt, e);
      if (rec.type == EV_GPS) gps_update(mc, u);
    }
    v = chain_det();
  }
  // the non-lean interface of the ring kernel (per-step outputs); the host dispatch keeps control-input
  // models with per-step outputs or the adaptive gate on the LDG kernel, so this only has to be correct
  __device__ __forceinline__ void predict(real dt, const ModelConst& mc) {
    const real e = real(mc.c2) * (dt * dt);
    predict_cov(dt, e, mc);
    predict_state(dt, e);
  }
  __device__ __forceinline__ double update(real dt, const ModelConst& mc, int ty, const real* u, bool& ok) {
    if (ty == EV_IMU) {
      const real e = real(mc.c2) * (dt * dt);
      members([&](auto Bm, int m) {
        using BM = decltype(Bm);
        x0[m] = fma(e, u[BM::pz(0)], x0[m]);
        x1[m] = fma(dt, u[BM::pz(0)], x1[m]);
      });
    } else if (ty == EV_GPS) {
      gps_update(mc, u);
    }
    return det(ok);
  }
};

// --------------------------------------------------------------------------------------
// Sym2CtrlFilter: the two-state form of Sym3CtrlFilter -- M identical (value, rate) chains whose IMU event is
// the control input x = F x + B u, B = [c2 dt^2, dt] (the north star's predict step as written; CVBU6), GPS
// measures the value.  One 2 x 2 covariance per filter: 6 (covariance) + 2 (det) + 3 per axis FP64 per IMU step.
// --------------------------------------------------------------------------------------
template <class B>
constexpr bool is_vr2c() {  // [value, rate]: value += dt rate; control on value (c2 dt^2 u) and rate (dt u)
  if (is_group_tag<B>::value) return false;
  if (B::n != 2) return false;
  return B::l1(0) == 1 && B::l2(0) < 0 && B::s1(0) == 1 && B::l1(1) < 0 && B::l2(1) < 0 && B::cu(0) == 1 &&
         B::cu(1) == 2 && B::gps(0) >= 0 && B::gps(1) < 0 && B::zk(0) == ZK_NONE && B::zk(1) == ZK_NONE;
}

template <typename real, class B0, class... Bs>
struct Sym2CtrlFilter {
  static constexpr bool rec_api = false;  // predict(dt, mc) / update(dt, mc, ...) interface
  static constexpr int M = 1 + sizeof...(Bs);
  static constexpr bool eligible = (is_vr2c<B0>() && ... && is_vr2c<Bs>());
  static constexpr bool lean_ok = eligible;
  using Lead = B0;
  real p00, p01, p11;  // the shared chain covariance (upper triangle)
  real x0[M], x1[M];
  real r0r1;  // unused scale of the lean aggregate (v is the chain determinant itself)

  using LeanRec = LeanRecT<real>;  // only (dt, dt^2, type) are read
  template <class F>
  __device__ __forceinline__ static void members(F&& fn) {
    int m = 0;
    fn(B0{}, m++);
    (fn(Bs{}, m++), ...);
  }
  __device__ __forceinline__ void load(const real* xg, const real* Pg, int n, const ModelConst&) {
    p00 = Pg[B0::gidx(0) * n + B0::gidx(0)];
    p01 = real(0.5) * (Pg[B0::gidx(0) * n + B0::gidx(1)] + Pg[B0::gidx(1) * n + B0::gidx(0)]);
    p11 = Pg[B0::gidx(1) * n + B0::gidx(1)];
    members([&](auto Bm, int m) {
      using BM = decltype(Bm);
      x0[m] = xg[BM::gidx(0)];
      x1[m] = xg[BM::gidx(1)];
    });
    r0r1 = real(1);
  }
  __device__ __forceinline__ void store(real* xg, real* Pg, int n) const {
    members([&](auto Bm, int m) {
      using BM = decltype(Bm);
      xg[BM::gidx(0)] = x0[m];
      xg[BM::gidx(1)] = x1[m];
      Pg[BM::gidx(0) * n + BM::gidx(0)] = p00;
      Pg[BM::gidx(0) * n + BM::gidx(1)] = p01;
      Pg[BM::gidx(1) * n + BM::gidx(0)] = p01;
      Pg[BM::gidx(1) * n + BM::gidx(1)] = p11;
    });
  }
  template <class F>
  __device__ __forceinline__ void for_each_state(F&& fn) const {
    members([&](auto Bm, int m) {
      using BM = decltype(Bm);
      fn(std::integral_constant<int, BM::gidx(0)>{}, x0[m]);
      fn(std::integral_constant<int, BM::gidx(1)>{}, x1[m]);
    });
  }
  __device__ __forceinline__ void to_normalized(const ModelConst&) {}
  __device__ __forceinline__ void from_normalized(const ModelConst&) {}
  __device__ __forceinline__ real chain_det() const { return fma(p00, p11, -(p01 * p01)); }
  __device__ __forceinline__ double det(bool& ok) const {
    const real d = chain_det();
    ok = ok && (d > real(0));
    double r = double(d);
#pragma unroll
    for (int i = 1; i < M; ++i) r *= double(d);
    return r;
  }
  // P = F P F^T + Q, entry by entry
  __device__ __forceinline__ void predict_cov(real dt, const ModelConst& mc) {
    const real a00 = fma(dt, p01, p00);
    const real a01 = fma(dt, p11, p01);
    p00 = fma(real(mc.q[B0::gidx(0)]), dt, fma(dt, a01, a00));
    p01 = a01;
    p11 = fma(real(mc.q[B0::gidx(1)]), dt, p11);
  }
  __device__ __forceinline__ void predict_state(real dt) {
#pragma unroll
    for (int m = 0; m < M; ++m) x0[m] = fma(dt, x1[m], x0[m]);
  }
  __device__ __forceinline__ void gps_update(const ModelConst& mc, const real* u) {
    const real r = real(mc.r_gps[B0::gps(0)]);
    const real inv = fast_rcp3(p00 + r);
    const real k0 = p00 * inv, k1 = p01 * inv, g = r * inv;
    members([&](auto Bm, int m) {
      using BM = decltype(Bm);
      const real y = u[BM::gps(0)] - x0[m];
      x0[m] = fma(k0, y, x0[m]);
      x1[m] = fma(k1, y, x1[m]);
    });
    p11 = fma(-k1, p01, p11);
    p00 *= g;  // row 0 of (I - K H) P is row * r / (P00 + r)
    p01 *= g;
  }
  // fused predict (+ control or GPS update); v = det of the chain covariance after the step
  __device__ __forceinline__ void step_lean(const LeanRec& rec, const ModelConst& mc, const real* u, real& v) {
    const real dt = real(rec.dt);
    const real e = real(mc.c2) * real(rec.dt2);
    predict_cov(dt, mc);
    if (rec.type == EV_IMU) {  // x = F x + B u: the control enters value and rate
      members([&](auto Bm, int m) {
        using BM = decltype(Bm);
        const real acc = u[BM::pz(0)];
        x0[m] = fma(e, acc, fma(dt, x1[m], x0[m]));
        x1[m] = fma(dt, acc, x1[m]);
      });
    } else {
      predict_state(dt);
      if (rec.type == EV_GPS) gps_update(mc, u);
    }
    v = chain_det();
  }
  // the non-lean interface of the ring kernel (per-step outputs)
  __device__ __forceinline__ void predict(real dt, const ModelConst& mc) {
    predict_cov(dt, mc);
    predict_state(dt);
  }
  __device__ __forceinline__ double update(real dt, const ModelConst& mc, int ty, const real* u, bool& ok) {
    if (ty == EV_IMU) {
      const real e = real(mc.c2) * (dt * dt);
      members([&](auto Bm, int m) {
        using BM = decltype(Bm);
        x0[m] = fma(e, u[BM::pz(0)], x0[m]);
        x1[m] = fma(dt, u[BM::pz(0)], x1[m]);
      });
    } else if (ty == EV_GPS) {
      gps_update(mc, u);
    }
    return det(ok);
  }
};

// the control-input ring filter of a model: the three-state or the two-state form, by the shape of its first chain
template <typename real, class B0, class... Bs>
struct CtrlSymSelect {
  using type = std::conditional_t<is_vr2c<B0>(), Sym2CtrlFilter<real, B0, Bs...>, Sym3CtrlFilter<real, B0, Bs...>>;
};

// --------------------------------------------------------------------------------------
// SymCAFilter: the lean form of the reference's OWN models (kf_workers.py:493-614 CA15, hw5_2.py:219-304 CA8):
// ML identical (position, velocity, acceleration) chains + MA identical (angle, rate) chains, every state
// measured by the IMU event (pseudo-measurement [X, V, a] on the linear chains, direct on the angular ones,
// kf_workers.py:698-706), position measured by the GPS event.  One 3 x 3 and one 2 x 2 covariance per filter
// (axis-symmetric P, checked per launch), the axis states in registers.
//
// The covariances are carried in R-NORMALISED coordinates P~ = D^-1 P D^-1, D = diag(sqrt r_imu): the H = I
// update is then P~+ = P~ (P~ + I)^-1, written through the cofactors of S~ = P~ + I --
//   diagonal      P~+_ii = (sum_k P~_ik C_ki) / det S~        off-diagonal  P~+_ij = -C_ij / det S~
// -- one reciprocal, no multiplication by r and no cancelling subtraction; det P~+ = det P~ / det S~.  The state
// stays in physical units: K = D P~+ D^-1, i.e. K_ii = P~+_ii and the off-diagonals scale by sqrt(r_i / r_j)
// (launch constants).  Everything that depends only on (dt, q, r) arrives with the step record (CARecT).
// One IMU step: 15 + 37 + 10 (3 x 3 covariance: predict, update, determinant) + 5 + 14 (2 x 2) + 8 (gain
// scaling) + 16 per linear axis + 7 per angular axis FP64 instructions.
// --------------------------------------------------------------------------------------
template <class... Cs>
struct ChainList {
  static constexpr int size = sizeof...(Cs);
};

template <typename real>
struct CARecT;
template <>
struct alignas(16) CARecT<double> {  // 112 bytes
  double dt, e;           // dt, c2 dt^2
  double k01, k02, k12;   // normalised F of the linear chain: dt sqrt(r1/r0), e sqrt(r2/r0), dt sqrt(r2/r1)
  double f0, f1, f2;      // q_i dt / r_i
  double ka, fa0, fa1;    // angular chain: dt sqrt(rb1/rb0), qb_i dt / rb_i
  double gam;             // r_gps / r0
  int type;
  int pad[3];
};
template <>
struct alignas(16) CARecT<float> {  // 64 bytes, same fields
  float dt, e, k01, k02, k12, f0, f1, f2, ka, fa0, fa1, gam;
  int type;
  int pad[3];
};

template <class B>
constexpr bool is_pva3() {  // [p, v, a]: p += dt v + c2 dt^2 a, v += dt a; IMU measures [X, V, a], GPS measures p
  if (is_group_tag<B>::value) return false;
  if (B::n != 3) return false;
  return B::l1(0) == 1 && B::l2(0) == 2 && B::s1(0) == 1 && B::s2(0) == 1 && B::l1(1) == 2 && B::l2(1) < 0 &&
         B::s1(1) == 1 && B::l1(2) < 0 && B::l2(2) < 0 && B::zk(0) == ZK_POS && B::zk(1) == ZK_VEL &&
         B::zk(2) == ZK_DIRECT && B::pz(0) == B::pz(1) && B::pz(1) == B::pz(2) && B::gps(0) >= 0 && B::gps(1) < 0 &&
         B::gps(2) < 0 && B::cu(0) == 0 && B::cu(1) == 0 && B::cu(2) == 0;
}
template <class B>
constexpr bool is_ar2() {  // [angle, rate]: angle += dt rate; both measured directly by the IMU event, no GPS
  if (is_group_tag<B>::value) return false;
  if (B::n != 2) return false;
  return B::l1(0) == 1 && B::l2(0) < 0 && B::s1(0) == 1 && B::l1(1) < 0 && B::l2(1) < 0 && B::zk(0) == ZK_DIRECT &&
         B::zk(1) == ZK_DIRECT && B::gps(0) < 0 && B::gps(1) < 0 && B::cu(0) == 0 && B::cu(1) == 0;
}

template <typename real, class LinList, class AngList>
struct SymCAFilter {  // models without the (linear chains, angular chains) shape
  static constexpr bool eligible = false;
  static constexpr bool lean_ok = false;
  static constexpr bool rec_api = false;
  static constexpr int M = 1;
  using LeanRec = LeanRecT<real>;
};

template <typename real, class L0, class... Ls, class A0, class... As>
struct SymCAFilter<real, ChainList<L0, Ls...>, ChainList<A0, As...>> {
  static constexpr int ML = 1 + sizeof...(Ls), MA = 1 + sizeof...(As);
  static constexpr bool eligible = (is_pva3<L0>() && ... && is_pva3<Ls>()) && (is_ar2<A0>() && ... && is_ar2<As>());
  static constexpr bool lean_ok = eligible;
  static constexpr bool rec_api = true;  // predict(rec) / update(rec, ...) take the step record, not (dt, mc)
  static constexpr int gcd_(int a, int b) { return b == 0 ? a : gcd_(b, a % b); }
  static constexpr int M = gcd_(ML, MA);          // multiplicity restored on the lean aggregate (kf_run_tma_kernel)
  static constexpr int EL = ML / M, EA = MA / M;  // v = (det lin)^EL (det ang)^EA per step
  using Lead = L0;
  using LeanRec = CARecT<real>;
  real l00, l01, l02, l11, l12, l22;  // linear-chain covariance (R-normalised between to_/from_normalized)
  real c00, c01, c11;                 // angular-chain covariance
  real xp[ML], xv[ML], xa[ML];
  real th[MA], om[MA];
  real s01, s02, s12, s10, s20, s21;  // sqrt(r_i / r_j) of the linear chain's IMU variances
  real t01, t10;                      // the same for the angular chain
  real r0r1;                          // (r0 r1 r2)^EL (rb0 rb1)^EA: constant factor of the lean aggregate's v

  template <class F>
  __device__ __forceinline__ static void lin_members(F&& fn) {
    int m = 0;
    fn(L0{}, m++);
    (fn(Ls{}, m++), ...);
  }
  template <class F>
  __device__ __forceinline__ static void ang_members(F&& fn) {
    int m = 0;
    fn(A0{}, m++);
    (fn(As{}, m++), ...);
  }
  __device__ __forceinline__ static real ipow(real v, int e) {
    real r = v;
    for (int i = 1; i < e; ++i) r *= v;
    return r;
  }

  __device__ __forceinline__ void load(const real* xg, const real* Pg, int n, const ModelConst& mc) {
    auto sl = [&](int i, int j) { return real(0.5) * (Pg[L0::gidx(i) * n + L0::gidx(j)] + Pg[L0::gidx(j) * n + L0::gidx(i)]); };
    auto sa = [&](int i, int j) { return real(0.5) * (Pg[A0::gidx(i) * n + A0::gidx(j)] + Pg[A0::gidx(j) * n + A0::gidx(i)]); };
    l00 = sl(0, 0), l01 = sl(0, 1), l02 = sl(0, 2), l11 = sl(1, 1), l12 = sl(1, 2), l22 = sl(2, 2);
    c00 = sa(0, 0), c01 = sa(0, 1), c11 = sa(1, 1);
    lin_members([&](auto Bm, int m) {
      using BM = decltype(Bm);
      xp[m] = xg[BM::gidx(0)];
      xv[m] = xg[BM::gidx(1)];
      xa[m] = xg[BM::gidx(2)];
    });
    ang_members([&](auto Bm, int m) {
      using BM = decltype(Bm);
      th[m] = xg[BM::gidx(0)];
      om[m] = xg[BM::gidx(1)];
    });
    const double r0 = mc.r_imu[L0::gidx(0)], r1 = mc.r_imu[L0::gidx(1)], r2 = mc.r_imu[L0::gidx(2)];
    const double b0 = mc.r_imu[A0::gidx(0)], b1 = mc.r_imu[A0::gidx(1)];
    s01 = real(sqrt(r0 / r1)), s02 = real(sqrt(r0 / r2)), s12 = real(sqrt(r1 / r2));
    s10 = real(sqrt(r1 / r0)), s20 = real(sqrt(r2 / r0)), s21 = real(sqrt(r2 / r1));
    t01 = real(sqrt(b0 / b1)), t10 = real(sqrt(b1 / b0));
    double rr = 1.0;
    for (int i = 0; i < EL; ++i) rr *= r0 * r1 * r2;
    for (int i = 0; i < EA; ++i) rr *= b0 * b1;
    r0r1 = real(rr);
  }
  __device__ __forceinline__ void store(real* xg, real* Pg, int n) const {
    const real L[3][3] = {{l00, l01, l02}, {l01, l11, l12}, {l02, l12, l22}};
    const real A[2][2] = {{c00, c01}, {c01, c11}};
    lin_members([&](auto Bm, int m) {
      using BM = decltype(Bm);
      xg[BM::gidx(0)] = xp[m];
      xg[BM::gidx(1)] = xv[m];
      xg[BM::gidx(2)] = xa[m];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) Pg[BM::gidx(i) * n + BM::gidx(j)] = L[i][j];
    });
    ang_members([&](auto Bm, int m) {
      using BM = decltype(Bm);
      xg[BM::gidx(0)] = th[m];
      xg[BM::gidx(1)] = om[m];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) Pg[BM::gidx(i) * n + BM::gidx(j)] = A[i][j];
    });
  }
  template <class F>
  __device__ __forceinline__ void for_each_state(F&& fn) const {
    lin_members([&](auto Bm, int m) {
      using BM = decltype(Bm);
      fn(std::integral_constant<int, BM::gidx(0)>{}, xp[m]);
      fn(std::integral_constant<int, BM::gidx(1)>{}, xv[m]);
      fn(std::integral_constant<int, BM::gidx(2)>{}, xa[m]);
    });
    ang_members([&](auto Bm, int m) {
      using BM = decltype(Bm);
      fn(std::integral_constant<int, BM::gidx(0)>{}, th[m]);
      fn(std::integral_constant<int, BM::gidx(1)>{}, om[m]);
    });
  }
  // P <-> P~ = D^-1 P D^-1 (exact scalings by launch constants, once per launch each way)
  __device__ __forceinline__ void to_normalized(const ModelConst& mc) {
    const double r0 = mc.r_imu[L0::gidx(0)], r1 = mc.r_imu[L0::gidx(1)], r2 = mc.r_imu[L0::gidx(2)];
    const double b0 = mc.r_imu[A0::gidx(0)], b1 = mc.r_imu[A0::gidx(1)];
    l00 = real(double(l00) / r0), l11 = real(double(l11) / r1), l22 = real(double(l22) / r2);
    l01 = real(double(l01) / sqrt(r0 * r1)), l02 = real(double(l02) / sqrt(r0 * r2)), l12 = real(double(l12) / sqrt(r1 * r2));
    c00 = real(double(c00) / b0), c11 = real(double(c11) / b1), c01 = real(double(c01) / sqrt(b0 * b1));
  }
  __device__ __forceinline__ void from_normalized(const ModelConst& mc) {
    const double r0 = mc.r_imu[L0::gidx(0)], r1 = mc.r_imu[L0::gidx(1)], r2 = mc.r_imu[L0::gidx(2)];
    const double b0 = mc.r_imu[A0::gidx(0)], b1 = mc.r_imu[A0::gidx(1)];
    l00 = real(double(l00) * r0), l11 = real(double(l11) * r1), l22 = real(double(l22) * r2);
    l01 = real(double(l01) * sqrt(r0 * r1)), l02 = real(double(l02) * sqrt(r0 * r2)), l12 = real(double(l12) * sqrt(r1 * r2));
    c00 = real(double(c00) * b0), c11 = real(double(c11) * b1), c01 = real(double(c01) * sqrt(b0 * b1));
  }
  // normalised determinants of the two covariances
  __device__ __forceinline__ real lin_det() const {
    return l00 * fma(l11, l22, -(l12 * l12)) - l01 * fma(l01, l22, -(l02 * l12)) + l02 * fma(l01, l12, -(l02 * l11));
  }
  __device__ __forceinline__ real ang_det() const { return fma(c00, c11, -(c01 * c01)); }
  __device__ __forceinline__ real vnorm() const { return ipow(lin_det(), EL) * ipow(ang_det(), EA); }
  // det P of the whole filter in physical units (normalised state assumed)
  __device__ __forceinline__ double det(bool& ok) const {
    const real dl = lin_det(), da = ang_det();
    ok = ok && (dl > real(0)) && (da > real(0));
    double r = 1.0;
    const double d = double(ipow(dl, EL)) * double(ipow(da, EA)) * double(r0r1);
#pragma unroll
    for (int i = 0; i < M; ++i) r *= d;
    return r;
  }
  // x = F x, P~ = F~ P~ F~^T + Q~ (entry by entry)
  __device__ __forceinline__ void predict(const LeanRec& r) {
    const real dt = r.dt, e = r.e;
#pragma unroll
    for (int m = 0; m < ML; ++m) {
      xp[m] = fma(e, xa[m], fma(dt, xv[m], xp[m]));
      xv[m] = fma(dt, xa[m], xv[m]);
    }
#pragma unroll
    for (int m = 0; m < MA; ++m) th[m] = fma(dt, om[m], th[m]);
    const real k01 = r.k01, k02 = r.k02, k12 = r.k12;
    const real a00 = fma(k02, l02, fma(k01, l01, l00));
    const real a01 = fma(k02, l12, fma(k01, l11, l01));
    const real a02 = fma(k02, l22, fma(k01, l12, l02));
    const real a11 = fma(k12, l12, l11);
    const real a12 = fma(k12, l22, l12);
    l00 = fma(k02, a02, fma(k01, a01, a00)) + real(r.f0);
    l01 = fma(k12, a02, a01);
    l02 = a02;
    l11 = fma(k12, a12, a11) + real(r.f1);
    l12 = a12;
    l22 = l22 + real(r.f2);
    const real ka = r.ka;
    const real bn = fma(ka, c11, c01);
    c00 = fma(ka, c01 + bn, c00 + real(r.fa0));
    c01 = bn;
    c11 = c11 + real(r.fa1);
  }
  // IMU event on the predicted filter: joint H = I update of both covariances and of every axis state
  __device__ __forceinline__ void imu_update(const LeanRec& r, const real* u) {
    const real dt = r.dt;
    {  // 3 x 3: P~+ = P~ (P~ + I)^-1 through the cofactors of S~
      const real q00 = l00 + real(1), q11 = l11 + real(1), q22 = l22 + real(1);
      const real C00 = fma(q11, q22, -(l12 * l12));
      const real C01 = fma(l02, l12, -(l01 * q22));
      const real C02 = fma(l01, l12, -(l02 * q11));
      const real C11 = fma(q00, q22, -(l02 * l02));
      const real C12 = fma(l01, l02, -(q00 * l12));
      const real C22 = fma(q00, q11, -(l01 * l01));
      const real dS = fma(q00, C00, fma(l01, C01, l02 * C02));
      const real n0 = fma(l00, C00, fma(l01, C01, l02 * C02));
      const real n1 = fma(l01, C01, fma(l11, C11, l12 * C12));
      const real n2 = fma(l02, C02, fma(l12, C12, l22 * C22));
      const real inv = fast_rcp3(dS);
      const real ninv = -inv;
      l00 = n0 * inv, l11 = n1 * inv, l22 = n2 * inv;
      l01 = C01 * ninv, l02 = C02 * ninv, l12 = C12 * ninv;
    }
    {  // state: K = D P~+ D^-1; y = [(v- + u dt) dt, u dt, u - a] (kf_workers.py:698-709)
      const real g01 = l01 * s01, g02 = l02 * s02, g12 = l12 * s12;
      const real g10 = l01 * s10, g20 = l02 * s20, g21 = l12 * s21;
      lin_members([&](auto Bm, int m) {
        using BM = decltype(Bm);
        const real acc = u[BM::pz(0)];
        const real w = acc * dt;
        const real y0 = (xv[m] + w) * dt, y2 = acc - xa[m];
        xp[m] = fma(l00, y0, fma(g01, w, fma(g02, y2, xp[m])));
        xv[m] = fma(g10, y0, fma(l11, w, fma(g12, y2, xv[m])));
        xa[m] = fma(g20, y0, fma(g21, w, fma(l22, y2, xa[m])));
      });
    }
    {  // 2 x 2 (as Sym2Filter::step_lean)
      const real detp = fma(c00, c11, -(c01 * c01));
      const real u0 = detp + c00, u1 = detp + c11;
      const real inv = fast_rcp3((u0 + real(1)) + c11);
      c00 = u0 * inv;
      c11 = u1 * inv;
      c01 = c01 * inv;
      const real h01 = c01 * t01, h10 = c01 * t10;
      ang_members([&](auto Bm, int m) {
        using BM = decltype(Bm);
        const real y0 = u[BM::pz(0)] - th[m], y1 = u[BM::pz(1)] - om[m];
        th[m] = fma(c00, y0, fma(h01, y1, th[m]));
        om[m] = fma(h10, y0, fma(c11, y1, om[m]));
      });
    }
  }
  // GPS event: scalar update of the position on the linear chains (gam = r_gps / r0); the angular chains only predict
  __device__ __forceinline__ void gps_update(const LeanRec& r, const real* u) {
    const real inv = fast_rcp3(l00 + real(r.gam));
    const real k0 = l00 * inv, k1 = l01 * inv, k2 = l02 * inv, g = real(r.gam) * inv;
    const real K1 = k1 * s10, K2 = k2 * s20;
    lin_members([&](auto Bm, int m) {
      using BM = decltype(Bm);
      const real y = u[BM::gps(0)] - xp[m];
      xp[m] = fma(k0, y, xp[m]);
      xv[m] = fma(K1, y, xv[m]);
      xa[m] = fma(K2, y, xa[m]);
    });
    l11 = fma(-k1, l01, l11);
    l12 = fma(-k1, l02, l12);
    l22 = fma(-k2, l02, l22);
    l00 *= g;
    l01 *= g;
    l02 *= g;
  }
  // record-driven interface of the ring kernel (rec_api)
  __device__ __forceinline__ double update(const LeanRec& r, int ty, const real* u, bool& ok) {
    if (ty == EV_IMU) imu_update(r, u);
    else if (ty == EV_GPS) gps_update(r, u);
    return det(ok);
  }
  // fused step of the aggregate-only loop; v = (det lin~)^EL (det ang~)^EA after the step
  __device__ __forceinline__ void step_lean(const LeanRec& r, const ModelConst&, const real* u, real& v) {
    predict(r);
    if (r.type == EV_IMU) imu_update(r, u);
    else if (r.type == EV_GPS) gps_update(r, u);
    v = vnorm();
  }
};

// linear / angular chain lists of the models that have the SymCAFilter shape
template <int MODEL_ID>
struct CALists {
  using Lin = void;
  using Ang = void;
};

// --------------------------------------------------------------------------------------
// Model registry: decoupled chain list and dense union per model id
// --------------------------------------------------------------------------------------
template <int MODEL_ID, bool DENSE, typename real>
struct ModelTraits;

// index of the chain that owns global state g (for the block-compatibility check)
template <class... Bs>
struct ChainIndex {
  static constexpr int of(int g) {
    int r = -1, k = 0;
    ((r = (r < 0 && Union<Bs...>::template local_of<Bs>(g) >= 0) ? k : r, ++k), ...);
    return r;
  }
};

#define B200KF_DEFINE_MODEL(ID, N, D, MGPS, CTRL, ...)                         \
  template <typename real>                                                     \
  struct ModelTraits<ID, false, real> {                                        \
    static constexpr int n = N, d = D, m_gps = MGPS;                           \
    static constexpr bool imu_control = CTRL;                                  \
    using Filter = FilterT<real, CTRL, __VA_ARGS__>;                           \
    using Dec2 = Dec2Filter<real, __VA_ARGS__>;                                \
    using Sym2 = Sym2Filter<real, __VA_ARGS__>;                                \
    using Sym3 = typename CtrlSymSelect<real, __VA_ARGS__>::type;              \
    using SymCA = SymCAFilter<real, typename CALists<ID>::Lin, typename CALists<ID>::Ang>; \
    static constexpr int chain_of(int g) { return ChainIndex<__VA_ARGS__>::of(g); } \
  };                                                                           \
  template <typename real>                                                     \
  struct ModelTraits<ID, true, real> {                                         \
    static constexpr int n = N, d = D, m_gps = MGPS;                           \
    static constexpr bool imu_control = CTRL;                                  \
    using Filter = FilterT<real, CTRL, Union<__VA_ARGS__>>;                    \
  };

enum : int { MODEL_CV6 = 0, MODEL_CA8 = 1, MODEL_BIAS9 = 2, MODEL_CA15 = 3, MODEL_CVBU6 = 4, MODEL_COUNT = 5 };

template <>
struct CALists<MODEL_CA8> {
  using Lin = ChainList<CA8LinChain<0>, CA8LinChain<1>>;
  using Ang = ChainList<CA8YawChain>;
};
template <>
struct CALists<MODEL_CA15> {
  using Lin = ChainList<CA15LinChain<0>, CA15LinChain<1>, CA15LinChain<2>>;
  using Ang = ChainList<CA15AngChain<0>, CA15AngChain<1>, CA15AngChain<2>>;
};

B200KF_DEFINE_MODEL(MODEL_CV6, 6, 3, 3, false, CV6Chain<0>, CV6Chain<1>, CV6Chain<2>)
B200KF_DEFINE_MODEL(MODEL_CA8, 8, 4, 2, false, CA8LinChain<0>, CA8LinChain<1>, CA8YawChain)
B200KF_DEFINE_MODEL(MODEL_BIAS9, 9, 3, 3, true, BIAS9Chain<0>, BIAS9Chain<1>, BIAS9Chain<2>)
B200KF_DEFINE_MODEL(MODEL_CVBU6, 6, 3, 3, true, CVBU6Chain<0>, CVBU6Chain<1>, CVBU6Chain<2>)
B200KF_DEFINE_MODEL(MODEL_CA15, 15, 9, 3, false, CA15LinChain<0>, CA15LinChain<1>, CA15LinChain<2>,
                    CA15AngChain<0>, CA15AngChain<1>, CA15AngChain<2>)

// Axis-symmetric form of each model: chains that are the same model on different axes grouped
// into SymBlocks (one covariance per group).
template <int MODEL_ID, typename real>
struct SymTraits;
template <typename real>
struct SymTraits<MODEL_CV6, real> {
  using Filter = FilterT<real, false, SymGroup<CV6Chain<0>, CV6Chain<1>, CV6Chain<2>>>;
};
template <typename real>
struct SymTraits<MODEL_CA8, real> {
  using Filter = FilterT<real, false, SymGroup<CA8LinChain<0>, CA8LinChain<1>>, CA8YawChain>;
};
template <typename real>
struct SymTraits<MODEL_BIAS9, real> {
  using Filter = FilterT<real, true, SymGroup<BIAS9Chain<0>, BIAS9Chain<1>, BIAS9Chain<2>>>;
};
template <typename real>
struct SymTraits<MODEL_CVBU6, real> {
  using Filter = FilterT<real, true, SymGroup<CVBU6Chain<0>, CVBU6Chain<1>, CVBU6Chain<2>>>;
};
template <typename real>
struct SymTraits<MODEL_CA15, real> {
  using Filter = FilterT<real, false, SymGroup<CA15LinChain<0>, CA15LinChain<1>, CA15LinChain<2>>,
                         SymGroup<CA15AngChain<0>, CA15AngChain<1>, CA15AngChain<2>>>;
};

// --------------------------------------------------------------------------------------
// Gain records: what a covariance recursion publishes per step for filters that only apply gains
// (shared-covariance kernels, precomputed gain schedules).  g holds, block after block of the model's
// axis-symmetric form, K = P+ H^T R^-1 of the event's measured columns (row-major [n][n_meas]).
// --------------------------------------------------------------------------------------
template <class Flt>
struct GainSlots {
  static constexpr int value = 0;
  static constexpr int cov = 0;
};
template <typename real, bool IC, class B0, class... Bs>
struct GainSlots<FilterT<real, IC, B0, Bs...>> {
  using Blk = typename BlockObj<B0, real>::type;
  static constexpr int value = Blk::kslots + GainSlots<FilterT<real, IC, Bs...>>::value;
  static constexpr int cov = Blk::NP + GainSlots<FilterT<real, IC, Bs...>>::cov;
};

constexpr int kMaxGainBlocks = 4;
// GainRecG::pad == kCovRecord: a COVARIANCE record -- g[0 .. cov) holds P+ of every block (upper triangle, block after
// block) instead of the gains, det / ldet are not written: the consumer forms K = P+ H^T R^-1 and det P itself.  The
// greedy selection's warp walk writes these (a record costs its serial walk 7 stores instead of ~80 instructions).
constexpr int kCovRecord = 1;
template <int NG>
struct alignas(16) GainRecG {    // a multiple of 16 bytes: records travel by bulk copy
  double dt;
  double det[kMaxGainBlocks];    // det of each block (with its axis multiplicity); unused entries are 1
  double ldet[kMaxGainBlocks];   // log of it (per-step log-det output only); unused entries are 0
  int type;
  int pad;
  double g[NG > 0 ? NG : 1];
};

// kernel variants: 0 = one covariance per chain, 1 = dense, 2 = one covariance per symmetric group
enum : int { VAR_DECOUPLED = 0, VAR_DENSE = 1, VAR_SYMMETRIC = 2 };
template <int MODEL_ID, int VAR, typename real>
struct VariantTraits : ModelTraits<MODEL_ID, VAR == VAR_DENSE, real> {};
template <int MODEL_ID, typename real>
struct VariantTraits<MODEL_ID, VAR_SYMMETRIC, real> : ModelTraits<MODEL_ID, false, real> {
  using Filter = typename SymTraits<MODEL_ID, real>::Filter;
};

// Running aggregate of log det P over the steps of one launch, kept in product form so the
// hot loop needs no log: sum log d_k = log(mantissa product) + exponent sum * ln 2.
struct LogDetAcc {
  double mant = 1.0;
  int expo = 0;
  double maxdet = 0.0;
  double last = 1.0;
  int count = 0;
  __device__ __forceinline__ void add(double det) {
    last = det;
    maxdet = fmax(maxdet, det);
    count += 1;
    double m = mant * det;
    // renormalise m to [1, 2): valid for positive normal m (det > 0 is checked by the caller)
    const int hi = __double2hiint(m);
    const int e = ((hi >> 20) & 0x7ff) - 1023;
    expo += e;
    mant = __hiloint2double(hi - (e << 20), __double2loint(m));
  }
  __device__ __forceinline__ double sum_log() const {
    return log(mant) + double(expo) * 0.6931471805599453094;
  }
};

}  // namespace b200kf
