// kf_axis.cuh -- dense covariance of a three-axis model split over THREE LANES per filter.
//
// An arbitrary dense P0 (the reference accepts any initial_pt, kf_workers.py:642-645) couples every state
// of the 15-state filter: 120 symmetric entries = 240 registers per thread, and the thread-per-filter
// dense kernel spills (~7 KB of local traffic per step).  The model itself is the same five-state table
// (p, v, a, angle, rate) on three axes, so the state is viewed as 3 x 5 and P as a 3 x 3 grid of 5 x 5
// blocks B[a][b]; lane a of a filter keeps B[a][a] (symmetric, 15 entries) and B[a][a+1] (25 entries) in
// registers -- 40 reals instead of 120, the same code on every lane:
//   predict   B[a][b] <- F5 B[a][b] F5^T (+ Q on the diagonal blocks): F never couples axes, no exchange
//   update    the event's measured local states t = 0..4 are processed as JOINT three-measurement stages
//             (t on all three axes at once; exact for selector H and diagonal R, like the scalar stages of
//             kf_core.cuh): S = P_mm + R is 3 x 3, inverted by cofactors on every lane; the three columns
//             of P the stage needs come from the neighbour lanes with 19 warp shuffles per lane
//   metric    det P+ = det P- * prod_t det R_t / det S_t; det P- from a block elimination (3 x 3 pivots)
//             on a copy, the same stage code without the measurement noise
// Ten filters per warp (30 lanes), no local memory.
#pragma once
#include "kf_core.cuh"

namespace b200kf {

// Local state table of ONE axis: the linear chain followed by the angular chain of axis 0; global state,
// payload column and GPS column of axis a are those of axis 0 plus a.
template <template <int> class Lin, template <int> class Ang>
struct AxisTable {
  using L = Lin<0>;
  using A = Ang<0>;
  static constexpr int nl = L::n;
  static constexpr int n = L::n + A::n;
  static constexpr int gidx0(int i) { return i < nl ? L::gidx(i) : A::gidx(i - nl); }
  static constexpr int l1(int i) { return i < nl ? L::l1(i) : (A::l1(i - nl) >= 0 ? A::l1(i - nl) + nl : -1); }
  static constexpr int l2(int i) { return i < nl ? L::l2(i) : (A::l2(i - nl) >= 0 ? A::l2(i - nl) + nl : -1); }
  static constexpr int s1(int i) { return i < nl ? L::s1(i) : A::s1(i - nl); }
  static constexpr int s2(int i) { return i < nl ? L::s2(i) : A::s2(i - nl); }
  static constexpr int gps0(int i) { return i < nl ? L::gps(i) : A::gps(i - nl); }
  static constexpr int zk(int i) { return i < nl ? L::zk(i) : A::zk(i - nl); }
  static constexpr int pz0(int i) { return i < nl ? L::pz(i) : A::pz(i - nl); }
  static constexpr bool axis_linear() {  // every index of axis 1 / 2 is the axis-0 index + 1 / + 2
    bool ok = true;
    for (int i = 0; i < Lin<0>::n; ++i) {
      ok = ok && Lin<1>::gidx(i) == Lin<0>::gidx(i) + 1 && Lin<2>::gidx(i) == Lin<0>::gidx(i) + 2;
      ok = ok && Lin<1>::pz(i) == Lin<0>::pz(i) + 1 && Lin<2>::pz(i) == Lin<0>::pz(i) + 2;
      ok = ok && (Lin<0>::gps(i) < 0 || (Lin<1>::gps(i) == Lin<0>::gps(i) + 1 && Lin<2>::gps(i) == Lin<0>::gps(i) + 2));
      ok = ok && Lin<0>::cu(i) == 0;
    }
    for (int i = 0; i < Ang<0>::n; ++i) {
      ok = ok && Ang<1>::gidx(i) == Ang<0>::gidx(i) + 1 && Ang<2>::gidx(i) == Ang<0>::gidx(i) + 2;
      ok = ok && Ang<1>::pz(i) == Ang<0>::pz(i) + 1 && Ang<2>::pz(i) == Ang<0>::pz(i) + 2;
      ok = ok && (Ang<0>::gps(i) < 0 || (Ang<1>::gps(i) == Ang<0>::gps(i) + 1 && Ang<2>::gps(i) == Ang<0>::gps(i) + 2));
      ok = ok && Ang<0>::cu(i) == 0;
    }
    return ok;
  }
  // distinct payload columns (axis 0) one lane reads per step: GPS columns first, then the IMU columns
  static constexpr int ncols() {
    int c = 0;
    for (int i = 0; i < n; ++i) c += first_use(i, true) ? 1 : 0;
    for (int i = 0; i < n; ++i) c += first_use(i, false) ? 1 : 0;
    return c;
  }
  static constexpr int col_of(int i, bool gps) { return gps ? gps0(i) : (zk(i) != ZK_NONE ? pz0(i) : -1); }
  static constexpr bool first_use(int i, bool gps) {
    const int c = col_of(i, gps);
    if (c < 0) return false;
    if (!gps)
      for (int k = 0; k < n; ++k)
        if (col_of(k, true) == c) return false;
    for (int k = 0; k < i; ++k)
      if (col_of(k, gps) == c) return false;
    return true;
  }
  static constexpr int col(int slot) {  // axis-0 payload column of slot
    int c = 0;
    for (int i = 0; i < n; ++i)
      if (first_use(i, true)) {
        if (c == slot) return col_of(i, true);
        ++c;
      }
    for (int i = 0; i < n; ++i)
      if (first_use(i, false)) {
        if (c == slot) return col_of(i, false);
        ++c;
      }
    return 0;
  }
  static constexpr int slot(int i, bool gps) {
    const int want = col_of(i, gps);
    for (int s = 0; s < ncols(); ++s)
      if (col(s) == want) return s;
    return 0;
  }
};

template <class AT, typename real>
struct AxisDense {
  static constexpr int n = AT::n;
  static constexpr int ND = n * (n + 1) / 2;
  static constexpr int NC = AT::ncols();
  real x[n];
  real D[ND];     // B[a][a], upper triangle
  real G[n * n];  // B[a][a+1], row-major: G[i*n+j] = P[(i, a)][(j, a+1)]
  real q[n], ri[n], rg;  // this axis' process / IMU / GPS variances
  int a, nx, pv;         // my axis; the lanes that hold the next / previous axis of this filter
  unsigned gmask;        // the lanes of this filter

  static constexpr int di(int i, int j) {
    return i <= j ? (i * n - i * (i - 1) / 2 + (j - i)) : (j * n - j * (j - 1) / 2 + (i - j));
  }
  // the whole warp steps together (the event type is warp-uniform; a per-filter gate is resolved by the caller)
  __device__ __forceinline__ real from_next(real v) const { return __shfl_sync(0xffffffffu, v, nx); }
  __device__ __forceinline__ real from_prev(real v) const { return __shfl_sync(0xffffffffu, v, pv); }

  __device__ __forceinline__ void init_lane(int lane, const ModelConst& mc) {
    const int grp = lane / 3;
    a = lane - 3 * grp;
    const int base = 3 * grp;
    nx = min(base + (a + 1) % 3, 31);
    pv = min(base + (a + 2) % 3, 31);
    gmask = (grp < 10) ? (7u << base) : (3u << 30);
    static_for<0, n>([&](auto I) {
      constexpr int g = AT::gidx0(I);
      q[I] = real(a == 0 ? mc.q[g] : (a == 1 ? mc.q[g + 1] : mc.q[g + 2]));
      ri[I] = real(a == 0 ? mc.r_imu[g] : (a == 1 ? mc.r_imu[g + 1] : mc.r_imu[g + 2]));
    });
    rg = real(0);
    static_for<0, n>([&](auto I) {
      if constexpr (AT::gps0(I) >= 0) {
        constexpr int c = AT::gps0(I);
        rg = real(a == 0 ? mc.r_gps[c] : (a == 1 ? mc.r_gps[c + 1] : mc.r_gps[c + 2]));
      }
    });
  }

  __device__ __forceinline__ void load(const real* __restrict__ xg, const real* __restrict__ Pg, int nt) {
    const int an = (a + 1) % 3;
    static_for<0, n>([&](auto I) {
      constexpr int i = I;
      const int gi = AT::gidx0(i) + a;
      x[i] = xg[gi];
      static_for<i, n>([&](auto J) {
        const int gj = AT::gidx0(J) + a;
        D[di(i, J)] = real(0.5) * (Pg[gi * nt + gj] + Pg[gj * nt + gi]);
      });
      static_for<0, n>([&](auto J) {
        const int gj = AT::gidx0(J) + an;
        G[i * n + J] = real(0.5) * (Pg[gi * nt + gj] + Pg[gj * nt + gi]);
      });
    });
  }
  __device__ __forceinline__ void store(real* __restrict__ xg, real* __restrict__ Pg, int nt) const {
    const int an = (a + 1) % 3;
    static_for<0, n>([&](auto I) {
      constexpr int i = I;
      const int gi = AT::gidx0(i) + a;
      xg[gi] = x[i];
      static_for<i, n>([&](auto J) {
        const int gj = AT::gidx0(J) + a;
        Pg[gi * nt + gj] = D[di(i, J)];
        Pg[gj * nt + gi] = D[di(i, J)];
      });
      static_for<0, n>([&](auto J) {
        const int gj = AT::gidx0(J) + an;
        Pg[gi * nt + gj] = G[i * n + J];
        Pg[gj * nt + gi] = G[i * n + J];
      });
    });
  }
  template <class F>
  __device__ __forceinline__ void for_each_cov(F&& fn) const {  // fn(global row, global col, value), both triangles
    const int an = (a + 1) % 3;
    static_for<0, n>([&](auto I) {
      constexpr int i = I;
      const int gi = AT::gidx0(i) + a;
      static_for<0, n>([&](auto J) {
        fn(gi, AT::gidx0(J) + a, D[di(i, J)]);
        fn(gi, AT::gidx0(J) + an, G[i * n + J]);
        fn(AT::gidx0(J) + an, gi, G[i * n + J]);
      });
    });
  }

  // x = F x, P = F P F^T + Q (kf_workers.py:546-549): rows then columns, links point forward
  __device__ __forceinline__ void predict(real dt, real c2dt2) {
    static_for<0, n>([&](auto I) {
      constexpr int i = I;
      constexpr int L1 = AT::l1(i), L2 = AT::l2(i);
      static_assert(L1 < 0 || L1 > i, "links must point forward");
      static_assert(L2 < 0 || L2 > i, "links must point forward");
      if constexpr (L1 >= 0 && L2 >= 0) {
        const real c1 = real(AT::s1(i)) * dt, c2 = real(AT::s2(i)) * c2dt2;
        x[i] += c1 * x[L1] + c2 * x[L2];
        D[di(i, i)] += real(2) * (c1 * D[di(L1, i)] + c2 * D[di(L2, i)]) + c1 * c1 * D[di(L1, L1)] +
                       real(2) * (c1 * c2) * D[di(L1, L2)] + c2 * c2 * D[di(L2, L2)];
        static_for<0, n>([&](auto J) {
          if constexpr (J != i) D[di(i, J)] += c1 * D[di(L1, J)] + c2 * D[di(L2, J)];
        });
        static_for<0, n>([&](auto J) { G[i * n + J] += c1 * G[L1 * n + J] + c2 * G[L2 * n + J]; });
      } else if constexpr (L1 >= 0) {
        const real c1 = real(AT::s1(i)) * dt;
        x[i] += c1 * x[L1];
        D[di(i, i)] += real(2) * (c1 * D[di(L1, i)]) + c1 * c1 * D[di(L1, L1)];
        static_for<0, n>([&](auto J) {
          if constexpr (J != i) D[di(i, J)] += c1 * D[di(L1, J)];
        });
        static_for<0, n>([&](auto J) { G[i * n + J] += c1 * G[L1 * n + J]; });
      }
    });
    static_for<0, n>([&](auto Jc) {  // the next axis runs the same table: column operations of G
      constexpr int j = Jc;
      constexpr int L1 = AT::l1(j), L2 = AT::l2(j);
      if constexpr (L1 >= 0 && L2 >= 0) {
        const real c1 = real(AT::s1(j)) * dt, c2 = real(AT::s2(j)) * c2dt2;
        static_for<0, n>([&](auto I) { G[I * n + j] += c1 * G[I * n + L1] + c2 * G[I * n + L2]; });
      } else if constexpr (L1 >= 0) {
        const real c1 = real(AT::s1(j)) * dt;
        static_for<0, n>([&](auto I) { G[I * n + j] += c1 * G[I * n + L1]; });
      }
    });
    static_for<0, n>([&](auto I) { D[di(I, I)] += q[I] * dt; });
  }

  // IMU pseudo-measurement from the PREDICTED state (kf_workers.py:698-706); u[] = this lane's payload slots
  __device__ __forceinline__ void build_imu_z(real dt, const real* u, real* z) const {
    static_for<0, n>([&](auto I) {
      constexpr int i = I;
      if constexpr (AT::zk(i) == ZK_POS) {
        const real V = x[AT::l1(i)] + u[AT::slot(i, false)] * dt;
        z[i] = x[i] + V * dt;
      } else if constexpr (AT::zk(i) == ZK_VEL) {
        z[i] = x[i] + u[AT::slot(i, false)] * dt;
      } else if constexpr (AT::zk(i) == ZK_DIRECT) {
        z[i] = u[AT::slot(i, false)];
      } else {
        z[i] = real(0);
      }
    });
  }

  // One joint stage on local state T of the three axes.
  //   ELIM = false: measurement update with variance r (this lane's axis) and value z; returns det S.
  //   ELIM = true : block elimination step on (Dd, Gd) = a copy: rows / columns > T only, no noise, no state;
  //                 returns the determinant of the 3 x 3 pivot block.
  template <int T, bool ELIM>
  __device__ __forceinline__ real stage(real* __restrict__ Dd, real* __restrict__ Gd, real r, real z) {
    constexpr int I0 = ELIM ? T + 1 : 0;
    // S in the lane-relative order (a, next, prev)
    const real s_aa = ELIM ? Dd[di(T, T)] : Dd[di(T, T)] + r;
    const real s_an = Gd[T * n + T];
    const real s_nn = from_next(s_aa), s_np = from_next(s_an);
    const real s_pp = from_prev(s_aa), s_pa = from_prev(s_an);
    const real c00 = s_nn * s_pp - s_np * s_np;
    const real c01 = s_np * s_pa - s_an * s_pp;
    const real c02 = s_an * s_np - s_nn * s_pa;
    const real det = s_aa * c00 + s_an * c01 + s_pa * c02;
    if constexpr (ELIM && T == n - 1) return det;
    const real c11 = s_aa * s_pp - s_pa * s_pa;
    const real c12 = s_an * s_pa - s_aa * s_np;
    const real c22 = s_aa * s_nn - s_an * s_an;
    const real ninv = -fast_rcp(det);  // Kn = -(C S^-1): the downdates below are plain FMA chains
    const real w00 = c00 * ninv, w01 = c01 * ninv, w02 = c02 * ninv, w11 = c11 * ninv, w12 = c12 * ninv, w22 = c22 * ninv;
    // columns (T, a), (T, next), (T, prev) of P: rows of my axis (ca) and of the next axis (cn)
    real ca[n][3], cn[n][3], Kn[n][3];
    static_for<I0, n>([&](auto I) {
      constexpr int i = I;
      ca[i][0] = Dd[di(i, T)];
      ca[i][1] = Gd[i * n + T];
      ca[i][2] = from_prev(Gd[T * n + i]);
      cn[i][0] = Gd[T * n + i];
      cn[i][1] = from_next(Dd[di(i, T)]);
      cn[i][2] = from_next(Gd[i * n + T]);
      Kn[i][0] = fma(ca[i][2], w02, fma(ca[i][1], w01, ca[i][0] * w00));
      Kn[i][1] = fma(ca[i][2], w12, fma(ca[i][1], w11, ca[i][0] * w01));
      Kn[i][2] = fma(ca[i][2], w22, fma(ca[i][1], w12, ca[i][0] * w02));
    });
    if constexpr (!ELIM) {
      const real y_a = x[T] - z;  // minus the innovation (Kn is minus the gain)
      const real y_n = from_next(y_a), y_p = from_prev(y_a);
      static_for<0, n>([&](auto I) { x[I] = fma(Kn[I][2], y_p, fma(Kn[I][1], y_n, fma(Kn[I][0], y_a, x[I]))); });
    }
    static_for<I0, n>([&](auto I) {
      constexpr int i = I;
      if constexpr (i != T) {
        static_for<i, n>([&](auto J) {
          constexpr int j = J;
          if constexpr (j != T)
            Dd[di(i, j)] = fma(Kn[i][2], ca[j][2], fma(Kn[i][1], ca[j][1], fma(Kn[i][0], ca[j][0], Dd[di(i, j)])));
        });
        static_for<I0, n>([&](auto J) {
          constexpr int j = J;
          if constexpr (j != T)
            Gd[i * n + j] = fma(Kn[i][2], cn[j][2], fma(Kn[i][1], cn[j][1], fma(Kn[i][0], cn[j][0], Gd[i * n + j])));
        });
      }
    });
    if constexpr (!ELIM) {
      // measured rows / columns: P+[m, :] = R S^-1 P[m, :] = R K^T, no cancelling subtraction
      const real mr = -r, mr_n = from_next(mr);
      static_for<0, n>([&](auto J) {
        constexpr int j = J;
        Dd[di(T, j)] = mr * Kn[j][0];
        Gd[j * n + T] = mr_n * Kn[j][1];
        if constexpr (j != T) Gd[T * n + j] = mr * fma(cn[j][2], w02, fma(cn[j][1], w01, cn[j][0] * w00));
      });
    }
    return det;
  }

  // det P by block elimination on a copy; ok is cleared on a non-positive pivot block
  __device__ __forceinline__ double det(bool& ok) {
    real Dc[ND], Gc[n * n];
    static_for<0, ND>([&](auto I) { Dc[I] = D[I]; });
    static_for<0, n * n>([&](auto I) { Gc[I] = G[I]; });
    double prod = 1.0;
    static_for<0, n>([&](auto Tc) {
      const real d3 = stage<Tc, true>(Dc, Gc, real(0), real(0));
      ok = ok && (d3 > real(0));
      prod *= double(d3);
    });
    return prod;
  }

  // measurement update of the event; returns prod_t det S_t (float64) of its stages
  template <bool GPS_EVENT>
  __device__ __forceinline__ double update(const real* z) {
    double sprod = 1.0;
    static_for<0, n>([&](auto Tc) {
      constexpr int t = Tc;
      if constexpr (GPS_EVENT ? (AT::gps0(t) >= 0) : (AT::zk(t) != ZK_NONE)) {
        const real d3 = stage<t, false>(D, G, GPS_EVENT ? rg : ri[t], z[t]);
        sprod *= double(d3);
      }
    });
    return sprod;
  }
};

// models with the three-axis shape
template <int MODEL_ID>
struct AxisTraits {
  static constexpr bool eligible = false;
  using Table = void;
};
template <>
struct AxisTraits<MODEL_CA15> {
  static constexpr bool eligible = true;
  using Table = AxisTable<CA15LinChain, CA15AngChain>;
  static_assert(Table::axis_linear(), "axis a = axis 0 + a");
  static_assert(3 * Table::n == 15, "three axes cover the state");
};

}  // namespace b200kf
