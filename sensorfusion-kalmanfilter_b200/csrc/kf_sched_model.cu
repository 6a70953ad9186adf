// kf_sched_model.cu -- brute-force schedule scoring for ONE model id (-DB200KF_INST_MODEL=<id>).
//
// Thread <-> candidate schedule (a bitmask over the <= 32 events of a window).  The covariance
// recursion of a linear KF never reads a measurement (kf_workers.py:44,64,67,70 touch no z), so
// the kernel carries P only: no state, no payload, no global loads at all in the hot loop -- the
// window (types, times), P0 and the rate constants arrive as kernel parameters.  Per schedule it
// reproduces evaluate_combo_chunk_worker's log_det trace (kf_workers.py:22-97): initial term,
// one term per used event with gaps merged into one predict, and the end-of-window propagation;
// the trace is reduced on the fly to (max, sum, last, n_terms).  Block- then grid-level
// reductions give (a) argmin of the sum with lowest-index ties and (b) the reference's selection
// rule -- first schedule in (cardinality, lexicographic) order whose max stays below the
// threshold (:1325-1358).
#include "kf_internal.h"

#ifndef B200KF_INST_MODEL
#error "compile with -DB200KF_INST_MODEL=<model id>"
#endif

namespace b200kf {

constexpr int kScoreTpb = 128;

__device__ __forceinline__ unsigned long long lex_key(unsigned long long s, int L) {
  // Within one cardinality the lexicographically smaller index tuple has the LARGER bit-reversed
  // mask (event 0 -> most significant bit), hence the complement.
  const unsigned long long rev = __brevll(s) >> (64 - L);
  const unsigned long long full = (L >= 64) ? ~0ULL : ((1ULL << L) - 1ULL);
  return (static_cast<unsigned long long>(__popcll(s)) << 32) | (full - rev);
}

// The same key for the searches' 32-bit schedule indices (windows hold at most 32 events)
static_assert(kMaxEvents <= 32, "the search kernels carry schedule indices in 32 bits");
__device__ __forceinline__ unsigned long long lex_key32(unsigned s, int L) {
  const unsigned rev = __brev(s) >> (32 - L);
  const unsigned full = (L >= 32) ? ~0u : ((1u << L) - 1u);
  return (static_cast<unsigned long long>(__popc(s)) << 32) | (full - rev);
}

__device__ __forceinline__ void better_sum(double& sc, unsigned long long& ix, double osc,
                                           unsigned long long oix) {
  if (osc < sc || (osc == sc && oix < ix)) {
    sc = osc;
    ix = oix;
  }
}

// warp shuffle + shared-memory reduction of a block's results into partials[blockIdx.x]
__device__ __forceinline__ void score_block_reduce(double best_sc, unsigned long long best_ix,
                                                   unsigned long long best_key, unsigned long long n_feas,
                                                   unsigned long long n_near, ScorePartial* partials) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double osc = __shfl_down_sync(0xffffffffu, best_sc, off);
    const unsigned long long oix = __shfl_down_sync(0xffffffffu, best_ix, off);
    better_sum(best_sc, best_ix, osc, oix);
    const unsigned long long ok2 = __shfl_down_sync(0xffffffffu, best_key, off);
    best_key = ok2 < best_key ? ok2 : best_key;
    n_feas += __shfl_down_sync(0xffffffffu, n_feas, off);
    n_near += __shfl_down_sync(0xffffffffu, n_near, off);
  }
  __shared__ ScorePartial sh[kScoreTpb / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sh[warp] = ScorePartial{best_sc, best_ix, best_key, n_feas, n_near};
  __syncthreads();
  if (threadIdx.x == 0) {
    ScorePartial r = sh[0];
    for (int w = 1; w < kScoreTpb / 32; ++w) {
      better_sum(r.score, r.index, sh[w].score, sh[w].index);
      r.key = sh[w].key < r.key ? sh[w].key : r.key;
      r.n_feasible += sh[w].n_feasible;
      r.n_near += sh[w].n_near;
    }
    partials[blockIdx.x] = r;
  }
}

// ---- one scored schedule -> the thread's running results ----------------------------------------
// The decisions are taken on the determinants themselves, without a logarithm per schedule: feasibility
// max(log det) < threshold (kf_workers.py:1353) as maxdet < exp(threshold), the undecided-to-rounding zone
// |max - threshold| <= 1e-9 as exp(threshold -/+ tol) <= maxdet <= ..., both bounds formed once on the host
// (the two forms can only differ inside that zone, which the caller re-scores in the reference's arithmetic);
// the sum of log-dets is formed -- log(mantissa) + exponent * ln 2, as before -- only for a schedule whose running
// product (exponent, mantissa) is not above the thread's current best by more than 1e-9 relative, i.e. for the few
// schedules that can win or tie: same winner, same tie order (lowest index), same reported score.
struct ThreadBest {
  double sc = INFINITY;
  unsigned long long ix = ~0ULL, key = ~0ULL;
  unsigned n_feas = 0, n_near = 0;  // per thread: far below 2^32 schedules
  double q_limit = INFINITY;  // exponent + mantissa of the best so far, plus the slack
};

__device__ __forceinline__ void score_leaf(const ScoreParams& p, const LogDetAcc& acc, bool ok, bool host_ld0,
                                           unsigned s, unsigned j, int L, ThreadBest& tb) {
  if (p.per_schedule != nullptr) {
    double mx = (acc.count > 0) ? log(acc.maxdet) : -INFINITY;
    double sm = acc.sum_log();
    double lastv = (acc.count > 0) ? log(acc.last) : p.initial_logdet;
    int cnt = acc.count;
    if (host_ld0) {  // term 0 is the same for every schedule: take the caller's value verbatim
      mx = fmax(mx, p.initial_logdet);
      sm += p.initial_logdet;
      cnt += 1;
    }
    if (!ok) mx = sm = lastv = NAN;
    double* o = p.per_schedule + (static_cast<unsigned long long>(j) - p.out_base) * 4;
    o[0] = mx;
    o[1] = sm;
    o[2] = lastv;
    o[3] = static_cast<double>(cnt);
  }
  if (s != 0 && ok) {  // the reference never enumerates the empty set (:1325)
    const double q = static_cast<double>(acc.expo) + acc.mant;
    if (q <= tb.q_limit) {
      double sm = acc.sum_log();
      if (host_ld0) sm += p.initial_logdet;
      if (sm < tb.sc || (sm == tb.sc && s < tb.ix)) {
        tb.sc = sm;
        tb.ix = s;
        tb.q_limit = q + p.q_slack * (fabs(q) + 1.0);
      }
    }
    // max over the terms computed HERE (the ones a rounding difference can move); the caller's exact term 0
    // decides on its own (it does so in the reference's arithmetic too)
    const bool ld0_ok = !(host_ld0 && p.initial_logdet >= p.threshold);
    const bool dev_ok = (acc.count > 0) ? (acc.maxdet < p.det_thr) : (-INFINITY < p.threshold);
    if (dev_ok && ld0_ok) {  // :1353
      tb.n_feas += 1;
      const unsigned long long key = lex_key32(s, L);
      tb.key = key < tb.key ? key : tb.key;
    }
    if (acc.count > 0 && acc.maxdet >= p.det_lo && acc.maxdet <= p.det_hi && ld0_ok) tb.n_near += 1;
  }
}

template <int MODEL, int VAR>
__global__ void __launch_bounds__(kScoreTpb) score_kernel(const ScoreParams p) {
  using MT = VariantTraits<MODEL, VAR, double>;
  constexpr int n = MT::n;

  ThreadBest tb;
  const int L = p.n_events;

  for (unsigned long long idx = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
       idx < p.count; idx += static_cast<unsigned long long>(gridDim.x) * blockDim.x) {
    const unsigned j = static_cast<unsigned>(p.first + idx);  // < 2^32: at most 32 events
    const unsigned s = static_cast<unsigned>(p.low_bits) | (p.n_low < 32 ? j << p.n_low : 0u);
    typename MT::Filter flt;
    flt.each([&](auto& b) { b.load_cov(p.P0, n); });
    bool ok = true;
    LogDetAcc acc;
    const bool host_ld0 = !isnan(p.initial_logdet);
    if (!host_ld0) acc.add(filter_det(flt, ok));  // log_det[0], kf_workers.py:32
    double tcur = p.t_prev;
    for (int i = 0; i < L; ++i) {
      if (!((s >> i) & 1u)) continue;
      const double dt = p.times[i] - tcur;  // merged gap, kf_workers.py:37
      if (dt < 0) continue;                 // :38-40
      const int ty = p.types[i];
      filter_predict<false>(flt, dt, p.mc, ty, static_cast<const double*>(nullptr));
      filter_update<false>(flt, dt, p.mc, ty, static_cast<const double*>(nullptr));
      tcur = p.times[i];
      acc.add(filter_det(flt, ok));
    }
    if (tcur < p.t_end - 1e-8) {  // kf_workers.py:74-82
      const double dt = p.t_end - tcur;
      filter_predict<false>(flt, dt, p.mc, EV_PREDICT_ONLY, static_cast<const double*>(nullptr));
      acc.add(filter_det(flt, ok));
    }
    score_leaf(p, acc, ok, host_ld0, s, j, L, tb);
  }

  score_block_reduce(tb.sc, tb.ix, tb.key, tb.n_feas, tb.n_near, static_cast<ScorePartial*>(p.partials));
}

// ---- prefix-sharing variant ----------------------------------------------------------------------
// Schedules that agree on the first events share the filter up to there.  For an aligned block
// [first, first + 2^a) the low a bits are free: thread <-> the low a-k bits (the EARLY events, walked
// once), then a depth-first walk over the next k bits with the covariance, the running log-det
// aggregate and the current time saved per level (local memory, L1-resident).  Leaves are visited
// in bit-reversed order so the earliest DFS event changes least often: going from one leaf to the
// next re-enters the tree at the level of the lowest set bit of the counter, which costs one filter
// step per leaf on average instead of one per used event (L/2).  Bits >= a (fixed by `first`) and
// the end-of-window propagation are applied per leaf.  Same arithmetic, order of operations and
// reductions as score_kernel, so the two agree bit for bit.  Strided enumeration (n_low > 0) fixes the
// n_low EARLIEST events instead: they are walked once per thread before the prefix, so a sharded search
// pays nothing per schedule for the bits that name the shard (fixed LATE events cost every leaf a step).
constexpr int kTreeMaxK = 8;

template <class Flt>
struct CovCount {
  static constexpr int value = 0;
};
template <typename real, bool IC>
struct CovCount<FilterT<real, IC>> {
  static constexpr int value = 0;
};
template <typename real, bool IC, class B0, class... Bs>
struct CovCount<FilterT<real, IC, B0, Bs...>> {
  static constexpr int value = BlockObj<B0, real>::type::NP + CovCount<FilterT<real, IC, Bs...>>::value;
};

template <int MODEL, int VAR>
__global__ void __launch_bounds__(kScoreTpb) score_tree_kernel(const ScoreParams p) {
  using MT = VariantTraits<MODEL, VAR, double>;
  using Flt = typename MT::Filter;
  constexpr int n = MT::n;
  constexpr int NC = CovCount<Flt>::value;

  ThreadBest tb;
  const int L = p.n_events;
  const int a = p.tree_a, k = p.tree_k;
  const int npre = a - k;  // events decided by the thread index
  const int nl = p.n_low;  // events fixed for the whole call (strided enumeration): walked once per thread, first
  const bool host_ld0 = !isnan(p.initial_logdet);

  struct Saved {
    double cov[NC];
    double mant, maxdet, last, tcur;
    int expo, count, ok;
  };
  Saved stk[kTreeMaxK + 1];

  Flt flt;
  LogDetAcc acc;
  bool ok;
  double tcur;
  // The deepest levels are the busy ones (level k-1 is restored by every other leaf, k-2 by every fourth): they live
  // in shared memory, [level][word][thread] -- no write-through to L2, no local-memory queue -- the rest in local memory.
#ifndef B200KF_TREE_SMEM_LEVELS
#define B200KF_TREE_SMEM_LEVELS 2
#endif
  constexpr int NS = (NC <= 16) ? B200KF_TREE_SMEM_LEVELS : 0;
  constexpr int NW = NC + 5;  // covariance entries, mant, maxdet, last, tcur, (expo, count | ok << 30)
  __shared__ double sst[NS > 0 ? NS : 1][NS > 0 ? NW : 1][NS > 0 ? kScoreTpb : 1];
  const int tix = threadIdx.x;
  auto save = [&](int lvl) {
    if (NS > 0 && lvl >= k - NS) {
      auto& sl = sst[lvl - (k - NS)];
      int o = 0;
      flt.each([&](const auto& b) {
        using BT = std::remove_cv_t<std::remove_reference_t<decltype(b)>>;
#pragma unroll
        for (int e = 0; e < BT::NP; ++e) sl[o + e][tix] = b.P[e];
        o += BT::NP;
      });
      sl[NC][tix] = acc.mant;
      sl[NC + 1][tix] = acc.maxdet;
      sl[NC + 2][tix] = acc.last;
      sl[NC + 3][tix] = tcur;
      sl[NC + 4][tix] = __hiloint2double(acc.expo, acc.count | (ok ? (1 << 30) : 0));
      return;
    }
    Saved& sv = stk[lvl];
    int o = 0;
    flt.each([&](const auto& b) {
      using BT = std::remove_cv_t<std::remove_reference_t<decltype(b)>>;
#pragma unroll
      for (int e = 0; e < BT::NP; ++e) sv.cov[o + e] = b.P[e];
      o += BT::NP;
    });
    sv.mant = acc.mant;
    sv.maxdet = acc.maxdet;
    sv.last = acc.last;
    sv.expo = acc.expo;
    sv.count = acc.count;
    sv.tcur = tcur;
    sv.ok = ok ? 1 : 0;
  };
  auto restore = [&](int lvl) {
    if (NS > 0 && lvl >= k - NS) {
      const auto& sl = sst[lvl - (k - NS)];
      int o = 0;
      flt.each([&](auto& b) {
        using BT = std::remove_cv_t<std::remove_reference_t<decltype(b)>>;
#pragma unroll
        for (int e = 0; e < BT::NP; ++e) b.P[e] = sl[o + e][tix];
        o += BT::NP;
      });
      acc.mant = sl[NC][tix];
      acc.maxdet = sl[NC + 1][tix];
      acc.last = sl[NC + 2][tix];
      tcur = sl[NC + 3][tix];
      const double w = sl[NC + 4][tix];
      acc.expo = __double2hiint(w);
      const int cw = __double2loint(w);
      acc.count = cw & ~(1 << 30);
      ok = (cw >> 30) & 1;
      return;
    }
    const Saved& sv = stk[lvl];
    int o = 0;
    flt.each([&](auto& b) {
      using BT = std::remove_cv_t<std::remove_reference_t<decltype(b)>>;
#pragma unroll
      for (int e = 0; e < BT::NP; ++e) b.P[e] = sv.cov[o + e];
      o += BT::NP;
    });
    acc.mant = sv.mant;
    acc.maxdet = sv.maxdet;
    acc.last = sv.last;
    acc.expo = sv.expo;
    acc.count = sv.count;
    tcur = sv.tcur;
    ok = sv.ok != 0;
  };
  auto apply = [&](int i) {  // use event i (kf_workers.py:36-71)
    const double dt = p.times[i] - tcur;  // merged gap, :37
    if (dt < 0) return;                   // :38-40
    const int ty = p.types[i];
    filter_predict<false>(flt, dt, p.mc, ty, static_cast<const double*>(nullptr));
    filter_update<false>(flt, dt, p.mc, ty, static_cast<const double*>(nullptr));
    tcur = p.times[i];
    acc.add(filter_det(flt, ok));
  };

  // 32-bit indices throughout: a window holds at most 32 events, an aligned tree block at most 2^32 schedules and
  // 2^(32 - 3) prefixes
  const unsigned n_prefix = 1u << npre;
  const unsigned low = static_cast<unsigned>(p.low_bits), first = static_cast<unsigned>(p.first);
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < n_prefix; t += gridDim.x * blockDim.x) {
    flt.each([&](auto& b) { b.load_cov(p.P0, n); });
    acc = LogDetAcc();
    ok = true;
    if (!host_ld0) acc.add(filter_det(flt, ok));  // log_det[0], kf_workers.py:32
    tcur = p.t_prev;
    for (int i = 0; i < nl; ++i)
      if ((low >> i) & 1u) apply(i);
    for (int i = 0; i < npre; ++i)
      if ((t >> i) & 1u) apply(nl + i);
    save(0);
    const unsigned n_leaf = 1u << k;
    for (unsigned c = 0; c < n_leaf; ++c) {
      const int lvl0 = (c == 0) ? 0 : (k - 1 - (__ffs(c) - 1));  // first level whose decision changed
      if (c != 0) restore(lvl0);
      for (int lv = lvl0; lv < k; ++lv) {
        if ((c >> (k - 1 - lv)) & 1u) apply(nl + npre + lv);
        if (lv + 1 < k) save(lv + 1);
      }
      const unsigned m = (k > 0) ? (__brev(c) >> (32 - k)) : 0u;
      const unsigned j = first | t | (m << npre);
      const unsigned s = low | (j << nl);
      for (int i = a + nl; i < L; ++i)
        if ((s >> i) & 1u) apply(i);
      if (tcur < p.t_end - 1e-8) {  // kf_workers.py:74-82
        const double dt = p.t_end - tcur;
        filter_predict<false>(flt, dt, p.mc, EV_PREDICT_ONLY, static_cast<const double*>(nullptr));
        acc.add(filter_det(flt, ok));
      }
      score_leaf(p, acc, ok, host_ld0, s, j, L, tb);
    }
  }
  score_block_reduce(tb.sc, tb.ix, tb.key, tb.n_feas, tb.n_near, static_cast<ScorePartial*>(p.partials));
}

// ---- worker with state: full filter per schedule over one shared window ----------------------
template <int MODEL, bool DENSE>
__global__ void __launch_bounds__(kScoreTpb) masked_run_kernel(const MaskedParams p) {
  using MT = ModelTraits<MODEL, DENSE, double>;
  constexpr int n = MT::n;
  constexpr int D = MT::d;
  const long long f = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (f >= p.n_schedules) return;
  const unsigned long long s = p.masks[f];
  const int L = p.n_events;
  const int rows = L + 2;
  const int tw = 1 + p.traj_ncomp;
  typename MT::Filter flt;
  flt.each([&](auto& b) {
    b.load_cov(p.P0, n);
    using BT = std::remove_reference_t<decltype(b)>;
    static_for<0, BT::n>([&](auto I) { b.x[I] = p.x0[BT::gidx(I)]; });
  });
  bool ok = true;
  int term = 0;
  auto emit = [&](double t) {
    if (p.logdet != nullptr) p.logdet[f * rows + term] = log(filter_det(flt, ok));
    if (p.traj != nullptr) {
      double* tr = p.traj + (f * rows + term) * tw;
      tr[0] = t;
      const int nc = p.traj_ncomp;
      flt.each([&](const auto& b) {
        using BT = std::remove_cv_t<std::remove_reference_t<decltype(b)>>;
        static_for<0, BT::n>([&](auto I) {
          constexpr int g = BT::gidx(I);
          if (g < nc) tr[1 + g] = b.x[I];
        });
      });
    }
    term += 1;
  };
  emit(p.t_prev);  // traj[0], log_det[0]  (kf_workers.py:31-32)
  double tcur = p.t_prev;
  for (int i = 0; i < L; ++i) {
    if (!((s >> i) & 1ULL)) continue;
    const double dt = p.times[i] - tcur;  // :37
    if (dt < 0) continue;                 // :38-40
    const int ty = p.types[i];
    double u[D];
#pragma unroll
    for (int c = 0; c < D; ++c) u[c] = p.payload[i * D + c];
    filter_predict<true>(flt, dt, p.mc, ty, u);
    filter_update<true>(flt, dt, p.mc, ty, u);
    tcur = p.times[i];
    emit(tcur);
  }
  if (tcur < p.t_end - 1e-8) {  // :74-82
    const double dt = p.t_end - tcur;
    filter_predict<true>(flt, dt, p.mc, EV_PREDICT_ONLY, static_cast<const double*>(nullptr));
    emit(p.t_end);
  }
  if (p.n_terms != nullptr) p.n_terms[f] = ok ? term : -term;
  flt.each([&](const auto& b) {
    using BT = std::remove_cv_t<std::remove_reference_t<decltype(b)>>;
    static_for<0, BT::n>([&](auto I) { p.x_final[f * n + BT::gidx(I)] = b.x[I]; });
  });
  if (p.P_final != nullptr) {
    double* Pg = p.P_final + f * (n * n);
    if constexpr (!DENSE) {
      for (int e = 0; e < n * n; ++e) Pg[e] = 0.0;
    }
    double xs[n];
    flt.each([&](const auto& b) { b.store(xs, Pg, n); });
  }
}

template <>
int launch_masked_m<B200KF_INST_MODEL>(const b200kf_model* m, const MaskedParams& p, unsigned flags,
                                       cudaStream_t st) {
  (void)m;
  using MT = ModelTraits<B200KF_INST_MODEL, false, double>;
  bool dense = (flags & B200KF_RUN_FORCE_DENSE) != 0;
  if (!dense && !(flags & B200KF_RUN_FORCE_DECOUPLED)) {
    for (int i = 0; i < MT::n; ++i)
      for (int j = 0; j < MT::n; ++j)
        if (MT::chain_of(i) != MT::chain_of(j) && p.P0[i * MT::n + j] != 0.0) dense = true;
  }
  const unsigned blocks = static_cast<unsigned>((p.n_schedules + kScoreTpb - 1) / kScoreTpb);
  if (dense)
    masked_run_kernel<B200KF_INST_MODEL, true><<<blocks, kScoreTpb, 0, st>>>(p);
  else
    masked_run_kernel<B200KF_INST_MODEL, false><<<blocks, kScoreTpb, 0, st>>>(p);
  B200KF_CUDA_TRY(cudaGetLastError());
  return B200KF_OK;
}

// ---- greedy one-per-window scheduling (kf_workers.py:877-947 + Scheduler :174-213) -------------
// Thread <-> instance (processing period, P0).  The selection never reads a measurement: gain()
// uses only row 1 of (H, R) -- state 0 for both sensors -- so per window there are two candidate
// gains, trace(P) - sum_i P[i][0]^2 / (P[0][0] + r_type), and greedy_schedule keeps the FIRST
// queue entry that attains the maximum (strict '>' at :205).  The thread walks the event list
// once, emitting the selected event index per window and advancing P through predict + update.
template <int MODEL, int VAR>
__global__ void select_greedy_kernel(const SelectParams p) {
  using MT = VariantTraits<MODEL, VAR, double>;
  constexpr bool DENSE = (VAR == VAR_DENSE);
  constexpr int n = MT::n;
  // Instances close their windows at different events, so instances sharing a warp execute the union of their
  // heavy iterations (12 sampling rates in one warp: a window closes at 99.7 % of the events).  With few
  // instances each gets a warp of its own (lane_stride = 32) and walks only its own windows.
  const int gt = blockIdx.x * blockDim.x + threadIdx.x;
  if (gt % p.lane_stride != 0) return;
  const int inst = gt / p.lane_stride;
  if (inst >= p.n_inst) return;
  typename MT::Filter flt;
  flt.each([&](auto& b) { b.load_cov(p.P0 + static_cast<long long>(inst) * n * n, n); });
  double ir_gps[kMaxN], ir_imu[kMaxN];  // 1 / r per measured column, block after block (gain records only)
  if constexpr (VAR == VAR_SYMMETRIC) {
    int og = 0, oi = 0;
    flt.each([&](const auto& b) {
      using BT = std::remove_cv_t<std::remove_reference_t<decltype(b)>>;
      BT::template inv_noise<true>(p.mc, ir_gps + og);
      og += BT::template n_meas<true>();
      if constexpr (!MT::imu_control) {
        BT::template inv_noise<false>(p.mc, ir_imu + oi);
        oi += BT::template n_meas<false>();
      }
    });
  }
  const double period = p.periods[inst];
  double prev = p.t_prev[inst];
  int* sel = p.sel + static_cast<long long>(inst) * p.max_sel;
  long long ns = 0;
  long long first_gps = -1, first_imu = -1, q_first = -1;  // queue summary
  double t_nx = (p.n_events > 0) ? __ldg(p.times) : 0.0;  // fetched one event ahead of the serial chain
  int ty_nx = (p.n_events > 0) ? __ldg(p.types) : 0;
  for (long long e = 0; e < p.n_events; ++e) {
    const double t = t_nx;
    const int ty = ty_nx;
    if (e + 1 < p.n_events) {
      t_nx = __ldg(p.times + e + 1);
      ty_nx = __ldg(p.types + e + 1);
    }
    if (t - prev < period) {  // :880-882
      if (q_first < 0) q_first = e;
      if (ty == EV_GPS) { if (first_gps < 0) first_gps = e; }
      else { if (first_imu < 0) first_imu = e; }
      continue;
    }
    if (q_first < 0) {  // empty queue: the triggering event is the only candidate (:884-886)
      q_first = e;
      if (ty == EV_GPS) first_gps = e; else first_imu = e;
    }
    // gains of the two sensor types for the current P (Scheduler.gain, :174-185)
    double tr = 0.0, col2 = 0.0, p00 = 0.0;
    flt.each([&](const auto& b) {
      using BT = std::remove_cv_t<std::remove_reference_t<decltype(b)>>;
      using TB = typename BT::Traits;  // the chain tables (a symmetric group's first member)
      static_for<0, BT::n>([&](auto I) {
        constexpr int i = I;
        tr += static_cast<double>(BT::M) * b.P[BT::pidx(i, i)];  // every member of a group carries this block
        static_for<0, BT::n>([&](auto J) {
          constexpr int j = J;
          if constexpr (TB::gidx(j) == 0) {  // column of global state 0: only the chain that owns it
            col2 += b.P[BT::pidx(i, j)] * b.P[BT::pidx(i, j)];
            if constexpr (i == j) p00 = b.P[BT::pidx(i, i)];
          }
        });
      });
    });
    const double gain_gps = tr - col2 / (p00 + p.mc.r_gps[0]);
    const double gain_imu = tr - col2 / (p00 + p.mc.r_imu[0]);
    long long pick;
    if (first_gps < 0) pick = first_imu;
    else if (first_imu < 0) pick = first_gps;
    else {
      // both present: the first entry attaining the max; on equal gains the earlier entry
      const long long a = first_gps < first_imu ? first_gps : first_imu;
      const double ga = (a == first_gps) ? gain_gps : gain_imu;
      const double gb = (a == first_gps) ? gain_imu : gain_gps;
      const long long bsel = (a == first_gps) ? first_imu : first_gps;
      pick = (gb > ga) ? bsel : a;
    }
    const double ts = p.times[pick];
    const int tys = p.types[pick];
    const double dt = ts - prev;  // :914
    filter_predict<false>(flt, dt, p.mc, tys, static_cast<const double*>(nullptr));
    filter_update<false, true>(flt, dt, p.mc, tys, static_cast<const double*>(nullptr));
    prev = ts;
    if (ns < p.max_sel) sel[ns] = static_cast<int>(pick);
    if constexpr (VAR == VAR_SYMMETRIC) {
      // the selection walks the very covariance recursion the selected events' filter needs: publish its gains
      // (K = P+ H^T R^-1 per block, det P) so that filter can run as a pure stream (b200kf_run_gains)
      if (p.gain_recs != nullptr && ns < p.max_sel) {
        using Rec = GainRecG<GainSlots<typename MT::Filter>::value>;
        Rec* rec = static_cast<Rec*>(p.gain_recs) + static_cast<long long>(inst) * p.max_sel + ns;
        int o = 0, oc = 0;
        flt.each([&](const auto& b) {
          using BT = std::remove_cv_t<std::remove_reference_t<decltype(b)>>;
          if (tys == EV_GPS) {
            b.template extract_gain<true>(ir_gps + oc, rec->g + o);
            oc += BT::template n_meas<true>();
          } else if (tys == EV_IMU) {
            if constexpr (!MT::imu_control) b.template extract_gain<false>(ir_imu + oc, rec->g + o);
            oc += BT::template n_meas<false>();
          }
          o += BT::kslots;
        });
        bool okd = true;
        const double det = filter_det(flt, okd);
        rec->dt = dt;
        rec->type = tys;
        rec->pad = 0;
#pragma unroll
        for (int q = 0; q < kMaxGainBlocks; ++q) {
          rec->det[q] = (q == 0) ? det : 1.0;
          rec->ldet[q] = 0.0;  // the consumer takes log(det) itself when a per-step log det is asked for
        }
      }
    }
    ns += 1;
    first_gps = first_imu = q_first = -1;
  }
  p.n_sel[inst] = ns;
  if (p.P_final != nullptr) {
    double* Pg = p.P_final + static_cast<long long>(inst) * n * n;
    if constexpr (!DENSE) {
      for (int e = 0; e < n * n; ++e) Pg[e] = 0.0;
    }
    double xs[n];
    flt.each([&](const auto& b) { b.store(xs, Pg, n); });
  }
}

// 16-byte global store under a predicate, without a branch (see emit_pending below)
__device__ __forceinline__ void st_pred(bool pred, double2* addr, double a, double b) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %0, 0;\n\t@p st.global.v2.f64 [%1], {%2, %3};\n\t}"
      :
      : "r"(static_cast<int>(pred)), "l"(addr), "d"(a), "d"(b)
      : "memory");
}

// Warp-per-instance form of the same walk (few instances: one per sampling rate in the scheduled sweep).  The
// thread-per-instance loop above spends ~0.25 us per EVENT on a dependent global load of (time, type) -- 590 000
// events of the 2 808 s stream = 150 ms for every rate, however few windows it has.  Here the warp loads 32 events
// per iteration (coalesced, one batch ahead), finds the event that closes the window and the queue's first GPS /
// IMU entries with ballots, and keeps the candidates' times in registers (no load on the decision path); all
// lanes run the covariance recursion redundantly (warp-uniform, no divergence), lane 0 writes the outputs.
// Same arithmetic, same order: selections and covariances are bit-identical to the thread-per-instance kernel.
template <int MODEL, int VAR, bool RECS>
__global__ void __launch_bounds__(32) select_greedy_warp_kernel(const SelectParams p) {
  using MT = VariantTraits<MODEL, VAR, double>;
  constexpr bool DENSE = (VAR == VAR_DENSE);
  constexpr int n = MT::n;
  constexpr unsigned kAll = 0xffffffffu;
  const int inst = blockIdx.x;
  const int lane = threadIdx.x;
  if (inst >= p.n_inst) return;
  typename MT::Filter flt;
  double* cr = (p.carry != nullptr) ? p.carry + static_cast<long long>(inst) * kSelectCarry : nullptr;
  const bool resume = cr != nullptr && cr[7] != 0.0;  // a later chunk of a chunked walk: state from the carry
  flt.each([&](auto& b) { b.load_cov((resume ? p.P_final : p.P0) + static_cast<long long>(inst) * n * n, n); });
  double ir_gps[kMaxN], ir_imu[kMaxN];
  if constexpr (VAR == VAR_SYMMETRIC) {
    int og = 0, oi = 0;
    flt.each([&](const auto& b) {
      using BT = std::remove_cv_t<std::remove_reference_t<decltype(b)>>;
      BT::template inv_noise<true>(p.mc, ir_gps + og);
      og += BT::template n_meas<true>();
      if constexpr (!MT::imu_control) {
        BT::template inv_noise<false>(p.mc, ir_imu + oi);
        oi += BT::template n_meas<false>();
      }
    });
  }
  const double period = p.periods[inst];
  double prev = p.t_prev[inst];
  int* sel = p.sel + static_cast<long long>(inst) * p.max_sel;
  const long long E = p.n_events;
  long long ns = 0;
  long long first_gps = -1, first_imu = -1, q_first = -1;  // queue summary
  double t_gps = 0.0, t_imu = 0.0;                         // times of those candidates
  int ty_imu = EV_IMU;                                     // type of the first non-GPS candidate
  if (resume) {
    prev = cr[0];
    q_first = static_cast<long long>(cr[1]), first_gps = static_cast<long long>(cr[2]);
    first_imu = static_cast<long long>(cr[3]);
    t_gps = cr[4], t_imu = cr[5];
    ty_imu = static_cast<int>(cr[6]);
    ns = p.n_sel[inst];
  }
  // Deferred COVARIANCE record (VAR_SYMMETRIC, kCovRecord): (dt, type, P+ of every block) of a selected event, cut
  // from the covariance that window left in `flt` at the start of the NEXT window (or after the loop), in the same
  // straight-line block as that window's decision and predict.  The consumer (b200kf_run_gains) forms the gains
  // K = P+ H^T R^-1 and det P+ itself: a record costs this serial walk seven predicated warp-uniform 16-byte stores
  // (same address, same value from every lane: one write) instead of ~80 instructions -- one warp issues about one
  // instruction per three cycles here whatever their dependencies, so the walk's time is its instruction count.
  long long pend_idx = -1;
  double pend_dt = 0.0;
  int pend_ty = 0;
  auto emit_pending = [&]() {
    if constexpr (VAR == VAR_SYMMETRIC && RECS) {
      using Rec = GainRecG<GainSlots<typename MT::Filter>::value>;
      constexpr int NG = GainSlots<typename MT::Filter>::value;
      constexpr int NCV = GainSlots<typename MT::Filter>::cov;
      static_assert(sizeof(Rec) % 16 == 0 && offsetof(Rec, g) == 80 && offsetof(Rec, type) == 72, "record layout");
      // (control-input models -- covariance larger than their gain slots -- never reach the selection: rejected by the API)
      const bool valid = pend_idx >= 0 && NCV <= NG;
      double g[NCV + 1];
      g[NCV] = 0.0;
      int oc = 0;
      flt.each([&](const auto& b) {
        using BT = std::remove_cv_t<std::remove_reference_t<decltype(b)>>;
        static_for<0, BT::NP>([&](auto I) { g[oc + I] = double(b.P[I]); });
        oc += BT::NP;
      });
      double2* out = reinterpret_cast<double2*>(static_cast<Rec*>(p.gain_recs) + static_cast<long long>(inst) * p.max_sel +
                                                (valid ? pend_idx : 0));
      st_pred(valid, out + 0, pend_dt, 0.0);                                  // dt, (det[0] unused)
      st_pred(valid, out + 4, 0.0, __hiloint2double(kCovRecord, pend_ty));    // (ldet[3] unused), (type, pad = kCovRecord)
#pragma unroll
      for (int q = 0; q < (NCV + 1) / 2; ++q) st_pred(valid, out + 5 + q, g[2 * q], g[2 * q + 1]);
      pend_idx = -1;
    }
  };
  const long long eb = p.e_begin;
  double t_nx = (eb + lane < E) ? __ldg(p.times + eb + lane) : 0.0;
  int ty_nx = (eb + lane < E) ? __ldg(p.types + eb + lane) : -1;
  for (long long e0 = eb; e0 < E; e0 += 32) {
    const double t_l = t_nx;
    const int ty_l = ty_nx;
    if (e0 + 32 + lane < E) {  // next batch, off the serial chain
      t_nx = __ldg(p.times + e0 + 32 + lane);
      ty_nx = __ldg(p.types + e0 + 32 + lane);
    }
    const unsigned valid = __ballot_sync(kAll, e0 + lane < E);
    const unsigned gps_m = __ballot_sync(kAll, ty_l == EV_GPS) & valid;
    int pos = 0;  // first lane of the batch not yet consumed
    while (true) {
      const unsigned from = (pos >= 32) ? 0u : (kAll << pos);
      const unsigned closes = __ballot_sync(kAll, !(t_l - prev < period)) & valid & from;  // :880-882
      const int c = closes ? (__ffs(closes) - 1) : 32;
      const unsigned range = valid & from & ((c >= 32) ? kAll : ((1u << c) - 1u));  // queued before the close
      if (range) {
        if (q_first < 0) q_first = e0 + __ffs(range) - 1;
        const unsigned g = range & gps_m, im = range & ~gps_m;
        if (first_gps < 0 && g) {
          const int l = __ffs(g) - 1;
          first_gps = e0 + l;
          t_gps = __shfl_sync(kAll, t_l, l);
        }
        if (first_imu < 0 && im) {
          const int l = __ffs(im) - 1;
          first_imu = e0 + l;
          t_imu = __shfl_sync(kAll, t_l, l);
          ty_imu = __shfl_sync(kAll, ty_l, l);
        }
      }
      if (c >= 32) break;  // batch exhausted, the queue carries over
      if (q_first < 0) {   // empty queue: the triggering event is the only candidate (:884-886); else it is dropped
        q_first = e0 + c;
        const double tc = __shfl_sync(kAll, t_l, c);
        const int tyc = __shfl_sync(kAll, ty_l, c);
        if (tyc == EV_GPS) {
          first_gps = q_first;
          t_gps = tc;
        } else {
          first_imu = q_first;
          t_imu = tc;
          ty_imu = tyc;
        }
      }
      // gains of the two sensor types for the current P (Scheduler.gain, :174-185)
      double tr = 0.0, col2 = 0.0, p00 = 0.0;
      flt.each([&](const auto& b) {
        using BT = std::remove_cv_t<std::remove_reference_t<decltype(b)>>;
        using TB = typename BT::Traits;
        static_for<0, BT::n>([&](auto I) {
          constexpr int i = I;
          tr += static_cast<double>(BT::M) * b.P[BT::pidx(i, i)];
          static_for<0, BT::n>([&](auto J) {
            constexpr int j = J;
            if constexpr (TB::gidx(j) == 0) {
              col2 += b.P[BT::pidx(i, j)] * b.P[BT::pidx(i, j)];
              if constexpr (i == j) p00 = b.P[BT::pidx(i, i)];
            }
          });
        });
      });
      bool pick_gps;
      if (first_gps < 0) pick_gps = false;
      else if (first_imu < 0) pick_gps = true;
      else {
        const double gain_gps = tr - col2 / (p00 + p.mc.r_gps[0]);
        const double gain_imu = tr - col2 / (p00 + p.mc.r_imu[0]);
        // both present: the first entry attaining the max; on equal gains the earlier entry (strict '>' at :205)
        const bool gps_first = first_gps < first_imu;
        const double ga = gps_first ? gain_gps : gain_imu;
        const double gb = gps_first ? gain_imu : gain_gps;
        pick_gps = (gb > ga) ? !gps_first : gps_first;
      }
      const long long pick = pick_gps ? first_gps : first_imu;
      const double ts = pick_gps ? t_gps : t_imu;
      const int tys = pick_gps ? static_cast<int>(EV_GPS) : ty_imu;
      const double dt = ts - prev;  // :914
      // the gain record of the PREVIOUS window is cut from the covariance it left behind, here, in the same
      // straight-line block as this window's decision and predict: its ~45 independent instructions fill the
      // issue slots the serial chain leaves empty instead of extending the chain
      if constexpr (VAR == VAR_SYMMETRIC) emit_pending();
      filter_predict<false>(flt, dt, p.mc, tys, static_cast<const double*>(nullptr));
      filter_update<false, true>(flt, dt, p.mc, tys, static_cast<const double*>(nullptr));
      prev = ts;
      if (lane == 0 && ns < p.max_sel) sel[ns] = static_cast<int>(pick);
      if constexpr (VAR == VAR_SYMMETRIC && RECS) {
        if (ns < p.max_sel) {
          pend_idx = ns;
          pend_dt = dt;
          pend_ty = tys;
        }
      }
      ns += 1;
      first_gps = first_imu = q_first = -1;
      pos = c + 1;
    }
  }
  if constexpr (VAR == VAR_SYMMETRIC) emit_pending();  // the last window's record
  if (lane != 0) return;
  p.n_sel[inst] = ns;
  if (cr != nullptr) {  // the walk's state at the end of this chunk (the open queue carries over)
    cr[0] = prev;
    cr[1] = static_cast<double>(q_first), cr[2] = static_cast<double>(first_gps), cr[3] = static_cast<double>(first_imu);
    cr[4] = t_gps, cr[5] = t_imu;
    cr[6] = static_cast<double>(ty_imu);
    cr[7] = 1.0;
  }
  if (p.P_final != nullptr) {
    double* Pg = p.P_final + static_cast<long long>(inst) * n * n;
    if constexpr (!DENSE) {
      for (int e = 0; e < n * n; ++e) Pg[e] = 0.0;
    }
    double xs[n];
    flt.each([&](const auto& b) { b.store(xs, Pg, n); });
  }
}

template <>
int launch_select_m<B200KF_INST_MODEL>(const b200kf_model* m, const SelectParams& p_in, unsigned flags,
                                       cudaStream_t st) {
  using SymF = typename SymTraits<B200KF_INST_MODEL, double>::Filter;
  const int tpb = 32;
  SelectParams p = p_in;
  p.lane_stride = (p.n_inst <= 2048) ? 32 : 1;
  const unsigned blocks = static_cast<unsigned>((static_cast<long long>(p.n_inst) * p.lane_stride + tpb - 1) / tpb);
  // P0 lives on the device, so the caller vouches for its structure: FORCE_SYMMETRIC = chain-compatible with
  // identical blocks across the axes of a group (one covariance per group), FORCE_DECOUPLED = chain-compatible;
  // without a flag the dense form, which is always valid
  const bool sym_ok = (flags & B200KF_RUN_FORCE_SYMMETRIC) && m->use_sym && SymF::constants_equal(m->mc);
  if (p.gain_recs != nullptr && !sym_ok)
    return set_error(B200KF_ERR_UNSUPPORTED, "gain records need B200KF_RUN_FORCE_SYMMETRIC (an axis-symmetric P0 and equal rates)");
  if (p.carry != nullptr && !(p.lane_stride == 32 && m->use_select_warp))
    return set_error(B200KF_ERR_UNSUPPORTED, "chunked selection runs on the warp-per-instance kernel (<= 2048 instances)");
  if (p.lane_stride == 32 && m->use_select_warp) {  // few instances: a warp each, events loaded 32 at a time
    const unsigned wb = static_cast<unsigned>(p.n_inst);
    if (sym_ok && p.gain_recs != nullptr)
      select_greedy_warp_kernel<B200KF_INST_MODEL, VAR_SYMMETRIC, true><<<wb, 32, 0, st>>>(p);
    else if (sym_ok)
      select_greedy_warp_kernel<B200KF_INST_MODEL, VAR_SYMMETRIC, false><<<wb, 32, 0, st>>>(p);
    else if (flags & (B200KF_RUN_FORCE_DECOUPLED | B200KF_RUN_FORCE_SYMMETRIC))
      select_greedy_warp_kernel<B200KF_INST_MODEL, VAR_DECOUPLED, false><<<wb, 32, 0, st>>>(p);
    else
      select_greedy_warp_kernel<B200KF_INST_MODEL, VAR_DENSE, false><<<wb, 32, 0, st>>>(p);
    B200KF_CUDA_TRY(cudaGetLastError());
    return B200KF_OK;
  }
  if (sym_ok)
    select_greedy_kernel<B200KF_INST_MODEL, VAR_SYMMETRIC><<<blocks, tpb, 0, st>>>(p);
  else if (flags & (B200KF_RUN_FORCE_DECOUPLED | B200KF_RUN_FORCE_SYMMETRIC))
    select_greedy_kernel<B200KF_INST_MODEL, VAR_DECOUPLED><<<blocks, tpb, 0, st>>>(p);
  else
    select_greedy_kernel<B200KF_INST_MODEL, VAR_DENSE><<<blocks, tpb, 0, st>>>(p);
  B200KF_CUDA_TRY(cudaGetLastError());
  return B200KF_OK;
}

template <int VAR>
static void launch_score_var(const ScoreParams& p, bool tree, unsigned blocks, cudaStream_t st) {
  if (tree)
    score_tree_kernel<B200KF_INST_MODEL, VAR><<<blocks, kScoreTpb, 0, st>>>(p);
  else
    score_kernel<B200KF_INST_MODEL, VAR><<<blocks, kScoreTpb, 0, st>>>(p);
}

// The range [first, first + count) is cut into aligned power-of-two blocks: blocks of at least
// 2^kTreeMinA schedules go to the prefix-sharing kernel, the ragged head and tail to the flat one.
// Every launch appends its per-block partials to the same array; returns their total number.
constexpr int kTreeMinA = 14;
constexpr int kMaxFlatBlocks = 2048;  // grid-stride beyond that
constexpr int kMaxTreeBlocks = 1184;  // upper bound (148 SMs x 8); a launch takes ONE resident wave (tree_wave_blocks), fewer when a call needs many tree launches

// Blocks of the prefix-sharing kernel that are resident at once on this device (6 per SM at 78 registers; forcing 72
// registers for 7 measured 6 % slower).  The kernel grid-strides over its threads, so a grid of exactly one wave
// keeps every SM equally loaded: 1184 blocks on 888 slots ran a second, one-third-full wave.
template <int VAR>
static int tree_wave_blocks() {
  static int cached = 0;
  if (cached == 0) {
    int per_sm = 0, dev = 0, sms = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, score_tree_kernel<B200KF_INST_MODEL, VAR>, kScoreTpb, 0) != cudaSuccess ||
        cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || per_sm < 1 || sms < 1) {
      cudaGetLastError();
      return kMaxTreeBlocks;
    }
    const int wave = per_sm * sms;
    cached = wave < kMaxTreeBlocks ? wave : kMaxTreeBlocks;
  }
  return cached;
}

template <>
int launch_score_m<B200KF_INST_MODEL>(const b200kf_model* m, ScoreParams p, unsigned flags,
                                      b200kf_best* best, cudaStream_t st) {
  (void)best;
  // P0 is a host matrix, so the variant is chosen on the host: dense when P0 couples chains,
  // one covariance per symmetric group when the member blocks of P0 are identical and the rates
  // agree (the reference's P0 and any P it produces, kf_workers.py:651), else one per chain.
  using MT = ModelTraits<B200KF_INST_MODEL, false, double>;
  using SymF = typename SymTraits<B200KF_INST_MODEL, double>::Filter;
  int var = VAR_DECOUPLED;
  if (flags & B200KF_RUN_FORCE_DENSE) {
    var = VAR_DENSE;
  } else if (flags & B200KF_RUN_FORCE_DECOUPLED) {
    var = VAR_DECOUPLED;
  } else {
    bool coupled = false;
    for (int i = 0; i < MT::n; ++i)
      for (int j = 0; j < MT::n; ++j)
        if (MT::chain_of(i) != MT::chain_of(j) && p.P0[i * MT::n + j] != 0.0) coupled = true;
    if (coupled) var = VAR_DENSE;
    else if (m->use_sym && SymF::constants_equal(m->mc) && SymF::blocks_identical_host(p.P0, MT::n)) var = VAR_SYMMETRIC;
  }
  if ((flags & B200KF_RUN_FORCE_SYMMETRIC) && !(flags & B200KF_RUN_FORCE_DENSE)) var = VAR_SYMMETRIC;
  p.dense_flag = nullptr;
  p.out_base = p.first;
  const bool tree_ok = m->use_tree != 0 && var != VAR_DENSE;  // the dense covariance is too large for the DFS stack
  ScorePartial* parts = static_cast<ScorePartial*>(p.partials);
  const unsigned long long end = p.first + p.count;
  unsigned long long cur = p.first;
  int total_blocks = 0;
  bool overflow = false;
  unsigned long long tree_cap = kMaxTreeBlocks;
  auto launch = [&](unsigned long long lo, unsigned long long cnt, int a) {  // a >= 0: aligned tree block
    ScoreParams q = p;
    q.first = lo;
    q.count = cnt;
    q.partials = parts + total_blocks;
    unsigned long long threads = cnt;
    q.tree_a = q.tree_k = 0;
    if (a >= 0) {
      // threads = 2^(a - k): 2^16 up to a = 20 (config 4: 40.9 us against 43.0 with 2^17), 2^17 beyond (a = 22 / 24:
      // 98 / 307 us against 107 / 364 with 2^16; tools/gpu_round2_zza.sh), until k reaches its maximum at a = 25
      static const int env_shift = []() { const char* e = getenv("B200KF_TREE_SHIFT"); return e ? atoi(e) : 0; }();
      const int shift = env_shift > 0 ? env_shift : (a >= 21 ? 17 : 16);
      int k = a - shift;
      k = k < 3 ? 3 : (k > kTreeMaxK ? kTreeMaxK : k);
      q.tree_a = a;
      q.tree_k = k;
      threads = 1ULL << (a - k);
    }
    // (64-thread blocks for launches below one wave -- config 4: 6.9 blocks per SM instead of 3.5 of 128 threads --
    // measured no faster: 47.1 vs 45.6 us per 2^20 search, tools/gpu_round2_zq.sh)
    unsigned long long blocks = (threads + kScoreTpb - 1) / kScoreTpb;
    const unsigned long long cap = (a >= 0) ? tree_cap : kMaxFlatBlocks;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    const unsigned nb = static_cast<unsigned>(blocks);
    if (total_blocks + static_cast<int>(nb) > kMaxSearchPartials) {
      overflow = true;
      return;
    }
    if (var == VAR_DENSE) launch_score_var<VAR_DENSE>(q, false, nb, st);
    else if (var == VAR_SYMMETRIC) launch_score_var<VAR_SYMMETRIC>(q, a >= 0, nb, st);
    else launch_score_var<VAR_DECOUPLED>(q, a >= 0, nb, st);
    total_blocks += static_cast<int>(nb);
  };
  // number of tree launches of this call -> blocks per tree launch such that all partials fit in the scratch
  int n_tree = 0;
  if (tree_ok) {
    for (unsigned long long c2 = p.first; c2 < end;) {
      const int align = (c2 == 0) ? 63 : __builtin_ctzll(c2);
      int a2 = 63 - __builtin_clzll(end - c2);
      if (a2 > align) a2 = align;
      if (a2 >= kTreeMinA) ++n_tree;
      c2 += 1ULL << a2;
    }
  }
  if (n_tree > 0) {
    const int fit = (kMaxSearchPartials - 2 * kMaxFlatBlocks) / n_tree;
    const int wave = (var == VAR_SYMMETRIC) ? tree_wave_blocks<VAR_SYMMETRIC>() : tree_wave_blocks<VAR_DECOUPLED>();
    tree_cap = fit < wave ? fit : wave;
  }
  if (p.count == 0) {
    launch(p.first, 0, -1);  // one block of idle threads writes the neutral partial
  }
  unsigned long long flat_lo = cur;  // start of the pending flat (ragged) piece
  while (cur < end) {
    const int align = (cur == 0) ? 63 : __builtin_ctzll(cur);
    int a = 63 - __builtin_clzll(end - cur);
    if (a > align) a = align;
    if (tree_ok && a >= kTreeMinA) {
      if (cur > flat_lo) launch(flat_lo, cur - flat_lo, -1);
      launch(cur, 1ULL << a, a);
      cur += 1ULL << a;
      flat_lo = cur;
    } else {
      cur += 1ULL << a;
    }
  }
  if (end > flat_lo) launch(flat_lo, end - flat_lo, -1);
  if (overflow) return set_error(B200KF_ERR_UNSUPPORTED, "schedule range needs more partial results than the scratch holds");
  B200KF_CUDA_TRY(cudaGetLastError());
  return total_blocks;
}

// ---- product-of-queues brute force, RMSE objective (KF_SensorFusion.ipynb:804-954) -------------
// Thread <-> candidate sequence: one event out of each window queue, index in itertools.product
// order (last queue fastest, :873-875).  Every sequence runs the full filter WITH state from
// (x0, P0) -- first event dt = 0 (:898) -- and is scored by the RMSE of the Euclidean position
// error against the track interpolated at the event times (calculate_accuracy_metrics, :766-802;
// the interpolation depends on the event only, so `ref` holds it per flattened event).  Events are
// gathered from small L1-resident tables; the block/grid reduction keeps the FIRST minimum
// (strict '<' at :933) and the runner-up score.  Same arithmetic per sequence whatever the sharing.
__device__ __forceinline__ void product_merge(double& sc, unsigned long long& ix, double& ru, double osc,
                                              unsigned long long oix, double oru) {
  const bool take = osc < sc || (osc == sc && oix < ix);
  const double loser = take ? sc : osc;
  ru = fmin(fmin(ru, oru), loser);
  if (take) {
    sc = osc;
    ix = oix;
  }
}

template <int MODEL, int VAR>
__global__ void __launch_bounds__(kScoreTpb) product_kernel(const ProductParams p) {
  using MT = VariantTraits<MODEL, VAR, double>;
  using Flt = typename MT::Filter;
  constexpr int n = MT::n;
  constexpr int D = MT::d;
  const int W = p.n_windows;
  double best_sc = INFINITY, best_ru = INFINITY;
  unsigned long long best_ix = ~0ULL;

  // Prefix sharing: consecutive indices differ in the LAST queues (itertools.product order), so a
  // thread owns the q[W-2] * q[W-1] sequences that agree on windows 0..W-3, filters that prefix
  // once and branches over the last two windows from saved copies of (filter, squared error, time).
  const int qa = (W >= 2) ? p.q_len[W - 2] : 1;  // second-to-last queue (1 pseudo-entry when W == 1)
  const int qb = p.q_len[W - 1];
  const unsigned long long B = static_cast<unsigned long long>(qa) * qb;
  const unsigned long long end = p.first + p.count;
  const unsigned long long blk0 = p.first / B;
  const unsigned long long nblk = (p.count == 0) ? 0 : ((end - 1) / B - blk0 + 1);

  struct St {
    Flt flt;
    double sse, tprev;
  };
  auto step = [&](St& s, int w, int e) {
    const double t = __ldg(p.times + e);
    const double dt = (w == 0) ? 0.0 : t - s.tprev;  // :898
    const int ty = __ldg(p.types + e);
    double u[D];
#pragma unroll
    for (int j = 0; j < D; ++j) u[j] = __ldg(p.payload + static_cast<long long>(e) * D + j);
    filter_predict<true>(s.flt, dt, p.mc, ty, u);
    filter_update<true>(s.flt, dt, p.mc, ty, u);
    double d2 = 0.0;  // |x[0:3] - ref|^2, then norm and square as numpy does (:788-790)
    s.flt.each([&](const auto& b) {
      b.for_each_state([&](auto G, double v) {
        constexpr int g = G;
        if constexpr (g < 3) {
          const double d = v - __ldg(p.ref + static_cast<long long>(e) * 3 + g);
          d2 += d * d;
        }
      });
    });
    const double nrm = sqrt(d2);
    s.sse += nrm * nrm;
    s.tprev = t;
  };

  for (unsigned long long bi = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x; bi < nblk;
       bi += static_cast<unsigned long long>(gridDim.x) * blockDim.x) {
    const unsigned long long c0 = (blk0 + bi) * B;
    St s0;
    s0.flt.each([&](auto& b) {
      b.load_cov(p.P0, n);
      b.load_state(p.x0);
    });
    s0.sse = 0.0;
    s0.tprev = 0.0;
    unsigned long long rem = c0;
    for (int w = 0; w + 2 < W; ++w) {  // the shared prefix
      const unsigned long long digit = rem / p.stride[w];
      rem -= digit * p.stride[w];
      step(s0, w, p.q_off[w] + static_cast<int>(digit));
    }
    for (int da = 0; da < qa; ++da) {
      St s1 = s0;
      if (W >= 2) step(s1, W - 2, p.q_off[W - 2] + da);
      for (int db = 0; db < qb; ++db) {
        const unsigned long long c = c0 + static_cast<unsigned long long>(da) * qb + db;
        if (c < p.first || c >= end) continue;
        St s2 = s1;
        step(s2, W - 1, p.q_off[W - 1] + db);
        const double metric = sqrt(s2.sse / static_cast<double>(W));
        if (p.per_combo != nullptr) p.per_combo[c - p.first] = metric;
        if (!isnan(metric)) product_merge(best_sc, best_ix, best_ru, metric, c, INFINITY);  // NaN < best is false (:933)
      }
    }
  }

#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double osc = __shfl_down_sync(0xffffffffu, best_sc, off);
    const unsigned long long oix = __shfl_down_sync(0xffffffffu, best_ix, off);
    const double oru = __shfl_down_sync(0xffffffffu, best_ru, off);
    product_merge(best_sc, best_ix, best_ru, osc, oix, oru);
  }
  __shared__ ProductPartial sh[kScoreTpb / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sh[warp] = ProductPartial{best_sc, best_ix, best_ru};
  __syncthreads();
  if (threadIdx.x == 0) {
    ProductPartial r = sh[0];
    for (int w = 1; w < kScoreTpb / 32; ++w) product_merge(r.score, r.index, r.runner_up, sh[w].score, sh[w].index, sh[w].runner_up);
    static_cast<ProductPartial*>(p.partials)[blockIdx.x] = r;
  }
}

template <>
int launch_product_m<B200KF_INST_MODEL>(const b200kf_model* m, const ProductParams& p, unsigned flags,
                                        cudaStream_t st) {
  const int W = p.n_windows;
  const unsigned long long B = static_cast<unsigned long long>(W >= 2 ? p.q_len[W - 2] : 1) * p.q_len[W - 1];
  const unsigned long long nthreads = (p.count == 0) ? 0 : ((p.first + p.count - 1) / B - p.first / B + 1);
  unsigned long long blocks = (nthreads + kScoreTpb - 1) / kScoreTpb;
  if (blocks > static_cast<unsigned long long>(kMaxSearchBlocks)) blocks = kMaxSearchBlocks;
  if (blocks == 0) blocks = 1;
  using MT = ModelTraits<B200KF_INST_MODEL, false, double>;
  using SymF = typename SymTraits<B200KF_INST_MODEL, double>::Filter;
  int var = VAR_DECOUPLED;
  if (flags & B200KF_RUN_FORCE_DENSE) {
    var = VAR_DENSE;
  } else if (!(flags & B200KF_RUN_FORCE_DECOUPLED)) {
    bool coupled = false;
    for (int i = 0; i < MT::n; ++i)
      for (int j = 0; j < MT::n; ++j)
        if (MT::chain_of(i) != MT::chain_of(j) && p.P0[i * MT::n + j] != 0.0) coupled = true;
    if (coupled) var = VAR_DENSE;
    else if (m->use_sym && SymF::constants_equal(m->mc) && SymF::blocks_identical_host(p.P0, MT::n)) var = VAR_SYMMETRIC;
  }
  if ((flags & B200KF_RUN_FORCE_SYMMETRIC) && !(flags & B200KF_RUN_FORCE_DENSE)) var = VAR_SYMMETRIC;
  if (var == VAR_DENSE)
    product_kernel<B200KF_INST_MODEL, VAR_DENSE><<<static_cast<unsigned>(blocks), kScoreTpb, 0, st>>>(p);
  else if (var == VAR_SYMMETRIC)
    product_kernel<B200KF_INST_MODEL, VAR_SYMMETRIC><<<static_cast<unsigned>(blocks), kScoreTpb, 0, st>>>(p);
  else
    product_kernel<B200KF_INST_MODEL, VAR_DECOUPLED><<<static_cast<unsigned>(blocks), kScoreTpb, 0, st>>>(p);
  B200KF_CUDA_TRY(cudaGetLastError());
  return static_cast<int>(blocks);
}

}  // namespace b200kf
