// kf_internal.h -- library-internal declarations shared by the translation units.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <cuda_runtime.h>

#include "../../include/b200kf.h"
#include "kf_core.cuh"

// b200kf_run_batch_host's persistent resources (created on first use, freed with the model)
struct HostWorkspace {
  std::mutex mu;  // b200kf_run_batch_host calls on one model serialise (they share the slab, streams and events)
  int ready;
  cudaStream_t s_copy, s_comp;
  cudaEvent_t ev_copied[2], ev_done[2];
  void* slab;
  size_t cap;
  // pinned staging for callers whose payload lives in pageable memory (plain numpy arrays): two chunk buffers
  // filled by a few host threads while the previous chunk's host->device copy is in flight
  void* h_stage[2];
  size_t h_cap;
};

struct b200kf_model {
  b200kf_model_desc desc;
  b200kf::ModelConst mc;
  int n, d, m_gps;
  int device;
  int sm_count;
  // Per-call device scratch (dispatch flags, packed schedules, reduction partials) comes from this
  // stream-ordered pool: allocated on the caller's stream, freed on it behind the last kernel that reads it.
  // Calls on different streams / threads therefore never share scratch and one model can be in flight on
  // many streams at once; the pool keeps what it ever held (release threshold = max), so a steady-state
  // call costs two ~1 us driver calls and no device synchronisation.
  cudaMemPool_t pool;
  int tpb;      // threads per block for the filter kernels (env B200KF_TPB)
  HostWorkspace host_ws;
  int use_sym;  // 1: axis-symmetric kernel when every filter's chains share one covariance (env B200KF_SYM=0 disables)
  int use_shared;  // 1: shared-covariance kernel when every filter of a launch carries the same P (env B200KF_SHARED=0 disables)
  int ring_f_cap;  // largest filters-per-CTA the ring kernels may pick: 896 / 448 / 224 (env B200KF_RING_F)
  int use_tree;  // 1: prefix-sharing schedule-search kernel for aligned blocks (env B200KF_TREE=0 disables)
  int use_select_warp;  // 1: warp-per-instance greedy selection for few instances (env B200KF_SELECT_WARP=0: thread per instance)
  int use_ca_ring;  // 1: per-filter ring kernel for the constant-acceleration models CA15 / CA8 (env B200KF_CA_RING=0: LDG kernels)
  int use_bulk_out;  // 1: per-step outputs of the ring kernels leave through the shared-memory output ring + bulk stores (env B200KF_BULK_OUT=0: per-thread stores)
  int use_axis_dense;  // 1: dense covariance of the three-axis model (CA15) on three lanes per filter (env B200KF_AXIS_DENSE=0: thread per filter)
  int use_tma;  // 1: TMA-staged kernel for the closed-form filters (env B200KF_TMA=0 disables)
};

namespace b200kf {

constexpr int kMaxSearchBlocks = 8192;
constexpr int kMaxSearchPartials = 24000;  // 40-byte ScorePartial entries in the search scratch (< 1 MB)

int set_error(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define B200KF_CUDA_TRY(expr)                                   \
  do {                                                          \
    cudaError_t _e = (expr);                                    \
    if (_e != cudaSuccess) return ::b200kf::cuda_fail(_e, #expr); \
  } while (0)

struct RunParams {
  ModelConst mc;
  long long n_filters;
  int n_steps;
  const uint8_t* types;
  const double* dts;
  const void* payload;
  void* x;
  void* P;
  double* agg;
  void* traj;
  int traj_ncomp;
  void* ptraj;  // [T, n*n, N] covariance after every step, or null
  void* logdet;
  uint8_t* used;
  int* status;
  double gate_threshold;
  const double* gate_thresholds;  // DEVICE [N] or null: per-filter thresholds (b200kf_run_batch_gated)
  int gate_enabled;
  int agg_continue;
  const int* dense_flag;  // AUTO mode: int[3] {coupled, asymmetric, P differs between filters}; a kernel runs only if they select its mode
  const void* gain_recs;  // b200kf_run_gains: DEVICE [T] precomputed gain records (GainRecG of the model)
  const int* step_rows;   // DEVICE [T] payload row of every step, or null (step k reads row k)
  int small_tiles;        // b200kf_run_gains: 128 filters per CTA where 224 would be picked (B200KF_RUN_SMALL_TILES)
};

constexpr int kMaxEvents = 32;

struct ScoreParams {
  ModelConst mc;
  int n_events;
  unsigned char types[kMaxEvents];
  double times[kMaxEvents];
  double t_prev, t_end;
  double P0[kMaxN * kMaxN];  // [n, n] row-major
  unsigned long long first, count;
  double threshold;
  double initial_logdet;  // NaN: compute term 0 on the device; else the host's slogdet(P0)
  double* per_schedule;  // [count of the API call, 4] or null; row = schedule index - out_base
  void* partials;        // per-block results
  const int* dense_flag;
  unsigned long long out_base;  // first schedule of the API call (a call may be split into several launches)
  int tree_a, tree_k;           // prefix-sharing kernel: [first, first + 2^tree_a) aligned block, DFS over its top tree_k bits
  // Strided enumeration (b200kf_score_schedules_strided): first / count / out_base / tree_a index j, and the schedule
  // is s = low_bits | (j << n_low) -- the EARLIEST n_low events are fixed, the way a sharded search splits the space
  unsigned long long low_bits;
  int n_low;
  // the threshold in the determinant's own scale (score_leaf): exp(threshold), exp(threshold -/+ 1e-9 max(1, |threshold|));
  // q_slack: relative slack of the (exponent + mantissa) pre-filter of the sum score
  double det_thr, det_lo, det_hi, q_slack;
};

constexpr int kMaxMaskedEvents = 64;
constexpr int kMaxPayload = 9;

// evaluate_combo_chunk_worker (kf_workers.py:22-97) with state: every schedule (bitmask over one
// shared window) runs the full filter; window, payload, x0, P0 travel as kernel parameters.
struct MaskedParams {
  ModelConst mc;
  int n_events;
  unsigned char types[kMaxMaskedEvents];
  double times[kMaxMaskedEvents];
  double payload[kMaxMaskedEvents * kMaxPayload];  // [L, d]
  double t_prev, t_end;
  double x0[kMaxN];
  double P0[kMaxN * kMaxN];
  long long n_schedules;
  const unsigned long long* masks;  // DEVICE [N]
  double* x_final;                  // DEVICE [N, n]
  double* P_final;                  // DEVICE [N, n, n] or null
  double* traj;                     // DEVICE [N, L+2, 1+traj_ncomp] or null
  int traj_ncomp;
  double* logdet;                   // DEVICE [N, L+2] or null
  int* n_terms;                     // DEVICE [N] or null
};

// run_kalman_filter_scheduled's window/selection logic (kf_workers.py:877-947), covariance-only.
struct SelectParams {
  ModelConst mc;
  long long n_events;
  const unsigned char* types;  // DEVICE [E]
  const double* times;         // DEVICE [E]
  int n_inst;
  const double* periods;       // DEVICE [I]  1 / processing_frequency
  const double* t_prev;        // DEVICE [I]
  const double* P0;            // DEVICE [I, n, n]
  long long max_sel;
  int* sel;                    // DEVICE [I, max_sel] selected event indices
  long long* n_sel;            // DEVICE [I]
  double* P_final;             // DEVICE [I, n, n] or null
  void* gain_recs;             // DEVICE [I, max_sel] gain records of the selected events (axis-symmetric form) or null
  int lane_stride;             // set by the launcher: 32 = one instance per warp (few instances), 1 = packed
  // chunked walk (b200kf_select_greedy_chunk): events [e_begin, n_events) of the stream, the walk's state between
  // calls in `carry` (DEVICE f64 [I, 8]: prev, q_first, first_gps, first_imu, t_gps, t_imu, ty_imu, started), the
  // covariance in P_final (in/out) and the selection count in n_sel (in/out).  carry == null: a whole walk.
  long long e_begin;
  double* carry;
};
constexpr int kSelectCarry = 8;

// run_brute_force_kalman_filter (KF_SensorFusion.ipynb:804-954): one event per window queue,
// full filter with state per sequence, RMSE-vs-track objective.
constexpr int kMaxWindows = 64;

struct ProductParams {
  ModelConst mc;
  int n_windows;
  int q_len[kMaxWindows];                 // events in queue w
  int q_off[kMaxWindows];                 // first flattened event of queue w
  unsigned long long stride[kMaxWindows]; // index stride of digit w (itertools.product: last fastest)
  const unsigned char* types;             // DEVICE [E]
  const double* times;                    // DEVICE [E]
  const double* payload;                  // DEVICE [E, d]
  const double* ref;                      // DEVICE [E, 3] track position at each event time
  double x0[kMaxN];
  double P0[kMaxN * kMaxN];
  unsigned long long first, count;
  double* per_combo;                      // DEVICE [count] or null
  void* partials;                         // per-block ProductPartial
};

struct ProductPartial {
  double score;            // min RMSE
  unsigned long long index;
  double runner_up;        // second-smallest RMSE seen
};

struct ScorePartial {
  double score;
  unsigned long long index;
  unsigned long long key;
  unsigned long long n_feasible;
  unsigned long long n_near;
};

// defined once per model id in kf_run_model.cu / kf_sched_model.cu
template <int MODEL>
int launch_run_m(const b200kf_model* m, const RunParams& p, unsigned flags, cudaStream_t st);
template <int MODEL>
int launch_score_m(const b200kf_model* m, ScoreParams p, unsigned flags, b200kf_best* best,
                   cudaStream_t st);

template <int MODEL>
int launch_masked_m(const b200kf_model* m, const MaskedParams& p, unsigned flags, cudaStream_t st);

template <int MODEL>
int launch_gains_m(const b200kf_model* m, const RunParams& p, cudaStream_t st);
template <int MODEL>
size_t gain_record_bytes_m();

template <int MODEL>
int launch_product_m(const b200kf_model* m, const ProductParams& p, unsigned flags, cudaStream_t st);

template <int MODEL>
int launch_select_m(const b200kf_model* m, const SelectParams& p, unsigned flags, cudaStream_t st);

// One per-call scratch allocation from the model's pool, returned to the pool on the same stream when the
// guard goes out of scope (i.e. behind every kernel the call enqueued).
struct ScratchBuf {
  void* p = nullptr;
  cudaStream_t st = nullptr;
  ScratchBuf() = default;
  ScratchBuf(const ScratchBuf&) = delete;
  ScratchBuf& operator=(const ScratchBuf&) = delete;
  ~ScratchBuf() {
    if (p != nullptr) cudaFreeAsync(p, st);
  }
  cudaError_t alloc(const b200kf_model* m, size_t bytes, cudaStream_t s) {
    st = s;
    return cudaMallocFromPoolAsync(&p, bytes > 0 ? bytes : 16, m->pool, s);
  }
};

}  // namespace b200kf
