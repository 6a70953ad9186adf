/* b200kf.h -- C ABI of libb200kf.so, the B200-native batched GPS/IMU Kalman-filter engine.
 *
 * The reference (IseanB/SensorFusion-KalmanFilter) is pure Python and has no FFI; the boundary
 * it offers is the Python call surface of kf_workers.py / hw5_2.py.  Each entry
 * point below names the reference code it replaces; INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add.  Conventions:
 *
 *   - every function returns an int status: 0 = OK, negative = error (b200kf_last_error gives the
 *     thread-local message); no C++ exception crosses the boundary;
 *   - pointers documented DEVICE are caller-owned CUDA allocations of the current device, all
 *     others are host pointers; the library never frees caller memory and keeps no pointer past
 *     the call;
 *   - work is enqueued on the caller's stream (a cudaStream_t passed as void*, NULL = default
 *     stream); functions are asynchronous unless stated;
 *   - a model object may be used from several host threads and on several streams at once: every call takes
 *     its device scratch (dispatch flags, packed schedules, reduction partials) from the model's stream-ordered
 *     pool on the caller's stream and returns it behind its last kernel, so concurrent calls share nothing;
 *     b200kf_run_batch_host (synchronous, owns a copy/compute stream pair) serialises per model;
 *   - "real" buffers hold float64 or float32 according to the model's precision; time steps
 *     `dt` are ALWAYS float64 and are computed by the host as t - t_prev from float64 epoch
 *     seconds (kf_workers.py:682), never from float32 timestamps.
 */
#ifndef B200KF_H_
#define B200KF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200KF_VERSION_MAJOR 0
#define B200KF_VERSION_MINOR 1

/* status codes */
#define B200KF_OK 0
#define B200KF_ERR_INVALID (-1)     /* bad argument */
#define B200KF_ERR_UNSUPPORTED (-2) /* shape / model outside the compiled family */
#define B200KF_ERR_CUDA (-3)        /* CUDA runtime error, text in b200kf_last_error */
#define B200KF_ERR_NO_DEVICE (-4)   /* no CUDA device: there is NO CPU fallback */

/* model ids: state layouts and F(dt) sparsity are compiled in, rate constants are data */
#define B200KF_MODEL_CV6 0   /* [p(3),v(3)]: kf_workers.py model restricted to rows {0,1,2,6,7,8} */
#define B200KF_MODEL_CA8 1   /* hw5_2.py:219-304  [x,y,th,vx,vy,thdot,ax,ay] */
#define B200KF_MODEL_BIAS9 2 /* [p(3),v(3),b(3)], IMU acceleration as control input */
#define B200KF_MODEL_CA15 3  /* kf_workers.py:493-614  [p,rpy,v,w,a] */
#define B200KF_MODEL_CVBU6 4 /* [p(3),v(3)], x = F x + B u: IMU acceleration as control input, GPS-only updates */

#define B200KF_F64 0
#define B200KF_F32 1

/* event types of a schedule (uint8) */
#define B200KF_EV_GPS 0          /* predict + GPS update   kf_workers.py:694-697 */
#define B200KF_EV_IMU 1          /* predict + IMU pseudo-measurement update (or control) :698-706 */
#define B200KF_EV_PREDICT_ONLY 2 /* predict, no update     kf_workers.py:74-82, 1060-1159 */

/* b200kf_run_batch flags */
#define B200KF_RUN_AUTO 0u            /* the cheapest valid variant, decided on the device per launch:
                                         shared > symmetric > decoupled > dense */
#define B200KF_RUN_FORCE_DENSE 1u     /* always the dense (all-states-coupled) kernel */
#define B200KF_RUN_FORCE_DECOUPLED 2u /* caller guarantees chain-compatible P (no check) */
#define B200KF_RUN_AGG_CONTINUE 4u    /* agg holds the running aggregates of earlier chunks */
#define B200KF_RUN_FORCE_SYMMETRIC 8u /* caller guarantees chain-compatible P whose chain blocks are
                                         identical across axes (and the model's rates agree): one
                                         covariance recursion per filter; falls back to
                                         FORCE_DECOUPLED when the model / launch cannot use it */

#define B200KF_RUN_FORCE_SHARED 16u   /* caller guarantees FORCE_SYMMETRIC's conditions AND that every
                                         filter of the launch carries the same P: one covariance /
                                         gain recursion per launch; falls back to FORCE_SYMMETRIC
                                         when the model / launch cannot use it */
#define B200KF_RUN_SMALL_TILES 32u    /* b200kf_run_gains only: 128 filters per CTA where 224 would be used.  For launches
                                         that share the GPU with others (one per sampling rate of a sweep,
                                         kf_workers.py:826-957 run once per rate): a CTA streams ~40 GB/s (ring bytes over
                                         memory latency), so the rate with the longest schedule needs more CTAs */

typedef struct b200kf_model b200kf_model; /* opaque, library-owned */

/* Replaces the eight model getters of KF_SensorFusion (kf_workers.py:493-621, hw5_2.py:219-311):
 * their VALUES travel here, their sparsity pattern is fixed by model_id. */
typedef struct b200kf_model_desc {
  int32_t model_id;
  int32_t precision;  /* B200KF_F64 / B200KF_F32 */
  double q[16];       /* Q(dt) = dt * diag(q)           get_process_noise_covariance_matrix */
  double r_gps[4];    /* R_GPS = diag(r_gps)            get_gps_measurement_noise_covariance_matrix */
  double r_imu[16];   /* R_IMU = diag(r_imu)            get_imu_measurement_noise_covariance_matrix */
  double accel_coef;  /* coefficient of dt^2 in F (0.5) get_state_transition_matrix */
} b200kf_model_desc;

/* Results of a schedule search over one index range (one GPU's shard). */
typedef struct b200kf_best {
  double argmin_sum_score;       /* min over the range of sum(log det P); +inf if range empty */
  uint64_t argmin_sum_index;     /* its schedule index, ties -> lowest index */
  uint64_t first_feasible_key;   /* (popcount << 32) | order-in-cardinality; UINT64_MAX if none */
  uint64_t first_feasible_index; /* first schedule, in the reference's enumeration order
                                    (kf_workers.py:1325-1336), with max(log det) < threshold */
  uint64_t n_feasible;           /* schedules in range with max(log det) < threshold */
  uint64_t n_near_threshold;     /* schedules whose largest DEVICE-computed term lies within
                                    1e-9*max(1,|threshold|) of the threshold (and which the caller's exact
                                    initial_logdet does not already rule out): the feasibility decisions a
                                    1e-12 rounding difference could flip.  The host mirror re-scores exactly
                                    those in the reference's numpy arithmetic (kf_workers.py epsilon guard) */
} b200kf_best;

/* Result of a product-of-queues search over one index range. */
typedef struct b200kf_best_product {
  double score;           /* min RMSE over the range; +inf if the range is empty */
  uint64_t index;         /* its sequence index in itertools.product order, ties -> lowest index */
  double runner_up_score; /* second-smallest RMSE in the range (+inf if none): a gap below ~1e-9
                             relative means a rounding difference could flip the selection */
  uint64_t n_scored;
} b200kf_best_product;

int b200kf_version(int* major, int* minor);
size_t b200kf_last_error(char* buf, size_t n);

/* number of CUDA devices visible (0 = the library cannot compute) */
int b200kf_device_count(int* count);

/* fills desc with the reference's constants for model_id (kf_workers.py:519-614, hw5_2.py:233-304) */
int b200kf_model_default_desc(int32_t model_id, int32_t precision, b200kf_model_desc* desc);
int b200kf_model_create(const b200kf_model_desc* desc, b200kf_model** out);
int b200kf_model_destroy(b200kf_model* model);
/* n = states, d = payload reals per event, m_gps = GPS rows */
int b200kf_model_dims(const b200kf_model* model, int32_t* n, int32_t* d, int32_t* m_gps);

/* The filter loop: replaces the event loops of run_kalman_filter_full (kf_workers.py:681-721),
 * run_kalman_filter (:767-823), run_adaptive_threshold_kalman_filter (:1010-1053),
 * run_no_update_kalman_filter (:1111-1154), the inner loop of run_kalman_filter_scheduled
 * (:914-947) and hw5_2.run_kalman_filter (hw5_2.py:332-379), for n_filters independent filters
 * that share one event schedule.
 *
 *   types   DEVICE u8  [T]        event type per step
 *   dts     DEVICE f64 [T]        t - t_prev per step
 *   payload DEVICE real[T, d, N]  time-major, filter index fastest.  GPS rows: (E, N, alt) in
 *                                 columns 0..m_gps-1; IMU rows: CA15 (r,p,y,wx,wy,wz,ax,ay,az),
 *                                 CA8 (yaw, wz, ax, ay), CV6/BIAS9/CVBU6 (ax, ay, az)
 *   x       DEVICE real[N, n]     in: state before step 0, out: state after step T-1
 *   P       DEVICE real[N, n, n]  in/out covariance, row-major, full symmetric
 *   agg     DEVICE f64 [N, 4]     nullable; (max, sum, last) of log det P over the steps and the
 *                                 step count; with B200KF_RUN_AGG_CONTINUE max/sum/count continue
 *   traj    DEVICE real[T, traj_ncomp, N]  nullable; x[0:traj_ncomp] after each step (:714)
 *   ptraj   DEVICE real[T, n*n, N] nullable; full covariance after each step (:822, legacy loop)
 *   logdet  DEVICE real[T, N]     nullable; log det P after each step (:716-717)
 *   used    DEVICE u8  [T, N]     nullable; 1 where the update ran (adaptive gate, :1025-1027)
 *   status  DEVICE i32 [N]        nullable; 0 ok, 1 = a non-positive determinant was seen
 *   gate_threshold  NaN disables; else update only when log det of the PREDICTED P exceeds it
 */
int b200kf_run_batch(const b200kf_model* model, int64_t n_filters, int64_t n_steps,
                     const uint8_t* types, const double* dts, const void* payload, void* x, void* P,
                     double* agg, void* traj, int32_t traj_ncomp, void* ptraj, void* logdet,
                     uint8_t* used, int32_t* status, double gate_threshold, uint32_t flags,
                     void* stream);

/* b200kf_run_batch over a SELECTION of a longer stream without a gathered copy of the payload:
 * step k reads payload row step_rows[k] (DEVICE i32 [T], 0 <= row < rows of `payload`); types and dts
 * stay per step.  What run_kalman_filter_scheduled does with its selected events (kf_workers.py:914-947)
 * for many seeds that share one base stream.  step_rows = NULL is b200kf_run_batch. */
int b200kf_run_batch_rows(const b200kf_model* model, int64_t n_filters, int64_t n_steps,
                          const uint8_t* types, const double* dts, const void* payload,
                          const int32_t* step_rows, void* x, void* P, double* agg, void* traj,
                          int32_t traj_ncomp, void* ptraj, void* logdet, uint8_t* used, int32_t* status,
                          double gate_threshold, uint32_t flags, void* stream);

/* b200kf_run_batch with ONE GATE THRESHOLD PER FILTER (gate_thresholds: DEVICE f64 [N]; a filter updates
 * only when log det of its predicted P exceeds its threshold, kf_workers.py:1023-1025).  The reference's
 * experiment runs the adaptive filter from the start of the stream once per window, each with its own
 * R_threshold (kf_workers.py:2316): filters with different thresholds share one launch here. */
int b200kf_run_batch_gated(const b200kf_model* model, int64_t n_filters, int64_t n_steps,
                           const uint8_t* types, const double* dts, const void* payload,
                           const double* gate_thresholds, void* x, void* P, double* agg, void* traj,
                           int32_t traj_ncomp, void* ptraj, void* logdet, uint8_t* used, int32_t* status,
                           uint32_t flags, void* stream);

/* Same computation from HOST buffers (the call a drop-in user makes): streams `payload` to the
 * device in time chunks, overlapping host->device copies with the kernel, and copies x, P, agg
 * back.  Synchronous.  payload/x/P/agg are HOST pointers (pinned memory copies fastest; a payload
 * in pageable memory -- a plain numpy array -- is staged through two pinned chunk buffers by a few
 * host threads, env B200KF_STAGE_THREADS); types and dts are host arrays.  chunk_steps = 0 picks a
 * default. */
int b200kf_run_batch_host(const b200kf_model* model, int64_t n_filters, int64_t n_steps,
                          const uint8_t* types, const double* dts, const void* payload, void* x,
                          void* P, double* agg, int32_t* status, int64_t chunk_steps, uint32_t flags);

/* The brute-force schedule search: replaces evaluate_combo_chunk_worker's log_det trace
 * (kf_workers.py:22-97) and the enumeration + selection of
 * run_brute_force_kalman_filter_no_sampling_min_usage (:1325-1371) for the schedules
 * [first, first + count) of a window of n_events <= 32 events.  Schedule s uses event i iff bit i
 * of s is set; gaps over skipped events merge into one predict (:37) and the filter is propagated
 * to t_end at the end (:74-82).  Covariance-only: log det P never depends on measurement values.
 *
 *   types, times  HOST [n_events]   event types and float64 epoch times of the window
 *   P0            HOST f64 [n, n]   covariance at t_prev
 *   initial_logdet  NaN: term 0 (log det P0, :32) is computed on the device.  Otherwise the
 *                 caller's value is used verbatim for term 0 of every schedule -- the reference
 *                 compares max(log_det) < threshold strictly (:1353) and a threshold derived from
 *                 slogdet(P0) itself ties exactly with it
 *   per_schedule  DEVICE f64 [count, 4] nullable: max, sum, last of the log-det trace (the max
 *                                   and sum include the initial and the end-propagation terms,
 *                                   :32, :81-82, :1353) and the number of terms
 *   best          DEVICE b200kf_best, written when the stream reaches it
 */
int b200kf_score_schedules(const b200kf_model* model, int32_t n_events, const uint8_t* types,
                           const double* times, double t_prev, double t_end, const double* P0,
                           uint64_t first, uint64_t count, double threshold, double initial_logdet,
                           double* per_schedule, b200kf_best* best, uint32_t flags, void* stream);

/* The same search over the schedules that SHARE their n_low earliest decisions: schedule
 * s = low_bits | (j << n_low) for j in [first, first + count), j < 2^(n_events - n_low); per_schedule
 * rows are indexed by j - first, every index in `best` is the full schedule s.  This is how the
 * enumeration of run_brute_force_kalman_filter_no_sampling_min_usage (kf_workers.py:1325-1371)
 * is split over workers of EQUAL cost: the fixed early events are filtered once per thread,
 * whereas a contiguous index range fixes the LATEST events, which every schedule pays on its own
 * (8 ranges of a 2^32 search: 16.4 ... 23.9 ms).  n_low = 0 is b200kf_score_schedules. */
int b200kf_score_schedules_strided(const b200kf_model* model, int32_t n_events, const uint8_t* types,
                                   const double* times, double t_prev, double t_end, const double* P0,
                                   uint64_t low_bits, int32_t n_low, uint64_t first, uint64_t count,
                                   double threshold, double initial_logdet, double* per_schedule,
                                   b200kf_best* best, uint32_t flags, void* stream);

/* The brute-force worker WITH state: replaces evaluate_combo_chunk_worker (kf_workers.py:22-97)
 * for n_schedules combos that are subsets of one shared window of n_events <= 64 time-ordered
 * events.  Schedule f uses event i iff bit i of masks[f] is set.  Float64 only.
 *
 *   types, times   HOST [L]          payload  HOST f64 [L, d] (row layout as b200kf_run_batch)
 *   x0             HOST f64 [n]      P0       HOST f64 [n, n]
 *   masks          DEVICE u64 [N]
 *   x_final        DEVICE f64 [N, n]           state after the last term (:86)
 *   P_final        DEVICE f64 [N, n, n]        nullable
 *   traj           DEVICE f64 [N, L+2, 1+traj_ncomp] nullable; row r = (t, x[0:traj_ncomp]) of
 *                                              term r: initial, one per used event, end (:31,68,80)
 *   logdet         DEVICE f64 [N, L+2]         nullable; log det P of term r (:32,70-71,81-82)
 *   n_terms        DEVICE i32 [N]              nullable; rows written (negative: a non-positive
 *                                              determinant was seen)
 */
int b200kf_run_masked(const b200kf_model* model, int32_t n_events, const uint8_t* types,
                      const double* times, const double* payload, double t_prev, double t_end,
                      const double* x0, const double* P0, int64_t n_schedules, const uint64_t* masks,
                      double* x_final, double* P_final, double* traj, int32_t traj_ncomp,
                      double* logdet, int32_t* n_terms, uint32_t flags, void* stream);

/* The notebook's brute force: replaces the enumeration, the per-sequence filter and the RMSE
 * objective of run_brute_force_kalman_filter (KF_SensorFusion.ipynb:849-938) for the sequences
 * [first, first + count) of itertools.product(*queues) (last queue fastest, :873-875).  A sequence
 * takes ONE event from each of the n_windows <= 64 window queues, runs the full filter with state
 * from (x0, P0) with dt = 0 at its first event (:898), and is scored by
 * sqrt(mean_k |x_k[0:3] - ref(event_k)|^2) (calculate_accuracy_metrics, :766-802).  Float64 only.
 *
 *   queue_len  HOST  i32 [n_windows]   events per queue (each >= 1), E = sum
 *   types, times  DEVICE [E]           the queues' events, flattened queue after queue
 *   payload    DEVICE f64 [E, d]       row layout as b200kf_run_batch
 *   ref_pos    DEVICE f64 [E, 3]       track position interpolated at each event time
 *                                      (b200kf_track_lerp)
 *   x0, P0     HOST f64 [n], [n, n]
 *   per_combo  DEVICE f64 [count]      nullable: RMSE of every sequence
 *   best       DEVICE b200kf_best_product; first minimum in enumeration order (strict '<', :933)
 */
int b200kf_score_products(const b200kf_model* model, int32_t n_windows, const int32_t* queue_len,
                          const uint8_t* types, const double* times, const double* payload,
                          const double* ref_pos, const double* x0, const double* P0, uint64_t first,
                          uint64_t count, double* per_combo, b200kf_best_product* best,
                          uint32_t flags, void* stream);

/* scipy interp1d(kind='linear', fill_value='extrapolate') of a time-sorted track at n_query times
 * (kf_workers.py:1195-1201, KF_SensorFusion.ipynb:779-783).  track_t DEVICE f64 [n_track],
 * track_p DEVICE f64 [n_track, k], query_t DEVICE f64 [n_query], out DEVICE f64 [n_query, k].
 * n_track = 1 returns the single point. */
int b200kf_track_lerp(int64_t n_track, const double* track_t, const double* track_p, int32_t k,
                      int64_t n_query, const double* query_t, double* out, void* stream);

/* total_position_rmse of calculate_accuracy_metrics (kf_workers.py:1204-1208) for every filter of
 * a time-major trajectory as written by b200kf_run_batch: traj DEVICE real [T, traj_ncomp, N]
 * (components 0..2 = position), ref_pos DEVICE f64 [T, 3] the track at the step times,
 * rmse DEVICE f64 [N].  Deterministic (fixed summation order). */
int b200kf_rmse_vs_track(const b200kf_model* model, int64_t n_filters, int64_t n_steps,
                         const void* traj, int32_t traj_ncomp, const double* ref_pos, double* rmse,
                         void* stream);

/* Greedy one-event-per-window scheduling: replaces the queueing + Scheduler.greedy_schedule
 * selection of run_kalman_filter_scheduled (kf_workers.py:877-910, :174-213) for n_inst
 * independent instances (processing period, start time, P0) walking ONE shared event list.  The
 * selection reads P but never a measurement, so this is a covariance-only pass; feed the selected
 * events to b200kf_run_batch for the states.
 *
 *   types, times  DEVICE [E]       events AFTER the initialising one (:877)
 *   periods       DEVICE f64 [I]   1 / processing_frequency
 *   t_prev        DEVICE f64 [I]   previous_time at entry
 *   P0            DEVICE f64 [I, n, n]
 *   sel           DEVICE i32 [I, max_sel]  selected event index per window (first max_sel kept)
 *   n_sel         DEVICE i64 [I]   number of windows (may exceed max_sel: call again larger)
 *   P_final       DEVICE f64 [I, n, n] nullable
 */
int b200kf_select_greedy(const b200kf_model* model, int64_t n_events, const uint8_t* types,
                         const double* times, int32_t n_inst, const double* periods,
                         const double* t_prev, const double* P0, int64_t max_sel, int32_t* sel,
                         int64_t* n_sel, double* P_final, uint32_t flags, void* stream);

/* b200kf_select_greedy that also publishes the gains of the selected events.  The selection walks the
 * very covariance recursion the selected events' filter needs (kf_workers.py:914-947), so it can
 * hand that filter its gain schedule: gain_records DEVICE [I, max_sel] records of
 * b200kf_gain_record_bytes(model) bytes each (16-byte aligned), record j of instance i = (dt, type,
 * K = P+ H^T R^-1 of every block of the model's axis-symmetric form, det P) of its j-th selected
 * event -- or, from the warp-per-instance walk (<= 2048 instances), the same record in its COVARIANCE
 * form (dt, type, P+ of every block; flagged in the record): b200kf_run_gains forms K and det P from it,
 * the same products.  Records are opaque to the caller: hand them to b200kf_run_gains unchanged.  Needs B200KF_RUN_FORCE_SYMMETRIC (the caller vouches for a chain-compatible P0 with
 * identical blocks across axes; the reference's P0 and every P it produces).  NULL = no records. */
int b200kf_select_greedy_gains(const b200kf_model* model, int64_t n_events, const uint8_t* types,
                               const double* times, int32_t n_inst, const double* periods,
                               const double* t_prev, const double* P0, int64_t max_sel, int32_t* sel,
                               int64_t* n_sel, double* P_final, void* gain_records, uint32_t flags,
                               void* stream);
int b200kf_gain_record_bytes(const b200kf_model* model, size_t* bytes);

/* The same walk in CHUNKS of the event stream (run_kalman_filter_scheduled's loop, kf_workers.py:877-947, cut at
 * arbitrary event indices): events [e_begin, e_end) of types/times; the walk's state between calls -- previous
 * selected time, the open window's queue summary -- lives in `carry` (DEVICE f64 [I, 8], zero-filled before the
 * first chunk), the covariance in P_state (DEVICE f64 [I, n, n], written by every call, read by all but the
 * first, which starts from P0 / t_prev) and the running selection count in n_sel (in/out).  sel / gain_records
 * are filled at their global positions [I, max_sel].  Chunk after chunk gives bit for bit the selections,
 * covariances and records of one b200kf_select_greedy_gains call over the whole stream; a caller can so run
 * the selection of chunk c+1 on one stream while the filters of chunk c consume its records on another.
 * At most 2048 instances (one warp each). */
int b200kf_select_greedy_chunk(const b200kf_model* model, int64_t e_begin, int64_t e_end, const uint8_t* types,
                               const double* times, int32_t n_inst, const double* periods,
                               const double* t_prev, const double* P0, double* carry, double* P_state,
                               int64_t max_sel, int32_t* sel, int64_t* n_sel, void* gain_records,
                               uint32_t flags, void* stream);

/* The filter loop for n_filters filters that share a schedule AND a covariance, given that
 * schedule's precomputed gain records (Monte-Carlo seeds of one configuration: the records are
 * computed once, every batch of seeds is then a pure stream -- no covariance recursion, no serial
 * work in the launch).  x += K (z - H x) per event with the reference's pseudo-measurement
 * (kf_workers.py:698-711); the covariance never enters (its producer returned P_final).
 *
 *   gain_records DEVICE [T] records (b200kf_select_greedy_gains)      payload DEVICE real [T, d, N]
 *   step_rows    DEVICE i32 [T] nullable: payload row of step k (as b200kf_run_batch_rows)
 *   x       DEVICE real [N, n] in/out        agg DEVICE f64 [N, 4] nullable (from the records' det P)
 *   traj / logdet / used / status as b200kf_run_batch; flags: B200KF_RUN_AGG_CONTINUE, B200KF_RUN_SMALL_TILES
 * Rows must be 16-byte aligned: N even (float64) or a multiple of 4 (float32). */
int b200kf_run_gains(const b200kf_model* model, int64_t n_filters, int64_t n_steps,
                     const void* gain_records, const void* payload, const int32_t* step_rows, void* x,
                     double* agg, void* traj, int32_t traj_ncomp, void* logdet, uint8_t* used,
                     int32_t* status, uint32_t flags, void* stream);

/* Batched Scheduler.cov_matrix (kf_workers.py:112-171; the reference's only device='cuda'
 * branch): Sigma_post = Sigma - K H^ Sigma with H^ = the k selected rows of H, for `batch`
 * covariances.  H HOST f64 [m, n], R HOST f64 [m, m], rows HOST i32 [k] (0-based),
 * Sigma / Sigma_post DEVICE f64 [batch, n, n], trace DEVICE f64 [batch] nullable (gain, :174-185).
 * n <= 16, k <= 16. */
int b200kf_cov_update(int32_t n, int32_t m, int32_t k, const int32_t* rows, const double* H,
                      const double* R, int64_t batch, const double* Sigma, double* Sigma_post,
                      double* trace, void* stream);

/* Synthetic payload of BASELINE config 3/5 generated on the device (counter-based hash keyed by
 * (seed, step, column, filter)): GPS rows ~ N(0, 3) in the first m_gps columns, other rows
 * ~ N(0, 0.1^2).  payload DEVICE real [T, d, N]. */
int b200kf_fill_payload(const b200kf_model* model, int64_t n_filters, int64_t n_steps,
                        const uint8_t* types /* DEVICE u8 [T] */, uint64_t seed, void* payload,
                        void* stream);

/* Measures the FP64 (or FP32) FMA peak of the current device with a register-only kernel; the
 * FP-roofline denominator (MEASURED_PEAKS.json has none).  precision bit 0: B200KF_F64/F32; bit 1
 * set: every FMA reads three distinct, changing registers (the register-file-limited rate a
 * filter step sees) instead of two loop-invariant ones.  Synchronous. */
int b200kf_probe_fma_peak(int32_t precision, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* B200KF_H_ */
