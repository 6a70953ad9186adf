"""Batched entry points over the C ABI (the calls the reference never had: many filters, many
schedules at once).  PyTorch is used only to own device memory and streams; every pointer handed
to libb200kf.so is a `tensor.data_ptr()`.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, byref, c_double, c_int32, c_uint8, c_void_p

import os

import numpy as np

from . import _lib
from .models import ModelSpec, default_spec

_MODES = {"auto": _lib.RUN_AUTO, "dense": _lib.RUN_FORCE_DENSE, "decoupled": _lib.RUN_FORCE_DECOUPLED,
          "symmetric": _lib.RUN_FORCE_SYMMETRIC, "shared": _lib.RUN_FORCE_SHARED}


def _torch():
    import torch
    return torch


def _require_cuda():
    torch = _torch()
    if not torch.cuda.is_available() or _lib.device_count() == 0:
        raise _lib.B200KFError(_lib.ERR_NO_DEVICE, "no CUDA device visible: b200kf has no CPU fallback")
    return torch


def _ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


class Engine:
    """One model instance on one device (wraps b200kf_model_create/_destroy)."""

    def __init__(self, spec: ModelSpec | str = "CA15", precision: str | None = None, device=None):
        if isinstance(spec, str):
            spec = default_spec(spec, precision or "f64")
        elif precision is not None and precision != spec.precision:
            spec = ModelSpec(spec.name, precision, spec.q, spec.r_gps, spec.r_imu, spec.accel_coef)
        self.spec = spec
        self.lib = _lib.load()
        torch = _require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.real = torch.float32 if spec.precision == "f32" else torch.float64
        self._h = c_void_p(0)
        with torch.cuda.device(self.device):
            desc = spec.to_desc()
            _lib.check(self.lib.b200kf_model_create(byref(desc), byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.b200kf_model_destroy(self._h)
            self._h = c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _dev(self, a, dtype):
        torch = _torch()
        if a is None:
            return None
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(self.device)

    def _stream(self, stream):
        torch = _torch()
        s = torch.cuda.current_stream(self.device) if stream is None else stream
        return c_void_p(s.cuda_stream)

    # ------------------------------------------------------------------ filter loop
    def run_batch(self, types, dts, payload, x, P, *, want_agg=True, traj_ncomp=0, want_logdet=False,
                  want_used=False, want_ptraj=False, gate_threshold=None, mode="auto", agg=None,
                  stream=None, inplace=False, step_rows=None):
        """Run N filters over a shared schedule (b200kf_run_batch).

        types [T] u8, dts [T] f64, payload [T, d, N], x [N, n], P [N, n, n]; numpy arrays or torch
        tensors.  Returns a dict of torch CUDA tensors: x, P, agg [N,4], traj [T,traj_ncomp,N],
        logdet [T,N], used [T,N], status [N].  `agg` continues running aggregates of a previous
        chunk.  With inplace=True torch inputs x/P on the right device/dtype are updated in place."""
        torch = _torch()
        sp = self.spec
        types_d = self._dev(types, torch.uint8)
        dts_d = self._dev(dts, torch.float64)
        pay_d = self._dev(payload, self.real)
        T = int(types_d.numel())
        x_d = self._dev(x, self.real)
        P_d = self._dev(P, self.real)
        if not inplace:
            if isinstance(x, torch.Tensor) and x_d.data_ptr() == x.data_ptr():
                x_d = x_d.clone()
            if isinstance(P, torch.Tensor) and P_d.data_ptr() == P.data_ptr():
                P_d = P_d.clone()
        N = int(x_d.shape[0])
        if tuple(x_d.shape) != (N, sp.n) or tuple(P_d.shape) != (N, sp.n, sp.n):
            raise ValueError(f"x must be [N,{sp.n}] and P [N,{sp.n},{sp.n}]")
        rows_d = None
        if step_rows is not None:  # step k reads payload row step_rows[k]: a selection of a longer stream, no gather
            rows_d = self._dev(step_rows, torch.int32)
            if int(rows_d.numel()) != T or tuple(pay_d.shape[1:]) != (sp.d, N):
                raise ValueError(f"step_rows must be [T={T}] and payload [rows, d={sp.d}, N={N}]")
        elif T and tuple(pay_d.shape) != (T, sp.d, N):
            raise ValueError(f"payload must be [T={T}, d={sp.d}, N={N}], got {tuple(pay_d.shape)}")
        if dts_d.numel() != T:
            raise ValueError("types and dts must have the same length")
        flags = _MODES[mode]
        agg_d = None
        if agg is not None:
            agg_d = self._dev(agg, torch.float64)
            flags |= _lib.RUN_AGG_CONTINUE
        elif want_agg:
            agg_d = torch.empty((N, 4), dtype=torch.float64, device=self.device)
        traj_d = torch.empty((T, traj_ncomp, N), dtype=self.real, device=self.device) if traj_ncomp else None
        ld_d = torch.empty((T, N), dtype=self.real, device=self.device) if want_logdet else None
        used_d = torch.empty((T, N), dtype=torch.uint8, device=self.device) if want_used else None
        ptraj_d = torch.empty((T, sp.n * sp.n, N), dtype=self.real, device=self.device) if want_ptraj else None
        status_d = torch.zeros((N,), dtype=torch.int32, device=self.device)
        if torch.is_tensor(gate_threshold) and gate_threshold.numel() == 1:
            gate_threshold = float(gate_threshold.item())  # a 0-dim / one-element tensor is a scalar threshold
        elif gate_threshold is not None and not torch.is_tensor(gate_threshold) and np.ndim(gate_threshold) > 0 \
                and np.size(gate_threshold) == 1 and N != 1:
            gate_threshold = float(np.asarray(gate_threshold).reshape(()))
        if gate_threshold is not None and (torch.is_tensor(gate_threshold) or np.ndim(gate_threshold) > 0):
            # one threshold per filter (b200kf_run_batch_gated)
            if rows_d is not None:
                raise ValueError("per-filter gate thresholds and step_rows cannot be combined")
            gates_d = self._dev(gate_threshold, torch.float64)
            if int(gates_d.numel()) != N:
                raise ValueError(f"gate_threshold must be a scalar or [N={N}]")
            with torch.cuda.device(self.device):
                _lib.check(self.lib.b200kf_run_batch_gated(
                    self._h, N, T, _ptr(types_d), _ptr(dts_d), _ptr(pay_d), _ptr(gates_d), _ptr(x_d), _ptr(P_d),
                    _ptr(agg_d), _ptr(traj_d), int(traj_ncomp), _ptr(ptraj_d), _ptr(ld_d), _ptr(used_d), _ptr(status_d),
                    flags, self._stream(stream)))
            return {"x": x_d, "P": P_d, "agg": agg_d, "traj": traj_d, "logdet": ld_d, "used": used_d,
                    "ptraj": ptraj_d, "status": status_d}
        gate = float("nan") if gate_threshold is None else float(gate_threshold)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b200kf_run_batch_rows(
                self._h, N, T, _ptr(types_d), _ptr(dts_d), _ptr(pay_d), _ptr(rows_d), _ptr(x_d), _ptr(P_d), _ptr(agg_d),
                _ptr(traj_d), int(traj_ncomp), _ptr(ptraj_d), _ptr(ld_d), _ptr(used_d), _ptr(status_d), gate, flags,
                self._stream(stream)))
        return {"x": x_d, "P": P_d, "agg": agg_d, "traj": traj_d, "logdet": ld_d, "used": used_d,
                "ptraj": ptraj_d, "status": status_d}

    def run_batch_host(self, types, dts, payload, x, P, *, chunk_steps=0, mode="auto", inplace=False):
        """Host-buffer path (b200kf_run_batch_host): numpy (or pinned torch CPU) arrays in, numpy
        out; payload is streamed to the device in time chunks overlapped with the kernels.  With
        inplace=True, x and P (contiguous arrays / CPU tensors of the model precision) are updated
        in place instead of being copied first."""
        torch = _require_cuda()
        sp = self.spec
        npreal = np.float32 if sp.precision == "f32" else np.float64

        def host(a, dt):
            if isinstance(a, torch.Tensor):
                assert a.device.type == "cpu" and a.is_contiguous()
                return a, a.data_ptr()
            arr = np.ascontiguousarray(a, dtype=dt)
            return arr, arr.ctypes.data

        types_h, types_p = host(types, np.uint8)
        dts_h, dts_p = host(dts, np.float64)
        pay_h, pay_p = host(payload, npreal)
        if inplace:
            x_h = x.numpy() if isinstance(x, torch.Tensor) else x
            P_h = P.numpy() if isinstance(P, torch.Tensor) else P
            if not (x_h.dtype == npreal and P_h.dtype == npreal and x_h.flags.c_contiguous and P_h.flags.c_contiguous):
                raise ValueError("inplace=True needs C-contiguous x, P of the model precision")
        else:
            x_h = np.array(x, dtype=npreal, order="C", copy=True)
            P_h = np.array(P, dtype=npreal, order="C", copy=True)
        N = x_h.shape[0]
        T = int(np.prod(types_h.shape))
        agg_h = np.zeros((N, 4), dtype=np.float64)
        status_h = np.zeros((N,), dtype=np.int32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b200kf_run_batch_host(
                self._h, N, T, c_void_p(types_p), c_void_p(dts_p), c_void_p(pay_p),
                c_void_p(x_h.ctypes.data), c_void_p(P_h.ctypes.data), c_void_p(agg_h.ctypes.data),
                c_void_p(status_h.ctypes.data), int(chunk_steps), _MODES[mode]))
        return {"x": x_h, "P": P_h, "agg": agg_h, "status": status_h}

    # ------------------------------------------------------------------ schedule search
    def score_schedules(self, types, times, t_prev, t_end, P0, *, first=0, count=None,
                        threshold=float("inf"), initial_logdet=None, want_per_schedule=False, mode="auto",
                        stream=None, device_result=False, low_bits=0, n_low=0):
        """Score the subset schedules [first, first+count) of a window (b200kf_score_schedules).
        With n_low > 0 (b200kf_score_schedules_strided) the range indexes j and the schedules are
        low_bits | (j << n_low): the slice of the search whose n_low earliest decisions are low_bits.
        initial_logdet: the caller's sign*logabsdet of P0 used verbatim as term 0 (exact ties with
        a threshold derived from it resolve as in the reference); None computes it on the device.
        Returns (best: dict, per_schedule: torch [count,4] or None).  With device_result=True `best` is the
        raw 48-byte b200kf_best as a CUDA uint8 tensor and nothing is copied to the host (no synchronisation:
        the sharded search all-gathers it as it is, distributed.search_sharded)."""
        torch = _torch()
        types = np.ascontiguousarray(types, dtype=np.uint8)
        times = np.ascontiguousarray(times, dtype=np.float64)
        P0 = np.ascontiguousarray(P0, dtype=np.float64)
        L = int(types.shape[0])
        if count is None:
            count = (1 << (L - int(n_low))) - first
        per = torch.empty((count, 4), dtype=torch.float64, device=self.device) if want_per_schedule else None
        best_d = torch.empty((ctypes.sizeof(_lib.Best),), dtype=torch.uint8, device=self.device)  # every field is written
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b200kf_score_schedules_strided(
                self._h, L, types.ctypes.data_as(POINTER(c_uint8)), times.ctypes.data_as(POINTER(c_double)),
                float(t_prev), float(t_end), P0.ctypes.data_as(POINTER(c_double)), int(low_bits), int(n_low),
                int(first), int(count),
                float(threshold), float("nan") if initial_logdet is None else float(initial_logdet),
                _ptr(per), _ptr(best_d), _MODES[mode], self._stream(stream)))
        if device_result:
            return best_d, per
        return self.best_from_bytes(best_d.cpu().numpy().tobytes()), per

    @staticmethod
    def best_from_bytes(raw):
        """b200kf_best (include/b200kf.h) as a dict."""
        b = _lib.Best.from_buffer_copy(bytes(raw))
        none = (1 << 64) - 1
        return {
            "argmin_sum_score": b.argmin_sum_score,
            "argmin_sum_index": None if b.argmin_sum_index == none else int(b.argmin_sum_index),
            "first_feasible_key": int(b.first_feasible_key),
            "first_feasible_index": None if b.first_feasible_key == none else int(b.first_feasible_index),
            "n_feasible": int(b.n_feasible),
            "n_near_threshold": int(b.n_near_threshold),
        }


    def run_masked(self, types, times, payload, t_prev, t_end, x0, P0, masks, *, traj_ncomp=6,
                   want_P=False, mode="auto", stream=None):
        """Full filter per subset schedule of one shared window (b200kf_run_masked): what
        evaluate_combo_chunk_worker computes per combo.  payload [L, d]; masks: iterable of ints.
        Returns dict of torch CUDA tensors x_final [N,n], P_final, traj [N,L+2,1+ncomp],
        logdet [N,L+2], n_terms [N]."""
        torch = _torch()
        sp = self.spec
        types = np.ascontiguousarray(types, dtype=np.uint8)
        times = np.ascontiguousarray(times, dtype=np.float64)
        payload = np.ascontiguousarray(payload, dtype=np.float64)
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        P0 = np.ascontiguousarray(P0, dtype=np.float64)
        L = int(types.shape[0])
        masks_d = torch.as_tensor(np.asarray(masks, dtype=np.uint64).view(np.int64)).to(self.device)
        N = int(masks_d.numel())
        x_f = torch.empty((N, sp.n), dtype=torch.float64, device=self.device)
        P_f = torch.empty((N, sp.n, sp.n), dtype=torch.float64, device=self.device) if want_P else None
        traj = torch.zeros((N, L + 2, 1 + traj_ncomp), dtype=torch.float64, device=self.device)
        ld = torch.full((N, L + 2), float("nan"), dtype=torch.float64, device=self.device)
        nt = torch.zeros((N,), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b200kf_run_masked(
                self._h, L, types.ctypes.data_as(POINTER(c_uint8)), times.ctypes.data_as(POINTER(c_double)),
                payload.ctypes.data_as(POINTER(c_double)), float(t_prev), float(t_end),
                x0.ctypes.data_as(POINTER(c_double)), P0.ctypes.data_as(POINTER(c_double)), N, _ptr(masks_d),
                _ptr(x_f), _ptr(P_f), _ptr(traj), int(traj_ncomp), _ptr(ld), _ptr(nt), _MODES[mode],
                self._stream(stream)))
        return {"x_final": x_f, "P_final": P_f, "traj": traj, "logdet": ld, "n_terms": nt}


    # ------------------------------------------------------------------ product search / accuracy
    def track_lerp(self, track_t, track_p, query_t, *, stream=None):
        """Track positions at query times, linear with extrapolation (b200kf_track_lerp; the
        interp1d calls of calculate_accuracy_metrics).  Returns torch CUDA f64 [n_query, k]."""
        torch = _torch()
        tt = self._dev(track_t, torch.float64)
        tp = self._dev(track_p, torch.float64)
        if tp.dim() == 1:
            tp = tp[:, None].contiguous()
        qt = self._dev(query_t, torch.float64)
        out = torch.empty((int(qt.numel()), int(tp.shape[1])), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b200kf_track_lerp(int(tt.numel()), _ptr(tt), _ptr(tp), int(tp.shape[1]),
                                                  int(qt.numel()), _ptr(qt), _ptr(out), self._stream(stream)))
        return out

    def rmse_vs_track(self, traj, ref_pos, *, stream=None):
        """Per-filter total_position_rmse of a time-major trajectory [T, ncomp>=3, N] against
        ref_pos [T, 3] (b200kf_rmse_vs_track).  Returns torch CUDA f64 [N]."""
        torch = _torch()
        tr = self._dev(traj, self.real)
        ref = self._dev(ref_pos, torch.float64)
        T, nc, N = (int(v) for v in tr.shape)
        if tuple(ref.shape) != (T, 3):
            raise ValueError(f"ref_pos must be [T={T}, 3]")
        out = torch.empty((N,), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b200kf_rmse_vs_track(self._h, N, T, _ptr(tr), nc, _ptr(ref), _ptr(out),
                                                     self._stream(stream)))
        return out

    def score_products(self, queue_len, types, times, payload, ref_pos, x0, P0, *, first=0, count=None,
                       want_per_combo=False, mode="auto", stream=None):
        """Score the sequences [first, first+count) of itertools.product(*queues) by RMSE against
        the track (b200kf_score_products).  types/times [E], payload [E, d], ref_pos [E, 3]: the
        queues' events flattened.  Returns (best: dict, per_combo: torch [count] or None)."""
        torch = _torch()
        sp = self.spec
        ql = np.ascontiguousarray(queue_len, dtype=np.int32)
        total = 1
        for q in ql:
            total *= int(q)
        if count is None:
            count = total - first
        types_d = self._dev(types, torch.uint8)
        times_d = self._dev(times, torch.float64)
        pay_d = self._dev(payload, torch.float64)
        ref_d = self._dev(ref_pos, torch.float64)
        E = int(ql.sum())
        if int(types_d.numel()) != E or int(times_d.numel()) != E or tuple(pay_d.shape) != (E, sp.d) \
                or tuple(ref_d.shape) != (E, 3):
            raise ValueError(f"need types/times [E={E}], payload [E, {sp.d}], ref_pos [E, 3]")
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        P0 = np.ascontiguousarray(P0, dtype=np.float64)
        per = torch.empty((count,), dtype=torch.float64, device=self.device) if want_per_combo else None
        best_d = torch.zeros((ctypes.sizeof(_lib.BestProduct),), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b200kf_score_products(
                self._h, int(ql.shape[0]), ql.ctypes.data_as(POINTER(c_int32)), _ptr(types_d), _ptr(times_d),
                _ptr(pay_d), _ptr(ref_d), x0.ctypes.data_as(POINTER(c_double)), P0.ctypes.data_as(POINTER(c_double)),
                int(first), int(count), _ptr(per), _ptr(best_d), _MODES[mode], self._stream(stream)))
        b = _lib.BestProduct.from_buffer_copy(bytes(best_d.cpu().numpy().tobytes()))
        none = (1 << 64) - 1
        best = {"score": b.score, "index": None if b.index == none else int(b.index),
                "runner_up_score": b.runner_up_score, "n_scored": int(b.n_scored)}
        return best, per

    def gain_record_bytes(self):
        n = ctypes.c_size_t(0)
        _lib.check(self.lib.b200kf_gain_record_bytes(self._h, byref(n)))
        return int(n.value)

    def select_greedy(self, types, times, periods, t_prev, P0, *, max_sel=None, mode="auto", stream=None,
                      want_gains=False):
        """Greedy one-event-per-window selection (b200kf_select_greedy) for I instances over one
        event list.  Returns (list of int32 numpy index arrays, P_final torch [I,n,n]); with
        want_gains=True a third item: the selected events' gain records, torch uint8 [I, max_sel, bytes]
        (None when P0 is not axis-symmetric: the records exist in the symmetric form only)."""
        torch = _torch()
        sp = self.spec
        types_d = self._dev(types, torch.uint8)
        times_d = self._dev(times, torch.float64)
        E = int(types_d.numel())
        per_d = self._dev(np.atleast_1d(periods), torch.float64)
        tp_d = self._dev(np.atleast_1d(t_prev), torch.float64)
        I = int(per_d.numel())
        if mode == "auto" and not isinstance(P0, torch.Tensor):
            # P0 still on the host: a chain-compatible covariance takes the per-chain form of the kernel
            # (an order of magnitude faster than the dense 15 x 15 one for the reference's P0, :851)
            from .models import axis_symmetric, chain_compatible, rates_axis_symmetric
            P0h = np.asarray(P0, dtype=np.float64).reshape(I, sp.n, sp.n)
            if axis_symmetric(sp.name, P0h) and rates_axis_symmetric(sp):
                mode = "symmetric"  # one covariance per group of axes (falls back to per-chain if the rates differ)
            elif chain_compatible(sp.name, P0h):
                mode = "decoupled"
        P0_d = self._dev(np.asarray(P0, dtype=np.float64).reshape(I, sp.n, sp.n), torch.float64)
        if max_sel is None:
            max_sel = E
        max_sel = max(1, int(max_sel))
        sel = torch.empty((I, max_sel), dtype=torch.int32, device=self.device)
        n_sel = torch.zeros((I,), dtype=torch.int64, device=self.device)
        P_f = torch.empty((I, sp.n, sp.n), dtype=torch.float64, device=self.device)
        recs = None
        if want_gains and mode == "symmetric":
            from .models import rates_axis_symmetric
            if rates_axis_symmetric(sp):  # the records exist in the symmetric form only; else: per-chain, no records
                recs = torch.empty((I, max_sel, self.gain_record_bytes()), dtype=torch.uint8, device=self.device)
            else:
                mode = "decoupled"
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b200kf_select_greedy_gains(
                self._h, E, _ptr(types_d), _ptr(times_d), I, _ptr(per_d), _ptr(tp_d), _ptr(P0_d), max_sel,
                _ptr(sel), _ptr(n_sel), _ptr(P_f), _ptr(recs), _MODES[mode], self._stream(stream)))
        ns = n_sel.cpu().numpy()
        sel_h = sel.cpu().numpy()
        sels = [sel_h[i, :min(int(ns[i]), max_sel)].copy() for i in range(I)]
        return (sels, P_f, recs) if want_gains else (sels, P_f)

    def scheduled_sweep_pipelined(self, types, times, rates, x0, P0, t_prev, payload_chunk, chunk_events, *,
                                  payload_depth=1):
        """scheduled_sweep_chunked with the selection OFF the critical path: the greedy selection is itself walked
        in time chunks (b200kf_select_greedy_chunk, one warp per rate, on its own stream) one chunk AHEAD of the
        filters, which consume its gain records as pure streams (b200kf_run_gains).  Per chunk the host waits only
        for that chunk's twelve selection counts (a 96-byte pinned copy).  `payload_chunk(e_lo, e1)` is asked for
        ranges that reach back one window period before the chunk: an event queued before the boundary may be selected
        after it.  The reference's P0 shape (axis-symmetric, kf_workers.py:851) is assumed, as for use_gains.
        payload_depth: how many consecutive chunks' payload may be alive at once -- 1 (default): `payload_chunk` may
        reuse one buffer, so a chunk's launches all finish before the next chunk's payload is asked for; k > 1: the
        callback cycles through k buffers (the caller's promise), and the rates' streams run up to k - 1 chunks apart
        instead of meeting at every chunk boundary (the highest rate's launch is the longest of every chunk).
        Returns {rate: dict(sel (device), n_sel, x, P, agg)}."""
        return _sweep_pipelined(self, types, times, rates, x0, P0, t_prev, payload_chunk, chunk_events, payload_depth)

    def select_state(self, types, times, periods, t_prev, P0, *, max_sel=None, want_gains=True):
        """A chunked greedy selection (b200kf_select_greedy_chunk): returns a SelectState whose advance(e_end)
        enqueues the walk over the next events; selections, covariances and gain records come out bit for bit as
        from one select_greedy call.  The axis-symmetric form only (P0 as the reference's, kf_workers.py:851)."""
        return SelectState(self, types, times, periods, t_prev, P0, max_sel, want_gains)

    def run_gains(self, gain_records, payload, x, *, want_agg=True, traj_ncomp=0, want_logdet=False, agg=None,
                  stream=None, inplace=False, step_rows=None, small_tiles=False):
        """N filters that share a schedule and a covariance, from that schedule's precomputed gain
        records (b200kf_run_gains): gain_records torch uint8 [T, bytes] (one instance's slice of
        select_greedy(..., want_gains=True)), payload [T, d, N], x [N, n].  Returns dict of torch CUDA
        tensors x, agg [N, 4], traj, logdet, status.  small_tiles: 128 filters per CTA (B200KF_RUN_SMALL_TILES) for a
        launch that runs beside others and carries the longest schedule of them."""
        torch = _torch()
        sp = self.spec
        rec = gain_records.contiguous()
        T = int(rec.shape[0])
        pay_d = self._dev(payload, self.real)
        x_d = self._dev(x, self.real)
        if not inplace and isinstance(x, torch.Tensor) and x_d.data_ptr() == x.data_ptr():
            x_d = x_d.clone()
        N = int(x_d.shape[0])
        rows_d = None
        if step_rows is not None:
            rows_d = self._dev(step_rows, torch.int32)
            if int(rows_d.numel()) != T or tuple(pay_d.shape[1:]) != (sp.d, N):
                raise ValueError(f"step_rows must be [T={T}] and payload [rows, d={sp.d}, N={N}]")
        elif T and tuple(pay_d.shape) != (T, sp.d, N):
            raise ValueError(f"payload must be [T={T}, d={sp.d}, N={N}], got {tuple(pay_d.shape)}")
        flags = _lib.RUN_SMALL_TILES if small_tiles else 0
        agg_d = None
        if agg is not None:
            agg_d = self._dev(agg, torch.float64)
            flags |= _lib.RUN_AGG_CONTINUE
        elif want_agg:
            agg_d = torch.empty((N, 4), dtype=torch.float64, device=self.device)
        traj_d = torch.empty((T, traj_ncomp, N), dtype=self.real, device=self.device) if traj_ncomp else None
        ld_d = torch.empty((T, N), dtype=self.real, device=self.device) if want_logdet else None
        status_d = torch.zeros((N,), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b200kf_run_gains(self._h, N, T, _ptr(rec), _ptr(pay_d), _ptr(rows_d), _ptr(x_d), _ptr(agg_d), _ptr(traj_d),
                                                 int(traj_ncomp), _ptr(ld_d), None, _ptr(status_d), flags,
                                                 self._stream(stream)))
        return {"x": x_d, "agg": agg_d, "traj": traj_d, "logdet": ld_d, "status": status_d}


    def scheduled_sweep(self, types, times, payload, rates, x0, P0, t_prev, *, traj_ncomp=0, mode="auto",
                        concurrent=True, sels=None, use_gains=False):
        """Sampling-rate sweep x Monte-Carlo seeds (BASELINE configs[1]; run_kalman_filter_scheduled
        with selection_method='greedy', kf_workers.py:826-957, for many seeds at once).

        types/times [E]: the base event stream AFTER the initialising event; payload [E, d, N]: each
        seed's measurements of those events; rates: processing frequencies (Hz); x0 [N, n], P0 [n, n]
        and t_prev: state at the initialising event.  The greedy selection never reads a measurement,
        so it runs once per rate (all rates in one covariance-only launch); the N seeds of a rate then
        share the selected schedule AND the covariance (same P0): one b200kf_run_batch per rate, which
        AUTO resolves to the shared-covariance kernel.  A rate's batch of a few thousand seeds fills
        only part of the GPU, so with concurrent=True every rate runs on its own stream (and its own
        model object: the per-model scratch is not shared between in-flight launches).  Returns
        {rate: dict(sel, times, x, P, agg, traj)}; `sels={rate: sel}` skips the selection pass (further
        batches of seeds over the same stream).  With use_gains (and an axis-symmetric P0) the selection pass
        also publishes the selected events' gains, and each rate's batch runs as a pure stream
        (b200kf_run_gains) with no covariance recursion in the launch."""
        torch = _torch()
        sp = self.spec
        rates = list(rates)
        types_d = self._dev(types, torch.uint8)
        times_d = self._dev(times, torch.float64)
        pay_d = self._dev(payload, self.real)
        N = int(pay_d.shape[2])
        P0 = np.asarray(P0, dtype=np.float64)
        recs = P_fin = None
        if sels is None:  # pass the `sel` arrays of an earlier call to reuse them: they depend on (events, rates, P0) only
            got = self.select_greedy(types_d, times_d, [1.0 / f for f in rates], [t_prev] * len(rates),
                                     np.repeat(P0[None], len(rates), axis=0), want_gains=use_gains)
            sels, P_fin, recs = got if use_gains else (got[0], got[1], None)
        else:
            sels = [np.asarray(sels[f]) for f in rates]
        if recs is not None and (N * (4 if sp.precision == "f32" else 8)) % 16 != 0:
            recs = None  # bulk-copy rows need 16-byte alignment: odd batches take the recursion-per-launch path
        x0_d = self._dev(x0, self.real)
        P0_d = self._dev(np.repeat(P0[None], N, axis=0), self.real)
        engines = [self]
        side = []
        if concurrent and len(rates) > 1:
            pool = getattr(self, "_sweep_engines", [])
            while len(pool) < len(rates) - 1:
                pool.append(Engine(self.spec, device=self.device.index))
            self._sweep_engines = pool
            engines = [self] + pool[:len(rates) - 1]
            spool = getattr(self, "_sweep_streams", [])  # reused: the caching allocator keeps blocks per stream
            while len(spool) < len(rates):
                spool.append(torch.cuda.Stream(device=self.device))
            self._sweep_streams = spool
            side = spool[:len(rates)]
        cur = torch.cuda.current_stream(self.device)
        out = {}
        # longest schedule first: its launch is the critical path, the host work of the others overlaps it
        order = sorted(range(len(rates)), key=lambda j: -len(sels[j]))
        for i in order:
            f, sel = rates[i], sels[i]
            eng = engines[i] if side else self
            st_ = side[i] if side else cur
            if side:
                st_.wait_stream(cur)
            with torch.cuda.stream(st_):
                sel_d = torch.from_numpy(sel.astype(np.int64)).to(self.device)
                st = times_d.index_select(0, sel_d)
                dts = torch.diff(st, prepend=torch.tensor([t_prev], dtype=torch.float64, device=self.device))
                rows = sel_d.to(torch.int32)  # the kernels read the selected rows in place: no gathered payload copy
                if recs is not None:  # the selection pass already walked this covariance recursion: pure stream
                    r = eng.run_gains(recs[i, :len(sel)], pay_d, x0_d, traj_ncomp=traj_ncomp, stream=st_, step_rows=rows,
                                      small_tiles=bool(side) and _small_tiles(len(sel), max(len(v) for v in sels)))
                    r["P"] = P_fin[i].to(self.real).expand(N, sp.n, sp.n)
                else:
                    r = eng.run_batch(types_d.index_select(0, sel_d), dts, pay_d, x0_d, P0_d, traj_ncomp=traj_ncomp,
                                      mode=mode, stream=st_, step_rows=rows)
            r["sel"] = sel
            r["times"] = st
            out[f] = r
        for st_ in side:
            cur.wait_stream(st_)
        return {f: out[f] for f in rates}


    def scheduled_sweep_chunked(self, types, times, rates, x0, P0, t_prev, payload_chunk, chunk_events, *,
                                use_gains=False, on_chunk=None):
        """scheduled_sweep over a stream whose per-seed payload does not fit the device at once (BASELINE
        configs[1] at full length: 589 679 events x 9 channels x 4096 seeds = 174 GB): ONE selection pass over
        the whole stream (it never reads a measurement), then time chunks of `chunk_events` events;
        `payload_chunk(e0, e1)` returns that chunk's payload [e1 - e0, d, N] on the device (the caller
        generates / uploads it), and every rate's filters carry x, P and the log-det aggregates from chunk to
        chunk.  Returns {rate: dict(sel, x, P, agg)}; `on_chunk(e0, e1, fn)` -- if given -- is called with the
        chunk's launch work as a callable so that a caller can time the launches apart from the payload
        generation."""
        torch = _torch()
        sp = self.spec
        rates = list(rates)
        types = np.ascontiguousarray(types, dtype=np.uint8)
        times = np.ascontiguousarray(times, dtype=np.float64)
        E = len(types)
        types_d = self._dev(types, torch.uint8)
        times_d = self._dev(times, torch.float64)
        P0 = np.asarray(P0, dtype=np.float64)
        got = self.select_greedy(types_d, times_d, [1.0 / f for f in rates], [t_prev] * len(rates),
                                 np.repeat(P0[None], len(rates), axis=0), want_gains=use_gains)
        sels, P_fin, recs = got if use_gains else (got[0], got[1], None)
        x0_d = self._dev(x0, self.real)
        N = int(x0_d.shape[0])
        if recs is not None and (N * (4 if sp.precision == "f32" else 8)) % 16 != 0:
            recs = None
        pool = getattr(self, "_sweep_engines", [])
        while len(pool) < len(rates) - 1:
            pool.append(Engine(self.spec, device=self.device.index))
        self._sweep_engines = pool
        engines = [self] + pool[:len(rates) - 1]
        spool = getattr(self, "_sweep_streams", [])
        while len(spool) < len(rates):
            spool.append(torch.cuda.Stream(device=self.device))
        self._sweep_streams = spool
        cur = torch.cuda.current_stream(self.device)
        sel_d = [torch.from_numpy(sel.astype(np.int64)).to(self.device) for sel in sels]
        seltimes = [times[sel] for sel in sels]
        st = [{"x": x0_d.clone(), "P": None if recs is not None else self._dev(np.repeat(P0[None], N, axis=0), self.real),
               "agg": None, "done": 0, "prev": float(t_prev)} for _ in rates]
        order = sorted(range(len(rates)), key=lambda j: -len(sels[j]))
        for e0 in range(0, E, int(chunk_events)):
            e1 = min(E, e0 + int(chunk_events))
            pay_d = payload_chunk(e0, e1)

            def launches(e0=e0, e1=e1, pay_d=pay_d):
                longest = max(int(np.searchsorted(sels[i], e1, side="left")) - st[i]["done"] for i in range(len(rates)))
                for i in order:
                    s_i = st[i]
                    j0 = s_i["done"]
                    j1 = int(np.searchsorted(sels[i], e1, side="left"))  # selected events inside this chunk
                    if j1 == j0:
                        continue
                    stream = spool[i]
                    stream.wait_stream(cur)
                    with torch.cuda.stream(stream):
                        rows = (sel_d[i][j0:j1] - e0).to(torch.int32)
                        if recs is not None:
                            r = engines[i].run_gains(recs[i, j0:j1], pay_d, s_i["x"], agg=s_i["agg"], stream=stream,
                                                     step_rows=rows, inplace=True, small_tiles=_small_tiles(j1 - j0, longest))
                        else:
                            tsel = seltimes[i][j0:j1]
                            dts = np.diff(tsel, prepend=s_i["prev"])  # float64 on the host, as the reference (:914)
                            r = engines[i].run_batch(types[sels[i][j0:j1]], dts, pay_d, s_i["x"], s_i["P"], agg=s_i["agg"],
                                                     stream=stream, step_rows=rows, inplace=True)
                            s_i["P"] = r["P"]
                        s_i["x"], s_i["agg"] = r["x"], r["agg"]
                        s_i["prev"] = float(seltimes[i][j1 - 1])
                        s_i["done"] = j1
                for stream in spool[:len(rates)]:
                    cur.wait_stream(stream)

            if on_chunk is not None:
                on_chunk(e0, e1, launches)
            else:
                launches()
            pay_d.record_stream(cur)
            del pay_d
        out = {}
        for i, f in enumerate(rates):
            P = st[i]["P"] if recs is None else P_fin[i].to(self.real).expand(N, sp.n, sp.n)
            out[f] = {"sel": sels[i], "x": st[i]["x"], "P": P, "agg": st[i]["agg"]}
        return out


def _small_tiles(steps, longest):
    """Concurrent gain-applying launches of a sweep: the schedules above half the longest one take 128-filter CTAs
    (B200KF_SWEEP_SMALL_TILES=0 disables, =all applies it to every launch)."""
    mode = os.environ.get("B200KF_SWEEP_SMALL_TILES", "1")
    if mode == "0":
        return False
    return mode == "all" or 2 * steps > longest


def _sweep_pipelined(self, types, times, rates, x0, P0, t_prev, payload_chunk, chunk_events, payload_depth=1):
    """Engine.scheduled_sweep_pipelined (defined below the class for length)."""
    torch = _torch()
    sp = self.spec
    rates = list(rates)
    R = len(rates)
    types = np.ascontiguousarray(types, dtype=np.uint8)
    times = np.ascontiguousarray(times, dtype=np.float64)
    E = len(types)
    bounds = list(range(0, E, int(chunk_events))) + [E]
    nch = len(bounds) - 1
    P0 = np.asarray(P0, dtype=np.float64)
    st = self.select_state(types, times, [1.0 / f for f in rates], [t_prev] * R, np.repeat(P0[None], R, axis=0),
                           want_gains=True)
    x0_d = self._dev(x0, self.real)
    N = int(x0_d.shape[0])
    if (N * (4 if sp.precision == "f32" else 8)) % 16 != 0:
        raise ValueError("the pipelined sweep streams the payload with bulk copies: N must make 16-byte rows")
    pool = getattr(self, "_sweep_engines", [])
    while len(pool) < R - 1:
        pool.append(Engine(self.spec, device=self.device.index))
    self._sweep_engines = pool
    engines = [self] + pool[:R - 1]
    spool = getattr(self, "_sweep_streams", [])
    while len(spool) < R:
        spool.append(torch.cuda.Stream(device=self.device))
    self._sweep_streams = spool
    sel_stream = getattr(self, "_select_stream", None)
    if sel_stream is None:
        sel_stream = self._select_stream = torch.cuda.Stream(device=self.device)
    cur = torch.cuda.current_stream(self.device)
    nsel_host = torch.zeros((nch, R), dtype=torch.int64).pin_memory()
    evs = [torch.cuda.Event() for _ in range(nch)]
    sel_stream.wait_stream(cur)

    def enqueue_selection(c):
        with torch.cuda.stream(sel_stream):
            st.advance(bounds[c + 1], stream=sel_stream)
            nsel_host[c].copy_(st.n_sel, non_blocking=True)
            evs[c].record(sel_stream)

    enqueue_selection(0)
    max_period = 1.0 / min(rates)
    xs = [x0_d.clone() for _ in rates]
    aggs = [None] * R
    done = [0] * R
    depth = max(1, int(payload_depth))
    chunk_done = []  # per chunk: one event per launch stream, recorded behind the chunk's launch on it
    order = sorted(range(R), key=lambda j: -rates[j])  # the highest rate has the longest launches: first
    if os.environ.get("B200KF_SWEEP_ORDER"):  # diagnostic: launch order of a chunk's rates, e.g. "50,60,...,120,40,...,10"
        want = [float(v) for v in os.environ["B200KF_SWEEP_ORDER"].split(",")]
        order = [rates.index(v) for v in want if v in rates]
        order += [j for j in range(R) if j not in order]
    for c in range(nch):
        if c + 1 < nch:
            enqueue_selection(c + 1)  # runs while this chunk's filters stream
        e0, e1 = bounds[c], bounds[c + 1]
        # A window that closes in this chunk may select an event queued before e0 (the open queue at the chunk
        # boundary); queue entries are younger than one period, so the payload rows reach back that far
        e_lo = max(0, int(np.searchsorted(times, times[e0 - 1] - max_period, side="left")) - 1) if e0 > 0 else 0
        if c >= depth:  # the buffer this chunk's payload lands in was last read by chunk c - depth
            for ev_ in chunk_done[c - depth]:
                cur.wait_event(ev_)
        pay_d = payload_chunk(e_lo, e1)
        evs[c].synchronize()  # this chunk's selection counts are on the host
        cur.wait_event(evs[c])
        longest = max(int(nsel_host[c, i]) - done[i] for i in range(R))
        for i in order:
            j0, j1 = done[i], int(nsel_host[c, i])
            if j1 == j0:
                continue
            stream = spool[i]
            stream.wait_stream(cur)
            with torch.cuda.stream(stream):
                rows = st.sel[i, j0:j1] - e_lo
                # the chunk lasts as long as its longest schedule's launch: those get twice the CTAs
                r = engines[i].run_gains(st.recs[i, j0:j1], pay_d, xs[i], agg=aggs[i], stream=stream, step_rows=rows,
                                         inplace=True, small_tiles=_small_tiles(j1 - j0, longest))
            xs[i], aggs[i] = r["x"], r["agg"]
            done[i] = j1
        evd = []
        for stream in spool[:R]:
            ev_ = torch.cuda.Event()
            ev_.record(stream)
            evd.append(ev_)
        chunk_done.append(evd)
        del pay_d
    for stream in spool[:R]:
        cur.wait_stream(stream)
    cur.wait_stream(sel_stream)
    out = {}
    for i, f in enumerate(rates):
        out[f] = {"sel": st.sel[i, :done[i]], "n_sel": done[i], "x": xs[i], "agg": aggs[i],
                  "P": st.P[i].to(self.real).expand(N, sp.n, sp.n)}
    return out


class SelectState:
    """State of a chunked greedy selection on the device: event arrays, per-instance carry, covariance,
    selection list, counts and (optionally) gain records."""

    def __init__(self, eng, types, times, periods, t_prev, P0, max_sel, want_gains):
        torch = _torch()
        self.eng = eng
        sp = eng.spec
        self.types = eng._dev(types, torch.uint8)
        self.times = eng._dev(times, torch.float64)
        self.E = int(self.types.numel())
        self.periods = eng._dev(np.atleast_1d(periods), torch.float64)
        self.I = int(self.periods.numel())
        self.t_prev = eng._dev(np.broadcast_to(np.atleast_1d(t_prev), (self.I,)).astype(np.float64), torch.float64)
        self.P0 = eng._dev(np.asarray(P0, dtype=np.float64).reshape(self.I, sp.n, sp.n), torch.float64)
        self.max_sel = int(self.E if max_sel is None else max_sel)
        dev = eng.device
        self.carry = torch.zeros((self.I, 8), dtype=torch.float64, device=dev)
        self.P = torch.empty((self.I, sp.n, sp.n), dtype=torch.float64, device=dev)
        self.sel = torch.empty((self.I, max(1, self.max_sel)), dtype=torch.int32, device=dev)
        self.n_sel = torch.zeros((self.I,), dtype=torch.int64, device=dev)
        self.recs = (torch.empty((self.I, max(1, self.max_sel), eng.gain_record_bytes()), dtype=torch.uint8, device=dev)
                     if want_gains else None)
        self.done = 0

    def advance(self, e_end, stream=None):
        """Walk events [done, e_end) on `stream` (default: the current stream)."""
        e_end = min(int(e_end), self.E)
        eng = self.eng
        torch = _torch()
        with torch.cuda.device(eng.device):
            _lib.check(eng.lib.b200kf_select_greedy_chunk(
                eng._h, self.done, e_end, _ptr(self.types), _ptr(self.times), self.I, _ptr(self.periods), _ptr(self.t_prev),
                _ptr(self.P0), _ptr(self.carry), _ptr(self.P), max(1, self.max_sel), _ptr(self.sel), _ptr(self.n_sel),
                _ptr(self.recs), _MODES["symmetric"], eng._stream(stream)))
        self.done = e_end
        return self


def cov_update(Sigma, rows, H, R, device=None, stream=None):
    """Batched Scheduler.cov_matrix (b200kf_cov_update).  Sigma [B, n, n] (numpy or torch), rows:
    0-based selected rows of (H, R).  Returns (Sigma_post [B,n,n], trace [B]) torch CUDA f64."""
    torch = _require_cuda()
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    S = Sigma if isinstance(Sigma, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(Sigma))
    S = S.to(device=dev, dtype=torch.float64).contiguous()
    B, n = int(S.shape[0]), int(S.shape[-1])
    H = np.ascontiguousarray(H, dtype=np.float64)
    R = np.ascontiguousarray(R, dtype=np.float64)
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    post = torch.empty_like(S)
    tr = torch.empty((B,), dtype=torch.float64, device=dev)
    s = torch.cuda.current_stream(dev) if stream is None else stream
    with torch.cuda.device(dev):
        _lib.check(lib.b200kf_cov_update(
            n, int(H.shape[0]), int(rows.shape[0]), rows.ctypes.data_as(POINTER(c_int32)),
            H.ctypes.data_as(POINTER(c_double)), R.ctypes.data_as(POINTER(c_double)), B, _ptr(S),
            _ptr(post), _ptr(tr), c_void_p(s.cuda_stream)))
    return post, tr


def fill_payload(engine: Engine, types_dev, n_filters, seed=1234, stream=None):
    """Device-side synthetic payload [T, d, N] of BASELINE config 3/5 (b200kf_fill_payload)."""
    torch = _torch()
    T = int(types_dev.numel())
    out = torch.empty((T, engine.spec.d, n_filters), dtype=engine.real, device=engine.device)
    with torch.cuda.device(engine.device):
        _lib.check(engine.lib.b200kf_fill_payload(engine._h, int(n_filters), T, _ptr(types_dev), int(seed),
                                                  _ptr(out), engine._stream(stream)))
    return out


def probe_fma_peak(precision="f64", distinct_operands=False) -> float:
    """Measured FMA peak of the current device in TFLOP/s (b200kf_probe_fma_peak).  With
    distinct_operands every FMA reads three different, changing registers."""
    _require_cuda()
    lib = _lib.load()
    v = c_double(0.0)
    code = (_lib.F32 if precision == "f32" else _lib.F64) | (2 if distinct_operands else 0)
    _lib.check(lib.b200kf_probe_fma_peak(code, byref(v)))
    return v.value
