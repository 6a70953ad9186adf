"""CPU-only tests of the host logic: model extraction from the getters, stream packing, dt rules,
sharding / reduce helpers."""
import numpy as np
import pytest

from oracle import kf_oracle as ko
from sensorfusion_kalmanfilter_b200 import models, streams, synth
from sensorfusion_kalmanfilter_b200.distributed import merge_best, shard_range


class _FromOracle:
    """Getter object built from an oracle model (same formulas as the reference's getters)."""

    def __init__(self, name):
        self.m = ko.get_model(name)

    def get_state_transition_matrix(self, dt):
        return self.m.F(dt)

    def get_process_noise_covariance_matrix(self, dt):
        return self.m.Q(dt)

    def predict_covariance(self, P, F, Q):
        return F @ P @ F.T + Q

    def get_gps_observation_matrix(self):
        return self.m.H_gps()

    def get_gps_measurement_noise_covariance_matrix(self):
        return self.m.R_gps()

    def get_imu_observation_matrix(self):
        return self.m.H_imu()

    def get_imu_measurement_noise_covariance_matrix(self):
        return self.m.R_imu()

    def calculate_kalman_gain(self, P, H, R):
        return P @ H.T @ np.linalg.inv(H @ P @ H.T + R)


@pytest.mark.parametrize("name", ["CA15", "CA8", "CV6"])
def test_spec_from_getters_recovers_reference_constants(name):
    spec = models.spec_from_getters(_FromOracle(name), name)
    dflt = models.default_spec(name)
    np.testing.assert_allclose(spec.q, dflt.q, rtol=1e-12)
    np.testing.assert_allclose(spec.r_imu, dflt.r_imu, rtol=1e-12)
    np.testing.assert_allclose(spec.r_gps, dflt.r_gps, rtol=1e-12)
    assert abs(spec.accel_coef - 0.5) < 1e-12 or name == "CV6"


def test_spec_from_getters_rejects_models_outside_the_family():
    g = _FromOracle("CA15")
    F0 = g.get_state_transition_matrix

    def F_coupled(dt):
        F = F0(dt)
        F[0, 4] = dt
        return F
    g.get_state_transition_matrix = F_coupled
    with pytest.raises(NotImplementedError):
        models.spec_from_getters(g, "CA15")
    g = _FromOracle("CA15")
    g.get_gps_measurement_noise_covariance_matrix = lambda: np.array([[3, 1, 0], [1, 3, 0], [0, 0, 3.0]])
    with pytest.raises(NotImplementedError):
        models.spec_from_getters(g, "CA15")
    g = _FromOracle("CA15")
    g.calculate_kalman_gain = lambda P, H, R: 0.5 * (P @ H.T @ np.linalg.inv(H @ P @ H.T + R))
    with pytest.raises(NotImplementedError):
        models.spec_from_getters(g, "CA15")


def test_chain_compatible():
    for name in ("CV6", "CA8", "BIAS9", "CA15", "CVBU6"):
        m = ko.get_model(name)
        assert models.chain_compatible(name, m.P0())
        P = m.P0().copy()
        P[0, 1] = P[1, 0] = 1.0
        assert not models.chain_compatible(name, P)
        # the oracle never couples chains: after a few dense steps the off-chain entries stay 0
        x, P = np.zeros(m.n), m.P0()
        rng = np.random.default_rng(0)
        for k in range(30):
            x, P = ko.step(m, x, P, 0 if k % 5 == 0 else 1, 0.01, rng.normal(size=m.d))
        assert models.chain_compatible(name, P)


def test_packer_matches_oracle_payload():
    ev = synth.make_indexed_sensor_data(120, seed=3)
    ps = streams.pack_events(ev)
    assert streams.pack_events(ev) is ps  # cached
    for name in ("CA15", "CA8", "CV6", "BIAS9", "CVBU6"):
        m = ko.get_model(name)
        pay = ps.payload(name)
        for k in (0, 3, 4, 57, 119):
            np.testing.assert_array_equal(pay[k], ko.event_payload(m, ev[k][1], ev[k][3]))


def test_packer_cache_is_identity_safe():
    """A regenerated list of the same shape (same length and end times, different measurements) must be
    packed afresh even when CPython hands it the id of the freed one, and so must an in-place edit."""
    import gc
    streams.invalidate()
    stale = 0
    for seed in range(20):
        ev = synth.make_indexed_sensor_data(90, seed=seed)
        ps = streams.pack_events(ev)
        want = streams.PackedStream.from_events(ev)
        stale += not (np.array_equal(ps.data, want.data) and np.array_equal(ps.times, want.times))
        del ev
        gc.collect()
    assert stale == 0
    ev = synth.make_indexed_sensor_data(90, seed=1)
    ps = streams.pack_events(ev)
    k = next(i for i, e in enumerate(ev) if e[1] == "GPS")
    g = dict(ev[k][3])
    g["easting"] += 1.0
    ev[k] = (ev[k][0], ev[k][1], ev[k][2], g)  # in-place edit of one event (every event is a probe at this size)
    ps2 = streams.pack_events(ev)
    assert ps2 is not ps and ps2.data[k, 0] == g["easting"]
    streams.invalidate(ev)
    assert streams.pack_events(ev) is not ps2


def test_schedule_dts_rules():
    t = np.array([10.0, 10.5, 10.2, 10.6, 11.0])
    keep, dts, prev = streams.schedule_dts(t, 10.0, skip_advances_prev=True)  # kf_workers.py:682-685
    np.testing.assert_array_equal(keep, [0, 1, 3, 4])
    np.testing.assert_allclose(dts, [0.0, 0.5, 10.6 - 10.2, 11.0 - 10.6])
    keep, dts, prev = streams.schedule_dts(t, 10.0, skip_advances_prev=False)  # :1013-1015
    np.testing.assert_array_equal(keep, [0, 1, 3, 4])
    np.testing.assert_allclose(dts, [0.0, 0.5, 10.6 - 10.5, 11.0 - 10.6])
    keep, dts, prev = streams.schedule_dts(np.array([1.0, 2.0]), 0.5)
    assert prev == 2.0 and list(keep) == [0, 1]


def test_shard_range_and_merge():
    for total, world in ((1 << 20, 8), (1023, 5), (7, 8)):
        got = [shard_range(total, world, r) for r in range(world)]
        assert got[0][0] == 0 and sum(c for _, c in got) == total
        for (f0, c0), (f1, _) in zip(got, got[1:]):
            assert f1 == f0 + c0
    a = {"argmin_sum_score": 1.0, "argmin_sum_index": 9, "first_feasible_key": 50, "first_feasible_index": 3,
         "n_feasible": 2, "n_near_threshold": 0}
    b = {"argmin_sum_score": 1.0, "argmin_sum_index": 4, "first_feasible_key": (1 << 64) - 1,
         "first_feasible_index": None, "n_feasible": 0, "n_near_threshold": 1}
    c = {"argmin_sum_score": float("inf"), "argmin_sum_index": None, "first_feasible_key": 20,
         "first_feasible_index": 77, "n_feasible": 5, "n_near_threshold": 0}
    m = merge_best([a, b, c])
    assert m["argmin_sum_index"] == 4 and m["first_feasible_index"] == 77 and m["n_feasible"] == 7
    assert m["n_near_threshold"] == 1
    assert merge_best([c, b, a]) == m


def test_axis_symmetric():
    from sensorfusion_kalmanfilter_b200.models import axis_symmetric, default_spec
    for name in ("CV6", "CA8", "BIAS9", "CA15", "CVBU6"):
        P = default_spec(name).default_P0()
        assert axis_symmetric(name, P) and axis_symmetric(name, np.stack([P, 2 * P]))
        Q = P.copy()
        Q[1, 1] *= 1.0000001  # y differs from x: per-chain covariances needed
        assert not axis_symmetric(name, Q)
        Q = P.copy()
        Q[0, 1] = Q[1, 0] = 0.5  # couples two chains: not even chain-compatible
        assert not axis_symmetric(name, Q)


def test_shard_strided():
    """Sharded subset search: power-of-two worlds fix the log2(world) earliest decisions (rank = their value);
    other world sizes and windows shorter than that keep contiguous ranges."""
    from sensorfusion_kalmanfilter_b200.distributed import shard_strided
    for L, W in ((32, 8), (20, 2), (26, 4), (3, 8)):
        got = [shard_strided(L, W, r) for r in range(W)]
        n_low = W.bit_length() - 1
        assert got == [(r, n_low) for r in range(W)]
        covered = sorted(low | (j << n_low) for low, _ in got for j in range(1 << min(L - n_low, 6)))
        assert covered == list(range(len(covered))) or L - n_low > 6
    for L, W in ((32, 1), (32, 3), (32, 6), (2, 8)):
        assert all(shard_strided(L, W, r) is None for r in range(W))
