"""BASELINE configs[0] on the reference's ONE real input, gps_data.csv (SURVEY.md section 4 data pins:
30 758 rows, 21 871 valid fixes, first valid row 2 735, first fix at 1697739552.3362827, zone 19 T,
altitude -32.6).  The log travels as four float64 columns (tests/golden/gps_data_reference.npz, written by
oracle/make_golden_long.py cfg1); synth.write_gps_csv rebuilds the CSV byte for byte (sha256 checked).
The IMU log is not in the reference's repository, so it is synthesised to the hw5_1.py:29-31 schema.
The GPU test runs the whole front end + both filters and holds them to the UNMODIFIED reference's
outputs on the same files (tests/golden/config1_real_gps.npz)."""
import hashlib
import os

import numpy as np
import pytest

from sensorfusion_kalmanfilter_b200 import synth

REF_CSV = "/root/reference/gps_data.csv"


@pytest.fixture(scope="module")
def real_gps(golden, tmp_path_factory):
    g = golden("gps_data_reference.npz")
    d = tmp_path_factory.mktemp("cfg1")
    path = synth.write_gps_csv(str(d / "gps_data.csv"), g["time"], g["lat"], g["lon"], g["alt"])
    return g, path, d


def _front_end(mod, gps_path, imu_path):
    k = mod.KF_SensorFusion(gps_path, imu_path)
    k.load_data()
    k.gps_to_modified_utm() if hasattr(k, "gps_to_modified_utm") else k.gps_to_utm()
    wb, ab, fvi = k.compute_imu_biases(k.gps_data, k.imu_data)
    k.unbias_imu_data(wb, ab)
    k.combine_sensor_data()
    return k, wb, ab, fvi


def test_gps_log_rebuilds_byte_for_byte_and_loader_pins(real_gps):
    from sensorfusion_kalmanfilter_b200 import kf_workers
    g, path, d = real_gps
    raw = open(path, "rb").read()
    assert len(raw) == int(g["n_bytes"])
    assert hashlib.sha256(raw).digest() == g["sha256"].tobytes()
    if os.path.isfile(REF_CSV):  # build container only: the file itself
        assert open(REF_CSV, "rb").read() == raw
    k = kf_workers.KF_SensorFusion(path, path)
    k.gps_data = k.load_data_from_csv(path)
    assert len(k.gps_data) == 30758
    k.gps_to_modified_utm()
    assert len(k.utm_data) == 21871
    first_valid = next(i for i, e in enumerate(k.gps_data) if "nan" not in e[1].lower())
    assert first_valid == 2735
    u0 = k.utm_data[0]
    assert u0["time"] == 1697739552.3362827  # the constant kf_workers.py:1993 hard-codes
    assert (u0["zone_number"], u0["zone_letter"]) == (19, "T") and u0["altitude"] == -32.6
    assert u0["easting"] == 0.0 and u0["northing"] == 0.0  # relative to the first fix (kf_workers.py:321-327)
    assert abs(np.median(np.diff(g["time"])) - 0.0999) < 2e-4  # ~10 Hz
    assert abs(g["time"][0] - 1697739278.76) < 0.01 and abs(g["time"][-1] - 1697742360.93) < 0.01  # 3 082 s


@pytest.mark.gpu
def test_config1_end_to_end_on_the_real_gps_log_vs_reference(real_gps, golden):
    from sensorfusion_kalmanfilter_b200 import hw5_2, kf_workers
    g, gps_path, d = real_gps
    c = golden("config1_real_gps.npz")
    stride = int(c["stride"])
    _, imu_path = synth.write_config1_csvs(str(d), n_imu=int(c["imu_rows"]), seed=0, gps_csv=gps_path)

    def sample(a):
        a = np.asarray(a, dtype=np.float64)
        return np.concatenate([a[:-8:stride], a[-8:]], axis=0)

    def rel(a, b):
        return np.max(np.abs(np.asarray(a) - b)) / max(np.max(np.abs(b)), 1e-300)

    for tag, mod in (("kw", kf_workers), ("hw", hw5_2)):
        k, wb, ab, fvi = _front_end(mod, gps_path, imu_path)
        assert len(k.indexed_sensor_data) == int(c[f"{tag}_n_events"]) and len(k.utm_data) == int(c[f"{tag}_n_fixes"])
        assert fvi == int(c[f"{tag}_first_valid"]) == 2735
        np.testing.assert_array_equal(wb, c[f"{tag}_wbias"])
        np.testing.assert_array_equal(ab, c[f"{tag}_abias"])
        assert k.utm_data[0]["time"] == float(c[f"{tag}_first_fix_time"])
        if tag == "kw":
            st, ld, P, prev = k.run_kalman_filter_full(0, len(k.indexed_sensor_data))
            assert len(st) == int(c["kw_n_states"])
            st = sample(np.array([[float(v) for v in s] for s in st]))
            np.testing.assert_array_equal(st[:, 0], c["kw_states"][:, 0])
            assert rel(st[:, 1:], c["kw_states"][:, 1:]) < 1e-9
            lda = np.array(ld, dtype=np.float64)
            assert np.max(np.abs(sample(lda) - c["kw_logdet"]) / np.maximum(1.0, np.abs(c["kw_logdet"]))) < 1e-9
            agg = np.array([lda.max(), lda.sum(), lda[-1]])
            assert np.max(np.abs(agg - c["kw_agg"]) / np.maximum(1.0, np.abs(c["kw_agg"]))) < 1e-9
            assert rel(P, c["kw_P"]) < 1e-9
        else:
            st = np.array(k.run_kalman_filter(), dtype=np.float64)
            assert len(st) == int(c["hw_n_states"])
            assert rel(sample(st), c["hw_states"]) < 1e-9
