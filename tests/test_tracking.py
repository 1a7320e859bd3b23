"""Tracking channel: oracle KATs from the reference's own tests (gnss/tracking.rs:458-519), oracle
behaviour on a synthesised signal (CPU), and GPU parity of r4wb_track_process with the oracle (pytest -m gpu)."""
import numpy as np
import pytest

from tests.conftest import config_path


def _gps_signal(oracle, n_ms, prn=7, doppler=1234.5, cn0=50.0, nav=True, seed=42):
    """n_ms of a single GPS L1 C/A satellite at 5 MHz from the oracle's scenario generator (noise on, cf32-rounded)"""
    from r4w_b200.config import load_config
    cfg = load_config(config_path("e1c_prn3_20s_withdoppler"), cli_elevation_mask_deg=5.0).copy()
    s = cfg.satellites[0]
    s.signal, s.prn, s.nav_data, s.doppler_hz, s.cn0_dbhz = "GpsL1Ca", prn, nav, doppler, cn0
    s.plane, s.slot = min(s.plane, 5), min(s.slot, 5)
    cfg.output.duration_s = n_ms * 1e-3
    cfg.output.seed = seed
    x = oracle.OracleScenario(cfg, noise=True).generate_range(0, n_ms * 5000)
    return x.astype(np.complex64), cfg


def _code_phase_of(oracle, x, prn):
    """coarse acquisition with the oracle's PCPS on the first ms -> (code phase in chips, doppler)"""
    idx = (np.arange(5000) * 1.023e6 / 5e6).astype(np.int64) % 1023
    code = oracle.gps_ca_code(prn)
    acq = oracle.OraclePcps(5000, 5e6).with_doppler_range(5000.0, 250.0)
    r = acq.acquire(x[:5000].astype(np.complex128), code[idx].astype(np.int8), prn)
    # the replica is aligned `lag` samples into the snapshot: sample 0 sits (5000 - lag) samples into the code
    return ((5000 - r.code_phase) % 5000) * 1.023e6 / 5e6, r.doppler_hz


def test_loop_filter_kats(oracle):
    """test_loop_filter_2nd_converges / test_loop_filter_3rd_converges (tracking.rs:462-488)"""
    out = oracle.loop_filter_2nd_run(1.0, 0.001, 0.1, 1000)
    first = oracle.loop_filter_2nd_run(1.0, 0.001, 0.1, 1)
    assert out > 0.0 and out > first
    # closed form: k1 d + n k2 d
    wn = 8.0 / 3.0
    assert first == pytest.approx((2 * wn / np.sqrt(2.0) * 1e-3 + wn * wn * 1e-6) * 0.1, rel=1e-12)
    assert out == pytest.approx((2 * wn / np.sqrt(2.0) * 1e-3 + 1000 * wn * wn * 1e-6) * 0.1, rel=1e-9)
    assert oracle.loop_filter_3rd_run(15.0, 0.001, 0.1, 1000) > 0.0


def test_dll_s_curve_shape(oracle):
    """test_dll_s_curve_shape (tracking.rs:490-518)"""
    e, d = oracle.dll_s_curve(0.5, 201)
    mid = len(e) // 2
    assert abs(e[mid]) < 0.02 and abs(d[mid]) < 0.1
    neg = d[(e > -0.4) & (e < -0.1)][0]
    pos = d[(e > 0.1) & (e < 0.4)][0]
    assert neg * pos < 0.0


def test_oracle_tracks_synthesised_gps(oracle):
    """the restated loop locks onto the oracle's own GPS L1 C/A signal: carrier frequency pulled to the true Doppler,
    code lock, C/N0 estimate reported, one nav bit per 20 ms"""
    x, _ = _gps_signal(oracle, 400, prn=7, doppler=1234.5, cn0=50.0)
    cp0, dop0 = _code_phase_of(oracle, x, 7)
    assert abs(dop0 - 1234.5) <= 250.0
    ch = oracle.OracleTrackingChannel(7, 1023, 5e6, 1.023e6, cp0, dop0)
    st = ch.run(x, oracle.gps_ca_code(7), 5000, 400)
    assert st["ms_count"][-1] == 400
    assert np.all(st["code_lock"][50:] == 1)
    assert abs(np.median(st["carrier_freq_hz"][300:]) - 1234.5) < 25.0
    assert len(ch.nav_bits()) == 20
    assert np.all(np.abs(ch.nav_bits()) == 1)
    final = ch.state()
    assert final["ms_count"] == 400 and final["code_phase"] == st["code_phase"][-1]


def _assert_states_close(got, want, code_length=1023.0, tol=1e-8):
    for f in ("carrier_freq_hz", "prompt_i", "prompt_q"):
        scale = max(1.0, float(np.max(np.abs(want[f]))))
        assert np.max(np.abs(got[f] - want[f])) <= tol * scale, f
    # phases are modular quantities
    d = np.abs(got["code_phase"] - want["code_phase"])
    assert np.max(np.minimum(d, code_length - d)) <= tol * code_length
    d = np.abs(got["carrier_phase_rad"] - want["carrier_phase_rad"])
    assert np.max(np.minimum(d, 2 * np.pi - d)) <= 1e-7
    assert np.max(np.abs(got["cn0_dbhz"] - want["cn0_dbhz"])) <= 1e-6
    for f in ("ms_count", "prn", "carrier_lock", "code_lock", "bit_sync"):
        assert np.array_equal(got[f], want[f]), f


@pytest.mark.gpu
def test_gpu_tracking_matches_oracle_gps(gpu, oracle):
    """three GPS channels (two present PRNs at different Dopplers are not available in one file, so: the true PRN, the
    true PRN started off-frequency, and an absent PRN) over 300 ms: every TrackingState of every call equal to the oracle's"""
    x, _ = _gps_signal(oracle, 300, prn=7, doppler=-2210.0, cn0=47.0)
    cp0, dop0 = _code_phase_of(oracle, x, 7)
    chans = [dict(prn=7, code_length=1023, sample_rate=5e6, chipping_rate=1.023e6, initial_code_phase=cp0, initial_doppler=dop0),
             dict(prn=7, code_length=1023, sample_rate=5e6, chipping_rate=1.023e6, initial_code_phase=cp0 + 0.2, initial_doppler=dop0 + 40.0,
                  dll_bandwidth_hz=2.0, pll_bandwidth_hz=18.0),
             dict(prn=19, code_length=1023, sample_rate=5e6, chipping_rate=1.023e6, initial_code_phase=100.0, initial_doppler=500.0)]
    codes = np.stack([oracle.gps_ca_code(c["prn"]) for c in chans]).astype(np.int8)
    bank = gpu.TrackerBank(chans)
    got = np.concatenate([bank.process(x[:500_000], codes, 5000, 100), bank.process(x[500_000:], codes, 5000, 200)])   # state persists
    for k, c in enumerate(chans):
        o = oracle.OracleTrackingChannel(c["prn"], 1023, 5e6, 1.023e6, c["initial_code_phase"], c["initial_doppler"])
        if "dll_bandwidth_hz" in c:
            o.with_dll_bandwidth(c["dll_bandwidth_hz"]).with_pll_bandwidth(c["pll_bandwidth_hz"])
        want = o.run(x, codes[k], 5000, 300)
        _assert_states_close(got[:, k], want)
        assert np.array_equal(bank.nav_bits(k), o.nav_bits())
        fin, ofin = bank.state()[k], o.state()
        assert fin.ms_count == 300 and fin.cn0_dbhz == pytest.approx(float(ofin["cn0_dbhz"]), abs=1e-6)
    # (the reference loop settles ~90 Hz off the true Doppler here: its carrier NCO advances the phase by one sample per
    # period, tracking.rs:251 — reproduced, not corrected)
    assert got["code_lock"][-1, 0] == 1 and got["carrier_lock"][-1, 0] == 1


@pytest.mark.gpu
def test_gpu_tracking_single_channel_api_and_e1c(gpu, oracle):
    """the reference call shape (one period per call, f64 samples) and a Galileo E1C channel (4 092 chips, 4 ms periods
    of 20 000 samples) on device-resident cf32 input"""
    import torch
    x, _ = _gps_signal(oracle, 40, prn=3, doppler=800.0, cn0=52.0, nav=False)
    code = oracle.gps_ca_code(3).astype(np.int8)
    cp0, dop0 = _code_phase_of(oracle, x, 3)
    ch = gpu.TrackingChannel(3, 1023, 5e6, 1.023e6, cp0, dop0).with_pll_bandwidth(12.0)
    o = oracle.OracleTrackingChannel(3, 1023, 5e6, 1.023e6, cp0, dop0).with_pll_bandwidth(12.0)
    for p in range(40):
        seg = x[p * 5000:(p + 1) * 5000].astype(np.complex128)
        g, w = ch.process(seg, code), o.process(seg, code)
        assert g.ms_count == w["ms_count"] and g.carrier_freq_hz == pytest.approx(float(w["carrier_freq_hz"]), abs=1e-7)
        assert g.prompt_i == pytest.approx(float(w["prompt_i"]), rel=1e-9, abs=1e-6)
    from r4w_b200.config import load_config
    cfg = load_config(config_path("e1c_prn3_20s_withdoppler"), cli_elevation_mask_deg=5.0)
    n_per = 20_000
    xe = gpu.GnssScenario(cfg, noise=True).generate_range(0, 50 * n_per)
    e1c = oracle.e1_code(1, 3).astype(np.int8)
    chans = [dict(prn=3, code_length=4092, sample_rate=5e6, chipping_rate=1.023e6, initial_code_phase=float(cpx), initial_doppler=-457.0)
             for cpx in (2941.0, 2941.3)]
    codes = np.stack([e1c, e1c])
    bank = gpu.TrackerBank(chans)
    got = bank.process(torch.from_numpy(xe).cuda(), codes, n_per, 50)
    for k, c in enumerate(chans):
        want = oracle.OracleTrackingChannel(3, 4092, 5e6, 1.023e6, c["initial_code_phase"], -457.0).run(xe, e1c, n_per, 50)
        _assert_states_close(got[:, k], want, 4092.0)


def _golden():
    import os
    from tests.conftest import GOLDEN_DIR
    g = np.load(os.path.join(GOLDEN_DIR, "track_case.npz"))
    return g["x"], g["code"], float(g["start"][0]), float(g["start"][1]), g["states"]


def test_oracle_reproduces_golden_tracking_case(oracle):
    """committed vectors (tools/make_golden_track.py): the restated tracking loop reproduces them bit for bit"""
    x, code, cp0, dop0, want = _golden()
    got = oracle.OracleTrackingChannel(11, 1023, 5e6, 1.023e6, cp0, dop0).with_dll_bandwidth(2.0).run(x, code, 5000, 20)
    for f in want.dtype.names:
        assert np.array_equal(got[f], want[f]), f


@pytest.mark.gpu
def test_gpu_tracking_golden_case(gpu):
    """the CUDA channel against the committed oracle states (no oracle on this path)"""
    x, code, cp0, dop0, want = _golden()
    bank = gpu.TrackerBank([dict(prn=11, code_length=1023, sample_rate=5e6, chipping_rate=1.023e6, initial_code_phase=cp0,
                                 initial_doppler=dop0, dll_bandwidth_hz=2.0)])
    got = bank.process(x, code[None, :], 5000, 20)[:, 0]
    _assert_states_close(got, want)
