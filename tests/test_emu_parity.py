"""Host replay of the DEVICE arithmetic (tests/emu/, the same __host__ __device__ functions the sm_100a kernels
call) against the oracle.  Lets the index arithmetic be checked in a container without a GPU; the `-m gpu`
tests repeat the comparisons with the real kernels through the C-ABI."""
import numpy as np
import pytest

from tests.conftest import config_path

TOL = 1e-5          # north_star: clean-scenario IQ within 1e-5 relative RMS


def _cfg(name):
    from r4w_b200.config import load_config
    return load_config(config_path(name), cli_elevation_mask_deg=5.0)


def _relrms(a, b):
    return float(np.sqrt(np.sum(np.abs(a - b) ** 2) / np.sum(np.abs(b) ** 2)))


@pytest.mark.parametrize("name,first,n", [
    ("e1c_prn3_20s_withdoppler", 0, 12000),
    ("e1c_prn3_20s_withdoppler", 99_990_000, 10000),          # tail of the file: sequential f64 phase drift is largest
    ("e1c_8prn_20s_clean", 0, 20000),
    ("e1c_8prn_20s_clean", 4993, 10014),                      # ragged window across block boundaries
    ("e1c_8prn_20s_clean", 60_000_000, 6000),
    ("e1c_60s_all_prns", 0, 11000),                           # range ramp: code phase steps every 1 ms block
    ("e1c_60s_all_prns", 123_455_000, 7000),
    ("e1c_8prn_60s_cn34_orbital", 0, 11000),                  # orbital Doppler / range per block
    ("e1c_8prn_60s_cn34_orbital", 2_500_000, 6000),
    ("e1c_prn3_20s_30ms_delay", 0, 11000),
    ("e1c_60s_cn34_effects", 0, 11000),                       # Klobuchar + Saastamoinen delays evaluated per block
    ("e1c_8prn_60s_mach3_ftwayne_berne", 149_995_000, 10000), # receiver trajectory (Mach 3)
])
def test_synthesis_replay_matches_oracle(oracle, emu, name, first, n):
    cfg = _cfg(name)
    got = emu.EmuScenario(cfg, noise=False).generate_range(first, n)
    want = oracle.OracleScenario(cfg, noise=False).generate_range(first, n)
    assert got.size == want.size == n
    assert _relrms(got, want) <= TOL


def test_closed_form_phase_drifts_from_reference(oracle, emu):
    """why the binade-segment phase model exists: (i+1)*inc differs from the reference's sequential f64 accumulation late in the file"""
    cfg = _cfg("e1c_8prn_20s_clean")
    first, n = 99_990_000, 5000
    want = oracle.OracleScenario(cfg, noise=False).generate_range(first, n)
    exact = emu.EmuScenario(cfg, noise=False).generate_range(first, n)
    closed = emu.EmuScenario(cfg, noise=False, closed_form_phase=True).generate_range(first, n)
    assert _relrms(exact, want) <= TOL < _relrms(closed, want)


def test_per_satellite_and_chip_boundaries(oracle, emu):
    """each satellite alone (amplitude-normalised) so a single flipped half-chip would show as an O(1) sample error"""
    cfg = _cfg("e1c_8prn_20s_clean")
    for k in (0, 4):
        one = cfg.copy()
        one.satellites = [cfg.satellites[k]]
        got = emu.EmuScenario(one, noise=False).generate_range(10_000_000, 20000)
        want = oracle.OracleScenario(one, noise=False).generate_range(10_000_000, 20000)
        amp = 10 ** ((one.satellites[0].cn0_dbhz - 44.0) / 20.0)
        assert np.abs(got - want).max() / amp < 2e-5
        full = emu.EmuScenario(cfg, noise=False).generate_range(10_000_000, 2000, only_sat=k)
        assert np.abs(full - got[:2000]).max() / amp < 1e-6


def test_sequential_generate_block_replay(oracle, emu):
    """generate_block with caller-chosen sizes (one reference block each, FIR history across blocks)"""
    cfg = _cfg("e1c_60s_all_prns")
    cfg.output.duration_s = 0.006
    e, o = emu.EmuScenario(cfg, noise=False), oracle.OracleScenario(cfg, noise=False)
    for bs in (5000, 1234, 8000, 5000, 20000):
        a, b = e.generate_block(bs), o.generate_block(bs)
        assert a.size == b.size
        if a.size:
            assert _relrms(a, b) <= TOL
    assert e.generate_block(5000).size == 0


def test_block_params_match_oracle_phase1(oracle, emu):
    """prologue entries (k_block_params arithmetic) vs generate_block Phase 1 (scenario.rs:378-454)"""
    cfg = _cfg("e1c_8prn_60s_cn34_orbital")
    e, o = emu.EmuScenario(cfg), oracle.OracleScenario(cfg)
    for block in (0, 1, 777):
        o.reset(); o.skip_to(block * 5000)
        params = o.peek_params()
        for sat in range(8):
            bp = e.block_params(block, sat)
            p = params[sat]
            assert bp[0] == p.visible == 1
            assert abs(bp[5] - p.initial_code_phase) < 1e-9 and bp[6] == p.initial_epoch_offset
            assert abs(bp[7] / p.rx_amplitude - 1.0) < 1e-6
            assert abs(bp[3] * 5e6 - p.doppler_start_hz) < 1e-6          # f (cycles/sample) * fs
            cyc = p.phase_before / (2.0 * np.pi)
            d = (bp[2] - cyc) % 1.0
            assert min(d, 1.0 - d) < 1e-6                                 # carried Doppler phase (mod 1 cycle)


def test_noise_replay_statistics(emu):
    cfg = _cfg("e1c_prn3_20s_withdoppler")
    n = 400_000
    noisy = emu.EmuScenario(cfg, noise=True).generate_range(0, n)
    clean = emu.EmuScenario(cfg, noise=False).generate_range(0, n)
    w = (noisy - clean).astype(np.complex128)
    sigma = 12.595361729330076
    assert abs(w.real.std() / sigma - 1.0) < 0.01 and abs(w.imag.std() / sigma - 1.0) < 0.01
    assert abs(w.mean()) < 0.1 and abs(np.mean(w.real * w.imag)) / sigma ** 2 < 0.01
    assert abs(np.mean(w[1:] * np.conj(w[:-1]))) / (2 * sigma ** 2) < 0.01          # white
    k = np.mean(w.real ** 4) / np.mean(w.real ** 2) ** 2
    assert abs(k - 3.0) < 0.05                                                     # Gaussian kurtosis
    again = emu.EmuScenario(cfg, noise=True).generate_range(1000, 50)              # counter-based: any window reproduces
    assert np.array_equal(again, noisy[1000:1050])


@pytest.mark.parametrize("log_n,log_m", [(0, 0), (1, 1), (3, 3), (4, 4), (6, 6), (10, 10), (11, 11), (13, 12), (15, 14), (15, 13)])
def test_fft_engine_replay(emu, log_n, log_m):
    rng = np.random.default_rng(log_n * 31 + log_m)
    n = 1 << log_n
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    for inverse in (False, True):
        ref = np.fft.ifft(x) * n if inverse else np.fft.fft(x)
        scale = max(np.abs(ref).max(), 1e-30)
        assert np.abs(emu.fft(x, log_m, inverse, double=True) - ref).max() / scale < 1e-13
        assert np.abs(emu.fft(x, log_m, inverse, double=False) - ref).max() / scale < 5e-6


def test_pcps_replay_kat_and_anchors(oracle, emu):
    code = oracle.gps_ca_code(1)
    i = np.arange(1023)
    sig = code[(i + 1023 - 100) % 1023] * np.exp(2j * np.pi * 1000.0 * (i / 1023.0))
    for dbl in (False, True):
        best, second, total, lin = emu.pcps(1023, 1023.0, 5000.0, 500.0, sig, code, double=dbl)
        assert lin % 1023 == 100 and abs(-5000.0 + (lin // 1023) * 500.0 - 1000.0) <= 500.0
    cfg = _cfg("e1c_8prn_20s_clean")
    x = oracle.to_cf32(oracle.OracleScenario(cfg, noise=False).generate_range(0, 20000))
    oacq = oracle.OraclePcps(20000, 5e6).with_doppler_range(5000.0, 250.0)
    for prn in (3, 5, 1):
        rep = oracle.e1c_replica(prn, 5e6, 20000)
        grid, olin = oacq.acquire_grid(x.astype(np.complex128), rep)
        best, second, total, lin = emu.pcps(20000, 5e6, 5000.0, 250.0, x, rep)
        assert lin == olin
        assert abs(best / grid.max() - 1.0) < 1e-5 and abs(total / grid.sum() - 1.0) < 1e-5
        assert abs(second / np.sort(grid.ravel())[-2] - 1.0) < 1e-5
    rep = oracle.e1c_replica(25, 5e6, 20000)
    grid, olin = oacq.acquire_grid(x.astype(np.complex128), rep)
    best, second, total, lin, g64 = emu.pcps(20000, 5e6, 5000.0, 250.0, x.astype(np.complex128), rep, double=True, want_grid=True)
    assert lin == olin and np.abs(g64 - grid).max() / grid.max() < 1e-11


def test_peak_merge_tie_break(emu):
    """equal maxima: the lowest linear index (Doppler ascending, then lag) wins, as the reference's strict `>` scan does"""
    L = 8
    code = np.ones(L, np.int8)
    x = np.zeros(L, np.complex128); x[0] = 1.0     # corr[k] = sum_n x[n+k] c[n] -> |corr|^2 = 1 at lag 0 for every bin
    best, second, total, lin = emu.pcps(L, 8.0, 2.0, 1.0, x, code, double=True)
    assert lin == 0 and best == pytest.approx(1.0) and second == pytest.approx(1.0)


def test_golden_fixtures_cpu(oracle, emu):
    """committed vectors (tools/make_golden.py): the oracle still reproduces them bit for bit, the replay within tolerance"""
    import os
    from tests.conftest import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, "synth_windows.npz"))
    keys = [k for k in z.files if k.endswith("_iq")]
    assert len(keys) == 5
    for key in keys:
        name, first = key[:-3].rsplit("@", 1)
        cfg = _cfg(name)
        want = z[key]
        assert np.array_equal(oracle.to_cf32(oracle.OracleScenario(cfg, noise=False).generate_range(int(first), want.size)), want)
        assert _relrms(emu.EmuScenario(cfg, noise=False).generate_range(int(first), want.size), want) <= TOL
    a = np.load(os.path.join(GOLDEN_DIR, "acq_cases.npz"))
    x, rows = a["x"], a["results"]
    oacq = oracle.OraclePcps(20000, 5e6).with_doppler_range(5000.0, 250.0)
    for prn, lag, dop, metric, det in rows[[0, 2, 24, 49]]:
        r = oacq.acquire(x.astype(np.complex128), oracle.e1c_replica(int(prn), 5e6, 20000), int(prn))
        assert (r.code_phase, r.doppler_hz, bool(r.detected)) == (lag, dop, bool(det))
        best, second, total, lin = emu.pcps(20000, 5e6, 5000.0, 250.0, x, oracle.e1c_replica(int(prn), 5e6, 20000))
        assert (lin % 20000, -5000.0 + (lin // 20000) * 250.0) == (lag, dop)


def test_class_table_equals_arithmetic_classes(emu, monkeypatch):
    """the fraction-indexed boundary-age table (synth_model.cpp, lut_den) and the arithmetic floor sums name the same
    class for every sample: bit-identical replay with the table switched off"""
    cfg = _cfg("e1c_8prn_60s_cn34_orbital")
    with_table = emu.EmuScenario(cfg, noise=False).generate_range(7_495_000, 12000)
    monkeypatch.setenv("R4WB_SYNTH_NO_LUT", "1")
    without = emu.EmuScenario(cfg, noise=False).generate_range(7_495_000, 12000)
    assert np.array_equal(with_table, without)


def _signal_variant(signals):
    """e1c_8prn_20s_clean with the first len(signals) satellites re-typed: (signal, prn, nav_data)"""
    cfg = _cfg("e1c_8prn_20s_clean").copy()
    cfg.satellites = cfg.satellites[: len(signals)]
    for s, (sig, prn, nav) in zip(cfg.satellites, signals):
        s.signal, s.prn, s.nav_data = sig, prn, nav
        if sig in ("GpsL1Ca", "GpsL5"):
            s.plane, s.slot = min(s.plane, 5), min(s.slot, 5)
        if sig == "GlonassL1of":
            s.plane, s.slot = min(s.plane, 2), min(s.slot, 7)
    return cfg


SIGNAL_CASES = {
    "gps_l1ca": [("GpsL1Ca", 3, True), ("GpsL1Ca", 17, False), ("GpsL1Ca", 32, True)],       # BPSK, 1 ms code, 20 ms nav bits
    "galileo_e1b": [("GalileoE1", 3, True), ("GalileoE1", 25, False)],                       # BOC(1,1), nav bit per 4 ms epoch
    "galileo_e1os": [("GalileoE1OS", 8, True), ("GalileoE1OS", 2, False), ("GalileoE1C", 5, False)],   # (e1b - e1c) / sqrt 2
    "mixed": [("GpsL1Ca", 5, True), ("GalileoE1", 11, True), ("GalileoE1C", 12, False), ("GalileoE1OS", 13, True)],
    # chip rates other than 1.023 MHz: the direct path (k_synth_direct)
    "glonass": [("GlonassL1of", 3, True), ("GlonassL1of", 0, False)],                       # 0.511 Mchip/s, 511 chips, PRN = frequency channel
    "gps_l5": [("GpsL5", 7, True), ("GpsL5", 32, False)],                                   # 10.23 Mchip/s, 10 230 chips (aliased at 5 MHz)
    "all_signals": [("GlonassL1of", 3, True), ("GpsL5", 7, True), ("GpsL1Ca", 9, True), ("GalileoE1C", 12, False), ("GalileoE1OS", 2, True)],
}


@pytest.mark.parametrize("case", sorted(SIGNAL_CASES))
def test_other_signals_replay_matches_oracle(oracle, emu, case):
    """the other emitter branches of generate_baseband_iq (satellite_emitter.rs:248-343) that chip at
    1.023 MHz — GPS L1 C/A, Galileo E1B, the E1OS composite — through the same kernels (per-satellite code structure)"""
    cfg = _signal_variant(SIGNAL_CASES[case])
    for first, n in ((0, 11000), (24_995_000, 12000)):
        got = emu.EmuScenario(cfg, noise=False).generate_range(first, n)
        want = oracle.OracleScenario(cfg, noise=False).generate_range(first, n)
        assert _relrms(got, want) <= TOL


@pytest.mark.parametrize("fs", [4_000_000.0, 4_092_000.0, 8_000_000.0, 12_500_000.0])
def test_other_sample_rates_replay_matches_oracle(oracle, emu, fs):
    """sample rates other than the configs' 5 MHz: samples per half-chip S = 4 fs / 1.023e6 changes the number of boundary-age
    classes (4 S: 63 at 4 MHz, 196 at 12.5 MHz; row stride of the collapsed-FIR table follows), the code period in samples
    and, at 4.092 MHz (exactly 4 samples per chip), puts every sample ON a chip boundary (literal-expression path)"""
    for name in ("e1c_8prn_20s_clean", "e1c_8prn_60s_cn34_orbital"):
        cfg = _cfg(name).copy()
        cfg.output.sample_rate = fs
        cfg.output.lpf_cutoff_hz = 0.0                     # -> fs / 2 (scenario.rs:210-214)
        B = int(np.ceil(fs * 0.001))
        first, n = 500 * B - 7, B + 13
        got = emu.EmuScenario(cfg, noise=False).generate_range(first, n)
        want = oracle.OracleScenario(cfg, noise=False).generate_range(first, n)
        assert _relrms(got, want) <= TOL


@pytest.mark.parametrize("name", ["OpenSky", "UrbanCanyon", "Driving", "Walking", "HighDynamics", "MultiConstellation"])
def test_presets_replay_matches_oracle(oracle, emu, name):
    """GnssScenarioPreset::to_config (scenario_config.rs:581-700): GPS L1 C/A (and Galileo E1B) satellites with no overrides,
    so geometry, Doppler, C/N0 (link budget + patch antenna), Klobuchar and Saastamoinen all come from the nominal orbits"""
    from r4w_b200.config import preset_config
    cfg = preset_config(name)
    n = cfg.total_samples()
    got = emu.EmuScenario(cfg, noise=False).generate_range(0, n)
    want = oracle.OracleScenario(cfg, noise=False).generate_range(0, n)
    assert got.size == n and np.abs(want).max() > 0
    assert _relrms(got, want) <= TOL
    longer = cfg.copy()
    longer.output.duration_s = 0.5
    got = emu.EmuScenario(longer, noise=False).generate_range(2_400_000, 7000)
    want = oracle.OracleScenario(longer, noise=False).generate_range(2_400_000, 7000)
    assert _relrms(got, want) <= TOL


@pytest.mark.parametrize("rate", [0.3, 40.0, 5000.0, -20000.0])
def test_doppler_rate_replay_matches_oracle(oracle, emu, rate):
    """`doppler_rate_hz_per_s` (scenario.rs:416-421): Doppler ramps linearly inside and across blocks; small rates take the
    linearised phasor recurrence, large ones the per-sample sincos path"""
    cfg = _cfg("e1c_8prn_20s_clean").copy()
    for k, s in enumerate(cfg.satellites):
        s.doppler_rate_hz_per_s = rate * (1 + 0.1 * k) * (1 if k % 2 == 0 else -1)
    for first, n in ((0, 11000), (49_995_000, 12000)):
        got = emu.EmuScenario(cfg, noise=False).generate_range(first, n)
        want = oracle.OracleScenario(cfg, noise=False).generate_range(first, n)
        assert _relrms(got, want) <= TOL


@pytest.mark.parametrize("kind", ["Isotropic", "Hemispherical", "Patch", "ChokeRing"])
def test_antenna_patterns_replay_matches_oracle(oracle, emu, kind):
    """AntennaPattern::gain_dbi (gnss/environment/antenna.rs:35-62) inside the link budget of satellites without a cn0 override"""
    from r4w_b200.config import preset_config, AntennaPattern
    cfg = preset_config("MultiConstellation")
    cfg.receiver.antenna = AntennaPattern(kind, 4.0, 120.0)
    cfg.output.duration_s = 0.004
    got = emu.EmuScenario(cfg, noise=False).generate_range(0, 20000)
    want = oracle.OracleScenario(cfg, noise=False).generate_range(0, 20000)
    assert _relrms(got, want) <= TOL


@pytest.mark.parametrize("span_hz,k_lo,k_hi,n,max_steps", [(2e-3, 14, 26, 5000, 1024), (5e-2, 17, 30, 5000, 1024), (2e-3, 8, 13, 5000, 1024),
                                                            (1.0, 20, 24, 5000, 64), (2e-3, 16, 24, 1234, 1024), (0.5, 30, 40, 5000, 1024)])
def test_phase_q_step_form_equals_per_sample_sum(emu, span_hz, k_lo, k_hi, n, max_steps):
    """k_phase_q's block sum Q = sum_i rint(inc_i / ulp) is formed from the step positions of the monotone sequence rint(inc_i / ulp)
    (block_phase_q_steps: each step located by interpolation and proven with the exact increment chain on both sides) instead of
    visiting the 5000 samples; it must equal the per-sample sum (block_phase_q) bit for bit, tie flag included — rising, falling
    and constant Doppler, zero crossings, coarse-mantissa Dopplers (plateaus), both division forms, few and many levels"""
    bad, declined, levels, ties = emu.phase_q_steps_check(20260, 3000, span_hz, k_lo, k_hi, n, max_steps)
    assert bad == 0
    if (span_hz, k_lo) == (2e-3, 14):
        assert declined == 0 and levels >= 50        # the production regime: every block answered, dozens of levels exercised
    if max_steps == 64:
        assert declined > 0                          # blocks with more levels than allowed are handed back, not guessed


@pytest.mark.parametrize("d0,rate,jerk,blocks", [(2779.31, -0.42, 0.0, 20000), (-2831.7, 0.35, 0.0, 20000), (35.0, -0.9, 0.0, 60000),
                                                 (-457.3938, 0.0, 1e-4, 20000), (0.0, 0.0, 0.0, 50)])
def test_exact_phase_model_equals_sequential_accumulation(emu, d0, rate, jerk, blocks):
    """the reference adds one f64 increment per sample for the whole run (scenario.rs:516-527); the product reproduces that sum
    block by block from integer sums of the increments rounded to the binade's ulp, walking only the blocks that cross a binade
    (or hold a tie): both must give the same f64 at EVERY block boundary — positive and negative phase, a Doppler zero crossing
    (phase turns around), a curved Doppler, and zero Doppler"""
    mism, walked, final, drift = emu.phase_model_check(d0, rate, jerk, blocks)
    assert mism == 0
    if d0 != 0.0:
        assert walked < blocks // 10          # almost every block takes the integer-sum path
        assert drift != 0.0                   # the reference's sum is not the real-number sum: that drift is what is reproduced


def test_sequential_blocks_equal_random_access_dynamic(emu):
    """dynamic satellites: the sequential API walks the reference's f64 phase sample by sample on the host (SeqState::advance),
    the random-access path rebuilds it from integer sums per block (k_phase_q / k_phase_exact) — the two must hand the kernel
    the same start phase for every block, so canonical blocks rendered one by one equal one render of the range bit for bit"""
    cfg = _cfg("e1c_8prn_60s_cn34_orbital")
    a = emu.EmuScenario(cfg, noise=False)
    seq = np.concatenate([a.generate_block(5000) for _ in range(24)])
    rnd = emu.EmuScenario(cfg, noise=False).generate_range(0, 120000)
    assert np.array_equal(seq, rnd)
