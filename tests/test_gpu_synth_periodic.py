"""Period-resident synthesis kernels (r4w_b200/csrc/synth_periodic.cu) against the oracle and against k_synth.

Static scenarios (constant code delay and Doppler: e1c_prn3_20s_withdoppler, e1c_8prn_20s_clean, ...) repeat every
primary-code period up to the epoch sign, so the bulk of a long render takes k_synth_periodic + k_periodic_fix; the
partial periods at the ends take k_synth.  R4WB_SYNTH_PERIODIC=0 forces k_synth everywhere (A/B).  pytest -m gpu."""
import numpy as np
import pytest

from tests.test_gpu_synth import TOL, _cfg, _relrms

pytestmark = pytest.mark.gpu


def _render(gpu, monkeypatch, cfg, first, n, noise, periodic):
    monkeypatch.setenv("R4WB_SYNTH_PERIODIC", "1" if periodic else "0")
    sc = gpu.GnssScenario(cfg, noise=noise)
    x = sc.generate_range(first, n)
    return x, sc.last_power_sum(), sc.last_path()


@pytest.mark.parametrize("name,first,n", [
    ("e1c_8prn_20s_clean", 0, 700_003),              # period 0 (zero FIR history) + ragged tail through k_synth
    ("e1c_8prn_20s_clean", 50_005_000, 815_000),     # head of 15 000 samples, whole periods, no tail
    ("e1c_prn3_20s_withdoppler", 99_000_000, 1_000_000),   # one satellite, up to the last sample of the run
    ("e1c_60s_clean", 123_460_000, 700_000),
])
def test_periodic_matches_oracle_and_general(gpu, oracle, monkeypatch, name, first, n):
    cfg = _cfg(name)
    fast, p_fast, path = _render(gpu, monkeypatch, cfg, first, n, False, True)
    assert path == 1, "period-resident kernels were not used"
    want = oracle.OracleScenario(cfg, noise=False).generate_range(first, n)
    assert _relrms(fast, want) <= TOL
    # worst single sample, relative to the signal RMS (a wrong epoch sign or class would be O(1))
    assert np.abs(fast - want).max() <= 2e-5 * np.sqrt(np.mean(np.abs(want) ** 2)) * 8
    gen, p_gen, path = _render(gpu, monkeypatch, cfg, first, n, False, False)
    assert path == 0
    assert _relrms(fast, gen) <= 2e-6
    assert p_fast == pytest.approx(float(np.sum(np.abs(fast.astype(np.complex128)) ** 2)), rel=1e-6)
    assert p_fast == pytest.approx(p_gen, rel=1e-6)


def test_periodic_noise_is_the_same_stream(gpu, monkeypatch):
    """noise on: both kernel families draw sample m's noise from Philox counter m >> 1, so the outputs differ only by the
    f32 rounding of the signal part"""
    cfg = _cfg("e1c_8prn_20s_clean")
    first, n = 20_000_000, 1_000_000
    fast, p_fast, path = _render(gpu, monkeypatch, cfg, first, n, True, True)
    assert path == 1
    gen, p_gen, _ = _render(gpu, monkeypatch, cfg, first, n, True, False)
    assert np.abs(fast - gen).max() <= 2e-4            # amplitude 11.2 x 8 satellites, sigma 12.6
    assert p_fast == pytest.approx(p_gen, rel=1e-6)


@pytest.mark.parametrize("case", ["gps_l1ca", "galileo_e1b", "galileo_e1os"])
def test_periodic_other_signals(gpu, oracle, monkeypatch, case):
    """nav-bit epochs (GPS L1 C/A: 5 000-sample period, a bit per 20 epochs; E1B: a bit per epoch) and the E1OS composite"""
    from tests.test_emu_parity import SIGNAL_CASES, _signal_variant
    cfg = _signal_variant(SIGNAL_CASES[case])
    first, n = 30_000_000, 700_000
    fast, _, path = _render(gpu, monkeypatch, cfg, first, n, False, True)
    assert path == 1
    want = oracle.OracleScenario(cfg, noise=False).generate_range(first, n)
    assert _relrms(fast, want) <= TOL
    assert np.abs(fast - want).max() <= 2e-5 * np.sqrt(np.mean(np.abs(want) ** 2)) * 8


def test_periodic_not_used_where_it_does_not_apply(gpu, oracle, monkeypatch):
    """orbital / range-ramp scenarios, mixed code lengths, short or misaligned ranges: k_synth, same results as before"""
    from tests.test_emu_parity import SIGNAL_CASES, _signal_variant
    monkeypatch.setenv("R4WB_SYNTH_PERIODIC", "1")
    for cfg, first, n in ((_cfg("e1c_8prn_60s_cn34_orbital"), 0, 800_000), (_cfg("e1c_60s_all_prns"), 1_000_000, 800_000),
                          (_signal_variant(SIGNAL_CASES["mixed"]), 0, 800_000), (_cfg("e1c_8prn_20s_clean"), 4993, 900_001),
                          (_cfg("e1c_8prn_20s_clean"), 0, 100_000)):
        sc = gpu.GnssScenario(cfg, noise=False)
        x = sc.generate_range(first, n)
        assert sc.last_path() == 0
        m = min(n, 60_000)
        want = oracle.OracleScenario(cfg, noise=False).generate_range(first + n - m, m)
        assert _relrms(x[n - m:], want) <= TOL


def test_periodic_device_output_full_config(gpu, monkeypatch):
    """BASELINE config 2 at full size into HBM: every primary-code period of the noise-free render is the first one
    times a unit phasor per satellite -> check the structure through a size-independent property: the power of every
    1 ms block equals the general kernel's on the same range, and the two renders agree sample by sample"""
    import torch
    cfg = _cfg("e1c_8prn_20s_clean")
    n = 100_000_000
    monkeypatch.setenv("R4WB_SYNTH_PERIODIC", "1")
    a = torch.empty(n, dtype=torch.complex64, device="cuda")
    sa = gpu.GnssScenario(cfg, noise=True)
    sa.generate_device(0, n, a)
    torch.cuda.synchronize()
    assert sa.last_path() == 1
    pa = sa.last_power_sum()
    monkeypatch.setenv("R4WB_SYNTH_PERIODIC", "0")
    b = torch.empty(n, dtype=torch.complex64, device="cuda")
    sb = gpu.GnssScenario(cfg, noise=True)
    sb.generate_device(0, n, b)
    torch.cuda.synchronize()
    assert sb.last_path() == 0
    assert float((a - b).abs().max()) <= 2e-4
    assert pa == pytest.approx(sb.last_power_sum(), rel=1e-6)
    assert float(torch.view_as_real(a).square().sum(dtype=torch.float64)) == pytest.approx(pa, rel=1e-6)


@pytest.mark.parametrize("fs,periodic", [(4_000_000.0, True), (4_092_000.0, True), (8_000_000.0, True), (12_500_000.0, True)])
def test_other_sample_rates(gpu, oracle, monkeypatch, fs, periodic):
    """both kernel families at other sample rates (code period 16 000 / 16 368 / 32 000 / 50 000 samples; 63-196 boundary-age
    classes; 4.092 MHz = exactly 4 samples per chip).  The orbital config at 4 / 8 MHz is also the regression test for the
    FMA-contraction bug of the f64 prologue (synth_prologue.cu): an oversample 5e-7 half-chips from a boundary."""
    for name, dynamic in (("e1c_8prn_20s_clean", False), ("e1c_8prn_60s_cn34_orbital", True)):
        cfg = _cfg(name).copy()
        cfg.output.sample_rate = fs
        cfg.output.lpf_cutoff_hz = 0.0
        L = int(round(fs * 0.004))
        first, n = 100 * L, 40 * L
        x, _, path = _render(gpu, monkeypatch, cfg, first, n, False, True)
        assert path == (1 if periodic and not dynamic else 0)
        m = 3 * L
        want = oracle.OracleScenario(cfg, noise=False).generate_range(first + n - m, m)
        assert _relrms(x[n - m:], want) <= TOL
        want = oracle.OracleScenario(cfg, noise=False).generate_range(first, L)
        assert _relrms(x[:L], want) <= TOL


def test_periodic_kernel_matches_oracle_across_the_file(gpu, oracle, monkeypatch):
    """The benchmarked static configuration at 64 period-aligned offsets over the whole 20 s file — half of them the periods in
    which a satellite's f64 carrier phase crosses a binade (a new PhaseSegment of the exact-phase emulation),
    the rest random.  Each offset is rendered as 40 periods (so the period-resident kernels take it) and the middle period
    is compared with the oracle."""
    monkeypatch.setenv("R4WB_SYNTH_PERIODIC", "1")
    cfg = _cfg("e1c_8prn_20s_clean")
    L, fs = 20000, 5e6
    n_periods = int(cfg.output.duration_s * fs) // L
    rng = np.random.default_rng(7)
    cross = []
    for s in cfg.satellites:                                   # |phase| = 2 pi |f| m / fs reaches 2^k at sample m
        k = 4
        while True:
            m = (2.0 ** k) * fs / (2 * np.pi * abs(s.doppler_hz))
            if m >= n_periods * L:
                break
            cross.append(int(m) // L)
            k += 1
    cross = sorted(set(p for p in cross if 20 <= p < n_periods - 20))
    picks = list(rng.choice(cross, size=min(32, len(cross)), replace=False)) + list(rng.integers(20, n_periods - 20, 64))
    picks = sorted(set(int(p) for p in picks))[:64]
    assert len(picks) >= 48
    sc = gpu.GnssScenario(cfg, noise=False)
    orc = oracle.OracleScenario(cfg, noise=False, threads=8)
    worst = 0.0
    for p in picks:
        got = sc.generate_range((p - 20) * L, 40 * L)
        assert sc.last_path() == 1
        want = orc.generate_range(p * L, L)
        worst = max(worst, _relrms(got[20 * L:21 * L], want))
    assert worst <= 1e-5, worst
