"""Parity of the CUDA PCPS acquisition path (through the C-ABI) with the oracle: the detected (code phase,
Doppler bin) indices are the bit-exact contract.  Run on the B200 box: pytest -m gpu."""
import os
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from tests.conftest import config_path

pytestmark = pytest.mark.gpu
ALL8 = [3, 25, 8, 2, 5, 16, 13, 15]


def _cfg(name):
    from r4w_b200.config import load_config
    return load_config(config_path(name), cli_elevation_mask_deg=5.0)


def _same(r, o, metric_rtol=2e-4):
    assert (r.code_phase, r.doppler_hz, bool(r.detected)) == (o.code_phase, o.doppler_hz, bool(o.detected)), (r, o.code_phase, o.doppler_hz)
    assert r.peak_metric == pytest.approx(o.peak_metric, rel=metric_rtol)
    assert (r.cn0_estimate is None) == (not o.has_cn0)
    if o.has_cn0:
        assert r.cn0_estimate == pytest.approx(o.cn0_estimate, abs=1e-3)


def test_reference_kats(gpu, oracle):
    """test_acquisition_no_noise / _wrong_prn / _grid (acquisition.rs:294-377) through the CUDA path"""
    code = oracle.gps_ca_code(1)
    i = np.arange(1023)
    sig = code[(i + 1023 - 100) % 1023] * np.exp(2j * np.pi * 1000.0 * (i / 1023.0))
    acq = gpu.PcpsAcquisition(1023, 1023.0).with_doppler_range(5000.0, 500.0).with_threshold(2.0)
    assert acq.fft_size() == 1024 and acq.num_doppler_bins() == 21
    r = acq.acquire(sig, code, 1)
    assert r.detected and int(r.code_phase) == 100 and abs(r.doppler_hz - 1000.0) <= 500.0
    _same(r, oracle.OraclePcps(1023, 1023.0).with_doppler_range(5000.0, 500.0).with_threshold(2.0).acquire(sig, code, 1))
    c7 = oracle.gps_ca_code(7)
    r = gpu.PcpsAcquisition(1023, 1023.0).with_threshold(2.5).acquire(code.astype(np.complex128), c7, 7)
    assert (not r.detected) or r.peak_metric < 10.0
    _same(r, oracle.OraclePcps(1023, 1023.0).with_threshold(2.5).acquire(code.astype(np.complex128), c7, 7))
    g = gpu.PcpsAcquisition(1023, 1023.0).with_doppler_range(2000.0, 500.0).acquire_grid(code.astype(np.complex128), code)
    dop, ph, pw = g.find_peak()
    assert abs(dop) <= 500.0 and (ph <= 1.0 or abs(ph - 1023.0) <= 1.0)
    og, olin = oracle.OraclePcps(1023, 1023.0).with_doppler_range(2000.0, 500.0).acquire_grid(code.astype(np.complex128), code)
    assert g.power.shape == og.shape == (9, 1023) and np.abs(g.power - og).max() / og.max() < 1e-11
    assert int(np.argmax(g.power)) == olin


def test_scenario_kat_gps_2046(gpu, oracle):
    """the shape of test_acquisition_on_scenario (scenario.rs:793-855): 2046 samples/code at 2.046 MHz, N = 2048"""
    rng = np.random.default_rng(7)
    code = np.repeat(oracle.gps_ca_code(5), 2)
    i = np.arange(2046)
    sig = np.roll(code, 321) * np.exp(2j * np.pi * 1750.0 * i / 2.046e6) + 0.7 * (rng.standard_normal(2046) + 1j * rng.standard_normal(2046))
    a = gpu.PcpsAcquisition(2046, 2.046e6).with_threshold(1.5)
    o = oracle.OraclePcps(2046, 2.046e6).with_threshold(1.5)
    assert a.fft_size() == 2048
    r = a.acquire(sig, code, 5)
    _same(r, o.acquire(sig, code, 5))
    assert int(r.code_phase) == 321 and r.doppler_hz == 1500.0      # corr[k] = sum x[n+k] c[n] peaks at the roll amount


@pytest.mark.parametrize("name", ["e1c_8prn_20s_clean", "e1c_prn3_20s_withdoppler", "e1c_60s_all_prns", "e1c_8prn_60s_cn34_orbital",
                                  "e1c_8prn_600s_cn34_orbital", "e1c_60s_clean", "e1c_60s_cn34", "e1c_60s_cn34_effects",
                                  "e1c_8prn_20s_cn34_orbital", "e1c_8prn_60s_mach3_ftwayne_berne", "e1c_prn3_20s_30ms_delay"])
def test_e1c_indices_match_oracle_all_prns(gpu, oracle, name):
    """every PRN 1-50 (present, absent, and the wrap-around-lag cases) on oracle-generated noisy input: (lag, bin) identical"""
    cfg = _cfg(name)
    cfg.output.duration_s = 0.008
    x = oracle.to_cf32(oracle.OracleScenario(cfg, noise=True, threads=8).generate_range(0, 40000))
    prns = list(range(1, 51))
    codes = np.stack([gpu.e1c_replica(p, 5e6, 20000) for p in prns])
    acq = gpu.PcpsAcquisition(20000, 5e6).with_doppler_range(5000.0, 250.0)
    assert acq.fft_size() == 32768 and acq.num_doppler_bins() == 41
    res = acq.acquire_batch(x, 2, 20000, 20000, codes, prns)
    oacq = oracle.OraclePcps(20000, 5e6).with_doppler_range(5000.0, 250.0)
    from concurrent.futures import ThreadPoolExecutor
    x64 = x.astype(np.complex128)
    jobs = [(s, c) for s in range(2) for c in range(50)]
    with ThreadPoolExecutor(8) as ex:
        want = list(ex.map(lambda j: oacq.acquire(x64[j[0] * 20000:(j[0] + 1) * 20000], codes[j[1]], prns[j[1]]), jobs))
    for (s, c), o in zip(jobs, want):
        _same(res[s][c], o)
        assert res[s][c].prn == prns[c]


def test_clean_anchor_values(gpu, oracle):
    """SURVEY.md §8c anchors through synth + acquire on the GPU end to end (noise off)"""
    cfg = _cfg("e1c_8prn_20s_clean")
    x = gpu.GnssScenario(cfg, noise=False).generate_range(0, 20000)
    acq = gpu.PcpsAcquisition(20000, 5e6).with_doppler_range(5000.0, 250.0)
    want = {3: (5625, 18), 25: (12965, 22), 8: (6074, 26), 2: (15397, 32), 13: (16269, 15), 15: (3532, 9),
            5: (4040, 32), 16: (227, 38), 1: (731, 20)}
    for prn, (lag, dbin) in want.items():
        r = acq.acquire(x, gpu.e1c_replica(prn, 5e6, 20000), prn)
        assert (int(r.code_phase), int(round((r.doppler_hz + 5000.0) / 250.0))) == (lag, dbin), prn


def test_acquire_grid_e1c_size_matches_oracle(gpu, oracle):
    """acquire_grid (acquisition.rs:199-249) at the E1C shape: fft_size 32 768, 41 x 20 000 surface, noisy 34 dB-Hz input —
    every cell against the oracle's f64 surface, the same first maximum"""
    cfg = _cfg("e1c_8prn_60s_cn34_orbital")
    x = oracle.OracleScenario(cfg, noise=True).generate_range(40_000, 20000)                  # f64
    a = gpu.PcpsAcquisition(20000, 5e6).with_doppler_range(5000.0, 250.0)
    o = oracle.OraclePcps(20000, 5e6).with_doppler_range(5000.0, 250.0)
    assert a.fft_size() == 32768
    for prn in (3, 16):
        rep = gpu.e1c_replica(prn, 5e6, 20000)
        g = a.acquire_grid(x, rep)
        og, olin = o.acquire_grid(x, rep)
        assert g.power.shape == og.shape == (41, 20000)
        assert np.max(np.abs(g.power - og)) <= 1e-9 * np.max(og)
        dop, lag, best = g.find_peak()
        d, p = divmod(olin, 20000)
        assert (lag, dop) == (float(p), -5000.0 + 250.0 * d) and best == pytest.approx(float(og[d, p]), rel=1e-9)


def test_default_doppler_grid_and_cf64_input(gpu, oracle):
    """reference defaults (+-5 kHz / 500 Hz = 21 bins, threshold 2.5) and Complex64 input as the Rust API passes it"""
    cfg = _cfg("e1c_8prn_20s_clean")
    x = oracle.OracleScenario(cfg, noise=True).generate_range(0, 20000)          # f64
    a, o = gpu.PcpsAcquisition(20000, 5e6), oracle.OraclePcps(20000, 5e6)
    assert a.num_doppler_bins() == 21
    for prn in (3, 15, 7):
        rep = gpu.e1c_replica(prn, 5e6, 20000)
        _same(a.acquire(x, rep, prn), o.acquire(x, rep, prn))


def test_edge_cases(gpu, oracle):
    rng = np.random.default_rng(3)
    a, o = gpu.PcpsAcquisition(20000, 5e6), oracle.OraclePcps(20000, 5e6)
    rep = gpu.e1c_replica(3, 5e6, 20000)
    # short input: `take(code_length)` just correlates fewer samples (acquisition.rs:134)
    x = (rng.standard_normal(12345) + 1j * rng.standard_normal(12345)).astype(np.complex64)
    _same(a.acquire(x, rep, 3), o.acquire(x.astype(np.complex128), rep, 3))
    # long input: only the first code_length samples are read
    x = (rng.standard_normal(30000) + 1j * rng.standard_normal(30000)).astype(np.complex64)
    _same(a.acquire(x, rep, 3), o.acquire(x.astype(np.complex128), rep, 3))
    # replica shorter / longer than the FFT size (zero-padded / truncated by resize, acquisition.rs:112)
    _same(a.acquire(x, rep[:7777], 3), o.acquire(x.astype(np.complex128), rep[:7777], 3))
    long_rep = np.concatenate([rep, rep])
    _same(a.acquire(x, long_rep, 3), o.acquire(x.astype(np.complex128), long_rep, 3))
    # all-zero input: nothing exceeds best_peak = 0 -> phase 0, doppler 0.0, metric 0, not detected
    z = a.acquire(np.zeros(20000, np.complex64), rep, 3)
    zo = o.acquire(np.zeros(20000, np.complex128), rep, 3)
    assert (z.code_phase, z.doppler_hz, z.peak_metric, z.detected) == (zo.code_phase, zo.doppler_hz, zo.peak_metric, bool(zo.detected)) == (0.0, 0.0, 0.0, False)
    # tiny transforms (fft_size 1, 2, 4, 8, 16, 32)
    for L in (1, 2, 3, 5, 13, 31):
        code = rng.choice(np.array([-1, 1], np.int8), L)
        sig = rng.standard_normal(L) + 1j * rng.standard_normal(L)
        _same(gpu.PcpsAcquisition(L, 1000.0).with_doppler_range(100.0, 50.0).acquire(sig, code, 9),
              oracle.OraclePcps(L, 1000.0).with_doppler_range(100.0, 50.0).acquire(sig, code, 9))
    with pytest.raises(gpu.R4wB200Error):
        gpu.PcpsAcquisition(20000, 5e6).with_doppler_range(5000.0, 0.0)


def test_exact_ties_take_lowest_index(gpu):
    """equal maxima in every Doppler row: the reference's ascending strict-`>` scan keeps the first (acquisition.rs:159)"""
    L = 8
    x = np.zeros(L, np.complex128); x[0] = 1.0
    r = gpu.PcpsAcquisition(L, 8.0).with_doppler_range(2.0, 1.0).acquire(x, np.ones(L, np.int8), 1)
    assert (r.code_phase, r.doppler_hz) == (0.0, -2.0)


def test_near_tie_guard_runs_f64(gpu, oracle):
    """a Doppler 1 mHz off the midpoint of two bins: the two candidate cells differ by 2.4e-6 relative, the size of the f32
    pipeline's own error, so the f64 re-run must decide — and must decide as the oracle does"""
    L, fs = 1023, 1.023e6
    code = oracle.gps_ca_code(3)
    i = np.arange(L)
    sig = (np.roll(code, 200) * np.exp(2j * np.pi * 250.001 * i / fs)).astype(np.complex128)
    acq = gpu.PcpsAcquisition(L, fs).with_doppler_range(500.0, 500.0)
    r = acq.acquire(sig, code, 3)
    o = oracle.OraclePcps(L, fs).with_doppler_range(500.0, 500.0).acquire(sig, code, 3)
    assert acq.guard_count() == 1
    assert (r.code_phase, r.doppler_hz) == (o.code_phase, o.doppler_hz) == (200.0, 500.0)
    assert r.peak_metric == pytest.approx(o.peak_metric, rel=1e-9)
    # a clear winner does not take the guard
    acq.acquire((np.roll(code, 200) * np.exp(2j * np.pi * 480.0 * i / fs)).astype(np.complex128), code, 3)
    assert acq.guard_count() == 0


def test_batch_layouts_and_device_input(gpu, oracle):
    """snapshot stride != n_input, device-resident input, results in [snapshot][code] order"""
    import torch
    cfg = _cfg("e1c_8prn_20s_clean")
    n = 5 * 20000
    d = torch.empty(n, dtype=torch.complex64, device="cuda")
    gpu.GnssScenario(cfg, noise=True).generate_device(0, n, d)
    host = d.cpu().numpy()
    codes = np.stack([gpu.e1c_replica(p, 5e6, 20000) for p in ALL8])
    acq = gpu.PcpsAcquisition(20000, 5e6).with_doppler_range(5000.0, 250.0)
    dev_res = acq.acquire_batch(d, 5, 20000, 20000, codes, ALL8)
    host_res = acq.acquire_batch(host, 5, 20000, 20000, codes, ALL8)
    strided = acq.acquire_batch(host, 3, 40000, 20000, codes, ALL8)
    for s in range(5):
        for c in range(8):
            assert dev_res[s][c] == host_res[s][c]
            single = acq.acquire(host[s * 20000:(s + 1) * 20000], codes[c], ALL8[c])
            assert single == host_res[s][c]
    for k, s in enumerate((0, 2, 4)):
        assert strided[k] == host_res[s]
    # present PRNs stay within one Doppler bin of each other in consecutive snapshots (static scenario; a Doppler between
    # two bins may land on either with noise)
    for c, prn in enumerate(ALL8):
        if prn in (3, 25, 8, 15):
            dops = [host_res[s][c].doppler_hz for s in range(5)]
            assert max(dops) - min(dops) <= 250.0
    with pytest.raises(ValueError):
        acq.acquire_batch(host, 6, 20000, 20000, codes, ALL8)


def test_batch_vs_oracle_on_gpu_generated_scenario(gpu, oracle):
    """synth -> acquire pipeline on the GPU vs the oracle acquiring the SAME GPU-generated samples"""
    cfg = _cfg("e1c_60s_all_prns")
    x = gpu.GnssScenario(cfg, noise=True).generate_range(150_000_000, 3 * 20000)
    codes = np.stack([gpu.e1c_replica(p, 5e6, 20000) for p in ALL8])
    acq = gpu.PcpsAcquisition(20000, 5e6).with_doppler_range(5000.0, 250.0)
    res = acq.acquire_batch(x, 3, 20000, 20000, codes, ALL8)
    oacq = oracle.OraclePcps(20000, 5e6).with_doppler_range(5000.0, 250.0)
    x64 = x.astype(np.complex128)
    for s in range(3):
        for c, prn in enumerate(ALL8):
            _same(res[s][c], oacq.acquire(x64[s * 20000:(s + 1) * 20000], codes[c], prn))


def test_golden_acquisition_cases(gpu):
    """committed oracle results for PRN 1-50 on a committed noisy snapshot (tests/golden/acq_cases.npz)"""
    import os
    from tests.conftest import GOLDEN_DIR
    a = np.load(os.path.join(GOLDEN_DIR, "acq_cases.npz"))
    x, rows = a["x"], a["results"]
    prns = [int(p) for p in rows[:, 0]]
    codes = np.stack([gpu.e1c_replica(p, 5e6, 20000) for p in prns])
    res = gpu.PcpsAcquisition(20000, 5e6).with_doppler_range(5000.0, 250.0).acquire_batch(x, 1, 20000, 20000, codes, prns)[0]
    for r, (prn, lag, dop, metric, det) in zip(res, rows):
        assert (r.prn, r.code_phase, r.doppler_hz, r.detected) == (int(prn), lag, dop, bool(det))
        assert r.peak_metric == pytest.approx(metric, rel=2e-4)


_ACQ_DUMP = """
import numpy as np, r4w_b200 as R
from tests.conftest import config_path
R.init(0)
cfg = R.load_config(config_path('e1c_8prn_60s_cn34_orbital'), cli_elevation_mask_deg=5.0)
x = R.GnssScenario(cfg, noise=True).generate_range(40_000_000, 6 * 20000)
prns = list(range(1, 17))
codes = np.stack([R.e1c_replica(p, 5e6, 20000) for p in prns])
acq = R.PcpsAcquisition(20000, 5e6).with_doppler_range(5000.0, 250.0)
res = acq.acquire_batch(x, 6, 20000, 20000, codes, prns)
print(';'.join(f'{int(r.code_phase)},{r.doppler_hz},{int(r.detected)}' for s in res for r in s))
"""


def test_register_engine_equals_shared_memory_engine(gpu):
    """fft_size 32768: the register-resident engine (rfft.cuh) and the in-shared-memory engine (fft.cuh) report the
    same (lag, Doppler bin, detected) for every (snapshot, PRN) of a noisy 34 dB-Hz scenario, present PRNs or not"""
    import subprocess
    import sys

    def run(extra):
        env = dict(os.environ)
        env.update(extra)
        out = subprocess.run([sys.executable, "-c", _ACQ_DUMP], capture_output=True, text=True, env=env, timeout=600,
                             cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        assert out.returncode == 0, out.stderr[-2000:]
        return out.stdout.strip().splitlines()[-1]

    a, b = run({}), run({"R4WB_ACQ_ENGINE": "smem"})
    assert a == b and a.count(";") == 6 * 16 - 1


def test_shared_handle_from_several_threads():
    """`acquire(&self)` is re-entrant in the reference: one handle searched from four host threads at once (ctypes drops the
    GIL during the call) gives each thread the sequential answer"""
    import threading
    import r4w_b200 as gpu
    cfg = _cfg("e1c_8prn_20s_clean")
    x = gpu.GnssScenario(cfg, noise=False).generate_range(0, 20000)
    acq = gpu.PcpsAcquisition(20000, 5e6).with_doppler_range(5000.0, 250.0)
    prns = [3, 25, 8, 2, 13, 15, 1, 7]
    codes = {p: gpu.e1c_replica(p, 5e6, 20000) for p in prns}
    want = {p: acq.acquire(x, codes[p], p) for p in prns}
    got, errs = {}, []

    def work(mine):
        try:
            for _ in range(3):
                for p in mine:
                    got[p] = acq.acquire(x, codes[p], p)
        except Exception as e:          # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=work, args=(prns[k::4],)) for k in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs
    for p in prns:
        assert (got[p].code_phase, got[p].doppler_hz, got[p].detected) == (want[p].code_phase, want[p].doppler_hz, want[p].detected)
        assert abs(got[p].peak_metric - want[p].peak_metric) <= 1e-9 * want[p].peak_metric
