"""Host replay of the lattice synthesis kernel (r4w_b200/csrc/synth_lattice.cuh, k_synth_lat) against the oracle and against
the literal per-sample evaluation of k_synth — runs without a GPU."""
import os

import numpy as np
import pytest

from tests.conftest import config_path

TOL = 1e-5


def _cfg(name):
    from r4w_b200.config import load_config
    return load_config(config_path(name), cli_elevation_mask_deg=5.0)


def _relrms(a, b):
    return float(np.sqrt(np.sum(np.abs(a - b) ** 2) / np.sum(np.abs(b) ** 2)))


@pytest.mark.parametrize("name,first,n", [
    ("e1c_prn3_20s_withdoppler", 0, 12000),
    ("e1c_8prn_20s_clean", 4993, 10014),                  # partial blocks at both ends: same kernel, masked stores
    ("e1c_8prn_60s_cn34_orbital", 0, 20000),              # block starts with a changed delay (yfix), varying Doppler
    ("e1c_8prn_60s_cn34_orbital", 7_500_000, 10000),
    ("e1c_60s_all_prns", 299_990_000, 10000),             # range ramp; the scenario's last blocks
    ("e1c_60s_cn34_effects", 100_000_000, 10000),
])
def test_lattice_replay_matches_oracle(oracle, emu, name, first, n):
    cfg = _cfg(name)
    e = emu.EmuScenario(cfg, noise=False)
    assert e.set_lattice(True)                            # every e1c_*.yaml sits on the 5 MHz lattice (q = 2500, p = 1023)
    got = e.generate_range(first, n)
    want = oracle.OracleScenario(cfg, noise=False).generate_range(first, n)
    assert _relrms(got, want) <= TOL
    assert np.max(np.abs(got - want)) < 1e-4 * np.max(np.abs(want))


def test_lattice_noise_is_the_same_stream(emu):
    """both kernels draw the noise of global sample m from Philox counter m >> 1"""
    cfg = _cfg("e1c_8prn_60s_cn34_orbital")
    a = emu.EmuScenario(cfg, noise=True); a.set_lattice(True)
    b = emu.EmuScenario(cfg, noise=True)
    x, y = a.generate_range(1_000_000, 10_000), b.generate_range(1_000_000, 10_000)
    assert np.max(np.abs(x - y)) < 2e-5                   # signal part differs by f32 rounding only; sigma = 12.6


def test_lattice_model_does_not_apply_to_other_lattices(emu):
    cfg = _cfg("e1c_8prn_20s_clean")
    cfg.output.sample_rate = 6e6                          # 6000-sample blocks = 6 q: not the two-half lattice
    assert not emu.EmuScenario(cfg, noise=False).set_lattice(True)
    cfg.output.sample_rate = 4e6                          # q = 2000, K = 4
    e = emu.EmuScenario(cfg, noise=False)
    assert e.set_lattice(True)


def test_lattice_4mhz_matches_oracle(oracle, emu):
    cfg = _cfg("e1c_8prn_60s_cn34_orbital")
    cfg.output.sample_rate = 4e6
    e = emu.EmuScenario(cfg, noise=False)
    assert e.set_lattice(True)
    got = e.generate_range(1_000_000, 12_000)
    want = oracle.OracleScenario(cfg, noise=False).generate_range(1_000_000, 12_000)
    assert _relrms(got, want) <= TOL


@pytest.mark.slow
def test_lattice_patches_equal_literal_evaluation(emu, monkeypatch):
    """530 s into the 600 s config ~2 % of the (block, satellite) entries have an oversample inside the f64 rounding band of a
    half-chip boundary.  k_synth re-evaluates those windows tap by tap with the reference's f64 expression; the lattice kernel
    resolves the disputed oversample once per block and patches +-2 h[g - q*] into the windows that hold it.  Same samples."""
    cfg = _cfg("e1c_8prn_600s_cn34_orbital")
    first, n = 2_650_000_000, 500_000
    a = emu.EmuScenario(cfg, noise=False, closed_form_phase=True); a.set_lattice(True)
    b = emu.EmuScenario(cfg, noise=False, closed_form_phase=True)
    x, y = a.generate_range(first, n), b.generate_range(first, n)
    assert a.patched() > 0 and b.n_ambiguous > 0
    assert np.max(np.abs(x - y)) < 2e-5 and _relrms(x, y) < 1e-6
    monkeypatch.setenv("R4WB_EMU_LAT_NO_PATCH", "1")      # without the patches the disputed windows are off by 2 h[k] A
    c = emu.EmuScenario(cfg, noise=False, closed_form_phase=True); c.set_lattice(True)
    z = c.generate_range(first, n)
    assert np.max(np.abs(z - y)) > 1e-3
