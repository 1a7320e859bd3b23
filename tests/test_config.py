"""Host logic: the YAML schema mirror (GnssScenarioConfig, gnss/scenario_config.rs) and the sharding helpers."""
import glob
import os

import numpy as np
import pytest

from tests.conftest import CONFIG_DIR, config_path


def test_all_reference_yamls_parse():
    from r4w_b200.config import load_config
    files = sorted(glob.glob(os.path.join(CONFIG_DIR, "e1c_*.yaml")))
    assert len(files) == 11
    for f in files:
        cfg = load_config(f)
        assert cfg.output.sample_rate == 5e6 and cfg.block_size() == 5000
        assert all(s.signal == "GalileoE1C" for s in cfg.satellites)
        assert cfg.receiver.antenna.kind in ("Patch", "Isotropic", "Hemispherical", "ChokeRing")


def test_named_configs():
    from r4w_b200.config import load_config, HAS_DOPPLER, HAS_RANGE_RATE
    c1 = load_config(config_path("e1c_prn3_20s_withdoppler"))
    assert [s.prn for s in c1.satellites] == [3] and c1.total_samples() == 100_000_000
    assert abs(c1.satellites[0].doppler_hz - (-457.3938)) < 1e-3 and c1.satellites[0].cn0_dbhz == 65.0
    c2 = load_config(config_path("e1c_8prn_20s_clean"))
    assert [s.prn for s in c2.satellites] == [3, 25, 8, 2, 5, 16, 13, 15]
    assert not any(s.orbital_dynamics for s in c2.satellites)
    c3 = load_config(config_path("e1c_8prn_60s_cn34_orbital"))
    assert all(s.orbital_dynamics and s.cn0_dbhz == 34.0 for s in c3.satellites) and c3.total_samples() == 300_000_000
    c4 = load_config(config_path("e1c_60s_all_prns"))
    pod, keep = c4.to_pod()
    assert all(pod.sats[k].has & HAS_RANGE_RATE and pod.sats[k].has & HAS_DOPPLER for k in range(pod.n_sats))
    c5 = load_config(config_path("e1c_8prn_600s_cn34_orbital"))
    assert c5.total_samples() == 3_000_000_000 and c5.output.seed == 42
    assert c5.receiver.antenna.kind == "Patch"      # `antenna: !Patch {..}` externally tagged enum


def test_cli_elevation_mask_override():
    """crates/r4w-cli/src/main.rs:4133: the CLI always overwrites the YAML's elevation mask with its default 5.0"""
    from r4w_b200.config import load_config
    assert load_config(config_path("e1c_8prn_20s_clean"), cli_elevation_mask_deg=5.0).receiver.elevation_mask_deg == 5.0


def test_yaml_dialect():
    from r4w_b200.config import loads_config
    text = open(config_path("e1c_prn3_20s_withdoppler")).read()
    cfg = loads_config(text)
    assert cfg.environment.ephemeris_source == "Nominal"
    t2 = text.replace("  orbital_dynamics: false\n", "").replace("  iono_delay_m:", "  unknown_key: 1\n  iono_delay_m:")
    assert t2 != text
    cfg2 = loads_config(t2)                          # unknown keys ignored, missing orbital_dynamics -> false
    assert cfg2.satellites[0].orbital_dynamics is False
    tri = loads_config(open(config_path("e1c_8prn_60s_mach3_ftwayne_berne")).read())
    assert tri.receiver.trajectory is not None and tri.receiver.trajectory.start.lat_deg != tri.receiver.trajectory.end.lat_deg


def test_segments_cover_and_align():
    from r4w_b200.dist import segment_for_rank, snapshots_for_rank
    for total, align in [(100_000_000, 20000), (3_000_000_000, 5000), (100_001, 5000), (12345, 1), (4999, 5000)]:
        for world in (1, 2, 3, 4, 8):
            segs = [segment_for_rank(total, align, r, world) for r in range(world)]
            assert segs[0][0] == 0 and sum(n for _, n in segs) == total
            for (f0, n0), (f1, _) in zip(segs, segs[1:]):
                assert f0 + n0 == f1 and f1 % align == 0
            sizes = [n for _, n in segs[:-1]]
            if sizes:
                assert max(sizes) - min(sizes) <= align
    assert [snapshots_for_rank(10, r, 4) for r in range(4)] == [(0, 3), (3, 3), (6, 2), (8, 2)]
    with pytest.raises(ValueError):
        segment_for_rank(10, 1, 2, 2)


def test_presets_and_constellation_tables():
    """GnssScenarioPreset::to_config + lookup_prn (scenario_config.rs:199-241, 581-700): PRNs as the reference's comments state,
    default start time = 2026-02-04 20:00 UTC in GPS seconds"""
    from r4w_b200.config import PRESETS, preset_config, gps_time_from_utc, gps_l1ca, galileo_e1
    assert gps_time_from_utc(2026, 2, 4, 20, 0, 0.0) == 1454270418.0
    assert gps_time_from_utc(1980, 1, 6, 0, 0, 0.0) == 18.0
    assert [gps_l1ca(p, s).prn for p, s in ((4, 2), (2, 3), (1, 4), (3, 3), (2, 4), (1, 3), (4, 1), (5, 2))] == [5, 17, 26, 6, 19, 12, 10, 9]
    assert [galileo_e1(p, s).prn for p, s in ((0, 7), (1, 5), (0, 0), (1, 6), (1, 4), (0, 6))] == [16, 12, 31, 33, 11, 25]
    with pytest.raises(ValueError):
        gps_l1ca(0, 4)                                      # no satellite in slot A-5
    sizes = {n: (len(preset_config(n).satellites), preset_config(n).total_samples()) for n in PRESETS}
    assert sizes == {"OpenSky": (8, 5000), "UrbanCanyon": (8, 5000), "Driving": (7, 50000), "Walking": (6, 50000),
                     "HighDynamics": (8, 5000), "MultiConstellation": (11, 5000)}
    assert preset_config("UrbanCanyon").receiver.elevation_mask_deg == 15.0
    assert preset_config("OpenSky").to_pod()[0].output.start_time_gps_s == 1454270418.0
