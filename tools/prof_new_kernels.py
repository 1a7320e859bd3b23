#!/usr/bin/env python
"""One launch each of the round's newer kernels (k_track with 8-CTA clusters, k_synth_direct, k_compose) for an ncu capture."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import r4w_b200 as R
from r4w_b200 import sim
from tests.conftest import config_path
R.init(0)
cfg = R.load_config(config_path("e1c_8prn_20s_clean"), cli_elevation_mask_deg=5.0)
x = torch.empty(250 * 20000, dtype=torch.complex64, device="cuda")
R.GnssScenario(cfg, noise=True).generate_device(0, x.numel(), x)
prns = [s.prn for s in cfg.satellites]
e1c = np.stack([np.repeat(R.e1_code(1, p).astype(np.int8), 2) * np.tile(np.array([1, -1], np.int8), 4092) for p in prns])
chans = [dict(prn=p, code_length=8184, sample_rate=5e6, chipping_rate=2.046e6, initial_code_phase=0.0, initial_doppler=float(s.doppler_hz)) for p, s in zip(prns, cfg.satellites)]
R.TrackerBank(chans).process(x, e1c, 20000, 250)
mixed = cfg.copy(); mixed.satellites = mixed.satellites[:4]
for s, (sig, prn) in zip(mixed.satellites, (("GlonassL1of", 3), ("GpsL5", 7), ("GpsL1Ca", 9), ("GalileoE1C", 12))):
    s.signal, s.prn, s.nav_data, s.plane, s.slot = sig, prn, True, min(s.plane, 2), min(s.slot, 5)
y = torch.empty(5_000_000, dtype=torch.complex64, device="cuda")
R.GnssScenario(mixed, noise=True).generate_device(0, y.numel(), y)
class Tone:
    def __init__(s, pos, f): s.pos, s.f = pos, f
    def state_at(s, t): return sim.EmitterState(s.pos, np.zeros(3), 50.0, True)
    def generate_iq(s, t, n, fs): ph = 2 * np.pi * s.f * (t + np.arange(n) / fs); return np.cos(ph) + 1j * np.sin(ph)
    def carrier_frequency_hz(s): return 1.5e9
    def nominal_power_dbm(s): return 50.0
    def id(s): return "tone"
ecfg = sim.ScenarioConfig(duration_s=0.5, sample_rate=4e6, block_size=1_000_000, seed=3)
eng = sim.ScenarioEngine(ecfg, [Tone(sim.lla_to_ecef(0, k, 2e7), 1e3 * (k + 1)) for k in range(8)], sim.Trajectory("Static", position=(0.0, 0.0, 0.0)))
eng.generate_all()
torch.cuda.synchronize()
print("ok")
