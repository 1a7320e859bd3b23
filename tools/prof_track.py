#!/usr/bin/env python
"""Tracking-bank throughput: C channels x P code periods of GPS L1 C/A (5 000 samples each) on one device-resident stream."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import r4w_b200 as R
from r4w_b200.config import load_config
from tests.conftest import config_path
R.init(0)
C_ = int(sys.argv[1]) if len(sys.argv) > 1 else 296
P_ = int(sys.argv[2]) if len(sys.argv) > 2 else 500
cfg = load_config(config_path("e1c_prn3_20s_withdoppler"), cli_elevation_mask_deg=5.0).copy()
s = cfg.satellites[0]; s.signal, s.prn, s.nav_data = "GpsL1Ca", 7, True; s.plane, s.slot = min(s.plane, 5), min(s.slot, 5)
x = torch.empty(P_ * 5000, dtype=torch.complex64, device="cuda")
R.GnssScenario(cfg, noise=True).generate_device(0, P_ * 5000, x)
code = R.gps_ca_code(7).astype(np.int8)
chans = [dict(prn=7, code_length=1023, sample_rate=5e6, chipping_rate=1.023e6, initial_code_phase=float(k % 1023), initial_doppler=-457.0 + k) for k in range(C_)]
codes = np.tile(code, (C_, 1))
for rep in range(3):
    bank = R.TrackerBank(chans)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st = bank.process(x, codes, 5000, P_)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{C_} channels x {P_} periods: {dt*1e3:.1f} ms  -> {C_*P_/dt/1e3:.1f} k channel-periods/s, {C_*P_*5000/dt/1e9:.2f} Gsample-correlations/s")
