#!/usr/bin/env python
"""Render e1c_8prn_20s_clean (1e8 samples, noise on) into HBM a few times; CUDA-event time per render.  Used plain and under ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import r4w_b200 as R
from tests.conftest import config_path
name = sys.argv[1] if len(sys.argv) > 1 else "e1c_8prn_20s_clean"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
R.init(0)
cfg = R.load_config(config_path(name), cli_elevation_mask_deg=5.0)
n = min(100_000_000, int(round(cfg.output.duration_s * cfg.output.sample_rate)))
sc = R.GnssScenario(cfg, noise=True)
out = torch.empty(n, dtype=torch.complex64, device="cuda")
for _ in range(2):
    sc.generate_device(0, n, out)
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sc.generate_device(0, n, out); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(name, "path", sc.last_path(), "ms", [round(t, 4) for t in ts], "Gsamples/s", round(n / min(ts) / 1e6, 1))
