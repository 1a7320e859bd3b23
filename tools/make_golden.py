"""Regenerates tests/golden/synth_windows.npz and acq_cases.npz from the oracle (run in the build container).
The reference is Rust and cannot run here, so golden vectors are oracle outputs; the oracle itself is pinned by
tests/test_oracle_kats.py against the reference's own known-answer tests."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O
from r4w_b200.config import load_config

out = {}
for name, first, n in [("e1c_prn3_20s_withdoppler", 99_995_000, 5000), ("e1c_8prn_20s_clean", 0, 6000),
                       ("e1c_8prn_20s_clean", 73_004_990, 5020), ("e1c_60s_all_prns", 200_000_000, 5000),
                       ("e1c_8prn_60s_cn34_orbital", 1_000_000, 5000)]:
    cfg = load_config(os.path.join(ROOT, "configs", name + ".yaml"), cli_elevation_mask_deg=5.0)
    out[f"{name}@{first}_iq"] = O.OracleScenario(cfg, noise=False).generate_range(first, n).astype(np.complex64)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "synth_windows.npz"), **out)

# acquisition: noisy oracle snapshot + oracle results for PRN 1-50
cfg = load_config(os.path.join(ROOT, "configs", "e1c_8prn_20s_clean.yaml"), cli_elevation_mask_deg=5.0)
x = O.to_cf32(O.OracleScenario(cfg, noise=True).generate_range(0, 20000))
acq = O.OraclePcps(20000, 5e6).with_doppler_range(5000.0, 250.0)
rows = []
for prn in range(1, 51):
    r = acq.acquire(x.astype(np.complex128), O.e1c_replica(prn, 5e6, 20000), prn)
    rows.append((prn, r.code_phase, r.doppler_hz, r.peak_metric, r.detected))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "acq_cases.npz"), x=x, results=np.array(rows, np.float64))
print("golden written", {k: v.shape for k, v in out.items()})
