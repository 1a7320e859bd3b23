"""A short acquire_batch on device-resident input (296 snapshots x 8 PRNs x 41 bins) for ncu captures of the PCPS kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import r4w_b200 as R
from tests.conftest import config_path
R.init(0)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 296
cfg = R.load_config(config_path("e1c_8prn_60s_cn34_orbital"), cli_elevation_mask_deg=5.0)
prns = [s.prn for s in cfg.satellites]
codes = np.stack([R.e1c_replica(p, 5e6, 20000) for p in prns])
x = torch.empty(S * 20000, dtype=torch.complex64, device="cuda")
R.GnssScenario(cfg, noise=True).generate_device(0, S * 20000, x)
acq = R.PcpsAcquisition(20000, 5e6).with_doppler_range(5000.0, 250.0); acq.set_profiling(True)
for _ in range(2):
    acq.acquire_batch_raw(x, S, 20000, 20000, codes, prns)
torch.cuda.synchronize()
print({k: round(v[0], 2) for k, v in acq.last_profile().items()})
