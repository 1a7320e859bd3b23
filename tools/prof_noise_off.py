"""k_synth_lat with and without the noise term (what the Philox / Box-Muller phase costs on top of the satellite loop)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from r4w_b200 import config as c
from r4w_b200.scenario import GnssScenario
cfg = c.load_config("configs/e1c_8prn_60s_cn34_orbital.yaml", 5.0)
n = 50_000_000
out = torch.empty(n, dtype=torch.complex64, device="cuda")
for noise in (True, False):
    s = GnssScenario(cfg, noise=noise)
    s.set_profiling(True)
    ms = []
    for r in range(4):
        s.generate_device(0, n, out); torch.cuda.synchronize()
        ms.append(s.last_profile()["k_synth_lat"][0])
    print(f"noise={noise}: k_synth_lat {min(ms):.3f} ms per {n} samples = {n / min(ms) / 1e6:.1f} Gs/s", flush=True)
