"""Times the general synthesis kernel on an orbital config (device-resident output); used under ncu for the source page."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from r4w_b200 import config as c
from r4w_b200.scenario import GnssScenario

name = sys.argv[1] if len(sys.argv) > 1 else "configs/e1c_8prn_60s_cn34_orbital.yaml"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 50_000_000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
cfg = c.load_config(name, 5.0)
s = GnssScenario(cfg)
out = torch.empty(n, dtype=torch.complex64, device="cuda")
s.set_profiling(True)
for r in range(reps):
    torch.cuda.synchronize(); t = time.time()
    s.generate_device(0, n, out)
    torch.cuda.synchronize(); dt = time.time() - t
    print(r, "path", s.last_path(), round(dt * 1e3, 3), "ms wall", round(n / dt / 1e9, 1), "Gs/s", s.last_profile())
