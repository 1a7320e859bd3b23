"""Cold prologue of a scenario handle: renders a short window at sample `first` of a config on a fresh handle, so the block table /
exact-phase pass / tile records for [0, block of first] are built.  Run plain (prints the wall clock) or under
`ncu --metrics gpu__time_duration.sum` for the per-kernel launch list."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import r4w_b200 as R

name = sys.argv[1] if len(sys.argv) > 1 else "e1c_8prn_600s_cn34_orbital.yaml"
first = int(float(sys.argv[2])) if len(sys.argv) > 2 else 2_990_000_000
n = int(float(sys.argv[3])) if len(sys.argv) > 3 else 1_000_000
R.init(0)
cfg = R.load_config(os.path.join(ROOT, "configs", name), cli_elevation_mask_deg=5.0)
buf = torch.empty(n, dtype=torch.complex64, device="cuda")
w = R.GnssScenario(cfg, noise=True); w.generate_device(0, 100_000, buf); w.close()     # kernel images loaded
torch.cuda.synchronize()
sc = R.GnssScenario(cfg, noise=True)
t = time.perf_counter()
sc.generate_device(first, n, buf)
torch.cuda.synchronize()
cold = time.perf_counter() - t
t = time.perf_counter()
sc.generate_device(first, n, buf)
torch.cuda.synchronize()
warm = time.perf_counter() - t
print(f"{name} first={first} n={n}: cold {cold*1e3:.1f} ms, warm {warm*1e3:.2f} ms")
