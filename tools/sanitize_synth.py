"""A short render through every synthesis kernel family for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import r4w_b200 as R
R.init(0)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for name, first, n in (("e1c_8prn_60s_cn34_orbital", 4_993, 123_456), ("e1c_8prn_20s_clean", 0, 700_003)):
    cfg = R.load_config(os.path.join(root, "configs", name + ".yaml"), cli_elevation_mask_deg=5.0)
    sc = R.GnssScenario(cfg, noise=True)
    x = sc.generate_range(first, n)
    print(name, "ok", float(np.abs(x).mean()), sc.last_path())
cfg = R.load_config(os.path.join(root, "configs", "e1c_8prn_60s_cn34_orbital.yaml"), cli_elevation_mask_deg=5.0)
cfg.output.duration_s = 0.0403
sc = R.GnssScenario(cfg, noise=True)
k = 0
while not sc.is_done():
    k += sc.generate_block(5000).size
print("block loop ok", k)
