"""Kernel time of k_synth on 10 s windows at several offsets into the 600 s orbital config, with the share of (block,
satellite) entries that carry the ambiguity flag (bit 1) in each window — the f64 rounding band of the reference's chip
index widens with elapsed time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from r4w_b200 import config as c
from r4w_b200.scenario import GnssScenario

cfg = c.load_config("configs/e1c_8prn_600s_cn34_orbital.yaml", 5.0)
s = GnssScenario(cfg)
n = 50_000_000
out = torch.empty(n, dtype=torch.complex64, device="cuda")
s.set_profiling(True)
s.generate_device(3_000_000_000 - n, n, out)                       # builds the whole block table once
torch.cuda.synchronize()
offsets = [int(a) for a in sys.argv[1:]] or [0, 100, 200, 260, 270, 400, 520, 530, 590]
for t0 in offsets:
    first = t0 * 5_000_000
    ms = []
    for r in range(3):
        s.generate_device(first, n, out)
        torch.cuda.synchronize()
        pr = s.last_profile()
        kn = max(pr, key=lambda k: pr[k][0])
        ms.append(pr[kn][0])
    print(f"t0 {t0:4d} s: {kn} {min(ms):.3f} ms for {n} samples = {n / min(ms) / 1e6:.1f} Gs/s", flush=True)
rng = np.random.default_rng(1)
for t0 in ((0, 270, 530) if len(sys.argv) == 1 else ()):
    flags = []
    for b in rng.integers(t0 * 1000, t0 * 1000 + 10000, 40):
        for sat in range(8):
            flags.append(int(s._debug_block_params(int(b), sat)[8]))
    flags = np.array(flags)
    print(f"t0 {t0:4d} s: flag histogram {dict(zip(*np.unique(flags, return_counts=True)))}", flush=True)
