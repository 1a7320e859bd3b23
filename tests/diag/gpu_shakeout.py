"""First-contact GPU check: synthesis and acquisition against the oracle + quick timings.  Diagnostic, not a test."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import r4w_b200 as R
from r4w_b200.config import load_config
from oracle import oracle as O

R.init(0)
print("version", R.version(), torch.cuda.get_device_name(0))

def relrms(a, b):
    return float(np.sqrt(np.sum(np.abs(a - b) ** 2) / np.sum(np.abs(b) ** 2)))

for name, first, n in [("e1c_8prn_20s_clean", 0, 20000), ("e1c_8prn_20s_clean", 4993, 10014), ("e1c_8prn_20s_clean", 99_980_000, 20000),
                       ("e1c_prn3_20s_withdoppler", 50_000_000, 15000), ("e1c_60s_all_prns", 10_000_000, 10000),
                       ("e1c_8prn_60s_cn34_orbital", 5_000_000, 10000)]:
    cfg = load_config(os.path.join(ROOT, "configs", name + ".yaml"), cli_elevation_mask_deg=5.0)
    sc = R.GnssScenario(cfg, noise=False)
    g = sc.generate_range(first, n)
    o = O.OracleScenario(cfg, noise=False).generate_range(first, n)
    print(f"synth {name} [{first},+{n}) relrms {relrms(g, o):.3e} max {np.abs(g-o).max():.3e} power {sc.last_power_sum():.6e} vs {np.sum(np.abs(O.to_cf32(o))**2):.6e}")

# sequential API
cfg = load_config(os.path.join(ROOT, "configs", "e1c_8prn_20s_clean.yaml"), cli_elevation_mask_deg=5.0)
cfg.output.duration_s = 0.01
sc = R.GnssScenario(cfg, noise=False); orc = O.OracleScenario(cfg, noise=False)
for bs in (5000, 5000, 1234, 8000, 5000):
    g = sc.generate_block(bs); o = orc.generate_block(bs)
    print("  generate_block", bs, g.size, o.size, "relrms %.3e" % relrms(g, o))

# noise statistics
cfg = load_config(os.path.join(ROOT, "configs", "e1c_prn3_20s_withdoppler.yaml"), cli_elevation_mask_deg=5.0)
sn = R.GnssScenario(cfg, noise=True); sq = R.GnssScenario(cfg, noise=False)
a = sn.generate_range(0, 1_000_000) - sq.generate_range(0, 1_000_000)
print("noise std re/im", a.real.std(), a.imag.std(), "expected 12.5954; mean", a.mean(), "corr", np.mean(a.real*a.imag))

# acquisition
cfg = load_config(os.path.join(ROOT, "configs", "e1c_8prn_20s_clean.yaml"), cli_elevation_mask_deg=5.0)
x = O.to_cf32(O.OracleScenario(cfg, noise=False).generate_range(0, 20000))
acq = R.PcpsAcquisition(20000, 5e6).with_doppler_range(5000.0, 250.0)
oacq = O.OraclePcps(20000, 5e6).with_doppler_range(5000.0, 250.0)
for prn in (3, 25, 8, 2, 13, 15, 5, 16, 1):
    rep = O.e1c_replica(prn, 5e6, 20000)
    r = acq.acquire(x, rep, prn); o = oacq.acquire(x.astype(np.complex128), rep, prn)
    print(f"acq prn {prn}: gpu ({r.code_phase:.0f},{r.doppler_hz:.0f},{r.peak_metric:.4f},{r.detected}) oracle ({o.code_phase:.0f},{o.doppler_hz:.0f},{o.peak_metric:.4f},{bool(o.detected)}) guard {acq.guard_count()}")
# KAT
code = O.gps_ca_code(1); i = np.arange(1023)
sig = code[(i + 1023 - 100) % 1023] * np.exp(2j * np.pi * 1000.0 * (i / 1023.0))
r = R.PcpsAcquisition(1023, 1023.0).with_doppler_range(5000.0, 500.0).with_threshold(2.0).acquire(sig, code, 1)
print("KAT no_noise:", r)
gr = R.PcpsAcquisition(1023, 1023.0).with_doppler_range(2000.0, 500.0).acquire_grid(code.astype(np.complex128), code)
print("KAT grid peak:", gr.find_peak())
og, olin = O.OraclePcps(1023, 1023.0).with_doppler_range(2000.0, 500.0).acquire_grid(code.astype(np.complex128), code)
print("grid max rel err vs oracle", np.abs(gr.power - og).max() / og.max())

# timings
cfg = load_config(os.path.join(ROOT, "configs", "e1c_8prn_20s_clean.yaml"), cli_elevation_mask_deg=5.0)
sc = R.GnssScenario(cfg, noise=True)
n = 50_000_000
buf = torch.empty(n, dtype=torch.complex64, device="cuda")
for it in range(3):
    torch.cuda.synchronize(); t = time.time()
    sc.generate_device(0, n, buf)
    torch.cuda.synchronize(); dt = time.time() - t
    print(f"synth 8prn {n/1e6:.0f} Msamples in {dt*1e3:.2f} ms -> {n/dt/1e6:.1f} Msamples/s, {8*n/dt/1e9:.1f} GB/s")
acq = R.PcpsAcquisition(20000, 5e6).with_doppler_range(5000.0, 250.0)
prns = [3, 25, 8, 2, 5, 16, 13, 15]
codes = np.stack([O.e1c_replica(p, 5e6, 20000) for p in prns])
S = 64
for it in range(3):
    torch.cuda.synchronize(); t = time.time()
    res = acq.acquire_batch_raw(buf, S, 20000, 20000, codes, prns)
    torch.cuda.synchronize(); dt = time.time() - t
    cells = S * len(prns) * 41 * 20000
    print(f"acq {S} snapshots x 8 prn: {dt*1e3:.2f} ms -> {cells/dt/1e9:.2f} Gcells/s guard {acq.guard_count()}")
print("first snapshot:", [(int(res[c].code_phase), res[c].doppler_hz, round(res[c].peak_metric, 2)) for c in range(8)])
print("launches", R.kernel_launches())
