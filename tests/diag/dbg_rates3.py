import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import r4w_b200 as R
from oracle import oracle as O
from tests.test_gpu_synth import _cfg, _relrms
R.init(0); O.build()
fs = float(sys.argv[1]) if len(sys.argv) > 1 else 4e6
os.environ["R4WB_SYNTH_PERIODIC"] = "0"
cfg = _cfg("e1c_8prn_60s_cn34_orbital").copy(); cfg.output.sample_rate = fs; cfg.output.lpf_cutoff_hz = 0.0
L = int(round(fs * 0.004)); B = L // 4
spc = 8 * fs / 1.023e6
for per in (105, 111):
    first, n = per * L, L
    for k in range(8):
        one = cfg.copy(); one.satellites = [cfg.satellites[k]]
        sc = R.GnssScenario(one, noise=False)
        x = sc.generate_range(first, n)
        w = O.OracleScenario(one, noise=False).generate_range(first, n)
        d = np.abs(x - w)
        bad = np.nonzero(d > 1e-4)[0]
        if bad.size:
            m = first + bad[0]
            blk, i = m // B, m % B
            bp = sc._debug_block_params(int(blk), 0)
            phase0 = bp[5]
            print(f"period {per} sat {k} prn {one.satellites[0].prn}: bad slots {bad[:8]} (n={bad.size}) block {blk} i {i}  flags {int(bp[8])} U_hc {bp[1]:.9f} phase0 {phase0!r} e0 {bp[6]}")
            for q in range(8 * i - 70, 8 * i + 1, 1):
                g = blk * B * 8 + q
                cf = phase0 + g / spc
                fr = (2 * cf) % 1.0
                if fr < 2e-6 or fr > 1 - 2e-6:
                    print(f"     oversample q={q} cf={cf!r} half-chip frac={fr:.3e}")
