import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import r4w_b200 as R
from oracle import oracle as O
from tests.test_gpu_synth import _cfg, _relrms
R.init(0); O.build()
fs = float(sys.argv[1]) if len(sys.argv) > 1 else 4e6
order = sys.argv[2].split(",") if len(sys.argv) > 2 else ["e1c_8prn_20s_clean", "e1c_8prn_60s_cn34_orbital"]
os.environ["R4WB_SYNTH_PERIODIC"] = sys.argv[3] if len(sys.argv) > 3 else "1"
for name in order:
    cfg = _cfg(name).copy(); cfg.output.sample_rate = fs; cfg.output.lpf_cutoff_hz = 0.0
    L = int(round(fs * 0.004)); B = L // 4
    first, n = 100 * L, 40 * L
    sc = R.GnssScenario(cfg, noise=False)
    x = sc.generate_range(first, n)
    want = O.OracleScenario(cfg, noise=False).generate_range(first, n)
    d = np.abs(x - want)
    per = np.sqrt(np.mean(d.reshape(40, L) ** 2, axis=1)) / np.sqrt(np.mean(np.abs(want) ** 2))
    print(name, fs, "path", sc.last_path(), "relrms %.2e" % _relrms(x, want), "per-period relrms", " ".join("%.0e" % v for v in per))
    bad = np.nonzero(d > 1e-3 * np.sqrt(np.mean(np.abs(want) ** 2)))[0]
    print("   bad samples:", bad.size, bad[:20] % L if bad.size else "")
