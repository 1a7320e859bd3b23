import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import r4w_b200 as R
from oracle import oracle as O
from tests.test_gpu_synth import _cfg, _relrms
R.init(0); O.build()
fs = float(sys.argv[1]) if len(sys.argv) > 1 else 4e6
for name in ("e1c_8prn_60s_cn34_orbital", "e1c_8prn_20s_clean"):
    cfg = _cfg(name).copy(); cfg.output.sample_rate = fs; cfg.output.lpf_cutoff_hz = 0.0
    B = int(np.ceil(fs * 1e-3)); L = 4 * B
    for first, n in ((0, 2 * B), (100 * L, 2 * B), (100 * L, 40 * L)):
        os.environ["R4WB_SYNTH_PERIODIC"] = "0"
        sc = R.GnssScenario(cfg, noise=False)
        x = sc.generate_range(first, n)
        m = min(n, 2 * B)
        want = O.OracleScenario(cfg, noise=False).generate_range(first + n - m, m)
        d = np.abs(x[n - m:] - want)
        print(name, fs, first, n, "path", sc.last_path(), "relrms %.2e" % _relrms(x[n - m:], want), "max at", int(np.argmax(d)), "%.3e" % d.max(),
              "rms by quarter block", ["%.1e" % np.sqrt(np.mean(d[k * B // 4:(k + 1) * B // 4] ** 2)) for k in range(4)])
    one = cfg.copy(); one.satellites = [cfg.satellites[0]]
    x = R.GnssScenario(one, noise=False).generate_range(100 * L, 2 * B)
    want = O.OracleScenario(one, noise=False).generate_range(100 * L, 2 * B)
    r = x / want
    print("  single sat: relrms %.2e" % _relrms(x, want), "median ratio", np.median(r.real), np.median(r.imag), "phase err rad", float(np.median(np.angle(r))))
