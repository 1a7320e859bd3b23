cd "$(dirname "$0")/.."; timeout 900 python -m pytest tests/test_sink.py tests/test_cabi.py -q -m gpu 2>&1 | tail -15
python -m r4w_b200.sink --config configs/e1c_8prn_20s_clean.yaml --output /tmp/e1c.cf32 2>&1 | tail -8
python - <<'PY'
import time, os
from r4w_b200 import config as c, _lib
from r4w_b200.scenario import GnssScenario
cfg = c.load_config('configs/e1c_8prn_20s_clean.yaml', 5.0)
for fmt, name in ((_lib.FMT_CF32, 'cf32'), (_lib.FMT_CI8, 'ci8')):
    s = GnssScenario(cfg)
    for rep in range(2):
        s.reset(); t = time.time(); n, b, p = s.write_file('/dev/shm/x.' + name, fmt); dt = time.time() - t
        print(name, 'tmpfs', n, b, round(dt, 3), 's', round(n / dt / 1e6, 1), 'Msamples/s', round(b / dt / 1e9, 2), 'GB/s')
    s.close(); os.remove('/dev/shm/x.' + name)
PY
