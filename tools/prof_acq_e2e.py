#!/usr/bin/env python
"""acquire_batch from pinned host memory: wall time per call (e2e leg of the bench), for A/B of the H2D pipeline."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import r4w_b200 as R
from r4w_b200 import _lib
from tests.conftest import config_path
R.init(0)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
cfg = R.load_config(config_path("e1c_8prn_20s_clean"), cli_elevation_mask_deg=5.0)
prns = [s.prn for s in cfg.satellites]
codes = np.stack([R.e1c_replica(p, 5e6, 20000) for p in prns])
n = S * 20000
host = C.c_void_p(); _lib.check(_lib.lib().r4wb_host_alloc(C.byref(host), n * 8))
host_np = np.ctypeslib.as_array(C.cast(host, C.POINTER(C.c_float)), shape=(2 * n,)).view(np.complex64)
sc = R.GnssScenario(cfg, noise=True); sc.generate_range_into(0, n, host.value)
acq = R.PcpsAcquisition(20000, 5e6).with_doppler_range(5000.0, 250.0); acq.set_profiling(True)
dev = torch.from_numpy(host_np).cuda()
side = torch.cuda.Stream()
for src, name in ((dev, "device"), (host_np, "host"), (host_np, "host on a non-default stream")):
    if name.endswith("stream"):
        _lib.set_stream(side.cuda_stream)
    acq.acquire_batch_raw(src, min(S, 64), 20000, 20000, codes, prns)
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        acq.acquire_batch_raw(src, S, 20000, 20000, codes, prns)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(name, f"{dt*1e3:.1f} ms", {k: round(v[0], 1) for k, v in acq.last_profile().items()})
