"""End-to-end through the C-ABI with HOST buffers from ONE process driving every GPU of the box (r4wb_init_devices): synthesis of
the first 60 s of the 600 s config into pinned host memory, then PCPS over 8 x 2960 snapshots of it.  Prints one JSON line."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import r4w_b200 as R
from r4w_b200 import _lib

out = {}
cfg = R.load_config(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs", "e1c_8prn_600s_cn34_orbital.yaml"), cli_elevation_mask_deg=5.0)
n = 300_000_000
host = C.c_void_p()
for nd in (1, 0):
    got = R.init_devices(nd)
    if host.value is None:
        _lib.check(_lib.lib().r4wb_host_alloc(C.byref(host), n * 8))
        host_np = np.ctypeslib.as_array(C.cast(host, C.POINTER(C.c_float)), shape=(2 * n,)).view(np.complex64)
    sc = R.GnssScenario(cfg, noise=True)
    sc.generate_range_into(0, n, host.value)              # tables, staging buffers, peers
    t = time.perf_counter()
    for _ in range(2):
        sc.generate_range_into(0, n, host.value)
    dt = (time.perf_counter() - t) / 2
    prns = [s.prn for s in cfg.satellites]
    codes = np.stack([R.e1c_replica(p, 5e6, 20000) for p in prns])
    acq = R.PcpsAcquisition(20000, 5e6).with_doppler_range(5000.0, 250.0)
    ns = min(2960 * got, n // 20000)
    acq.acquire_batch_raw(host_np, ns, 20000, 20000, codes, prns)
    t = time.perf_counter()
    acq.acquire_batch_raw(host_np, ns, 20000, 20000, codes, prns)
    da = time.perf_counter() - t
    out[f"devices_{got}"] = {"synth_e2e_msamples_per_s": n / dt / 1e6, "acq_e2e_cells_per_s": ns * len(prns) * 41 * 20000 / da, "snapshots": ns}
    sc.close(); acq.close()
R.init_devices(1)
print(json.dumps(out))
