"""The reference's block loop through the C-ABI (generate_block(block_size) until done) for several ring chunk sizes."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import r4w_b200 as R
from r4w_b200 import _lib
R.init(0)
cfg = R.load_config(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs", "e1c_8prn_60s_cn34_orbital.yaml"), cli_elevation_mask_deg=5.0)
cfg.output.duration_s = 10.0
for chunk in [int(a) for a in sys.argv[1:]] or [1 << 20]:
    os.environ["R4WB_RING_CHUNK"] = str(chunk)
    sc = R.GnssScenario(cfg, noise=True)
    nb = sc.block_size()
    sc.generate_block(nb); sc.reset()
    L = _lib.lib(); gb, done, h = L.r4wb_scenario_generate_block, L.r4wb_scenario_is_done, sc._h
    buf = np.empty(nb, np.complex64); p = C.c_void_p(buf.ctypes.data); wr = C.c_uint64(0); pw = C.byref(wr)
    best = 1e9
    for rep in range(3):
        sc.reset()
        t = time.perf_counter(); got = 0
        while not done(h):
            gb(h, nb, p, 0, 0, pw); got += wr.value
        best = min(best, time.perf_counter() - t)
    print(f"chunk {chunk:8d} samples: {got / best / 1e6:8.1f} Msamples/s, {best / (got // nb) * 1e6:.2f} us per call", flush=True)
    sc.close()
