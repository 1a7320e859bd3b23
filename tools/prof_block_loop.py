"""The generate_block loop from the compiled driver (tools/ubench/block_loop.c) for one setting of the ring's environment hooks:
   R4WB_RING_CHUNK=<samples> python tools/prof_block_loop.py [seconds]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import r4w_b200 as R
from r4w_b200 import _lib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sec = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
from r4w_b200.config import load_config
cfg = load_config(os.path.join(ROOT, "configs", "e1c_8prn_60s_cn34_orbital.yaml"), cli_elevation_mask_deg=5.0)
cfg.output.duration_s = sec
sc = R.GnssScenario(cfg, noise=True)
nb = sc.block_size()
sc.generate_block(nb)
L = _lib.lib()
BL = C.CDLL(os.path.join(ROOT, "tools", "ubench", "libblock_loop.so")).r4wb_block_loop
BL.restype = C.c_int
BL.argtypes = [C.c_void_p] * 4 + [C.c_uint64, C.c_void_p, C.c_int, C.c_uint64] + [C.POINTER(C.c_uint64)] * 2 + [C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
fp = lambda f: C.cast(f, C.c_void_p)
buf = np.empty(nb, np.complex64)
out = []
for mode in (0, 1, 2, 0):
    sc.reset()
    got, calls, s_, fold = C.c_uint64(0), C.c_uint64(0), C.c_double(0), C.c_uint64(0)
    _lib.check(BL(fp(L.r4wb_scenario_generate_block), fp(L.r4wb_scenario_generate_block_view), fp(L.r4wb_scenario_is_done), sc._h, nb,
                  buf.ctypes.data, mode, 8, C.byref(got), C.byref(calls), C.byref(s_), C.byref(fold)))
    out.append(f"mode {mode}: {s_.value / calls.value * 1e6:.3f} us/call {got.value / s_.value / 1e6:.0f} Ms/s")
print(f"chunk={os.environ.get('R4WB_RING_CHUNK', 'default')}: " + "; ".join(out))
