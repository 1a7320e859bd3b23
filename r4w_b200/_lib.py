"""ctypes binding of libr4w_b200.so (include/r4w_b200.h) — the C-ABI boundary of the B200 path.

There is no CPU fallback: if the shared library is missing or no CUDA device is usable, every compute call
raises.  Nothing in this package imports `oracle/` (test infrastructure).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from .config import AcqResultPod, SatStatusPod, ScenarioCfgPod

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libr4w_b200.so")

ERRORS = {0: "Ok", 1: "NullPointer", 2: "InvalidSize", 3: "BufferFull", 4: "BufferEmpty", 5: "InvalidParameter",
          6: "AllocationFailed", 7: "NotSupported", 100: "Cuda"}

MEM_HOST, MEM_DEVICE = 0, 1
FMT_CF32, FMT_CF64, FMT_CI16, FMT_CI8, FMT_CU8 = 0, 1, 2, 3, 4


class R4wB200Error(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"r4w_b200: {ERRORS.get(code, code)} ({code}): {message}")
        self.code = code


def build(force: bool = False) -> str:
    """Compile csrc/ for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    if force:
        cmd.append("-B")
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    return LIB_PATH


# every symbol include/r4w_b200.h declares: name -> (restype, argtypes)
_vp, _u64, _u32, _u8, _dbl, _int = C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint8, C.c_double, C.c_int
SYMBOLS = {
    "r4wb_version": (C.c_char_p, []),
    "r4wb_last_error": (C.c_char_p, []),
    "r4wb_init": (_int, [_int]),
    "r4wb_device_count": (_int, [C.POINTER(_int)]),
    "r4wb_init_devices": (_int, [_int]),
    "r4wb_devices_initialised": (_int, []),
    "r4wb_set_stream": (_int, [_vp]),
    "r4wb_host_alloc": (_int, [C.POINTER(_vp), C.c_size_t]),
    "r4wb_host_free": (_int, [_vp]),
    "r4wb_kernel_launches": (_u64, []),
    "r4wb_scenario_create": (_int, [C.POINTER(ScenarioCfgPod), C.POINTER(_vp)]),
    "r4wb_scenario_destroy": (None, [_vp]),
    "r4wb_scenario_total_samples": (_u64, [_vp]),
    "r4wb_scenario_block_size": (_u64, [_vp]),
    "r4wb_scenario_is_done": (_int, [_vp]),
    "r4wb_scenario_progress": (_dbl, [_vp]),
    "r4wb_scenario_reset": (_int, [_vp]),
    "r4wb_scenario_current_sample": (_u64, [_vp]),
    "r4wb_scenario_generate_block": (_int, [_vp, _u64, _vp, _int, _int, C.POINTER(_u64)]),
    "r4wb_scenario_generate_block_view": (_int, [_vp, _u64, _int, C.POINTER(_vp), C.POINTER(_u64)]),
    "r4wb_scenario_generate": (_int, [_vp, _u64, _u64, _vp, _int, _int]),
    "r4wb_scenario_generate_rest": (_int, [_vp, _vp, _u64, _int, _int, C.POINTER(_u64)]),
    "r4wb_scenario_write_file": (_int, [_vp, C.c_char_p, _int, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_dbl)]),
    "r4wb_scenario_last_power_sum": (_int, [_vp, C.POINTER(_dbl)]),
    "r4wb_scenario_last_path": (C.c_uint32, [_vp]),
    "r4wb_scenario_set_profiling": (_int, [_vp, _int]),
    "r4wb_scenario_last_profile": (_int, [_vp, _vp, _vp]),
    "r4wb_scenario_status": (_int, [_vp, C.POINTER(SatStatusPod), _u32, C.POINTER(_u32)]),
    "r4wb_e1_code": (_int, [_u32, _u8, _vp, _u64]),
    "r4wb_gps_ca_code": (_int, [_u8, _vp, _u64]),
    "r4wb_gps_l5_code": (_int, [_u8, _vp, _u64]),
    "r4wb_glonass_code": (_int, [_vp, _u64]),
    "r4wb_e1c_secondary": (_int, [_vp, _u64]),
    "r4wb_e1c_replica": (_int, [_u8, _dbl, _vp, _u64]),
    "r4wb_pcps_create": (_int, [_u64, _dbl, C.POINTER(_vp)]),
    "r4wb_pcps_destroy": (None, [_vp]),
    "r4wb_pcps_set_doppler_range": (_int, [_vp, _dbl, _dbl]),
    "r4wb_pcps_set_threshold": (_int, [_vp, _dbl]),
    "r4wb_pcps_set_coherent_periods": (_int, [_vp, _u64]),
    "r4wb_pcps_fft_size": (_u64, [_vp]),
    "r4wb_pcps_num_doppler_bins": (_u32, [_vp]),
    "r4wb_pcps_acquire": (_int, [_vp, _vp, _int, _u64, _vp, _u64, _u8, C.POINTER(AcqResultPod)]),
    "r4wb_pcps_acquire_batch": (_int, [_vp, _vp, _int, _int, _u64, _u64, _u64, _vp, _u64, _vp, _u32, _vp]),
    "r4wb_pcps_acquire_grid": (_int, [_vp, _vp, _int, _u64, _vp, _u64, _vp, _u64]),
    "r4wb_pcps_guard_count": (_u64, [_vp]),
    "r4wb_composer_create": (_int, [_u32, _dbl, _dbl, _u64, C.POINTER(_vp)]),
    "r4wb_composer_destroy": (None, [_vp]),
    "r4wb_composer_reset": (_int, [_vp]),
    "r4wb_composer_block": (_int, [_vp, _vp, _int, _int, _u64, _vp, _vp, _vp, _vp, _int, _int]),
    "r4wb_composer_phases": (_int, [_vp, _vp, _u32]),
    "r4wb_track_create": (_int, [_vp, _u32, C.POINTER(_vp)]),
    "r4wb_track_destroy": (None, [_vp]),
    "r4wb_track_process": (_int, [_vp, _vp, _int, _int, _u64, _u64, _u64, _vp, _u64, _vp]),
    "r4wb_track_state_get": (_int, [_vp, _vp, _u32]),
    "r4wb_track_nav_bits": (_int, [_vp, _u32, _vp, _u64, C.POINTER(_u64)]),
    "r4wb_pcps_set_profiling": (_int, [_vp, _int]),
    "r4wb_pcps_last_profile": (_int, [_vp, _vp, _vp]),
    # test hook, not part of the drop-in surface
    "r4wb_debug_block_params": (_int, [_vp, _u64, _u32, _vp]),
}

_lib = None


def lib():
    """The loaded library.  Raises if it has not been built — there is no fallback implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise R4wB200Error(7, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                  f"(make -C r4w_b200/csrc); r4w_b200 has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise R4wB200Error(rc, (lib().r4wb_last_error() or b"").decode("utf-8", "replace"))


_initialised = False


def init(device: int = -1):
    """cudaSetDevice(device) + context warm-up.  Raises R4wB200Error(Cuda) when no device is usable."""
    global _initialised
    check(lib().r4wb_init(device))
    _initialised = True


def init_devices(n_gpus: int = 0) -> int:
    """Multi-GPU inside the library: host-buffer batch calls (acquire_batch, generate into host memory) shard over devices
    0 .. n_gpus-1 (0 = every visible device).  Returns the number of devices in use."""
    global _initialised
    check(lib().r4wb_init_devices(int(n_gpus)))
    _initialised = True
    return int(lib().r4wb_devices_initialised())


def ensure_init():
    if not _initialised:
        init(-1)


def set_stream(cuda_stream: int):
    check(lib().r4wb_set_stream(C.c_void_p(cuda_stream)))


class on_stream:
    """`with on_stream(s):` — the library's thread-local stream is `s` for the calls inside and the default stream again
    afterwards, so a later host-path call never runs on a stale (possibly destroyed) torch stream."""

    def __init__(self, cuda_stream: int):
        self._s = int(cuda_stream)

    def __enter__(self):
        set_stream(self._s)
        return self

    def __exit__(self, *exc):
        lib().r4wb_set_stream(C.c_void_p(0))
        return False


def kernel_launches() -> int:
    return int(lib().r4wb_kernel_launches())


def device_count() -> int:
    n = C.c_int(0)
    rc = lib().r4wb_device_count(C.byref(n))
    return int(n.value) if rc == 0 else 0


def version() -> str:
    return lib().r4wb_version().decode()
