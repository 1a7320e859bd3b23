"""ctypes wrapper over tests/emu/libr4w_emu.so — host replay of the device arithmetic (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from r4w_b200.config import GnssScenarioConfig, ScenarioCfgPod

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libr4w_emu.so")
_lib = None


def build(force: bool = False) -> str:
    cmd = ["make", "-C", _HERE]
    if force:
        cmd.append("-B")
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, u64 = C.c_void_p, C.c_uint64
        L.emu_last_error.restype = C.c_char_p
        L.emu_scenario_create.argtypes = [C.POINTER(ScenarioCfgPod), C.POINTER(vp)]
        L.emu_scenario_destroy.argtypes = [vp]
        L.emu_scenario_total_samples.argtypes = [vp]; L.emu_scenario_total_samples.restype = u64
        L.emu_scenario_block_size.argtypes = [vp]; L.emu_scenario_block_size.restype = u64
        L.emu_scenario_segments.argtypes = [vp]; L.emu_scenario_segments.restype = C.c_uint32
        L.emu_scenario_set_lattice.argtypes = [vp, C.c_int]
        L.emu_scenario_patched.argtypes = [vp]; L.emu_scenario_patched.restype = u64
        L.emu_scenario_generate.argtypes = [vp, u64, u64, vp, C.c_int, C.POINTER(u64)]
        L.emu_scenario_generate_block.argtypes = [vp, u64, vp, C.POINTER(u64)]
        L.emu_block_params.argtypes = [vp, u64, C.c_uint32, vp]
        L.emu_phase_model_check.argtypes = [C.c_double, C.c_double, C.c_double, u64, vp]; L.emu_phase_model_check.restype = None
        L.emu_phase_q_steps_check.argtypes = [u64, C.c_uint32, C.c_double, C.c_int, C.c_int, C.c_uint32, C.c_uint32, vp]; L.emu_phase_q_steps_check.restype = None
        L.emu_fft.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
        L.emu_pcps.argtypes = [C.c_int, u64, C.c_double, C.c_double, C.c_double, vp, C.c_int, u64, vp, u64, vp, vp]
        _lib = L
    return _lib


class EmuScenario:
    def __init__(self, cfg: GnssScenarioConfig, noise: bool = True, closed_form_phase: bool = False):
        pod, self._keep = cfg.to_pod(flags=(0 if noise else 1) | (2 if closed_form_phase else 0))
        h = C.c_void_p()
        rc = lib().emu_scenario_create(C.byref(pod), C.byref(h))
        if rc:
            raise RuntimeError(f"emu create failed rc={rc}: {lib().emu_last_error().decode()}")
        self._h = h
        self.n_ambiguous = 0

    def __del__(self):
        if getattr(self, "_h", None):
            lib().emu_scenario_destroy(self._h)
            self._h = None

    def set_lattice(self, on: bool = True) -> bool:
        """replay k_synth_lat (synth_lattice.cuh) for full blocks; False when the scenario does not qualify"""
        return bool(lib().emu_scenario_set_lattice(self._h, int(on)))

    def patched(self) -> int:
        """windows-with-a-disputed-oversample records met so far by the lattice replay"""
        return int(lib().emu_scenario_patched(self._h))

    def total_samples(self): return int(lib().emu_scenario_total_samples(self._h))
    def block_size(self): return int(lib().emu_scenario_block_size(self._h))
    def segments(self): return int(lib().emu_scenario_segments(self._h))

    def generate_range(self, first: int, n: int, only_sat: int = -1) -> np.ndarray:
        out = np.zeros(n, np.complex64)
        amb = C.c_uint64(0)
        rc = lib().emu_scenario_generate(self._h, first, n, out.ctypes.data_as(C.c_void_p), only_sat, C.byref(amb))
        if rc:
            raise RuntimeError(f"emu generate failed rc={rc}: {lib().emu_last_error().decode()}")
        self.n_ambiguous = int(amb.value)
        return out

    def generate_block(self, n: int) -> np.ndarray:
        out = np.zeros(n, np.complex64)
        w = C.c_uint64(0)
        rc = lib().emu_scenario_generate_block(self._h, n, out.ctypes.data_as(C.c_void_p), C.byref(w))
        if rc:
            raise RuntimeError(f"emu generate_block failed rc={rc}: {lib().emu_last_error().decode()}")
        return out[: w.value]

    def block_params(self, block: int, sat: int) -> np.ndarray:
        out = np.zeros(12)
        rc = lib().emu_block_params(self._h, block, sat, out.ctypes.data_as(C.c_void_p))
        if rc:
            raise RuntimeError(f"emu block_params failed rc={rc}")
        return out


def fft(x: np.ndarray, log_m: int, inverse: bool, double: bool) -> np.ndarray:
    a = np.ascontiguousarray(x, np.complex128)
    n = a.size
    log_n = int(round(np.log2(n))) if n > 1 else 0
    out = np.zeros(n, np.complex128)
    rc = lib().emu_fft(log_n, log_m, 1 if inverse else -1, int(double), a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return out


def pcps(code_length: int, fs: float, dmax: float, dstep: float, x: np.ndarray, code: np.ndarray, double: bool = False,
         want_grid: bool = False):
    """returns (best, second, sum, lin[, grid])"""
    if x.dtype != np.complex64:
        x = np.ascontiguousarray(x, np.complex128)
    code = np.ascontiguousarray(code, np.int8)
    out = np.zeros(4)
    bins = int(2.0 * dmax / dstep) + 1
    grid = np.zeros((bins, code_length)) if want_grid else None
    lib().emu_pcps(int(double), code_length, fs, dmax, dstep, x.ctypes.data_as(C.c_void_p), int(x.dtype == np.complex128), x.size,
                   code.ctypes.data_as(C.c_void_p), code.size, out.ctypes.data_as(C.c_void_p),
                   grid.ctypes.data_as(C.c_void_p) if want_grid else None)
    res = (out[0], out[1], out[2], int(out[3]))
    return res + (grid,) if want_grid else res


def phase_q_steps_check(seed: int, cases: int, span_hz: float, k_lo: int, k_hi: int, n: int = 5000, max_steps: int = 1024):
    """-> (cases where the step-function form of the block's integer phase sum differs from the per-sample sum, cases it
    declined, largest number of levels, cases with an exact tie)"""
    out = np.zeros(4, np.float64)
    lib().emu_phase_q_steps_check(seed, cases, span_hz, k_lo, k_hi, n, max_steps, out.ctypes.data_as(C.c_void_p))
    return int(out[0]), int(out[1]), int(out[2]), int(out[3])


def phase_model_check(d0: float, rate: float, jerk: float, blocks: int):
    """-> (blocks where the integer-sum phase model differs from the literal accumulation, blocks walked sample by sample,
    final phase, final phase - real-number sum)"""
    out = np.zeros(4, np.float64)
    lib().emu_phase_model_check(d0, rate, jerk, blocks, out.ctypes.data_as(C.c_void_p))
    return int(out[0]), int(out[1]), float(out[2]), float(out[3])
