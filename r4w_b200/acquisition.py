"""PcpsAcquisition — host-side mirror of r4w's FFT-based parallel-code-phase search on the B200 kernels.

Mirrors `PcpsAcquisition` (crates/r4w-core/src/waveform/gnss/acquisition.rs:40-255): builder methods
`new / with_doppler_range / with_threshold / with_coherent_periods` (:63-93), `acquire` (:104-195),
`acquire_grid` (:199-249), `fft_size` (:252), and `AcquisitionResult` (gnss/types.rs:168-183) /
`AcquisitionGrid` (acquisition.rs:259-286).  `acquire_batch` is the throughput entry point: many snapshots x
many local replicas in one call.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from .config import AcqResultPod


@dataclass
class AcquisitionResult:
    prn: int
    detected: bool
    code_phase: float          # integer lag in samples, as f64 (acquisition.rs:189)
    doppler_hz: float
    peak_metric: float
    threshold: float
    cn0_estimate: Optional[float]


@dataclass
class AcquisitionGrid:
    """AcquisitionGrid, acquisition.rs:259-286."""
    power: np.ndarray          # [doppler_bins][code_phases]
    doppler_bins: np.ndarray
    code_phases: int

    def find_peak(self):
        """(doppler_hz, code_phase, power) of the first maximum in scan order (acquisition.rs:270-285)."""
        lin = int(np.argmax(self.power))           # first occurrence in row-major order == the reference's strict `>`
        d, p = divmod(lin, self.power.shape[1])
        best = float(self.power[d, p])
        if not best > 0.0:
            return 0.0, 0.0, 0.0
        return float(self.doppler_bins[d]), float(p), best


def _result(p: AcqResultPod) -> AcquisitionResult:
    return AcquisitionResult(prn=int(p.prn), detected=bool(p.detected), code_phase=float(p.code_phase),
                             doppler_hz=float(p.doppler_hz), peak_metric=float(p.peak_metric), threshold=float(p.threshold),
                             cn0_estimate=float(p.cn0_estimate) if p.has_cn0 else None)


def _as_samples(x):
    """host array -> (ndarray, fmt); complex64 stays cf32, everything else becomes Complex64 (cf64)."""
    a = np.asarray(x)
    if a.dtype == np.complex64:
        return np.ascontiguousarray(a), _lib.FMT_CF32
    return np.ascontiguousarray(a, np.complex128), _lib.FMT_CF64


class PcpsAcquisition:
    def __init__(self, code_length: int, sample_rate: float):
        """PcpsAcquisition::new (acquisition.rs:63-74): defaults +-5000 Hz / 500 Hz, threshold 2.5."""
        _lib.ensure_init()
        h = C.c_void_p()
        _lib.check(_lib.lib().r4wb_pcps_create(int(code_length), float(sample_rate), C.byref(h)))
        self._h = h
        self.code_length = int(code_length)
        self.sample_rate = float(sample_rate)
        self.doppler_max_hz, self.doppler_step_hz, self.threshold, self.coherent_periods = 5000.0, 500.0, 2.5, 1

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().r4wb_pcps_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- builders (by value in the reference; here they mutate and return self)
    def with_doppler_range(self, max_hz: float, step_hz: float) -> "PcpsAcquisition":
        _lib.check(_lib.lib().r4wb_pcps_set_doppler_range(self._h, float(max_hz), float(step_hz)))
        self.doppler_max_hz, self.doppler_step_hz = float(max_hz), float(step_hz)
        return self

    def with_threshold(self, threshold: float) -> "PcpsAcquisition":
        _lib.check(_lib.lib().r4wb_pcps_set_threshold(self._h, float(threshold)))
        self.threshold = float(threshold)
        return self

    def with_coherent_periods(self, periods: int) -> "PcpsAcquisition":
        _lib.check(_lib.lib().r4wb_pcps_set_coherent_periods(self._h, max(int(periods), 1)))   # periods.max(1), :91
        self.coherent_periods = max(int(periods), 1)
        return self

    def fft_size(self) -> int:
        return int(_lib.lib().r4wb_pcps_fft_size(self._h))

    def num_doppler_bins(self) -> int:
        return int(_lib.lib().r4wb_pcps_num_doppler_bins(self._h))

    def guard_count(self) -> int:
        """(snapshot, code) pairs of the last batch that were re-run in f64 (near-tie / near-threshold guard)."""
        return int(_lib.lib().r4wb_pcps_guard_count(self._h))

    def set_profiling(self, enabled: bool = True):
        _lib.check(_lib.lib().r4wb_pcps_set_profiling(self._h, int(bool(enabled))))

    def last_profile(self):
        """{kernel kind: (summed CUDA-event ms, launches)} of the last acquire_batch (profiling must be enabled)."""
        ms = np.zeros(4, np.float64)
        n = np.zeros(4, np.uint64)
        _lib.check(_lib.lib().r4wb_pcps_last_profile(self._h, ms.ctypes.data_as(C.c_void_p), n.ctypes.data_as(C.c_void_p)))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(("code_spectra", "forward_fft", "inverse_fft_peak", "pair_reduce"))}

    # ---- searches
    def acquire(self, input_, code, prn: int) -> AcquisitionResult:
        x, fmt = _as_samples(input_)
        c = np.ascontiguousarray(code, np.int8)
        out = AcqResultPod()
        _lib.check(_lib.lib().r4wb_pcps_acquire(self._h, x.ctypes.data_as(C.c_void_p), fmt, x.size,
                                                c.ctypes.data_as(C.c_void_p), c.size, int(prn), C.byref(out)))
        return _result(out)

    def acquire_batch(self, input_, n_snapshots: int, snapshot_stride: int, n_input: int, codes, prns: Sequence[int]
                      ) -> List[List[AcquisitionResult]]:
        """Snapshot s = input[s*stride : s*stride + n_input]; codes is [n_codes][code_len] int8.
        `input_` is a host array or a CUDA torch tensor (complex64/complex128).  Returns [n_snapshots][n_codes]."""
        pods = self.acquire_batch_raw(input_, n_snapshots, snapshot_stride, n_input, codes, prns)
        P = len(prns)
        return [[_result(pods[s * P + c]) for c in range(P)] for s in range(n_snapshots)]

    def acquire_batch_raw(self, input_, n_snapshots: int, snapshot_stride: int, n_input: int, codes, prns: Sequence[int]):
        stream = None
        cs = np.ascontiguousarray(codes, np.int8)
        if cs.ndim != 2 or cs.shape[0] != len(prns):
            raise ValueError("codes must be [n_codes][code_len]")
        pr = np.ascontiguousarray(prns, np.uint8)
        n_snapshots, snapshot_stride, n_input = int(n_snapshots), int(snapshot_stride), int(n_input)
        need = (n_snapshots - 1) * snapshot_stride + n_input if n_snapshots > 0 else 0
        if hasattr(input_, "is_cuda") and input_.is_cuda:
            import torch
            from .scenario import _device_ptr
            ptr, cap, fmt = _device_ptr(input_)
            where = _lib.MEM_DEVICE
            stream = torch.cuda.current_stream(input_.device).cuda_stream
        else:
            x, fmt = _as_samples(input_)
            ptr, cap, where = x.ctypes.data, x.size, _lib.MEM_HOST
        if cap < need:
            raise ValueError(f"input holds {cap} samples, batch needs {need}")
        out = (AcqResultPod * max(n_snapshots * len(prns), 1))()
        args = (self._h, C.c_void_p(ptr), fmt, where, n_snapshots, snapshot_stride, n_input, cs.ctypes.data_as(C.c_void_p), cs.shape[1],
                pr.ctypes.data_as(C.c_void_p), len(prns), out)
        if stream is None:
            _lib.check(_lib.lib().r4wb_pcps_acquire_batch(*args))
        else:
            with _lib.on_stream(stream):          # device input: run on torch's current stream, then back to the default
                _lib.check(_lib.lib().r4wb_pcps_acquire_batch(*args))
        return out

    def acquire_grid(self, input_, code) -> AcquisitionGrid:
        x, fmt = _as_samples(input_)
        c = np.ascontiguousarray(code, np.int8)
        bins = self.num_doppler_bins()
        power = np.zeros((bins, self.code_length), np.float64)
        _lib.check(_lib.lib().r4wb_pcps_acquire_grid(self._h, x.ctypes.data_as(C.c_void_p), fmt, x.size,
                                                     c.ctypes.data_as(C.c_void_p), c.size,
                                                     power.ctypes.data_as(C.c_void_p), power.size))
        dop = -self.doppler_max_hz + np.arange(bins, dtype=np.float64) * self.doppler_step_hz
        return AcquisitionGrid(power=power, doppler_bins=dop, code_phases=self.code_length)


# ---- codes (GalileoE1CodeGenerator, gnss/prn.rs:268-327)
def e1_code(channel: int, prn: int) -> np.ndarray:
    """4092 chips (+1/-1) of Galileo E1B (channel 0) or E1C (channel 1), PRN 1-50.  Host-side table unpack."""
    out = np.zeros(4092, np.int8)
    _lib.check(_lib.lib().r4wb_e1_code(int(channel), int(prn), out.ctypes.data_as(C.c_void_p), out.size))
    return out


def gps_ca_code(prn: int) -> np.ndarray:
    """1023 chips (+1/-1) of the GPS L1 C/A Gold code, PRN 1-32 (GpsCaCodeGenerator, gnss/prn.rs:34-162)."""
    out = np.zeros(1023, np.int8)
    _lib.check(_lib.lib().r4wb_gps_ca_code(int(prn), out.ctypes.data_as(C.c_void_p), out.size))
    return out


def e1c_secondary() -> np.ndarray:
    out = np.zeros(25, np.int8)
    _lib.check(_lib.lib().r4wb_e1c_secondary(out.ctypes.data_as(C.c_void_p), out.size))
    return out


def e1c_replica(prn: int, sample_rate: float, n: int) -> np.ndarray:
    """Sampled E1C x BOC(1,1) local replica (no secondary code): the `code` argument of acquire() for an E1C scenario."""
    out = np.zeros(int(n), np.int8)
    _lib.check(_lib.lib().r4wb_e1c_replica(int(prn), float(sample_rate), out.ctypes.data_as(C.c_void_p), out.size))
    return out


def gps_l5_code(prn: int) -> np.ndarray:
    """10230 chips (+1/-1) of the GPS L5 I5 code as the reference generates it (gnss/prn.rs:345-397), PRN 1-32"""
    out = np.zeros(10230, np.int8)
    _lib.check(_lib.lib().r4wb_gps_l5_code(int(prn), out.ctypes.data_as(C.c_void_p), out.size))
    return out


def glonass_code() -> np.ndarray:
    """the 511-chip GLONASS L1OF ranging code (gnss/prn.rs:170-216)"""
    out = np.zeros(511, np.int8)
    _lib.check(_lib.lib().r4wb_glonass_code(out.ctypes.data_as(C.c_void_p), out.size))
    return out
