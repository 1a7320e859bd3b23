"""Host-side mirror of the reference's tracking channel (crates/r4w-core/src/waveform/gnss/tracking.rs) over the C-ABI.

`TrackingChannel` keeps the reference surface — `new(prn, code_length, sample_rate, chipping_rate, initial_code_phase,
initial_doppler)`, `with_dll_bandwidth`, `with_pll_bandwidth`, `process(samples, code) -> TrackingState`, `nav_bits()`,
`state()` — and `TrackerBank` is the throughput form: many independent channels, many code periods per call, one CTA
per channel on the GPU (r4wb_track_process).  No CPU fallback."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Sequence

import numpy as np

from . import _lib


class TrackCfgPod(C.Structure):
    _fields_ = [("sample_rate", C.c_double), ("chipping_rate", C.c_double), ("initial_code_phase", C.c_double),
                ("initial_doppler", C.c_double), ("dll_bandwidth_hz", C.c_double), ("pll_bandwidth_hz", C.c_double),
                ("code_length", C.c_uint64), ("prn", C.c_uint8), ("pad", C.c_uint8 * 7)]


class TrackStatePod(C.Structure):
    _fields_ = [("code_phase", C.c_double), ("carrier_freq_hz", C.c_double), ("carrier_phase_rad", C.c_double),
                ("prompt_i", C.c_double), ("prompt_q", C.c_double), ("cn0_dbhz", C.c_double), ("ms_count", C.c_uint64),
                ("prn", C.c_uint8), ("carrier_lock", C.c_uint8), ("code_lock", C.c_uint8), ("bit_sync", C.c_uint8),
                ("pad", C.c_uint8 * 4)]


TRACK_STATE_DTYPE = np.dtype([("code_phase", "<f8"), ("carrier_freq_hz", "<f8"), ("carrier_phase_rad", "<f8"), ("prompt_i", "<f8"),
                              ("prompt_q", "<f8"), ("cn0_dbhz", "<f8"), ("ms_count", "<u8"), ("prn", "u1"), ("carrier_lock", "u1"),
                              ("code_lock", "u1"), ("bit_sync", "u1"), ("pad", "u1", (4,))])
assert TRACK_STATE_DTYPE.itemsize == C.sizeof(TrackStatePod) == 64


@dataclass
class TrackingState:                    # gnss/types.rs:187-210
    prn: int
    code_phase: float
    carrier_freq_hz: float
    carrier_phase_rad: float
    prompt_i: float
    prompt_q: float
    cn0_dbhz: float
    carrier_lock: bool
    code_lock: bool
    bit_sync: bool
    ms_count: int


def _state(r) -> TrackingState:
    return TrackingState(int(r["prn"]), float(r["code_phase"]), float(r["carrier_freq_hz"]), float(r["carrier_phase_rad"]),
                         float(r["prompt_i"]), float(r["prompt_q"]), float(r["cn0_dbhz"]), bool(r["carrier_lock"]),
                         bool(r["code_lock"]), bool(r["bit_sync"]), int(r["ms_count"]))


class TrackerBank:
    """n independent channels; `channels` is a list of dicts with the TrackingChannel::new arguments
    (prn, code_length, sample_rate, chipping_rate, initial_code_phase, initial_doppler[, dll_bandwidth_hz, pll_bandwidth_hz])."""

    def __init__(self, channels: Sequence[dict]):
        n = len(channels)
        pods = (TrackCfgPod * n)()
        for p, c in zip(pods, channels):
            p.prn = int(c["prn"]); p.code_length = int(c["code_length"]); p.sample_rate = float(c["sample_rate"])
            p.chipping_rate = float(c["chipping_rate"]); p.initial_code_phase = float(c["initial_code_phase"])
            p.initial_doppler = float(c["initial_doppler"])
            p.dll_bandwidth_hz = float(c.get("dll_bandwidth_hz", 0.0)); p.pll_bandwidth_hz = float(c.get("pll_bandwidth_hz", 0.0))
        self._n = n
        self._code_length = max(int(c["code_length"]) for c in channels)
        self._h = C.c_void_p()
        _lib.check(_lib.lib().r4wb_track_create(pods, n, C.byref(self._h)))

    def __del__(self):
        try:
            h = getattr(self, "_h", None)
            if h:
                _lib.lib().r4wb_track_destroy(h)
                self._h = None
        except Exception:       # interpreter shutdown
            pass

    def channels(self) -> int:
        return self._n

    def process(self, samples, codes: np.ndarray, n_per_period: int, n_periods: int, per_channel_input: bool = False) -> np.ndarray:
        """n_periods consecutive `process` calls on every channel -> structured array [n_periods][n_channels] of TrackingState.
        samples: complex64 / complex128 numpy array, or a CUDA complex64 torch tensor; [n_periods * n_per_period] shared by all
        channels, or [n_channels][n_periods * n_per_period] with per_channel_input.  codes: int8 [n_channels][code_length]."""
        codes = np.ascontiguousarray(codes, np.int8).reshape(self._n, -1)
        assert codes.shape[1] >= self._code_length
        span = int(n_per_period) * int(n_periods)
        stream = 0
        if isinstance(samples, np.ndarray):
            fmt = _lib.FMT_CF64 if samples.dtype == np.complex128 else _lib.FMT_CF32
            x = np.ascontiguousarray(samples, np.complex128 if fmt == _lib.FMT_CF64 else np.complex64)
            assert x.size >= (span * self._n if per_channel_input else span)
            ptr, where = x.ctypes.data_as(C.c_void_p), _lib.MEM_HOST
        else:
            import torch
            assert samples.is_cuda and samples.dtype == torch.complex64 and samples.is_contiguous()
            assert samples.numel() >= (span * self._n if per_channel_input else span)
            stream = torch.cuda.current_stream(samples.device).cuda_stream
            ptr, where, fmt = C.c_void_p(samples.data_ptr()), _lib.MEM_DEVICE, _lib.FMT_CF32
        out = np.zeros((int(n_periods), self._n), TRACK_STATE_DTYPE)
        with _lib.on_stream(stream):
            _lib.check(_lib.lib().r4wb_track_process(self._h, ptr, fmt, where, int(n_per_period), int(n_periods),
                                                     span if per_channel_input else 0, codes.ctypes.data_as(C.c_void_p), codes.shape[1],
                                                     out.ctypes.data_as(C.c_void_p)))
        return out

    def state(self) -> List[TrackingState]:
        out = np.zeros(self._n, TRACK_STATE_DTYPE)
        _lib.check(_lib.lib().r4wb_track_state_get(self._h, out.ctypes.data_as(C.c_void_p), self._n))
        return [_state(r) for r in out]

    def nav_bits(self, channel: int = 0) -> np.ndarray:
        n = C.c_uint64(0)
        _lib.check(_lib.lib().r4wb_track_nav_bits(self._h, int(channel), None, 0, C.byref(n)))
        out = np.zeros(int(n.value), np.int8)
        if out.size:
            _lib.check(_lib.lib().r4wb_track_nav_bits(self._h, int(channel), out.ctypes.data_as(C.c_void_p), out.size, C.byref(n)))
        return out


class TrackingChannel:
    """Single channel with the reference's call shapes (tracking.rs:107-358)."""

    def __init__(self, prn: int, code_length: int, sample_rate: float, chipping_rate: float, initial_code_phase: float,
                 initial_doppler: float):
        self._cfg = dict(prn=prn, code_length=code_length, sample_rate=sample_rate, chipping_rate=chipping_rate,
                         initial_code_phase=initial_code_phase, initial_doppler=initial_doppler)
        self._bank = None

    def with_dll_bandwidth(self, bw_hz: float) -> "TrackingChannel":
        self._cfg["dll_bandwidth_hz"] = float(bw_hz); self._bank = None
        return self

    def with_pll_bandwidth(self, bw_hz: float) -> "TrackingChannel":
        self._cfg["pll_bandwidth_hz"] = float(bw_hz); self._bank = None
        return self

    def _b(self) -> TrackerBank:
        if self._bank is None:
            self._bank = TrackerBank([self._cfg])
        return self._bank

    def process(self, samples: np.ndarray, code: np.ndarray) -> TrackingState:
        """one code period of samples -> updated state (tracking.rs:177-313)"""
        return _state(self._b().process(samples, np.asarray(code, np.int8)[None, :], len(samples), 1)[0, 0])

    def nav_bits(self) -> np.ndarray:
        return self._b().nav_bits(0)

    def state(self) -> TrackingState:
        return self._b().state()[0]
