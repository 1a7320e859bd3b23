"""GnssScenario — host-side mirror of r4w's multi-satellite IQ scenario generator on the B200 kernels.

Mirrors `GnssScenario` (crates/r4w-core/src/waveform/gnss/scenario.rs:51-705): same constructor input
(`GnssScenarioConfig`, the e1c_*.yaml schema), same method names and meaning:
`new/generate_block/generate/satellite_status/reset/is_done/progress/config/total_samples/block_size`
(scenario.rs:78,308,549,564,636,646,651,657,662,667).  Every sample is rendered by libr4w_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import _lib
from .config import (FLAG_CLOSED_FORM_PHASE, FLAG_NOISE_OFF, SIGNALS, GnssScenarioConfig, SatStatusPod, load_config)


@dataclass
class SatelliteStatus:
    """SatelliteStatus, gnss/satellite_emitter.rs:19-34."""
    prn: int
    signal: str
    elevation_deg: float
    azimuth_deg: float
    range_m: float
    range_rate_mps: float
    doppler_hz: float
    cn0_dbhz: float
    iono_delay_m: float
    tropo_delay_m: float
    antenna_gain_dbi: float
    clock_correction_s: float
    visible: bool


def _device_ptr(t):
    """(data_ptr, n_complex_samples, fmt) of a CUDA torch tensor holding complex samples."""
    import torch
    if not t.is_cuda or not t.is_contiguous():
        raise ValueError("device output must be a contiguous CUDA tensor")
    if t.dtype == torch.complex64:
        return t.data_ptr(), t.numel(), _lib.FMT_CF32
    if t.dtype == torch.complex128:
        return t.data_ptr(), t.numel(), _lib.FMT_CF64
    if t.dtype == torch.float32 and t.shape[-1] == 2:
        return t.data_ptr(), t.numel() // 2, _lib.FMT_CF32
    raise ValueError(f"unsupported tensor dtype {t.dtype}")


class GnssScenario:
    def __init__(self, config: GnssScenarioConfig, noise: bool = True, closed_form_phase: bool = False):
        """GnssScenario::new (scenario.rs:78-237).  `noise=False` drops the thermal-noise term (parity runs);
        `closed_form_phase=True` uses (i+1)*inc for the Doppler phase instead of emulating the reference's
        sequential f64 accumulation."""
        _lib.ensure_init()
        self._config = config
        flags = (0 if noise else FLAG_NOISE_OFF) | (FLAG_CLOSED_FORM_PHASE if closed_form_phase else 0)
        pod, self._keep = config.to_pod(flags=flags)
        h = C.c_void_p()
        _lib.check(_lib.lib().r4wb_scenario_create(C.byref(pod), C.byref(h)))
        self._h = h

    @classmethod
    def from_yaml(cls, path, elevation_mask_deg: Optional[float] = 5.0, **kw) -> "GnssScenario":
        """What `r4w gnss scenario --config <path>` builds (crates/r4w-cli/src/main.rs:4108-4133, 4442): the CLI
        overwrites receiver.elevation_mask_deg with its own default of 5.0."""
        return cls(load_config(path, cli_elevation_mask_deg=elevation_mask_deg), **kw)

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().r4wb_scenario_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- reference surface
    def config(self) -> GnssScenarioConfig:
        return self._config

    def total_samples(self) -> int:
        return int(_lib.lib().r4wb_scenario_total_samples(self._h))

    def block_size(self) -> int:
        return int(_lib.lib().r4wb_scenario_block_size(self._h))

    def current_sample(self) -> int:
        return int(_lib.lib().r4wb_scenario_current_sample(self._h))

    def is_done(self) -> bool:
        return bool(_lib.lib().r4wb_scenario_is_done(self._h))

    def progress(self) -> float:
        return float(_lib.lib().r4wb_scenario_progress(self._h))

    def reset(self):
        _lib.check(_lib.lib().r4wb_scenario_reset(self._h))

    def generate_block(self, block_size: int, dtype=np.complex64) -> np.ndarray:
        """generate_block (scenario.rs:308-546): the next min(block_size, remaining) samples as ONE reference
        block; an empty array when done.  dtype complex128 returns the reference's Vec<Complex64> layout
        (values are the f32-rendered samples widened)."""
        fmt = _lib.FMT_CF64 if np.dtype(dtype) == np.complex128 else _lib.FMT_CF32
        out = np.empty(int(block_size), np.complex128 if fmt == _lib.FMT_CF64 else np.complex64)
        written = C.c_uint64(0)
        _lib.check(_lib.lib().r4wb_scenario_generate_block(self._h, int(block_size), out.ctypes.data_as(C.c_void_p),
                                                           _lib.MEM_HOST, fmt, C.byref(written)))
        return out[: written.value]

    def generate_block_view(self, block_size: int, dtype=np.complex64) -> np.ndarray:
        """generate_block without the host copy: a read-only array over the library's pinned block (for canonical
        block sizes the render-ahead ring itself), valid until the next call on this scenario.  For sinks that
        only read the block (file / socket writers: main.rs:4488-4500)."""
        fmt = _lib.FMT_CF64 if np.dtype(dtype) == np.complex128 else _lib.FMT_CF32
        dt = np.complex128 if fmt == _lib.FMT_CF64 else np.complex64
        written, ptr = C.c_uint64(0), C.c_void_p(0)
        _lib.check(_lib.lib().r4wb_scenario_generate_block_view(self._h, int(block_size), fmt, C.byref(ptr), C.byref(written)))
        if written.value == 0:
            return np.empty(0, dt)
        buf = (C.c_char * (written.value * np.dtype(dt).itemsize)).from_address(ptr.value)
        out = np.frombuffer(buf, dtype=dt, count=written.value)
        out.flags.writeable = False
        return out

    def generate(self, dtype=np.complex64) -> np.ndarray:
        """generate (scenario.rs:549-561): `while !is_done {generate_block(block_size())}` — everything from
        current_sample to the end, rendered in one call; leaves the scenario done."""
        fmt = _lib.FMT_CF64 if np.dtype(dtype) == np.complex128 else _lib.FMT_CF32
        out = np.empty(self.total_samples() - self.current_sample(), np.complex128 if fmt == _lib.FMT_CF64 else np.complex64)
        written = C.c_uint64(0)
        _lib.check(_lib.lib().r4wb_scenario_generate_rest(self._h, out.ctypes.data_as(C.c_void_p), out.size, _lib.MEM_HOST, fmt,
                                                          C.byref(written)))
        return out[: written.value]

    def satellite_status(self) -> List[SatelliteStatus]:
        n = len(self._config.satellites)
        arr = (SatStatusPod * max(n, 1))()
        got = C.c_uint32(0)
        _lib.check(_lib.lib().r4wb_scenario_status(self._h, arr, n, C.byref(got)))
        return [SatelliteStatus(prn=a.prn, signal=SIGNALS[a.signal], elevation_deg=a.elevation_deg, azimuth_deg=a.azimuth_deg,
                                range_m=a.range_m, range_rate_mps=a.range_rate_mps, doppler_hz=a.doppler_hz,
                                cn0_dbhz=a.cn0_dbhz, iono_delay_m=a.iono_delay_m, tropo_delay_m=a.tropo_delay_m,
                                antenna_gain_dbi=a.antenna_gain_dbi, clock_correction_s=a.clock_correction_s,
                                visible=bool(a.visible)) for a in arr[: got.value]]

    # ---- random access (the throughput entry points)
    def generate_range(self, first: int, n: int, dtype=np.complex64, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Samples [first, first+n) of the canonical stream into host memory."""
        fmt = _lib.FMT_CF64 if np.dtype(dtype) == np.complex128 else _lib.FMT_CF32
        if out is None:
            out = np.empty(int(n), np.complex128 if fmt == _lib.FMT_CF64 else np.complex64)
        assert out.size >= n and out.flags.c_contiguous
        _lib.check(_lib.lib().r4wb_scenario_generate(self._h, int(first), int(n), out.ctypes.data_as(C.c_void_p),
                                                     _lib.MEM_HOST, fmt))
        return out[:n]

    def generate_range_format(self, first: int, n: int, fmt: str) -> np.ndarray:
        """Samples [first, first+n) in one of the CLI's integer sink formats ("ci16", "ci8", "cu8": IqFormat::write_sample,
        core/io/format.rs:203-222), converted in the kernel's store epilogue -> [n][2] integers (re, im)."""
        code, dt = {"ci16": (_lib.FMT_CI16, np.int16), "ci8": (_lib.FMT_CI8, np.int8), "cu8": (_lib.FMT_CU8, np.uint8)}[fmt]
        out = np.empty((int(n), 2), dt)
        _lib.check(_lib.lib().r4wb_scenario_generate(self._h, int(first), int(n), out.ctypes.data_as(C.c_void_p), _lib.MEM_HOST, code))
        return out

    def generate_range_into(self, first: int, n: int, host_ptr: int, fmt: int = _lib.FMT_CF32):
        """Same, into caller-owned host memory (e.g. pinned, r4wb_host_alloc)."""
        _lib.check(_lib.lib().r4wb_scenario_generate(self._h, int(first), int(n), C.c_void_p(host_ptr), _lib.MEM_HOST, fmt))

    def generate_device(self, first: int, n: int, out) -> None:
        """Samples [first, first+n) into a CUDA torch tensor (complex64 / complex128), on torch's current stream."""
        import torch
        ptr, cap, fmt = _device_ptr(out)
        if cap < n:
            raise ValueError("output tensor too small")
        with _lib.on_stream(torch.cuda.current_stream(out.device).cuda_stream):
            _lib.check(_lib.lib().r4wb_scenario_generate(self._h, int(first), int(n), C.c_void_p(ptr), _lib.MEM_DEVICE, fmt))

    def write_file(self, path, fmt: int = _lib.FMT_CF32):
        """The CLI's file sink (main.rs:4483-4509): the whole scenario streamed into `path` in `fmt` -> (samples written,
        bytes written, sum |s|^2 of the pre-conversion samples).  Leaves the scenario done."""
        n, b, p = C.c_uint64(0), C.c_uint64(0), C.c_double(0.0)
        _lib.check(_lib.lib().r4wb_scenario_write_file(self._h, str(path).encode(), int(fmt), C.byref(n), C.byref(b), C.byref(p)))
        return int(n.value), int(b.value), float(p.value)

    def last_power_sum(self) -> float:
        """Sum |s|^2 over the last generate call (the CLI's avg-power line, main.rs:4494-4509)."""
        v = C.c_double(0.0)
        _lib.check(_lib.lib().r4wb_scenario_last_power_sum(self._h, C.byref(v)))
        return float(v.value)

    def last_path(self) -> int:
        """0: k_synth / k_synth_lat rendered the last generate call, 1: the period-resident kernels did (diagnostic)."""
        return int(_lib.lib().r4wb_scenario_last_path(self._h))

    def set_profiling(self, enabled: bool = True):
        _lib.check(_lib.lib().r4wb_scenario_set_profiling(self._h, int(bool(enabled))))

    def last_profile(self):
        """{kernel: (summed CUDA-event ms, launches)} of the last generate call (profiling must be enabled)."""
        ms = np.zeros(4, np.float64)
        n = np.zeros(4, np.uint64)
        _lib.check(_lib.lib().r4wb_scenario_last_profile(self._h, ms.ctypes.data_as(C.c_void_p), n.ctypes.data_as(C.c_void_p)))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(("k_synth", "k_synth_periodic", "k_periodic_fix", "k_synth_lat"))}

    def _debug_block_params(self, block: int, sat: int) -> np.ndarray:
        out = np.zeros(12, np.float64)
        _lib.check(_lib.lib().r4wb_debug_block_params(self._h, int(block), int(sat), out.ctypes.data_as(C.c_void_p)))
        return out
