"""ctypes wrapper over oracle/libr4w_oracle.so — the CPU restatement of the reference path.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under r4w_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from r4w_b200.config import (AcqResultPod, GnssScenarioConfig, SatStatusPod, ScenarioCfgPod)

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libr4w_oracle.so")


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, a second or two)."""
    src = os.path.join(_HERE, "r4w_oracle.c")
    stale = (not os.path.exists(_SO)) or os.path.getmtime(_SO) < max(
        os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "r4w_oracle_track.c")),
        os.path.getmtime(os.path.join(_HERE, "r4w_oracle_sim.c")),
        os.path.getmtime(os.path.join(_HERE, "r4w_oracle.h")))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libr4w_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


class BlockParamsPod(C.Structure):
    _fields_ = [("visible", C.c_int32), ("_pad", C.c_int32), ("range_m", C.c_double), ("iono_delay_s", C.c_double),
                ("tropo_delay_s", C.c_double), ("rx_amplitude", C.c_double), ("doppler_start_hz", C.c_double),
                ("doppler_end_hz", C.c_double), ("initial_code_phase", C.c_double),
                ("initial_epoch_offset", C.c_uint64), ("phase_before", C.c_double)]


class PcpsPod(C.Structure):
    _fields_ = [("fft_size", C.c_uint64), ("code_length", C.c_uint64), ("doppler_max_hz", C.c_double),
                ("doppler_step_hz", C.c_double), ("threshold", C.c_double), ("sample_rate", C.c_double),
                ("coherent_periods", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        vp, u64, sz, dbl, i32 = C.c_void_p, C.c_uint64, C.c_size_t, C.c_double, C.c_int
        L.orc_e1_code.argtypes = [i32, i32, vp]; L.orc_e1_code.restype = i32
        L.orc_e1c_secondary.argtypes = [vp]
        L.orc_gps_ca_code.argtypes = [i32, vp]; L.orc_gps_ca_code.restype = i32
        L.orc_glonass_code.argtypes = [i32, vp]; L.orc_glonass_code.restype = i32
        L.orc_gps_l5_code.argtypes = [i32, i32, vp]; L.orc_gps_l5_code.restype = i32
        L.orc_e1c_replica.argtypes = [i32, dbl, vp, sz]
        L.orc_blackman_window.argtypes = [sz, vp]
        L.orc_lowpass_taps.argtypes = [dbl, dbl, sz, vp, sz]; L.orc_lowpass_taps.restype = sz
        L.orc_lla_to_ecef.argtypes = [vp, vp]
        L.orc_look_angle.argtypes = [vp, vp, vp, vp, vp, vp]
        L.orc_range_rate.argtypes = [vp, vp, vp, vp]; L.orc_range_rate.restype = dbl
        L.orc_fspl_db.argtypes = [dbl, dbl]; L.orc_fspl_db.restype = dbl
        L.orc_galileo_position_velocity.argtypes = [i32, i32, dbl, vp, vp]
        L.orc_gps_position_velocity.argtypes = [i32, i32, dbl, vp, vp]
        L.orc_kepler_period.argtypes = [dbl]; L.orc_kepler_period.restype = dbl
        L.orc_solve_kepler.argtypes = [dbl, dbl]; L.orc_solve_kepler.restype = dbl
        L.orc_antenna_gain_dbi.argtypes = [C.c_uint32, dbl, dbl, dbl]; L.orc_antenna_gain_dbi.restype = dbl
        L.orc_emitter_baseband.argtypes = [C.c_uint32, i32, i32, sz, dbl, dbl, dbl, dbl, sz, vp]
        L.orc_emitter_baseband.restype = i32
        L.orc_scenario_new.argtypes = [C.POINTER(ScenarioCfgPod), C.POINTER(vp)]; L.orc_scenario_new.restype = i32
        L.orc_scenario_free.argtypes = [vp]
        for name in ("total_samples", "block_size", "current_sample"):
            f = getattr(L, "orc_scenario_" + name); f.argtypes = [vp]; f.restype = u64
        L.orc_scenario_is_done.argtypes = [vp]; L.orc_scenario_is_done.restype = i32
        L.orc_scenario_reset.argtypes = [vp]
        L.orc_scenario_generate_block.argtypes = [vp, sz, vp]; L.orc_scenario_generate_block.restype = sz
        L.orc_scenario_set_threads.argtypes = [vp, i32]
        L.orc_scenario_skip_to.argtypes = [vp, u64]; L.orc_scenario_skip_to.restype = i32
        L.orc_scenario_peek_params.argtypes = [vp, sz, vp, sz]; L.orc_scenario_peek_params.restype = i32
        L.orc_scenario_status.argtypes = [vp, vp, sz]; L.orc_scenario_status.restype = i32
        L.orc_scenario_noise_std.argtypes = [vp]; L.orc_scenario_noise_std.restype = dbl
        L.orc_fft.argtypes = [vp, sz, i32]
        L.orc_pcps_init.argtypes = [C.POINTER(PcpsPod), u64, dbl]
        L.orc_pcps_num_bins.argtypes = [C.POINTER(PcpsPod)]; L.orc_pcps_num_bins.restype = i32
        L.orc_pcps_acquire.argtypes = [C.POINTER(PcpsPod), vp, sz, vp, sz, C.c_uint8, C.POINTER(AcqResultPod)]
        L.orc_pcps_acquire_grid.argtypes = [C.POINTER(PcpsPod), vp, sz, vp, sz, vp]
        L.orc_pcps_acquire_grid.restype = C.c_int64
        L.orc_to_cf32.argtypes = [vp, sz, vp]
        d = C.c_double
        L.orc_klobuchar_delay_s.argtypes = [vp, vp, d, d, d, d, d]
        L.orc_klobuchar_delay_s.restype = d
        L.orc_saastamoinen_delay_m.argtypes = [d, d, d, d, d]
        L.orc_saastamoinen_delay_m.restype = d
        L.orc_saastamoinen_zenith_m.argtypes = [d, d, d, d, C.c_int]
        L.orc_saastamoinen_zenith_m.restype = d
        L.orc_to_int_format.argtypes = [vp, sz, C.c_int, vp]
        L.orc_to_int_format.restype = C.c_int
        L.orc_sim_compose_emitter.argtypes = [vp, sz, d, d, d, vp, vp]
        L.orc_sim_link.argtypes = [vp, vp, vp, vp, d, d, vp]
        L.orc_sim_noise_power.argtypes = [d, d]; L.orc_sim_noise_power.restype = d
        L.orc_sim_trajectory.argtypes = [i32, vp, d, vp]
        L.orc_sim_waypoints.argtypes = [vp, sz, d, vp]
        L.orc_track_new.argtypes = [C.c_uint8, sz, d, d, d, d]; L.orc_track_new.restype = vp
        L.orc_track_free.argtypes = [vp]
        L.orc_track_set_dll_bandwidth.argtypes = [vp, d]
        L.orc_track_set_pll_bandwidth.argtypes = [vp, d]
        L.orc_track_process.argtypes = [vp, vp, sz, vp, vp]
        L.orc_track_state_get.argtypes = [vp, vp]
        L.orc_track_nav_bits.argtypes = [vp, vp, sz]; L.orc_track_nav_bits.restype = sz
        L.orc_loop_filter_2nd_run.argtypes = [d, d, d, sz]; L.orc_loop_filter_2nd_run.restype = d
        L.orc_loop_filter_3rd_run.argtypes = [d, d, d, sz]; L.orc_loop_filter_3rd_run.restype = d
        L.orc_dll_s_curve.argtypes = [d, sz, vp, vp]
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------ codes / filters / geometry
def e1_code(channel: int, prn: int) -> np.ndarray:
    out = np.zeros(4092, np.int8)
    if lib().orc_e1_code(channel, prn, _ptr(out)):
        raise ValueError("bad Galileo PRN/channel")
    return out


def e1c_secondary() -> np.ndarray:
    out = np.zeros(25, np.int8)
    lib().orc_e1c_secondary(_ptr(out))
    return out


def gps_ca_code(prn: int) -> np.ndarray:
    out = np.zeros(1023, np.int8)
    if lib().orc_gps_ca_code(prn, _ptr(out)):
        raise ValueError("bad GPS PRN")
    return out


def glonass_code(frequency_channel: int = 0) -> np.ndarray:
    out = np.zeros(511, np.int8)
    if lib().orc_glonass_code(int(frequency_channel), _ptr(out)):
        raise ValueError("GLONASS frequency channel must be -7 to +6")
    return out


def gps_l5_code(prn: int, q_channel: bool = False) -> np.ndarray:
    out = np.zeros(10230, np.int8)
    if lib().orc_gps_l5_code(int(prn), int(bool(q_channel)), _ptr(out)):
        raise ValueError("GPS L5 PRN must be 1-32")
    return out


def e1c_replica(prn: int, sample_rate: float, n: int) -> np.ndarray:
    out = np.zeros(n, np.int8)
    lib().orc_e1c_replica(prn, sample_rate, _ptr(out), n)
    return out


def lowpass_taps(cutoff_hz: float, sample_rate: float, num_taps: int) -> np.ndarray:
    out = np.zeros(num_taps + 1, np.float64)
    n = lib().orc_lowpass_taps(cutoff_hz, sample_rate, num_taps, _ptr(out), out.size)
    return out[:n].copy()


def blackman_window(n: int) -> np.ndarray:
    out = np.zeros(n, np.float64)
    lib().orc_blackman_window(n, _ptr(out))
    return out


def lla_to_ecef(lat, lon, alt) -> np.ndarray:
    lla = np.array([lat, lon, alt], np.float64); out = np.zeros(3)
    lib().orc_lla_to_ecef(_ptr(lla), _ptr(out))
    return out


def look_angle(obs_ecef, obs_lla, tgt_ecef):
    o = np.asarray(obs_ecef, np.float64); l = np.asarray(obs_lla, np.float64); t = np.asarray(tgt_ecef, np.float64)
    el, az, rg = C.c_double(), C.c_double(), C.c_double()
    lib().orc_look_angle(_ptr(o), _ptr(l), _ptr(t), C.addressof(el), C.addressof(az), C.addressof(rg))
    return el.value, az.value, rg.value


def range_rate(op, ov, tp, tv) -> float:
    a = [np.asarray(x, np.float64) for x in (op, ov, tp, tv)]
    return lib().orc_range_rate(*[_ptr(x) for x in a])


def galileo_position_velocity(plane, slot, t):
    p, v = np.zeros(3), np.zeros(3)
    lib().orc_galileo_position_velocity(plane, slot, t, _ptr(p), _ptr(v))
    return p, v


def gps_position_velocity(plane, slot, t):
    p, v = np.zeros(3), np.zeros(3)
    lib().orc_gps_position_velocity(plane, slot, t, _ptr(p), _ptr(v))
    return p, v


def emitter_baseband(signal: int, prn: int, nav_data: bool, n: int, sample_rate: float, range_m: float,
                     iono_s: float = 0.0, tropo_s: float = 0.0, sample_offset: int = 0) -> np.ndarray:
    out = np.zeros(n, np.float64)
    rc = lib().orc_emitter_baseband(signal, prn, int(nav_data), n, sample_rate, range_m, iono_s, tropo_s,
                                    sample_offset, _ptr(out))
    if rc:
        raise ValueError(f"emitter init failed ({rc})")
    return out


def fft(x: np.ndarray, inverse: bool = False) -> np.ndarray:
    a = np.ascontiguousarray(x, np.complex128).copy()
    lib().orc_fft(_ptr(a), a.size, int(inverse))
    return a


# ------------------------------------------------------------------ scenario
class OracleScenario:
    """GnssScenario restated (gnss/scenario.rs:51-705)."""

    def __init__(self, cfg: GnssScenarioConfig, noise: bool = True, threads: int = 1):
        self.cfg = cfg
        pod, self._keep = cfg.to_pod(flags=0 if noise else 1)
        h = C.c_void_p()
        rc = lib().orc_scenario_new(C.byref(pod), C.byref(h))
        if rc:
            raise ValueError(f"oracle cannot build this scenario (rc={rc})")
        self._h = h
        if threads > 1:
            lib().orc_scenario_set_threads(h, threads)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().orc_scenario_free(self._h)
                self._h = None
        except Exception:      # interpreter teardown
            pass

    def total_samples(self) -> int: return int(lib().orc_scenario_total_samples(self._h))
    def block_size(self) -> int: return int(lib().orc_scenario_block_size(self._h))
    def current_sample(self) -> int: return int(lib().orc_scenario_current_sample(self._h))
    def is_done(self) -> bool: return bool(lib().orc_scenario_is_done(self._h))
    def reset(self): lib().orc_scenario_reset(self._h)
    def noise_std(self) -> float: return float(lib().orc_scenario_noise_std(self._h))

    def generate_block(self, block_size: int) -> np.ndarray:
        out = np.zeros(block_size, np.complex128)
        n = lib().orc_scenario_generate_block(self._h, block_size, _ptr(out))
        return out[:n]

    def skip_to(self, sample: int):
        rc = lib().orc_scenario_skip_to(self._h, sample)
        if rc:
            raise ValueError(f"skip_to({sample}) failed rc={rc}")

    def generate_range(self, first: int, n: int) -> np.ndarray:
        """Samples [first, first+n) of the canonical `while !is_done {generate_block(block_size())}` stream."""
        bs = self.block_size()
        start_block = (first // bs) * bs
        if start_block < self.current_sample():
            self.reset()
        self.skip_to(start_block)
        chunks, have = [], 0
        need = first - start_block + n
        while have < need and not self.is_done():
            b = self.generate_block(bs)
            chunks.append(b); have += b.size
        allv = np.concatenate(chunks) if chunks else np.zeros(0, np.complex128)
        return allv[first - start_block: first - start_block + n]

    def peek_params(self, block_size=None):
        n = len(self.cfg.satellites)
        arr = (BlockParamsPod * max(n, 1))()
        lib().orc_scenario_peek_params(self._h, block_size or self.block_size(), arr, n)
        return [arr[k] for k in range(n)]

    def status(self):
        n = len(self.cfg.satellites)
        arr = (SatStatusPod * max(n, 1))()
        lib().orc_scenario_status(self._h, arr, n)
        return [arr[k] for k in range(n)]


KLOBUCHAR_DEFAULT = ([0.1118e-7, 0.7451e-8, -0.5961e-7, -0.1192e-6], [0.1167e6, -0.4267e5, -0.2621e6, 0.1311e6])


def klobuchar_delay_s(el_rad, az_rad, lat_rad, lon_rad, gps_t, alpha=None, beta=None) -> float:
    """KlobucharModel::delay_seconds (environment/ionosphere.rs:46-108), default_broadcast coefficients unless given"""
    a = np.asarray(alpha if alpha is not None else KLOBUCHAR_DEFAULT[0], np.float64)
    b = np.asarray(beta if beta is not None else KLOBUCHAR_DEFAULT[1], np.float64)
    return float(lib().orc_klobuchar_delay_s(_ptr(a), _ptr(b), el_rad, az_rad, lat_rad, lon_rad, gps_t))


def saastamoinen(height_m=0.0, temperature_k=288.15, pressure_hpa=1013.25, relative_humidity=0.5):
    """SaastamoinenModel (environment/troposphere.rs): returns (zenith dry m, zenith wet m, slant-delay function of el_rad)"""
    L = lib()
    args = (height_m, temperature_k, pressure_hpa, relative_humidity)
    return (float(L.orc_saastamoinen_zenith_m(*args, 0)), float(L.orc_saastamoinen_zenith_m(*args, 1)),
            lambda el: float(L.orc_saastamoinen_delay_m(*args, el)))


INT_FORMATS = {"ci16": (2, np.int16), "ci8": (3, np.int8), "cu8": (4, np.uint8)}


def to_int_format(x: np.ndarray, fmt: str) -> np.ndarray:
    """IqFormat::{Ci16, Ci8, Cu8}.write_sample (core/io/format.rs:203-222) -> [n][2] integers (re, im)"""
    code, dt = INT_FORMATS[fmt]
    a = np.ascontiguousarray(x, np.complex128)
    out = np.empty((a.size, 2), dt)
    assert lib().orc_to_int_format(_ptr(a), a.size, code, _ptr(out)) == 0
    return out


def to_cf32(x: np.ndarray) -> np.ndarray:
    """IqFormat::Cf32 sink cast (core/io/format.rs:197-200)."""
    a = np.ascontiguousarray(x, np.complex128)
    out = np.zeros(a.size, np.complex64)
    lib().orc_to_cf32(_ptr(a), a.size, _ptr(out))
    return out


# ------------------------------------------------------------------ PCPS
class OraclePcps:
    """PcpsAcquisition restated (gnss/acquisition.rs:40-255)."""

    def __init__(self, code_length: int, sample_rate: float):
        self.p = PcpsPod()
        lib().orc_pcps_init(C.byref(self.p), code_length, sample_rate)

    def with_doppler_range(self, max_hz, step_hz):
        self.p.doppler_max_hz, self.p.doppler_step_hz = max_hz, step_hz
        return self

    def with_threshold(self, t):
        self.p.threshold = t
        return self

    def with_coherent_periods(self, n):
        self.p.coherent_periods = max(int(n), 1)
        return self

    def fft_size(self) -> int: return int(self.p.fft_size)
    def num_bins(self) -> int: return int(lib().orc_pcps_num_bins(C.byref(self.p)))

    def acquire(self, input_: np.ndarray, code: np.ndarray, prn: int) -> AcqResultPod:
        x = np.ascontiguousarray(input_, np.complex128); c = np.ascontiguousarray(code, np.int8)
        out = AcqResultPod()
        lib().orc_pcps_acquire(C.byref(self.p), _ptr(x), x.size, _ptr(c), c.size, prn, C.byref(out))
        return out

    def acquire_grid(self, input_: np.ndarray, code: np.ndarray):
        x = np.ascontiguousarray(input_, np.complex128); c = np.ascontiguousarray(code, np.int8)
        grid = np.zeros((self.num_bins(), int(self.p.code_length)), np.float64)
        lin = lib().orc_pcps_acquire_grid(C.byref(self.p), _ptr(x), x.size, _ptr(c), c.size, _ptr(grid))
        return grid, int(lin)


# ---- tracking channel (r4w_oracle_track.c; gnss/tracking.rs) ------------------------------------------------------
TRACK_STATE_DTYPE = np.dtype([("code_phase", "<f8"), ("carrier_freq_hz", "<f8"), ("carrier_phase_rad", "<f8"), ("prompt_i", "<f8"),
                              ("prompt_q", "<f8"), ("cn0_dbhz", "<f8"), ("ms_count", "<u8"), ("prn", "u1"), ("carrier_lock", "u1"),
                              ("code_lock", "u1"), ("bit_sync", "u1"), ("pad", "u1", (4,))])


class OracleTrackingChannel:
    """TrackingChannel (tracking.rs:107-358), f64, one code period per `process` call"""

    def __init__(self, prn, code_length, sample_rate, chipping_rate, initial_code_phase, initial_doppler):
        self._h = lib().orc_track_new(int(prn), int(code_length), float(sample_rate), float(chipping_rate),
                                      float(initial_code_phase), float(initial_doppler))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_track_free(self._h)
            self._h = None

    def with_dll_bandwidth(self, bw):
        lib().orc_track_set_dll_bandwidth(self._h, float(bw)); return self

    def with_pll_bandwidth(self, bw):
        lib().orc_track_set_pll_bandwidth(self._h, float(bw)); return self

    def process(self, samples: np.ndarray, code: np.ndarray):
        x = np.ascontiguousarray(samples, np.complex128)
        c = np.ascontiguousarray(code, np.int8)
        out = np.zeros(1, TRACK_STATE_DTYPE)
        lib().orc_track_process(self._h, _ptr(x), x.size, _ptr(c), _ptr(out))
        return out[0]

    def run(self, samples: np.ndarray, code: np.ndarray, n_per_period: int, n_periods: int) -> np.ndarray:
        """n_periods consecutive process calls -> structured array of the returned states"""
        x = np.ascontiguousarray(samples, np.complex128)
        c = np.ascontiguousarray(code, np.int8)
        out = np.zeros(n_periods, TRACK_STATE_DTYPE)
        for p in range(n_periods):
            seg = x[p * n_per_period:(p + 1) * n_per_period]
            lib().orc_track_process(self._h, _ptr(seg), seg.size, _ptr(c), C.c_void_p(out.ctypes.data + p * TRACK_STATE_DTYPE.itemsize))
        return out

    def state(self):
        out = np.zeros(1, TRACK_STATE_DTYPE)
        lib().orc_track_state_get(self._h, _ptr(out))
        return out[0]

    def nav_bits(self) -> np.ndarray:
        n = lib().orc_track_nav_bits(self._h, None, 0)
        out = np.zeros(int(n), np.int8)
        if n:
            lib().orc_track_nav_bits(self._h, _ptr(out), out.size)
        return out


def loop_filter_2nd_run(bw, T, disc, n) -> float:
    return float(lib().orc_loop_filter_2nd_run(float(bw), float(T), float(disc), int(n)))


def loop_filter_3rd_run(bw, T, disc, n) -> float:
    return float(lib().orc_loop_filter_3rd_run(float(bw), float(T), float(disc), int(n)))


def dll_s_curve(el_spacing: float, num_points: int):
    e = np.zeros(num_points); d = np.zeros(num_points)
    lib().orc_dll_s_curve(float(el_spacing), int(num_points), _ptr(e), _ptr(d))
    return e, d


# ---- r4w-sim generic scenario engine (r4w_oracle_sim.c; crates/r4w-sim/src/scenario/) -----------------------------------
def sim_trajectory(kind: str, params, t: float) -> np.ndarray:
    """Trajectory::state_at -> [x, y, z, vx, vy, vz] (ECEF).  kind: Static (lat, lon, alt), Linear (+ ve, vn, vu),
    Circular (+ radius_m, omega_rad_s, initial_bearing_deg), Waypoints (rows of t, lat, lon, alt)"""
    out = np.zeros(6)
    if kind == "Waypoints":
        p = np.ascontiguousarray(params, np.float64).reshape(-1, 4)
        lib().orc_sim_waypoints(_ptr(p), p.shape[0], float(t), _ptr(out))
    else:
        p = np.zeros(6); q = np.asarray(params, np.float64).ravel(); p[:q.size] = q
        lib().orc_sim_trajectory({"Static": 0, "Linear": 1, "Circular": 3}[kind], _ptr(p), float(t), _ptr(out))
    return out


def sim_link(rx6, em6, carrier_hz: float, power_dbm: float):
    """engine.rs:82-98 -> (range_m, doppler_hz, path_loss_db, rx_amplitude)"""
    rx = np.ascontiguousarray(rx6, np.float64); em = np.ascontiguousarray(em6, np.float64)
    out = np.zeros(4)
    lib().orc_sim_link(_ptr(rx[:3].copy()), _ptr(rx[3:].copy()), _ptr(em[:3].copy()), _ptr(em[3:].copy()), float(carrier_hz), float(power_dbm), _ptr(out))
    return tuple(float(v) for v in out)


def sim_noise_power(noise_floor_dbw_hz: float, sample_rate: float) -> float:
    return float(lib().orc_sim_noise_power(float(noise_floor_dbw_hz), float(sample_rate)))


class OracleComposer:
    """the Doppler / amplitude / sum loops of ScenarioEngine::generate_block (engine.rs:105-122), noise-free"""

    def __init__(self, n_emitters: int, sample_rate: float):
        self.fs = float(sample_rate)
        self.phases = np.zeros(n_emitters)

    def block(self, baseband: np.ndarray, doppler_hz, amplitude, active=None) -> np.ndarray:
        bb = np.ascontiguousarray(baseband, np.complex128)
        out = np.zeros(bb.shape[1], np.complex128)
        for e in range(bb.shape[0]):
            if active is not None and not active[e]:
                continue
            ph = C.c_double(self.phases[e])
            lib().orc_sim_compose_emitter(_ptr(bb[e]), bb.shape[1], float(doppler_hz[e]), self.fs, float(amplitude[e]), C.byref(ph), _ptr(out))
            self.phases[e] = ph.value
        return out
