"""Host-side mirror of r4w-sim's generic scenario engine (crates/r4w-sim/src/scenario/) over the C-ABI composer.

`ScenarioConfig` (config.rs), `Trajectory` (trajectory.rs: Static / Linear / Waypoints / Circular), the `Emitter` protocol
(emitter.rs: state_at, generate_iq, carrier_frequency_hz, nominal_power_dbm, id) and `ScenarioEngine` (engine.rs:
generate_block, generate_all, reset, emitter_status, progress, is_done) keep the reference's names and meaning.  The
per-block geometry (receiver state at the block midpoint, range, range rate, free-space path loss: engine.rs:68-98) is f64
host arithmetic here as in the reference; the per-sample work — Doppler rotation with the continuously accumulated carrier
phase, amplitude, sum over emitters, receiver noise — runs on the GPU (r4wb_composer_block, csrc/compose.cu).
Noise is the library's Philox stream: the reference's StdRng draw cannot be reproduced, only its statistics."""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib

SPEED_OF_LIGHT = 299_792_458.0
_WGS84_A = 6_378_137.0
_WGS84_F = 1.0 / 298.257_223_563
_WGS84_E2 = 2.0 * _WGS84_F - _WGS84_F * _WGS84_F


def lla_to_ecef(lat_deg: float, lon_deg: float, alt_m: float) -> np.ndarray:
    """r4w_core::coordinates::lla_to_ecef (coordinates.rs:129-144)"""
    lat, lon = math.radians(lat_deg), math.radians(lon_deg)
    sin_lat, cos_lat, sin_lon, cos_lon = math.sin(lat), math.cos(lat), math.sin(lon), math.cos(lon)
    n = _WGS84_A / math.sqrt(1.0 - _WGS84_E2 * sin_lat * sin_lat)
    return np.array([(n + alt_m) * cos_lat * cos_lon, (n + alt_m) * cos_lat * sin_lon, (n * (1.0 - _WGS84_E2) + alt_m) * sin_lat])


def range_rate(obs_pos, obs_vel, tgt_pos, tgt_vel) -> float:
    """coordinates.rs:225-237: relative velocity projected on the unit vector observer -> target"""
    d = np.asarray(tgt_pos, float) - np.asarray(obs_pos, float)
    r = math.sqrt(float(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]))
    if r < 1e-10:
        return 0.0
    u = d / r
    rv = np.asarray(tgt_vel, float) - np.asarray(obs_vel, float)
    return float(rv[0] * u[0] + rv[1] * u[1] + rv[2] * u[2])


def fspl_db(distance_m: float, frequency_hz: float) -> float:
    """coordinates.rs:240-246"""
    if distance_m <= 0.0 or frequency_hz <= 0.0:
        return 0.0
    return 20.0 * math.log10(4.0 * math.pi * distance_m * frequency_hz / SPEED_OF_LIGHT)


@dataclass
class ScenarioConfig:                      # config.rs:10-37
    duration_s: float = 0.001
    sample_rate: float = 2_046_000.0
    center_frequency_hz: float = 1_575_420_000.0
    block_size: int = 2046
    noise_floor_dbw_hz: float = -204.0
    seed: int = 42

    def total_samples(self) -> int:
        return int(math.ceil(self.duration_s * self.sample_rate))

    def num_blocks(self) -> int:
        return (self.total_samples() + self.block_size - 1) // self.block_size

    def noise_power_linear(self) -> float:
        return 10.0 ** (self.noise_floor_dbw_hz / 10.0) * self.sample_rate


@dataclass
class TrajectoryState:
    position: np.ndarray
    velocity: np.ndarray
    time_s: float


@dataclass
class Trajectory:
    """trajectory.rs:20-43.  kind: "Static" (position), "Linear" (start, velocity_enu), "Waypoints" (points = [(t, (lat, lon,
    alt)), ...]), "Circular" (center, radius_m, omega_rad_s, initial_bearing_deg); positions are (lat_deg, lon_deg, alt_m)."""
    kind: str = "Static"
    position: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    velocity_enu: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    points: Sequence = ()
    radius_m: float = 0.0
    omega_rad_s: float = 0.0
    initial_bearing_deg: float = 0.0

    def state_at(self, t: float) -> TrajectoryState:          # trajectory.rs:45-170
        zero = np.zeros(3)
        if self.kind == "Static":
            return TrajectoryState(lla_to_ecef(*self.position), zero, t)
        if self.kind == "Waypoints":
            pts = list(self.points)
            if not pts:
                return TrajectoryState(zero.copy(), zero, t)
            if len(pts) == 1 or t <= pts[0][0]:
                return TrajectoryState(lla_to_ecef(*pts[0][1]), zero, t)
            if t >= pts[-1][0]:
                return TrajectoryState(lla_to_ecef(*pts[-1][1]), zero, t)
            idx = next(k for k, (pt, _) in enumerate(pts) if pt > t) - 1
            (t0, p0), (t1, p1) = pts[idx], pts[idx + 1]
            dt = t1 - t0
            frac = (t - t0) / dt
            e0, e1 = lla_to_ecef(*p0), lla_to_ecef(*p1)
            return TrajectoryState(e0 + (e1 - e0) * frac, (e1 - e0) / dt, t)
        lat, lon = math.radians(self.position[0]), math.radians(self.position[1])
        sin_lat, cos_lat, sin_lon, cos_lon = math.sin(lat), math.cos(lat), math.sin(lon), math.cos(lon)
        e = lla_to_ecef(*self.position)
        if self.kind == "Linear":
            ve, vn, vu = self.velocity_enu
            v = np.array([-sin_lon * ve - sin_lat * cos_lon * vn + cos_lat * cos_lon * vu,
                          cos_lon * ve - sin_lat * sin_lon * vn + cos_lat * sin_lon * vu,
                          cos_lat * vn + sin_lat * vu])
            return TrajectoryState(e + v * t, v, t)
        if self.kind == "Circular":
            bearing = math.radians(self.initial_bearing_deg) + self.omega_rad_s * t
            east, north = self.radius_m * math.sin(bearing), self.radius_m * math.cos(bearing)
            pos = e + np.array([-sin_lon * east - sin_lat * cos_lon * north, cos_lon * east - sin_lat * sin_lon * north, cos_lat * north])
            d_east = self.radius_m * self.omega_rad_s * math.cos(bearing)
            d_north = -self.radius_m * self.omega_rad_s * math.sin(bearing)
            vel = np.array([-sin_lon * d_east - sin_lat * cos_lon * d_north, cos_lon * d_east - sin_lat * sin_lon * d_north, cos_lat * d_north])
            return TrajectoryState(pos, vel, t)
        raise ValueError(f"unknown trajectory kind {self.kind!r}")


@dataclass
class EmitterState:                          # emitter.rs:10-19
    position: np.ndarray
    velocity: np.ndarray
    power_dbm: float
    active: bool = True


@dataclass
class EmitterStatus:                         # engine.rs:17-28
    id: str
    range_m: float
    doppler_hz: float
    path_loss_db: float
    received_power_dbm: float
    active: bool


class ScenarioEngine:
    """engine.rs:30-214.  `emitters` are objects with the Emitter trait's methods: state_at(t) -> EmitterState,
    generate_iq(t, num_samples, sample_rate) -> complex array, carrier_frequency_hz(), nominal_power_dbm(), id()."""

    def __init__(self, config: ScenarioConfig, emitters: List, trajectory: Trajectory, noise: bool = True):
        _lib.ensure_init()
        self.config, self.emitters, self.trajectory = config, list(emitters), trajectory
        self.current_sample = 0
        noise_std = math.sqrt(config.noise_power_linear() / 2.0) if noise else 0.0        # engine.rs:126-127
        self._h = C.c_void_p()
        _lib.check(_lib.lib().r4wb_composer_create(len(self.emitters), float(config.sample_rate), noise_std, int(config.seed), C.byref(self._h)))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.lib().r4wb_composer_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _link(self, rx: TrajectoryState, em, st: EmitterState):
        d = st.position - rx.position
        range_m = math.sqrt(float(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]))
        carrier_hz = em.carrier_frequency_hz()
        doppler_hz = -range_rate(rx.position, rx.velocity, st.position, st.velocity) * carrier_hz / SPEED_OF_LIGHT
        pl_db = fspl_db(range_m, carrier_hz)
        return range_m, doppler_hz, pl_db, st.power_dbm - pl_db

    def generate_block(self) -> np.ndarray:
        """engine.rs:61-137: the next block of composite IQ (complex128 like Vec<Complex64>; empty when done)"""
        cfg = self.config
        remaining = max(cfg.total_samples() - self.current_sample, 0)
        n = min(remaining, cfg.block_size)
        if n == 0:
            return np.zeros(0, np.complex128)
        t_start = self.current_sample / cfg.sample_rate
        t_mid = t_start + (n / 2.0) / cfg.sample_rate
        rx = self.trajectory.state_at(t_mid)
        E = len(self.emitters)
        bb = np.zeros((max(E, 1), n), np.complex128)
        dop, amp, act = np.zeros(max(E, 1)), np.zeros(max(E, 1)), np.zeros(max(E, 1), np.uint8)
        for k, em in enumerate(self.emitters):
            st = em.state_at(t_mid)
            if not st.active:
                continue
            _, dop[k], _, rx_power_dbm = self._link(rx, em, st)
            amp[k] = 10.0 ** ((rx_power_dbm - 30.0) / 20.0)                         # dBm -> linear voltage, engine.rs:97-98
            bb[k] = np.asarray(em.generate_iq(t_start, n, cfg.sample_rate), np.complex128)
            act[k] = 1
        out = np.zeros(n, np.complex128)
        _lib.check(_lib.lib().r4wb_composer_block(self._h, bb.ctypes.data_as(C.c_void_p), _lib.FMT_CF64, _lib.MEM_HOST, n,
                                                  dop.ctypes.data_as(C.c_void_p), amp.ctypes.data_as(C.c_void_p),
                                                  act.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), _lib.FMT_CF64, _lib.MEM_HOST))
        self.current_sample += n
        return out

    def generate_all(self) -> np.ndarray:
        parts = []
        while self.current_sample < self.config.total_samples():
            parts.append(self.generate_block())
        return np.concatenate(parts) if parts else np.zeros(0, np.complex128)

    def reset(self):
        self.current_sample = 0
        _lib.check(_lib.lib().r4wb_composer_reset(self._h))

    def carrier_phases(self) -> np.ndarray:
        out = np.zeros(max(len(self.emitters), 1))
        _lib.check(_lib.lib().r4wb_composer_phases(self._h, out.ctypes.data_as(C.c_void_p), out.size))
        return out[:len(self.emitters)]

    def emitter_status(self, t: float) -> List[EmitterStatus]:
        rx = self.trajectory.state_at(t)
        res = []
        for em in self.emitters:
            st = em.state_at(t)
            range_m, doppler_hz, pl_db, rx_power = self._link(rx, em, st)
            res.append(EmitterStatus(em.id(), range_m, doppler_hz, pl_db, rx_power, bool(st.active)))
        return res

    def progress(self) -> float:
        total = self.config.total_samples()
        return 1.0 if total == 0 else self.current_sample / total

    def is_done(self) -> bool:
        return self.current_sample >= self.config.total_samples()
