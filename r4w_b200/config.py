"""Host-side mirror of r4w's GNSS scenario configuration (the drop-in input format).

Mirrors `GnssScenarioConfig` and friends (crates/r4w-core/src/waveform/gnss/scenario_config.rs:137-191,
304-315, 383-401, 417-437, 455-487, 537-547) and reads the same YAML the reference CLI feeds to
`serde_yaml::from_str::<GnssScenarioConfig>` (crates/r4w-cli/src/main.rs:4108-4113):

* `antenna: !Patch {peak_gain_dbi, beamwidth_deg}` — serde_yaml externally tagged enum
  (gnss/environment/antenna.rs:11-31);
* `ionosphere_source: {type: Klobuchar}` / `ephemeris_source: {type: Nominal}` — adjacently tagged;
* optional per-satellite keys absent or null -> None; `orbital_dynamics` absent -> false;
* unknown keys are ignored (no deny_unknown_fields).

`to_pod()` produces the ctypes structure that both `libr4w_b200.so` (include/r4w_b200.h,
`r4wb_scenario_cfg`) and the test oracle (`orc_scenario_cfg`, same layout) accept.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from dataclasses import dataclass, field
from typing import List, Optional

import yaml

# ---------------------------------------------------------------------------- enums
SIGNALS = ["GpsL1Ca", "GpsL5", "GlonassL1of", "GalileoE1", "GalileoE1C", "GalileoE1OS"]  # gnss/types.rs:33-47
ANTENNAS = ["Isotropic", "Hemispherical", "Patch", "ChokeRing"]
MULTIPATH_PRESETS = ["OpenSky", "Suburban", "UrbanCanyon", "Indoor"]

HAS_ELEVATION, HAS_AZIMUTH, HAS_RANGE, HAS_RANGE_RATE = 1, 2, 4, 8
HAS_DOPPLER, HAS_DOPPLER_RATE, HAS_CN0, HAS_IONO, HAS_TROPO = 16, 32, 64, 128, 256

FLAG_NOISE_OFF = 1
FLAG_CLOSED_FORM_PHASE = 2


# ---------------------------------------------------------------------------- ctypes PODs (include/r4w_b200.h)
class LlaPod(C.Structure):
    _fields_ = [("lat_deg", C.c_double), ("lon_deg", C.c_double), ("alt_m", C.c_double)]


class SatCfgPod(C.Structure):
    _fields_ = [
        ("signal", C.c_uint32), ("has", C.c_uint32),
        ("prn", C.c_uint8), ("plane", C.c_uint8), ("slot", C.c_uint8), ("nav_data", C.c_uint8),
        ("orbital_dynamics", C.c_uint8), ("_pad", C.c_uint8 * 3),
        ("tx_power_dbw", C.c_double),
        ("elevation_deg", C.c_double), ("azimuth_deg", C.c_double), ("range_m", C.c_double),
        ("range_rate_mps", C.c_double), ("doppler_hz", C.c_double), ("doppler_rate_hz_per_s", C.c_double),
        ("cn0_dbhz", C.c_double), ("iono_delay_m", C.c_double), ("tropo_delay_m", C.c_double),
    ]


class ReceiverCfgPod(C.Structure):
    _fields_ = [
        ("position", LlaPod), ("antenna", C.c_uint32), ("has_trajectory", C.c_uint32),
        ("antenna_peak_gain_dbi", C.c_double), ("antenna_beamwidth_deg", C.c_double),
        ("elevation_mask_deg", C.c_double), ("noise_figure_db", C.c_double), ("bandwidth_hz", C.c_double),
        ("traj_start", LlaPod), ("traj_end", LlaPod), ("traj_has_speed", C.c_uint32), ("_pad", C.c_uint32),
        ("traj_speed_mps", C.c_double),
    ]


class EnvironmentCfgPod(C.Structure):
    _fields_ = [("ionosphere_enabled", C.c_uint32), ("troposphere_enabled", C.c_uint32),
                ("multipath_enabled", C.c_uint32), ("multipath_preset", C.c_uint32),
                ("klobuchar_alpha", C.c_double * 4), ("klobuchar_beta", C.c_double * 4),
                ("tropo_height_m", C.c_double), ("tropo_temperature_k", C.c_double), ("tropo_pressure_hpa", C.c_double),
                ("tropo_relative_humidity", C.c_double)]


class OutputCfgPod(C.Structure):
    _fields_ = [("sample_rate", C.c_double), ("duration_s", C.c_double), ("block_size", C.c_uint64),
                ("seed", C.c_uint64), ("start_time_gps_s", C.c_double), ("lpf_cutoff_hz", C.c_double)]


class ScenarioCfgPod(C.Structure):
    _fields_ = [("n_sats", C.c_uint32), ("flags", C.c_uint32), ("sats", C.POINTER(SatCfgPod)),
                ("receiver", ReceiverCfgPod), ("environment", EnvironmentCfgPod), ("output", OutputCfgPod)]


class SatStatusPod(C.Structure):
    _fields_ = [
        ("signal", C.c_uint32), ("prn", C.c_uint8), ("visible", C.c_uint8), ("_pad", C.c_uint8 * 2),
        ("elevation_deg", C.c_double), ("azimuth_deg", C.c_double), ("range_m", C.c_double),
        ("range_rate_mps", C.c_double), ("doppler_hz", C.c_double), ("cn0_dbhz", C.c_double),
        ("iono_delay_m", C.c_double), ("tropo_delay_m", C.c_double), ("antenna_gain_dbi", C.c_double),
        ("clock_correction_s", C.c_double),
    ]


class AcqResultPod(C.Structure):
    _fields_ = [("prn", C.c_uint8), ("detected", C.c_uint8), ("has_cn0", C.c_uint8), ("_pad", C.c_uint8 * 5),
                ("code_phase", C.c_double), ("doppler_hz", C.c_double), ("peak_metric", C.c_double),
                ("threshold", C.c_double), ("cn0_estimate", C.c_double)]


# ---------------------------------------------------------------------------- dataclasses (reference field names)
@dataclass
class LlaPosition:
    lat_deg: float
    lon_deg: float
    alt_m: float


@dataclass
class AntennaPattern:
    kind: str = "Patch"
    peak_gain_dbi: float = 5.0      # AntennaPattern::default_patch
    beamwidth_deg: float = 150.0


@dataclass
class SatelliteConfig:
    signal: str
    prn: int
    plane: int
    slot: int
    tx_power_dbw: float
    nav_data: bool
    elevation_deg: Optional[float] = None
    azimuth_deg: Optional[float] = None
    range_m: Optional[float] = None
    range_rate_mps: Optional[float] = None
    doppler_hz: Optional[float] = None
    doppler_rate_hz_per_s: Optional[float] = None
    orbital_dynamics: bool = False
    cn0_dbhz: Optional[float] = None
    iono_delay_m: Optional[float] = None
    tropo_delay_m: Optional[float] = None


@dataclass
class ReceiverTrajectory:
    start: LlaPosition
    end: LlaPosition
    speed_mps: Optional[float] = None
    description: Optional[str] = None


@dataclass
class ReceiverConfig:
    position: LlaPosition = field(default_factory=lambda: LlaPosition(41.08, -85.14, 240.0))
    antenna: AntennaPattern = field(default_factory=AntennaPattern)
    elevation_mask_deg: float = 5.0
    noise_figure_db: float = 2.0
    bandwidth_hz: float = 5_000_000.0
    trajectory: Optional[ReceiverTrajectory] = None


@dataclass
class EnvironmentConfig:
    ionosphere_enabled: bool = True
    ionosphere_model: Optional[dict] = None
    ionosphere_source: str = "Klobuchar"
    troposphere_enabled: bool = True
    troposphere_model: Optional[dict] = None
    multipath_preset: str = "OpenSky"
    multipath_enabled: bool = False
    ephemeris_source: str = "Nominal"


def _julian_day(year: int, month: int, day: int) -> int:
    """scenario_config.rs:509-517 (integer division truncates like Rust's for these positive operands)"""
    a = (14 - month) // 12
    y2 = year + 4800 - a
    m2 = month + 12 * a - 3
    return day + (153 * m2 + 2) // 5 + 365 * y2 + y2 // 4 - y2 // 100 + y2 // 400 - 32045


def gps_time_from_utc(year: int, month: int, day: int, hour: int, minute: int, sec: float) -> float:
    """gps_time_from_utc, scenario_config.rs:498-506: seconds since the GPS epoch, 18 leap seconds"""
    days = float(_julian_day(year, month, day) - _julian_day(1980, 1, 6))
    return days * 86400.0 + hour * 3600.0 + minute * 60.0 + sec + 18.0


@dataclass
class OutputConfig:
    sample_rate: float = 5_000_000.0
    duration_s: float = 0.001
    block_size: int = 0
    seed: int = 42
    start_time_gps_s: float = field(default_factory=lambda: gps_time_from_utc(2026, 2, 4, 20, 0, 0.0))   # OutputConfig::default, :519-533
    format: str = "cf32"
    lpf_cutoff_hz: float = 0.0
    output_path: Optional[str] = None


@dataclass
class GnssScenarioConfig:
    satellites: List[SatelliteConfig]
    receiver: ReceiverConfig = field(default_factory=ReceiverConfig)
    environment: EnvironmentConfig = field(default_factory=EnvironmentConfig)
    output: OutputConfig = field(default_factory=OutputConfig)

    # ---- reference-shaped helpers
    def total_samples(self) -> int:
        import math
        return int(math.ceil(self.output.duration_s * self.output.sample_rate))   # scenario.rs:80

    def block_size(self) -> int:
        import math
        if self.output.block_size > 0:
            return int(self.output.block_size)
        return int(math.ceil(self.output.sample_rate * 0.001))                    # scenario.rs:667-674

    def copy(self) -> "GnssScenarioConfig":
        return dataclasses.replace(
            self,
            satellites=[dataclasses.replace(s) for s in self.satellites],
            receiver=dataclasses.replace(self.receiver),
            environment=dataclasses.replace(self.environment),
            output=dataclasses.replace(self.output),
        )

    # ---- POD for the C ABI / oracle
    def to_pod(self, flags: int = 0):
        """Returns (pod, keepalive); keepalive owns the satellite array the pod points to."""
        n = len(self.satellites)
        arr = (SatCfgPod * max(n, 1))()
        for k, s in enumerate(self.satellites):
            p = arr[k]
            if s.signal not in SIGNALS:
                raise ValueError(f"unknown signal {s.signal!r}")
            p.signal = SIGNALS.index(s.signal)
            p.prn, p.plane, p.slot = int(s.prn), int(s.plane), int(s.slot)
            p.nav_data = 1 if s.nav_data else 0
            p.orbital_dynamics = 1 if s.orbital_dynamics else 0
            p.tx_power_dbw = float(s.tx_power_dbw)
            has = 0
            for bit, name in ((HAS_ELEVATION, "elevation_deg"), (HAS_AZIMUTH, "azimuth_deg"), (HAS_RANGE, "range_m"),
                              (HAS_RANGE_RATE, "range_rate_mps"), (HAS_DOPPLER, "doppler_hz"),
                              (HAS_DOPPLER_RATE, "doppler_rate_hz_per_s"), (HAS_CN0, "cn0_dbhz"),
                              (HAS_IONO, "iono_delay_m"), (HAS_TROPO, "tropo_delay_m")):
                v = getattr(s, name)
                if v is not None:
                    has |= bit
                    setattr(p, name, float(v))
            p.has = has
        pod = ScenarioCfgPod()
        pod.n_sats = n
        pod.flags = flags
        pod.sats = C.cast(arr, C.POINTER(SatCfgPod))
        r = self.receiver
        pod.receiver.position = LlaPod(r.position.lat_deg, r.position.lon_deg, r.position.alt_m)
        pod.receiver.antenna = ANTENNAS.index(r.antenna.kind)
        pod.receiver.antenna_peak_gain_dbi = r.antenna.peak_gain_dbi
        pod.receiver.antenna_beamwidth_deg = r.antenna.beamwidth_deg
        pod.receiver.elevation_mask_deg = r.elevation_mask_deg
        pod.receiver.noise_figure_db = r.noise_figure_db
        pod.receiver.bandwidth_hz = r.bandwidth_hz
        if r.trajectory is not None:
            t = r.trajectory
            pod.receiver.has_trajectory = 1
            pod.receiver.traj_start = LlaPod(t.start.lat_deg, t.start.lon_deg, t.start.alt_m)
            pod.receiver.traj_end = LlaPod(t.end.lat_deg, t.end.lon_deg, t.end.alt_m)
            if t.speed_mps is not None:
                pod.receiver.traj_has_speed = 1
                pod.receiver.traj_speed_mps = float(t.speed_mps)
        e = self.environment
        pod.environment.ionosphere_enabled = 1 if (e.ionosphere_enabled and e.ionosphere_source != "Disabled") else 0
        pod.environment.troposphere_enabled = 1 if e.troposphere_enabled else 0
        # KlobucharModel::default_broadcast (environment/ionosphere.rs:28-33) / SaastamoinenModel::standard_atmosphere
        # (troposphere.rs:28-35) when the YAML leaves the model null (scenario.rs:141-150)
        im = e.ionosphere_model or {}
        alpha = [float(v) for v in im.get("alpha", [0.1118e-7, 0.7451e-8, -0.5961e-7, -0.1192e-6])]
        beta = [float(v) for v in im.get("beta", [0.1167e6, -0.4267e5, -0.2621e6, 0.1311e6])]
        pod.environment.klobuchar_alpha = (C.c_double * 4)(*alpha)
        pod.environment.klobuchar_beta = (C.c_double * 4)(*beta)
        tm = e.troposphere_model or {}
        pod.environment.tropo_height_m = float(tm.get("height_m", 0.0))
        pod.environment.tropo_temperature_k = float(tm.get("temperature_k", 288.15))
        pod.environment.tropo_pressure_hpa = float(tm.get("pressure_hpa", 1013.25))
        pod.environment.tropo_relative_humidity = float(tm.get("relative_humidity", 0.5))
        pod.environment.multipath_enabled = 1 if e.multipath_enabled else 0
        pod.environment.multipath_preset = MULTIPATH_PRESETS.index(e.multipath_preset) if e.multipath_preset in MULTIPATH_PRESETS else 0
        o = self.output
        pod.output.sample_rate = o.sample_rate
        pod.output.duration_s = o.duration_s
        pod.output.block_size = int(o.block_size)
        pod.output.seed = int(o.seed)
        pod.output.start_time_gps_s = o.start_time_gps_s
        pod.output.lpf_cutoff_hz = o.lpf_cutoff_hz
        return pod, arr


# ---------------------------------------------------------------------------- YAML
class _Loader(yaml.SafeLoader):
    pass


def _tagged(loader, suffix, node):
    """serde_yaml externally tagged enum: `!Variant {fields}` or `!Variant` (unit)."""
    if isinstance(node, yaml.MappingNode):
        value = loader.construct_mapping(node, deep=True)
    elif isinstance(node, yaml.SequenceNode):
        value = loader.construct_sequence(node, deep=True)
    else:
        value = loader.construct_scalar(node)
    return {"__tag__": suffix, "value": value}


_Loader.add_multi_constructor("!", _tagged)


def _opt_float(d, key):
    v = d.get(key)
    return None if v is None else float(v)


def _lla(d) -> LlaPosition:
    return LlaPosition(float(d["lat_deg"]), float(d["lon_deg"]), float(d["alt_m"]))


def _antenna(v) -> AntennaPattern:
    if v is None:
        return AntennaPattern()
    if isinstance(v, str):                      # unit variant written as a plain string
        return AntennaPattern(kind=v, peak_gain_dbi=0.0, beamwidth_deg=0.0)
    if isinstance(v, dict) and "__tag__" in v:
        kind, body = v["__tag__"], v["value"] if isinstance(v["value"], dict) else {}
    elif isinstance(v, dict) and len(v) == 1:   # `Patch: {..}` map form
        kind, body = next(iter(v.items()))
        body = body or {}
    else:
        raise ValueError(f"cannot parse antenna {v!r}")
    if kind not in ANTENNAS:
        raise ValueError(f"unknown antenna pattern {kind!r}")
    return AntennaPattern(kind=kind, peak_gain_dbi=float(body.get("peak_gain_dbi", 0.0)),
                          beamwidth_deg=float(body.get("beamwidth_deg", 0.0)))


def _adjacent_tag(v, default):
    if v is None:
        return default
    if isinstance(v, dict):
        if "__tag__" in v:
            return v["__tag__"]
        return str(v.get("type", default))
    return str(v)


def config_from_dict(doc: dict) -> GnssScenarioConfig:
    sats = []
    for s in doc["satellites"]:
        sats.append(SatelliteConfig(
            signal=str(s["signal"]), prn=int(s["prn"]), plane=int(s["plane"]), slot=int(s["slot"]),
            tx_power_dbw=float(s["tx_power_dbw"]), nav_data=bool(s["nav_data"]),
            elevation_deg=_opt_float(s, "elevation_deg"), azimuth_deg=_opt_float(s, "azimuth_deg"),
            range_m=_opt_float(s, "range_m"), range_rate_mps=_opt_float(s, "range_rate_mps"),
            doppler_hz=_opt_float(s, "doppler_hz"), doppler_rate_hz_per_s=_opt_float(s, "doppler_rate_hz_per_s"),
            orbital_dynamics=bool(s.get("orbital_dynamics", False)),
            cn0_dbhz=_opt_float(s, "cn0_dbhz"), iono_delay_m=_opt_float(s, "iono_delay_m"),
            tropo_delay_m=_opt_float(s, "tropo_delay_m")))
    r = doc["receiver"]
    traj = None
    if r.get("trajectory") is not None:
        t = r["trajectory"]
        traj = ReceiverTrajectory(start=_lla(t["start"]), end=_lla(t["end"]),
                                  speed_mps=_opt_float(t, "speed_mps"), description=t.get("description"))
    receiver = ReceiverConfig(position=_lla(r["position"]), antenna=_antenna(r.get("antenna")),
                              elevation_mask_deg=float(r["elevation_mask_deg"]),
                              noise_figure_db=float(r["noise_figure_db"]), bandwidth_hz=float(r["bandwidth_hz"]),
                              trajectory=traj)
    e = doc["environment"]
    environment = EnvironmentConfig(
        ionosphere_enabled=bool(e["ionosphere_enabled"]), ionosphere_model=e.get("ionosphere_model"),
        ionosphere_source=_adjacent_tag(e.get("ionosphere_source"), "Klobuchar"),
        troposphere_enabled=bool(e["troposphere_enabled"]), troposphere_model=e.get("troposphere_model"),
        multipath_preset=_adjacent_tag(e.get("multipath_preset"), "OpenSky"),
        multipath_enabled=bool(e["multipath_enabled"]),
        ephemeris_source=_adjacent_tag(e.get("ephemeris_source"), "Nominal"))
    o = doc["output"]
    output = OutputConfig(sample_rate=float(o["sample_rate"]), duration_s=float(o["duration_s"]),
                          block_size=int(o["block_size"]), seed=int(o["seed"]),
                          start_time_gps_s=float(o["start_time_gps_s"]), format=str(o.get("format", "cf32")),
                          lpf_cutoff_hz=float(o.get("lpf_cutoff_hz", 0.0) or 0.0), output_path=o.get("output_path"))
    return GnssScenarioConfig(satellites=sats, receiver=receiver, environment=environment, output=output)


def load_config(path, cli_elevation_mask_deg: Optional[float] = None) -> GnssScenarioConfig:
    """Parse an r4w scenario YAML.  `cli_elevation_mask_deg` reproduces the CLI's unconditional
    `config.receiver.elevation_mask_deg = elevation_mask` (crates/r4w-cli/src/main.rs:4133, default 5.0)."""
    with open(path, "r") as f:
        doc = yaml.load(f, Loader=_Loader)
    cfg = config_from_dict(doc)
    if cli_elevation_mask_deg is not None:
        cfg.receiver.elevation_mask_deg = float(cli_elevation_mask_deg)
    return cfg


def loads_config(text: str) -> GnssScenarioConfig:
    return config_from_dict(yaml.load(text, Loader=_Loader))


# ---- constellation tables and presets (scenario_config.rs:36-135, 199-241, 549-700) ----------------------------------------
# (PRN, plane, slot): GPS from NAVCEN, Galileo from the GSC (the reference's lookup tables, restated as data)
GPS_CONSTELLATION = [
    (24, 0, 0), (31, 0, 1), (30, 0, 2), (7, 0, 3), (28, 0, 5),
    (16, 1, 0), (25, 1, 1), (22, 1, 2), (12, 1, 3), (26, 1, 4), (14, 1, 5),
    (29, 2, 0), (27, 2, 1), (8, 2, 2), (17, 2, 3), (19, 2, 4),
    (2, 3, 0), (1, 3, 1), (6, 3, 3), (11, 3, 4), (18, 3, 5),
    (3, 4, 0), (10, 4, 1), (5, 4, 2), (23, 4, 4), (21, 4, 5),
    (32, 5, 0), (15, 5, 1), (9, 5, 2), (4, 5, 3), (13, 5, 5),
]
GALILEO_CONSTELLATION = [
    (31, 0, 0), (23, 0, 1), (21, 0, 2), (27, 0, 3), (30, 0, 4), (2, 0, 5), (25, 0, 6), (16, 0, 7),
    (13, 1, 0), (15, 1, 1), (34, 1, 2), (36, 1, 3), (11, 1, 4), (12, 1, 5), (33, 1, 6), (26, 1, 7),
    (5, 2, 0), (9, 2, 1), (4, 2, 2), (19, 2, 3), (29, 2, 4), (7, 2, 5), (8, 2, 6), (3, 2, 7),
]


def lookup_prn(table, plane: int, slot: int) -> int:
    """SatelliteConfig::lookup_prn (:199-211): 0 when no satellite sits in the slot"""
    for prn, p, s in table:
        if p == plane and s == slot:
            return prn
    return 0


def gps_l1ca(plane: int, slot: int) -> SatelliteConfig:
    """SatelliteConfig::gps_l1ca (:214-226): everything else from the nominal orbit and the link budget"""
    prn = lookup_prn(GPS_CONSTELLATION, plane, slot)
    if prn == 0:
        raise ValueError(f"No GPS satellite at plane={plane} slot={slot}")
    return SatelliteConfig("GpsL1Ca", prn, plane, slot, 14.3, True)


def galileo_e1(plane: int, slot: int) -> SatelliteConfig:
    """SatelliteConfig::galileo_e1 (:229-241)"""
    prn = lookup_prn(GALILEO_CONSTELLATION, plane, slot)
    if prn == 0:
        raise ValueError(f"No Galileo satellite at plane={plane} slot={slot}")
    return SatelliteConfig("GalileoE1", prn, plane, slot, 15.0, True)


PRESETS = ("OpenSky", "UrbanCanyon", "Driving", "Walking", "HighDynamics", "MultiConstellation")
_GPS8 = [(4, 2), (2, 3), (1, 4), (3, 3), (2, 4), (1, 3), (4, 1), (5, 2)]


def preset_config(name: str) -> GnssScenarioConfig:
    """GnssScenarioPreset::to_config (:581-700): the six presets of `r4w gnss scenario --preset`"""
    gps = lambda n: [gps_l1ca(p, s) for p, s in _GPS8[:n]]           # noqa: E731
    if name == "OpenSky":
        return GnssScenarioConfig(gps(8))
    if name == "UrbanCanyon":
        return GnssScenarioConfig(gps(8), receiver=ReceiverConfig(elevation_mask_deg=15.0),
                                  environment=EnvironmentConfig(multipath_preset="UrbanCanyon", multipath_enabled=True))
    if name == "Driving":
        return GnssScenarioConfig(gps(7), environment=EnvironmentConfig(multipath_preset="Suburban", multipath_enabled=True),
                                  output=OutputConfig(duration_s=0.01))
    if name == "Walking":
        return GnssScenarioConfig(gps(6), output=OutputConfig(duration_s=0.01))
    if name == "HighDynamics":
        return GnssScenarioConfig(gps(8), output=OutputConfig(duration_s=0.001))
    if name == "MultiConstellation":
        gal = [galileo_e1(p, s) for p, s in ((0, 7), (1, 5), (0, 0), (1, 6), (1, 4), (0, 6))]
        return GnssScenarioConfig(gps(5) + gal)
    raise ValueError(f"unknown preset {name!r}; one of {PRESETS}")


# ---- serde_yaml-shaped export (the CLI's companion `<output>.yaml`, crates/r4w-cli/src/main.rs:4511-4531, and --export-preset) ----
def _ryu(x: float) -> str:
    """f64 the way serde_yaml 0.9 prints it (ryu shortest round-trip; `.inf`/`.nan` spellings)"""
    x = float(x)
    if x != x:
        return ".nan"
    if x in (float("inf"), float("-inf")):
        return ".inf" if x > 0 else "-.inf"
    r = repr(x)
    if "e" in r:
        mant, exp = r.split("e")
        e = int(exp)
        if e == -5:                                       # ryu stays positional down to 1e-5; Python switches at 1e-5
            return f"{x:.{len(mant.replace('.', '').replace('-', '')) + 4}f}".rstrip("0")
        return f"{mant}e{e}"
    return r


def _yaml_str(s: str) -> str:
    plain = bool(s) and s[0] not in "{}[]&*!|>'\"%@`#,?:- " and ": " not in s and " #" not in s and not s.endswith((" ", ":")) \
        and s.lower() not in ("null", "true", "false", "~", "yes", "no", "on", "off")
    if plain:
        try:
            float(s)
            plain = False
        except ValueError:
            pass
    return s if plain else "'" + s.replace("'", "''") + "'"


def _scalar(v) -> str:
    if v is None:
        return "null"
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, int):
        return str(v)
    if isinstance(v, float):
        return _ryu(v)
    return _yaml_str(str(v))


class _Tagged:
    """externally tagged enum with a struct body: `key: !Variant` + the body's fields one level in"""
    def __init__(self, tag, body):
        self.tag, self.body = tag, body


def _emit(node, ind: int, out: list):
    pad = "  " * ind
    if isinstance(node, dict):
        for k, v in node.items():
            if isinstance(v, _Tagged):
                out.append(f"{pad}{k}: !{v.tag}")
                _emit(v.body, ind + 1, out)
            elif isinstance(v, dict) and v:
                out.append(f"{pad}{k}:")
                _emit(v, ind + 1, out)
            elif isinstance(v, (list, tuple)) and len(v):
                out.append(f"{pad}{k}:")
                _emit(list(v), ind, out)                  # libyaml: a sequence sits at its key's own indent
            elif isinstance(v, (dict, list, tuple)):
                out.append(f"{pad}{k}: " + ("{}" if isinstance(v, dict) else "[]"))
            else:
                out.append(f"{pad}{k}: {_scalar(v)}")
    elif isinstance(node, list):
        for item in node:
            if isinstance(item, dict) and item:
                sub: list = []
                _emit(item, ind + 1, sub)
                out.append(f"{pad}- {sub[0][2 * (ind + 1):]}")
                out.extend(sub[1:])
            else:
                out.append(f"{pad}- {_scalar(item)}")


def config_to_dict(cfg: GnssScenarioConfig) -> dict:
    """Field order and `skip_serializing_if = Option::is_none` as the reference's derive(Serialize) (scenario_config.rs:137-191,
    304-315, 383-401, 417-437, 455-487)"""
    lla = lambda p: {"lat_deg": float(p.lat_deg), "lon_deg": float(p.lon_deg), "alt_m": float(p.alt_m)}       # noqa: E731
    sats = []
    for s in cfg.satellites:
        d = {"signal": s.signal, "prn": int(s.prn), "plane": int(s.plane), "slot": int(s.slot),
             "tx_power_dbw": float(s.tx_power_dbw), "nav_data": bool(s.nav_data)}
        for key in ("elevation_deg", "azimuth_deg", "range_m", "range_rate_mps", "doppler_hz", "doppler_rate_hz_per_s"):
            if getattr(s, key) is not None:
                d[key] = float(getattr(s, key))
        d["orbital_dynamics"] = bool(s.orbital_dynamics)
        for key in ("cn0_dbhz", "iono_delay_m", "tropo_delay_m"):
            if getattr(s, key) is not None:
                d[key] = float(getattr(s, key))
        sats.append(d)
    r, a = cfg.receiver, cfg.receiver.antenna
    if a.kind == "Isotropic":
        antenna = "Isotropic"
    elif a.kind == "Patch":
        antenna = _Tagged("Patch", {"peak_gain_dbi": float(a.peak_gain_dbi), "beamwidth_deg": float(a.beamwidth_deg)})
    else:
        antenna = _Tagged(a.kind, {"peak_gain_dbi": float(a.peak_gain_dbi)})
    recv = {"position": lla(r.position), "antenna": antenna, "elevation_mask_deg": float(r.elevation_mask_deg),
            "noise_figure_db": float(r.noise_figure_db), "bandwidth_hz": float(r.bandwidth_hz)}
    if r.trajectory is not None:
        t = r.trajectory
        recv["trajectory"] = {"start": lla(t.start), "end": lla(t.end),
                              "speed_mps": None if t.speed_mps is None else float(t.speed_mps), "description": t.description}
    e = cfg.environment
    fl = lambda m: None if m is None else {k: ([float(x) for x in v] if isinstance(v, (list, tuple)) else float(v)) for k, v in m.items()}  # noqa: E731
    env = {"ionosphere_enabled": bool(e.ionosphere_enabled), "ionosphere_model": fl(e.ionosphere_model),
           "ionosphere_source": {"type": e.ionosphere_source}, "troposphere_enabled": bool(e.troposphere_enabled),
           "troposphere_model": fl(e.troposphere_model), "multipath_preset": e.multipath_preset,
           "multipath_enabled": bool(e.multipath_enabled), "ephemeris_source": {"type": e.ephemeris_source}}
    o = cfg.output
    outp = {"sample_rate": float(o.sample_rate), "duration_s": float(o.duration_s), "block_size": int(o.block_size),
            "seed": int(o.seed), "start_time_gps_s": float(o.start_time_gps_s), "format": o.format,
            "lpf_cutoff_hz": float(o.lpf_cutoff_hz), "output_path": o.output_path}
    return {"satellites": sats, "receiver": recv, "environment": env, "output": outp}


def dumps_config(cfg: GnssScenarioConfig) -> str:
    """`serde_yaml::to_string(&config)`: the text of the effective-config companion file / `--export-preset`.  Round-trips
    through loads_config."""
    out: list = []
    _emit(config_to_dict(cfg), 0, out)
    return "\n".join(out) + "\n"
