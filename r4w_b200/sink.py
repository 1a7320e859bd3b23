"""The file-sink half of `r4w gnss scenario` (crates/r4w-cli/src/main.rs:3929-4534) over the C-ABI: resolve the sample format
and the output path, stream the scenario to disk (r4wb_scenario_write_file: device staging -> pinned buffers -> writer
thread), print the reference's summary lines (satellite table, sample/byte counts, average power) and save the effective
config as the companion `<output>.yaml`.

    python -m r4w_b200.sink --config configs/e1c_8prn_20s_clean.yaml --output /tmp/e1c.cf32 [--format ci8] [--duration 2]

Flags keep the reference's names and meaning.  The ones that trigger satellite discovery or external data in the reference
(--lat/--lon/--alt/--time/--signals, --ephemeris/--sp3/--ionex) are not part of the hot path and are not offered."""
from __future__ import annotations

import argparse
import math
import os
import sys
import time
from typing import List, Optional, Sequence

from . import _lib
from .config import GnssScenarioConfig, PRESETS, dumps_config, load_config, preset_config

_SPEED_OF_LIGHT = 299_792_458.0
# IqFormat::from_str (core/io/format.rs:137-156) and display_name (:101-109)
_FORMATS = {
    "cf64": (_lib.FMT_CF64, "cf64 (Complex Float64)", ("f64", "cf64", "cf64_le", "complex64")),
    "cf32": (_lib.FMT_CF32, "cf32 (Complex Float32)", ("f32", "cf32", "cf32_le", "ettus", "float32", "float")),
    "ci16": (_lib.FMT_CI16, "ci16 (Complex Signed Int16)", ("i16", "ci16", "ci16_le", "sc16", "int16", "short")),
    "ci8": (_lib.FMT_CI8, "ci8 (Complex Signed Int8)", ("i8", "ci8", "int8")),
    "cu8": (_lib.FMT_CU8, "cu8 (Complex Unsigned Int8)", ("u8", "cu8", "uint8", "rtlsdr")),
}
_PRESET_FLAGS = {"open-sky": "OpenSky", "urban-canyon": "UrbanCanyon", "driving": "Driving", "walking": "Walking",
                 "high-dynamics": "HighDynamics", "multi-constellation": "MultiConstellation"}
_SIGNAL_DISPLAY = {"GpsL1Ca": "GPS-L1CA", "GpsL5": "GPS-L5", "GlonassL1of": "GLONASS-L1OF", "GalileoE1": "Galileo-E1B",
                   "GalileoE1C": "Galileo-E1C", "GalileoE1OS": "Galileo-E1OS"}                  # gnss/types.rs:49-60
_CHIPPING = {"GpsL1Ca": (1.023e6, 1023), "GpsL5": (10.23e6, 10230), "GlonassL1of": (0.511e6, 511), "GalileoE1": (1.023e6, 4092),
             "GalileoE1C": (1.023e6, 4092), "GalileoE1OS": (1.023e6, 4092)}                     # gnss/types.rs:62-127


def iq_format_from_str(s: str):
    """-> (short name, r4wb_fmt, display name); ValueError with the CLI's message for an unknown name (main.rs:4398-4403)"""
    low = s.lower()
    for short, (code, display, aliases) in _FORMATS.items():
        if low in aliases:
            return short, code, display
    raise ValueError(f"Unknown format '{s}'. Options: cf32/f32/ettus, cf64/f64, sc16/ci16, ci8, cu8/rtlsdr")


def parse_prn(s: str) -> int:
    """main.rs:1094-1104: "3", "prn3", "PRN3", "prn 3" """
    t = s.strip().lower()
    num = t[3:].strip() if t.startswith("prn") else t
    if not num.isdigit() or not 0 <= int(num) <= 255:
        raise ValueError(f"Invalid PRN '{t}': expected number like '3' or 'prn3'")
    return int(num)


def expand_output_template(template: str, config: GnssScenarioConfig, format_str: str, now: Optional[time.struct_time] = None) -> str:
    """main.rs:3881-3926: {ts} {date} {time} {duration} {n_sats} {signal} {format} {sr_mhz}"""
    now = now or time.localtime()
    signal = "unknown"
    if config.satellites:
        signal = _SIGNAL_DISPLAY[config.satellites[0].signal].lower().replace("-", "").replace(" ", "")
    for key in ("e1c", "e1os", "e1", "l1ca", "l5", "l1of"):
        if key in signal:
            signal = key
            break
    dur = config.output.duration_s
    duration = f"{int(dur)}s" if dur >= 1.0 else f"{int(dur * 1000.0)}ms"
    sr = config.output.sample_rate / 1e6
    sr_str = str(int(sr)) if sr == math.floor(sr) else f"{sr:.1f}"
    return (template.replace("{ts}", time.strftime("%Y%m%d_%H%M", now)).replace("{date}", time.strftime("%Y%m%d", now))
            .replace("{time}", time.strftime("%H%M", now)).replace("{duration}", duration)
            .replace("{n_sats}", str(len(config.satellites))).replace("{signal}", signal).replace("{format}", format_str)
            .replace("{sr_mhz}", sr_str))


def status_table(statuses) -> List[str]:
    """the satellite table the CLI prints before generating (main.rs:4445-4472)"""
    lines = [f"Satellites ({len(statuses)} configured):",
             "  {:>4} {:>10} {:>7} {:>7} {:>11} {:>10} {:>10} {:>7} {:>10} {:>9} {:>8}".format(
                 "PRN", "Signal", "El(°)", "Az(°)", "Range(km)", "Rate(m/s)", "Dopp(Hz)", "C/N0", "Delay(ms)", "CodePhase", "SecEpoch")]
    for s in statuses:
        delay_s = s.range_m / _SPEED_OF_LIGHT + s.iono_delay_m / _SPEED_OF_LIGHT + s.tropo_delay_m / _SPEED_OF_LIGHT
        rate, length = _CHIPPING[s.signal]
        chips = delay_s * rate
        lines.append("  {:>4} {:>10} {:>7.2f} {:>7.2f} {:>11.1f} {:>10.2f} {:>10.1f} {:>7.1f} {:>10.3f} {:>9.1f} {:>8}".format(
            s.prn, _SIGNAL_DISPLAY[s.signal], s.elevation_deg, s.azimuth_deg, s.range_m / 1000.0, s.range_rate_mps, s.doppler_hz,
            s.cn0_dbhz, delay_s * 1000.0, math.fmod(chips, length), int(chips / length) % 25))
    return lines


def write_scenario(config: GnssScenarioConfig, output, format_str: Optional[str] = None, log=print):
    """Generate `config` into `output` and write the companion YAML next to it.  -> dict(samples, bytes, avg_power_db,
    config_path).  `format_str` None = the config's `output.format` (CLI flag > YAML > cf32, main.rs:4397)."""
    from .scenario import GnssScenario
    format_str = format_str or config.output.format
    short, code, display = iq_format_from_str(format_str)
    scen = GnssScenario(config)
    try:
        for line in status_table(scen.satellite_status()):
            log(line)
        log("")
        log(f"Generating {scen.total_samples()} samples...")
        samples, nbytes, power_sum = scen.write_file(output, code)
    finally:
        scen.close()
    log(f"Generated {samples} IQ samples")
    log(f"Written {nbytes} bytes to {output} ({display})")
    avg = power_sum / samples if samples > 0 else 0.0
    avg_db = 10.0 * math.log10(avg) if avg > 0.0 else float("-inf")
    log(f"Average power: {avg_db:.2f} dB")
    effective = config.copy()
    effective.output.format = format_str
    config_path = os.path.splitext(str(output))[0] + ".yaml"            # PathBuf::with_extension("yaml")
    try:
        with open(config_path, "w") as f:
            f.write(dumps_config(effective))
        log(f"Config:        {config_path}")
    except OSError as e:
        print(f"Warning: Failed to write config file '{config_path}': {e}", file=sys.stderr)
    return {"samples": samples, "bytes": nbytes, "avg_power_db": avg_db, "config_path": config_path}


def build_config(args) -> GnssScenarioConfig:
    """the CLI's override order (main.rs:4107-4134, 4270-4280, 4415-4419)"""
    if args.preset is not None and args.preset not in _PRESET_FLAGS:
        raise ValueError(f"Unknown preset '{args.preset}'. Use --list-presets to see options.")
    config = load_config(args.config) if args.config else preset_config(_PRESET_FLAGS[args.preset or "open-sky"])
    if args.duration is not None:
        config.output.duration_s = float(args.duration)
    if args.sample_rate is not None:
        config.output.sample_rate = float(args.sample_rate)
    config.receiver.elevation_mask_deg = float(args.elevation_mask)
    if args.limit_prns:
        prns = [parse_prn(p) for p in args.limit_prns.split(",")]
        config.satellites = [s for s in config.satellites if s.prn in prns]
        if not config.satellites:
            raise ValueError(f"No satellites match the specified PRNs: {args.limit_prns.split(',')}")
    if args.lpf_cutoff is not None:
        config.output.lpf_cutoff_hz = float(args.lpf_cutoff)
    return config


def parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(prog="python -m r4w_b200.sink", description="GNSS scenario -> IQ file (r4w gnss scenario)")
    p.add_argument("-P", "--preset")
    p.add_argument("-c", "--config")
    p.add_argument("-o", "--output")
    p.add_argument("-d", "--duration", type=float)
    p.add_argument("--sample-rate", type=float)
    p.add_argument("--list-presets", action="store_true")
    p.add_argument("--export-preset", nargs="?", const="-")
    p.add_argument("--format")
    p.add_argument("--elevation-mask", type=float, default=5.0)
    p.add_argument("--lpf-cutoff", type=float)
    p.add_argument("--limit-prns")
    return p


def main(argv: Optional[Sequence[str]] = None) -> int:
    args = parser().parse_args(argv)
    if args.list_presets:
        print("Available GNSS scenario presets:\n")
        for flag, name in _PRESET_FLAGS.items():
            cfg = preset_config(name)
            print(f"  {flag:20} - {len(cfg.satellites)} satellites, {cfg.environment.multipath_preset} environment")
        return 0
    try:
        config = build_config(args)
        if args.export_preset is not None:
            text = dumps_config(config)
            if args.export_preset == "-":
                sys.stdout.write(text)
            else:
                with open(args.export_preset, "w") as f:
                    f.write(text)
                print(f"Wrote config to {args.export_preset}", file=sys.stderr)
            return 0
        format_str = args.format or config.output.format
        _, _, display = iq_format_from_str(format_str)
        output = args.output or expand_output_template(
            config.output.output_path or "{ts}_{signal}_{n_sats}prn_{duration}.sigmf-data", config, format_str)
        pos = config.receiver.position
        print("GNSS Scenario Generator\n=======================")
        print(f"Config:      {args.config or (args.preset or 'open-sky')}")
        print(f"Receiver:    {pos.lat_deg:.4f}°N, {-pos.lon_deg:.4f}°W, {pos.alt_m:.0f}m")
        print(f"Start time:  GPS {config.output.start_time_gps_s:.0f} s")
        print(f"Duration:    {config.output.duration_s * 1000.0} ms")
        print(f"Sample rate: {config.output.sample_rate / 1e6} MHz")
        print(f"Format:      {display}")
        print(f"Output:      {output}\n")
        write_scenario(config, output, format_str)
    except (ValueError, OSError, RuntimeError) as e:
        print(f"Error: {e}", file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
