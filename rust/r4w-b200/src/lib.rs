//! Source-compatible replacements for the three hot-path types of `r4w_core::waveform::gnss`:
//!   GnssScenario     gnss/scenario.rs:51-674      -> r4wb_scenario_*
//!   PcpsAcquisition  gnss/acquisition.rs:40-254   -> r4wb_pcps_*
//!   TrackingChannel  gnss/tracking.rs:107-358     -> r4wb_track_*
//! Same constructors, builders, methods and result types; the arithmetic runs on the GPU (sm_100a).  As in the
//! reference nothing here returns an error: a failure of the library panics with its message, like the `assert!`s of
//! gnss/prn.rs:283.  NOT COMPILED in the build image of this repository (no Rust toolchain) — see rust/README.md.
use std::ffi::{CStr, CString};
use std::os::raw::c_int;

use r4w_b200_sys as sys;
use r4w_core::coordinates::LlaPosition;
use r4w_core::types::IQSample;
use r4w_core::waveform::gnss::environment::antenna::AntennaPattern;
use r4w_core::waveform::gnss::environment::ionosphere::KlobucharModel;
use r4w_core::waveform::gnss::environment::multipath::GnssMultipathPreset;
use r4w_core::waveform::gnss::environment::troposphere::SaastamoinenModel;
use r4w_core::waveform::gnss::satellite_emitter::SatelliteStatus;
use r4w_core::waveform::gnss::scenario_config::{
    EnvironmentConfig, GnssScenarioConfig, GnssScenarioPreset, IonosphereSource, OutputConfig, ReceiverConfig, SatelliteConfig,
};
use r4w_core::waveform::gnss::types::{AcquisitionResult, GnssSignal, TrackingState};

fn check(code: c_int) {
    if code != sys::R4WB_OK {
        let msg = unsafe { CStr::from_ptr(sys::r4wb_last_error()) }.to_string_lossy().into_owned();
        panic!("libr4w_b200: error {code}: {msg}");
    }
}

/// Select the GPU of this process (one process per GPU); -1 keeps the current device.
pub fn init(device: i32) {
    check(unsafe { sys::r4wb_init(device) });
}

fn lla(p: &LlaPosition) -> sys::r4wb_lla {
    sys::r4wb_lla { lat_deg: p.lat_deg, lon_deg: p.lon_deg, alt_m: p.alt_m }
}

fn signal_index(s: GnssSignal) -> u32 {
    match s {
        GnssSignal::GpsL1Ca => sys::R4WB_SIG_GPS_L1CA,
        GnssSignal::GpsL5 => sys::R4WB_SIG_GPS_L5,
        GnssSignal::GlonassL1of => sys::R4WB_SIG_GLONASS_L1OF,
        GnssSignal::GalileoE1 => sys::R4WB_SIG_GALILEO_E1,
        GnssSignal::GalileoE1C => sys::R4WB_SIG_GALILEO_E1C,
        GnssSignal::GalileoE1OS => sys::R4WB_SIG_GALILEO_E1OS,
    }
}

fn signal_from_index(i: u32) -> GnssSignal {
    match i {
        sys::R4WB_SIG_GPS_L1CA => GnssSignal::GpsL1Ca,
        sys::R4WB_SIG_GPS_L5 => GnssSignal::GpsL5,
        sys::R4WB_SIG_GLONASS_L1OF => GnssSignal::GlonassL1of,
        sys::R4WB_SIG_GALILEO_E1 => GnssSignal::GalileoE1,
        sys::R4WB_SIG_GALILEO_E1C => GnssSignal::GalileoE1C,
        _ => GnssSignal::GalileoE1OS,
    }
}

fn sat_to_pod(s: &SatelliteConfig) -> sys::r4wb_sat_cfg {
    let mut p = sys::r4wb_sat_cfg {
        signal: signal_index(s.signal),
        prn: s.prn,
        plane: s.plane,
        slot: s.slot,
        nav_data: s.nav_data as u8,
        orbital_dynamics: s.orbital_dynamics as u8,
        tx_power_dbw: s.tx_power_dbw,
        ..Default::default()
    };
    let mut opt = |v: Option<f64>, bit: u32, dst: &mut f64| {
        if let Some(x) = v {
            p.has |= bit;
            *dst = x;
        }
    };
    let (mut el, mut az, mut rg, mut rr, mut dp, mut dr, mut cn, mut io, mut tr) = (0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0);
    opt(s.elevation_deg, sys::R4WB_HAS_ELEVATION, &mut el);
    opt(s.azimuth_deg, sys::R4WB_HAS_AZIMUTH, &mut az);
    opt(s.range_m, sys::R4WB_HAS_RANGE, &mut rg);
    opt(s.range_rate_mps, sys::R4WB_HAS_RANGE_RATE, &mut rr);
    opt(s.doppler_hz, sys::R4WB_HAS_DOPPLER, &mut dp);
    opt(s.doppler_rate_hz_per_s, sys::R4WB_HAS_DOPPLER_RATE, &mut dr);
    opt(s.cn0_dbhz, sys::R4WB_HAS_CN0, &mut cn);
    opt(s.iono_delay_m, sys::R4WB_HAS_IONO, &mut io);
    opt(s.tropo_delay_m, sys::R4WB_HAS_TROPO, &mut tr);
    drop(opt);
    p.elevation_deg = el;
    p.azimuth_deg = az;
    p.range_m = rg;
    p.range_rate_mps = rr;
    p.doppler_hz = dp;
    p.doppler_rate_hz_per_s = dr;
    p.cn0_dbhz = cn;
    p.iono_delay_m = io;
    p.tropo_delay_m = tr;
    p
}

fn rx_to_pod(r: &ReceiverConfig) -> sys::r4wb_receiver_cfg {
    let (antenna, gain, beam) = match r.antenna {
        AntennaPattern::Isotropic => (sys::R4WB_ANT_ISOTROPIC, 0.0, 0.0),
        AntennaPattern::Hemispherical { peak_gain_dbi } => (sys::R4WB_ANT_HEMISPHERICAL, peak_gain_dbi, 0.0),
        AntennaPattern::Patch { peak_gain_dbi, beamwidth_deg } => (sys::R4WB_ANT_PATCH, peak_gain_dbi, beamwidth_deg),
        AntennaPattern::ChokeRing { peak_gain_dbi } => (sys::R4WB_ANT_CHOKE_RING, peak_gain_dbi, 0.0),
    };
    let mut p = sys::r4wb_receiver_cfg {
        position: lla(&r.position),
        antenna,
        antenna_peak_gain_dbi: gain,
        antenna_beamwidth_deg: beam,
        elevation_mask_deg: r.elevation_mask_deg,
        noise_figure_db: r.noise_figure_db,
        bandwidth_hz: r.bandwidth_hz,
        ..Default::default()
    };
    if let Some(t) = &r.trajectory {
        p.has_trajectory = 1;
        p.traj_start = lla(&t.start);
        p.traj_end = lla(&t.end);
        if let Some(v) = t.speed_mps {
            p.traj_has_speed = 1;
            p.traj_speed_mps = v;
        }
    }
    p
}

fn env_to_pod(e: &EnvironmentConfig) -> sys::r4wb_environment_cfg {
    // scenario.rs:141-150: a null model means the broadcast default / the standard atmosphere
    let k = e.ionosphere_model.clone().unwrap_or_else(KlobucharModel::default_broadcast);
    let t = e.troposphere_model.clone().unwrap_or_else(SaastamoinenModel::standard_atmosphere);
    sys::r4wb_environment_cfg {
        ionosphere_enabled: (e.ionosphere_enabled && !matches!(e.ionosphere_source, IonosphereSource::Disabled)) as u32,
        troposphere_enabled: e.troposphere_enabled as u32,
        multipath_enabled: e.multipath_enabled as u32,
        multipath_preset: match e.multipath_preset {
            GnssMultipathPreset::OpenSky => 0,
            GnssMultipathPreset::Suburban => 1,
            GnssMultipathPreset::UrbanCanyon => 2,
            GnssMultipathPreset::Indoor => 3,
        },
        klobuchar_alpha: k.alpha,
        klobuchar_beta: k.beta,
        tropo_height_m: t.height_m,
        tropo_temperature_k: t.temperature_k,
        tropo_pressure_hpa: t.pressure_hpa,
        tropo_relative_humidity: t.relative_humidity,
    }
}

fn out_to_pod(o: &OutputConfig) -> sys::r4wb_output_cfg {
    sys::r4wb_output_cfg {
        sample_rate: o.sample_rate,
        duration_s: o.duration_s,
        block_size: o.block_size as u64,
        seed: o.seed,
        start_time_gps_s: o.start_time_gps_s,
        lpf_cutoff_hz: o.lpf_cutoff_hz,
    }
}

// ------------------------------------------------------------------------------------------------ GnssScenario
/// gnss/scenario.rs:51-74.  `&mut self` methods: one thread at a time, like the reference.
pub struct GnssScenario {
    h: *mut sys::r4wb_scenario,
    config: GnssScenarioConfig,
}

impl GnssScenario {
    /// scenario.rs:78
    pub fn new(config: GnssScenarioConfig) -> Self {
        let sats: Vec<sys::r4wb_sat_cfg> = config.satellites.iter().map(sat_to_pod).collect();
        let pod = sys::r4wb_scenario_cfg {
            n_sats: sats.len() as u32,
            flags: 0,
            sats: sats.as_ptr(),
            receiver: rx_to_pod(&config.receiver),
            environment: env_to_pod(&config.environment),
            output: out_to_pod(&config.output),
        };
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::r4wb_scenario_create(&pod, &mut h) });
        Self { h, config }
    }

    /// scenario.rs:240
    pub fn from_preset(preset: GnssScenarioPreset) -> Self {
        Self::new(preset.to_config())
    }

    /// scenario.rs:308: one reference block at the current position; empty when done
    pub fn generate_block(&mut self, block_size: usize) -> Vec<IQSample> {
        let mut out = vec![IQSample::new(0.0, 0.0); block_size];
        let mut written = 0u64;
        check(unsafe {
            sys::r4wb_scenario_generate_block(self.h, block_size as u64, out.as_mut_ptr().cast(), sys::R4WB_MEM_HOST, sys::R4WB_FMT_CF64, &mut written)
        });
        out.truncate(written as usize);
        out
    }

    /// The sink form of `generate_block`: the block as the interleaved f32 pairs the CLI's cf32 writer emits
    /// (main.rs:4488-4500, core/io/format.rs:197-200), borrowed from the library's pinned render-ahead ring: no host
    /// copy.  The slice lives until the next call on this scenario (the `&mut self` borrow enforces it).
    pub fn generate_block_cf32(&mut self, block_size: usize) -> &[[f32; 2]] {
        let mut written = 0u64;
        let mut block: *const std::ffi::c_void = std::ptr::null();
        check(unsafe { sys::r4wb_scenario_generate_block_view(self.h, block_size as u64, sys::R4WB_FMT_CF32, &mut block, &mut written) });
        if written == 0 {
            return &[];
        }
        unsafe { std::slice::from_raw_parts(block.cast::<[f32; 2]>(), written as usize) }
    }

    /// scenario.rs:549: everything from the current position to the end (the CLI loop's concatenation), rendered in one
    /// call; leaves the scenario done like the reference's `while !self.is_done()` loop
    pub fn generate(&mut self) -> Vec<IQSample> {
        let n = self.total_samples() - unsafe { sys::r4wb_scenario_current_sample(self.h) } as usize;
        let mut out = vec![IQSample::new(0.0, 0.0); n];
        let mut written = 0u64;
        check(unsafe {
            sys::r4wb_scenario_generate_rest(self.h, out.as_mut_ptr().cast(), n as u64, sys::R4WB_MEM_HOST, sys::R4WB_FMT_CF64, &mut written)
        });
        out.truncate(written as usize);
        out
    }

    /// The CLI's sink (main.rs:4483-4509): renders [0, total_samples) straight into `path` in `format`
    /// (any alias IqFormat::from_str accepts); returns (samples, bytes, sum |s|^2).
    pub fn write_file(&mut self, path: &std::path::Path, format: &str) -> (u64, u64, f64) {
        let fmt = match format.to_lowercase().as_str() {
            // IqFormat::from_str, core/io/format.rs:137-156
            "f64" | "cf64" | "cf64_le" | "complex64" => sys::R4WB_FMT_CF64,
            "f32" | "cf32" | "cf32_le" | "ettus" | "float32" | "float" => sys::R4WB_FMT_CF32,
            "i16" | "ci16" | "ci16_le" | "sc16" | "int16" | "short" => sys::R4WB_FMT_CI16,
            "i8" | "ci8" | "int8" => sys::R4WB_FMT_CI8,
            "u8" | "cu8" | "uint8" | "rtlsdr" => sys::R4WB_FMT_CU8,
            other => panic!("unknown IQ format {other:?}"),
        };
        let c = CString::new(path.to_string_lossy().as_bytes()).expect("path contains NUL");
        let (mut samples, mut bytes, mut power) = (0u64, 0u64, 0f64);
        check(unsafe { sys::r4wb_scenario_write_file(self.h, c.as_ptr(), fmt, &mut samples, &mut bytes, &mut power) });
        (samples, bytes, power)
    }

    /// scenario.rs:564
    pub fn satellite_status(&self) -> Vec<SatelliteStatus> {
        let mut raw = vec![sys::r4wb_sat_status::default(); self.config.satellites.len().max(1) * 2];   // E1OS renders as two rows
        let mut n = 0u32;
        check(unsafe { sys::r4wb_scenario_status(self.h, raw.as_mut_ptr(), raw.len() as u32, &mut n) });
        raw.truncate(n as usize);
        raw.iter()
            .map(|s| SatelliteStatus {
                prn: s.prn,
                signal: signal_from_index(s.signal),
                elevation_deg: s.elevation_deg,
                azimuth_deg: s.azimuth_deg,
                range_m: s.range_m,
                range_rate_mps: s.range_rate_mps,
                doppler_hz: s.doppler_hz,
                cn0_dbhz: s.cn0_dbhz,
                iono_delay_m: s.iono_delay_m,
                tropo_delay_m: s.tropo_delay_m,
                antenna_gain_dbi: s.antenna_gain_dbi,
                visible: s.visible != 0,
                clock_correction_s: s.clock_correction_s,
            })
            .collect()
    }

    pub fn reset(&mut self) {
        check(unsafe { sys::r4wb_scenario_reset(self.h) })
    }
    pub fn is_done(&self) -> bool {
        unsafe { sys::r4wb_scenario_is_done(self.h) != 0 }
    }
    pub fn progress(&self) -> f64 {
        unsafe { sys::r4wb_scenario_progress(self.h) }
    }
    pub fn config(&self) -> &GnssScenarioConfig {
        &self.config
    }
    pub fn total_samples(&self) -> usize {
        unsafe { sys::r4wb_scenario_total_samples(self.h) as usize }
    }
    pub fn block_size(&self) -> usize {
        unsafe { sys::r4wb_scenario_block_size(self.h) as usize }
    }
}

impl Drop for GnssScenario {
    fn drop(&mut self) {
        unsafe { sys::r4wb_scenario_destroy(self.h) }
    }
}

// ------------------------------------------------------------------------------------------------ PcpsAcquisition
/// gnss/acquisition.rs:40-55.  `acquire(&self)` is re-entrant in the reference; the library lets threads share a
/// configured handle (searches take turns on its device scratch), hence Send + Sync.
pub struct PcpsAcquisition {
    h: *mut sys::r4wb_pcps,
    code_length: usize,
    doppler_max_hz: f64,
    doppler_step_hz: f64,
}
unsafe impl Send for PcpsAcquisition {}
unsafe impl Sync for PcpsAcquisition {}

/// gnss/acquisition.rs:257-286: the reference's own fields
#[derive(Debug, Clone)]
pub struct AcquisitionGrid {
    /// Doppler frequency bins (Hz): `-doppler_max + d * doppler_step`
    pub doppler_bins: Vec<f64>,
    /// Code phase bins (samples, `0 .. code_length`)
    pub code_phases: Vec<f64>,
    /// Correlation power [doppler][code_phase]
    pub correlation_power: Vec<Vec<f64>>,
}

impl AcquisitionGrid {
    /// acquisition.rs:270-285: first strict maximum in scan order -> (doppler_hz, code_phase, power)
    pub fn find_peak(&self) -> (f64, f64, f64) {
        let (mut best_power, mut best_doppler, mut best_phase) = (0.0_f64, 0.0, 0.0);
        for (d, row) in self.correlation_power.iter().enumerate() {
            for (p, &power) in row.iter().enumerate() {
                if power > best_power {
                    best_power = power;
                    best_doppler = self.doppler_bins[d];
                    best_phase = self.code_phases[p];
                }
            }
        }
        (best_doppler, best_phase, best_power)
    }
}

impl PcpsAcquisition {
    /// acquisition.rs:63 (defaults +-5 kHz / 500 Hz, threshold 2.5, one coherent period)
    pub fn new(code_length: usize, sample_rate: f64) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::r4wb_pcps_create(code_length as u64, sample_rate, &mut h) });
        Self { h, code_length, doppler_max_hz: 5000.0, doppler_step_hz: 500.0 }
    }
    pub fn with_doppler_range(mut self, max_hz: f64, step_hz: f64) -> Self {
        check(unsafe { sys::r4wb_pcps_set_doppler_range(self.h, max_hz, step_hz) });
        self.doppler_max_hz = max_hz;
        self.doppler_step_hz = step_hz;
        self
    }
    pub fn with_threshold(self, threshold: f64) -> Self {
        check(unsafe { sys::r4wb_pcps_set_threshold(self.h, threshold) });
        self
    }
    pub fn with_coherent_periods(self, periods: usize) -> Self {
        check(unsafe { sys::r4wb_pcps_set_coherent_periods(self.h, periods as u64) });
        self
    }
    pub fn fft_size(&self) -> usize {
        unsafe { sys::r4wb_pcps_fft_size(self.h) as usize }
    }

    /// acquisition.rs:104
    pub fn acquire(&self, input: &[IQSample], code: &[i8], prn: u8) -> AcquisitionResult {
        let mut r = sys::r4wb_acq_result::default();
        check(unsafe {
            sys::r4wb_pcps_acquire(self.h, input.as_ptr().cast(), sys::R4WB_FMT_CF64, input.len() as u64, code.as_ptr(), code.len() as u64, prn, &mut r)
        });
        result_from_pod(&r)
    }

    /// Many snapshots x many replicas in one call (no counterpart in the reference: its callers loop over `acquire`).
    /// `input` holds `n_snapshots` windows of `n_input` samples, `stride` samples apart; result is [snapshot][code].
    pub fn acquire_batch(&self, input: &[IQSample], n_snapshots: usize, stride: usize, n_input: usize, codes: &[&[i8]], prns: &[u8]) -> Vec<Vec<AcquisitionResult>> {
        assert!(codes.len() == prns.len() && !codes.is_empty());
        assert!(n_snapshots == 0 || (n_snapshots - 1) * stride + n_input <= input.len());
        let code_len = codes[0].len();
        let flat: Vec<i8> = codes.iter().flat_map(|c| { assert_eq!(c.len(), code_len); c.iter().copied() }).collect();
        let mut raw = vec![sys::r4wb_acq_result::default(); n_snapshots * codes.len()];
        check(unsafe {
            sys::r4wb_pcps_acquire_batch(self.h, input.as_ptr().cast(), sys::R4WB_FMT_CF64, sys::R4WB_MEM_HOST, n_snapshots as u64, stride as u64,
                                         n_input as u64, flat.as_ptr(), code_len as u64, prns.as_ptr(), codes.len() as u32, raw.as_mut_ptr())
        });
        raw.chunks(codes.len()).map(|row| row.iter().map(result_from_pod).collect()).collect()
    }

    /// acquisition.rs:199
    pub fn acquire_grid(&self, input: &[IQSample], code: &[i8]) -> AcquisitionGrid {
        let bins = unsafe { sys::r4wb_pcps_num_doppler_bins(self.h) } as usize;
        let mut flat = vec![0f64; bins * self.code_length];
        check(unsafe {
            sys::r4wb_pcps_acquire_grid(self.h, input.as_ptr().cast(), sys::R4WB_FMT_CF64, input.len() as u64, code.as_ptr(), code.len() as u64,
                                        flat.as_mut_ptr(), flat.len() as u64)
        });
        AcquisitionGrid {
            doppler_bins: (0..bins).map(|d| -self.doppler_max_hz + d as f64 * self.doppler_step_hz).collect(),      // acquisition.rs:213-220
            code_phases: (0..self.code_length).map(|i| i as f64).collect(),
            correlation_power: flat.chunks(self.code_length).map(|r| r.to_vec()).collect(),
        }
    }
}

fn result_from_pod(r: &sys::r4wb_acq_result) -> AcquisitionResult {
    AcquisitionResult {
        prn: r.prn,
        detected: r.detected != 0,
        code_phase: r.code_phase,
        doppler_hz: r.doppler_hz,
        peak_metric: r.peak_metric,
        threshold: r.threshold,
        cn0_estimate: (r.has_cn0 != 0).then_some(r.cn0_estimate),
    }
}

impl Drop for PcpsAcquisition {
    fn drop(&mut self) {
        unsafe { sys::r4wb_pcps_destroy(self.h) }
    }
}

// ------------------------------------------------------------------------------------------------ TrackingChannel
/// gnss/tracking.rs:107-358: one channel (a bank of one).  A receiver with many satellites should build ONE bank
/// through `sys::r4wb_track_create` with a config per channel and hand it whole seconds of samples.
pub struct TrackingChannel {
    h: *mut sys::r4wb_tracker,
    cfg: sys::r4wb_track_cfg,
}

impl TrackingChannel {
    /// tracking.rs:113
    pub fn new(prn: u8, code_length: usize, sample_rate: f64, chipping_rate: f64, initial_code_phase: f64, initial_doppler: f64) -> Self {
        let cfg = sys::r4wb_track_cfg {
            sample_rate,
            chipping_rate,
            initial_code_phase,
            initial_doppler,
            dll_bandwidth_hz: 0.0, // <= 0: the reference defaults (1 Hz DLL, 15 Hz PLL)
            pll_bandwidth_hz: 0.0,
            code_length: code_length as u64,
            prn,
            pad: [0; 7],
        };
        Self::build(cfg)
    }
    fn build(cfg: sys::r4wb_track_cfg) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::r4wb_track_create(&cfg, 1, &mut h) });
        Self { h, cfg }
    }
    /// tracking.rs:156 (builders run before the first `process`, so the channel is simply rebuilt).  `bw_hz` must be positive:
    /// the C-ABI reads a bandwidth <= 0 as "the reference default" (1 Hz DLL / 15 Hz PLL), it does not build a zero-gain loop.
    pub fn with_dll_bandwidth(self, bw_hz: f64) -> Self {
        assert!(bw_hz > 0.0, "loop bandwidth must be positive");
        let mut cfg = self.cfg;
        cfg.dll_bandwidth_hz = bw_hz;
        Self::build(cfg)
    }
    /// tracking.rs:163 (same rule)
    pub fn with_pll_bandwidth(self, bw_hz: f64) -> Self {
        assert!(bw_hz > 0.0, "loop bandwidth must be positive");
        let mut cfg = self.cfg;
        cfg.pll_bandwidth_hz = bw_hz;
        Self::build(cfg)
    }
    /// tracking.rs:177: one code period per call
    pub fn process(&mut self, samples: &[IQSample], code: &[i8]) -> TrackingState {
        let mut s = sys::r4wb_track_state::default();
        check(unsafe {
            sys::r4wb_track_process(self.h, samples.as_ptr().cast(), sys::R4WB_FMT_CF64, sys::R4WB_MEM_HOST, samples.len() as u64, 1, 0,
                                    code.as_ptr(), code.len() as u64, &mut s)
        });
        state_from_pod(&s)
    }
    /// tracking.rs:340 (owned: the bits live on the device)
    pub fn nav_bits(&self) -> Vec<i8> {
        let mut n = 0u64;
        check(unsafe { sys::r4wb_track_nav_bits(self.h, 0, std::ptr::null_mut(), 0, &mut n) });
        let mut out = vec![0i8; n as usize];
        if n > 0 {
            check(unsafe { sys::r4wb_track_nav_bits(self.h, 0, out.as_mut_ptr(), n, &mut n) });
        }
        out
    }
    /// tracking.rs:345
    pub fn state(&self) -> TrackingState {
        let mut s = sys::r4wb_track_state::default();
        check(unsafe { sys::r4wb_track_state_get(self.h, &mut s, 1) });
        state_from_pod(&s)
    }
}

fn state_from_pod(s: &sys::r4wb_track_state) -> TrackingState {
    TrackingState {
        prn: s.prn,
        code_phase: s.code_phase,
        carrier_freq_hz: s.carrier_freq_hz,
        carrier_phase_rad: s.carrier_phase_rad,
        prompt_i: s.prompt_i,
        prompt_q: s.prompt_q,
        cn0_dbhz: s.cn0_dbhz,
        carrier_lock: s.carrier_lock != 0,
        code_lock: s.code_lock != 0,
        bit_sync: s.bit_sync != 0,
        ms_count: s.ms_count,
    }
}

impl Drop for TrackingChannel {
    fn drop(&mut self) {
        unsafe { sys::r4wb_track_destroy(self.h) }
    }
}

// ------------------------------------------------------------------------------------------------ r4w-sim composer
/// The per-sample loops of `ScenarioEngine::generate_block` (crates/r4w-sim/src/scenario/engine.rs:105-135): Doppler rotation
/// with the continuously accumulated `carrier_phases`, amplitude, sum over the emitters and receiver noise.  The engine keeps
/// evaluating the block geometry (`EmitterState`, engine.rs:68-101) on the host and hands each block's per-emitter Doppler,
/// amplitude and baseband here; `reset()` mirrors `ScenarioEngine::reset` (engine.rs:149-153).
pub struct Composer {
    h: *mut sys::r4wb_composer,
    n_emitters: usize,
}

impl Composer {
    /// `noise_std = sqrt(ScenarioConfig::noise_power_linear() / 2)` (engine.rs:126-127), `seed = ScenarioConfig::seed`
    pub fn new(n_emitters: usize, sample_rate: f64, noise_std: f64, seed: u64) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::r4wb_composer_create(n_emitters as u32, sample_rate, noise_std, seed, &mut h) });
        Self { h, n_emitters }
    }

    /// One block: `baseband[e]` = `Emitter::generate_iq` of emitter `e` (all the same length), `active[e]` = the emitter passed the
    /// engine's visibility test.  Returns the composite block; the carrier phases advance as engine.rs:108-113.
    pub fn block(&mut self, baseband: &[Vec<IQSample>], doppler_hz: &[f64], amplitude: &[f64], active: &[bool]) -> Vec<IQSample> {
        assert!(baseband.len() == self.n_emitters && doppler_hz.len() == self.n_emitters && amplitude.len() == self.n_emitters && active.len() == self.n_emitters);
        let n = baseband.first().map_or(0, |b| b.len());
        let flat: Vec<IQSample> = baseband.iter().flat_map(|b| { assert_eq!(b.len(), n); b.iter().copied() }).collect();
        let act: Vec<u8> = active.iter().map(|&a| a as u8).collect();
        let mut out = vec![IQSample::new(0.0, 0.0); n];
        check(unsafe {
            sys::r4wb_composer_block(self.h, flat.as_ptr().cast(), sys::R4WB_FMT_CF64, sys::R4WB_MEM_HOST, n as u64, doppler_hz.as_ptr(), amplitude.as_ptr(),
                                     act.as_ptr(), out.as_mut_ptr().cast(), sys::R4WB_FMT_CF64, sys::R4WB_MEM_HOST)
        });
        out
    }

    /// `ScenarioEngine::carrier_phases` (radians) after the last block
    pub fn carrier_phases(&self) -> Vec<f64> {
        let mut out = vec![0f64; self.n_emitters];
        check(unsafe { sys::r4wb_composer_phases(self.h, out.as_mut_ptr(), out.len() as u32) });
        out
    }

    pub fn reset(&mut self) {
        check(unsafe { sys::r4wb_composer_reset(self.h) })
    }
}

impl Drop for Composer {
    fn drop(&mut self) {
        unsafe { sys::r4wb_composer_destroy(self.h) }
    }
}
