//! Raw declarations of `libr4w_b200.so`, one for one with `include/r4w_b200.h` (checked by `tests/test_rust_decls.py`).
//! Every fallible call returns an `r4wb_error` as `c_int`: 0-7 are r4w-ffi's `R4wError` values
//! (crates/r4w-ffi/src/lib.rs, include/r4w.h:8-25), 100 = CUDA.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const R4WB_OK: c_int = 0;
pub const R4WB_ERR_NULL_POINTER: c_int = 1;
pub const R4WB_ERR_INVALID_SIZE: c_int = 2;
pub const R4WB_ERR_BUFFER_FULL: c_int = 3;
pub const R4WB_ERR_BUFFER_EMPTY: c_int = 4;
pub const R4WB_ERR_INVALID_PARAMETER: c_int = 5;
pub const R4WB_ERR_ALLOCATION_FAILED: c_int = 6;
pub const R4WB_ERR_NOT_SUPPORTED: c_int = 7;
pub const R4WB_ERR_CUDA: c_int = 100;

pub const R4WB_MEM_HOST: c_int = 0;
pub const R4WB_MEM_DEVICE: c_int = 1;

pub const R4WB_FMT_CF32: c_int = 0;
pub const R4WB_FMT_CF64: c_int = 1;
pub const R4WB_FMT_CI16: c_int = 2;
pub const R4WB_FMT_CI8: c_int = 3;
pub const R4WB_FMT_CU8: c_int = 4;

/// declaration index of GnssSignal (gnss/types.rs:33-47)
pub const R4WB_SIG_GPS_L1CA: u32 = 0;
pub const R4WB_SIG_GPS_L5: u32 = 1;
pub const R4WB_SIG_GLONASS_L1OF: u32 = 2;
pub const R4WB_SIG_GALILEO_E1: u32 = 3;
pub const R4WB_SIG_GALILEO_E1C: u32 = 4;
pub const R4WB_SIG_GALILEO_E1OS: u32 = 5;
/// declaration index of AntennaPattern (gnss/environment/antenna.rs:11-31)
pub const R4WB_ANT_ISOTROPIC: u32 = 0;
pub const R4WB_ANT_HEMISPHERICAL: u32 = 1;
pub const R4WB_ANT_PATCH: u32 = 2;
pub const R4WB_ANT_CHOKE_RING: u32 = 3;

pub const R4WB_HAS_ELEVATION: u32 = 1 << 0;
pub const R4WB_HAS_AZIMUTH: u32 = 1 << 1;
pub const R4WB_HAS_RANGE: u32 = 1 << 2;
pub const R4WB_HAS_RANGE_RATE: u32 = 1 << 3;
pub const R4WB_HAS_DOPPLER: u32 = 1 << 4;
pub const R4WB_HAS_DOPPLER_RATE: u32 = 1 << 5;
pub const R4WB_HAS_CN0: u32 = 1 << 6;
pub const R4WB_HAS_IONO: u32 = 1 << 7;
pub const R4WB_HAS_TROPO: u32 = 1 << 8;

pub const R4WB_FLAG_NOISE_OFF: u32 = 1 << 0;
pub const R4WB_FLAG_CLOSED_FORM_PHASE: u32 = 1 << 1;

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct r4wb_lla {
    pub lat_deg: f64,
    pub lon_deg: f64,
    pub alt_m: f64,
}

/// SatelliteConfig, gnss/scenario_config.rs:137-191 (`Option<f64>` -> bit in `has` + value)
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct r4wb_sat_cfg {
    pub signal: u32,
    pub has: u32,
    pub prn: u8,
    pub plane: u8,
    pub slot: u8,
    pub nav_data: u8,
    pub orbital_dynamics: u8,
    pub _pad: [u8; 3],
    pub tx_power_dbw: f64,
    pub elevation_deg: f64,
    pub azimuth_deg: f64,
    pub range_m: f64,
    pub range_rate_mps: f64,
    pub doppler_hz: f64,
    pub doppler_rate_hz_per_s: f64,
    pub cn0_dbhz: f64,
    pub iono_delay_m: f64,
    pub tropo_delay_m: f64,
}

/// ReceiverConfig + ReceiverTrajectory, gnss/scenario_config.rs:304-315, 383-401
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct r4wb_receiver_cfg {
    pub position: r4wb_lla,
    pub antenna: u32,
    pub has_trajectory: u32,
    pub antenna_peak_gain_dbi: f64,
    pub antenna_beamwidth_deg: f64,
    pub elevation_mask_deg: f64,
    pub noise_figure_db: f64,
    pub bandwidth_hz: f64,
    pub traj_start: r4wb_lla,
    pub traj_end: r4wb_lla,
    pub traj_has_speed: u32,
    pub _pad: u32,
    pub traj_speed_mps: f64,
}

/// EnvironmentConfig, gnss/scenario_config.rs:417-437
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct r4wb_environment_cfg {
    pub ionosphere_enabled: u32,
    pub troposphere_enabled: u32,
    pub multipath_enabled: u32,
    pub multipath_preset: u32,
    pub klobuchar_alpha: [f64; 4],
    pub klobuchar_beta: [f64; 4],
    pub tropo_height_m: f64,
    pub tropo_temperature_k: f64,
    pub tropo_pressure_hpa: f64,
    pub tropo_relative_humidity: f64,
}

/// OutputConfig, gnss/scenario_config.rs:455-487
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct r4wb_output_cfg {
    pub sample_rate: f64,
    pub duration_s: f64,
    pub block_size: u64,
    pub seed: u64,
    pub start_time_gps_s: f64,
    pub lpf_cutoff_hz: f64,
}

/// GnssScenarioConfig, gnss/scenario_config.rs:537-547
#[repr(C)]
pub struct r4wb_scenario_cfg {
    pub n_sats: u32,
    pub flags: u32,
    pub sats: *const r4wb_sat_cfg,
    pub receiver: r4wb_receiver_cfg,
    pub environment: r4wb_environment_cfg,
    pub output: r4wb_output_cfg,
}

/// SatelliteStatus, gnss/satellite_emitter.rs:19-34
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct r4wb_sat_status {
    pub signal: u32,
    pub prn: u8,
    pub visible: u8,
    pub _pad: [u8; 2],
    pub elevation_deg: f64,
    pub azimuth_deg: f64,
    pub range_m: f64,
    pub range_rate_mps: f64,
    pub doppler_hz: f64,
    pub cn0_dbhz: f64,
    pub iono_delay_m: f64,
    pub tropo_delay_m: f64,
    pub antenna_gain_dbi: f64,
    pub clock_correction_s: f64,
}

/// AcquisitionResult, gnss/types.rs:168-183
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct r4wb_acq_result {
    pub prn: u8,
    pub detected: u8,
    pub has_cn0: u8,
    pub _pad: [u8; 5],
    pub code_phase: f64,
    pub doppler_hz: f64,
    pub peak_metric: f64,
    pub threshold: f64,
    pub cn0_estimate: f64,
}

/// TrackingChannel::new arguments + the bandwidth builders, gnss/tracking.rs:107-167
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct r4wb_track_cfg {
    pub sample_rate: f64,
    pub chipping_rate: f64,
    pub initial_code_phase: f64,
    pub initial_doppler: f64,
    pub dll_bandwidth_hz: f64,
    pub pll_bandwidth_hz: f64,
    pub code_length: u64,
    pub prn: u8,
    pub pad: [u8; 7],
}

/// TrackingState, gnss/types.rs:187-210
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct r4wb_track_state {
    pub code_phase: f64,
    pub carrier_freq_hz: f64,
    pub carrier_phase_rad: f64,
    pub prompt_i: f64,
    pub prompt_q: f64,
    pub cn0_dbhz: f64,
    pub ms_count: u64,
    pub prn: u8,
    pub carrier_lock: u8,
    pub code_lock: u8,
    pub bit_sync: u8,
    pub pad: [u8; 4],
}

#[repr(C)]
pub struct r4wb_scenario {
    _opaque: [u8; 0],
}
#[repr(C)]
pub struct r4wb_pcps {
    _opaque: [u8; 0],
}
#[repr(C)]
pub struct r4wb_tracker {
    _opaque: [u8; 0],
}
#[repr(C)]
pub struct r4wb_composer {
    _opaque: [u8; 0],
}

extern "C" {
    // ---- library
    pub fn r4wb_version() -> *const c_char;
    pub fn r4wb_last_error() -> *const c_char;
    pub fn r4wb_init(device: c_int) -> c_int;
    pub fn r4wb_init_devices(n_gpus: c_int) -> c_int;
    pub fn r4wb_devices_initialised() -> c_int;
    pub fn r4wb_device_count(n: *mut c_int) -> c_int;
    pub fn r4wb_set_stream(cuda_stream: *mut c_void) -> c_int;
    pub fn r4wb_host_alloc(p: *mut *mut c_void, bytes: usize) -> c_int;
    pub fn r4wb_host_free(p: *mut c_void) -> c_int;
    pub fn r4wb_kernel_launches() -> u64;

    // ---- scenario synthesis (GnssScenario, gnss/scenario.rs)
    pub fn r4wb_scenario_create(cfg: *const r4wb_scenario_cfg, out: *mut *mut r4wb_scenario) -> c_int;
    pub fn r4wb_scenario_destroy(h: *mut r4wb_scenario);
    pub fn r4wb_scenario_total_samples(h: *const r4wb_scenario) -> u64;
    pub fn r4wb_scenario_block_size(h: *const r4wb_scenario) -> u64;
    pub fn r4wb_scenario_is_done(h: *const r4wb_scenario) -> c_int;
    pub fn r4wb_scenario_progress(h: *const r4wb_scenario) -> f64;
    pub fn r4wb_scenario_reset(h: *mut r4wb_scenario) -> c_int;
    pub fn r4wb_scenario_current_sample(h: *const r4wb_scenario) -> u64;
    pub fn r4wb_scenario_generate_block(h: *mut r4wb_scenario, n: u64, dst: *mut c_void, mem: c_int, fmt: c_int, written: *mut u64) -> c_int;
    pub fn r4wb_scenario_generate_block_view(h: *mut r4wb_scenario, n: u64, fmt: c_int, block: *mut *const c_void, written: *mut u64) -> c_int;
    pub fn r4wb_scenario_generate(h: *mut r4wb_scenario, first: u64, n: u64, dst: *mut c_void, mem: c_int, fmt: c_int) -> c_int;
    pub fn r4wb_scenario_generate_rest(h: *mut r4wb_scenario, dst: *mut c_void, cap: u64, mem: c_int, fmt: c_int, written: *mut u64) -> c_int;
    pub fn r4wb_scenario_write_file(h: *mut r4wb_scenario, path: *const c_char, fmt: c_int, samples: *mut u64, bytes: *mut u64, power_sum: *mut f64) -> c_int;
    pub fn r4wb_scenario_last_power_sum(h: *const r4wb_scenario, power_sum: *mut f64) -> c_int;
    pub fn r4wb_scenario_last_path(h: *const r4wb_scenario) -> u32;
    pub fn r4wb_scenario_set_profiling(h: *mut r4wb_scenario, enabled: c_int) -> c_int;
    pub fn r4wb_scenario_last_profile(h: *mut r4wb_scenario, ms: *mut f64, launches: *mut u64) -> c_int;
    pub fn r4wb_scenario_status(h: *const r4wb_scenario, out: *mut r4wb_sat_status, cap: u32, n: *mut u32) -> c_int;

    // ---- codes (gnss/prn.rs, galileo_e1_codes.rs)
    pub fn r4wb_e1_code(channel: u32, prn: u8, out: *mut i8, cap: u64) -> c_int;
    pub fn r4wb_gps_ca_code(prn: u8, out: *mut i8, cap: u64) -> c_int;
    pub fn r4wb_gps_l5_code(prn: u8, out: *mut i8, cap: u64) -> c_int;
    pub fn r4wb_glonass_code(out: *mut i8, cap: u64) -> c_int;
    pub fn r4wb_e1c_secondary(out: *mut i8, cap: u64) -> c_int;
    pub fn r4wb_e1c_replica(prn: u8, sample_rate: f64, out: *mut i8, n: u64) -> c_int;

    // ---- PCPS acquisition (PcpsAcquisition, gnss/acquisition.rs)
    pub fn r4wb_pcps_create(code_length: u64, sample_rate: f64, out: *mut *mut r4wb_pcps) -> c_int;
    pub fn r4wb_pcps_destroy(h: *mut r4wb_pcps);
    pub fn r4wb_pcps_set_doppler_range(h: *mut r4wb_pcps, max_hz: f64, step_hz: f64) -> c_int;
    pub fn r4wb_pcps_set_threshold(h: *mut r4wb_pcps, threshold: f64) -> c_int;
    pub fn r4wb_pcps_set_coherent_periods(h: *mut r4wb_pcps, periods: u64) -> c_int;
    pub fn r4wb_pcps_fft_size(h: *const r4wb_pcps) -> u64;
    pub fn r4wb_pcps_num_doppler_bins(h: *const r4wb_pcps) -> u32;
    pub fn r4wb_pcps_acquire(h: *mut r4wb_pcps, input: *const c_void, fmt: c_int, n_input: u64, code: *const i8, code_len: u64, prn: u8, out: *mut r4wb_acq_result) -> c_int;
    pub fn r4wb_pcps_acquire_batch(h: *mut r4wb_pcps, input: *const c_void, fmt: c_int, mem: c_int, n_snapshots: u64, snapshot_stride: u64, n_input: u64, codes: *const i8, code_len: u64, prns: *const u8, n_codes: u32, out: *mut r4wb_acq_result) -> c_int;
    pub fn r4wb_pcps_acquire_grid(h: *mut r4wb_pcps, input: *const c_void, fmt: c_int, n_input: u64, code: *const i8, code_len: u64, power_out: *mut f64, cap: u64) -> c_int;
    pub fn r4wb_pcps_guard_count(h: *const r4wb_pcps) -> u64;
    pub fn r4wb_pcps_set_profiling(h: *mut r4wb_pcps, enabled: c_int) -> c_int;
    pub fn r4wb_pcps_last_profile(h: *const r4wb_pcps, ms: *mut f64, launches: *mut u64) -> c_int;

    // ---- tracking channels (TrackingChannel, gnss/tracking.rs)
    pub fn r4wb_track_create(cfgs: *const r4wb_track_cfg, n_channels: u32, out: *mut *mut r4wb_tracker) -> c_int;
    pub fn r4wb_track_destroy(h: *mut r4wb_tracker);
    pub fn r4wb_track_process(h: *mut r4wb_tracker, samples: *const c_void, fmt: c_int, mem: c_int, n_per_period: u64, n_periods: u64, channel_stride: u64, codes: *const i8, code_stride: u64, out: *mut r4wb_track_state) -> c_int;
    pub fn r4wb_track_state_get(h: *const r4wb_tracker, out: *mut r4wb_track_state, cap: u32) -> c_int;
    pub fn r4wb_track_nav_bits(h: *const r4wb_tracker, channel: u32, out: *mut i8, cap: u64, n: *mut u64) -> c_int;

    // ---- r4w-sim composer (ScenarioEngine::generate_block inner loops, crates/r4w-sim/src/scenario/engine.rs:105-135)
    pub fn r4wb_composer_create(n_emitters: u32, sample_rate: f64, noise_std: f64, seed: u64, out: *mut *mut r4wb_composer) -> c_int;
    pub fn r4wb_composer_destroy(h: *mut r4wb_composer);
    pub fn r4wb_composer_reset(h: *mut r4wb_composer) -> c_int;
    pub fn r4wb_composer_block(h: *mut r4wb_composer, baseband: *const c_void, in_fmt: c_int, in_mem: c_int, n: u64, doppler_hz: *const f64, amplitude: *const f64, active: *const u8, out: *mut c_void, out_fmt: c_int, out_mem: c_int) -> c_int;
    pub fn r4wb_composer_phases(h: *const r4wb_composer, out: *mut f64, cap: u32) -> c_int;
}
