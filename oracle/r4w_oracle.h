/*
 * r4w_oracle.h — CPU restatement (f64, plain C) of r4w's GNSS hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and only
 * as the checker or the timed CPU baseline.  libr4w_b200.so never links or calls it.
 *
 * The reference (Rust) cannot be compiled in this image (no cargo/rustc), so this file restates the
 * algorithm line by line; each function cites the reference file:line it follows (paths relative to
 * /root/reference; gnss/ = crates/r4w-core/src/waveform/gnss/, core/ = crates/r4w-core/src/).
 * Pinning: the reference's own known-answer tests are reproduced in
 * tests/test_oracle_kats.py.  The FFT inside PCPS lives in the un-vendored crate rustfft 6.4.1
 * (Cargo.lock:5249-5252); its rounding is parity-unpinned, any correct f64 FFT agrees to ~1e-13.
 *
 * The scenario config structs are layout-identical to include/r4w_b200.h's r4wb_* PODs so one
 * host-side (ctypes) structure feeds both the product and the oracle.
 */
#ifndef R4W_ORACLE_H
#define R4W_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_c64 { double re, im; } orc_c64;
typedef struct orc_lla { double lat_deg, lon_deg, alt_m; } orc_lla;

enum { ORC_HAS_ELEVATION = 1, ORC_HAS_AZIMUTH = 2, ORC_HAS_RANGE = 4, ORC_HAS_RANGE_RATE = 8,
       ORC_HAS_DOPPLER = 16, ORC_HAS_DOPPLER_RATE = 32, ORC_HAS_CN0 = 64, ORC_HAS_IONO = 128,
       ORC_HAS_TROPO = 256 };
enum { ORC_FLAG_NOISE_OFF = 1 };

typedef struct orc_sat_cfg {
    uint32_t signal, has;
    uint8_t prn, plane, slot, nav_data, orbital_dynamics, _pad[3];
    double tx_power_dbw;
    double elevation_deg, azimuth_deg, range_m, range_rate_mps;
    double doppler_hz, doppler_rate_hz_per_s, cn0_dbhz, iono_delay_m, tropo_delay_m;
} orc_sat_cfg;

typedef struct orc_receiver_cfg {
    orc_lla position;
    uint32_t antenna, has_trajectory;
    double antenna_peak_gain_dbi, antenna_beamwidth_deg;
    double elevation_mask_deg, noise_figure_db, bandwidth_hz;
    orc_lla traj_start, traj_end;
    uint32_t traj_has_speed, _pad;
    double traj_speed_mps;
} orc_receiver_cfg;

typedef struct orc_environment_cfg {
    uint32_t ionosphere_enabled, troposphere_enabled, multipath_enabled, multipath_preset;
    double klobuchar_alpha[4], klobuchar_beta[4];   /* KlobucharModel, environment/ionosphere.rs:18-33 */
    double tropo_height_m, tropo_temperature_k, tropo_pressure_hpa, tropo_relative_humidity;   /* SaastamoinenModel, troposphere.rs:14-36 */
} orc_environment_cfg;

typedef struct orc_output_cfg {
    double sample_rate, duration_s;
    uint64_t block_size, seed;
    double start_time_gps_s, lpf_cutoff_hz;
} orc_output_cfg;

typedef struct orc_scenario_cfg {
    uint32_t n_sats, flags;
    const orc_sat_cfg* sats;
    orc_receiver_cfg receiver;
    orc_environment_cfg environment;
    orc_output_cfg output;
} orc_scenario_cfg;

typedef struct orc_sat_status {
    uint32_t signal;
    uint8_t prn, visible, _pad[2];
    double elevation_deg, azimuth_deg, range_m, range_rate_mps, doppler_hz, cn0_dbhz;
    double iono_delay_m, tropo_delay_m, antenna_gain_dbi, clock_correction_s;
} orc_sat_status;

typedef struct orc_acq_result {
    uint8_t prn, detected, has_cn0, _pad[5];
    double code_phase, doppler_hz, peak_metric, threshold, cn0_estimate;
} orc_acq_result;

/* per-(block, satellite) quantities of generate_block Phase 1 — exposed for tests */
typedef struct orc_block_params {
    int32_t visible, _pad;
    double range_m, iono_delay_s, tropo_delay_s, rx_amplitude, doppler_start_hz, doppler_end_hz;
    double initial_code_phase;     /* chips, satellite_emitter.rs:235 */
    uint64_t initial_epoch_offset; /* satellite_emitter.rs:242 */
    double phase_before;           /* doppler_phases[idx] entering the block, radians */
} orc_block_params;

typedef struct orc_scenario orc_scenario;

/* codes */
int orc_e1_code(int channel, int prn, int8_t* out4092);
void orc_e1c_secondary(int8_t* out25);
int orc_gps_ca_code(int prn, int8_t* out1023);
int orc_glonass_code(int frequency_channel, int8_t* out511);          /* gnss/prn.rs:170-216 */
int orc_gps_l5_code(int prn, int q_channel, int8_t* out10230);       /* gnss/prn.rs:376-397 */
void orc_e1c_replica(int prn, double sample_rate, int8_t* out, size_t n);

/* filters */
void orc_blackman_window(size_t length, double* out);
size_t orc_lowpass_taps(double cutoff_hz, double sample_rate, size_t num_taps, double* out, size_t cap);

/* geometry */
void orc_lla_to_ecef(const orc_lla* lla, double* xyz);
void orc_look_angle(const double* obs_xyz, const orc_lla* obs_lla, const double* tgt_xyz,
                    double* elevation_deg, double* azimuth_deg, double* range_m);
double orc_range_rate(const double* obs_pos, const double* obs_vel, const double* tgt_pos, const double* tgt_vel);
double orc_fspl_db(double distance_m, double frequency_hz);
void orc_galileo_position_velocity(int plane, int slot, double t, double* pos, double* vel);
void orc_gps_position_velocity(int plane, int slot, double t, double* pos, double* vel);
double orc_kepler_period(double a);
double orc_solve_kepler(double m, double e);
double orc_antenna_gain_dbi(uint32_t kind, double peak_gain_dbi, double beamwidth_deg, double elevation_deg);

/* emitter (single satellite baseband, no filter) */
int orc_emitter_baseband(uint32_t signal, int prn, int nav_data, size_t num_samples, double sample_rate,
                         double range_m, double iono_delay_s, double tropo_delay_s, size_t sample_offset,
                         double* out_re);

/* scenario */
int orc_scenario_new(const orc_scenario_cfg* cfg, orc_scenario** out);
void orc_scenario_free(orc_scenario* s);
uint64_t orc_scenario_total_samples(const orc_scenario* s);
uint64_t orc_scenario_block_size(const orc_scenario* s);
uint64_t orc_scenario_current_sample(const orc_scenario* s);
int orc_scenario_is_done(const orc_scenario* s);
void orc_scenario_reset(orc_scenario* s);
/* generate_block: returns samples written */
size_t orc_scenario_generate_block(orc_scenario* s, size_t block_size, orc_c64* out);
/* threads > 1: satellites of a block are generated by that many worker threads (rayon `parallel` analogue) */
void orc_scenario_set_threads(orc_scenario* s, int threads);
/* Advance to `sample` (a multiple of the canonical block size) without producing output: the sequential
 * states (Doppler phase, xorshift64, FIR delay lines) end up exactly as if every block had been generated. */
int orc_scenario_skip_to(orc_scenario* s, uint64_t sample);
/* Phase-1 parameters the NEXT generate_block(block_size) call would use, one entry per configured satellite */
int orc_scenario_peek_params(orc_scenario* s, size_t block_size, orc_block_params* out, size_t cap);
int orc_scenario_status(const orc_scenario* s, orc_sat_status* out, size_t cap);
double orc_scenario_noise_std(const orc_scenario* s);
/* unit-amplitude, noise-free contribution of satellite `sat_idx` for the next block (does not advance) */

/* FFT */
void orc_fft(orc_c64* buf, size_t n, int inverse_scaled);

/* PCPS */
typedef struct orc_pcps {
    uint64_t fft_size, code_length;
    double doppler_max_hz, doppler_step_hz, threshold, sample_rate;
    uint64_t coherent_periods;
} orc_pcps;
void orc_pcps_init(orc_pcps* p, uint64_t code_length, double sample_rate);
int orc_pcps_num_bins(const orc_pcps* p);
void orc_pcps_acquire(const orc_pcps* p, const orc_c64* input, size_t n_input, const int8_t* code,
                      size_t code_len, uint8_t prn, orc_acq_result* out);
/* power_out[bins][code_length]; also returns the linear index (d*code_length+phase) of the first maximum */
int64_t orc_pcps_acquire_grid(const orc_pcps* p, const orc_c64* input, size_t n_input, const int8_t* code,
                              size_t code_len, double* power_out);

/* IqFormat::Cf32 sink cast, core/io/format.rs:197-200 */
void orc_to_cf32(const orc_c64* in, size_t n, float* out_interleaved);
/* KlobucharModel::delay_seconds, environment/ionosphere.rs:46-108 */
double orc_klobuchar_delay_s(const double alpha[4], const double beta[4], double elevation_rad, double azimuth_rad,
                             double user_lat_rad, double user_lon_rad, double gps_time_s);
/* SaastamoinenModel::delay_meters (zenith hydrostatic + wet, Chao mapping), environment/troposphere.rs:52-97 */
double orc_saastamoinen_delay_m(double height_m, double temperature_k, double pressure_hpa, double relative_humidity, double elevation_rad);
double orc_saastamoinen_zenith_m(double height_m, double temperature_k, double pressure_hpa, double relative_humidity, int component);
/* core/io/format.rs:203-222, fmt 2 = ci16, 3 = ci8, 4 = cu8; out = interleaved (re, im) integers */
int orc_to_int_format(const orc_c64* in, size_t n, int fmt, void* out);

/* ---- tracking channel, r4w_oracle_track.c: gnss/tracking.rs:107-337, 365-456 ---- */
typedef struct orc_track_state {      /* TrackingState, gnss/types.rs:187-210 */
    double code_phase, carrier_freq_hz, carrier_phase_rad, prompt_i, prompt_q, cn0_dbhz;
    uint64_t ms_count;
    uint8_t prn, carrier_lock, code_lock, bit_sync, pad[4];
} orc_track_state;
typedef struct orc_tracker orc_tracker;
orc_tracker* orc_track_new(uint8_t prn, size_t code_length, double sample_rate, double chipping_rate, double initial_code_phase,
                           double initial_doppler);
void orc_track_free(orc_tracker* t);
void orc_track_set_dll_bandwidth(orc_tracker* t, double bw_hz);
void orc_track_set_pll_bandwidth(orc_tracker* t, double bw_hz);
void orc_track_process(orc_tracker* t, const orc_c64* samples, size_t n, const int8_t* code, orc_track_state* out);
void orc_track_state_get(const orc_tracker* t, orc_track_state* out);
size_t orc_track_nav_bits(const orc_tracker* t, int8_t* out, size_t cap);
double orc_loop_filter_2nd_run(double bw, double T, double disc, size_t n_updates);
double orc_loop_filter_3rd_run(double bw, double T, double disc, size_t n_updates);
void orc_dll_s_curve(double el_spacing, size_t num_points, double* err, double* disc);

/* ---- r4w-sim generic scenario engine, r4w_oracle_sim.c: crates/r4w-sim/src/scenario/ ---- */
void orc_sim_compose_emitter(const orc_c64* baseband, size_t n, double doppler_hz, double sample_rate, double rx_amplitude,
                             double* carrier_phase, orc_c64* composite);
void orc_sim_link(const double* rx_pos, const double* rx_vel, const double* em_pos, const double* em_vel, double carrier_hz,
                  double power_dbm, double* out4);
double orc_sim_noise_power(double noise_floor_dbw_hz, double sample_rate);
void orc_sim_trajectory(int kind, const double* p, double t, double* out6);
void orc_sim_waypoints(const double* points, size_t n, double t, double* out6);

#ifdef __cplusplus
}
#endif
#endif
