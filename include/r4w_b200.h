/*
 * r4w_b200.h — C ABI of libr4w_b200.so, the B200 (sm_100a) drop-in for r4w's GNSS hot path:
 * multi-satellite IQ scenario synthesis and FFT-based PCPS acquisition.
 *
 * Every entry point cites the reference interface it replaces (paths relative to the r4w repo;
 * gnss/ = crates/r4w-core/src/waveform/gnss/).  Conventions follow r4w's own C FFI
 * (crates/r4w-ffi/include/r4w.h:8-25, crates/r4w-ffi/src/lib.rs:255-314): an error enum is returned
 * from every fallible call, handles are opaque heap objects released by NULL-safe *_destroy, output
 * buffers are caller-allocated with explicit lengths, complex samples are {re, im} pairs, nothing
 * unwinds across the boundary.  No torch / C++ types appear in any signature.
 *
 * There is NO CPU fallback: every compute entry point needs a CUDA device and returns
 * R4WB_ERR_CUDA (with r4wb_last_error() text) when none is usable.
 */
#ifndef R4W_B200_H
#define R4W_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- errors: values 0-7 are r4w.h's R4wError (r4w.h:8-25); 100+ are ours ---- */
typedef enum r4wb_error {
    R4WB_OK = 0,
    R4WB_ERR_NULL_POINTER = 1,
    R4WB_ERR_INVALID_SIZE = 2,
    R4WB_ERR_BUFFER_FULL = 3,
    R4WB_ERR_BUFFER_EMPTY = 4,
    R4WB_ERR_INVALID_PARAMETER = 5,
    R4WB_ERR_ALLOCATION_FAILED = 6,
    R4WB_ERR_NOT_SUPPORTED = 7,
    R4WB_ERR_CUDA = 100
} r4wb_error;

/* ---- plain data types ---- */
typedef struct r4wb_cf32 { float re, im; } r4wb_cf32;   /* IqFormat::Cf32 sample, core/io/format.rs:197-200 */
typedef struct r4wb_cf64 { double re, im; } r4wb_cf64;  /* num_complex::Complex64 layout (r4w.h:50-55)       */

typedef enum r4wb_mem { R4WB_MEM_HOST = 0, R4WB_MEM_DEVICE = 1 } r4wb_mem;
/* IqFormat (core/io/format.rs:44-60, write_sample :191-227).  CF32/CF64 are accepted everywhere; the integer sink formats
 * only as synthesis OUTPUT, converted in the store epilogue with the reference's expressions:
 *   CI16  (x * 32767).clamp(-32768, 32767) as i16      CI8  (x * 127).clamp(-128, 127) as i8
 *   CU8   ((x + 1) * 127.5).clamp(0, 255) as u8        (`as` truncates toward zero), interleaved re, im. */
typedef enum r4wb_fmt { R4WB_FMT_CF32 = 0, R4WB_FMT_CF64 = 1, R4WB_FMT_CI16 = 2, R4WB_FMT_CI8 = 3, R4WB_FMT_CU8 = 4 } r4wb_fmt;

/* gnss/types.rs:33-47 (declaration order) */
typedef enum r4wb_signal {
    R4WB_SIG_GPS_L1CA = 0,
    R4WB_SIG_GPS_L5 = 1,
    R4WB_SIG_GLONASS_L1OF = 2,
    R4WB_SIG_GALILEO_E1 = 3,
    R4WB_SIG_GALILEO_E1C = 4,
    R4WB_SIG_GALILEO_E1OS = 5
} r4wb_signal;

/* gnss/environment/antenna.rs:11-31 */
typedef enum r4wb_antenna {
    R4WB_ANT_ISOTROPIC = 0,
    R4WB_ANT_HEMISPHERICAL = 1,
    R4WB_ANT_PATCH = 2,
    R4WB_ANT_CHOKE_RING = 3
} r4wb_antenna;

/* presence bits for the Option<f64> fields of SatelliteConfig (gnss/scenario_config.rs:153-190) */
enum {
    R4WB_HAS_ELEVATION = 1u << 0,
    R4WB_HAS_AZIMUTH = 1u << 1,
    R4WB_HAS_RANGE = 1u << 2,
    R4WB_HAS_RANGE_RATE = 1u << 3,
    R4WB_HAS_DOPPLER = 1u << 4,
    R4WB_HAS_DOPPLER_RATE = 1u << 5,
    R4WB_HAS_CN0 = 1u << 6,
    R4WB_HAS_IONO = 1u << 7,
    R4WB_HAS_TROPO = 1u << 8
};

typedef struct r4wb_lla { double lat_deg, lon_deg, alt_m; } r4wb_lla;   /* core/coordinates.rs:93-101 */

/* POD mirror of SatelliteConfig, gnss/scenario_config.rs:137-191 */
typedef struct r4wb_sat_cfg {
    uint32_t signal;            /* r4wb_signal */
    uint32_t has;               /* R4WB_HAS_* presence mask */
    uint8_t prn, plane, slot, nav_data;
    uint8_t orbital_dynamics;   /* #[serde(default)] false */
    uint8_t _pad[3];
    double tx_power_dbw;
    double elevation_deg, azimuth_deg, range_m, range_rate_mps;
    double doppler_hz, doppler_rate_hz_per_s, cn0_dbhz, iono_delay_m, tropo_delay_m;
} r4wb_sat_cfg;

/* POD mirror of ReceiverConfig (+ ReceiverTrajectory), gnss/scenario_config.rs:304-315, 383-401 */
typedef struct r4wb_receiver_cfg {
    r4wb_lla position;
    uint32_t antenna;           /* r4wb_antenna */
    uint32_t has_trajectory;
    double antenna_peak_gain_dbi, antenna_beamwidth_deg;
    double elevation_mask_deg, noise_figure_db, bandwidth_hz;
    r4wb_lla traj_start, traj_end;
    uint32_t traj_has_speed;
    uint32_t _pad;
    double traj_speed_mps;
} r4wb_receiver_cfg;

/* POD mirror of EnvironmentConfig, gnss/scenario_config.rs:417-437.  The Klobuchar / Saastamoinen models
 * (gnss/environment/ionosphere.rs:26-108, troposphere.rs:26-104) are consulted per block for a satellite that lacks
 * the iono_delay_m / tropo_delay_m override (gnss/scenario.rs:430-439 via SatelliteEmitter::status_at,
 * satellite_emitter.rs:165-178).  The host fills the model fields with the YAML's `ionosphere_model` /
 * `troposphere_model` or, when those are null, KlobucharModel::default_broadcast / SaastamoinenModel::standard_atmosphere. */
typedef struct r4wb_environment_cfg {
    uint32_t ionosphere_enabled, troposphere_enabled, multipath_enabled, multipath_preset;
    double klobuchar_alpha[4], klobuchar_beta[4];
    double tropo_height_m, tropo_temperature_k, tropo_pressure_hpa, tropo_relative_humidity;
} r4wb_environment_cfg;

/* POD mirror of OutputConfig, gnss/scenario_config.rs:455-487 (format/output_path are host-side only) */
typedef struct r4wb_output_cfg {
    double sample_rate, duration_s;
    uint64_t block_size;        /* 0 = ceil(sample_rate * 1 ms), gnss/scenario.rs:667-674 */
    uint64_t seed;
    double start_time_gps_s;
    double lpf_cutoff_hz;       /* 0 = sample_rate / 2, gnss/scenario.rs:210-214 */
} r4wb_output_cfg;

enum {
    R4WB_FLAG_NOISE_OFF = 1u << 0,      /* skip the thermal-noise term of gnss/scenario.rs:530-542 (parity runs) */
    R4WB_FLAG_CLOSED_FORM_PHASE = 1u << 1 /* do not emulate the reference's sequential f64 `phase +=` drift */
};

/* POD mirror of GnssScenarioConfig, gnss/scenario_config.rs:537-547 */
typedef struct r4wb_scenario_cfg {
    uint32_t n_sats;
    uint32_t flags;             /* R4WB_FLAG_* */
    const r4wb_sat_cfg* sats;   /* [n_sats] */
    r4wb_receiver_cfg receiver;
    r4wb_environment_cfg environment;
    r4wb_output_cfg output;
} r4wb_scenario_cfg;

/* SatelliteStatus, gnss/satellite_emitter.rs:19-34 */
typedef struct r4wb_sat_status {
    uint32_t signal;
    uint8_t prn, visible, _pad[2];
    double elevation_deg, azimuth_deg, range_m, range_rate_mps, doppler_hz, cn0_dbhz;
    double iono_delay_m, tropo_delay_m, antenna_gain_dbi, clock_correction_s;
} r4wb_sat_status;

/* AcquisitionResult, gnss/types.rs:168-183 (cn0_estimate: Option<f64> -> has_cn0) */
typedef struct r4wb_acq_result {
    uint8_t prn, detected, has_cn0, _pad[5];
    double code_phase;          /* integer lag in samples, as f64 (gnss/acquisition.rs:189) */
    double doppler_hz;
    double peak_metric;
    double threshold;
    double cn0_estimate;
} r4wb_acq_result;

typedef struct r4wb_scenario r4wb_scenario;   /* opaque: GnssScenario, gnss/scenario.rs:51-74 */
typedef struct r4wb_pcps r4wb_pcps;           /* opaque: PcpsAcquisition, gnss/acquisition.rs:40-55 */

/* Threading.  A scenario, tracker or composer handle mirrors a `&mut self` object: one host thread at a time.  A pcps
 * handle mirrors `PcpsAcquisition::acquire(&self)` (re-entrant in the reference): after configuration it may be shared
 * between threads; searches on one handle take turns on its device scratch.  Every call returns with its results complete
 * (the calling thread's stream is synchronised before a host result is handed back).  Errors never unwind through the
 * boundary: a code is returned and the text is kept per thread (r4wb_last_error). */

/* ---- library ---- */
const char* r4wb_version(void);                 /* static NUL-terminated, like r4w_version (r4w-ffi/src/lib.rs:119-123) */
const char* r4wb_last_error(void);              /* thread-local text of the last failure */
r4wb_error r4wb_init(int device);               /* cudaSetDevice(device) + context warm-up; -1 keeps the current device */
r4wb_error r4wb_device_count(int* n);
/* Multi-GPU inside the library (the caller is one host process, e.g. `r4w gnss scenario`,
 * crates/r4w-cli/src/main.rs:4442): initialise devices 0 .. n_gpus-1 (n_gpus <= 0: every visible device).  From then on the
 * host-buffer batch calls shard their independent units over those devices — r4wb_pcps_acquire_batch by snapshot,
 * r4wb_scenario_generate(..., R4WB_MEM_HOST) by time segment — each device working into its slice of the caller's buffers;
 * nothing is exchanged between devices (the 32-byte results / samples go straight to the host).  Handles stay bound to the
 * device that was current at creation for everything else.  r4wb_init_devices(1) switches the sharding off again. */
r4wb_error r4wb_init_devices(int n_gpus);
int r4wb_devices_initialised(void);             /* devices the batch calls shard over (1 until r4wb_init_devices) */
/* Launch all subsequent work of this thread's handles on `cuda_stream` (a cudaStream_t; NULL = default stream). */
r4wb_error r4wb_set_stream(void* cuda_stream);
/* pinned host memory for callers that want full-rate D2H/H2D (optional) */
r4wb_error r4wb_host_alloc(void** p, size_t bytes);
r4wb_error r4wb_host_free(void* p);
/* kernels launched by this library since process start (all threads) */
uint64_t r4wb_kernel_launches(void);

/* ---- scenario synthesis ---- */
/* GnssScenario::new, gnss/scenario.rs:78-237 */
r4wb_error r4wb_scenario_create(const r4wb_scenario_cfg* cfg, r4wb_scenario** out);
/* Drop */
void r4wb_scenario_destroy(r4wb_scenario* h);
/* total_samples / block_size / is_done / progress / reset, gnss/scenario.rs:636-674 */
uint64_t r4wb_scenario_total_samples(const r4wb_scenario* h);
uint64_t r4wb_scenario_block_size(const r4wb_scenario* h);
int r4wb_scenario_is_done(const r4wb_scenario* h);
double r4wb_scenario_progress(const r4wb_scenario* h);
r4wb_error r4wb_scenario_reset(r4wb_scenario* h);
uint64_t r4wb_scenario_current_sample(const r4wb_scenario* h);
/* GnssScenario::generate_block, gnss/scenario.rs:308-546.  Produces min(n, remaining) samples as ONE
 * reference block starting at the handle's current_sample (per-block geometry, Doppler ramp over the
 * n samples), writes them as cf32 (the CLI's sink cast, core/io/format.rs:197-200; crates/r4w-cli/src/
 * main.rs:4488-4500) and advances.  *written = 0 when done (the reference returns an empty Vec). */
r4wb_error r4wb_scenario_generate_block(r4wb_scenario* h, uint64_t n, void* dst, r4wb_mem where,
                                        r4wb_fmt fmt, uint64_t* written);
/* The same call for a host consumer that only READS the block (the CLI's sink: IqFormat::write_samples into a BufWriter,
 * crates/r4w-cli/src/main.rs:4488-4500; a `&[IQSample]` on the Rust side): *block points at the next min(n, remaining)
 * samples in pinned host memory owned by the handle — for canonical block sizes the render-ahead ring itself, so the call
 * copies nothing — and stays valid until the next call on this handle.  *written = 0 and *block = NULL when done.
 * Same samples, same advance, same last_power_sum as r4wb_scenario_generate_block. */
r4wb_error r4wb_scenario_generate_block_view(r4wb_scenario* h, uint64_t n, r4wb_fmt fmt, const void** block,
                                             uint64_t* written);
/* Random access to the stream the CLI loop (`while !is_done { generate_block(block_size()) }`,
 * main.rs:4488-4500 / GnssScenario::generate, scenario.rs:549-561) would produce: samples
 * [first, first+n) of the canonical block partition.  Does not move current_sample.  This is the
 * throughput entry point (time-sharding across GPUs = disjoint [first, first+n) ranges). */
r4wb_error r4wb_scenario_generate(r4wb_scenario* h, uint64_t first, uint64_t n, void* dst,
                                  r4wb_mem where, r4wb_fmt fmt);
/* GnssScenario::generate, gnss/scenario.rs:549-561: `while !is_done { generate_block(block_size()) }` — everything from
 * current_sample to the end into dst (capacity `cap` samples; InvalidSize when smaller than the remainder), *written =
 * samples produced (0 when already done).  Leaves the handle done.  When every earlier generate_block call used the
 * canonical block size the remainder is rendered in one pass (same samples as r4wb_scenario_generate); after odd-sized
 * blocks it walks the reference's own block partition from current_sample. */
r4wb_error r4wb_scenario_generate_rest(r4wb_scenario* h, void* dst, uint64_t cap, r4wb_mem where, r4wb_fmt fmt,
                                       uint64_t* written);
/* The CLI's file sink (`r4w gnss scenario --output`, crates/r4w-cli/src/main.rs:4483-4509: BufWriter + IqFormat::
 * write_samples per block, core/io/format.rs:191-227): renders the whole scenario [0, total_samples) in `fmt` and streams
 * it into `path` (created/truncated) — device staging -> pinned host buffers -> a writer thread, all three overlapped.
 * *samples / *bytes = what was written, *power_sum = sum |s|^2 of the pre-conversion samples (the avg-power line);
 * any of the three may be NULL.  Leaves the handle done (is_done() = 1) as the CLI loop does.  InvalidParameter when the
 * file cannot be created, BufferFull on a short write. */
r4wb_error r4wb_scenario_write_file(r4wb_scenario* h, const char* path, r4wb_fmt fmt, uint64_t* samples, uint64_t* bytes,
                                    double* power_sum);
/* Sum of |s|^2 over the last generate / generate_block call (the CLI's avg-power line, main.rs:4494-4509) */
r4wb_error r4wb_scenario_last_power_sum(const r4wb_scenario* h, double* power_sum);
/* Which synthesis kernels rendered the bulk of the last generate call: 0 = k_synth (general), 1 = the period-resident
 * kernels (constant-delay, constant-Doppler scenarios; R4WB_SYNTH_PERIODIC=0 in the environment disables them).
 * Diagnostic for A/B parity tests; no counterpart in the reference. */
uint32_t r4wb_scenario_last_path(const r4wb_scenario* h);
/* Optional device-side timing of the last generate call (measurement aid, no reference counterpart): when enabled, CUDA
 * events bracket every synthesis-kernel launch on its stream.  ms[3] / launches[3] are the summed durations and launch
 * counts of {k_synth, k_synth_periodic, k_periodic_fix, k_synth_lat} (ms / launches: 4 entries each); last_profile waits
 * for the events. */
r4wb_error r4wb_scenario_set_profiling(r4wb_scenario* h, int enabled);
r4wb_error r4wb_scenario_last_profile(r4wb_scenario* h, double* ms, uint64_t* launches);
/* GnssScenario::satellite_status, gnss/scenario.rs:564-633 */
r4wb_error r4wb_scenario_status(const r4wb_scenario* h, r4wb_sat_status* out, uint32_t cap, uint32_t* n);

/* ---- codes ---- */
/* GalileoE1CodeGenerator::new_e1b / new_e1c + unpack_code, gnss/prn.rs:268-292, galileo_e1_codes.rs:17-25.
 * channel 0 = E1B, 1 = E1C; out[4092] = +1/-1. */
r4wb_error r4wb_e1_code(uint32_t channel, uint8_t prn, int8_t* out, uint64_t cap);
/* GpsCaCodeGenerator::generate_code, gnss/prn.rs:34-162; PRN 1-32; out[1023] = +1/-1 */
r4wb_error r4wb_gps_ca_code(uint8_t prn, int8_t* out, uint64_t cap);
/* GpsL5CodeGenerator::new_i5, gnss/prn.rs:345-397 (the reference's simplified XA / XB pair); PRN 1-32; out[10230] = +1/-1 */
r4wb_error r4wb_gps_l5_code(uint8_t prn, int8_t* out, uint64_t cap);
/* GlonassCodeGenerator::generate_code, gnss/prn.rs:170-216: the 511-chip m-sequence shared by every satellite; out[511] */
r4wb_error r4wb_glonass_code(int8_t* out, uint64_t cap);
/* GalileoE1CodeGenerator::secondary_code, galileo_e1_codes.rs:29-31; out[25] */
r4wb_error r4wb_e1c_secondary(int8_t* out, uint64_t cap);
/* Sampled local replica code[floor(i*1.023e6/fs) mod 4092] * BOC(1,1) for i in [0, n) (no secondary code) —
 * the `code: &[i8]` argument a caller passes to PcpsAcquisition::acquire for an E1C scenario. */
r4wb_error r4wb_e1c_replica(uint8_t prn, double sample_rate, int8_t* out, uint64_t n);

/* ---- PCPS acquisition ---- */
/* PcpsAcquisition::new / with_doppler_range / with_threshold / with_coherent_periods / fft_size,
 * gnss/acquisition.rs:63-93, 252-254 */
r4wb_error r4wb_pcps_create(uint64_t code_length, double sample_rate, r4wb_pcps** out);
void r4wb_pcps_destroy(r4wb_pcps* h);
r4wb_error r4wb_pcps_set_doppler_range(r4wb_pcps* h, double max_hz, double step_hz);
r4wb_error r4wb_pcps_set_threshold(r4wb_pcps* h, double threshold);
r4wb_error r4wb_pcps_set_coherent_periods(r4wb_pcps* h, uint64_t periods);
uint64_t r4wb_pcps_fft_size(const r4wb_pcps* h);
uint32_t r4wb_pcps_num_doppler_bins(const r4wb_pcps* h);
/* PcpsAcquisition::acquire, gnss/acquisition.rs:104-195: one (input, code) pair, host buffers. */
r4wb_error r4wb_pcps_acquire(r4wb_pcps* h, const void* input, r4wb_fmt fmt, uint64_t n_input,
                             const int8_t* code, uint64_t code_len, uint8_t prn, r4wb_acq_result* out);
/* The same search for n_snapshots inputs x n_codes replicas in one call: snapshot s starts at
 * input + s*snapshot_stride samples and holds n_input samples; codes is [n_codes][code_len];
 * out is [n_snapshots][n_codes].  `where` says where `input` lives; codes/prns/out are host memory. */
r4wb_error r4wb_pcps_acquire_batch(r4wb_pcps* h, const void* input, r4wb_fmt fmt, r4wb_mem where,
                                   uint64_t n_snapshots, uint64_t snapshot_stride, uint64_t n_input,
                                   const int8_t* codes, uint64_t code_len, const uint8_t* prns,
                                   uint32_t n_codes, r4wb_acq_result* out);
/* PcpsAcquisition::acquire_grid, gnss/acquisition.rs:199-249: power_out is [num_doppler_bins][code_length] f64 */
r4wb_error r4wb_pcps_acquire_grid(r4wb_pcps* h, const void* input, r4wb_fmt fmt, uint64_t n_input,
                                  const int8_t* code, uint64_t code_len, double* power_out, uint64_t cap);
/* statistics of the last batch: rows run through the f64 near-tie guard */
uint64_t r4wb_pcps_guard_count(const r4wb_pcps* h);
/* Optional device-side timing of the last acquire_batch (measurement aid, no reference counterpart): when
 * enabled, CUDA events bracket every kernel launch on the launching stream.  ms[4] / launches[4] are the summed
 * durations and launch counts of {code spectra, forward FFT, inverse FFT + peak, pair reduce}. */
r4wb_error r4wb_pcps_set_profiling(r4wb_pcps* h, int enabled);
r4wb_error r4wb_pcps_last_profile(const r4wb_pcps* h, double* ms, uint64_t* launches);

/* ---- tracking channels: TrackingChannel, crates/r4w-core/src/waveform/gnss/tracking.rs ---- */
/* TrackingChannel::new(prn, code_length, sample_rate, chipping_rate, initial_code_phase, initial_doppler) :107-153 plus the
 * with_dll_bandwidth / with_pll_bandwidth builders :156-167 (<= 0: the reference defaults, 1 Hz and 15 Hz) */
typedef struct r4wb_track_cfg {
    double sample_rate, chipping_rate, initial_code_phase, initial_doppler, dll_bandwidth_hz, pll_bandwidth_hz;
    uint64_t code_length;
    uint8_t prn, pad[7];
} r4wb_track_cfg;
/* TrackingState, gnss/types.rs:187-210 */
typedef struct r4wb_track_state {
    double code_phase, carrier_freq_hz, carrier_phase_rad, prompt_i, prompt_q, cn0_dbhz;
    uint64_t ms_count;
    uint8_t prn, carrier_lock, code_lock, bit_sync, pad[4];
} r4wb_track_state;
typedef struct r4wb_tracker r4wb_tracker;   /* a bank of independent channels, state resident on the GPU */
r4wb_error r4wb_track_create(const r4wb_track_cfg* cfgs, uint32_t n_channels, r4wb_tracker** out);
void r4wb_track_destroy(r4wb_tracker* h);
/* n_periods consecutive TrackingChannel::process(samples, code) calls (:177-313) on every channel.  Channel c reads
 * samples[c * channel_stride + p * n_per_period ...] for period p (channel_stride = 0: all channels track the same
 * stream) and the +-1 chips codes[c * code_stride ...]; out is [n_periods][n_channels], the state returned by each call. */
r4wb_error r4wb_track_process(r4wb_tracker* h, const void* samples, r4wb_fmt fmt, r4wb_mem where, uint64_t n_per_period,
                              uint64_t n_periods, uint64_t channel_stride, const int8_t* codes, uint64_t code_stride,
                              r4wb_track_state* out);
/* TrackingChannel::state :344-358 for every channel */
r4wb_error r4wb_track_state_get(const r4wb_tracker* h, r4wb_track_state* out, uint32_t cap);
/* TrackingChannel::nav_bits :339-342: *n = bits detected so far, the first min(cap, *n) copied to out */
r4wb_error r4wb_track_nav_bits(const r4wb_tracker* h, uint32_t channel, int8_t* out, uint64_t cap, uint64_t* n);

/* ---- r4w-sim generic scenario engine: the per-sample loops of ScenarioEngine::generate_block,
 * crates/r4w-sim/src/scenario/engine.rs:61-137.  The emitters are caller objects (trait Emitter) and the per-block geometry
 * (:68-101) is the caller's f64 work; the composer owns ScenarioEngine::carrier_phases and the noise stream. ---- */
typedef struct r4wb_composer r4wb_composer;
/* noise_std = sqrt(ScenarioConfig::noise_power_linear() / 2) (engine.rs:126-127), seed = ScenarioConfig::seed */
r4wb_error r4wb_composer_create(uint32_t n_emitters, double sample_rate, double noise_std, uint64_t seed, r4wb_composer** out);
void r4wb_composer_destroy(r4wb_composer* h);
/* ScenarioEngine::reset, engine.rs:149-153 */
r4wb_error r4wb_composer_reset(r4wb_composer* h);
/* One block: baseband is [n_emitters][n] (Emitter::generate_iq of every emitter; rows of inactive emitters are ignored),
 * doppler_hz / amplitude / active are per emitter (active may be NULL = all active).  out[n] = sum of the Doppler-rotated,
 * scaled emitters + receiver noise; the carrier phases advance as engine.rs:108-113 (f64, exactly).  The per-sample products
 * and the sum over emitters are formed in f32: a CF64 output carries f32-precision values widened to f64 (the reference sums
 * Complex64), and n = 0 (no samples) writes nothing — pass the block length even when no emitter is active to get the noise. */
r4wb_error r4wb_composer_block(r4wb_composer* h, const void* baseband, r4wb_fmt in_fmt, r4wb_mem in_where, uint64_t n,
                               const double* doppler_hz, const double* amplitude, const uint8_t* active, void* out,
                               r4wb_fmt out_fmt, r4wb_mem out_where);
/* ScenarioEngine::carrier_phases (radians) after the last block */
r4wb_error r4wb_composer_phases(const r4wb_composer* h, double* out, uint32_t cap);

#ifdef __cplusplus
}
#endif
#endif /* R4W_B200_H */
