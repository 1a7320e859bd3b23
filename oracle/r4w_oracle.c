/*
 * r4w_oracle.c — CPU restatement (f64) of r4w's GNSS scenario synthesis + PCPS acquisition.
 * TEST INFRASTRUCTURE ONLY (see r4w_oracle.h).  Built with plain gcc, no fast-math, so that every
 * expression keeps the reference's evaluation order.  Citations: paths relative to /root/reference,
 * gnss/ = crates/r4w-core/src/waveform/gnss/, core/ = crates/r4w-core/src/.
 */
#include "r4w_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static const uint8_t E1_PACKED[2 * 50 * 512] = {
#include "../data/galileo_e1_codes.inc"
};

/* ------------------------------------------------------------------ codes */

/* unpack_code, gnss/galileo_e1_codes.rs:17-25; new_e1b/new_e1c, gnss/prn.rs:268-292 */
int orc_e1_code(int channel, int prn, int8_t* out)
{
    if (prn < 1 || prn > 50 || channel < 0 || channel > 1) return -1;
    const uint8_t* packed = E1_PACKED + ((size_t)channel * 50 + (size_t)(prn - 1)) * 512;
    for (int i = 0; i < 4092; ++i) {
        int byte_idx = i / 8;
        int bit_idx = 7 - (i % 8);
        out[i] = (((packed[byte_idx] >> bit_idx) & 1) == 0) ? 1 : -1;
    }
    return 0;
}

/* E1C_SECONDARY, gnss/galileo_e1_codes.rs:27-31 (binary 0011100000000101011110001, 0 -> +1) */
static const int8_t E1C_SECONDARY[25] = {1, 1, -1, -1, -1, 1, 1, 1, 1, 1, 1, 1, 1, -1, 1, -1, 1, -1, -1, -1, -1, 1, 1, 1, -1};

void orc_e1c_secondary(int8_t* out25) { memcpy(out25, E1C_SECONDARY, 25); }

/* Lfsr::clock / tap_output, core/spreading/lfsr.rs:58-83 */
typedef struct { uint32_t state, polynomial; int degree; } lfsr_t;
static int lfsr_clock(lfsr_t* l)
{
    int output = (int)((l->state >> (l->degree - 1)) & 1u);
    uint32_t feedback = (uint32_t)__builtin_popcount(l->state & l->polynomial) & 1u;
    uint32_t mask = (1u << l->degree) - 1u;
    l->state = ((l->state << 1) | feedback) & mask;
    return output;
}
static int lfsr_tap(const lfsr_t* l, int tap) { return (int)((l->state >> (tap - 1)) & 1u); }

/* GPS_CA_TAPS + GpsCaCodeGenerator, gnss/prn.rs:34-162 */
static const uint8_t GPS_CA_TAPS[32][2] = {
    {2, 6}, {3, 7}, {4, 8}, {5, 9}, {1, 9}, {2, 10}, {1, 8}, {2, 9}, {3, 10}, {2, 3}, {3, 4},
    {5, 6}, {6, 7}, {7, 8}, {8, 9}, {9, 10}, {1, 4}, {2, 5}, {3, 6}, {4, 7}, {5, 8}, {6, 9},
    {1, 3}, {4, 6}, {5, 7}, {6, 8}, {7, 9}, {8, 10}, {1, 6}, {2, 7}, {3, 8}, {4, 9}};

int orc_gps_ca_code(int prn, int8_t* out)
{
    if (prn < 1 || prn > 32) return -1;
    lfsr_t g1 = {0x3FF, 0x204, 10}, g2 = {0x3FF, 0x3A6, 10};
    int ta = GPS_CA_TAPS[prn - 1][0], tb = GPS_CA_TAPS[prn - 1][1];
    for (int i = 0; i < 1023; ++i) {
        int g1_out = lfsr_clock(&g1);
        int g2_tap = lfsr_tap(&g2, ta) ^ lfsr_tap(&g2, tb);
        lfsr_clock(&g2);
        out[i] = ((g1_out ^ g2_tap) == 0) ? 1 : -1;
    }
    return 0;
}

/* GlonassCodeGenerator, gnss/prn.rs:170-216: 9-stage LFSR 1 + x^5 + x^9, all-ones start, the same 511 chips for every
 * satellite; the constructor asserts the frequency channel -7..6 (the scenario passes `prn as i8`, so PRN 0..6) */
int orc_glonass_code(int frequency_channel, int8_t* out)
{
    if (frequency_channel < -7 || frequency_channel > 6) return -1;
    lfsr_t l = {0x1FF, 0x110, 9};
    for (int i = 0; i < 511; ++i) out[i] = lfsr_clock(&l) == 0 ? 1 : -1;
    return 0;
}

/* GpsL5CodeGenerator::generate_l5_code, gnss/prn.rs:376-397 (the reference's simplified XA / XB pair: no short cycle) */
int orc_gps_l5_code(int prn, int q_channel, int8_t* out)
{
    if (prn < 1 || prn > 32) return -1;
    uint32_t xb_init = ((uint32_t)prn * 0x2468u + (q_channel ? 0xACE0u : 0x1357u)) & 0x1FFFu;
    if (xb_init < 1u) xb_init = 1u;
    lfsr_t xa = {0x1FFF, 0x1E01, 13}, xb = {xb_init, q_channel ? 0x1B4Fu : 0x1AE3u, 13};
    for (int i = 0; i < 10230; ++i) out[i] = ((lfsr_clock(&xa) ^ lfsr_clock(&xb)) == 0) ? 1 : -1;
    return 0;
}

/* ------------------------------------------------------------------ signal constants, gnss/types.rs:62-127 */
enum { SIG_GPS_L1CA = 0, SIG_GPS_L5 = 1, SIG_GLONASS_L1OF = 2, SIG_GAL_E1 = 3, SIG_GAL_E1C = 4, SIG_GAL_E1OS = 5 };

static double sig_carrier_hz(uint32_t s)
{
    switch (s) { case SIG_GPS_L5: return 1176450000.0; case SIG_GLONASS_L1OF: return 1602000000.0; default: return 1575420000.0; }
}
static double sig_chipping_rate(uint32_t s)
{
    switch (s) { case SIG_GPS_L5: return 10230000.0; case SIG_GLONASS_L1OF: return 511000.0; default: return 1023000.0; }
}
static size_t sig_code_length(uint32_t s)
{
    switch (s) { case SIG_GPS_L1CA: return 1023; case SIG_GPS_L5: return 10230; case SIG_GLONASS_L1OF: return 511; default: return 4092; }
}
static double sig_code_period_s(uint32_t s) { return (double)sig_code_length(s) / sig_chipping_rate(s); }
static double sig_nav_rate(uint32_t s)
{
    switch (s) { case SIG_GAL_E1: case SIG_GAL_E1OS: return 250.0; case SIG_GAL_E1C: return 0.0; default: return 50.0; }
}
static int sig_is_galileo(uint32_t s) { return s == SIG_GAL_E1 || s == SIG_GAL_E1C || s == SIG_GAL_E1OS; }

/* ------------------------------------------------------------------ filters */

/* blackman_window, core/filters/windows.rs:137-151 */
void orc_blackman_window(size_t length, double* out)
{
    if (length == 0) return;
    if (length == 1) { out[0] = 1.0; return; }
    double n_minus_1 = (double)(length - 1);
    for (size_t n = 0; n < length; ++n) {
        double x = 2.0 * M_PI * (double)n / n_minus_1;
        out[n] = 0.42 - 0.5 * cos(x) + 0.08 * cos(2.0 * x);
    }
}

/* FirFilter::lowpass -> design_lowpass_windowed(.., Blackman), core/filters/fir.rs:72-75, 458-499 */
size_t orc_lowpass_taps(double cutoff_hz, double sample_rate, size_t num_taps, double* out, size_t cap)
{
    if (num_taps % 2 == 0) num_taps += 1;
    if (num_taps > cap) return 0;
    double fc = cutoff_hz / sample_rate;
    double m = (double)(num_taps - 1);
    double mid = m / 2.0;
    double* w = (double*)malloc(num_taps * sizeof(double));
    orc_blackman_window(num_taps, w);
    for (size_t i = 0; i < num_taps; ++i) {
        double n = (double)i;
        double sinc = (fabs(n - mid) < 1e-10) ? 2.0 * M_PI * fc : sin(2.0 * M_PI * fc * (n - mid)) / (n - mid);
        out[i] = sinc * w[i];
    }
    free(w);
    double sum = 0.0;
    for (size_t i = 0; i < num_taps; ++i) sum += out[i];
    if (fabs(sum) > 1e-10)
        for (size_t i = 0; i < num_taps; ++i) out[i] /= sum;
    return num_taps;
}

/* FirFilter (complex delay line; only the real lane is ever non-zero on this path), core/filters/fir.rs:392-409 */
#define FIR_TAPS 63
typedef struct { double coeffs[FIR_TAPS]; double delay_re[FIR_TAPS]; size_t delay_idx; } fir_t;

static void fir_reset(fir_t* f) { memset(f->delay_re, 0, sizeof f->delay_re); f->delay_idx = 0; }
static inline double fir_process(fir_t* f, double input)
{
    f->delay_re[f->delay_idx] = input;
    double output = 0.0;
    const size_t len = FIR_TAPS;
    for (size_t i = 0; i < len; ++i) {
        size_t delay_pos = (f->delay_idx + len - i) % len;
        output += f->delay_re[delay_pos] * f->coeffs[i];
    }
    f->delay_idx = (f->delay_idx + 1) % len;
    return output;
}

/* ------------------------------------------------------------------ geometry, core/coordinates.rs */
static const double WGS84_A = 6378137.0;
static const double WGS84_F = 1.0 / 298.257223563;
#define WGS84_E2 (2.0 * WGS84_F - WGS84_F * WGS84_F)
static const double SPEED_OF_LIGHT = 299792458.0;

static inline double to_radians(double d) { return d * (M_PI / 180.0); }
static inline double to_degrees(double r) { return r * (180.0 / M_PI); }

/* lla_to_ecef, core/coordinates.rs:129-144 */
void orc_lla_to_ecef(const orc_lla* lla, double* xyz)
{
    double lat = to_radians(lla->lat_deg), lon = to_radians(lla->lon_deg);
    double sin_lat = sin(lat), cos_lat = cos(lat), sin_lon = sin(lon), cos_lon = cos(lon);
    double n = WGS84_A / sqrt(1.0 - WGS84_E2 * sin_lat * sin_lat);
    xyz[0] = (n + lla->alt_m) * cos_lat * cos_lon;
    xyz[1] = (n + lla->alt_m) * cos_lat * sin_lon;
    xyz[2] = (n * (1.0 - WGS84_E2) + lla->alt_m) * sin_lat;
}

/* look_angle, core/coordinates.rs:191-222 */
void orc_look_angle(const double* obs, const orc_lla* obs_lla, const double* tgt,
                    double* elevation_deg, double* azimuth_deg, double* range_m)
{
    double dx = tgt[0] - obs[0], dy = tgt[1] - obs[1], dz = tgt[2] - obs[2];
    *range_m = sqrt(dx * dx + dy * dy + dz * dz);
    double lat = to_radians(obs_lla->lat_deg), lon = to_radians(obs_lla->lon_deg);
    double sin_lat = sin(lat), cos_lat = cos(lat), sin_lon = sin(lon), cos_lon = cos(lon);
    double east = -sin_lon * dx + cos_lon * dy;
    double north = -sin_lat * cos_lon * dx - sin_lat * sin_lon * dy + cos_lat * dz;
    double up = cos_lat * cos_lon * dx + cos_lat * sin_lon * dy + sin_lat * dz;
    *elevation_deg = to_degrees(atan2(up, sqrt(east * east + north * north)));
    double az = to_degrees(atan2(east, north));
    if (az < 0.0) az += 360.0;
    *azimuth_deg = az;
}

/* range_rate + direction_to + dot_direction, core/coordinates.rs:41-51, 86-88, 225-238 */
double orc_range_rate(const double* op, const double* ov, const double* tp, const double* tv)
{
    double dx = tp[0] - op[0], dy = tp[1] - op[1], dz = tp[2] - op[2];
    double r = sqrt(dx * dx + dy * dy + dz * dz);
    double dir[3] = {0.0, 0.0, 0.0};
    if (!(r < 1e-10)) { dir[0] = dx / r; dir[1] = dy / r; dir[2] = dz / r; }
    double rvx = tv[0] - ov[0], rvy = tv[1] - ov[1], rvz = tv[2] - ov[2];
    return rvx * dir[0] + rvy * dir[1] + rvz * dir[2];
}

/* fspl_db, core/coordinates.rs:241-246 */
double orc_fspl_db(double distance_m, double frequency_hz)
{
    if (distance_m <= 0.0 || frequency_hz <= 0.0) return 0.0;
    return 20.0 * log10(4.0 * M_PI * distance_m * frequency_hz / SPEED_OF_LIGHT);
}

/* KeplerianOrbit, gnss/environment/orbit.rs */
static const double GM_EARTH = 3.986004418e14;
static const double OMEGA_E = 7.2921150e-5;
typedef struct { double a, e, i, omega_0, omega, m0, t_epoch, omega_dot; } orbit_t;

/* solve_kepler, orbit.rs:203-213 */
double orc_solve_kepler(double m, double e)
{
    double ecc = m;
    for (int k = 0; k < 20; ++k) {
        double de = (ecc - e * sin(ecc) - m) / (1.0 - e * cos(ecc));
        ecc -= de;
        if (fabs(de) < 1e-14) break;
    }
    return ecc;
}

double orc_kepler_period(double a) { return 2.0 * M_PI / sqrt(GM_EARTH / (a * a * a)); }

/* position_velocity_at, orbit.rs:49-119 */
static void orbit_pv(const orbit_t* o, double t, double* pos, double* vel)
{
    double dt = t - o->t_epoch;
    double n = sqrt(GM_EARTH / (o->a * o->a * o->a));
    double m = fmod(o->m0 + n * dt, 2.0 * M_PI);
    double ecc_anom = orc_solve_kepler(m, o->e);
    double sin_e = sin(ecc_anom), cos_e = cos(ecc_anom);
    double sqrt_1_e2 = sqrt(1.0 - o->e * o->e);
    double true_anom = atan2(sqrt_1_e2 * sin_e, cos_e - o->e);
    double r = o->a * (1.0 - o->e * cos_e);
    double x_orb = r * cos(true_anom), y_orb = r * sin(true_anom);
    double h = sqrt(GM_EARTH * o->a * (1.0 - o->e * o->e));
    double vx_orb = -GM_EARTH / h * sin(true_anom);
    double vy_orb = GM_EARTH / h * (o->e + cos(true_anom));
    double omega_t = o->omega_0 + o->omega_dot * dt;
    double cos_omega = cos(o->omega), sin_omega = sin(o->omega);
    double cos_raan = cos(omega_t), sin_raan = sin(omega_t);
    double cos_i = cos(o->i), sin_i = sin(o->i);
    double x_eci = (cos_raan * cos_omega - sin_raan * sin_omega * cos_i) * x_orb
                 + (-cos_raan * sin_omega - sin_raan * cos_omega * cos_i) * y_orb;
    double y_eci = (sin_raan * cos_omega + cos_raan * sin_omega * cos_i) * x_orb
                 + (-sin_raan * sin_omega + cos_raan * cos_omega * cos_i) * y_orb;
    double z_eci = (sin_omega * sin_i) * x_orb + (cos_omega * sin_i) * y_orb;
    double vx_eci = (cos_raan * cos_omega - sin_raan * sin_omega * cos_i) * vx_orb
                  + (-cos_raan * sin_omega - sin_raan * cos_omega * cos_i) * vy_orb;
    double vy_eci = (sin_raan * cos_omega + cos_raan * sin_omega * cos_i) * vx_orb
                  + (-sin_raan * sin_omega + cos_raan * cos_omega * cos_i) * vy_orb;
    double vz_eci = (sin_omega * sin_i) * vx_orb + (cos_omega * sin_i) * vy_orb;
    double theta = OMEGA_E * t;
    double cos_t = cos(theta), sin_t = sin(theta);
    double x_ecef = cos_t * x_eci + sin_t * y_eci;
    double y_ecef = -sin_t * x_eci + cos_t * y_eci;
    double z_ecef = z_eci;
    pos[0] = x_ecef; pos[1] = y_ecef; pos[2] = z_ecef;
    vel[0] = cos_t * vx_eci + sin_t * vy_eci + OMEGA_E * y_ecef;
    vel[1] = -sin_t * vx_eci + cos_t * vy_eci - OMEGA_E * x_ecef;
    vel[2] = vz_eci;
}

/* gps_nominal / galileo_nominal / glonass_nominal, orbit.rs:125-199 */
static orbit_t orbit_gps(int plane, int slot)
{
    orbit_t o = {26559700.0, 0.0, to_radians(55.0), (double)plane * to_radians(60.0), 0.0,
                 (double)slot * to_radians(60.0), 0.0, 0.0};
    return o;
}
static orbit_t orbit_galileo(int plane, int slot)
{
    double raan_spacing = to_radians(120.0), slot_spacing = to_radians(45.0), walker_phase = to_radians(15.0);
    double raan_cal = to_radians(118.0), m0_cal = to_radians(176.0);
    orbit_t o = {29600318.0, 0.0, to_radians(56.0), (double)plane * raan_spacing + raan_cal, 0.0,
                 (double)slot * slot_spacing + (double)plane * walker_phase + m0_cal, 0.0, 0.0};
    return o;
}
static orbit_t orbit_glonass(int plane, int slot)
{
    orbit_t o = {25508000.0, 0.0, to_radians(64.8), (double)plane * to_radians(120.0), 0.0,
                 (double)slot * to_radians(45.0), 0.0, 0.0};
    return o;
}
/* create_orbit, gnss/scenario.rs:718-725 */
static orbit_t orbit_for(uint32_t signal, int plane, int slot)
{
    if (sig_is_galileo(signal)) return orbit_galileo(plane, slot);
    if (signal == SIG_GLONASS_L1OF) return orbit_glonass(plane, slot);
    return orbit_gps(plane, slot);
}
void orc_galileo_position_velocity(int plane, int slot, double t, double* pos, double* vel)
{
    orbit_t o = orbit_galileo(plane, slot);
    orbit_pv(&o, t, pos, vel);
}
void orc_gps_position_velocity(int plane, int slot, double t, double* pos, double* vel)
{
    orbit_t o = orbit_gps(plane, slot);
    orbit_pv(&o, t, pos, vel);
}

/* AntennaPattern::gain_dbi, gnss/environment/antenna.rs:35-78 */
double orc_antenna_gain_dbi(uint32_t kind, double peak, double beamwidth_deg, double elevation_deg)
{
    switch (kind) {
    case 0: return 0.0;
    case 1: return elevation_deg >= 0.0 ? peak : -30.0;
    case 2: {
        if (elevation_deg < -5.0) return -30.0;
        double theta = to_radians(90.0 - elevation_deg);
        double half_bw = to_radians(beamwidth_deg / 2.0);
        double n = log10(3.0) / log10(1.0 / cos(half_bw));
        double gain_lin = pow(fabs(cos(theta)), n);
        double gain_db = gain_lin > 1e-6 ? 10.0 * log10(gain_lin) : -30.0;
        return peak + gain_db;
    }
    default: {
        if (elevation_deg < 0.0) return -40.0;
        double theta = to_radians(90.0 - elevation_deg);
        double gain_lin = pow(fabs(cos(theta)), 1.5);
        double gain_db = gain_lin > 1e-6 ? 10.0 * log10(gain_lin) : -40.0;
        return peak + gain_db;
    }
    }
}

/* ReceiverTrajectory::position_at / distance_m, gnss/scenario_config.rs:319-368 */
static orc_lla traj_position_at(const orc_lla* start, const orc_lla* end, double frac)
{
    if (frac < 0.0) frac = 0.0;
    if (frac > 1.0) frac = 1.0;
    double lat1 = to_radians(start->lat_deg), lon1 = to_radians(start->lon_deg);
    double lat2 = to_radians(end->lat_deg), lon2 = to_radians(end->lon_deg);
    double d_lat = lat2 - lat1, d_lon = lon2 - lon1;
    double s1 = sin(d_lat / 2.0), s2 = sin(d_lon / 2.0);
    double a = s1 * s1 + cos(lat1) * cos(lat2) * (s2 * s2);
    double angular_dist = 2.0 * asin(sqrt(a));
    double lat, lon;
    if (fabs(angular_dist) < 1e-12) { lat = lat1; lon = lon1; }
    else {
        double a_coeff = sin((1.0 - frac) * angular_dist) / sin(angular_dist);
        double b_coeff = sin(frac * angular_dist) / sin(angular_dist);
        double x = a_coeff * cos(lat1) * cos(lon1) + b_coeff * cos(lat2) * cos(lon2);
        double y = a_coeff * cos(lat1) * sin(lon1) + b_coeff * cos(lat2) * sin(lon2);
        double z = a_coeff * sin(lat1) + b_coeff * sin(lat2);
        lat = atan2(z, sqrt(x * x + y * y));
        lon = atan2(y, x);
    }
    orc_lla r = {to_degrees(lat), to_degrees(lon), start->alt_m + frac * (end->alt_m - start->alt_m)};
    return r;
}
static double traj_distance_m(const orc_lla* start, const orc_lla* end)
{
    double r = 6371000.0;
    double lat1 = to_radians(start->lat_deg), lon1 = to_radians(start->lon_deg);
    double lat2 = to_radians(end->lat_deg), lon2 = to_radians(end->lon_deg);
    double d_lat = lat2 - lat1, d_lon = lon2 - lon1;
    double s1 = sin(d_lat / 2.0), s2 = sin(d_lon / 2.0);
    double a = s1 * s1 + cos(lat1) * cos(lat2) * (s2 * s2);
    double c = 2.0 * asin(sqrt(a));
    return r * c;
}

/* ------------------------------------------------------------------ emitter */

typedef struct {
    uint32_t signal;
    int prn, nav_data;
    size_t code_len;
    int8_t code[10230];
    int8_t code_e1c[4092];
    orbit_t orbit;
    double tx_power_dbw;
} emitter_t;

/* SatelliteEmitter::new + generate_prn_code, gnss/satellite_emitter.rs:71-120, 361-393 */
static int emitter_init(emitter_t* e, uint32_t signal, int prn, int plane, int slot, double tx_power_dbw, int nav_data)
{
    memset(e, 0, sizeof *e);
    e->signal = signal; e->prn = prn; e->nav_data = nav_data; e->tx_power_dbw = tx_power_dbw;
    e->orbit = orbit_for(signal, plane, slot);
    switch (signal) {
    case SIG_GPS_L1CA: e->code_len = 1023; return orc_gps_ca_code(prn, e->code);
    case SIG_GAL_E1: e->code_len = 4092; return orc_e1_code(0, prn, e->code);
    case SIG_GAL_E1C: e->code_len = 4092; return orc_e1_code(1, prn, e->code);
    case SIG_GAL_E1OS: e->code_len = 4092; if (orc_e1_code(0, prn, e->code)) return -1; return orc_e1_code(1, prn, e->code_e1c);
    case SIG_GPS_L5: e->code_len = 10230; return orc_gps_l5_code(prn, 0, e->code);                 /* new_i5, :369-372 */
    case SIG_GLONASS_L1OF: e->code_len = 511; return orc_glonass_code((int)(int8_t)(uint8_t)prn, e->code);   /* new(prn as i8), :373-376 */
    default: return -2;
    }
}

static inline size_t f64_as_usize(double x) { return (x > 0.0) ? (size_t)x : 0; } /* Rust `as usize`: saturating, NaN -> 0 */

/* SatelliteEmitter::generate_baseband_iq, gnss/satellite_emitter.rs:218-347.  Output is real (imag is 0). */
static void emitter_baseband(const emitter_t* e, size_t num_samples, double sample_rate, double geometric_range_m,
                             double iono_delay_s, double tropo_delay_s, size_t sample_offset, double* out)
{
    double chipping_rate = sig_chipping_rate(e->signal);
    double total_delay_s = geometric_range_m / SPEED_OF_LIGHT + iono_delay_s + tropo_delay_s;
    double chips_delay = total_delay_s * chipping_rate;
    double code_length = (double)e->code_len;
    double initial_code_phase = fmod(chips_delay, code_length);
    size_t initial_epoch_offset = f64_as_usize(chips_delay / code_length);
    double samples_per_chip = sample_rate / chipping_rate;
    double nav_data_rate = sig_nav_rate(e->signal);
    size_t code_periods_per_bit = nav_data_rate > 0.0 ? f64_as_usize(1.0 / (nav_data_rate * sig_code_period_s(e->signal))) : 1;
    int is_composite = e->signal == SIG_GAL_E1OS;
    int is_e1c = e->signal == SIG_GAL_E1C;

    for (size_t i = 0; i < num_samples; ++i) {
        size_t global_i = sample_offset + i;
        double chip_idx_f = initial_code_phase + (double)global_i / samples_per_chip;
        size_t chip_idx = f64_as_usize(fmod(chip_idx_f, code_length));
        double chip_phase = chip_idx_f - floor(chip_idx_f);
        size_t code_epoch_idx = initial_epoch_offset + f64_as_usize(chip_idx_f / code_length);
        size_t ci = chip_idx < e->code_len - 1 ? chip_idx : e->code_len - 1;
        double code_val_e1b = (double)e->code[ci];
        double nav_bit = 1.0;
        if (e->nav_data && nav_data_rate > 0.0) {
            size_t bit_idx = code_epoch_idx / code_periods_per_bit;
            nav_bit = ((bit_idx + (size_t)e->prn) % 2 == 0) ? 1.0 : -1.0;
        }
        double secondary_chip = (double)E1C_SECONDARY[code_epoch_idx % 25];
        double boc11 = 1.0;
        if (is_composite || is_e1c || e->signal == SIG_GAL_E1) {
            double sub_phase = fmod(chip_phase * 2.0, 2.0);
            boc11 = sub_phase < 1.0 ? 1.0 : -1.0;
        }
        double v;
        if (is_composite) {
            double code_val_e1c = (double)e->code_e1c[ci];
            double e1b_val = code_val_e1b * nav_bit * boc11;
            double e1c_val = code_val_e1c * secondary_chip * boc11;
            double scale = 1.0 / sqrt(2.0);
            v = (e1b_val - e1c_val) * scale;
        } else if (is_e1c) {
            v = code_val_e1b * boc11 * secondary_chip;
        } else if (e->signal == SIG_GAL_E1) {
            v = code_val_e1b * nav_bit * boc11;
        } else {
            v = code_val_e1b * nav_bit;
        }
        out[i] = v;
    }
}

int orc_emitter_baseband(uint32_t signal, int prn, int nav_data, size_t num_samples, double sample_rate,
                         double range_m, double iono_delay_s, double tropo_delay_s, size_t sample_offset, double* out_re)
{
    emitter_t* e = (emitter_t*)malloc(sizeof *e);
    int rc = emitter_init(e, signal, prn, 0, 0, 15.0, nav_data);
    if (rc == 0) emitter_baseband(e, num_samples, sample_rate, range_m, iono_delay_s, tropo_delay_s, sample_offset, out_re);
    free(e);
    return rc;
}

/* sampled E1C x BOC(1,1) replica = the emitter with zero delay and no secondary code */
void orc_e1c_replica(int prn, double sample_rate, int8_t* out, size_t n)
{
    int8_t code[4092];
    if (orc_e1_code(1, prn, code)) { memset(out, 0, n); return; }
    double samples_per_chip = sample_rate / 1023000.0;
    for (size_t i = 0; i < n; ++i) {
        double chip_idx_f = (double)i / samples_per_chip;
        size_t chip_idx = f64_as_usize(fmod(chip_idx_f, 4092.0));
        double chip_phase = chip_idx_f - floor(chip_idx_f);
        int boc = fmod(chip_phase * 2.0, 2.0) < 1.0 ? 1 : -1;
        out[i] = (int8_t)(code[chip_idx < 4091 ? chip_idx : 4091] * boc);
    }
}

/* ------------------------------------------------------------------ scenario, gnss/scenario.rs */
#define OS 8 /* BASEBAND_OVERSAMPLE, scenario.rs:47 */

struct orc_scenario {
    orc_scenario_cfg cfg;
    orc_sat_cfg* sats;
    emitter_t* emitters;
    size_t n;
    uint64_t current_sample, total_samples;
    double sample_rate;
    double* doppler_phases;
    double* orbital_doppler_t0;
    double* orbital_range_t0;
    fir_t* lpfs;
    uint64_t rng_state;
    int threads;
};

static void rx_state(const orc_scenario* s, double elapsed_s, orc_lla* lla, double* ecef, double* vel);

/* GnssScenario::new, scenario.rs:78-237 */
int orc_scenario_new(const orc_scenario_cfg* cfg, orc_scenario** out)
{
    orc_scenario* s = (orc_scenario*)calloc(1, sizeof *s);
    s->cfg = *cfg;
    s->n = cfg->n_sats;
    s->sats = (orc_sat_cfg*)malloc(sizeof(orc_sat_cfg) * (s->n ? s->n : 1));
    memcpy(s->sats, cfg->sats, sizeof(orc_sat_cfg) * s->n);
    s->cfg.sats = s->sats;
    s->sample_rate = cfg->output.sample_rate;
    s->total_samples = (uint64_t)ceil(cfg->output.duration_s * s->sample_rate);
    s->emitters = (emitter_t*)calloc(s->n ? s->n : 1, sizeof(emitter_t));
    s->doppler_phases = (double*)calloc(s->n ? s->n : 1, sizeof(double));
    s->orbital_doppler_t0 = (double*)calloc(s->n ? s->n : 1, sizeof(double));
    s->orbital_range_t0 = (double*)calloc(s->n ? s->n : 1, sizeof(double));
    s->lpfs = (fir_t*)calloc(s->n ? s->n : 1, sizeof(fir_t));
    s->threads = 1;
    for (size_t k = 0; k < s->n; ++k) {
        const orc_sat_cfg* sc = &s->sats[k];
        int rc = emitter_init(&s->emitters[k], sc->signal, sc->prn, sc->plane, sc->slot, sc->tx_power_dbw, sc->nav_data);
        if (rc) { orc_scenario_free(s); return rc; }
    }
    s->rng_state = cfg->output.seed > 1 ? cfg->output.seed : 1;

    double t0 = cfg->output.start_time_gps_s;
    orc_lla rx_lla_t0; double rx_ecef_t0[3], rx_vel_t0[3];
    rx_state(s, 0.0, &rx_lla_t0, rx_ecef_t0, rx_vel_t0);
    if (cfg->receiver.has_trajectory) {
        /* velocity at t = 0, scenario.rs:168-183 (same finite difference, without the frac < 1 test) */
        const orc_receiver_cfg* r = &cfg->receiver;
        double dist = traj_distance_m(&r->traj_start, &r->traj_end);
        double speed = r->traj_has_speed ? r->traj_speed_mps : dist / cfg->output.duration_s;
        double travel_time_s = dist / speed;
        double dt = fmin(0.01, travel_time_s * 0.001);
        double frac_dt = dt / travel_time_s;
        orc_lla pos_dt = traj_position_at(&r->traj_start, &r->traj_end, frac_dt);
        double ecef_dt[3];
        orc_lla_to_ecef(&pos_dt, ecef_dt);
        for (int k = 0; k < 3; ++k) rx_vel_t0[k] = (ecef_dt[k] - rx_ecef_t0[k]) / dt;
    }
    for (size_t k = 0; k < s->n; ++k) {
        double sp[3], sv[3], el, az, rg;
        orbit_pv(&s->emitters[k].orbit, t0, sp, sv);
        orc_look_angle(rx_ecef_t0, &rx_lla_t0, sp, &el, &az, &rg);
        double rr = orc_range_rate(rx_ecef_t0, rx_vel_t0, sp, sv);
        s->orbital_doppler_t0[k] = -rr * sig_carrier_hz(s->emitters[k].signal) / SPEED_OF_LIGHT;
        s->orbital_range_t0[k] = rg;
    }
    double oversampled_rate = s->sample_rate * (double)OS;
    double lpf_cutoff = cfg->output.lpf_cutoff_hz > 0.0 ? cfg->output.lpf_cutoff_hz : s->sample_rate / 2.0;
    for (size_t k = 0; k < s->n; ++k) {
        orc_lowpass_taps(lpf_cutoff, oversampled_rate, FIR_TAPS, s->lpfs[k].coeffs, FIR_TAPS);
        fir_reset(&s->lpfs[k]);
    }
    *out = s;
    return 0;
}

void orc_scenario_free(orc_scenario* s)
{
    if (!s) return;
    free(s->sats); free(s->emitters); free(s->doppler_phases); free(s->orbital_doppler_t0);
    free(s->orbital_range_t0); free(s->lpfs); free(s);
}

uint64_t orc_scenario_total_samples(const orc_scenario* s) { return s->total_samples; }
uint64_t orc_scenario_current_sample(const orc_scenario* s) { return s->current_sample; }
int orc_scenario_is_done(const orc_scenario* s) { return s->current_sample >= s->total_samples; }
void orc_scenario_set_threads(orc_scenario* s, int threads) { s->threads = threads > 1 ? threads : 1; }

/* block_size, scenario.rs:667-674 */
uint64_t orc_scenario_block_size(const orc_scenario* s)
{
    if (s->cfg.output.block_size > 0) return s->cfg.output.block_size;
    return (uint64_t)ceil(s->sample_rate * 0.001);
}

/* reset, scenario.rs:636-643 */
void orc_scenario_reset(orc_scenario* s)
{
    s->current_sample = 0;
    s->rng_state = s->cfg.output.seed > 1 ? s->cfg.output.seed : 1;
    for (size_t k = 0; k < s->n; ++k) { s->doppler_phases[k] = 0.0; fir_reset(&s->lpfs[k]); }
}

/* xorshift64 / box_muller_pair, scenario.rs:688-704 */
static inline double xorshift64(uint64_t* st)
{
    uint64_t x = *st;
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    *st = x;
    return (double)x / 18446744073709551616.0; /* u64::MAX as f64 == 2^64 */
}
static inline void box_muller_pair(uint64_t* st, double* g1, double* g2)
{
    double u1 = xorshift64(st); if (!(u1 > 1e-15)) u1 = 1e-15;
    double u2 = xorshift64(st);
    double r = sqrt(-2.0 * log(u1));
    double theta = 2.0 * M_PI * u2;
    *g1 = r * cos(theta); *g2 = r * sin(theta);
}

/* noise_std, scenario.rs:532-537 */
double orc_scenario_noise_std(const orc_scenario* s)
{
    double noise_figure_linear = pow(10.0, s->cfg.receiver.noise_figure_db / 10.0);
    double n0 = 1.380649e-23 * 290.0 * noise_figure_linear;
    double noise_power = n0 * s->sample_rate;
    return sqrt(noise_power / 2.0) * 1e8;
}

/* receiver position / velocity for a block, scenario.rs:320-353 */
static void rx_state(const orc_scenario* s, double elapsed_s, orc_lla* lla, double* ecef, double* vel)
{
    const orc_receiver_cfg* r = &s->cfg.receiver;
    if (r->has_trajectory) {
        double duration_s = s->cfg.output.duration_s;
        double dist = traj_distance_m(&r->traj_start, &r->traj_end);
        double speed = r->traj_has_speed ? r->traj_speed_mps : dist / duration_s;
        double travel_time_s = dist / speed;
        double frac = elapsed_s / travel_time_s;
        if (frac < 0.0) frac = 0.0;
        if (frac > 1.0) frac = 1.0;
        *lla = traj_position_at(&r->traj_start, &r->traj_end, frac);
        orc_lla_to_ecef(lla, ecef);
        if (frac < 1.0) {
            double dt = fmin(0.01, travel_time_s * 0.001);
            double frac_dt = (elapsed_s + dt) / travel_time_s;
            if (frac_dt < 0.0) frac_dt = 0.0;
            if (frac_dt > 1.0) frac_dt = 1.0;
            orc_lla lla_dt = traj_position_at(&r->traj_start, &r->traj_end, frac_dt);
            double ecef_dt[3];
            orc_lla_to_ecef(&lla_dt, ecef_dt);
            for (int k = 0; k < 3; ++k) vel[k] = (ecef_dt[k] - ecef[k]) / dt;
        } else {
            vel[0] = vel[1] = vel[2] = 0.0;
        }
    } else {
        *lla = r->position;
        orc_lla_to_ecef(lla, ecef);
        vel[0] = vel[1] = vel[2] = 0.0;
    }
}

/* generate_block Phase 1, scenario.rs:378-454: parameters of every configured satellite for the block
 * [current_sample, current_sample + n) */
static void phase1(const orc_scenario* s, size_t n, orc_block_params* w)
{
    double t_offset = s->cfg.output.start_time_gps_s;
    double t_start = t_offset + (double)s->current_sample / s->sample_rate;
    double elapsed_s = (double)s->current_sample / s->sample_rate;
    orc_lla rx_lla; double rx_ecef[3], rx_vel[3];
    rx_state(s, elapsed_s, &rx_lla, rx_ecef, rx_vel);
    double t_end = t_start + (double)n / s->sample_rate;
    double elapsed_end = elapsed_s + (double)n / s->sample_rate;

    for (size_t k = 0; k < s->n; ++k) {
        const orc_sat_cfg* sc = &s->sats[k];
        const emitter_t* em = &s->emitters[k];
        orc_block_params* p = &w[k];
        memset(p, 0, sizeof *p);
        p->phase_before = s->doppler_phases[k];
        double ps[3], vs[3], pe[3], ve[3], la_el, la_az, la_range;
        orbit_pv(&em->orbit, t_start, ps, vs);
        orbit_pv(&em->orbit, t_end, pe, ve);
        orc_look_angle(rx_ecef, &rx_lla, ps, &la_el, &la_az, &la_range);
        double elevation_deg = (sc->has & ORC_HAS_ELEVATION) ? sc->elevation_deg : la_el;
        if (elevation_deg < s->cfg.receiver.elevation_mask_deg) { p->visible = 0; continue; }
        p->visible = 1;
        double carrier_hz = sig_carrier_hz(em->signal);
        double rr_orbital_start = orc_range_rate(rx_ecef, rx_vel, ps, vs);
        double rr_orbital_end = orc_range_rate(rx_ecef, rx_vel, pe, ve);
        double orbital_doppler_start = -rr_orbital_start * carrier_hz / SPEED_OF_LIGHT;
        double orbital_doppler_end = -rr_orbital_end * carrier_hz / SPEED_OF_LIGHT;

        double range_m;
        if (sc->orbital_dynamics) {
            if (sc->has & ORC_HAS_RANGE) range_m = sc->range_m + (la_range - s->orbital_range_t0[k]);
            else range_m = la_range;
        } else if ((sc->has & ORC_HAS_RANGE) && (sc->has & ORC_HAS_RANGE_RATE)) {
            range_m = sc->range_m + sc->range_rate_mps * elapsed_s;
        } else {
            range_m = (sc->has & ORC_HAS_RANGE) ? sc->range_m : la_range;
        }

        double d0, d1;
        if (sc->orbital_dynamics) {
            if (sc->has & ORC_HAS_DOPPLER) {
                double delta_start = orbital_doppler_start - s->orbital_doppler_t0[k];
                double delta_end = orbital_doppler_end - s->orbital_doppler_t0[k];
                d0 = sc->doppler_hz + delta_start; d1 = sc->doppler_hz + delta_end;
            } else { d0 = orbital_doppler_start; d1 = orbital_doppler_end; }
        } else if (sc->has & ORC_HAS_DOPPLER) {
            if (sc->has & ORC_HAS_DOPPLER_RATE) {
                d0 = sc->doppler_hz + sc->doppler_rate_hz_per_s * elapsed_s;
                d1 = sc->doppler_hz + sc->doppler_rate_hz_per_s * elapsed_end;
            } else { d0 = sc->doppler_hz; d1 = sc->doppler_hz; }
        } else if (sc->has & ORC_HAS_RANGE_RATE) {
            double doppler = -sc->range_rate_mps * carrier_hz / SPEED_OF_LIGHT;
            d0 = doppler; d1 = doppler;
        } else { d0 = orbital_doppler_start; d1 = orbital_doppler_end; }

        /* scenario.rs:430-439: the override, else SatelliteEmitter::status_at(t_start) (satellite_emitter.rs:165-178): the
         * models see the ORBIT's look angle (not the YAML elevation override) and GPS seconds of week; 0 m when disabled */
        const orc_environment_cfg* ev = &s->cfg.environment;
        const double DEG = 3.14159265358979323846 / 180.0;
        double iono_delay_m = 0.0, tropo_delay_m = 0.0;
        if (sc->has & ORC_HAS_IONO) iono_delay_m = sc->iono_delay_m;
        else if (ev->ionosphere_enabled)
            iono_delay_m = orc_klobuchar_delay_s(ev->klobuchar_alpha, ev->klobuchar_beta, la_el * DEG, la_az * DEG, rx_lla.lat_deg * DEG,
                                                 rx_lla.lon_deg * DEG, fmod(t_start, 604800.0)) * SPEED_OF_LIGHT;
        if (sc->has & ORC_HAS_TROPO) tropo_delay_m = sc->tropo_delay_m;
        else if (ev->troposphere_enabled)
            tropo_delay_m = orc_saastamoinen_delay_m(ev->tropo_height_m, ev->tropo_temperature_k, ev->tropo_pressure_hpa,
                                                     ev->tropo_relative_humidity, la_el * DEG);
        double cn0_dbhz;
        if (sc->has & ORC_HAS_CN0) cn0_dbhz = sc->cn0_dbhz;
        else {
            double fspl = orc_fspl_db(range_m, carrier_hz);
            double g = orc_antenna_gain_dbi(s->cfg.receiver.antenna, s->cfg.receiver.antenna_peak_gain_dbi,
                                            s->cfg.receiver.antenna_beamwidth_deg, elevation_deg);
            cn0_dbhz = sc->tx_power_dbw - fspl + g + 204.0;
        }
        double rx_power_dbw = cn0_dbhz - 204.0;
        p->range_m = range_m;
        p->iono_delay_s = iono_delay_m / SPEED_OF_LIGHT;
        p->tropo_delay_s = tropo_delay_m / SPEED_OF_LIGHT;
        p->rx_amplitude = pow(10.0, (rx_power_dbw + 160.0) / 20.0);
        p->doppler_start_hz = d0; p->doppler_end_hz = d1;
        /* informational: satellite_emitter.rs:228-242 */
        double total_delay_s = range_m / SPEED_OF_LIGHT + p->iono_delay_s + p->tropo_delay_s;
        double chips_delay = total_delay_s * sig_chipping_rate(em->signal);
        double code_length = (double)em->code_len;
        p->initial_code_phase = fmod(chips_delay, code_length);
        p->initial_epoch_offset = f64_as_usize(chips_delay / code_length);
    }
}

int orc_scenario_peek_params(orc_scenario* s, size_t block_size, orc_block_params* out, size_t cap)
{
    if (cap < s->n) return -1;
    uint64_t remaining = s->total_samples > s->current_sample ? s->total_samples - s->current_sample : 0;
    size_t n = remaining < block_size ? (size_t)remaining : block_size;
    if (n == 0) return 0;
    phase1(s, n, out);
    return (int)s->n;
}

/* generate_block, scenario.rs:308-546 */
size_t orc_scenario_generate_block(orc_scenario* s, size_t block_size, orc_c64* composite)
{
    uint64_t remaining = s->total_samples > s->current_sample ? s->total_samples - s->current_sample : 0;
    size_t n = remaining < block_size ? (size_t)remaining : block_size;
    if (n == 0) return 0;
    double oversampled_rate = s->sample_rate * (double)OS;
    size_t oversampled_n = n * OS;
    size_t oversampled_offset = (size_t)s->current_sample * OS;

    orc_block_params* work = (orc_block_params*)malloc(sizeof(orc_block_params) * (s->n ? s->n : 1));
    phase1(s, n, work);
    for (size_t i = 0; i < n; ++i) { composite[i].re = 0.0; composite[i].im = 0.0; }

    /* Phase 2 (scenario.rs:461-513): baseband at 8x, LPF, keep every 8th.  One buffer per satellite;
     * with threads > 1 satellites run concurrently (rayon `parallel` analogue; LPF state is per satellite). */
    double* baseband = (double*)malloc(sizeof(double) * n * (s->n ? s->n : 1));
    int threads = s->threads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) if (threads > 1)
    for (size_t k = 0; k < s->n; ++k) {
        if (!work[k].visible) continue;
        double* os_buf = (double*)malloc(sizeof(double) * oversampled_n);
        emitter_baseband(&s->emitters[k], oversampled_n, oversampled_rate, work[k].range_m, work[k].iono_delay_s,
                         work[k].tropo_delay_s, oversampled_offset, os_buf);
        fir_t* f = &s->lpfs[k];
        for (size_t g = 0; g < oversampled_n; ++g) os_buf[g] = fir_process(f, os_buf[g]);
        double* bb = baseband + k * n;
        for (size_t i = 0; i < n; ++i) bb[i] = os_buf[i * OS];
        free(os_buf);
    }

    /* Phase 3 (scenario.rs:516-528) */
    for (size_t k = 0; k < s->n; ++k) {
        if (!work[k].visible) continue;
        const double* bb = baseband + k * n;
        double n_f64 = (double)n;
        double phase = s->doppler_phases[k];
        double ds = work[k].doppler_start_hz, de = work[k].doppler_end_hz, amp = work[k].rx_amplitude;
        for (size_t i = 0; i < n; ++i) {
            double frac = (double)i / n_f64;
            double doppler_hz = ds + frac * (de - ds);
            double phase_inc = 2.0 * M_PI * doppler_hz / s->sample_rate;
            phase += phase_inc;
            double c = cos(phase), sn = sin(phase);
            /* sample * doppler_shift * rx_amplitude with sample = (bb, 0) */
            double pr = bb[i] * c - 0.0 * sn, pi = bb[i] * sn + 0.0 * c;
            composite[i].re += pr * amp;
            composite[i].im += pi * amp;
        }
        s->doppler_phases[k] = phase;
    }
    free(baseband);
    free(work);

    /* thermal noise (scenario.rs:530-542) */
    double noise_std = orc_scenario_noise_std(s);
    if (!(s->cfg.flags & ORC_FLAG_NOISE_OFF)) {
        for (size_t i = 0; i < n; ++i) {
            double g1, g2;
            box_muller_pair(&s->rng_state, &g1, &g2);
            composite[i].re += g1 * noise_std;
            composite[i].im += g2 * noise_std;
        }
    }
    s->current_sample += n;
    return n;
}

/* Fast-forward: same state evolution as generating every block of the canonical partition up to `sample`. */
int orc_scenario_skip_to(orc_scenario* s, uint64_t sample)
{
    uint64_t bs = orc_scenario_block_size(s);
    if (sample < s->current_sample || sample > s->total_samples) return -1;
    if ((sample - s->current_sample) % bs != 0 && sample != s->total_samples) return -2;
    orc_block_params* work = (orc_block_params*)malloc(sizeof(orc_block_params) * (s->n ? s->n : 1));
    double tail[FIR_TAPS];
    double oversampled_rate = s->sample_rate * (double)OS;
    while (s->current_sample < sample) {
        uint64_t remaining = s->total_samples - s->current_sample;
        size_t n = remaining < bs ? (size_t)remaining : (size_t)bs;
        phase1(s, n, work);
        for (size_t k = 0; k < s->n; ++k) {
            if (!work[k].visible) continue;
            double n_f64 = (double)n, phase = s->doppler_phases[k];
            double ds = work[k].doppler_start_hz, de = work[k].doppler_end_hz;
            for (size_t i = 0; i < n; ++i) {
                double frac = (double)i / n_f64;
                double doppler_hz = ds + frac * (de - ds);
                double phase_inc = 2.0 * M_PI * doppler_hz / s->sample_rate;
                phase += phase_inc;
            }
            s->doppler_phases[k] = phase;
            /* the delay line only remembers the last 63 inputs: push the block's last min(63, 8n) oversamples */
            size_t os_n = n * OS;
            size_t cnt = os_n < FIR_TAPS ? os_n : FIR_TAPS;
            emitter_baseband(&s->emitters[k], cnt, oversampled_rate, work[k].range_m, work[k].iono_delay_s,
                             work[k].tropo_delay_s, (size_t)s->current_sample * OS + (os_n - cnt), tail);
            fir_t* f = &s->lpfs[k];
            for (size_t g = 0; g < cnt; ++g) {
                f->delay_re[f->delay_idx] = tail[g];
                f->delay_idx = (f->delay_idx + 1) % FIR_TAPS;
            }
        }
        if (!(s->cfg.flags & ORC_FLAG_NOISE_OFF)) {
            for (size_t i = 0; i < 2 * n; ++i) (void)xorshift64(&s->rng_state);
        }
        s->current_sample += n;
    }
    free(work);
    return 0;
}

/* satellite_status, scenario.rs:564-633 (no SP3: clock correction 0) */
int orc_scenario_status(const orc_scenario* s, orc_sat_status* out, size_t cap)
{
    if (cap < s->n) return -1;
    double t = s->cfg.output.start_time_gps_s + (double)s->current_sample / s->sample_rate;
    const orc_lla* rx_lla = &s->cfg.receiver.position;
    double rx_ecef[3], zero[3] = {0.0, 0.0, 0.0};
    orc_lla_to_ecef(rx_lla, rx_ecef);
    for (size_t k = 0; k < s->n; ++k) {
        const orc_sat_cfg* sc = &s->sats[k];
        double sp[3], sv[3], el, az, rg;
        orbit_pv(&s->emitters[k].orbit, t, sp, sv);
        orc_look_angle(rx_ecef, rx_lla, sp, &el, &az, &rg);
        orc_sat_status* o = &out[k];
        memset(o, 0, sizeof *o);
        o->prn = sc->prn; o->signal = sc->signal;
        o->range_m = (sc->has & ORC_HAS_RANGE) ? sc->range_m : rg;
        o->elevation_deg = (sc->has & ORC_HAS_ELEVATION) ? sc->elevation_deg : el;
        o->azimuth_deg = (sc->has & ORC_HAS_AZIMUTH) ? sc->azimuth_deg : az;
        double rr = (sc->has & ORC_HAS_RANGE_RATE) ? sc->range_rate_mps : orc_range_rate(rx_ecef, zero, sp, sv);
        double carrier_hz = sig_carrier_hz(sc->signal);
        o->range_rate_mps = rr;
        o->doppler_hz = (sc->has & ORC_HAS_DOPPLER) ? sc->doppler_hz : -rr * carrier_hz / SPEED_OF_LIGHT;
        o->antenna_gain_dbi = orc_antenna_gain_dbi(s->cfg.receiver.antenna, s->cfg.receiver.antenna_peak_gain_dbi,
                                                   s->cfg.receiver.antenna_beamwidth_deg, o->elevation_deg);
        if (sc->has & ORC_HAS_CN0) o->cn0_dbhz = sc->cn0_dbhz;
        else o->cn0_dbhz = sc->tx_power_dbw - orc_fspl_db(o->range_m, carrier_hz) + o->antenna_gain_dbi + 204.0;
        /* scenario.rs:604-613: override, else the emitter's models at the ORBIT's look angle (status_at, t) */
        {
            const orc_environment_cfg* ev = &s->cfg.environment;
            const double DEG = 3.14159265358979323846 / 180.0;
            o->iono_delay_m = 0.0; o->tropo_delay_m = 0.0;
            if (sc->has & ORC_HAS_IONO) o->iono_delay_m = sc->iono_delay_m;
            else if (ev->ionosphere_enabled)
                o->iono_delay_m = orc_klobuchar_delay_s(ev->klobuchar_alpha, ev->klobuchar_beta, el * DEG, az * DEG, rx_lla->lat_deg * DEG,
                                                        rx_lla->lon_deg * DEG, fmod(t, 604800.0)) * SPEED_OF_LIGHT;
            if (sc->has & ORC_HAS_TROPO) o->tropo_delay_m = sc->tropo_delay_m;
            else if (ev->troposphere_enabled)
                o->tropo_delay_m = orc_saastamoinen_delay_m(ev->tropo_height_m, ev->tropo_temperature_k, ev->tropo_pressure_hpa,
                                                            ev->tropo_relative_humidity, el * DEG);
        }
        o->visible = o->elevation_deg > 0.0;
        o->clock_correction_s = 0.0;
    }
    return (int)s->n;
}

/* ------------------------------------------------------------------ FFT (stands in for rustfft 6.4.1, f64) */
/* FftProcessor::fft_inplace / ifft_inplace, core/fft_utils.rs:85-108: forward unscaled, inverse scaled by 1/N */
void orc_fft(orc_c64* a, size_t n, int inverse)
{
    if (n < 2) return;
    size_t lg = 0; while (((size_t)1 << lg) < n) ++lg;
    for (size_t i = 0, j = 0; i < n; ++i) {
        if (i < j) { orc_c64 t = a[i]; a[i] = a[j]; a[j] = t; }
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
    }
    double* tw = (double*)malloc(sizeof(double) * n);
    for (size_t k = 0; k < n / 2; ++k) {
        double ang = (inverse ? 2.0 : -2.0) * M_PI * (double)k / (double)n;
        tw[2 * k] = cos(ang); tw[2 * k + 1] = sin(ang);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        size_t half = len >> 1, step = n / len;
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < half; ++k) {
                double wr = tw[2 * k * step], wi = tw[2 * k * step + 1];
                orc_c64 u = a[i + k], v = a[i + k + half];
                double xr = v.re * wr - v.im * wi, xi = v.re * wi + v.im * wr;
                a[i + k].re = u.re + xr; a[i + k].im = u.im + xi;
                a[i + k + half].re = u.re - xr; a[i + k + half].im = u.im - xi;
            }
    }
    free(tw);
    if (inverse) {
        double scale = 1.0 / (double)n;
        for (size_t i = 0; i < n; ++i) { a[i].re *= scale; a[i].im *= scale; }
    }
}

/* ------------------------------------------------------------------ PCPS, gnss/acquisition.rs */
/* PcpsAcquisition::new, acquisition.rs:63-74 */
void orc_pcps_init(orc_pcps* p, uint64_t code_length, double sample_rate)
{
    uint64_t f = 1; while (f < code_length) f <<= 1; /* usize::next_power_of_two (0 -> 1) */
    p->fft_size = f; p->code_length = code_length;
    p->doppler_max_hz = 5000.0; p->doppler_step_hz = 500.0; p->threshold = 2.5;
    p->sample_rate = sample_rate; p->coherent_periods = 1;
}
int orc_pcps_num_bins(const orc_pcps* p) { return (int)(2.0 * p->doppler_max_hz / p->doppler_step_hz) + 1; }

/* shared loop of acquire (acquisition.rs:104-165) and acquire_grid (:199-242) */
static void pcps_search(const orc_pcps* p, const orc_c64* input, size_t n_input, const int8_t* code, size_t code_len,
                        double* power_out, double* best_peak_o, size_t* best_phase_o, double* best_doppler_o,
                        double* noise_sum_o, size_t* total_bins_o, int64_t* best_lin_o)
{
    size_t N = (size_t)p->fft_size, L = (size_t)p->code_length;
    orc_c64* code_fft = (orc_c64*)calloc(N, sizeof(orc_c64));
    for (size_t i = 0; i < code_len && i < N; ++i) code_fft[i].re = (double)code[i];
    orc_fft(code_fft, N, 0);
    for (size_t i = 0; i < N; ++i) code_fft[i].im = -code_fft[i].im;
    orc_c64* mixed = (orc_c64*)malloc(sizeof(orc_c64) * N);
    double best_peak = 0.0, best_doppler = 0.0, noise_floor = 0.0;
    size_t best_phase = 0, total_bins = 0; int64_t best_lin = -1;
    int bins = orc_pcps_num_bins(p);
    double doppler_start = -p->doppler_max_hz;
    size_t take = n_input < L ? n_input : L;
    for (int d = 0; d < bins; ++d) {
        double doppler = doppler_start + (double)d * p->doppler_step_hz;
        size_t m = 0;
        for (; m < take && m < N; ++m) {
            double t = (double)m / p->sample_rate;
            double phase = -2.0 * M_PI * doppler * t;
            double cr = cos(phase), ci = sin(phase);
            mixed[m].re = input[m].re * cr - input[m].im * ci;
            mixed[m].im = input[m].re * ci + input[m].im * cr;
        }
        for (; m < N; ++m) { mixed[m].re = 0.0; mixed[m].im = 0.0; }
        orc_fft(mixed, N, 0);
        for (size_t i = 0; i < N; ++i) {
            double xr = mixed[i].re * code_fft[i].re - mixed[i].im * code_fft[i].im;
            double xi = mixed[i].re * code_fft[i].im + mixed[i].im * code_fft[i].re;
            mixed[i].re = xr; mixed[i].im = xi;
        }
        orc_fft(mixed, N, 1);
        for (size_t ph = 0; ph < L && ph < N; ++ph) {
            double mag = mixed[ph].re * mixed[ph].re + mixed[ph].im * mixed[ph].im;
            if (power_out) power_out[(size_t)d * L + ph] = mag;
            total_bins += 1;
            noise_floor += mag;
            if (mag > best_peak) { best_peak = mag; best_phase = ph; best_doppler = doppler; best_lin = (int64_t)d * (int64_t)L + (int64_t)ph; }
        }
    }
    free(mixed); free(code_fft);
    *best_peak_o = best_peak; *best_phase_o = best_phase; *best_doppler_o = best_doppler;
    *noise_sum_o = noise_floor; *total_bins_o = total_bins; *best_lin_o = best_lin;
}

/* PcpsAcquisition::acquire, acquisition.rs:104-195 */
void orc_pcps_acquire(const orc_pcps* p, const orc_c64* input, size_t n_input, const int8_t* code, size_t code_len,
                      uint8_t prn, orc_acq_result* out)
{
    double best_peak, best_doppler, noise_floor; size_t best_phase, total_bins; int64_t lin;
    pcps_search(p, input, n_input, code, code_len, NULL, &best_peak, &best_phase, &best_doppler, &noise_floor, &total_bins, &lin);
    size_t denom = total_bins > 1 ? total_bins - 1 : 1; /* (total_bins - 1).max(1) */
    noise_floor = (noise_floor - best_peak) / (double)denom;
    double peak_metric = noise_floor > 0.0 ? best_peak / noise_floor : best_peak;
    int detected = peak_metric > p->threshold;
    memset(out, 0, sizeof *out);
    out->prn = prn; out->detected = (uint8_t)detected;
    out->code_phase = (double)best_phase; out->doppler_hz = best_doppler;
    out->peak_metric = peak_metric; out->threshold = p->threshold;
    if (detected) {
        double code_period = (double)p->code_length / p->sample_rate;
        out->has_cn0 = 1; out->cn0_estimate = 10.0 * log10(peak_metric / code_period);
    }
}

/* PcpsAcquisition::acquire_grid, acquisition.rs:199-249 */
int64_t orc_pcps_acquire_grid(const orc_pcps* p, const orc_c64* input, size_t n_input, const int8_t* code,
                              size_t code_len, double* power_out)
{
    double best_peak, best_doppler, noise_floor; size_t best_phase, total_bins; int64_t lin;
    pcps_search(p, input, n_input, code, code_len, power_out, &best_peak, &best_phase, &best_doppler, &noise_floor, &total_bins, &lin);
    return lin;
}

/* IqFormat::Cf32 write_sample, core/io/format.rs:197-200: `(x as f32)` rounds to nearest even */
/* ---- environment models ------------------------------------------------------------------------------ */
/* KlobucharModel::delay_seconds, environment/ionosphere.rs:46-108 (f64::powi(n) = repeated squaring) */
double orc_klobuchar_delay_s(const double alpha[4], const double beta[4], double elevation_rad, double azimuth_rad,
                             double user_lat_rad, double user_lon_rad, double gps_time_s)
{
    const double PI = 3.14159265358979323846;
    double el_sc = elevation_rad / PI, az_sc = azimuth_rad / PI, lat_sc = user_lat_rad / PI, lon_sc = user_lon_rad / PI;
    double psi = 0.0137 / (el_sc + 0.11) - 0.022;
    double lat_ipp = lat_sc + psi * cos(az_sc) * PI;          /* sic: cos of the semicircle value, times pi (:60) */
    if (lat_ipp > 0.416) lat_ipp = 0.416;
    if (lat_ipp < -0.416) lat_ipp = -0.416;
    double lon_ipp = lon_sc + psi * sin(az_sc * PI) / cos(lat_ipp * PI);
    double lat_mag = lat_ipp + 0.064 * cos(lon_ipp - 1.617);  /* sic: no pi here (:72) */
    double t_local = 43200.0 * lon_ipp + gps_time_s;
    t_local = fmod(t_local, 86400.0);
    double d = 0.53 - el_sc;
    double f_obl = 1.0 + 16.0 * (d * d * d);
    double lm2 = lat_mag * lat_mag, lm3 = lm2 * lat_mag;
    double amp = alpha[0] + alpha[1] * lat_mag + alpha[2] * lm2 + alpha[3] * lm3;
    if (!(amp > 0.0)) amp = 0.0;                              /* f64::max(0.0) */
    double per = beta[0] + beta[1] * lat_mag + beta[2] * lm2 + beta[3] * lm3;
    if (!(per > 72000.0)) per = 72000.0;
    double x = 2.0 * PI * (t_local - 50400.0) / per;
    if (fabs(x) < 1.57) {
        double x2 = x * x, x4 = x2 * x2;
        return f_obl * (5.0e-9 + amp * (1.0 - x * x / 2.0 + x4 / 24.0));
    }
    return f_obl * 5.0e-9;
}

/* SaastamoinenModel, environment/troposphere.rs:52-97.  component: 0 = hydrostatic, 1 = wet, 2 = total */
double orc_saastamoinen_zenith_m(double height_m, double temperature_k, double pressure_hpa, double relative_humidity, int component)
{
    const double PI = 3.14159265358979323846;
    double h_rad = height_m * (PI / 180.0);                   /* sic: height_m.to_radians() (:54) */
    double dry = 0.002277 * pressure_hpa / (1.0 - 0.00266 * cos(2.0 * h_rad) - 0.00028 * height_m / 1000.0);
    double t_c = temperature_k - 273.15;
    double es = 6.1121 * exp((18.678 - t_c / 234.5) * t_c / (257.14 + t_c));
    double e = relative_humidity * es;
    double wet = 0.002277 * (1255.0 / temperature_k + 0.05) * e;
    return component == 0 ? dry : component == 1 ? wet : dry + wet;
}
double orc_saastamoinen_delay_m(double height_m, double temperature_k, double pressure_hpa, double relative_humidity, double elevation_rad)
{
    double el = elevation_rad > 0.05 ? elevation_rad : 0.05;  /* f64::max(0.05) */
    double sin_el = sin(el);
    double mapping = 1.0 / (sin_el + 0.00143 / tan(0.0455 + sin_el));
    return orc_saastamoinen_zenith_m(height_m, temperature_k, pressure_hpa, relative_humidity, 2) * mapping;
}

/* IqFormat::write_sample for the integer sink formats (core/io/format.rs:203-222): f64 scale, clamp, Rust `as`
 * (truncation toward zero; NaN -> 0).  fmt: 2 = ci16, 3 = ci8, 4 = cu8 (the r4wb_fmt numbering). */
static double orc_clamp(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }
int orc_to_int_format(const orc_c64* in, size_t n, int fmt, void* out)
{
    for (size_t i = 0; i < n; ++i) {
        const double c[2] = {in[i].re, in[i].im};
        for (int k = 0; k < 2; ++k) {
            const double x = c[k];
            if (fmt == 2) {
                const double v = orc_clamp(x * 32767.0, -32768.0, 32767.0);
                ((int16_t*)out)[2 * i + k] = (x != x) ? 0 : (int16_t)v;
            } else if (fmt == 3) {
                const double v = orc_clamp(x * 127.0, -128.0, 127.0);
                ((int8_t*)out)[2 * i + k] = (x != x) ? 0 : (int8_t)v;
            } else if (fmt == 4) {
                const double v = orc_clamp((x + 1.0) * 127.5, 0.0, 255.0);
                ((uint8_t*)out)[2 * i + k] = (x != x) ? 0 : (uint8_t)v;
            } else {
                return -1;
            }
        }
    }
    return 0;
}

void orc_to_cf32(const orc_c64* in, size_t n, float* out)
{
    for (size_t i = 0; i < n; ++i) { out[2 * i] = (float)in[i].re; out[2 * i + 1] = (float)in[i].im; }
}
