/*
 * r4w_oracle_sim.c — CPU restatement (f64, plain C) of r4w-sim's generic scenario engine.
 * TEST INFRASTRUCTURE ONLY (see r4w_oracle.h).  Follows crates/r4w-sim/src/scenario/:
 *   engine.rs:61-137    ScenarioEngine::generate_block — geometry at the block midpoint, Doppler rotation with the
 *                       continuously accumulated (and wrapped) carrier phase, amplitude, sum; the AWGN draw is NOT
 *                       restated (rand's StdRng / rand_distr::Normal are un-vendored crates; the noise is statistical)
 *   config.rs:41-58     total_samples, noise_power_linear
 *   trajectory.rs:45-170 Trajectory::state_at (Static, Linear, Waypoints, Circular)
 * and crates/r4w-core/src/coordinates.rs:129-144 (lla_to_ecef), :225-246 (range_rate, fspl_db).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#include "r4w_oracle.h"

#define ORC_PI 3.14159265358979323846
#define ORC_C 299792458.0

/* the per-sample loops of engine.rs:105-122 for one emitter: phase accumulates before use, wraps once past 1e6 */
void orc_sim_compose_emitter(const orc_c64* baseband, size_t n, double doppler_hz, double sample_rate, double rx_amplitude,
                             double* carrier_phase, orc_c64* composite)
{
    double ph = *carrier_phase;
    for (size_t i = 0; i < n; ++i) {
        ph += 2.0 * ORC_PI * doppler_hz / sample_rate;
        if (fabs(ph) > 1e6) ph = fmod(ph, 2.0 * ORC_PI);
        const double c = cos(ph), s = sin(ph);
        /* sample * doppler_shift * rx_amplitude (num-complex: (a + jb)(c + js), then scaled) */
        const double re = baseband[i].re * c - baseband[i].im * s, im = baseband[i].re * s + baseband[i].im * c;
        composite[i].re += re * rx_amplitude;
        composite[i].im += im * rx_amplitude;
    }
    *carrier_phase = ph;
}

/* geometry of one emitter at the block midpoint, engine.rs:82-98: out = {range_m, doppler_hz, path_loss_db, rx_amplitude} */
void orc_sim_link(const double* rx_pos, const double* rx_vel, const double* em_pos, const double* em_vel, double carrier_hz,
                  double power_dbm, double* out4)
{
    const double dx = em_pos[0] - rx_pos[0], dy = em_pos[1] - rx_pos[1], dz = em_pos[2] - rx_pos[2];
    const double range_m = sqrt(dx * dx + dy * dy + dz * dz);
    const double rr = orc_range_rate(rx_pos, rx_vel, em_pos, em_vel);
    const double pl_db = orc_fspl_db(range_m, carrier_hz);
    out4[0] = range_m;
    out4[1] = -rr * carrier_hz / ORC_C;
    out4[2] = pl_db;
    out4[3] = pow(10.0, ((power_dbm - pl_db) - 30.0) / 20.0);
}

/* ScenarioConfig::noise_power_linear, config.rs:52-57 */
double orc_sim_noise_power(double noise_floor_dbw_hz, double sample_rate) { return pow(10.0, noise_floor_dbw_hz / 10.0) * sample_rate; }

/* Trajectory::state_at.  kind 0 Static {lla}, 1 Linear {lla, venu[3]}, 3 Circular {lla, radius, omega, bearing_deg}; p = parameters,
 * out6 = ECEF position and velocity.  (Waypoints, kind 2: orc_sim_waypoints.) */
void orc_sim_trajectory(int kind, const double* p, double t, double* out6)
{
    orc_lla lla = {p[0], p[1], p[2]};
    double e[3];
    orc_lla_to_ecef(&lla, e);
    const double lat = lla.lat_deg * ORC_PI / 180.0, lon = lla.lon_deg * ORC_PI / 180.0;
    const double sin_lat = sin(lat), cos_lat = cos(lat), sin_lon = sin(lon), cos_lon = cos(lon);
    if (kind == 0) {
        out6[0] = e[0]; out6[1] = e[1]; out6[2] = e[2]; out6[3] = out6[4] = out6[5] = 0.0;
    } else if (kind == 1) {
        const double ve = p[3], vn = p[4], vu = p[5];
        const double vx = -sin_lon * ve - sin_lat * cos_lon * vn + cos_lat * cos_lon * vu;
        const double vy = cos_lon * ve - sin_lat * sin_lon * vn + cos_lat * sin_lon * vu;
        const double vz = cos_lat * vn + sin_lat * vu;
        out6[0] = e[0] + vx * t; out6[1] = e[1] + vy * t; out6[2] = e[2] + vz * t;
        out6[3] = vx; out6[4] = vy; out6[5] = vz;
    } else {
        const double radius = p[3], omega = p[4];
        const double bearing = p[5] * ORC_PI / 180.0 + omega * t;
        const double east = radius * sin(bearing), north = radius * cos(bearing);
        out6[0] = e[0] + (-sin_lon * east - sin_lat * cos_lon * north);
        out6[1] = e[1] + (cos_lon * east - sin_lat * sin_lon * north);
        out6[2] = e[2] + (cos_lat * north);
        const double d_east = radius * omega * cos(bearing), d_north = -radius * omega * sin(bearing);
        out6[3] = -sin_lon * d_east - sin_lat * cos_lon * d_north;
        out6[4] = cos_lon * d_east - sin_lat * sin_lon * d_north;
        out6[5] = cos_lat * d_north;
    }
}

/* Trajectory::Waypoints, trajectory.rs:86-128: points = n rows of {t, lat, lon, alt}, linear interpolation in ECEF */
void orc_sim_waypoints(const double* points, size_t n, double t, double* out6)
{
    for (int k = 0; k < 6; ++k) out6[k] = 0.0;
    if (n == 0) return;
    orc_lla a, b;
    if (n == 1 || t <= points[0]) { a = (orc_lla){points[1], points[2], points[3]}; orc_lla_to_ecef(&a, out6); return; }
    if (t >= points[4 * (n - 1)]) { a = (orc_lla){points[4 * (n - 1) + 1], points[4 * (n - 1) + 2], points[4 * (n - 1) + 3]}; orc_lla_to_ecef(&a, out6); return; }
    size_t idx = n;
    for (size_t k = 0; k < n; ++k) if (points[4 * k] > t) { idx = k; break; }
    idx -= 1;
    const double t0 = points[4 * idx], t1 = points[4 * (idx + 1)];
    a = (orc_lla){points[4 * idx + 1], points[4 * idx + 2], points[4 * idx + 3]};
    b = (orc_lla){points[4 * (idx + 1) + 1], points[4 * (idx + 1) + 2], points[4 * (idx + 1) + 3]};
    double e0[3], e1[3];
    orc_lla_to_ecef(&a, e0);
    orc_lla_to_ecef(&b, e1);
    const double dt = t1 - t0, frac = (t - t0) / dt;
    for (int k = 0; k < 3; ++k) { out6[k] = e0[k] + (e1[k] - e0[k]) * frac; out6[3 + k] = (e1[k] - e0[k]) / dt; }
}
