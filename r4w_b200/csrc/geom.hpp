// geom.hpp — per-block receiver/satellite geometry in f64, callable from host and device.
//
// Product code (libr4w_b200.so).  Re-derivation of the reference's "Phase 1" inputs:
//   WGS-84 LLA->ECEF, look angle, range rate, FSPL     core/coordinates.rs:129-144, 191-246
//   Keplerian two-body propagation -> ECEF pos/vel     gnss/environment/orbit.rs:49-119, 125-199
//   great-circle receiver trajectory                   gnss/scenario_config.rs:319-356
//   antenna gain patterns                              gnss/environment/antenna.rs:35-78
// f64 is mandatory: the Earth-rotation angle is OMEGA_E * t with t ~ 1.44e9 s.
#pragma once
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define R4WB_HD __host__ __device__ __forceinline__
#else
#define R4WB_HD inline
#endif

namespace r4wb {

constexpr double kPi = 3.14159265358979323846;
constexpr double kC = 299792458.0;            // speed of light, m/s
constexpr double kGmEarth = 3.986004418e14;   // m^3/s^2
constexpr double kOmegaE = 7.2921150e-5;      // rad/s
constexpr double kWgsA = 6378137.0;
constexpr double kWgsF = 1.0 / 298.257223563;
constexpr double kWgsE2 = 2.0 * kWgsF - kWgsF * kWgsF;
constexpr double kDeg = kPi / 180.0;          // f64::to_radians multiplier
constexpr double kRad = 180.0 / kPi;          // f64::to_degrees multiplier

struct Vec3 { double x, y, z; };
struct Lla { double lat_deg, lon_deg, alt_m; };
struct Look { double elevation_deg, azimuth_deg, range_m; };

R4WB_HD Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }

R4WB_HD Vec3 ecef_of(const Lla& p)
{
    const double lat = p.lat_deg * kDeg, lon = p.lon_deg * kDeg;
    const double sl = sin(lat), cl = cos(lat), so = sin(lon), co = cos(lon);
    const double n = kWgsA / sqrt(1.0 - kWgsE2 * sl * sl);
    return {(n + p.alt_m) * cl * co, (n + p.alt_m) * cl * so, (n * (1.0 - kWgsE2) + p.alt_m) * sl};
}

// elevation / azimuth / slant range of `tgt` seen from `obs` (ENU rotation at the observer's geodetic lat/lon)
R4WB_HD Look look_from(const Vec3& obs, const Lla& obs_lla, const Vec3& tgt)
{
    const Vec3 d = tgt - obs;
    const double lat = obs_lla.lat_deg * kDeg, lon = obs_lla.lon_deg * kDeg;
    const double sl = sin(lat), cl = cos(lat), so = sin(lon), co = cos(lon);
    const double e = -so * d.x + co * d.y;
    const double n = -sl * co * d.x - sl * so * d.y + cl * d.z;
    const double u = cl * co * d.x + cl * so * d.y + sl * d.z;
    Look r;
    r.range_m = sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
    r.elevation_deg = atan2(u, sqrt(e * e + n * n)) * kRad;
    double az = atan2(e, n) * kRad;
    if (az < 0.0) az += 360.0;
    r.azimuth_deg = az;
    return r;
}

// d(range)/dt: relative velocity projected on the line of sight (positive = receding)
R4WB_HD double los_rate(const Vec3& op, const Vec3& ov, const Vec3& tp, const Vec3& tv)
{
    const Vec3 d = tp - op;
    const double r = sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
    if (r < 1e-10) return 0.0;
    const Vec3 rv = tv - ov;
    return rv.x * (d.x / r) + rv.y * (d.y / r) + rv.z * (d.z / r);
}

// KlobucharModel::delay_seconds, gnss/environment/ionosphere.rs:46-108 (semicircle quirks of the reference kept as written)
R4WB_HD double klobuchar_delay_s(const double* alpha, const double* beta, double elevation_rad, double azimuth_rad, double lat_rad,
                                 double lon_rad, double gps_time_s)
{
    const double el_sc = elevation_rad / kPi, az_sc = azimuth_rad / kPi, lat_sc = lat_rad / kPi, lon_sc = lon_rad / kPi;
    const double psi = 0.0137 / (el_sc + 0.11) - 0.022;
    double lat_ipp = lat_sc + psi * cos(az_sc) * kPi;
    if (lat_ipp > 0.416) lat_ipp = 0.416;
    if (lat_ipp < -0.416) lat_ipp = -0.416;
    const double lon_ipp = lon_sc + psi * sin(az_sc * kPi) / cos(lat_ipp * kPi);
    const double lat_mag = lat_ipp + 0.064 * cos(lon_ipp - 1.617);
    const double t_local = fmod(43200.0 * lon_ipp + gps_time_s, 86400.0);
    const double d = 0.53 - el_sc;
    const double f_obl = 1.0 + 16.0 * (d * d * d);
    const double lm2 = lat_mag * lat_mag, lm3 = lm2 * lat_mag;
    double amp = alpha[0] + alpha[1] * lat_mag + alpha[2] * lm2 + alpha[3] * lm3;
    if (!(amp > 0.0)) amp = 0.0;
    double per = beta[0] + beta[1] * lat_mag + beta[2] * lm2 + beta[3] * lm3;
    if (!(per > 72000.0)) per = 72000.0;
    const double x = 2.0 * kPi * (t_local - 50400.0) / per;
    if (fabs(x) < 1.57) {
        const double x2 = x * x;
        return f_obl * (5.0e-9 + amp * (1.0 - x * x / 2.0 + (x2 * x2) / 24.0));
    }
    return f_obl * 5.0e-9;
}

// SaastamoinenModel::delay_meters, gnss/environment/troposphere.rs:52-97
R4WB_HD double saastamoinen_delay_m(double height_m, double temperature_k, double pressure_hpa, double relative_humidity, double elevation_rad)
{
    const double dry = 0.002277 * pressure_hpa / (1.0 - 0.00266 * cos(2.0 * (height_m * kDeg)) - 0.00028 * height_m / 1000.0);
    const double t_c = temperature_k - 273.15;
    const double es = 6.1121 * exp((18.678 - t_c / 234.5) * t_c / (257.14 + t_c));
    const double wet = 0.002277 * (1255.0 / temperature_k + 0.05) * (relative_humidity * es);
    const double el = elevation_rad > 0.05 ? elevation_rad : 0.05;
    const double sin_el = sin(el);
    return (dry + wet) * (1.0 / (sin_el + 0.00143 / tan(0.0455 + sin_el)));
}

R4WB_HD double fspl_db(double dist_m, double freq_hz)
{
    if (dist_m <= 0.0 || freq_hz <= 0.0) return 0.0;
    return 20.0 * log10(4.0 * kPi * dist_m * freq_hz / kC);
}

struct Orbit { double a, e, inc, raan0, argp, m0, t_epoch, raan_dot; };

R4WB_HD double kepler_E(double m, double e)
{
    double E = m;
    for (int k = 0; k < 20; ++k) {
        const double dE = (E - e * sin(E) - m) / (1.0 - e * cos(E));
        E -= dE;
        if (fabs(dE) < 1e-14) break;
    }
    return E;
}

// ECEF position and velocity at GPS time t (two-body, Earth rotation theta = OMEGA_E * t)
R4WB_HD void orbit_state(const Orbit& o, double t, Vec3& pos, Vec3& vel)
{
    const double dt = t - o.t_epoch;
    const double n = sqrt(kGmEarth / (o.a * o.a * o.a));
    const double M = fmod(o.m0 + n * dt, 2.0 * kPi);
    const double E = kepler_E(M, o.e);
    const double sE = sin(E), cE = cos(E);
    const double nu = atan2(sqrt(1.0 - o.e * o.e) * sE, cE - o.e);
    const double r = o.a * (1.0 - o.e * cE);
    const double snu = sin(nu), cnu = cos(nu);
    const double xo = r * cnu, yo = r * snu;
    const double h = sqrt(kGmEarth * o.a * (1.0 - o.e * o.e));
    const double vxo = -kGmEarth / h * snu, vyo = kGmEarth / h * (o.e + cnu);
    const double raan = o.raan0 + o.raan_dot * dt;
    const double cw = cos(o.argp), sw = sin(o.argp), cO = cos(raan), sO = sin(raan), ci = cos(o.inc), si = sin(o.inc);
    const double r11 = cO * cw - sO * sw * ci, r12 = -cO * sw - sO * cw * ci;
    const double r21 = sO * cw + cO * sw * ci, r22 = -sO * sw + cO * cw * ci;
    const double r31 = sw * si, r32 = cw * si;
    const double xi = r11 * xo + r12 * yo, yi = r21 * xo + r22 * yo, zi = r31 * xo + r32 * yo;
    const double vxi = r11 * vxo + r12 * vyo, vyi = r21 * vxo + r22 * vyo, vzi = r31 * vxo + r32 * vyo;
    const double th = kOmegaE * t, ct = cos(th), st = sin(th);
    pos.x = ct * xi + st * yi;
    pos.y = -st * xi + ct * yi;
    pos.z = zi;
    vel.x = ct * vxi + st * vyi + kOmegaE * pos.y;
    vel.y = -st * vxi + ct * vyi - kOmegaE * pos.x;
    vel.z = vzi;
}

inline Orbit nominal_orbit(uint32_t signal, int plane, int slot)
{
    // signal enum: 0 GpsL1Ca, 1 GpsL5, 2 GlonassL1of, 3..5 Galileo (gnss/scenario.rs:718-725)
    if (signal >= 3) {   // Galileo Walker 24/3/1 with the reference's RAAN / M0 calibration offsets
        return {29600318.0, 0.0, 56.0 * kDeg, plane * (120.0 * kDeg) + 118.0 * kDeg, 0.0,
                slot * (45.0 * kDeg) + plane * (15.0 * kDeg) + 176.0 * kDeg, 0.0, 0.0};
    }
    if (signal == 2) return {25508000.0, 0.0, 64.8 * kDeg, plane * (120.0 * kDeg), 0.0, slot * (45.0 * kDeg), 0.0, 0.0};
    return {26559700.0, 0.0, 55.0 * kDeg, plane * (60.0 * kDeg), 0.0, slot * (60.0 * kDeg), 0.0, 0.0};
}

R4WB_HD double antenna_gain_dbi(uint32_t kind, double peak, double beamwidth_deg, double el_deg)
{
    if (kind == 0) return 0.0;
    if (kind == 1) return el_deg >= 0.0 ? peak : -30.0;
    if (kind == 2) {
        if (el_deg < -5.0) return -30.0;
        const double theta = (90.0 - el_deg) * kDeg;
        const double n = log10(3.0) / log10(1.0 / cos(beamwidth_deg / 2.0 * kDeg));
        const double g = pow(fabs(cos(theta)), n);
        return peak + (g > 1e-6 ? 10.0 * log10(g) : -30.0);
    }
    if (el_deg < 0.0) return -40.0;
    const double g = pow(fabs(cos((90.0 - el_deg) * kDeg)), 1.5);
    return peak + (g > 1e-6 ? 10.0 * log10(g) : -40.0);
}

// great-circle slerp between two LLA points, altitude linear
R4WB_HD double gc_angle(const Lla& a, const Lla& b)
{
    const double la1 = a.lat_deg * kDeg, lo1 = a.lon_deg * kDeg, la2 = b.lat_deg * kDeg, lo2 = b.lon_deg * kDeg;
    const double s1 = sin((la2 - la1) / 2.0), s2 = sin((lo2 - lo1) / 2.0);
    return 2.0 * asin(sqrt(s1 * s1 + cos(la1) * cos(la2) * (s2 * s2)));
}
R4WB_HD Lla gc_point(const Lla& a, const Lla& b, double frac)
{
    frac = frac < 0.0 ? 0.0 : (frac > 1.0 ? 1.0 : frac);
    const double la1 = a.lat_deg * kDeg, lo1 = a.lon_deg * kDeg, la2 = b.lat_deg * kDeg, lo2 = b.lon_deg * kDeg;
    const double ang = gc_angle(a, b);
    double lat = la1, lon = lo1;
    if (!(fabs(ang) < 1e-12)) {
        const double ka = sin((1.0 - frac) * ang) / sin(ang), kb = sin(frac * ang) / sin(ang);
        const double x = ka * cos(la1) * cos(lo1) + kb * cos(la2) * cos(lo2);
        const double y = ka * cos(la1) * sin(lo1) + kb * cos(la2) * sin(lo2);
        const double z = ka * sin(la1) + kb * sin(la2);
        lat = atan2(z, sqrt(x * x + y * y));
        lon = atan2(y, x);
    }
    return {lat * kRad, lon * kRad, a.alt_m + frac * (b.alt_m - a.alt_m)};
}

struct RxModel {
    Lla position;
    int has_trajectory;
    Lla traj_start, traj_end;
    double travel_time_s;   // distance / speed (speed defaults to distance / duration)
    double fd_dt;           // finite-difference step min(0.01, travel_time * 0.001)
};

struct RxState { Lla lla; Vec3 pos, vel; };

// receiver position/velocity at `elapsed_s` into the scenario (gnss/scenario.rs:320-353)
R4WB_HD RxState rx_at(const RxModel& m, double elapsed_s)
{
    RxState s;
    if (!m.has_trajectory) {
        s.lla = m.position;
        s.pos = ecef_of(s.lla);
        s.vel = {0.0, 0.0, 0.0};
        return s;
    }
    double frac = elapsed_s / m.travel_time_s;
    frac = frac < 0.0 ? 0.0 : (frac > 1.0 ? 1.0 : frac);
    s.lla = gc_point(m.traj_start, m.traj_end, frac);
    s.pos = ecef_of(s.lla);
    if (frac < 1.0) {
        const Vec3 p2 = ecef_of(gc_point(m.traj_start, m.traj_end, (elapsed_s + m.fd_dt) / m.travel_time_s));
        s.vel = {(p2.x - s.pos.x) / m.fd_dt, (p2.y - s.pos.y) / m.fd_dt, (p2.z - s.pos.z) / m.fd_dt};
    } else {
        s.vel = {0.0, 0.0, 0.0};
    }
    return s;
}

}  // namespace r4wb
