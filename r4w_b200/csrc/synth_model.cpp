// synth_model.cpp — host-only model of GnssScenario::new (gnss/scenario.rs:78-237): constants, code bits,
// filter tables, the piecewise-exact Doppler phase table, sequential-block bookkeeping, satellite_status.
// No CUDA runtime calls in this file.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "synth_math.cuh"

namespace r4wb {

static const uint8_t kE1Packed[2 * 50 * 512] = {
#include "../../data/galileo_e1_codes.inc"
};

// chip i of PRN `prn` on `channel` (0 = E1B, 1 = E1C) as a sign bit (1 <=> -1):
// MSB-first packing, bit 0 -> +1 (gnss/galileo_e1_codes.rs:17-25)
static inline uint32_t e1_sign_bit(uint32_t channel, uint32_t prn, uint32_t i)
{
    const uint8_t* p = kE1Packed + ((size_t)channel * 50 + (prn - 1)) * 512;
    return (p[i >> 3] >> (7 - (i & 7))) & 1u;
}

void e1_code_chips(uint32_t channel, uint32_t prn, int8_t* out)
{
    for (uint32_t i = 0; i < (uint32_t)kCodeLen; ++i) out[i] = e1_sign_bit(channel, prn, i) ? -1 : 1;
}

// GPS L1 C/A Gold code (gnss/prn.rs:34-162: G1 = x^10 + x^3 + 1, G2 = x^10 + x^9 + x^8 + x^6 + x^3 + x^2 + 1, both all-ones,
// G2 read through the PRN's two phase-selector taps; bit 0 -> chip +1)
static const uint8_t kGpsCaTaps[32][2] = {{2, 6}, {3, 7}, {4, 8}, {5, 9}, {1, 9}, {2, 10}, {1, 8}, {2, 9}, {3, 10}, {2, 3}, {3, 4},
                                          {5, 6}, {6, 7}, {7, 8}, {8, 9}, {9, 10}, {1, 4}, {2, 5}, {3, 6}, {4, 7}, {5, 8}, {6, 9},
                                          {1, 3}, {4, 6}, {5, 7}, {6, 8}, {7, 9}, {8, 10}, {1, 6}, {2, 7}, {3, 8}, {4, 9}};
void gps_ca_code_chips(uint32_t prn, int8_t* out /*[1023]*/)
{
    if (prn < 1 || prn > 32) fail(R4WB_ERR_INVALID_PARAMETER, "GPS PRN must be 1-32, got %u", prn);
    uint32_t g1[11], g2[11];                     // stage i = 1..10
    for (int i = 1; i <= 10; ++i) g1[i] = g2[i] = 1u;
    const int ta = kGpsCaTaps[prn - 1][0], tb = kGpsCaTaps[prn - 1][1];
    for (int n = 0; n < 1023; ++n) {
        const uint32_t bit = g1[10] ^ g2[ta] ^ g2[tb];
        out[n] = bit ? -1 : 1;
        const uint32_t f1 = g1[3] ^ g1[10];
        const uint32_t f2 = g2[2] ^ g2[3] ^ g2[6] ^ g2[8] ^ g2[9] ^ g2[10];
        for (int i = 10; i > 1; --i) { g1[i] = g1[i - 1]; g2[i] = g2[i - 1]; }
        g1[1] = f1; g2[1] = f2;
    }
}

// Shift-register sequences in the product's own form (bit n of `taps` set <=> stage n+1 is fed back; the output is the last
// stage; the register shifts toward the last stage) — what Lfsr::clock (core/spreading/lfsr.rs:58-72) computes.
static uint32_t shift_register_step(uint32_t& state, uint32_t taps, int stages)
{
    const uint32_t out = (state >> (stages - 1)) & 1u;
    uint32_t fb = state & taps;
    fb ^= fb >> 16; fb ^= fb >> 8; fb ^= fb >> 4; fb ^= fb >> 2; fb ^= fb >> 1;
    state = ((state << 1) | (fb & 1u)) & ((1u << stages) - 1u);
    return out;
}

// GLONASS L1OF ranging code (gnss/prn.rs:170-216): 9 stages, 1 + x^5 + x^9, all ones; 511 chips, identical for every satellite
void glonass_code_chips(int8_t* out /*[511]*/)
{
    uint32_t r = 0x1FFu;
    for (int n = 0; n < 511; ++n) out[n] = shift_register_step(r, 0x110u, 9) ? -1 : 1;
}

// GPS L5 I5 code as the reference generates it (gnss/prn.rs:376-397): XA (x^13 + x^12 + x^10 + x^9 + 1, all ones) xor XB
// (I5 taps 0x1AE3, start state derived from the PRN), 10 230 chips, no short cycle
void gps_l5_i5_code_chips(uint32_t prn, int8_t* out /*[10230]*/)
{
    if (prn < 1 || prn > 32) fail(R4WB_ERR_INVALID_PARAMETER, "GPS L5 PRN must be 1-32, got %u", prn);
    uint32_t xa = 0x1FFFu, xb = (prn * 0x2468u + 0x1357u) & 0x1FFFu;
    if (xb == 0u) xb = 1u;
    for (int n = 0; n < 10230; ++n) {
        const uint32_t a = shift_register_step(xa, 0x1E01u, 13), b = shift_register_step(xb, 0x1AE3u, 13);
        out[n] = (a ^ b) ? -1 : 1;
    }
}

// 63-tap Blackman windowed-sinc low-pass, unity DC gain (core/filters/fir.rs:458-499, windows.rs:137-151)
static void design_lowpass(double cutoff_hz, double rate_hz, double* h /*[63]*/)
{
    const double fc = cutoff_hz / rate_hz, mid = (kTaps - 1) / 2.0;
    double sum = 0.0;
    for (int i = 0; i < kTaps; ++i) {
        const double x = 2.0 * kPi * (double)i / (double)(kTaps - 1);
        const double w = 0.42 - 0.5 * cos(x) + 0.08 * cos(2.0 * x);
        const double n = (double)i - mid;
        const double sinc = fabs(n) < 1e-10 ? 2.0 * kPi * fc : sin(2.0 * kPi * fc * n) / n;
        h[i] = sinc * w;
    }
    for (int i = 0; i < kTaps; ++i) sum += h[i];
    if (fabs(sum) > 1e-10)
        for (int i = 0; i < kTaps; ++i) h[i] /= sum;
}

// Piecewise-exact model of `phase += inc` repeated `steps` times in f64 (gnss/scenario.rs:518-527):
// inside one binade every add moves the phase by the same multiple of the binade's ulp.
static void build_phase_segments(double inc, uint64_t steps, std::vector<PhaseSegment>& out)
{
    double x = 0.0;
    uint64_t i = 0;
    auto push = [&](uint64_t i0, double x0, double step) { out.push_back(PhaseSegment{i0, x0, step}); };
    if (inc == 0.0 || steps == 0) { out.push_back(PhaseSegment{0, 0.0, inc}); return; }
    while (i < steps) {
        const double x1 = x + inc, s1 = x1 - x;
        const double x2 = x1 + inc, s2 = x2 - x1;
        if (x != 0.0 && s1 == s2 && std::ilogb(x) == std::ilogb(x1) && std::ilogb(x1) == std::ilogb(x2)) {
            const int e = std::ilogb(x);
            const long double limit = std::ldexp(1.0L, e + 1);
            const long double room = (limit - fabsl((long double)x)) / fabsl((long double)s1);
            long long fit = (long long)floorl(room) - 2;
            if (fit >= 3) {
                const uint64_t n = std::min<uint64_t>((uint64_t)fit, steps - i);
                push(i, x, s1);
                const long long X0 = (long long)std::scalbn(x, 52 - e), S = (long long)std::scalbn(s1, 52 - e);
                x = std::scalbn((double)(X0 + (long long)n * S), e - 52);
                i += n;
                continue;
            }
        }
        push(i, x, s1);
        x = x1;
        i += 1;
    }
    if ((int)out.size() > kMaxSegments) fail(R4WB_ERR_NOT_SUPPORTED, "phase model needs %zu segments", out.size());
}

static inline uint64_t gcd_u64(uint64_t a, uint64_t b) { while (b) { uint64_t t = a % b; a = b; b = t; } return a; }

ScenarioModel::ScenarioModel(const r4wb_scenario_cfg& c_in) : cfg(c_in)
{
    if (cfg.n_sats > (uint32_t)kMaxSats) fail(R4WB_ERR_NOT_SUPPORTED, "at most %d satellites per scenario", kMaxSats);
    if (cfg.n_sats && !c_in.sats) fail(R4WB_ERR_NULL_POINTER, "sats is NULL");
    cfg_sats.assign(c_in.sats, c_in.sats + c_in.n_sats);
    cfg.sats = cfg_sats.data();
    const r4wb_output_cfg& oc = cfg.output;
    if (!(oc.sample_rate > 0.0) || !(oc.duration_s >= 0.0)) fail(R4WB_ERR_INVALID_PARAMETER, "sample_rate/duration_s");

    sc.fs = oc.sample_rate;
    sc.t0_gps = oc.start_time_gps_s;
    sc.duration_s = oc.duration_s;
    sc.total = (uint64_t)std::ceil(oc.duration_s * oc.sample_rate);                       // scenario.rs:80
    sc.B = oc.block_size > 0 ? oc.block_size : (uint64_t)std::ceil(oc.sample_rate * 0.001);   // scenario.rs:667-674
    if (sc.B < 8 || sc.B > 65536) fail(R4WB_ERR_NOT_SUPPORTED, "block size %llu outside [8, 65536]", (unsigned long long)sc.B);
    sc.n_sats = cfg.n_sats;
    sc.flags = cfg.flags;
    sc.antenna = cfg.receiver.antenna;
    sc.ant_peak = cfg.receiver.antenna_peak_gain_dbi;
    sc.ant_bw = cfg.receiver.antenna_beamwidth_deg;
    sc.elev_mask_deg = cfg.receiver.elevation_mask_deg;
    sc.seed = oc.seed;
    sc.iono_enabled = cfg.environment.ionosphere_enabled ? 1u : 0u;
    sc.tropo_enabled = cfg.environment.troposphere_enabled ? 1u : 0u;
    for (int i = 0; i < 4; ++i) { sc.klob_alpha[i] = cfg.environment.klobuchar_alpha[i]; sc.klob_beta[i] = cfg.environment.klobuchar_beta[i]; }
    sc.tropo_height_m = cfg.environment.tropo_height_m; sc.tropo_temperature_k = cfg.environment.tropo_temperature_k;
    sc.tropo_pressure_hpa = cfg.environment.tropo_pressure_hpa; sc.tropo_relative_humidity = cfg.environment.tropo_relative_humidity;

    // receiver model (scenario.rs:160-183, 320-353)
    RxModel& rx = sc.rx;
    rx.position = Lla{cfg.receiver.position.lat_deg, cfg.receiver.position.lon_deg, cfg.receiver.position.alt_m};
    rx.has_trajectory = cfg.receiver.has_trajectory ? 1 : 0;
    rx.traj_start = Lla{cfg.receiver.traj_start.lat_deg, cfg.receiver.traj_start.lon_deg, cfg.receiver.traj_start.alt_m};
    rx.traj_end = Lla{cfg.receiver.traj_end.lat_deg, cfg.receiver.traj_end.lon_deg, cfg.receiver.traj_end.alt_m};
    rx.travel_time_s = 1.0;
    rx.fd_dt = 0.01;
    if (rx.has_trajectory) {
        const double dist = 6371000.0 * gc_angle(rx.traj_start, rx.traj_end);           // scenario_config.rs:358-368
        const double speed = cfg.receiver.traj_has_speed ? cfg.receiver.traj_speed_mps : dist / oc.duration_s;
        rx.travel_time_s = dist / speed;
        rx.fd_dt = std::min(0.01, rx.travel_time_s * 0.001);
    }

    // code NCO constants.  The reference computes samples_per_chip = (8 fs) / chipping_rate in f64.
    sc.chip_rate = 1023000.0;
    const double os_rate = oc.sample_rate * (double)kOversample;
    sc.spc = os_rate / sc.chip_rate;
    if (os_rate != std::floor(os_rate) || os_rate >= 9007199254740992.0)
        fail(R4WB_ERR_NOT_SUPPORTED, "sample_rate must be an integer number of Hz");
    const uint64_t os_i = (uint64_t)os_rate, cr_i = 1023000ull, g = gcd_u64(os_i, cr_i);
    sc.ratA = os_i / g;
    sc.ratB = cr_i / g;
    if (sc.ratA >= (1ull << 40)) fail(R4WB_ERR_NOT_SUPPORTED, "sample_rate / chip rate ratio too fine");
    sc.delta = (double)((long double)sc.spc * (long double)sc.ratB / (long double)sc.ratA - 1.0L);
    sc.delta46 = (uint64_t)floorl(140737488355328.0L / (long double)sc.spc);
    const double S = sc.spc * 0.5;   // oversamples per half-chip
    if (S * 4.0 < (double)(kTaps - 1) || S >= 60.0)
        fail(R4WB_ERR_NOT_SUPPORTED, "sample_rate %.0f Hz outside the supported 3.97-15.3 MHz span", oc.sample_rate);
    sc.kmul = (uint32_t)llround(S * 16777216.0);
    for (int j = 0; j < 8; ++j) sc.cj[j] = (uint32_t)std::min<long long>(llround((double)j * S * 16777216.0), 0xffffffffll);
    {
        const uint64_t D = sc.ratA / gcd_u64(2 * sc.ratB, sc.ratA);
        sc.lattice_den = D <= (1ull << 17) ? D : 0;
    }
    {
        const double nf_lin = std::pow(10.0, cfg.receiver.noise_figure_db / 10.0);       // scenario.rs:532-537
        const double n0 = 1.380649e-23 * 290.0 * nf_lin;
        sc.noise_std = (float)(std::sqrt(n0 * oc.sample_rate / 2.0) * 1e8);
    }

    // satellites
    const RxState rx0 = rx_at(rx, 0.0);
    // Virtual satellites: one per configured satellite, two for GalileoE1OS ((e1b - e1c) / sqrt 2 is a sum of two binary
    // sequences with the same delay and Doppler, and the FIR is linear; satellite_emitter.rs:307-321)
    struct Virt { uint32_t cfg; uint32_t kind; double scale; };      // kind: 0 GPS C/A, 1 E1B, 2 E1C, 3 GPS L5 (I5), 4 GLONASS L1OF
    std::vector<Virt> virt;
    for (uint32_t k = 0; k < cfg.n_sats; ++k) {
        const r4wb_sat_cfg& c = cfg_sats[k];
        const bool galileo = c.signal == R4WB_SIG_GALILEO_E1 || c.signal == R4WB_SIG_GALILEO_E1C || c.signal == R4WB_SIG_GALILEO_E1OS;
        if ((unsigned)c.signal > (unsigned)R4WB_SIG_GALILEO_E1OS) fail(R4WB_ERR_INVALID_PARAMETER, "satellite %u: unknown signal %u", k, (unsigned)c.signal);
        if (c.signal == R4WB_SIG_GLONASS_L1OF) {
            // the reference builds GlonassCodeGenerator::new(prn as i8), which asserts a frequency channel of -7..6 (prn.rs:181-183)
            if (c.prn > 6) fail(R4WB_ERR_INVALID_PARAMETER, "GLONASS PRN (used as frequency channel) must be 0-6, got %u", c.prn);
            if (c.plane >= 3 || c.slot >= 8) fail(R4WB_ERR_INVALID_PARAMETER, "GLONASS plane 0-2 / slot 0-7");
        } else if (galileo) {
            if (c.prn < 1 || c.prn > 50) fail(R4WB_ERR_INVALID_PARAMETER, "Galileo PRN must be 1-50, got %u", c.prn);
            if (c.plane >= 3 || c.slot >= 10) fail(R4WB_ERR_INVALID_PARAMETER, "Galileo plane 0-2 / slot 0-9");
        } else {
            if (c.prn < 1 || c.prn > 32) fail(R4WB_ERR_INVALID_PARAMETER, "GPS PRN must be 1-32, got %u", c.prn);
            if (c.plane >= 6 || c.slot >= 6) fail(R4WB_ERR_INVALID_PARAMETER, "GPS plane 0-5 / slot 0-5");
        }
        if (c.signal == R4WB_SIG_GPS_L1CA) virt.push_back({k, 0u, 1.0});
        else if (c.signal == R4WB_SIG_GPS_L5) virt.push_back({k, 3u, 1.0});
        else if (c.signal == R4WB_SIG_GLONASS_L1OF) virt.push_back({k, 4u, 1.0});
        else if (c.signal == R4WB_SIG_GALILEO_E1) virt.push_back({k, 1u, 1.0});
        else if (c.signal == R4WB_SIG_GALILEO_E1C) virt.push_back({k, 2u, 1.0});
        else { const double sc2 = 1.0 / std::sqrt(2.0); virt.push_back({k, 1u, sc2}); virt.push_back({k, 2u, -sc2}); }
    }
    if (virt.size() > (size_t)kMaxSats) fail(R4WB_ERR_NOT_SUPPORTED, "at most %d (virtual) satellites per scenario", kMaxSats);
    sc.n_sats = (uint32_t)virt.size();
    sats.resize(virt.size());
    satcode.assign(std::max<size_t>(1, virt.size()), SatCode{});
    cfg_index.assign(virt.size(), 0u);
    codebits.assign(std::max<size_t>(1, virt.size()) * 128, 0u);
    perbits.assign(std::max<size_t>(1, virt.size()) * kPerWords, 0u);
    dsat.assign(std::max<size_t>(1, virt.size()), DirectSat{});
    dcodebits.assign(std::max<size_t>(1, virt.size()) * kDirectWords, 0u);
    const double os_rate_d = oc.sample_rate * (double)kOversample;
    for (uint32_t k = 0; k < (uint32_t)virt.size(); ++k) {
        const r4wb_sat_cfg& c = cfg_sats[virt[k].cfg];
        cfg_index[k] = virt[k].cfg;
        SatConst& s = sats[k];
        std::memset(&s, 0, sizeof s);
        s.amp_scale = virt[k].scale;
        s.orbit = nominal_orbit(c.signal, c.plane, c.slot);
        // GnssSignal::carrier_frequency_hz / chipping_rate (gnss/types.rs:73-90)
        s.carrier_hz = c.signal == R4WB_SIG_GPS_L5 ? 1176450000.0 : (c.signal == R4WB_SIG_GLONASS_L1OF ? 1602000000.0 : 1575420000.0);
        s.chip_rate = virt[k].kind == 3 ? 10230000.0 : (virt[k].kind == 4 ? 511000.0 : sc.chip_rate);
        s.direct = virt[k].kind >= 3 ? 1u : 0u;
        if (s.direct) any_direct = true;
        s.has = c.has;
        s.orbital_dynamics = c.orbital_dynamics ? 1u : 0u;
        s.tx_power_dbw = c.tx_power_dbw;
        s.elevation_deg = c.elevation_deg; s.range_m = c.range_m; s.range_rate_mps = c.range_rate_mps;
        s.doppler_hz = c.doppler_hz; s.doppler_rate_hz_per_s = c.doppler_rate_hz_per_s; s.cn0_dbhz = c.cn0_dbhz;
        s.iono_delay_m = c.iono_delay_m; s.tropo_delay_m = c.tropo_delay_m;
        const bool doppler_from_orbit = c.orbital_dynamics || (!(c.has & R4WB_HAS_DOPPLER) && !(c.has & R4WB_HAS_RANGE_RATE));
        const bool range_from_orbit = c.orbital_dynamics || !(c.has & R4WB_HAS_RANGE);
        const bool env_from_orbit = (!(c.has & R4WB_HAS_IONO) && cfg.environment.ionosphere_enabled) ||
                                    (!(c.has & R4WB_HAS_TROPO) && cfg.environment.troposphere_enabled);   // the models need the look angle
        s.needs_orbit = (doppler_from_orbit || range_from_orbit || env_from_orbit || !(c.has & R4WB_HAS_ELEVATION)) ? 1u : 0u;
        const bool const_doppler = !c.orbital_dynamics && (((c.has & R4WB_HAS_DOPPLER) && !(c.has & R4WB_HAS_DOPPLER_RATE)) ||
                                                            (!(c.has & R4WB_HAS_DOPPLER) && (c.has & R4WB_HAS_RANGE_RATE)));
        s.static_phase = (const_doppler && (c.has & R4WB_HAS_ELEVATION)) ? 1u : 0u;
        if (!(c.has & R4WB_HAS_ELEVATION)) any_var_visibility = true;
        if (!s.static_phase) any_dynamic = true;
        // orbital anchors at t0 (scenario.rs:195-204)
        Vec3 sp, sv;
        orbit_state(s.orbit, sc.t0_gps, sp, sv);
        s.orb_range_t0 = look_from(rx0.pos, rx0.lla, sp).range_m;
        s.orb_doppler_t0 = -los_rate(rx0.pos, rx0.vel, sp, sv) * s.carrier_hz / kC;
        // constant-Doppler satellites: segments of the sequential f64 phase accumulation
        s.seg_begin = (int32_t)segments.size();
        if (s.static_phase) {
            const double dop = (c.has & R4WB_HAS_DOPPLER) ? c.doppler_hz : -c.range_rate_mps * s.carrier_hz / kC;
            const double inc = 2.0 * kPi * dop / sc.fs;                                  // scenario.rs:522
            std::vector<PhaseSegment> segs;
            build_phase_segments(inc, sc.total, segs);
            segments.insert(segments.end(), segs.begin(), segs.end());
        }
        s.seg_count = (int32_t)segments.size() - s.seg_begin;

        // code structure (satellite_emitter.rs:248-343)
        SatCode& cd = s.code;
        int8_t chips[10230];
        if (virt[k].kind == 0) { gps_ca_code_chips(c.prn, chips); cd.code_len = 1023; cd.has_boc = 0; }
        else if (virt[k].kind == 3) { gps_l5_i5_code_chips(c.prn, chips); cd.code_len = 10230; cd.has_boc = 0; }
        else if (virt[k].kind == 4) { glonass_code_chips(chips); cd.code_len = 511; cd.has_boc = 0; }
        else { e1_code_chips(virt[k].kind == 1 ? 0u : 1u, c.prn, chips); cd.code_len = (uint32_t)kCodeLen; cd.has_boc = 1; }
        cd.per_len = 2u * cd.code_len;
        cd.epoch_period = 1; cd.epoch_bits = 0;
        if (virt[k].kind == 2) {                              // E1C: 25-chip secondary code, one chip per 4 ms epoch
            cd.epoch_period = (uint32_t)kSecLen; cd.epoch_bits = kSecBits;
        } else if (c.nav_data) {                              // nav bit (bit_idx + prn) % 2, bit_idx = epoch / periods_per_bit (:286-292)
            const uint32_t ppb = virt[k].kind == 1 ? 1u : 20u;   // 1 / (250 bps x 4 ms); 1 / (50 bps x 1 ms) for GPS L1 C/A, L5, GLONASS
            cd.epoch_period = 2u * ppb;
            for (uint32_t e = 0; e < cd.epoch_period; ++e)
                if (((e / ppb) + c.prn) % 2u != 0u) cd.epoch_bits |= 1ull << e;
        }
        cd.hc_mod = cd.per_len * cd.epoch_period;
        satcode[k] = cd;
        if (s.direct) {      // direct path: packed chips + the per-satellite samples_per_chip of satellite_emitter.rs:245
            dsat[k] = DirectSat{os_rate_d / s.chip_rate, cd.epoch_bits, 1u, cd.code_len, cd.epoch_period, 0u};
            for (uint32_t i = 0; i < cd.code_len; ++i)
                dcodebits[(size_t)k * kDirectWords + (i >> 5)] |= (chips[i] < 0 ? 1u : 0u) << (i & 31);
            continue;        // no half-chip tables: k_synth skips this satellite
        }
        for (uint32_t i = 0; i < cd.code_len; ++i)
            codebits[(size_t)k * 128 + (i >> 5)] |= (chips[i] < 0 ? 1u : 0u) << (i & 31);
        // half-chip signs of one primary-code period (with BOC(1,1): second half of every chip inverted), followed by a
        // copy of its first 64 bits so a 32-bit window may start anywhere in the period
        uint32_t* per = perbits.data() + (size_t)k * kPerWords;
        for (uint32_t n = 0; n < cd.per_len + 64u; ++n) {
            const uint32_t hc = n % cd.per_len;
            per[n >> 5] |= ((chips[hc >> 1] < 0 ? 1u : 0u) ^ (cd.has_boc ? (hc & 1u) : 0u)) << (n & 31);
        }
    }
    if (segments.empty()) segments.push_back(PhaseSegment{0, 0.0, 0.0});

    // filter: FirFilter::lowpass(lpf_cutoff or fs/2, 8 fs, 63)  (scenario.rs:209-217)
    double h[kTaps];
    design_lowpass(oc.lpf_cutoff_hz > 0.0 ? oc.lpf_cutoff_hz : oc.sample_rate / 2.0, os_rate, h);
    double etab[64], run = 0.0;
    for (int d = 0; d < 64; ++d) {
        taps_f[d] = 0.0f;
        if (d < kTaps) { taps_f[d] = (float)h[d]; run += h[d]; }
        etab[d] = d >= kTaps - 1 ? 1.0 : run;                // E[d] = sum_{k<=d} h[k]; window fully covered -> 1
        etab_f[d] = (float)etab[d];
    }
    // Collapsed-FIR table.  With t0 = oversamples since the newest half-chip boundary (2^-24 fixed point), the
    // boundary ages are d_j = (t0 + cj[j]) >> 24; every time one of them steps, d_0+d_1+d_2+d_3 steps by one, so
    // that sum (minus its minimum) names the age class.  Entry = s_4 + sum_j (s_j - s_{j+1}) E[min(d_j, 62)],
    // evaluated in f64 and rounded once.
    sc.dsum0 = (sc.cj[1] >> kTBits) + (sc.cj[2] >> kTBits) + (sc.cj[3] >> kTBits);
    {
        std::vector<uint64_t> starts{0};
        const uint64_t one = 1ull << kTBits;
        for (int j = 0; j < 4; ++j)
            for (uint64_t n = 1; n * one < (uint64_t)sc.kmul + sc.cj[j] + one; ++n)
                if (n * one >= sc.cj[j] && n * one - sc.cj[j] < sc.kmul) starts.push_back(n * one - sc.cj[j]);
        std::sort(starts.begin(), starts.end());
        // classes are numbered 0 .. (steps of d_0 + d_1 + d_2 + d_3 over one half-chip) <= 4 S; odd row stride spreads the
        // sign patterns over the shared-memory banks (81 at 5 MHz)
        uint32_t max_cls = 0;
        for (uint64_t t0 : starts) {
            uint32_t dsum = 0;
            for (int j = 0; j < 4; ++j) dsum += (uint32_t)((t0 + sc.cj[j]) >> kTBits);
            max_cls = std::max(max_cls, dsum - sc.dsum0);
        }
        if (max_cls >= (uint32_t)kMaxYStride) fail(R4WB_ERR_NOT_SUPPORTED, "FIR age class %u out of range", max_cls);
        sc.ystride = (max_cls + 2u) | 1u;
        ytab.assign((size_t)32 * sc.ystride, 0.0f);
        for (uint64_t t0 : starts) {
            uint32_t d[4], dsum = 0;
            for (int j = 0; j < 4; ++j) { d[j] = (uint32_t)((t0 + sc.cj[j]) >> kTBits); dsum += d[j]; }
            dsum -= sc.dsum0;
            for (uint32_t pat = 0; pat < 32; ++pat) {
                double sgn[kJ + 1];
                for (int j = 0; j <= kJ; ++j) sgn[j] = ((pat >> (kJ - j)) & 1u) ? -1.0 : 1.0;
                double y = sgn[kJ];
                for (int j = 0; j < kJ; ++j) y += (sgn[j] - sgn[j + 1]) * etab[d[j] < 62u ? d[j] : 62u];
                ytab[(size_t)pat * sc.ystride + dsum] = (float)y;
            }
        }
    }

    // Boundary-age class as a table over the half-chip fraction.  Every class boundary sits at a multiple of 1/D
    // (D = lattice_den: fractions n / S mod 1 with S = ratA / (2 ratB) oversamples per half-chip), so bin q = floor(frac D)
    // names the class exactly for any sample that is not within the ambiguity band of a lattice point (those blocks
    // are flagged and take the arithmetic path).  Bin centre evaluated in extended precision.
    sc.lut_den = 0;
    if (sc.lattice_den != 0 && sc.lattice_den <= 40960 && !std::getenv("R4WB_SYNTH_NO_LUT")) {
        const uint32_t D = (uint32_t)sc.lattice_den;
        sc.lut_den = D;
        clslut.assign(((size_t)D + 15) & ~(size_t)15, 0);
        const long double Sl = (long double)sc.ratA / (2.0L * (long double)sc.ratB);
        for (uint32_t q = 0; q < D; ++q) {
            const long double t0 = ((long double)q + 0.5L) / (long double)D * Sl;
            uint32_t dsum = 0;
            for (int j = 0; j < 4; ++j) dsum += (uint32_t)floorl(t0 + (long double)j * Sl);
            dsum -= sc.dsum0;
            if (dsum >= sc.ystride) fail(R4WB_ERR_NOT_SUPPORTED, "FIR age class %u out of range", dsum);
            clslut[cls_lut_index(q)] = (uint8_t)dsum;
        }
    } else {
        clslut.assign(16, 0);
    }

    // Lattice kernel (synth_lattice.cuh): a block of B = 2 q samples whose half-chip position advances by p / q per sample
    // (direct-path satellites are skipped by it like by k_synth; k_synth_direct adds them afterwards)
    lat = LatConst{};
    {
        const char* env = std::getenv("R4WB_SYNTH_LATTICE");                 // A/B hook: 0 keeps every launch on k_synth
        const uint32_t D = sc.lut_den;
        const uint64_t g2 = gcd_u64(2 * sc.ratB, sc.ratA);
        const uint64_t pov = 2 * sc.ratB / g2;                                // oversample step = pov / D half-chips, gcd(pov, D) = 1
        const uint32_t q = D / 8u;
        if (!(env && env[0] == '0') && D != 0 && D % 8u == 0 && q % 2u == 0 && sc.B == 2ull * q && pov < D && sc.ystride <= 200u &&
            sc.n_sats >= 1 && sc.n_sats <= 16) {
            const uint32_t K = (q + 2u * kSynthThreads - 1u) / (2u * kSynthThreads);
            auto inv_mod = [](uint64_t a, uint64_t m) {                       // a^-1 mod m (gcd = 1), extended Euclid
                long long t = 0, nt = 1, r = (long long)m, nr = (long long)(a % m);
                while (nr != 0) { const long long qq = r / nr; long long x = t - qq * nt; t = nt; nt = x; x = r - qq * nr; r = nr; nr = x; }
                return (uint64_t)(t < 0 ? t + (long long)m : t);
            };
            if ((K == 4 || K == 5) && gcd_u64(pov, q) == 1) {
                lat.q = q; lat.p = (uint32_t)pov; lat.pov = (uint32_t)pov;
                lat.pinv = (uint32_t)inv_mod(pov, q); lat.povinv = (uint32_t)inv_mod(pov, D);
                lat.K = K; lat.cls_len = (q + 2u * kSynthThreads * K + 2u + 15u) & ~15u;
                lat.d8_20 = (uint32_t)llround((double)pov / (double)q * 1048576.0);
                lat.step20 = (uint32_t)llround((double)(2 * kSynthThreads) * (double)pov / (double)q * 1048576.0);
                // class of sample index n for sub-residue b0: bin 8 (p n mod q) + b0 (bin centre, extended precision as above)
                const long double Sl = (long double)sc.ratA / (2.0L * (long double)sc.ratB);
                clsn.assign((size_t)8 * lat.cls_len, 0);
                for (uint32_t b0 = 0; b0 < 8; ++b0)
                    for (uint32_t n = 0; n < lat.cls_len; ++n) {
                        const uint32_t bin = 8u * (uint32_t)(((uint64_t)pov * (n % q)) % q) + b0;
                        const long double t0 = ((long double)bin + 0.5L) / (long double)D * Sl;
                        uint32_t dsum = 0;
                        for (int j = 0; j < 4; ++j) dsum += (uint32_t)floorl(t0 + (long double)j * Sl);
                        clsn[(size_t)b0 * lat.cls_len + n] = (uint8_t)(dsum - sc.dsum0);
                    }
                // the collapsed-FIR table transposed: row = class (256 B), column = sign pattern
                ytab2.assign(((size_t)sc.ystride * kLatYStride + 3) & ~(size_t)3, 0.0f);
                for (uint32_t c = 0; c < sc.ystride; ++c)
                    for (uint32_t pat = 0; pat < 32; ++pat) ytab2[(size_t)c * kLatYStride + pat] = ytab[(size_t)pat * sc.ystride + c];
            }
        }
    }
    if (clsn.empty()) clsn.assign(16, 0);
    if (ytab2.empty()) ytab2.assign(64, 0.0f);
    tile_k = 10;
    if (const char* e = std::getenv("R4WB_SYNTH_TILE_K")) { if (std::atoi(e) == 5) tile_k = 5; }   // tuning hook
    {
        const double span = std::ceil((double)synth_tile_samples(tile_k) * kOversample * (2.0 / sc.spc)) + 2.0;
        nw64 = (uint32_t)std::ceil((span + 8.0) / 32.0) + 1u;
    }
}

// GnssScenario::satellite_status (scenario.rs:564-633): static receiver position, zero receiver velocity
void ScenarioModel::status(uint64_t current, r4wb_sat_status* out, uint32_t cap, uint32_t* n_out) const
{
    if (cap < cfg.n_sats) fail(R4WB_ERR_INVALID_SIZE, "status buffer holds %u of %u satellites", cap, cfg.n_sats);
    const double t = sc.t0_gps + (double)current / sc.fs;
    const Lla rx_lla = sc.rx.position;
    const Vec3 rx_pos = ecef_of(rx_lla), zero{0.0, 0.0, 0.0};
    for (uint32_t k = 0; k < cfg.n_sats; ++k) {      // configured satellites (a GalileoE1OS one is two virtual satellites)
        const r4wb_sat_cfg& c = cfg_sats[k];
        Vec3 sp, sv;
        orbit_state(nominal_orbit(c.signal, c.plane, c.slot), t, sp, sv);
        const double carrier_hz = c.signal == R4WB_SIG_GPS_L5 ? 1176450000.0 : (c.signal == R4WB_SIG_GLONASS_L1OF ? 1602000000.0 : 1575420000.0);
        const Look la = look_from(rx_pos, rx_lla, sp);
        r4wb_sat_status& o = out[k];
        std::memset(&o, 0, sizeof o);
        o.signal = c.signal; o.prn = c.prn;
        o.range_m = (c.has & R4WB_HAS_RANGE) ? c.range_m : la.range_m;
        o.elevation_deg = (c.has & R4WB_HAS_ELEVATION) ? c.elevation_deg : la.elevation_deg;
        o.azimuth_deg = (c.has & R4WB_HAS_AZIMUTH) ? c.azimuth_deg : la.azimuth_deg;
        o.range_rate_mps = (c.has & R4WB_HAS_RANGE_RATE) ? c.range_rate_mps : los_rate(rx_pos, zero, sp, sv);
        o.doppler_hz = (c.has & R4WB_HAS_DOPPLER) ? c.doppler_hz : -o.range_rate_mps * carrier_hz / kC;
        o.antenna_gain_dbi = antenna_gain_dbi(sc.antenna, sc.ant_peak, sc.ant_bw, o.elevation_deg);
        o.cn0_dbhz = (c.has & R4WB_HAS_CN0) ? c.cn0_dbhz
                                            : c.tx_power_dbw - fspl_db(o.range_m, carrier_hz) + o.antenna_gain_dbi + 204.0;
        // scenario.rs:604-613: override, else the emitter's models at the ORBIT's look angle
        o.iono_delay_m = 0.0; o.tropo_delay_m = 0.0;
        if (c.has & R4WB_HAS_IONO) o.iono_delay_m = c.iono_delay_m;
        else if (sc.iono_enabled)
            o.iono_delay_m = klobuchar_delay_s(sc.klob_alpha, sc.klob_beta, la.elevation_deg * kDeg, la.azimuth_deg * kDeg, rx_lla.lat_deg * kDeg,
                                               rx_lla.lon_deg * kDeg, std::fmod(t, 604800.0)) * kC;
        if (c.has & R4WB_HAS_TROPO) o.tropo_delay_m = c.tropo_delay_m;
        else if (sc.tropo_enabled)
            o.tropo_delay_m = saastamoinen_delay_m(sc.tropo_height_m, sc.tropo_temperature_k, sc.tropo_pressure_hpa, sc.tropo_relative_humidity,
                                                   la.elevation_deg * kDeg);
        o.visible = o.elevation_deg > 0.0 ? 1 : 0;
        o.clock_correction_s = 0.0;
    }
    if (n_out) *n_out = cfg.n_sats;
}

// ---------------------------------------------------------------------------------------------- sequential API
void SeqState::reset(size_t n_sats)
{
    m.assign(n_sats, 0);
    phi.assign(n_sats, 0);
    ph.assign(n_sats, 0.0);
    dop.assign(2 * n_sats, 0.0);
    prev.assign(n_sats, BlockSat{});
    has_prev.assign(n_sats, 0);
}

void SeqState::make_table(const ScenarioModel& md, uint64_t first, uint32_t n, std::vector<BlockSat>& tab, BlockHdr hdr[2])
{
    const uint32_t ns = md.sc.n_sats;
    // dynamic satellites start the block at the reference's own f64 phase (walked sample by sample in advance()) unless the
    // caller asked for the closed form
    const bool exact = !(md.sc.flags & R4WB_FLAG_CLOSED_FORM_PHASE);
    tab.assign((size_t)2 * std::max(1u, ns), BlockSat{});
    for (uint32_t s = 0; s < ns; ++s) {
        BlockSat cur;
        const uint64_t dyn_phi = exact ? cycles_to_fixed(ph[s] / (2.0 * kPi)) : phi[s];
        fill_block_sat(md.sc, md.sats[s], md.segments.data(), first, n, md.sats[s].static_phase ? m[s] : dyn_phi, cur, &dop[2 * s]);
        cur.prev = has_prev[s] ? (int32_t)s : -1;
        tab[s] = prev[s];
        tab[ns + s] = cur;
    }
    hdr[0] = BlockHdr{0, 0, 0};
    hdr[1] = BlockHdr{first, n, 0};
}

void SeqState::advance(const ScenarioModel& md, const std::vector<BlockSat>& tab, uint32_t n)
{
    const uint32_t ns = md.sc.n_sats;
    for (uint32_t s = 0; s < ns; ++s) {
        const BlockSat& cur = tab[ns + s];
        if (!(cur.flags & 1u)) continue;
        m[s] += n;
        if (!md.sats[s].static_phase) {
            phi[s] += block_advance(cur);
            if (!(md.sc.flags & R4WB_FLAG_CLOSED_FORM_PHASE)) ph[s] = phase_walk(ph[s], dop[2 * s], dop[2 * s + 1], n, md.sc.fs);   // scenario.rs:520-523
        }
        prev[s] = cur;
        prev[s].prev = -1;
        has_prev[s] = 1;
    }
}

size_t synth_smem_bytes(uint32_t n_sats, uint32_t nw64, uint32_t lut_den, uint32_t ystride)
{
    size_t b = ((size_t)(32 * ystride + 128) * 4 + 15) & ~(size_t)15;   // ytab, taps, etab
    b += (size_t)n_sats * kPerWords * 4;
    b += (size_t)n_sats * sizeof(TileRec);
    b += (size_t)n_sats * nw64 * 8;
    b += (size_t)n_sats * (nw64 + 1) * 4;
    b = (b + 15) & ~(size_t)15;
    b += ((size_t)lut_den + 15) & ~(size_t)15;
    return b;
}

int synth_tile_samples(int K) { return kSynthThreads * 2 * K; }

}  // namespace r4wb
