// synth_math.cuh — per-(block, satellite) and per-sample arithmetic of the synthesis path.
//
// Everything here is __host__ __device__: the sm_100a kernels (synth_kernels.cu) call it per thread, and
// tests/emu/ replays the very same functions in plain host loops so the index arithmetic can be checked
// against the oracle in a container without a GPU.  The host build is a debugging aid for tests only;
// libr4w_b200.so never renders samples on the CPU.
//
// What it replaces per sample:
//   SatelliteEmitter::generate_baseband_iq   gnss/satellite_emitter.rs:218-347  (E1C branch :325-330)
//   FirFilter::process + step_by(8)          core/filters/fir.rs:392-409, gnss/scenario.rs:486-489
//   Doppler rotation + amplitude             gnss/scenario.rs:516-528
//   thermal noise                            gnss/scenario.rs:530-542   (Philox4x32-10 instead of xorshift64)
//
// Algebra: the 8x-oversampled baseband is +-1 and constant over half-chips, so the 63-tap
// FIR at a decimated output is  y = s_old + sum_j (s_j - s_{j+1}) * E[d_j]  with E the running sum of the
// taps and d_j the age (in oversamples) of the j-th half-chip boundary inside the window.  Boundaries
// come from a 64-bit fixed-point code NCO; whenever a sample lies within the rounding wobble of the
// reference's own f64 expression the reference expression is evaluated literally (exact path).
#pragma once
#include <cmath>
#include <cstdint>

#include "synth.cuh"

#ifdef __CUDACC__
#define R4WB_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define R4WB_HD_NOINLINE static
#endif

namespace r4wb {

// ----------------------------------------------------------------------------------------------
// strict f64 helpers (never contracted into FMAs; host g++ has no FMA target so plain ops are strict)
R4WB_HD double mul_rn(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
R4WB_HD double add_rn(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
R4WB_HD double div_rn(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
R4WB_HD uint32_t umulhi32(uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}
// low 32 bits of ((hi:lo) >> (sh & 31))
R4WB_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh)
{
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, sh);
#else
    return (uint32_t)(((((uint64_t)hi) << 32) | (uint64_t)lo) >> (sh & 31u));
#endif
}
R4WB_HD void fast_sincos(float x, float* s, float* c)
{
#ifdef __CUDA_ARCH__
    __sincosf(x, s, c);
#else
    *s = sinf(x);
    *c = cosf(x);
#endif
}
R4WB_HD float fast_log2(float x)
{
#ifdef __CUDA_ARCH__
    return __log2f(x);
#else
    return log2f(x);
#endif
}

constexpr uint64_t kFracMask = (1ull << kFracBits) - 1ull;
constexpr uint64_t kUMod = (uint64_t)kHalfChipsPerSec << kFracBits;   // 204600 * 2^46 < 2^64
constexpr double kTwo64 = 18446744073709551616.0;
constexpr int kJ = 4;   // half-chip boundaries that can fall inside one 63-tap window

// E1C secondary code as a bit mask (bit e set <=> chip -1), galileo_e1_codes.rs:27-31
constexpr uint32_t kSecBits = (1u << 2) | (1u << 3) | (1u << 4) | (1u << 13) | (1u << 15) | (1u << 17) |
                              (1u << 18) | (1u << 19) | (1u << 20) | (1u << 24);

// ----------------------------------------------------------------------------------------------
// reference phase accumulation, piecewise exact (see PhaseSegment)
R4WB_HD void static_phase_at(const PhaseSegment* segs, int count, uint64_t m, double& x, double& step)
{
    int lo = 0, hi = count - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (segs[mid].i0 <= m) lo = mid; else hi = mid - 1;
    }
    const PhaseSegment sg = segs[lo];
    step = sg.step;
    if (m == sg.i0 || sg.x0 == 0.0) { x = sg.x0 + (double)(m - sg.i0) * sg.step; return; }
    const int e = ilogb(sg.x0);
    const long long X0 = (long long)scalbn(sg.x0, 52 - e);
    const long long S = (long long)scalbn(sg.step, 52 - e);
    const long long X = X0 + (long long)(m - sg.i0) * S;
    x = scalbn((double)X, e - 52);
}

R4WB_HD uint64_t cycles_to_fixed(double cycles)
{
    double fr = cycles - floor(cycles);
    if (!(fr < 1.0)) fr = 0.0;
    const double v = fr * kTwo64;
    return v >= kTwo64 ? 0ull : (uint64_t)v;
}

// Phase 1 of generate_block for one (block, satellite): gnss/scenario.rs:378-454 +
// satellite_emitter.rs:228-242, plus the fixed-point NCO start values the kernel needs.
// m_or_phi: visible-sample count (static-phase satellites) or phase in cycles 0.64 (dynamic, explicit mode).
R4WB_HD void fill_block_sat(const ScenConst& sc, const SatConst& st, const PhaseSegment* segs, uint64_t first,
                            uint32_t n, uint64_t m_or_phi, BlockSat& o, double* dop_out = nullptr /* [2]: doppler_start_hz, doppler_end_hz */)
{
    const double fs = sc.fs;
    const double elapsed = (double)first / fs;
    const double t_start = sc.t0_gps + elapsed;
    const double t_end = t_start + (double)n / fs;
    const double elapsed_end = elapsed + (double)n / fs;

    double la_el = 0.0, la_az = 0.0, la_range = 0.0, dop_s = 0.0, dop_e = 0.0, rx_lat = 0.0, rx_lon = 0.0;
    if (st.needs_orbit) {
        const RxState rx = rx_at(sc.rx, elapsed);
        Vec3 ps, vs, pe, ve;
        orbit_state(st.orbit, t_start, ps, vs);
        orbit_state(st.orbit, t_end, pe, ve);
        const Look la = look_from(rx.pos, rx.lla, ps);
        la_el = la.elevation_deg;
        la_az = la.azimuth_deg;
        la_range = la.range_m;
        rx_lat = rx.lla.lat_deg; rx_lon = rx.lla.lon_deg;
        dop_s = -los_rate(rx.pos, rx.vel, ps, vs) * st.carrier_hz / kC;
        dop_e = -los_rate(rx.pos, rx.vel, pe, ve) * st.carrier_hz / kC;
    }
    const double elevation = (st.has & R4WB_HAS_ELEVATION) ? st.elevation_deg : la_el;
    const bool visible = !(elevation < sc.elev_mask_deg);

    double range_m;
    if (st.orbital_dynamics) {
        range_m = (st.has & R4WB_HAS_RANGE) ? st.range_m + (la_range - st.orb_range_t0) : la_range;
    } else if ((st.has & R4WB_HAS_RANGE) && (st.has & R4WB_HAS_RANGE_RATE)) {
        range_m = add_rn(st.range_m, mul_rn(st.range_rate_mps, elapsed));
    } else {
        range_m = (st.has & R4WB_HAS_RANGE) ? st.range_m : la_range;
    }
    double ds, de;
    if (st.orbital_dynamics) {
        if (st.has & R4WB_HAS_DOPPLER) {
            ds = st.doppler_hz + (dop_s - st.orb_doppler_t0);
            de = st.doppler_hz + (dop_e - st.orb_doppler_t0);
        } else { ds = dop_s; de = dop_e; }
    } else if (st.has & R4WB_HAS_DOPPLER) {
        if (st.has & R4WB_HAS_DOPPLER_RATE) {
            ds = add_rn(st.doppler_hz, mul_rn(st.doppler_rate_hz_per_s, elapsed));
            de = add_rn(st.doppler_hz, mul_rn(st.doppler_rate_hz_per_s, elapsed_end));
        } else { ds = st.doppler_hz; de = st.doppler_hz; }
    } else if (st.has & R4WB_HAS_RANGE_RATE) {
        ds = de = -st.range_rate_mps * st.carrier_hz / kC;
    } else { ds = dop_s; de = dop_e; }
    if (dop_out) { dop_out[0] = ds; dop_out[1] = de; }

    // scenario.rs:430-439: the override, else SatelliteEmitter::status_at(t_start) (satellite_emitter.rs:165-178): the models see
    // the ORBIT's look angle (not the YAML elevation) and GPS seconds of week; 0 m when the model is disabled
    double iono_m = 0.0, tropo_m = 0.0;
    if (st.has & R4WB_HAS_IONO) iono_m = st.iono_delay_m;
    else if (sc.iono_enabled)
        iono_m = klobuchar_delay_s(sc.klob_alpha, sc.klob_beta, la_el * kDeg, la_az * kDeg, rx_lat * kDeg, rx_lon * kDeg, fmod(t_start, 604800.0)) * kC;
    if (st.has & R4WB_HAS_TROPO) tropo_m = st.tropo_delay_m;
    else if (sc.tropo_enabled)
        tropo_m = saastamoinen_delay_m(sc.tropo_height_m, sc.tropo_temperature_k, sc.tropo_pressure_hpa, sc.tropo_relative_humidity, la_el * kDeg);
    const double iono_s = iono_m / kC;
    const double tropo_s = tropo_m / kC;
    double cn0 = st.cn0_dbhz;
    if (!(st.has & R4WB_HAS_CN0))
        cn0 = st.tx_power_dbw - fspl_db(range_m, st.carrier_hz) + antenna_gain_dbi(sc.antenna, sc.ant_peak, sc.ant_bw, elevation) + 204.0;
    const double amp = pow(10.0, ((cn0 - 204.0) + 160.0) / 20.0);

    // code phase of the block (satellite_emitter.rs:228-242)
    const double total_delay_s = add_rn(add_rn(div_rn(range_m, kC), iono_s), tropo_s);
    const double chips_delay = mul_rn(total_delay_s, st.chip_rate);
    const SatCode& cd = st.code;
    const double phase0 = fmod(chips_delay, (double)cd.code_len);
    const double eq = div_rn(chips_delay, (double)cd.code_len);
    const uint64_t e0 = eq > 0.0 ? (uint64_t)eq : 0ull;

    const uint64_t G = first * (uint64_t)kOversample;
    if (st.direct) {     // other chip rates: k_synth_direct evaluates the reference expression per tap from phase0 / e0 / G
        uint64_t phi_d;
        long long f_d, df_d;
        if (st.static_phase) {
            if (sc.flags & R4WB_FLAG_CLOSED_FORM_PHASE) {
                f_d = (long long)llrint(ds / fs * kTwo64);
                phi_d = (uint64_t)f_d * m_or_phi;
            } else {
                double x, step;
                static_phase_at(segs + st.seg_begin, st.seg_count, m_or_phi, x, step);
                f_d = (long long)llrint(step / (2.0 * kPi) * kTwo64);
                phi_d = cycles_to_fixed(x / (2.0 * kPi));
            }
            df_d = 0;
        } else {
            f_d = (long long)llrint(ds / fs * kTwo64);
            df_d = (long long)llrint((de - ds) / ((double)n * fs) * kTwo64);
            phi_d = m_or_phi;
        }
        o.U = 0; o.phi = phi_d; o.f = f_d; o.df = df_d; o.phase0 = phase0; o.G = G; o.n = n; o.e0 = (uint32_t)e0;
        o.amp = (float)(amp * st.amp_scale); o.flags = (visible ? 1u : 0u) | 32u; o.prev = -1; o.eps46 = 0; o.eps_t = 0; o.pad = 0;
        return;
    }
    // fixed-point half-chip position of the block's first oversample: exact rational part + f64 corrections
    const unsigned __int128 n1 = (unsigned __int128)G * sc.ratB;
    const uint64_t ci = (uint64_t)(n1 / sc.ratA);
    const uint64_t cr = (uint64_t)(n1 % sc.ratA);
    const uint64_t fracfx = (uint64_t)((((unsigned __int128)cr) << (kFracBits + 1)) / sc.ratA);
    const double chips_exact = (double)ci + (double)cr / (double)sc.ratA;
    const double corr_chips = -chips_exact * (sc.delta / (1.0 + sc.delta));
    const long long corrfx = llrint(corr_chips * 140737488355328.0 /* 2^47 */);
    const uint64_t hc_int = ((e0 % cd.epoch_period) * (uint64_t)cd.per_len + 2ull * (ci % ((uint64_t)cd.code_len * cd.epoch_period))) % cd.hc_mod;
    const __int128 umod = (__int128)((uint64_t)cd.hc_mod << kFracBits);
    const uint64_t p0fx = (uint64_t)(phase0 * 140737488355328.0);
    __int128 Uw = ((__int128)hc_int << kFracBits) + (__int128)p0fx + (__int128)fracfx + (__int128)corrfx;
    while (Uw < 0) Uw += umod;
    while (Uw >= umod) Uw -= umod;
    const uint64_t U = (uint64_t)Uw;

    // ambiguity band: the reference's cf = fl(phase0 + fl(g/spc)) is within ulp(cf) chips of the real value
    const double cf_max = phase0 + (double)(G + (uint64_t)kOversample * n) / sc.spc + 2.0;
    const double ulp = scalbn(1.0, ilogb(cf_max) - 52);
    const double eps_hc = 4.0 * ulp + 2.9802322387695312e-08 /* 2^-25: fixed-point error of the fast path */;
    double e46 = ceil(eps_hc * 70368744177664.0 /* 2^46 */);
    if (e46 > 2147483648.0) e46 = 2147483648.0;
    double et = ceil(eps_hc * (sc.spc * 0.5) * 16777216.0 /* 2^24 */) + 8.0;
    if (et > 4194304.0) et = 4194304.0;
    uint32_t flags = visible ? 1u : 0u;
    {
        const uint64_t D = sc.lattice_den;
        bool near = true;
        if (D != 0) {
            const uint64_t r = ((U & kFracMask) * D) & kFracMask;
            const uint64_t dist = r < (kFracMask + 1 - r) ? r : (kFracMask + 1 - r);
            near = (double)dist < e46 * (double)D + 65536.0 * (double)D;
        }
        if (near) flags |= 2u;
    }

    // carrier NCO
    uint64_t phi;
    long long f, df;
    if (st.static_phase) {
        double fcyc;
        if (sc.flags & R4WB_FLAG_CLOSED_FORM_PHASE) {
            fcyc = ds / fs;
            f = (long long)llrint(fcyc * kTwo64);
            phi = (uint64_t)f * m_or_phi;
        } else {
            double x, step;
            static_phase_at(segs + st.seg_begin, st.seg_count, m_or_phi, x, step);
            f = (long long)llrint(step / (2.0 * kPi) * kTwo64);
            phi = cycles_to_fixed(x / (2.0 * kPi));
        }
        df = 0;
    } else {
        f = (long long)llrint(ds / fs * kTwo64);
        df = (long long)llrint((de - ds) / ((double)n * fs) * kTwo64);
        phi = m_or_phi;
    }

    o.U = U; o.phi = phi; o.f = f; o.df = df; o.phase0 = phase0; o.G = G; o.n = n; o.e0 = (uint32_t)e0;
    o.amp = (float)(amp * st.amp_scale); o.flags = flags; o.prev = -1; o.eps46 = (uint32_t)e46; o.eps_t = (uint32_t)et; o.pad = 0;
}

// phase advance of one block (cycles 0.64, wrapping): sum_{i<n} (f + i*df)
R4WB_HD uint64_t block_advance(const BlockSat& b)
{
    const uint64_t n = b.n;
    return n * (uint64_t)b.f + (uint64_t)b.df * (n * (n - 1) / 2);
}

// ---- the reference's carrier phase of a satellite whose Doppler changes (gnss/scenario.rs:516-527), exactly --------------
// The reference keeps ONE f64 per satellite and adds a freshly computed increment per sample for the whole run; by 600 s the
// rounding of those 3e9 additions has drifted ~2e-5 rad away from the real sum.  The drift is reproducible without walking
// the samples in order: while |phase| stays inside one binade [2^k, 2^(k+1)) it is a multiple of u = 2^(k-52), and
// fl(phase + inc) = phase + rint(inc / u) u (unless inc / u is an exact tie, when the parity of the running value decides).
// So a block's advance is the INTEGER sum of its rounded increments (block_phase_q, one thread per block and satellite), and
// only blocks that cross a binade, hold a tie, or start near zero are walked sample by sample (phase_walk).

// the per-sample increment, literally: frac = i / n; doppler = ds + frac (de - ds); inc = 2.0 * PI * doppler / fs
R4WB_HD double ref_phase_inc(double ds, double de, uint32_t i, double n_f64, double fs)
{
    const double frac = div_rn((double)i, n_f64);
    const double dop = add_rn(ds, mul_rn(frac, add_rn(de, -ds)));
    return div_rn(mul_rn(6.283185307179586, dop), fs);
}

// Smallest binade the integer-sum model is formed for.  Below it the phase crosses a binade (or an increment lands exactly
// half-way between two multiples of the ulp: probability ~2^(-8.6-k) per sample) in most blocks anyway, so they are walked
// sample by sample.  x = inc 2^(52-k) stays below 2^41 (rint trick: 2^51; the per-sample sum is formed in int64).
constexpr int kPhaseMinBinade = 5;

struct PhaseQ {
    long long Q;        // sum over the block of rint(inc_i / 2^(k-52))
    double approx;      // real-number sum of the block's increments (0 when the satellite is not visible)
    double span;        // bound of |partial sums| inside the block
    int k;              // binade the sum was formed for
    uint32_t ok;        // 1: Q usable when the phase stays in binade k over the block; 0: walk the samples
};

// approximate (real-number) advance of a block and a bound on the excursion of the partial sums
R4WB_HD void block_phase_approx(double ds, double de, uint32_t n, double fs, double* approx, double* span)
{
    const double w = 6.283185307179586 / fs;
    *approx = w * ((double)n * ds + (de - ds) * ((double)n - 1.0) * 0.5);
    const double m = fmax(fabs(ds), fabs(de));
    *span = ((ds < 0.0) != (de < 0.0)) ? w * m * (double)n : fabs(*approx);
}

R4WB_HD void block_phase_q(double ds, double de, uint32_t n, double fs, int k, long long* Q, bool* tie)
{
    long long q = 0;
    bool t = false;
    const double nf = (double)n;
    for (uint32_t i = 0; i < n; ++i) {
        const double x = scalbn(ref_phase_inc(ds, de, i, nf, fs), 52 - k);     // exact (power of two)
        const double fl = floor(x), r = x - fl;                                 // exact
        if (r == 0.5) t = true;
        q += (long long)fl + (r > 0.5 ? 1 : 0);
    }
    *Q = q; *tie = t;
}

// The same Q without visiting every sample.  Every operation of ref_phase_inc is a correctly rounded, monotone function of
// its argument, so x_i = inc_i 2^(52-k) is monotone in i (direction = sign of de - ds) and y_i = rint(x_i) is a monotone STEP
// function: over one 1 ms block the Doppler moves by ~1e-3 Hz and x by 0..a few dozen units.  With j running in the direction
// of non-decreasing y, and i_v = the first j with y_j >= v for every level v in (y_0, y_last],
//     Q = sum_j y_j = n y_0 + sum_v (n - i_v).
// Each i_v is estimated by linear interpolation and then PROVEN with the exact increment chain at i_v - 1 and i_v (moved until
// both hold), so Q is the brute-force sum bit for bit while the chain runs 2 + ~2 steps times instead of n.  A tie
// (x exactly half-way) can only sit on a sample next to a step or on the first / last sample — all of them are evaluated — and
// marks the block for the sample-by-sample walk, as in block_phase_q.  Returns false when the block has more than
// max_steps levels (start of the file, where the ulp is tiny): the caller then sums sample by sample.
// div: 0 = a / fs by division, 1 = Markstein with inv_fs (verified by the host for this fs; see k_phase_q)
R4WB_HD bool block_phase_q_steps(double ds, double de, uint32_t n, double fs, double inv_fs, int div, int k, uint32_t max_steps,
                                 long long* Q, bool* tie)
{
    if (n == 0) { *Q = 0; *tie = false; return true; }
    const double nf = (double)n;
    const double dd = add_rn(de, -ds);
    const double scale = scalbn(1.0, 52 - k), magic = 6755399441055744.0;    // 1.5 2^52: rint for |x| < 2^51
    const bool rev = dd < 0.0;
    bool t = false;
    double x_last_eval = 0.0;
    auto Y = [&](uint32_t j) -> long long {
        const uint32_t i = rev ? n - 1u - j : j;
        const double frac = div_rn((double)i, nf);
        const double a = mul_rn(6.283185307179586, add_rn(ds, mul_rn(frac, dd)));
        double inc;
        if (div) {
            const double q0 = mul_rn(a, inv_fs);
            inc = fma(fma(-q0, fs, a), inv_fs, q0);
        } else {
            inc = div_rn(a, fs);
        }
        const double x = mul_rn(inc, scale);
        const double y = add_rn(add_rn(x, magic), -magic);
        if (fabs(add_rn(x, -y)) == 0.5) t = true;
        x_last_eval = x;
        return (long long)y;
    };
    const long long y0 = Y(0);
    const double x0 = x_last_eval;
    if (!(fabs(x0) < 1125899906842624.0)) return false;                       // 2^50: outside the rint trick (never in practice)
    const long long yl = Y(n - 1u);
    const double xl = x_last_eval;
    if (yl < y0 || (unsigned long long)(yl - y0) > (unsigned long long)max_steps) return false;
    long long q = (long long)n * y0;
    if (yl > y0) {
        const double slope = ((double)(n - 1u)) / (xl - x0);                  // xl - x0 > 0 here
        for (long long v = y0 + 1; v <= yl; ++v) {
            double est = ceil(((double)v - 0.5 - x0) * slope);
            if (!(est >= 1.0)) est = 1.0;
            if (est > (double)(n - 1u)) est = (double)(n - 1u);
            uint32_t j = (uint32_t)est;
            while (Y(j) < v) ++j;                                             // y_{n-1} = yl >= v: stops by n - 1
            while (Y(j - 1u) >= v) --j;                                       // y_0 < v: stops by j = 1
            q += (long long)(n - j);
        }
    }
    *Q = q; *tie = t;
    return true;
}

// true when a phase that starts the block at `ph` provably stays inside binade k for every sample of the block
R4WB_HD bool phase_stays_in_binade(double ph, const PhaseQ& r)
{
    if (!r.ok || ph == 0.0 || ilogb(ph) != r.k) return false;
    const double a = fabs(ph), lo = scalbn(1.0, r.k), hi = scalbn(1.0, r.k + 1);
    const double margin = 1e-3 + 1e-9 * a;
    // partial sums lie between ph and ph + approx (same-sign Doppler) or within +-span of ph
    const double e = (ph < 0.0) ? -r.approx : r.approx;                          // advance of |phase|
    const double amin = fmin(a, a + e), amax = fmax(a, a + e);
    const bool mono = r.span == fabs(r.approx);
    const double l = mono ? amin : a - r.span, h = mono ? amax : a + r.span;
    return l - margin > lo && h + margin < hi;
}

// the block's samples one by one (the reference's own loop)
R4WB_HD double phase_walk(double ph, double ds, double de, uint32_t n, double fs)
{
    const double nf = (double)n;
    for (uint32_t i = 0; i < n; ++i) ph = add_rn(ph, ref_phase_inc(ds, de, i, nf, fs));
    return ph;
}

// phase at the end of a block that starts at ph
R4WB_HD double phase_after_block(double ph, const PhaseQ& r, double ds, double de, uint32_t n, double fs)
{
    if (phase_stays_in_binade(ph, r)) return add_rn(ph, scalbn((double)r.Q, r.k - 52));   // both multiples of u, sum inside the binade: exact
    return phase_walk(ph, ds, de, n, fs);
}

// ----------------------------------------------------------------------------------------------
// packed pairs: a float2 holds the same quantity for the two adjacent samples (a, b) a thread owns; on sm_100a
// one FFMA2 / FMUL2 / FADD2 processes both
R4WB_HD float2 pk_fma(float2 a, float2 b, float2 c)
{
#ifdef __CUDA_ARCH__
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d)
        : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)), "l"(reinterpret_cast<unsigned long long&>(c)));
    return reinterpret_cast<float2&>(d);
#else
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
R4WB_HD float2 pk_mul(float2 a, float2 b)
{
#ifdef __CUDA_ARCH__
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
    return reinterpret_cast<float2&>(d);
#else
    return make_float2(a.x * b.x, a.y * b.y);
#endif
}
R4WB_HD float2 pk_add(float2 a, float2 b)
{
#ifdef __CUDA_ARCH__
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
    return reinterpret_cast<float2&>(d);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
// low 32 bits of ((hi:lo) >> min(sh, 32))
R4WB_HD uint32_t funnel_rc(uint32_t lo, uint32_t hi, uint32_t sh)
{
#ifdef __CUDA_ARCH__
    return __funnelshift_rc(lo, hi, sh);
#else
    return (uint32_t)(((((uint64_t)hi) << 32) | (uint64_t)lo) >> (sh < 32u ? sh : 32u));
#endif
}
R4WB_HD void accurate_sincos_cycles(uint64_t ph, float* s, float* c)
{
    // phase in cycles (0.64 fixed) -> sin/cos with full f32 accuracy (used once per tile and satellite)
    const float x = (float)(int32_t)(ph >> 32) * 4.656612873077393e-10f;   // * 2^-31 -> half-cycles in [-1, 1)
#ifdef __CUDA_ARCH__
    sincospif(x, s, c);
#else
    *s = (float)sin(3.141592653589793 * (double)x);
    *c = (float)cos(3.141592653589793 * (double)x);
#endif
}

// ----------------------------------------------------------------------------------------------
// constants of one launch, derived from SynthArgs once per thread
struct SynthK {
    uint64_t delta46, d8;       // half-chips per oversample / per output sample, 2^-46 units
    uint64_t step32;            // half-chips per 2*kSynthThreads samples, 32.32 fixed
    uint32_t d8_32;             // half-chips per output sample, 0.32 fixed (< 1)
    uint32_t kmul, c1, c2, c3, dsum0;
    uint32_t lut_den, ystride;
    double spc;
};

// position of fraction bin q in the class table: bits 2-3 are xor-ed with bits 8-9 so the bins a warp touches
// (a fixed stride of 16 * 227 mod 20000 between lanes at 5 MHz) spread over the banks
R4WB_HD uint32_t cls_lut_index(uint32_t q) { return q ^ ((q >> 6) & 0xcu); }

R4WB_HD SynthK make_synth_k(uint64_t delta46, uint32_t kmul, const uint32_t* cj, uint32_t dsum0, double spc, uint32_t lut_den, uint32_t ystride)
{
    SynthK k;
    k.delta46 = delta46;
    k.d8 = delta46 * (uint64_t)kOversample;
    k.step32 = (k.d8 * (uint64_t)(2 * kSynthThreads)) >> (kFracBits - 32);
    k.d8_32 = (uint32_t)(k.d8 >> (kFracBits - 32));
    k.kmul = kmul; k.c1 = cj[1]; k.c2 = cj[2]; k.c3 = cj[3]; k.dsum0 = dsum0; k.lut_den = lut_den; k.ystride = ystride;
    k.spc = spc;
    return k;
}

R4WB_HD bool continues_prev(const BlockSat& cur, const BlockSat* __restrict__ tab);

R4WB_HD TileSat tile_sat(const BlockSat& b, const BlockSat* __restrict__ tab, uint32_t i_begin, uint64_t d8)
{
    TileSat t;
    t.u0 = b.U + (uint64_t)i_begin * d8;
    const uint32_t h0 = (uint32_t)(t.u0 >> kFracBits);
    t.hb = h0 - (uint32_t)(kJ + 2);            // may go "negative": all uses are modulo differences
    t.phi = b.phi; t.f = b.f; t.df = b.df; t.amp = b.amp; t.eps_t = b.eps_t;
    uint32_t fl = b.flags & 3u;
    if (b.flags & 32u) fl &= ~1u;                 // direct-path satellite: not rendered by k_synth
    // rotation over one step of 2*kSynthThreads samples, starting at sample i: phase(i + n) - phase(i) with
    // phase(i) = phi + (i+1) f + i(i+1)/2 df
    const uint64_t n = (uint64_t)(2 * kSynthThreads);
    const uint64_t step = n * (uint64_t)b.f + (uint64_t)b.df * (n * (uint64_t)i_begin + n * (n + 1) / 2);
    accurate_sincos_cycles(step, &t.wi, &t.wr);
    const float two_pi_cyc = 3.4061215800865545e-19f;      // 2 pi / 2^64
    t.th1 = (float)((double)b.df * (double)n) * two_pi_cyc;
    t.th2 = (float)((double)b.df * (double)(n * n)) * two_pi_cyc;
    if (b.df != 0) {
        fl |= 16u;
        // the linearised step phasor is good to (th1 * tile)^2 / 2 and (th2 * K)^2 / 2: fall back to per-sample sincos beyond 1e-4
        if (fabsf(t.th1) * 8192.0f > 1e-4f || fabsf(t.th2) * 16.0f > 1e-4f) fl |= 4u;
    }
    // the first eight windows of a block reach into its predecessor: collapsed form across the boundary only when the code
    // sequence continues exactly AND neither block has a boundary inside the f64 rounding band (fir_block_start is literal then)
    if (i_begin == 0 && (!continues_prev(b, tab) || (b.flags & 2u) || (tab[b.prev].flags & 2u))) fl |= 8u;
    t.flags = fl;
    return t;
}

// sign bit (1 <=> -1) of half-chip hc in [0, 204600): primary code x BOC(1,1) from the period table, x secondary code
R4WB_HD uint32_t halfchip_sign(const uint32_t* __restrict__ per, uint32_t hc, const SatCode& cd)
{
    const uint32_t e = hc / cd.per_len, p = hc - e * cd.per_len;
    return ((per[p >> 5] >> (p & 31u)) ^ (uint32_t)(cd.epoch_bits >> e)) & 1u;
}
R4WB_HD uint32_t code_bit(const uint32_t* __restrict__ per, uint32_t c) { return (per[c >> 4] >> ((2u * c) & 31u)) & 1u; }

// word w of the per-tile half-chip sign table: bit n <-> half-chip hb + 32 w + n (1 <=> -1)
R4WB_HD uint32_t sign_word(const uint32_t* __restrict__ per, uint32_t hb, uint32_t w, const SatCode& cd)
{
    int64_t hh = (int64_t)(int32_t)hb + 32 * (int64_t)w;
    hh %= (int64_t)cd.hc_mod;
    if (hh < 0) hh += cd.hc_mod;
    const uint32_t h = (uint32_t)hh, e = h / cd.per_len, p = h - e * cd.per_len;
    const uint32_t bits = funnel_r(per[p >> 5], per[(p >> 5) + 1], p);       // the table repeats its first 64 bits at the end
    const uint32_t e1 = e + 1 == cd.epoch_period ? 0u : e + 1;
    const uint32_t se = 0u - ((uint32_t)(cd.epoch_bits >> e) & 1u), se1 = 0u - ((uint32_t)(cd.epoch_bits >> e1) & 1u);
    const uint32_t nlow = cd.per_len - p;                                     // bits of this word inside epoch e
    const uint32_t lowmask = nlow >= 32u ? 0xffffffffu : ((1u << nlow) - 1u);
    return bits ^ ((se & lowmask) | (se1 & ~lowmask));
}

// the same word for a table origin given as (epoch e0, half-chip p0 inside the period): no division
R4WB_HD uint32_t sign_word_ep(const uint32_t* __restrict__ per, uint32_t e0, uint32_t p0, uint32_t w, const SatCode& cd)
{
    uint32_t p = p0 + 32u * w, e = e0;
    while (p >= cd.per_len) { p -= cd.per_len; e = e + 1u == cd.epoch_period ? 0u : e + 1u; }
    const uint32_t bits = funnel_r(per[p >> 5], per[(p >> 5) + 1], p);
    const uint32_t e1 = e + 1u == cd.epoch_period ? 0u : e + 1u;
    const uint32_t se = 0u - ((uint32_t)(cd.epoch_bits >> e) & 1u), se1 = 0u - ((uint32_t)(cd.epoch_bits >> e1) & 1u);
    const uint32_t nlow = cd.per_len - p;
    const uint32_t lowmask = nlow >= 32u ? 0xffffffffu : ((1u << nlow) - 1u);
    return bits ^ ((se & lowmask) | (se1 & ~lowmask));
}

// sign bit (1 <=> -1) of oversample q (>= 0, relative to the block start) of block entry bs;
// evaluates the reference expression literally when q is within the rounding band of a boundary.
R4WB_HD_NOINLINE uint32_t chip_sign_exact(const BlockSat& bs, long long q, const uint32_t* __restrict__ per, double spc, const SatCode& cd)
{
    const double g = (double)(bs.G + (uint64_t)q);
    const double cf = add_rn(bs.phase0, div_rn(g, spc));                       // satellite_emitter.rs:268
    // `cf % code_length` (:269, Rust % = fmod) is exact for 0 <= cf < 2^53: floor(cf) mod code_length + frac(cf), so the chip
    // index `(..) as usize` is an integer remainder (always <= code_length - 1, the clamp of :281 never binds)
    const double fl = floor(cf);
    const uint32_t c = cf > 0.0 ? (uint32_t)((uint64_t)fl % cd.code_len) : 0u;
    const double cp = cf - fl;                                                   // :270
    const double eq = div_rn(cf, (double)cd.code_len);
    const uint64_t ep = (uint64_t)bs.e0 + (eq > 0.0 ? (uint64_t)eq : 0ull);      // :278 (the f64 quotient, literally: it may round up to an integer)
    const uint32_t boc = (cd.has_boc && !(cp < 0.5)) ? 1u : 0u;                  // :303-305: (cp * 2) % 2 < 1 with cp in [0, 1) <=> cp < 0.5, exactly
    return code_bit(per, c) ^ boc ^ ((uint32_t)(cd.epoch_bits >> (uint32_t)(ep % cd.epoch_period)) & 1u);
}

R4WB_HD uint32_t chip_sign(const BlockSat& bs, long long q, const uint32_t* __restrict__ per, uint64_t delta46, double spc, const SatCode& cd)
{
    const uint64_t u = bs.U + (uint64_t)q * delta46;
    const uint64_t fr = u & kFracMask;
    if (fr < bs.eps46 || fr > kFracMask - bs.eps46) return chip_sign_exact(bs, q, per, spc, cd);
    uint32_t h = (uint32_t)(u >> kFracBits);
    if (h >= cd.hc_mod) h -= cd.hc_mod;
    return halfchip_sign(per, h, cd);
}

// sign of tap k of the window whose newest oversample is g (relative to the start of `cur`): 0 <=> +1, 1 <=> -1, -1 <=> the
// tap reaches into the zero-initialised delay line (or history older than one block: not modelled) and contributes nothing
R4WB_HD int tap_sign(const BlockSat& cur, const BlockSat* __restrict__ tab, const uint32_t* __restrict__ per, long long g, int k,
                     uint64_t delta46, double spc, const SatCode& cd)
{
    long long q = g - k;
    if (q >= 0) return (int)chip_sign(cur, q, per, delta46, spc, cd);
    if (cur.prev < 0) return -1;
    const BlockSat& pb = tab[cur.prev];
    q += (long long)kOversample * pb.n;
    if (q < 0) return -1;
    return (int)chip_sign(pb, q, per, delta46, spc, cd);
}

// direct 63-tap evaluation of output sample i of block entry `cur` (history from `prev`): the reference's own loop
// (fir.rs:392-409 over satellite_emitter.rs:264-330), used for samples whose window touches an ambiguous boundary.
R4WB_HD_NOINLINE float fir_direct(const BlockSat& cur, const BlockSat* __restrict__ tab, const uint32_t* __restrict__ per,
                                  const float* __restrict__ taps, int i, uint64_t delta46, double spc, const SatCode& cd)
{
    float acc = 0.0f;
    const long long g = (long long)kOversample * i;
    for (int k = 0; k < kTaps; ++k) {
        const int sgn = tap_sign(cur, tab, per, g, k, delta46, spc, cd);
        if (sgn < 0) continue;
        acc += sgn ? -taps[k] : taps[k];
    }
    return acc;
}

#ifdef __CUDACC__
// fir_direct for the lanes of a warp that need it, evaluated by the WHOLE warp.  A lane owns the adjacent samples i and i + 1
// (`need_a`, `need_b`).  The lanes that meet an ambiguous boundary own consecutive samples (the boundary stays inside the
// 63-tap window for eight samples), so their windows overlap: requesters whose windows fit into one span of 128 oversamples
// form a group, lane l resolves the signs of oversamples l, l + 32, l + 64, l + 96 of the span (the expensive part: the
// reference's literal f64 chip index for the one oversample inside the rounding band — once per group, not once per requester
// and tap), eight ballots publish them, and every requester adds its 63 taps in the reference's order: the same f32 sum as
// fir_direct, bit for bit.  Every lane of the warp must call this (converged); lane order = sample order.
static __device__ __noinline__ float2 fir_direct_warp(bool need_a, bool need_b, int i, float2 y, const BlockSat& cur,
                                                      const BlockSat* __restrict__ tab, const uint32_t* __restrict__ per,
                                                      const float* __restrict__ taps, uint64_t delta46, double spc, const SatCode& cd)
{
    unsigned m = __ballot_sync(0xffffffffu, need_a || need_b);
    const int lane = (int)(threadIdx.x & 31u);
    const long long g_lo = (long long)kOversample * (need_a ? i : i + 1), g_hi = (long long)kOversample * (need_b ? i + 1 : i);
    while (m) {
        const int src = __ffs((int)m) - 1;
        const long long g0 = __shfl_sync(0xffffffffu, g_lo, src);              // the group's oldest window
        const bool mine = ((m >> lane) & 1u) && g_lo >= g0 && g_hi - g0 <= 128 - kTaps;
        const unsigned grp = __ballot_sync(0xffffffffu, mine);                  // contains src
        const long long base = g0 - (kTaps - 1);                                // span = oversamples [base, base + 128)
        unsigned neg[4], val[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int sgn = tap_sign(cur, tab, per, base + lane + 32 * r, 0, delta46, spc, cd);
            neg[r] = __ballot_sync(0xffffffffu, sgn == 1);
            val[r] = __ballot_sync(0xffffffffu, sgn >= 0);
        }
        if (mine) {
            const unsigned long long negA = neg[0] | ((unsigned long long)neg[1] << 32), negB = neg[2] | ((unsigned long long)neg[3] << 32);
            const unsigned long long valA = val[0] | ((unsigned long long)val[1] << 32), valB = val[2] | ((unsigned long long)val[3] << 32);
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                if (!(h ? need_b : need_a)) continue;
                // the window's 63 signs as one word: bit j <-> span index lo + j <-> tap 62 - j
                const int lo = (int)((long long)kOversample * (i + h) - g0);    // 0 .. 128 - kTaps
                unsigned long long wn, wv;
                if (lo == 0) { wn = negA; wv = valA; }
                else if (lo < 64) { wn = (negA >> lo) | (negB << (64 - lo)); wv = (valA >> lo) | (valB << (64 - lo)); }
                else { wn = negB >> (lo - 64); wv = valB >> (lo - 64); }
                float acc = 0.0f;
#pragma unroll
                for (int k = 0; k < kTaps; ++k) {                               // the reference's order; an absent tap adds +0 (exact)
                    const int j = kTaps - 1 - k;
                    const float t = taps[k];
                    acc += ((wv >> j) & 1ull) ? (((wn >> j) & 1ull) ? -t : t) : 0.0f;
                }
                if (h) y.y = acc; else y.x = acc;
            }
        }
        m &= ~grp;
    }
    return y;
}
#endif

// ---- direct path (chip rates other than 1.023 MHz: GPS L5, GLONASS L1OF) -------------------------------------------------
// sign bit (1 <=> -1) of oversample q of block entry bs: the reference expression, literally (satellite_emitter.rs:264-292,
// plain BPSK branch :334-336): chip = (cf % code_length) as usize, epoch = e0 + (cf / code_length) as usize, nav bit per epoch
R4WB_HD uint32_t direct_sign(const BlockSat& bs, long long q, const uint32_t* __restrict__ code, const DirectSat& d)
{
    const double g = (double)(bs.G + (uint64_t)q);
    const double cf = add_rn(bs.phase0, div_rn(g, d.spc));
    const double cl = (double)d.code_len;
    const double cm = fmod(cf, cl);
    uint32_t c = cm > 0.0 ? (uint32_t)cm : 0u;
    if (c > d.code_len - 1u) c = d.code_len - 1u;
    const double eq = div_rn(cf, cl);
    const uint64_t ep = (uint64_t)bs.e0 + (eq > 0.0 ? (uint64_t)eq : 0ull);
    return ((code[c >> 5] >> (c & 31u)) ^ (uint32_t)(d.epoch_bits >> (uint32_t)(ep % d.epoch_period))) & 1u;
}

// the reference's 63-tap loop (fir.rs:392-409) for output sample i of block entry `cur`, FIR history from its predecessor.
// The chip index and the primary-code epoch are carried from tap to tap (consecutive taps are 1 / spc < 1 chips apart, so
// the chip index steps down by 0 or 1); one 64-bit division per block entry touched.  The fast position phase0 + (G + q) *
// (1 / spc) is within ~5e-8 chips of the reference's f64 value (cf < 2^28 chips): when its fraction keeps 1e-6 clear of a
// chip boundary its floor IS the reference's floor; an oversample closer than that takes direct_sign (the literal
// expression, also for the epoch), so the result is the reference's for every tap.
R4WB_HD float direct_fir(const BlockSat& cur, const BlockSat* __restrict__ tab, const uint32_t* __restrict__ code,
                         const float* __restrict__ taps, int i, const DirectSat& d)
{
    double acc = 0.0;
    const long long g = (long long)kOversample * i;
    const double inv_spc = 1.0 / d.spc;
    const BlockSat* bs = &cur;
    long long shift = 0;                 // q (relative to bs) = g - k + shift
    bool primed = false;
    double c_prev = 0.0;
    uint32_t chip = 0, ep = 0;
    for (int k = 0; k < kTaps; ++k) {
        long long q = g - k + shift;
        if (q < 0) {
            if (bs != &cur || cur.prev < 0) break;           // zero-initialised delay line / history older than one block: not modelled
            bs = &tab[cur.prev];
            shift = (long long)kOversample * bs->n;
            q += shift;
            if (q < 0) break;
            primed = false;
        }
        const double gq = (double)(bs->G + (uint64_t)q);
        const double cf = fma(gq, inv_spc, bs->phase0);
        const double fl = floor(cf), fr = cf - fl;
        uint32_t sgn;
        if (!(fr > 1e-6 && fr < 1.0 - 1e-6)) {               // too close to a chip boundary for the fast form: literal expression
            sgn = direct_sign(*bs, q, code, d);
            primed = false;
        } else {
            if (!primed) {
                const uint64_t c = (uint64_t)fl;
                chip = (uint32_t)(c % d.code_len);
                ep = (uint32_t)(((uint64_t)bs->e0 + c / d.code_len) % d.epoch_period);
                primed = true;
            } else if (fl != c_prev) {                        // one chip back
                if (chip == 0u) { chip = d.code_len - 1u; ep = ep == 0u ? d.epoch_period - 1u : ep - 1u; }
                else --chip;
            }
            c_prev = fl;
            sgn = ((code[chip >> 5] >> (chip & 31u)) ^ (uint32_t)(d.epoch_bits >> ep)) & 1u;
        }
        acc += sgn ? -(double)taps[k] : (double)taps[k];
    }
    return (float)acc;
}

// one direct-path satellite's contribution to output sample i of its block: amp * y * exp(j phase(i))
R4WB_HD float2 direct_sample(const BlockSat& cur, const BlockSat* __restrict__ tab, const uint32_t* __restrict__ code,
                             const float* __restrict__ taps, uint32_t i, const DirectSat& d)
{
    const float y = direct_fir(cur, tab, code, taps, (int)i, d);
    const uint64_t ph = cur.phi + (uint64_t)(i + 1) * (uint64_t)cur.f + ((uint64_t)i * (i + 1) / 2) * (uint64_t)cur.df;
    float sn, cs;
    accurate_sincos_cycles(ph, &sn, &cs);
    return make_float2(cur.amp * y * cs, cur.amp * y * sn);
}

// true when block `cur` continues its predecessor's code sequence exactly (same delay, adjacent in time): the
// collapsed FIR may then look back across the block boundary through the sign table
R4WB_HD bool continues_prev(const BlockSat& cur, const BlockSat* __restrict__ tab)
{
    if (cur.prev < 0) return false;
    const BlockSat& pb = tab[cur.prev];
    return (pb.flags & 1u) && pb.phase0 == cur.phase0 && pb.e0 == cur.e0 && pb.G + (uint64_t)kOversample * pb.n == cur.G;
}

// boundary ages d_j = floor(t0 + j S) of the half-chip boundaries inside the window (t0 = oversamples since the newest one)
R4WB_HD void boundary_ages(uint32_t frac32, const SynthK& K, uint32_t d[4])
{
    const uint32_t t0 = umulhi32(frac32, K.kmul);
    d[0] = t0 >> kTBits; d[1] = (t0 + K.c1) >> kTBits; d[2] = (t0 + K.c2) >> kTBits; d[3] = (t0 + K.c3) >> kTBits;
}

// Collapsed FIR restricted to taps k <= g of the window whose newest oversample sits at half-chip position u of the
// code sequence: s_4 E[g] + sum_j (s_j - s_{j+1}) E[min(d_j, g)]   (g = 62 gives the whole window)
R4WB_HD float fir_collapsed_upto(uint64_t u, uint32_t g, const uint32_t* __restrict__ per, const float* __restrict__ etab, const SynthK& K,
                                 const SatCode& cd)
{
    uint32_t h = (uint32_t)(u >> kFracBits);
    uint32_t d[4];
    boundary_ages((uint32_t)(u >> (kFracBits - 32)), K, d);
    float s[kJ + 1];
#pragma unroll
    for (int j = 0; j <= kJ; ++j) {
        const uint32_t hc = (h + cd.hc_mod - (uint32_t)j) % cd.hc_mod;
        s[j] = halfchip_sign(per, hc, cd) ? -1.0f : 1.0f;
    }
    float y = s[kJ] * etab[g];
#pragma unroll
    for (int j = 0; j < kJ; ++j) y = fmaf(s[j] - s[j + 1], etab[d[j] < g ? d[j] : g], y);
    return y;
}

// Output sample i (< 8) of a block whose code delay differs from its predecessor's: taps reaching oversamples >= 0 see
// the block's own sequence, older taps the tail of the previous visible block (the FIR delay line persists across
// blocks, core/filters/fir.rs:392-409 + gnss/scenario.rs:486-489).  Both parts in collapsed form; the literal 63-tap
// loop only when either block has a boundary inside the f64 rounding band.
R4WB_HD float fir_block_start(const BlockSat& cur, const BlockSat* __restrict__ tab, const uint32_t* __restrict__ per,
                              const float* __restrict__ taps, const float* __restrict__ etab, int i, const SynthK& K, const SatCode& cd)
{
    const bool has_prev = cur.prev >= 0;
    if ((cur.flags & 2u) || (has_prev && (tab[cur.prev].flags & 2u))) return fir_direct(cur, tab, per, taps, i, K.delta46, K.spc, cd);
    const uint32_t g = (uint32_t)(kOversample * i);
    float y = fir_collapsed_upto(cur.U + (uint64_t)g * K.delta46, g, per, etab, K, cd);
    if (has_prev) {
        const BlockSat& pb = tab[cur.prev];
        const uint64_t up = pb.U + ((uint64_t)kOversample * pb.n + g) * K.delta46;   // the old sequence, continued
        y += fir_collapsed_upto(up, kTaps - 1, per, etab, K, cd) - fir_collapsed_upto(up, g, per, etab, K, cd);
    }
    return y;
}

// carrier phase (cycles 0.64) of sample i of a block: phi + (i+1) f + i(i+1)/2 df   (scenario.rs:519-524)
R4WB_HD uint64_t carrier_phase(const TileSat& ts, uint32_t i)
{
    return ts.phi + (uint64_t)(i + 1) * (uint64_t)ts.f + ((uint64_t)i * (i + 1) / 2) * (uint64_t)ts.df;
}

R4WB_HD void phasor(uint64_t ph, float* s, float* c)
{
    fast_sincos((float)(int32_t)(ph >> 32) * 1.4629180792671596e-09f /* 2 pi / 2^32 */, s, c);
}

// what the slow paths of sat_accumulate need
struct SlowCtx {
    const BlockSat* cur;           // this block's entry for the satellite
    const BlockSat* tab;           // whole table (history links)
    const uint32_t* per;           // the satellite's period table
    const float* taps;
    const SatCode* code;           // the satellite's code structure
};

// One satellite's contribution to the NK sample pairs thread `tid` owns in a tile: pair k = samples
// (i_begin + 2 tid + 2 kSynthThreads k, +1).  ar[k] / ai[k] accumulate (re_a, re_b) / (im_a, im_b).
//   t64   [nw64] per-tile sign table as overlapping 64-bit windows (word k = bits 32k .. 32k+63)
//   ytab  [32][K.ystride] collapsed-FIR outputs per (sign pattern, boundary-age class)
//   yfix  first 8 outputs of the block when its window reaches into a block with another delay (flags bit3)
//   GENERAL = false: no ambiguity checks, phasor recurrence, DYN compile-time (the common case, straight-line code)
//   GENERAL = true:  checks / per-sample sincos / varying Doppler selected at run time from ts.flags
//   LUT: boundary-age classes from the fraction table (only with GENERAL = false)
template <int NK, bool GENERAL, bool DYN, bool LUT>
R4WB_HD void sat_accumulate_t(const TileSat& ts, const SynthK& K, const uint2* __restrict__ t64, const float* __restrict__ ytab,
                            const uint8_t* __restrict__ clslut, const float* __restrict__ yfix, const SlowCtx& slow, uint32_t tid, uint32_t i_begin, uint32_t i_end,
                            float2 (&ar)[NK], float2 (&ai)[NK], uint64_t* n_ambiguous)
{
    const uint32_t ia0 = i_begin + 2u * tid;
    // code position of sample a, 32.32 fixed relative to bit 0 of the pattern field: hi = h - hb - J, lo = fraction
    const uint64_t ua = ts.u0 + (uint64_t)(2u * tid) * K.d8;
    uint64_t pos = (ua - ((uint64_t)(ts.hb + (uint32_t)kJ) << kFracBits)) >> (kFracBits - 32);

    // carrier: exact 64-bit start phases, then a phasor recurrence over the strided pairs
    const uint64_t pha = carrier_phase(ts, ia0);
    const uint64_t gb = (uint64_t)ts.f + (uint64_t)ts.df * (uint64_t)(ia0 + 1);      // increment applied at sample ia0 + 1
    float sa, ca, sb, cb;
    phasor(pha, &sa, &ca);
    phasor(pha + gb, &sb, &cb);
    float2 zr = make_float2(ca * ts.amp, cb * ts.amp), zi = make_float2(sa * ts.amp, sb * ts.amp);
    const bool dyn = GENERAL ? (ts.flags & 16u) != 0u : DYN;
    const bool exact_rot = GENERAL && (ts.flags & 4u) != 0u, check = GENERAL && (ts.flags & 2u) != 0u;
    float wr = ts.wr, wi = ts.wi, dr = 0.0f, di = 0.0f;
    if (dyn) {
        const float dl = ts.th1 * (float)(2u * tid);
        wr = fmaf(-dl, ts.wi, ts.wr);
        wi = fmaf(dl, ts.wr, ts.wi);
        dr = -ts.th2 * wi;
        di = ts.th2 * wr;
    }
    float2 w_r = make_float2(wr, wr), w_i = make_float2(wi, wi), w_ni = make_float2(-wi, -wi);
    const float2 d_r = make_float2(dr, dr), d_i = make_float2(di, di), d_ni = make_float2(-di, -di);
    // exact-rotation mode state
    uint64_t ph_k = pha, gb_k = gb;
    uint64_t step = (uint64_t)(2 * kSynthThreads) * (uint64_t)ts.f +
                    (uint64_t)ts.df * ((uint64_t)(2 * kSynthThreads) * ia0 + (uint64_t)(2 * kSynthThreads) * (2 * kSynthThreads + 1) / 2);
    const uint64_t step2 = (uint64_t)ts.df * (uint64_t)(2 * kSynthThreads) * (uint64_t)(2 * kSynthThreads);
    const uint64_t g_n = (uint64_t)ts.df * (uint64_t)(2 * kSynthThreads);

#pragma unroll
    for (int k = 0; k < NK; ++k) {
        const uint32_t hi = (uint32_t)(pos >> 32), lo = (uint32_t)pos;
        const uint32_t lob = lo + K.d8_32, hib = hi + (lob < lo ? 1u : 0u);
        // boundary-age classes
        uint32_t ta0 = 0, ta1 = 0, ta2 = 0, ta3 = 0, tb0 = 0, tb1 = 0, tb2 = 0, tb3 = 0, cls_a, cls_b;
        if (LUT && !GENERAL) {
            cls_a = clslut[cls_lut_index(umulhi32(lo, K.lut_den))];
            cls_b = clslut[cls_lut_index(umulhi32(lob, K.lut_den))];
        } else {
            ta0 = umulhi32(lo, K.kmul); ta1 = ta0 + K.c1; ta2 = ta0 + K.c2; ta3 = ta0 + K.c3;
            tb0 = umulhi32(lob, K.kmul); tb1 = tb0 + K.c1; tb2 = tb0 + K.c2; tb3 = tb0 + K.c3;
            cls_a = (ta0 >> kTBits) + (ta1 >> kTBits) + (ta2 >> kTBits) + (ta3 >> kTBits) - K.dsum0;
            cls_b = (tb0 >> kTBits) + (tb1 >> kTBits) + (tb2 >> kTBits) + (tb3 >> kTBits) - K.dsum0;
        }
        // 5-sign patterns out of one 64-bit window
        const uint2 w = *reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned char*>(t64) + ((hi >> 2) & 0x3ffffff8u));
        const uint32_t pa = funnel_r(w.x, w.y, hi) & 31u;
        const uint32_t pb = funnel_rc(w.x, w.y, (hi & 31u) + (hib - hi)) & 31u;
        float2 y = make_float2(ytab[pa * K.ystride + cls_a], ytab[pb * K.ystride + cls_b]);
        if (check) {
            const uint32_t m = (1u << kTBits) - 1u, e = ts.eps_t;
            const bool amb_a = ((ta0 + e) & m) < 2 * e || ((ta1 + e) & m) < 2 * e || ((ta2 + e) & m) < 2 * e || ((ta3 + e) & m) < 2 * e;
            const bool amb_b = ((tb0 + e) & m) < 2 * e || ((tb1 + e) & m) < 2 * e || ((tb2 + e) & m) < 2 * e || ((tb3 + e) & m) < 2 * e;
            const uint32_t ia = ia0 + 2u * (uint32_t)kSynthThreads * (uint32_t)k;
#ifdef __CUDA_ARCH__
            // device: the warp shares the literal evaluation (all lanes are here: `check` and k are uniform)
            const bool need_a = amb_a && ia < i_end, need_b = amb_b && ia + 1 < i_end;
            if (__any_sync(0xffffffffu, need_a || need_b))
                y = fir_direct_warp(need_a, need_b, (int)ia, y, *slow.cur, slow.tab, slow.per, slow.taps, K.delta46, K.spc, *slow.code);
#else
            if (amb_a && ia < i_end) {
                y.x = fir_direct(*slow.cur, slow.tab, slow.per, slow.taps, (int)ia, K.delta46, K.spc, *slow.code);
                if (n_ambiguous) ++*n_ambiguous;
            }
            if (amb_b && ia + 1 < i_end) {
                y.y = fir_direct(*slow.cur, slow.tab, slow.per, slow.taps, (int)ia + 1, K.delta46, K.spc, *slow.code);
                if (n_ambiguous) ++*n_ambiguous;
            }
#endif
        }
        if (k == 0 && (ts.flags & 8u) && tid < 4u) y = make_float2(yfix[2u * tid], yfix[2u * tid + 1u]);
        if (exact_rot) {
            float s0, c0, s1, c1;
            phasor(ph_k, &s0, &c0);
            phasor(ph_k + gb_k, &s1, &c1);
            zr = make_float2(c0 * ts.amp, c1 * ts.amp);
            zi = make_float2(s0 * ts.amp, s1 * ts.amp);
            ph_k += step; step += step2; gb_k += g_n;
        }
        ar[k] = pk_fma(y, zr, ar[k]);
        ai[k] = pk_fma(y, zi, ai[k]);
        if (!exact_rot && k + 1 < NK) {
            const float2 nr = pk_fma(zr, w_r, pk_mul(zi, w_ni));
            const float2 ni = pk_fma(zr, w_i, pk_mul(zi, w_r));
            zr = nr; zi = ni;
            if (dyn) { w_r = pk_add(w_r, d_r); w_i = pk_add(w_i, d_i); w_ni = pk_add(w_ni, d_ni); }
        }
        pos += K.step32;
    }
}

template <int NK>
R4WB_HD void sat_accumulate(const TileSat& ts, const SynthK& K, const uint2* __restrict__ t64, const float* __restrict__ ytab,
                            const uint8_t* __restrict__ clslut, const float* __restrict__ yfix, const SlowCtx& slow, uint32_t tid,
                            uint32_t i_begin, uint32_t i_end, float2 (&ar)[NK], float2 (&ai)[NK], uint64_t* n_ambiguous)
{
    if (ts.flags & 6u) sat_accumulate_t<NK, true, false, false>(ts, K, t64, ytab, clslut, yfix, slow, tid, i_begin, i_end, ar, ai, n_ambiguous);
    else if (K.lut_den) {
        if (ts.flags & 16u) sat_accumulate_t<NK, false, true, true>(ts, K, t64, ytab, clslut, yfix, slow, tid, i_begin, i_end, ar, ai, n_ambiguous);
        else sat_accumulate_t<NK, false, false, true>(ts, K, t64, ytab, clslut, yfix, slow, tid, i_begin, i_end, ar, ai, n_ambiguous);
    } else {
        if (ts.flags & 16u) sat_accumulate_t<NK, false, true, false>(ts, K, t64, ytab, clslut, yfix, slow, tid, i_begin, i_end, ar, ai, n_ambiguous);
        else sat_accumulate_t<NK, false, false, false>(ts, K, t64, ytab, clslut, yfix, slow, tid, i_begin, i_end, ar, ai, n_ambiguous);
    }
}

// ----------------------------------------------------------------------------------------------
// noise: Philox4x32-10 keyed by the scenario seed, counter = global sample index / 2 (one draw -> two samples)
struct PhiloxKeys { uint32_t k0[10], k1[10]; };

R4WB_HD PhiloxKeys philox_keys(uint64_t seed)
{
    PhiloxKeys K;
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) { K.k0[r] = a; K.k1[r] = b; a += 0x9E3779B9u; b += 0xBB67AE85u; }
    return K;
}

R4WB_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& K,
                           uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = umulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = umulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ K.k0[r]; c1 = lo1; c2 = hi0 ^ c3 ^ K.k1[r]; c3 = lo0;
    }
    r0 = c0; r1 = c1; r2 = c2; r3 = c3;
}

// one round of Philox4x32 (key words of the round passed in)
R4WB_HD void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1)
{
    const uint32_t hi0 = umulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = umulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
}

// uint32 -> float, rounded toward zero (never reaches 2^32)
R4WB_HD float u32_to_float_rz(uint32_t a)
{
#ifdef __CUDA_ARCH__
    return __uint2float_rz(a);
#else
    if (a >> 24) {
        int top = 31;
        while (!((a >> top) & 1u)) --top;
        const int sh = top - 23;
        a = (a >> sh) << sh;
    }
    return (float)a;
#endif
}
R4WB_HD float approx_lg2(float x)
{
#ifdef __CUDA_ARCH__
    float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
    return log2f(x);
#endif
}
R4WB_HD float approx_sqrt(float x)
{
#ifdef __CUDA_ARCH__
    float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
    return sqrtf(x);
#endif
}
R4WB_HD void approx_sincos(float x, float* s, float* c)
{
#ifdef __CUDA_ARCH__
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(*s) : "f"(x));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(*c) : "f"(x));
#else
    *s = sinf(x);
    *c = cosf(x);
#endif
}

// Box-Muller on two 32-bit words: u1 = (a + 256) 2^-32 in [2^-24, 1], angle = 2 pi (b 2^-32 - 1/2)
R4WB_HD float2 gauss_pair(uint32_t a, uint32_t b)
{
    const float u1 = fmaf(u32_to_float_rz(a), 2.3283064365386963e-10f, 5.9604644775390625e-08f);
    const float ang = fmaf(u32_to_float_rz(b), 1.4629180792671596e-09f /* 2 pi 2^-32 */, -3.14159265358979f);
    const float r = approx_sqrt(-1.3862943611198906f * approx_lg2(u1));
    float sn, cs;
    approx_sincos(ang, &sn, &cs);
    return make_float2(r * cs, r * sn);
}

// unit-variance noise pairs (re, im) of the two samples 2c and 2c+1 sharing Philox counter c
R4WB_HD void noise_of_counter(uint64_t ctr, const PhiloxKeys& K, float2& g0, float2& g1)
{
    uint32_t r0, r1, r2, r3;
    philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u, K, r0, r1, r2, r3);
    g0 = gauss_pair(r0, r1);
    g1 = gauss_pair(r2, r3);
}

// the (re, im) unit-variance noise pair of global sample m
R4WB_HD float2 noise_of_sample(uint64_t m, uint64_t seed)
{
    const PhiloxKeys K = philox_keys(seed);
    float2 g0, g1;
    noise_of_counter(m >> 1, K, g0, g1);
    return (m & 1ull) ? g1 : g0;
}

// test hook: a BlockSat entry as 12 doubles
inline void block_sat_debug(const BlockSat& e, double* o)
{
    o[0] = (double)(e.flags & 1u);
    o[1] = (double)e.U / 70368744177664.0;              // half-chips
    o[2] = (double)e.phi / 18446744073709551616.0;      // cycles
    o[3] = (double)e.f / 18446744073709551616.0;        // cycles / sample
    o[4] = (double)e.df / 18446744073709551616.0;
    o[5] = e.phase0;
    o[6] = (double)e.e0;
    o[7] = (double)e.amp;
    o[8] = (double)e.flags;
    o[9] = (double)e.eps46 / 70368744177664.0;
    o[10] = (double)e.G;
    o[11] = (double)e.n;
}

}  // namespace r4wb
