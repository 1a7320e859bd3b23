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
                            uint32_t n, uint64_t m_or_phi, BlockSat& o)
{
    const double fs = sc.fs;
    const double elapsed = (double)first / fs;
    const double t_start = sc.t0_gps + elapsed;
    const double t_end = t_start + (double)n / fs;
    const double elapsed_end = elapsed + (double)n / fs;

    double la_el = 0.0, la_range = 0.0, dop_s = 0.0, dop_e = 0.0;
    if (st.needs_orbit) {
        const RxState rx = rx_at(sc.rx, elapsed);
        Vec3 ps, vs, pe, ve;
        orbit_state(st.orbit, t_start, ps, vs);
        orbit_state(st.orbit, t_end, pe, ve);
        const Look la = look_from(rx.pos, rx.lla, ps);
        la_el = la.elevation_deg;
        la_range = la.range_m;
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

    const double iono_s = ((st.has & R4WB_HAS_IONO) ? st.iono_delay_m : 0.0) / kC;
    const double tropo_s = ((st.has & R4WB_HAS_TROPO) ? st.tropo_delay_m : 0.0) / kC;
    double cn0 = st.cn0_dbhz;
    if (!(st.has & R4WB_HAS_CN0))
        cn0 = st.tx_power_dbw - fspl_db(range_m, st.carrier_hz) + antenna_gain_dbi(sc.antenna, sc.ant_peak, sc.ant_bw, elevation) + 204.0;
    const double amp = pow(10.0, ((cn0 - 204.0) + 160.0) / 20.0);

    // code phase of the block (satellite_emitter.rs:228-242)
    const double total_delay_s = add_rn(add_rn(div_rn(range_m, kC), iono_s), tropo_s);
    const double chips_delay = mul_rn(total_delay_s, sc.chip_rate);
    const double phase0 = fmod(chips_delay, (double)kCodeLen);
    const double eq = div_rn(chips_delay, (double)kCodeLen);
    const uint64_t e0 = eq > 0.0 ? (uint64_t)eq : 0ull;

    // fixed-point half-chip position of the block's first oversample: exact rational part + f64 corrections
    const uint64_t G = first * (uint64_t)kOversample;
    const unsigned __int128 n1 = (unsigned __int128)G * sc.ratB;
    const uint64_t ci = (uint64_t)(n1 / sc.ratA);
    const uint64_t cr = (uint64_t)(n1 % sc.ratA);
    const uint64_t fracfx = (uint64_t)((((unsigned __int128)cr) << (kFracBits + 1)) / sc.ratA);
    const double chips_exact = (double)ci + (double)cr / (double)sc.ratA;
    const double corr_chips = -chips_exact * (sc.delta / (1.0 + sc.delta));
    const long long corrfx = llrint(corr_chips * 140737488355328.0 /* 2^47 */);
    const uint64_t hc_int = ((e0 % kSecLen) * (uint64_t)(2 * kCodeLen) + 2ull * (ci % (uint64_t)(kCodeLen * kSecLen))) % kHalfChipsPerSec;
    const uint64_t p0fx = (uint64_t)(phase0 * 140737488355328.0);
    __int128 Uw = ((__int128)hc_int << kFracBits) + (__int128)p0fx + (__int128)fracfx + (__int128)corrfx;
    while (Uw < 0) Uw += (__int128)kUMod;
    while (Uw >= (__int128)kUMod) Uw -= (__int128)kUMod;
    const uint64_t U = (uint64_t)Uw;

    // ambiguity band: the reference's cf = fl(phase0 + fl(g/spc)) is within ulp(cf) chips of the real value
    const double cf_max = phase0 + (double)(G + (uint64_t)kOversample * n) / sc.spc + 2.0;
    const double ulp = scalbn(1.0, ilogb(cf_max) - 52);
    const double eps_hc = 4.0 * ulp + 1.4901161193847656e-08 /* 2^-26 */;
    double e46 = ceil(eps_hc * 70368744177664.0 /* 2^46 */);
    if (e46 > 2147483648.0) e46 = 2147483648.0;
    double et = ceil(eps_hc * (sc.spc * 0.5) * 16777216.0 /* 2^24 */) + 2.0;
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
    o.amp = (float)amp; o.flags = flags; o.prev = -1; o.eps46 = (uint32_t)e46; o.eps_t = (uint32_t)et; o.pad = 0;
}

// phase advance of one block (cycles 0.64, wrapping): sum_{i<n} (f + i*df)
R4WB_HD uint64_t block_advance(const BlockSat& b)
{
    const uint64_t n = b.n;
    return n * (uint64_t)b.f + (uint64_t)b.df * (n * (n - 1) / 2);
}

// ----------------------------------------------------------------------------------------------
// per-tile, per-satellite state
struct TileSat {
    uint64_t u0;        // half-chip position at the newest oversample of the tile's first sample
    uint64_t phi;
    long long f, df;
    uint32_t hb;        // half-chip index of bit 0 of the sign table
    float amp;
    uint32_t flags;
    uint32_t eps_t;
};

R4WB_HD TileSat tile_sat(const BlockSat& b, uint32_t i_begin, uint64_t d8)
{
    TileSat t;
    t.u0 = b.U + (uint64_t)i_begin * d8;
    const uint32_t h0 = (uint32_t)(t.u0 >> kFracBits);
    t.hb = h0 - (uint32_t)(kJ + 2);            // may go "negative": all uses are modulo differences
    t.phi = b.phi; t.f = b.f; t.df = b.df; t.amp = b.amp; t.flags = b.flags; t.eps_t = b.eps_t;
    return t;
}

R4WB_HD uint32_t code_bit(const uint32_t* __restrict__ code, uint32_t c) { return (code[c >> 5] >> (c & 31)) & 1u; }

// word w of the per-tile half-chip sign table: bit n <-> half-chip hb + 32 w + n (1 <=> -1)
R4WB_HD uint32_t sign_word(const uint32_t* __restrict__ code, uint32_t hb, uint32_t w)
{
    int64_t hh = (int64_t)(int32_t)hb + 32 * (int64_t)w;
    hh %= (int64_t)kHalfChipsPerSec;
    if (hh < 0) hh += kHalfChipsPerSec;
    uint32_t h = (uint32_t)hh, chip = h >> 1, e = chip / kCodeLen, c = chip - e * kCodeLen, word = 0;
#pragma unroll 4
    for (int n = 0; n < 32; ++n) {
        word |= (code_bit(code, c) ^ (h & 1u) ^ ((kSecBits >> e) & 1u)) << n;
        if (h & 1u) { if (++c == (uint32_t)kCodeLen) { c = 0; if (++e == (uint32_t)kSecLen) e = 0; } }
        if (++h == kHalfChipsPerSec) { h = 0; }
    }
    return word;
}

// sign bit (1 <=> -1) of oversample q (>= 0, relative to the block start) of block entry bs;
// evaluates the reference expression literally when q is within the rounding band of a boundary.
R4WB_HD_NOINLINE uint32_t chip_sign_exact(const BlockSat& bs, long long q, const uint32_t* __restrict__ code, double spc)
{
    const double g = (double)(bs.G + (uint64_t)q);
    const double cf = add_rn(bs.phase0, div_rn(g, spc));                       // satellite_emitter.rs:268
    const double cm = fmod(cf, (double)kCodeLen);
    uint32_t c = cm > 0.0 ? (uint32_t)cm : 0u;                                   // :269
    if (c > (uint32_t)(kCodeLen - 1)) c = kCodeLen - 1;                          // :281
    const double cp = cf - floor(cf);                                            // :270
    const double eq = div_rn(cf, (double)kCodeLen);
    const uint64_t ep = (uint64_t)bs.e0 + (eq > 0.0 ? (uint64_t)eq : 0ull);      // :278
    const uint32_t boc = fmod(mul_rn(cp, 2.0), 2.0) < 1.0 ? 0u : 1u;             // :303-305
    return code_bit(code, c) ^ boc ^ ((kSecBits >> (uint32_t)(ep % kSecLen)) & 1u);
}

R4WB_HD uint32_t chip_sign(const BlockSat& bs, long long q, const uint32_t* __restrict__ code, uint64_t delta46, double spc)
{
    const uint64_t u = bs.U + (uint64_t)q * delta46;
    const uint64_t fr = u & kFracMask;
    if (fr < bs.eps46 || fr > kFracMask - bs.eps46) return chip_sign_exact(bs, q, code, spc);
    uint32_t h = (uint32_t)(u >> kFracBits);
    if (h >= kHalfChipsPerSec) h -= kHalfChipsPerSec;
    const uint32_t chip = h >> 1;
    const uint32_t e = chip / kCodeLen, c = chip - e * kCodeLen;
    return code_bit(code, c) ^ (h & 1u) ^ ((kSecBits >> e) & 1u);
}

// direct 63-tap evaluation of output sample i of block entry `cur` (history from `prev`): the reference's
// own loop (fir.rs:392-409 over satellite_emitter.rs:264-330), used for the first 8 samples of a block and
// for samples whose window touches an ambiguous boundary.
R4WB_HD_NOINLINE float fir_direct(const BlockSat& cur, const BlockSat* __restrict__ tab, const uint32_t* __restrict__ code,
                                  const float* __restrict__ taps, int i, uint64_t delta46, double spc)
{
    float acc = 0.0f;
    const long long g = (long long)kOversample * i;
    for (int k = 0; k < kTaps; ++k) {
        long long q = g - k;
        uint32_t sgn;
        if (q >= 0) {
            sgn = chip_sign(cur, q, code, delta46, spc);
        } else {
            if (cur.prev < 0) continue;                      // zero-initialised delay line
            const BlockSat& pb = tab[cur.prev];
            q += (long long)kOversample * pb.n;
            if (q < 0) continue;                             // history older than one block: not modelled
            sgn = chip_sign(pb, q, code, delta46, spc);
        }
        acc += sgn ? -taps[k] : taps[k];
    }
    return acc;
}

// collapsed FIR: four table look-ups + one 5-bit sign pattern.
//   u     half-chip position (18.46) of the newest oversample of the output sample
//   t64   [nw64] per-tile sign table as overlapping 64-bit windows (word k = bits 32k .. 32k+63)
//   erep  [63][32] E[d] replicated over the 32 banks, coef [32][8] (s_j - s_{j+1}) per sign pattern
R4WB_HD float fir_fast(uint64_t u, const TileSat& ts, const uint2* __restrict__ t64, const float* __restrict__ erep,
                       const float4* __restrict__ coef, uint32_t kmul, const uint32_t* cj, uint32_t lane, bool& ambiguous)
{
    const uint32_t h = (uint32_t)(u >> kFracBits);
    const uint32_t f32 = (uint32_t)(u >> (kFracBits - 32));
    const uint32_t t0 = umulhi32(f32, kmul);            // oversamples since the newest boundary, 2^-24 units
    const uint32_t t1 = t0 + cj[1], t2 = t0 + cj[2], t3 = t0 + cj[3];
    const uint32_t d0 = t0 >> kTBits, d1 = t1 >> kTBits, d2 = t2 >> kTBits;
    uint32_t d3 = t3 >> kTBits;
    if (d3 > 62u) d3 = 62u;
    if (ts.flags & 2u) {
        const uint32_t m = (1u << kTBits) - 1u, e = ts.eps_t;
        ambiguous = ((t0 + e) & m) < 2 * e || ((t1 + e) & m) < 2 * e || ((t2 + e) & m) < 2 * e || ((t3 + e) & m) < 2 * e;
    }
    const float e0 = erep[(d0 < 62u ? d0 : 62u) * 32 + lane], e1 = erep[(d1 < 62u ? d1 : 62u) * 32 + lane];
    const float e2 = erep[(d2 < 62u ? d2 : 62u) * 32 + lane], e3 = erep[d3 * 32 + lane];
    const uint32_t idx = h - ts.hb - (uint32_t)kJ;       // bit of half-chip h-J
    const uint2 w = t64[idx >> 5];
    const uint32_t pat = funnel_r(w.x, w.y, idx) & 31u;  // bit m <-> half-chip h-J+m
    const float4 c = coef[pat * 8 + (lane & 7)];
    const uint32_t so_bits = 0x3f800000u | (pat << 31);
#ifdef __CUDA_ARCH__
    const float s_old = __int_as_float((int)so_bits);
#else
    float s_old;
    { union { uint32_t u; float f; } cv; cv.u = so_bits; s_old = cv.f; }
#endif
    return fmaf(c.x, e0, fmaf(c.y, e1, fmaf(c.z, e2, fmaf(c.w, e3, s_old))));
}

// coefficient table entry for sign pattern `pat` (bit m <-> half-chip h-J+m  =>  s_j (age index j) is bit J-j)
R4WB_HD float4 coef_entry(uint32_t pat)
{
    float s[kJ + 1];
#pragma unroll
    for (int j = 0; j <= kJ; ++j) s[j] = ((pat >> (kJ - j)) & 1u) ? -1.0f : 1.0f;
    return make_float4(s[0] - s[1], s[1] - s[2], s[2] - s[3], s[3] - s[4]);
}

// carrier phase (cycles 0.64) of sample i of a block: phi + (i+1) f + i(i+1)/2 df   (scenario.rs:519-524)
R4WB_HD uint64_t carrier_phase(const TileSat& ts, uint32_t i)
{
    return ts.phi + (uint64_t)(i + 1) * (uint64_t)ts.f + ((uint64_t)i * (i + 1) / 2) * (uint64_t)ts.df;
}

R4WB_HD void rotate_acc(float y, uint64_t ph, float& re, float& im)
{
    float s, c;
    fast_sincos((float)(int32_t)(ph >> 32) * 1.4629180792671596e-09f /* 2 pi / 2^32 */, &s, &c);
    re = fmaf(y, c, re);
    im = fmaf(y, s, im);
}

// ----------------------------------------------------------------------------------------------
// noise: Philox4x32-10 keyed by the scenario seed, counter = global sample index / 2 (one draw -> two samples)
R4WB_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                           uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = umulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = umulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    r0 = c0; r1 = c1; r2 = c2; r3 = c3;
}

// Box-Muller on two 32-bit words: u1 = ((a>>8)+1) 2^-24 in (0,1], angle = 2 pi ((b>>8) 2^-24 - 1/2)
R4WB_HD float2 gauss_pair(uint32_t a, uint32_t b)
{
    const float u1 = (float)((a >> 8) + 1u) * 5.9604644775390625e-08f;
    const float ang = ((float)(b >> 8) * 5.9604644775390625e-08f - 0.5f) * 6.283185307179586f;
    const float r = sqrtf(-1.3862943611198906f * fast_log2(u1));
    float sn, cs;
    fast_sincos(ang, &sn, &cs);
    return make_float2(r * cs, r * sn);
}

// the (re, im) unit-variance noise pair of global sample m
R4WB_HD float2 noise_of_sample(uint64_t m, uint64_t seed)
{
    uint32_t r0, r1, r2, r3;
    const uint64_t ctr = m >> 1;
    philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r0, r1, r2, r3);
    return (m & 1ull) ? gauss_pair(r2, r3) : gauss_pair(r0, r1);
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
