// synth_kernels.cu — sm_100a kernels of the scenario-synthesis path.
//
// Replaces GnssScenario::generate_block (gnss/scenario.rs:308-546) and what it calls per sample:
//   SatelliteEmitter::generate_baseband_iq   gnss/satellite_emitter.rs:218-347  (E1C branch :325-330)
//   FirFilter::process + step_by(8)          core/filters/fir.rs:392-409, gnss/scenario.rs:486-489
//   Doppler rotation + amplitude + sum       gnss/scenario.rs:516-528
//   thermal noise                            gnss/scenario.rs:530-542   (Philox4x32-10 instead of xorshift64)
//   cf32 sink cast                           core/io/format.rs:197-200
//
// Algebra: the 8x-oversampled baseband is +-1 and constant over half-chips, so the 63-tap
// FIR at a decimated output is  y = s_old + sum_j (s_j - s_{j+1}) * E[d_j]  with E the running sum of the
// taps and d_j the age (in oversamples) of the j-th half-chip boundary inside the window.  Boundaries
// come from a 64-bit fixed-point code NCO; whenever a sample lies within the rounding wobble of the
// reference's own f64 expression the kernel evaluates that expression literally (exact path).
#include <cuda_runtime.h>

#include "synth.cuh"

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

constexpr uint64_t kFracMask = (1ull << kFracBits) - 1ull;
constexpr uint64_t kUMod = (uint64_t)kHalfChipsPerSec << kFracBits;   // 204600 * 2^46 < 2^64
constexpr double kTwo64 = 18446744073709551616.0;

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

// host-visible wrapper used by the sequential (explicit block) API and by tests
void host_fill_block_sat(const ScenConst& sc, const SatConst& st, const PhaseSegment* segs, uint64_t first,
                         uint32_t n, uint64_t m_or_phi, BlockSat& o)
{
    fill_block_sat(sc, st, segs, first, n, m_or_phi, o);
}

// phase advance of one block (cycles 0.64, wrapping): sum_{i<n} (f + i*df)
R4WB_HD uint64_t block_advance(const BlockSat& b)
{
    const uint64_t n = b.n;
    return n * (uint64_t)b.f + (uint64_t)b.df * (n * (n - 1) / 2);
}
uint64_t host_block_advance(const BlockSat& b) { return block_advance(b); }

// ----------------------------------------------------------------------------------------------
// prologue kernels
__global__ void k_block_params(ScenConst sc, const SatConst* __restrict__ sats, const PhaseSegment* __restrict__ segs,
                               uint64_t blk0, uint32_t nblk, BlockSat* __restrict__ tab, BlockHdr* __restrict__ hdr)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)nblk * sc.n_sats) return;
    const uint32_t tb = (uint32_t)(idx / sc.n_sats), s = (uint32_t)(idx % sc.n_sats);
    const uint64_t first = (blk0 + tb) * sc.B;
    const uint64_t rem = sc.total - first;
    const uint32_t n = (uint32_t)(rem < sc.B ? rem : sc.B);
    BlockSat o;
    fill_block_sat(sc, sats[s], segs, first, n, /*visible samples so far (constant visibility)*/ first, o);
    o.prev = tb > 0 ? (int32_t)((tb - 1) * sc.n_sats + s) : -1;
    if (!sats[s].static_phase) o.phi = (o.flags & 1u) ? block_advance(o) : 0ull;   // scanned by k_phase_scan
    tab[idx] = o;
    if (s == 0) hdr[tb] = BlockHdr{first, n, 0};
}

// One CTA per satellite: exclusive scan of the per-block phase advance (dynamic satellites) and of the
// "last visible block" pointer (satellites whose visibility can change).  Table must start at block 0.
__global__ void __launch_bounds__(1024) k_phase_scan(const SatConst* __restrict__ sats, uint32_t n_sats, uint32_t nblk,
                                                      BlockSat* __restrict__ tab)
{
    __shared__ uint64_t s_sum[1024];
    __shared__ int s_last[1024];
    const uint32_t s = blockIdx.x, t = threadIdx.x;
    const bool dynamic = !sats[s].static_phase;
    const uint32_t chunk = (nblk + 1023u) / 1024u;
    const uint32_t lo = t * chunk, hi = min(nblk, lo + chunk);
    uint64_t sum = 0;
    int last = -1;
    for (uint32_t b = lo; b < hi; ++b) {
        const BlockSat& e = tab[(size_t)b * n_sats + s];
        if (dynamic) sum += e.phi;
        if (e.flags & 1u) last = (int)b;
    }
    s_sum[t] = sum;
    s_last[t] = last;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        uint64_t v = 0;
        int l = -1;
        if ((int)t >= off) { v = s_sum[t - off]; l = s_last[t - off]; }
        __syncthreads();
        if ((int)t >= off) { s_sum[t] += v; s_last[t] = max(s_last[t], l); }
        __syncthreads();
    }
    uint64_t run = t > 0 ? s_sum[t - 1] : 0ull;
    int prev = t > 0 ? s_last[t - 1] : -1;
    for (uint32_t b = lo; b < hi; ++b) {
        BlockSat& e = tab[(size_t)b * n_sats + s];
        const uint64_t adv = e.phi;
        if (dynamic) { e.phi = run; run += adv; }
        e.prev = prev >= 0 ? (int32_t)((uint32_t)prev * n_sats + s) : -1;
        if (e.flags & 1u) prev = (int)b;
    }
}

// ----------------------------------------------------------------------------------------------
// synthesis kernel
constexpr int kThreads = 256;

struct TileSat {
    uint64_t u0;        // half-chip position at the newest oversample of the tile's first sample
    uint64_t phi;
    long long f, df;
    uint32_t hb;        // half-chip index of bit 0 of the sign table
    float amp;
    uint32_t flags;
    uint32_t eps_t;
};

__device__ __forceinline__ uint32_t code_bit(const uint32_t* __restrict__ code, uint32_t c) { return (code[c >> 5] >> (c & 31)) & 1u; }

// sign bit (1 <=> -1) of oversample q (>= 0, relative to the block start) of block entry bs;
// evaluates the reference expression literally when q is within the rounding band of a boundary.
__device__ __noinline__ uint32_t chip_sign_exact(const BlockSat& bs, long long q, const uint32_t* __restrict__ code, double spc)
{
    const double g = (double)(bs.G + (uint64_t)q);
    const double cf = __dadd_rn(bs.phase0, __ddiv_rn(g, spc));                 // satellite_emitter.rs:268
    const double cm = fmod(cf, (double)kCodeLen);
    uint32_t c = cm > 0.0 ? (uint32_t)cm : 0u;                                   // :269
    if (c > (uint32_t)(kCodeLen - 1)) c = kCodeLen - 1;                          // :281
    const double cp = cf - floor(cf);                                            // :270
    const double eq = __ddiv_rn(cf, (double)kCodeLen);
    const uint64_t ep = (uint64_t)bs.e0 + (eq > 0.0 ? (uint64_t)eq : 0ull);      // :278
    const uint32_t boc = fmod(__dmul_rn(cp, 2.0), 2.0) < 1.0 ? 0u : 1u;          // :303-305
    return code_bit(code, c) ^ boc ^ ((kSecBits >> (uint32_t)(ep % kSecLen)) & 1u);
}

__device__ __forceinline__ uint32_t chip_sign(const BlockSat& bs, long long q, const uint32_t* __restrict__ code,
                                              uint64_t delta46, double spc)
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
__device__ __noinline__ float fir_direct(const BlockSat& cur, const BlockSat* __restrict__ tab, const uint32_t* __restrict__ code,
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

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                               uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    r0 = c0; r1 = c1; r2 = c2; r3 = c3;
}

// Box-Muller on two 32-bit words: u1 = ((a>>8)+1) 2^-24 in (0,1], angle = 2 pi ((b>>8) 2^-24 - 1/2)
__device__ __forceinline__ float2 gauss_pair(uint32_t a, uint32_t b)
{
    const float u1 = (float)((a >> 8) + 1u) * 5.9604644775390625e-08f;
    const float ang = ((float)(b >> 8) * 5.9604644775390625e-08f - 0.5f) * 6.283185307179586f;
    const float r = sqrtf(-1.3862943611198906f * __log2f(u1));
    float sn, cs;
    __sincosf(ang, &sn, &cs);
    return make_float2(r * cs, r * sn);
}

template <int J>
__device__ __forceinline__ float fir_fast(uint64_t u, const TileSat& ts, const uint2* __restrict__ t64, const float* __restrict__ erep,
                                          const float4* __restrict__ coef, uint32_t kmul, const uint32_t* cj, uint32_t lane,
                                          bool& ambiguous)
{
    static_assert(J == 4, "coefficient table is built for four boundaries");
    const uint32_t h = (uint32_t)(u >> kFracBits);
    const uint32_t f32 = (uint32_t)(u >> (kFracBits - 32));
    const uint32_t t0 = __umulhi(f32, kmul);            // oversamples since the newest boundary, 2^-24 units
    const uint32_t t1 = t0 + cj[1], t2 = t0 + cj[2], t3 = t0 + cj[3];
    const uint32_t d0 = t0 >> kTBits, d1 = t1 >> kTBits, d2 = t2 >> kTBits, d3 = min(t3 >> kTBits, 62u);
    if (ts.flags & 2u) {
        const uint32_t m = (1u << kTBits) - 1u, e = ts.eps_t;
        ambiguous = ((t0 + e) & m) < 2 * e || ((t1 + e) & m) < 2 * e || ((t2 + e) & m) < 2 * e || ((t3 + e) & m) < 2 * e;
    }
    const float e0 = erep[min(d0, 62u) * 32 + lane], e1 = erep[min(d1, 62u) * 32 + lane];
    const float e2 = erep[min(d2, 62u) * 32 + lane], e3 = erep[d3 * 32 + lane];
    const uint32_t idx = h - ts.hb - J;                  // bit of half-chip h-J
    const uint2 w = t64[idx >> 5];
    const uint32_t pat = __funnelshift_r(w.x, w.y, idx) & 31u;   // bit m <-> half-chip h-J+m
    const float4 c = coef[pat * 8 + (lane & 7)];
    const float s_old = __int_as_float(0x3f800000u | (pat << 31));
    return fmaf(c.x, e0, fmaf(c.y, e1, fmaf(c.z, e2, fmaf(c.w, e3, s_old))));
}

template <int K, bool CF64>
__global__ void __launch_bounds__(kThreads) k_synth(SynthArgs a)
{
    constexpr int J = 4;
    constexpr int TILE = kThreads * 2 * K;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* erep = reinterpret_cast<float*>(smem_raw);                          // [63][32]
    float4* coef = reinterpret_cast<float4*>(erep + 63 * 32);                  // [32][8]
    float* taps = reinterpret_cast<float*>(coef + 32 * 8);                     // [64]
    uint32_t* codes = reinterpret_cast<uint32_t*>(taps + 64);                  // [n_sats][128]
    TileSat* tsat = reinterpret_cast<TileSat*>(codes + a.n_sats * 128);        // [n_sats]
    uint2* t64 = reinterpret_cast<uint2*>(tsat + a.n_sats);                    // [n_sats][nw64]
    uint32_t* w32 = reinterpret_cast<uint32_t*>(t64 + a.n_sats * a.nw64);      // [n_sats][nw64+1]
    float* yfix = reinterpret_cast<float*>(w32 + a.n_sats * (a.nw64 + 1));     // [n_sats][8]
    __shared__ float s_pow[kThreads / 32];

    const uint32_t tid = threadIdx.x, lane = tid & 31u;

    // kernel-lifetime tables
    for (uint32_t k = tid; k < 63 * 32; k += kThreads) erep[k] = a.etab[k >> 5];
    for (uint32_t k = tid; k < 32 * 8; k += kThreads) {
        const uint32_t pat = k >> 3;     // bit m <-> half-chip h-J+m  => s_j (j = age index) is bit J-j
        float s[J + 1];
#pragma unroll
        for (int j = 0; j <= J; ++j) s[j] = ((pat >> (J - j)) & 1u) ? -1.0f : 1.0f;
        coef[k] = make_float4(s[0] - s[1], s[1] - s[2], s[2] - s[3], s[3] - s[4]);
    }
    for (uint32_t k = tid; k < 64; k += kThreads) taps[k] = a.taps[k];
    for (uint32_t k = tid; k < a.n_sats * 128; k += kThreads) codes[k] = a.codebits[k];

    float pow_acc = 0.0f;
    const uint32_t n_tiles = a.tb_count * a.tiles_per_block;
    const uint64_t d8 = a.delta46 * (uint64_t)kOversample;

    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t tb = a.tb_begin + tile / a.tiles_per_block;
        const uint32_t chunk = tile % a.tiles_per_block;
        const BlockHdr hd = a.hdr[tb];
        const uint32_t i_begin = chunk * TILE;
        if (i_begin >= hd.n) continue;
        const uint32_t i_end = min(hd.n, i_begin + TILE);
        // skip tiles entirely outside the requested output range
        if (hd.first + i_end <= a.out_first || hd.first + i_begin >= a.out_first + a.out_n) continue;
        const BlockSat* row = a.tab + (size_t)tb * a.n_sats;

        __syncthreads();   // previous tile's readers are done
        if (tid < a.n_sats) {
            const BlockSat& b = row[tid];
            TileSat t;
            t.u0 = b.U + (uint64_t)i_begin * d8;
            uint32_t h0 = (uint32_t)(t.u0 >> kFracBits);
            t.hb = h0 - (J + 2);                    // may go "negative": all uses are modulo differences
            t.phi = b.phi; t.f = b.f; t.df = b.df; t.amp = b.amp; t.flags = b.flags; t.eps_t = b.eps_t;
            tsat[tid] = t;
        }
        __syncthreads();
        // half-chip sign words: bit n of word w <-> half-chip hb + 32 w + n
        for (uint32_t k = tid; k < a.n_sats * (a.nw64 + 1); k += kThreads) {
            const uint32_t s = k / (a.nw64 + 1), w = k - s * (a.nw64 + 1);
            const uint32_t* code = codes + s * 128;
            int64_t hh = (int64_t)(int32_t)tsat[s].hb + 32 * (int64_t)w;
            hh %= (int64_t)kHalfChipsPerSec;
            if (hh < 0) hh += kHalfChipsPerSec;
            uint32_t h = (uint32_t)hh, chip = h >> 1, e = chip / kCodeLen, c = chip - e * kCodeLen, word = 0;
#pragma unroll 4
            for (int n = 0; n < 32; ++n) {
                word |= (code_bit(code, c) ^ (h & 1u) ^ ((kSecBits >> e) & 1u)) << n;
                if (h & 1u) { if (++c == kCodeLen) { c = 0; if (++e == kSecLen) e = 0; } }
                if (++h == kHalfChipsPerSec) h = 0;
            }
            w32[k] = word;
        }
        __syncthreads();
        for (uint32_t k = tid; k < a.n_sats * a.nw64; k += kThreads) {
            const uint32_t s = k / a.nw64, w = k - s * a.nw64;
            t64[k] = make_uint2(w32[s * (a.nw64 + 1) + w], w32[s * (a.nw64 + 1) + w + 1]);
        }
        // first 8 samples of a block: their window reaches into the previous block
        if (chunk == 0) {
            for (uint32_t k = tid; k < a.n_sats * 8; k += kThreads) {
                const uint32_t s = k >> 3, i = k & 7u;
                float y = 0.0f;
                if ((row[s].flags & 1u) && i < hd.n) y = fir_direct(row[s], a.tab, codes + s * 128, taps, (int)i, a.delta46, a.spc);
                yfix[k] = y;
            }
        }
        __syncthreads();

        float2 acc[K][2];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k][0] = acc[k][1] = make_float2(0.0f, 0.0f);

        for (uint32_t s = 0; s < a.n_sats; ++s) {
            const TileSat ts = tsat[s];
            if (!(ts.flags & 1u)) continue;
            const uint2* tw = t64 + s * a.nw64;
            const uint32_t ia0 = i_begin + 2 * tid;                          // sample index within the block
            uint64_t ua = ts.u0 + (uint64_t)(2 * tid) * d8;
            // phase of sample i: phi + (i+1) f + i(i+1)/2 df
            uint64_t pha = ts.phi + (uint64_t)(ia0 + 1) * (uint64_t)ts.f + ((uint64_t)ia0 * (ia0 + 1) / 2) * (uint64_t)ts.df;
            uint64_t gb = (uint64_t)ts.f + (uint64_t)ts.df * (uint64_t)(ia0 + 1);       // increment applied at sample i+1
            uint64_t step = (uint64_t)(2 * kThreads) * (uint64_t)ts.f +
                            (uint64_t)ts.df * ((uint64_t)(2 * kThreads) * ia0 + (uint64_t)(2 * kThreads) * (2 * kThreads + 1) / 2);
            const uint64_t step2 = (uint64_t)ts.df * (uint64_t)(2 * kThreads) * (uint64_t)(2 * kThreads);
            const uint64_t g512 = (uint64_t)ts.df * (uint64_t)(2 * kThreads);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const uint32_t ia = ia0 + 2 * kThreads * k;
                bool amb_a = false, amb_b = false;
                float ya = fir_fast<J>(ua, ts, tw, erep, coef, a.kmul, a.cj, lane, amb_a);
                float yb = fir_fast<J>(ua + d8, ts, tw, erep, coef, a.kmul, a.cj, lane, amb_b);
                if (amb_a && ia < i_end) ya = fir_direct(row[s], a.tab, codes + s * 128, taps, (int)ia, a.delta46, a.spc);
                if (amb_b && ia + 1 < i_end) yb = fir_direct(row[s], a.tab, codes + s * 128, taps, (int)ia + 1, a.delta46, a.spc);
                if (k == 0 && chunk == 0 && tid < 4) { ya = yfix[s * 8 + 2 * tid]; yb = yfix[s * 8 + 2 * tid + 1]; }
                const uint64_t phb = pha + gb;
                float sa, ca, sb, cb;
                __sincosf((float)(int32_t)(pha >> 32) * 1.4629180792671596e-09f /* 2 pi / 2^32 */, &sa, &ca);
                __sincosf((float)(int32_t)(phb >> 32) * 1.4629180792671596e-09f, &sb, &cb);
                ya *= ts.amp; yb *= ts.amp;
                acc[k][0].x = fmaf(ya, ca, acc[k][0].x); acc[k][0].y = fmaf(ya, sa, acc[k][0].y);
                acc[k][1].x = fmaf(yb, cb, acc[k][1].x); acc[k][1].y = fmaf(yb, sb, acc[k][1].y);
                ua += d8 * (uint64_t)(2 * kThreads);
                pha += step; step += step2; gb += g512;
            }
        }

        // noise, power, store
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const uint32_t ia = i_begin + 2 * tid + 2 * kThreads * k;
            if (ia >= i_end) continue;
            const uint64_t m = hd.first + ia;                    // global sample index
            const bool has_b = ia + 1 < i_end;
            float2 va = acc[k][0], vb = acc[k][1];
            if (!(a.flags & R4WB_FLAG_NOISE_OFF)) {
                uint32_t r0, r1, r2, r3;
                const uint64_t ctr = m >> 1;
                philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u, (uint32_t)a.seed, (uint32_t)(a.seed >> 32), r0, r1, r2, r3);
                float2 ga, gb2;
                if ((m & 1ull) == 0) { ga = gauss_pair(r0, r1); gb2 = gauss_pair(r2, r3); }
                else {
                    ga = gauss_pair(r2, r3);
                    const uint64_t c2 = ctr + 1;
                    philox4x32_10((uint32_t)c2, (uint32_t)(c2 >> 32), 0u, 0u, (uint32_t)a.seed, (uint32_t)(a.seed >> 32), r0, r1, r2, r3);
                    gb2 = gauss_pair(r0, r1);
                }
                va.x = fmaf(ga.x, a.noise_std, va.x); va.y = fmaf(ga.y, a.noise_std, va.y);
                vb.x = fmaf(gb2.x, a.noise_std, vb.x); vb.y = fmaf(gb2.y, a.noise_std, vb.y);
            }
            const bool wa = m >= a.out_first && m < a.out_first + a.out_n;
            const bool wb = has_b && (m + 1) >= a.out_first && (m + 1) < a.out_first + a.out_n;
            if (wa) pow_acc += va.x * va.x + va.y * va.y;
            if (wb) pow_acc += vb.x * vb.x + vb.y * vb.y;
            const uint64_t o = m - a.out_first;                  // only meaningful when wa (wraps otherwise)
            if (CF64) {
                double2* out = reinterpret_cast<double2*>(a.out);
                if (wa) out[o] = make_double2((double)va.x, (double)va.y);
                if (wb) out[o + 1] = make_double2((double)vb.x, (double)vb.y);
            } else {
                float2* out = reinterpret_cast<float2*>(a.out);
                if (wa && wb && ((o & 1ull) == 0)) {
                    *reinterpret_cast<float4*>(out + o) = make_float4(va.x, va.y, vb.x, vb.y);
                } else {
                    if (wa) out[o] = va;
                    if (wb) out[o + 1] = vb;
                }
            }
        }
    }

    if (a.power_sum) {
        for (int off = 16; off > 0; off >>= 1) pow_acc += __shfl_xor_sync(0xffffffffu, pow_acc, off);
        if (lane == 0) s_pow[tid >> 5] = pow_acc;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < kThreads / 32; ++w) t += (double)s_pow[w];
            atomicAdd(a.power_sum, t);
        }
    }
}

// ----------------------------------------------------------------------------------------------
// launchers (called from synth_host.cu)
size_t synth_smem_bytes(uint32_t n_sats, uint32_t nw64)
{
    size_t b = 63 * 32 * 4 + 32 * 8 * 16 + 64 * 4;
    b += (size_t)n_sats * 128 * 4;
    b += (size_t)n_sats * sizeof(TileSat);
    b += (size_t)n_sats * nw64 * 8;
    b += (size_t)n_sats * (nw64 + 1) * 4;
    b += (size_t)n_sats * 8 * 4;
    return (b + 15) & ~(size_t)15;
}

int synth_tile_samples(int K) { return kThreads * 2 * K; }

template <int K, bool CF64>
static void launch_synth_t(const SynthArgs& a, int grid, size_t smem, cudaStream_t st)
{
    static bool attr_done = false;
    if (!attr_done) {
        R4WB_CUDA(cudaFuncSetAttribute(k_synth<K, CF64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    k_synth<K, CF64><<<grid, kThreads, smem, st>>>(a);
    R4WB_LAUNCH_CHECK();
}

void launch_synth_kernel(const SynthArgs& a, int K, bool cf64, int grid, cudaStream_t st)
{
    const size_t smem = synth_smem_bytes(a.n_sats, a.nw64);
    if (K == 5) { cf64 ? launch_synth_t<5, true>(a, grid, smem, st) : launch_synth_t<5, false>(a, grid, smem, st); }
    else if (K == 10) { cf64 ? launch_synth_t<10, true>(a, grid, smem, st) : launch_synth_t<10, false>(a, grid, smem, st); }
    else fail(R4WB_ERR_INVALID_PARAMETER, "unsupported tile factor %d", K);
}

int synth_max_blocks_per_sm(int K, bool cf64, size_t smem)
{
    int nb = 0;
    if (K == 5) {
        if (cf64) R4WB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_synth<5, true>, kThreads, smem));
        else R4WB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_synth<5, false>, kThreads, smem));
    } else {
        if (cf64) R4WB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_synth<10, true>, kThreads, smem));
        else R4WB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_synth<10, false>, kThreads, smem));
    }
    return nb;
}

void launch_block_params(const ScenConst& sc, const SatConst* d_sats, const PhaseSegment* d_segs, uint64_t blk0, uint32_t nblk,
                         BlockSat* d_tab, BlockHdr* d_hdr, cudaStream_t st)
{
    const uint64_t total = (uint64_t)nblk * sc.n_sats;
    if (total == 0) return;
    const int threads = 128;
    k_block_params<<<(unsigned)((total + threads - 1) / threads), threads, 0, st>>>(sc, d_sats, d_segs, blk0, nblk, d_tab, d_hdr);
    R4WB_LAUNCH_CHECK();
}

void launch_phase_scan(const SatConst* d_sats, uint32_t n_sats, uint32_t nblk, BlockSat* d_tab, cudaStream_t st)
{
    if (nblk == 0 || n_sats == 0) return;
    k_phase_scan<<<n_sats, 1024, 0, st>>>(d_sats, n_sats, nblk, d_tab);
    R4WB_LAUNCH_CHECK();
}

}  // namespace r4wb
