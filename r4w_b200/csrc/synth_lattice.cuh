// synth_lattice.cuh — per-thread arithmetic of the lattice synthesis kernel (k_synth_lat, synth_lattice.cu).
//
// Same path as synth_math.cuh (GnssScenario::generate_block, gnss/scenario.rs:308-546; emitter satellite_emitter.rs:218-347;
// FIR + decimation fir.rs:392-409, scenario.rs:486-489; Doppler rotation scenario.rs:516-528), same collapsed-FIR algebra
// (y = ytab[boundary-age class][5-sign pattern]), organised around the SAMPLE LATTICE of the reference:
//
//   The reference has no code Doppler: inside one 1 ms block the half-chip position of output sample m is
//   U_b + m p / q (p / q = 1023 / 2500 half-chips per sample at 5 MHz, exactly), so
//   * samples m and m + q of a block have the same half-chip fraction — the same boundary-age class — and sit exactly
//     p half-chips apart: a thread owns the quad {m, m + 1, m + q, m + q + 1} and looks the two classes up once;
//   * the class of sample m is clsn[b0][m + moff]: CONSECUTIVE samples read CONSECUTIVE bytes of a table chosen by the
//     block's sub-residue b0 (8 tables), so the class look-up is a conflict-free shared-memory byte load with an
//     immediate offset per step — no per-sample multiply, swizzle or fraction arithmetic;
//   * the four 5-sign patterns of a quad come out of ONE 16-byte shared-memory load: two overlapping 64-bit windows of the
//     block's half-chip sign sequence, the second pre-shifted by p bits, shared by both halves' shift amounts.
//   A block whose lattice comes within the f64 rounding band of a half-chip boundary (TileSat.flags bit 1) is rendered by
//   the same straight-line code; the (at most two) oversamples of the block inside the band are resolved once per block
//   by the reference's literal f64 expression in k_tile_params and, if they fall on the other side, patched into the
//   (at most 16) windows that hold them: +-2 h[g - q*].
//
// Everything is __host__ __device__: tests/emu replays it on the CPU against the oracle.
#pragma once
#include <cstdlib>

#include "synth_math.cuh"

namespace r4wb {

constexpr int kLatMaxSats = 16;
constexpr int kLatMaxK = 5;

// entries of a satellite's window table / sign words behind them
R4WB_HD uint32_t lat_n_ent(const LatConst& L)
{
    const uint64_t span = ((uint64_t)(2 * kSynthThreads) * L.K * L.p + L.q - 1) / L.q + 2;   // half-chips a first-half quad can reach
    return (uint32_t)(span >> 5) + 2u;
}
R4WB_HD uint32_t lat_n_words(const LatConst& L) { return lat_n_ent(L) + (L.p >> 5) + 2u; }

// entry j: bits [32 j, 32 j + 64) of the sign sequence and bits [32 j + p, 32 j + p + 64) (the samples q later)
R4WB_HD uint4 lat_entry(const uint32_t* __restrict__ W, uint32_t j, uint32_t p)
{
    const uint32_t o = j + (p >> 5), s = p & 31u;
    return make_uint4(W[j], W[j + 1], funnel_r(W[o], W[o + 1], s), funnel_r(W[o + 1], W[o + 2], s));
}

// Lattice part of a tile record (k_tile_params).  `row` is the block's entry, tile = the whole block (i_begin = 0).
R4WB_HD TileLat tile_lat(const BlockSat& b, const LatConst& L, const uint32_t* __restrict__ per, double spc, const SatCode& cd)
{
    TileLat t;
    const uint32_t D = 8u * L.q;
    // rotation over one sample and over q samples at the block start: phase(i) = phi + (i+1) f + i(i+1)/2 df
    {
        const double c64 = 5.421010862427522e-20;     // 2^-64
        const double x1 = (double)(int64_t)((uint64_t)b.f + (uint64_t)b.df) * c64;             // phase(1) - phase(0), cycles
        const uint64_t q = L.q;
        const double xq = (double)(int64_t)(q * (uint64_t)b.f + (uint64_t)b.df * (q * (q + 1) / 2)) * c64;
        t.r1r = (float)cos(6.283185307179586 * x1); t.r1i = (float)sin(6.283185307179586 * x1);
        t.rqr = (float)cos(6.283185307179586 * xq); t.rqi = (float)sin(6.283185307179586 * xq);
    }
    // bin of the first sample's half-chip fraction, centred fraction, class-table offset
    const uint64_t fr = b.U & kFracMask;
    const uint32_t bin = (uint32_t)(((unsigned __int128)fr * D) >> kFracBits);           // floor(frac D) < D
    const uint32_t b1 = bin >> 3;
    const uint32_t moff = (uint32_t)(((uint64_t)b1 * L.pinv) % L.q);
    t.latb = bin | (moff << 16);
    t.fc32 = (uint32_t)((((uint64_t)(2u * bin + 1u)) << 32) / (2ull * D));                // (bin + 1/2) / D, 0.32 fixed (floor)
    t.patch = 0u;
    {   // origin of the sign table (half-chip H - 6, the hb of tile_sat) as (epoch, position inside the period)
        const uint32_t H0 = (uint32_t)(b.U >> kFracBits);
        const uint32_t hbm = (H0 + cd.hc_mod - (uint32_t)(kJ + 2)) % cd.hc_mod;
        const uint32_t e0 = hbm / cd.per_len;
        t.ep0 = e0 | ((hbm - e0 * cd.per_len) << 8);
    }
    if ((b.flags & 3u) == 3u) {
        // the two oversamples of the block whose lattice point is the one next to a half-chip boundary
        const uint64_t rD = ((unsigned __int128)fr * D) & kFracMask;                       // frac(frac D), 0.46 fixed
        const uint32_t target = rD < (1ull << (kFracBits - 1)) ? 0u : D - 1u;              // lattice just above / just below the boundary
        const uint32_t q1 = (uint32_t)(((uint64_t)((target + D - bin) % D) * L.povinv) % D);
        const uint32_t H = (uint32_t)(b.U >> kFracBits);
        for (int e = 0; e < 2; ++e) {
            const uint32_t qs = q1 + (uint32_t)e * D;
            const uint32_t s_exact = chip_sign_exact(b, (long long)qs, per, spc, cd);
            const uint64_t adv = ((uint64_t)bin + (uint64_t)L.pov * qs) / D;               // whole half-chips the lattice model has advanced
            const uint32_t hm = (uint32_t)((H + adv) % cd.hc_mod);
            const uint32_t s_model = halfchip_sign(per, hm, cd);
            uint32_t code = s_exact == s_model ? 0u : (s_exact == 0u ? 1u : 2u);            // exact +1, model -1: add +2 h
#ifdef R4WB_EMU_BUILD
            if (code && std::getenv("R4WB_EMU_LAT_NO_PATCH")) code = 3u;                    // host replay only: proves the patches are live (tests)
#endif
            t.patch |= code << (16 + 2 * e);
        }
        t.patch |= q1;
    }
    return t;
}

// True when the lattice kernel's carrier model holds for this record: step phasor linearised in the Doppler slope, the
// second-order growth over the K steps of a thread folded into a constant (error <= 2 th2 rad, see lat_sat_accumulate)
R4WB_HD bool lat_rotation_ok(const TileSat& ts) { return !(ts.flags & 4u) && fabsf(ts.th2) * 2.0f <= 5e-7f; }

// One satellite's contribution to the K quads thread `tid` owns in a block: quad k = samples
// (2 tid + 2 kSynthThreads k, +1) and the same + q.  arA/aiA accumulate (re, re) / (im, im) of the first-half pair,
// arB/aiB of the second-half pair.
//   ent    [lat_n_ent] window table of this (block, satellite)
//   ytab2  [classes][kLatYStride]: row = class, column = 5-sign pattern (bit 0 = oldest half-chip)
//   clsn   [8][L.cls_len] class of sample index n per sub-residue
template <int K>
R4WB_HD void lat_sat_accumulate(const TileRec& rec, const LatConst& L, uint64_t d8_46, const uint4* __restrict__ ent,
                                const float* __restrict__ ytab2, const uint8_t* __restrict__ clsn, const float* __restrict__ taps,
                                uint32_t tid, float2 (&arA)[K], float2 (&aiA)[K], float2 (&arB)[K], float2 (&aiB)[K])
{
    const TileSat& ts = rec.ts;
    const uint32_t bin = rec.lat.latb & 0xffffu, moff = rec.lat.latb >> 16;
    const uint8_t* __restrict__ cls = clsn + (size_t)(bin & 7u) * L.cls_len + moff + 2u * tid;
    // position of sample 2 tid relative to the block's first half-chip, 12.20 fixed point (the fraction is bin-centred: the
    // accumulated rounding of the K steps stays two orders below the half bin that separates it from a wrong floor)
    uint32_t pos = (uint32_t)((((uint64_t)rec.lat.fc32 << (kFracBits - 32)) + (uint64_t)(2u * tid) * d8_46) >> (kFracBits - 20));

    // carrier: exact 64-bit phase of sample 2 tid, the other three samples of the quad by rotation (f64-accurate phasors of the
    // tile record, linearised in the Doppler slope: the neglected terms are below 1e-10 rad)
    const uint32_t i0 = 2u * tid;
    const float eps = ts.th1 * (1.0f / (float)(2 * kSynthThreads));           // 2 pi df: growth of the per-sample advance per sample
    float s0, c0;
    phasor(carrier_phase(ts, i0), &s0, &c0);
    c0 *= ts.amp; s0 *= ts.amp;
    // sample i0 + 1: advance = 2 pi (f + df (i0 + 1)) = angle(r1) + eps i0
    const float e1 = eps * (float)i0;
    const float r1r = fmaf(-e1, rec.lat.r1i, rec.lat.r1r), r1i = fmaf(e1, rec.lat.r1r, rec.lat.r1i);
    const float c1 = fmaf(c0, r1r, -s0 * r1i), s1 = fmaf(c0, r1i, s0 * r1r);
    float2 zrA = make_float2(c0, c1), ziA = make_float2(s0, s1);
    // samples + q: advance = angle(rq) + eps q i0
    const float eq = eps * (float)L.q * (float)i0;
    const float rqr = fmaf(-eq, rec.lat.rqi, rec.lat.rqr), rqi = fmaf(eq, rec.lat.rqr, rec.lat.rqi);
    const float2 q_r = make_float2(rqr, rqr), q_i = make_float2(rqi, rqi), q_ni = make_float2(-rqi, -rqi);
    float2 zrB = pk_fma(zrA, q_r, pk_mul(ziA, q_ni)), ziB = pk_fma(zrA, q_i, pk_mul(ziA, q_r));
    // step phasor over 2 kSynthThreads samples from sample i: angle(w) + th1 i, growing by th2 per step; the growth over the
    // K steps of a thread, th2 k (k - 1) / 2, is replaced by its chord th2 k (K - 2) / 2 (exact at k = 0 and K - 1,
    // off by at most th2 (K - 2)^2 / 8 in between)
    const float chord = ts.th2 * (0.5f * (float)(K - 2));
    const float dA = fmaf(ts.th1, (float)i0, chord), dB = fmaf(ts.th1, (float)(i0 + L.q), chord);
    const float wrA = fmaf(-dA, ts.wi, ts.wr), wiA = fmaf(dA, ts.wr, ts.wi);
    const float wrB = fmaf(-dB, ts.wi, ts.wr), wiB = fmaf(dB, ts.wr, ts.wi);
    const float2 wA_r = make_float2(wrA, wrA), wA_i = make_float2(wiA, wiA), wA_ni = make_float2(-wiA, -wiA);
    const float2 wB_r = make_float2(wrB, wrB), wB_i = make_float2(wiB, wiB), wB_ni = make_float2(-wiB, -wiB);

#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint32_t posb = pos + L.d8_20;
        const uint32_t ha = pos >> 20, hb = posb >> 20;
        const uint4 w = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(ent) + ((ha >> 1) & 0xfffffff0u));
        const uint32_t sb = hb - (ha & ~31u);                                   // <= 32
        const uint32_t ta0 = funnel_r(w.x, w.y, ha), ta1 = funnel_rc(w.x, w.y, sb);
        const uint32_t tb0 = funnel_r(w.z, w.w, ha), tb1 = funnel_rc(w.z, w.w, sb);
        const uint32_t ca = (uint32_t)cls[2 * kSynthThreads * k], cb = (uint32_t)cls[2 * kSynthThreads * k + 1];
        const unsigned char* yt = reinterpret_cast<const unsigned char*>(ytab2);
        constexpr uint32_t RS = 4u * kLatYStride;                              // row stride in bytes: one multiply-add per look-up
        float2 yA = make_float2(*reinterpret_cast<const float*>(yt + (ca * RS + (ta0 & 0x7cu))), *reinterpret_cast<const float*>(yt + (cb * RS + (ta1 & 0x7cu))));
        const float2 yB = make_float2(*reinterpret_cast<const float*>(yt + (ca * RS + (tb0 & 0x7cu))), *reinterpret_cast<const float*>(yt + (cb * RS + (tb1 & 0x7cu))));
        if (k == 0 && (ts.flags & 8u) && tid < 4u) yA = make_float2(rec.yfix[2u * tid], rec.yfix[2u * tid + 1u]);
        arA[k] = pk_fma(yA, zrA, arA[k]);
        aiA[k] = pk_fma(yA, ziA, aiA[k]);
        arB[k] = pk_fma(yB, zrB, arB[k]);
        aiB[k] = pk_fma(yB, ziB, aiB[k]);
        if (k + 1 < K) {
            const float2 nrA = pk_fma(zrA, wA_r, pk_mul(ziA, wA_ni)), niA = pk_fma(zrA, wA_i, pk_mul(ziA, wA_r));
            const float2 nrB = pk_fma(zrB, wB_r, pk_mul(ziB, wB_ni)), niB = pk_fma(zrB, wB_i, pk_mul(ziB, wB_r));
            zrA = nrA; ziA = niA; zrB = nrB; ziB = niB;
            pos += L.step20;
        }
    }

    // flagged block: the windows that hold an oversample the reference's f64 expression puts on the other side of a boundary
    if (ts.flags & 2u) {
#pragma unroll 1
        for (int e = 0; e < 2; ++e) {
            const uint32_t code = (rec.lat.patch >> (16 + 2 * e)) & 3u;
            if (!code || code == 3u) continue;
            const uint32_t qs = (rec.lat.patch & 0xffffu) + (e ? 8u * L.q : 0u);
            const float d2 = code == 1u ? 2.0f : -2.0f;
            const uint32_t i_min = (ts.flags & 8u) ? 8u : 0u;                   // the first eight samples come from yfix (exact already)
#pragma unroll
            for (int k = 0; k < K; ++k) {
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const uint32_t i = i0 + (uint32_t)(2 * kSynthThreads * k) + (uint32_t)(v & 1) + ((v & 2) ? L.q : 0u);
                    const uint32_t tap = (uint32_t)kOversample * i - qs;
                    if (tap <= (uint32_t)(kTaps - 1) && i >= i_min) {
                        float sn, cs;
                        phasor(carrier_phase(ts, i), &sn, &cs);
                        const float dy = d2 * taps[tap] * ts.amp;
                        if (v == 0) { arA[k].x = fmaf(dy, cs, arA[k].x); aiA[k].x = fmaf(dy, sn, aiA[k].x); }
                        if (v == 1) { arA[k].y = fmaf(dy, cs, arA[k].y); aiA[k].y = fmaf(dy, sn, aiA[k].y); }
                        if (v == 2) { arB[k].x = fmaf(dy, cs, arB[k].x); aiB[k].x = fmaf(dy, sn, aiB[k].x); }
                        if (v == 3) { arB[k].y = fmaf(dy, cs, arB[k].y); aiB[k].y = fmaf(dy, sn, aiB[k].y); }
                    }
                }
            }
        }
    }
}

}  // namespace r4wb
