// synth_periodic.cuh — data structures of the period-resident synthesis path (synth_periodic.cu).
//
// When every satellite of a scenario has a constant code delay and a constant Doppler (the `doppler_hz` / `range_m`
// overrides without `orbital_dynamics` or `range_rate_mps`: e1c_prn3_20s_withdoppler, e1c_8prn_20s_clean, e1c_60s_clean,
// ...), the reference's per-satellite baseband after the LPF + decimator (satellite_emitter.rs:264-330, fir.rs:392-409,
// scenario.rs:486-489) repeats EXACTLY every primary-code period of L output samples (20 000 at 5 MHz for E1), up to
// the sign of the primary-code epoch (secondary code / nav bit): there is no code Doppler in the reference model
// (`phase0` is re-derived from the same range every block).  So sample m of period k is
//     sum_s  sign_s[k + o_s(m)] * Y_s[m] * A_s * exp(j phi_s(k L + m))  +  noise
// and a thread that owns the same four slots m in many consecutive periods keeps Y_s[m] * exp(j 2 pi m f_s) in
// registers; per period and satellite it only multiplies by one per-(period, satellite) phasor that carries the epoch
// sign, the amplitude and the reference's exact start-of-period phase.  The few slots whose FIR window straddles an
// epoch boundary (and the slots of the one warp per satellite that contains the boundary) are patched afterwards by
// k_periodic_fix.  Everything is derived from the same block table the general kernel uses, and a device-side check
// (k_static_check) proves the preconditions; scenarios that fail it take the general kernel.
#pragma once
#include <cstdint>

#include "synth.cuh"

namespace r4wb {

constexpr int kPerThreads = 256;          // threads per CTA of k_synth_periodic
constexpr int kPerSlotsPerThread = 4;     // adjacent output samples a thread owns (one 32-byte store per period)
constexpr int kPerWarpSlots = 32 * kPerSlotsPerThread;
constexpr int kPerMaxSats = 8;            // more (virtual) satellites: general kernel
constexpr int kPerMaxCands = 4096;        // slots patched by k_periodic_fix

struct PerSat {
    long long f;          // carrier increment per sample, cycles 0.64
    uint32_t mstar;       // first slot of a period whose primary-code epoch is e_ref + 1 (L when there is none)
    uint32_t e_ref;       // epoch (0 .. epoch_period-1) of slot 0 in the reference period
    uint32_t visible;
    float amp;
    float wr, wi;         // exp(j 2 pi f): rotation between adjacent samples
};
static_assert(sizeof(PerSat) == 32, "PerSat layout");

struct PeriodicArgs {
    const float* ys;        // [NS][L] FIR output of slot m with the epoch sign removed (single-epoch windows: exact)
    const float* yb;        // [NS][L] part of ys that belongs to the previous epoch (0 for all but <= 8 slots per satellite)
    const float4* T;        // [n_periods][NS][2]: (re, re, im, im) of sign[e + j] * amp * exp(j phi(k L)), j = 0, 1
    const PerSat* sat;      // [NS]
    const uint2* cands;     // [n_cands] (slot, satellite mask) patched by k_periodic_fix
    const uint32_t* n_cands;
    const BlockSat* tab;    // canonical block table, row 0 = block tab_blk0
    const SatCode* satcode; // [n_sats]
    float2* out;            // out[0] <-> sample (k0 * L)
    double* power_sum;
    uint64_t k0;            // first period rendered (T[0])
    uint64_t k_ref;         // reference period of e_ref
    uint64_t tab_blk0, B;
    uint32_t n_periods, L, tile_len, n_tiles, KI, n_chunks;
    uint32_t n_sats;        // real count (<= NS)
    uint32_t flags;
    float noise_std;
    uint64_t seed;
};

// arguments of the once-per-table prologue kernels
struct PeriodTableArgs {
    const BlockSat* tab; const uint32_t* perbits; const SatCode* satcode; const float* taps;
    float* ys; float* yb; PerSat* sat; uint2* cands; uint32_t* n_cands;
    uint64_t tab_blk0, B, k_ref, delta46;
    uint32_t n_sats, L, tile_len;
};

}  // namespace r4wb
