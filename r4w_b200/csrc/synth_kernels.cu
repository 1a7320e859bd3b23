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

#include <algorithm>

#include "synth_lattice.cuh"

namespace r4wb {

// ----------------------------------------------------------------------------------------------
// (k_block_params / k_phase_scan: synth_prologue.cu, compiled without FMA contraction)

// Per (table block, chunk, satellite): everything k_synth needs about the tile that costs latency to derive (f64 phasor
// of the rotation step, link to the previous block, the collapsed block-start fix-up), so the synthesis kernel's per-tile
// prologue is one coalesced 96-byte read per satellite.
__global__ void k_tile_params(SynthArgs a, uint32_t tb_begin, uint32_t tb_count, uint32_t tile_samples, TileRec* __restrict__ out)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t total = (uint64_t)tb_count * a.tiles_per_block * a.n_sats;
    const bool live = idx < total;                      // no early exit: the literal-FIR part below is warp-cooperative
    const uint32_t lane = threadIdx.x & 31u;
    const SynthK KK = make_synth_k(a.delta46, a.kmul, a.cj, a.dsum0, a.spc, a.lut_den, a.ystride);
    uint32_t s = 0, tb = 0, chunk = 0;
    TileRec r;
    bool literal = false;                               // first 8 outputs by the reference's own 63-tap loop
    if (live) {
        s = (uint32_t)(idx % a.n_sats);
        const uint32_t tile = (uint32_t)(idx / a.n_sats);
        tb = tb_begin + tile / a.tiles_per_block; chunk = tile % a.tiles_per_block;
        const BlockSat* row = a.tab + (size_t)tb * a.n_sats;
        const BlockHdr hd = a.hdr[tb];
        r.ts = tile_sat(row[s], a.tab, chunk * tile_samples, KK.d8);
        const bool fix = chunk == 0 && (r.ts.flags & 9u) == 9u;
        literal = fix && ((row[s].flags & 2u) || (row[s].prev >= 0 && (a.tab[row[s].prev].flags & 2u)));   // fir_block_start's own test
#pragma unroll 1
        for (int i = 0; i < 8; ++i) {
            float y = 0.0f;
            if (fix && !literal && (uint32_t)i < hd.n)
                y = fir_block_start(row[s], a.tab, a.perbits + (size_t)s * kPerWords, a.taps, a.etab, i, KK, a.satcode[s]);
            r.yfix[i] = y;
        }
        r.lat = TileLat{};
        if (a.lat.q != 0 && chunk == 0) {
            r.lat = tile_lat(row[s], a.lat, a.perbits + (size_t)s * kPerWords, a.spc, a.satcode[s]);
            if (a.stats && (r.ts.flags & 1u) && !lat_rotation_ok(r.ts)) atomicAdd(a.stats, 1u);
        }
    }
    // Records with a half-chip boundary inside the f64 rounding band (about one in 150) need fir_direct for their 8 outputs:
    // 8 x 63 literal chip-index evaluations.  Done per lane, one such record made its whole warp wait for ~120 000
    // instructions (93 % of this kernel's time).  The 8 windows span oversamples [-62, 56] of the block: the warp resolves
    // those 119 signs once (lane l: oversamples l - 62 + 32 r), ballots publish them, and lanes 0..7 add the 63 taps of output
    // i in the reference's order — the same f32 sums as fir_direct, bit for bit.
    unsigned m = __ballot_sync(0xffffffffu, literal);
    while (m) {
        const int src = __ffs((int)m) - 1;
        m &= m - 1u;
        const uint32_t s_src = __shfl_sync(0xffffffffu, s, src), tb_src = __shfl_sync(0xffffffffu, tb, src);
        const BlockSat& cur = a.tab[(size_t)tb_src * a.n_sats + s_src];
        const uint32_t* per = a.perbits + (size_t)s_src * kPerWords;
        const SatCode cd = a.satcode[s_src];
        const uint32_t n_blk = a.hdr[tb_src].n;
        const long long base = -(long long)(kTaps - 1);
        unsigned neg[4], val[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int sgn = tap_sign(cur, a.tab, per, base + (long long)lane + 32 * q, 0, a.delta46, a.spc, cd);
            neg[q] = __ballot_sync(0xffffffffu, sgn == 1);
            val[q] = __ballot_sync(0xffffffffu, sgn >= 0);
        }
        float acc = 0.0f;
        if (lane < 8u && lane < n_blk) {
            const unsigned long long negA = neg[0] | ((unsigned long long)neg[1] << 32), negB = neg[2] | ((unsigned long long)neg[3] << 32);
            const unsigned long long valA = val[0] | ((unsigned long long)val[1] << 32), valB = val[2] | ((unsigned long long)val[3] << 32);
            const int lo = kOversample * (int)lane;                            // window of output `lane`: span indices lo .. lo + 62
            unsigned long long wn, wv;
            if (lo == 0) { wn = negA; wv = valA; }
            else { wn = (negA >> lo) | (negB << (64 - lo)); wv = (valA >> lo) | (valB << (64 - lo)); }   // lo <= 56
            for (int k = 0; k < kTaps; ++k) {                                  // the reference's order; an absent tap adds nothing
                const int j = kTaps - 1 - k;
                if (!((wv >> j) & 1ull)) continue;
                const float t = a.taps[k];
                acc += ((wn >> j) & 1ull) ? -t : t;
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float v = __shfl_sync(0xffffffffu, acc, i);
            if ((int)lane == src) r.yfix[i] = v;
        }
    }
    if (live) out[((size_t)tb * a.tiles_per_block + chunk) * a.n_sats + s] = r;
}

void launch_tile_params(const SynthArgs& a, uint32_t tb_begin, uint32_t tb_count, uint32_t tile_samples, TileRec* out, cudaStream_t st)
{
    const uint64_t total = (uint64_t)tb_count * a.tiles_per_block * a.n_sats;
    if (total == 0) return;
    k_tile_params<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(a, tb_begin, tb_count, tile_samples, out);
    R4WB_LAUNCH_CHECK();
}

// ----------------------------------------------------------------------------------------------
// synthesis kernel: persistent CTAs, one tile (<= 256*2*K consecutive samples of one 1 ms block) at a time.
// Thread t owns the sample pairs (2t, 2t+1) + 512 k of the tile, so every store is a 16-byte float4 and a
// warp writes 512 contiguous bytes.  Per satellite and pair: one 64-bit code-NCO step, two boundary-age classes,
// one shared-memory window load for both 5-sign patterns, two table look-ups, and a packed (FFMA2) phasor recurrence.
constexpr int kThreads = kSynthThreads;

struct SynthSmem {
    float* ytab; float* taps; float* etab; uint32_t* per; TileRec* trec; uint2* t64; uint32_t* w32; uint8_t* clslut;
};

__device__ __forceinline__ SynthSmem carve_smem(unsigned char* raw, uint32_t n_sats, uint32_t nw64, uint32_t ystride)
{
    SynthSmem m;
    m.ytab = reinterpret_cast<float*>(raw);                                   // [32][ystride]
    m.taps = m.ytab + 32 * ystride;                                          // [64]
    m.etab = m.taps + 64;                                                     // [64]
    size_t off = ((size_t)(32 * ystride + 128) * 4 + 15) & ~(size_t)15;
    m.per = reinterpret_cast<uint32_t*>(raw + off);                           // [n_sats][kPerWords]
    off += (size_t)n_sats * kPerWords * 4;
    m.trec = reinterpret_cast<TileRec*>(raw + off);                           // [n_sats]
    off += (size_t)n_sats * sizeof(TileRec);
    m.t64 = reinterpret_cast<uint2*>(raw + off);                              // [n_sats][nw64]
    off += (size_t)n_sats * nw64 * 8;
    m.w32 = reinterpret_cast<uint32_t*>(raw + off);                           // [n_sats][nw64 + 1]
    off += (size_t)n_sats * (nw64 + 1) * 4;
    off = (off + 15) & ~(size_t)15;
    m.clslut = raw + off;                                                     // [lut_den padded to 16]
    return m;
}

// sink conversions of core/io/format.rs:203-222 (Rust `as` = truncation toward zero after the clamp)
// in f64 like the reference, so the code is exactly what write_sample produces for the f32 sample widened to f64
__device__ __forceinline__ int to_i16(float x) { return __double2int_rz(fmin(fmax((double)x * 32767.0, -32768.0), 32767.0)); }
__device__ __forceinline__ int to_i8(float x) { return __double2int_rz(fmin(fmax((double)x * 127.0, -128.0), 127.0)); }
__device__ __forceinline__ int to_u8(float x) { return __double2int_rz(fmin(fmax(((double)x + 1.0) * 127.5, 0.0), 255.0)); }

// one complex sample (re, im) in output format FMT at sample index o
template <int FMT>
__device__ __forceinline__ void store_sample(void* out, uint64_t o, float re, float im)
{
    if (FMT == R4WB_FMT_CF64) reinterpret_cast<double2*>(out)[o] = make_double2((double)re, (double)im);
    else if (FMT == R4WB_FMT_CF32) reinterpret_cast<float2*>(out)[o] = make_float2(re, im);
    else if (FMT == R4WB_FMT_CI16) reinterpret_cast<uint32_t*>(out)[o] = ((uint32_t)to_i16(re) & 0xffffu) | ((uint32_t)to_i16(im) << 16);
    else if (FMT == R4WB_FMT_CI8) reinterpret_cast<uint16_t*>(out)[o] = (uint16_t)(((uint32_t)to_i8(re) & 0xffu) | (((uint32_t)to_i8(im) & 0xffu) << 8));
    else reinterpret_cast<uint16_t*>(out)[o] = (uint16_t)((uint32_t)to_u8(re) | ((uint32_t)to_u8(im) << 8));
}
// two adjacent samples at even sample index o (out aligned to the pair)
template <int FMT>
__device__ __forceinline__ void store_pair(void* out, uint64_t o, float4 v)
{
    if (FMT == R4WB_FMT_CF32) reinterpret_cast<float4*>(out)[o >> 1] = v;
    else if (FMT == R4WB_FMT_CI16)
        reinterpret_cast<uint2*>(out)[o >> 1] = make_uint2(((uint32_t)to_i16(v.x) & 0xffffu) | ((uint32_t)to_i16(v.y) << 16),
                                                           ((uint32_t)to_i16(v.z) & 0xffffu) | ((uint32_t)to_i16(v.w) << 16));
    else if (FMT == R4WB_FMT_CI8)
        reinterpret_cast<uint32_t*>(out)[o >> 1] = ((uint32_t)to_i8(v.x) & 0xffu) | (((uint32_t)to_i8(v.y) & 0xffu) << 8) |
                                                   (((uint32_t)to_i8(v.z) & 0xffu) << 16) | ((uint32_t)to_i8(v.w) << 24);
    else if (FMT == R4WB_FMT_CU8)
        reinterpret_cast<uint32_t*>(out)[o >> 1] = (uint32_t)to_u8(v.x) | ((uint32_t)to_u8(v.y) << 8) | ((uint32_t)to_u8(v.z) << 16) |
                                                   ((uint32_t)to_u8(v.w) << 24);
    else { store_sample<FMT>(out, o, v.x, v.y); store_sample<FMT>(out, o + 1, v.z, v.w); }
}

template <int K, int FMT>
__global__ void __launch_bounds__(kThreads, K <= 5 ? 3 : 2) k_synth(SynthArgs a)
{
    constexpr int TILE = kThreads * 2 * K;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SynthSmem sm = carve_smem(smem_raw, a.n_sats, a.nw64, a.ystride);
    __shared__ float s_pow[kThreads / 32];

    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const SynthK KK = make_synth_k(a.delta46, a.kmul, a.cj, a.dsum0, a.spc, a.lut_den, a.ystride);
    const PhiloxKeys PK = philox_keys(a.seed);

    // kernel-lifetime tables
    for (uint32_t k = tid; k < 32 * a.ystride; k += kThreads) sm.ytab[k] = a.ytab[k];
    for (uint32_t k = tid; k < 64; k += kThreads) { sm.taps[k] = a.taps[k]; sm.etab[k] = a.etab[k]; }
    for (uint32_t k = tid; k < a.n_sats * kPerWords; k += kThreads) sm.per[k] = a.perbits[k];
    for (uint32_t k = tid; k < ((a.lut_den + 15u) >> 4); k += kThreads)
        reinterpret_cast<uint4*>(sm.clslut)[k] = reinterpret_cast<const uint4*>(a.clslut)[k];

    float pow_acc = 0.0f;
    const uint32_t n_tiles = a.tb_count * a.tiles_per_block;
    const uint32_t rec_f4 = a.n_sats * (uint32_t)(sizeof(TileRec) / 16);        // float4s of one tile's records
    const float4* recs = reinterpret_cast<const float4*>(a.tiles + (size_t)a.tb_begin * a.tiles_per_block * a.n_sats);
    // the next tile's records travel in registers while the current tile is rendered
    float4 pre[2];
    {
        const uint32_t tile = blockIdx.x;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const uint32_t k = tid + (uint32_t)u * kThreads;
            pre[u] = (tile < n_tiles && k < rec_f4) ? __ldg(recs + (size_t)tile * rec_f4 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }

    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t tb = a.tb_begin + tile / a.tiles_per_block;
        const uint32_t chunk = tile % a.tiles_per_block;
        const BlockHdr hd = a.hdr[tb];
        const uint32_t i_begin = chunk * TILE;
        const uint32_t i_end = min(hd.n, i_begin + TILE);
        // tiles past the block's end or entirely outside the requested output range are skipped (uniform per CTA)
        const bool skip = i_begin >= hd.n || hd.first + i_end <= a.out_first || hd.first + i_begin >= a.out_first + a.out_n;
        const BlockSat* row = a.tab + (size_t)tb * a.n_sats;

        __syncthreads();   // previous tile's readers are done (and the kernel-lifetime tables are in place)
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const uint32_t k = tid + (uint32_t)u * kThreads;
            if (k < rec_f4) reinterpret_cast<float4*>(sm.trec)[k] = pre[u];
        }
        {
            const uint32_t nt = tile + gridDim.x;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const uint32_t k = tid + (uint32_t)u * kThreads;
                if (nt < n_tiles && k < rec_f4) pre[u] = __ldg(recs + (size_t)nt * rec_f4 + k);
            }
        }
        if (skip) continue;
        __syncthreads();
        // half-chip sign words: bit n of word w <-> half-chip hb + 32 w + n
        for (uint32_t k = tid; k < a.n_sats * (a.nw64 + 1); k += kThreads) {
            const uint32_t s = k / (a.nw64 + 1), w = k - s * (a.nw64 + 1);
            sm.w32[k] = sign_word(sm.per + s * kPerWords, sm.trec[s].ts.hb, w, a.satcode[s]);
        }
        __syncthreads();
        for (uint32_t k = tid; k < a.n_sats * a.nw64; k += kThreads) {
            const uint32_t s = k / a.nw64, w = k - s * a.nw64;
            sm.t64[k] = make_uint2(sm.w32[s * (a.nw64 + 1) + w], sm.w32[s * (a.nw64 + 1) + w + 1]);
        }
        __syncthreads();

        float2 ar[K], ai[K];      // (re_a, re_b), (im_a, im_b) of pair k
#pragma unroll
        for (int k = 0; k < K; ++k) ar[k] = ai[k] = make_float2(0.0f, 0.0f);

        for (uint32_t s = 0; s < a.n_sats; ++s) {
            const TileSat ts = sm.trec[s].ts;
            if (!(ts.flags & 1u)) continue;
            const SlowCtx slow{row + s, a.tab, sm.per + s * kPerWords, sm.taps, a.satcode + s};
            sat_accumulate<K>(ts, KK, sm.t64 + s * a.nw64, sm.ytab, sm.clslut, sm.trec[s].yfix, slow, tid, i_begin, i_end, ar, ai, nullptr);
        }

        // noise, power, store.  A tile that lies inside the requested range and starts on a 16-byte boundary of the
        // output (the usual case) stores float4 pairs without per-sample range checks.
        const bool noise_on = !(a.flags & R4WB_FLAG_NOISE_OFF);
        const uint64_t m0 = hd.first + i_begin;
        const bool tile_plain = a.out_aligned16 && m0 >= a.out_first && hd.first + i_end <= a.out_first + a.out_n &&
                                (((m0 - a.out_first) | (uint64_t)i_end) & 1ull) == 0;
        if (tile_plain) {
            const uint64_t o0 = (m0 - a.out_first) + 2u * tid;  // even
            const uint64_t ctr0 = (m0 >> 1) + tid;               // m0 is even when out_first is; odd handled below
            const bool m_even = (m0 & 1ull) == 0;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (i_begin + 2 * tid + 2 * kThreads * k >= i_end) continue;
                float4 v = make_float4(ar[k].x, ai[k].x, ar[k].y, ai[k].y);
                if (noise_on) {
                    float2 ga, gb2;
                    const uint64_t ctr = ctr0 + (uint64_t)(kThreads * k);
                    if (m_even) {
                        noise_of_counter(ctr, PK, ga, gb2);
                    } else {
                        float2 t0, t1;
                        noise_of_counter(ctr, PK, t0, ga);
                        noise_of_counter(ctr + 1, PK, gb2, t1);
                    }
                    v.x = fmaf(ga.x, a.noise_std, v.x); v.y = fmaf(ga.y, a.noise_std, v.y);
                    v.z = fmaf(gb2.x, a.noise_std, v.z); v.w = fmaf(gb2.y, a.noise_std, v.w);
                }
                pow_acc += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
                store_pair<FMT>(a.out, o0 + (uint64_t)(2 * kThreads * k), v);
            }
            continue;
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const uint32_t ia = i_begin + 2 * tid + 2 * kThreads * k;
            if (ia >= i_end) continue;
            const uint64_t m = hd.first + ia;                    // global sample index
            const bool has_b = ia + 1 < i_end;
            float2 va = make_float2(ar[k].x, ai[k].x), vb = make_float2(ar[k].y, ai[k].y);
            if (noise_on) {
                float2 ga, gb2;
                if ((m & 1ull) == 0) {      // one Philox draw covers both samples of the pair
                    noise_of_counter(m >> 1, PK, ga, gb2);
                } else {
                    float2 t0, t1;
                    noise_of_counter(m >> 1, PK, t0, ga);
                    noise_of_counter((m >> 1) + 1, PK, gb2, t1);
                }
                va.x = fmaf(ga.x, a.noise_std, va.x); va.y = fmaf(ga.y, a.noise_std, va.y);
                vb.x = fmaf(gb2.x, a.noise_std, vb.x); vb.y = fmaf(gb2.y, a.noise_std, vb.y);
            }
            const bool wa = m >= a.out_first && m < a.out_first + a.out_n;
            const bool wb = has_b && (m + 1) >= a.out_first && (m + 1) < a.out_first + a.out_n;
            if (wa) pow_acc += va.x * va.x + va.y * va.y;
            if (wb) pow_acc += vb.x * vb.x + vb.y * vb.y;
            const uint64_t o = m - a.out_first;                  // only meaningful when wa (wraps otherwise)
            if (wa && wb && ((o & 1ull) == 0) && a.out_aligned16) {
                store_pair<FMT>(a.out, o, make_float4(va.x, va.y, vb.x, vb.y));
            } else {
                if (wa) store_sample<FMT>(a.out, o, va.x, va.y);
                if (wb) store_sample<FMT>(a.out, o + 1, vb.x, vb.y);
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
// Direct path: satellites whose chip rate is not 1.023 MHz (GPS L5 at 10.23 Mchip/s: up to 17 chip boundaries inside one
// 63-tap window; GLONASS L1OF at 0.511 Mchip/s: another sample lattice).  One thread per output sample evaluates the
// reference's 63-tap sum literally per satellite (f64 chip index per tap, synth_math.cuh::direct_fir) and ADDS the rotated
// contribution to the sample k_synth has already written (noise and the 1.023 MHz satellites).  Slow next to the collapsed
// path (63 f64 divisions per satellite-sample) but the same arithmetic as the reference; float formats only.
template <typename OutT>
__global__ void __launch_bounds__(256) k_synth_direct(SynthArgs a)
{
    __shared__ double s_dp[8];
    const uint32_t tb = a.tb_begin + blockIdx.y + a.direct_y0, i = blockIdx.x * blockDim.x + threadIdx.x;
    const BlockHdr hd = a.hdr[tb];
    const uint64_t m = hd.first + i;
    double dp = 0.0;
    if (i < hd.n && m >= a.out_first && m < a.out_first + a.out_n) {
        const BlockSat* row = a.tab + (size_t)tb * a.n_sats;
        float re = 0.0f, im = 0.0f;
        bool any = false;
        for (uint32_t s = 0; s < a.n_sats; ++s) {
            const DirectSat d = a.dsat[s];
            if (!d.direct || !(row[s].flags & 1u)) continue;
            const float2 v = direct_sample(row[s], a.tab, a.dcode + (size_t)s * kDirectWords, a.taps, i, d);
            re += v.x; im += v.y;
            any = true;
        }
        if (any) {
            OutT* o = reinterpret_cast<OutT*>(a.out) + (m - a.out_first);
            const OutT old = *o;
            OutT nw;
            nw.x = old.x + re; nw.y = old.y + im;
            *o = nw;
            const float ox = (float)old.x, oy = (float)old.y, nx = (float)nw.x, ny = (float)nw.y;     // power of the cf32 sample, as k_synth
            dp = ((double)nx * nx + (double)ny * ny) - ((double)ox * ox + (double)oy * oy);
        }
    }
    if (a.power_sum) {
        for (int off = 16; off > 0; off >>= 1) dp += __shfl_xor_sync(0xffffffffu, dp, off);
        if ((threadIdx.x & 31u) == 0) s_dp[threadIdx.x >> 5] = dp;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w) t += s_dp[w];
            if (t != 0.0) atomicAdd(a.power_sum, t);
        }
    }
}

void launch_synth_direct(const SynthArgs& a, r4wb_fmt fmt, cudaStream_t st)
{
    if (fmt != R4WB_FMT_CF32 && fmt != R4WB_FMT_CF64)
        fail(R4WB_ERR_NOT_SUPPORTED, "GPS L5 / GLONASS satellites are rendered in the float formats only (cf32, cf64)");
    if (a.tb_count == 0 || a.max_block_n == 0) return;
    // gridDim.y is limited to 65 535: long renders go in slabs of table blocks
    for (uint32_t y0 = 0; y0 < a.tb_count; y0 += 65535u) {
        SynthArgs b = a;
        b.direct_y0 = y0;
        const dim3 grid((a.max_block_n + 255) / 256, std::min<uint32_t>(65535u, a.tb_count - y0));
        if (fmt == R4WB_FMT_CF32) k_synth_direct<float2><<<grid, 256, 0, st>>>(b);
        else k_synth_direct<double2><<<grid, 256, 0, st>>>(b);
        R4WB_LAUNCH_CHECK();
    }
}

// ----------------------------------------------------------------------------------------------
// launchers (called from synth_host.cu)
template <int K, int FMT>
static void launch_synth_t(const SynthArgs& a, int grid, size_t smem, cudaStream_t st)
{
    static PerDeviceOnce attr_done;
    if (attr_done.first()) {
        R4WB_CUDA(cudaFuncSetAttribute(k_synth<K, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    k_synth<K, FMT><<<grid, kThreads, smem, st>>>(a);
    R4WB_LAUNCH_CHECK();
}

template <int K, int FMT>
static int max_blocks_t(size_t smem)
{
    int nb = 0;
    R4WB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_synth<K, FMT>, kThreads, smem));
    return nb;
}

// K = 5 (tuning hook) exists for the float formats only
void launch_synth_kernel(const SynthArgs& a, int K, r4wb_fmt fmt, int grid, cudaStream_t st)
{
    const size_t smem = synth_smem_bytes(a.n_sats, a.nw64, a.lut_den, a.ystride);
    if (smem > 200 * 1024) fail(R4WB_ERR_NOT_SUPPORTED, "scenario needs %zu bytes of shared memory per CTA", smem);
    if (K == 5 && fmt == R4WB_FMT_CF32) return launch_synth_t<5, R4WB_FMT_CF32>(a, grid, smem, st);
    if (K == 5 && fmt == R4WB_FMT_CF64) return launch_synth_t<5, R4WB_FMT_CF64>(a, grid, smem, st);
    if (K != 10) fail(R4WB_ERR_INVALID_PARAMETER, "unsupported tile factor %d for this format", K);
    switch (fmt) {
    case R4WB_FMT_CF32: return launch_synth_t<10, R4WB_FMT_CF32>(a, grid, smem, st);
    case R4WB_FMT_CF64: return launch_synth_t<10, R4WB_FMT_CF64>(a, grid, smem, st);
    case R4WB_FMT_CI16: return launch_synth_t<10, R4WB_FMT_CI16>(a, grid, smem, st);
    case R4WB_FMT_CI8: return launch_synth_t<10, R4WB_FMT_CI8>(a, grid, smem, st);
    case R4WB_FMT_CU8: return launch_synth_t<10, R4WB_FMT_CU8>(a, grid, smem, st);
    }
    fail(R4WB_ERR_INVALID_PARAMETER, "unknown sample format %d", (int)fmt);
}

int synth_max_blocks_per_sm(int K, r4wb_fmt fmt, size_t smem)
{
    if (K == 5 && fmt == R4WB_FMT_CF32) return max_blocks_t<5, R4WB_FMT_CF32>(smem);
    if (K == 5 && fmt == R4WB_FMT_CF64) return max_blocks_t<5, R4WB_FMT_CF64>(smem);
    switch (fmt) {
    case R4WB_FMT_CF64: return max_blocks_t<10, R4WB_FMT_CF64>(smem);
    case R4WB_FMT_CI16: return max_blocks_t<10, R4WB_FMT_CI16>(smem);
    case R4WB_FMT_CI8: return max_blocks_t<10, R4WB_FMT_CI8>(smem);
    case R4WB_FMT_CU8: return max_blocks_t<10, R4WB_FMT_CU8>(smem);
    default: return max_blocks_t<10, R4WB_FMT_CF32>(smem);
    }
}

}  // namespace r4wb
