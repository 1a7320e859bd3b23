// synth_periodic.cu — period-resident synthesis kernels for scenarios whose satellites all have a constant code
// delay and Doppler (see synth_periodic.cuh for the algebra).  Same reference code as synth_kernels.cu:
//   SatelliteEmitter::generate_baseband_iq   gnss/satellite_emitter.rs:218-347
//   FirFilter::process + step_by(8)          core/filters/fir.rs:392-409, gnss/scenario.rs:486-489
//   Doppler rotation + amplitude + sum       gnss/scenario.rs:516-528
//   thermal noise                            gnss/scenario.rs:530-542   (same Philox stream as k_synth)
//   cf32 sink cast                           core/io/format.rs:197-200
// Prologue (once per block table): k_static_check proves the preconditions from the block table itself,
// k_period_tables evaluates the 63-tap FIR of every slot of one reference period literally (split by primary-code
// epoch), k_period_cands lists the slots whose epoch sign needs a patch.  Per render: k_period_phasors (one f64 sincos
// per period and satellite), k_synth_periodic, k_periodic_fix (patch pass).
#include <cuda_runtime.h>

#include "synth_math.cuh"
#include "synth_periodic.cuh"

namespace r4wb {

// ----------------------------------------------------------------------------------------------
// preconditions, checked on the table the general kernel would use
__global__ void k_static_check(const BlockSat* __restrict__ tab, const SatConst* __restrict__ sats, uint32_t nblk, uint32_t n_sats,
                               uint64_t B, uint32_t* __restrict__ bad)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)nblk * n_sats) return;
    const uint32_t b = (uint32_t)(idx / n_sats), s = (uint32_t)(idx % n_sats);
    const BlockSat& r = tab[idx];
    const BlockSat& r0 = tab[s];
    bool ok = (r.flags & 1u) == (r0.flags & 1u);
    if (r0.flags & 1u) {
        ok = ok && sats[s].static_phase && r.df == 0 && r.phase0 == r0.phase0 && r.e0 == r0.e0 && r.amp == r0.amp && !(r.flags & 2u);
        ok = ok && (r.n == B || b + 1 == nblk);
        if (b > 0) ok = ok && r.prev == (int32_t)((b - 1) * n_sats + s);
    }
    if (!ok) atomicOr(bad, 1u);
}


// epoch (0 .. epoch_period-1) and position inside the primary-code period of the half-chip at fixed-point position u
__device__ __forceinline__ void epoch_of(uint64_t u, const SatCode& cd, uint32_t& e, uint32_t& p)
{
    uint32_t h = (uint32_t)(u >> kFracBits);
    if (h >= cd.hc_mod) h -= cd.hc_mod;
    e = h / cd.per_len;
    p = h - e * cd.per_len;
}

// One thread per (satellite, slot): the reference's 63-tap sum over the +-1 oversampled baseband of sample
// k_ref * L + m (fir.rs:392-409 over satellite_emitter.rs:264-330) with the epoch sign left out, split into the
// taps inside the sample's own primary-code epoch (A) and those of the epoch before (B).  f64 accumulation.
__global__ void k_period_tables(PeriodTableArgs a)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)a.n_sats * a.L) return;
    const uint32_t s = (uint32_t)(idx / a.L), m = (uint32_t)(idx % a.L);
    const SatCode cd = a.satcode[s];
    const uint32_t* per = a.perbits + (size_t)s * kPerWords;
    const uint64_t g0 = a.k_ref * a.L, g = g0 + m;
    const BlockSat& first = a.tab[(size_t)(g0 / a.B - a.tab_blk0) * a.n_sats + s];
    const BlockSat& cur = a.tab[(size_t)(g / a.B - a.tab_blk0) * a.n_sats + s];
    const uint32_t i = (uint32_t)(g % a.B);
    uint32_t e0, em, p;
    epoch_of(first.U + (uint64_t)kOversample * (uint64_t)(g0 % a.B) * a.delta46, cd, e0, p);
    epoch_of(cur.U + (uint64_t)kOversample * i * a.delta46, cd, em, p);
    double A = 0.0, Bv = 0.0;
    for (int k = 0; k < kTaps; ++k) {
        long long q = (long long)kOversample * i - k;
        const BlockSat* bs = &cur;
        if (q < 0) {
            if (cur.prev < 0) continue;
            bs = &a.tab[cur.prev];
            q += (long long)kOversample * bs->n;
            if (q < 0) continue;
        }
        uint32_t eq, pq;
        epoch_of(bs->U + (uint64_t)q * a.delta46, cd, eq, pq);
        const double v = ((per[pq >> 5] >> (pq & 31u)) & 1u) ? -(double)a.taps[k] : (double)a.taps[k];
        if (eq == em) A += v; else Bv += v;
    }
    a.ys[idx] = (float)(A + Bv);
    a.yb[idx] = (float)Bv;
    const uint32_t o = (em + cd.epoch_period - e0) % cd.epoch_period;
    if (o != 0) atomicMin(&a.sat[s].mstar, m);
    if (m == 0) a.sat[s].e_ref = e0;
}

__global__ void k_period_finish(PeriodTableArgs a, uint32_t ns_padded)
{
    const uint32_t s = threadIdx.x;
    if (s >= ns_padded) return;
    PerSat ps = a.sat[s];
    if (s >= a.n_sats) {
        ps = PerSat{0, a.L, 0u, 0u, 0.0f, 1.0f, 0.0f};
    } else {
        const BlockSat& r = a.tab[(size_t)(a.k_ref * a.L / a.B - a.tab_blk0) * a.n_sats + s];
        ps.f = r.f;
        ps.mstar = min(ps.mstar, a.L);
        ps.visible = r.flags & 1u;
        ps.amp = r.amp;
        double sn, cs;
        sincospi((double)r.f * 1.0842021724855044e-19 /* 2^-63 */, &sn, &cs);
        ps.wr = (float)cs; ps.wi = (float)sn;
    }
    a.sat[s] = ps;
}

// first slot of the warp of k_synth_periodic that owns slot m
__host__ __device__ __forceinline__ uint32_t period_warp_first(uint32_t m, uint32_t tile_len)
{
    const uint32_t tile = m / tile_len, r = m - tile * tile_len;
    return tile * tile_len + (r / (uint32_t)kPerWarpSlots) * (uint32_t)kPerWarpSlots;
}

// slots k_periodic_fix must patch: the FIR window reaches into the previous epoch, or the slot's epoch differs from
// the one its warp was resolved to
__global__ void k_period_cands(PeriodTableArgs a)
{
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= a.L) return;
    const uint32_t wf = period_warp_first(m, a.tile_len);
    uint32_t smask = 0;                               // satellites that need a patch at this slot
    for (uint32_t s = 0; s < a.n_sats; ++s) {
        const PerSat ps = a.sat[s];
        if (!ps.visible) continue;
        if (a.yb[(size_t)s * a.L + m] != 0.0f || ((m >= ps.mstar) != (wf >= ps.mstar))) smask |= 1u << s;
    }
    if (smask) {
        const uint32_t k = atomicAdd(a.n_cands, 1u);
        if (k < (uint32_t)kPerMaxCands) a.cands[k] = make_uint2(m, smask);
    }
}

// ----------------------------------------------------------------------------------------------
// per-(period, satellite) phasors: the reference's exact start-of-period carrier phase (64-bit cycles from the block
// table), the amplitude and the epoch sign for the two epoch offsets a warp can be resolved to
__global__ void k_period_phasors(PeriodicArgs a, uint32_t ns_padded, uint64_t tab_blk1, float4* __restrict__ T)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)a.n_periods * ns_padded) return;
    const uint32_t kk = (uint32_t)(idx / ns_padded), s = (uint32_t)(idx % ns_padded);
    float4* o = T + idx * 2;
    const PerSat ps = a.sat[s];
    if (s >= a.n_sats || !ps.visible) {
        o[0] = o[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    const uint64_t k = a.k0 + kk, M = k * a.L;
    uint64_t b = M / a.B;
    b = b < a.tab_blk0 ? a.tab_blk0 : (b >= tab_blk1 ? tab_blk1 - 1 : b);
    const long long i = (long long)M - (long long)(b * a.B);
    const BlockSat& r = a.tab[(size_t)(b - a.tab_blk0) * a.n_sats + s];
    const uint64_t ph = r.phi + (uint64_t)((i + 1) * r.f);           // phase of sample M (scenario.rs:519-524), df == 0
    double sn, cs;
    sincospi((double)(long long)ph * 1.0842021724855044e-19 /* 2^-63 */, &sn, &cs);
    const SatCode cd = a.satcode[s];
    const long long P = (long long)cd.epoch_period;
    long long dk = ((long long)k - (long long)a.k_ref) % P;
    if (dk < 0) dk += P;
    for (int j = 0; j < 2; ++j) {
        const uint32_t e = (uint32_t)(((long long)ps.e_ref + dk + j) % P);
        const double sg = ((cd.epoch_bits >> e) & 1ull) ? -(double)ps.amp : (double)ps.amp;
        const float re = (float)(sg * cs), im = (float)(sg * sn);
        o[j] = make_float4(re, re, im, im);
    }
}

// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 pk_sub(float2 a, float2 b)
{
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
    return reinterpret_cast<float2&>(d);
}

// One CTA = (tile of <= 1024 slots, chunk of KI periods).  Thread t owns slots tile*tile_len + 4t .. +3 in every
// period of the chunk: NS x 4 complex registers q = Y_s[m] exp(j 2 pi m f_s) stay resident; per period and satellite
// one 16-byte shared-memory read of the warp's phasor and 8 packed FMAs, then the noise pair of each sample and one
// 32-byte store.  A warp writes 1 KiB of contiguous output per period.
template <int NS>
__device__ __forceinline__ float periodic_render(const PeriodicArgs& a, float4* s_T, uint32_t tile, uint32_t kk0, uint32_t kk1)
{
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t wfirst = tile * a.tile_len + (uint32_t)kPerWarpSlots * warp;
    if ((uint32_t)kPerWarpSlots * warp >= a.tile_len || wfirst >= a.L) return 0.0f;               // whole warp idle
    const uint32_t m0 = wfirst + (uint32_t)kPerSlotsPerThread * lane;
    const bool active = (uint32_t)kPerWarpSlots * warp + (uint32_t)kPerSlotsPerThread * lane < a.tile_len && m0 < a.L;

    // the warp's phasors: epoch offset resolved per satellite from the warp's first slot
    float4* Tw = s_T + (size_t)warp * a.KI * NS;
    {
        const uint32_t n = (kk1 - kk0) * NS;
        for (uint32_t k = lane; k < n; k += 32u) {
            const uint32_t s = k % NS;
            const uint32_t ow = wfirst >= a.sat[s].mstar ? 1u : 0u;
            Tw[k] = __ldg(a.T + ((size_t)kk0 * NS + k) * 2 + ow);
        }
        __syncwarp();
    }

    float2 qr[NS][2], qi[NS][2];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const PerSat ps = a.sat[s];
        const float4 y = active ? __ldg(reinterpret_cast<const float4*>(a.ys + (size_t)s * a.L + m0)) : make_float4(0.f, 0.f, 0.f, 0.f);
        float s0, c0;
        accurate_sincos_cycles((uint64_t)m0 * (uint64_t)ps.f, &s0, &c0);
        const float c1 = c0 * ps.wr - s0 * ps.wi, s1 = fmaf(c0, ps.wi, s0 * ps.wr);
        const float c2 = c1 * ps.wr - s1 * ps.wi, s2 = fmaf(c1, ps.wi, s1 * ps.wr);
        const float c3 = c2 * ps.wr - s2 * ps.wi, s3 = fmaf(c2, ps.wi, s2 * ps.wr);
        qr[s][0] = make_float2(y.x * c0, y.y * c1); qi[s][0] = make_float2(y.x * s0, y.y * s1);
        qr[s][1] = make_float2(y.z * c2, y.w * c3); qi[s][1] = make_float2(y.z * s2, y.w * s3);
    }

    const PhiloxKeys PK = philox_keys(a.seed);
    const bool noise_on = !(a.flags & R4WB_FLAG_NOISE_OFF);
    const float sigma = noise_on ? a.noise_std : 0.0f;
    const float2 sig2 = make_float2(sigma, sigma);
    float2 pw = make_float2(0.0f, 0.0f);
    uint64_t g = (a.k0 + kk0) * (uint64_t)a.L + m0;                       // global sample index, multiple of 4
    float* op = reinterpret_cast<float*>(a.out + ((uint64_t)kk0 * a.L + m0));
    const float4* Tp = Tw;

    // The noise of period k+1 is drawn while period k is summed: the ten Philox rounds of the two counters are spread
    // over the satellite loop (integer work between the packed FMAs), the Box-Muller transforms (MUFU) run at the end
    // of the iteration and complete under the next iteration's FMAs.
    float2 n0 = make_float2(0.f, 0.f), n1 = n0, n2 = n0, n3 = n0;         // (re, im) noise of the four samples
    if (noise_on) {
        noise_of_counter(g >> 1, PK, n0, n1);
        noise_of_counter((g >> 1) + 1, PK, n2, n3);
    }

#pragma unroll 1
    for (uint32_t kk = kk0; kk < kk1; ++kk, op += 2 * (size_t)a.L, Tp += NS) {
        g += a.L;
        const uint64_t ctr = g >> 1;                                      // next period's first counter (even)
        uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0u, c3 = 0u;
        uint32_t d0 = c0 | 1u, d1 = c1, d2 = 0u, d3 = 0u;
        float2 xr0 = make_float2(0.f, 0.f), xi0 = xr0, yr0 = xr0, yi0 = xr0, xr1 = xr0, xi1 = xr0, yr1 = xr0, yi1 = xr0;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const float4 t = Tp[s];
            const float2 trr = make_float2(t.x, t.y), tii = make_float2(t.z, t.w);
            xr0 = pk_fma(qr[s][0], trr, xr0); xi0 = pk_fma(qr[s][0], tii, xi0);
            yr0 = pk_fma(qi[s][0], trr, yr0); yi0 = pk_fma(qi[s][0], tii, yi0);
            xr1 = pk_fma(qr[s][1], trr, xr1); xi1 = pk_fma(qr[s][1], tii, xi1);
            yr1 = pk_fma(qi[s][1], trr, yr1); yi1 = pk_fma(qi[s][1], tii, yi1);
            if (noise_on) {
#pragma unroll
                for (int r = (10 * s) / NS; r < (10 * (s + 1)) / NS; ++r) {
                    philox_round(c0, c1, c2, c3, PK.k0[r], PK.k1[r]);
                    philox_round(d0, d1, d2, d3, PK.k0[r], PK.k1[r]);
                }
            }
        }
        // (qr + j qi)(tr + j ti): re = qr tr - qi ti, im = qr ti + qi tr
        float2 re0 = pk_sub(xr0, yi0), im0 = pk_add(xi0, yr0), re1 = pk_sub(xr1, yi1), im1 = pk_add(xi1, yr1);
        re0 = pk_fma(make_float2(n0.x, n1.x), sig2, re0); im0 = pk_fma(make_float2(n0.y, n1.y), sig2, im0);
        re1 = pk_fma(make_float2(n2.x, n3.x), sig2, re1); im1 = pk_fma(make_float2(n2.y, n3.y), sig2, im1);
        pw = pk_fma(re0, re0, pw); pw = pk_fma(im0, im0, pw); pw = pk_fma(re1, re1, pw); pw = pk_fma(im1, im1, pw);
        if (active)
            asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(op), "f"(re0.x), "f"(im0.x), "f"(re0.y), "f"(im0.y),
                         "f"(re1.x), "f"(im1.x), "f"(re1.y), "f"(im1.y) : "memory");
        if (noise_on) {
            n0 = gauss_pair(c0, c1); n1 = gauss_pair(c2, c3);
            n2 = gauss_pair(d0, d1); n3 = gauss_pair(d2, d3);
        }
    }

    return active ? pw.x + pw.y : 0.0f;
}

// One CTA = (tile of <= 1024 slots, chunk of KI periods).
template <int NS>
__global__ void __launch_bounds__(kPerThreads, 2) k_synth_periodic(PeriodicArgs a)
{
    extern __shared__ float4 s_T[];                       // [warps][KI][NS]
    const uint32_t tile = blockIdx.x % a.n_tiles, chunk = blockIdx.x / a.n_tiles;
    const uint32_t kk0 = chunk * a.KI, kk1 = min(a.n_periods, kk0 + a.KI);
    if (kk0 >= kk1) return;
    float p = periodic_render<NS>(a, s_T, tile, kk0, kk1);
    if (a.power_sum) {
        for (int off = 16; off > 0; off >>= 1) p += __shfl_xor_sync(0xffffffffu, p, off);
        if ((threadIdx.x & 31u) == 0 && p != 0.0f) atomicAdd(a.power_sum, (double)p);
    }
}

// Patch pass.  A thread owns one listed slot in kFixPeriods consecutive periods and adds the (true - rendered)
// contribution of the satellites whose epoch sign the render kernel could not resolve per warp (slot on the other side
// of the epoch boundary than the warp's first slot) or whose FIR window straddles the epoch boundary (two signs inside
// one window).  32-bit index arithmetic, one division per thread; all table reads first, then the loads of the rendered
// samples, then the stores, so the periods overlap their latency.
constexpr int kFixPeriods = 8;

__global__ void __launch_bounds__(128) k_periodic_fix(PeriodicArgs a)
{
    __shared__ double s_dp[4];
    const uint32_t ci = blockIdx.y * blockDim.x + threadIdx.x, kkb = blockIdx.x * (uint32_t)kFixPeriods;
    double dp = 0.0;
    if (ci < min(*a.n_cands, (uint32_t)kPerMaxCands)) {
        const uint2 cand = a.cands[ci];
        const uint32_t m = cand.x, wf = period_warp_first(m, a.tile_len);
        const uint32_t B = (uint32_t)a.B, Lq = a.L / B, Lr = a.L % B;
        const uint64_t g0 = (a.k0 + kkb) * (uint64_t)a.L + m;
        uint32_t b[kFixPeriods], i[kFixPeriods];
        b[0] = (uint32_t)(g0 / B - a.tab_blk0); i[0] = (uint32_t)(g0 % B);
#pragma unroll
        for (int u = 1; u < kFixPeriods; ++u) {
            i[u] = i[u - 1] + Lr; b[u] = b[u - 1] + Lq;
            if (i[u] >= B) { i[u] -= B; ++b[u]; }
        }
        float dre[kFixPeriods], dim[kFixPeriods];
        float2 old[kFixPeriods];
        bool any[kFixPeriods];
#pragma unroll
        for (int u = 0; u < kFixPeriods; ++u) { dre[u] = dim[u] = 0.0f; any[u] = false; }
        for (uint32_t rest = cand.y; rest; rest &= rest - 1u) {
            const uint32_t s = (uint32_t)__ffs((int)rest) - 1u;
            const uint32_t mstar = __ldg(&a.sat[s].mstar), e_ref = __ldg(&a.sat[s].e_ref);
            const float amp = __ldg(&a.sat[s].amp);
            const uint32_t ot = m >= mstar ? 1u : 0u, ow = wf >= mstar ? 1u : 0u;
            const unsigned long long bits = __ldg(reinterpret_cast<const unsigned long long*>(&a.satcode[s].epoch_bits));
            const uint32_t P = __ldg(&a.satcode[s].epoch_period);
            const float Bv = __ldg(a.yb + (size_t)s * a.L + m), Y = __ldg(a.ys + (size_t)s * a.L + m);
            // epoch of the period's slot 0: e_ref + (k - k_ref) mod P, stepped per period
            const long long dk0 = (long long)(a.k0 + kkb) - (long long)a.k_ref;
            uint32_t ek = (uint32_t)(dk0 >= 0 ? (uint64_t)dk0 % P : (P - (uint32_t)((uint64_t)(-dk0) % P)) % P);
            ek = (ek + e_ref) % P;
#pragma unroll
            for (int u = 0; u < kFixPeriods; ++u) {
                const uint32_t eT = ek + ot >= P ? ek + ot - P : ek + ot, eM = ek + ow >= P ? ek + ow - P : ek + ow;
                const uint32_t eP = eT == 0u ? P - 1u : eT - 1u;
                const float sT = ((bits >> eT) & 1ull) ? -1.0f : 1.0f, sP = ((bits >> eP) & 1ull) ? -1.0f : 1.0f;
                const float sM = ((bits >> eM) & 1ull) ? -1.0f : 1.0f;
                ek = ek + 1u == P ? 0u : ek + 1u;
                if (kkb + u >= a.n_periods || (sT == sM && sP == sM)) continue;
                const float coef = (sT * (Y - Bv) + sP * Bv) - sM * Y;
                if (coef == 0.0f) continue;
                const BlockSat* r = a.tab + (size_t)b[u] * a.n_sats + s;
                const uint64_t ph = __ldg(reinterpret_cast<const unsigned long long*>(&r->phi)) +
                                    (uint64_t)(i[u] + 1u) * (uint64_t)__ldg(reinterpret_cast<const long long*>(&r->f));
                float sn, cs;
                accurate_sincos_cycles(ph, &sn, &cs);
                dre[u] = fmaf(coef * amp, cs, dre[u]);
                dim[u] = fmaf(coef * amp, sn, dim[u]);
                any[u] = true;
            }
        }
        float2* o = a.out + ((uint64_t)kkb * a.L + m);
#pragma unroll
        for (int u = 0; u < kFixPeriods; ++u)
            if (any[u]) old[u] = o[(uint64_t)u * a.L];
#pragma unroll
        for (int u = 0; u < kFixPeriods; ++u)
            if (any[u]) {
                const float2 nw = make_float2(old[u].x + dre[u], old[u].y + dim[u]);
                o[(uint64_t)u * a.L] = nw;
                dp += ((double)nw.x * nw.x + (double)nw.y * nw.y) - ((double)old[u].x * old[u].x + (double)old[u].y * old[u].y);
            }
    }
    if (a.power_sum) {
        for (int off = 16; off > 0; off >>= 1) dp += __shfl_xor_sync(0xffffffffu, dp, off);
        if ((threadIdx.x & 31u) == 0) s_dp[threadIdx.x >> 5] = dp;
        __syncthreads();
        if (threadIdx.x == 0) {
            const double t = (s_dp[0] + s_dp[1]) + (s_dp[2] + s_dp[3]);
            if (t != 0.0) atomicAdd(a.power_sum, t);
        }
    }
}

// ----------------------------------------------------------------------------------------------
// launchers (called from synth_host.cu)
void launch_static_check(const BlockSat* tab, const SatConst* sats, uint32_t nblk, uint32_t n_sats, uint64_t B, uint32_t* bad, cudaStream_t st)
{
    const uint64_t total = (uint64_t)nblk * n_sats;
    if (total == 0) return;
    k_static_check<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(tab, sats, nblk, n_sats, B, bad);
    R4WB_LAUNCH_CHECK();
}

void launch_period_tables(const PeriodTableArgs& a, uint32_t ns_padded, cudaStream_t st)
{
    const uint64_t total = (uint64_t)a.n_sats * a.L;
    k_period_tables<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(a);
    R4WB_LAUNCH_CHECK();
    k_period_finish<<<1, 32, 0, st>>>(a, ns_padded);
    R4WB_LAUNCH_CHECK();
    k_period_cands<<<(a.L + 127) / 128, 128, 0, st>>>(a);
    R4WB_LAUNCH_CHECK();
}

void launch_period_phasors(const PeriodicArgs& a, uint32_t ns_padded, uint64_t tab_blk1, float4* T, cudaStream_t st)
{
    const uint64_t total = (uint64_t)a.n_periods * ns_padded;
    if (total == 0) return;
    k_period_phasors<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(a, ns_padded, tab_blk1, T);
    R4WB_LAUNCH_CHECK();
}

template <int NS>
static void launch_periodic_t(const PeriodicArgs& a, cudaStream_t st)
{
    const size_t smem = (size_t)(kPerThreads / 32) * a.KI * NS * sizeof(float4);
    static PerDeviceOnce attr_done;
    if (attr_done.first()) {
        R4WB_CUDA(cudaFuncSetAttribute(k_synth_periodic<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    }
    k_synth_periodic<NS><<<a.n_tiles * a.n_chunks, kPerThreads, smem, st>>>(a);
    R4WB_LAUNCH_CHECK();
}

// smallest instantiated satellite count >= n (0: none)
uint32_t periodic_padded_sats(uint32_t n)
{
    return n <= 1 ? 1u : n <= 2 ? 2u : n <= 4 ? 4u : n <= 8 ? 8u : 0u;
}

void launch_synth_periodic(const PeriodicArgs& a, uint32_t ns_padded, cudaStream_t st)
{
    switch (ns_padded) {
    case 1: launch_periodic_t<1>(a, st); break;
    case 2: launch_periodic_t<2>(a, st); break;
    case 4: launch_periodic_t<4>(a, st); break;
    case 8: launch_periodic_t<8>(a, st); break;
    default: fail(R4WB_ERR_INVALID_PARAMETER, "periodic path: %u satellites", ns_padded);
    }
}

void launch_periodic_fix(const PeriodicArgs& a, uint32_t n_cands, cudaStream_t st)
{
    if (n_cands > 0 && a.n_periods > 0) {
        k_periodic_fix<<<dim3((a.n_periods + kFixPeriods - 1) / kFixPeriods, (n_cands + 127) / 128), 128, 0, st>>>(a);
        R4WB_LAUNCH_CHECK();
    }
}

}  // namespace r4wb
