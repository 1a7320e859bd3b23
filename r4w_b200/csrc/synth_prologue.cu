// synth_prologue.cu — f64 prologue of the synthesis path: Phase 1 of generate_block per (block, satellite)
// (gnss/scenario.rs:378-454, environment/orbit.rs:49-119, core/coordinates.rs:129-238) and the phase scan.
//
// This translation unit is compiled with -fmad=false.  The reference (Rust) never contracts a*b+c into an FMA, and the
// geometry is evaluated at t ~ 1.44e9 s: the mean anomaly M0 + n dt is ~1.8e5 rad, one ulp of it moves the satellite
// 0.9 mm, i.e. ~1e-6 half-chips of code delay.  A contracted FMA there rounds differently from the reference in some
// blocks, and an oversample that lies within that distance of a half-chip boundary then lands on the other side of it
// (seen at 4 / 8 MHz as bursts of 6-7 slightly wrong samples, 3 in 0.16 s of e1c_8prn_60s_cn34_orbital).  Without
// contraction the device evaluates the same IEEE operations in the same order as the oracle and the reference.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>

#include <mutex>

#include "synth_math.cuh"

namespace r4wb {

// dop (nullable): [nblk * n_sats][2] Doppler at the start / end of the block; papprox: [nblk * n_sats] real-number phase advance
// of the block (dynamic, visible entries; 0 otherwise) — inputs of the exact phase model below
__global__ void k_block_params(ScenConst sc, const SatConst* __restrict__ sats, const PhaseSegment* __restrict__ segs,
                               uint64_t blk0, uint32_t nblk, BlockSat* __restrict__ tab, BlockHdr* __restrict__ hdr,
                               double* __restrict__ dop, double* __restrict__ papprox)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)nblk * sc.n_sats) return;
    const uint32_t tb = (uint32_t)(idx / sc.n_sats), s = (uint32_t)(idx % sc.n_sats);
    const uint64_t first = (blk0 + tb) * sc.B;
    const uint64_t rem = sc.total - first;
    const uint32_t n = (uint32_t)(rem < sc.B ? rem : sc.B);
    BlockSat o;
    double d2[2];
    fill_block_sat(sc, sats[s], segs, first, n, /*visible samples so far (constant visibility)*/ first, o, d2);
    if (dop) {
        dop[2 * idx] = d2[0]; dop[2 * idx + 1] = d2[1];
        double a = 0.0, span = 0.0;
        if (!sats[s].static_phase && (o.flags & 1u)) block_phase_approx(d2[0], d2[1], n, sc.fs, &a, &span);
        papprox[idx] = a;
    }
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

// ---- exact carrier phase of dynamic satellites (synth_math.cuh: ref_phase_inc .. phase_after_block) ----------------------
// The reference adds 3e9 f64 increments in order.  A block whose phase provably stays inside one binade advances by an INTEGER
// multiple of that binade's ulp (PhaseQ.Q) — associative — so only the few blocks per satellite that cross a binade, hold an
// exact tie or start near zero ("walk" blocks, on the order of a hundred of 600 000) are sequential:
//   k_phase_prefix   CTA per satellite: scan of the real-number block advances -> predicted start phase / binade of every block
//   k_phase_q        thread per (block, satellite): Q = sum over the block of rint(inc_i / ulp), from the positions of the steps of
//                    that monotone sequence (block_phase_q_steps); walk flag from the prediction
//   k_phase_runs     CTA per satellite: wrapping prefix sum of Q over the non-walk blocks + ordered list of the walk blocks
//   k_phase_chain    warp per satellite: hops from walk block to walk block (run advance = difference of two prefix entries,
//                    exact), walks each walk block sample by sample (the reference's own loop) -> phase at the start of every run
//   k_phase_fill     thread per (block, satellite): phase = run start + (prefix difference) ulp; VERIFIES with that exact phase
//                    that the block really stays inside its binade (the predicate of the serial walk).  Any violation is
//                    counted and the host falls back to the serial kernel (k_phase_exact) — the first violating block of a
//                    satellite has a correct phase by induction, so none can go unnoticed.

// CTA per satellite: exclusive prefix of the approximate block advances
__global__ void __launch_bounds__(1024) k_phase_prefix(uint32_t n_sats, uint32_t nblk, const double* __restrict__ papprox, double* __restrict__ pstart)
{
    __shared__ double s_sum[1024];
    const uint32_t s = blockIdx.x, t = threadIdx.x;
    const uint32_t chunk = (nblk + 1023u) / 1024u;
    const uint32_t lo = min(nblk, t * chunk), hi = min(nblk, lo + chunk);
    double sum = 0.0;
    for (uint32_t b = lo; b < hi; ++b) sum += papprox[(size_t)b * n_sats + s];
    s_sum[t] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        double v = 0.0;
        if ((int)t >= off) v = s_sum[t - off];
        __syncthreads();
        if ((int)t >= off) s_sum[t] += v;
        __syncthreads();
    }
    double run = t > 0 ? s_sum[t - 1] : 0.0;
    for (uint32_t b = lo; b < hi; ++b) {
        pstart[(size_t)b * n_sats + s] = run;
        run += papprox[(size_t)b * n_sats + s];
    }
}

// frac[i] = fl(i / n) for i < n: the first division of ref_phase_inc, shared by every block of n samples
__global__ void k_phase_frac(uint32_t n, double* __restrict__ frac)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) frac[i] = div_rn((double)i, (double)n);
}

// One thread per (block, satellite): the integer sum of the block's increments rounded to the ulp of the predicted binade.
// PhaseQ.ok: bit0 = Q usable while the phase stays in binade k, bit1 = "walk" block (predicted phase does not provably stay inside
// the binade over the block, a tie, or a start below 2^kPhaseMinBinade rad)
// fast_div: a / fs as q = a y, r = fma(-q, fs, a), q' = fma(r, y, q) with y = RN(1 / fs) — the correctly rounded quotient
// (Markstein), three FP64 instructions instead of the ~35 of a division; the host verifies it for the scenario's fs first
constexpr uint32_t kPhaseQMaxSteps = 256;

__global__ void k_phase_q(double fs, double inv_fs, int fast_div, int step_form, const SatConst* __restrict__ sats, uint32_t n_sats, uint32_t nblk,
                          const BlockSat* __restrict__ tab, const double* __restrict__ dop, const double* __restrict__ pstart,
                          const double* __restrict__ frac, uint32_t frac_n, PhaseQ* __restrict__ out)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)nblk * n_sats) return;
    const uint32_t s = (uint32_t)(idx % n_sats);
    const BlockSat& e = tab[idx];
    PhaseQ r;
    r.Q = 0; r.approx = 0.0; r.span = 0.0; r.k = 0; r.ok = 0u;
    if (!sats[s].static_phase && (e.flags & 1u)) {
        const double ds = dop[2 * idx], de = dop[2 * idx + 1];
        block_phase_approx(ds, de, e.n, fs, &r.approx, &r.span);
        const double p = pstart[idx];
        if (p != 0.0) {
            r.k = ilogb(p);
            if (r.k >= kPhaseMinBinade) {
                bool tie = false;
                // Q from the step positions of the monotone sequence rint(inc_i / ulp) (block_phase_q_steps: ~2 + 2 levels exact
                // increment chains instead of n); blocks with more than kPhaseQMaxSteps levels — the first second or two of a
                // file, where the ulp is tiny — are summed sample by sample
                if (!step_form || !block_phase_q_steps(ds, de, e.n, fs, inv_fs, fast_div, r.k, kPhaseQMaxSteps, &r.Q, &tie)) {
                    tie = false;
                    if (e.n == frac_n) {
                        // The same increments as block_phase_q (i / n from the table), rounded to the ulp of binade k:
                        // x = inc 2^(52-k) (exact), y = rint(x) by the 1.5 2^52 trick (|x| < 2^41 here), summed in int64.
                        // rint differs from block_phase_q's floor + (r > 0.5) only at an exact tie, and a block with a tie
                        // is walked sample by sample anyway.
                        long long qd = 0;
                        const double dd = add_rn(de, -ds);
                        const double scale = scalbn(1.0, 52 - r.k), magic = 6755399441055744.0;
                        for (uint32_t i = 0; i < e.n; ++i) {
                            const double dopp = add_rn(ds, mul_rn(frac[i], dd));
                            const double a = mul_rn(6.283185307179586, dopp);
                            double inc;
                            if (fast_div) {
                                const double q0 = mul_rn(a, inv_fs);
                                inc = fma(fma(-q0, fs, a), inv_fs, q0);
                            } else {
                                inc = div_rn(a, fs);
                            }
                            const double x = mul_rn(inc, scale);
                            const double y = add_rn(add_rn(x, magic), -magic);
                            if (fabs(add_rn(x, -y)) == 0.5) tie = true;
                            qd += (long long)y;
                        }
                        r.Q = qd;
                    } else {
                        block_phase_q(ds, de, e.n, fs, r.k, &r.Q, &tie);
                    }
                }
                r.ok = tie ? 0u : 1u;
            }
        }
        if (!phase_stays_in_binade(p, r)) r.ok |= 2u;
    }
    out[idx] = r;
}

// CTA per dynamic satellite: prefQ[b] = wrapping sum of Q over the non-walk blocks before b; walk blocks listed in order
__global__ void __launch_bounds__(1024) k_phase_runs(const SatConst* __restrict__ sats, uint32_t n_sats, uint32_t nblk, const PhaseQ* __restrict__ pq,
                                                      unsigned long long* __restrict__ prefq, uint32_t* __restrict__ wlist, uint32_t* __restrict__ nwalk)
{
    __shared__ unsigned long long s_sum[1024];
    __shared__ uint32_t s_cnt[1024];
    const uint32_t s = blockIdx.x, t = threadIdx.x;
    if (sats[s].static_phase) { if (t == 0) nwalk[s] = 0; return; }
    const uint32_t chunk = (nblk + 1023u) / 1024u;
    const uint32_t lo = min(nblk, t * chunk), hi = min(nblk, lo + chunk);
    unsigned long long sum = 0;
    uint32_t cnt = 0;
    for (uint32_t b = lo; b < hi; ++b) {
        const PhaseQ& r = pq[(size_t)b * n_sats + s];
        if (r.ok & 2u) ++cnt; else sum += (unsigned long long)r.Q;
    }
    s_sum[t] = sum; s_cnt[t] = cnt;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        unsigned long long v = 0; uint32_t c = 0;
        if ((int)t >= off) { v = s_sum[t - off]; c = s_cnt[t - off]; }
        __syncthreads();
        if ((int)t >= off) { s_sum[t] += v; s_cnt[t] += c; }
        __syncthreads();
    }
    unsigned long long run = t > 0 ? s_sum[t - 1] : 0ull;
    uint32_t w = t > 0 ? s_cnt[t - 1] : 0u;
    for (uint32_t b = lo; b < hi; ++b) {
        const PhaseQ& r = pq[(size_t)b * n_sats + s];
        prefq[(size_t)b * n_sats + s] = run;
        if (r.ok & 2u) wlist[(size_t)s * nblk + w++] = b; else run += (unsigned long long)r.Q;
    }
    if (t == 1023) nwalk[s] = s_cnt[1023];
}

// exact advance of a run: (prefix difference) x ulp of the binade of the run's start phase (both multiples of that ulp, the sum
// stays inside the binade: exact).  dq = 0 leaves the phase untouched (runs of invisible blocks, empty runs).
__device__ __forceinline__ double run_phase(double ph0, unsigned long long dq)
{
    if (dq == 0ull) return ph0;
    return ph0 + scalbn((double)(long long)dq, ilogb(ph0) - 52);
}

// One warp per dynamic satellite: from walk block to walk block.  runph[s][i] = exact phase at the first block after walk block
// i - 1 (i = 0: block 0), i.e. at the start of run i, which ends with walk block i (the last run has no walk block).
__global__ void __launch_bounds__(32) k_phase_chain(double fs, const SatConst* __restrict__ sats, uint32_t n_sats, uint32_t nblk,
                                                     const BlockSat* __restrict__ tab, const double* __restrict__ dop,
                                                     const unsigned long long* __restrict__ prefq, const uint32_t* __restrict__ wlist,
                                                     const uint32_t* __restrict__ nwalk, double* __restrict__ runph)
{
    const uint32_t s = blockIdx.x, lane = threadIdx.x;
    if (sats[s].static_phase) return;
    const uint32_t nw = nwalk[s];
    const uint32_t* wl = wlist + (size_t)s * nblk;
    double* rp = runph + (size_t)s * ((size_t)nblk + 1);
    double ph = 0.0;
    uint32_t start = 0;                                       // first block of the current run
    for (uint32_t i = 0; i < nw; ++i) {
        if (lane == 0) rp[i] = ph;
        const uint32_t w = wl[i];
        const size_t idx = (size_t)w * n_sats + s;
        ph = run_phase(ph, prefq[idx] - prefq[(size_t)start * n_sats + s]);
        const double ds = dop[2 * idx], de = dop[2 * idx + 1];
        const uint32_t n = tab[idx].n;
        const double nf = (double)n;
        for (uint32_t i0 = 0; i0 < n; i0 += 32) {
            const uint32_t ii = i0 + lane;
            const double inc = ii < n ? ref_phase_inc(ds, de, ii, nf, fs) : 0.0;       // + 0.0 leaves the phase as it is
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) ph = ph + __shfl_sync(0xffffffffu, inc, jj);
        }
        start = w + 1;
    }
    if (lane == 0) rp[nw] = ph;
}

// One thread per (block, satellite): phase before the block's first increment -> tab[].phi; verification (see above)
__global__ void k_phase_fill(const SatConst* __restrict__ sats, uint32_t n_sats, uint32_t nblk, BlockSat* __restrict__ tab,
                             const PhaseQ* __restrict__ pq, const unsigned long long* __restrict__ prefq, const uint32_t* __restrict__ wlist,
                             const uint32_t* __restrict__ nwalk, const double* __restrict__ runph, uint32_t* __restrict__ bad)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)nblk * n_sats) return;
    const uint32_t s = (uint32_t)(idx % n_sats), b = (uint32_t)(idx / n_sats);
    if (sats[s].static_phase) return;
    const uint32_t nw = nwalk[s];
    const uint32_t* wl = wlist + (size_t)s * nblk;
    uint32_t lo = 0, hi = nw;                                 // i = number of walk blocks before b = index of b's run
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (wl[mid] < b) lo = mid + 1; else hi = mid;
    }
    const uint32_t start = lo > 0 ? wl[lo - 1] + 1 : 0u;
    const double ph = run_phase(runph[(size_t)s * ((size_t)nblk + 1) + lo], prefq[idx] - prefq[(size_t)start * n_sats + s]);
    tab[idx].phi = cycles_to_fixed(ph / (2.0 * kPi));
    const PhaseQ r = pq[idx];
    if ((tab[idx].flags & 1u) && !(r.ok & 2u) && !phase_stays_in_binade(ph, r)) atomicAdd(bad, 1u);
}

// Serial reference implementation of the same result (fallback when k_phase_fill reports a violation; also the A/B hook
// R4WB_PHASE_SERIAL=1): one warp per dynamic satellite walks the blocks in order with the exact f64 phase.
__global__ void __launch_bounds__(32) k_phase_exact(double fs, const SatConst* __restrict__ sats, uint32_t n_sats, uint32_t nblk,
                                                     BlockSat* __restrict__ tab, const double* __restrict__ dop, const PhaseQ* __restrict__ pq)
{
    __shared__ PhaseQ s_rec[32];
    __shared__ double s_ds[32], s_de[32];
    __shared__ uint32_t s_n[32], s_vis[32];
    const uint32_t s = blockIdx.x, lane = threadIdx.x;
    if (sats[s].static_phase) return;
    double ph = 0.0;
    for (uint32_t b0 = 0; b0 < nblk; b0 += 32) {
        const uint32_t b = b0 + lane;
        if (b < nblk) {
            const size_t idx = (size_t)b * n_sats + s;
            s_rec[lane] = pq[idx];
            s_rec[lane].ok &= 1u;
            s_ds[lane] = dop[2 * idx]; s_de[lane] = dop[2 * idx + 1];
            s_n[lane] = tab[idx].n; s_vis[lane] = tab[idx].flags & 1u;
        }
        __syncwarp();
        const uint32_t cnt = min(32u, nblk - b0);
        for (uint32_t j = 0; j < cnt; ++j) {
            if (lane == 0) tab[(size_t)(b0 + j) * n_sats + s].phi = cycles_to_fixed(ph / (2.0 * kPi));
            if (!s_vis[j]) continue;
            const PhaseQ r = s_rec[j];
            if (phase_stays_in_binade(ph, r)) {
                ph = ph + scalbn((double)r.Q, r.k - 52);
            } else {
                const double ds = s_ds[j], de = s_de[j], nf = (double)s_n[j];
                for (uint32_t i0 = 0; i0 < s_n[j]; i0 += 32) {
                    const uint32_t i = i0 + lane;
                    const double inc = i < s_n[j] ? ref_phase_inc(ds, de, i, nf, fs) : 0.0;       // + 0.0 leaves the phase as it is
                    for (int jj = 0; jj < 32; ++jj) ph = ph + __shfl_sync(0xffffffffu, inc, jj);
                }
            }
        }
        __syncwarp();
    }
}


// ---- chunked scans over the blocks of every satellite ---------------------------------------------------------------------------
// The three per-satellite scans of the prologue (phase advance / last visible block, approximate phase, integer run sums) each
// ran as one 1024-thread CTA per satellite: 8 SMs busy, every thread striding through its own 1/1024 of the table.  Here the
// blocks are cut into chunks of 1024: k_chunk_reduce (grid chunks x satellites) -> per-chunk totals, k_chunk_carry (one CTA per
// satellite) -> exclusive carries of the chunks, k_chunk_apply (grid chunks x satellites) -> exclusive scan inside the chunk plus
// the carry.  An Op supplies the value type T, identity(), comb() (associative), load(block, sat), store(block, sat, exclusive
// prefix) and finish(sat, total).  Integer sums wrap identically in any order; the one floating-point scan (approximate phase)
// only predicts binades and is verified by k_phase_fill.
constexpr uint32_t kScanChunk = 1024;

template <class Op>
__global__ void __launch_bounds__(kScanChunk) k_chunk_reduce(Op op, uint32_t n_sats, uint32_t nblk, typename Op::T* __restrict__ partial)
{
    using T = typename Op::T;
    __shared__ T sh[kScanChunk];
    const uint32_t c = blockIdx.x, s = blockIdx.y, t = threadIdx.x;
    const uint32_t b = c * kScanChunk + t;
    sh[t] = b < nblk ? op.load(b, s) : Op::identity();
    __syncthreads();
    for (uint32_t half = kScanChunk / 2; half > 0; half >>= 1) {
        if (t < half) sh[t] = Op::comb(sh[t], sh[t + half]);
        __syncthreads();
    }
    if (t == 0) partial[(size_t)c * n_sats + s] = sh[0];
}

// partial[c][s]: totals in, exclusive carries out
template <class Op>
__global__ void __launch_bounds__(1024) k_chunk_carry(Op op, uint32_t n_sats, uint32_t n_chunks, typename Op::T* __restrict__ partial)
{
    using T = typename Op::T;
    __shared__ T sh[1024];
    const uint32_t s = blockIdx.x, t = threadIdx.x;
    const uint32_t per = (n_chunks + 1023u) / 1024u;
    const uint32_t lo = min(n_chunks, t * per), hi = min(n_chunks, lo + per);
    T sum = Op::identity();
    for (uint32_t c = lo; c < hi; ++c) sum = Op::comb(sum, partial[(size_t)c * n_sats + s]);
    sh[t] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        T v = Op::identity();
        if ((int)t >= off) v = sh[t - off];
        __syncthreads();
        if ((int)t >= off) sh[t] = Op::comb(v, sh[t]);
        __syncthreads();
    }
    T run = t > 0 ? sh[t - 1] : Op::identity();
    for (uint32_t c = lo; c < hi; ++c) {
        const T v = partial[(size_t)c * n_sats + s];
        partial[(size_t)c * n_sats + s] = run;
        run = Op::comb(run, v);
    }
    if (t == 1023) op.finish(s, sh[1023]);
}

template <class Op>
__global__ void __launch_bounds__(kScanChunk) k_chunk_apply(Op op, uint32_t n_sats, uint32_t nblk, const typename Op::T* __restrict__ carry)
{
    using T = typename Op::T;
    __shared__ T sh[kScanChunk];
    const uint32_t c = blockIdx.x, s = blockIdx.y, t = threadIdx.x;
    const uint32_t b = c * kScanChunk + t;
    const T mine = b < nblk ? op.load(b, s) : Op::identity();
    sh[t] = mine;
    __syncthreads();
    for (int off = 1; off < (int)kScanChunk; off <<= 1) {
        T v = Op::identity();
        if ((int)t >= off) v = sh[t - off];
        __syncthreads();
        if ((int)t >= off) sh[t] = Op::comb(v, sh[t]);
        __syncthreads();
    }
    if (b >= nblk) return;
    T ex = carry[(size_t)c * n_sats + s];
    if (t > 0) ex = Op::comb(ex, sh[t - 1]);
    op.store(b, s, ex, mine);
}

template <class Op>
static void chunk_scan(const Op& op, uint32_t n_sats, uint32_t nblk, void* scratch, cudaStream_t st)
{
    using T = typename Op::T;
    if (nblk == 0 || n_sats == 0) return;
    const uint32_t n_chunks = (nblk + kScanChunk - 1) / kScanChunk;
    T* partial = reinterpret_cast<T*>(scratch);
    k_chunk_reduce<Op><<<dim3(n_chunks, n_sats), kScanChunk, 0, st>>>(op, n_sats, nblk, partial);
    R4WB_LAUNCH_CHECK();
    k_chunk_carry<Op><<<n_sats, 1024, 0, st>>>(op, n_sats, n_chunks, partial);
    R4WB_LAUNCH_CHECK();
    k_chunk_apply<Op><<<dim3(n_chunks, n_sats), kScanChunk, 0, st>>>(op, n_sats, nblk, partial);
    R4WB_LAUNCH_CHECK();
}

// bytes chunk_scan needs for any of the Ops below (T is at most 16 bytes)
size_t scan_scratch_bytes(uint32_t n_sats, uint32_t nblk)
{
    const size_t n_chunks = ((size_t)nblk + kScanChunk - 1) / kScanChunk;
    return ((n_chunks * std::max(1u, n_sats) * 16) + 255) & ~(size_t)255;
}

// k_phase_scan as an Op: exclusive sum of the block advances (dynamic satellites) and the last visible block before each block
struct ScanAdvanceOp {
    struct T { uint64_t sum; int last; int pad; };
    const SatConst* sats; BlockSat* tab; uint32_t n_sats;
    __device__ static T identity() { return T{0ull, -1, 0}; }
    __device__ static T comb(const T& a, const T& b) { return T{a.sum + b.sum, max(a.last, b.last), 0}; }
    __device__ T load(uint32_t b, uint32_t s) const
    {
        const BlockSat& e = tab[(size_t)b * n_sats + s];
        return T{sats[s].static_phase ? 0ull : e.phi, (e.flags & 1u) ? (int)b : -1, 0};
    }
    __device__ void store(uint32_t b, uint32_t s, const T& ex, const T&) const
    {
        BlockSat& e = tab[(size_t)b * n_sats + s];
        if (!sats[s].static_phase) e.phi = ex.sum;
        e.prev = ex.last >= 0 ? (int32_t)((uint32_t)ex.last * n_sats + s) : -1;
    }
    __device__ void finish(uint32_t, const T&) const {}
};

// k_phase_prefix as an Op: predicted (real-number) phase at the start of every block
struct ScanApproxOp {
    using T = double;
    const double* papprox; double* pstart; uint32_t n_sats;
    __device__ static T identity() { return 0.0; }
    __device__ static T comb(const T& a, const T& b) { return a + b; }
    __device__ T load(uint32_t b, uint32_t s) const { return papprox[(size_t)b * n_sats + s]; }
    __device__ void store(uint32_t b, uint32_t s, const T& ex, const T&) const { pstart[(size_t)b * n_sats + s] = ex; }
    __device__ void finish(uint32_t, const T&) const {}
};

// k_phase_runs as an Op: wrapping sum of Q over the non-walk blocks before each block, ordered list of the walk blocks
struct ScanRunsOp {
    struct T { unsigned long long sum; uint32_t cnt; uint32_t pad; };
    const SatConst* sats; const PhaseQ* pq; unsigned long long* prefq; uint32_t* wlist; uint32_t* nwalk; uint32_t n_sats, nblk;
    __device__ static T identity() { return T{0ull, 0u, 0u}; }
    __device__ static T comb(const T& a, const T& b) { return T{a.sum + b.sum, a.cnt + b.cnt, 0u}; }
    __device__ T load(uint32_t b, uint32_t s) const
    {
        if (sats[s].static_phase) return identity();
        const PhaseQ& r = pq[(size_t)b * n_sats + s];
        return (r.ok & 2u) ? T{0ull, 1u, 0u} : T{(unsigned long long)r.Q, 0u, 0u};
    }
    __device__ void store(uint32_t b, uint32_t s, const T& ex, const T& mine) const
    {
        if (sats[s].static_phase) return;
        prefq[(size_t)b * n_sats + s] = ex.sum;
        if (mine.cnt) wlist[(size_t)s * nblk + ex.cnt] = b;
    }
    __device__ void finish(uint32_t s, const T& total) const { nwalk[s] = total.cnt; }
};

static bool scan_per_satellite()
{
    static const bool v = [] { const char* e = std::getenv("R4WB_SCAN_PER_SAT"); return e && e[0] == '1'; }();   // A/B hook: round 2's first scans
    return v;
}

// scratch of the exact-phase pass (caller-owned, sized by phase_exact_scratch_bytes)
struct PhaseScratch {
    double* pstart; PhaseQ* pq; double* frac; unsigned long long* prefq; uint32_t* wlist; uint32_t* nwalk; double* runph; uint32_t* bad;
};

size_t phase_exact_scratch_bytes(uint32_t n_sats, uint32_t nblk, uint32_t B)
{
    const size_t ne = (size_t)nblk * n_sats;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    return al(ne * 8) + al(ne * sizeof(PhaseQ)) + al((size_t)B * 8) + al(ne * 8) + al(ne * 4) + al((size_t)n_sats * 4 + 4) + al((ne + n_sats) * 8) + 256 +
           scan_scratch_bytes(n_sats, nblk);
}

// returns false when the parallel pass found a violation of its own premise and the serial kernel was run instead
bool launch_phase_exact(const ScenConst& sc, const SatConst* d_sats, uint32_t nblk, BlockSat* d_tab, const double* d_dop,
                        const double* d_papprox, unsigned char* scratch, cudaStream_t st)
{
    if (nblk == 0 || sc.n_sats == 0) return true;
    const uint32_t ns = sc.n_sats;
    const size_t ne = (size_t)nblk * ns;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    PhaseScratch S;
    unsigned char* p = scratch;
    S.pstart = reinterpret_cast<double*>(p); p += al(ne * 8);
    S.pq = reinterpret_cast<PhaseQ*>(p); p += al(ne * sizeof(PhaseQ));
    S.frac = reinterpret_cast<double*>(p); p += al((size_t)sc.B * 8);
    S.prefq = reinterpret_cast<unsigned long long*>(p); p += al(ne * 8);
    S.wlist = reinterpret_cast<uint32_t*>(p); p += al(ne * 4);
    S.nwalk = reinterpret_cast<uint32_t*>(p); p += al((size_t)ns * 4 + 4);
    S.runph = reinterpret_cast<double*>(p); p += al((ne + ns) * 8);
    S.bad = reinterpret_cast<uint32_t*>(p); p += 256;
    void* scan_scratch = p;

    if (scan_per_satellite()) {
        k_phase_prefix<<<ns, 1024, 0, st>>>(ns, nblk, d_papprox, S.pstart);
        R4WB_LAUNCH_CHECK();
    } else {
        chunk_scan(ScanApproxOp{d_papprox, S.pstart, ns}, ns, nblk, scan_scratch, st);
    }
    k_phase_frac<<<(unsigned)((sc.B + 255) / 256), 256, 0, st>>>((uint32_t)sc.B, S.frac);
    R4WB_LAUNCH_CHECK();
    // exact-division shortcut, checked on the host for this sample rate over the magnitudes 2 pi |Doppler| can take
    static std::mutex checked_mu;                     // tables are built from several host threads (r4wb_init_devices)
    static double checked_fs = 0.0;
    static int checked_ok = 0;
    std::unique_lock<std::mutex> checked_lock(checked_mu);
    if (checked_fs != sc.fs) {
        const double y = 1.0 / sc.fs;
        uint64_t s = 88172645463325252ull;
        int ok = 1;
        for (int i = 0; i < 200000 && ok; ++i) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            double a = ldexp(1.0 + (double)(s >> 12) * 2.220446049250313e-16, (int)((s >> 3) % 40) - 16);
            if (s & 1) a = -a;
            const double q0 = a * y;
            if (fma(fma(-q0, sc.fs, a), y, q0) != a / sc.fs) ok = 0;
        }
        checked_ok = ok; checked_fs = sc.fs;
    }
    const int fast_div = checked_ok;
    checked_lock.unlock();
    // A/B hook: R4WB_PHASE_Q_BRUTE=1 sums every block's increments sample by sample instead of locating the steps
    static const int step_form = [] { const char* e = std::getenv("R4WB_PHASE_Q_BRUTE"); return (e && e[0] == '1') ? 0 : 1; }();
    k_phase_q<<<(unsigned)((ne + 127) / 128), 128, 0, st>>>(sc.fs, 1.0 / sc.fs, fast_div, step_form, d_sats, ns, nblk, d_tab, d_dop, S.pstart, S.frac,
                                                             (uint32_t)sc.B, S.pq);
    R4WB_LAUNCH_CHECK();
    static const bool serial = [] { const char* e = std::getenv("R4WB_PHASE_SERIAL"); return e && e[0] == '1'; }();
    uint32_t bad = 0;
    if (!serial) {
        R4WB_CUDA(cudaMemsetAsync(S.bad, 0, sizeof(uint32_t), st));
        if (scan_per_satellite()) {
            k_phase_runs<<<ns, 1024, 0, st>>>(d_sats, ns, nblk, S.pq, S.prefq, S.wlist, S.nwalk);
            R4WB_LAUNCH_CHECK();
        } else {
            chunk_scan(ScanRunsOp{d_sats, S.pq, S.prefq, S.wlist, S.nwalk, ns, nblk}, ns, nblk, scan_scratch, st);
        }
        k_phase_chain<<<ns, 32, 0, st>>>(sc.fs, d_sats, ns, nblk, d_tab, d_dop, S.prefq, S.wlist, S.nwalk, S.runph);
        R4WB_LAUNCH_CHECK();
        k_phase_fill<<<(unsigned)((ne + 127) / 128), 128, 0, st>>>(d_sats, ns, nblk, d_tab, S.pq, S.prefq, S.wlist, S.nwalk, S.runph, S.bad);
        R4WB_LAUNCH_CHECK();
        R4WB_CUDA(cudaMemcpyAsync(&bad, S.bad, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        R4WB_CUDA(cudaStreamSynchronize(st));
        if (bad == 0) return true;
    }
    k_phase_exact<<<ns, 32, 0, st>>>(sc.fs, d_sats, ns, nblk, d_tab, d_dop, S.pq);
    R4WB_LAUNCH_CHECK();
    R4WB_CUDA(cudaStreamSynchronize(st));
    return serial;
}

void launch_block_params(const ScenConst& sc, const SatConst* d_sats, const PhaseSegment* d_segs, uint64_t blk0, uint32_t nblk,
                         BlockSat* d_tab, BlockHdr* d_hdr, double* d_dop, double* d_papprox, cudaStream_t st)
{
    const uint64_t total = (uint64_t)nblk * sc.n_sats;
    if (total == 0) return;
    const int threads = 128;
    k_block_params<<<(unsigned)((total + threads - 1) / threads), threads, 0, st>>>(sc, d_sats, d_segs, blk0, nblk, d_tab, d_hdr, d_dop, d_papprox);
    R4WB_LAUNCH_CHECK();
}

// scratch: scan_scratch_bytes(n_sats, nblk) of device memory
void launch_phase_scan(const SatConst* d_sats, uint32_t n_sats, uint32_t nblk, BlockSat* d_tab, void* scratch, cudaStream_t st)
{
    if (nblk == 0 || n_sats == 0) return;
    if (scan_per_satellite()) {
        k_phase_scan<<<n_sats, 1024, 0, st>>>(d_sats, n_sats, nblk, d_tab);
        R4WB_LAUNCH_CHECK();
    } else {
        chunk_scan(ScanAdvanceOp{d_sats, d_tab, n_sats}, n_sats, nblk, scratch, st);
    }
}

}  // namespace r4wb
