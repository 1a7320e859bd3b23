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
// One warp per satellite: exclusive prefix of the approximate block advances (predicts the binade of the phase at each block).
__global__ void __launch_bounds__(32) k_phase_prefix(uint32_t n_sats, uint32_t nblk, const double* __restrict__ papprox, double* __restrict__ pstart)
{
    const uint32_t s = blockIdx.x, lane = threadIdx.x;
    double carry = 0.0;
    for (uint32_t b0 = 0; b0 < nblk; b0 += 32) {
        const uint32_t b = b0 + lane;
        const double v = b < nblk ? papprox[(size_t)b * n_sats + s] : 0.0;
        double x = v;
        for (int off = 1; off < 32; off <<= 1) {
            const double y = __shfl_up_sync(0xffffffffu, x, off);
            if ((int)lane >= off) x += y;
        }
        if (b < nblk) pstart[(size_t)b * n_sats + s] = carry + (x - v);
        carry += __shfl_sync(0xffffffffu, x, 31);
    }
}

// One thread per (block, satellite): the integer sum of the block's increments rounded to the ulp of the predicted binade.
__global__ void k_phase_q(double fs, const SatConst* __restrict__ sats, uint32_t n_sats, uint32_t nblk, const BlockSat* __restrict__ tab,
                          const double* __restrict__ dop, const double* __restrict__ pstart, PhaseQ* __restrict__ out)
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
            if (r.k >= 8) {
                bool tie;
                block_phase_q(ds, de, e.n, fs, r.k, &r.Q, &tie);
                r.ok = tie ? 0u : 1u;
            }
        }
    }
    out[idx] = r;
}

// One warp per dynamic satellite walks the blocks in order with the exact f64 phase: a block whose phase stays inside the
// predicted binade advances by its integer sum, any other block is walked sample by sample (lanes evaluate 32 increments at a
// time, the additions stay sequential).  Writes the phase before the block's first increment into the table.
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

void launch_phase_exact(const ScenConst& sc, const SatConst* d_sats, uint32_t nblk, BlockSat* d_tab, const double* d_dop,
                        const double* d_papprox, double* d_pstart, PhaseQ* d_pq, cudaStream_t st)
{
    if (nblk == 0 || sc.n_sats == 0) return;
    k_phase_prefix<<<sc.n_sats, 32, 0, st>>>(sc.n_sats, nblk, d_papprox, d_pstart);
    R4WB_LAUNCH_CHECK();
    const uint64_t total = (uint64_t)nblk * sc.n_sats;
    k_phase_q<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(sc.fs, d_sats, sc.n_sats, nblk, d_tab, d_dop, d_pstart, d_pq);
    R4WB_LAUNCH_CHECK();
    k_phase_exact<<<sc.n_sats, 32, 0, st>>>(sc.fs, d_sats, sc.n_sats, nblk, d_tab, d_dop, d_pq);
    R4WB_LAUNCH_CHECK();
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

void launch_phase_scan(const SatConst* d_sats, uint32_t n_sats, uint32_t nblk, BlockSat* d_tab, cudaStream_t st)
{
    if (nblk == 0 || n_sats == 0) return;
    k_phase_scan<<<n_sats, 1024, 0, st>>>(d_sats, n_sats, nblk, d_tab);
    R4WB_LAUNCH_CHECK();
}

}  // namespace r4wb
