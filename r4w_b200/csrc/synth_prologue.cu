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
