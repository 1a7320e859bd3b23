// compose.cu — the per-sample loop of r4w-sim's generic scenario engine on the GPU.
//
// Replaces the inner loops of ScenarioEngine::generate_block (crates/r4w-sim/src/scenario/engine.rs:61-137):
//   per emitter: carrier_phase += 2 pi doppler / fs (before use, wrapped `%= 2 pi` once |phase| > 1e6), then
//                composite[i] += sample * (cos phase + j sin phase) * rx_amplitude                       (:105-122)
//   receiver noise: composite[i] += Normal(0, sqrt(noise_power / 2)) per component                      (:125-135)
// The emitters themselves are user objects (trait Emitter: generate_iq / state_at) and the per-block geometry (trajectory,
// range rate, path loss at the block midpoint, :68-101) is host f64 work done by the caller (r4w_b200/sim.py mirrors it);
// this file takes one block's baseband of every active emitter plus its Doppler and amplitude and produces the composite.
// Noise is the library's Philox4x32-10 / Box-Muller stream (the reference draws from rand's ChaCha-based StdRng through
// rand_distr's ziggurat: only statistical parity is possible, as SURVEY.md §8 f5 states).
#include <cuda_runtime.h>

#include <cmath>
#include <vector>

#include "compose.cuh"
#include "synth_math.cuh"

namespace r4wb {


template <typename InT>
__global__ void __launch_bounds__(256) k_compose(const InT* __restrict__ baseband, const ComposeEmitter* __restrict__ em, uint32_t n_emitters,
                                                 uint64_t n, uint64_t sample0, float noise_std, uint64_t seed, int out_f64, void* __restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float re = 0.0f, im = 0.0f;
    for (uint32_t e = 0; e < n_emitters; ++e) {
        const ComposeEmitter p = em[e];
        if (!p.active) continue;
        // phase of sample i: base + (steps since the base) * inc, f64; then reduced to a fraction of a cycle
        const double ph = i >= p.wrap_at ? __dadd_rn(p.phase_wrap, __dmul_rn((double)(i - p.wrap_at), p.inc))
                                         : __dadd_rn(p.phase0, __dmul_rn((double)(i + 1), p.inc));
        const double cyc = ph * 0.15915494309189535;                  // / (2 pi)
        const float fr = (float)(cyc - rint(cyc));                    // [-0.5, 0.5]
        float sn, cs;
        sincospif(2.0f * fr, &sn, &cs);
        const InT s = baseband[(size_t)e * n + i];
        const float sr = (float)s.x, si = (float)s.y;
        re = fmaf(p.amp, sr * cs - si * sn, re);
        im = fmaf(p.amp, sr * sn + si * cs, im);
    }
    if (noise_std > 0.0f) {
        const float2 g = noise_of_sample(sample0 + i, seed);
        re = fmaf(g.x, noise_std, re);
        im = fmaf(g.y, noise_std, im);
    }
    if (out_f64) reinterpret_cast<double2*>(out)[i] = make_double2((double)re, (double)im);
    else reinterpret_cast<float2*>(out)[i] = make_float2(re, im);
}

Composer::Composer(uint32_t n_emitters, double sample_rate, double noise_std, uint64_t seed)
    : fs_(sample_rate), noise_std_(noise_std), seed_(seed)
{
    if (n_emitters > 4096u) fail(R4WB_ERR_INVALID_SIZE, "composer: %u emitters", n_emitters);
    if (!(sample_rate > 0.0) || !(noise_std >= 0.0)) fail(R4WB_ERR_INVALID_PARAMETER, "composer: sample_rate / noise_std");
    phases_.assign(n_emitters, 0.0);
}

void Composer::reset()
{
    std::fill(phases_.begin(), phases_.end(), 0.0);
    current_ = 0;
}

void Composer::block(const void* baseband, r4wb_fmt in_fmt, r4wb_mem in_where, uint64_t n, const double* doppler_hz, const double* amplitude,
                     const uint8_t* active, void* out, r4wb_fmt out_fmt, r4wb_mem out_where)
{
    const uint32_t E = (uint32_t)phases_.size();
    if (n == 0) return;
    if (!out || (E && (!baseband || !doppler_hz || !amplitude))) fail(R4WB_ERR_NULL_POINTER, "composer: NULL argument");
    if ((in_fmt != R4WB_FMT_CF32 && in_fmt != R4WB_FMT_CF64) || (out_fmt != R4WB_FMT_CF32 && out_fmt != R4WB_FMT_CF64))
        fail(R4WB_ERR_INVALID_PARAMETER, "composer: cf32 / cf64 only");
    if (n > 0xffffffffull) fail(R4WB_ERR_INVALID_SIZE, "composer: block of %llu samples", (unsigned long long)n);
    cudaStream_t st = current_stream();
    const size_t in_bps = in_fmt == R4WB_FMT_CF64 ? 16 : 8, out_bps = out_fmt == R4WB_FMT_CF64 ? 16 : 8;
    const double two_pi = 2.0 * 3.14159265358979323846;
    std::vector<ComposeEmitter> em(std::max<uint32_t>(E, 1));
    for (uint32_t e = 0; e < E; ++e) {
        ComposeEmitter& p = em[e];
        p.active = (!active || active[e]) ? 1u : 0u;
        p.phase0 = phases_[e];
        p.inc = two_pi * doppler_hz[e] / fs_;                                    // engine.rs:109
        p.amp = (float)amplitude[e];
        p.wrap_at = 0xffffffffu; p.phase_wrap = 0.0; p.pad = 0;
        if (!p.active) continue;                                                 // inactive emitters keep their phase (:78-80)
        // end-of-block phase, with the reference's wrap (:111-113) at the first sample whose phase exceeds 1e6 in magnitude
        double ph = p.phase0;
        uint64_t done = 0;
        if (p.inc != 0.0 && std::fabs(ph + (double)n * p.inc) > 1.0e6) {
            const double room = 1.0e6 - (p.inc > 0.0 ? ph : -ph);
            uint64_t k = room <= 0.0 ? 1 : (uint64_t)std::floor(room / std::fabs(p.inc)) + 1;    // first index (1-based) past the limit
            while (k > 1 && std::fabs(ph + (double)(k - 1) * p.inc) > 1.0e6) --k;
            while (k <= n && !(std::fabs(ph + (double)k * p.inc) > 1.0e6)) ++k;
            if (k <= n) {
                p.wrap_at = (uint32_t)(k - 1);                                   // sample index k-1 uses the wrapped phase
                p.phase_wrap = std::fmod(ph + (double)k * p.inc, two_pi);
                ph = p.phase_wrap;
                done = k;
            }
        }
        phases_[e] = ph + (double)(n - done) * p.inc;
    }
    const void* d_in = baseband;
    if (E && in_where != R4WB_MEM_DEVICE) {
        d_in_.reserve((size_t)E * n * in_bps);
        R4WB_CUDA(cudaMemcpyAsync(d_in_.p, baseband, (size_t)E * n * in_bps, cudaMemcpyHostToDevice, st));
        d_in = d_in_.p;
    }
    d_em_.reserve(em.size());
    R4WB_CUDA(cudaMemcpyAsync(d_em_.p, em.data(), em.size() * sizeof(ComposeEmitter), cudaMemcpyHostToDevice, st));
    void* d_out = out;
    if (out_where != R4WB_MEM_DEVICE) d_out = d_out_.reserve((size_t)n * out_bps);
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (in_fmt == R4WB_FMT_CF64)
        k_compose<double2><<<grid, 256, 0, st>>>((const double2*)d_in, d_em_.p, E, n, current_, (float)noise_std_, seed_,
                                                 out_fmt == R4WB_FMT_CF64, d_out);
    else
        k_compose<float2><<<grid, 256, 0, st>>>((const float2*)d_in, d_em_.p, E, n, current_, (float)noise_std_, seed_,
                                                out_fmt == R4WB_FMT_CF64, d_out);
    R4WB_LAUNCH_CHECK();
    if (out_where != R4WB_MEM_DEVICE) R4WB_CUDA(cudaMemcpyAsync(out, d_out, (size_t)n * out_bps, cudaMemcpyDeviceToHost, st));
    R4WB_CUDA(cudaStreamSynchronize(st));      // `em` is a host temporary
    current_ += n;
}

}  // namespace r4wb
