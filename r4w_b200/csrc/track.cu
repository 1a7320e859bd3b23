// track.cu — bank of DLL/PLL tracking channels on the GPU.
//
// Replaces TrackingChannel (crates/r4w-core/src/waveform/gnss/tracking.rs):
//   new / with_dll_bandwidth / with_pll_bandwidth   :107-167
//   process (E/P/L correlators, discriminators, loop filters, NCO updates, C/N0, lock, nav bits)   :177-313
//   estimate_cn0 :316-337, LoopFilter2nd :365-397, LoopFilter3rd :399-437
// The reference runs one channel at a time, one code period per call; the loop feedback makes the periods of a channel
// sequential, the channels are independent.  k_track: one thread-block cluster (1-8 CTAs) per channel walks its periods in
// order; inside a period the cluster's threads split the samples (carrier wipe-off + the three code look-ups in f64 with the reference's own
// expressions, never contracted into FMAs, so the chip indices are the reference's), a fixed-order tree reduces the six
// correlator sums, thread 0 runs the loop update exactly as :219-313 and publishes the state for the next period.
// Only the summation order differs from the reference (pairwise instead of sequential): states agree to ~1e-12.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "track.cuh"

namespace cg = cooperative_groups;

namespace r4wb {

constexpr int kTrackThreads = 512;
constexpr double kTwoPi = 6.283185307179586476925286766559;

__device__ __forceinline__ double rem_euclid_d(double a, double b)
{
    // f64::rem_euclid: r = a % b (fmod), r < 0 -> r + |b|.  For 0 <= a < 2b (every chip position of a code period) fmod is
    // a or a - b, and that subtraction is exact (Sterbenz), so the shortcut is bit-identical to the library routine.
    double r;
    if (a >= 0.0 && a < b) r = a;
    else if (a >= b && a < __dadd_rn(b, b)) r = __dadd_rn(a, -b);
    else r = fmod(a, b);
    return r < 0.0 ? __dadd_rn(r, fabs(b)) : r;
}
__device__ __forceinline__ uint32_t as_index(double x, uint32_t code_length)
{
    const unsigned long long v = x > 0.0 ? (unsigned long long)x : 0ull;     // Rust `as usize`: truncation, negative -> 0
    return (uint32_t)(v % code_length);
}

__device__ double track_estimate_cn0(const TrackChan& c)      // tracking.rs:316-337
{
    if (c.cn0_n < 2u) return 0.0;
    const double n = (double)c.cn0_n;
    double sum = 0.0;
    for (uint32_t i = 0; i < c.cn0_n; ++i) sum = __dadd_rn(sum, c.cn0_buf[i]);
    const double mean = sum / n;
    double vs = 0.0;
    for (uint32_t i = 0; i < c.cn0_n; ++i) { const double d = __dadd_rn(c.cn0_buf[i], -mean); vs = __dadd_rn(vs, __dmul_rn(d, d)); }
    const double variance = vs / (n - 1.0);
    if (variance <= 0.0 || mean <= 0.0) return 0.0;
    const double snr = __dmul_rn(mean, mean) / variance;
    const double x = snr - 1.0;
    const double lin = __dmul_rn(1.0 / 0.001, x > 0.01 ? x : 0.01);
    return __dmul_rn(10.0, log10(lin));
}

template <typename SampleT>
__global__ void __launch_bounds__(kTrackThreads) k_track(TrackChan* __restrict__ chans, const SampleT* __restrict__ samples, uint64_t n_per_period,
                                                          uint64_t n_periods, uint64_t channel_stride, const int8_t* __restrict__ codes,
                                                          uint64_t code_stride, r4wb_track_state* __restrict__ out, uint32_t n_channels,
                                                          int8_t* __restrict__ nav_out, uint32_t nav_cap, uint32_t* __restrict__ nav_n)
{
    // A channel is one thread-block CLUSTER (1, 2, 4 or 8 CTAs, chosen by the host from the bank size): the CTAs split the
    // period's samples, leave their six partial sums in their own shared memory, and after one cluster barrier every CTA
    // reads all partials through distributed shared memory in rank order and runs the (identical) loop update itself, so
    // nothing has to be broadcast back.  The partial slots are double-buffered: one cluster barrier per period.
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t cs = cluster.num_blocks(), rank = cluster.block_rank();
    extern __shared__ int8_t s_code[];
    __shared__ double s_red[kTrackThreads / 32][6];
    __shared__ double s_part[2][6];
    __shared__ TrackChan s_c;
    const uint32_t ch = blockIdx.x / cs, tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_c = chans[ch];
    __syncthreads();
    const uint32_t code_length = s_c.code_length;
    for (uint32_t k = tid; k < code_length; k += kTrackThreads) s_code[k] = codes[(size_t)ch * code_stride + k];
    const SampleT* x = samples + (size_t)ch * channel_stride;
    uint32_t nav_count = 0;
    __syncthreads();

    for (uint64_t p = 0; p < n_periods; ++p) {
        const double fs = s_c.sample_rate, cl = (double)code_length;
        const double spc = fs / s_c.code_freq;                                   // :178
        const double cf = s_c.carrier_freq, cph = s_c.carrier_phase, cp = s_c.code_phase, half = s_c.el_spacing / 2.0;
        double ei = 0.0, eq = 0.0, pi = 0.0, pq = 0.0, li = 0.0, lq = 0.0;
        const SampleT* xp = x + p * n_per_period;
        for (uint64_t i = (uint64_t)rank * kTrackThreads + tid; i < n_per_period; i += (uint64_t)kTrackThreads * cs) {   // :186-217
            const double t = (double)i / fs;
            const double arg = __dmul_rn(__dmul_rn(-2.0, 3.14159265358979323846), __dadd_rn(__dmul_rn(cf, t), cph));
            double sn, cs_;
            sincos(arg, &sn, &cs_);
            const double re = (double)xp[i].x, im = (double)xp[i].y;
            const double sre = __dadd_rn(__dmul_rn(re, cs_), -__dmul_rn(im, sn));
            const double sim = __dadd_rn(__dmul_rn(re, sn), __dmul_rn(im, cs_));
            const double chip = __dadd_rn(cp, (double)i / spc);
            const double ec = (double)s_code[as_index(rem_euclid_d(__dadd_rn(chip, -half), cl), code_length)];
            const double pc = (double)s_code[as_index(rem_euclid_d(chip, cl), code_length)];
            const double lc = (double)s_code[as_index(rem_euclid_d(__dadd_rn(chip, half), cl), code_length)];
            ei += sre * ec; eq += sim * ec;          // +-1 codes: the products are exact
            pi += sre * pc; pq += sim * pc;
            li += sre * lc; lq += sim * lc;
        }
        double v[6] = {ei, eq, pi, pq, li, lq};
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            for (int off = 16; off > 0; off >>= 1) v[j] += __shfl_xor_sync(0xffffffffu, v[j], off);
            if (lane == 0) s_red[warp][j] = v[j];
        }
        __syncthreads();
        if (cs > 1) {
            if (tid < 6) {
                double a = 0.0;
                for (int w = 0; w < kTrackThreads / 32; ++w) a += s_red[w][tid];
                s_part[p & 1][tid] = a;
            }
            cluster.sync();                           // every CTA's partial sums of this period are in place
        }
        if (tid == 0) {
            TrackChan& c = s_c;
            double s[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            if (cs > 1) {
                for (uint32_t r = 0; r < cs; ++r) {   // rank order: the same totals in every CTA
                    const double* part = cluster.map_shared_rank(&s_part[p & 1][0], r);
                    for (int j = 0; j < 6; ++j) s[j] += part[j];
                }
            } else {
                for (int j = 0; j < 6; ++j)
                    for (int w = 0; w < kTrackThreads / 32; ++w) s[j] += s_red[w][j];
            }
            c.p_i = s[2]; c.p_q = s[3];
            const double early_power = sqrt(__dadd_rn(__dmul_rn(s[0], s[0]), __dmul_rn(s[1], s[1])));
            const double late_power = sqrt(__dadd_rn(__dmul_rn(s[4], s[4]), __dmul_rn(s[5], s[5])));
            const double el = __dadd_rn(early_power, late_power);
            const double dll_disc = el > 0.0 ? __dadd_rn(early_power, -late_power) / el : 0.0;
            const double pll_disc = fabs(c.p_i) > 1e-10 ? atan2(c.p_q, c.p_i) / kTwoPi : 0.0;
            const double fll_disc = pll_disc;
            // LoopFilter2nd::update :389-392
            c.dll_int = __dadd_rn(c.dll_int, __dmul_rn(c.dll_k2, dll_disc));
            const double code_correction = __dadd_rn(__dmul_rn(c.dll_k1, dll_disc), c.dll_int);
            // LoopFilter3rd::update :426-430
            c.pll_i1 = __dadd_rn(c.pll_i1, __dmul_rn(c.pll_k2, pll_disc));
            c.pll_i2 = __dadd_rn(c.pll_i2, __dmul_rn(c.pll_k3, pll_disc));
            const double pll_out = __dadd_rn(__dadd_rn(__dmul_rn(c.pll_k1, pll_disc), c.pll_i1), c.pll_i2);
            double carrier_correction;
            if (c.fll_assist && c.ms_count < 100ull)
                carrier_correction = __dadd_rn(__dmul_rn(pll_out, 0.5), __dmul_rn(__dmul_rn(fll_disc, c.fll_bandwidth), 0.5));
            else
                carrier_correction = pll_out;
            c.code_phase = rem_euclid_d(__dadd_rn(c.code_phase, __dmul_rn(code_correction, c.el_spacing)), cl);
            c.carrier_freq = __dadd_rn(c.carrier_freq, carrier_correction);
            c.carrier_phase = rem_euclid_d(__dadd_rn(c.carrier_phase, c.carrier_freq / c.sample_rate), 1.0);
            const double code_doppler = __dmul_rn(c.carrier_freq, c.chipping_rate) / 1575420000.0;
            c.code_freq = __dadd_rn(c.chipping_rate, code_doppler);
            const double prompt_power = __dadd_rn(__dmul_rn(c.p_i, c.p_i), __dmul_rn(c.p_q, c.p_q));
            if (c.cn0_n == 20u) {
                for (int j = 0; j < 19; ++j) c.cn0_buf[j] = c.cn0_buf[j + 1];
                c.cn0_n = 19u;
            }
            c.cn0_buf[c.cn0_n++] = prompt_power;
            const double cn0 = track_estimate_cn0(c);
            c.carrier_lock = (cn0 > 25.0 && c.ms_count > 10ull) ? 1u : 0u;
            c.code_lock = cn0 > 20.0 ? 1u : 0u;
            c.nav_acc = __dadd_rn(c.nav_acc, c.p_i);
            c.nav_bit_count += 1u;
            const int sign = c.p_i >= 0.0 ? 1 : -1;
            if (c.prev_sign != 0 && sign != c.prev_sign && !c.bit_sync && c.ms_count > 20ull) c.bit_sync = 1u;
            c.prev_sign = sign;
            if (c.nav_bit_count >= 20u) {
                if (rank == 0 && nav_count < nav_cap) nav_out[(size_t)ch * nav_cap + nav_count] = c.nav_acc >= 0.0 ? 1 : -1;
                ++nav_count;
                c.nav_acc = 0.0;
                c.nav_bit_count = 0u;
            }
            c.ms_count += 1ull;
            if (c.ms_count > 200ull) c.fll_assist = 0u;
            r4wb_track_state o;
            o.code_phase = c.code_phase; o.carrier_freq_hz = c.carrier_freq; o.carrier_phase_rad = __dmul_rn(__dmul_rn(c.carrier_phase, 2.0), 3.14159265358979323846);
            o.prompt_i = c.p_i; o.prompt_q = c.p_q; o.cn0_dbhz = cn0; o.ms_count = c.ms_count;
            o.prn = (uint8_t)c.prn; o.carrier_lock = (uint8_t)c.carrier_lock; o.code_lock = (uint8_t)c.code_lock; o.bit_sync = (uint8_t)c.bit_sync;
            o.pad[0] = o.pad[1] = o.pad[2] = o.pad[3] = 0;
            if (rank == 0) out[p * n_channels + ch] = o;
        }
        __syncthreads();
    }
    if (cs > 1) cluster.sync();                       // no CTA leaves while another may still read its partial sums
    if (tid == 0 && rank == 0) {
        chans[ch] = s_c;
        nav_n[ch] = nav_count;
    }
}

// ----------------------------------------------------------------------------------------------
static void loop2_gains(double bw, double T, double& k1, double& k2)      // LoopFilter2nd::new :375-385
{
    const double omega_n = bw * 8.0 / 3.0, zeta = 1.0 / std::sqrt(2.0);
    k1 = 2.0 * zeta * omega_n * T;
    k2 = (omega_n * omega_n) * (T * T);
}
static void loop3_gains(double bw, double T, double& k1, double& k2, double& k3)   // LoopFilter3rd::new :412-424
{
    const double omega_n = bw * 2.4, a3 = 1.1, b3 = 2.4;
    k1 = b3 * omega_n * T;
    k2 = a3 * (omega_n * omega_n) * (T * T);
    k3 = (omega_n * omega_n * omega_n) * (T * T * T);
}

TrackerBank::TrackerBank(const r4wb_track_cfg* cfgs, uint32_t n)
{
    if (n == 0 || n > 65535u) fail(R4WB_ERR_INVALID_SIZE, "tracker bank: %u channels", n);
    host_.resize(n);
    nav_.resize(n);
    for (uint32_t k = 0; k < n; ++k) {
        const r4wb_track_cfg& c = cfgs[k];
        if (c.code_length == 0 || c.code_length > 65536ull) fail(R4WB_ERR_INVALID_SIZE, "channel %u: code_length %llu", k, (unsigned long long)c.code_length);
        if (!(c.sample_rate > 0.0) || !(c.chipping_rate > 0.0)) fail(R4WB_ERR_INVALID_PARAMETER, "channel %u: rates must be positive", k);
        TrackChan t{};
        t.prn = c.prn; t.code_length = (uint32_t)c.code_length; t.sample_rate = c.sample_rate; t.chipping_rate = c.chipping_rate;
        const double code_doppler = c.initial_doppler * c.chipping_rate / 1575420000.0;            // :122
        t.code_phase = c.initial_code_phase; t.code_freq = c.chipping_rate + code_doppler; t.el_spacing = 0.5;
        loop2_gains(c.dll_bandwidth_hz > 0.0 ? c.dll_bandwidth_hz : 1.0, 0.001, t.dll_k1, t.dll_k2);
        loop3_gains(c.pll_bandwidth_hz > 0.0 ? c.pll_bandwidth_hz : 15.0, 0.001, t.pll_k1, t.pll_k2, t.pll_k3);
        t.carrier_phase = 0.0; t.carrier_freq = c.initial_doppler; t.fll_bandwidth = 50.0; t.fll_assist = 1u;
        host_[k] = t;
    }
    d_chan_.reserve(n);
    cudaStream_t st = current_stream();
    R4WB_CUDA(cudaMemcpyAsync(d_chan_.p, host_.data(), n * sizeof(TrackChan), cudaMemcpyHostToDevice, st));
    R4WB_CUDA(cudaStreamSynchronize(st));
}

void TrackerBank::process(const void* samples, r4wb_fmt fmt, r4wb_mem where, uint64_t n_per_period, uint64_t n_periods, uint64_t channel_stride,
                          const int8_t* codes, uint64_t code_stride, r4wb_track_state* out)
{
    const uint32_t n = (uint32_t)host_.size();
    if (n_periods == 0) return;
    if (!samples || !codes || !out) fail(R4WB_ERR_NULL_POINTER, "samples/codes/out is NULL");
    if (fmt != R4WB_FMT_CF32 && fmt != R4WB_FMT_CF64) fail(R4WB_ERR_INVALID_PARAMETER, "tracking input must be cf32 or cf64");
    if (n_per_period == 0 || n_per_period > (1ull << 24)) fail(R4WB_ERR_INVALID_SIZE, "samples per period %llu", (unsigned long long)n_per_period);
    uint32_t max_cl = 0;
    for (const TrackChan& c : host_) max_cl = std::max(max_cl, c.code_length);
    if (code_stride < max_cl) fail(R4WB_ERR_INVALID_SIZE, "code_stride %llu < code_length %u", (unsigned long long)code_stride, max_cl);
    cudaStream_t st = current_stream();
    const size_t bps = fmt == R4WB_FMT_CF64 ? 16 : 8;
    const uint64_t span = n_per_period * n_periods;                               // samples per channel
    const void* d_x = samples;
    if (where != R4WB_MEM_DEVICE) {
        const uint64_t total = channel_stride ? channel_stride * (n - 1) + span : span;
        d_in_.reserve((size_t)total * bps);
        R4WB_CUDA(cudaMemcpyAsync(d_in_.p, samples, (size_t)total * bps, cudaMemcpyHostToDevice, st));
        d_x = d_in_.p;
    }
    d_codes_.reserve((size_t)n * code_stride);
    R4WB_CUDA(cudaMemcpyAsync(d_codes_.p, codes, (size_t)n * code_stride, cudaMemcpyHostToDevice, st));
    d_out_.reserve((size_t)n_periods * n);
    const uint32_t nav_cap = (uint32_t)(n_periods / 20 + 2);
    d_nav_.reserve((size_t)n * nav_cap);
    d_nav_n_.reserve(n);
    const size_t smem = (max_cl + 15u) & ~15u;
    {   // codes above ~47 K chips need the opt-in dynamic shared-memory limit (code_length is accepted up to 65 536)
        static PerDeviceOnce once;
        if (once.first()) {
            R4WB_CUDA(cudaFuncSetAttribute(k_track<float2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
            R4WB_CUDA(cudaFuncSetAttribute(k_track<double2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
        }
    }
    // cluster size: as many CTAs per channel as keep the whole bank resident (148 SMs), up to the portable maximum of 8
    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        R4WB_CUDA(cudaGetDevice(&dev));
        R4WB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    unsigned cs = 1;
    while (cs < 8 && (uint64_t)n * cs * 2 <= (uint64_t)sm_count && (uint64_t)kTrackThreads * cs * 2 <= n_per_period) cs *= 2;
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(n * cs); lc.blockDim = dim3(kTrackThreads); lc.dynamicSmemBytes = smem; lc.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    lc.attrs = attr; lc.numAttrs = 1;
    const uint32_t nch = n;
    if (cs == 1 && fmt == R4WB_FMT_CF32)             // plain launch: a cluster attribute of 1 costs occupancy on a full bank
        k_track<float2><<<n, kTrackThreads, smem, st>>>(d_chan_.p, (const float2*)d_x, n_per_period, n_periods, channel_stride, d_codes_.p,
                                                        code_stride, d_out_.p, nch, d_nav_.p, nav_cap, d_nav_n_.p);
    else if (cs == 1)
        k_track<double2><<<n, kTrackThreads, smem, st>>>(d_chan_.p, (const double2*)d_x, n_per_period, n_periods, channel_stride, d_codes_.p,
                                                         code_stride, d_out_.p, nch, d_nav_.p, nav_cap, d_nav_n_.p);
    else if (fmt == R4WB_FMT_CF32)
        R4WB_CUDA(cudaLaunchKernelEx(&lc, k_track<float2>, d_chan_.p, (const float2*)d_x, n_per_period, n_periods, channel_stride,
                                     (const int8_t*)d_codes_.p, code_stride, d_out_.p, nch, d_nav_.p, nav_cap, d_nav_n_.p));
    else
        R4WB_CUDA(cudaLaunchKernelEx(&lc, k_track<double2>, d_chan_.p, (const double2*)d_x, n_per_period, n_periods, channel_stride,
                                     (const int8_t*)d_codes_.p, code_stride, d_out_.p, nch, d_nav_.p, nav_cap, d_nav_n_.p));
    R4WB_LAUNCH_CHECK();
    std::vector<int8_t> nav((size_t)n * nav_cap);
    std::vector<uint32_t> nav_n(n);
    R4WB_CUDA(cudaMemcpyAsync(out, d_out_.p, (size_t)n_periods * n * sizeof(r4wb_track_state), cudaMemcpyDeviceToHost, st));
    R4WB_CUDA(cudaMemcpyAsync(host_.data(), d_chan_.p, n * sizeof(TrackChan), cudaMemcpyDeviceToHost, st));
    R4WB_CUDA(cudaMemcpyAsync(nav.data(), d_nav_.p, nav.size(), cudaMemcpyDeviceToHost, st));
    R4WB_CUDA(cudaMemcpyAsync(nav_n.data(), d_nav_n_.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    R4WB_CUDA(cudaStreamSynchronize(st));
    for (uint32_t k = 0; k < n; ++k)
        nav_[k].insert(nav_[k].end(), nav.begin() + (size_t)k * nav_cap, nav.begin() + (size_t)k * nav_cap + std::min(nav_n[k], nav_cap));
}

// TrackingChannel::state :344-358 (C/N0 re-estimated from the buffer)
static double host_estimate_cn0(const TrackChan& c)
{
    if (c.cn0_n < 2u) return 0.0;
    const double n = (double)c.cn0_n;
    double sum = 0.0;
    for (uint32_t i = 0; i < c.cn0_n; ++i) sum += c.cn0_buf[i];
    const double mean = sum / n;
    double vs = 0.0;
    for (uint32_t i = 0; i < c.cn0_n; ++i) { const double d = c.cn0_buf[i] - mean; vs += d * d; }
    const double variance = vs / (n - 1.0);
    if (variance <= 0.0 || mean <= 0.0) return 0.0;
    const double snr = (mean * mean) / variance;
    const double x = snr - 1.0;
    return 10.0 * std::log10((1.0 / 0.001) * (x > 0.01 ? x : 0.01));
}

void TrackerBank::state(r4wb_track_state* out, uint32_t cap) const
{
    const uint32_t n = std::min<uint32_t>(cap, (uint32_t)host_.size());
    for (uint32_t k = 0; k < n; ++k) {
        const TrackChan& c = host_[k];
        r4wb_track_state o{};
        o.code_phase = c.code_phase; o.carrier_freq_hz = c.carrier_freq; o.carrier_phase_rad = c.carrier_phase * 2.0 * 3.14159265358979323846;
        o.prompt_i = c.p_i; o.prompt_q = c.p_q; o.cn0_dbhz = host_estimate_cn0(c); o.ms_count = c.ms_count;
        o.prn = (uint8_t)c.prn; o.carrier_lock = (uint8_t)c.carrier_lock; o.code_lock = (uint8_t)c.code_lock; o.bit_sync = (uint8_t)c.bit_sync;
        out[k] = o;
    }
}

uint64_t TrackerBank::nav_bits(uint32_t channel, int8_t* out, uint64_t cap) const
{
    if (channel >= nav_.size()) fail(R4WB_ERR_INVALID_PARAMETER, "channel %u out of range", channel);
    const std::vector<int8_t>& v = nav_[channel];
    const uint64_t n = std::min<uint64_t>(cap, v.size());
    if (out && n) std::copy(v.begin(), v.begin() + n, out);
    return v.size();
}

}  // namespace r4wb
