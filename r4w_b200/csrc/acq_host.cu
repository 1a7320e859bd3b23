// acq_host.cu — host side of PcpsAcquisition (gnss/acquisition.rs:40-255): batches (snapshot, code) searches
// through the FFT kernels, re-runs near-ties in f64, and finishes the AcquisitionResult in f64 on the host.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "acq.cuh"

namespace r4wb {

// launchers (acq_kernels.cu)
template <typename T> void launch_twiddles(cx<T>* W, uint32_t N, cudaStream_t st);
template <typename T>
void launch_fwd_input(const AcqGeom& g, uint32_t rows, const void* input, uint32_t in64, uint64_t stride, uint32_t take,
                      const cx<T>* W, cx<T>* out, cudaStream_t st);
template <typename T>
void launch_fwd_codes(const AcqGeom& g, uint32_t n_codes, const int8_t* codes, uint64_t code_len, uint32_t take, const cx<T>* W,
                      cx<T>* out, cudaStream_t st);
template <typename T>
void launch_inv_peak(const AcqGeom& g, uint32_t rows, const cx<T>* X, const cx<T>* C, const cx<T>* W, RowPeak* peaks, double* grid,
                     cudaStream_t st);
void launch_pair_reduce(const AcqGeom& g, uint32_t n_snap, const RowPeak* peaks, PairPeak* out, cudaStream_t st);
// register-resident f32 fast path for fft_size 32768 (acq_rf_kernels.cu)
bool rf_supported(const AcqGeom& g);
bool rf_use_tmem();
void launch_rf_fwd_input(const AcqGeom& g, uint32_t rows, const void* input, uint32_t in64, uint64_t stride, uint32_t take,
                         const cx<float>* W, cx<float>* out, cudaStream_t st);
void launch_rf_fwd_codes(const AcqGeom& g, uint32_t n_codes, const int8_t* codes, uint64_t code_len, uint32_t take, const cx<float>* W,
                         cx<float>* out, cudaStream_t st);
void launch_rf_inv_peak(const AcqGeom& g, uint32_t rows, const cx<float>* X, const cx<float>* C, const cx<float>* W, RowPeak* peaks,
                        cudaStream_t st);

// tuning / A-B hook: R4WB_ACQ_ENGINE=smem forces the in-shared-memory engine for every size
static bool rf_enabled()
{
    static int v = -1;
    if (v < 0) { const char* e = std::getenv("R4WB_ACQ_ENGINE"); v = (e && std::strcmp(e, "smem") == 0) ? 0 : 1; }
    return v == 1;
}

// f32 results whose two best cells are closer than this (relative) are re-run in f64 so the reported
// (lag, Doppler bin) is the one an f64 evaluation (the reference's arithmetic) picks
static constexpr double kNearTie = 1e-4;
static constexpr double kNearThreshold = 1e-3;
static constexpr size_t kSpectraChunkBytes = 96u << 20;   // forward spectra kept per chunk: stays inside the 126 MB L2
// With the forward spectrum parked in Tensor Memory every row is read from memory once, so L2 residency no longer matters
// and a chunk is sized to fill many waves of one-CTA-per-row launches instead
static constexpr size_t kSpectraChunkBytesTmem = (size_t)2 << 30;

Pcps::Pcps(uint64_t code_length, double sample_rate) : code_length_(code_length), fs_(sample_rate)
{
    if (!(sample_rate > 0.0)) fail(R4WB_ERR_INVALID_PARAMETER, "sample_rate must be positive");
    if (code_length > (1ull << 17)) fail(R4WB_ERR_NOT_SUPPORTED, "code_length %llu: fft_size above 131072 is not supported", (unsigned long long)code_length);
    uint64_t f = 1;                       // usize::next_power_of_two (acquisition.rs:64)
    while (f < code_length) f <<= 1;
    fft_size_ = f;
    logn_ = 0;
    while ((1ull << logn_) < f) ++logn_;
}

Pcps::~Pcps()
{
    for (cudaEvent_t e : event_pool_) cudaEventDestroy(e);
    for (cudaEvent_t e : h2d_events_) cudaEventDestroy(e);
    if (copy_stream_) cudaStreamDestroy(copy_stream_);
}

// Host input of acquire_batch: copied in 32 MB pieces on a copy stream (after the work already queued on st, which may
// still read the buffer), one event per piece.
void Pcps::h2d_start(const void* host, size_t bytes, cudaStream_t st, const int8_t* codes, size_t code_bytes)
{
    if (!copy_stream_) {
        int lo = 0, hi = 0;
        R4WB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        R4WB_CUDA(cudaStreamCreateWithPriority(&copy_stream_, cudaStreamNonBlocking, hi));
    }
    h2d_piece_ = (size_t)32 << 20;
    h2d_pieces_ = (bytes + h2d_piece_ - 1) / h2d_piece_;
    h2d_waited_ = 0;
    while (h2d_events_.size() < h2d_pieces_ + 2) {
        cudaEvent_t e;
        R4WB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h2d_events_.push_back(e);
    }
    R4WB_CUDA(cudaEventRecord(h2d_events_[h2d_pieces_], st));
    R4WB_CUDA(cudaStreamWaitEvent(copy_stream_, h2d_events_[h2d_pieces_], 0));
    if (code_bytes) {
        R4WB_CUDA(cudaMemcpyAsync(d_codes_.p, codes, code_bytes, cudaMemcpyHostToDevice, copy_stream_));
        R4WB_CUDA(cudaEventRecord(h2d_events_[h2d_pieces_ + 1], copy_stream_));
        R4WB_CUDA(cudaStreamWaitEvent(st, h2d_events_[h2d_pieces_ + 1], 0));
    }
    for (size_t k = 0; k < h2d_pieces_; ++k) {
        const size_t off = k * h2d_piece_, len = std::min(h2d_piece_, bytes - off);
        R4WB_CUDA(cudaMemcpyAsync(d_in_.p + off, static_cast<const unsigned char*>(host) + off, len, cudaMemcpyHostToDevice, copy_stream_));
        R4WB_CUDA(cudaEventRecord(h2d_events_[k], copy_stream_));
    }
}

void Pcps::h2d_wait(size_t end_byte, cudaStream_t st)
{
    if (h2d_pieces_ == 0) return;
    const size_t need = std::min(h2d_pieces_, (end_byte + h2d_piece_ - 1) / h2d_piece_);
    for (; h2d_waited_ < need; ++h2d_waited_) R4WB_CUDA(cudaStreamWaitEvent(st, h2d_events_[h2d_waited_], 0));
}

void Pcps::prof_begin(int kind)
{
    if (!profiling_) return;
    while (event_pool_.size() < event_next_ + 2) {
        cudaEvent_t e;
        R4WB_CUDA(cudaEventCreate(&e));
        event_pool_.push_back(e);
    }
    Timed t{event_pool_[event_next_], event_pool_[event_next_ + 1], kind};
    event_next_ += 2;
    R4WB_CUDA(cudaEventRecord(t.a, current_stream()));
    timed_.push_back(t);
}

void Pcps::prof_end()
{
    if (!profiling_) return;
    R4WB_CUDA(cudaEventRecord(timed_.back().b, current_stream()));
}

void Pcps::prof_collect()   // after the stream has been synchronised
{
    for (const Timed& t : timed_) {
        float ms = 0.0f;
        R4WB_CUDA(cudaEventElapsedTime(&ms, t.a, t.b));
        prof_ms_[t.kind] += (double)ms;
        prof_n_[t.kind] += 1;
    }
    timed_.clear();
    event_next_ = 0;
}

void Pcps::last_profile(double* ms4, uint64_t* launches4) const
{
    for (int k = 0; k < 4; ++k) { ms4[k] = prof_ms_[k]; launches4[k] = prof_n_[k]; }
}

void Pcps::set_doppler_range(double max_hz, double step_hz)
{
    if (!(step_hz > 0.0) || !(max_hz >= 0.0) || !std::isfinite(max_hz) || 2.0 * max_hz / step_hz > 1.0e6)
        fail(R4WB_ERR_INVALID_PARAMETER, "doppler range +-%g Hz step %g Hz", max_hz, step_hz);
    dmax_ = max_hz;
    dstep_ = step_hz;
}

uint32_t Pcps::num_bins() const { return (uint32_t)((int32_t)(2.0 * dmax_ / dstep_) + 1); }   // acquisition.rs:126

AcqGeom Pcps::geom(int max_logM) const
{
    AcqGeom g{};
    g.logN = logn_;
    g.logM = std::min(logn_, max_logM);
    g.logF = g.logN - g.logM;
    if (g.logF > 3) fail(R4WB_ERR_NOT_SUPPORTED, "fft_size %llu too large", (unsigned long long)fft_size_);
    g.N = (uint32_t)fft_size_;
    g.L = (uint32_t)code_length_;
    g.D = num_bins();
    g.P = 0;
    g.fs = fs_; g.dmax = dmax_; g.dstep = dstep_;
    return g;
}

template <typename T>
void Pcps::run(AcqWork<T>& w, const void* d_input, r4wb_fmt fmt, uint64_t s0, uint64_t ns, uint64_t stride, uint64_t n_input,
               const int8_t* d_codes, uint64_t code_len, uint32_t c0, uint32_t nc, PairPeak* d_out, double* d_grid)
{
    cudaStream_t st = current_stream();
    AcqGeom g = geom(fft_max_logM<T>());
    g.P = nc;
    const uint32_t F = 1u << g.logF;
    if (!w.tw_ready) {
        launch_twiddles<T>(w.tw.reserve(g.N), g.N, st);
        w.tw_ready = true;
    }
    const uint32_t take = (uint32_t)std::min<uint64_t>(n_input, g.L);               // input.iter().take(samples_per_code)
    const uint32_t code_take = (uint32_t)std::min<uint64_t>(code_len, g.N);         // code_fft.resize(fft_size)
    // f32, fft_size 32768, no surface dump: the register-resident engine (one CTA per row / per (row, code))
    const bool fast = sizeof(T) == 4 && d_grid == nullptr && rf_enabled() && rf_supported(g) && code_take <= 16384u + 4096u;
    AcqGeom gr = g;                                                                 // geometry of the RowPeak table
    if (fast) gr.logF = 0;
    const uint32_t Fr = 1u << gr.logF;
    w.c.reserve((size_t)nc * g.N);
    prof_begin(0);
    if (fast)
        launch_rf_fwd_codes(g, nc, d_codes + (size_t)c0 * code_len, code_len, code_take, reinterpret_cast<const cx<float>*>(w.tw.p),
                            reinterpret_cast<cx<float>*>(w.c.p), st);
    else
        launch_fwd_codes<T>(g, nc, d_codes + (size_t)c0 * code_len, code_len, code_take, w.tw.p, w.c.p, st);
    prof_end();

    const size_t row_bytes = (size_t)g.N * sizeof(cx<T>);
    const size_t chunk_bytes = fast && rf_use_tmem() ? kSpectraChunkBytesTmem : kSpectraChunkBytes;
    uint64_t chunk = std::max<uint64_t>(1, chunk_bytes / (row_bytes * std::max(1u, g.D)));
    chunk = std::min<uint64_t>(chunk, std::max<uint64_t>(1, 0x3fffffffull / ((uint64_t)std::max(1u, g.D) * nc * F)));
    chunk = std::min(chunk, ns);
    if (fast && rf_use_tmem() && chunk < ns && chunk > 8) {
        // one CTA per (snapshot, Doppler) row and one resident CTA per SM: pick the chunk (within 25 % of the byte budget)
        // whose row count fills whole waves of the SMs best (195 snapshots x 41 rows = 54.02 waves of 148 -> 148 x 41 = 41.00)
        static int sm_count = 0;
        if (!sm_count) {
            int dev = 0;
            R4WB_CUDA(cudaGetDevice(&dev));
            R4WB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
        }
        uint64_t best = chunk;
        double best_eff = 0.0;
        for (uint64_t c = chunk; c >= chunk - chunk / 4 && c >= 1; --c) {
            const uint64_t rows = c * g.D, waves = (rows + sm_count - 1) / sm_count;
            const double eff = (double)rows / (double)(waves * sm_count);
            if (eff > best_eff + 1e-9) { best_eff = eff; best = c; }
        }
        chunk = best;
    }
    w.x.reserve((size_t)chunk * g.D * g.N);
    d_rowpeaks_.reserve((size_t)chunk * g.D * nc * Fr);
    const size_t bps = fmt == R4WB_FMT_CF64 ? 16 : 8;
    for (uint64_t s = 0; s < ns; s += chunk) {
        const uint32_t cs = (uint32_t)std::min<uint64_t>(chunk, ns - s);
        const unsigned char* in = static_cast<const unsigned char*>(d_input) + (s0 + s) * stride * bps;
        if (d_input == d_in_.p) h2d_wait(((s0 + s + cs - 1) * stride + n_input) * bps, st);
        prof_begin(1);
        if (fast)
            launch_rf_fwd_input(g, cs * g.D, in, fmt == R4WB_FMT_CF64 ? 1u : 0u, stride, take, reinterpret_cast<const cx<float>*>(w.tw.p),
                                reinterpret_cast<cx<float>*>(w.x.p), st);
        else
            launch_fwd_input<T>(g, cs * g.D, in, fmt == R4WB_FMT_CF64 ? 1u : 0u, stride, take, w.tw.p, w.x.p, st);
        prof_end();
        prof_begin(2);
        if (fast)
            launch_rf_inv_peak(g, cs * g.D, reinterpret_cast<const cx<float>*>(w.x.p), reinterpret_cast<const cx<float>*>(w.c.p),
                               reinterpret_cast<const cx<float>*>(w.tw.p), d_rowpeaks_.p, st);
        else
            launch_inv_peak<T>(g, cs * g.D, w.x.p, w.c.p, w.tw.p, d_rowpeaks_.p, d_grid, st);
        prof_end();
        prof_begin(3);
        launch_pair_reduce(gr, cs, d_rowpeaks_.p, d_out + s * nc, st);
        prof_end();
    }
}

// acquisition.rs:167-194
void Pcps::finish(const PairPeak& pk, uint8_t prn, r4wb_acq_result& r) const
{
    const uint64_t L = code_length_;
    const uint64_t total_bins = (uint64_t)num_bins() * L;
    double best = pk.best, phase = 0.0, doppler = 0.0;
    if (!(best > 0.0) || pk.lin == 0xffffffffu) {
        best = 0.0;                                        // `mag > best_peak` never fired: initial values survive
    } else {
        const uint64_t d = pk.lin / L;
        phase = (double)(pk.lin - d * L);
        doppler = -dmax_ + (double)d * dstep_;
    }
    const uint64_t denom = total_bins > 1 ? total_bins - 1 : 1;
    const double noise_floor = (pk.sum - best) / (double)denom;
    const double metric = noise_floor > 0.0 ? best / noise_floor : best;
    std::memset(&r, 0, sizeof r);
    r.prn = prn;
    r.detected = metric > threshold_ ? 1 : 0;
    r.code_phase = phase;
    r.doppler_hz = doppler;
    r.peak_metric = metric;
    r.threshold = threshold_;
    if (r.detected) {
        r.has_cn0 = 1;
        r.cn0_estimate = 10.0 * std::log10(metric / ((double)L / fs_));
    }
}

void Pcps::acquire_batch(const void* input, r4wb_fmt fmt, r4wb_mem where, uint64_t S, uint64_t stride, uint64_t n_input,
                         const int8_t* codes, uint64_t code_len, const uint8_t* prns, uint32_t P, r4wb_acq_result* out)
{
    if (S == 0 || P == 0) return;
    if (S * (uint64_t)P > 0x7fffffffull) fail(R4WB_ERR_INVALID_SIZE, "too many (snapshot, code) pairs in one call");
    if (fmt != R4WB_FMT_CF32 && fmt != R4WB_FMT_CF64) fail(R4WB_ERR_INVALID_PARAMETER, "acquisition input must be cf32 or cf64");
    if (code_length_ == 0) fail(R4WB_ERR_INVALID_SIZE, "code_length is 0");
    cudaStream_t st = current_stream();
    const size_t bps = fmt == R4WB_FMT_CF64 ? 16 : 8;
    d_codes_.reserve(std::max<size_t>((size_t)P * code_len, 16));
    bool codes_copied = false;
    const void* d_input = input;
    if (where == R4WB_MEM_HOST) {
        const size_t bytes = ((S - 1) * stride + n_input) * bps;
        d_in_.reserve(std::max<size_t>(bytes, 16));
        h2d_pieces_ = 0;
        static const bool pipelined = [] { const char* e = std::getenv("R4WB_ACQ_H2D_PIPELINE"); return !(e && e[0] == '0'); }();
        if (bytes && pipelined) {
            // The compute stream must not touch the H2D copy engine in this call: after an H2D copy on st the driver places
            // st's next event record on that engine's queue, behind the input pieces, and the first kernel then starts only
            // when the whole input has arrived (measured: 14.6 ms).  So the codes travel on the copy stream too, ahead of the
            // pieces, and st waits for their event.
            h2d_start(input, bytes, st, codes, (size_t)P * code_len);
            codes_copied = true;
        }
        else if (bytes) R4WB_CUDA(cudaMemcpyAsync(d_in_.p, input, bytes, cudaMemcpyHostToDevice, st));
        d_input = d_in_.p;
    }
    if (!codes_copied && code_len) R4WB_CUDA(cudaMemcpyAsync(d_codes_.p, codes, (size_t)P * code_len, cudaMemcpyHostToDevice, st));
    const size_t pairs = (size_t)S * P;
    const size_t guard_slots = 64;
    d_pairpeaks_.reserve(pairs + guard_slots);

    for (int k = 0; k < 4; ++k) { prof_ms_[k] = 0.0; prof_n_[k] = 0; }
    run<float>(w32_, d_input, fmt, 0, S, stride, n_input, d_codes_.p, code_len, 0, P, d_pairpeaks_.p, nullptr);
    std::vector<PairPeak> pk(pairs);
    R4WB_CUDA(cudaMemcpyAsync(pk.data(), d_pairpeaks_.p, pairs * sizeof(PairPeak), cudaMemcpyDeviceToHost, st));
    R4WB_CUDA(cudaStreamSynchronize(st));
    prof_collect();

    // Near-ties and near-threshold metrics: re-run those (snapshot, code) pairs through the f64 engine with the reference's
    // literal wipe-off expression.  The reruns are queued back to back (they share the f64 work buffers in stream order) and
    // read back once per batch.
    guard_count_ = 0;
    std::vector<size_t> guards;
    for (size_t i = 0; i < pairs; ++i) {
        const uint32_t c = (uint32_t)(i % P);
        finish(pk[i], prns ? prns[c] : 0, out[i]);
        const bool near_tie = pk[i].best > 0.0 && (pk[i].best - pk[i].second) <= kNearTie * pk[i].best;
        const bool near_thr = threshold_ > 0.0 && std::fabs(out[i].peak_metric / threshold_ - 1.0) < kNearThreshold;
        if (near_tie || near_thr) guards.push_back(i);
    }
    std::vector<PairPeak> g(guard_slots);
    for (size_t b = 0; b < guards.size(); b += guard_slots) {
        const size_t nb = std::min(guard_slots, guards.size() - b);
        for (size_t j = 0; j < nb; ++j) {
            const size_t i = guards[b + j];
            run<double>(w64_, d_input, fmt, i / P, 1, stride, n_input, d_codes_.p, code_len, (uint32_t)(i % P), 1, d_pairpeaks_.p + pairs + j, nullptr);
        }
        R4WB_CUDA(cudaMemcpyAsync(g.data(), d_pairpeaks_.p + pairs, nb * sizeof(PairPeak), cudaMemcpyDeviceToHost, st));
        R4WB_CUDA(cudaStreamSynchronize(st));
        prof_collect();
        for (size_t j = 0; j < nb; ++j) {
            const size_t i = guards[b + j];
            finish(g[j], prns ? prns[(uint32_t)(i % P)] : 0, out[i]);
            ++guard_count_;
        }
    }
}

void Pcps::acquire_grid(const void* input, r4wb_fmt fmt, uint64_t n_input, const int8_t* code, uint64_t code_len,
                        double* power_out, uint64_t cap)
{
    if (fmt != R4WB_FMT_CF32 && fmt != R4WB_FMT_CF64) fail(R4WB_ERR_INVALID_PARAMETER, "acquisition input must be cf32 or cf64");
    const uint64_t cells = (uint64_t)num_bins() * code_length_;
    if (cap < cells) fail(R4WB_ERR_INVALID_SIZE, "grid buffer holds %llu of %llu cells", (unsigned long long)cap, (unsigned long long)cells);
    if (cells == 0) return;
    cudaStream_t st = current_stream();
    const size_t bps = fmt == R4WB_FMT_CF64 ? 16 : 8;
    d_in_.reserve(std::max<size_t>(n_input * bps, 16));
    h2d_pieces_ = 0;                                       // plain in-stream copy: nothing to wait for
    if (n_input) R4WB_CUDA(cudaMemcpyAsync(d_in_.p, input, n_input * bps, cudaMemcpyHostToDevice, st));
    d_codes_.reserve(std::max<size_t>(code_len, 16));
    if (code_len) R4WB_CUDA(cudaMemcpyAsync(d_codes_.p, code, code_len, cudaMemcpyHostToDevice, st));
    d_grid_.reserve(cells);
    d_pairpeaks_.reserve(2);
    run<double>(w64_, d_in_.p, fmt, 0, 1, 0, n_input, d_codes_.p, code_len, 0, 1, d_pairpeaks_.p, d_grid_.p);
    R4WB_CUDA(cudaMemcpyAsync(power_out, d_grid_.p, cells * sizeof(double), cudaMemcpyDeviceToHost, st));
    R4WB_CUDA(cudaStreamSynchronize(st));
    prof_collect();
}

}  // namespace r4wb
