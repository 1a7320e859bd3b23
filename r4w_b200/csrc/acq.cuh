// acq.cuh — PcpsAcquisition (gnss/acquisition.rs:40-255) on top of the FFT engine: host model, kernel
// argument blocks and the __host__ __device__ load/epilogue arithmetic shared with tests/emu/.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "common.hpp"
#include "fft.cuh"

namespace r4wb {

// running (best, second, sum) over |corr|^2 cells; `idx` breaks ties toward the lowest index, which is what
// the reference's strict `>` scan in ascending (Doppler, lag) order does (gnss/acquisition.rs:154-164)
template <typename V>
struct PeakAcc {
    V best, second, sum;
    uint32_t idx;
};
template <typename V> R4WB_HD void peak_init(PeakAcc<V>& a) { a.best = (V)-1; a.second = (V)-1; a.sum = (V)0; a.idx = 0xffffffffu; }
template <typename V> R4WB_HD void peak_push(PeakAcc<V>& a, V mag, uint32_t idx)
{
    a.sum += mag;
    if (mag > a.best || (mag == a.best && idx < a.idx)) { a.second = a.best; a.best = mag; a.idx = idx; }
    else if (mag > a.second) a.second = mag;
}
template <typename V> R4WB_HD void peak_merge(PeakAcc<V>& a, const PeakAcc<V>& o)
{
    a.sum += o.sum;
    if (o.best > a.best || (o.best == a.best && o.idx < a.idx)) {
        const V s = a.best > o.second ? a.best : o.second;
        a.best = o.best; a.idx = o.idx; a.second = s;
    } else {
        const V s = o.best > a.second ? o.best : a.second;
        a.second = s;
    }
}

// per-(row, PRN, residue) result of the fused IFFT + |.|^2 + arg-max epilogue
struct RowPeak {
    double best, second, sum;
    uint32_t lag, pad;
};
// per-(snapshot, PRN) reduction over Doppler rows
struct PairPeak {
    double best, second, sum;
    uint32_t lin;        // d * code_length + lag
    uint32_t pad;
};

// carrier wipe-off fused into the forward-FFT load (gnss/acquisition.rs:133-140): sample n of the snapshot
// times exp(-j 2 pi doppler n / fs), zero beyond `take` samples.
template <typename T>
struct WipeLoad {
    const void* in;     // cf32 or cf64 snapshot
    uint32_t in64, take;
    double doppler, fs;
    R4WB_HD cx<T> operator()(uint32_t n) const
    {
        if (n >= take) return cx<T>{(T)0, (T)0};
        T xr, xi;
        if (in64) { const cx<double> v = static_cast<const cx<double>*>(in)[n]; xr = (T)v.re; xi = (T)v.im; }
        else { const cx<float> v = static_cast<const cx<float>*>(in)[n]; xr = (T)v.re; xi = (T)v.im; }
        T c, s;
        if (sizeof(T) == 8) {       // the reference's own expression, f64
            const double t = (double)n / fs;
            const double phase = ((-2.0 * kPi) * doppler) * t;
            c = (T)cos(phase); s = (T)sin(phase);
        } else {                    // exact cycle reduction in f64, then an f32 sincos of a small argument
            double cyc = doppler * ((double)n / fs);
            cyc -= rint(cyc);
#ifdef __CUDA_ARCH__
            float sf, cf;
            sincospif(-2.0f * (float)cyc, &sf, &cf);
            c = (T)cf; s = (T)sf;
#else
            c = (T)cosf(-6.283185307179586f * (float)cyc); s = (T)sinf(-6.283185307179586f * (float)cyc);
#endif
        }
        return cx<T>{xr * c - xi * s, xr * s + xi * c};
    }
};

// local replica as +-1 int8, zero-padded / truncated to N (gnss/acquisition.rs:109-112)
template <typename T>
struct CodeLoad {
    const int8_t* code;
    uint32_t take;
    R4WB_HD cx<T> operator()(uint32_t n) const { return cx<T>{n < take ? (T)code[n] : (T)0, (T)0}; }
};

// spectrum product fused into the inverse-FFT load (gnss/acquisition.rs:146-148).  The 1/N of ifft_inplace
// (core/fft_utils.rs:104-107) is folded into the stored code spectrum: c = conj(FFT(code)) / N.
template <typename T>
struct ProductLoad {
    const cx<T>* x;      // forward spectrum of the wiped-off snapshot
    const cx<T>* c;      // conj(code spectrum) / N
    R4WB_HD cx<T> operator()(uint32_t n) const { return x[n] * c[n]; }
};

struct AcqGeom {
    int logN, logM, logF;
    uint32_t N, L;           // fft size, code_length (lags scanned)
    uint32_t D, P;           // Doppler bins, codes
    double fs, dmax, dstep;
};

// per-precision device buffers: twiddle table, forward spectra of the current snapshot chunk, code spectra
template <typename T>
struct AcqWork {
    DevBuf<cx<T>> tw, x, c;
    bool tw_ready = false;
};

class Pcps {
public:
    Pcps(uint64_t code_length, double sample_rate);
    ~Pcps();
    void set_doppler_range(double max_hz, double step_hz);
    void set_threshold(double t) { threshold_ = t; }
    void set_coherent_periods(uint64_t p) { coherent_ = p < 1 ? 1 : p; }
    uint64_t fft_size() const { return fft_size_; }
    uint32_t num_bins() const;
    uint64_t guard_count() const { return guard_count_; }
    void set_profiling(bool on) { profiling_ = on; }
    void last_profile(double* ms4, uint64_t* launches4) const;

    void acquire_batch(const void* input, r4wb_fmt fmt, r4wb_mem where, uint64_t n_snapshots, uint64_t stride,
                       uint64_t n_input, const int8_t* codes, uint64_t code_len, const uint8_t* prns, uint32_t n_codes,
                       r4wb_acq_result* out);
    void acquire_grid(const void* input, r4wb_fmt fmt, uint64_t n_input, const int8_t* code, uint64_t code_len,
                      double* power_out, uint64_t cap);

private:
    // search of snapshots [s0, s0+ns) x codes [c0, c0+nc): peaks land in d_pairpeaks_[(s - s0) * nc + (c - c0)]
    template <typename T>
    void run(AcqWork<T>& w, const void* d_input, r4wb_fmt fmt, uint64_t s0, uint64_t ns, uint64_t stride, uint64_t n_input,
             const int8_t* d_codes, uint64_t code_len, uint32_t c0, uint32_t nc, PairPeak* d_out, double* d_grid);
    void finish(const PairPeak& pk, uint8_t prn, r4wb_acq_result& r) const;
    AcqGeom geom(int max_logM) const;

    uint64_t code_length_, fft_size_;
    double fs_, dmax_ = 5000.0, dstep_ = 500.0, threshold_ = 2.5;
    uint64_t coherent_ = 1;
    uint64_t guard_count_ = 0;
    int logn_ = 0;

    // optional per-launch CUDA-event timing (kind 0..3 = code spectra, forward, inverse+peak, pair reduce)
    struct Timed { cudaEvent_t a, b; int kind; };
    void prof_begin(int kind);
    void prof_end();
    void prof_collect();
    bool profiling_ = false;
    std::vector<Timed> timed_;
    std::vector<cudaEvent_t> event_pool_;
    size_t event_next_ = 0;
    double prof_ms_[4] = {0, 0, 0, 0};
    uint64_t prof_n_[4] = {0, 0, 0, 0};

    AcqWork<float> w32_;
    AcqWork<double> w64_;
    DevBuf<unsigned char> d_in_;
    // host input: the H2D copy runs in pieces on its own stream and the snapshot chunks wait only for the pieces they read
    cudaStream_t copy_stream_ = nullptr;
    std::vector<cudaEvent_t> h2d_events_;
    size_t h2d_piece_ = 0, h2d_pieces_ = 0, h2d_waited_ = 0;
    void h2d_start(const void* host, size_t bytes, cudaStream_t st, const int8_t* codes, size_t code_bytes);
    void h2d_wait(size_t end_byte, cudaStream_t st);     // stream st may read bytes [0, end_byte) of d_in_ afterwards
    DevBuf<int8_t> d_codes_;
    DevBuf<RowPeak> d_rowpeaks_;
    DevBuf<PairPeak> d_pairpeaks_;
    DevBuf<double> d_grid_;
};

}  // namespace r4wb
