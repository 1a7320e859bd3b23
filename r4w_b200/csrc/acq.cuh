// acq.cuh — host model of PcpsAcquisition (gnss/acquisition.rs:40-255) on top of the FFT kernels.
#pragma once
#include <cstdint>
#include <vector>

#include "common.hpp"

namespace r4wb {

// per-(row, residue) partial result of the fused IFFT + |.|^2 + arg-max epilogue
struct RowPeak {
    float best;        // largest |corr|^2 (already scaled by 1/N^2)
    uint32_t lag;      // its lag (lowest lag among equals)
    float second;      // second-largest value of the row part
    float sum;         // sum of |corr|^2 over lags < code_length
};

struct PairPeak {      // per-(snapshot, code) reduction over Doppler rows
    double best;
    double sum;
    uint32_t lin;      // d * code_length + lag
    uint32_t near_tie; // top-2 closer than the f32 guard band
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

    void acquire_batch(const void* input, r4wb_fmt fmt, r4wb_mem where, uint64_t n_snapshots, uint64_t stride,
                       uint64_t n_input, const int8_t* codes, uint64_t code_len, const uint8_t* prns, uint32_t n_codes,
                       r4wb_acq_result* out);
    void acquire_grid(const void* input, r4wb_fmt fmt, uint64_t n_input, const int8_t* code, uint64_t code_len,
                      double* power_out, uint64_t cap);

private:
    template <typename T>
    void run_pairs(const void* d_input, r4wb_fmt fmt, uint64_t n_snapshots, uint64_t stride, uint64_t n_input,
                   const int8_t* d_codes, uint64_t code_len, uint32_t n_codes, const uint32_t* pair_list,
                   uint32_t n_pairs_listed, PairPeak* h_out, float* d_grid);
    void ensure_twiddles();
    void finish(const PairPeak& pk, uint8_t prn, r4wb_acq_result& r) const;

    uint64_t code_length_, fft_size_;
    double fs_, dmax_ = 5000.0, dstep_ = 500.0, threshold_ = 2.5;
    uint64_t coherent_ = 1;
    uint64_t guard_count_ = 0;
    int logn_ = 0;

    DevBuf<float2> d_tw32_;
    DevBuf<double2> d_tw64_;
    bool tw_ready_ = false;
    DevBuf<unsigned char> d_in_, d_x_, d_c_;
    DevBuf<int8_t> d_codes_;
    DevBuf<RowPeak> d_rowpeaks_;
    DevBuf<PairPeak> d_pairpeaks_;
    DevBuf<uint32_t> d_pairlist_;
    DevBuf<float> d_grid_;
};

}  // namespace r4wb
