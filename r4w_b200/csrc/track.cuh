// track.cuh — tracking-channel bank (track.cu): device-resident state of TrackingChannel (gnss/tracking.rs:36-101).
#pragma once
#include <cstdint>
#include <vector>

#include "common.hpp"

namespace r4wb {

// one channel's state, f64 like the reference; lives in HBM between process calls (mirrored on the host after each)
struct TrackChan {
    double sample_rate, chipping_rate;
    double code_phase, code_freq, el_spacing;
    double dll_k1, dll_k2, dll_int;                          // LoopFilter2nd
    double carrier_phase, carrier_freq, fll_bandwidth;
    double pll_k1, pll_k2, pll_k3, pll_i1, pll_i2;          // LoopFilter3rd
    double p_i, p_q;                                         // last prompt correlator
    double cn0_buf[20];
    double nav_acc;
    unsigned long long ms_count;
    uint32_t code_length, prn;
    uint32_t cn0_n, nav_bit_count;
    uint32_t fll_assist, carrier_lock, code_lock, bit_sync;
    int32_t prev_sign;
    uint32_t pad;
};

class TrackerBank {
public:
    TrackerBank(const r4wb_track_cfg* cfgs, uint32_t n);
    uint32_t channels() const { return (uint32_t)host_.size(); }
    // n_periods consecutive TrackingChannel::process calls per channel (tracking.rs:177-313)
    void process(const void* samples, r4wb_fmt fmt, r4wb_mem where, uint64_t n_per_period, uint64_t n_periods, uint64_t channel_stride,
                 const int8_t* codes, uint64_t code_stride, r4wb_track_state* out);
    void state(r4wb_track_state* out, uint32_t cap) const;
    uint64_t nav_bits(uint32_t channel, int8_t* out, uint64_t cap) const;

private:
    std::vector<TrackChan> host_;
    std::vector<std::vector<int8_t>> nav_;
    DevBuf<TrackChan> d_chan_;
    DevBuf<unsigned char> d_in_;
    DevBuf<int8_t> d_codes_, d_nav_;
    DevBuf<uint32_t> d_nav_n_;
    DevBuf<r4wb_track_state> d_out_;
};

}  // namespace r4wb
