// compose.cuh — multi-emitter composition of r4w-sim's ScenarioEngine (compose.cu)
#pragma once
#include <cstdint>
#include <vector>

#include "common.hpp"

namespace r4wb {

struct ComposeEmitter {
    double phase0;        // carrier phase before the block's first increment (rad)
    double inc;           // 2 pi doppler / fs
    double phase_wrap;    // phase after the `%= 2 pi` wrap (taken at sample wrap_at), when the block crosses |phase| > 1e6
    float amp;
    uint32_t active;
    uint32_t wrap_at;     // first sample index that uses phase_wrap as its base (0xffffffff: no wrap inside the block)
    uint32_t pad;
};

class Composer {
public:
    Composer(uint32_t n_emitters, double sample_rate, double noise_std, uint64_t seed);
    void reset();
    uint32_t emitters() const { return (uint32_t)phases_.size(); }
    uint64_t current_sample() const { return current_; }
    const std::vector<double>& phases() const { return phases_; }
    // one generate_block: baseband[e][n] of every emitter, its Doppler (Hz), linear amplitude and active flag
    void block(const void* baseband, r4wb_fmt in_fmt, r4wb_mem in_where, uint64_t n, const double* doppler_hz, const double* amplitude,
               const uint8_t* active, void* out, r4wb_fmt out_fmt, r4wb_mem out_where);

private:
    double fs_, noise_std_;
    uint64_t seed_, current_ = 0;
    std::vector<double> phases_;          // ScenarioEngine::carrier_phases
    DevBuf<unsigned char> d_in_, d_out_;
    DevBuf<ComposeEmitter> d_em_;
};

}  // namespace r4wb
