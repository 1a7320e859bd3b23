// synth.cuh — data structures of the scenario-synthesis path (host model + kernel arguments).
#pragma once
#include <cstdint>
#include <vector>

#include "common.hpp"
#include "geom.hpp"

namespace r4wb {

constexpr int kOversample = 8;        // BASEBAND_OVERSAMPLE, gnss/scenario.rs:47
constexpr int kTaps = 63;             // FirFilter::lowpass(.., 63), gnss/scenario.rs:216
constexpr int kCodeLen = 4092;        // Galileo E1 primary code
constexpr int kSecLen = 25;           // E1C secondary code
constexpr uint32_t kHalfChipsPerSec = 2u * kCodeLen * kSecLen;   // 204 600 half-chips = 100 ms
constexpr int kFracBits = 46;         // fixed-point fraction of the half-chip position
constexpr int kTBits = 24;            // fraction bits of "oversamples since boundary" (8 integer bits: up to 256 oversamples)
constexpr int kMaxSats = 64;
constexpr int kMaxSegments = 512;
constexpr int kMaxYStride = 241;      // boundary-age classes per sign pattern: < 4 S + 1 (S <= 59.9 oversamples per half-chip at 15.3 MHz)
constexpr int kPerBits = 2 * kCodeLen;   // half-chips of one Galileo E1 primary-code period (the longest supported)
constexpr int kPerWords = 260;        // 8184 sign bits + 64 wrap-around bits, padded to a multiple of 4 words
constexpr int kSynthThreads = 256;
constexpr int kLatYStride = 33;       // row stride (words) of the lattice kernel's [class][pattern] table
constexpr int kDirectWords = 320;     // packed chips of a "direct" satellite's primary code (GPS L5: 10 230 chips)

// One piece of the reference's sequential f64 `phase += phase_inc` (gnss/scenario.rs:518-527) for a
// constant-Doppler satellite: from visible-sample count i0 on, phase = x0 + (m - i0) * step EXACTLY
// (x0 and step are multiples of the binade's ulp), until the next segment starts.
struct PhaseSegment {
    uint64_t i0;
    double x0;
    double step;
};

// Code structure of one (virtual) satellite: what SatelliteEmitter::generate_baseband_iq modulates on the real axis
// (gnss/satellite_emitter.rs:264-343) besides the primary code: BOC(1,1) or not, and a sign per primary-code epoch
// (E1C: the 25-chip secondary code; E1B / GPS L1 C/A with nav_data: the reference's deterministic nav bit
// `(epoch / periods_per_bit + prn) % 2`, :286-292).  All supported signals chip at 1.023 MHz.
struct SatCode {
    uint64_t epoch_bits;    // bit e (e < epoch_period) set: primary-code epoch e of the cycle is inverted
    uint32_t code_len;      // chips per primary-code period (4092 Galileo E1, 1023 GPS L1 C/A)
    uint32_t per_len;       // half-chips per primary-code period = 2 * code_len
    uint32_t epoch_period;  // epochs after which epoch_bits repeats (25, 2, 40 or 1)
    uint32_t hc_mod;        // per_len * epoch_period: half-chip positions are kept modulo this (<= 204 600)
    uint32_t has_boc;       // BOC(1,1): second half of every chip inverted
    uint32_t pad;
};

// per-satellite constants (device copy lives in Scenario::d_sat)
struct SatConst {
    SatCode code;
    double chip_rate;            // 1.023e6 for the collapsed-FIR signals; 10.23e6 (GPS L5) / 0.511e6 (GLONASS) are "direct"
    uint32_t direct;             // 1: rendered by k_synth_direct (literal 63-tap evaluation), skipped by k_synth
    uint32_t pad0;
    double amp_scale;            // +1; the two halves of a GalileoE1OS satellite carry +-1/sqrt(2) (satellite_emitter.rs:316-321)
    Orbit orbit;
    double carrier_hz;
    uint32_t has;
    uint32_t orbital_dynamics;
    uint32_t needs_orbit;        // geometry must be evaluated per block
    uint32_t static_phase;       // constant Doppler + constant visibility: emulate the f64 accumulation via segments
    double tx_power_dbw;
    double elevation_deg, range_m, range_rate_mps, doppler_hz, doppler_rate_hz_per_s, cn0_dbhz;
    double iono_delay_m, tropo_delay_m;
    double orb_doppler_t0, orb_range_t0;
    int32_t seg_begin, seg_count;   // into Scenario::d_segments
};

// scenario-wide constants handed to kernels by value
struct ScenConst {
    double fs, t0_gps, duration_s;
    uint64_t total, B;
    uint32_t n_sats, flags;
    RxModel rx;
    uint32_t antenna;
    double ant_peak, ant_bw, elev_mask_deg;
    uint32_t iono_enabled, tropo_enabled;    // Klobuchar / Saastamoinen models for satellites without an override
    double klob_alpha[4], klob_beta[4];
    double tropo_height_m, tropo_temperature_k, tropo_pressure_hpa, tropo_relative_humidity;
    double chip_rate, spc;          // spc = (8 fs) / chip_rate exactly as the reference computes it (f64)
    uint64_t ratA, ratB;            // (8 fs) / chip_rate == ratA / ratB as exact integers
    double delta;                   // spc == (ratA/ratB) * (1 + delta)
    uint64_t delta46;               // half-chips per oversample, 2^-46 units (floor)
    uint64_t lattice_den;           // D: half-chip fractions of a block lie on offset + k/D (0 = dense)
    uint32_t kmul;                  // round(S * 2^24), S = oversamples per half-chip
    uint32_t cj[8];                 // round(j * S * 2^24)
    uint32_t dsum0;                 // (cj[1] >> 24) + (cj[2] >> 24) + (cj[3] >> 24)
    uint32_t lut_den;               // D when the boundary-age class can be read from a D-entry table (0: compute it)
    uint32_t ystride;               // row stride of ytab: smallest odd number >= the number of boundary-age classes (81 at 5 MHz)
    float noise_std;
    uint64_t seed;
};

// per-(block, satellite) entry produced by the prologue and consumed by the synthesis kernel
struct BlockSat {
    uint64_t U;        // half-chip position (18.46 fixed, mod 204600) of the block's first oversample
    uint64_t phi;      // carrier phase before the block's first increment, cycles 0.64
    int64_t f;         // first per-sample increment, cycles 0.64
    int64_t df;        // growth of the increment per sample
    double phase0;     // initial_code_phase (chips), gnss/satellite_emitter.rs:235
    uint64_t G;        // global oversample index of the block's first oversample
    uint32_t n;        // samples in the block
    uint32_t e0;       // initial_epoch_offset, gnss/satellite_emitter.rs:242
    float amp;         // rx_amplitude
    uint32_t flags;    // bit0 visible, bit1 lattice comes within eps of a half-chip boundary
    int32_t prev;      // flat entry index whose tail fills the FIR history (-1: zeros)
    uint32_t eps46;    // ambiguity half-width in 2^-46 half-chips
    uint32_t eps_t;    // same in 2^-24 oversamples
    uint32_t pad;
};
static_assert(sizeof(BlockSat) == 80, "BlockSat layout");

struct BlockHdr { uint64_t first; uint32_t n; uint32_t pad; };

// per-satellite constants of the direct path (k_synth_direct): signals whose chip rate is not 1.023 MHz
struct DirectSat {
    double spc;             // oversamples per chip = (8 fs) / chipping_rate, f64 like the reference (satellite_emitter.rs:245)
    uint64_t epoch_bits;    // nav-bit sign per primary-code epoch (as SatCode)
    uint32_t direct, code_len, epoch_period, pad;
};

// per-tile, per-satellite state (built by k_tile_params, read by k_synth)
struct TileSat {
    uint64_t u0;        // half-chip position (18.46) at the newest oversample of the tile's first sample
    uint64_t phi;
    long long f, df;
    uint32_t hb;        // half-chip index of bit 0 of the sign table
    float amp;
    uint32_t flags;     // bit0 visible, bit1 ambiguity checks needed, bit2 per-sample sincos (large Doppler rate),
                        // bit3 first 8 samples of the block come from yfix, bit4 Doppler varies inside the block
    uint32_t eps_t;
    float wr, wi;       // e^{j phase advance over 2*kSynthThreads samples}, at the tile's first sample
    float th1, th2;     // radians: growth of that advance per sample index, and per step of 2*kSynthThreads samples
};
// what the lattice kernel (k_synth_lat, synth_lattice.cuh) needs on top of TileSat, per (block, satellite)
struct TileLat {
    float r1r, r1i;     // e^{j 2 pi f}: rotation from sample i to i + 1 at the block start (f64-evaluated)
    float rqr, rqi;     // e^{j 2 pi (q f + df q (q + 1) / 2)}: rotation from sample i to i + q at the block start
    uint32_t patch;     // flagged blocks (TileSat.flags bit1): bits 0-15 = oversample index q* < D of the first of the two oversamples
                        // of the block (q*, q* + D) that lie inside the f64 rounding band of a half-chip boundary; bits 16-17 / 18-19:
                        // their codes, 0 = the reference's own f64 expression agrees with the lattice model, 1 = add +2 h[g - q*],
                        // 2 = add -2 h[g - q*] to the windows that hold it
    uint32_t ep0;       // origin of the block's sign table (half-chip hb) inside the satellite's code cycle: bits 0-7 primary-code
                        // epoch, bits 8-31 half-chip inside the period (so the kernel builds sign words without a division)
    uint32_t latb;      // bits 0-15: bin b = floor(frac(U) D) of the block's first sample, bits 16-31: sample offset `moff` of the
                        // class table (class of sample m = clsn[b & 7][m + moff])
    uint32_t fc32;      // (b + 1/2) / D as a 0.32 fraction: the bin-centred half-chip fraction of the block's first sample
};
// one record per (table block, chunk, satellite): 128 bytes = 8 x 16
struct TileRec {
    TileSat ts;
    float yfix[8];      // first 8 FIR outputs of the block when its delay differs from its predecessor's (chunk 0, flags bit3)
    TileLat lat;
};
static_assert(sizeof(TileSat) == 64 && sizeof(TileLat) == 32 && sizeof(TileRec) == 128, "TileRec layout");

// constants of the sample lattice (ScenarioModel fills them when the lattice kernel applies, q = 0 otherwise): the
// half-chip position advances by p / q half-chips per output sample (lowest terms), a block holds exactly 2 q samples,
// half-chip fractions of oversamples lie on offset + k / D with D = 8 q
struct LatConst {
    uint32_t q, p;          // 2500, 1023 at 5 MHz
    uint32_t pinv;          // p^-1 mod q
    uint32_t pov;           // oversample step in units of 1 / D half-chips (2 ratB D / ratA; 1023 at 5 MHz)
    uint32_t povinv;        // pov^-1 mod D
    uint32_t K;             // quad steps per thread: ceil(q / (2 kSynthThreads))
    uint32_t cls_len;       // entries per sub-residue table: q + 2 kSynthThreads K + 2
    uint32_t d8_20, step20; // half-chips per sample / per 2 kSynthThreads samples, 2^-20 units (rounded)
};

struct SynthArgs {
    const BlockSat* tab;       // [n_tab_blocks][n_sats]
    const BlockHdr* hdr;       // [n_tab_blocks]
    const TileRec* tiles;      // [n_tab_blocks][tiles_per_block][n_sats]
    const uint32_t* perbits;   // [n_sats][kPerWords] half-chip signs of one primary-code period (code x BOC(1,1)), bit=1 -> -1
    const SatCode* satcode;    // [n_sats]
    const float* taps;         // [64] h[k] (f32), [63] = 0
    const float* etab;         // [64] E[d] = sum_{k<=d} h[k]  (E[62] = E[63] = 1)
    const float* ytab;         // [32][ystride] collapsed-FIR outputs per (sign pattern, boundary-age class)
    const uint8_t* clslut;     // [lut_den, padded to 16] boundary-age class of half-chip fraction bin q (swizzled, see cls_lut_index)
    uint32_t lut_den;
    uint32_t ystride;
    void* out;                 // cf32 or cf64, out[0] <-> sample out_first
    double* power_sum;         // optional accumulator of |s|^2
    uint64_t out_first, out_n; // only samples in [out_first, out_first + out_n) are written
    uint32_t tb_begin, tb_count;   // table blocks to render
    uint32_t tiles_per_block;
    uint32_t n_sats;
    uint32_t nw64;             // 64-bit words of the per-satellite half-chip sign table
    uint32_t flags;
    const DirectSat* dsat;     // [n_sats] (direct path)
    const uint32_t* dcode;     // [n_sats][kDirectWords] packed chips, bit = 1 -> -1
    uint32_t any_direct;       // some satellite is rendered by k_synth_direct
    uint32_t max_block_n;      // longest block of the table (grid of k_synth_direct)
    uint32_t direct_y0;        // first table block (relative to tb_begin) of this k_synth_direct launch slab
    uint32_t out_aligned16;    // out is aligned to two samples of the output format: sample pairs may be stored as one vector
    LatConst lat;              // lattice kernel constants (lat.q = 0: not applicable)
    const uint8_t* clsn;       // [8][lat.cls_len] boundary-age class of sample index n per sub-residue b0 (synth_lattice.cuh)
    const float* ytab2;        // [ystride][33] collapsed-FIR outputs, row = class, column = sign pattern; the odd row stride makes the
                               // bank (class + pattern) mod 32: a warp's classes differ, so equal patterns no longer share a bank
    uint32_t* stats;           // [2] written by k_tile_params: [0] = records that need per-sample sincos or a varying step phasor
                               // beyond the lattice kernel's model (it then leaves the launch to k_synth)
    uint64_t delta46;
    uint32_t kmul;
    uint32_t cj[8];
    uint32_t dsum0;            // floor(S) + floor(2S) + floor(3S): offset of the boundary-age class index
    double spc;
    float noise_std;
    uint64_t seed;
};

// Host-only model of a scenario: everything GnssScenario::new derives from the config (gnss/scenario.rs:78-237)
// plus the constant tables the kernels read.  No CUDA calls — tests/emu/ builds it without a GPU.
struct ScenarioModel {
    explicit ScenarioModel(const r4wb_scenario_cfg& cfg);

    ScenConst sc{};
    std::vector<SatConst> sats;
    std::vector<r4wb_sat_cfg> cfg_sats;
    r4wb_scenario_cfg cfg{};
    std::vector<PhaseSegment> segments;
    std::vector<uint32_t> codebits;     // [n_sats][128] packed primary code (bit=1 -> chip -1)
    std::vector<uint32_t> perbits;      // [n_sats][kPerWords]
    std::vector<SatCode> satcode;       // [n_sats]
    std::vector<DirectSat> dsat;        // [n_sats]
    std::vector<uint32_t> dcodebits;    // [n_sats][kDirectWords]
    bool any_direct = false;
    std::vector<uint32_t> cfg_index;    // virtual satellite -> index into cfg_sats (GalileoE1OS expands to two)
    float taps_f[64];
    float etab_f[64];
    std::vector<float> ytab;            // [32][sc.ystride]
    std::vector<uint8_t> clslut;        // [lut_den padded to 16]
    LatConst lat{};                     // lattice kernel constants (q = 0: the scenario does not qualify)
    std::vector<uint8_t> clsn;          // [8][lat.cls_len]
    std::vector<float> ytab2;           // [sc.ystride][kLatYStride]
    int tile_k = 10;                    // samples per tile = 256 threads * 2 * tile_k
    uint32_t nw64 = 0;
    bool any_dynamic = false, any_var_visibility = false;

    uint64_t n_blocks() const { return sc.B ? (sc.total + sc.B - 1) / sc.B : 0; }
    // first canonical block a table must start at so that block b0's entries are complete
    uint64_t table_begin(uint64_t b0) const { return (any_dynamic || any_var_visibility) ? 0 : (b0 > 0 ? b0 - 1 : 0); }
    void status(uint64_t current, r4wb_sat_status* out, uint32_t cap, uint32_t* n) const;
};

// sequential-API state (generate_block with caller-chosen block sizes)
struct SeqState {
    std::vector<uint64_t> m;            // visible samples so far (static-phase satellites)
    std::vector<uint64_t> phi;          // carrier phase, cycles 0.64 (dynamic satellites, closed-form scan)
    std::vector<double> ph;             // the reference's own f64 phase accumulator (dynamic satellites; scenario.rs:516-527)
    std::vector<double> dop;            // [2 n_sats] Doppler at the start / end of the block make_table built last
    std::vector<BlockSat> prev;         // last visible block per satellite
    std::vector<uint8_t> has_prev;
    void reset(size_t n_sats);
    // two-row table (row 0 = each satellite's last visible block, row 1 = the block [first, first+n))
    void make_table(const ScenarioModel& md, uint64_t first, uint32_t n, std::vector<BlockSat>& tab, BlockHdr hdr[2]);
    void advance(const ScenarioModel& md, const std::vector<BlockSat>& tab, uint32_t n);
};

// bytes per complex sample of an output format (IqFormat::bytes_per_sample, core/io/format.rs)
inline size_t fmt_bytes(r4wb_fmt fmt)
{
    switch (fmt) { case R4WB_FMT_CF64: return 16; case R4WB_FMT_CF32: return 8; case R4WB_FMT_CI16: return 4; default: return 2; }
}

size_t synth_smem_bytes(uint32_t n_sats, uint32_t nw64, uint32_t lut_den, uint32_t ystride);
int synth_tile_samples(int K);

class Scenario {
public:
    explicit Scenario(const r4wb_scenario_cfg& cfg);
    ~Scenario();

    const r4wb_scenario_cfg& config() const { return md_.cfg; }      // what the handle was created from (satellite array owned by the model)
    uint64_t total_samples() const { return md_.sc.total; }
    uint64_t block_size() const { return md_.sc.B; }
    uint64_t current_sample() const { return current_; }
    bool is_done() const { return current_ >= md_.sc.total; }
    double progress() const { return md_.sc.total == 0 ? 1.0 : (double)current_ / (double)md_.sc.total; }
    void reset();

    // canonical-partition random access; dst is device or host memory
    void generate(uint64_t first, uint64_t n, void* dst, r4wb_mem where, r4wb_fmt fmt);
    // one reference block of min(n, remaining) samples at current_sample
    uint64_t generate_block(uint64_t n, void* dst, r4wb_mem where, r4wb_fmt fmt);
    // the same block, borrowed: pointer into pinned host memory of the handle, valid until the next call on it
    const void* generate_block_view(uint64_t n, r4wb_fmt fmt, uint64_t* n_out);
    // GnssScenario::generate (scenario.rs:549-561): [current_sample, total) and done
    uint64_t generate_rest(void* dst, uint64_t cap, r4wb_mem where, r4wb_fmt fmt);
    // the CLI's file sink (main.rs:4483-4509): the whole scenario, in `fmt`, streamed into `path`; returns sum |s|^2
    double write_file(const char* path, r4wb_fmt fmt, uint64_t* samples, uint64_t* bytes);
    double last_power_sum();
    uint32_t last_path() const { return last_path_; }
    void set_profiling(bool on) { profiling_ = on; }
    void last_profile(double* ms3, uint64_t* launches3);   // of the last generate call: {k_synth, k_synth_periodic, k_periodic_fix, k_synth_lat} (4 entries)
    void status(r4wb_sat_status* out, uint32_t cap, uint32_t* n) const { md_.status(current_, out, cap, n); }
    // test hook: entry of canonical block `block`, satellite `sat` -> 12 doubles
    void debug_block(uint64_t block, uint32_t sat, double* out12);

private:
    void launch_synth(const BlockSat* tab, const BlockHdr* hdr, const TileRec* tiles, uint32_t tb_begin, uint32_t tb_count,
                      uint64_t out_first, uint64_t out_n, void* d_out, r4wb_fmt fmt, uint64_t max_block_n, cudaStream_t st);
    // samples [first, first + n) into device memory: period-resident kernels where the scenario allows, else k_synth
    void render_device(uint64_t first, uint64_t n, void* d_out, r4wb_fmt fmt);
    bool plan_periodic();                                                // once per block table
    bool render_periodic(uint64_t first, uint64_t n, void* d_out);      // false: not applicable to this range
    SynthArgs base_args(const BlockSat* tab, const BlockHdr* hdr, uint64_t max_block_n) const;
    void build_tiles(const SynthArgs& a, uint32_t tb_begin, uint32_t tb_count, TileRec* out);
    void build_canonical_table(uint64_t blk_begin, uint64_t blk_end, uint64_t need_begin = ~0ull);   // fills d_tab_/d_hdr_ for [blk_begin, blk_end)
    void render_to(uint64_t first, uint64_t n, void* dst, r4wb_mem where, r4wb_fmt fmt);

    ScenarioModel md_;
    uint64_t current_ = 0;
    bool seq_canonical_ = true;     // every generate_block so far was a canonical block (start on a multiple of B, B samples)
    SeqState seq_;

    // device state
    DevBuf<SatConst> d_sat_;
    DevBuf<PhaseSegment> d_segments_;
    DevBuf<uint32_t> d_perbits_;
    DevBuf<SatCode> d_satcode_;
    DevBuf<DirectSat> d_dsat_;
    DevBuf<uint32_t> d_dcode_;
    DevBuf<float> d_taps_, d_etab_, d_ytab_;
    DevBuf<uint8_t> d_clslut_, d_clsn_;
    DevBuf<float> d_ytab2_;
    DevBuf<uint32_t> d_stats_;
    uint32_t tab_lat_bad_ = 0;              // records of the canonical table the lattice kernel cannot render
    uint64_t tiles_lo_ = 0;                 // d_tiles_ holds records for table blocks [tiles_lo_, table end) (relative to tab_blk0_)
    bool phase_parallel_ = true;            // the last exact-phase pass ran the parallel kernels (false: serial fallback)
    DevBuf<BlockSat> d_tab_, d_seq_tab_;
    DevBuf<BlockHdr> d_hdr_, d_seq_hdr_;
    DevBuf<TileRec> d_tiles_, d_seq_tiles_;
    uint32_t tiles_per_block_cached_ = 0;   // d_tiles_ holds records for this tiling of the canonical table (0: none)
    DevBuf<double> d_power_;
    DevBuf<unsigned char> d_stage_;
    uint64_t tab_blk0_ = 0, tab_blk1_ = 0;   // canonical block range currently held by d_tab_
    bool tab_valid_ = false;

    // period-resident path (synth_periodic.cu): plan + tables, valid for the block table they were derived from
    struct PeriodicState;
    PeriodicState* per_ = nullptr;
    uint32_t last_path_ = 0;
    struct Timed { cudaEvent_t a, b; int kind; };
    bool profiling_ = false;
    std::vector<cudaEvent_t> event_pool_;
    std::vector<Timed> timed_;
    void prof_begin(int kind, cudaStream_t st);
    void prof_end(cudaStream_t st);
    DevBuf<unsigned char> d_stage2_;          // second staging buffer of the host-destination pipeline
    cudaStream_t side_stream_ = nullptr;      // head / tail launches next to the periodic kernel
    cudaStream_t copy_stream_ = nullptr;      // D2H copies of the host-destination pipeline
    cudaEvent_t ev_fork_ = nullptr, ev_join_ = nullptr, ev_render_[2] = {nullptr, nullptr}, ev_copy_[2] = {nullptr, nullptr};
    void ensure_side_stream();
    // ---- render-ahead ring of the sequential API (generate_block with the canonical block size, host destination): whole
    // chunks of canonical blocks are rendered from the cached canonical table and copied into a pinned host ring ahead of the
    // caller, so one generate_block call is a host memcpy (crates/r4w-cli/src/main.rs:4488-4500 is this loop)
    struct BlockRing;
    BlockRing* ring_ = nullptr;
    uint64_t seq_pos_ = 0;                    // sample position the host-side SeqState corresponds to
    uint64_t ring_block(uint64_t n, void* dst, r4wb_fmt fmt);
    void ring_schedule(uint64_t chunk);
    void ring_drop();
    void seq_sync();                          // replay SeqState up to current_ (after ring-served blocks)
    void* view_buf_ = nullptr;                // pinned bounce buffer of generate_block_view for blocks the ring does not serve
    size_t view_cap_ = 0;
    const void* last_block_ = nullptr;        // last ring-served block (lazy power sum)
    uint64_t last_block_n_ = 0;
    r4wb_fmt last_block_fmt_ = R4WB_FMT_CF32;

    // Cached device tables (block table, tile records, period tables, phasors) are produced on whatever stream the call
    // that built them ran on.  Every public call ends by recording ev_done_ on its stream; a later call on ANOTHER stream
    // waits for that event first, so it never reads a table whose prologue kernels are still in flight (and never touches
    // the old stream handle again, which may be gone by then).
    struct StreamScope;
    cudaEvent_t ev_done_ = nullptr;
    cudaStream_t last_stream_ = nullptr;
    bool has_last_ = false;
};

}  // namespace r4wb
