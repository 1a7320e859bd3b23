// synth_host.cu — host model of GnssScenario (gnss/scenario.rs:51-705) on top of the synthesis kernels.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

#include "synth.cuh"

namespace r4wb {

// kernels / launchers (synth_kernels.cu)
void host_fill_block_sat(const ScenConst&, const SatConst&, const PhaseSegment*, uint64_t, uint32_t, uint64_t, BlockSat&);
uint64_t host_block_advance(const BlockSat&);
size_t synth_smem_bytes(uint32_t n_sats, uint32_t nw64);
int synth_tile_samples(int K);
void launch_synth_kernel(const SynthArgs& a, int K, bool cf64, int grid, cudaStream_t st);
int synth_max_blocks_per_sm(int K, bool cf64, size_t smem);
void launch_block_params(const ScenConst&, const SatConst*, const PhaseSegment*, uint64_t, uint32_t, BlockSat*, BlockHdr*, cudaStream_t);
void launch_phase_scan(const SatConst*, uint32_t, uint32_t, BlockSat*, cudaStream_t);

static const uint8_t kE1Packed[2 * 50 * 512] = {
#include "../../data/galileo_e1_codes.inc"
};

// chip i of PRN `prn` on `channel` (0 = E1B, 1 = E1C) as a sign bit (1 <=> -1):
// MSB-first packing, bit 0 -> +1 (gnss/galileo_e1_codes.rs:17-25)
static inline uint32_t e1_sign_bit(uint32_t channel, uint32_t prn, uint32_t i)
{
    const uint8_t* p = kE1Packed + ((size_t)channel * 50 + (prn - 1)) * 512;
    return (p[i >> 3] >> (7 - (i & 7))) & 1u;
}

void e1_code_chips(uint32_t channel, uint32_t prn, int8_t* out)
{
    for (uint32_t i = 0; i < (uint32_t)kCodeLen; ++i) out[i] = e1_sign_bit(channel, prn, i) ? -1 : 1;
}

// 63-tap Blackman windowed-sinc low-pass, unity DC gain (core/filters/fir.rs:458-499, windows.rs:137-151)
static void design_lowpass(double cutoff_hz, double rate_hz, double* h /*[63]*/)
{
    const double fc = cutoff_hz / rate_hz, mid = (kTaps - 1) / 2.0;
    double sum = 0.0;
    for (int i = 0; i < kTaps; ++i) {
        const double x = 2.0 * kPi * (double)i / (double)(kTaps - 1);
        const double w = 0.42 - 0.5 * cos(x) + 0.08 * cos(2.0 * x);
        const double n = (double)i - mid;
        const double sinc = fabs(n) < 1e-10 ? 2.0 * kPi * fc : sin(2.0 * kPi * fc * n) / n;
        h[i] = sinc * w;
    }
    for (int i = 0; i < kTaps; ++i) sum += h[i];
    if (fabs(sum) > 1e-10)
        for (int i = 0; i < kTaps; ++i) h[i] /= sum;
}

// Piecewise-exact model of `phase += inc` repeated `steps` times in f64 (SURVEY.md §7 hard part 1):
// inside one binade every add moves the phase by the same multiple of the binade's ulp.
static void build_phase_segments(double inc, uint64_t steps, std::vector<PhaseSegment>& out)
{
    double x = 0.0;
    uint64_t i = 0;
    auto push = [&](uint64_t i0, double x0, double step) { out.push_back(PhaseSegment{i0, x0, step}); };
    if (inc == 0.0 || steps == 0) { out.push_back(PhaseSegment{0, 0.0, inc}); return; }
    while (i < steps) {
        const double x1 = x + inc, s1 = x1 - x;
        const double x2 = x1 + inc, s2 = x2 - x1;
        if (x != 0.0 && s1 == s2 && std::ilogb(x) == std::ilogb(x1) && std::ilogb(x1) == std::ilogb(x2)) {
            const int e = std::ilogb(x);
            const long double limit = std::ldexp(1.0L, e + 1);
            const long double room = (limit - fabsl((long double)x)) / fabsl((long double)s1);
            long long fit = (long long)floorl(room) - 2;
            if (fit >= 3) {
                const uint64_t n = std::min<uint64_t>((uint64_t)fit, steps - i);
                push(i, x, s1);
                const long long X0 = (long long)std::scalbn(x, 52 - e), S = (long long)std::scalbn(s1, 52 - e);
                x = std::scalbn((double)(X0 + (long long)n * S), e - 52);
                i += n;
                continue;
            }
        }
        push(i, x, s1);
        x = x1;
        i += 1;
    }
    if ((int)out.size() > kMaxSegments) fail(R4WB_ERR_NOT_SUPPORTED, "phase model needs %zu segments", out.size());
}

static inline uint64_t gcd_u64(uint64_t a, uint64_t b) { while (b) { uint64_t t = a % b; a = b; b = t; } return a; }

Scenario::Scenario(const r4wb_scenario_cfg& cfg) : cfg_(cfg)
{
    if (cfg.n_sats > (uint32_t)kMaxSats) fail(R4WB_ERR_NOT_SUPPORTED, "at most %d satellites per scenario", kMaxSats);
    if (cfg.n_sats && !cfg.sats) fail(R4WB_ERR_NULL_POINTER, "sats is NULL");
    cfg_sats_.assign(cfg.sats, cfg.sats + cfg.n_sats);
    cfg_.sats = cfg_sats_.data();
    const r4wb_output_cfg& oc = cfg.output;
    if (!(oc.sample_rate > 0.0) || !(oc.duration_s >= 0.0)) fail(R4WB_ERR_INVALID_PARAMETER, "sample_rate/duration_s");

    sc_.fs = oc.sample_rate;
    sc_.t0_gps = oc.start_time_gps_s;
    sc_.duration_s = oc.duration_s;
    sc_.total = (uint64_t)std::ceil(oc.duration_s * oc.sample_rate);                       // scenario.rs:80
    sc_.B = oc.block_size > 0 ? oc.block_size : (uint64_t)std::ceil(oc.sample_rate * 0.001);   // scenario.rs:667-674
    if (sc_.B < 8 || sc_.B > 65536) fail(R4WB_ERR_NOT_SUPPORTED, "block size %llu outside [8, 65536]", (unsigned long long)sc_.B);
    sc_.n_sats = cfg.n_sats;
    sc_.flags = cfg.flags;
    sc_.antenna = cfg.receiver.antenna;
    sc_.ant_peak = cfg.receiver.antenna_peak_gain_dbi;
    sc_.ant_bw = cfg.receiver.antenna_beamwidth_deg;
    sc_.elev_mask_deg = cfg.receiver.elevation_mask_deg;
    sc_.seed = oc.seed;

    // receiver model (scenario.rs:160-183, 320-353)
    RxModel& rx = sc_.rx;
    rx.position = Lla{cfg.receiver.position.lat_deg, cfg.receiver.position.lon_deg, cfg.receiver.position.alt_m};
    rx.has_trajectory = cfg.receiver.has_trajectory ? 1 : 0;
    rx.traj_start = Lla{cfg.receiver.traj_start.lat_deg, cfg.receiver.traj_start.lon_deg, cfg.receiver.traj_start.alt_m};
    rx.traj_end = Lla{cfg.receiver.traj_end.lat_deg, cfg.receiver.traj_end.lon_deg, cfg.receiver.traj_end.alt_m};
    rx.travel_time_s = 1.0;
    rx.fd_dt = 0.01;
    if (rx.has_trajectory) {
        const double dist = 6371000.0 * gc_angle(rx.traj_start, rx.traj_end);           // scenario_config.rs:358-368
        const double speed = cfg.receiver.traj_has_speed ? cfg.receiver.traj_speed_mps : dist / oc.duration_s;
        rx.travel_time_s = dist / speed;
        rx.fd_dt = std::min(0.01, rx.travel_time_s * 0.001);
    }

    // code NCO constants.  The reference computes samples_per_chip = (8 fs) / chipping_rate in f64.
    sc_.chip_rate = 1023000.0;
    const double os_rate = oc.sample_rate * (double)kOversample;
    sc_.spc = os_rate / sc_.chip_rate;
    if (os_rate != std::floor(os_rate) || os_rate >= 9007199254740992.0)
        fail(R4WB_ERR_NOT_SUPPORTED, "sample_rate must be an integer number of Hz");
    const uint64_t os_i = (uint64_t)os_rate, cr_i = 1023000ull, g = gcd_u64(os_i, cr_i);
    sc_.ratA = os_i / g;
    sc_.ratB = cr_i / g;
    if (sc_.ratA >= (1ull << 40)) fail(R4WB_ERR_NOT_SUPPORTED, "sample_rate / chip rate ratio too fine");
    sc_.delta = (double)((long double)sc_.spc * (long double)sc_.ratB / (long double)sc_.ratA - 1.0L);
    sc_.delta46 = (uint64_t)floorl(140737488355328.0L / (long double)sc_.spc);
    const double S = sc_.spc * 0.5;   // oversamples per half-chip
    if (S * 4.0 < (double)(kTaps - 1) || S >= 60.0)
        fail(R4WB_ERR_NOT_SUPPORTED, "sample_rate %.0f Hz outside the supported 3.97-15.3 MHz span", oc.sample_rate);
    sc_.kmul = (uint32_t)llround(S * 16777216.0);
    for (int j = 0; j < 8; ++j) sc_.cj[j] = (uint32_t)std::min<long long>(llround((double)j * S * 16777216.0), 0xffffffffll);
    {
        const uint64_t D = sc_.ratA / gcd_u64(2 * sc_.ratB, sc_.ratA);
        sc_.lattice_den = D <= (1ull << 17) ? D : 0;
    }
    {
        const double nf_lin = std::pow(10.0, cfg.receiver.noise_figure_db / 10.0);       // scenario.rs:532-537
        const double n0 = 1.380649e-23 * 290.0 * nf_lin;
        sc_.noise_std = (float)(std::sqrt(n0 * oc.sample_rate / 2.0) * 1e8);
    }

    // satellites
    const Vec3 rx_pos0 = rx_at(rx, 0.0).pos;
    const RxState rx0 = rx_at(rx, 0.0);
    sats_.resize(cfg.n_sats);
    std::vector<uint32_t> codebits((size_t)cfg.n_sats * 128, 0u);
    for (uint32_t k = 0; k < cfg.n_sats; ++k) {
        const r4wb_sat_cfg& c = cfg_sats_[k];
        if (c.signal != R4WB_SIG_GALILEO_E1C)
            fail(R4WB_ERR_NOT_SUPPORTED, "satellite %u: only GalileoE1C is implemented on the GPU path", k);
        if (c.prn < 1 || c.prn > 50) fail(R4WB_ERR_INVALID_PARAMETER, "Galileo PRN must be 1-50, got %u", c.prn);
        if (c.plane >= 3 || c.slot >= 10) fail(R4WB_ERR_INVALID_PARAMETER, "Galileo plane 0-2 / slot 0-9");
        if (!(c.has & R4WB_HAS_IONO) && cfg.environment.ionosphere_enabled)
            fail(R4WB_ERR_NOT_SUPPORTED, "satellite %u: Klobuchar model needed (no iono_delay_m override)", k);
        if (!(c.has & R4WB_HAS_TROPO) && cfg.environment.troposphere_enabled)
            fail(R4WB_ERR_NOT_SUPPORTED, "satellite %u: Saastamoinen model needed (no tropo_delay_m override)", k);
        SatConst& s = sats_[k];
        std::memset(&s, 0, sizeof s);
        s.orbit = nominal_orbit(c.signal, c.plane, c.slot);
        s.carrier_hz = 1575420000.0;
        s.has = c.has;
        s.orbital_dynamics = c.orbital_dynamics ? 1u : 0u;
        s.tx_power_dbw = c.tx_power_dbw;
        s.elevation_deg = c.elevation_deg; s.range_m = c.range_m; s.range_rate_mps = c.range_rate_mps;
        s.doppler_hz = c.doppler_hz; s.doppler_rate_hz_per_s = c.doppler_rate_hz_per_s; s.cn0_dbhz = c.cn0_dbhz;
        s.iono_delay_m = c.iono_delay_m; s.tropo_delay_m = c.tropo_delay_m;
        const bool doppler_from_orbit = c.orbital_dynamics || (!(c.has & R4WB_HAS_DOPPLER) && !(c.has & R4WB_HAS_RANGE_RATE));
        const bool range_from_orbit = c.orbital_dynamics || !(c.has & R4WB_HAS_RANGE);
        s.needs_orbit = (doppler_from_orbit || range_from_orbit || !(c.has & R4WB_HAS_ELEVATION)) ? 1u : 0u;
        const bool const_doppler = !c.orbital_dynamics && (((c.has & R4WB_HAS_DOPPLER) && !(c.has & R4WB_HAS_DOPPLER_RATE)) ||
                                                            (!(c.has & R4WB_HAS_DOPPLER) && (c.has & R4WB_HAS_RANGE_RATE)));
        s.static_phase = (const_doppler && (c.has & R4WB_HAS_ELEVATION)) ? 1u : 0u;
        if (!(c.has & R4WB_HAS_ELEVATION)) any_var_visibility_ = true;
        if (!s.static_phase) any_dynamic_ = true;
        // orbital anchors at t0 (scenario.rs:195-204)
        Vec3 sp, sv;
        orbit_state(s.orbit, sc_.t0_gps, sp, sv);
        s.orb_range_t0 = look_from(rx_pos0, rx0.lla, sp).range_m;
        s.orb_doppler_t0 = -los_rate(rx0.pos, rx0.vel, sp, sv) * s.carrier_hz / kC;
        // constant-Doppler satellites: segments of the sequential f64 phase accumulation
        s.seg_begin = (int32_t)segments_.size();
        if (s.static_phase) {
            const double dop = (c.has & R4WB_HAS_DOPPLER) ? c.doppler_hz : -c.range_rate_mps * s.carrier_hz / kC;
            const double inc = 2.0 * kPi * dop / sc_.fs;                                  // scenario.rs:522
            std::vector<PhaseSegment> segs;
            build_phase_segments(inc, sc_.total, segs);
            segments_.insert(segments_.end(), segs.begin(), segs.end());
        }
        s.seg_count = (int32_t)segments_.size() - s.seg_begin;
        for (uint32_t i = 0; i < (uint32_t)kCodeLen; ++i)
            codebits[(size_t)k * 128 + (i >> 5)] |= e1_sign_bit(1, c.prn, i) << (i & 31);
    }
    if (segments_.empty()) segments_.push_back(PhaseSegment{0, 0.0, 0.0});

    // filter: FirFilter::lowpass(lpf_cutoff or fs/2, 8 fs, 63)  (scenario.rs:209-217)
    double h[kTaps];
    design_lowpass(oc.lpf_cutoff_hz > 0.0 ? oc.lpf_cutoff_hz : oc.sample_rate / 2.0, os_rate, h);
    float taps_f[64] = {0}, etab_f[64];
    double run = 0.0;
    for (int d = 0; d < 64; ++d) {
        if (d < kTaps) { taps_f[d] = (float)h[d]; run += h[d]; }
        etab_f[d] = d >= kTaps - 1 ? 1.0f : (float)run;      // window fully covered -> sum h = 1
    }

    tile_k_ = 5;
    {
        const double span = std::ceil((double)synth_tile_samples(tile_k_) * kOversample * (2.0 / sc_.spc)) + 2.0;
        nw64_ = (uint32_t)std::ceil((span + 8.0) / 32.0) + 1u;
    }

    seq_m_.assign(cfg.n_sats, 0);
    seq_phi_.assign(cfg.n_sats, 0);
    seq_prev_.assign(cfg.n_sats, BlockSat{});
    seq_has_prev_.assign(cfg.n_sats, 0);

    // device constants
    cudaStream_t st = current_stream();
    R4WB_CUDA(cudaMemcpyAsync(d_sat_.reserve(std::max<size_t>(1, sats_.size())), sats_.data(), sats_.size() * sizeof(SatConst), cudaMemcpyHostToDevice, st));
    R4WB_CUDA(cudaMemcpyAsync(d_segments_.reserve(segments_.size()), segments_.data(), segments_.size() * sizeof(PhaseSegment), cudaMemcpyHostToDevice, st));
    R4WB_CUDA(cudaMemcpyAsync(d_codebits_.reserve(std::max<size_t>(128, codebits.size())), codebits.data(), codebits.size() * 4, cudaMemcpyHostToDevice, st));
    R4WB_CUDA(cudaMemcpyAsync(d_taps_.reserve(64), taps_f, sizeof taps_f, cudaMemcpyHostToDevice, st));
    R4WB_CUDA(cudaMemcpyAsync(d_etab_.reserve(64), etab_f, sizeof etab_f, cudaMemcpyHostToDevice, st));
    d_power_.reserve(1);
    R4WB_CUDA(cudaMemsetAsync(d_power_.p, 0, sizeof(double), st));
    R4WB_CUDA(cudaStreamSynchronize(st));
}

Scenario::~Scenario() {}

void Scenario::reset()
{
    current_ = 0;
    std::fill(seq_m_.begin(), seq_m_.end(), 0);
    std::fill(seq_phi_.begin(), seq_phi_.end(), 0);
    std::fill(seq_has_prev_.begin(), seq_has_prev_.end(), 0);
}

void Scenario::build_canonical_table(uint64_t blk_begin, uint64_t blk_end)
{
    const uint64_t nblk = blk_end - blk_begin;
    if (nblk > 0xffffffffull / std::max(1u, sc_.n_sats)) fail(R4WB_ERR_INVALID_SIZE, "too many blocks in one call");
    cudaStream_t st = current_stream();
    d_tab_.reserve(std::max<size_t>(1, (size_t)nblk * sc_.n_sats));
    d_hdr_.reserve(std::max<size_t>(1, nblk));
    tab_blk0_ = blk_begin;
    launch_block_params(sc_, d_sat_.p, d_segments_.p, blk_begin, (uint32_t)nblk, d_tab_.p, d_hdr_.p, st);
    if (sc_.n_sats == 0) {   // headers still needed
        std::vector<BlockHdr> h(nblk);
        for (uint64_t b = 0; b < nblk; ++b) {
            const uint64_t first = (blk_begin + b) * sc_.B;
            h[b] = BlockHdr{first, (uint32_t)std::min<uint64_t>(sc_.B, sc_.total - first), 0};
        }
        R4WB_CUDA(cudaMemcpyAsync(d_hdr_.p, h.data(), nblk * sizeof(BlockHdr), cudaMemcpyHostToDevice, st));
        R4WB_CUDA(cudaStreamSynchronize(st));
    }
    if (any_dynamic_ || any_var_visibility_) launch_phase_scan(d_sat_.p, sc_.n_sats, (uint32_t)nblk, d_tab_.p, st);
}

void Scenario::launch_synth(const BlockSat* tab, const BlockHdr* hdr, uint32_t tb_begin, uint32_t tb_count,
                            uint64_t out_first, uint64_t out_n, void* d_out, r4wb_fmt fmt, uint64_t max_block_n)
{
    SynthArgs a{};
    a.tab = tab; a.hdr = hdr; a.codebits = d_codebits_.p; a.taps = d_taps_.p; a.etab = d_etab_.p;
    a.out = d_out; a.power_sum = d_power_.p;
    a.out_first = out_first; a.out_n = out_n;
    a.tb_begin = tb_begin; a.tb_count = tb_count;
    const uint32_t tile = (uint32_t)synth_tile_samples(tile_k_);
    a.tiles_per_block = (uint32_t)((max_block_n + tile - 1) / tile);
    a.n_sats = sc_.n_sats; a.nw64 = nw64_; a.flags = sc_.flags;
    a.delta46 = sc_.delta46; a.kmul = sc_.kmul;
    for (int j = 0; j < 8; ++j) a.cj[j] = sc_.cj[j];
    a.spc = sc_.spc; a.noise_std = sc_.noise_std; a.seed = sc_.seed;

    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        R4WB_CUDA(cudaGetDevice(&dev));
        R4WB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    const bool cf64 = fmt == R4WB_FMT_CF64;
    const size_t smem = synth_smem_bytes(a.n_sats, a.nw64);
    const int per_sm = std::max(1, synth_max_blocks_per_sm(tile_k_, cf64, smem));
    const uint64_t n_tiles = (uint64_t)tb_count * a.tiles_per_block;
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(n_tiles, (uint64_t)sm_count * per_sm));
    launch_synth_kernel(a, tile_k_, cf64, grid, current_stream());
}

void Scenario::render_to(uint64_t first, uint64_t n, void* dst, r4wb_mem where, r4wb_fmt fmt)
{
    cudaStream_t st = current_stream();
    const size_t bps = fmt == R4WB_FMT_CF64 ? 16 : 8;
    R4WB_CUDA(cudaMemsetAsync(d_power_.p, 0, sizeof(double), st));

    const uint64_t b0 = first / sc_.B, b1 = (first + n - 1) / sc_.B;
    const uint64_t tbl_begin = (any_dynamic_ || any_var_visibility_) ? 0 : (b0 > 0 ? b0 - 1 : 0);
    build_canonical_table(tbl_begin, b1 + 1);

    if (where == R4WB_MEM_DEVICE) {
        launch_synth(d_tab_.p, d_hdr_.p, (uint32_t)(b0 - tbl_begin), (uint32_t)(b1 - b0 + 1), first, n, dst, fmt, sc_.B);
        return;
    }
    // host destination: render chunk by chunk into a device staging buffer and copy out
    const uint64_t chunk_blocks = std::max<uint64_t>(1, (uint64_t)(64u << 20) / sc_.B);   // ~64 Msamples per chunk
    unsigned char* stage = d_stage_.reserve((size_t)std::min<uint64_t>(n, chunk_blocks * sc_.B + sc_.B) * bps);
    for (uint64_t cb = b0; cb <= b1; cb += chunk_blocks) {
        const uint64_t ce = std::min(b1 + 1, cb + chunk_blocks);
        const uint64_t f = std::max(first, cb * sc_.B), l = std::min(first + n, ce * sc_.B);
        launch_synth(d_tab_.p, d_hdr_.p, (uint32_t)(cb - tbl_begin), (uint32_t)(ce - cb), f, l - f, stage, fmt, sc_.B);
        R4WB_CUDA(cudaMemcpyAsync((unsigned char*)dst + (f - first) * bps, stage, (l - f) * bps, cudaMemcpyDeviceToHost, st));
        R4WB_CUDA(cudaStreamSynchronize(st));
    }
}

void Scenario::generate(uint64_t first, uint64_t n, void* dst, r4wb_mem where, r4wb_fmt fmt)
{
    if (first > sc_.total || n > sc_.total - first) fail(R4WB_ERR_INVALID_SIZE, "range [%llu, +%llu) exceeds total_samples %llu",
                                                          (unsigned long long)first, (unsigned long long)n, (unsigned long long)sc_.total);
    if (n == 0) return;
    if (!dst) fail(R4WB_ERR_NULL_POINTER, "dst is NULL");
    render_to(first, n, dst, where, fmt);
}

uint64_t Scenario::generate_block(uint64_t n_req, void* dst, r4wb_mem where, r4wb_fmt fmt)
{
    const uint64_t remaining = sc_.total > current_ ? sc_.total - current_ : 0;
    const uint64_t n = std::min(remaining, n_req);
    if (n == 0) return 0;
    if (!dst) fail(R4WB_ERR_NULL_POINTER, "dst is NULL");
    if (n > 65536) fail(R4WB_ERR_NOT_SUPPORTED, "generate_block: at most 65536 samples per reference block");
    cudaStream_t st = current_stream();
    const uint32_t ns = sc_.n_sats;
    // explicit two-row table: row 0 = each satellite's last visible block, row 1 = this block
    std::vector<BlockSat> tab((size_t)2 * std::max(1u, ns));
    for (uint32_t s = 0; s < ns; ++s) {
        BlockSat cur;
        host_fill_block_sat(sc_, sats_[s], segments_.data(), current_, (uint32_t)n,
                            sats_[s].static_phase ? seq_m_[s] : seq_phi_[s], cur);
        cur.prev = seq_has_prev_[s] ? (int32_t)s : -1;
        tab[s] = seq_prev_[s];
        tab[ns + s] = cur;
    }
    BlockHdr hdr[2] = {{0, 0, 0}, {current_, (uint32_t)n, 0}};
    R4WB_CUDA(cudaMemcpyAsync(d_seq_tab_.reserve(tab.size()), tab.data(), tab.size() * sizeof(BlockSat), cudaMemcpyHostToDevice, st));
    R4WB_CUDA(cudaMemcpyAsync(d_seq_hdr_.reserve(2), hdr, sizeof hdr, cudaMemcpyHostToDevice, st));
    R4WB_CUDA(cudaMemsetAsync(d_power_.p, 0, sizeof(double), st));
    const size_t bps = fmt == R4WB_FMT_CF64 ? 16 : 8;
    void* d_out = where == R4WB_MEM_DEVICE ? dst : (void*)d_stage_.reserve((size_t)n * bps);
    launch_synth(d_seq_tab_.p, d_seq_hdr_.p, 1, 1, current_, n, d_out, fmt, n);
    if (where != R4WB_MEM_DEVICE) R4WB_CUDA(cudaMemcpyAsync(dst, d_out, (size_t)n * bps, cudaMemcpyDeviceToHost, st));
    R4WB_CUDA(cudaStreamSynchronize(st));   // `tab`/`hdr` are stack/heap temporaries
    for (uint32_t s = 0; s < ns; ++s) {
        const BlockSat& cur = tab[ns + s];
        if (!(cur.flags & 1u)) continue;
        seq_m_[s] += n;
        if (!sats_[s].static_phase) seq_phi_[s] += host_block_advance(cur);
        seq_prev_[s] = cur;
        seq_prev_[s].prev = -1;
        seq_has_prev_[s] = 1;
    }
    current_ += n;
    return n;
}

double Scenario::last_power_sum()
{
    double v = 0.0;
    cudaStream_t st = current_stream();
    R4WB_CUDA(cudaMemcpyAsync(&v, d_power_.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    R4WB_CUDA(cudaStreamSynchronize(st));
    return v;
}

// GnssScenario::satellite_status (scenario.rs:564-633): static receiver position, zero receiver velocity
void Scenario::status(r4wb_sat_status* out, uint32_t cap, uint32_t* n_out) const
{
    if (cap < sc_.n_sats) fail(R4WB_ERR_INVALID_SIZE, "status buffer holds %u of %u satellites", cap, sc_.n_sats);
    const double t = sc_.t0_gps + (double)current_ / sc_.fs;
    const Lla rx_lla = sc_.rx.position;
    const Vec3 rx_pos = ecef_of(rx_lla), zero{0.0, 0.0, 0.0};
    for (uint32_t k = 0; k < sc_.n_sats; ++k) {
        const r4wb_sat_cfg& c = cfg_sats_[k];
        Vec3 sp, sv;
        orbit_state(sats_[k].orbit, t, sp, sv);
        const Look la = look_from(rx_pos, rx_lla, sp);
        r4wb_sat_status& o = out[k];
        std::memset(&o, 0, sizeof o);
        o.signal = c.signal; o.prn = c.prn;
        o.range_m = (c.has & R4WB_HAS_RANGE) ? c.range_m : la.range_m;
        o.elevation_deg = (c.has & R4WB_HAS_ELEVATION) ? c.elevation_deg : la.elevation_deg;
        o.azimuth_deg = (c.has & R4WB_HAS_AZIMUTH) ? c.azimuth_deg : la.azimuth_deg;
        o.range_rate_mps = (c.has & R4WB_HAS_RANGE_RATE) ? c.range_rate_mps : los_rate(rx_pos, zero, sp, sv);
        o.doppler_hz = (c.has & R4WB_HAS_DOPPLER) ? c.doppler_hz : -o.range_rate_mps * sats_[k].carrier_hz / kC;
        o.antenna_gain_dbi = antenna_gain_dbi(sc_.antenna, sc_.ant_peak, sc_.ant_bw, o.elevation_deg);
        o.cn0_dbhz = (c.has & R4WB_HAS_CN0) ? c.cn0_dbhz
                                            : c.tx_power_dbw - fspl_db(o.range_m, sats_[k].carrier_hz) + o.antenna_gain_dbi + 204.0;
        o.iono_delay_m = (c.has & R4WB_HAS_IONO) ? c.iono_delay_m : 0.0;
        o.tropo_delay_m = (c.has & R4WB_HAS_TROPO) ? c.tropo_delay_m : 0.0;
        o.visible = o.elevation_deg > 0.0 ? 1 : 0;
        o.clock_correction_s = 0.0;
    }
    if (n_out) *n_out = sc_.n_sats;
}

void Scenario::debug_block(uint64_t block, uint32_t sat, double* o)
{
    const uint64_t nb = (sc_.total + sc_.B - 1) / sc_.B;
    if (block >= nb || sat >= sc_.n_sats) fail(R4WB_ERR_INVALID_PARAMETER, "block/sat out of range");
    const uint64_t tbl_begin = (any_dynamic_ || any_var_visibility_) ? 0 : block;
    build_canonical_table(tbl_begin, block + 1);
    BlockSat e;
    cudaStream_t st = current_stream();
    R4WB_CUDA(cudaMemcpyAsync(&e, d_tab_.p + (size_t)(block - tbl_begin) * sc_.n_sats + sat, sizeof e, cudaMemcpyDeviceToHost, st));
    R4WB_CUDA(cudaStreamSynchronize(st));
    o[0] = (double)(e.flags & 1u);
    o[1] = (double)e.U / 70368744177664.0;              // half-chips
    o[2] = (double)e.phi / 18446744073709551616.0;      // cycles
    o[3] = (double)e.f / 18446744073709551616.0;        // cycles / sample
    o[4] = (double)e.df / 18446744073709551616.0;
    o[5] = e.phase0;
    o[6] = (double)e.e0;
    o[7] = (double)e.amp;
    o[8] = (double)e.flags;
    o[9] = (double)e.eps46 / 70368744177664.0;
    o[10] = (double)e.G;
    o[11] = (double)e.n;
}

}  // namespace r4wb
