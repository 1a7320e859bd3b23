// synth_host.cu — device side of GnssScenario (gnss/scenario.rs:51-705): owns the HBM-resident tables and
// launches the prologue + synthesis kernels.  The host-only model lives in synth_model.cpp.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <condition_variable>
#include <cerrno>
#include <fcntl.h>
#include <unistd.h>
#include <mutex>
#include <thread>

#include "synth_math.cuh"
#include "synth_periodic.cuh"

namespace r4wb {

// launchers (synth_kernels.cu)
void launch_synth_kernel(const SynthArgs& a, int K, r4wb_fmt fmt, int grid, cudaStream_t st);
int synth_max_blocks_per_sm(int K, r4wb_fmt fmt, size_t smem);
void launch_synth_direct(const SynthArgs& a, r4wb_fmt fmt, cudaStream_t st);
void launch_block_params(const ScenConst&, const SatConst*, const PhaseSegment*, uint64_t, uint32_t, BlockSat*, BlockHdr*, double*, double*, cudaStream_t);
bool launch_phase_exact(const ScenConst&, const SatConst*, uint32_t, BlockSat*, const double*, const double*, unsigned char* scratch, cudaStream_t);
size_t phase_exact_scratch_bytes(uint32_t n_sats, uint32_t nblk, uint32_t B);
void launch_phase_scan(const SatConst*, uint32_t, uint32_t, BlockSat*, void* scratch, cudaStream_t);
size_t scan_scratch_bytes(uint32_t n_sats, uint32_t nblk);
void launch_tile_params(const SynthArgs& a, uint32_t tb_begin, uint32_t tb_count, uint32_t tile_samples, TileRec* out, cudaStream_t st);
// lattice kernel (synth_lattice.cu)
bool lat_supported(const SynthArgs& a);
void launch_synth_lat(const SynthArgs& a, r4wb_fmt fmt, int sm_count, cudaStream_t st);
// launchers (synth_periodic.cu)
void launch_static_check(const BlockSat*, const SatConst*, uint32_t nblk, uint32_t n_sats, uint64_t B, uint32_t* bad, cudaStream_t);
void launch_period_tables(const PeriodTableArgs&, uint32_t ns_padded, cudaStream_t);
void launch_period_phasors(const PeriodicArgs&, uint32_t ns_padded, uint64_t tab_blk1, float4* T, cudaStream_t);
void launch_synth_periodic(const PeriodicArgs&, uint32_t ns_padded, cudaStream_t);
void launch_periodic_fix(const PeriodicArgs&, uint32_t n_cands, cudaStream_t);
uint32_t periodic_padded_sats(uint32_t n);

struct Scenario::PeriodicState {
    bool planned = false, ok = false, enabled = true;
    uint64_t blk0 = 0, blk1 = 0;            // block table the plan belongs to
    uint32_t L = 0, ns = 0, tile_len = 0, n_tiles = 0, n_cands = 0;
    uint64_t k_ref = 0;
    uint64_t T_k0 = 0;
    uint32_t T_n = 0;
    DevBuf<float> ys, yb;
    DevBuf<PerSat> sat;
    DevBuf<uint2> cands;
    DevBuf<uint32_t> counters;              // counters[0] = precondition violations, [1] = number of patch slots
    DevBuf<float4> T;
};

// ---- render-ahead ring -------------------------------------------------------------------------------------------------
struct Scenario::BlockRing {
    static constexpr uint32_t kSlots = 4;           // pinned chunks
    uint32_t C = 0;                                 // canonical blocks per chunk
    r4wb_fmt fmt = R4WB_FMT_CF32;
    size_t chunk_bytes = 0;
    unsigned char* pin = nullptr;                   // kSlots * chunk_bytes, pinned
    DevBuf<unsigned char> stage[2];
    cudaEvent_t ev_done[kSlots] = {};               // chunk in slot is in host memory
    cudaEvent_t ev_render[2] = {};
    uint64_t chunk_of_slot[kSlots];
    uint64_t synced_chunk = ~0ull;                  // chunk whose event the host has already waited for
    uint64_t next_chunk = 0;                        // next chunk to schedule
    uint64_t n_chunks = 0;
    ~BlockRing()
    {
        for (cudaEvent_t e : ev_done) if (e) cudaEventDestroy(e);
        for (cudaEvent_t e : ev_render) if (e) cudaEventDestroy(e);
        if (pin) cudaFreeHost(pin);
    }
};

Scenario::Scenario(const r4wb_scenario_cfg& cfg) : md_(cfg)
{
    seq_.reset(md_.sc.n_sats);
    cudaStream_t st = current_stream();
    const auto& sats = md_.sats;
    R4WB_CUDA(cudaMemcpyAsync(d_sat_.reserve(std::max<size_t>(1, sats.size())), sats.data(), sats.size() * sizeof(SatConst), cudaMemcpyHostToDevice, st));
    R4WB_CUDA(cudaMemcpyAsync(d_segments_.reserve(md_.segments.size()), md_.segments.data(), md_.segments.size() * sizeof(PhaseSegment), cudaMemcpyHostToDevice, st));
    R4WB_CUDA(cudaMemcpyAsync(d_perbits_.reserve(md_.perbits.size()), md_.perbits.data(), md_.perbits.size() * 4, cudaMemcpyHostToDevice, st));
    R4WB_CUDA(cudaMemcpyAsync(d_satcode_.reserve(md_.satcode.size()), md_.satcode.data(), md_.satcode.size() * sizeof(SatCode), cudaMemcpyHostToDevice, st));
    R4WB_CUDA(cudaMemcpyAsync(d_dsat_.reserve(md_.dsat.size()), md_.dsat.data(), md_.dsat.size() * sizeof(DirectSat), cudaMemcpyHostToDevice, st));
    R4WB_CUDA(cudaMemcpyAsync(d_dcode_.reserve(md_.dcodebits.size()), md_.dcodebits.data(), md_.dcodebits.size() * 4, cudaMemcpyHostToDevice, st));
    R4WB_CUDA(cudaMemcpyAsync(d_ytab_.reserve(md_.ytab.size()), md_.ytab.data(), md_.ytab.size() * 4, cudaMemcpyHostToDevice, st));
    R4WB_CUDA(cudaMemcpyAsync(d_taps_.reserve(64), md_.taps_f, sizeof md_.taps_f, cudaMemcpyHostToDevice, st));
    R4WB_CUDA(cudaMemcpyAsync(d_etab_.reserve(64), md_.etab_f, sizeof md_.etab_f, cudaMemcpyHostToDevice, st));
    R4WB_CUDA(cudaMemcpyAsync(d_clslut_.reserve(md_.clslut.size()), md_.clslut.data(), md_.clslut.size(), cudaMemcpyHostToDevice, st));
    R4WB_CUDA(cudaMemcpyAsync(d_clsn_.reserve(md_.clsn.size()), md_.clsn.data(), md_.clsn.size(), cudaMemcpyHostToDevice, st));
    R4WB_CUDA(cudaMemcpyAsync(d_ytab2_.reserve(md_.ytab2.size()), md_.ytab2.data(), md_.ytab2.size() * 4, cudaMemcpyHostToDevice, st));
    d_stats_.reserve(2);
    R4WB_CUDA(cudaMemsetAsync(d_stats_.p, 0, 2 * sizeof(uint32_t), st));
    d_power_.reserve(1);
    R4WB_CUDA(cudaMemsetAsync(d_power_.p, 0, sizeof(double), st));
    R4WB_CUDA(cudaStreamSynchronize(st));
}

struct Scenario::StreamScope {
    Scenario& s;
    cudaStream_t st;
    explicit StreamScope(Scenario& sc) : s(sc), st(current_stream())
    {
        if (!s.ev_done_) R4WB_CUDA(cudaEventCreateWithFlags(&s.ev_done_, cudaEventDisableTiming));
        if (s.has_last_ && s.last_stream_ != st) R4WB_CUDA(cudaStreamWaitEvent(st, s.ev_done_, 0));
    }
    ~StreamScope()
    {
        if (cudaEventRecord(s.ev_done_, st) == cudaSuccess) { s.last_stream_ = st; s.has_last_ = true; }
    }
};

Scenario::~Scenario()
{
    ring_drop();
    if (view_buf_) cudaFreeHost(view_buf_);
    delete per_;
    if (ev_done_) cudaEventDestroy(ev_done_);
    for (cudaEvent_t e : {ev_fork_, ev_join_, ev_render_[0], ev_render_[1], ev_copy_[0], ev_copy_[1]})
        if (e) cudaEventDestroy(e);
    if (side_stream_) cudaStreamDestroy(side_stream_);
    if (copy_stream_) cudaStreamDestroy(copy_stream_);
    for (cudaEvent_t e : event_pool_) cudaEventDestroy(e);
}

// optional CUDA-event timing of the synthesis kernels (measurement aid): kinds 0 = k_synth, 1 = k_synth_periodic, 2 = k_periodic_fix
void Scenario::prof_begin(int kind, cudaStream_t st)
{
    if (!profiling_) return;
    while (event_pool_.size() < timed_.size() * 2 + 2) {
        cudaEvent_t e;
        R4WB_CUDA(cudaEventCreate(&e));
        event_pool_.push_back(e);
    }
    Timed t{event_pool_[timed_.size() * 2], event_pool_[timed_.size() * 2 + 1], kind};
    R4WB_CUDA(cudaEventRecord(t.a, st));
    timed_.push_back(t);
}

void Scenario::prof_end(cudaStream_t st)
{
    if (!profiling_) return;
    R4WB_CUDA(cudaEventRecord(timed_.back().b, st));
}

void Scenario::last_profile(double* ms3, uint64_t* launches3)
{
    for (int k = 0; k < 4; ++k) { ms3[k] = 0.0; launches3[k] = 0; }
    for (const Timed& t : timed_) {
        R4WB_CUDA(cudaEventSynchronize(t.b));
        float ms = 0.0f;
        R4WB_CUDA(cudaEventElapsedTime(&ms, t.a, t.b));
        ms3[t.kind] += (double)ms;
        launches3[t.kind] += 1;
    }
}

void Scenario::ensure_side_stream()
{
    if (side_stream_) return;
    int prio_lo = 0, prio_hi = 0;
    R4WB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    R4WB_CUDA(cudaStreamCreateWithPriority(&side_stream_, cudaStreamNonBlocking, prio_hi));   // its few CTAs go ahead of the bulk's queue
    R4WB_CUDA(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
    for (cudaEvent_t* e : {&ev_fork_, &ev_join_, &ev_render_[0], &ev_render_[1], &ev_copy_[0], &ev_copy_[1]})
        R4WB_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
}

void Scenario::reset()
{
    current_ = 0;
    seq_canonical_ = true;
    seq_.reset(md_.sc.n_sats);
    seq_pos_ = 0;
    last_block_ = nullptr;
    if (ring_) { ring_->synced_chunk = ~0ull; }      // chunks stay valid (same samples); the ring repositions on the next call
}

// Prologue: Phase-1 parameters + NCO start values of canonical blocks [blk_begin, blk_end) into d_tab_.
// The table is kept between calls (a bench loop re-rendering the same range pays for it once).
// `need_begin` (absolute block index, default blk_begin): first block the caller is going to render.  The block table of a
// dynamic scenario must start at block 0 (the phase prefix), the per-tile records are only needed from one block before
// `need_begin` on — the last rank of a time-sharded run builds an N-th of them.
void Scenario::build_canonical_table(uint64_t blk_begin, uint64_t blk_end, uint64_t need_begin)
{
    const ScenConst& sc = md_.sc;
    if (need_begin == ~0ull || need_begin < blk_begin) need_begin = blk_begin;
    if (tab_valid_ && blk_begin >= tab_blk0_ && blk_end <= tab_blk1_ && (blk_begin == tab_blk0_ || !(md_.any_dynamic || md_.any_var_visibility))) {
        // cached table: make sure the tile records reach down to the block before the first one rendered
        const uint64_t lo = need_begin > tab_blk0_ ? need_begin - tab_blk0_ - 1 : 0;
        if (lo < tiles_lo_) {
            cudaStream_t st = current_stream();
            SynthArgs a = base_args(d_tab_.p, d_hdr_.p, sc.B);
            if (md_.lat.q != 0) { a.stats = d_stats_.p; R4WB_CUDA(cudaMemsetAsync(d_stats_.p, 0, 2 * sizeof(uint32_t), st)); }
            build_tiles(a, (uint32_t)lo, (uint32_t)(tiles_lo_ - lo), d_tiles_.p);
            if (md_.lat.q != 0) {
                uint32_t bad = 0;
                R4WB_CUDA(cudaMemcpyAsync(&bad, d_stats_.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
                R4WB_CUDA(cudaStreamSynchronize(st));
                tab_lat_bad_ += bad;
            }
            tiles_lo_ = lo;
        }
        return;
    }
    const uint64_t nblk = blk_end - blk_begin;
    if (nblk > 0x7fffffffull / std::max(1u, sc.n_sats)) fail(R4WB_ERR_INVALID_SIZE, "too many blocks in one call");
    cudaStream_t st = current_stream();
    tab_valid_ = false;
    if (per_) per_->planned = false;
    // R4WB_PROLOGUE_TRACE=1: wall clock of the stages on stderr (measurement aid; cudaMalloc / cudaFree of the ~GB tables of
    // a 600 s file vary from a few ms to over a second between processes on the same box, the kernels do not)
    static const bool trace = [] { const char* e = std::getenv("R4WB_PROLOGUE_TRACE"); return e && e[0] == '1'; }();
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    const auto t_begin = now();
    auto t_alloc = t_begin, t_phase = t_begin;
    d_tab_.reserve(std::max<size_t>(1, (size_t)nblk * sc.n_sats));
    d_hdr_.reserve(std::max<size_t>(1, nblk));
    // dynamic satellites: the reference's sequentially accumulated f64 carrier phase, reproduced exactly (synth_math.cuh:
    // block_phase_q / phase_after_block) unless the caller asked for the closed form; needs the table to start at block 0
    const bool exact_phase = md_.any_dynamic && !(sc.flags & R4WB_FLAG_CLOSED_FORM_PHASE) && blk_begin == 0 && sc.n_sats > 0;
    // The scratch of the exact-phase pass (per-block Doppler pair, approximate phase, the scan's work arrays: ~75 B per table
    // entry) lives only until the tile records (128 B per entry) are written, on the same stream: it is carved out of the
    // tile-record buffer, so a cold table costs three device allocations and no free (cudaFree synchronises the device, and
    // both calls were the unpredictable part of the cold path: 2-20 ms usually, over a second on a busy host)
    const SynthArgs a_tiles = base_args(d_tab_.p, d_hdr_.p, sc.B);
    const size_t ne = (size_t)nblk * sc.n_sats;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t scan_bytes = scan_scratch_bytes(sc.n_sats, (uint32_t)nblk);           // k_phase_scan's chunk carries: first in the buffer
    const size_t scratch_bytes = scan_bytes + (exact_phase ? al(2 * ne * sizeof(double)) + al(ne * sizeof(double)) + phase_exact_scratch_bytes(sc.n_sats, (uint32_t)nblk, (uint32_t)sc.B) : 0);
    d_tiles_.reserve(std::max<size_t>(std::max<size_t>(1, (size_t)nblk * a_tiles.tiles_per_block * sc.n_sats), (scratch_bytes + sizeof(TileRec) - 1) / sizeof(TileRec)));
    double *dop = nullptr, *papprox = nullptr;
    unsigned char* pscratch = nullptr;
    if (exact_phase) {
        unsigned char* sp = reinterpret_cast<unsigned char*>(d_tiles_.p) + scan_bytes;
        dop = reinterpret_cast<double*>(sp); sp += al(2 * ne * sizeof(double));
        papprox = reinterpret_cast<double*>(sp); sp += al(ne * sizeof(double));
        pscratch = sp;
    }
    if (trace) t_alloc = now();
    launch_block_params(sc, d_sat_.p, d_segments_.p, blk_begin, (uint32_t)nblk, d_tab_.p, d_hdr_.p, dop, papprox, st);
    if (sc.n_sats == 0) {   // headers still needed
        std::vector<BlockHdr> h(nblk);
        for (uint64_t b = 0; b < nblk; ++b) {
            const uint64_t first = (blk_begin + b) * sc.B;
            h[b] = BlockHdr{first, (uint32_t)std::min<uint64_t>(sc.B, sc.total - first), 0};
        }
        R4WB_CUDA(cudaMemcpyAsync(d_hdr_.p, h.data(), nblk * sizeof(BlockHdr), cudaMemcpyHostToDevice, st));
        R4WB_CUDA(cudaStreamSynchronize(st));
    }
    if (md_.any_dynamic || md_.any_var_visibility) launch_phase_scan(d_sat_.p, sc.n_sats, (uint32_t)nblk, d_tab_.p, d_tiles_.p, st);
    if (exact_phase) phase_parallel_ = launch_phase_exact(sc, d_sat_.p, (uint32_t)nblk, d_tab_.p, dop, papprox, pscratch, st);
    if (trace) { cudaStreamSynchronize(st); t_phase = now(); }
    tab_blk0_ = blk_begin;
    tab_blk1_ = blk_end;
    tab_valid_ = true;
    // per-tile records for the canonical tiling (block size B); k_tile_params counts the records whose carrier model the
    // lattice kernel cannot follow (large Doppler rate), read back once per table
    {
        SynthArgs a = a_tiles;                      // the records overwrite the phase scratch (stream order)
        tab_lat_bad_ = 0;
        if (md_.lat.q != 0) {
            a.stats = d_stats_.p;
            R4WB_CUDA(cudaMemsetAsync(d_stats_.p, 0, 2 * sizeof(uint32_t), st));
        }
        tiles_lo_ = need_begin > blk_begin ? need_begin - blk_begin - 1 : 0;
        build_tiles(a, (uint32_t)tiles_lo_, (uint32_t)(nblk - tiles_lo_), d_tiles_.p);
        if (md_.lat.q != 0) {
            R4WB_CUDA(cudaMemcpyAsync(&tab_lat_bad_, d_stats_.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            R4WB_CUDA(cudaStreamSynchronize(st));
        }
        if (trace) {
            cudaStreamSynchronize(st);
            std::fprintf(stderr, "[r4wb prologue] %llu blocks: allocations %.1f ms, block params + phase %.1f ms, tile records %.1f ms\n",
                         (unsigned long long)nblk, ms(t_begin, t_alloc), ms(t_alloc, t_phase), ms(t_phase, now()));
        }
    }
}

SynthArgs Scenario::base_args(const BlockSat* tab, const BlockHdr* hdr, uint64_t max_block_n) const
{
    const ScenConst& sc = md_.sc;
    SynthArgs a{};
    a.tab = tab; a.hdr = hdr; a.perbits = d_perbits_.p; a.satcode = d_satcode_.p; a.taps = d_taps_.p; a.etab = d_etab_.p; a.ytab = d_ytab_.p; a.clslut = d_clslut_.p; a.lut_den = sc.lut_den; a.ystride = sc.ystride;
    a.dsat = d_dsat_.p; a.dcode = d_dcode_.p; a.any_direct = md_.any_direct ? 1u : 0u; a.max_block_n = (uint32_t)max_block_n;
    a.lat = md_.lat; a.clsn = d_clsn_.p; a.ytab2 = d_ytab2_.p; a.stats = nullptr;
    const uint32_t tile = (uint32_t)synth_tile_samples(md_.tile_k);
    a.tiles_per_block = (uint32_t)((max_block_n + tile - 1) / tile);
    a.n_sats = sc.n_sats; a.nw64 = md_.nw64; a.flags = sc.flags;
    a.delta46 = sc.delta46; a.kmul = sc.kmul; a.dsum0 = sc.dsum0;
    for (int j = 0; j < 8; ++j) a.cj[j] = sc.cj[j];
    a.spc = sc.spc; a.noise_std = sc.noise_std; a.seed = sc.seed;
    return a;
}

// per-tile records of table blocks [tb_begin, tb_begin + tb_count) (k_tile_params)
void Scenario::build_tiles(const SynthArgs& a, uint32_t tb_begin, uint32_t tb_count, TileRec* out)
{
    launch_tile_params(a, tb_begin, tb_count, (uint32_t)synth_tile_samples(md_.tile_k), out, current_stream());
}

void Scenario::launch_synth(const BlockSat* tab, const BlockHdr* hdr, const TileRec* tiles, uint32_t tb_begin, uint32_t tb_count,
                            uint64_t out_first, uint64_t out_n, void* d_out, r4wb_fmt fmt, uint64_t max_block_n, cudaStream_t st)
{
    SynthArgs a = base_args(tab, hdr, max_block_n);
    const uint32_t tb_begin_all = tb_begin, tb_count_all = tb_count;
    a.tiles = tiles;
    a.out = d_out; a.power_sum = d_power_.p;
    a.out_first = out_first; a.out_n = out_n;
    a.tb_begin = tb_begin; a.tb_count = tb_count;
    a.out_aligned16 = ((uintptr_t)d_out % (2 * fmt_bytes(fmt))) == 0 ? 1u : 0u;

    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        R4WB_CUDA(cudaGetDevice(&dev));
        R4WB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    // Lattice kernel for the run of full canonical blocks; a partial last block of the scenario (or a caller-sized block of
    // the sequential API) keeps k_synth.  Both write disjoint samples of the same output.
    if (tiles == d_tiles_.p && max_block_n == md_.sc.B && tab_lat_bad_ == 0 && lat_supported(a)) {
        const uint64_t nb_total = md_.n_blocks();
        const bool last_partial = md_.sc.total % md_.sc.B != 0;
        const uint64_t abs_last = tab_blk0_ + tb_begin + tb_count - 1;
        uint32_t n_full = tb_count;
        if (last_partial && abs_last == nb_total - 1) n_full -= 1;
        if (n_full > 0) {
            SynthArgs f = a;
            f.tb_count = n_full;
            prof_begin(3, st);
            launch_synth_lat(f, fmt, sm_count, st);
            prof_end(st);
        }
        if (n_full == tb_count) {
            if (md_.any_direct) launch_synth_direct(a, fmt, st);
            return;
        }
        a.tb_begin = tb_begin + n_full;
        a.tb_count = tb_count - n_full;
        tb_count = a.tb_count;
    }
    const int tile_k = md_.tile_k;
    const size_t smem = synth_smem_bytes(a.n_sats, a.nw64, a.lut_den, a.ystride);
    thread_local int per_sm_cache[2][5] = {};                    // [tile_k == 5][fmt], for the smem size of the first query
    thread_local size_t per_sm_smem[2][5] = {};
    int& cached = per_sm_cache[tile_k == 5][(int)fmt];
    if (!cached || per_sm_smem[tile_k == 5][(int)fmt] != smem) {
        cached = std::max(1, synth_max_blocks_per_sm(tile_k, fmt, smem));
        per_sm_smem[tile_k == 5][(int)fmt] = smem;
    }
    const int per_sm = cached;
    const uint64_t n_tiles = (uint64_t)tb_count * a.tiles_per_block;
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(n_tiles, (uint64_t)sm_count * per_sm));
    prof_begin(0, st);
    launch_synth_kernel(a, tile_k, fmt, grid, st);
    prof_end(st);
    if (md_.any_direct) {                                     // GPS L5 / GLONASS: added to what the kernels above wrote
        SynthArgs d = a;
        d.tb_begin = tb_begin_all; d.tb_count = tb_count_all;
        launch_synth_direct(d, fmt, st);
    }
}

static bool periodic_enabled()
{
    const char* e = std::getenv("R4WB_SYNTH_PERIODIC");      // read per plan so that tests can A/B inside one process
    return !(e && e[0] == '0');
}

// Decide once per block table whether the period-resident kernels apply, and build their tables.
bool Scenario::plan_periodic()
{
    if (!per_) per_ = new PeriodicState;
    PeriodicState& P = *per_;
    const bool enabled = periodic_enabled();
    if (P.planned && P.enabled == enabled && P.blk0 == tab_blk0_ && P.blk1 == tab_blk1_) return P.ok;
    P.enabled = enabled;
    P.planned = true; P.ok = false; P.blk0 = tab_blk0_; P.blk1 = tab_blk1_; P.T_n = 0;
    const ScenConst& sc = md_.sc;
    if (!enabled || sc.n_sats == 0 || md_.any_dynamic || md_.any_var_visibility || md_.any_direct) return false;
    P.ns = periodic_padded_sats(sc.n_sats);
    if (P.ns == 0) return false;
    // one primary-code period must be a whole number of output samples, the same for every satellite
    const uint32_t code_len = md_.satcode[0].code_len;
    for (const SatCode& c : md_.satcode) if (c.code_len != code_len) return false;
    const unsigned __int128 num = (unsigned __int128)code_len * sc.ratA, den = (unsigned __int128)sc.ratB * kOversample;
    if (num % den != 0) return false;
    const uint64_t L = (uint64_t)(num / den);
    if (L < 1024 || L > (1u << 20) || L % kPerSlotsPerThread != 0) return false;
    P.L = (uint32_t)L;
    P.n_tiles = (uint32_t)((L + 1023) / 1024);
    P.tile_len = (uint32_t)(((L + P.n_tiles - 1) / P.n_tiles + 3) / 4 * 4);
    // reference period: the first one whose samples and FIR history lie inside the table
    P.k_ref = ((tab_blk0_ + 1) * sc.B + L - 1) / L;
    if (P.k_ref == 0) P.k_ref = 1;
    if ((P.k_ref + 1) * L > tab_blk1_ * sc.B || (P.k_ref + 1) * L > sc.total) return false;

    cudaStream_t st = current_stream();
    const uint32_t nblk = (uint32_t)(tab_blk1_ - tab_blk0_);
    P.counters.reserve(2);
    R4WB_CUDA(cudaMemsetAsync(P.counters.p, 0, 2 * sizeof(uint32_t), st));
    launch_static_check(d_tab_.p, d_sat_.p, nblk, sc.n_sats, sc.B, P.counters.p, st);
    P.ys.reserve((size_t)P.ns * L); P.yb.reserve((size_t)P.ns * L);
    P.sat.reserve(P.ns); P.cands.reserve(kPerMaxCands);
    R4WB_CUDA(cudaMemsetAsync(P.ys.p, 0, (size_t)P.ns * L * sizeof(float), st));
    R4WB_CUDA(cudaMemsetAsync(P.yb.p, 0, (size_t)P.ns * L * sizeof(float), st));
    R4WB_CUDA(cudaMemsetAsync(P.sat.p, 0xff, (size_t)P.ns * sizeof(PerSat), st));
    PeriodTableArgs ta{};
    ta.tab = d_tab_.p; ta.perbits = d_perbits_.p; ta.satcode = d_satcode_.p; ta.taps = d_taps_.p;
    ta.ys = P.ys.p; ta.yb = P.yb.p; ta.sat = P.sat.p; ta.cands = P.cands.p; ta.n_cands = P.counters.p + 1;
    ta.tab_blk0 = tab_blk0_; ta.B = sc.B; ta.k_ref = P.k_ref; ta.delta46 = sc.delta46;
    ta.n_sats = sc.n_sats; ta.L = P.L; ta.tile_len = P.tile_len;
    launch_period_tables(ta, P.ns, st);
    uint32_t h[2] = {1u, 0u};
    R4WB_CUDA(cudaMemcpyAsync(h, P.counters.p, sizeof h, cudaMemcpyDeviceToHost, st));
    R4WB_CUDA(cudaStreamSynchronize(st));
    P.n_cands = h[1];
    P.ok = h[0] == 0 && h[1] <= (uint32_t)kPerMaxCands;
    return P.ok;
}

// Whole primary-code periods inside [first, first + n) through the period-resident kernels, the partial periods at
// either end (and the first period of the run, whose FIR starts from a zero delay line) through k_synth on a side stream.
bool Scenario::render_periodic(uint64_t first, uint64_t n, void* d_out)
{
    if (!plan_periodic()) return false;
    PeriodicState& P = *per_;
    const ScenConst& sc = md_.sc;
    const uint64_t L = P.L;
    const uint64_t k_lo = std::max<uint64_t>(1, (first + L - 1) / L), k_hi = (first + n) / L;
    if (k_hi < k_lo + 32) return false;
    const uint64_t head = k_lo * L - first;                               // samples before the first whole period
    float2* out_fast = reinterpret_cast<float2*>(d_out) + head;
    if (head % kPerSlotsPerThread != 0 || (uintptr_t)out_fast % 32 != 0) return false;
    if (k_lo * L / sc.B < tab_blk0_ + (tab_blk0_ > 0 ? 1 : 0)) return false;

    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        R4WB_CUDA(cudaGetDevice(&dev));
        R4WB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    cudaStream_t st = current_stream();
    ensure_side_stream();
    const uint32_t n_periods = (uint32_t)(k_hi - k_lo);

    PeriodicArgs a{};
    a.ys = P.ys.p; a.yb = P.yb.p; a.sat = P.sat.p; a.cands = P.cands.p; a.n_cands = P.counters.p + 1;
    a.tab = d_tab_.p; a.satcode = d_satcode_.p;
    a.out = out_fast; a.power_sum = d_power_.p;
    a.k0 = k_lo; a.k_ref = P.k_ref; a.tab_blk0 = tab_blk0_; a.B = sc.B;
    a.n_periods = n_periods; a.L = P.L; a.tile_len = P.tile_len; a.n_tiles = P.n_tiles;
    a.n_sats = sc.n_sats; a.flags = sc.flags; a.noise_std = sc.noise_std; a.seed = sc.seed;
    // periods per CTA: about 64, adjusted so that the CTAs fill whole waves of 2 per SM
    {
        const uint64_t slots = (uint64_t)sm_count * 2;
        const uint64_t items0 = (uint64_t)P.n_tiles * ((n_periods + 63) / 64);
        const uint64_t waves = std::max<uint64_t>(1, (items0 + slots / 2) / slots);
        uint64_t n_chunks = std::max<uint64_t>(1, waves * slots / P.n_tiles);
        uint64_t KI = (n_periods + n_chunks - 1) / n_chunks;
        if (KI > 96) KI = 96;
        if (KI < 8) KI = std::min<uint64_t>(8, n_periods);
        a.KI = (uint32_t)KI;
        a.n_chunks = (uint32_t)((n_periods + KI - 1) / KI);
    }
    if (P.T_n != n_periods || P.T_k0 != k_lo) {
        P.T.reserve((size_t)n_periods * P.ns * 2);
        launch_period_phasors(a, P.ns, tab_blk1_, P.T.p, st);
        P.T_n = n_periods; P.T_k0 = k_lo;
    }
    a.T = P.T.p;

    // the periodic kernels first (the GPU starts on the bulk while the host enqueues the rest); the partial periods at the
    // ends go through the general kernel on the side stream, forked before and joined after
    const uint64_t tail_first = k_hi * L, tail_n = first + n - tail_first;
    if (head > 0 || tail_n > 0) R4WB_CUDA(cudaEventRecord(ev_fork_, st));
    prof_begin(1, st);
    launch_synth_periodic(a, P.ns, st);
    prof_end(st);
    prof_begin(2, st);
    launch_periodic_fix(a, P.n_cands, st);
    prof_end(st);
    if (head > 0 || tail_n > 0) {
        R4WB_CUDA(cudaStreamWaitEvent(side_stream_, ev_fork_, 0));
        if (head > 0) {
            const uint64_t hb0 = first / sc.B, hb1 = (first + head - 1) / sc.B;
            launch_synth(d_tab_.p, d_hdr_.p, d_tiles_.p, (uint32_t)(hb0 - tab_blk0_), (uint32_t)(hb1 - hb0 + 1), first, head, d_out, R4WB_FMT_CF32, sc.B, side_stream_);
        }
        if (tail_n > 0) {
            const uint64_t tb0 = tail_first / sc.B, tb1 = (first + n - 1) / sc.B;
            launch_synth(d_tab_.p, d_hdr_.p, d_tiles_.p, (uint32_t)(tb0 - tab_blk0_), (uint32_t)(tb1 - tb0 + 1), tail_first, tail_n,
                         reinterpret_cast<float2*>(d_out) + (tail_first - first), R4WB_FMT_CF32, sc.B, side_stream_);
        }
        R4WB_CUDA(cudaEventRecord(ev_join_, side_stream_));
    }
    if (head > 0 || tail_n > 0) R4WB_CUDA(cudaStreamWaitEvent(st, ev_join_, 0));
    return true;
}

void Scenario::render_device(uint64_t first, uint64_t n, void* d_out, r4wb_fmt fmt)
{
    const ScenConst& sc = md_.sc;
    if (fmt == R4WB_FMT_CF32 && render_periodic(first, n, d_out)) { last_path_ = 1; return; }
    last_path_ = 0;
    const uint64_t b0 = first / sc.B, b1 = (first + n - 1) / sc.B;
    launch_synth(d_tab_.p, d_hdr_.p, d_tiles_.p, (uint32_t)(b0 - tab_blk0_), (uint32_t)(b1 - b0 + 1), first, n, d_out, fmt, sc.B, current_stream());
}

void Scenario::render_to(uint64_t first, uint64_t n, void* dst, r4wb_mem where, r4wb_fmt fmt)
{
    const ScenConst& sc = md_.sc;
    cudaStream_t st = current_stream();
    const size_t bps = fmt_bytes(fmt);
    R4WB_CUDA(cudaMemsetAsync(d_power_.p, 0, sizeof(double), st));
    timed_.clear();

    const uint64_t b0 = first / sc.B, b1 = (first + n - 1) / sc.B;
    build_canonical_table(md_.table_begin(b0), b1 + 1, b0);

    if (where == R4WB_MEM_DEVICE) {
        render_device(first, n, dst, fmt);
        return;
    }
    // host destination: render chunk by chunk into two device staging buffers; the copy of one chunk (side stream)
    // overlaps the rendering of the next
    ensure_side_stream();
    const uint64_t chunk_blocks = std::max<uint64_t>(1, (uint64_t)(16u << 20) / sc.B);   // ~16 Msamples per chunk
    uint64_t bpp = 1;                                                              // blocks per primary-code period
    if (fmt == R4WB_FMT_CF32 && plan_periodic() && per_->L % sc.B == 0) bpp = per_->L / sc.B;
    ensure_side_stream();
    const size_t stage_bytes = (size_t)std::min<uint64_t>(n + sc.B, chunk_blocks * sc.B + sc.B) * bps;
    unsigned char* stage[2] = {d_stage_.reserve(stage_bytes), d_stage2_.reserve(stage_bytes)};
    uint32_t c = 0;
    for (uint64_t cb = b0, ce; cb <= b1; cb = ce, ++c) {
        // chunk edges on primary-code period boundaries (absolute multiples of bpp blocks), so that only the first and the
        // last chunk of a range carry a partial period
        ce = (cb + chunk_blocks) / bpp * bpp;
        if (ce <= cb) ce = cb + chunk_blocks;
        ce = std::min(b1 + 1, ce);
        const uint64_t f = std::max(first, cb * sc.B), l = std::min(first + n, ce * sc.B);
        const uint32_t k = c & 1u;
        if (c >= 2) R4WB_CUDA(cudaStreamWaitEvent(st, ev_copy_[k], 0));          // the buffer's previous copy is done
        render_device(f, l - f, stage[k], fmt);
        R4WB_CUDA(cudaEventRecord(ev_render_[k], st));
        R4WB_CUDA(cudaStreamWaitEvent(copy_stream_, ev_render_[k], 0));
        R4WB_CUDA(cudaMemcpyAsync((unsigned char*)dst + (f - first) * bps, stage[k], (l - f) * bps, cudaMemcpyDeviceToHost, copy_stream_));
        R4WB_CUDA(cudaEventRecord(ev_copy_[k], copy_stream_));
    }
    R4WB_CUDA(cudaStreamSynchronize(copy_stream_));
    R4WB_CUDA(cudaStreamSynchronize(st));
}

void Scenario::generate(uint64_t first, uint64_t n, void* dst, r4wb_mem where, r4wb_fmt fmt)
{
    const ScenConst& sc = md_.sc;
    if (first > sc.total || n > sc.total - first) fail(R4WB_ERR_INVALID_SIZE, "range [%llu, +%llu) exceeds total_samples %llu",
                                                       (unsigned long long)first, (unsigned long long)n, (unsigned long long)sc.total);
    if ((unsigned)fmt > (unsigned)R4WB_FMT_CU8) fail(R4WB_ERR_INVALID_PARAMETER, "unknown sample format %d", (int)fmt);
    if (n == 0) return;
    if (!dst) fail(R4WB_ERR_NULL_POINTER, "dst is NULL");
    StreamScope scope(*this);
    render_to(first, n, dst, where, fmt);
}

void Scenario::ring_drop()
{
    if (!ring_) return;
    cudaStreamSynchronize(current_stream());
    if (copy_stream_) cudaStreamSynchronize(copy_stream_);
    delete ring_;
    ring_ = nullptr;
    last_block_ = nullptr;
}

// chunk j = canonical blocks [j C, (j + 1) C) -> device staging buffer j & 1 -> pinned slot j % kSlots
void Scenario::ring_schedule(uint64_t j)
{
    BlockRing& R = *ring_;
    const ScenConst& sc = md_.sc;
    cudaStream_t st = current_stream();
    const uint32_t slot = (uint32_t)(j % BlockRing::kSlots), sb = (uint32_t)(j & 1u);
    const uint64_t first = j * R.C * sc.B, end = std::min<uint64_t>(sc.total, (j + 1) * R.C * sc.B);
    // the staging buffer's previous copy (chunk j - 2, slot (j - 2) % kSlots) must have left the device
    if (j >= 2) R4WB_CUDA(cudaStreamWaitEvent(st, R.ev_done[(j - 2) % BlockRing::kSlots], 0));
    render_device(first, end - first, R.stage[sb].p, R.fmt);
    R4WB_CUDA(cudaEventRecord(R.ev_render[sb], st));
    R4WB_CUDA(cudaStreamWaitEvent(copy_stream_, R.ev_render[sb], 0));
    R4WB_CUDA(cudaMemcpyAsync(R.pin + (size_t)slot * R.chunk_bytes, R.stage[sb].p, (size_t)(end - first) * fmt_bytes(R.fmt), cudaMemcpyDeviceToHost, copy_stream_));
    R4WB_CUDA(cudaEventRecord(R.ev_done[slot], copy_stream_));
    R.chunk_of_slot[slot] = j;
}

uint64_t Scenario::ring_block(uint64_t n, void* dst, r4wb_fmt fmt)
{
    const ScenConst& sc = md_.sc;
    const size_t bps = fmt_bytes(fmt);
    if (ring_ && ring_->fmt != fmt) ring_drop();
    if (!ring_) {
        StreamScope scope(*this);
        ensure_side_stream();
        build_canonical_table(0, md_.n_blocks());
        ring_ = new BlockRing;
        BlockRing& R = *ring_;
        R.fmt = fmt;
        // ~128 Ksamples (1 MB of cf32) per chunk: three chunks in flight stay inside the last-level cache the DMA writes
        // allocate in, so the caller's memcpy reads them from cache (measured: 1 Msample chunks 4.0 us per call, 80-160 K 3.1-3.4)
        uint64_t chunk_samples = 1u << 17;
        if (const char* e = std::getenv("R4WB_RING_CHUNK")) { const long v = std::atol(e); if (v >= 1024) chunk_samples = (uint64_t)v; }   // tuning hook
        R.C = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(md_.n_blocks(), chunk_samples / sc.B));
        R.chunk_bytes = (size_t)R.C * sc.B * bps;
        R.n_chunks = (md_.n_blocks() + R.C - 1) / R.C;
        R4WB_CUDA(cudaMallocHost((void**)&R.pin, R.chunk_bytes * BlockRing::kSlots));
        for (auto& b : R.stage) b.reserve(R.chunk_bytes);
        for (auto& e : R.ev_done) R4WB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& e : R.ev_render) R4WB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& c : R.chunk_of_slot) c = ~0ull;
        R.next_chunk = (current_ / sc.B) / R.C;
    }
    BlockRing& R = *ring_;
    const uint64_t blk = current_ / sc.B, c = blk / R.C;
    const uint32_t slot = (uint32_t)(c % BlockRing::kSlots);
    if (R.synced_chunk != c) {
        // keep kSlots - 1 chunks in flight ahead of the one being consumed (the slot of chunk c - 1 is free by now)
        if (R.chunk_of_slot[slot] != c && (R.next_chunk > c + BlockRing::kSlots - 1 || R.next_chunk < c)) R.next_chunk = c;   // repositioned
        StreamScope scope(*this);
        while (R.next_chunk < R.n_chunks && R.next_chunk < c + BlockRing::kSlots - 1) ring_schedule(R.next_chunk++);
        if (R.chunk_of_slot[slot] != c) fail(R4WB_ERR_CUDA, "block ring lost chunk %llu", (unsigned long long)c);
        R4WB_CUDA(cudaEventSynchronize(R.ev_done[slot]));
        R.synced_chunk = c;
    }
    const unsigned char* src = R.pin + (size_t)slot * R.chunk_bytes + (size_t)(blk - c * R.C) * sc.B * bps;
    if (dst) std::memcpy(dst, src, (size_t)n * bps);   // dst == NULL: generate_block_view hands out `src` itself
    last_block_ = src; last_block_n_ = n; last_block_fmt_ = fmt;
    current_ += n;
    return n;
}

// The host-side SeqState lags behind while the ring serves canonical blocks; an odd-sized block afterwards continues the
// reference's own partition, so the state is replayed block by block up to current_ (host f64 walk: ~0.3 ms per block)
void Scenario::seq_sync()
{
    const ScenConst& sc = md_.sc;
    std::vector<BlockSat> tab;
    BlockHdr hdr[2];
    while (seq_pos_ < current_) {
        const uint32_t n = (uint32_t)std::min<uint64_t>(sc.B, current_ - seq_pos_);
        seq_.make_table(md_, seq_pos_, n, tab, hdr);
        seq_.advance(md_, tab, n);
        seq_pos_ += n;
    }
}

static bool ring_enabled()
{
    const char* e = std::getenv("R4WB_BLOCK_RING");          // A/B hook: 0 renders every block on demand
    return !(e && e[0] == '0');
}

uint64_t Scenario::generate_block(uint64_t n_req, void* dst, r4wb_mem where, r4wb_fmt fmt)
{
    const ScenConst& sc = md_.sc;
    const uint64_t remaining = sc.total > current_ ? sc.total - current_ : 0;
    const uint64_t n = std::min(remaining, n_req);
    if (n == 0) return 0;
    if (!dst) fail(R4WB_ERR_NULL_POINTER, "dst is NULL");
    if (n > (1ull << 24)) fail(R4WB_ERR_NOT_SUPPORTED, "generate_block: at most 2^24 samples per reference block");
    if ((unsigned)fmt > (unsigned)R4WB_FMT_CU8) fail(R4WB_ERR_INVALID_PARAMETER, "unknown sample format %d", (int)fmt);
    // the CLI loop (canonical block size, host destination): served from the render-ahead ring
    if (where == R4WB_MEM_HOST && seq_canonical_ && current_ % sc.B == 0 && (n == sc.B || n == remaining) && n <= sc.B && ring_enabled())
        return ring_block(n, dst, fmt);
    ring_drop();
    seq_sync();
    StreamScope scope(*this);
    cudaStream_t st = current_stream();
    std::vector<BlockSat> tab;
    BlockHdr hdr[2];
    seq_.make_table(md_, current_, (uint32_t)n, tab, hdr);
    R4WB_CUDA(cudaMemcpyAsync(d_seq_tab_.reserve(tab.size()), tab.data(), tab.size() * sizeof(BlockSat), cudaMemcpyHostToDevice, st));
    R4WB_CUDA(cudaMemcpyAsync(d_seq_hdr_.reserve(2), hdr, sizeof(BlockHdr) * 2, cudaMemcpyHostToDevice, st));
    R4WB_CUDA(cudaMemsetAsync(d_power_.p, 0, sizeof(double), st));
    const size_t bps = fmt_bytes(fmt);
    void* d_out = where == R4WB_MEM_DEVICE ? dst : (void*)d_stage_.reserve((size_t)n * bps);
    {
        SynthArgs a = base_args(d_seq_tab_.p, d_seq_hdr_.p, n);
        d_seq_tiles_.reserve(std::max<size_t>(1, (size_t)2 * a.tiles_per_block * sc.n_sats));
        build_tiles(a, 1, 1, d_seq_tiles_.p);
    }
    launch_synth(d_seq_tab_.p, d_seq_hdr_.p, d_seq_tiles_.p, 1, 1, current_, n, d_out, fmt, n, st);
    if (where != R4WB_MEM_DEVICE) R4WB_CUDA(cudaMemcpyAsync(dst, d_out, (size_t)n * bps, cudaMemcpyDeviceToHost, st));
    R4WB_CUDA(cudaStreamSynchronize(st));   // `tab`/`hdr` are stack/heap temporaries
    seq_.advance(md_, tab, (uint32_t)n);
    if (current_ % sc.B != 0 || (n != sc.B && n != remaining)) seq_canonical_ = false;
    current_ += n;
    seq_pos_ = current_;
    last_block_ = nullptr;
    return n;
}

// Borrowed form of generate_block: the ring's pinned chunk is the caller's block (no host copy).  A chunk's slot is rewritten
// only after the consumer has moved two chunks on, so the pointer outlives the next call; the contract says "until the next
// call".  Blocks the ring does not serve (odd sizes) are rendered into a pinned bounce buffer of the handle.
const void* Scenario::generate_block_view(uint64_t n_req, r4wb_fmt fmt, uint64_t* n_out)
{
    const ScenConst& sc = md_.sc;
    const uint64_t remaining = sc.total > current_ ? sc.total - current_ : 0;
    const uint64_t n = std::min(remaining, n_req);
    *n_out = n;
    if (n == 0) return nullptr;
    if ((unsigned)fmt > (unsigned)R4WB_FMT_CU8) fail(R4WB_ERR_INVALID_PARAMETER, "unknown sample format %d", (int)fmt);
    if (seq_canonical_ && current_ % sc.B == 0 && (n == sc.B || n == remaining) && n <= sc.B && ring_enabled()) {
        ring_block(n, nullptr, fmt);
        return last_block_;
    }
    if (n > (1ull << 24)) fail(R4WB_ERR_NOT_SUPPORTED, "generate_block: at most 2^24 samples per reference block");
    const size_t need = (size_t)n * fmt_bytes(fmt);
    if (need > view_cap_) {
        if (view_buf_) cudaFreeHost(view_buf_);
        view_buf_ = nullptr; view_cap_ = 0;
        R4WB_CUDA(cudaMallocHost(&view_buf_, need));
        view_cap_ = need;
    }
    generate_block(n, view_buf_, R4WB_MEM_HOST, fmt);
    return view_buf_;
}

uint64_t Scenario::generate_rest(void* dst, uint64_t cap, r4wb_mem where, r4wb_fmt fmt)
{
    const ScenConst& sc = md_.sc;
    const uint64_t remaining = sc.total > current_ ? sc.total - current_ : 0;
    if (remaining == 0) return 0;
    if (!dst) fail(R4WB_ERR_NULL_POINTER, "dst is NULL");
    if (cap < remaining) fail(R4WB_ERR_INVALID_SIZE, "dst holds %llu samples, %llu remain", (unsigned long long)cap, (unsigned long long)remaining);
    if (seq_canonical_ && current_ % sc.B == 0) {
        generate(current_, remaining, dst, where, fmt);
        current_ = sc.total;
        return remaining;
    }
    // odd-sized blocks came before: the reference's partition continues from current_sample in block_size() steps
    const size_t bps = fmt_bytes(fmt);
    uint64_t done = 0;
    while (current_ < sc.total) done += generate_block(sc.B, static_cast<unsigned char*>(dst) + done * bps, where, fmt);
    return done;
}

// File sink: segments of ~16 Msamples are rendered into two pinned host buffers (render_to's own double-buffered D2H); a
// writer drains one buffer into the file (kSinkWriters threads, one pwrite range each) while the next segment renders and
// copies into the other.
namespace {
constexpr int kSinkWriters = 4;

bool pwrite_all(int fd, const unsigned char* p, size_t n, off_t at)
{
    while (n > 0) {
        const ssize_t w = ::pwrite(fd, p, std::min<size_t>(n, (size_t)64 << 20), at);
        if (w < 0) { if (errno == EINTR) continue; return false; }
        if (w == 0) return false;
        p += w; n -= (size_t)w; at += w;
    }
    return true;
}
}  // namespace

double Scenario::write_file(const char* path, r4wb_fmt fmt, uint64_t* samples, uint64_t* bytes)
{
    const ScenConst& sc = md_.sc;
    if ((unsigned)fmt > (unsigned)R4WB_FMT_CU8) fail(R4WB_ERR_INVALID_PARAMETER, "unknown sample format %d", (int)fmt);
    const int fd = ::open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) fail(R4WB_ERR_INVALID_PARAMETER, "cannot create '%s': %s", path, std::strerror(errno));
    const size_t bps = fmt_bytes(fmt);
    StreamScope scope(*this);
    uint64_t unit = sc.B;                                                          // segment edges on period boundaries
    if (fmt == R4WB_FMT_CF32 && plan_periodic() && per_->L % sc.B == 0) unit = per_->L;
    const uint64_t seg = std::max<uint64_t>(1, (uint64_t)(16u << 20) / unit) * unit;
    const uint64_t cap = std::min<uint64_t>(seg, std::max<uint64_t>(sc.total, 1));
    unsigned char* pin[2] = {nullptr, nullptr};
    struct Job { const unsigned char* p; size_t n; off_t at; };
    std::mutex mu;
    std::condition_variable cv;
    Job job{nullptr, 0, 0};
    bool busy = false, quit = false, io_failed = false;
    std::thread writer([&] {
        for (;;) {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return busy || quit; });
            if (!busy) return;
            const Job j = job;
            lk.unlock();
            bool ok[kSinkWriters];
            std::thread part[kSinkWriters];
            const size_t each = ((j.n + kSinkWriters - 1) / kSinkWriters + 4095) / 4096 * 4096;
            for (int t = 0; t < kSinkWriters; ++t) {
                const size_t lo = std::min(j.n, each * t), hi = std::min(j.n, each * (t + 1));
                ok[t] = true;
                part[t] = std::thread([&, t, lo, hi] { if (hi > lo) ok[t] = pwrite_all(fd, j.p + lo, hi - lo, j.at + (off_t)lo); });
            }
            bool all = true;
            for (int t = 0; t < kSinkWriters; ++t) { part[t].join(); all = all && ok[t]; }
            lk.lock();
            if (!all) io_failed = true;
            busy = false;
            cv.notify_all();
        }
    });
    auto drain = [&] { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return !busy; }); };
    auto finish = [&] {
        drain();
        { std::lock_guard<std::mutex> lk(mu); quit = true; }
        cv.notify_all();
        writer.join();
        for (auto*& p : pin) if (p) { cudaFreeHost(p); p = nullptr; }
    };
    double power = 0.0;
    uint64_t done = 0;
    try {
        // the block table of the whole file once, up front: a table that only ever grows from block 0 (dynamic scenarios) would
        // otherwise be rebuilt for every segment — prologue work quadratic in the file length
        if (sc.total > 0) build_canonical_table(md_.table_begin(0), md_.n_blocks());
        for (auto*& p : pin) R4WB_CUDA(cudaMallocHost((void**)&p, (size_t)cap * bps));
        for (uint32_t c = 0; done < sc.total; ++c) {
            const uint64_t n = std::min(seg, sc.total - done);
            unsigned char* dst = pin[c & 1u];
            render_to(done, n, dst, R4WB_MEM_HOST, fmt);                           // the other buffer may still be draining
            power += last_power_sum();
            drain();
            if (io_failed) break;
            { std::lock_guard<std::mutex> lk(mu); job = Job{dst, (size_t)n * bps, (off_t)(done * bps)}; busy = true; }
            cv.notify_all();
            done += n;
        }
    } catch (...) {
        finish();
        ::close(fd);
        throw;
    }
    finish();
    const bool closed = ::close(fd) == 0;
    if (io_failed || !closed) fail(R4WB_ERR_BUFFER_FULL, "short write to '%s'", path);
    current_ = sc.total;                                                           // the CLI loop leaves the scenario done
    if (samples) *samples = done;
    if (bytes) *bytes = done * bps;
    return power;
}

double Scenario::last_power_sum()
{
    if (last_block_) {                                   // ring-served block: sum |s|^2 of the samples handed out (float formats)
        double acc = 0.0;
        if (last_block_fmt_ == R4WB_FMT_CF32) {
            const float* p = static_cast<const float*>(last_block_);
            for (uint64_t i = 0; i < 2 * last_block_n_; ++i) acc += (double)p[i] * (double)p[i];
        } else if (last_block_fmt_ == R4WB_FMT_CF64) {
            const double* p = static_cast<const double*>(last_block_);
            for (uint64_t i = 0; i < 2 * last_block_n_; ++i) acc += p[i] * p[i];
        } else fail(R4WB_ERR_NOT_SUPPORTED, "power sum of a ring-served block is kept for the float formats only");
        return acc;
    }
    double v = 0.0;
    cudaStream_t st = current_stream();
    R4WB_CUDA(cudaMemcpyAsync(&v, d_power_.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    R4WB_CUDA(cudaStreamSynchronize(st));
    return v;
}

void Scenario::debug_block(uint64_t block, uint32_t sat, double* o)
{
    const ScenConst& sc = md_.sc;
    if (block >= md_.n_blocks() || sat >= sc.n_sats) fail(R4WB_ERR_INVALID_PARAMETER, "block/sat out of range");
    StreamScope scope(*this);
    build_canonical_table(md_.table_begin(block), block + 1);
    BlockSat e;
    cudaStream_t st = current_stream();
    R4WB_CUDA(cudaMemcpyAsync(&e, d_tab_.p + (size_t)(block - tab_blk0_) * sc.n_sats + sat, sizeof e, cudaMemcpyDeviceToHost, st));
    R4WB_CUDA(cudaStreamSynchronize(st));
    block_sat_debug(e, o);
}

}  // namespace r4wb
