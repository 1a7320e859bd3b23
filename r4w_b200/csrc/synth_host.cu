// synth_host.cu — device side of GnssScenario (gnss/scenario.rs:51-705): owns the HBM-resident tables and
// launches the prologue + synthesis kernels.  The host-only model lives in synth_model.cpp.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "synth_math.cuh"

namespace r4wb {

// launchers (synth_kernels.cu)
void launch_synth_kernel(const SynthArgs& a, int K, r4wb_fmt fmt, int grid, cudaStream_t st);
int synth_max_blocks_per_sm(int K, r4wb_fmt fmt, size_t smem);
void launch_block_params(const ScenConst&, const SatConst*, const PhaseSegment*, uint64_t, uint32_t, BlockSat*, BlockHdr*, cudaStream_t);
void launch_phase_scan(const SatConst*, uint32_t, uint32_t, BlockSat*, cudaStream_t);
void launch_tile_params(const SynthArgs& a, uint32_t tb_begin, uint32_t tb_count, uint32_t tile_samples, TileRec* out, cudaStream_t st);

Scenario::Scenario(const r4wb_scenario_cfg& cfg) : md_(cfg)
{
    seq_.reset(md_.sc.n_sats);
    cudaStream_t st = current_stream();
    const auto& sats = md_.sats;
    R4WB_CUDA(cudaMemcpyAsync(d_sat_.reserve(std::max<size_t>(1, sats.size())), sats.data(), sats.size() * sizeof(SatConst), cudaMemcpyHostToDevice, st));
    R4WB_CUDA(cudaMemcpyAsync(d_segments_.reserve(md_.segments.size()), md_.segments.data(), md_.segments.size() * sizeof(PhaseSegment), cudaMemcpyHostToDevice, st));
    R4WB_CUDA(cudaMemcpyAsync(d_perbits_.reserve(md_.perbits.size()), md_.perbits.data(), md_.perbits.size() * 4, cudaMemcpyHostToDevice, st));
    R4WB_CUDA(cudaMemcpyAsync(d_satcode_.reserve(md_.satcode.size()), md_.satcode.data(), md_.satcode.size() * sizeof(SatCode), cudaMemcpyHostToDevice, st));
    R4WB_CUDA(cudaMemcpyAsync(d_ytab_.reserve(md_.ytab.size()), md_.ytab.data(), md_.ytab.size() * 4, cudaMemcpyHostToDevice, st));
    R4WB_CUDA(cudaMemcpyAsync(d_taps_.reserve(64), md_.taps_f, sizeof md_.taps_f, cudaMemcpyHostToDevice, st));
    R4WB_CUDA(cudaMemcpyAsync(d_etab_.reserve(64), md_.etab_f, sizeof md_.etab_f, cudaMemcpyHostToDevice, st));
    R4WB_CUDA(cudaMemcpyAsync(d_clslut_.reserve(md_.clslut.size()), md_.clslut.data(), md_.clslut.size(), cudaMemcpyHostToDevice, st));
    d_power_.reserve(1);
    R4WB_CUDA(cudaMemsetAsync(d_power_.p, 0, sizeof(double), st));
    R4WB_CUDA(cudaStreamSynchronize(st));
}

Scenario::~Scenario() {}

void Scenario::reset()
{
    current_ = 0;
    seq_.reset(md_.sc.n_sats);
}

// Prologue: Phase-1 parameters + NCO start values of canonical blocks [blk_begin, blk_end) into d_tab_.
// The table is kept between calls (a bench loop re-rendering the same range pays for it once).
void Scenario::build_canonical_table(uint64_t blk_begin, uint64_t blk_end)
{
    if (tab_valid_ && blk_begin >= tab_blk0_ && blk_end <= tab_blk1_ && (blk_begin == tab_blk0_ || !(md_.any_dynamic || md_.any_var_visibility)))
        return;
    const ScenConst& sc = md_.sc;
    const uint64_t nblk = blk_end - blk_begin;
    if (nblk > 0x7fffffffull / std::max(1u, sc.n_sats)) fail(R4WB_ERR_INVALID_SIZE, "too many blocks in one call");
    cudaStream_t st = current_stream();
    tab_valid_ = false;
    d_tab_.reserve(std::max<size_t>(1, (size_t)nblk * sc.n_sats));
    d_hdr_.reserve(std::max<size_t>(1, nblk));
    launch_block_params(sc, d_sat_.p, d_segments_.p, blk_begin, (uint32_t)nblk, d_tab_.p, d_hdr_.p, st);
    if (sc.n_sats == 0) {   // headers still needed
        std::vector<BlockHdr> h(nblk);
        for (uint64_t b = 0; b < nblk; ++b) {
            const uint64_t first = (blk_begin + b) * sc.B;
            h[b] = BlockHdr{first, (uint32_t)std::min<uint64_t>(sc.B, sc.total - first), 0};
        }
        R4WB_CUDA(cudaMemcpyAsync(d_hdr_.p, h.data(), nblk * sizeof(BlockHdr), cudaMemcpyHostToDevice, st));
        R4WB_CUDA(cudaStreamSynchronize(st));
    }
    if (md_.any_dynamic || md_.any_var_visibility) launch_phase_scan(d_sat_.p, sc.n_sats, (uint32_t)nblk, d_tab_.p, st);
    tab_blk0_ = blk_begin;
    tab_blk1_ = blk_end;
    tab_valid_ = true;
    // per-tile records for the canonical tiling (block size B)
    {
        SynthArgs a = base_args(d_tab_.p, d_hdr_.p, sc.B);
        d_tiles_.reserve(std::max<size_t>(1, (size_t)nblk * a.tiles_per_block * sc.n_sats));
        build_tiles(a, 0, (uint32_t)nblk, d_tiles_.p);
    }
}

SynthArgs Scenario::base_args(const BlockSat* tab, const BlockHdr* hdr, uint64_t max_block_n) const
{
    const ScenConst& sc = md_.sc;
    SynthArgs a{};
    a.tab = tab; a.hdr = hdr; a.perbits = d_perbits_.p; a.satcode = d_satcode_.p; a.taps = d_taps_.p; a.etab = d_etab_.p; a.ytab = d_ytab_.p; a.clslut = d_clslut_.p; a.lut_den = sc.lut_den;
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
                            uint64_t out_first, uint64_t out_n, void* d_out, r4wb_fmt fmt, uint64_t max_block_n)
{
    SynthArgs a = base_args(tab, hdr, max_block_n);
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
    const int tile_k = md_.tile_k;
    const size_t smem = synth_smem_bytes(a.n_sats, a.nw64, a.lut_den);
    const int per_sm = std::max(1, synth_max_blocks_per_sm(tile_k, fmt, smem));
    const uint64_t n_tiles = (uint64_t)tb_count * a.tiles_per_block;
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(n_tiles, (uint64_t)sm_count * per_sm));
    launch_synth_kernel(a, tile_k, fmt, grid, current_stream());
}

void Scenario::render_to(uint64_t first, uint64_t n, void* dst, r4wb_mem where, r4wb_fmt fmt)
{
    const ScenConst& sc = md_.sc;
    cudaStream_t st = current_stream();
    const size_t bps = fmt_bytes(fmt);
    R4WB_CUDA(cudaMemsetAsync(d_power_.p, 0, sizeof(double), st));

    const uint64_t b0 = first / sc.B, b1 = (first + n - 1) / sc.B;
    build_canonical_table(md_.table_begin(b0), b1 + 1);
    const uint64_t tb0 = tab_blk0_;

    if (where == R4WB_MEM_DEVICE) {
        launch_synth(d_tab_.p, d_hdr_.p, d_tiles_.p, (uint32_t)(b0 - tb0), (uint32_t)(b1 - b0 + 1), first, n, dst, fmt, sc.B);
        return;
    }
    // host destination: render chunk by chunk into a device staging buffer and copy out
    const uint64_t chunk_blocks = std::max<uint64_t>(1, (uint64_t)(32u << 20) / sc.B);   // ~32 Msamples per chunk
    unsigned char* stage = d_stage_.reserve((size_t)std::min<uint64_t>(n + sc.B, chunk_blocks * sc.B + sc.B) * bps);
    for (uint64_t cb = b0; cb <= b1; cb += chunk_blocks) {
        const uint64_t ce = std::min(b1 + 1, cb + chunk_blocks);
        const uint64_t f = std::max(first, cb * sc.B), l = std::min(first + n, ce * sc.B);
        launch_synth(d_tab_.p, d_hdr_.p, d_tiles_.p, (uint32_t)(cb - tb0), (uint32_t)(ce - cb), f, l - f, stage, fmt, sc.B);
        R4WB_CUDA(cudaMemcpyAsync((unsigned char*)dst + (f - first) * bps, stage, (l - f) * bps, cudaMemcpyDeviceToHost, st));
        R4WB_CUDA(cudaStreamSynchronize(st));
    }
}

void Scenario::generate(uint64_t first, uint64_t n, void* dst, r4wb_mem where, r4wb_fmt fmt)
{
    const ScenConst& sc = md_.sc;
    if (first > sc.total || n > sc.total - first) fail(R4WB_ERR_INVALID_SIZE, "range [%llu, +%llu) exceeds total_samples %llu",
                                                       (unsigned long long)first, (unsigned long long)n, (unsigned long long)sc.total);
    if ((unsigned)fmt > (unsigned)R4WB_FMT_CU8) fail(R4WB_ERR_INVALID_PARAMETER, "unknown sample format %d", (int)fmt);
    if (n == 0) return;
    if (!dst) fail(R4WB_ERR_NULL_POINTER, "dst is NULL");
    render_to(first, n, dst, where, fmt);
}

uint64_t Scenario::generate_block(uint64_t n_req, void* dst, r4wb_mem where, r4wb_fmt fmt)
{
    const ScenConst& sc = md_.sc;
    const uint64_t remaining = sc.total > current_ ? sc.total - current_ : 0;
    const uint64_t n = std::min(remaining, n_req);
    if (n == 0) return 0;
    if (!dst) fail(R4WB_ERR_NULL_POINTER, "dst is NULL");
    if (n > 65536) fail(R4WB_ERR_NOT_SUPPORTED, "generate_block: at most 65536 samples per reference block");
    if ((unsigned)fmt > (unsigned)R4WB_FMT_CU8) fail(R4WB_ERR_INVALID_PARAMETER, "unknown sample format %d", (int)fmt);
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
    launch_synth(d_seq_tab_.p, d_seq_hdr_.p, d_seq_tiles_.p, 1, 1, current_, n, d_out, fmt, n);
    if (where != R4WB_MEM_DEVICE) R4WB_CUDA(cudaMemcpyAsync(dst, d_out, (size_t)n * bps, cudaMemcpyDeviceToHost, st));
    R4WB_CUDA(cudaStreamSynchronize(st));   // `tab`/`hdr` are stack/heap temporaries
    seq_.advance(md_, tab, (uint32_t)n);
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

void Scenario::debug_block(uint64_t block, uint32_t sat, double* o)
{
    const ScenConst& sc = md_.sc;
    if (block >= md_.n_blocks() || sat >= sc.n_sats) fail(R4WB_ERR_INVALID_PARAMETER, "block/sat out of range");
    build_canonical_table(md_.table_begin(block), block + 1);
    BlockSat e;
    cudaStream_t st = current_stream();
    R4WB_CUDA(cudaMemcpyAsync(&e, d_tab_.p + (size_t)(block - tab_blk0_) * sc.n_sats + sat, sizeof e, cudaMemcpyDeviceToHost, st));
    R4WB_CUDA(cudaStreamSynchronize(st));
    block_sat_debug(e, o);
}

}  // namespace r4wb
