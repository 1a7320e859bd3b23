// synth_emu.cpp — HOST REPLAY of the synthesis kernels' per-sample arithmetic (r4w_b200/csrc/synth_math.cuh).
//
// TEST INFRASTRUCTURE ONLY.  This container has no GPU; this file lets `-m "not gpu"` tests check the device
// index arithmetic (code NCO, collapsed FIR, carrier NCO, Philox noise) against the oracle by calling the same
// __host__ __device__ functions the sm_100a kernels call, in plain loops that mirror k_block_params,
// k_phase_scan and k_synth.  It is never linked into or loaded by libr4w_b200.so / r4w_b200/.
#include <algorithm>
#include <cstring>
#include <memory>
#include <vector>

#include "../../r4w_b200/csrc/synth_lattice.cuh"

using namespace r4wb;

namespace {

struct Emu {
    ScenarioModel md;
    SeqState seq;
    uint64_t current = 0;
    std::vector<BlockSat> tab;
    std::vector<BlockHdr> hdr;
    uint64_t tab_blk0 = 0, tab_blk1 = 0;
    bool tab_valid = false;
    explicit Emu(const r4wb_scenario_cfg& c) : md(c) { seq.reset(md.sc.n_sats); }

    // k_block_params + k_phase_scan
    void build_table(uint64_t blk_begin, uint64_t blk_end)
    {
        if (tab_valid && blk_begin >= tab_blk0 && blk_end <= tab_blk1 && (blk_begin == tab_blk0 || !(md.any_dynamic || md.any_var_visibility))) return;
        const ScenConst& sc = md.sc;
        const uint32_t ns = sc.n_sats;
        const uint64_t nblk = blk_end - blk_begin;
        tab.assign((size_t)nblk * std::max(1u, ns), BlockSat{});
        hdr.assign(nblk, BlockHdr{});
        // k_phase_prefix / k_phase_q / k_phase_exact: the reference's sequential f64 carrier phase of dynamic satellites
        const bool exact_phase = md.any_dynamic && !(sc.flags & R4WB_FLAG_CLOSED_FORM_PHASE) && blk_begin == 0 && ns > 0;
        std::vector<double> dop(exact_phase ? 2 * (size_t)nblk * ns : 0);
        for (uint64_t tb = 0; tb < nblk; ++tb) {
            const uint64_t first = (blk_begin + tb) * sc.B;
            const uint64_t rem = sc.total - first;
            const uint32_t n = (uint32_t)(rem < sc.B ? rem : sc.B);
            hdr[tb] = BlockHdr{first, n, 0};
            for (uint32_t s = 0; s < ns; ++s) {
                BlockSat o;
                double d2[2];
                fill_block_sat(sc, md.sats[s], md.segments.data(), first, n, first, o, d2);
                o.prev = tb > 0 ? (int32_t)((tb - 1) * ns + s) : -1;
                if (!md.sats[s].static_phase) o.phi = (o.flags & 1u) ? block_advance(o) : 0ull;
                tab[tb * ns + s] = o;
                if (exact_phase) { dop[2 * (tb * ns + s)] = d2[0]; dop[2 * (tb * ns + s) + 1] = d2[1]; }
            }
        }
        if (md.any_dynamic || md.any_var_visibility) {
            for (uint32_t s = 0; s < ns; ++s) {
                const bool dynamic = !md.sats[s].static_phase;
                uint64_t run = 0;
                int prev = -1;
                for (uint64_t b = 0; b < nblk; ++b) {
                    BlockSat& e = tab[b * ns + s];
                    const uint64_t adv = e.phi;
                    if (dynamic) { e.phi = run; run += adv; }
                    e.prev = prev >= 0 ? (int32_t)((uint32_t)prev * ns + s) : -1;
                    if (e.flags & 1u) prev = (int)b;
                }
            }
        }
        if (exact_phase) {
            const bool steps_diag = std::getenv("R4WB_EMU_PHASE_STEPS") != nullptr;
            std::vector<PhaseQ> pq((size_t)nblk * ns);
            for (uint32_t s = 0; s < ns; ++s) {
                if (md.sats[s].static_phase) continue;
                double run = 0.0;                                        // k_phase_prefix
                std::vector<double> start(nblk);
                for (uint64_t b = 0; b < nblk; ++b) {
                    const BlockSat& e = tab[b * ns + s];
                    PhaseQ& r = pq[b * ns + s];
                    r = PhaseQ{};
                    start[b] = run;
                    if (e.flags & 1u) block_phase_approx(dop[2 * (b * ns + s)], dop[2 * (b * ns + s) + 1], e.n, sc.fs, &r.approx, &r.span);
                    run += r.approx;
                }
#pragma omp parallel for schedule(dynamic, 256)
                for (long long b = 0; b < (long long)nblk; ++b) {        // k_phase_q
                    const BlockSat& e = tab[b * ns + s];
                    PhaseQ& r = pq[b * ns + s];
                    if (!(e.flags & 1u) || start[b] == 0.0) continue;
                    r.k = ilogb(start[b]);
                    if (r.k < kPhaseMinBinade) continue;
                    bool tie;
                    // the replay sums every sample (the independent check of the device's step form); R4WB_EMU_PHASE_STEPS=1
                    // switches to the step form so that a whole 600 s table can be replayed for statistics
                    if (!(steps_diag && block_phase_q_steps(dop[2 * (b * ns + s)], dop[2 * (b * ns + s) + 1], e.n, sc.fs, 1.0 / sc.fs, 0, r.k, 256u, &r.Q, &tie)))
                        block_phase_q(dop[2 * (b * ns + s)], dop[2 * (b * ns + s) + 1], e.n, sc.fs, r.k, &r.Q, &tie);
                    r.ok = tie ? 0u : 1u;
                }
                if (steps_diag) {                                        // why blocks are walked (diagnostic)
                    uint64_t n_small = 0, n_tie = 0, n_pred = 0, n_vis = 0;
                    for (uint64_t b = 0; b < nblk; ++b) {
                        const BlockSat& e = tab[b * ns + s];
                        if (!(e.flags & 1u)) continue;
                        ++n_vis;
                        const PhaseQ& r = pq[b * ns + s];
                        if (start[b] == 0.0 || r.k < kPhaseMinBinade) ++n_small;
                        else if (!r.ok) ++n_tie;
                        else if (!phase_stays_in_binade(start[b], r)) ++n_pred;
                    }
                    std::fprintf(stderr, "[emu phase] sat %u: %llu visible blocks, walk: start below the first binade %llu, tie %llu, binade margin %llu\n", s,
                                 (unsigned long long)n_vis, (unsigned long long)n_small, (unsigned long long)n_tie, (unsigned long long)n_pred);
                }
                double ph = 0.0;                                         // k_phase_exact
                for (uint64_t b = 0; b < nblk; ++b) {
                    BlockSat& e = tab[b * ns + s];
                    e.phi = cycles_to_fixed(ph / (2.0 * kPi));
                    if (!(e.flags & 1u)) continue;
                    const size_t k = b * ns + s;
                    if (phase_stays_in_binade(ph, pq[k])) ph = ph + scalbn((double)pq[k].Q, pq[k].k - 52);
                    else ph = phase_walk(ph, dop[2 * k], dop[2 * k + 1], e.n, sc.fs);
                }
            }
        }
        tab_blk0 = blk_begin; tab_blk1 = blk_end; tab_valid = true;
    }

    // k_synth, thread loops flattened.  only_sat >= 0 renders that satellite alone.
    template <int K>
    void render_t(const BlockSat* tb_tab, const BlockHdr* tb_hdr, uint32_t tb_begin, uint32_t tb_count, uint64_t out_first,
                  uint64_t out_n, float* out, uint64_t max_block_n, int only_sat, uint64_t* n_ambiguous)
    {
        const ScenConst& sc = md.sc;
        const uint32_t ns = sc.n_sats, nw64 = md.nw64;
        constexpr int kThreads = kSynthThreads;
        const uint32_t TILE = (uint32_t)synth_tile_samples(K);
        const uint32_t tiles_per_block = (uint32_t)((max_block_n + TILE - 1) / TILE);
        const SynthK KK = make_synth_k(sc.delta46, sc.kmul, sc.cj, sc.dsum0, sc.spc, sc.lut_den, sc.ystride);
        std::vector<TileSat> tsat(std::max(1u, ns));
        std::vector<uint32_t> w32((size_t)std::max(1u, ns) * (nw64 + 1));
        std::vector<uint2> t64((size_t)std::max(1u, ns) * nw64);
        std::vector<float> yfix((size_t)std::max(1u, ns) * 8);
        const bool noise = !(sc.flags & R4WB_FLAG_NOISE_OFF);

        for (uint32_t tile = 0; tile < tb_count * tiles_per_block; ++tile) {
            const uint32_t tb = tb_begin + tile / tiles_per_block, chunk = tile % tiles_per_block;
            const BlockHdr hd = tb_hdr[tb];
            const uint32_t i_begin = chunk * TILE;
            if (i_begin >= hd.n) continue;
            const uint32_t i_end = std::min(hd.n, i_begin + TILE);
            if (hd.first + i_end <= out_first || hd.first + i_begin >= out_first + out_n) continue;
            const BlockSat* row = tb_tab + (size_t)tb * ns;
            // k_tile_params
            for (uint32_t s = 0; s < ns; ++s) {
                tsat[s] = tile_sat(row[s], tb_tab, i_begin, KK.d8);
                for (uint32_t i = 0; i < 8; ++i) {
                    float y = 0.0f;
                    if (chunk == 0 && (tsat[s].flags & 9u) == 9u && i < hd.n)
                        y = fir_block_start(row[s], tb_tab, md.perbits.data() + s * kPerWords, md.taps_f, md.etab_f, (int)i, KK, md.satcode[s]);
                    yfix[s * 8 + i] = y;
                }
            }
            for (uint32_t k = 0; k < ns * (nw64 + 1); ++k) {
                const uint32_t s = k / (nw64 + 1), w = k - s * (nw64 + 1);
                w32[k] = sign_word(md.perbits.data() + s * kPerWords, tsat[s].hb, w, md.satcode[s]);
            }
            for (uint32_t k = 0; k < ns * nw64; ++k) {
                const uint32_t s = k / nw64, w = k - s * nw64;
                t64[k] = make_uint2(w32[s * (nw64 + 1) + w], w32[s * (nw64 + 1) + w + 1]);
            }
            for (uint32_t tid = 0; tid < (uint32_t)kThreads; ++tid) {
                float2 ar[K], ai[K];
                for (int k = 0; k < K; ++k) ar[k] = ai[k] = make_float2(0.0f, 0.0f);
                for (uint32_t s = 0; s < ns; ++s) {
                    if (only_sat >= 0 && (int)s != only_sat) continue;
                    const TileSat ts = tsat[s];
                    if (!(ts.flags & 1u)) continue;
                    const SlowCtx slow{row + s, tb_tab, md.perbits.data() + s * kPerWords, md.taps_f, md.satcode.data() + s};
                    sat_accumulate<K>(ts, KK, t64.data() + s * nw64, md.ytab.data(), md.clslut.data(), yfix.data() + s * 8, slow, tid, i_begin, i_end, ar, ai,
                                      n_ambiguous);
                }
                for (int k = 0; k < K; ++k) {
                    const uint32_t ia = i_begin + 2 * tid + 2 * kThreads * k;
                    for (uint32_t j = 0; j < 2; ++j) {
                        const uint32_t i = ia + j;
                        if (i >= i_end) continue;
                        const uint64_t m = hd.first + i;
                        if (m < out_first || m >= out_first + out_n) continue;
                        float2 val = j == 0 ? make_float2(ar[k].x, ai[k].x) : make_float2(ar[k].y, ai[k].y);
                        if (noise && only_sat < 0) {
                            const float2 g = noise_of_sample(m, sc.seed);
                            val.x = fmaf(g.x, sc.noise_std, val.x);
                            val.y = fmaf(g.y, sc.noise_std, val.y);
                        }
                        out[2 * (m - out_first)] = val.x;
                        out[2 * (m - out_first) + 1] = val.y;
                    }
                }
            }
        }
    }

    // k_synth_lat (synth_lattice.cu), thread loops flattened: full canonical blocks only.  Returns the number of patched windows.
    template <int K>
    uint64_t render_lat_t(const BlockSat* tb_tab, const BlockHdr* tb_hdr, uint32_t tb_begin, uint32_t tb_count, uint64_t out_first,
                          uint64_t out_n, float* out, int only_sat)
    {
        const ScenConst& sc = md.sc;
        const LatConst& L = md.lat;
        const uint32_t ns = sc.n_sats, n_ent = lat_n_ent(L), n_w = lat_n_words(L);
        const SynthK KK = make_synth_k(sc.delta46, sc.kmul, sc.cj, sc.dsum0, sc.spc, sc.lut_den, sc.ystride);
        const uint64_t d8_46 = sc.delta46 * (uint64_t)kOversample;
        const bool noise = !(sc.flags & R4WB_FLAG_NOISE_OFF);
        std::vector<TileRec> rec(ns);
        std::vector<uint32_t> W((size_t)ns * n_w);
        std::vector<uint4> ent((size_t)ns * n_ent);
        uint64_t patched = 0;
        for (uint32_t tile = 0; tile < tb_count; ++tile) {
            const uint32_t tb = tb_begin + tile;
            const BlockHdr hd = tb_hdr[tb];
            if (hd.n != 2 * L.q) fail(R4WB_ERR_INVALID_SIZE, "lattice replay: partial block");
            if (hd.first + hd.n <= out_first || hd.first >= out_first + out_n) continue;
            const BlockSat* row = tb_tab + (size_t)tb * ns;
            for (uint32_t s = 0; s < ns; ++s) {                              // k_tile_params
                rec[s].ts = tile_sat(row[s], tb_tab, 0, KK.d8);
                for (uint32_t i = 0; i < 8; ++i)
                    rec[s].yfix[i] = (rec[s].ts.flags & 9u) == 9u ? fir_block_start(row[s], tb_tab, md.perbits.data() + s * kPerWords, md.taps_f, md.etab_f, (int)i, KK, md.satcode[s]) : 0.0f;
                rec[s].lat = tile_lat(row[s], L, md.perbits.data() + s * kPerWords, sc.spc, md.satcode[s]);
                if ((rec[s].ts.flags & 1u) && !lat_rotation_ok(rec[s].ts)) fail(R4WB_ERR_NOT_SUPPORTED, "lattice replay: carrier model out of range");
                for (int e = 0; e < 2; ++e) if ((rec[s].lat.patch >> (16 + 2 * e)) & 3u) ++patched;
                for (uint32_t w = 0; w < n_w; ++w) {
                    W[s * n_w + w] = sign_word_ep(md.perbits.data() + s * kPerWords, rec[s].lat.ep0 & 0xffu, rec[s].lat.ep0 >> 8, w, md.satcode[s]);
                    if (W[s * n_w + w] != sign_word(md.perbits.data() + s * kPerWords, rec[s].ts.hb, w, md.satcode[s])) fail(R4WB_ERR_CUDA, "sign_word_ep != sign_word");
                }
                for (uint32_t j = 0; j < n_ent; ++j) ent[s * n_ent + j] = lat_entry(W.data() + s * n_w, j, L.p);
            }
            for (uint32_t tid = 0; tid < (uint32_t)kSynthThreads; ++tid) {
                float2 arA[K], aiA[K], arB[K], aiB[K];
                for (int k = 0; k < K; ++k) arA[k] = aiA[k] = arB[k] = aiB[k] = make_float2(0.0f, 0.0f);
                for (uint32_t s = 0; s < ns; ++s) {
                    if (only_sat >= 0 && (int)s != only_sat) continue;
                    if (!(rec[s].ts.flags & 1u)) continue;
                    lat_sat_accumulate<K>(rec[s], L, d8_46, ent.data() + s * n_ent, md.ytab2.data(), md.clsn.data(), md.taps_f, tid, arA, aiA, arB, aiB);
                }
                for (int k = 0; k < K; ++k) {
                    const uint32_t ia = 2 * tid + 2 * kSynthThreads * k;
                    if (ia >= L.q) continue;
                    for (uint32_t v = 0; v < 4; ++v) {
                        const uint32_t i = ia + (v & 1u) + ((v & 2u) ? L.q : 0u);
                        const uint64_t m = hd.first + i;
                        if (m < out_first || m >= out_first + out_n) continue;
                        const float2& r2 = (v & 2u) ? arB[k] : arA[k];
                        const float2& i2 = (v & 2u) ? aiB[k] : aiA[k];
                        float2 val = (v & 1u) ? make_float2(r2.y, i2.y) : make_float2(r2.x, i2.x);
                        if (noise && only_sat < 0) {
                            const float2 g = noise_of_sample(m, sc.seed);
                            val.x = fmaf(g.x, sc.noise_std, val.x);
                            val.y = fmaf(g.y, sc.noise_std, val.y);
                        }
                        out[2 * (m - out_first)] = val.x;
                        out[2 * (m - out_first) + 1] = val.y;
                    }
                }
            }
        }
        return patched;
    }

    bool use_lattice = false;      // emu_scenario_set_lattice: replay k_synth_lat for full blocks instead of k_synth
    uint64_t n_patched = 0;

    void render(const BlockSat* tb_tab, const BlockHdr* tb_hdr, uint32_t tb_begin, uint32_t tb_count, uint64_t out_first,
                uint64_t out_n, float* out, uint64_t max_block_n, int only_sat, uint64_t* n_ambiguous)
    {
        if (use_lattice && md.lat.q != 0 && max_block_n == md.sc.B && !md.any_direct) {
            uint32_t n_full = tb_count;
            if (tb_hdr[tb_begin + tb_count - 1].n != md.sc.B) n_full -= 1;
            if (n_full) n_patched += md.lat.K == 5 ? render_lat_t<5>(tb_tab, tb_hdr, tb_begin, n_full, out_first, out_n, out, only_sat)
                                                   : render_lat_t<4>(tb_tab, tb_hdr, tb_begin, n_full, out_first, out_n, out, only_sat);
            if (n_full == tb_count) return;
            tb_begin += n_full; tb_count -= n_full;
        }
        if (md.tile_k == 5) render_t<5>(tb_tab, tb_hdr, tb_begin, tb_count, out_first, out_n, out, max_block_n, only_sat, n_ambiguous);
        else render_t<10>(tb_tab, tb_hdr, tb_begin, tb_count, out_first, out_n, out, max_block_n, only_sat, n_ambiguous);
        if (!md.any_direct) return;
        // k_synth_direct, thread loop flattened: GPS L5 / GLONASS satellites added to what the main pass wrote
        const uint32_t ns = md.sc.n_sats;
        for (uint32_t tb = tb_begin; tb < tb_begin + tb_count; ++tb) {
            const BlockHdr hd = tb_hdr[tb];
            const BlockSat* row = tb_tab + (size_t)tb * ns;
            for (uint32_t i = 0; i < hd.n; ++i) {
                const uint64_t m = hd.first + i;
                if (m < out_first || m >= out_first + out_n) continue;
                float re = 0.0f, im = 0.0f;
                for (uint32_t s = 0; s < ns; ++s) {
                    if (only_sat >= 0 && (int)s != only_sat) continue;
                    const DirectSat d = md.dsat[s];
                    if (!d.direct || !(row[s].flags & 1u)) continue;
                    const float2 v = direct_sample(row[s], tb_tab, md.dcodebits.data() + (size_t)s * kDirectWords, md.taps_f, i, d);
                    re += v.x; im += v.y;
                }
                out[2 * (m - out_first)] += re;
                out[2 * (m - out_first) + 1] += im;
            }
        }
    }
};

thread_local std::string g_err;

template <typename F>
int guard(F&& f)
{
    try { f(); return 0; }
    catch (const Failure& e) { g_err = e.what; return (int)e.code; }
    catch (const std::exception& e) { g_err = e.what(); return -1; }
}

}  // namespace

extern "C" {

const char* emu_last_error(void) { return g_err.c_str(); }

// The binade / integer-sum model of the reference's sequential f64 carrier phase (synth_math.cuh: block_phase_q,
// phase_after_block) against the literal per-sample accumulation, for a Doppler d0 + rate t + jerk t^2 over `blocks` blocks of
// 5000 samples at 5 MHz.  The binade is predicted from the approximate prefix exactly as k_phase_prefix / k_phase_q do.
// out[0] = blocks where the two differ (must be 0), out[1] = blocks walked sample by sample, out[2] = final phase (rad),
// out[3] = final phase - real-number sum (the reference's rounding drift).
void emu_phase_model_check(double d0, double rate, double jerk, uint64_t blocks, double* out)
{
    const double fs = 5e6;
    const uint32_t n = 5000;
    double ph_ref = 0.0, ph = 0.0, approx_sum = 0.0;
    uint64_t walked = 0, mism = 0;
    for (uint64_t b = 0; b < blocks; ++b) {
        const double t0 = (double)b * 1e-3, t1 = t0 + 1e-3;
        const double ds = d0 + rate * t0 + jerk * t0 * t0, de = d0 + rate * t1 + jerk * t1 * t1;
        ph_ref = phase_walk(ph_ref, ds, de, n, fs);
        PhaseQ r{};
        block_phase_approx(ds, de, n, fs, &r.approx, &r.span);
        if (approx_sum != 0.0) {
            r.k = ilogb(approx_sum);
            if (r.k >= kPhaseMinBinade) {
                bool tie;
                block_phase_q(ds, de, n, fs, r.k, &r.Q, &tie);
                r.ok = tie ? 0u : 1u;
                long long q2 = 0; bool tie2 = false;                              // the step-function form must agree wherever it answers
                if (block_phase_q_steps(ds, de, n, fs, 1.0 / fs, (int)(b & 1u), r.k, 1024u, &q2, &tie2) && (tie2 != tie || (!tie && q2 != r.Q))) ++mism;
            }
        }
        if (!phase_stays_in_binade(ph, r)) ++walked;
        ph = phase_after_block(ph, r, ds, de, n, fs);
        approx_sum += r.approx;
        if (ph != ph_ref) { ++mism; ph = ph_ref; }
    }
    out[0] = (double)mism; out[1] = (double)walked; out[2] = ph_ref; out[3] = ph_ref - approx_sum;
}

// block_phase_q_steps (the step-function form k_phase_q uses) against block_phase_q (every sample) on `cases` random blocks:
// Doppler +-6 kHz, change over the block up to +-span_hz (both signs, zero crossings included), binade k in [k_lo, k_hi], both
// division forms.  out[0] = cases where Q or the tie flag differ (must be 0), out[1] = cases the step form declined (more
// than max_steps levels), out[2] = largest number of levels seen, out[3] = cases with a tie.
void emu_phase_q_steps_check(uint64_t seed, uint32_t cases, double span_hz, int k_lo, int k_hi, uint32_t n, uint32_t max_steps, double* out)
{
    const double fs = 5e6, inv_fs = 1.0 / fs;
    uint64_t st = seed * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) * (1.0 / 9007199254740992.0); };
    uint64_t bad = 0, declined = 0, ties = 0;
    long long max_levels = 0;
    for (uint32_t c = 0; c < cases; ++c) {
        double ds = (rnd() * 2.0 - 1.0) * 6000.0;
        if (c % 17 == 0) ds = (rnd() * 2.0 - 1.0) * span_hz * 0.5;                 // zero crossings
        double de = ds + (rnd() * 2.0 - 1.0) * span_hz;
        if (c % 23 == 0) de = ds;
        if (c % 29 == 0) { ds = ldexp(floor(ldexp(ds, 20)), -20); de = ds; }       // few mantissa bits: plateaus, ties more likely
        const int k = k_lo + (int)(rnd() * (double)(k_hi - k_lo + 1));
        long long q_ref = 0, q = 0;
        bool tie_ref = false, tie = false;
        block_phase_q(ds, de, n, fs, k, &q_ref, &tie_ref);
        if (tie_ref) ++ties;
        const int div = (int)(c & 1u);
        if (!block_phase_q_steps(ds, de, n, fs, inv_fs, div, k, max_steps, &q, &tie)) { ++declined; continue; }
        if (tie != tie_ref || (!tie && q != q_ref)) ++bad;                         // with a tie the block is walked and Q unused (rint vs round-half-down)
        // levels, for the statistics
        const double x0 = scalbn(ref_phase_inc(ds, de, 0, (double)n, fs), 52 - k), x1 = scalbn(ref_phase_inc(ds, de, n - 1, (double)n, fs), 52 - k);
        max_levels = std::max(max_levels, (long long)fabs(x1 - x0));
    }
    out[0] = (double)bad; out[1] = (double)declined; out[2] = (double)max_levels; out[3] = (double)ties;
}

int emu_scenario_create(const r4wb_scenario_cfg* cfg, void** out)
{
    *out = nullptr;
    return guard([&] { *out = new Emu(*cfg); });
}
void emu_scenario_destroy(void* h) { delete static_cast<Emu*>(h); }
uint64_t emu_scenario_total_samples(void* h) { return static_cast<Emu*>(h)->md.sc.total; }
uint64_t emu_scenario_block_size(void* h) { return static_cast<Emu*>(h)->md.sc.B; }
uint32_t emu_scenario_segments(void* h) { return (uint32_t)static_cast<Emu*>(h)->md.segments.size(); }
// 1: replay k_synth_lat for full blocks.  Returns whether the scenario qualifies (lattice constants present).
int emu_scenario_set_lattice(void* h, int on) { Emu* e = static_cast<Emu*>(h); e->use_lattice = on != 0; return e->md.lat.q != 0 ? 1 : 0; }
uint64_t emu_scenario_patched(void* h) { return static_cast<Emu*>(h)->n_patched; }

// canonical-partition random access (mirror of Scenario::generate)
int emu_scenario_generate(void* h, uint64_t first, uint64_t n, float* out_cf32, int only_sat, uint64_t* n_ambiguous)
{
    Emu* e = static_cast<Emu*>(h);
    return guard([&] {
        const ScenConst& sc = e->md.sc;
        if (first > sc.total || n > sc.total - first) fail(R4WB_ERR_INVALID_SIZE, "range exceeds total_samples");
        if (n == 0) return;
        if (n_ambiguous) *n_ambiguous = 0;
        const uint64_t b0 = first / sc.B, b1 = (first + n - 1) / sc.B;
        e->build_table(e->md.table_begin(b0), b1 + 1);
        e->render(e->tab.data(), e->hdr.data(), (uint32_t)(b0 - e->tab_blk0), (uint32_t)(b1 - b0 + 1), first, n, out_cf32, sc.B,
                  only_sat, n_ambiguous);
    });
}

// sequential API (mirror of Scenario::generate_block)
int emu_scenario_generate_block(void* h, uint64_t n_req, float* out_cf32, uint64_t* written)
{
    Emu* e = static_cast<Emu*>(h);
    *written = 0;
    return guard([&] {
        const ScenConst& sc = e->md.sc;
        const uint64_t remaining = sc.total > e->current ? sc.total - e->current : 0;
        const uint64_t n = std::min(remaining, n_req);
        if (n == 0) return;
        std::vector<BlockSat> tab;
        BlockHdr hdr[2];
        e->seq.make_table(e->md, e->current, (uint32_t)n, tab, hdr);
        e->render(tab.data(), hdr, 1, 1, e->current, n, out_cf32, n, -1, nullptr);
        e->seq.advance(e->md, tab, (uint32_t)n);
        e->current += n;
        *written = n;
    });
}

int emu_block_params(void* h, uint64_t block, uint32_t sat, double* out12)
{
    Emu* e = static_cast<Emu*>(h);
    return guard([&] {
        if (block >= e->md.n_blocks() || sat >= e->md.sc.n_sats) fail(R4WB_ERR_INVALID_PARAMETER, "block/sat out of range");
        e->build_table(e->md.table_begin(block), block + 1);
        block_sat_debug(e->tab[(size_t)(block - e->tab_blk0) * e->md.sc.n_sats + sat], out12);
    });
}

}  // extern "C"
