// synth_lattice.cu — k_synth_lat: the synthesis kernel of every scenario whose blocks sit on the reference's sample lattice
// (all e1c_*.yaml: 5 MHz, 1 ms blocks of 2 q = 5000 samples, 1.023 MHz chipping).  Replaces, per sample, the same reference code
// as k_synth (synth_kernels.cu): satellite_emitter.rs:218-347, fir.rs:392-409 + scenario.rs:486-489 (decimation),
// scenario.rs:516-528 (Doppler rotation, amplitude, sum), scenario.rs:530-542 (noise), io/format.rs:197-222 (sink cast).
//
// One CTA renders one whole 1 ms block at a time (persistent grid, 2 CTAs per SM).  Thread t owns, in each of its K steps,
// the quad {m, m + 1, m + q, m + q + 1}, m = 2 t + 512 k: the two halves of a block share the boundary-age classes (one byte
// load each, consecutive threads read consecutive bytes) and one 16-byte window load yields all four sign patterns
// (synth_lattice.cuh).  The next block's sign tables are built while this block is rendered (double-buffered window
// tables and tile records): two barriers per block.  Stores are 16-byte pairs, a warp writes 512 contiguous bytes.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>

#include "synth_lattice.cuh"

namespace r4wb {

// store helpers of synth_kernels.cu (sink formats)
__device__ __forceinline__ int lat_i16(float x) { return __double2int_rz(fmin(fmax((double)x * 32767.0, -32768.0), 32767.0)); }
__device__ __forceinline__ int lat_i8(float x) { return __double2int_rz(fmin(fmax((double)x * 127.0, -128.0), 127.0)); }
__device__ __forceinline__ int lat_u8(float x) { return __double2int_rz(fmin(fmax(((double)x + 1.0) * 127.5, 0.0), 255.0)); }

template <int FMT>
__device__ __forceinline__ void lat_store_sample(void* out, uint64_t o, float re, float im)
{
    if (FMT == R4WB_FMT_CF64) reinterpret_cast<double2*>(out)[o] = make_double2((double)re, (double)im);
    else if (FMT == R4WB_FMT_CF32) reinterpret_cast<float2*>(out)[o] = make_float2(re, im);
    else if (FMT == R4WB_FMT_CI16) reinterpret_cast<uint32_t*>(out)[o] = ((uint32_t)lat_i16(re) & 0xffffu) | ((uint32_t)lat_i16(im) << 16);
    else if (FMT == R4WB_FMT_CI8) reinterpret_cast<uint16_t*>(out)[o] = (uint16_t)(((uint32_t)lat_i8(re) & 0xffu) | (((uint32_t)lat_i8(im) & 0xffu) << 8));
    else reinterpret_cast<uint16_t*>(out)[o] = (uint16_t)((uint32_t)lat_u8(re) | ((uint32_t)lat_u8(im) << 8));
}
template <int FMT>
__device__ __forceinline__ void lat_store_pair(void* out, uint64_t o, float4 v)     // o even, out aligned to the pair
{
    if (FMT == R4WB_FMT_CF32) reinterpret_cast<float4*>(out)[o >> 1] = v;
    else if (FMT == R4WB_FMT_CI16)
        reinterpret_cast<uint2*>(out)[o >> 1] = make_uint2(((uint32_t)lat_i16(v.x) & 0xffffu) | ((uint32_t)lat_i16(v.y) << 16),
                                                           ((uint32_t)lat_i16(v.z) & 0xffffu) | ((uint32_t)lat_i16(v.w) << 16));
    else if (FMT == R4WB_FMT_CI8)
        reinterpret_cast<uint32_t*>(out)[o >> 1] = ((uint32_t)lat_i8(v.x) & 0xffu) | (((uint32_t)lat_i8(v.y) & 0xffu) << 8) |
                                                   (((uint32_t)lat_i8(v.z) & 0xffu) << 16) | ((uint32_t)lat_i8(v.w) << 24);
    else if (FMT == R4WB_FMT_CU8)
        reinterpret_cast<uint32_t*>(out)[o >> 1] = (uint32_t)lat_u8(v.x) | ((uint32_t)lat_u8(v.y) << 8) | ((uint32_t)lat_u8(v.z) << 16) |
                                                   ((uint32_t)lat_u8(v.w) << 24);
    else { lat_store_sample<FMT>(out, o, v.x, v.y); lat_store_sample<FMT>(out, o + 1, v.z, v.w); }
}

constexpr int kLatThreads = kSynthThreads;
constexpr int kLatCtasPerSm = 2;      // 16 warps per SM; a third CTA fits (80 registers, 64 KB) but measured 5 % slower: the kernel is bound by shared-memory wavefronts and issue slots, not latency

struct LatSmem {
    float* ytab2; float* taps; uint8_t* clsn; uint32_t* W; uint4* ent; TileRec* trec;
};

__host__ __device__ inline size_t lat_smem_layout(uint32_t n_sats, uint32_t ystride, const LatConst& L, LatSmem* m, unsigned char* raw)
{
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = (off + bytes + 15) & ~(size_t)15; return o; };
    const size_t o_ytab = take((size_t)ystride * kLatYStride * 4);
    const size_t o_taps = take(64 * 4);
    const size_t o_cls = take((size_t)8 * L.cls_len);
    const size_t o_w = take((size_t)n_sats * lat_n_words(L) * 4);
    const size_t o_ent = take((size_t)2 * n_sats * lat_n_ent(L) * 16);
    const size_t o_rec = take((size_t)2 * n_sats * sizeof(TileRec));
    if (m) {
        m->ytab2 = reinterpret_cast<float*>(raw + o_ytab); m->taps = reinterpret_cast<float*>(raw + o_taps);
        m->clsn = raw + o_cls; m->W = reinterpret_cast<uint32_t*>(raw + o_w);
        m->ent = reinterpret_cast<uint4*>(raw + o_ent); m->trec = reinterpret_cast<TileRec*>(raw + o_rec);
    }
    return off;
}

constexpr size_t kLatMaxSmem = 110 * 1024;     // beyond ~74 KB fewer than kLatCtasPerSm CTAs fit an SM (more than 8 satellites)

size_t lat_smem_bytes(uint32_t n_sats, uint32_t ystride, const LatConst& L) { return lat_smem_layout(n_sats, ystride, L, nullptr, nullptr); }

template <int K, int FMT>
__global__ void __launch_bounds__(kLatThreads, kLatCtasPerSm) k_synth_lat(SynthArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LatSmem sm;
    lat_smem_layout(a.n_sats, a.ystride, a.lat, &sm, smem_raw);
    __shared__ float s_pow[kLatThreads / 32];

    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const LatConst L = a.lat;
    const uint32_t n_ent = lat_n_ent(L), n_w = lat_n_words(L);
    const uint64_t d8_46 = a.delta46 * (uint64_t)kOversample;
    const PhiloxKeys PK = philox_keys(a.seed);
    const uint32_t q = L.q;

    // kernel-lifetime tables
    for (uint32_t k = tid; k < (a.ystride * (uint32_t)kLatYStride + 3u) / 4u; k += kLatThreads)
        reinterpret_cast<float4*>(sm.ytab2)[k] = reinterpret_cast<const float4*>(a.ytab2)[k];
    for (uint32_t k = tid; k < 64; k += kLatThreads) sm.taps[k] = a.taps[k];
    for (uint32_t k = tid; k < (8u * L.cls_len + 15u) / 16u; k += kLatThreads)
        reinterpret_cast<uint4*>(sm.clsn)[k] = reinterpret_cast<const uint4*>(a.clsn)[k];

    float pow_acc = 0.0f;
    const uint32_t n_tiles = a.tb_count;                                       // one tile = one block
    const uint32_t rec_f4 = a.n_sats * (uint32_t)(sizeof(TileRec) / 16);      // float4s of one block's records
    const float4* recs = reinterpret_cast<const float4*>(a.tiles + (size_t)a.tb_begin * a.tiles_per_block * a.n_sats);
    const bool noise_on = !(a.flags & R4WB_FLAG_NOISE_OFF);

    // records of tile `t` -> registers (rec_f4 <= 16 * 8 = 128 <= kLatThreads)
    auto load_rec = [&](uint32_t tile) {
        return (tile < n_tiles && tid < rec_f4) ? __ldg(recs + (size_t)tile * a.tiles_per_block * rec_f4 + tid) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    // Sign words of a block: a warp takes a satellite (two when there are more than eight), its lanes the words — no index
    // division.  The table origin `ep0` of the warp's satellites is fetched from the records in global memory one block ahead
    // (registers), the satellites' code structures once per kernel, so the prologue of a block waits for no global load.
    const TileRec* g_recs = a.tiles + (size_t)a.tb_begin * a.tiles_per_block * a.n_sats;
    constexpr uint32_t kWarps = kLatThreads / 32, kSatsPerWarp = (kLatMaxSats + kWarps - 1) / kWarps;
    SatCode my_cd[kSatsPerWarp];
#pragma unroll
    for (uint32_t u = 0; u < kSatsPerWarp; ++u) {
        const uint32_t s = (tid >> 5) + u * kWarps;
        my_cd[u] = a.satcode[s < a.n_sats ? s : 0];
    }
    auto load_ep0 = [&](uint32_t t, uint32_t (&ep)[kSatsPerWarp]) {
#pragma unroll
        for (uint32_t u = 0; u < kSatsPerWarp; ++u) {
            const uint32_t s = (tid >> 5) + u * kWarps;
            ep[u] = (t < n_tiles && s < a.n_sats) ? __ldg(&g_recs[(size_t)t * a.n_sats + s].lat.ep0) : 0u;
        }
    };
    auto build_words = [&](const uint32_t (&ep)[kSatsPerWarp]) {
#pragma unroll
        for (uint32_t u = 0; u < kSatsPerWarp; ++u) {
            const uint32_t s = (tid >> 5) + u * kWarps;
            if (s >= a.n_sats) break;
            for (uint32_t w = lane; w < n_w; w += 32)
                sm.W[s * n_w + w] = sign_word_ep(a.perbits + s * kPerWords, ep[u] & 0xffu, ep[u] >> 8, w, my_cd[u]);
        }
    };
    auto build_entries = [&](uint32_t buf) {
        for (uint32_t s = tid >> 5; s < a.n_sats; s += kLatThreads / 32)
            for (uint32_t j = lane; j < n_ent; j += 32) sm.ent[(buf * a.n_sats + s) * n_ent + j] = lat_entry(sm.W + s * n_w, j, L.p);
    };

    // pipeline prologue: first tile's tables
    uint32_t tile = blockIdx.x;
    {
        const float4 r0 = load_rec(tile);
        uint32_t ep_first[kSatsPerWarp];
        load_ep0(tile, ep_first);
        if (tid < rec_f4) reinterpret_cast<float4*>(sm.trec)[tid] = r0;
        __syncthreads();                                                       // kernel-lifetime tables in place
        if (tile < n_tiles) build_words(ep_first);
        __syncthreads();
        if (tile < n_tiles) build_entries(0);
    }
    float4 pre = load_rec(tile + gridDim.x);
    uint32_t ep_next[kSatsPerWarp];
    load_ep0(tile + gridDim.x, ep_next);

    for (uint32_t it = 0; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t cur = it & 1u, nxt = cur ^ 1u;
        // the next tile's records, sign words and window entries go into buffer nxt while this tile is rendered from buffer
        // cur: the barrier ends the previous iteration's reads of nxt (and of W), the second one publishes W
        const bool has_next = tile + gridDim.x < n_tiles;
        __syncthreads();
        if (tid < rec_f4) reinterpret_cast<float4*>(sm.trec + nxt * a.n_sats)[tid] = pre;
        pre = load_rec(tile + 2u * gridDim.x);
        if (has_next) build_words(ep_next);
        load_ep0(tile + 2u * gridDim.x, ep_next);
        __syncthreads();
        if (has_next) build_entries(nxt);

        const uint32_t tb = a.tb_begin + tile;
        const BlockHdr hd = a.hdr[tb];
        // blocks entirely outside the requested output range are skipped (uniform per CTA)
        if (hd.first + hd.n <= a.out_first || hd.first >= a.out_first + a.out_n) continue;

        float2 arA[K], aiA[K], arB[K], aiB[K];
#pragma unroll
        for (int k = 0; k < K; ++k) arA[k] = aiA[k] = arB[k] = aiB[k] = make_float2(0.0f, 0.0f);

        const TileRec* trec = sm.trec + cur * a.n_sats;
        for (uint32_t s = 0; s < a.n_sats; ++s) {
            if (!(trec[s].ts.flags & 1u)) continue;
            lat_sat_accumulate<K>(trec[s], L, d8_46, sm.ent + (cur * a.n_sats + s) * n_ent, sm.ytab2, sm.clsn, sm.taps, tid, arA, aiA, arB, aiB);
        }

        // noise, power, store.  The usual case (block inside the requested range, output aligned to sample pairs) is one straight
        // run of 16-byte stores; the masked variant for ragged range ends lives in its own loop so that it stays out of the hot
        // instruction stream (the per-block code path is ~40 KB, close to what the instruction cache holds).
        const uint64_t m0 = hd.first;                                         // even (blocks of 2 q samples, q even)
        const bool plain = a.out_aligned16 && m0 >= a.out_first && m0 + hd.n <= a.out_first + a.out_n && ((m0 - a.out_first) & 1ull) == 0;
        if (plain) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const uint32_t ia = 2u * tid + (uint32_t)(2 * kLatThreads * k);
                if (ia >= q) continue;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint64_t m = m0 + ia + (h ? q : 0u);
                    float4 v = h ? make_float4(arB[k].x, aiB[k].x, arB[k].y, aiB[k].y) : make_float4(arA[k].x, aiA[k].x, arA[k].y, aiA[k].y);
                    if (noise_on) {
                        float2 ga, gb2;
                        noise_of_counter(m >> 1, PK, ga, gb2);
                        v.x = fmaf(ga.x, a.noise_std, v.x); v.y = fmaf(ga.y, a.noise_std, v.y);
                        v.z = fmaf(gb2.x, a.noise_std, v.z); v.w = fmaf(gb2.y, a.noise_std, v.w);
                    }
                    pow_acc += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
                    lat_store_pair<FMT>(a.out, m - a.out_first, v);
                }
            }
        } else {
#pragma unroll 1
            for (int kk = 0; kk < 2 * K; ++kk) {
                const int k = kk >> 1, h = kk & 1;
                const uint32_t ia = 2u * tid + (uint32_t)(2 * kLatThreads * k);
                if (ia >= q) continue;
                const uint64_t m = m0 + ia + (h ? q : 0u);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int c = 0; c < K; ++c)                                   // static register indices: select, do not index
                    if (c == k) v = h ? make_float4(arB[c].x, aiB[c].x, arB[c].y, aiB[c].y) : make_float4(arA[c].x, aiA[c].x, arA[c].y, aiA[c].y);
                if (noise_on) {
                    float2 ga, gb2;
                    noise_of_counter(m >> 1, PK, ga, gb2);
                    v.x = fmaf(ga.x, a.noise_std, v.x); v.y = fmaf(ga.y, a.noise_std, v.y);
                    v.z = fmaf(gb2.x, a.noise_std, v.z); v.w = fmaf(gb2.y, a.noise_std, v.w);
                }
                const bool wa = m >= a.out_first && m < a.out_first + a.out_n;
                const bool wb = (m + 1) >= a.out_first && (m + 1) < a.out_first + a.out_n;
                if (wa) pow_acc += v.x * v.x + v.y * v.y;
                if (wb) pow_acc += v.z * v.z + v.w * v.w;
                const uint64_t o = m - a.out_first;                           // only meaningful when wa (wraps otherwise)
                if (wa && wb && ((o & 1ull) == 0) && a.out_aligned16) lat_store_pair<FMT>(a.out, o, v);
                else {
                    if (wa) lat_store_sample<FMT>(a.out, o, v.x, v.y);
                    if (wb) lat_store_sample<FMT>(a.out, o + 1, v.z, v.w);
                }
            }
        }
    }

    if (a.power_sum) {
        for (int off = 16; off > 0; off >>= 1) pow_acc += __shfl_xor_sync(0xffffffffu, pow_acc, off);
        if (lane == 0) s_pow[tid >> 5] = pow_acc;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < kLatThreads / 32; ++w) t += (double)s_pow[w];
            atomicAdd(a.power_sum, t);
        }
    }
}

// ----------------------------------------------------------------------------------------------
template <int K, int FMT>
static void launch_lat_t(const SynthArgs& a, int grid, size_t smem, cudaStream_t st)
{
    static PerDeviceOnce attr_done;
    if (attr_done.first()) {
        R4WB_CUDA(cudaFuncSetAttribute(k_synth_lat<K, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLatMaxSmem));
    }
    k_synth_lat<K, FMT><<<grid, kLatThreads, smem, st>>>(a);
    R4WB_LAUNCH_CHECK();
}

template <int K>
static void launch_lat_k(const SynthArgs& a, r4wb_fmt fmt, int grid, size_t smem, cudaStream_t st)
{
    switch (fmt) {
    case R4WB_FMT_CF32: return launch_lat_t<K, R4WB_FMT_CF32>(a, grid, smem, st);
    case R4WB_FMT_CF64: return launch_lat_t<K, R4WB_FMT_CF64>(a, grid, smem, st);
    case R4WB_FMT_CI16: return launch_lat_t<K, R4WB_FMT_CI16>(a, grid, smem, st);
    case R4WB_FMT_CI8: return launch_lat_t<K, R4WB_FMT_CI8>(a, grid, smem, st);
    case R4WB_FMT_CU8: return launch_lat_t<K, R4WB_FMT_CU8>(a, grid, smem, st);
    }
    fail(R4WB_ERR_INVALID_PARAMETER, "unknown sample format %d", (int)fmt);
}

// true when the lattice kernel can run this scenario's full blocks
bool lat_supported(const SynthArgs& a)
{
    if (a.lat.q == 0 || a.n_sats == 0 || a.n_sats > (uint32_t)kLatMaxSats || a.tiles_per_block != 1) return false;
    if (a.lat.K != 4 && a.lat.K != 5) return false;
    return lat_smem_bytes(a.n_sats, a.ystride, a.lat) <= kLatMaxSmem;
}

void launch_synth_lat(const SynthArgs& a, r4wb_fmt fmt, int sm_count, cudaStream_t st)
{
    const size_t smem = lat_smem_bytes(a.n_sats, a.ystride, a.lat);
    static std::atomic<int> per_sm[kLatMaxSats + 1];
    int ps = per_sm[a.n_sats].load();
    if (!ps) {
        int nb = 0;
        R4WB_CUDA(cudaFuncSetAttribute(k_synth_lat<5, R4WB_FMT_CF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLatMaxSmem));
        R4WB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_synth_lat<5, R4WB_FMT_CF32>, kLatThreads, smem));
        ps = std::max(1, nb);
        per_sm[a.n_sats].store(ps);
    }
    const int grid = (int)std::max<uint32_t>(1u, std::min<uint32_t>(a.tb_count, (uint32_t)(sm_count * ps)));
    if (a.lat.K == 5) launch_lat_k<5>(a, fmt, grid, smem, st);
    else launch_lat_k<4>(a, fmt, grid, smem, st);
}

}  // namespace r4wb
