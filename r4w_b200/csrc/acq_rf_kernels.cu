// acq_rf_kernels.cu — the f32 fast path of PCPS acquisition for fft_size 32768 (Galileo E1C at 5 MHz: code_length
// 20 000, gnss/acquisition.rs:63-66), built on the register-resident 16384-point transform of rfft.cuh.
//
//   k_rf_fwd        one CTA per (snapshot, Doppler) row [or per local code]: carrier wipe-off as a phasor recurrence
//                   re-anchored in f64 every 8 points (acquisition.rs:133-140), zero padding, the top radix-2
//                   decimation-in-frequency fold, two 16384-point forward transforms (even / odd bins), spectrum
//                   stored in slot order with 16-byte coalesced stores            (acquisition.rs:109-117, 141-143)
//   k_rf_inv_peak   one CTA per (row, code): X * conj(C)/N with 16-byte coalesced loads (each operand read once),
//                   inverse transform of the even bins -> E (parked in shared memory, thread-private), inverse of
//                   the odd bins -> O, corr[n] = E + w^-n O and corr[n + 16384] = E - w^-n O, |.|^2, running sum,
//                   first-max / second-max for lags < code_length, warp-shuffle + CTA reduction -> one RowPeak
//                   (acquisition.rs:146-164).  The correlation surface never leaves the SM.
//
// Not tensor-core work (no dense contraction); the bound is FP32 issue with the spectra L2-resident.
#include <cuda_runtime.h>

#include <cstdlib>

#include "acq.cuh"
#include "rfft.cuh"

namespace r4wb {

using rf::cf;

// exp(+2 pi i J / 64), J = 0..31, as compile-time constants
__host__ __device__ constexpr float w64_cos(int j)
{
    constexpr float c[17] = {1.00000000000000000000f, 0.99518472667219692873f, 0.98078528040323043058f, 0.95694033573220882438f,
                             0.92387953251128673848f, 0.88192126434835504956f, 0.83146961230254523567f, 0.77301045336273699338f,
                             0.70710678118654757274f, 0.63439328416364548779f, 0.55557023301960228867f, 0.47139673682599780857f,
                             0.38268343236508983729f, 0.29028467725446233105f, 0.19509032201612833135f, 0.09801714032956077016f, 0.0f};
    return j <= 16 ? c[j] : -c[32 - j];
}
__host__ __device__ constexpr float w64_sin(int j) { return j <= 16 ? w64_cos(16 - j) : w64_cos(j - 16); }
template <int J> struct W64 { static constexpr float c = w64_cos(J), s = w64_sin(J); };

// exp(-2 pi i cyc), cyc reduced to [-1/2, 1/2] in f64 first
__device__ __forceinline__ cf phasor_cycles(double cyc)
{
    cyc -= rint(cyc);
    float s, c;
    sincospif(-2.0f * (float)cyc, &s, &c);
    return cf{c, s};
}

// MODE 0: rows are (snapshot, Doppler) pairs of the wiped-off input; MODE 1: rows are local replicas (+-1 int8), output
// conjugated and scaled by 1/N (the ifft normalisation of core/fft_utils.rs:104-107 folded into the code spectrum)
template <int MODE>
__global__ void __launch_bounds__(rf::kNT, 1)
k_rf_fwd(AcqGeom g, const void* __restrict__ input, uint32_t in64, uint64_t stride, uint32_t take, const int8_t* __restrict__ codes,
         uint64_t code_len, const cf* __restrict__ W, cf* __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* xb = reinterpret_cast<float*>(smem_raw);
    const uint32_t t = threadIdx.x, row = blockIdx.x;
    constexpr uint32_t M = rf::kM, N = 2u * rf::kM;
    const rf::Consts K = rf::load_consts(W, t);

    double dop_cyc = 0.0;                 // Doppler in cycles per sample
    const unsigned char* in = nullptr;
    const int8_t* code = nullptr;
    if (MODE == 0) {
        const uint32_t snap = row / g.D, d = row - snap * g.D;
        const double doppler = -g.dmax + (double)d * g.dstep;       // acquisition.rs:127-130
        dop_cyc = doppler / g.fs;
        in = static_cast<const unsigned char*>(input) + (size_t)snap * stride * (in64 ? 16 : 8);
    } else {
        code = codes + (size_t)row * code_len;
    }
    // x[k + M] enters with exp(-j phi M) relative to x[k]
    const cf qM = MODE == 0 ? phasor_cycles(dop_cyc * (double)M) : cf{1.0f, 0.0f};

    float4* o4 = reinterpret_cast<float4*>(out + (size_t)row * N);
#pragma unroll 1
    for (uint32_t r = 0; r < 2; ++r) {
        // z_r[k] = exp(-2 pi i (dop_cyc + r/N) k) * (x[k] + (-1)^r exp(-j phi M) x[k + M]),  k = t + 512 j
        const double nu = dop_cyc + (double)r / (double)N;
        const cf Q = phasor_cycles(nu * 512.0);
        const cf qs = r ? cf{-qM.re, -qM.im} : qM;
        cf a[32];
        // pass 1: every load of the residue is issued before the first use (x[k] lands in a[j], x[k + M] in hi[j])
        cf hi[8];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const uint32_t k = t + 512u * (uint32_t)j;
            cf x0 = cf{0.0f, 0.0f}, x1 = cf{0.0f, 0.0f};
            if (MODE == 0) {
                if (k < take) {
                    if (in64) { const cx<double> v = reinterpret_cast<const cx<double>*>(in)[k]; x0 = cf{(float)v.re, (float)v.im}; }
                    else x0 = ld_cx(reinterpret_cast<const cf*>(in) + k);
                }
                if (j < 8 && k + M < take) {                        // only reachable while code_length > 16384
                    if (in64) { const cx<double> v = reinterpret_cast<const cx<double>*>(in)[k + M]; x1 = cf{(float)v.re, (float)v.im}; }
                    else x1 = ld_cx(reinterpret_cast<const cf*>(in) + k + M);
                }
            } else {
                if (k < take) x0.re = (float)code[k];
                if (j < 8 && k + M < take) x1.re = (float)code[k + M];
            }
            a[j] = x0;
            if (j < 8) hi[j] = x1;
        }
        // pass 2: wipe-off / fold phasor, re-anchored in f64 every 8 points
        cf P = cf{1.0f, 0.0f};
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const uint32_t k = t + 512u * (uint32_t)j;
            if ((j & 7) == 0) P = phasor_cycles(nu * (double)k);    // f64 cycle reduction as the reference's f64 phase
            else P = rf::cmul(P, Q);
            a[j] = rf::cmul(P, j < 8 ? rf::cadd(a[j], rf::cmul(qs, hi[j])) : a[j]);
        }
        rf::forward<true>(a, xb, K, t);
        float4* o = o4 + (size_t)r * (M / 2);
        const float inv_n = 1.0f / (float)N;
#pragma unroll
        for (int s = 0; s < 16; ++s) {
            cf v0 = a[2 * s], v1 = a[2 * s + 1];
            if (MODE == 1) { v0 = cf{v0.re * inv_n, -v0.im * inv_n}; v1 = cf{v1.re * inv_n, -v1.im * inv_n}; }
            o[(uint32_t)s * 512u + t] = make_float4(v0.re, v0.im, v1.re, v1.im);
        }
    }
}

__device__ __forceinline__ PeakAcc<float> rf_peak_shfl_xor(const PeakAcc<float>& a, int off)
{
    PeakAcc<float> o;
    o.best = __shfl_xor_sync(0xffffffffu, a.best, off);
    o.second = __shfl_xor_sync(0xffffffffu, a.second, off);
    o.sum = __shfl_xor_sync(0xffffffffu, a.sum, off);
    o.idx = __shfl_xor_sync(0xffffffffu, a.idx, off);
    return o;
}

// a cell visited in ascending lag order inside a thread: strict `>` keeps the lowest index on ties
__device__ __forceinline__ void rf_peak_push(PeakAcc<float>& a, float mag, uint32_t idx)
{
    a.sum += mag;
    if (mag > a.best) { a.second = a.best; a.best = mag; a.idx = idx; }
    else a.second = fmaxf(a.second, mag);
}

// lags n = t + 512 J below M: corr[n] = E[n] + w_N^{-n} O[n]; w_N^{-t} is already inside O, exp(2 pi i J / 64) is applied here
template <int J>
struct RfEpilogue {
    static __device__ __forceinline__ void run(const cf* a, cf* hi, const float* e_re, const float* e_im, uint32_t t, uint32_t L,
                                               PeakAcc<float>& acc)
    {
        const uint32_t n = t + 512u * (uint32_t)J;
        const cf e = cf{e_re[(uint32_t)J * 512u + t], e_im[(uint32_t)J * 512u + t]};
        const cf o = J == 0 ? a[0] : rf::cmul(a[rf::bitrev5(J)], cf{W64<J>::c, W64<J>::s});
        const cf c0 = rf::cadd(e, o);
        if (J < 8) hi[J < 8 ? J : 0] = rf::csub(e, o);
        if (n < L) rf_peak_push(acc, c0.re * c0.re + c0.im * c0.im, n);
        RfEpilogue<J + 1>::run(a, hi, e_re, e_im, t, L, acc);
    }
};
template <>
struct RfEpilogue<32> {
    static __device__ __forceinline__ void run(const cf*, cf*, const float*, const float*, uint32_t, uint32_t, PeakAcc<float>&) {}
};

// blockIdx.x = row * P + code, row = snapshot * D + d
__global__ void __launch_bounds__(rf::kNT, 1)
k_rf_inv_peak(AcqGeom g, const cf* __restrict__ X, const cf* __restrict__ C, const cf* __restrict__ W, RowPeak* __restrict__ peaks)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* xb = reinterpret_cast<float*>(smem_raw);                 // exchange buffer
    float* e_re = xb + rf::kXbufFloats;                             // E parked per thread: [j][t]
    float* e_im = e_re + rf::kM;
    __shared__ PeakAcc<float> s_red[rf::kNT / 32];
    constexpr uint32_t M = rf::kM, N = 2u * rf::kM;
    const uint32_t t = threadIdx.x;
    const uint32_t row = blockIdx.x / g.P, code = blockIdx.x - row * g.P;
    const rf::Consts K = rf::load_consts(W, t);
    const cf wNt = cconj(W[t]);                                     // exp(+2 pi i t / N)

    const float4* X4 = reinterpret_cast<const float4*>(X + (size_t)row * N);
    const float4* C4 = reinterpret_cast<const float4*>(C + (size_t)code * N);

    cf a[32];
    // even bins: E[n] = sum_m Y[2m] w_M^{-nm}
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        const float4 x = __ldg(X4 + (uint32_t)s * 512u + t), c = __ldg(C4 + (uint32_t)s * 512u + t);
        a[2 * s] = rf::cmul(cf{x.x, x.y}, cf{c.x, c.y});
        a[2 * s + 1] = rf::cmul(cf{x.z, x.w}, cf{c.z, c.w});
    }
    rf::inverse<false, true>(a, xb, K, t, cf{1.0f, 0.0f});   // complex exchange: the E park is still free
    __syncthreads();                                                 // the complex exchange reached into the park area
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        e_re[(uint32_t)j * 512u + t] = a[rf::bitrev5(j)].re;
        e_im[(uint32_t)j * 512u + t] = a[rf::bitrev5(j)].im;
    }
    // odd bins: O[n] = sum_m Y[2m+1] w_M^{-nm}; w_N^{-n} O[n] with n = t + 512 j: w_N^{-t} goes into the last twiddle
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        const float4 x = __ldg(X4 + (M / 2) + (uint32_t)s * 512u + t), c = __ldg(C4 + (M / 2) + (uint32_t)s * 512u + t);
        a[2 * s] = rf::cmul(cf{x.x, x.y}, cf{c.x, c.y});
        a[2 * s + 1] = rf::cmul(cf{x.z, x.w}, cf{c.z, c.w});
    }
    rf::inverse<true, false>(a, xb, K, t, wNt);

    PeakAcc<float> acc;
    peak_init(acc);
    cf hi[8];                                                        // corr[n + M] for j < 8, pushed after all lags below M
    RfEpilogue<0>::run(a, hi, e_re, e_im, t, g.L, acc);
    if (g.L > M) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {                                // lags M + n < L <= M + 4096
            const uint32_t n = M + t + 512u * (uint32_t)j;
            if (n < g.L) rf_peak_push(acc, hi[j].re * hi[j].re + hi[j].im * hi[j].im, n);
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) peak_merge(acc, rf_peak_shfl_xor(acc, off));
    if ((t & 31u) == 0) s_red[t >> 5] = acc;
    __syncthreads();
    if (t < 32) {
        PeakAcc<float> b;
        peak_init(b);
        if (t < rf::kNT / 32) b = s_red[t];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) peak_merge(b, rf_peak_shfl_xor(b, off));
        if (t == 0) {
            RowPeak o;
            o.best = (double)b.best; o.second = (double)b.second; o.sum = (double)b.sum; o.lag = b.idx; o.pad = 0;
            peaks[blockIdx.x] = o;
        }
    }
}

// ---- Tensor Memory as a per-thread park (sm_100a).  TMEM is 128 lanes x 512 columns x 32 bit per SM; a warp reaches
// the 32 lanes of its quarter (warp % 4) and tcgen05.ld/st.32x32b hands every thread its own lane, so 512 threads get 128
// private 32-bit words each (columns [(warp / 4) * 128, +128)): exactly one forward-spectrum row in slot order.
__device__ __forceinline__ void tmem_alloc_512(uint32_t* smem_dst)   // one full warp
{
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(dst) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_512(uint32_t base)      // one full warp
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
                   "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
                   "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// One CTA per (snapshot, Doppler) row, all P codes: the row's forward spectrum is read from L2 once and parked in Tensor
// Memory (all 256 KB), so every further code costs one code-spectrum read (256 KB) instead of two operand reads.
// peaks[row * P + code] as k_rf_inv_peak.
__global__ void __launch_bounds__(rf::kNT, 1)
k_rf_inv_peak_tm(AcqGeom g, const cf* __restrict__ X, const cf* __restrict__ C, const cf* __restrict__ W, RowPeak* __restrict__ peaks)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* xb = reinterpret_cast<float*>(smem_raw);
    float* e_re = xb + rf::kXbufFloats;
    float* e_im = e_re + rf::kM;
    __shared__ PeakAcc<float> s_red[rf::kNT / 32];
    __shared__ uint32_t s_tmem;
    constexpr uint32_t M = rf::kM, N = 2u * rf::kM;
    const uint32_t t = threadIdx.x, warp = t >> 5, row = blockIdx.x;
    const rf::Consts K = rf::load_consts(W, t);
    const cf wNt = cconj(W[t]);

    if (warp == 0) tmem_alloc_512(&s_tmem);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = s_tmem + (((warp & 3u) * 32u) << 16) + (warp >> 2) * 128u;

    // park the row: half h, register pair s -> columns h * 64 + 4 s .. + 3
    {
        const float4* X4 = reinterpret_cast<const float4*>(X + (size_t)row * N);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                float v[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 x = __ldg(X4 + (uint32_t)h * (M / 2) + (uint32_t)(4 * c4 + q) * 512u + t);
                    v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
                }
                tmem_st16(tbase + (uint32_t)(h * 64 + c4 * 16), v);
            }
        }
        tmem_wait_st();
    }

#pragma unroll 1
    for (uint32_t code = 0; code < g.P; ++code) {
        const float4* C4 = reinterpret_cast<const float4*>(C + (size_t)code * N);
        cf a[32];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
            float v[16];
            tmem_ld16(tbase + (uint32_t)(c4 * 16), v);
            float4 c[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) c[q] = __ldg(C4 + (uint32_t)(4 * c4 + q) * 512u + t);
            tmem_wait_ld();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                a[8 * c4 + 2 * q] = rf::cmul(cf{v[4 * q], v[4 * q + 1]}, cf{c[q].x, c[q].y});
                a[8 * c4 + 2 * q + 1] = rf::cmul(cf{v[4 * q + 2], v[4 * q + 3]}, cf{c[q].z, c[q].w});
            }
        }
        rf::inverse<false, true>(a, xb, K, t, cf{1.0f, 0.0f});
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            e_re[(uint32_t)j * 512u + t] = a[rf::bitrev5(j)].re;
            e_im[(uint32_t)j * 512u + t] = a[rf::bitrev5(j)].im;
        }
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
            float v[16];
            tmem_ld16(tbase + (uint32_t)(64 + c4 * 16), v);
            float4 c[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) c[q] = __ldg(C4 + (M / 2) + (uint32_t)(4 * c4 + q) * 512u + t);
            tmem_wait_ld();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                a[8 * c4 + 2 * q] = rf::cmul(cf{v[4 * q], v[4 * q + 1]}, cf{c[q].x, c[q].y});
                a[8 * c4 + 2 * q + 1] = rf::cmul(cf{v[4 * q + 2], v[4 * q + 3]}, cf{c[q].z, c[q].w});
            }
        }
        rf::inverse<true, false>(a, xb, K, t, wNt);

        PeakAcc<float> acc;
        peak_init(acc);
        cf hi[8];
        RfEpilogue<0>::run(a, hi, e_re, e_im, t, g.L, acc);
        if (g.L > M) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t n = M + t + 512u * (uint32_t)j;
                if (n < g.L) rf_peak_push(acc, hi[j].re * hi[j].re + hi[j].im * hi[j].im, n);
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) peak_merge(acc, rf_peak_shfl_xor(acc, off));
        if ((t & 31u) == 0) s_red[t >> 5] = acc;
        __syncthreads();
        if (t < 32) {
            PeakAcc<float> b;
            peak_init(b);
            if (t < rf::kNT / 32) b = s_red[t];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) peak_merge(b, rf_peak_shfl_xor(b, off));
            if (t == 0) {
                RowPeak o;
                o.best = (double)b.best; o.second = (double)b.second; o.sum = (double)b.sum; o.lag = b.idx; o.pad = 0;
                peaks[(size_t)row * g.P + code] = o;
            }
        }
        // s_red is rewritten only after the next code's transforms, which synchronise the CTA many times
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc_512(s_tmem);
}

// ---------------------------------------------------------------------------------------------- launchers
bool rf_supported(const AcqGeom& g) { return g.logN == rf::kLogM + 1 && g.L <= (uint32_t)rf::kM + 4096u; }

static size_t rf_fwd_smem() { return (size_t)rf::kXbufFloats * 8; }   // complex exchange buffer
static size_t rf_inv_smem() { return (size_t)(rf::kXbufFloats + 2 * rf::kM) * 4; }

void launch_rf_fwd_input(const AcqGeom& g, uint32_t rows, const void* input, uint32_t in64, uint64_t stride, uint32_t take, const cf* W,
                         cf* out, cudaStream_t st)
{
    static PerDeviceOnce attr;
    if (attr.first()) {
        R4WB_CUDA(cudaFuncSetAttribute(k_rf_fwd<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rf_fwd_smem()));
    }
    if (rows == 0) return;
    k_rf_fwd<0><<<rows, rf::kNT, rf_fwd_smem(), st>>>(g, input, in64, stride, take, nullptr, 0, W, out);
    R4WB_LAUNCH_CHECK();
}

void launch_rf_fwd_codes(const AcqGeom& g, uint32_t n_codes, const int8_t* codes, uint64_t code_len, uint32_t take, const cf* W, cf* out,
                         cudaStream_t st)
{
    static PerDeviceOnce attr;
    if (attr.first()) {
        R4WB_CUDA(cudaFuncSetAttribute(k_rf_fwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rf_fwd_smem()));
    }
    if (n_codes == 0) return;
    k_rf_fwd<1><<<n_codes, rf::kNT, rf_fwd_smem(), st>>>(g, nullptr, 0, 0, take, codes, code_len, W, out);
    R4WB_LAUNCH_CHECK();
}

// R4WB_ACQ_TMEM=0 keeps the forward spectrum out of Tensor Memory (one CTA per (row, code), both operands from L2)
bool rf_use_tmem()
{
    static int v = -1;
    if (v < 0) { const char* e = std::getenv("R4WB_ACQ_TMEM"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

void launch_rf_inv_peak(const AcqGeom& g, uint32_t rows, const cf* X, const cf* C, const cf* W, RowPeak* peaks, cudaStream_t st)
{
    if (rf_use_tmem()) {
        static PerDeviceOnce attr_tm;
        if (attr_tm.first()) {
            R4WB_CUDA(cudaFuncSetAttribute(k_rf_inv_peak_tm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rf_inv_smem()));
        }
        if (rows == 0) return;
        k_rf_inv_peak_tm<<<rows, rf::kNT, rf_inv_smem(), st>>>(g, X, C, W, peaks);
        R4WB_LAUNCH_CHECK();
        return;
    }
    static PerDeviceOnce attr;
    if (attr.first()) {
        R4WB_CUDA(cudaFuncSetAttribute(k_rf_inv_peak, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rf_inv_smem()));
    }
    const uint64_t items = (uint64_t)rows * g.P;
    if (items == 0) return;
    if (items > 0x7fffffffull) fail(R4WB_ERR_INVALID_SIZE, "too many FFT rows in one launch");
    k_rf_inv_peak<<<(unsigned)items, rf::kNT, rf_inv_smem(), st>>>(g, X, C, W, peaks);
    R4WB_LAUNCH_CHECK();
}

}  // namespace r4wb
