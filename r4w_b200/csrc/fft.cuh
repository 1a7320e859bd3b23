// fft.cuh — power-of-two complex FFT engine of the acquisition path (stands in for rustfft behind
// FftProcessor::fft_inplace / ifft_inplace, core/fft_utils.rs:85-108).
//
// Shape: a length-N transform is cut into F independent work items by one decimation-in-frequency radix-F
// stage that is fused into the load ("fold"): item r (0 <= r < F) gathers  z[k] = w_N^{kr} * sum_j
// w_F^{jr} y[k + j M],  M = N/F, runs an M-point transform entirely in shared memory and ends up holding the
// outputs with index  n = F m + r.  The M-point transform is an in-place DIF with radix-16 passes (a last
// radix-2/4/8 pass mops up), each butterfly living in registers; results are left in digit-reversed
// positions (pos_to_nat maps a position back to m), which the fused epilogues do not mind.
// Everything is __host__ __device__ so tests/emu/ can replay the exact same arithmetic without a GPU.
#pragma once
#include <cstdint>

#include "geom.hpp"   // R4WB_HD

namespace r4wb {

template <typename T>
struct alignas(2 * sizeof(T)) cx {
    T re, im;
};

template <typename T> R4WB_HD cx<T> operator+(cx<T> a, cx<T> b) { return {a.re + b.re, a.im + b.im}; }
template <typename T> R4WB_HD cx<T> operator-(cx<T> a, cx<T> b) { return {a.re - b.re, a.im - b.im}; }
template <typename T> R4WB_HD cx<T> operator*(cx<T> a, cx<T> b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
template <typename T> R4WB_HD cx<T> cconj(cx<T> a) { return {a.re, -a.im}; }

// one 8-byte (cf32) / 16-byte (cf64) access per complex element (the plain struct copy is split into two scalar
// accesses by the compiler)
R4WB_HD cx<float> ld_cx(const cx<float>* p)
{
#ifdef __CUDA_ARCH__
    const float2 v = *reinterpret_cast<const float2*>(p);
    return cx<float>{v.x, v.y};
#else
    return *p;
#endif
}
R4WB_HD cx<double> ld_cx(const cx<double>* p)
{
#ifdef __CUDA_ARCH__
    const double2 v = *reinterpret_cast<const double2*>(p);
    return cx<double>{v.x, v.y};
#else
    return *p;
#endif
}
R4WB_HD void st_cx(cx<float>* p, cx<float> v)
{
#ifdef __CUDA_ARCH__
    *reinterpret_cast<float2*>(p) = make_float2(v.re, v.im);
#else
    *p = v;
#endif
}
R4WB_HD void st_cx(cx<double>* p, cx<double> v)
{
#ifdef __CUDA_ARCH__
    *reinterpret_cast<double2*>(p) = make_double2(v.re, v.im);
#else
    *p = v;
#endif
}

// shared-memory position of logical element p: one pad element per 16 keeps every pass's half-warp
// accesses on distinct 8-byte banks (strides 2^k collide otherwise)
R4WB_HD uint32_t fft_pad(uint32_t p) { return p + (p >> 4); }
R4WB_HD uint32_t fft_padded_len(uint32_t M) { return M + (M >> 4) + 1; }

// digit-reversed position -> natural index m of an M = 2^logM point transform done with radix-16 passes
// followed by one radix-2^(logM % 4) pass
R4WB_HD uint32_t pos_to_nat(uint32_t p, int logM)
{
    uint32_t m = 0;
    int rem = logM, out = 0;
    while (rem > 0) {
        const int bits = rem >= 4 ? 4 : rem;
        rem -= bits;
        m |= ((p >> rem) & ((1u << bits) - 1u)) << out;
        out += bits;
    }
    return m;
}
R4WB_HD uint32_t nat_to_pos(uint32_t m, int logM)
{
    uint32_t p = 0;
    int rem = logM, in = 0;
    while (rem > 0) {
        const int bits = rem >= 4 ? 4 : rem;
        rem -= bits;
        p |= ((m >> in) & ((1u << bits) - 1u)) << rem;
        in += bits;
    }
    return p;
}

// exp(SIGN * 2 pi i t / 16), t = 0..7
template <typename T, int SIGN>
R4WB_HD cx<T> w16(int t)
{
    constexpr double c[8] = {1.0, 0.92387953251128673848, 0.70710678118654752440, 0.38268343236508977173,
                             0.0, -0.38268343236508977173, -0.70710678118654752440, -0.92387953251128673848};
    constexpr double s[8] = {0.0, 0.38268343236508977173, 0.70710678118654752440, 0.92387953251128673848,
                             1.0, 0.92387953251128673848, 0.70710678118654752440, 0.38268343236508977173};
    return {(T)c[t], (T)(SIGN * s[t])};
}

// In-register R-point DFT (R = 2, 4, 8, 16), decimation in frequency: on return X[i] sits in a[bitrev_R(i)].
template <int R, int SIGN, typename T>
R4WB_HD void dft_bitrev(cx<T>* a)
{
#pragma unroll
    for (int len = R; len >= 2; len >>= 1) {
        const int half = len >> 1;
#pragma unroll
        for (int base = 0; base < R; base += len) {
#pragma unroll
            for (int k = 0; k < half; ++k) {
                const cx<T> u = a[base + k], v = a[base + k + half];
                a[base + k] = u + v;
                const cx<T> d = u - v;
                const int t = k * (16 / len);      // twiddle exp(SIGN 2 pi i k / len)
                if (t == 0) a[base + k + half] = d;
                else if (t == 4) a[base + k + half] = cx<T>{(T)(-SIGN) * d.im, (T)SIGN * d.re};
                else a[base + k + half] = d * w16<T, SIGN>(t);
            }
        }
    }
}

template <int R>
R4WB_HD constexpr int bitrev_r(int i)
{
    int r = 0;
    for (int b = 1, o = R >> 1; b < R; b <<= 1, o >>= 1)
        if (i & b) r |= o;
    return r;
}

// One radix-R butterfly of an in-place DIF pass over s[0..M) (padded layout).
//   b      butterfly index in [0, M/R)
//   logL   log2 of the current sub-transform length L (this pass splits it into R pieces of L/R)
//   W      table exp(-2 pi i t / N), t in [0, N); logN - logL turns a twiddle exponent into a table index
template <int R, int SIGN, typename T>
R4WB_HD void fft_butterfly(cx<T>* s, uint32_t b, int logL, int logN, const cx<T>* __restrict__ W)
{
    constexpr int LOGR = R == 16 ? 4 : R == 8 ? 3 : R == 4 ? 2 : 1;
    const int logst = logL - LOGR;
    const uint32_t st = 1u << logst;
    const uint32_t g = b >> logst, q = b & (st - 1u);
    const uint32_t base = (g << logL) + q;
    cx<T> a[R];
#pragma unroll
    for (int j = 0; j < R; ++j) a[j] = ld_cx(s + fft_pad(base + (uint32_t)j * st));
    dft_bitrev<R, SIGN, T>(a);
    if (logst == 0) {           // last pass: all twiddles are 1
#pragma unroll
        for (int i = 0; i < R; ++i) st_cx(s + fft_pad(base + (uint32_t)i * st), a[bitrev_r<R>(i)]);
        return;
    }
    // w[i] = w_L^{i q}; built as a product tree (depth <= 4) so rounding does not pile up
    cx<T> w[R];
    {
        cx<T> w1 = W[(q << (logN - logL)) & ((1u << logN) - 1u)];
        if (SIGN > 0) w1 = cconj(w1);
        w[0] = cx<T>{(T)1, (T)0};
        w[1] = w1;
        if (R > 2) { w[2 % R] = w1 * w1; w[3 % R] = w[2 % R] * w1; }
        if (R > 4) {
            w[4 % R] = w[2 % R] * w[2 % R];
#pragma unroll
            for (int i = 1; i < 4; ++i) w[(4 + i) % R] = w[4 % R] * w[i];
        }
        if (R > 8) {
            w[8 % R] = w[4 % R] * w[4 % R];
#pragma unroll
            for (int i = 1; i < 8; ++i) w[(8 + i) % R] = w[8 % R] * w[i];
        }
    }
    st_cx(s + fft_pad(base), a[0]);
#pragma unroll
    for (int i = 1; i < R; ++i) st_cx(s + fft_pad(base + (uint32_t)i * st), a[bitrev_r<R>(i)] * w[i]);
}

// All butterflies of pass `pass_idx` that thread `tid` of `nthreads` owns.  Returns false once the
// transform is complete (no such pass).  The caller synchronises between passes.
template <int SIGN, typename T>
R4WB_HD bool fft_pass(cx<T>* s, int logM, int logN, const cx<T>* __restrict__ W, int pass_idx, uint32_t tid, uint32_t nthreads)
{
    const int logL = logM - 4 * pass_idx;
    if (logL <= 0) return false;
    const uint32_t M = 1u << logM;
    if (logL >= 4) {
        for (uint32_t b = tid; b < (M >> 4); b += nthreads) fft_butterfly<16, SIGN, T>(s, b, logL, logN, W);
    } else if (logL == 3) {
        for (uint32_t b = tid; b < (M >> 3); b += nthreads) fft_butterfly<8, SIGN, T>(s, b, logL, logN, W);
    } else if (logL == 2) {
        for (uint32_t b = tid; b < (M >> 2); b += nthreads) fft_butterfly<4, SIGN, T>(s, b, logL, logN, W);
    } else {
        for (uint32_t b = tid; b < (M >> 1); b += nthreads) fft_butterfly<2, SIGN, T>(s, b, logL, logN, W);
    }
    return true;
}
R4WB_HD int fft_num_passes(int logM) { return (logM + 3) / 4; }

// Fused radix-F fold: z[k] for item r.  `load(n)` returns y[n], n in [0, N).
template <int SIGN, typename T, typename Load>
R4WB_HD cx<T> fft_fold_point(Load&& load, uint32_t k, int logM, int logF, uint32_t r, int logN, const cx<T>* __restrict__ W)
{
    const uint32_t F = 1u << logF, nmask = (1u << logN) - 1u;
    cx<T> acc = load(k);
    for (uint32_t j = 1; j < F; ++j) {
        cx<T> v = load(k + (j << logM));
        if (r != 0) {
            cx<T> wf = W[((j * r) << (logN - logF)) & nmask];     // w_F^{jr}
            if (SIGN > 0) wf = cconj(wf);
            v = v * wf;
        }
        acc = acc + v;
    }
    if (r != 0) {
        cx<T> wk = W[(k * r) & nmask];                             // w_N^{kr}
        if (SIGN > 0) wk = cconj(wk);
        acc = acc * wk;
    }
    return acc;
}

// largest M the engine keeps in one CTA's shared memory (padded): 16384 cf32 / 8192 cf64 = 136 KiB
template <typename T> R4WB_HD constexpr int fft_max_logM() { return sizeof(T) == 4 ? 14 : 13; }

}  // namespace r4wb
