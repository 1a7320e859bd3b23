// rfft.cuh — register-resident 16384-point complex FFT for one 512-thread CTA (the E1C shape: fft_size 32768 = 2 x 16384).
//
// Stands in for rustfft behind FftProcessor::fft_inplace / ifft_inplace (core/fft_utils.rs:85-108) on the f32 fast
// path of PcpsAcquisition (gnss/acquisition.rs:104-195).  Every thread keeps 32 points in registers; the transform is
// 32 x 32 x 16 with two exchanges through shared memory (real and imaginary parts in turn, 66 KB), so a point crosses
// shared memory twice instead of the nine times of the in-shared-memory engine (fft.cuh), and all butterflies are
// straight-line register code (packed FADD2 for the complex adds).
//
//   forward  (DIF): natural-order input  z[t + 512 j] in a[j]  ->  spectrum left in registers, "slot" order
//   inverse  (DIT): slot-order input in registers  ->  natural-order output z[t + 512 j] in a[bitrev5(j)]
// The slot order is whatever the forward transform ends with: thread t, register rho = 16 gi + i holds
// X[k1 + 32 k2 + 1024 k3] with k1 = 2 (t >> 5) + gi, k2 = t & 31, k3 = bitrev4(i).  Spectra are stored in HBM/L2
// in that order (element (rho, t) at ((rho >> 1) * 512 + t) * 2 + (rho & 1): one float4 per thread and register
// pair), the code spectra too, so the element-wise product needs no reordering and the inverse transform is the
// exact transpose of the forward one.
#pragma once
#include <cstdint>

#include "fft.cuh"

namespace r4wb {
namespace rf {

typedef cx<float> cf;

constexpr int kNT = 512;                       // threads per CTA
constexpr int kLogM = 14;
constexpr int kM = 1 << kLogM;                 // points per transform
constexpr int kRow = 528;                      // floats per k1 row of the exchange buffer: 512 + 16 puts the next row on the other 16 banks
constexpr int kXbufFloats = 32 * kRow;         // 16 896 floats = 67 584 bytes (twice that for the complex layout)

#ifdef __CUDACC__

__device__ __forceinline__ cf cadd(cf a, cf b)
{
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
    return reinterpret_cast<cf&>(d);
}
__device__ __forceinline__ cf csub(cf a, cf b)
{
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
    return reinterpret_cast<cf&>(d);
}
__device__ __forceinline__ cf cmul(cf a, cf b) { return cf{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ cf csqr(cf a) { return cf{a.re * a.re - a.im * a.im, 2.0f * a.re * a.im}; }

__host__ __device__ constexpr int bitrev5(int i) { return ((i & 1) << 4) | ((i & 2) << 2) | (i & 4) | ((i & 8) >> 2) | ((i & 16) >> 4); }
__host__ __device__ constexpr int bitrev4(int i) { return ((i & 1) << 3) | ((i & 2) << 1) | ((i & 4) >> 1) | ((i & 8) >> 3); }
template <int R> __host__ __device__ constexpr int bitrev(int i) { return R == 32 ? bitrev5(i) : bitrev4(i); }

// exp(SIGN 2 pi i T / 32), T = 0..15, as compile-time constants (template argument: always folded into immediates)
template <int T> struct W32 {
    static constexpr float c = T == 0 ? 1.0f : T == 1 ? 0.98078528040323044913f : T == 2 ? 0.92387953251128675613f : T == 3 ? 0.83146961230254523708f :
                               T == 4 ? 0.70710678118654752440f : T == 5 ? 0.55557023301960222474f : T == 6 ? 0.38268343236508977173f :
                               T == 7 ? 0.19509032201612826785f : T == 8 ? 0.0f : T == 9 ? -0.19509032201612826785f : T == 10 ? -0.38268343236508977173f :
                               T == 11 ? -0.55557023301960222474f : T == 12 ? -0.70710678118654752440f : T == 13 ? -0.83146961230254523708f :
                               T == 14 ? -0.92387953251128675613f : -0.98078528040323044913f;
    static constexpr float s = T == 0 ? 0.0f : T == 1 ? 0.19509032201612826785f : T == 2 ? 0.38268343236508977173f : T == 3 ? 0.55557023301960222474f :
                               T == 4 ? 0.70710678118654752440f : T == 5 ? 0.83146961230254523708f : T == 6 ? 0.92387953251128675613f :
                               T == 7 ? 0.98078528040323044913f : T == 8 ? 1.0f : T == 9 ? 0.98078528040323044913f : T == 10 ? 0.92387953251128675613f :
                               T == 11 ? 0.83146961230254523708f : T == 12 ? 0.70710678118654752440f : T == 13 ? 0.55557023301960222474f :
                               T == 14 ? 0.38268343236508977173f : 0.19509032201612826785f;
};

// v * exp(SIGN 2 pi i T / 32)
template <int SIGN, int T>
__device__ __forceinline__ cf mul_w32(cf v)
{
    if (T == 0) return v;
    if (T == 8) return cf{(float)(-SIGN) * v.im, (float)SIGN * v.re};
    return cmul(v, cf{W32<T>::c, (float)SIGN * W32<T>::s});
}

// butterflies (base + K, base + K + LEN/2) of one stage, K and BASE unrolled at compile time
template <int R, int SIGN, int LEN, int BASE, int K, bool DIT>
struct Bfly {
    static __device__ __forceinline__ void run(cf* a)
    {
        constexpr int half = LEN / 2, i0 = BASE + K, i1 = BASE + K + half, T = K * (32 / LEN);
        if (DIT) {
            const cf u = a[i0], tv = mul_w32<SIGN, T>(a[i1]);
            a[i0] = cadd(u, tv);
            a[i1] = csub(u, tv);
        } else {
            const cf u = a[i0], v = a[i1];
            a[i0] = cadd(u, v);
            a[i1] = mul_w32<SIGN, T>(csub(u, v));
        }
        if (K + 1 < half) Bfly<R, SIGN, LEN, BASE, (K + 1 < half ? K + 1 : 0), DIT>::run(a);
        else if (BASE + LEN < R) Bfly<R, SIGN, LEN, (BASE + LEN < R ? BASE + LEN : 0), 0, DIT>::run(a);
    }
};

// R-point DFT in registers, decimation in frequency: natural-order input, X[k] left in a[bitrev<R>(k)]
template <int R, int SIGN, int LEN = R>
struct Dif {
    static __device__ __forceinline__ void run(cf* a)
    {
        Bfly<R, SIGN, LEN, 0, 0, false>::run(a);
        Dif<R, SIGN, LEN / 2>::run(a);
    }
};
template <int R, int SIGN> struct Dif<R, SIGN, 1> { static __device__ __forceinline__ void run(cf*) {} };
template <int R, int SIGN> __device__ __forceinline__ void dif(cf* a) { Dif<R, SIGN>::run(a); }

// R-point DFT in registers, decimation in time: input x[n] in a[bitrev<R>(n)], natural-order output
template <int R, int SIGN, int LEN = 2>
struct Dit {
    static __device__ __forceinline__ void run(cf* a)
    {
        Bfly<R, SIGN, LEN, 0, 0, true>::run(a);
        Dit<R, SIGN, LEN * 2>::run(a);
    }
};
template <int SIGN> struct Dit<16, SIGN, 32> { static __device__ __forceinline__ void run(cf*) {} };
template <int SIGN> struct Dit<32, SIGN, 64> { static __device__ __forceinline__ void run(cf*) {} };
template <int R, int SIGN> __device__ __forceinline__ void dit(cf* a) { Dit<R, SIGN>::run(a); }

// a[POS(k)] *= f * w^k for k = 0..R-1 (f = 1 when HAS_F is false): powers by a two-level product tree (depth <= 5)
template <int R, bool BITREV, bool HAS_F>
__device__ __forceinline__ void apply_powers(cf* a, cf w, cf f)
{
    constexpr int LO = R == 32 ? 8 : 4;                 // low digit
    cf p[LO];                                           // f * w^lo
    p[1] = w;
    p[2] = csqr(w);
    p[3] = cmul(p[2], w);
    cf q1;                                              // w^LO
    if (LO == 8) {
        p[4] = csqr(p[2]);
        p[5] = cmul(p[4], p[1]); p[6] = cmul(p[4], p[2]); p[7] = cmul(p[4], p[3]);
        q1 = csqr(p[4]);
    } else {
        q1 = csqr(p[2]);
    }
    const cf q2 = csqr(q1), q3 = cmul(q2, q1);
    if (HAS_F) {
        p[0] = f;
#pragma unroll
        for (int i = 1; i < LO; ++i) p[i] = cmul(p[i], f);
    }
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const int lo = k % LO, hi = k / LO;
        const int pos = BITREV ? bitrev<R>(k) : k;
        if (hi == 0) {
            if (lo != 0 || HAS_F) a[pos] = cmul(a[pos], p[lo]);
        } else {
            const cf q = hi == 1 ? q1 : hi == 2 ? q2 : q3;
            if (lo == 0 && !HAS_F) a[pos] = cmul(a[pos], q);
            else a[pos] = cmul(a[pos], cmul(q, p[lo]));
        }
    }
}

// ---- exchange buffer addressing (floats)
// exchange 1: element (k1, t)           t = 0..511
__device__ __forceinline__ uint32_t x1_addr(uint32_t k1, uint32_t t) { return k1 * (uint32_t)kRow + t; }
// exchange 2: element (g = k1 * 32 + k2, tl), tl = 0..15: 16-float groups, 4-float chunks xor-swizzled so both the
// (k1, tl)-threads' scalar accesses and the g-threads' 16-byte accesses are conflict-free
__device__ __forceinline__ uint32_t x2_addr(uint32_t g, uint32_t tl)
{
    return (g >> 5) * (uint32_t)kRow + (g & 31u) * 16u + ((((tl >> 2) ^ (g >> 1)) & 3u) << 2) + (tl & 3u);
}

// thread constants of the transform
struct Consts {
    cf wM;        // exp(-2 pi i t / 16384)
    cf w512;      // exp(-2 pi i (t & 15) / 512)
    cf w512k;     // exp(-2 pi i (t & 31) / 512)
};

__device__ __forceinline__ Consts load_consts(const cf* __restrict__ W /* exp(-2 pi i n / 32768) */, uint32_t t)
{
    Consts c;
    c.wM = W[2u * t];
    c.w512 = W[64u * (t & 15u)];
    c.w512k = W[64u * (t & 31u)];
    return c;
}

// Exchange 1 moves element (k1, t) between the thread that owns t and the thread that owns (k1, t & 15).  With a
// complex buffer (CPLX: 32 x kRow float2 = 135 KB) it takes one round, with the 66 KB float buffer two (re, then im).
// Exchange 2 only moves data inside a warp: warp w owns rows k1 = 2w and 2w + 1 of the buffer in both layouts, so it
// needs __syncwarp() only.
//
// Forward 16384-point DFT.  In: a[j] = z[t + 512 j].  Out (slot order): a[16 gi + i] = X[k1 + 32 k2 + 1024 bitrev4(i)]
// with k1 = 2 (t >> 5) + gi, k2 = t & 31.
template <bool CPLX>
__device__ __forceinline__ void forward(cf* a, float* xb, const Consts& c, uint32_t t)
{
    dif<32, -1>(a);                                      // over j -> k1 at a[bitrev5(k1)]
    apply_powers<32, true, false>(a, c.wM, cf{1.0f, 0.0f});     // w_M^{t k1}
    cf b[32];
    const uint32_t k1p = t >> 4, tl = t & 15u;
    if (CPLX) {
        cf* xc = reinterpret_cast<cf*>(xb);
        __syncthreads();
#pragma unroll
        for (int k1 = 0; k1 < 32; ++k1) st_cx(xc + x1_addr(k1, t), a[bitrev5(k1)]);
        __syncthreads();
#pragma unroll
        for (int j2 = 0; j2 < 32; ++j2) b[j2] = ld_cx(xc + x1_addr(k1p, tl + 16u * j2));
    } else {
#pragma unroll
        for (int comp = 0; comp < 2; ++comp) {
            __syncthreads();
#pragma unroll
            for (int k1 = 0; k1 < 32; ++k1) xb[x1_addr(k1, t)] = comp == 0 ? a[bitrev5(k1)].re : a[bitrev5(k1)].im;
            __syncthreads();
#pragma unroll
            for (int j2 = 0; j2 < 32; ++j2) {
                const float v = xb[x1_addr(k1p, tl + 16u * j2)];
                if (comp == 0) b[j2].re = v; else b[j2].im = v;
            }
        }
    }
    dif<32, -1>(b);                                      // over j2 -> k2 at b[bitrev5(k2)]
    apply_powers<32, true, false>(b, c.w512, cf{1.0f, 0.0f});   // w_512^{tl k2}
    const uint32_t gbase = (t >> 5) * 64u + (t & 31u);   // group of gi = 0; gi = 1 is 32 further (k1 + 1)
#pragma unroll
    for (int comp = 0; comp < 2; ++comp) {
        if (CPLX && comp == 0) __syncthreads();          // complex rows just read overlap other warps' float rows
        else __syncwarp();
#pragma unroll
        for (int k2 = 0; k2 < 32; ++k2) xb[x2_addr(k1p * 32u + k2, tl)] = comp == 0 ? b[bitrev5(k2)].re : b[bitrev5(k2)].im;
        __syncwarp();
#pragma unroll
        for (int gi = 0; gi < 2; ++gi) {
            const uint32_t g = gbase + 32u * gi;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                const float4 v = *reinterpret_cast<const float4*>(xb + x2_addr(g, 4u * ch));
                if (comp == 0) { a[16 * gi + 4 * ch].re = v.x; a[16 * gi + 4 * ch + 1].re = v.y; a[16 * gi + 4 * ch + 2].re = v.z; a[16 * gi + 4 * ch + 3].re = v.w; }
                else { a[16 * gi + 4 * ch].im = v.x; a[16 * gi + 4 * ch + 1].im = v.y; a[16 * gi + 4 * ch + 2].im = v.z; a[16 * gi + 4 * ch + 3].im = v.w; }
            }
        }
    }
    dif<16, -1>(a);                                      // over tl -> k3 at a[16 gi + bitrev4(k3)]
    dif<16, -1>(a + 16);
}

// Inverse (unnormalised) 16384-point DFT, the transpose of forward().  In: slot order.  Out: a[bitrev5(j)] = z[t + 512 j].
// `f` is folded into the last twiddle (a[k1] *= f * conj(w_M)^{t k1}): the caller's per-thread output factor.
template <bool HAS_F, bool CPLX>
__device__ __forceinline__ void inverse(cf* a, float* xb, const Consts& c, uint32_t t, cf f)
{
    dit<16, +1>(a);                                      // over k3 -> tl, natural
    dit<16, +1>(a + 16);
    const cf u = cconj(c.w512k);                         // conj(w_512)^{k2 tl}, k2 = t & 31 for both groups
    apply_powers<16, false, false>(a, u, cf{1.0f, 0.0f});
    apply_powers<16, false, false>(a + 16, u, cf{1.0f, 0.0f});
    cf b[32];
    const uint32_t k1p = t >> 4, tl = t & 15u;
    const uint32_t gbase = (t >> 5) * 64u + (t & 31u);
    __syncthreads();                                     // the buffer's previous readers (any row) are done
#pragma unroll
    for (int comp = 0; comp < 2; ++comp) {
        __syncwarp();
#pragma unroll
        for (int gi = 0; gi < 2; ++gi) {
            const uint32_t g = gbase + 32u * gi;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                float4 v;
                if (comp == 0) v = make_float4(a[16 * gi + 4 * ch].re, a[16 * gi + 4 * ch + 1].re, a[16 * gi + 4 * ch + 2].re, a[16 * gi + 4 * ch + 3].re);
                else v = make_float4(a[16 * gi + 4 * ch].im, a[16 * gi + 4 * ch + 1].im, a[16 * gi + 4 * ch + 2].im, a[16 * gi + 4 * ch + 3].im);
                *reinterpret_cast<float4*>(xb + x2_addr(g, 4u * ch)) = v;
            }
        }
        __syncwarp();
#pragma unroll
        for (int k2 = 0; k2 < 32; ++k2) {
            const float v = xb[x2_addr(k1p * 32u + k2, tl)];
            if (comp == 0) b[k2].re = v; else b[k2].im = v;
        }
    }
    dif<32, +1>(b);                                      // over k2 -> j2 at b[bitrev5(j2)]
    if (CPLX) {
        cf* xc = reinterpret_cast<cf*>(xb);
        __syncthreads();                                 // complex rows overlap other warps' float rows
#pragma unroll
        for (int j2 = 0; j2 < 32; ++j2) st_cx(xc + x1_addr(k1p, tl + 16u * j2), b[bitrev5(j2)]);
        __syncthreads();
#pragma unroll
        for (int k1 = 0; k1 < 32; ++k1) a[k1] = ld_cx(xc + x1_addr(k1, t));
    } else {
#pragma unroll
        for (int comp = 0; comp < 2; ++comp) {
            if (comp == 0) __syncwarp(); else __syncthreads();   // rows 2w, 2w+1 were last read by this warp / by everyone
#pragma unroll
            for (int j2 = 0; j2 < 32; ++j2) xb[x1_addr(k1p, tl + 16u * j2)] = comp == 0 ? b[bitrev5(j2)].re : b[bitrev5(j2)].im;
            __syncthreads();
#pragma unroll
            for (int k1 = 0; k1 < 32; ++k1) {
                const float v = xb[x1_addr(k1, t)];
                if (comp == 0) a[k1].re = v; else a[k1].im = v;
            }
        }
    }
    apply_powers<32, false, HAS_F>(a, cconj(c.wM), f);   // f * conj(w_M)^{t k1}
    dif<32, +1>(a);                                      // over k1 -> j at a[bitrev5(j)]
}

#endif  // __CUDACC__

}  // namespace rf
}  // namespace r4wb
