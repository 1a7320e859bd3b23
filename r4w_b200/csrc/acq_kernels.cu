// acq_kernels.cu — sm_100a kernels of PCPS acquisition (gnss/acquisition.rs:104-249).
//
//   k_twiddles      exp(-2 pi i t / N) table (f64 sincospi, rounded once to the working precision)
//   k_fwd           per (snapshot, Doppler) or per code: [wipe-off | int8 replica] -> fold -> M-point FFT in
//                   shared memory -> natural-order spectrum to HBM/L2           (acquisition.rs:109-117, 133-143)
//   k_inv_peak      per (snapshot, Doppler, code, residue): X * conj(C) / N fused into the fold load ->
//                   M-point inverse FFT in shared memory -> |.|^2, sum and first-max for lags < code_length,
//                   warp-shuffle + CTA reduction -> one RowPeak                  (acquisition.rs:146-164)
//   k_pair_reduce   per (snapshot, code): merge the Doppler x residue RowPeaks honouring the reference's
//                   scan order (lowest d * code_length + lag wins ties)
//
// No tensor cores: a length-32768 FFT row is not a dense contraction.  The bound is shared-memory / FP32
// throughput with the spectra L2-resident.
#include <cuda_runtime.h>

#include "acq.cuh"

namespace r4wb {

template <typename T>
__global__ void k_twiddles(cx<T>* __restrict__ W, uint32_t N)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N) return;
    double s, c;
    sincospi(-2.0 * (double)t / (double)N, &s, &c);
    W[t] = cx<T>{(T)c, (T)s};
}

template <typename T> struct FftThreads { static constexpr int value = sizeof(T) == 4 ? 512 : 256; };

// fold + all passes; on return s[] holds the item's outputs at digit-reversed positions (CTA-synchronised)
template <int SIGN, typename T, typename Load>
__device__ __forceinline__ void transform_item(cx<T>* s, const Load& load, const AcqGeom& g, uint32_t r, const cx<T>* __restrict__ W)
{
    constexpr uint32_t NT = FftThreads<T>::value;
    const uint32_t M = 1u << g.logM;
#pragma unroll 4
    for (uint32_t k = threadIdx.x; k < M; k += NT) st_cx(s + fft_pad(k), fft_fold_point<SIGN, T>(load, k, g.logM, g.logF, r, g.logN, W));
    __syncthreads();
    const int np = fft_num_passes(g.logM);
    for (int p = 0; p < np; ++p) {
        fft_pass<SIGN, T>(s, g.logM, g.logN, W, p, threadIdx.x, NT);
        __syncthreads();
    }
}

// MODE 0: rows are (snapshot, Doppler) pairs of the wiped-off input; MODE 1: rows are local replicas, output conjugated
template <typename T, int MODE>
__global__ void __launch_bounds__(FftThreads<T>::value)
k_fwd(AcqGeom g, const void* __restrict__ input, uint32_t in64, uint64_t stride, uint32_t take, const int8_t* __restrict__ codes,
      uint64_t code_len, const cx<T>* __restrict__ W, cx<T>* __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<T>* s = reinterpret_cast<cx<T>*>(smem_raw);
    constexpr uint32_t NT = FftThreads<T>::value;
    const uint32_t F = 1u << g.logF, M = 1u << g.logM;
    const uint32_t row = blockIdx.x >> g.logF, r = blockIdx.x & (F - 1u);
    if (MODE == 0) {
        const uint32_t snap = row / g.D, d = row - snap * g.D;
        WipeLoad<T> ld;
        const size_t bps = in64 ? 16 : 8;
        ld.in = static_cast<const unsigned char*>(input) + (size_t)snap * stride * bps;
        ld.in64 = in64; ld.take = take;
        ld.doppler = -g.dmax + (double)d * g.dstep;       // acquisition.rs:127-130
        ld.fs = g.fs;
        transform_item<-1, T>(s, ld, g, r, W);
    } else {
        CodeLoad<T> ld;
        ld.code = codes + (size_t)row * code_len;
        ld.take = take;
        transform_item<-1, T>(s, ld, g, r, W);
    }
    cx<T>* o = out + (size_t)row * g.N;
    for (uint32_t m = threadIdx.x; m < M; m += NT) {
        cx<T> v = ld_cx(s + fft_pad(nat_to_pos(m, g.logM)));
        if (MODE == 1) { const T inv_n = (T)1 / (T)g.N; v = cx<T>{v.re * inv_n, -v.im * inv_n}; }   // conj, and ifft's 1/N
        o[(m << g.logF) + r] = v;
    }
}

template <typename V>
__device__ __forceinline__ PeakAcc<V> peak_shfl_xor(const PeakAcc<V>& a, int off)
{
    PeakAcc<V> o;
    o.best = __shfl_xor_sync(0xffffffffu, a.best, off);
    o.second = __shfl_xor_sync(0xffffffffu, a.second, off);
    o.sum = __shfl_xor_sync(0xffffffffu, a.sum, off);
    o.idx = __shfl_xor_sync(0xffffffffu, a.idx, off);
    return o;
}

// Fold of the spectrum product for F = 2, f32 (the E1C shape, N = 32768): z[k] = (Y[k] +- Y[k+M]) * w^{kr} / N with
// Y = X * C.  Each thread owns element PAIRS so every global load is 16 bytes, and all loads of U iterations are
// issued before the first use (the generic fold stores to shared memory between loads, which the compiler must
// keep ordered against the un-restricted global pointers — that serialised every L2 round trip).
__device__ __forceinline__ void fold_product_f2(cx<float>* s, const cx<float>* __restrict__ X, const cx<float>* __restrict__ C,
                                                const cx<float>* __restrict__ W, int logM, uint32_t r)
{
    constexpr uint32_t NT = FftThreads<float>::value;
    constexpr int U = 4;
    const uint32_t M = 1u << logM;
    const float4* X4 = reinterpret_cast<const float4*>(X);
    const float4* C4 = reinterpret_cast<const float4*>(C);
    const float4* W4 = reinterpret_cast<const float4*>(W);
    const uint32_t half = M >> 1;                      // float4 index of element M
    for (uint32_t i0 = threadIdx.x; i0 < half; i0 += U * NT) {
        float4 x0[U], x1[U], c0[U], c1[U], w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t i = i0 + (uint32_t)u * NT;
            if (i < half) {
                x0[u] = __ldg(X4 + i); x1[u] = __ldg(X4 + i + half);
                c0[u] = __ldg(C4 + i); c1[u] = __ldg(C4 + i + half);
                if (r) w[u] = __ldg(W4 + i);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t i = i0 + (uint32_t)u * NT;
            if (i >= half) break;
            // two complex products per operand pair
            const float a0r = x0[u].x * c0[u].x - x0[u].y * c0[u].y, a0i = x0[u].x * c0[u].y + x0[u].y * c0[u].x;
            const float a1r = x0[u].z * c0[u].z - x0[u].w * c0[u].w, a1i = x0[u].z * c0[u].w + x0[u].w * c0[u].z;
            const float b0r = x1[u].x * c1[u].x - x1[u].y * c1[u].y, b0i = x1[u].x * c1[u].y + x1[u].y * c1[u].x;
            const float b1r = x1[u].z * c1[u].z - x1[u].w * c1[u].w, b1i = x1[u].z * c1[u].w + x1[u].w * c1[u].z;
            cx<float> z0, z1;
            if (r == 0) {
                z0 = cx<float>{a0r + b0r, a0i + b0i};
                z1 = cx<float>{a1r + b1r, a1i + b1i};
            } else {                                   // (Y[k] - Y[k+M]) * conj(W[k])
                const float d0r = a0r - b0r, d0i = a0i - b0i;
                const float d1r = a1r - b1r, d1i = a1i - b1i;
                z0 = cx<float>{d0r * w[u].x + d0i * w[u].y, d0i * w[u].x - d0r * w[u].y};
                z1 = cx<float>{d1r * w[u].z + d1i * w[u].w, d1i * w[u].z - d1r * w[u].w};
            }
            const uint32_t p = fft_pad(2u * i);        // 2i and 2i+1 share a pad group
            st_cx(s + p, z0);
            st_cx(s + p + 1, z1);
        }
    }
}

// blockIdx.x = ((row * P) + code) * F + r, row = snapshot * D + d
template <typename T>
__global__ void __launch_bounds__(FftThreads<T>::value)
k_inv_peak(AcqGeom g, const cx<T>* __restrict__ X, const cx<T>* __restrict__ C, const cx<T>* __restrict__ W,
           RowPeak* __restrict__ peaks, double* __restrict__ grid)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<T>* s = reinterpret_cast<cx<T>*>(smem_raw);
    constexpr uint32_t NT = FftThreads<T>::value;
    __shared__ PeakAcc<T> s_red[NT / 32];
    __shared__ uint32_t s_hi[64];
    const uint32_t F = 1u << g.logF, M = 1u << g.logM;
    const uint32_t r = blockIdx.x & (F - 1u);
    const uint32_t rc = blockIdx.x >> g.logF;
    const uint32_t row = rc / g.P, code = rc - row * g.P;

    // digit reversal is a bit permutation: pos_to_nat(tid | i * NT) = pos_to_nat(tid) | pos_to_nat(i * NT)
    if (threadIdx.x < 64 && threadIdx.x * NT < M) s_hi[threadIdx.x] = pos_to_nat(threadIdx.x * NT, g.logM);

    ProductLoad<T> ld;
    ld.x = X + (size_t)row * g.N;
    ld.c = C + (size_t)code * g.N;
    if (sizeof(T) == 4 && g.logF == 1 && g.logM >= 2) {
        fold_product_f2(reinterpret_cast<cx<float>*>(s), reinterpret_cast<const cx<float>*>(ld.x), reinterpret_cast<const cx<float>*>(ld.c),
                        reinterpret_cast<const cx<float>*>(W), g.logM, r);
        __syncthreads();
        const int np = fft_num_passes(g.logM);
        for (int p = 0; p < np; ++p) {
            fft_pass<+1, T>(s, g.logM, g.logN, W, p, threadIdx.x, NT);
            __syncthreads();
        }
    } else {
        transform_item<+1, T>(s, ld, g, r, W);
    }

    PeakAcc<T> acc;
    peak_init(acc);
    const uint32_t lo = threadIdx.x < M ? pos_to_nat(threadIdx.x, g.logM) : 0u;
    for (uint32_t p = threadIdx.x, i = 0; p < M; p += NT, ++i) {
        const uint32_t m = M > NT ? (lo | s_hi[i]) : lo;
        const uint32_t n = (m << g.logF) + r;                              // lag
        if (n < g.L) {
            const cx<T> v = ld_cx(s + fft_pad(p));
            const T mag = v.re * v.re + v.im * v.im;
            peak_push(acc, mag, n);
            if (grid) grid[(size_t)(row % g.D) * g.L + n] = (double)mag;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) peak_merge(acc, peak_shfl_xor(acc, off));
    if ((threadIdx.x & 31u) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        PeakAcc<T> a;
        peak_init(a);
        if (threadIdx.x < NT / 32) a = s_red[threadIdx.x];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) peak_merge(a, peak_shfl_xor(a, off));
        if (threadIdx.x == 0) {
            RowPeak o;
            o.best = (double)a.best; o.second = (double)a.second; o.sum = (double)a.sum; o.lag = a.idx; o.pad = 0;
            peaks[blockIdx.x] = o;
        }
    }
}

// one warp per (snapshot, code)
__global__ void k_pair_reduce(AcqGeom g, uint32_t n_snap, const RowPeak* __restrict__ peaks, PairPeak* __restrict__ out)
{
    const uint32_t pair = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pair >= n_snap * g.P) return;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t snap = pair / g.P, code = pair - snap * g.P;
    const uint32_t F = 1u << g.logF;
    PeakAcc<double> acc;
    peak_init(acc);
    for (uint32_t it = lane; it < g.D * F; it += 32) {
        const uint32_t d = it >> g.logF, r = it & (F - 1u);
        const RowPeak rp = peaks[(((size_t)snap * g.D + d) * g.P + code) * F + r];
        PeakAcc<double> o;
        o.best = rp.best; o.second = rp.second; o.sum = rp.sum;
        o.idx = rp.lag == 0xffffffffu ? 0xffffffffu : d * g.L + rp.lag;
        peak_merge(acc, o);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) peak_merge(acc, peak_shfl_xor(acc, off));
    if (lane == 0) {
        PairPeak o;
        o.best = acc.best; o.second = acc.second; o.sum = acc.sum; o.lin = acc.idx; o.pad = 0;
        out[pair] = o;
    }
}

// ---------------------------------------------------------------------------------------------- launchers
template <typename T>
size_t fft_smem_bytes(int logM) { return (size_t)fft_padded_len(1u << logM) * sizeof(cx<T>); }

template <typename T>
void launch_twiddles(cx<T>* W, uint32_t N, cudaStream_t st)
{
    k_twiddles<T><<<(N + 255) / 256, 256, 0, st>>>(W, N);
    R4WB_LAUNCH_CHECK();
}

template <typename T, int MODE>
static void launch_fwd_t(const AcqGeom& g, uint32_t rows, const void* input, uint32_t in64, uint64_t stride, uint32_t take,
                         const int8_t* codes, uint64_t code_len, const cx<T>* W, cx<T>* out, cudaStream_t st)
{
    static PerDeviceOnce attr;
    if (attr.first()) {
        R4WB_CUDA(cudaFuncSetAttribute(k_fwd<T, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    if (rows == 0) return;
    k_fwd<T, MODE><<<rows << g.logF, FftThreads<T>::value, fft_smem_bytes<T>(g.logM), st>>>(g, input, in64, stride, take, codes, code_len, W, out);
    R4WB_LAUNCH_CHECK();
}

template <typename T>
void launch_fwd_input(const AcqGeom& g, uint32_t rows, const void* input, uint32_t in64, uint64_t stride, uint32_t take,
                      const cx<T>* W, cx<T>* out, cudaStream_t st)
{
    launch_fwd_t<T, 0>(g, rows, input, in64, stride, take, nullptr, 0, W, out, st);
}
template <typename T>
void launch_fwd_codes(const AcqGeom& g, uint32_t n_codes, const int8_t* codes, uint64_t code_len, uint32_t take, const cx<T>* W,
                      cx<T>* out, cudaStream_t st)
{
    launch_fwd_t<T, 1>(g, n_codes, nullptr, 0, 0, take, codes, code_len, W, out, st);
}

template <typename T>
void launch_inv_peak(const AcqGeom& g, uint32_t rows, const cx<T>* X, const cx<T>* C, const cx<T>* W, RowPeak* peaks, double* grid,
                     cudaStream_t st)
{
    static PerDeviceOnce attr;
    if (attr.first()) {
        R4WB_CUDA(cudaFuncSetAttribute(k_inv_peak<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    const uint64_t items = ((uint64_t)rows * g.P) << g.logF;
    if (items == 0) return;
    if (items > 0x7fffffffull) fail(R4WB_ERR_INVALID_SIZE, "too many FFT rows in one launch");
    k_inv_peak<T><<<(unsigned)items, FftThreads<T>::value, fft_smem_bytes<T>(g.logM), st>>>(g, X, C, W, peaks, grid);
    R4WB_LAUNCH_CHECK();
}

void launch_pair_reduce(const AcqGeom& g, uint32_t n_snap, const RowPeak* peaks, PairPeak* out, cudaStream_t st)
{
    const uint32_t pairs = n_snap * g.P;
    if (pairs == 0) return;
    k_pair_reduce<<<(pairs + 3) / 4, 128, 0, st>>>(g, n_snap, peaks, out);
    R4WB_LAUNCH_CHECK();
}

#define R4WB_INST(T)                                                                                                             \
    template void launch_twiddles<T>(cx<T>*, uint32_t, cudaStream_t);                                                            \
    template void launch_fwd_input<T>(const AcqGeom&, uint32_t, const void*, uint32_t, uint64_t, uint32_t, const cx<T>*, cx<T>*, \
                                      cudaStream_t);                                                                             \
    template void launch_fwd_codes<T>(const AcqGeom&, uint32_t, const int8_t*, uint64_t, uint32_t, const cx<T>*, cx<T>*,        \
                                      cudaStream_t);                                                                             \
    template void launch_inv_peak<T>(const AcqGeom&, uint32_t, const cx<T>*, const cx<T>*, const cx<T>*, RowPeak*, double*,      \
                                     cudaStream_t);
R4WB_INST(float)
R4WB_INST(double)

}  // namespace r4wb
