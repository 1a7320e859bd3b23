// acq_emu.cpp — HOST REPLAY of the acquisition kernels' arithmetic (r4w_b200/csrc/fft.cuh, acq.cuh).
// TEST INFRASTRUCTURE ONLY (see synth_emu.cpp): same __host__ __device__ functions as k_fwd / k_inv_peak /
// k_pair_reduce, thread loops flattened.  Never linked into or loaded by libr4w_b200.so.
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../r4w_b200/csrc/acq.cuh"

using namespace r4wb;

namespace {

template <int SIGN, typename T, typename Load>
void transform_item(std::vector<cx<T>>& s, const Load& load, const AcqGeom& g, uint32_t r, const cx<T>* W, uint32_t NT)
{
    const uint32_t M = 1u << g.logM;
    for (uint32_t tid = 0; tid < NT; ++tid)
        for (uint32_t k = tid; k < M; k += NT) s[fft_pad(k)] = fft_fold_point<SIGN, T>(load, k, g.logM, g.logF, r, g.logN, W);
    const int np = fft_num_passes(g.logM);
    for (int p = 0; p < np; ++p)
        for (uint32_t tid = 0; tid < NT; ++tid) fft_pass<SIGN, T>(s.data(), g.logM, g.logN, W, p, tid, NT);
}

template <typename T>
std::vector<cx<T>> twiddles(uint32_t N)
{
    std::vector<cx<T>> W(N);
    for (uint32_t t = 0; t < N; ++t) {
        const double a = -2.0 * kPi * (double)t / (double)N;
        W[t] = cx<T>{(T)cos(a), (T)sin(a)};
    }
    return W;
}

template <typename T>
int fft_t(int logN, int logM, int sign, const double* in, double* out)
{
    AcqGeom g{};
    g.logN = logN; g.logM = logM; g.logF = logN - logM; g.N = 1u << logN;
    const uint32_t F = 1u << g.logF, M = 1u << logM, NT = sizeof(T) == 4 ? 512 : 256;
    const auto W = twiddles<T>(g.N);
    std::vector<cx<T>> x(g.N), s(fft_padded_len(M));
    for (uint32_t n = 0; n < g.N; ++n) x[n] = cx<T>{(T)in[2 * n], (T)in[2 * n + 1]};
    auto load = [&](uint32_t n) { return x[n]; };
    for (uint32_t r = 0; r < F; ++r) {
        if (sign < 0) transform_item<-1, T>(s, load, g, r, W.data(), NT);
        else transform_item<+1, T>(s, load, g, r, W.data(), NT);
        for (uint32_t m = 0; m < M; ++m) {
            const cx<T> v = s[fft_pad(nat_to_pos(m, logM))];
            if (pos_to_nat(nat_to_pos(m, logM), logM) != m) return -1;
            out[2 * ((m << g.logF) + r)] = (double)v.re;
            out[2 * ((m << g.logF) + r) + 1] = (double)v.im;
        }
    }
    return 0;
}

template <typename T>
int pcps_t(uint64_t code_length, double fs, double dmax, double dstep, const void* input, int in64, uint64_t n_input,
           const int8_t* code, uint64_t code_len, double* out4, double* grid)
{
    AcqGeom g{};
    uint64_t f = 1;
    while (f < code_length) f <<= 1;
    g.N = (uint32_t)f;
    while ((1u << g.logN) < g.N) ++g.logN;
    g.logM = std::min(g.logN, fft_max_logM<T>());
    g.logF = g.logN - g.logM;
    g.L = (uint32_t)code_length;
    g.D = (uint32_t)((int32_t)(2.0 * dmax / dstep) + 1);
    g.P = 1; g.fs = fs; g.dmax = dmax; g.dstep = dstep;
    const uint32_t F = 1u << g.logF, M = 1u << g.logM, NT = sizeof(T) == 4 ? 512 : 256;
    const auto W = twiddles<T>(g.N);
    std::vector<cx<T>> s(fft_padded_len(M)), C(g.N), X(g.N);
    CodeLoad<T> cl{code, (uint32_t)std::min<uint64_t>(code_len, g.N)};
    for (uint32_t r = 0; r < F; ++r) {
        transform_item<-1, T>(s, cl, g, r, W.data(), NT);
        for (uint32_t m = 0; m < M; ++m) {
            const cx<T> v = s[fft_pad(nat_to_pos(m, g.logM))];
            const T inv_n = (T)1 / (T)g.N;
            C[(m << g.logF) + r] = cx<T>{v.re * inv_n, -v.im * inv_n};
        }
    }
    PeakAcc<double> pair;
    peak_init(pair);
    for (uint32_t d = 0; d < g.D; ++d) {
        WipeLoad<T> wl;
        wl.in = input; wl.in64 = (uint32_t)in64; wl.take = (uint32_t)std::min<uint64_t>(n_input, g.L);
        wl.doppler = -dmax + (double)d * dstep; wl.fs = fs;
        for (uint32_t r = 0; r < F; ++r) {
            transform_item<-1, T>(s, wl, g, r, W.data(), NT);
            for (uint32_t m = 0; m < M; ++m) X[(m << g.logF) + r] = s[fft_pad(nat_to_pos(m, g.logM))];
        }
        ProductLoad<T> pl{X.data(), C.data()};
        for (uint32_t r = 0; r < F; ++r) {
            transform_item<+1, T>(s, pl, g, r, W.data(), NT);
            // per-thread accumulators merged in the kernel's order: lanes by xor butterflies, then warps
            std::vector<PeakAcc<T>> acc(NT);
            for (uint32_t tid = 0; tid < NT; ++tid) {
                peak_init(acc[tid]);
                for (uint32_t p = tid; p < M; p += NT) {
                    const uint32_t n = (pos_to_nat(p, g.logM) << g.logF) + r;
                    if (n < g.L) {
                        const cx<T> v = s[fft_pad(p)];
                        const T mag = v.re * v.re + v.im * v.im;
                        peak_push(acc[tid], mag, n);
                        if (grid) grid[(size_t)d * g.L + n] = (double)mag;
                    }
                }
            }
            for (int off = 16; off > 0; off >>= 1) {
                std::vector<PeakAcc<T>> nx = acc;
                for (uint32_t tid = 0; tid < NT; ++tid) peak_merge(nx[tid], acc[tid ^ (uint32_t)off]);
                acc = nx;
            }
            std::vector<PeakAcc<T>> wacc(32);
            for (uint32_t w = 0; w < 32; ++w) { peak_init(wacc[w]); if (w < NT / 32) wacc[w] = acc[w * 32]; }
            for (int off = 16; off > 0; off >>= 1) {
                std::vector<PeakAcc<T>> nx = wacc;
                for (uint32_t w = 0; w < 32; ++w) peak_merge(nx[w], wacc[w ^ (uint32_t)off]);
                wacc = nx;
            }
            PeakAcc<double> o;
            o.best = (double)wacc[0].best; o.second = (double)wacc[0].second; o.sum = (double)wacc[0].sum;
            o.idx = wacc[0].idx == 0xffffffffu ? 0xffffffffu : d * g.L + wacc[0].idx;
            peak_merge(pair, o);
        }
    }
    out4[0] = pair.best; out4[1] = pair.second; out4[2] = pair.sum; out4[3] = (double)pair.idx;
    return 0;
}

}  // namespace

extern "C" {

int emu_fft(int logN, int logM, int sign, int is_double, const double* in, double* out)
{
    return is_double ? fft_t<double>(logN, logM, sign, in, out) : fft_t<float>(logN, logM, sign, in, out);
}

// out4 = {best, second, sum, linear index d * code_length + lag}; grid (optional) is [bins][code_length]
int emu_pcps(int is_double, uint64_t code_length, double fs, double dmax, double dstep, const void* input, int in64,
             uint64_t n_input, const int8_t* code, uint64_t code_len, double* out4, double* grid)
{
    return is_double ? pcps_t<double>(code_length, fs, dmax, dstep, input, in64, n_input, code, code_len, out4, grid)
                     : pcps_t<float>(code_length, fs, dmax, dstep, input, in64, n_input, code, code_len, out4, grid);
}

}  // extern "C"
