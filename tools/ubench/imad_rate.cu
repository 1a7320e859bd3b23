// Microbenchmark: integer multiply forms on sm_100a (cycles per warp-instruction per SMSP, 4 warps/SMSP, 8 chains/thread)
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
template <int MODE>
__global__ void __launch_bounds__(512) k(unsigned* out, long long* cyc, unsigned m, int iters)
{
    const int t = threadIdx.x;
    unsigned x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = t * 7 + i; y[i] = t * 13 + i; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) { u64 w = (u64)x[i] * 0xD2511F53u; x[i] = (unsigned)(w >> 32); y[i] = (unsigned)w; x[i] += y[i]; }  // WIDE + IADD
                if (MODE == 1) x[i] = __umulhi(x[i], 0xD2511F53u);          // IMAD.HI
                if (MODE == 2) x[i] = x[i] * 0xD2511F53u + 1u;              // IMAD lo
                if (MODE == 3) x[i] = x[i] * m + y[i];                      // IMAD lo reg
                if (MODE == 5) x[i] = (x[i] ^ y[i]) + (x[i] >> 3);          // LOP3 + SHF/IADD (alu)
            }
    }
    const long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + t] = s;
    if (t == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, int instr_per_iter)
{
    unsigned* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 20000;
    k<MODE><<<148, 512>>>(out, cyc, 0xCD9E8D57u, 10);
    k<MODE><<<148, 512>>>(out, cyc, 0xCD9E8D57u, iters);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    printf("%-44s %.3f cycles per group per SMSP-warp slot (4 warps/SMSP)\n", name, c / ((double)iters * instr_per_iter * 4));
}
int main()
{
    run<0>("IMAD.WIDE + IADD (per pair)", 32);
    run<1>("IMAD.HI", 32);
    run<2>("IMAD lo imm", 32);
    run<3>("IMAD lo reg", 32);
    run<5>("LOP3 + SHF + IADD (per triple)", 32);
    return 0;
}
