// Microbenchmark: Philox-like integer work (IMAD.WIDE + LOP3) interleaved with FP32 accumulation as FFMA or FFMA2
// (sm_100a; cycles per group per SMSP with 4 warps/SMSP).  Group = 2 IMAD.WIDE + 2 LOP3 + 8 real-FMA-pairs worth of FP.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, long long* cyc, int iters)
{
    const int t = threadIdx.x;
    float a[16], b[4];
    u64 A[8], B[2];
    unsigned c0 = t, c1 = t * 3, c2 = t * 5, c3 = t * 7, d0 = t + 1, d1 = t * 3 + 1, d2 = t * 5 + 1, d3 = t * 7 + 1;
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = t + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) A[i] = ((u64)__float_as_uint(a[i]) << 32) | i;
    b[0] = 1.0f + t * 1e-7f; b[1] = 0.5f; b[2] = 0.25f + t * 1e-7f; b[3] = 0.75f;
    B[0] = ((u64)__float_as_uint(b[0]) << 32) | __float_as_uint(b[1]); B[1] = ((u64)__float_as_uint(b[2]) << 32) | __float_as_uint(b[3]);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (MODE != 3) {
                {   const u64 w0 = (u64)c0 * 0xD2511F53u, w1 = (u64)c2 * 0xCD9E8D57u;
                    c0 = (unsigned)(w1 >> 32) ^ c1 ^ 0x1234567u; c1 = (unsigned)w1; c2 = (unsigned)(w0 >> 32) ^ c3 ^ 0x89abcdeu; c3 = (unsigned)w0; }
                {   const u64 w0 = (u64)d0 * 0xD2511F53u, w1 = (u64)d2 * 0xCD9E8D57u;
                    d0 = (unsigned)(w1 >> 32) ^ d1 ^ 0x1234567u; d1 = (unsigned)w1; d2 = (unsigned)(w0 >> 32) ^ d3 ^ 0x89abcdeu; d3 = (unsigned)w0; }
            }
            if (MODE == 0 || MODE == 3) {
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fma1(b[(i + r) & 3], b[(i + r + 1) & 3], a[i]);
            }
            if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < 8; ++i) A[i] = fma2(B[(i + r) & 1], B[(i + r + 1) & 1], A[i]);
            }
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += (float)(A[i] >> 32) + (float)(unsigned)A[i];
    s += (float)(c0 ^ c1 ^ c2 ^ c3 ^ d0 ^ d1 ^ d2 ^ d3);
    out[blockIdx.x * blockDim.x + t] = s;
    if (t == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name)
{
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 20000;
    k<MODE><<<148, 512>>>(out, cyc, 10);
    k<MODE><<<148, 512>>>(out, cyc, iters);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    printf("%-60s %.2f cycles per group per warp slot (4 warps/SMSP)\n", name, c / ((double)iters * 8 * 4));
}
int main()
{
    run<0>("2 Philox rounds x2 chains (4 IMAD.WIDE+4 LOP3) + 16 FFMA");
    run<1>("2 Philox rounds x2 chains (4 IMAD.WIDE+4 LOP3) + 8 FFMA2");
    run<2>("4 IMAD.WIDE + 4 LOP3 only");
    run<3>("16 FFMA only");
    return 0;
}
