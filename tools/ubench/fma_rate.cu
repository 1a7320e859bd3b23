// Microbenchmark: issue rate of FFMA / FFMA2 / IMAD.WIDE forms on sm_100a (cycles per warp-instruction per SM sub-partition).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_rate fma_rate.cu && ./fma_rate
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, long long* cyc, float p0, float p1, u64 q0, int iters)
{
    const int t = threadIdx.x;
    float a[8], b = 1.0f + t * 1e-7f, c = 0.5f + t * 1e-7f;
    u64 A[8], B = ((u64)__float_as_uint(b) << 32) | __float_as_uint(c), C = B ^ 0x100000001ull;
    unsigned x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = t + i; A[i] = ((u64)__float_as_uint(a[i]) << 32) | i; x[i] = t * 7 + i; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) a[i] = fma1(a[i], b, c);                 // FFMA 3 distinct regs (acc as multiplicand)
                if (MODE == 1) a[i] = fma1(b, c, a[i]);                 // FFMA acc as addend, b c shared
                if (MODE == 2) a[i] = fma1(a[i], p0, p1);               // FFMA with param (uniform/const) operands
                if (MODE == 3) A[i] = fma2(B, C, A[i]);                 // FFMA2 3 regs
                if (MODE == 4) A[i] = fma2(A[i], q0, C);                // FFMA2 one param operand
                if (MODE == 5) { u64 w = (u64)x[i] * 0xD2511F53u; x[i] = (unsigned)(w >> 32) ^ (unsigned)w; }   // IMAD.WIDE + LOP3
                if (MODE == 6) a[i] = fma1(a[(i + 1) & 7], b, a[i]);    // FFMA 3 regs all varying
                if (MODE == 7) A[i] = fma2(A[(i + 1) & 7], B, A[i]);    // FFMA2 all varying
            }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + (float)(A[i] >> 32) + (float)x[i];
    out[blockIdx.x * blockDim.x + t] = s;
    if (t == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_iter)
{
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 20000;
    k<MODE><<<148, 512>>>(out, cyc, 1.0001f, 0.5f, 0x3f8000013f800001ull, 10);
    k<MODE><<<148, 512>>>(out, cyc, 1.0001f, 0.5f, 0x3f8000013f800001ull, iters);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    // 16 warps per SM = 4 per SMSP
    printf("%-44s %.3f cycles per warp-instruction per SMSP (4 warps/SMSP)\n", name, c / ((double)iters * instr_per_iter * 4));
}
int main()
{
    run<0>("FFMA  acc*b+c", 32);
    run<1>("FFMA  b*c+acc (shared b,c)", 32);
    run<2>("FFMA  acc*param+param", 32);
    run<3>("FFMA2 B*C+acc", 32);
    run<4>("FFMA2 acc*param+C", 32);
    run<5>("IMAD.WIDE + LOP3 pair", 64);
    run<6>("FFMA  a[i+1]*b+a[i]", 32);
    run<7>("FFMA2 A[i+1]*B+A[i]", 32);
    return 0;
}
