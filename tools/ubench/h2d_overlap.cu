// Does a kernel on one stream start while pinned H2D copies are queued on another?  (sm_100a box check)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void spin(float* p, int n) { float a = p[threadIdx.x]; for (int i = 0; i < n; ++i) a = a * 1.0001f + 0.5f; p[threadIdx.x] = a; }
int main()
{
    const size_t bytes = 800u << 20, piece = 32u << 20;
    void *h, *d; float* w;
    cudaHostAlloc(&h, bytes, cudaHostAllocDefault); cudaMalloc(&d, bytes); cudaMalloc(&w, 4096);
    cudaStream_t a, b; cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking);
    cudaEvent_t e0, e1, e2, c1; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventCreate(&c1);
    for (int mode = 0; mode < 3; ++mode) {
        cudaDeviceSynchronize();
        cudaEventRecord(e0, b);
        if (mode >= 1) for (size_t off = 0; off < bytes; off += piece) cudaMemcpyAsync((char*)d + off, (char*)h + off, piece < bytes - off ? piece : bytes - off, cudaMemcpyHostToDevice, a);
        cudaEventRecord(c1, a);
        if (mode == 2) { cudaEvent_t t; cudaEventCreate(&t); cudaEventRecord(t, b); }
        cudaEventRecord(e1, b);
        spin<<<148, 256, 0, b>>>(w, 100000);
        cudaEventRecord(e2, b);
        cudaDeviceSynchronize();
        float lead, run, cp;
        cudaEventElapsedTime(&lead, e0, e1); cudaEventElapsedTime(&run, e1, e2); cudaEventElapsedTime(&cp, e0, c1);
        printf("mode %d: kernel stream start delay %.3f ms, kernel %.3f ms, copies done after %.3f ms\n", mode, lead, run, cp);
    }
    return 0;
}
