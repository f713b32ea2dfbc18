// What a CLI process pays before its first kernel: context creation, properties query, big allocations.
#include <cstdio>
#include <chrono>
#include <cuda_runtime.h>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
__global__ void k(int* p) { if (p) p[0] = 1; }
int main() {
    double t0 = now(), t;
    int n = 0; cudaGetDeviceCount(&n); t = now(); printf("cudaGetDeviceCount        %.3f s\n", t - t0); t0 = t;
    cudaSetDevice(0); cudaFree(0); t = now(); printf("cudaSetDevice + cudaFree(0) %.3f s\n", t - t0); t0 = t;
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); t = now(); printf("cudaGetDeviceProperties   %.3f s\n", t - t0); t0 = t;
    int v; cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, 0); t = now(); printf("cudaDeviceGetAttribute    %.3f s\n", t - t0); t0 = t;
    cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking); t = now(); printf("cudaStreamCreate          %.3f s\n", t - t0); t0 = t;
    k<<<1, 32, 0, s>>>(nullptr); cudaStreamSynchronize(s); t = now(); printf("first kernel launch       %.3f s\n", t - t0); t0 = t;
    void* a; cudaMalloc(&a, 13ull << 30); t = now(); printf("cudaMalloc 13 GB          %.3f s\n", t - t0); t0 = t;
    void* b; cudaMallocAsync(&b, 1400ull << 20, s); cudaStreamSynchronize(s); t = now(); printf("cudaMallocAsync 1.4 GB    %.3f s\n", t - t0); t0 = t;
    void* h; cudaMallocHost(&h, 256ull << 20); t = now(); printf("cudaMallocHost 256 MB     %.3f s\n", t - t0); t0 = t;
    cudaFree(a); t = now(); printf("cudaFree 13 GB            %.3f s\n", t - t0); t0 = t;
    return 0;
}
