// Microbenchmark deciding the partition design (DESIGN.md "partition"): how fast can B200 do
//   (a) RED.ADD histogram over 2^P random global counters,
//   (b) ATOM.ADD cursor + 8-byte scattered store into 2^P append regions (L2 write combining?)
// for n = 2^29 records.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scatter_mb scatter_mb.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }

__global__ void hist_kernel(uint64_t n, int P, int skew, unsigned int* hist) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        uint64_t h = mix(i);
        if (skew) { uint64_t h2 = mix(i ^ 0x1234567); h = h < h2 ? h : h2; }   // min of two uniforms: canonical-k-mer-like skew
        atomicAdd(&hist[h >> (64 - P)], 1u);
    }
}
__global__ void scatter_kernel(uint64_t n, int P, int skew, unsigned int* cursor, const uint64_t* base, uint64_t* out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        uint64_t h = mix(i);
        if (skew) { uint64_t h2 = mix(i ^ 0x1234567); h = h < h2 ? h : h2; }
        uint32_t b = (uint32_t)(h >> (64 - P));
        uint32_t pos = atomicAdd(&cursor[b], 1u);
        out[base[b] + pos] = h;
    }
}
__global__ void linear_kernel(uint64_t n, uint64_t* out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = mix(i);
}
__global__ void scan_kernel(const unsigned int* hist, uint32_t nb, uint64_t* base) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { uint64_t run = 0; for (uint32_t b = 0; b < nb; b++) { base[b] = run; run += hist[b]; } }
}
int main(int argc, char** argv) {
    uint64_t n = 1ull << 29;
    unsigned int *hist, *cursor; uint64_t *base, *out;
    CK(cudaMalloc(&out, n * 8));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int skew = 0; skew < 2; skew++)
    for (int P = 9; P <= 19; P += 2) {
        uint32_t nb = 1u << P;
        CK(cudaMalloc(&hist, nb * 4)); CK(cudaMalloc(&cursor, nb * 4)); CK(cudaMalloc(&base, nb * 8));
        float th = 0, ts = 0, tl = 0;
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaMemset(hist, 0, nb * 4)); CK(cudaMemset(cursor, 0, nb * 4));
            cudaEventRecord(e0); hist_kernel<<<148 * 8, 256>>>(n, P, skew, hist); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            cudaEventElapsedTime(&th, e0, e1);
            scan_kernel<<<1, 32>>>(hist, nb, base);
            cudaEventRecord(e0); scatter_kernel<<<148 * 8, 256>>>(n, P, skew, cursor, base, out); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            cudaEventElapsedTime(&ts, e0, e1);
            cudaEventRecord(e0); linear_kernel<<<148 * 8, 256>>>(n, out); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            cudaEventElapsedTime(&tl, e0, e1);
        }
        printf("skew=%d P=%2d  hist %.3f ms (%.1f G/s)  scatter %.3f ms (%.1f G rec/s, %.0f GB/s)  linear-store %.3f ms (%.0f GB/s)\n", skew, P, th,
               n / th / 1e6, ts, n / ts / 1e6, n * 8 / ts / 1e6, tl, n * 8 / tl / 1e6);
        cudaFree(hist); cudaFree(cursor); cudaFree(base);
    }
    return 0;
}
