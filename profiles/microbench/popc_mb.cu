// Microbenchmark behind K5's roofline ("report K5 as popc/s against a measured popc microbenchmark peak"):
//   popc32 / lop3 issue rates per SM, and the rate of the b1 AND.POPC mma.sync (emulated on sm_100a with IMMA.16832.U8).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o popc_mb popc_mb.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE> __global__ void __launch_bounds__(256) rate_kernel(int iters, uint32_t seed, uint32_t* out) {
    uint32_t x[8], acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B9u; acc[i] = 0; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) { acc[i] += __popc(x[i]); x[i] += acc[i]; }                      // 1 popc + 2 iadd (dependent chain per i, 8 chains)
            else if (MODE == 1) { acc[i] += __popc(x[i] ^ acc[(i + 1) & 7]); }              // 1 popc + lop + iadd
            else if (MODE == 2) { x[i] = (x[i] ^ acc[i]) | (x[(i + 1) & 7] & seed); acc[i] ^= x[i]; }   // lop3 only
            else { acc[i] += __popc(x[i]) ; acc[i] += __popc(x[i] ^ seed); acc[i] += __popc(x[i] & (seed >> 3)); x[i] ^= acc[i]; }   // 3 popc per step
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r ^= acc[i] ^ x[i];
    if (r == 0x12345u) out[0] = r;
}

__global__ void __launch_bounds__(256) bmma_kernel(int iters, const uint32_t* in, int* out) {
    uint32_t a0 = in[threadIdx.x], a1 = in[threadIdx.x + 32], a2 = in[threadIdx.x + 64], a3 = in[threadIdx.x + 96];
    uint32_t b0 = in[threadIdx.x + 128], b1 = in[threadIdx.x + 160];
    int c[4][4] = {};
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int q = 0; q < 4; q++)
            asm volatile("mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.and.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+r"(c[q][0]), "+r"(c[q][1]), "+r"(c[q][2]), "+r"(c[q][3]) : "r"(a0 + q), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1 + it));
    }
    int r = 0;
    for (int q = 0; q < 4; q++) for (int i = 0; i < 4; i++) r += c[q][i];
    if (r == 0x1234567) out[0] = r;
}
__global__ void __launch_bounds__(256) imma_kernel(int iters, const uint32_t* in, int* out) {
    uint32_t a0 = in[threadIdx.x], a1 = in[threadIdx.x + 32], a2 = in[threadIdx.x + 64], a3 = in[threadIdx.x + 96];
    uint32_t b0 = in[threadIdx.x + 128], b1 = in[threadIdx.x + 160];
    int c[4][4] = {};
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int q = 0; q < 4; q++)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+r"(c[q][0]), "+r"(c[q][1]), "+r"(c[q][2]), "+r"(c[q][3]) : "r"(a0 + q), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1 + it));
    }
    int r = 0;
    for (int q = 0; q < 4; q++) for (int i = 0; i < 4; i++) r += c[q][i];
    if (r == 0x1234567) out[0] = r;
}

int main() {
    uint32_t* d; CK(cudaMalloc(&d, 4096)); CK(cudaMemset(d, 0x5A, 4096));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8, iters = 20000;
    const char* names[] = {"popc+2iadd", "popc+lop+iadd", "2 lop3", "3 popc + 3 iadd + 3 lop"};
    const double ops_per_step[] = {1, 1, 2, 3};
    for (int mode = 0; mode < 4; mode++) {
        float ms = 0;
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) rate_kernel<0><<<grid, 256>>>(iters, 12345u, d);
            else if (mode == 1) rate_kernel<1><<<grid, 256>>>(iters, 12345u, d);
            else if (mode == 2) rate_kernel<2><<<grid, 256>>>(iters, 12345u, d);
            else rate_kernel<3><<<grid, 256>>>(iters, 12345u, d);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        }
        const double ops = (double)grid * 256 * iters * 8 * ops_per_step[mode];
        printf("%-28s %.3f ms  %.2f T%s/s  = %.1f lanes/clk/SM at 1.965 GHz\n", names[mode], ms, ops / ms / 1e9, mode == 2 ? "lop3" : "popc",
               ops / (ms * 1e-3) / 148 / 1.965e9);
    }
    {
        float ms = 0; const int it2 = 4000;
        for (int rep = 0; rep < 2; rep++) { cudaEventRecord(e0); bmma_kernel<<<grid, 256>>>(it2, d, (int*)d + 512); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); }
        const double macs = (double)grid * 8 * it2 * 4 * 16 * 8 * 256;
        printf("mma.sync b1 and.popc m16n8k256: %.3f ms  %.1f Tbit-MAC/s = %.0f bit-MAC/clk/SM (a popc32 does 32: x%.1f the popc rate of 16 lanes/clk)\n", ms, macs / ms / 1e9,
               macs / (ms * 1e-3) / 148 / 1.965e9, macs / (ms * 1e-3) / 148 / 1.965e9 / 512.0);
        for (int rep = 0; rep < 2; rep++) { cudaEventRecord(e0); imma_kernel<<<grid, 256>>>(it2, d, (int*)d + 512); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); }
        const double macs8 = (double)grid * 8 * it2 * 4 * 16 * 8 * 32;
        printf("mma.sync u8 m16n8k32:           %.3f ms  %.1f TMAC/s = %.0f MAC/clk/SM\n", ms, macs8 / ms / 1e9, macs8 / (ms * 1e-3) / 148 / 1.965e9);
    }
    return 0;
}
