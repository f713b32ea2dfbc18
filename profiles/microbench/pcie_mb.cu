// PCIe microbenchmark for the e2e arm: H2D alone, D2H alone, both at once (pinned memory, two streams).
#include <cstdio>
#include <cuda_runtime.h>
#include <chrono>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main() {
    const size_t n = 512ull << 20;
    char *h1, *h2, *d1, *d2;
    CK(cudaMallocHost(&h1, n)); CK(cudaMallocHost(&h2, n)); CK(cudaMalloc(&d1, n)); CK(cudaMalloc(&d2, n));
    cudaStream_t s1, s2; cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    auto run = [&](bool up, bool down, size_t chunk) {
        double best = 1e9;
        for (int rep = 0; rep < 3; rep++) {
            cudaDeviceSynchronize();
            auto t0 = std::chrono::steady_clock::now();
            for (size_t o = 0; o < n; o += chunk) {
                if (up) cudaMemcpyAsync(d1 + o, h1 + o, chunk, cudaMemcpyHostToDevice, s1);
                if (down) cudaMemcpyAsync(h2 + o, d2 + o, chunk, cudaMemcpyDeviceToHost, s2);
            }
            cudaDeviceSynchronize();
            double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (ms < best) best = ms;
        }
        return best;
    };
    for (size_t chunk : {n, (size_t)8 << 20, (size_t)1 << 20}) {
        double a = run(true, false, chunk), b = run(false, true, chunk), c = run(true, true, chunk);
        printf("chunk %4zu MB: H2D %.1f GB/s, D2H %.1f GB/s, both at once %.1f + %.1f GB/s (%.2f ms for 2 x 512 MB)\n", chunk >> 20, n / a / 1e6, n / b / 1e6,
               n / c / 1e6, n / c / 1e6, c);
    }
    return 0;
}
