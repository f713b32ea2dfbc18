// Microbenchmark for the partition scatter (round 2): what bounds part_scatter_kernel at 0.42 of the HBM peak?
//   (1) raw shared-memory atomic throughput (ATOMS.ADD with return / RED-style without / plain LDS+STS),
//   (2) one 8-bit scatter pass over n random u64 records with different ranking schemes and CTA shapes:
//         atom<T,I,B>   shared-memory atomicAdd per record on a per-digit counter (the round-1 kernel)
//         match<T,I,B>  8 unrolled ballots per record-group + warp-private counters (no atomics)
//         many<T,I,B>   __match_any_sync + warp-private counters
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scatter_mb2 scatter_mb2.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }

// ------------------------------------------------------------------------------------------------ (1) atomics
template <int MODE>   // 0: atomicAdd with return, 1: atomicAdd result unused, 2: LDS + STS (racy, for the rate only), 3: atomicOr no return, 4: atomicCAS
__global__ void __launch_bounds__(256) atoms_kernel(int iters, int n_addr_mask, unsigned* sink) {
    __shared__ unsigned s[8192];
    for (int i = threadIdx.x; i < 8192; i += 256) s[i] = 0;
    __syncthreads();
    unsigned x = (unsigned)mix(blockIdx.x * 256 + threadIdx.x), acc = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x = x * 1664525u + 1013904223u;
            const unsigned a = (x >> 10) & n_addr_mask;
            if (MODE == 0) acc += atomicAdd(&s[a], 1u);
            else if (MODE == 1) atomicAdd(&s[a], 1u);
            else if (MODE == 2) { unsigned v = s[a]; s[a] = v + 1; acc += v; }
            else if (MODE == 3) atomicOr(&s[a], 1u << (x & 31));
            else acc += atomicCAS(&s[a], 0u, x);
        }
    }
    __syncthreads();
    if (acc == 0x12345678u || s[threadIdx.x] == 0xFFFFFFFFu) sink[0] = acc;
}

// ------------------------------------------------------------------------------------------------ (2) scatter
constexpr int BITS = 8;
constexpr int D = 1 << BITS;
__device__ __forceinline__ uint32_t digit_of(uint64_t r) { return (uint32_t)(r >> 52) & (D - 1); }

__global__ void gen_kernel(uint64_t n, uint64_t* out, int skew) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        uint64_t h = mix(i + 1) >> 4;                       // 60-bit "key"
        if (skew) { const uint64_t h2 = mix(i ^ 0x1234567) >> 4; h = h < h2 ? h : h2; }
        out[i] = h;
    }
}
__global__ void hist_kernel(const uint64_t* src, uint64_t n, unsigned long long* hist) {
    __shared__ unsigned s[D];
    for (int i = threadIdx.x; i < D; i += blockDim.x) s[i] = 0;
    __syncthreads();
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) atomicAdd(&s[digit_of(src[i])], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += blockDim.x) if (s[i]) atomicAdd(&hist[i], (unsigned long long)s[i]);
}
__global__ void scan_kernel(const unsigned long long* hist, unsigned long long* cursor, unsigned long long* off) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { unsigned long long run = 0; for (int b = 0; b < D; b++) { off[b] = run; cursor[b] = run; run += hist[b]; } off[D] = run; }
}
__global__ void check_kernel(const uint64_t* dst, uint64_t n, const unsigned long long* off, unsigned long long* bad, unsigned long long* xsum) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long x = 0, b = 0;
    for (; i < n; i += stride) {
        const uint64_t r = dst[i]; const uint32_t d = digit_of(r);
        if (i < off[d] || i >= off[d + 1]) b++;
        x += mix(r);
    }
    if (b) atomicAdd(bad, b);
    atomicAdd(xsum, x);
}

// common tail: scan the per-digit tile counts, reserve, stage, write
template <int T, int I, int WC, class RankFn>
__device__ __forceinline__ void scatter_body(const uint64_t* __restrict__ src, uint64_t n, unsigned long long* __restrict__ cursor,
                                             uint64_t* __restrict__ dst, RankFn rank) {
    constexpr int TILE = T * I;
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t* s_stage = reinterpret_cast<uint64_t*>(smem);                                    // [TILE]
    unsigned long long* s_gbase = reinterpret_cast<unsigned long long*>(s_stage + TILE);      // [D]
    uint32_t* s_doff = reinterpret_cast<uint32_t*>(s_gbase + D);                              // [D + 1]
    uint32_t* s_cnt = s_doff + D + 1;                                                         // [D] (+ per-warp counters after it)
    uint16_t* s_sdig = reinterpret_cast<uint16_t*>(s_cnt + D + WC * D);                 // [TILE]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t beg = (uint64_t)blockIdx.x * TILE, end = min(n, beg + TILE);
    uint64_t rr[I]; uint32_t dp[I];
#pragma unroll
    for (int j = 0; j < I; j++) { const uint64_t i = beg + (uint64_t)j * T + threadIdx.x; rr[j] = i < end ? src[i] : ~0ULL; }
    rank(rr, dp, s_cnt, beg, end);            // dp[j] = digit << 16 | rank among the tile's records of that digit; leaves totals in s_cnt[0..D)
    __syncthreads();
    if (warp == 0) {
        constexpr int per = D / 32; const int b0 = lane * per;
        uint32_t v[per], sum = 0;
#pragma unroll
        for (int q = 0; q < per; q++) { v[q] = s_cnt[b0 + q]; sum += v[q]; }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
        uint32_t run = inc - sum;
#pragma unroll
        for (int q = 0; q < per; q++) { s_doff[b0 + q] = run; run += v[q]; }
        if (lane == 31) s_doff[D] = inc;
    }
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += T) {
        const uint32_t tot = s_doff[d + 1] - s_doff[d];
        if (tot) s_gbase[d] = atomicAdd(&cursor[d], (unsigned long long)tot) - s_doff[d];
    }
#pragma unroll
    for (int j = 0; j < I; j++) {
        if (dp[j] == 0xFFFFFFFFu) continue;
        const uint32_t d = dp[j] >> 16, slot = s_doff[d] + (dp[j] & 0xFFFFu);
        s_stage[slot] = rr[j]; s_sdig[slot] = (uint16_t)d;
    }
    __syncthreads();
    const uint32_t nv = s_doff[D];
    for (uint32_t slot = threadIdx.x; slot < nv; slot += T) dst[s_gbase[s_sdig[slot]] + slot] = s_stage[slot];
}

template <int T, int I> struct RankAtom {
    __device__ __forceinline__ void operator()(const uint64_t (&rr)[I], uint32_t (&dp)[I], uint32_t* s_cnt, uint64_t beg, uint64_t end) const {
        for (int i = threadIdx.x; i < D; i += T) s_cnt[i] = 0;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < I; j++) {
            const uint64_t i = beg + (uint64_t)j * T + threadIdx.x;
            if (i < end) { const uint32_t d = digit_of(rr[j]); dp[j] = (d << 16) | atomicAdd(&s_cnt[d], 1u); } else dp[j] = 0xFFFFFFFFu;
        }
    }
};
// ballot matching, warp-private counters: wc[warp][D] u32 after the D totals
template <int T, int I, bool MATCH_ANY> struct RankMatch {
    __device__ __forceinline__ void operator()(const uint64_t (&rr)[I], uint32_t (&dp)[I], uint32_t* s_cnt, uint64_t beg, uint64_t end) const {
        constexpr int W = T / 32;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        uint32_t* wc = s_cnt + D;                                    // [W][D]
        for (int i = threadIdx.x; i < W * D; i += T) wc[i] = 0;
        __syncthreads();
        uint32_t* my = wc + warp * D;
        const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
        for (int j = 0; j < I; j++) {
            const uint64_t i = beg + (uint64_t)j * T + threadIdx.x;
            const bool valid = i < end;
            const uint32_t d = digit_of(rr[j]);
            uint32_t peers;
            if (MATCH_ANY) peers = __match_any_sync(0xFFFFFFFFu, valid ? d : 0xFFFFu);
            else {
                peers = __ballot_sync(0xFFFFFFFFu, valid);
#pragma unroll
                for (int b = 0; b < BITS; b++) {
                    const uint32_t m = __ballot_sync(0xFFFFFFFFu, (d >> b) & 1u);
                    peers &= ((d >> b) & 1u) ? m : ~m;
                }
            }
            const uint32_t old = my[d];                              // every peer reads the same counter (broadcast)
            __syncwarp();
            if (valid && (peers & lt) == 0) my[d] = old + __popc(peers);
            __syncwarp();
            dp[j] = valid ? ((d << 16) | (old + __popc(peers & lt))) : 0xFFFFFFFFu;
        }
        __syncthreads();
        // per digit: exclusive prefix over warps (in place), total to s_cnt[d]
        for (int d = threadIdx.x; d < D; d += T) {
            uint32_t tot = 0;
#pragma unroll
            for (int w = 0; w < W; w++) { const uint32_t t = wc[w * D + d]; wc[w * D + d] = tot; tot += t; }
            s_cnt[d] = tot;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < I; j++) if (dp[j] != 0xFFFFFFFFu) dp[j] += my[dp[j] >> 16];
    }
};

template <int T, int I, int B> __global__ void __launch_bounds__(T, B)
scatter_atom(const uint64_t* src, uint64_t n, unsigned long long* cursor, uint64_t* dst) { scatter_body<T, I, 0>(src, n, cursor, dst, RankAtom<T, I>()); }
template <int T, int I, int B> __global__ void __launch_bounds__(T, B)
scatter_match(const uint64_t* src, uint64_t n, unsigned long long* cursor, uint64_t* dst) { scatter_body<T, I, T / 32>(src, n, cursor, dst, RankMatch<T, I, false>()); }
template <int T, int I, int B> __global__ void __launch_bounds__(T, B)
scatter_many(const uint64_t* src, uint64_t n, unsigned long long* cursor, uint64_t* dst) { scatter_body<T, I, T / 32>(src, n, cursor, dst, RankMatch<T, I, true>()); }

// plain copy for the bandwidth reference
__global__ void copy_kernel(const uint64_t* src, uint64_t n, uint64_t* dst) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = src[i];
}

static uint64_t* g_src; static uint64_t* g_dst; static unsigned long long *g_hist, *g_cursor, *g_off, *g_chk;
static uint64_t g_n; static unsigned long long g_xsum_ref;
static cudaEvent_t e0, e1;

template <class K> static void run_scatter(const char* name, K kern, int T, int I, bool wc = false) {
    const int TILE = T * I;
    const size_t smem = (size_t)TILE * 8 + D * 8 + (D + 1) * 4 + (D + (wc ? (T / 32) * D : 0)) * 4 + TILE * 2 + 64;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
    float best = 1e9;
    for (int rep = 0; rep < 4; rep++) {
        scan_kernel<<<1, 32>>>(g_hist, g_cursor, g_off);
        CK(cudaEventRecord(e0));
        kern<<<(unsigned)((g_n + TILE - 1) / TILE), T, smem>>>(g_src, g_n, g_cursor, g_dst);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    CK(cudaMemset(g_chk, 0, 16));
    check_kernel<<<148 * 8, 256>>>(g_dst, g_n, g_off, g_chk, g_chk + 1);
    unsigned long long h[2]; CK(cudaMemcpy(h, g_chk, 16, cudaMemcpyDeviceToHost));
    printf("%-22s T=%4d I=%2d regs=%3d smem=%6zu occ=%d  %.3f ms  %.0f GB/s (r+w)  %s\n", name, T, I, fa.numRegs, smem, occ, best,
           g_n * 16 / best / 1e6, (h[0] == 0 && h[1] == g_xsum_ref) ? "ok" : "WRONG");
    fflush(stdout);
}

int main(int argc, char** argv) {
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    // ---- (1)
    unsigned* sink; CK(cudaMalloc(&sink, 4));
    const char* mode_name[] = {"atomicAdd+ret", "atomicAdd noret", "LDS+STS", "atomicOr noret", "atomicCAS"};
    if (argc <= 1) for (int mask : {255, 4095, 8191}) {
        for (int mode = 0; mode < 5; mode++) {
            const int iters = 2000, grid = 148 * 8;
            float ms = 0;
            for (int rep = 0; rep < 2; rep++) {
                CK(cudaEventRecord(e0));
                if (mode == 0) atoms_kernel<0><<<grid, 256>>>(iters, mask, sink);
                else if (mode == 1) atoms_kernel<1><<<grid, 256>>>(iters, mask, sink);
                else if (mode == 2) atoms_kernel<2><<<grid, 256>>>(iters, mask, sink);
                else if (mode == 3) atoms_kernel<3><<<grid, 256>>>(iters, mask, sink);
                else atoms_kernel<4><<<grid, 256>>>(iters, mask, sink);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                cudaEventElapsedTime(&ms, e0, e1);
            }
            const double ops = (double)grid * 256 * iters * 8;
            // SM-cycles per warp instruction at 1.965 GHz: time * f * 148 / (ops / 32)
            printf("atoms %-16s addr=%5d: %.3f ms, %.1f Gops/s, %.2f SM-cycles per warp-instruction\n", mode_name[mode], mask + 1, ms, ops / ms / 1e6,
                   ms * 1e-3 * 1.965e9 * 148 / (ops / 32));
        }
    }
    fflush(stdout);
    // ---- (2)
    g_n = argc > 1 ? strtoull(argv[1], nullptr, 10) : (1ull << 28);
    CK(cudaMalloc(&g_src, g_n * 8)); CK(cudaMalloc(&g_dst, g_n * 8));
    CK(cudaMalloc(&g_hist, D * 8)); CK(cudaMalloc(&g_cursor, D * 8)); CK(cudaMalloc(&g_off, (D + 1) * 8)); CK(cudaMalloc(&g_chk, 16));
    gen_kernel<<<148 * 8, 256>>>(g_n, g_src, argc > 2 ? atoi(argv[2]) : 0);
    CK(cudaMemset(g_hist, 0, D * 8));
    hist_kernel<<<148 * 8, 256>>>(g_src, g_n, g_hist);
    scan_kernel<<<1, 32>>>(g_hist, g_cursor, g_off);
    CK(cudaMemset(g_chk, 0, 16));
    check_kernel<<<148 * 8, 256>>>(g_src, g_n, g_off, g_chk, g_chk + 1);
    unsigned long long h[2]; CK(cudaMemcpy(h, g_chk, 16, cudaMemcpyDeviceToHost)); g_xsum_ref = h[1];
    {
        float best = 1e9;
        for (int rep = 0; rep < 4; rep++) {
            CK(cudaEventRecord(e0)); copy_kernel<<<148 * 16, 512>>>(g_src, g_n, g_dst); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("copy                   %.3f ms  %.0f GB/s (r+w), n = %llu records\n", best, g_n * 16 / best / 1e6, (unsigned long long)g_n);
    }
    run_scatter("atom 512x8 lb1", scatter_atom<512, 8, 1>, 512, 8);
    run_scatter("atom 512x8 lb2", scatter_atom<512, 8, 2>, 512, 8);
    run_scatter("atom 512x8 lb3", scatter_atom<512, 8, 3>, 512, 8);
    run_scatter("atom 512x8 lb4", scatter_atom<512, 8, 4>, 512, 8);
    run_scatter("atom 256x16 lb4", scatter_atom<256, 16, 4>, 256, 16);
    run_scatter("atom 256x16 lb3", scatter_atom<256, 16, 3>, 256, 16);
    run_scatter("atom 256x16 lb5", scatter_atom<256, 16, 5>, 256, 16);
    run_scatter("atom 128x32 lb8", scatter_atom<128, 32, 8>, 128, 32);
    run_scatter("atom 128x16 lb8", scatter_atom<128, 16, 8>, 128, 16);
    run_scatter("atom 256x24 lb2", scatter_atom<256, 24, 2>, 256, 24);
    run_scatter("atom 512x16 lb1", scatter_atom<512, 16, 1>, 512, 16);
    run_scatter("atom 256x8 lb8", scatter_atom<256, 8, 8>, 256, 8);
    run_scatter("atom 1024x4 lb2", scatter_atom<1024, 4, 2>, 1024, 4);
    run_scatter("atom 1024x8 lb1", scatter_atom<1024, 8, 1>, 1024, 8);
    run_scatter("match 512x8 lb2", scatter_match<512, 8, 2>, 512, 8, true);
    run_scatter("match 512x8 lb3", scatter_match<512, 8, 3>, 512, 8, true);
    run_scatter("match 256x16 lb4", scatter_match<256, 16, 4>, 256, 16, true);
    run_scatter("match 256x16 lb6", scatter_match<256, 16, 6>, 256, 16, true);
    run_scatter("match 128x32 lb8", scatter_match<128, 32, 8>, 128, 32, true);
    run_scatter("match 1024x8 lb1", scatter_match<1024, 8, 1>, 1024, 8, true);
    run_scatter("many 512x8 lb2", scatter_many<512, 8, 2>, 512, 8, true);
    run_scatter("many 256x16 lb4", scatter_many<256, 16, 4>, 256, 16, true);
    return 0;
}
