// sort128.cu -- ascending sort of 16-byte records (hi word = key, lo word = payload / tie-break), the sort behind the exact
// emulation of the reference's count filter (k1q_reads.cu: observations by (mixed hash, position), distinct hashes by
// (bloom word, first position); src/ska_dict/bloom_filter.rs:116-148).  Replaces the library radix sort of round 1.
//
// MSD: the records are grouped by their top 22 key bits with the two scatter passes of partition.cu (group_wide_records);
// keys here are hashes / bloom words, i.e. uniform, so a bucket holds n / 2^22 records on average.  One WARP then sorts one
// bucket in shared memory (bitonic, <= 256 records); the rare larger bucket (one k-mer repeated thousands of times) is
// sorted by a whole CTA in shared memory (<= 8192 records) or, beyond that, in place in global memory.
// Roofline: HBM; 16 B per record moved by each of two scatters and once by the bucket sort.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace ska {

typedef unsigned __int128 u128s;
constexpr int SW_PER_WARP = 256;          // records a warp sorts
constexpr int SM_MAX = 2048;              // records a 256-thread CTA sorts in shared memory (32 KB: several CTAs per SM)
constexpr int SB_MAX = 8192;              // records a 1024-thread CTA sorts in shared memory

__device__ __forceinline__ bool rec_gt(const ulonglong2& a, const ulonglong2& b) { return a.y > b.y || (a.y == b.y && a.x > b.x); }

// one warp per bucket; buckets larger than SW_PER_WARP go to the big list
__global__ void __launch_bounds__(256)
bucket_sort_small_kernel(uint64_t* __restrict__ rec, const uint64_t* __restrict__ off, uint32_t n_buckets, uint32_t* __restrict__ big_list,
                         uint32_t* __restrict__ n_big) {
    __shared__ ulonglong2 s_buf[8][SW_PER_WARP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t b = blockIdx.x * 8 + warp;
    if (b >= n_buckets) return;
    const uint64_t beg = off[b], end = off[b + 1];
    const uint32_t n = (uint32_t)min(end - beg, (uint64_t)0xFFFFFFFFu);
    if (n < 2) return;
    if (end - beg > SW_PER_WARP) { if (lane == 0) big_list[atomicAdd(n_big, 1u)] = b; return; }
    ulonglong2* r2 = reinterpret_cast<ulonglong2*>(rec) + beg;
    ulonglong2* s = s_buf[warp];
    uint32_t np2 = 2; while (np2 < n) np2 <<= 1;
    for (uint32_t i = lane; i < np2; i += 32) s[i] = i < n ? r2[i] : make_ulonglong2(~0ULL, ~0ULL);
    __syncwarp();
    for (uint32_t size = 2; size <= np2; size <<= 1)
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = lane; t < np2 / 2; t += 32) {
                const uint32_t a = 2 * t - (t & (stride - 1)), c = a + stride;
                const bool up = (a & size) == 0;
                const ulonglong2 x = s[a], y = s[c];
                if (rec_gt(x, y) == up) { s[a] = y; s[c] = x; }
            }
            __syncwarp();
        }
    for (uint32_t i = lane; i < n; i += 32) r2[i] = s[i];
}

// one CTA per large bucket: shared memory up to SB_MAX records; anything larger goes to the huge list (sorted by recursing
// on the next key bits from the host: a bucket that large is one hash seen thousands of times, told apart only by position)
template <int T, int MAXN>
__global__ void __launch_bounds__(T)
bucket_sort_big_kernel(uint64_t* __restrict__ rec, const uint64_t* __restrict__ off, const uint32_t* __restrict__ big_list,
                       uint32_t* __restrict__ huge_list, uint32_t* __restrict__ n_huge) {
    extern __shared__ __align__(16) uint8_t smem_sb[];
    ulonglong2* s = reinterpret_cast<ulonglong2*>(smem_sb);
    const uint32_t b = big_list[blockIdx.x];
    const uint64_t beg = off[b], n = off[b + 1] - beg;
    if (n > MAXN) { if (threadIdx.x == 0) huge_list[atomicAdd(n_huge, 1u)] = b; return; }
    ulonglong2* r2 = reinterpret_cast<ulonglong2*>(rec) + beg;
    uint32_t np2 = 2; while (np2 < n) np2 <<= 1;
    for (uint32_t i = threadIdx.x; i < np2; i += T) s[i] = i < n ? r2[i] : make_ulonglong2(~0ULL, ~0ULL);
    __syncthreads();
    for (uint32_t size = 2; size <= np2; size <<= 1)
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = threadIdx.x; t < np2 / 2; t += T) {
                const uint32_t a = 2 * t - (t & (stride - 1)), c = a + stride;
                const bool up = (a & size) == 0;
                const ulonglong2 x = s[a], y = s[c];
                if (rec_gt(x, y) == up) { s[a] = y; s[c] = x; }
            }
            __syncthreads();
        }
    for (uint32_t i = threadIdx.x; i < n; i += T) r2[i] = s[i];
}

// buckets the warps left (big_list, count in cnts[0]): 256-thread CTAs up to SM_MAX records, 1024-thread CTAs up to SB_MAX;
// what is larger still is returned as (bucket index list on the host, bucket offsets on the host)
static void sort_big_buckets(uint64_t* data, const uint64_t* off, uint32_t nb, uint32_t* big_list, uint32_t* mid_list, uint32_t* huge_list,
                             uint32_t* cnts /* [3], cnts[0] filled */, uint32_t n_big, std::vector<uint32_t>& h_huge, std::vector<uint64_t>& h_off) {
    Context& c = ctx();
    h_huge.clear();
    if (!n_big) return;
    SKA_LAUNCH((bucket_sort_big_kernel<256, SM_MAX>), n_big, 256, (size_t)SM_MAX * 16, data, off, big_list, mid_list, cnts + 1);
    uint32_t h_cnt[3] = {0, 0, 0};
    SKA_CUDA(cudaMemcpyAsync(h_cnt, cnts, 12, cudaMemcpyDeviceToHost, c.stream));
    SKA_CUDA(cudaStreamSynchronize(c.stream));
    if (!h_cnt[1]) return;
    const size_t smem = (size_t)SB_MAX * 16;
    SKA_CUDA(cudaFuncSetAttribute(bucket_sort_big_kernel<1024, SB_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SKA_LAUNCH((bucket_sort_big_kernel<1024, SB_MAX>), h_cnt[1], 1024, smem, data, off, mid_list, huge_list, cnts + 2);
    SKA_CUDA(cudaMemcpyAsync(h_cnt, cnts, 12, cudaMemcpyDeviceToHost, c.stream));
    SKA_CUDA(cudaStreamSynchronize(c.stream));
    if (!h_cnt[2]) return;
    h_huge.resize(h_cnt[2]);
    h_off.resize(nb + 1);
    SKA_CUDA(cudaMemcpyAsync(h_huge.data(), huge_list, (size_t)h_cnt[2] * 4, cudaMemcpyDeviceToHost, c.stream));
    SKA_CUDA(cudaMemcpyAsync(h_off.data(), off, (size_t)(nb + 1) * 8, cudaMemcpyDeviceToHost, c.stream));
    SKA_CUDA(cudaStreamSynchronize(c.stream));
}

// OR over i of (rec[i] ^ rec[0]): the bits in which the records of an oversized bucket differ at all
__global__ void diff_bits_kernel(const ulonglong2* __restrict__ rec, uint64_t n, unsigned long long* __restrict__ out) {
    const ulonglong2 r0 = rec[0];
    unsigned long long dx = 0, dy = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const ulonglong2 r = rec[i];
        dx |= r.x ^ r0.x; dy |= r.y ^ r0.y;
    }
    for (int o = 16; o > 0; o >>= 1) { dx |= __shfl_xor_sync(0xFFFFFFFFu, dx, o); dy |= __shfl_xor_sync(0xFFFFFFFFu, dy, o); }
    if ((threadIdx.x & 31) == 0) { if (dx) atomicOr(out, dx); if (dy) atomicOr(out + 1, dy); }
}

// sorts data[0, n) in place; bits at and above `top_bit` are equal in all records.  tmp: n records of scratch.
static void sort_range_wide(uint64_t* data, uint64_t n, int top_bit, uint64_t* tmp, int depth) {
    Context& c = ctx();
    if (n < 2 || top_bit <= 0) return;
    if (depth > 8) throw Error("internal: wide sort recursion too deep");
    {   // an oversized bucket is one key seen thousands of times: skip the bits its records share
        unsigned long long* d_diff = arena().take_n<unsigned long long>(2);
        SKA_CUDA(cudaMemsetAsync(d_diff, 0, 16, c.stream));
        SKA_LAUNCH(diff_bits_kernel, (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)c.sm_count * 8), 256, 0,
                   reinterpret_cast<const ulonglong2*>(data), n, d_diff);
        unsigned long long h_diff[2];
        SKA_CUDA(cudaMemcpyAsync(h_diff, d_diff, 16, cudaMemcpyDeviceToHost, c.stream));
        SKA_CUDA(cudaStreamSynchronize(c.stream));
        const int msb = h_diff[1] ? 64 + (63 - __builtin_clzll(h_diff[1])) : h_diff[0] ? 63 - __builtin_clzll(h_diff[0]) : -1;
        if (msb < 0) return;                                         // all records equal
        top_bit = std::min(top_bit, msb + 1);
    }
    int P = 0;
    while ((n >> P) > 96 && P < 22 && P < top_bit) P++;
    if (P == 0) P = std::min(1, top_bit);
    const int P1 = (P + 1) / 2, P2 = P - P1;
    uint64_t* grouped = nullptr; uint64_t* off = nullptr;
    group_wide_records(data, n, top_bit, P1, P2, tmp, data, &grouped, &off);
    if (grouped != data) SKA_CUDA(cudaMemcpyAsync(data, grouped, n * 16, cudaMemcpyDeviceToDevice, c.stream));
    const uint32_t nb = 1u << P;
    uint32_t* big_list = arena().take_n<uint32_t>(nb + 1);
    uint32_t* mid_list = arena().take_n<uint32_t>(nb + 1);
    uint32_t* huge_list = arena().take_n<uint32_t>(nb + 1);
    uint32_t* cnts = arena().take_n<uint32_t>(4);
    SKA_CUDA(cudaMemsetAsync(cnts, 0, 16, c.stream));
    SKA_LAUNCH(bucket_sort_small_kernel, (nb + 7) / 8, 256, 0, data, off, nb, big_list, cnts);
    uint32_t n_big = 0;
    SKA_CUDA(cudaMemcpyAsync(&n_big, cnts, 4, cudaMemcpyDeviceToHost, c.stream));
    SKA_CUDA(cudaStreamSynchronize(c.stream));
    std::vector<uint32_t> h_huge; std::vector<uint64_t> h_off;
    sort_big_buckets(data, off, nb, big_list, mid_list, huge_list, cnts, n_big, h_huge, h_off);
    for (uint32_t b : h_huge) sort_range_wide(data + 2 * h_off[b], h_off[b + 1] - h_off[b], top_bit - P, tmp, depth + 1);
}

uint64_t* sort_records_wide(const uint64_t* rec, uint64_t n, int key_bits, uint64_t* buf_a, uint64_t* buf_b, uint64_t* n_out) {
    Context& c = ctx();
    *n_out = 0;
    if (n == 0) return buf_b;
    // first level here (it also drops the all-ones records and tells how many are left), the rest through sort_range_wide
    int P = 0;
    while ((n >> P) > 96 && P < 22 && P < key_bits) P++;
    const int P1 = (P + 1) / 2, P2 = P - P1;
    uint64_t* grouped = nullptr; uint64_t* off = nullptr;
    group_wide_records(rec, n, 64 + key_bits, P1, P2, buf_a, buf_b, &grouped, &off);
    uint64_t* other = grouped == buf_a ? buf_b : buf_a;
    const uint32_t nb = 1u << P;
    uint32_t* big_list = arena().take_n<uint32_t>(nb + 1);
    uint32_t* mid_list = arena().take_n<uint32_t>(nb + 1);
    uint32_t* huge_list = arena().take_n<uint32_t>(nb + 1);
    uint32_t* cnts = arena().take_n<uint32_t>(4);
    SKA_CUDA(cudaMemsetAsync(cnts, 0, 16, c.stream));
    SKA_LAUNCH(bucket_sort_small_kernel, (nb + 7) / 8, 256, 0, grouped, off, nb, big_list, cnts);
    uint32_t n_big = 0; uint64_t kept = 0;
    SKA_CUDA(cudaMemcpyAsync(&n_big, cnts, 4, cudaMemcpyDeviceToHost, c.stream));
    SKA_CUDA(cudaMemcpyAsync(&kept, off + nb, 8, cudaMemcpyDeviceToHost, c.stream));
    SKA_CUDA(cudaStreamSynchronize(c.stream));
    *n_out = kept;
    std::vector<uint32_t> h_huge; std::vector<uint64_t> h_off;
    sort_big_buckets(grouped, off, nb, big_list, mid_list, huge_list, cnts, n_big, h_huge, h_off);
    for (uint32_t b : h_huge) sort_range_wide(grouped + 2 * h_off[b], h_off[b + 1] - h_off[b], 64 + key_bits - P, other, 1);
    return grouped;
}

}  // namespace ska
