// sort128.cu -- ascending sort of 16-byte records (hi word = key, lo word = payload / tie-break), the sort behind the exact
// emulation of the reference's count filter (k1q_reads.cu: observations by (mixed hash, position), distinct hashes by
// (bloom word, first position); src/ska_dict/bloom_filter.rs:116-148).  Replaces the library radix sort of round 1.
//
// MSD: the records are grouped by their top 22 key bits with the two scatter passes of partition.cu (group_wide_records);
// keys here are hashes / bloom words, i.e. uniform, so a bucket holds n / 2^22 records on average (~115 observations of a
// 2 M-pair read set).  One WARP then sorts one bucket IN REGISTERS: a bitonic network over 32 lanes x E records per lane
// (E = 2..16 by bucket size), compare-exchanges inside a lane for strides < E and through __shfl_xor above -- no shared
// memory traffic, 15 of the 36 steps of a 256-record sort cross lanes.  The rare larger bucket (one k-mer repeated thousands
// of times) is sorted by a CTA in shared memory (<= 8192 records) or, beyond that, by recursion on the next key bits.
//
// For the observations the same warp goes on to find the runs of equal hashes in its sorted bucket (a run never leaves its
// bucket: the bucket is a hash prefix) and writes one ENTRY per distinct hash -- count, first position, and the positions of
// observation number min_count - 1 and min_count, the only two the count filter can accept (bloom_filter.rs:131-146) -- so
// the sorted observations themselves are written back only for min_count == 2 / 65535, where every later one is accepted.
// Roofline: HBM; 16 B per record moved by each of two scatters and read once by the bucket sort.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace ska {

constexpr int SW_MAX_E = 16;                 // records per lane of the largest warp sort
constexpr int SW_PER_WARP = 32 * SW_MAX_E;   // records a warp sorts
constexpr int SM_MAX = 2048;                 // records a 256-thread CTA sorts in shared memory (32 KB: several CTAs per SM)
constexpr int SB_MAX = 8192;                 // records a 1024-thread CTA sorts in shared memory
constexpr uint32_t FULL = 0xFFFFFFFFu;

__device__ __forceinline__ bool rec_gt(const ulonglong2& a, const ulonglong2& b) { return a.y > b.y || (a.y == b.y && a.x > b.x); }

// ascending bitonic sort of 32 * E records held as v[r] = record (lane * E + r).  Only the loops over a lane's own records are
// unrolled (register indices must be constants); the merge sizes and the lane strides are run-time loops -- fully unrolled,
// the five bucket sizes came to 36 K instructions (575 KB) and the warps of a CTA, each in its own variant, starved on fetch.
__device__ __forceinline__ void cex(ulonglong2& a, ulonglong2& b, const bool up) {
    if (rec_gt(a, b) == up) { const ulonglong2 t = a; a = b; b = t; }
}
template <int E>
__device__ __forceinline__ void warp_bitonic(ulonglong2 (&v)[E], const int lane) {
    // sizes 2 .. E: inside the lane; the direction of the last one alternates with the lane, of the smaller ones with r
#pragma unroll
    for (int size = 2; size <= E; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
            for (int r = 0; r < E; r++)
                if ((r & stride) == 0) cex(v[r], v[r | stride], size == E ? (lane & 1) == 0 : (r & size) == 0);
        }
    }
    // sizes 2E .. 32E: strides >= E cross lanes (partner lane ^ stride / E, same r), then a merge inside the lane
#pragma unroll 1
    for (int size = 2 * E; size <= 32 * E; size <<= 1) {
        const bool up = ((lane * E) & size) == 0;
#pragma unroll 1
        for (int lm = size / (2 * E); lm > 0; lm >>= 1) {
            const bool keep_min = ((lane & lm) == 0) == up;
#pragma unroll
            for (int r = 0; r < E; r++) {
                ulonglong2 o;
                o.x = __shfl_xor_sync(FULL, v[r].x, lm);
                o.y = __shfl_xor_sync(FULL, v[r].y, lm);
                if (rec_gt(v[r], o) == keep_min) v[r] = o;
            }
        }
#pragma unroll
        for (int stride = E / 2; stride > 0; stride >>= 1) {
#pragma unroll
            for (int r = 0; r < E; r++)
                if ((r & stride) == 0) cex(v[r], v[r | stride], up);
        }
    }
}

// The observation sort works on 96 bits per record instead of 128: inside a bucket the top P hash bits are the bucket's, and
// the next KH (32) bits together with the position order the records unless two different hashes of one bucket agree in all
// of those P + 32 bits -- about once in three read sets.  So the network compares k = next 32 hash bits << 32 | position (one
// 64-bit compare) and carries the remaining hash bits t along; the epilogue sees such a pair (neighbours with equal k-part,
// different t) and hands the whole bucket to the full-width path.
struct Rec96 { uint64_t k; uint32_t t; };
__device__ __forceinline__ void cex96(Rec96& a, Rec96& b, const bool up) {
    if ((a.k > b.k) == up) { const Rec96 x = a; a = b; b = x; }
}
template <int E>
__device__ __forceinline__ void warp_bitonic96(Rec96 (&v)[E], const int lane) {
#pragma unroll
    for (int size = 2; size <= E; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
            for (int r = 0; r < E; r++)
                if ((r & stride) == 0) cex96(v[r], v[r | stride], size == E ? (lane & 1) == 0 : (r & size) == 0);
        }
    }
#pragma unroll 1
    for (int size = 2 * E; size <= 32 * E; size <<= 1) {
        const bool up = ((lane * E) & size) == 0;
#pragma unroll 1
        for (int lm = size / (2 * E); lm > 0; lm >>= 1) {
            const bool keep_min = ((lane & lm) == 0) == up;
#pragma unroll
            for (int r = 0; r < E; r++) {
                Rec96 o;
                o.k = __shfl_xor_sync(FULL, v[r].k, lm);
                o.t = __shfl_xor_sync(FULL, v[r].t, lm);
                if ((v[r].k > o.k) == keep_min) v[r] = o;              // positions are distinct: k never ties between two records
            }
        }
#pragma unroll
        for (int stride = E / 2; stride > 0; stride >>= 1) {
#pragma unroll
            for (int r = 0; r < E; r++)
                if ((r & stride) == 0) cex96(v[r], v[r | stride], up);
        }
    }
}

// what the run-finding epilogue writes (observation sort only)
struct RunsOut {
    ulonglong2* drec;                // per entry: x = count << 32 | entry index, y = bloom word << 32 | first position
    ulonglong2* ent;                 // per entry: x = mixed hash, y = position of observation min_count << 32 | of observation min_count - 1
    uint64_t* run_start;             // per entry: index of its first observation in the sorted array (write_sorted only)
    unsigned long long* n_ent;
    uint32_t min_count;
    uint32_t bloom_words;
    int write_sorted;
    int prefix_bits;                 // hash bits that are the bucket index
    int kh_bits;                     // hash bits below them that the 96-bit network compares (32; fewer only to test the fallback)
};
__device__ __forceinline__ uint32_t word_of(uint64_t mixed, uint32_t words) { return (uint32_t)__umul64hi(mixed, (uint64_t)words); }

constexpr int SW_WARPS = 8;                  // warps per CTA of the warp-sort kernels

// the runs of one sorted bucket, from its hashes and positions in shared memory (record i at s_key[i] / s_pos[i]): one entry
// per run.  Chunks of 32 records, last chunk first, so that every run head knows where the next run starts.
__device__ __forceinline__ bool warp_emit_runs(const uint64_t* __restrict__ s_key, const uint32_t* __restrict__ s_pos, const uint32_t n,
                                               const uint64_t beg, const int lane, const RunsOut& ro) {
    const uint32_t chunks = (n + 31) / 32;
    const int amb_shift = 64 - ro.kh_bits;                                // ((Ka ^ Kb) << P) >> amb_shift == 0: equal in every compared bit
    uint32_t total = 0;
    bool ambiguous = false;
    for (uint32_t j = 0; j < chunks; j++) {
        const uint32_t i = j * 32 + lane;
        const bool head = i < n && (i == 0 || s_key[i - 1] != s_key[i]);
        if (head && i > 0 && (((s_key[i - 1] ^ s_key[i]) << ro.prefix_bits) >> amb_shift) == 0) ambiguous = true;
        total += __popc(__ballot_sync(FULL, head));
    }
    if (__any_sync(FULL, ambiguous)) return false;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(ro.n_ent, (unsigned long long)total);
    base = __shfl_sync(FULL, base, 0);
    uint32_t next_head = n, remaining = total;
    const uint32_t mc = ro.min_count;
    for (uint32_t j = chunks; j-- > 0;) {
        const uint32_t i = j * 32 + lane;
        const uint64_t K = i < n ? s_key[i] : 0;
        const bool head = i < n && (i == 0 || s_key[i - 1] != K);
        const uint32_t m = __ballot_sync(FULL, head);
        remaining -= __popc(m);
        if (head) {
            const uint32_t above = lane == 31 ? 0u : m >> (lane + 1);
            const uint32_t end = above ? i + (uint32_t)__ffs(above) : next_head;
            const uint32_t c = end - i;
            const unsigned long long e = base + remaining + __popc(m & ((1u << lane) - 1u));
            const uint32_t pa = (mc >= 2 && mc - 1 <= c) ? s_pos[i + mc - 2] : 0xFFFFFFFFu;
            const uint32_t pb = (mc >= 1 && mc <= c) ? s_pos[i + mc - 1] : 0xFFFFFFFFu;
            ro.drec[e] = make_ulonglong2(((unsigned long long)c << 32) | e, ((unsigned long long)word_of(K, ro.bloom_words) << 32) | s_pos[i]);
            ro.ent[e] = make_ulonglong2(K, ((unsigned long long)pb << 32) | pa);
            if (ro.write_sorted) ro.run_start[e] = beg + i;
        }
        if (m) next_head = j * 32 + (uint32_t)__ffs(m) - 1;
    }
    return true;
}

template <int E, bool RUNS>
__device__ __forceinline__ void warp_sort_bucket(ulonglong2* __restrict__ r2, const uint32_t n, const int lane, uint64_t* __restrict__ s_key,
                                                 uint32_t* __restrict__ s_pos, const bool write_sorted, const uint32_t bucket, const int P, const int KH) {
    if (!RUNS) {
        ulonglong2 v[E];
#pragma unroll
        for (int r = 0; r < E; r++) {
            const uint32_t i = lane * E + r;
            v[r] = i < n ? r2[i] : make_ulonglong2(~0ULL, ~0ULL);
        }
        warp_bitonic<E>(v, lane);
#pragma unroll
        for (int r = 0; r < E; r++) { const uint32_t i = lane * E + r; if (i < n) r2[i] = v[r]; }
        return;
    }
    Rec96 v[E];
#pragma unroll
    for (int r = 0; r < E; r++) {
        const uint32_t i = lane * E + r;
        if (i < n) {
            const ulonglong2 x = r2[i];                                   // x = position, y = hash
            const uint64_t rem = x.y << P;
            v[r].k = ((rem >> (64 - KH)) << 32) | (uint32_t)x.x;
            v[r].t = (uint32_t)((rem << KH) >> 32);
        } else { v[r].k = ~0ULL; v[r].t = ~0u; }
    }
    warp_bitonic96<E>(v, lane);
    const uint64_t top = P ? (uint64_t)bucket << (64 - P) : 0ULL;
#pragma unroll
    for (int r = 0; r < E; r++) {
        const uint32_t i = lane * E + r;
        if (i < n) {
            const uint64_t rem = ((v[r].k >> 32) << (64 - KH)) | ((uint64_t)v[r].t << (32 - KH));
            const uint64_t K = top | (P ? rem >> P : rem);
            s_key[i] = K; s_pos[i] = (uint32_t)v[r].k;
            if (write_sorted) r2[i] = make_ulonglong2((unsigned long long)(uint32_t)v[r].k, K);
        }
    }
}

// Warps take buckets in turn (grid-stride: a warp that drew small buckets does not wait for its CTA's slowest).  EMAX = 8:
// every bucket (list == NULL); those above 256 records go to `over_list`.  EMAX = 16: the buckets of `list` (*n_list of them),
// 16 records per lane; those above 512 go to `over_list`.  Two kernels because the 16-per-lane network needs 128 registers.
template <int EMAX, bool RUNS>
__global__ void __launch_bounds__(32 * SW_WARPS, EMAX == 8 ? 3 : 2)
bucket_sort_warp_kernel(uint64_t* __restrict__ rec, const uint64_t* __restrict__ off, uint32_t n_buckets, const uint32_t* __restrict__ list,
                        const uint32_t* __restrict__ n_list, uint32_t* __restrict__ over_list, uint32_t* __restrict__ n_over,
                        uint32_t* __restrict__ big_list, uint32_t* __restrict__ n_big, RunsOut ro) {
    __shared__ uint64_t s_key_all[RUNS ? SW_WARPS * 32 * EMAX : 1];
    __shared__ uint32_t s_pos_all[RUNS ? SW_WARPS * 32 * EMAX : 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t* s_key = s_key_all + (RUNS ? warp * 32 * EMAX : 0);
    uint32_t* s_pos = s_pos_all + (RUNS ? warp * 32 * EMAX : 0);
    const uint32_t count = list ? *n_list : n_buckets;
    for (uint32_t idx = blockIdx.x * SW_WARPS + warp; idx < count; idx += gridDim.x * SW_WARPS) {
        const uint32_t b = list ? list[idx] : idx;
        const uint64_t beg = off[b], end = off[b + 1];
        if (end == beg || (!RUNS && end - beg < 2)) continue;
        if (end - beg > 32 * EMAX) { if (lane == 0) over_list[atomicAdd(n_over, 1u)] = b; continue; }
        const uint32_t n = (uint32_t)(end - beg);
        ulonglong2* r2 = reinterpret_cast<ulonglong2*>(rec) + beg;
        const bool ws = RUNS && ro.write_sorted;
        const int P = ro.prefix_bits, KH = ro.kh_bits;
        if (EMAX == 16) warp_sort_bucket<16, RUNS>(r2, n, lane, s_key, s_pos, ws, b, P, KH);
        else if (n <= 32) warp_sort_bucket<1, RUNS>(r2, n, lane, s_key, s_pos, ws, b, P, KH);
        else if (n <= 64) warp_sort_bucket<2, RUNS>(r2, n, lane, s_key, s_pos, ws, b, P, KH);
        else if (n <= 128) warp_sort_bucket<4, RUNS>(r2, n, lane, s_key, s_pos, ws, b, P, KH);
        else warp_sort_bucket<8, RUNS>(r2, n, lane, s_key, s_pos, ws, b, P, KH);
        if (RUNS) {
            __syncwarp();
            // two hashes that the 96-bit order cannot tell apart: the full-width path sorts this bucket again
            if (!warp_emit_runs(s_key, s_pos, n, beg, lane, ro) && lane == 0) big_list[atomicAdd(n_big, 1u)] = b;
            __syncwarp();
        }
    }
}

// both warp kernels over the buckets of one grouping; -> big_list / cnts[0]: the buckets above 512 records
static void launch_warp_sorts(uint64_t* data, const uint64_t* off, uint32_t nb, uint32_t* list16, uint32_t* big_list, uint32_t* cnts /* [0] big, [3] list16 */,
                              const RunsOut* ro) {
    Context& c = ctx();
    const unsigned g8 = (unsigned)std::min<uint64_t>(((uint64_t)nb + SW_WARPS - 1) / SW_WARPS, (uint64_t)c.sm_count * 3 * 8);
    const unsigned g16 = (unsigned)c.sm_count * 2;
    if (ro) {
        SKA_LAUNCH((bucket_sort_warp_kernel<8, true>), g8, 32 * SW_WARPS, 0, data, off, nb, (const uint32_t*)nullptr, (const uint32_t*)nullptr, list16, cnts + 3, big_list, cnts, *ro);
        SKA_LAUNCH((bucket_sort_warp_kernel<16, true>), g16, 32 * SW_WARPS, 0, data, off, nb, (const uint32_t*)list16, (const uint32_t*)(cnts + 3), big_list, cnts, big_list, cnts, *ro);
    } else {
        SKA_LAUNCH((bucket_sort_warp_kernel<8, false>), g8, 32 * SW_WARPS, 0, data, off, nb, (const uint32_t*)nullptr, (const uint32_t*)nullptr, list16, cnts + 3, big_list, cnts, RunsOut{});
        SKA_LAUNCH((bucket_sort_warp_kernel<16, false>), g16, 32 * SW_WARPS, 0, data, off, nb, (const uint32_t*)list16, (const uint32_t*)(cnts + 3), big_list, cnts, big_list, cnts, RunsOut{});
    }
}

// one CTA per large bucket: shared memory up to MAXN records; anything larger goes to the next list (the last one is sorted
// by recursing on the next key bits from the host: a bucket that large is one hash seen thousands of times)
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

// the runs of the large buckets, read back from their sorted records (one CTA per bucket; the end of a run by galloping)
__global__ void __launch_bounds__(256)
bucket_runs_big_kernel(const uint64_t* __restrict__ rec, const uint64_t* __restrict__ off, const uint32_t* __restrict__ big_list, RunsOut ro) {
    const uint32_t b = big_list[blockIdx.x];
    const uint64_t beg = off[b], n = off[b + 1] - beg;
    const ulonglong2* r2 = reinterpret_cast<const ulonglong2*>(rec) + beg;
    for (uint64_t i = threadIdx.x; i < n; i += 256) {
        const uint64_t K = r2[i].y;
        if (i > 0 && r2[i - 1].y == K) continue;
        uint64_t lo = i, step = 1;                                      // r2[lo].y == K; hi: n, or an index whose hash is larger
        while (lo + step < n && r2[lo + step].y == K) { lo += step; step <<= 1; }
        uint64_t hi = min(n, lo + step);
        while (hi - lo > 1) { const uint64_t m = (lo + hi) >> 1; if (r2[m].y == K) lo = m; else hi = m; }
        const uint64_t c = hi - i;
        const uint32_t mc = ro.min_count;
        const unsigned long long e = atomicAdd(ro.n_ent, 1ULL);
        const uint32_t pa = (mc >= 2 && (uint64_t)mc - 1 <= c) ? (uint32_t)r2[i + mc - 2].x : 0xFFFFFFFFu;
        const uint32_t pb = (mc >= 1 && (uint64_t)mc <= c) ? (uint32_t)r2[i + mc - 1].x : 0xFFFFFFFFu;
        ro.drec[e] = make_ulonglong2(((unsigned long long)min(c, (uint64_t)0xFFFFFFFFu) << 32) | e,
                                     ((unsigned long long)word_of(K, ro.bloom_words) << 32) | (uint32_t)r2[i].x);
        ro.ent[e] = make_ulonglong2(K, ((unsigned long long)pb << 32) | pa);
        if (ro.write_sorted) ro.run_start[e] = beg + i;
    }
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

static int sort_prefix_bits(uint64_t n, int max_bits) {
    int P = 0;
    while ((n >> P) > 96 && P < 22 && P < max_bits) P++;
    return P;
}

// sorts data[0, n) in place; bits at and above `top_bit` are equal in all records.  tmp: n records of scratch.
static void sort_range_wide(uint64_t* data, uint64_t n, int top_bit, uint64_t* tmp, int depth) {
    Context& c = ctx();
    if (n < 2 || top_bit <= 0) return;
    if (depth > 8) throw Error("internal: wide sort recursion too deep");
    {   // an oversized bucket is one key seen thousands of times: skip the bits its records share
        unsigned long long* d_diff = arena().take_n<unsigned long long>(2);
        dev_fill(d_diff, 0, 16);
        SKA_LAUNCH(diff_bits_kernel, (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)c.sm_count * 8), 256, 0,
                   reinterpret_cast<const ulonglong2*>(data), n, d_diff);
        unsigned long long h_diff[2];
        SKA_CUDA(cudaMemcpyAsync(h_diff, d_diff, 16, cudaMemcpyDeviceToHost, c.stream));
        SKA_CUDA(cudaStreamSynchronize(c.stream));
        const int msb = h_diff[1] ? 64 + (63 - __builtin_clzll(h_diff[1])) : h_diff[0] ? 63 - __builtin_clzll(h_diff[0]) : -1;
        if (msb < 0) return;                                         // all records equal
        top_bit = std::min(top_bit, msb + 1);
    }
    int P = sort_prefix_bits(n, top_bit);
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
    dev_fill(cnts, 0, 16);
    launch_warp_sorts(data, off, nb, mid_list, big_list, cnts, nullptr);
    uint32_t n_big = 0;
    SKA_CUDA(cudaMemcpyAsync(&n_big, cnts, 4, cudaMemcpyDeviceToHost, c.stream));
    SKA_CUDA(cudaStreamSynchronize(c.stream));
    std::vector<uint32_t> h_huge; std::vector<uint64_t> h_off;
    sort_big_buckets(data, off, nb, big_list, mid_list, huge_list, cnts, n_big, h_huge, h_off);
    for (uint32_t b : h_huge) sort_range_wide(data + 2 * h_off[b], h_off[b + 1] - h_off[b], top_bit - P, tmp, depth + 1);
}

// both entry points: group, sort the buckets (warps, then the CTA tiers and the recursion for what the warps left), and with
// `ro` the run entries of the observation sort (ro->drec is set here: the buffer the grouping left free)
static uint64_t* sort_impl(const uint64_t* rec, uint64_t n, int key_bits, uint64_t* buf_a, uint64_t* buf_b, uint64_t* n_out, RunsOut* ro) {
    Context& c = ctx();
    *n_out = 0;
    if (n == 0) { if (ro) ro->drec = reinterpret_cast<ulonglong2*>(buf_a); return buf_b; }
    const int P = sort_prefix_bits(n, key_bits);
    const int P1 = (P + 1) / 2, P2 = P - P1;
    uint64_t* grouped = nullptr; uint64_t* off = nullptr;
    if (ro) timer().begin("obs_group");                                // the observation sort reports its two halves as stages
    group_wide_records(rec, n, 64 + key_bits, P1, P2, buf_a, buf_b, &grouped, &off);
    if (ro) { timer().end(); timer().begin("obs_sort"); }
    uint64_t* other = grouped == buf_a ? buf_b : buf_a;
    const uint32_t nb = 1u << P;
    uint32_t* big_list = arena().take_n<uint32_t>(nb + 1);
    uint32_t* mid_list = arena().take_n<uint32_t>(nb + 1);
    uint32_t* huge_list = arena().take_n<uint32_t>(nb + 1);
    uint32_t* cnts = arena().take_n<uint32_t>(4);
    dev_fill(cnts, 0, 16);
    if (ro) {
        ro->drec = reinterpret_cast<ulonglong2*>(other);
        ro->prefix_bits = P;
        const char* e = getenv("SKA_SORT_KH_BITS");                     // tests: fewer compared bits, so that the fallback does run
        ro->kh_bits = std::min(32, std::max(e ? atoi(e) : 32, 32 - P));
    }
    launch_warp_sorts(grouped, off, nb, mid_list, big_list, cnts, ro);
    uint32_t n_big = 0; uint64_t kept = 0;
    SKA_CUDA(cudaMemcpyAsync(&n_big, cnts, 4, cudaMemcpyDeviceToHost, c.stream));
    SKA_CUDA(cudaMemcpyAsync(&kept, off + nb, 8, cudaMemcpyDeviceToHost, c.stream));
    SKA_CUDA(cudaStreamSynchronize(c.stream));
    *n_out = kept;
    std::vector<uint32_t> h_huge; std::vector<uint64_t> h_off;
    sort_big_buckets(grouped, off, nb, big_list, mid_list, huge_list, cnts, n_big, h_huge, h_off);
    if (!h_huge.empty()) {
        uint64_t* tmp = other;                                        // scratch of the recursion; with `ro` the other buffer holds the entries
        if (ro) {
            uint64_t longest = 0;
            for (uint32_t b : h_huge) longest = std::max(longest, h_off[b + 1] - h_off[b]);
            tmp = arena().take_n<uint64_t>(2 * longest);
        }
        for (uint32_t b : h_huge) sort_range_wide(grouped + 2 * h_off[b], h_off[b + 1] - h_off[b], 64 + key_bits - P, tmp, 1);
    }
    if (ro && n_big) SKA_LAUNCH(bucket_runs_big_kernel, n_big, 256, 0, (const uint64_t*)grouped, (const uint64_t*)off, (const uint32_t*)big_list, *ro);
    if (ro) timer().end();
    return grouped;
}

uint64_t* sort_records_wide(const uint64_t* rec, uint64_t n, int key_bits, uint64_t* buf_a, uint64_t* buf_b, uint64_t* n_out) {
    return sort_impl(rec, n, key_bits, buf_a, buf_b, n_out, nullptr);
}

// The observations of one read set: rec = n records (x = position, y = mixed hash; all-ones = none).  They are sorted by
// (hash, position) into one of buf_a / buf_b (returned; complete only with write_sorted), and one entry per distinct hash
// goes to *drec (the other of the two buffers) / ent (room for n entries) / run_start (n entries, write_sorted only):
//   drec[e] = (count << 32 | e, bloom word << 32 | first position)
//   ent[e]  = (hash, position of observation number min_count << 32 | of number min_count - 1; 0xFFFFFFFF = no such observation)
// *d_n_ent (device) must be zero on entry and holds the number of entries afterwards.
uint64_t* sort_observations(const uint64_t* rec, uint64_t n, uint64_t* buf_a, uint64_t* buf_b, uint64_t* n_obs, uint64_t** drec, uint64_t* ent,
                            uint64_t* run_start, unsigned long long* d_n_ent, uint32_t min_count, uint32_t bloom_words, bool write_sorted) {
    RunsOut ro;
    ro.drec = nullptr; ro.ent = reinterpret_cast<ulonglong2*>(ent); ro.run_start = run_start;
    ro.prefix_bits = 0; ro.kh_bits = 32;
    ro.n_ent = d_n_ent; ro.min_count = min_count; ro.bloom_words = bloom_words; ro.write_sorted = write_sorted ? 1 : 0;
    uint64_t* sorted = sort_impl(rec, n, 64, buf_a, buf_b, n_obs, &ro);
    *drec = reinterpret_cast<uint64_t*>(ro.drec);
    return sorted;
}

}  // namespace ska
