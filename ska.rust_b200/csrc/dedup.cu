// dedup.cu -- one read set's accepted windows -> its distinct (split k-mer, middle base) records.
//
// The count filter accepts EVERY observation of a k-mer once its count has reached min_count
// (src/ska_dict/bloom_filter.rs:116-148 -> SkaDict::add_to_dict, src/ska_dict/mod.rs:119-153), so at 120x coverage a sample
// hands ~100 identical records per k-mer to the dictionary.  Inserting them is idempotent (the middle base is OR-ed into the
// entry's IUPAC set), so the device drops the repeats sample by sample before the all-sample merge: 64 samples x 5e8
// accepted windows would be 256 GB of records, their distinct ones are 2.5 GB.
//
// The sample's records are partitioned by key prefix (partition.cu) into buckets of ~500 records; one CTA per bucket puts
// the packed records into a shared-memory set and writes the members back in K1's format (key << 4 | one-hot middle) through
// one global cursor -- the order does not matter, the merge partitions them again.  A bucket with more distinct records than
// the set holds sends the sample back with more prefix bits.
// Roofline: HBM; per accepted window 8 (16) B moved twice by the partition and read once here.
#include <algorithm>

#include "common.cuh"

namespace ska {

constexpr int DD_THREADS = 256;
constexpr int DD_SLOTS = 2048;
constexpr int DD_LOG_SLOTS = 11;
constexpr int DD_LIST = 1024;               // wide keys: distinct records per bucket

__global__ void __launch_bounds__(DD_THREADS)
bucket_unique_kernel(const uint64_t* __restrict__ rec, const uint64_t* __restrict__ off, int sb, int low_bits, int P2, uint64_t* __restrict__ out,
                     unsigned long long* __restrict__ cursor, uint32_t* __restrict__ overflow) {
    __shared__ uint64_t s_set[DD_SLOTS];
    __shared__ uint32_t s_n, s_over, s_w;
    __shared__ unsigned long long s_base;
    const uint64_t beg = off[blockIdx.x], end = off[blockIdx.x + 1];
    if (beg == end) return;
    for (int i = threadIdx.x; i < DD_SLOTS; i += DD_THREADS) s_set[i] = ~0ULL;
    if (threadIdx.x == 0) { s_n = 0; s_over = 0; s_w = 0; }
    __syncthreads();
    for (uint64_t i = beg + threadIdx.x; i < end; i += DD_THREADS) {
        const uint64_t v = __ldg(rec + i);                              // key_low << (sb + 2) | sample (0) << 2 | middle
        uint32_t slot = (uint32_t)((v * 0x9E3779B97F4A7C15ULL) >> (64 - DD_LOG_SLOTS));
        while (!*(volatile uint32_t*)&s_over) {
            unsigned long long cur = *(volatile unsigned long long*)&s_set[slot];
            if (cur == ~0ULL) {
                cur = atomicCAS((unsigned long long*)&s_set[slot], ~0ULL, v);
                if (cur == ~0ULL) { if (atomicAdd(&s_n, 1u) >= (uint32_t)(DD_SLOTS - DD_SLOTS / 4)) s_over = 1; break; }
            }
            if (cur == v) break;
            slot = (slot + 1) & (DD_SLOTS - 1);
        }
    }
    __syncthreads();
    if (s_over) { if (threadIdx.x == 0) atomicExch(overflow, 1u); return; }
    if (threadIdx.x == 0) s_base = atomicAdd(cursor, (unsigned long long)s_n);
    __syncthreads();
    const uint64_t prefix = (uint64_t)(blockIdx.x >> P2) << low_bits;
    for (int i = threadIdx.x; i < DD_SLOTS; i += DD_THREADS) {
        const uint64_t v = s_set[i];
        if (v == ~0ULL) continue;
        out[s_base + atomicAdd(&s_w, 1u)] = ((prefix | (v >> (sb + 2))) << 4) | (1ULL << (v & 3));
    }
}

// two-word records (k > 31): list + 16-bit index table, claim-then-publish as in merge128.cu
__global__ void __launch_bounds__(DD_THREADS)
bucket_unique_wide_kernel(const uint64_t* __restrict__ rec, const uint64_t* __restrict__ off, int sb, int low_bits, int P2,
                          uint64_t* __restrict__ out, unsigned long long* __restrict__ cursor, uint32_t* __restrict__ overflow) {
    typedef unsigned __int128 u128;
    constexpr uint32_t EMPTY = 0xFFFFu, LOCK = 0xFFFEu;
    __shared__ uint64_t s_lo[DD_LIST], s_hi[DD_LIST];
    __shared__ uint16_t s_slot[DD_SLOTS];
    __shared__ uint32_t s_n, s_over;
    __shared__ unsigned long long s_base;
    const uint64_t beg = off[blockIdx.x], end = off[blockIdx.x + 1];
    if (beg == end) return;
    for (int i = threadIdx.x; i < DD_SLOTS; i += DD_THREADS) s_slot[i] = (uint16_t)EMPTY;
    if (threadIdx.x == 0) { s_n = 0; s_over = 0; }
    __syncthreads();
    for (uint64_t base = beg; base < end; base += DD_THREADS) {
        const uint64_t i = base + threadIdx.x;
        bool pending = i < end;
        uint64_t lo = 0, hi = 0;
        if (pending) { const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(rec) + i); lo = v.x; hi = v.y; }
        uint32_t slot = (uint32_t)(((lo ^ (hi * 0xD6E8FEB86659FD93ULL)) * 0x9E3779B97F4A7C15ULL) >> (64 - DD_LOG_SLOTS));
        while (__any_sync(0xFFFFFFFFu, pending)) {
            if (pending && !*(volatile uint32_t*)&s_over) {
                const uint32_t idx = *(volatile uint16_t*)&s_slot[slot];
                if (idx == EMPTY) {
                    uint32_t* w = reinterpret_cast<uint32_t*>(s_slot) + (slot >> 1);
                    const uint32_t sh = (slot & 1u) * 16u;
                    const uint32_t cur = *(volatile uint32_t*)w;
                    if (((cur >> sh) & 0xFFFFu) == EMPTY && atomicCAS(w, cur, (cur & ~(0xFFFFu << sh)) | (LOCK << sh)) == cur) {
                        const uint32_t n = atomicAdd(&s_n, 1u);
                        if (n >= (uint32_t)DD_LIST) { s_over = 1; *(volatile uint16_t*)&s_slot[slot] = (uint16_t)EMPTY; }
                        else {
                            s_lo[n] = lo; s_hi[n] = hi;
                            __threadfence_block();
                            *(volatile uint16_t*)&s_slot[slot] = (uint16_t)n;
                        }
                        pending = false;
                    }
                } else if (idx != LOCK) {
                    if (*(volatile uint64_t*)&s_lo[idx] == lo && *(volatile uint64_t*)&s_hi[idx] == hi) pending = false;
                    else slot = (slot + 1) & (DD_SLOTS - 1);
                }
            } else pending = false;
            __syncwarp();
        }
    }
    __syncthreads();
    if (s_over) { if (threadIdx.x == 0) atomicExch(overflow, 1u); return; }
    const uint32_t n = s_n;
    if (threadIdx.x == 0) s_base = atomicAdd(cursor, (unsigned long long)n);
    __syncthreads();
    const u128 prefix = (u128)(blockIdx.x >> P2) << low_bits;
    for (uint32_t i = threadIdx.x; i < n; i += DD_THREADS) {
        const u128 v = ((u128)s_hi[i] << 64) | s_lo[i];
        const u128 r = ((prefix | (v >> (sb + 2))) << 4) | (u128)(1u << ((uint32_t)s_lo[i] & 3u));
        reinterpret_cast<ulonglong2*>(out)[s_base + i] = make_ulonglong2((unsigned long long)r, (unsigned long long)(r >> 64));
    }
}

// rec: n K1 records of ONE sample (sentinels are skipped); overwritten.  -> the distinct records (in `rec` or in arena
// memory), *n_out of them; *n_out = UINT64_MAX when a bucket overflowed at this many prefix bits (the caller re-creates rec
// and calls again with more).
uint64_t* unique_sample_records(uint64_t* rec, uint64_t n, uint32_t k, int prefix_bits, uint64_t* n_out) {
    Context& c = ctx();
    *n_out = 0;
    if (n == 0) return rec;
    const bool wide = k > 31;
    const PartPlan pl = make_part_plan(k, prefix_bits, 1);
    uint64_t h_off[2] = {0, n};
    uint64_t* d_soff = arena().take_n<uint64_t>(2);
    SKA_CUDA(cudaMemcpyAsync(d_soff, h_off, 16, cudaMemcpyHostToDevice, c.stream));
    uint64_t* prec = nullptr; uint64_t* off = nullptr; uint64_t* first = nullptr;
    timer().begin("sample_unique");                                     // one stage: the partition's own stages stay silent
    const bool was_muted = timer().muted;
    timer().muted = true;
    partition_records(rec, n, d_soff, 1, 0, pl, rec, &prec, &off, &first, nullptr);
    uint64_t* out = prec == rec ? first : rec;                          // whichever of the two buffers the grouped records are not in
    unsigned long long* cursor = arena().take_n<unsigned long long>(1);
    uint32_t* over = arena().take_n<uint32_t>(2);
    dev_fill(cursor, 0, 8);
    dev_fill(over, 0, 8);
    const unsigned n_buckets = 1u << pl.P;
    if (wide) SKA_LAUNCH(bucket_unique_wide_kernel, n_buckets, DD_THREADS, 0, prec, off, pl.sb, pl.low_bits, pl.P2, out, cursor, over);
    else SKA_LAUNCH(bucket_unique_kernel, n_buckets, DD_THREADS, 0, prec, off, pl.sb, pl.low_bits, pl.P2, out, cursor, over);
    timer().muted = was_muted;
    timer().end();
    unsigned long long h_n = 0; uint32_t h_over = 0;
    SKA_CUDA(cudaMemcpyAsync(&h_n, cursor, 8, cudaMemcpyDeviceToHost, c.stream));
    SKA_CUDA(cudaMemcpyAsync(&h_over, over, 4, cudaMemcpyDeviceToHost, c.stream));
    SKA_CUDA(cudaStreamSynchronize(c.stream));
    *n_out = h_over ? ~0ULL : (uint64_t)h_n;
    return out;
}

}  // namespace ska
