// cov.cu -- the counting half of `ska cov` / `--min-count auto`: how many distinct split k-mers
// of a read set occur once, twice, ... MAX_COUNT times.
//
// Replaces the dictionary pass of CoverageHistogram::new and the first loop of fit_histogram
// (src/coverage.rs:70-135, 156-163): every window of every read (no quality filter, SplitKmer::new(.., 0, NoFilter),
// coverage.rs:108-117) bumps `kmer_dict[canonical split k-mer]`; counts[c - 1] += 1 for every k-mer with count c <= 1000.
// The Poisson-mixture fit on that 1000-bin histogram stays on the host (ska_host.cpp).
//
// Here: K1 over the reads (ASCII batch without qualities: windows never span reads), the records are partitioned by key
// prefix into buckets of ~1 K records (partition.cu), and one CTA per bucket counts its keys in a shared-memory table --
// key -> occurrences -- then adds one to the histogram bin of every distinct key.  A bucket whose distinct keys overflow the
// table sends the whole pass back with more prefix bits (never seen on read sets: 120x coverage means ~10 keys per bucket).
// Roofline: HBM; 8 (16) B per window written by K1 and moved twice by the partition, read once here.
#include <algorithm>

#include "common.cuh"

namespace ska {

constexpr int CV_THREADS = 256;
constexpr int CV_SLOTS = 2048;              // table slots (static shared memory: 24 KB narrow, 24 KB wide)
constexpr int CV_LOG_SLOTS = 11;
constexpr int CV_LIST = 1024;               // wide keys: distinct keys per bucket
constexpr uint32_t CV_MAX_COUNT = 1000;     // coverage.rs:30

__global__ void __launch_bounds__(CV_THREADS)
bucket_hist_kernel(const uint64_t* __restrict__ rec, const uint64_t* __restrict__ off, int sb, unsigned long long* __restrict__ hist,
                   uint32_t* __restrict__ overflow) {
    __shared__ uint64_t s_key[CV_SLOTS];
    __shared__ uint32_t s_cnt[CV_SLOTS];
    __shared__ uint32_t s_n, s_over;
    const uint64_t beg = off[blockIdx.x], end = off[blockIdx.x + 1];
    if (beg == end) return;
    for (int i = threadIdx.x; i < CV_SLOTS; i += CV_THREADS) { s_key[i] = ~0ULL; s_cnt[i] = 0; }
    if (threadIdx.x == 0) { s_n = 0; s_over = 0; }
    __syncthreads();
    for (uint64_t i = beg + threadIdx.x; i < end; i += CV_THREADS) {
        const uint64_t key = __ldg(rec + i) >> (sb + 2);                 // key without its top P1 bits: unique inside the bucket
        uint32_t slot = (uint32_t)((key * 0x9E3779B97F4A7C15ULL) >> (64 - CV_LOG_SLOTS));
        while (!*(volatile uint32_t*)&s_over) {
            unsigned long long cur = *(volatile unsigned long long*)&s_key[slot];
            if (cur == ~0ULL) {
                cur = atomicCAS((unsigned long long*)&s_key[slot], ~0ULL, key);
                if (cur == ~0ULL) { if (atomicAdd(&s_n, 1u) >= (uint32_t)(CV_SLOTS - CV_SLOTS / 4)) s_over = 1; cur = key; }
            }
            if (cur == key) { atomicAdd(&s_cnt[slot], 1u); break; }
            slot = (slot + 1) & (CV_SLOTS - 1);
        }
    }
    __syncthreads();
    if (s_over) { if (threadIdx.x == 0) atomicExch(overflow, 1u); return; }
    for (int i = threadIdx.x; i < CV_SLOTS; i += CV_THREADS) {
        const uint32_t c = s_cnt[i];
        if (c && c <= CV_MAX_COUNT) atomicAdd(&hist[c - 1], 1ULL);
    }
}

// two-word keys (k > 31): list + index table with the claim-then-publish protocol of merge128.cu
__global__ void __launch_bounds__(CV_THREADS)
bucket_hist_wide_kernel(const uint64_t* __restrict__ rec, const uint64_t* __restrict__ off, int sb, unsigned long long* __restrict__ hist,
                        uint32_t* __restrict__ overflow) {
    typedef unsigned __int128 u128;
    constexpr uint32_t EMPTY = 0xFFFFu, LOCK = 0xFFFEu;
    __shared__ uint64_t s_klo[CV_LIST], s_khi[CV_LIST];
    __shared__ uint32_t s_cnt[CV_LIST];
    __shared__ uint16_t s_slot[CV_SLOTS];
    __shared__ uint32_t s_n, s_over;
    const uint64_t beg = off[blockIdx.x], end = off[blockIdx.x + 1];
    if (beg == end) return;
    for (int i = threadIdx.x; i < CV_SLOTS; i += CV_THREADS) s_slot[i] = (uint16_t)EMPTY;
    for (int i = threadIdx.x; i < CV_LIST; i += CV_THREADS) s_cnt[i] = 0;
    if (threadIdx.x == 0) { s_n = 0; s_over = 0; }
    __syncthreads();
    for (uint64_t base = beg; base < end; base += CV_THREADS) {
        const uint64_t i = base + threadIdx.x;
        bool pending = i < end;
        uint64_t klo = 0, khi = 0;
        if (pending) {
            const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(rec) + i);
            const u128 key = ((((u128)v.y) << 64) | v.x) >> (sb + 2);
            klo = (uint64_t)key; khi = (uint64_t)(key >> 64);
        }
        uint32_t slot = (uint32_t)(((klo ^ (khi * 0xD6E8FEB86659FD93ULL)) * 0x9E3779B97F4A7C15ULL) >> (64 - CV_LOG_SLOTS));
        while (__any_sync(0xFFFFFFFFu, pending)) {
            if (pending && !*(volatile uint32_t*)&s_over) {
                const uint32_t idx = *(volatile uint16_t*)&s_slot[slot];
                if (idx == EMPTY) {
                    uint32_t* w = reinterpret_cast<uint32_t*>(s_slot) + (slot >> 1);
                    const uint32_t sh = (slot & 1u) * 16u;
                    const uint32_t cur = *(volatile uint32_t*)w;
                    if (((cur >> sh) & 0xFFFFu) == EMPTY && atomicCAS(w, cur, (cur & ~(0xFFFFu << sh)) | (LOCK << sh)) == cur) {
                        const uint32_t n = atomicAdd(&s_n, 1u);
                        if (n >= (uint32_t)CV_LIST) { s_over = 1; *(volatile uint16_t*)&s_slot[slot] = (uint16_t)EMPTY; }
                        else {
                            s_klo[n] = klo; s_khi[n] = khi;
                            atomicAdd(&s_cnt[n], 1u);
                            __threadfence_block();
                            *(volatile uint16_t*)&s_slot[slot] = (uint16_t)n;
                        }
                        pending = false;
                    }
                } else if (idx != LOCK) {
                    if (*(volatile uint64_t*)&s_klo[idx] == klo && *(volatile uint64_t*)&s_khi[idx] == khi) { atomicAdd(&s_cnt[idx], 1u); pending = false; }
                    else slot = (slot + 1) & (CV_SLOTS - 1);
                }
            } else pending = false;
            __syncwarp();
        }
    }
    __syncthreads();
    if (s_over) { if (threadIdx.x == 0) atomicExch(overflow, 1u); return; }
    const uint32_t n = s_n;
    for (uint32_t i = threadIdx.x; i < n; i += CV_THREADS) {
        const uint32_t c = s_cnt[i];
        if (c && c <= CV_MAX_COUNT) atomicAdd(&hist[c - 1], 1ULL);
    }
}

// reads: ONE sample (both FASTQ files' records), uploaded without qualities.  h_hist: CV_MAX_COUNT bins.
void kmer_count_histogram(const SeqBatch& reads, uint32_t k, bool rc, uint64_t* h_hist) {
    Context& c = ctx();
    if (reads.packed || reads.has_qual) throw Error("internal: the k-mer counter takes a plain ASCII batch");
    const bool wide = k > 31;
    const int key_bits = 2 * ((int)k - 1);
    int P = 0;
    while ((reads.n_pos >> P) > 512 && P < key_bits && P < 22) P++;
    DevBuf<unsigned long long> hist(CV_MAX_COUNT);
    DevBuf<uint32_t> over(1);
    for (int attempt = 0;; attempt++) {
        arena().reset();
        uint64_t* rec = arena().take_n<uint64_t>(std::max<uint64_t>(reads.n_pos, 1) * (wide ? 2 : 1));
        timer().begin("k1_extract");
        launch_k1(reads, k, rc, false, 0, rec, nullptr, nullptr);
        timer().end();
        const PartPlan pl = make_part_plan(k, P, 1);
        uint64_t* prec = nullptr; uint64_t* off = nullptr;
        partition_records(rec, reads.n_pos, reads.sample_off.p, (uint32_t)reads.n_samples, 0, pl, rec, &prec, &off, nullptr, nullptr);
        dev_fill(hist.p, 0, CV_MAX_COUNT * 8);
        dev_fill(over.p, 0, 4);
        timer().begin("count_hist");
        const unsigned n_buckets = 1u << pl.P;
        if (wide) SKA_LAUNCH(bucket_hist_wide_kernel, n_buckets, CV_THREADS, 0, prec, off, pl.sb, hist.p, over.p);
        else SKA_LAUNCH(bucket_hist_kernel, n_buckets, CV_THREADS, 0, prec, off, pl.sb, hist.p, over.p);
        timer().end();
        uint32_t h_over = 0;
        SKA_CUDA(cudaMemcpyAsync(&h_over, over.p, 4, cudaMemcpyDeviceToHost, c.stream));
        SKA_CUDA(cudaMemcpyAsync(h_hist, hist.p, CV_MAX_COUNT * 8, cudaMemcpyDeviceToHost, c.stream));
        SKA_CUDA(cudaStreamSynchronize(c.stream));
        if (!h_over) return;
        if (pl.P >= std::min(22, key_bits) || attempt > 8) throw Error("k-mer counting: a key range is too crowded for the bucket tables");
        P = std::min(std::min(22, key_bits), pl.P + 2);
    }
}

}  // namespace ska
