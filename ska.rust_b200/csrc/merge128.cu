// merge128.cu -- K2/K3 for k > 31: the bucket merge of merge.cu on split k-mers of up to 124 bits (the reference's u128
// instantiation: src/lib.rs:607-622, UInt for u128 in src/ska_dict/bit_encoding.rs:239-303).
//
// Same reduction as the 64-bit path -- OR the one-hot middle base of every record into cell (key, sample), rows out in
// ascending key order as IUPAC letters -- but a 128-bit key cannot be claimed with one shared-memory CAS.  Per bucket:
//   count pass  distinct keys go to a LIST in shared memory (two words each); an open-addressing table holds 16-bit
//               indices into the list.  A slot is claimed by CAS(EMPTY -> LOCK), the key is written to the list, and the
//               index is published; a thread that meets LOCK retries after the warp reconverges (the owner never waits).
//               The list is then bitonic-sorted (only DISTINCT keys are sorted, not records) and left in the spare
//               partition buffer together with the bucket's row count.
//   emit pass   the sorted keys are reloaded (rank = index), inserted into the index table (no lock: every key once), the
//               records are streamed again and OR-ed into 4-bit sets in shared memory, rows leave as letters.
// Buckets are smaller than in the 64-bit path (~512 records) so that list, table and sets fit; a bucket with more distinct
// keys than fit is split by key range into items that re-scan it (same host loop as merge.cu).
// Roofline: HBM, 16 B per record read twice + rows written once; in practice bound by shared-memory probes like merge.cu.
#include <algorithm>

#include "common.cuh"

namespace ska {

typedef unsigned __int128 u128;
constexpr int W_THREADS = 256;
constexpr int W_LIST = 2048;                 // distinct keys one item may hold
constexpr int W_SLOTS = 4096;                // index table
constexpr uint32_t W_EMPTY = 0xFFFFu, W_LOCK = 0xFFFEu;
constexpr uint32_t W_OVERFLOW = 0xFFFFFFFFu;

// item: records [beg, end) of one prefix bucket restricted to the inclusive range [lo, hi] of key_low (the key without its
// top P1 bits, as the packed records hold it); key_top = those top bits in place
struct ItemW { uint64_t beg, end; uint64_t lo_lo, lo_hi, hi_lo, hi_hi; uint64_t top_lo, top_hi; uint64_t scratch; };
constexpr uint64_t W_NO_SCRATCH = ~0ULL;

__device__ __forceinline__ u128 mk128(uint64_t lo, uint64_t hi) { return ((u128)hi << 64) | lo; }
__device__ __forceinline__ u128 ld128(const uint64_t* __restrict__ p, uint64_t i) {
    const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(p) + i);
    return mk128(v.x, v.y);
}
__device__ __forceinline__ uint32_t slot_of(u128 k) {
    const uint64_t x = ((uint64_t)k ^ ((uint64_t)(k >> 64) * 0xD6E8FEB86659FD93ULL)) * 0x9E3779B97F4A7C15ULL;
    return (uint32_t)(x >> 52);               // 12 bits: W_SLOTS
}

__global__ void make_items_wide_kernel(const uint64_t* __restrict__ off, uint32_t n_buckets, int key_bits, int P, int P2, int low_bits,
                                       ItemW* __restrict__ items) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_buckets) return;
    const int low = key_bits - P;                           // key bits below the bucket prefix
    const u128 full_lo = (u128)j << low;                    // smallest key of the bucket
    const u128 low_mask = (((u128)1) << low_bits) - 1;      // low_bits = key_bits - P1 < 128
    const u128 width = low >= 128 ? ~(u128)0 : ((((u128)1) << low) - 1);
    const u128 kl_lo = full_lo & low_mask, kl_hi = kl_lo | width;
    const u128 top = ((u128)(j >> P2)) << low_bits;
    ItemW it;
    it.beg = off[j]; it.end = off[j + 1];
    it.lo_lo = (uint64_t)kl_lo; it.lo_hi = (uint64_t)(kl_lo >> 64); it.hi_lo = (uint64_t)kl_hi; it.hi_hi = (uint64_t)(kl_hi >> 64);
    it.top_lo = (uint64_t)top; it.top_hi = (uint64_t)(top >> 64);
    it.scratch = it.beg;
    items[j] = it;
}

// ---- count pass
// sorted distinct keys of item `it` live at index it.scratch of a two-part space: [0, split) in the spare partition buffer
// (a bucket's own record range), [split, ...) in an extra region handed to the children of split buckets
struct ScratchW { uint64_t* a; uint64_t* b; uint64_t split; };
__device__ __forceinline__ uint64_t* scratch_at(const ScratchW& sc, uint64_t idx) { return idx < sc.split ? sc.a + 2 * idx : sc.b + 2 * (idx - sc.split); }

__global__ void __launch_bounds__(W_THREADS)
bucket_count_wide_kernel(const uint64_t* __restrict__ rec, const ItemW* __restrict__ items, int sb, uint32_t limit,
                         uint32_t* __restrict__ n_rows, uint32_t* __restrict__ need_split, ScratchW sc) {
    __shared__ uint64_t s_klo[W_LIST], s_khi[W_LIST];
    __shared__ uint16_t s_slot[W_SLOTS];
    __shared__ uint32_t s_n, s_over;
    const ItemW it = items[blockIdx.x];
    if (it.beg == it.end) { if (threadIdx.x == 0) n_rows[blockIdx.x] = 0; return; }
    for (int i = threadIdx.x; i < W_SLOTS; i += W_THREADS) s_slot[i] = (uint16_t)W_EMPTY;
    if (threadIdx.x == 0) { s_n = 0; s_over = 0; }
    __syncthreads();
    const u128 lo = mk128(it.lo_lo, it.lo_hi), hi = mk128(it.hi_lo, it.hi_hi);
    for (uint64_t base = it.beg; base < it.end; base += W_THREADS) {
        const uint64_t i = base + threadIdx.x;
        bool pending = false;
        u128 key = 0;
        if (i < it.end) { key = ld128(rec, i) >> (sb + 2); pending = key >= lo && key <= hi; }
        uint32_t slot = slot_of(key);
        const uint64_t klo = (uint64_t)key, khi = (uint64_t)(key >> 64);
        // every lane takes part in every round, so that a lane that met a locked slot sees it published in the next one
        while (__any_sync(0xFFFFFFFFu, pending)) {
            if (pending && !*(volatile uint32_t*)&s_over) {
                uint32_t idx = *(volatile uint16_t*)&s_slot[slot];
                if (idx == W_EMPTY) {
                    // 16-bit CAS on the slot (aligned 32-bit word holds two slots)
                    uint32_t* w = reinterpret_cast<uint32_t*>(s_slot) + (slot >> 1);
                    const uint32_t sh = (slot & 1u) * 16u;
                    const uint32_t cur = *(volatile uint32_t*)w;
                    if (((cur >> sh) & 0xFFFFu) == W_EMPTY) {
                        const uint32_t want = (cur & ~(0xFFFFu << sh)) | (W_LOCK << sh);
                        if (atomicCAS(w, cur, want) == cur) {
                            const uint32_t n = atomicAdd(&s_n, 1u);
                            if (n >= (uint32_t)W_LIST) { s_over = 1; *(volatile uint16_t*)&s_slot[slot] = (uint16_t)W_EMPTY; }
                            else {
                                s_klo[n] = klo; s_khi[n] = khi;
                                __threadfence_block();
                                *(volatile uint16_t*)&s_slot[slot] = (uint16_t)n;
                            }
                            pending = false;
                        }
                    }
                    // lost the race or the neighbour half-word changed: look again next round
                } else if (idx != W_LOCK) {
                    if (*(volatile uint64_t*)&s_klo[idx] == klo && *(volatile uint64_t*)&s_khi[idx] == khi) pending = false;
                    else slot = (slot + 1) & (W_SLOTS - 1);
                }
            } else pending = false;
            __syncwarp();
        }
    }
    __syncthreads();
    const uint32_t n = s_over ? W_OVERFLOW : s_n;
    if (threadIdx.x == 0) {
        n_rows[blockIdx.x] = n;
        if (n > limit) atomicExch(need_split, 1u);
    }
    if (n > limit || it.scratch == W_NO_SCRATCH || !sc.a) return;
    // sort the distinct keys (bitonic, padded with all-ones)
    uint32_t np2 = 1; while (np2 < n) np2 <<= 1;
    for (uint32_t i = n + threadIdx.x; i < np2; i += W_THREADS) { s_klo[i] = ~0ULL; s_khi[i] = ~0ULL; }
    __syncthreads();
    for (uint32_t size = 2; size <= np2; size <<= 1)
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = threadIdx.x; t < np2 / 2; t += W_THREADS) {
                const uint32_t a = 2 * t - (t & (stride - 1)), b = a + stride;
                const bool up = (a & size) == 0;
                const uint64_t alo = s_klo[a], ahi = s_khi[a], blo = s_klo[b], bhi = s_khi[b];
                const bool a_gt_b = ahi > bhi || (ahi == bhi && alo > blo);
                if (a_gt_b == up) { s_klo[a] = blo; s_khi[a] = bhi; s_klo[b] = alo; s_khi[b] = ahi; }
            }
            __syncthreads();
        }
    uint64_t* dst = scratch_at(sc, it.scratch);
    for (uint32_t i = threadIdx.x; i < n; i += W_THREADS) { dst[2 * i] = s_klo[i]; dst[2 * i + 1] = s_khi[i]; }
}

// reverse complement of a (k-1)-base key held in 128 bits (bit_encoding.rs:241-261)
__device__ __forceinline__ u128 revcomp128(u128 x, int n_bases) {
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    lo = __brevll(lo); hi = __brevll(hi);                     // reversing all 128 bits swaps the halves
    lo = ((lo & 0x5555555555555555ULL) << 1) | ((lo >> 1) & 0x5555555555555555ULL);
    hi = ((hi & 0x5555555555555555ULL) << 1) | ((hi >> 1) & 0x5555555555555555ULL);
    u128 r = mk128(hi, lo);
    r ^= mk128(0xAAAAAAAAAAAAAAAAULL, 0xAAAAAAAAAAAAAAAAULL);
    return r >> (128 - 2 * n_bases);
}
__device__ __forceinline__ uint32_t iupac4_w(uint32_t v) {   // four base sets (one per byte) -> four IUPAC letters ('-' for empty)
    const uint32_t T0 = '-' | ('A' << 8) | ('C' << 16) | ('M' << 24), T1 = 'T' | ('W' << 8) | ('Y' << 16) | ('H' << 24);
    const uint32_t T2 = 'G' | ('R' << 8) | ('S' << 16) | ('V' << 24), T3 = 'K' | ('D' << 8) | ('B' << 16) | ('N' << 24);
    const uint32_t sel = (v & 0x7u) | ((v >> 4) & 0x70u) | ((v >> 8) & 0x700u) | ((v >> 12) & 0x7000u);
    const uint32_t lo = __byte_perm(T0, T1, sel), hi = __byte_perm(T2, T3, sel);
    const uint32_t m = ((v >> 3) & 0x01010101u) * 0xFFu;
    return (hi & m) | (lo & ~m);
}
__device__ __forceinline__ uint32_t pal_letter_w(uint32_t m) {   // ska_dict.rs:85-113
    if (!m) return '-';
    const bool at = m & 5u, cg = m & 10u;
    return at && cg ? 'N' : (at ? 'W' : 'S');
}

// ---- emit pass.  dynamic smem: klo u64[max_u] | khi u64[max_u] | sets u32[max_u * stride_w] | cnt u32[max_u] | slot u16[W_SLOTS] | pal u8[max_u]
__global__ void __launch_bounds__(W_THREADS)
bucket_emit_wide_kernel(const uint64_t* __restrict__ rec, const ItemW* __restrict__ items, int sb, const uint64_t* __restrict__ row_off,
                        uint32_t max_u, uint32_t n_samples, uint32_t stride_w, int k, int rc, uint64_t* __restrict__ out_keys,
                        uint8_t* __restrict__ out_var, uint32_t* __restrict__ out_cnt, ScratchW sc) {
    extern __shared__ __align__(16) uint8_t smem_w[];
    uint64_t* s_klo = reinterpret_cast<uint64_t*>(smem_w);
    uint64_t* s_khi = s_klo + max_u;
    uint32_t* s_sets = reinterpret_cast<uint32_t*>(s_khi + max_u);
    uint32_t* s_cnt = s_sets + (size_t)max_u * stride_w;
    uint16_t* s_slot = reinterpret_cast<uint16_t*>(s_cnt + max_u);
    uint8_t* s_pal = reinterpret_cast<uint8_t*>(s_slot + W_SLOTS);
    const ItemW it = items[blockIdx.x];
    const uint64_t row_base = row_off[blockIdx.x];
    const uint32_t u = (uint32_t)(row_off[blockIdx.x + 1] - row_base);
    if (u == 0) return;
    for (int i = threadIdx.x; i < W_SLOTS; i += W_THREADS) s_slot[i] = (uint16_t)W_EMPTY;
    for (uint32_t i = threadIdx.x; i < u * stride_w; i += W_THREADS) s_sets[i] = 0u;
    __syncthreads();
    const u128 top = mk128(it.top_lo, it.top_hi);
    const uint64_t* src_keys = scratch_at(sc, it.scratch);
    for (uint32_t i = threadIdx.x; i < u; i += W_THREADS) {
        const uint64_t klo = src_keys[2 * i], khi = src_keys[2 * i + 1];
        s_klo[i] = klo; s_khi[i] = khi; s_cnt[i] = 0;
        const u128 full = top | mk128(klo, khi);
        out_keys[2 * (row_base + i)] = (uint64_t)full;
        out_keys[2 * (row_base + i) + 1] = (uint64_t)(full >> 64);
        s_pal[i] = (uint8_t)(rc && revcomp128(full, k - 1) == full);
        // every key once: claim a free slot with a 16-bit CAS (no lock needed, the key is already in the list)
        uint32_t slot = slot_of(mk128(klo, khi));
        while (true) {
            uint32_t* w = reinterpret_cast<uint32_t*>(s_slot) + (slot >> 1);
            const uint32_t sh = (slot & 1u) * 16u;
            const uint32_t cur = *(volatile uint32_t*)w;
            if (((cur >> sh) & 0xFFFFu) == W_EMPTY) {
                if (atomicCAS(w, cur, (cur & ~(0xFFFFu << sh)) | (i << sh)) == cur) break;
            } else slot = (slot + 1) & (W_SLOTS - 1);
        }
    }
    __syncthreads();
    const u128 lo = mk128(it.lo_lo, it.lo_hi), hi = mk128(it.hi_lo, it.hi_hi);
    const uint32_t sid_mask = (1u << sb) - 1u;
    for (uint64_t i = it.beg + threadIdx.x; i < it.end; i += W_THREADS) {
        const u128 r = ld128(rec, i);
        const u128 key = r >> (sb + 2);
        if (key < lo || key > hi) continue;
        const uint64_t klo = (uint64_t)key, khi = (uint64_t)(key >> 64);
        uint32_t slot = slot_of(key), idx;
        uint32_t probes = 0;
        while ((idx = s_slot[slot]) != W_EMPTY && !(s_klo[idx] == klo && s_khi[idx] == khi) && probes < (uint32_t)W_SLOTS) { slot = (slot + 1) & (W_SLOTS - 1); probes++; }
        if (idx == W_EMPTY || probes >= (uint32_t)W_SLOTS) continue;        // cannot happen: the count pass saw every key of the range
        const uint32_t smp = (uint32_t)((uint64_t)r >> 2) & sid_mask;
        atomicOr(&s_sets[idx * stride_w + (smp >> 3)], (1u << ((uint32_t)r & 3u)) << (4 * (smp & 7)));
    }
    __syncthreads();
    uint8_t* outb = out_var + row_base * n_samples;
    const uint32_t total = u * n_samples;
    for (uint32_t o = threadIdx.x; o < total; o += W_THREADS) {
        const uint32_t row = o / n_samples, col = o - row * n_samples;
        const uint32_t m = (s_sets[row * stride_w + (col >> 3)] >> (4 * (col & 7))) & 0xFu;
        if (m) atomicAdd(&s_cnt[row], 1u);
        outb[o] = (uint8_t)(s_pal[row] ? pal_letter_w(m) : (iupac4_w(m) & 0xFFu));
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < u; i += W_THREADS) out_cnt[row_base + i] = s_cnt[i];
}

static size_t emit_wide_smem(uint32_t max_u, uint32_t stride_w) {
    return (size_t)max_u * 16 + (size_t)max_u * stride_w * 4 + (size_t)max_u * 4 + W_SLOTS * 2 + max_u + 32;
}

int merge_prefix_bits_wide(uint64_t n_records_total, uint32_t k) {
    const int key_bits = 2 * ((int)k - 1);
    int p = 0;
    while ((n_records_total >> p) > 512 && p < key_bits && p < 22) p++;   // ~512 records per bucket: a crowded key range still fits the list
    if (const char* e = getenv("SKA_PREFIX_BITS_WIDE")) { int v = atoi(e); if (v >= 0 && v <= 22) p = v; }   // tests: force bucket splitting
    return p;
}

void build_array_from_records_wide(uint64_t* rec, uint64_t n_rec, const uint64_t* d_sample_off, uint32_t n_samples_local, uint32_t sid_base,
                                   uint64_t n_samples_total, uint32_t k, bool rc, Array& out, int prefix_bits,
                                   const unsigned long long* k1_hist) {
    Context& c = ctx();
    const PartPlan pl = make_part_plan(k, prefix_bits >= 0 ? prefix_bits : merge_prefix_bits_wide(n_rec, k), n_samples_total);
    if (!pl.wide) throw Error("internal: wide merge on narrow keys");
    const uint32_t n_buckets = 1u << pl.P;
    uint64_t* prec = nullptr; uint64_t* off = nullptr; uint64_t* part_a = nullptr;
    partition_records(rec, n_rec, d_sample_off, n_samples_local, sid_base, pl, rec, &prec, &off, &part_a, k1_hist);
    ScratchW sc{(prec == rec) ? part_a : rec, nullptr, n_rec};   // the buffer the partition no longer needs: sorted distinct keys, 16 B each

    const uint32_t stride_w = (uint32_t)((n_samples_total + 7) / 8) | 1u;
    const size_t budget = std::min<size_t>((size_t)150 * 1024, c.smem_optin - 2048) - (W_SLOTS * 2 + 64);
    uint32_t limit = (uint32_t)std::min<size_t>(W_LIST - 1, budget / (16 + (size_t)stride_w * 4 + 5));
    if (const char* e = getenv("SKA_WIDE_LIMIT")) limit = std::max(2, std::min<int>((int)limit, atoi(e)));   // tests: force bucket splitting
    if (limit < 2) throw Error("too many samples for the shared-memory rows of the k > 31 merge");

    timer().begin("count");
    ItemW* items = arena().take_n<ItemW>(n_buckets);
    SKA_LAUNCH(make_items_wide_kernel, (n_buckets + 255) / 256, 256, 0, off, n_buckets, pl.key_bits, pl.P, pl.P2, pl.low_bits, items);
    uint32_t n_items = n_buckets;
    uint32_t* n_rows = arena().take_n<uint32_t>(std::max<uint32_t>(n_items, 1));
    uint32_t* flags = arena().take_n<uint32_t>(2);
    uint32_t h_flags[2] = {0, 0};
    for (int round = 0;; round++) {
        dev_fill(flags, 0, 8);
        SKA_LAUNCH(bucket_count_wide_kernel, n_items, W_THREADS, 0, prec, items, pl.sb, limit, n_rows, flags, sc);
        SKA_CUDA(cudaMemcpyAsync(h_flags, flags, 8, cudaMemcpyDeviceToHost, c.stream));
        SKA_CUDA(cudaStreamSynchronize(c.stream));
        if (!h_flags[0]) break;
        if (round > 130) throw Error("bucket splitting did not converge");
        // split the items that hold too many distinct keys by key range (they re-scan their bucket); every item then gets
        // private scratch room for its sorted keys
        std::vector<ItemW> h_items(n_items), next;
        std::vector<uint32_t> h_rows(n_items);
        SKA_CUDA(cudaMemcpyAsync(h_items.data(), items, n_items * sizeof(ItemW), cudaMemcpyDeviceToHost, c.stream));
        SKA_CUDA(cudaMemcpyAsync(h_rows.data(), n_rows, n_items * 4, cudaMemcpyDeviceToHost, c.stream));
        SKA_CUDA(cudaStreamSynchronize(c.stream));
        uint64_t extra = 0;                                 // keys of room handed out in the extra region so far (this round)
        for (uint32_t i = 0; i < n_items; i++) {
            const ItemW& it = h_items[i];
            if (h_rows[i] <= limit) {
                ItemW keep = it;
                if (keep.scratch >= sc.split) { keep.scratch = sc.split + extra; extra += std::min<uint64_t>(limit, it.end - it.beg); }   // re-packed below
                next.push_back(keep);
                continue;
            }
            const u128 lo = ((u128)it.lo_hi << 64) | it.lo_lo, hi = ((u128)it.hi_hi << 64) | it.hi_lo;
            if (lo == hi) throw Error("internal: single-key item overflow");
            int parts = 16;
            if (h_rows[i] != W_OVERFLOW) { parts = 2; while ((uint64_t)parts * limit < 2ull * h_rows[i]) parts *= 2; }
            const u128 step = (hi - lo) / parts + 1;
            for (u128 a = lo;; a += step) {
                ItemW ch = it;
                const bool last = (hi - a) < step;
                const u128 b = last ? hi : a + step - 1;
                ch.lo_lo = (uint64_t)a; ch.lo_hi = (uint64_t)(a >> 64); ch.hi_lo = (uint64_t)b; ch.hi_hi = (uint64_t)(b >> 64);
                // children share their bucket's records, so each gets room of its own in the extra region: it cannot hold
                // more distinct keys than the bucket has records, nor (once it passes) more than `limit`
                ch.scratch = sc.split + extra; extra += std::min<uint64_t>(limit, it.end - it.beg);
                next.push_back(ch);
                if (last) break;
            }
        }
        n_items = (uint32_t)next.size();
        sc.b = arena().take_n<uint64_t>((size_t)extra * 2 + 2);     // the whole round is recounted, so the previous extra region is dead
        items = arena().take_n<ItemW>(n_items); n_rows = arena().take_n<uint32_t>(n_items);
        SKA_CUDA(cudaMemcpyAsync(items, next.data(), n_items * sizeof(ItemW), cudaMemcpyHostToDevice, c.stream));
        SKA_CUDA(cudaStreamSynchronize(c.stream));
    }
    uint64_t* row_off = arena().take_n<uint64_t>(n_items + 1);
    uint64_t total_rows = 0;
    launch_scan_counts(n_rows, n_items, row_off, flags + 1);
    SKA_CUDA(cudaMemcpyAsync(h_flags, flags, 8, cudaMemcpyDeviceToHost, c.stream));
    SKA_CUDA(cudaMemcpyAsync(&total_rows, row_off + n_items, 8, cudaMemcpyDeviceToHost, c.stream));
    SKA_CUDA(cudaStreamSynchronize(c.stream));
    timer().end();
    const uint32_t max_u = std::max<uint32_t>(h_flags[1], 1);

    timer().begin("emit");
    out.k = k; out.rc = rc; out.key_bits = 128; out.n_samples = n_samples_total; out.sorted = true;
    out.n_rows = out.n_keys = total_rows;
    out.keys.alloc(total_rows * 2); out.variants.alloc(total_rows * n_samples_total); out.counts.alloc(total_rows);
    if (total_rows) {
        const size_t smem = emit_wide_smem(max_u, stride_w);
        SKA_CUDA(cudaFuncSetAttribute(bucket_emit_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SKA_LAUNCH(bucket_emit_wide_kernel, n_items, W_THREADS, smem, prec, items, pl.sb, row_off, max_u, (uint32_t)n_samples_total, stride_w,
                   (int)k, rc ? 1 : 0, out.keys.p, out.variants.p, out.counts.p, sc);
    }
    SKA_CUDA(cudaStreamSynchronize(c.stream));
    timer().end();
}

}  // namespace ska
