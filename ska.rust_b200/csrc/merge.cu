// merge.cu -- K2/K3: per-sample de-duplication with IUPAC merge + multi-sample merge into the dense
// (n_kmers x n_samples) middle-base array, in ONE step.
//
// Replaces SkaDict::add_to_dict / add_palindrome_to_dict (src/ska_dict.rs:76-113, one hash-map probe
// per base), MergeSkaDict::append / merge (src/merge_ska_dict.rs:77-151, one Vec<u8> per k-mer) and
// MergeSkaArray::new (src/merge_ska_array.rs:166-186).
//
// Why one step: the per-sample IUPAC merge is a set union (bit_encoding.rs:354-369) and the
// multi-sample merge is a byte-OR of disjoint columns, so both are the same commutative reduction
// "OR the 4-bit base set of every window into cell (key, sample)".  The records K1 wrote
// (key << 4 | one-hot middle, plus the sample id) are
//   (1) partitioned by the top P bits of the key into 2^P buckets that are contiguous in HBM,
//   (2) each bucket is reduced by one CTA in shared memory: open-addressing table key -> slot
//       (atomicCAS), 4-bit set matrix slot x sample (atomicOr), then the bucket's distinct keys are
//       sorted (bitonic) and written out as finished rows: key, N letters, non-missing count.
// Because buckets are key ranges and rows inside a bucket are sorted, the array comes out in
// ascending key order -- the canonical row order -- with no global sort of rows.
// A first pass (count) sizes the output exactly and gives each bucket its row offset.
//
// Step (1) is partition.cu (two MSD passes, packed 8-byte records that carry the sample id).
#include <cstdlib>

#include "common.cuh"

namespace ska {

constexpr int MG_THREADS = 256;
constexpr int MG_UNROLL = 4;               // records probed together by a thread (independent shared-memory loads overlap)
constexpr int MG_PREFETCH = 8;             // records a thread keeps in flight from HBM while it works on the previous batch
#ifndef SKA_MG_CHUNK_LOG
#define SKA_MG_CHUNK_LOG 31
#endif
constexpr uint64_t MG_CHUNK = 1ULL << SKA_MG_CHUNK_LOG;   // records walked with 32-bit positions (tests build a small value to cross chunks)
constexpr int COUNT_CAP = 4096;             // home slots of the counting pass
constexpr int COUNT_LOG_CAP = 12;
constexpr int COUNT_TAIL = 256;             // slots past the last home slot: probing never wraps, so slot order follows key order
constexpr int COUNT_SLOTS = COUNT_CAP + COUNT_TAIL;
constexpr int COUNT_LIST = COUNT_CAP - COUNT_CAP / 4;   // most distinct keys one item may hold, plus one (3072)
constexpr int COUNT_GROUPS = COUNT_SLOTS / 32;
constexpr int COUNT_GPW = COUNT_GROUPS / (MG_THREADS / 32);      // slot groups per warp in the compaction
static_assert(COUNT_GROUPS % (MG_THREADS / 32) == 0, "slot groups must divide evenly among the warps");
constexpr size_t COUNT_SMEM = (size_t)COUNT_SLOTS * 8 + (size_t)COUNT_LIST * 2;
constexpr uint64_t KEY_EMPTY = ~0ULL;

// slot of a key in a table of 2^log_cap slots (log_cap >= 1): the top bits of a 32-bit multiplicative mix of both halves
// (three instructions; the 64-bit golden-ratio product this replaces took five and was 5 % of the emit kernel)
__device__ __forceinline__ uint32_t hash_slot(uint64_t key, int log_cap) {
    return ((uint32_t)key * 0x9E3779B1u + (uint32_t)(key >> 32) * 0x85EBCA77u) >> (32 - log_cap);
}

// A work item = a contiguous record range (one prefix bucket) restricted to an inclusive key range.
// Level-0 items cover their whole bucket; a bucket with too many distinct keys for shared memory is
// split by further key bits into several items that each re-scan the bucket (rare path).
struct Item { uint64_t beg, end, key_lo, key_hi; uint64_t key_base; uint64_t scratch; };   // key = key_base | (prec >> (sb+2))
constexpr uint64_t NO_SCRATCH = ~0ULL;   // scratch: where the count pass leaves this item's sorted distinct keys for the emit pass
constexpr uint32_t COUNT_OVERFLOW = 0xFFFFFFFFu;

// bucket j of this array is prefix bucket b0 + j of the whole key space (b0 > 0 on a row shard)
__global__ void make_items_kernel(const uint64_t* __restrict__ off, uint32_t n_buckets, uint64_t b0, int key_bits, int P, int P2, int low_bits,
                                  Item* __restrict__ items) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_buckets) return;
    const uint64_t b = b0 + j;
    const int low = key_bits - P;                          // key bits below the bucket prefix
    Item it;
    it.beg = off[j]; it.end = off[j + 1];
    it.key_lo = b << low;
    it.key_hi = it.key_lo | ((low >= 64) ? ~0ULL : ((1ULL << low) - 1ULL));
    it.key_base = (b >> P2) << low_bits;
    it.scratch = it.beg;                                   // distinct keys <= records: the bucket's own offset is room enough
    items[j] = it;
}

// ---- pass 1: distinct keys per item, left SORTED in `scratch_keys` for the emit pass.
// The table is order-preserving: a key's home slot is its offset in the item's key range scaled to
// COUNT_CAP slots, and probing never wraps (COUNT_TAIL spare slots at the end), so reading the occupied
// slots in slot order gives the keys in key order except inside a run of adjacent occupied slots, where
// linear probing may have displaced them.  Every key of a run has its home inside the run and homes are
// monotone in the key, so runs are ordered among themselves: a key's rank is its position in slot order,
// minus the larger keys before it in its run, plus the smaller keys after it in its run.  No sort.
// dynamic smem: tab u64[COUNT_SLOTS] (compacted in place into the key list) | lslot u16[COUNT_LIST]
__global__ void __launch_bounds__(MG_THREADS, 4)
bucket_count_kernel(const uint64_t* __restrict__ rec, const Item* __restrict__ items, int sb, uint32_t limit,
                    uint32_t* __restrict__ n_rows, uint32_t* __restrict__ need_split, uint64_t* __restrict__ scratch_keys) {
    extern __shared__ __align__(16) uint8_t smem_c[];
    uint64_t* s_tab = reinterpret_cast<uint64_t*>(smem_c);
    uint64_t* s_list = s_tab;                               // the compaction moves keys down in place
    uint16_t* s_lslot = reinterpret_cast<uint16_t*>(s_tab + COUNT_SLOTS);
    __shared__ uint32_t s_n, s_over;
    __shared__ uint32_t s_goff[COUNT_GROUPS + 1];
    const Item it = items[blockIdx.x];
    if (it.beg == it.end) { if (threadIdx.x == 0) n_rows[blockIdx.x] = 0; return; }
    // HBM latency is hidden by depth, not occupancy: the next MG_PREFETCH records of a thread are in flight while it
    // probes the previous ones
    // The records are walked in chunks of at most 2^31 (one chunk, in practice), so that the positions inside a chunk are
    // 32-bit: the 64-bit index arithmetic was 18 % of this kernel's instructions.
    uint64_t nx[MG_PREFETCH];
    const uint64_t* __restrict__ rc = rec + it.beg;
    uint32_t cn = (uint32_t)min(it.end - it.beg, MG_CHUNK);
    uint64_t rem = it.end - it.beg - cn;                     // records after this chunk
#pragma unroll
    for (int j = 0; j < MG_PREFETCH; j++) {
        const uint32_t o = (uint32_t)j * MG_THREADS + threadIdx.x;
        nx[j] = o < cn ? rc[o] : ~0ULL;
    }
    {
        uint4* t4 = reinterpret_cast<uint4*>(s_tab);
        const uint4 e4 = make_uint4(~0u, ~0u, ~0u, ~0u);
        for (int i = threadIdx.x; i < COUNT_SLOTS / 2; i += MG_THREADS) t4[i] = e4;
    }
    if (threadIdx.x == 0) { s_n = 0; s_over = 0; }
    // keys are handled without their top P1 bits (constant over the item): kl = record >> (sb+2)
    const uint64_t lo_l = it.key_lo - it.key_base, width = it.key_hi - it.key_lo;
    const int range_bits = width ? 64 - __clzll((long long)width) : 0;
    const int shift = max(0, range_bits - COUNT_LOG_CAP);
    __syncthreads();
    while (true) {
    for (uint32_t o0 = 0; o0 < cn; o0 += MG_PREFETCH * MG_THREADS) {
        uint64_t rr[MG_PREFETCH];
#pragma unroll
        for (int j = 0; j < MG_PREFETCH; j++) rr[j] = nx[j];
        const uint32_t o1 = o0 + MG_PREFETCH * MG_THREADS;
        if (o1 < cn) {
#pragma unroll
            for (int j = 0; j < MG_PREFETCH; j++) {
                const uint32_t o = o1 + (uint32_t)j * MG_THREADS + threadIdx.x;
                nx[j] = o < cn ? rc[o] : ~0ULL;
            }
        }
#pragma unroll
        for (int h = 0; h < MG_PREFETCH; h += MG_UNROLL) {
            // probe MG_UNROLL records together: independent shared loads overlap their latency; a key
            // already present (the common case: one record per sample per key) needs nothing else
            uint64_t kk[MG_UNROLL], cur[MG_UNROLL]; uint32_t sl[MG_UNROLL]; bool ok[MG_UNROLL];
#pragma unroll
            for (int j = 0; j < MG_UNROLL; j++) {
                kk[j] = rr[h + j] >> (sb + 2);
                const uint64_t d = kk[j] - lo_l;                    // one unsigned compare covers both ends of the key range
                ok[j] = (o0 + (uint32_t)(h + j) * MG_THREADS + threadIdx.x < cn) && d <= width;
                sl[j] = ok[j] ? (uint32_t)(d >> shift) : 0u;
            }
            // a present key sits at its home slot or, displaced once, right after it (home + 1 < COUNT_SLOTS always): both are
            // read up front so that only new keys and rarer displacements take the divergent path below
            uint64_t cur2[MG_UNROLL];
#pragma unroll
            for (int j = 0; j < MG_UNROLL; j++) {
                cur[j] = *(volatile unsigned long long*)&s_tab[sl[j]];
                cur2[j] = *(volatile unsigned long long*)&s_tab[sl[j] + 1];
            }
#pragma unroll
            for (int j = 0; j < MG_UNROLL; j++) {
                const uint64_t key = kk[j];
                if (ok[j] && cur[j] != key && cur2[j] != key && !*(volatile uint32_t*)&s_over) {
                    uint32_t slot = sl[j];
                    while (true) {
                        if (slot >= (uint32_t)COUNT_SLOTS) { s_over = 1; break; }     // ran off the tail: too dense up here, split the item
                        unsigned long long old = *(volatile unsigned long long*)&s_tab[slot];
                        if (old == key) break;
                        if (old == KEY_EMPTY) {
                            old = atomicCAS((unsigned long long*)&s_tab[slot], KEY_EMPTY, key);
                            if (old == KEY_EMPTY) {
                                if (atomicAdd(&s_n, 1u) >= (uint32_t)COUNT_LIST) s_over = 1;   // keep the table sparse enough to probe
                                break;
                            }
                            if (old == key) break;
                        }
                        slot++;
                    }
                }
                __syncwarp();          // lanes that inserted rejoin the others before the next record (they would otherwise stay apart)
            }
        }
    }
        if (rem == 0) break;                                 // next chunk (items of 2^31 records and more only)
        rc += cn; cn = (uint32_t)min(rem, MG_CHUNK); rem -= cn;
#pragma unroll
        for (int j = 0; j < MG_PREFETCH; j++) {
            const uint32_t o = (uint32_t)j * MG_THREADS + threadIdx.x;
            nx[j] = o < cn ? rc[o] : ~0ULL;
        }
    }
    __syncthreads();
    // (counting the keys at insertion is measured faster than counting the occupied slots afterwards: 2.90 vs 4.26 ms)
    const uint32_t n = s_over ? COUNT_OVERFLOW : s_n;
    if (threadIdx.x == 0) {
        n_rows[blockIdx.x] = n;
        if (n > limit) atomicExch(need_split, 1u);
    }
    if (n > limit || it.scratch == NO_SCRATCH || !scratch_keys) return;
    // occupied slots in slot order -> list (order-preserving compaction: ballot per 32 slots, scan of the group counts)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = MG_THREADS / 32;
    uint64_t keep[COUNT_GPW];                                // every slot is read before any list entry is written (same memory)
#pragma unroll
    for (int q = 0; q < COUNT_GPW; q++) {
        const int g = warp + q * NW;
        keep[q] = s_tab[g * 32 + lane];
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, keep[q] != KEY_EMPTY);
        if (lane == 0) s_goff[g] = __popc(m);
    }
    __syncthreads();
    if (warp == 0) {
        constexpr int PER = (COUNT_GROUPS + 31) / 32;
        uint32_t v[PER], sum = 0;
#pragma unroll
        for (int j = 0; j < PER; j++) { int g = lane * PER + j; v[j] = g < COUNT_GROUPS ? s_goff[g] : 0u; sum += v[j]; }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
        uint32_t run = inc - sum;
#pragma unroll
        for (int j = 0; j < PER; j++) { int g = lane * PER + j; if (g < COUNT_GROUPS) s_goff[g] = run; run += v[j]; }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < COUNT_GPW; q++) {
        const int g = warp + q * NW;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, keep[q] != KEY_EMPTY);
        if (keep[q] != KEY_EMPTY) {
            const uint32_t pos = s_goff[g] + __popc(m & ((1u << lane) - 1u));
            s_list[pos] = keep[q];
            s_lslot[pos] = (uint16_t)(g * 32 + lane);
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += MG_THREADS) {
        const uint64_t key = s_list[i];
        uint32_t rank = i;
        for (uint32_t j = i; j > 0 && (uint32_t)s_lslot[j - 1] + 1u == (uint32_t)s_lslot[j]; j--) rank -= (s_list[j - 1] > key);
        for (uint32_t j = i; j + 1 < n && (uint32_t)s_lslot[j + 1] == (uint32_t)s_lslot[j] + 1u; j++) rank += (s_list[j + 1] < key);
        scratch_keys[it.scratch + rank] = it.key_base | key;
    }
}

// exclusive scan of n u32 counts into u64 offsets (single CTA; n is at most a few million).  Coalesced:
// the CTA walks tiles of 1024 x 4 elements, scans each with warp shuffles and carries the running total.
__global__ void __launch_bounds__(1024)
scan_counts_kernel(const uint32_t* __restrict__ cnt, uint32_t n, uint64_t* __restrict__ off, uint32_t* __restrict__ max_out) {
    __shared__ uint64_t s_warp[32];
    __shared__ uint64_t s_carry;
    __shared__ uint32_t s_max[32];
    const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) s_carry = 0;
    uint32_t mx = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 4096) {
        uint32_t v[4]; uint64_t sum = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) { uint32_t i = base + 4 * t + j; v[j] = i < n ? cnt[i] : 0u; sum += v[j]; mx = max(mx, v[j]); }
        uint64_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint64_t u = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= (uint32_t)o) inc += u; }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint64_t w = s_warp[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint64_t u = __shfl_up_sync(0xFFFFFFFFu, winc, o); if (lane >= (uint32_t)o) winc += u; }
            s_warp[lane] = winc - w;                         // exclusive prefix of the warp totals
        }
        __syncthreads();
        const uint64_t carry = s_carry;
        uint64_t run = carry + s_warp[warp] + (inc - sum);
#pragma unroll
        for (int j = 0; j < 4; j++) { uint32_t i = base + 4 * t + j; if (i < n) off[i] = run; run += v[j]; }
        __syncthreads();
        if (t == 1023) s_carry = run;                        // thread 1023 holds the tile's inclusive total
        __syncthreads();
    }
    // maximum
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
    if (lane == 0) s_max[warp] = mx;
    __syncthreads();
    if (t == 0) {
        uint32_t m = 0;
        for (int w = 0; w < 32; w++) m = max(m, s_max[w]);
        *max_out = m;
        off[n] = s_carry;
    }
}

void launch_scan_counts(const uint32_t* cnt, uint32_t n, uint64_t* off, uint32_t* max_out) {
    SKA_LAUNCH(scan_counts_kernel, 1, 1024, 0, cnt, n, off, max_out);
}

// reverse complement of a (k-1)-base packed key (bit_encoding.rs:182-195)
__device__ __forceinline__ uint64_t revcomp_key(uint64_t x, int n_bases) {
    x = __brevll(x);
    x = ((x & 0x5555555555555555ULL) << 1) | ((x >> 1) & 0x5555555555555555ULL);
    x ^= 0xAAAAAAAAAAAAAAAAULL;
    return x >> (64 - 2 * n_bases);
}

// 4-bit base set (bit = 2-bit code: A=1 C=2 T=4 G=8) -> IUPAC letter, 0 -> '-'   (bit_encoding.rs:354-453)
__device__ __forceinline__ uint32_t iupac_letter(uint32_t m) {
    // index:      0    1    2    3    4    5    6    7    8    9   10   11   12   13   14   15
    // letter:    '-'  'A'  'C'  'M'  'T'  'W'  'Y'  'H'  'G'  'R'  'S'  'V'  'K'  'D'  'B'  'N'
    const uint64_t t0 = ((uint64_t)'-') | ((uint64_t)'A' << 8) | ((uint64_t)'C' << 16) | ((uint64_t)'M' << 24) |
                        ((uint64_t)'T' << 32) | ((uint64_t)'W' << 40) | ((uint64_t)'Y' << 48) | ((uint64_t)'H' << 56);
    const uint64_t t1 = ((uint64_t)'G') | ((uint64_t)'R' << 8) | ((uint64_t)'S' << 16) | ((uint64_t)'V' << 24) |
                        ((uint64_t)'K' << 32) | ((uint64_t)'D' << 40) | ((uint64_t)'B' << 48) | ((uint64_t)'N' << 56);
    return (uint32_t)(((m & 8u) ? t1 : t0) >> (8 * (m & 7u))) & 0xFFu;
}
// four base sets (one per byte, low 4 bits) -> four IUPAC letters, with two byte-permutes over the 16-letter table
__device__ __forceinline__ uint32_t iupac_letters4(uint32_t v) {
    const uint32_t T0 = '-' | ('A' << 8) | ('C' << 16) | ('M' << 24), T1 = 'T' | ('W' << 8) | ('Y' << 16) | ('H' << 24);
    const uint32_t T2 = 'G' | ('R' << 8) | ('S' << 16) | ('V' << 24), T3 = 'K' | ('D' << 8) | ('B' << 16) | ('N' << 24);
    const uint32_t sel = (v & 0x7u) | ((v >> 4) & 0x70u) | ((v >> 8) & 0x700u) | ((v >> 12) & 0x7000u);   // one selector nibble per byte
    const uint32_t lo = __byte_perm(T0, T1, sel), hi = __byte_perm(T2, T3, sel);
    const uint32_t m = ((v >> 3) & 0x01010101u) * 0xFFu;          // 0xFF in bytes whose set contains G (bit 3)
    return (hi & m) | (lo & ~m);
}
// palindromic key: only {A,T} seen -> W, only {C,G} -> S, both -> N   (ska_dict.rs:85-113)
__device__ __forceinline__ uint32_t pal_letter(uint32_t m) {
    if (!m) return '-';
    bool at = m & 5u, cg = m & 10u;
    return at && cg ? 'N' : (at ? 'W' : 'S');
}

// ---- pass 2: reduce one item and emit its rows in key order.
// Phases (one CTA per item, all phases separated by CTA barriers):
//   1. insert the item's keys into a shared-memory open-addressing table (atomicCAS)
//   2. compact the occupied slots and bitonic-sort them by key -> rank of every slot
//   3. write the sorted keys; zero the item's rows of the output matrix (they are L2-resident:
//      a few hundred rows x N bytes)
//   4. stream the records again (L2 hits): probe the table read-only, OR the record's 4-bit base
//      set into byte (row, sample) of the OUTPUT matrix with a 32-bit global RED.OR
//   5. convert the item's rows in place from base sets to IUPAC letters ('-' for empty; W/S/N for
//      palindromic keys, ska_dict.rs:85-113) and count the non-missing cells per row.
// dynamic smem: tab u64[cap] | lkey u64[cap/2] | rank u16[cap] | lslot u16[cap/2] | cnt u32[cap/2]
__global__ void __launch_bounds__(MG_THREADS)
bucket_emit_kernel(const uint64_t* __restrict__ rec, const Item* __restrict__ items, int sb,
                   const uint64_t* __restrict__ row_off, int log_cap, uint32_t n_samples, int k, int rc,
                   uint64_t* __restrict__ out_keys, uint8_t* __restrict__ out_var, uint32_t* __restrict__ out_cnt,
                   const uint64_t* __restrict__ scratch_keys) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t cap = 1u << log_cap, half = cap >> 1;
    uint64_t* s_tab = reinterpret_cast<uint64_t*>(smem);
    uint64_t* s_lkey = s_tab + cap;
    uint32_t* s_cnt = reinterpret_cast<uint32_t*>(s_lkey + half);
    uint16_t* s_rank = reinterpret_cast<uint16_t*>(s_cnt + half);
    uint16_t* s_lslot = s_rank + cap;
    __shared__ uint32_t s_u;

    const uint32_t b = blockIdx.x;
    const Item it = items[b];
    const uint64_t beg = it.beg, end = it.end;
    const uint64_t row_base = row_off[b];
    const uint32_t u = (uint32_t)(row_off[b + 1] - row_base);     // distinct keys, from the count pass
    if (u == 0) return;

    for (uint32_t i = threadIdx.x; i < cap; i += MG_THREADS) s_tab[i] = KEY_EMPTY;
    if (threadIdx.x == 0) s_u = 0;
    __syncthreads();

    const bool from_scratch = scratch_keys && it.scratch != NO_SCRATCH;
    if (from_scratch) {
        // phases 1-2 were done by the count pass: load its sorted distinct keys, insert each once with its rank
        for (uint32_t i = threadIdx.x; i < u; i += MG_THREADS) {
            const uint64_t key = scratch_keys[it.scratch + i];
            s_lkey[i] = key;
            uint32_t slot = hash_slot(key, log_cap);
            while (atomicCAS((unsigned long long*)&s_tab[slot], KEY_EMPTY, key) != KEY_EMPTY) slot = (slot + 1) & (cap - 1);
            s_rank[slot] = (uint16_t)i;
        }
        __syncthreads();
    } else {
    // phase 1
    for (uint64_t base = beg + threadIdx.x; base < end; base += MG_UNROLL * MG_THREADS) {
        uint64_t rr[MG_UNROLL];
#pragma unroll
        for (int j = 0; j < MG_UNROLL; j++) {
            uint64_t i = base + (uint64_t)j * MG_THREADS;
            rr[j] = i < end ? rec[i] : 0;
        }
        uint64_t kk[MG_UNROLL], cur[MG_UNROLL]; uint32_t sl[MG_UNROLL];
#pragma unroll
        for (int j = 0; j < MG_UNROLL; j++) {
            kk[j] = it.key_base | (rr[j] >> (sb + 2));
            sl[j] = hash_slot(kk[j], log_cap);
            if (base + (uint64_t)j * MG_THREADS >= end || kk[j] < it.key_lo || kk[j] > it.key_hi) kk[j] = KEY_EMPTY;
        }
#pragma unroll
        for (int j = 0; j < MG_UNROLL; j++) cur[j] = *(volatile unsigned long long*)&s_tab[sl[j]];
#pragma unroll
        for (int j = 0; j < MG_UNROLL; j++) {
            const uint64_t key = kk[j];
            if (key == KEY_EMPTY || cur[j] == key) continue;
            uint32_t slot = sl[j];
            while (true) {
                unsigned long long old = *(volatile unsigned long long*)&s_tab[slot];
                if (old == key) break;
                if (old == KEY_EMPTY) old = atomicCAS((unsigned long long*)&s_tab[slot], KEY_EMPTY, key);
                if (old == KEY_EMPTY || old == key) break;
                slot = (slot + 1) & (cap - 1);
            }
        }
    }
    __syncthreads();

    // phase 2
    for (uint32_t i = threadIdx.x; i < cap; i += MG_THREADS) {
        uint64_t key = s_tab[i];
        if (key != KEY_EMPTY) {
            uint32_t pos = atomicAdd(&s_u, 1u);
            s_lkey[pos] = key;
            s_lslot[pos] = (uint16_t)i;
        }
    }
    __syncthreads();
    uint32_t np2 = 1; while (np2 < u) np2 <<= 1;
    for (uint32_t i = u + threadIdx.x; i < np2; i += MG_THREADS) { s_lkey[i] = KEY_EMPTY; s_lslot[i] = 0; }
    __syncthreads();
    for (uint32_t size = 2; size <= np2; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = threadIdx.x; t < np2 / 2; t += MG_THREADS) {
                uint32_t lo = 2 * t - (t & (stride - 1));
                uint32_t hi = lo + stride;
                bool up = ((lo & size) == 0);
                uint64_t a = s_lkey[lo], c = s_lkey[hi];
                if ((a > c) == up) {
                    s_lkey[lo] = c; s_lkey[hi] = a;
                    uint16_t sa = s_lslot[lo]; s_lslot[lo] = s_lslot[hi]; s_lslot[hi] = sa;
                }
            }
            __syncthreads();
        }
    }
    for (uint32_t i = threadIdx.x; i < u; i += MG_THREADS) s_rank[s_lslot[i]] = (uint16_t)i;
    __syncthreads();
    }   // !from_scratch

    // phase 3: keys, palindrome flags (in lslot), zero the rows
    for (uint32_t i = threadIdx.x; i < u; i += MG_THREADS) {
        uint64_t key = s_lkey[i];
        out_keys[row_base + i] = key;
        bool pal = rc && (revcomp_key(key, k - 1) == key);
        s_lslot[i] = (uint16_t)(pal ? 1u : 0u);
        s_cnt[i] = 0;
    }
    const uint64_t byte_beg = row_base * n_samples, byte_end = (row_base + u) * n_samples;
    const uint64_t w_beg = byte_beg >> 2, w_end = (byte_end + 3) >> 2;    // 32-bit words touching the rows
    uint32_t* out32 = reinterpret_cast<uint32_t*>(out_var);
    const bool aligned = ((n_samples & 3u) == 0);
    if (aligned) {
        for (uint64_t w = w_beg + threadIdx.x; w < w_end; w += MG_THREADS) out32[w] = 0u;
    } else {
        // rows of neighbouring items may share the boundary words: clear bytes with atomics there
        for (uint64_t o = byte_beg + threadIdx.x; o < byte_end; o += MG_THREADS)
            atomicAnd(&out32[o >> 2], ~(0xFFu << (8 * (o & 3))));
    }
    __threadfence();
    __syncthreads();

    // phase 4
    const uint32_t sid_mask = (1u << sb) - 1u;
    for (uint64_t base = beg + threadIdx.x; base < end; base += MG_UNROLL * MG_THREADS) {
        uint64_t rr[MG_UNROLL];
#pragma unroll
        for (int j = 0; j < MG_UNROLL; j++) {
            uint64_t i = base + (uint64_t)j * MG_THREADS;
            rr[j] = i < end ? rec[i] : 0;
        }
        uint64_t kk[MG_UNROLL], cur[MG_UNROLL]; uint32_t sl[MG_UNROLL], rk[MG_UNROLL];
#pragma unroll
        for (int j = 0; j < MG_UNROLL; j++) {
            kk[j] = it.key_base | (rr[j] >> (sb + 2));
            sl[j] = hash_slot(kk[j], log_cap);
            if (base + (uint64_t)j * MG_THREADS >= end || kk[j] < it.key_lo || kk[j] > it.key_hi) kk[j] = KEY_EMPTY;
        }
#pragma unroll
        for (int j = 0; j < MG_UNROLL; j++) cur[j] = s_tab[sl[j]];
#pragma unroll
        for (int j = 0; j < MG_UNROLL; j++)                       // collisions: walk on (rare)
            if (kk[j] != KEY_EMPTY) while (cur[j] != kk[j]) { sl[j] = (sl[j] + 1) & (cap - 1); cur[j] = s_tab[sl[j]]; }
#pragma unroll
        for (int j = 0; j < MG_UNROLL; j++) rk[j] = s_rank[sl[j]];
#pragma unroll
        for (int j = 0; j < MG_UNROLL; j++) {
            if (kk[j] == KEY_EMPTY) continue;
            const uint32_t smp = (uint32_t)(rr[j] >> 2) & sid_mask;
            uint64_t o = (row_base + rk[j]) * n_samples + smp;
            atomicOr(&out32[o >> 2], (1u << (uint32_t)(rr[j] & 3)) << (8 * (o & 3)));
        }
    }
    __threadfence();
    __syncthreads();

    // phase 5
    if (aligned) {
        const uint32_t wpr = n_samples >> 2;                      // words per row
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        constexpr int NW = MG_THREADS / 32;
        for (uint32_t i0 = warp; i0 < u; i0 += MG_UNROLL * NW) {   // one warp per row, MG_UNROLL rows in flight
            uint32_t nz[MG_UNROLL];
#pragma unroll
            for (int j = 0; j < MG_UNROLL; j++) nz[j] = 0;
            for (uint32_t w0 = lane; w0 < wpr; w0 += 32) {
                uint32_t vv[MG_UNROLL];
#pragma unroll
                for (int j = 0; j < MG_UNROLL; j++) {
                    uint32_t i = i0 + j * NW;
                    vv[j] = i < u ? __ldcg(&out32[w_beg + (uint64_t)i * wpr + w0]) : 0u;
                }
#pragma unroll
                for (int j = 0; j < MG_UNROLL; j++) {
                    uint32_t i = i0 + j * NW;
                    if (i >= u) break;
                    const bool pal = s_lslot[i];                  // warp-uniform: one row per warp
                    const uint32_t v = vv[j];
                    uint32_t word;
                    nz[j] += __popc((v + 0x7F7F7F7Fu) & 0x80808080u);   // non-empty cells (sets are 1..15)
                    if (!pal) word = iupac_letters4(v);
                    else {
                        word = 0;
#pragma unroll
                        for (int q = 0; q < 4; q++) word |= pal_letter((v >> (8 * q)) & 0xFu) << (8 * q);
                    }
                    out32[w_beg + (uint64_t)i * wpr + w0] = word;
                }
            }
#pragma unroll
            for (int j = 0; j < MG_UNROLL; j++) {
                uint32_t i = i0 + j * NW;
                uint32_t t = __reduce_add_sync(0xFFFFFFFFu, nz[j]);
                if (lane == 0 && i < u) s_cnt[i] = t;
            }
        }
    } else {
        for (uint64_t o = byte_beg + threadIdx.x; o < byte_end; o += MG_THREADS) {
            uint32_t i = (uint32_t)((o - byte_beg) / n_samples);
            uint32_t m = (__ldcg(&out32[o >> 2]) >> (8 * (o & 3))) & 0xFu;
            bool pal = s_lslot[i];
            if (m) atomicAdd(&s_cnt[i], 1u);
            uint32_t letter = pal ? pal_letter(m) : iupac_letter(m);
            // byte store; neighbouring bytes of the word may belong to another item's rows
            out_var[o] = (uint8_t)letter;
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < u; i += MG_THREADS) out_cnt[row_base + i] = s_cnt[i];
}

// ---- pass 2, shared-memory form (the default): the item's base sets are accumulated in SHARED memory, one 4-bit set
// per (row, sample), and every row leaves the CTA exactly once as finished IUPAC letters in coalesced 8-byte stores.
// Needs the sorted keys of the count pass (rank of a key = its index there) and  u * N/2  bytes of shared memory
// for the item's u rows; merge_items falls back to bucket_emit_kernel when that does not fit.
//   1. load the u sorted keys: write them out, insert each into the open-addressing table with its rank
//   2. stream the records once: probe (read-only), atomicOr the record's base into nibble (rank, sample)
//   3. rows out: 8 nibbles -> 8 letters ('-' for empty; W/S/N for palindromic keys, ska_dict.rs:85-113), non-missing counts
// dynamic smem: tab u64[cap] | sets u32[max_u * stride_w] | cnt u32[max_u] | rank u16[cap] | pal u8[max_u]
__device__ __forceinline__ uint32_t spread_nibbles4(uint32_t b16) {          // 4 nibbles -> 4 bytes
    uint32_t v = (b16 | (b16 << 8)) & 0x00FF00FFu;
    return (v | (v << 4)) & 0x0F0F0F0Fu;
}
__device__ __forceinline__ uint32_t letters4(uint32_t v, bool pal) {
    if (!pal) return iupac_letters4(v);
    uint32_t word = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) word |= pal_letter((v >> (8 * q)) & 0xFu) << (8 * q);
    return word;
}

__global__ void __launch_bounds__(MG_THREADS, 4)
bucket_emit_smem_kernel(const uint64_t* __restrict__ rec, const Item* __restrict__ items, int sb,
                        const uint64_t* __restrict__ row_off, int log_cap, uint32_t max_u, uint32_t n_samples, uint32_t stride_w,
                        int k, int rc, uint64_t* __restrict__ out_keys, uint8_t* __restrict__ out_var, uint32_t* __restrict__ out_cnt,
                        const uint64_t* __restrict__ scratch_keys, uint32_t u_min /* this launch takes items with u_min < u <= max_u */) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t cap = 1u << log_cap;
    uint64_t* s_tab = reinterpret_cast<uint64_t*>(smem);
    uint32_t* s_sets = reinterpret_cast<uint32_t*>(s_tab + cap);
    uint32_t* s_cnt = s_sets + (size_t)max_u * stride_w;
    uint16_t* s_rank = reinterpret_cast<uint16_t*>(s_cnt + max_u);
    uint8_t* s_pal = reinterpret_cast<uint8_t*>(s_rank + cap);

    __shared__ uint32_t s_anypal;
    const uint32_t b = blockIdx.x;
    const Item it = items[b];
    const uint64_t beg = it.beg, end = it.end;
    const uint64_t row_base = row_off[b];
    const uint32_t u = (uint32_t)(row_off[b + 1] - row_base);     // distinct keys, from the count pass
    if (u <= u_min || u > max_u) return;                          // empty, or another launch's size class
    if (threadIdx.x == 0) s_anypal = 0;

    uint64_t nx[MG_PREFETCH];                                     // first records: in flight while the table is set up
    const uint64_t* __restrict__ rp = rec + beg;                  // chunks of at most 2^31 records: 32-bit positions inside (see bucket_count_kernel)
    uint32_t n = (uint32_t)min(end - beg, MG_CHUNK);
    uint64_t rem = end - beg - n;
#pragma unroll
    for (int j = 0; j < MG_PREFETCH; j++) {
        const uint32_t o = (uint32_t)j * MG_THREADS + threadIdx.x;
        nx[j] = o < n ? rp[o] : ~0ULL;
    }
    {
        uint4* t4 = reinterpret_cast<uint4*>(s_tab);
        const uint4 e4 = make_uint4(~0u, ~0u, ~0u, ~0u);
        for (uint32_t i = threadIdx.x; i < cap / 2; i += MG_THREADS) t4[i] = e4;
    }
    for (uint32_t i = threadIdx.x; i < u * stride_w; i += MG_THREADS) s_sets[i] = 0u;
    __syncthreads();
    // 1. keys (sorted by the count pass); the table holds them without the item's constant top bits
    for (uint32_t i = threadIdx.x; i < u; i += MG_THREADS) {
        const uint64_t key = scratch_keys[it.scratch + i];
        out_keys[row_base + i] = key;
        const bool pal = rc && revcomp_key(key, k - 1) == key;
        s_pal[i] = (uint8_t)(pal ? 1 : 0);
        if (pal) s_anypal = 1;
        s_cnt[i] = 0;
        const uint64_t kl = key ^ it.key_base;
        uint32_t slot = hash_slot(kl, log_cap);
        while (atomicCAS((unsigned long long*)&s_tab[slot], KEY_EMPTY, kl) != KEY_EMPTY) slot = (slot + 1) & (cap - 1);
        s_rank[slot] = (uint16_t)i;
    }
    __syncthreads();

    // 2. records -> base sets
    const uint32_t sid_mask = (1u << sb) - 1u;
    const uint32_t word_mask = sid_mask >> 3, bit_mask = sb >= 3 ? 31u : (1u << (sb + 2)) - 1u;   // fewer than 8 samples: key bits follow at once
    while (true) {
    for (uint32_t o0 = 0; o0 < n; o0 += MG_PREFETCH * MG_THREADS) {
        uint64_t rr[MG_PREFETCH];
#pragma unroll
        for (int j = 0; j < MG_PREFETCH; j++) rr[j] = nx[j];
        const uint32_t o1 = o0 + MG_PREFETCH * MG_THREADS;
        if (o1 < n) {
#pragma unroll
            for (int j = 0; j < MG_PREFETCH; j++) {
                const uint32_t o = o1 + (uint32_t)j * MG_THREADS + threadIdx.x;
                nx[j] = o < n ? rp[o] : ~0ULL;
            }
        }
#pragma unroll
        for (int h = 0; h < MG_PREFETCH; h += MG_UNROLL) {
            uint64_t kk[MG_UNROLL], cur[MG_UNROLL]; uint32_t sl[MG_UNROLL], rk[MG_UNROLL];
#pragma unroll
            for (int j = 0; j < MG_UNROLL; j++) {
                kk[j] = rr[h + j] >> (sb + 2);
                sl[j] = hash_slot(kk[j], log_cap);
            }
#pragma unroll
            for (int j = 0; j < MG_UNROLL; j++) {                  // home slot and the next one, branch-free: covers ~98 % of the probes
                const uint32_t s1 = (sl[j] + 1) & (cap - 1);
                const uint64_t c0 = s_tab[sl[j]], c1 = s_tab[s1];
                const bool second = c0 != kk[j] && c1 == kk[j];
                cur[j] = second ? c1 : c0;
                sl[j] = second ? s1 : sl[j];
            }
#pragma unroll
            for (int j = 0; j < MG_UNROLL; j++) {                  // longer collision chains: walk on (rare); bounded so that a foreign key cannot hang the CTA
                const bool live = o0 + (uint32_t)(h + j) * MG_THREADS + threadIdx.x < n;
                uint32_t probes = 0;
                while (live && cur[j] != kk[j] && probes < cap) { sl[j] = (sl[j] + 1) & (cap - 1); cur[j] = s_tab[sl[j]]; probes++; }
                if (!live || cur[j] != kk[j]) kk[j] = KEY_EMPTY;
                __syncwarp();
            }
#pragma unroll
            for (int j = 0; j < MG_UNROLL; j++) rk[j] = s_rank[sl[j]];
#pragma unroll
            for (int j = 0; j < MG_UNROLL; j++) {
                if (kk[j] != KEY_EMPTY) {
                    // record = key_low | sample << 2 | middle: its low five bits are the bit of (sample & 7, middle) inside the set
                    // word, the next bits of the sample id the word of the row
                    const uint32_t lo = (uint32_t)rr[h + j];
                    atomicOr(&s_sets[rk[j] * stride_w + ((lo >> 5) & word_mask)], 1u << (lo & bit_mask));
                }
            }
            __syncwarp();
        }
    }
        if (rem == 0) break;                                      // next chunk (items of 2^31 records and more only)
        rp += n; n = (uint32_t)min(rem, MG_CHUNK); rem -= n;
#pragma unroll
        for (int j = 0; j < MG_PREFETCH; j++) {
            const uint32_t o = (uint32_t)j * MG_THREADS + threadIdx.x;
            nx[j] = o < n ? rp[o] : ~0ULL;
        }
    }
    __syncthreads();

    // 3. rows out
    if ((n_samples & 7u) == 0) {
        const uint32_t wpr = n_samples >> 3;                       // set words (8 samples each) per row
        const int lg = (wpr & (wpr - 1)) == 0 ? 31 - __clz(wpr) : -1;
        uint2* out8 = reinterpret_cast<uint2*>(out_var + row_base * n_samples);
        const uint32_t total = u * wpr;
        for (uint32_t idx = threadIdx.x; idx < total; idx += MG_THREADS) {
            const uint32_t row = lg >= 0 ? idx >> lg : idx / wpr;
            const uint32_t w = idx - row * wpr;
            const uint32_t x = s_sets[row * stride_w + w];
            // branch-free: an empty set maps to '-' by the same table; palindromic rows (rare) are redone below
            uint2 o;
            o.x = iupac_letters4(spread_nibbles4(x & 0xFFFFu));
            o.y = iupac_letters4(spread_nibbles4(x >> 16));
            const uint32_t m = (x | (x >> 1) | (x >> 2) | (x >> 3)) & 0x11111111u;
            if (m) atomicAdd(&s_cnt[row], (uint32_t)__popc(m));
            out8[idx] = o;
        }
        if (s_anypal) {                                            // W/S/N for palindromic keys (ska_dict.rs:85-113)
            __syncthreads();
            for (uint32_t idx = threadIdx.x; idx < total; idx += MG_THREADS) {
                const uint32_t row = lg >= 0 ? idx >> lg : idx / wpr;
                if (!s_pal[row]) continue;
                const uint32_t x = s_sets[row * stride_w + (idx - row * wpr)];
                out8[idx] = make_uint2(letters4(spread_nibbles4(x & 0xFFFFu), true), letters4(spread_nibbles4(x >> 16), true));
            }
        }
    } else {
        uint8_t* outb = out_var + row_base * n_samples;
        const uint32_t total = u * n_samples;
        for (uint32_t o = threadIdx.x; o < total; o += MG_THREADS) {
            const uint32_t row = o / n_samples, col = o - row * n_samples;
            const uint32_t m = (s_sets[row * stride_w + (col >> 3)] >> (4 * (col & 7))) & 0xFu;
            if (m) atomicAdd(&s_cnt[row], 1u);
            outb[o] = (uint8_t)(s_pal[row] ? pal_letter(m) : iupac_letter(m));
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < u; i += MG_THREADS) out_cnt[row_base + i] = s_cnt[i];
}

static size_t emit_smem_bytes(int log_cap, uint32_t max_u, uint32_t stride_w) {
    const size_t cap = (size_t)1 << log_cap;
    return cap * 8 + (size_t)max_u * stride_w * 4 + (size_t)max_u * 4 + cap * 2 + max_u + 16;
}

// ---- fused merge: ONE pass over the records ------------------------------------------------------------------------
// The count pass above exists only to size the output and to hand sorted keys to the emit pass.  When an item's
// table AND its base sets fit shared memory together, one CTA can do everything for its bucket in a single read of
// the records: order-preserving table (as in bucket_count_kernel) with the 4-bit base sets indexed by SLOT, then the
// rows leave in key order.  What is missing is the bucket's row offset = distinct keys of all earlier buckets; it
// comes from a decoupled look-back over per-bucket status words (buckets are taken in ticket order, so every
// predecessor is running or done and publishes its count without waiting for anyone).  The output is allocated from
// an estimate (distinct keys of every S-th bucket, counted by bucket_count_kernel) plus a margin; a table overflow,
// an exceeded capacity or a stalled look-back raise an error flag and merge_items falls back to the two passes.
// dynamic smem: tab u64[slots] (later: cnt u32[slots]) | sets u32[slots * wps] | lslot u16[slots] | order u16[slots]
constexpr int FM_THREADS = 512;
constexpr size_t FM_SMEM_BUDGET = 110 * 1024;             // two CTAs per SM
constexpr uint64_t ST_AGG = 1ULL << 62, ST_PREFIX = 2ULL << 62, ST_MASK = (1ULL << 62) - 1ULL;
enum { FM_ERR_TABLE = 1, FM_ERR_CAPACITY = 2, FM_ERR_STALL = 4 };
struct FusedCtl { uint32_t ticket, error; unsigned long long total_rows; };

__global__ void __launch_bounds__(FM_THREADS, 2)
bucket_merge_kernel(const uint64_t* __restrict__ rec, const Item* __restrict__ items, uint32_t n_items, int sb, uint32_t slots,
                    uint32_t home_slots, uint32_t max_keys, uint32_t n_samples, int k, int rc, uint64_t cap_rows,
                    uint64_t* __restrict__ out_keys, uint8_t* __restrict__ out_var, uint32_t* __restrict__ out_cnt,
                    unsigned long long* __restrict__ status, FusedCtl* __restrict__ ctl) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t wps = n_samples >> 3;                           // set words (8 samples each) per slot / row
    uint64_t* s_tab = reinterpret_cast<uint64_t*>(smem);
    uint32_t* s_sets = reinterpret_cast<uint32_t*>(s_tab + slots);
    uint16_t* s_lslot = reinterpret_cast<uint16_t*>(s_sets + (size_t)slots * wps);
    uint16_t* s_order = s_lslot + slots;
    __shared__ uint32_t s_goff[160];                               // slot groups (slots <= 4096 -> <= 128 groups)
    __shared__ uint32_t s_n, s_over, s_bid, s_anypal;
    __shared__ unsigned long long s_row_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = FM_THREADS / 32;

    if (tid == 0) { s_bid = atomicAdd(&ctl->ticket, 1u); s_n = 0; s_over = 0; s_anypal = 0; }
    __syncthreads();
    const uint32_t bid = s_bid;
    if (bid >= n_items) return;
    const Item it = items[bid];
    const uint64_t beg = it.beg, end = it.end;

    uint64_t nx[MG_PREFETCH];                                     // first records: in flight while the table is set up
#pragma unroll
    for (int j = 0; j < MG_PREFETCH; j++) {
        const uint64_t i = beg + (uint64_t)j * FM_THREADS + tid;
        nx[j] = i < end ? rec[i] : ~0ULL;
    }
    {
        uint4* t4 = reinterpret_cast<uint4*>(s_tab);
        const uint4 e4 = make_uint4(~0u, ~0u, ~0u, ~0u), z4 = make_uint4(0u, 0u, 0u, 0u);
        for (uint32_t i = tid; i < slots / 2; i += FM_THREADS) t4[i] = e4;
        uint4* z = reinterpret_cast<uint4*>(s_sets);
        for (uint32_t i = tid; i < slots * wps / 4; i += FM_THREADS) z[i] = z4;
    }
    const uint64_t lo_l = it.key_lo - it.key_base, width = it.key_hi - it.key_lo;
    const int range_bits = width ? 64 - __clzll((long long)width) : 0;
    const int shift = range_bits - 12;                             // offset in the key range -> 12 bits -> home slot (monotone)
    const uint32_t sid_mask = (1u << sb) - 1u;
    __syncthreads();

    // 1. records -> table + base sets
    for (uint64_t b0 = beg; b0 < end; b0 += (uint64_t)MG_PREFETCH * FM_THREADS) {
        uint64_t rr[MG_PREFETCH];
#pragma unroll
        for (int j = 0; j < MG_PREFETCH; j++) rr[j] = nx[j];
        const uint64_t b1 = b0 + (uint64_t)MG_PREFETCH * FM_THREADS;
        if (b1 < end) {
#pragma unroll
            for (int j = 0; j < MG_PREFETCH; j++) {
                const uint64_t i = b1 + (uint64_t)j * FM_THREADS + tid;
                nx[j] = i < end ? rec[i] : ~0ULL;
            }
        }
#pragma unroll
        for (int h = 0; h < MG_PREFETCH; h += MG_UNROLL) {
            uint64_t kk[MG_UNROLL], cur[MG_UNROLL]; uint32_t sl[MG_UNROLL]; bool ok[MG_UNROLL];
#pragma unroll
            for (int j = 0; j < MG_UNROLL; j++) {
                kk[j] = rr[h + j] >> (sb + 2);
                const uint64_t d = kk[j] - lo_l;
                ok[j] = (b0 + (uint64_t)(h + j) * FM_THREADS + tid < end) && d <= width;
                const uint32_t h12 = (uint32_t)(shift >= 0 ? d >> shift : d << -shift) & 0xFFFu;
                sl[j] = ok[j] ? (h12 * home_slots) >> 12 : 0u;
            }
#pragma unroll
            for (int j = 0; j < MG_UNROLL; j++) cur[j] = *(volatile unsigned long long*)&s_tab[sl[j]];
#pragma unroll
            for (int j = 0; j < MG_UNROLL; j++) {
                const uint64_t key = kk[j];
                uint32_t slot = sl[j];
                bool found = ok[j] && cur[j] == key;
                if (ok[j] && !found) {
                    while (true) {
                        if (slot >= slots || *(volatile uint32_t*)&s_over) { s_over = 1; break; }
                        unsigned long long old = *(volatile unsigned long long*)&s_tab[slot];
                        if (old == KEY_EMPTY) {
                            old = atomicCAS((unsigned long long*)&s_tab[slot], KEY_EMPTY, key);
                            if (old == KEY_EMPTY) { found = true; break; }
                        }
                        if (old == key) { found = true; break; }
                        slot++;
                    }
                }
                __syncwarp();          // lanes that probed or inserted rejoin the others before the next record
                sl[j] = slot; ok[j] = found;
            }
#pragma unroll
            for (int j = 0; j < MG_UNROLL; j++) {
                if (ok[j]) {
                    const uint32_t smp = (uint32_t)(rr[h + j] >> 2) & sid_mask;
                    atomicOr(&s_sets[sl[j] * wps + (smp >> 3)], (1u << (uint32_t)(rr[h + j] & 3)) << (4 * (smp & 7)));
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
    // 2. occupied slots in slot order (their number is the bucket's row count), then rank inside every run of
    //    adjacent occupied slots (see bucket_count_kernel)
    const uint32_t groups = slots >> 5;
    for (uint32_t g = warp; g < groups; g += NW) {
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, s_tab[g * 32 + lane] != KEY_EMPTY);
        if (lane == 0) s_goff[g] = __popc(m);
    }
    __syncthreads();
    if (warp == 0) {
        uint32_t v[5], sum = 0;                                    // 5 x 32 = 160 groups at most
#pragma unroll
        for (int j = 0; j < 5; j++) { const uint32_t g = lane * 5 + j; v[j] = g < groups ? s_goff[g] : 0u; sum += v[j]; }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
        uint32_t run = inc - sum;
#pragma unroll
        for (int j = 0; j < 5; j++) { const uint32_t g = lane * 5 + j; if (g < groups) s_goff[g] = run; run += v[j]; }
        if (lane == 31) {
            // publish this bucket's count at once: successors only need it to get past us
            const bool ov = *(volatile uint32_t*)&s_over != 0;
            const uint32_t cnt = ov ? 0u : inc;
            s_n = cnt;
            if (ov) atomicOr(&ctl->error, (uint32_t)FM_ERR_TABLE);
            __threadfence();
            atomicExch(&status[bid], (bid == 0 ? ST_PREFIX : ST_AGG) | (unsigned long long)cnt);
        }
    }
    __syncthreads();
    const uint32_t u = s_n;
    for (uint32_t g = warp; g < groups; g += NW) {
        const bool occ = s_tab[g * 32 + lane] != KEY_EMPTY;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, occ);
        if (occ) s_lslot[s_goff[g] + __popc(m & ((1u << lane) - 1u))] = (uint16_t)(g * 32 + lane);
    }
    __syncthreads();
    for (uint32_t i = tid; i < u; i += FM_THREADS) {
        const uint32_t slot = s_lslot[i];
        const uint64_t key = s_tab[slot];
        uint32_t rank = i;
        for (uint32_t j = slot; j > 0 && s_tab[j - 1] != KEY_EMPTY; j--) rank -= (s_tab[j - 1] > key);
        for (uint32_t j = slot + 1; j < slots && s_tab[j] != KEY_EMPTY; j++) rank += (s_tab[j] < key);
        s_order[rank] = (uint16_t)slot;
    }
    // 3. row offset: warp 0 looks back over the predecessors' status words, 32 at a time
    if (warp == 0) {
        unsigned long long excl = 0;
        long long j = (long long)bid - 1;                         // nearest predecessor not yet accounted for
        uint32_t spins = 0;
        bool stalled = false;
        while (j >= 0) {
            const long long idx = j - lane;
            const unsigned long long st = idx >= 0 ? *(volatile unsigned long long*)&status[idx] : ST_PREFIX;   // nothing before bucket 0
            const uint32_t flag = (uint32_t)(st >> 62);
            const uint32_t not_ready = __ballot_sync(0xFFFFFFFFu, flag == 0), prefix = __ballot_sync(0xFFFFFFFFu, flag == 2);
            const int first_nr = not_ready ? __ffs(not_ready) - 1 : 32, first_pf = prefix ? __ffs(prefix) - 1 : 32;
            const int take = first_pf < first_nr ? first_pf + 1 : first_nr;       // lanes 0..take-1 are usable in order
            unsigned long long v = lane < take ? (st & ST_MASK) : 0ULL;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
            excl += v;
            if (first_pf < first_nr) break;                       // reached an inclusive prefix: done
            j -= take;
            if (take == 0) {                                      // nearest predecessor has not published yet
                if (++spins > 20000000u || (*(volatile uint32_t*)&ctl->error & FM_ERR_STALL)) { stalled = true; break; }
                __nanosleep(40);
            } else spins = 0;
        }
        if (lane == 0) {
            if (stalled) atomicOr(&ctl->error, (uint32_t)FM_ERR_STALL);
            if (bid > 0) { __threadfence(); atomicExch(&status[bid], ST_PREFIX | (excl + u)); }
            if (bid + 1 == n_items) ctl->total_rows = excl + u;
            if (excl + u > cap_rows) atomicOr(&ctl->error, (uint32_t)FM_ERR_CAPACITY);
            s_row_base = excl;
        }
    }
    __syncthreads();
    const uint64_t row_base = s_row_base;
    if (u == 0 || row_base + u > cap_rows) return;

    // 4. keys, palindrome flags (in lslot, which is free now)
    for (uint32_t i = tid; i < u; i += FM_THREADS) {
        const uint64_t key = it.key_base | s_tab[s_order[i]];
        out_keys[row_base + i] = key;
        const bool pal = rc && revcomp_key(key, k - 1) == key;
        s_lslot[i] = (uint16_t)(pal ? 1 : 0);
        if (pal) s_anypal = 1;
    }
    __syncthreads();
    uint32_t* s_cnt = reinterpret_cast<uint32_t*>(s_tab);          // the table is no longer needed
    for (uint32_t i = tid; i < u; i += FM_THREADS) s_cnt[i] = 0;
    __syncthreads();
    // 5. rows out: 8 base sets -> 8 letters, one 8-byte store per thread and step
    {
        const int lg = (wps & (wps - 1)) == 0 ? 31 - __clz(wps) : -1;
        uint2* out8 = reinterpret_cast<uint2*>(out_var + row_base * n_samples);
        const uint32_t total = u * wps;
        for (uint32_t idx = tid; idx < total; idx += FM_THREADS) {
            const uint32_t row = lg >= 0 ? idx >> lg : idx / wps;
            const uint32_t w = idx - row * wps;
            const uint32_t x = s_sets[(uint32_t)s_order[row] * wps + w];
            uint2 o;                                               // branch-free; palindromic rows (rare) are redone below
            o.x = iupac_letters4(spread_nibbles4(x & 0xFFFFu));
            o.y = iupac_letters4(spread_nibbles4(x >> 16));
            const uint32_t m = (x | (x >> 1) | (x >> 2) | (x >> 3)) & 0x11111111u;
            if (m) atomicAdd(&s_cnt[row], (uint32_t)__popc(m));
            out8[idx] = o;
        }
        if (s_anypal) {
            __syncthreads();
            for (uint32_t idx = tid; idx < total; idx += FM_THREADS) {
                const uint32_t row = lg >= 0 ? idx >> lg : idx / wps;
                if (!s_lslot[row]) continue;
                const uint32_t x = s_sets[(uint32_t)s_order[row] * wps + (idx - row * wps)];
                out8[idx] = make_uint2(letters4(spread_nibbles4(x & 0xFFFFu), true), letters4(spread_nibbles4(x >> 16), true));
            }
        }
    }
    __syncthreads();
    for (uint32_t i = tid; i < u; i += FM_THREADS) out_cnt[row_base + i] = s_cnt[i];
}

__global__ void sample_items_kernel(const Item* __restrict__ items, uint32_t n_items, uint32_t stride, uint32_t n_out, Item* __restrict__ out) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_out) return;
    Item it = items[min((uint64_t)j * stride, (uint64_t)n_items - 1)];
    it.scratch = NO_SCRATCH;
    out[j] = it;
}

// -> true if `out` holds the merged array; false: nothing was produced, take the two-pass path
static bool merge_items_fused(const uint64_t* rec, const Item* items, uint32_t n_items, int sb, uint64_t n_samples, uint32_t k, bool rc,
                              Array& out) {
    Context& c = ctx();
    // Measured on 128 x 5 Mbp: 7.9 ms for the fused pass against 2.9 + 3.8 ms for count + emit -- reading the records
    // twice costs less than the slot-indexed sets and the look-back; the two passes stay the default.
    int mode = 0;                                                  // 0 never (default), 1 whenever it fits, 2 large builds only
    if (const char* e = getenv("SKA_MERGE_FUSED")) mode = atoi(e);
    if (mode == 0 || n_items == 0 || (n_samples & 7u) || n_samples > 65535) return false;
    if (mode == 2 && n_items < 1024) return false;                // small builds: the estimate pass would cost as much as it saves
    const uint32_t wps = (uint32_t)(n_samples >> 3);
    const size_t per_slot = 8 + 4 * (size_t)wps + 2 + 2;
    uint32_t slots = (uint32_t)std::min<size_t>(4096, FM_SMEM_BUDGET / per_slot) & ~31u;
    if (slots < 256) return false;
    const uint32_t tail = std::max(64u, (slots / 12) & ~31u);
    const uint32_t home_slots = slots - tail, max_keys = home_slots - home_slots / 8;
    const size_t smem = slots * per_slot;
    if (smem + 2048 > c.smem_optin) return false;

    // distinct keys of every `stride`-th bucket -> size of the output, largest bucket
    timer().begin("estimate");
    const uint32_t stride = std::max(1u, n_items / 2048u), ns = (n_items + stride - 1) / stride;
    Item* sitems = arena().take_n<Item>(ns);
    uint32_t* s_rows = arena().take_n<uint32_t>(ns);
    uint32_t* flags = arena().take_n<uint32_t>(2);
    dev_fill(flags, 0, 8);
    SKA_LAUNCH(sample_items_kernel, (ns + 255) / 256, 256, 0, items, n_items, stride, ns, sitems);
    SKA_CUDA(cudaFuncSetAttribute(bucket_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)COUNT_SMEM));
    SKA_LAUNCH(bucket_count_kernel, ns, MG_THREADS, COUNT_SMEM, rec, sitems, sb, (uint32_t)COUNT_LIST - 1, s_rows, flags, (uint64_t*)nullptr);
    std::vector<uint32_t> h_rows(ns);
    SKA_CUDA(cudaMemcpyAsync(h_rows.data(), s_rows, (size_t)ns * 4, cudaMemcpyDeviceToHost, c.stream));
    SKA_CUDA(cudaStreamSynchronize(c.stream));
    timer().end();
    uint64_t sum = 0; uint32_t mx = 0;
    const bool dbg = getenv("SKA_DEBUG_MERGE") != nullptr;
    for (uint32_t v : h_rows) {
        if (v == COUNT_OVERFLOW) { if (dbg) fprintf(stderr, "[skagpu] fused merge declined: a sampled bucket overflowed the count table\n"); return false; }
        sum += v; mx = std::max(mx, v);
    }
    // canonical keys crowd the low prefixes (density 2(1-p)): the sample includes bucket 0, so its maximum is close to the
    // real one; 15 % on top.  Fuller buckets belong to the two-pass path.
    if ((uint64_t)mx + mx * 15 / 100 + 16 > max_keys) {
        if (dbg) fprintf(stderr, "[skagpu] fused merge declined: largest sampled bucket %u keys, table holds %u\n", mx, max_keys);
        return false;
    }
    const double est = (double)sum * ((double)n_items / (double)ns);
    const uint64_t cap_rows = (uint64_t)(est * 1.06) + 8ull * mx + 4096;

    timer().begin("merge");
    out.k = k; out.rc = rc; out.key_bits = 64; out.n_samples = n_samples; out.sorted = true;
    out.keys.alloc(cap_rows); out.variants.alloc(cap_rows * n_samples); out.counts.alloc(cap_rows);
    unsigned long long* status = arena().take_n<unsigned long long>(n_items);
    FusedCtl* ctl = arena().take_n<FusedCtl>(1);
    dev_fill(status, 0, (size_t)n_items * 8);
    dev_fill(ctl, 0, sizeof(FusedCtl));
    SKA_CUDA(cudaFuncSetAttribute(bucket_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SKA_LAUNCH(bucket_merge_kernel, n_items, FM_THREADS, smem, rec, items, n_items, sb, slots, home_slots, max_keys, (uint32_t)n_samples,
               (int)k, rc ? 1 : 0, cap_rows, out.keys.p, out.variants.p, out.counts.p, status, ctl);
    FusedCtl h_ctl;
    SKA_CUDA(cudaMemcpyAsync(&h_ctl, ctl, sizeof(FusedCtl), cudaMemcpyDeviceToHost, c.stream));
    SKA_CUDA(cudaStreamSynchronize(c.stream));
    timer().end();
    if (h_ctl.error) {
        if (dbg) fprintf(stderr, "[skagpu] fused merge gave up (error %u, rows %llu, capacity %llu): two-pass path\n",
                                               h_ctl.error, h_ctl.total_rows, (unsigned long long)cap_rows);
        out.keys.release(); out.variants.release(); out.counts.release();
        return false;
    }
    out.n_rows = out.n_keys = h_ctl.total_rows;
    return true;
}

static int choose_prefix_bits(uint64_t n_pos, int key_bits, uint64_t target) {
    int p = 0;
    while ((n_pos >> p) > target && p < key_bits && p < 24) p++;
    return p;
}

// (2) of the file header: count distinct keys per item (splitting items that do not fit shared memory),
// scan to row offsets, emit rows.  `items` is a device array in ascending key order.
static void merge_items(const uint64_t* rec, Item* items, uint32_t n_items, int sb, uint64_t n_samples, uint32_t k, bool rc, Array& out, uint64_t* scratch_keys) {
    Context& c = ctx();
    if (merge_items_fused(rec, items, n_items, sb, n_samples, k, rc, out)) return;
    // bytes of shared memory per table slot in the emit kernel: tab 8 + rank 2 + (lkey 8 + lslot 2 + cnt 4)/2
    const size_t slot_bytes = 17;
    const int max_log_cap = 13;                                   // 8192 slots = 136 KB
    const uint32_t limit = COUNT_LIST - 1;                        // most distinct keys one item may hold (3071)
    bool all_scratch = scratch_keys != nullptr;                   // every item left its sorted keys for the emit pass

    timer().begin("count");
    SKA_CUDA(cudaFuncSetAttribute(bucket_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)COUNT_SMEM));
    uint32_t* n_rows = arena().take_n<uint32_t>(std::max<uint32_t>(n_items, 1));
    uint32_t* flags = arena().take_n<uint32_t>(2);
    uint32_t h_flags[2] = {0, 0};
    for (int round = 0; n_items > 0; round++) {
        dev_fill(flags, 0, 8);
        SKA_LAUNCH(bucket_count_kernel, n_items, MG_THREADS, COUNT_SMEM, rec, items, sb, limit, n_rows, flags, scratch_keys);
        read_back(h_flags, flags, 8);
        read_back_sync();
        if (!h_flags[0]) break;
        if (round > 40) throw Error("bucket splitting did not converge");
        std::vector<Item> h_items(n_items), next;
        std::vector<uint32_t> h_rows(n_items);
        SKA_CUDA(cudaMemcpyAsync(h_items.data(), items, n_items * sizeof(Item), cudaMemcpyDeviceToHost, c.stream));
        SKA_CUDA(cudaMemcpyAsync(h_rows.data(), n_rows, n_items * 4, cudaMemcpyDeviceToHost, c.stream));
        SKA_CUDA(cudaStreamSynchronize(c.stream));
        next.reserve(n_items + 64);
        for (uint32_t i = 0; i < n_items; i++) {
            const Item& it = h_items[i];
            if (h_rows[i] <= limit) { next.push_back(it); continue; }
            uint64_t width = it.key_hi - it.key_lo;                       // inclusive range is width+1 keys
            if (width == 0) throw Error("internal: single-key item overflow");
            int parts = 16;
            if (h_rows[i] != COUNT_OVERFLOW) { parts = 2; while ((uint64_t)parts * limit < 2ull * h_rows[i]) parts *= 2; }
            uint64_t step = width / parts + 1;                            // ceil((width+1)/parts) keeps children ordered, disjoint
            for (uint64_t lo = it.key_lo;; lo += step) {
                Item ch = it; ch.key_lo = lo; ch.scratch = NO_SCRATCH;    // children share the parent's records: no private scratch
                bool last = (it.key_hi - lo) < step;
                ch.key_hi = last ? it.key_hi : lo + step - 1;
                next.push_back(ch);
                if (last) break;
            }
        }
        n_items = (uint32_t)next.size();
        all_scratch = false;
        items = arena().take_n<Item>(n_items); n_rows = arena().take_n<uint32_t>(n_items);
        SKA_CUDA(cudaMemcpyAsync(items, next.data(), n_items * sizeof(Item), cudaMemcpyHostToDevice, c.stream));
        SKA_CUDA(cudaStreamSynchronize(c.stream));
    }
    uint64_t* row_off = arena().take_n<uint64_t>(n_items + 1);
    uint64_t total_rows = 0;
    if (n_items) {
        launch_scan_counts(n_rows, n_items, row_off, flags + 1);
        read_back(h_flags, flags, 8);
        read_back(&total_rows, row_off + n_items, 8);
        read_back_sync();
    }
    timer().end();
    const uint32_t max_u = h_flags[1];

    timer().begin("emit");
    out.k = k; out.rc = rc; out.key_bits = 64; out.n_samples = n_samples; out.sorted = true;
    out.n_rows = out.n_keys = total_rows;
    out.keys.alloc(total_rows); out.variants.alloc(total_rows * n_samples); out.counts.alloc(total_rows);
    if (total_rows) {
        // table slots per distinct key: at 2 a key sits further than one slot from its home in ~2.5 % of the cases, which sends
        // more than half of the warps' probes into the chain walk (16 % of the kernel's instructions, profiles/); sparser
        // tables cost 10 B of shared memory per slot
        uint32_t cap_factor = 4;
        if (const char* e = getenv("SKA_EMIT_CAP_FACTOR")) cap_factor = (uint32_t)std::max(2, atoi(e));
        int log_cap = 6;
        while ((1u << log_cap) < cap_factor * max_u && log_cap < max_log_cap) log_cap++;
        while ((1u << log_cap) < 2 * max_u) log_cap++;
        if (log_cap > max_log_cap) throw Error("internal: item larger than the shared-memory table");
        const uint32_t stride_w = (uint32_t)((n_samples + 7) / 8) | 1u;    // odd: rows fall on different banks
        const size_t smem_sets = emit_smem_bytes(log_cap, max_u, stride_w);
        const char* force_red = getenv("SKA_EMIT_GLOBAL");                  // tests: force the wide-row path
        if (all_scratch && smem_sets <= (size_t)160 * 1024 && smem_sets + 1024 <= c.smem_optin && !(force_red && atoi(force_red))) {
            SKA_CUDA(cudaFuncSetAttribute(bucket_emit_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sets));
            // Two size classes.  Canonical keys crowd the low prefixes (density 2(1-p) over the key range), so the largest
            // bucket has ~2-3x the rows of the median one, and shared memory sized for it leaves 2 CTAs per SM.  Buckets
            // of up to `u_small` rows (<= 48 KB: 4 CTAs per SM) go first in their own launch; the other CTAs of each
            // launch return at once.
            uint32_t u_small = 0;
            int log_cap_small = log_cap;
            size_t small_bytes = (size_t)48 * 1024;
            if (const char* e = getenv("SKA_EMIT_SMALL_BYTES")) small_bytes = (size_t)std::max(0, atoi(e));   // tests: classes at small sizes
            if (smem_sets > small_bytes + small_bytes / 6 && !(getenv("SKA_EMIT_ONE_CLASS") && atoi(getenv("SKA_EMIT_ONE_CLASS")))) {
                for (int lc = 6; lc <= log_cap; lc++) {
                    const uint32_t uu = (1u << lc) / cap_factor;               // table at most 1 / cap_factor full
                    if (uu < max_u && emit_smem_bytes(lc, uu, stride_w) <= small_bytes) { u_small = uu; log_cap_small = lc; }
                }
            }
            if (u_small) {
                const size_t smem_small = emit_smem_bytes(log_cap_small, u_small, stride_w);
                SKA_LAUNCH(bucket_emit_smem_kernel, n_items, MG_THREADS, smem_small, rec, items, sb, row_off, log_cap_small, u_small,
                           (uint32_t)n_samples, stride_w, (int)k, rc ? 1 : 0, out.keys.p, out.variants.p, out.counts.p,
                           (const uint64_t*)scratch_keys, 0u);
            }
            SKA_LAUNCH(bucket_emit_smem_kernel, n_items, MG_THREADS, smem_sets, rec, items, sb, row_off, log_cap, max_u,
                       (uint32_t)n_samples, stride_w, (int)k, rc ? 1 : 0, out.keys.p, out.variants.p, out.counts.p,
                       (const uint64_t*)scratch_keys, u_small);
        } else {
            size_t smem = ((size_t)1 << log_cap) * slot_bytes;
            SKA_CUDA(cudaFuncSetAttribute(bucket_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            SKA_LAUNCH(bucket_emit_kernel, n_items, MG_THREADS, smem, rec, items, sb, row_off, log_cap,
                       (uint32_t)n_samples, (int)k, rc ? 1 : 0, out.keys.p, out.variants.p, out.counts.p, (const uint64_t*)scratch_keys);
        }
    }
    timer().end();
}

// ---- union of sorted key runs (the owner side of the partial-row exchange, rowmerge.cu) ---------------------------
// Every run is ascending, so a key-range bucket is one contiguous piece of every run: its ends are found by binary
// search (no partition pass), the pieces are copied next to each other as packed records and the count pass above
// leaves every bucket's distinct keys sorted; a last copy closes the gaps.  Buckets are cut so that none holds more
// records than the count table can take, hence no splitting.
__global__ void run_cumulative_kernel(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ run_off, int n_runs, int shift,
                                      uint32_t n_cuts, uint64_t* __restrict__ cum /* [n_runs][n_cuts] */) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    const int w = blockIdx.y;
    if (b >= n_cuts || w >= n_runs) return;
    const uint64_t* a = keys + run_off[w];
    const uint64_t n = run_off[w + 1] - run_off[w];
    uint64_t lo = 0, hi = n;
    if (b + 1 == n_cuts) lo = n;
    else {
        const uint64_t cut = (uint64_t)b << shift;
        while (lo < hi) { uint64_t m = (lo + hi) >> 1; if (a[m] < cut) lo = m + 1; else hi = m; }
    }
    cum[(size_t)w * n_cuts + b] = lo;
}
__global__ void bucket_sizes_kernel(const uint64_t* __restrict__ cum, int n_runs, uint32_t n_buckets, uint32_t* __restrict__ sizes) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_buckets) return;
    uint64_t t = 0;
    for (int w = 0; w < n_runs; w++) t += cum[(size_t)w * (n_buckets + 1) + b + 1] - cum[(size_t)w * (n_buckets + 1) + b];
    sizes[b] = (uint32_t)min(t, (uint64_t)0xFFFFFFFFu);
}
// one warp per bucket: the runs' pieces -> packed records (key << 3: one sample, sb = 1, no middle base)
__global__ void __launch_bounds__(256)
gather_buckets_kernel(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ run_off, const uint64_t* __restrict__ cum, int n_runs,
                      uint32_t n_buckets, const uint64_t* __restrict__ off, uint64_t* __restrict__ prec) {
    const uint32_t b = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= n_buckets) return;
    uint64_t dst = off[b];
    for (int w = 0; w < n_runs; w++) {
        const uint64_t beg = cum[(size_t)w * (n_buckets + 1) + b], end = cum[(size_t)w * (n_buckets + 1) + b + 1];
        const uint64_t* src = keys + run_off[w];
        for (uint64_t i = beg + lane; i < end; i += 32) prec[dst + (i - beg)] = src[i] << 3;
        dst += end - beg;
    }
}
__global__ void __launch_bounds__(128)
compact_keys_kernel(const uint64_t* __restrict__ scratch_keys, const Item* __restrict__ items, const uint64_t* __restrict__ row_off,
                    uint64_t* __restrict__ out_keys) {
    const Item it = items[blockIdx.x];
    const uint64_t r0 = row_off[blockIdx.x];
    const uint32_t u = (uint32_t)(row_off[blockIdx.x + 1] - r0);
    for (uint32_t i = threadIdx.x; i < u; i += 128) out_keys[r0 + i] = scratch_keys[it.scratch + i];
}

bool union_sorted_runs(const uint64_t* keys, int n_runs, const uint64_t* h_run_off, uint32_t k, DevBuf<uint64_t>& out_keys, uint64_t& n_out) {
    Context& c = ctx();
    const int key_bits = 2 * ((int)k - 1);
    const uint64_t n_in = h_run_off[n_runs];
    n_out = 0;
    if (n_in == 0) return true;
    const int max_bits = std::min(key_bits, 24);
    int pbits = 0;
    while ((n_in >> pbits) > 1024 && pbits < max_bits) pbits++;           // ~1 K records per bucket: room for a 3x crowded key range
    uint64_t* d_run_off = arena().take_n<uint64_t>(n_runs + 1);
    SKA_CUDA(cudaMemcpyAsync(d_run_off, h_run_off, (size_t)(n_runs + 1) * 8, cudaMemcpyHostToDevice, c.stream));
    uint32_t* flags = arena().take_n<uint32_t>(2);
    uint32_t h_flags[2] = {0, 0};
    uint64_t *cum = nullptr, *off = nullptr;
    uint32_t NB = 0;
    while (true) {
        NB = 1u << pbits;
        cum = arena().take_n<uint64_t>((size_t)n_runs * (NB + 1));
        uint32_t* sizes = arena().take_n<uint32_t>(NB);
        off = arena().take_n<uint64_t>(NB + 1);
        dim3 grid((NB + 1 + 255) / 256, n_runs);
        SKA_LAUNCH(run_cumulative_kernel, grid, 256, 0, keys, d_run_off, n_runs, key_bits - pbits, NB + 1, cum);
        SKA_LAUNCH(bucket_sizes_kernel, (NB + 255) / 256, 256, 0, cum, n_runs, NB, sizes);
        launch_scan_counts(sizes, NB, off, flags + 1);
        read_back(h_flags, flags, 8);
        read_back_sync();
        if (h_flags[1] <= (uint32_t)COUNT_LIST - 1) break;               // every bucket fits the count table
        if (pbits >= max_bits) return false;                               // a key range this crowded: let the general merge split it
        pbits = std::min(max_bits, pbits + 2);
    }
    uint64_t* prec = arena().take_n<uint64_t>(n_in);
    uint64_t* scratch = arena().take_n<uint64_t>(n_in);
    SKA_LAUNCH(gather_buckets_kernel, (NB + 7) / 8, 256, 0, keys, d_run_off, cum, n_runs, NB, off, prec);
    Item* items = arena().take_n<Item>(NB);
    SKA_LAUNCH(make_items_kernel, (NB + 255) / 256, 256, 0, off, NB, (uint64_t)0, key_bits, pbits, pbits, key_bits, items);
    uint32_t* n_rows = arena().take_n<uint32_t>(NB);
    SKA_CUDA(cudaFuncSetAttribute(bucket_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)COUNT_SMEM));
    dev_fill(flags, 0, 8);
    SKA_LAUNCH(bucket_count_kernel, NB, MG_THREADS, COUNT_SMEM, prec, items, 1, (uint32_t)COUNT_LIST - 1, n_rows, flags, scratch);
    uint64_t* row_off = arena().take_n<uint64_t>(NB + 1);
    launch_scan_counts(n_rows, NB, row_off, flags + 1);
    uint64_t total = 0;
    read_back(h_flags, flags, 8);
    read_back(&total, row_off + NB, 8);
    read_back_sync();
    if (h_flags[0]) return false;                                          // probing ran off a table's tail (very crowded corner)
    out_keys.alloc(total);
    if (total) SKA_LAUNCH(compact_keys_kernel, NB, 128, 0, (const uint64_t*)scratch, items, row_off, out_keys.p);
    n_out = total;
    return true;
}

int merge_prefix_bits(uint64_t n_records_total, uint32_t k, uint64_t n_samples_total) {
    const int key_bits = 2 * ((int)k - 1);
    (void)n_samples_total;
    // ~8-16 K records per bucket whatever the number of samples (measured, profiles/README.md "wide rows": with 1024
    // columns the same bucket size keeps the rows a CTA ORs into L2-resident; larger buckets thrash, smaller ones
    // cost more in the partition).  Past 2^16 buckets a partition pass would need more than 8 bits per digit, which
    // costs more than it saves: let buckets grow to 32 K records first.
    int P = choose_prefix_bits(n_records_total, key_bits, 16384);
    if (P > 16) P = std::max(16, choose_prefix_bits(n_records_total, key_bits, 32768));
    if (const char* e = getenv("SKA_PREFIX_BITS")) { int v = atoi(e); if (v >= 0 && v <= key_bits && v <= 22) P = v; }   // tuning/tests
    return std::min(std::min(P, 22), key_bits);
}

void build_array_from_records(uint64_t* rec, uint64_t n_rec, const uint64_t* d_sample_off, uint32_t n_samples_local, uint32_t sid_base,
                              uint64_t n_samples_total, uint32_t k, bool rc, Array& out, int prefix_bits, const unsigned long long* k1_hist) {
    const PartPlan pl = make_part_plan(k, prefix_bits >= 0 ? prefix_bits : merge_prefix_bits(n_rec, k, n_samples_total), n_samples_total);
    const uint32_t n_buckets = 1u << pl.P;

    // (1) partition by key prefix (stages "partition1", "partition2")
    uint64_t* prec = nullptr; uint64_t* off = nullptr; uint64_t* part_a = nullptr;
    partition_records(rec, n_rec, d_sample_off, n_samples_local, sid_base, pl, rec, &prec, &off, &part_a, k1_hist);

    timer().begin("bucket_offsets");
    Item* items = arena().take_n<Item>(n_buckets);
    SKA_LAUNCH(make_items_kernel, (n_buckets + 255) / 256, 256, 0, off, n_buckets, (uint64_t)0, pl.key_bits, pl.P, pl.P2, pl.low_bits, items);
    timer().end();

    // the buffer the partition no longer needs doubles as the count -> emit key scratch (same length as the records)
    uint64_t* scratch = (prec == rec) ? part_a : rec;
    merge_items(prec, items, n_buckets, pl.sb, n_samples_total, k, rc, out, scratch);
}

// Row shard: the received records are `n_runs` runs (one per source rank) of first-pass output, each holding
// its pieces of the first-level regions [region_lo, region_hi) in order.  run_off[w * (nr + 1) + j] is the offset,
// inside the receive buffer, of region (region_lo + j) of run w (nr = region_hi - region_lo).  The second
// partition pass runs here, over all runs at once, so the merge below sees one run.  `prec` is overwritten
// (it becomes the count -> emit key scratch).
void build_array_from_runs(uint64_t* prec, uint64_t n_rec, int n_runs, const uint64_t* run_off, uint32_t region_lo, uint32_t region_hi,
                           const PartPlan& pl, uint64_t n_samples, uint32_t k, bool rc, Array& out) {
    if (n_runs < 1) throw Error("row shard needs at least one run");
    const uint32_t nr = region_hi - region_lo;
    const uint32_t nb = nr << pl.P2;
    uint64_t* part = arena().take_n<uint64_t>(std::max<uint64_t>(n_rec, 1));
    uint64_t* off = nullptr;
    partition_runs(prec, n_runs, run_off, nr, pl, part, &off);
    timer().begin("bucket_offsets");
    Item* items = arena().take_n<Item>(std::max<uint32_t>(nb, 1));
    if (nb)
        SKA_LAUNCH(make_items_kernel, (nb + 255) / 256, 256, 0, off, nb, (uint64_t)region_lo << pl.P2, pl.key_bits, pl.P, pl.P2, pl.low_bits,
                   items);
    timer().end();
    merge_items(part, items, nb, pl.sb, n_samples, k, rc, out, prec);
}

}  // namespace ska
