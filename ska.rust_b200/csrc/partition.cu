// partition.cu -- groups K1's records by the top P bits of the split-k-mer key: the step that lets one
// CTA merge one key range in shared memory (merge.cu).  Hand-written replacement for the library radix
// sort used earlier in this round (DESIGN.md "partition").
//
// Two MSD passes of P1 and P2 bits (any widths up to 11), each pass = histogram + scatter.  On several GPUs
// the exchange sits BETWEEN the passes: a rank runs pass 1 on its own samples, ships every first-level
// region to the rank that owns it, and pass 2 runs at the owner over the regions' pieces from all source
// ranks ("segments") -- so the second pass also concatenates the runs and the merge sees one run.
//   * records move as ONE packed u64: after the first digit is implied by the bucket, its bits hold the
//     sample id:   prec = key_low << (SB+2) | sample_id << 2 | middle_code,
//     key_low = key without its top P1 bits, SB = bits of a sample id.  16 B of traffic per record and pass
//     instead of 20 B for (u64 record, u16 sample id) pairs;
//   * the partition is UNSTABLE on purpose: order inside a bucket is irrelevant to the merge (a set
//     union), so a tile reserves its slice of every destination bucket with one global atomicAdd per
//     digit instead of a decoupled look-back;
//   * ranking inside a tile is one shared-memory atomicAdd per record on a per-digit counter (measured
//     faster on B200 than ballot-matching peers: 10.5 ms vs 17.5 ms for the whole partition of 6.4e8
//     records); records are staged in shared memory in digit order so the global stores are coalesced runs.
// Tried and dropped (profiles/README.md): other tile shapes (256 x 8, 256 x 12, 512 x 4 records: equal or slower --
// shorter runs per digit cost more than the extra CTAs per SM gain) and a persistent form that keeps the next tile's
// loads in flight (4.7 vs 4.6 ms per pass: the kernel does not wait on its loads).
// Roofline: HBM; 8 B read by each histogram, 8 B read + 8 B written by each scatter, per record.
#include "common.cuh"

namespace ska {

#ifndef SKA_PT_THREADS
#define SKA_PT_THREADS 512
#endif
constexpr int PT_THREADS = SKA_PT_THREADS;
#ifndef SKA_PT_ITEMS
#define SKA_PT_ITEMS 8
#endif
constexpr int PT_ITEMS = SKA_PT_ITEMS;
constexpr int PT_TILE = PT_THREADS * PT_ITEMS;      // records per tile
constexpr int PT_WARPS = PT_THREADS / 32;
constexpr int PT_MAX_BITS = 11;
// scatter CTA shape (same tile): threads x records per thread
#ifndef SKA_SC_THREADS
#define SKA_SC_THREADS 256
#endif
#ifndef SKA_SC_MIN_CTAS
#define SKA_SC_MIN_CTAS 4
#endif
constexpr int SC_T = SKA_SC_THREADS;
constexpr int SC_I = PT_TILE / SC_T;
constexpr int SC_MIN_CTAS = SKA_SC_MIN_CTAS;
static_assert(SC_T * SC_I == PT_TILE && PT_TILE <= 65536, "scatter tile shape");
constexpr uint32_t PT_NONE = 0xFFFFFFFFu;

// ---------------------------------------------------------------------------------------------- histograms
// src: K1 records (pass 1) or packed records (pass 2).  For pass 2 tiles are enumerated per SEGMENT (a contiguous
// piece of one first-level region) so that a tile never straddles two regions: tile_pref[s] = tiles before segment s.
// On one GPU a segment is a whole region (region == NULL: region = segment index).
struct DigitFn {
    int pass;            // 1: digit = rec >> shift (K1 record, REC_SENTINEL skipped); 2: digit = (prec >> shift) & mask
    int shift;
    uint32_t mask;
    __device__ __forceinline__ uint32_t operator()(uint64_t r) const {
        if (pass == 1) return r == REC_SENTINEL ? PT_NONE : (uint32_t)(r >> shift);
        return mask ? (uint32_t)(r >> shift) & mask : 0u;
    }
};
struct Segs { const uint64_t* tile_pref; const uint64_t* beg; const uint64_t* end; const uint32_t* region; uint32_t n; };

// consecutive tiles of one CTA: the segment index only moves forward, so after the first tile no search is needed
__device__ __forceinline__ void tile_range_next(uint64_t tile, const Segs& sg, uint64_t n_total, uint32_t& seg, uint64_t& beg, uint64_t& end, uint32_t& region) {
    if (!sg.tile_pref) { beg = tile * PT_TILE; end = min(n_total, beg + PT_TILE); region = 0; return; }
    if (seg == PT_NONE) {
        uint32_t lo = 0, hi = sg.n;
        while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (sg.tile_pref[m] <= tile) lo = m; else hi = m; }
        seg = lo;
    }
    while (seg + 1 < sg.n && sg.tile_pref[seg + 1] <= tile) seg++;
    region = sg.region ? sg.region[seg] : seg;
    beg = sg.beg[seg] + (tile - sg.tile_pref[seg]) * PT_TILE;
    end = min(sg.end[seg], beg + PT_TILE);
}
__device__ __forceinline__ void tile_range(uint64_t tile, const Segs& sg, uint64_t n_total, uint64_t& beg, uint64_t& end, uint32_t& region) {
    if (!sg.tile_pref) { beg = tile * PT_TILE; end = min(n_total, beg + PT_TILE); region = 0; return; }
    uint32_t lo = 0, hi = sg.n;                             // last segment with tile_pref[segment] <= tile
    while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (sg.tile_pref[m] <= tile) lo = m; else hi = m; }
    region = sg.region ? sg.region[lo] : lo;
    beg = sg.beg[lo] + (tile - sg.tile_pref[lo]) * PT_TILE;
    end = min(sg.end[lo], beg + PT_TILE);
}

// The tile's range found by the whole CTA: every thread looks at one segment, so the lookup is one round of loads (a binary
// search is log2(segments) dependent loads in front of the tile's own loads).  Contains a barrier: call it from every thread.
template <int THREADS>
__device__ __forceinline__ void tile_range_coop(uint64_t tile, const Segs& sg, uint64_t n_total, uint64_t& beg, uint64_t& end, uint32_t& region) {
    if (!sg.tile_pref) { beg = tile * PT_TILE; end = min(n_total, beg + PT_TILE); region = 0; return; }
    __shared__ uint64_t s_tr[2];
    __shared__ uint32_t s_trr;
    for (uint32_t m = threadIdx.x; m < sg.n; m += THREADS) {
        const uint64_t p0 = sg.tile_pref[m];
        if (p0 <= tile && tile < sg.tile_pref[m + 1]) {
            const uint64_t b0 = sg.beg[m] + (tile - p0) * PT_TILE;
            s_tr[0] = b0; s_tr[1] = min(sg.end[m], b0 + PT_TILE);
            s_trr = sg.region ? sg.region[m] : m;
        }
    }
    __syncthreads();
    beg = s_tr[0]; end = s_tr[1]; region = s_trr;
}

#ifndef SKA_HIST_MIN_CTAS
#define SKA_HIST_MIN_CTAS 2
#endif
#ifndef SKA_HIST_PREFETCH
#define SKA_HIST_PREFETCH 1
#endif
__global__ void __launch_bounds__(PT_THREADS, SKA_HIST_MIN_CTAS)
part_hist_kernel(const uint64_t* __restrict__ src, uint64_t n_total, uint64_t n_tiles, Segs sg, DigitFn fn, int bits,
                 unsigned long long* __restrict__ hist /* [n_regions or 1][1 << bits] */) {
    extern __shared__ uint32_t s_cnt[];                     // [D]
    const uint32_t D = 1u << bits;
    // a CTA takes a contiguous chunk of tiles; with regions it flushes whenever the region changes
    const uint64_t per = (n_tiles + gridDim.x - 1) / gridDim.x;
    const uint64_t t_beg = min(n_tiles, (uint64_t)blockIdx.x * per), t_end = min(n_tiles, t_beg + per);
    uint32_t cur_region = PT_NONE;
    for (uint32_t i = threadIdx.x; i < D; i += PT_THREADS) s_cnt[i] = 0;
    __syncthreads();
    auto flush = [&](uint32_t region) {
        __syncthreads();
        for (uint32_t d = threadIdx.x; d < D; d += PT_THREADS) {
            uint32_t tot = s_cnt[d];
            s_cnt[d] = 0;
            if (tot) atomicAdd(&hist[(size_t)region * D + d], (unsigned long long)tot);
        }
        __syncthreads();
    };
#if SKA_HIST_PREFETCH
    // the next tile's loads are issued before this tile's records are counted (a warp issues in order: without this the
    // loads of tile t + 1 wait behind the atomics of tile t, which wait for the loads of tile t)
    uint64_t nx[PT_ITEMS];
    uint64_t nbeg = 0, nend = 0; uint32_t nregion = 0, seg = PT_NONE;
    if (t_beg < t_end) {
        tile_range_next(t_beg, sg, n_total, seg, nbeg, nend, nregion);
#pragma unroll
        for (int j = 0; j < PT_ITEMS; j++) {
            const uint64_t i = nbeg + (uint64_t)j * PT_THREADS + threadIdx.x;
            nx[j] = i < nend ? __ldg(src + i) : REC_SENTINEL;
        }
    }
    for (uint64_t t = t_beg; t < t_end; t++) {
        const uint64_t beg = nbeg, end = nend; const uint32_t region = nregion;
        uint64_t rr[PT_ITEMS];
#pragma unroll
        for (int j = 0; j < PT_ITEMS; j++) rr[j] = nx[j];
        if (t + 1 < t_end) {
            tile_range_next(t + 1, sg, n_total, seg, nbeg, nend, nregion);
#pragma unroll
            for (int j = 0; j < PT_ITEMS; j++) {
                const uint64_t i = nbeg + (uint64_t)j * PT_THREADS + threadIdx.x;
                nx[j] = i < nend ? __ldg(src + i) : REC_SENTINEL;
            }
        }
        if (region != cur_region) { if (cur_region != PT_NONE) flush(cur_region); cur_region = region; }
#pragma unroll
        for (int j = 0; j < PT_ITEMS; j++) {
            const uint64_t i = beg + (uint64_t)j * PT_THREADS + threadIdx.x;
            const uint32_t d = i < end ? fn(rr[j]) : PT_NONE;
            if (d != PT_NONE) atomicAdd(&s_cnt[d], 1u);
        }
    }
#else
    for (uint64_t t = t_beg; t < t_end; t++) {
        uint64_t beg, end; uint32_t region;
        tile_range(t, sg, n_total, beg, end, region);
        if (region != cur_region) { if (cur_region != PT_NONE) flush(cur_region); cur_region = region; }
        uint64_t rr[PT_ITEMS];
#pragma unroll
        for (int j = 0; j < PT_ITEMS; j++) {
            uint64_t i = beg + (uint64_t)j * PT_THREADS + threadIdx.x;
            rr[j] = i < end ? src[i] : REC_SENTINEL;
        }
#pragma unroll
        for (int j = 0; j < PT_ITEMS; j++) {
            uint64_t i = beg + (uint64_t)j * PT_THREADS + threadIdx.x;
            uint32_t d = i < end ? fn(rr[j]) : PT_NONE;
            if (d != PT_NONE) atomicAdd(&s_cnt[d], 1u);
        }
    }
#endif
    if (cur_region != PT_NONE) flush(cur_region);
}

// The second pass's histogram from the digit bytes the first scatter left beside its records (one byte per record instead
// of the 8-byte record: 0.64 GB instead of 5.1 GB for 128 x 5 Mbp).  blockIdx.y = region, the CTAs of a region take its 16 KB
// chunks in turn; a thread has four aligned 16-byte loads in flight, bytes outside the region are skipped at its two ends.
constexpr int HB_THREADS = 256, HB_WORDS = 4;
__global__ void __launch_bounds__(HB_THREADS)
part_hist_bytes_kernel(const uint8_t* __restrict__ dig, const uint64_t* __restrict__ region_off, int bits, unsigned long long* __restrict__ hist) {
    extern __shared__ uint32_t s_cnt[];
    const uint32_t D = 1u << bits;
    const uint32_t region = blockIdx.y;
    const uint64_t beg = region_off[region], end = region_off[region + 1];
    if (beg >= end) return;
    for (uint32_t i = threadIdx.x; i < D; i += HB_THREADS) s_cnt[i] = 0;
    __syncthreads();
    const uint4* __restrict__ w4 = reinterpret_cast<const uint4*>(dig);
    const uint64_t w_beg = beg >> 4, w_end = (end + 15) >> 4;                     // 16-byte words that touch the region
    constexpr uint64_t CH = (uint64_t)HB_THREADS * HB_WORDS;
    for (uint64_t w0 = w_beg + (uint64_t)blockIdx.x * CH; w0 < w_end; w0 += (uint64_t)gridDim.x * CH) {
        uint4 v[HB_WORDS];
#pragma unroll
        for (int j = 0; j < HB_WORDS; j++) {
            const uint64_t w = w0 + (uint64_t)j * HB_THREADS + threadIdx.x;
            v[j] = w < w_end ? __ldg(w4 + w) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int j = 0; j < HB_WORDS; j++) {
            const uint64_t w = w0 + (uint64_t)j * HB_THREADS + threadIdx.x;
            if (w >= w_end) continue;
            const uint64_t p = w << 4;
            const uint32_t x[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
            if (p >= beg && p + 16 <= end) {
#pragma unroll
                for (int b = 0; b < 16; b++) atomicAdd(&s_cnt[(x[b >> 2] >> (8 * (b & 3))) & 0xFFu], 1u);
            } else {
#pragma unroll
                for (int b = 0; b < 16; b++)
                    if (p + b >= beg && p + b < end) atomicAdd(&s_cnt[(x[b >> 2] >> (8 * (b & 3))) & 0xFFu], 1u);
            }
        }
    }
    __syncthreads();
    for (uint32_t d = threadIdx.x; d < D; d += HB_THREADS) {
        const uint32_t tot = s_cnt[d];
        if (tot) atomicAdd(&hist[(size_t)region * D + d], (unsigned long long)tot);
    }
}

// ---------------------------------------------------------------------------------------------- scatter
struct PackFn {               // pass 1: K1 record -> packed record (needs the sample id of the position)
    int sb;                   // bits of a sample id
    uint64_t low_mask;        // key bits kept (key without its top P1 bits)
    uint32_t sid_base;
};

// Scatter of one tile of PT_TILE records.  SC_T threads x SC_I records each (256 x 16: four CTAs per SM at 64 registers;
// measured on B200 with profiles/microbench/scatter_mb2.cu: 2.29 ms for 6.4e8 records against 2.72 ms for 512 x 8 and
// 3.7 ms for the round-1 kernel, which zeroed its counters and waited at a barrier BEFORE issuing its loads).
// BITS > 0: digit width fixed at compile time (the 8 + 8 split of a 128-sample build); BITS == 0: run-time width.
template <int PASS, int BITS>
__global__ void __launch_bounds__(SC_T, SC_MIN_CTAS)
part_scatter_kernel(const uint64_t* __restrict__ src, uint64_t n_total, Segs sg, DigitFn fn, int bits_rt, PackFn pk,
                    const uint64_t* __restrict__ sample_off, uint32_t n_samples,
                    unsigned long long* __restrict__ cursor /* [n_regions or 1][1 << bits], preset to bucket starts */,
                    uint64_t* __restrict__ dst) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int bits = BITS ? BITS : bits_rt;
    const uint32_t D = 1u << bits;
    uint64_t* s_stage = reinterpret_cast<uint64_t*>(smem);                       // [PT_TILE]
    unsigned long long* s_gbase = reinterpret_cast<unsigned long long*>(s_stage + PT_TILE);   // [D]
    uint32_t* s_doff = reinterpret_cast<uint32_t*>(s_gbase + D);                 // [D + 1]
    uint32_t* s_acnt = s_doff + D + 1;                                           // [D] per-digit counters
    uint16_t* s_sdig = reinterpret_cast<uint16_t*>(s_acnt + D);                  // [PT_TILE]
    __shared__ uint32_t s_sid[2];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t beg, end; uint32_t region;
    tile_range(blockIdx.x, sg, n_total, beg, end, region);
    // 1. the tile's records: in flight while the counters are cleared
    uint64_t rr[SC_I]; uint32_t dp[SC_I];
    const uint32_t cnt = (uint32_t)(end - beg);
    const uint64_t* __restrict__ sp = src + beg;
#pragma unroll
    for (int j = 0; j < SC_I; j++) {
        const uint32_t li = j * SC_T + threadIdx.x;
        rr[j] = li < cnt ? __ldg(sp + li) : REC_SENTINEL;
    }
    for (uint32_t i = threadIdx.x; i < D; i += SC_T) s_acnt[i] = 0;
    if (PASS == 1 && threadIdx.x == 0) {                     // sample ids of the tile's first and last position
        uint64_t pf = beg, pl = end - 1;
        uint32_t lo = 0, hi = n_samples;
        while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (sample_off[m] <= pf) lo = m; else hi = m; }
        s_sid[0] = lo; hi = n_samples;
        while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (sample_off[m] <= pl) lo = m; else hi = m; }
        s_sid[1] = lo;
    }
    __syncthreads();

    // 2. digit, packed record, rank among the tile's records of the digit (dp = digit << 16 | rank; PT_TILE <= 65536)
#pragma unroll
    for (int j = 0; j < SC_I; j++) {
        const uint32_t li = j * SC_T + threadIdx.x;
        const uint64_t r = rr[j];
        // pass 1: padding carries the sentinel already; pass 2: packed records have no sentinel
        const uint32_t d = PASS == 1 ? (r == REC_SENTINEL ? PT_NONE : (uint32_t)(r >> fn.shift))
                                     : (li < cnt ? (fn.mask ? (uint32_t)(r >> fn.shift) & fn.mask : 0u) : PT_NONE);
        if (PASS == 1) {
            uint32_t s = s_sid[0];
            if (d != PT_NONE && s_sid[0] != s_sid[1]) {      // tile straddles samples (rare)
                const uint64_t i = beg + li;
                uint32_t lo = s_sid[0], hi = s_sid[1] + 1;
                while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (sample_off[m] <= i) lo = m; else hi = m; }
                s = lo;
            }
            const uint64_t key = r >> 4;
            const uint32_t mid = (uint32_t)__ffs((uint32_t)(r & 0xF)) - 1u;
            rr[j] = ((key & pk.low_mask) << (pk.sb + 2)) | ((uint64_t)(s + pk.sid_base) << 2) | mid;
        }
        dp[j] = d != PT_NONE ? ((d << 16) | atomicAdd(&s_acnt[d], 1u)) : PT_NONE;
    }
    __syncthreads();

    // 3. exclusive scan over digits (warp 0: 32 lanes x D/32 consecutive digits), global reservation
    if (warp == 0) {
        if (BITS) {
            constexpr int per = BITS ? (1 << BITS) / 32 : 1;
            const int b0 = lane * per;
            uint32_t v[per], sum = 0;
#pragma unroll
            for (int q = 0; q < per; q++) { v[q] = s_acnt[b0 + q]; sum += v[q]; }
            uint32_t inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
            uint32_t run = inc - sum;
#pragma unroll
            for (int q = 0; q < per; q++) { s_doff[b0 + q] = run; run += v[q]; }
            if (lane == 31) s_doff[D] = inc;
        } else {
            const uint32_t per = (D + 31) / 32, b0 = lane * per, e0 = min(D, b0 + per);
            uint32_t sum = 0;
            for (uint32_t d = b0; d < e0; d++) sum += s_acnt[d];
            uint32_t inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += v; }
            uint32_t run = inc - sum;
            for (uint32_t d = b0; d < e0; d++) { uint32_t t = s_acnt[d]; s_doff[d] = run; run += t; }
            if (lane == 31) s_doff[D] = inc;
        }
    }
    __syncthreads();
    for (uint32_t d = threadIdx.x; d < D; d += SC_T) {
        uint32_t tot = s_doff[d + 1] - s_doff[d];
        if (tot) s_gbase[d] = atomicAdd(&cursor[(size_t)region * D + d], (unsigned long long)tot) - s_doff[d];
    }
    // 4. stage in digit order
#pragma unroll
    for (int j = 0; j < SC_I; j++) {
        if (dp[j] == PT_NONE) continue;
        const uint32_t d = dp[j] >> 16, slot = s_doff[d] + (dp[j] & 0xFFFFu);
        s_stage[slot] = rr[j];
        s_sdig[slot] = (uint16_t)d;
    }
    __syncthreads();
    // 5. coalesced runs out
    const uint32_t n_valid = s_doff[D];
    for (uint32_t slot = threadIdx.x; slot < n_valid; slot += SC_T) dst[s_gbase[s_sdig[slot]] + slot] = s_stage[slot];
}

// Second form of the same scatter (default; SKA_SCATTER_V1=1 selects the one above).  What the ncu capture of the first
// form showed (profiles/kernels_r02_raw.csv): 0.35 instructions per cycle and scheduler, the warps waiting on shared-memory
// results (short scoreboard 6.7 + MIO throttle 3.2 of 22.6 cycles per instruction) and on the global atomics.  Here
//   * no digit array: the digit of a staged record is recomputed from the record when it leaves the CTA (pass 1 stages the
//     K1 record and repacks it on the way out), which removes one 2-byte store and one 2-byte load per record;
//   * with one thread per digit (D == SC_T) the reservation atomicAdd is issued before the scan and its result is used
//     after the staging stores, so its round trip to L2 is hidden; the scan runs on all warps (no 8-way bank conflicts);
//   * a pass-1 tile that straddles samples (one in a thousand) is scattered in one round per sample, so the sample id
//     is uniform in a round and needs no per-record lookup.
// Shared memory 36 KB per CTA instead of 45.
#ifndef SKA_SC_LOAD
#define SKA_SC_LOAD __ldg
#endif
// STR (pass 1 only): 0 = every tile that lies inside one sample (tiles that straddle samples return at once),
// 1 = the straddling tiles: CTA m - 1 looks at the boundary sample_off[m] and takes its tile if it is the first boundary
// strictly inside that tile.  Two launches, so that the common form carries no per-sample loop.
template <int PASS, int BITS, int STR>
__global__ void __launch_bounds__(SC_T, SC_MIN_CTAS)
part_scatter2_kernel(const uint64_t* __restrict__ src, uint64_t n_total, Segs sg, DigitFn fn, int bits_rt, PackFn pk,
                     const uint64_t* __restrict__ sample_off, uint32_t n_samples,
                     unsigned long long* __restrict__ cursor, uint64_t* __restrict__ dst,
                     uint8_t* __restrict__ next_digit /* pass 1, may be null: the record's pass-2 digit, same order as dst */,
                     int nd_shift, uint32_t nd_mask) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int bits = BITS ? BITS : bits_rt;
    const uint32_t D = 1u << bits;
    constexpr bool PER_THREAD = BITS >= 3 && (1 << BITS) <= SC_T;        // one thread per digit (threads beyond D idle in phase B)
    uint64_t* s_stage = reinterpret_cast<uint64_t*>(smem);                       // [PT_TILE]
    unsigned long long* s_gbase = reinterpret_cast<unsigned long long*>(s_stage + PT_TILE);   // [D]
    uint32_t* s_acnt = reinterpret_cast<uint32_t*>(s_gbase + D);                 // [D] records per digit (16-byte aligned)
    uint32_t* s_cur = s_acnt + D;                                                // [D] first slot of the digit, then its cursor
    __shared__ uint32_t s_sid[2], s_total;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t beg, end; uint32_t region = 0;
    for (uint32_t i = threadIdx.x; i < D; i += SC_T) s_acnt[i] = 0;     // visible after the first barrier below
    bool synced = false;
    if (PASS == 1) {
        uint64_t tile = blockIdx.x;
        if (STR) {
            const uint32_t m = blockIdx.x + 1;
            const uint64_t p = sample_off[m];
            tile = p / PT_TILE;
            if (p >= n_total || p == tile * PT_TILE || sample_off[m - 1] > tile * PT_TILE) return;
        }
        beg = tile * PT_TILE; end = min(n_total, beg + PT_TILE);
    } else {
        synced = sg.tile_pref != nullptr;                    // the lookup over segments contains a barrier
        tile_range_coop<SC_T>(blockIdx.x, sg, n_total, beg, end, region);
    }
    uint64_t rr[SC_I];
    const uint32_t cnt = (uint32_t)(end - beg);
    const uint64_t* __restrict__ sp = src + beg;
#pragma unroll
    for (int j = 0; j < SC_I; j++) {
        const uint32_t li = j * SC_T + threadIdx.x;
        rr[j] = li < cnt ? SKA_SC_LOAD(sp + li) : REC_SENTINEL;
    }
    uint32_t sid_lo = 0, sid_hi = 0;
    if (PASS == 1) {                                         // sample ids of the tile's first and last position
        const uint64_t pf = beg, pl = end - 1;
        if (!STR && n_samples <= SC_T) {                     // sample_off ascends: the id is the number of offsets <= position, minus one
            const uint64_t so = threadIdx.x < n_samples ? sample_off[threadIdx.x] : ~0ULL;
            sid_hi = (uint32_t)__syncthreads_count(so <= pl) - 1u;
            sid_lo = sample_off[sid_hi] <= pf ? sid_hi : sid_hi - 1u;     // only "the same sample or not" matters here
        } else {
            if (threadIdx.x == 0) {
                uint32_t lo = 0, hi = n_samples;
                while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (sample_off[m] <= pf) lo = m; else hi = m; }
                s_sid[0] = lo; hi = n_samples;
                while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (sample_off[m] <= pl) lo = m; else hi = m; }
                s_sid[1] = lo;
            }
            __syncthreads();
            sid_lo = s_sid[0]; sid_hi = s_sid[1];
        }
        if (!STR && sid_lo != sid_hi) return;                // the straddling launch takes this tile
    } else if (!synced) {
        __syncthreads();
    }

    for (uint32_t sid = sid_lo; sid <= sid_hi; sid++) {
        // positions of this round inside the tile (the whole tile unless it straddles samples)
        uint32_t a = 0, b = cnt;
        if (PASS == 1 && STR) {
            const uint64_t so = sample_off[sid], se = sample_off[sid + 1];
            a = so > beg ? (uint32_t)min((uint64_t)cnt, so - beg) : 0u;
            b = se > beg ? (uint32_t)min((uint64_t)cnt, se - beg) : 0u;
            if (sid != sid_lo) {                             // the previous round has left the staging buffer
                __syncthreads();
                for (uint32_t i = threadIdx.x; i < D; i += SC_T) s_acnt[i] = 0;
                __syncthreads();
            }
        }
        // the digit of record j of this thread, PT_NONE if it takes no part in this round (pass 1: "no window here" is the
        // all-ones record; at k = 31 its top bits are a legal digit, so the whole word is compared)
        auto digit = [&](int j) -> uint32_t {
            const uint32_t li = j * SC_T + threadIdx.x;
            const uint64_t r = rr[j];
            if (PASS == 1) return (r == REC_SENTINEL || (STR && (li < a || li >= b))) ? PT_NONE : (uint32_t)(r >> fn.shift);
            return li < cnt ? (fn.mask ? (uint32_t)(r >> fn.shift) & fn.mask : 0u) : PT_NONE;
        };
        // A. records per digit
#pragma unroll
        for (int j = 0; j < SC_I; j++) {
            const uint32_t d = digit(j);
            if (d != PT_NONE) atomicAdd(&s_acnt[d], 1u);
        }
        __syncthreads();
        // B. exclusive scan over digits -> s_cur (the slot cursors of phase C); slice of every destination bucket
        uint32_t n_valid;
        if (PER_THREAD) {
            // one thread per digit: the reservation is issued first and its result used last, so the round trip to L2
            // lies under the scan and the staging; every warp scans all group sums itself (no second barrier)
            const uint32_t v = threadIdx.x < D ? s_acnt[threadIdx.x] : 0u;
            unsigned long long g = 0;
            if (v) g = atomicAdd(&cursor[(size_t)region * D + threadIdx.x], (unsigned long long)v);
            uint32_t gs = 0;                                                                // digits 8 * lane .. 8 * lane + 7
            if (8u * lane < D) {
                const uint4 q0 = reinterpret_cast<const uint4*>(s_acnt)[2 * lane], q1 = reinterpret_cast<const uint4*>(s_acnt)[2 * lane + 1];
                gs = q0.x + q0.y + q0.z + q0.w + q1.x + q1.y + q1.z + q1.w;
            }
            uint32_t ginc = gs, inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, ginc, o), u = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if (lane >= o) { ginc += t; inc += u; }
            }
            const uint32_t base = __shfl_sync(0xFFFFFFFFu, ginc - gs, 4 * warp);            // digits below 32 * warp
            n_valid = __shfl_sync(0xFFFFFFFFu, ginc, 31);
            const uint32_t excl = base + inc - v;
            if (threadIdx.x < D) s_cur[threadIdx.x] = excl;
            __syncthreads();
            // C. stage in digit order
#pragma unroll
            for (int j = 0; j < SC_I; j++) {
                const uint32_t d = digit(j);
                if (d != PT_NONE) s_stage[atomicAdd(&s_cur[d], 1u)] = rr[j];
            }
            if (threadIdx.x < D) s_gbase[threadIdx.x] = g - excl;
        } else {
            if (warp == 0) {
                const uint32_t per = (D + 31) / 32, b0 = lane * per, e0 = min(D, b0 + per);
                uint32_t sum = 0;
                for (uint32_t d = b0; d < e0; d++) sum += s_acnt[d];
                uint32_t inc = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += v; }
                uint32_t run = inc - sum;
                for (uint32_t d = b0; d < e0; d++) { s_cur[d] = run; run += s_acnt[d]; }
                if (lane == 31) s_total = inc;
            }
            __syncthreads();
            n_valid = s_total;
            for (uint32_t d = threadIdx.x; d < D; d += SC_T) {
                const uint32_t tot = s_acnt[d];
                if (tot) s_gbase[d] = atomicAdd(&cursor[(size_t)region * D + d], (unsigned long long)tot) - s_cur[d];
            }
            __syncthreads();                                 // the reservations read s_cur before phase C moves it
#pragma unroll
            for (int j = 0; j < SC_I; j++) {
                const uint32_t d = digit(j);
                if (d != PT_NONE) s_stage[atomicAdd(&s_cur[d], 1u)] = rr[j];
            }
        }
        __syncthreads();
        // D. coalesced runs out; the digit comes back out of the record
        const uint64_t sid_bits = (uint64_t)(sid + pk.sid_base) << 2;
        for (uint32_t slot = threadIdx.x; slot < n_valid; slot += SC_T) {
            const uint64_t r = s_stage[slot];
            if (PASS == 1) {
                const uint32_t d = (uint32_t)(r >> fn.shift);
                const uint32_t mid = (uint32_t)__ffs((uint32_t)(r & 0xF)) - 1u;
                const uint64_t prec = (((r >> 4) & pk.low_mask) << (pk.sb + 2)) | sid_bits | mid;
                const unsigned long long g = s_gbase[d] + slot;
                dst[g] = prec;
                if (next_digit) next_digit[g] = (uint8_t)((uint32_t)(prec >> nd_shift) & nd_mask);
            } else {
                const uint32_t d = fn.mask ? (uint32_t)(r >> fn.shift) & fn.mask : 0u;
                dst[s_gbase[d] + slot] = r;
            }
        }
    }
}
static size_t scatter2_smem(int bits) {
    const size_t D = (size_t)1 << bits;
    return PT_TILE * 8 + D * 8 + D * 4 + D * 4 + 64;
}
static bool scatter_v1() { const char* e = getenv("SKA_SCATTER_V1"); return e && atoi(e) != 0; }   // read per build: tests switch it

// ---------------------------------------------------------------------------------------------- wide records (k > 31)
// The same two passes on 16-byte records (key of up to 124 bits << 4 | one-hot middle from K1; packed records carry
// key_low << (sb + 2) | sample << 2 | middle in 128 bits, which needs P1 >= sb - 2 exactly as the 64-bit form does).
typedef unsigned __int128 u128;
__device__ __forceinline__ u128 ld_rec128(const uint64_t* __restrict__ p, uint64_t i) {
    const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(p) + i);
    return ((u128)v.y << 64) | v.x;
}
__device__ __forceinline__ uint32_t digit128(u128 r, int pass, int shift, uint32_t mask) {
    if (pass == 1) return r == ~(u128)0 ? PT_NONE : (uint32_t)(r >> shift);
    if (pass == 3) return r == ~(u128)0 ? PT_NONE : (uint32_t)(r >> shift) & mask;         // first pass of a plain grouping (no repacking)
    return mask ? (uint32_t)(r >> shift) & mask : 0u;
}
__global__ void __launch_bounds__(PT_THREADS)
part_hist_wide_kernel(const uint64_t* __restrict__ src, uint64_t n_total, uint64_t n_tiles, Segs sg, DigitFn fn, int bits,
                      unsigned long long* __restrict__ hist) {
    extern __shared__ uint32_t s_cnt[];
    const uint32_t D = 1u << bits;
    const uint64_t per = (n_tiles + gridDim.x - 1) / gridDim.x;
    const uint64_t t_beg = min(n_tiles, (uint64_t)blockIdx.x * per), t_end = min(n_tiles, t_beg + per);
    uint32_t cur_region = PT_NONE;
    for (uint32_t i = threadIdx.x; i < D; i += PT_THREADS) s_cnt[i] = 0;
    __syncthreads();
    auto flush = [&](uint32_t region) {
        __syncthreads();
        for (uint32_t d = threadIdx.x; d < D; d += PT_THREADS) {
            uint32_t tot = s_cnt[d];
            s_cnt[d] = 0;
            if (tot) atomicAdd(&hist[(size_t)region * D + d], (unsigned long long)tot);
        }
        __syncthreads();
    };
    // as in part_hist_kernel: the next records are in flight while these are counted, and the segment index walks forward.
    // The unit is half a tile (four 16-byte records per thread in flight + four being counted: 32 registers).
    constexpr int WI = PT_ITEMS / 2;
    constexpr uint32_t HALF = PT_TILE / 2;
    static_assert(WI * PT_THREADS == (int)HALF, "half tiles");
    ulonglong2 nx[WI];
    uint64_t nbeg = 0, nend = 0; uint32_t nregion = 0, seg = PT_NONE;
    auto next_unit = [&](uint64_t u) {
        uint64_t b, e;
        tile_range_next(u >> 1, sg, n_total, seg, b, e, nregion);
        nbeg = min(e, b + (u & 1) * HALF); nend = min(e, nbeg + HALF);
#pragma unroll
        for (int j = 0; j < WI; j++) {
            const uint64_t i = nbeg + (uint64_t)j * PT_THREADS + threadIdx.x;
            nx[j] = i < nend ? __ldg(reinterpret_cast<const ulonglong2*>(src) + i) : make_ulonglong2(~0ULL, ~0ULL);
        }
    };
    const uint64_t u_beg = 2 * t_beg, u_end = 2 * t_end;
    if (u_beg < u_end) next_unit(u_beg);
    for (uint64_t u = u_beg; u < u_end; u++) {
        const uint64_t beg = nbeg, end = nend; const uint32_t region = nregion;
        ulonglong2 rr[WI];
#pragma unroll
        for (int j = 0; j < WI; j++) rr[j] = nx[j];
        if (u + 1 < u_end) next_unit(u + 1);
        if (region != cur_region) { if (cur_region != PT_NONE) flush(cur_region); cur_region = region; }
#pragma unroll
        for (int j = 0; j < WI; j++) {
            const uint64_t i = beg + (uint64_t)j * PT_THREADS + threadIdx.x;
            const uint32_t d = i < end ? digit128(((u128)rr[j].y << 64) | rr[j].x, fn.pass, fn.shift, fn.mask) : PT_NONE;
            if (d != PT_NONE) atomicAdd(&s_cnt[d], 1u);
        }
    }
    if (cur_region != PT_NONE) flush(cur_region);
}

struct PackWide { int sb; int low_bits; uint32_t sid_base; };
constexpr int SCW_T = 256;
constexpr int SCW_I = PT_TILE / SCW_T;
template <int PASS>
__global__ void __launch_bounds__(SCW_T, 2)
part_scatter_wide_kernel(const uint64_t* __restrict__ src, uint64_t n_total, Segs sg, DigitFn fn, int bits, PackWide pk,
                         const uint64_t* __restrict__ sample_off, uint32_t n_samples, unsigned long long* __restrict__ cursor,
                         uint64_t* __restrict__ dst) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t D = 1u << bits;
    ulonglong2* s_stage = reinterpret_cast<ulonglong2*>(smem);                   // [PT_TILE]
    unsigned long long* s_gbase = reinterpret_cast<unsigned long long*>(s_stage + PT_TILE);   // [D]
    uint32_t* s_doff = reinterpret_cast<uint32_t*>(s_gbase + D);                 // [D + 1]
    uint32_t* s_acnt = s_doff + D + 1;                                           // [D]
    uint16_t* s_sdig = reinterpret_cast<uint16_t*>(s_acnt + D);                  // [PT_TILE]
    __shared__ uint32_t s_sid[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t beg, end; uint32_t region;
    tile_range_coop<SCW_T>(blockIdx.x, sg, n_total, beg, end, region);
    const uint32_t cnt = (uint32_t)(end - beg);
    u128 rr[SCW_I]; uint32_t dp[SCW_I];
#pragma unroll
    for (int j = 0; j < SCW_I; j++) {
        const uint32_t li = j * SCW_T + threadIdx.x;
        rr[j] = li < cnt ? ld_rec128(src, beg + li) : ~(u128)0;
    }
    for (uint32_t i = threadIdx.x; i < D; i += SCW_T) s_acnt[i] = 0;
    if (PASS == 1 && threadIdx.x == 0) {
        uint64_t pf = beg, pl = end - 1;
        uint32_t lo = 0, hi = n_samples;
        while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (sample_off[m] <= pf) lo = m; else hi = m; }
        s_sid[0] = lo; hi = n_samples;
        while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (sample_off[m] <= pl) lo = m; else hi = m; }
        s_sid[1] = lo;
    }
    __syncthreads();
    const u128 low_mask = pk.low_bits >= 128 ? ~(u128)0 : (((u128)1 << pk.low_bits) - 1);
#pragma unroll
    for (int j = 0; j < SCW_I; j++) {
        const uint32_t li = j * SCW_T + threadIdx.x;
        const u128 r = rr[j];
        const uint32_t d = li < cnt ? digit128(r, fn.pass, fn.shift, fn.mask) : PT_NONE;
        if (PASS == 1 && d != PT_NONE) {
            uint32_t s = s_sid[0];
            if (s_sid[0] != s_sid[1]) {
                const uint64_t i = beg + li;
                uint32_t lo = s_sid[0], hi = s_sid[1] + 1;
                while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (sample_off[m] <= i) lo = m; else hi = m; }
                s = lo;
            }
            const u128 key = r >> 4;
            const uint32_t mid = (uint32_t)__ffs((uint32_t)r & 0xFu) - 1u;
            rr[j] = ((key & low_mask) << (pk.sb + 2)) | ((u128)(s + pk.sid_base) << 2) | mid;
        }
        dp[j] = d != PT_NONE ? ((d << 16) | atomicAdd(&s_acnt[d], 1u)) : PT_NONE;
    }
    __syncthreads();
    if (warp == 0) {
        const uint32_t per = (D + 31) / 32, b0 = lane * per, e0 = min(D, b0 + per);
        uint32_t sum = 0;
        for (uint32_t d = b0; d < e0; d++) sum += s_acnt[d];
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += v; }
        uint32_t run = inc - sum;
        for (uint32_t d = b0; d < e0; d++) { uint32_t t = s_acnt[d]; s_doff[d] = run; run += t; }
        if (lane == 31) s_doff[D] = inc;
    }
    __syncthreads();
    for (uint32_t d = threadIdx.x; d < D; d += SCW_T) {
        uint32_t tot = s_doff[d + 1] - s_doff[d];
        if (tot) s_gbase[d] = atomicAdd(&cursor[(size_t)region * D + d], (unsigned long long)tot) - s_doff[d];
    }
#pragma unroll
    for (int j = 0; j < SCW_I; j++) {
        if (dp[j] == PT_NONE) continue;
        const uint32_t d = dp[j] >> 16, slot = s_doff[d] + (dp[j] & 0xFFFFu);
        s_stage[slot] = make_ulonglong2((unsigned long long)rr[j], (unsigned long long)(rr[j] >> 64));
        s_sdig[slot] = (uint16_t)d;
    }
    __syncthreads();
    const uint32_t n_valid = s_doff[D];
    ulonglong2* dst2 = reinterpret_cast<ulonglong2*>(dst);
    for (uint32_t slot = threadIdx.x; slot < n_valid; slot += SCW_T) dst2[s_gbase[s_sdig[slot]] + slot] = s_stage[slot];
}
// Plain grouping of 16-byte records without the staging buffer: with 2^11 digits a tile of 4096 records holds two records per
// digit, so staging cannot make the global writes any wider than they are -- every record goes from its register straight to
// digit base + rank.  Shared memory holds the counters only (24 KB at 11 bits), 512 threads x 8 records.
constexpr int SCD_T = 512;
constexpr int SCD_I = PT_TILE / SCD_T;
__global__ void __launch_bounds__(SCD_T, 2)
part_scatter_wide_direct_kernel(const uint64_t* __restrict__ src, uint64_t n_total, Segs sg, DigitFn fn, int bits,
                                unsigned long long* __restrict__ cursor, uint64_t* __restrict__ dst) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t D = 1u << bits;
    unsigned long long* s_gbase = reinterpret_cast<unsigned long long*>(smem);   // [D]
    uint32_t* s_acnt = reinterpret_cast<uint32_t*>(s_gbase + D);                 // [D]
    uint64_t beg, end; uint32_t region;
    tile_range_coop<SCD_T>(blockIdx.x, sg, n_total, beg, end, region);
    const uint32_t cnt = (uint32_t)(end - beg);
    ulonglong2 rr[SCD_I]; uint32_t dp[SCD_I];
    const ulonglong2* src2 = reinterpret_cast<const ulonglong2*>(src) + beg;
#pragma unroll
    for (int j = 0; j < SCD_I; j++) {
        const uint32_t li = j * SCD_T + threadIdx.x;
        rr[j] = li < cnt ? __ldg(src2 + li) : make_ulonglong2(~0ULL, ~0ULL);
    }
    for (uint32_t i = threadIdx.x; i < D; i += SCD_T) s_acnt[i] = 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SCD_I; j++) {
        const u128 r = ((u128)rr[j].y << 64) | rr[j].x;
        const uint32_t d = j * SCD_T + threadIdx.x < cnt ? digit128(r, fn.pass, fn.shift, fn.mask) : PT_NONE;
        dp[j] = d != PT_NONE ? ((d << 16) | atomicAdd(&s_acnt[d], 1u)) : PT_NONE;
    }
    __syncthreads();
    for (uint32_t d = threadIdx.x; d < D; d += SCD_T) {
        const uint32_t tot = s_acnt[d];
        if (tot) s_gbase[d] = atomicAdd(&cursor[(size_t)region * D + d], (unsigned long long)tot);
    }
    __syncthreads();
    ulonglong2* dst2 = reinterpret_cast<ulonglong2*>(dst);
#pragma unroll
    for (int j = 0; j < SCD_I; j++)
        if (dp[j] != PT_NONE) dst2[s_gbase[dp[j] >> 16] + (dp[j] & 0xFFFFu)] = rr[j];
}
static bool scatter_direct_enabled() { const char* e = getenv("SKA_SCATTER_DIRECT"); return !e || atoi(e) != 0; }

static size_t scatter_wide_smem(int bits) {
    const size_t D = (size_t)1 << bits;
    return PT_TILE * 16 + D * 8 + (D + 1) * 4 + D * 4 + PT_TILE * 2 + 64;
}

// ---------------------------------------------------------------------------------------------- small helpers
// exclusive scan of u64 counts -> offsets[n+1]; also cursors[i] = offsets[i] and, for the second level, the
// number of tiles before every region.  Single CTA, coalesced tiles of 1024 x 4 elements.
__global__ void __launch_bounds__(1024)
part_scan_kernel(const unsigned long long* __restrict__ cnt, uint32_t n, uint64_t* __restrict__ off,
                 unsigned long long* __restrict__ cursor, uint64_t* __restrict__ tile_pref /* may be null */,
                 int n_copies /* cnt holds this many histograms of n counters to be added up */) {
    __shared__ uint64_t s_warp[2][32];
    __shared__ uint64_t s_carry[2];
    const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t < 2) s_carry[t] = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 4096) {
        uint64_t v[4], tl[4], sum = 0, tsum = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t i = base + 4 * t + j;
            v[j] = 0ull;
            if (i < n) for (int cpy = 0; cpy < n_copies; cpy++) v[j] += cnt[(size_t)cpy * n + i];
            tl[j] = (v[j] + PT_TILE - 1) / PT_TILE;
            sum += v[j]; tsum += tl[j];
        }
        uint64_t inc = sum, tinc = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t u = __shfl_up_sync(0xFFFFFFFFu, inc, o), w = __shfl_up_sync(0xFFFFFFFFu, tinc, o);
            if (lane >= (uint32_t)o) { inc += u; tinc += w; }
        }
        if (lane == 31) { s_warp[0][warp] = inc; s_warp[1][warp] = tinc; }
        __syncthreads();
        if (warp < 2) {
            uint64_t w = s_warp[warp][lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint64_t u = __shfl_up_sync(0xFFFFFFFFu, winc, o); if (lane >= (uint32_t)o) winc += u; }
            s_warp[warp][lane] = winc - w;
        }
        __syncthreads();
        uint64_t run = s_carry[0] + s_warp[0][warp] + (inc - sum), trun = s_carry[1] + s_warp[1][warp] + (tinc - tsum);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t i = base + 4 * t + j;
            if (i < n) { off[i] = run; if (cursor) cursor[i] = run; if (tile_pref) tile_pref[i] = trun; }
            run += v[j]; trun += tl[j];
        }
        __syncthreads();
        if (t == 1023) { s_carry[0] = run; s_carry[1] = trun; }
        __syncthreads();
    }
    if (t == 0) { off[n] = s_carry[0]; if (tile_pref) tile_pref[n] = s_carry[1]; }
}

static size_t scatter_smem(int bits) {
    const size_t D = (size_t)1 << bits;
    return PT_TILE * 8 + D * 8 + (D + 1) * 4 + D * 4 + PT_TILE * 2 + 64;
}

// The same scan for many counters (2^20 and more buckets): every CTA sums its 4096 counters, one CTA scans the sums, every
// CTA scans its counters again from its carry.  (The single CTA above walks 4 M counters in 1.7 ms.)
constexpr uint32_t SCAN_CHUNK = 4096;
__device__ __forceinline__ void cta_reduce2(uint64_t& a, uint64_t& b, uint64_t (*s_w)[32]) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xFFFFFFFFu, a, o); b += __shfl_xor_sync(0xFFFFFFFFu, b, o); }
    if (lane == 0) { s_w[0][warp] = a; s_w[1][warp] = b; }
    __syncthreads();
    a = s_w[0][lane]; b = s_w[1][lane];                      // 1024 threads: 32 warps
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xFFFFFFFFu, a, o); b += __shfl_xor_sync(0xFFFFFFFFu, b, o); }
}
__global__ void __launch_bounds__(1024)
part_scan_sums_kernel(const unsigned long long* __restrict__ cnt, uint32_t n, int n_copies, unsigned long long* __restrict__ sums /* [2][gridDim.x] */) {
    __shared__ uint64_t s_w[2][32];
    uint64_t sum = 0, tsum = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t i = blockIdx.x * SCAN_CHUNK + 4 * threadIdx.x + j;
        uint64_t v = 0;
        if (i < n) for (int cpy = 0; cpy < n_copies; cpy++) v += cnt[(size_t)cpy * n + i];
        sum += v; tsum += (v + PT_TILE - 1) / PT_TILE;
    }
    cta_reduce2(sum, tsum, s_w);
    if (threadIdx.x == 0) { sums[blockIdx.x] = sum; sums[gridDim.x + blockIdx.x] = tsum; }
}
__global__ void __launch_bounds__(1024)
part_scan_apply_kernel(const unsigned long long* __restrict__ cnt, uint32_t n, int n_copies, const uint64_t* __restrict__ carry /* [2][grid + 1] */,
                       uint64_t* __restrict__ off, unsigned long long* __restrict__ cursor, uint64_t* __restrict__ tile_pref) {
    __shared__ uint64_t s_warp[2][32];
    const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
    uint64_t v[4], tl[4], sum = 0, tsum = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t i = blockIdx.x * SCAN_CHUNK + 4 * t + j;
        v[j] = 0ull;
        if (i < n) for (int cpy = 0; cpy < n_copies; cpy++) v[j] += cnt[(size_t)cpy * n + i];
        tl[j] = (v[j] + PT_TILE - 1) / PT_TILE;
        sum += v[j]; tsum += tl[j];
    }
    uint64_t inc = sum, tinc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t u = __shfl_up_sync(0xFFFFFFFFu, inc, o), w = __shfl_up_sync(0xFFFFFFFFu, tinc, o);
        if (lane >= (uint32_t)o) { inc += u; tinc += w; }
    }
    if (lane == 31) { s_warp[0][warp] = inc; s_warp[1][warp] = tinc; }
    __syncthreads();
    if (warp < 2) {
        uint64_t w = s_warp[warp][lane], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint64_t u = __shfl_up_sync(0xFFFFFFFFu, winc, o); if (lane >= (uint32_t)o) winc += u; }
        s_warp[warp][lane] = winc - w;
    }
    __syncthreads();
    uint64_t run = carry[blockIdx.x] + s_warp[0][warp] + (inc - sum), trun = carry[gridDim.x + 1 + blockIdx.x] + s_warp[1][warp] + (tinc - tsum);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t i = blockIdx.x * SCAN_CHUNK + 4 * t + j;
        if (i < n) { off[i] = run; if (cursor) cursor[i] = run; if (tile_pref) tile_pref[i] = trun; }
        run += v[j]; trun += tl[j];
    }
    if (blockIdx.x == gridDim.x - 1 && t == 1023) { off[n] = run; if (tile_pref) tile_pref[n] = trun; }
}

static void launch_part_scan(const unsigned long long* cnt, uint32_t n, uint64_t* off, unsigned long long* cursor, uint64_t* tile_pref, int n_copies) {
    if (n <= 4 * SCAN_CHUNK) {
        SKA_LAUNCH(part_scan_kernel, 1, 1024, 0, cnt, n, off, cursor, tile_pref, n_copies);
        return;
    }
    const uint32_t g = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
    unsigned long long* sums = arena().take_n<unsigned long long>(2 * (size_t)g);
    uint64_t* carry = arena().take_n<uint64_t>(2 * ((size_t)g + 1));
    SKA_LAUNCH(part_scan_sums_kernel, g, 1024, 0, cnt, n, n_copies, sums);
    SKA_LAUNCH(part_scan_kernel, 1, 1024, 0, (const unsigned long long*)sums, g, carry, (unsigned long long*)nullptr, (uint64_t*)nullptr, 1);
    SKA_LAUNCH(part_scan_kernel, 1, 1024, 0, (const unsigned long long*)(sums + g), g, carry + g + 1, (unsigned long long*)nullptr, (uint64_t*)nullptr, 1);
    SKA_LAUNCH(part_scan_apply_kernel, g, 1024, 0, cnt, n, n_copies, (const uint64_t*)carry, off, cursor, tile_pref);
}

PartPlan make_part_plan(uint32_t k, int prefix_bits, uint64_t n_samples_total) {
    PartPlan pl;
    pl.key_bits = 2 * ((int)k - 1);
    pl.wide = k > 31;
    pl.sb = 1; while ((1ull << pl.sb) < n_samples_total) pl.sb++;
    const int need1 = std::max(0, pl.key_bits + pl.sb + 2 - (pl.wide ? 128 : 64));       // the first digit must free room for the sample id
    pl.P = std::min(std::max(prefix_bits, need1), pl.key_bits);
    if (pl.P > 2 * PT_MAX_BITS) pl.P = 2 * PT_MAX_BITS;
    pl.P1 = std::max((pl.P + 1) / 2, need1);
    if (const char* e = getenv("SKA_P1_BITS")) { const int v = atoi(e); if (v >= need1 && v <= pl.P && v <= PT_MAX_BITS && pl.P - v <= PT_MAX_BITS) pl.P1 = v; }   // tuning/tests
    if (pl.P1 > PT_MAX_BITS || pl.P1 > pl.P) throw Error("too many samples for the packed record format at this k");
    pl.P2 = pl.P - pl.P1;
    pl.low_bits = pl.key_bits - pl.P1;
    return pl;
}

static void set_scatter_attrs() {
    static bool done = false;
    if (done) return;
    SKA_CUDA(cudaFuncSetAttribute(part_scatter_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter_smem(PT_MAX_BITS)));
    SKA_CUDA(cudaFuncSetAttribute(part_scatter_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter_smem(PT_MAX_BITS)));
    SKA_CUDA(cudaFuncSetAttribute(part_scatter_kernel<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter_smem(8)));
    SKA_CUDA(cudaFuncSetAttribute(part_scatter_kernel<2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter_smem(8)));
    SKA_CUDA(cudaFuncSetAttribute(part_scatter2_kernel<1, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter2_smem(PT_MAX_BITS)));
    SKA_CUDA(cudaFuncSetAttribute(part_scatter2_kernel<1, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter2_smem(PT_MAX_BITS)));
    SKA_CUDA(cudaFuncSetAttribute(part_scatter2_kernel<2, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter2_smem(PT_MAX_BITS)));
    SKA_CUDA(cudaFuncSetAttribute(part_scatter2_kernel<1, 7, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter2_smem(8)));
    SKA_CUDA(cudaFuncSetAttribute(part_scatter2_kernel<1, 7, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter2_smem(8)));
    SKA_CUDA(cudaFuncSetAttribute(part_scatter2_kernel<2, 7, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter2_smem(8)));
    SKA_CUDA(cudaFuncSetAttribute(part_scatter2_kernel<1, 8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter2_smem(8)));
    SKA_CUDA(cudaFuncSetAttribute(part_scatter2_kernel<1, 8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter2_smem(8)));
    SKA_CUDA(cudaFuncSetAttribute(part_scatter2_kernel<2, 8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter2_smem(8)));
    SKA_CUDA(cudaFuncSetAttribute(part_scatter_wide_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter_wide_smem(PT_MAX_BITS)));
    SKA_CUDA(cudaFuncSetAttribute(part_scatter_wide_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter_wide_smem(PT_MAX_BITS)));
    done = true;
}

// pass 1.  rec: K1 output (n_pos entries, linear) -> packed records grouped by the top P1 key bits in `*out`
// (arena), region offsets `*d_off1` [2^P1 + 1] and tiles-before-region `*d_tile_pref` [2^P1 + 1] on the device.
void partition_coarse(const uint64_t* rec, uint64_t n_pos, const uint64_t* d_sample_off, uint32_t n_samples_local, uint32_t sid_base,
                      const PartPlan& pl, uint64_t** out, uint64_t** d_off1, uint64_t** d_tile_pref, const unsigned long long* k1_hist,
                      uint8_t** d_next_digit) {
    Context& c = ctx();
    timer().begin("hist1");
    const uint32_t D1 = 1u << pl.P1;
    unsigned long long* hist1 = arena().take_n<unsigned long long>(D1);
    unsigned long long* cur1 = arena().take_n<unsigned long long>(D1);
    uint64_t* off1 = arena().take_n<uint64_t>(D1 + 1);
    uint64_t* tile_pref = arena().take_n<uint64_t>(D1 + 1);
    if (!k1_hist) dev_fill(hist1, 0, D1 * 8);
    const uint64_t n_tiles1 = (n_pos + PT_TILE - 1) / PT_TILE;
    DigitFn f1{1, 4 + pl.key_bits - pl.P1, 0xFFFFFFFFu};
    PackFn pk{pl.sb, pl.low_bits >= 64 ? ~0ULL : ((1ULL << pl.low_bits) - 1ULL), sid_base};
    const unsigned hist_grid = (unsigned)std::min<uint64_t>(std::max<uint64_t>(n_tiles1, 1), (uint64_t)c.sm_count * 4);
    set_scatter_attrs();
    const Segs none{nullptr, nullptr, nullptr, nullptr, 1u};
    if (n_pos && !k1_hist) {
        if (pl.wide) SKA_LAUNCH(part_hist_wide_kernel, hist_grid, PT_THREADS, D1 * 4, rec, n_pos, n_tiles1, none, f1, pl.P1, hist1);
        else SKA_LAUNCH(part_hist_kernel, hist_grid, PT_THREADS, D1 * 4, rec, n_pos, n_tiles1, none, f1, pl.P1, hist1);
    }
    launch_part_scan(k1_hist ? k1_hist : (const unsigned long long*)hist1, D1, off1, cur1, tile_pref, k1_hist ? K1_HIST_COPIES : 1);
    timer().end();
    timer().begin("scatter1");
    uint64_t* a = arena().take_n<uint64_t>(std::max<uint64_t>(n_pos, 1) * (pl.wide ? 2 : 1));
    // the second pass's digit of every record, one byte each, in the order of `a`: its histogram then reads n bytes, not 8 n
    uint8_t* nd = nullptr;
    const int nd_shift = pl.sb + 2 + pl.low_bits - pl.P2;
    const uint32_t nd_mask = (1u << pl.P2) - 1u;
    const char* nd_env = getenv("SKA_HIST2_BYTES");
    const bool nd_enabled = !nd_env || atoi(nd_env) != 0;
    if (d_next_digit) {
        if (nd_enabled && !pl.wide && !scatter_v1() && pl.P2 >= 1 && pl.P2 <= 8 && n_pos) nd = arena().take_n<uint8_t>(n_pos + 16);     // + 16: the histogram reads whole 16-byte words
        *d_next_digit = nd;
    }
    if (n_pos && pl.wide) {
        const PackWide pw{pl.sb, pl.low_bits, sid_base};
        SKA_LAUNCH((part_scatter_wide_kernel<1>), (unsigned)n_tiles1, SCW_T, scatter_wide_smem(pl.P1), rec, n_pos, none, f1, pl.P1, pw,
                   d_sample_off, n_samples_local, cur1, a);
    } else if (n_pos) {
        if (scatter_v1()) {
            if (pl.P1 == 8)
                SKA_LAUNCH((part_scatter_kernel<1, 8>), (unsigned)n_tiles1, SC_T, scatter_smem(pl.P1), rec, n_pos, none, f1, pl.P1, pk,
                           d_sample_off, n_samples_local, cur1, a);
            else
                SKA_LAUNCH((part_scatter_kernel<1, 0>), (unsigned)n_tiles1, SC_T, scatter_smem(pl.P1), rec, n_pos, none, f1, pl.P1, pk,
                           d_sample_off, n_samples_local, cur1, a);
        } else if (pl.P1 == 8) {
            SKA_LAUNCH((part_scatter2_kernel<1, 8, 0>), (unsigned)n_tiles1, SC_T, scatter2_smem(pl.P1), rec, n_pos, none, f1, pl.P1, pk,
                       d_sample_off, n_samples_local, cur1, a, nd, nd_shift, nd_mask);
            if (n_samples_local > 1)
                SKA_LAUNCH((part_scatter2_kernel<1, 8, 1>), n_samples_local - 1, SC_T, scatter2_smem(pl.P1), rec, n_pos, none, f1, pl.P1, pk,
                           d_sample_off, n_samples_local, cur1, a, nd, nd_shift, nd_mask);
        } else if (pl.P1 == 7) {
            SKA_LAUNCH((part_scatter2_kernel<1, 7, 0>), (unsigned)n_tiles1, SC_T, scatter2_smem(pl.P1), rec, n_pos, none, f1, pl.P1, pk,
                       d_sample_off, n_samples_local, cur1, a, nd, nd_shift, nd_mask);
            if (n_samples_local > 1)
                SKA_LAUNCH((part_scatter2_kernel<1, 7, 1>), n_samples_local - 1, SC_T, scatter2_smem(pl.P1), rec, n_pos, none, f1, pl.P1, pk,
                           d_sample_off, n_samples_local, cur1, a, nd, nd_shift, nd_mask);
        } else {
            SKA_LAUNCH((part_scatter2_kernel<1, 0, 0>), (unsigned)n_tiles1, SC_T, scatter2_smem(pl.P1), rec, n_pos, none, f1, pl.P1, pk,
                       d_sample_off, n_samples_local, cur1, a, nd, nd_shift, nd_mask);
            if (n_samples_local > 1)
                SKA_LAUNCH((part_scatter2_kernel<1, 0, 1>), n_samples_local - 1, SC_T, scatter2_smem(pl.P1), rec, n_pos, none, f1, pl.P1, pk,
                           d_sample_off, n_samples_local, cur1, a, nd, nd_shift, nd_mask);
        }
    }
    *out = a; *d_off1 = off1; *d_tile_pref = tile_pref;
    timer().end();
}

// pass 2.  src: packed records; the segments (device arrays) say where the pieces of each of the `n_regions`
// regions lie.  -> dst grouped by (region, next P2 key bits), `*d_bucket_off` [n_regions * 2^P2 + 1] on the device.
void partition_fine(const uint64_t* src, const Segs& sg, uint64_t n_tiles2, uint32_t n_regions, const PartPlan& pl, uint64_t* dst,
                    uint64_t** d_bucket_off, const uint8_t* next_digit = nullptr /* the first scatter's digit bytes, in the order of src */) {
    Context& c = ctx();
    timer().begin("hist2");
    const uint32_t D2 = 1u << pl.P2;
    const size_t NB = (size_t)n_regions * D2;
    unsigned long long* hist2 = arena().take_n<unsigned long long>(std::max<size_t>(NB, 1));
    unsigned long long* cur2 = arena().take_n<unsigned long long>(std::max<size_t>(NB, 1));
    uint64_t* bucket_off = arena().take_n<uint64_t>(NB + 1);
    dev_fill(hist2, 0, std::max<size_t>(NB, 1) * 8);
    set_scatter_attrs();
    DigitFn f2{2, pl.sb + 2 + pl.low_bits - pl.P2, D2 - 1};
    PackFn pk{pl.sb, 0, 0};
    if (n_tiles2) {
        const unsigned g2 = (unsigned)std::min<uint64_t>(n_tiles2, (uint64_t)c.sm_count * 4);
        if (pl.wide) SKA_LAUNCH(part_hist_wide_kernel, g2, PT_THREADS, D2 * 4, src, (uint64_t)0, n_tiles2, sg, f2, pl.P2, hist2);
        else if (next_digit && sg.tile_pref && !sg.region && sg.end == sg.beg + 1)      // segments = whole regions, back to back
            SKA_LAUNCH(part_hist_bytes_kernel, dim3(8, n_regions), HB_THREADS, 256 * 4, next_digit, sg.beg, pl.P2, hist2);
        else SKA_LAUNCH(part_hist_kernel, g2, PT_THREADS, D2 * 4, src, (uint64_t)0, n_tiles2, sg, f2, pl.P2, hist2);
    }
    launch_part_scan((const unsigned long long*)hist2, (uint32_t)NB, bucket_off, cur2, (uint64_t*)nullptr, 1);
    timer().end();
    timer().begin("scatter2");
    if (n_tiles2 && pl.wide) {
        const PackWide pw{pl.sb, 0, 0};
        SKA_LAUNCH((part_scatter_wide_kernel<2>), (unsigned)n_tiles2, SCW_T, scatter_wide_smem(pl.P2), src, (uint64_t)0, sg, f2, pl.P2, pw,
                   (const uint64_t*)nullptr, 0u, cur2, dst);
    } else if (n_tiles2) {
        if (scatter_v1()) {
            if (pl.P2 == 8)
                SKA_LAUNCH((part_scatter_kernel<2, 8>), (unsigned)n_tiles2, SC_T, scatter_smem(pl.P2), src, (uint64_t)0, sg, f2, pl.P2, pk,
                           (const uint64_t*)nullptr, 0u, cur2, dst);
            else
                SKA_LAUNCH((part_scatter_kernel<2, 0>), (unsigned)n_tiles2, SC_T, scatter_smem(pl.P2), src, (uint64_t)0, sg, f2, pl.P2, pk,
                           (const uint64_t*)nullptr, 0u, cur2, dst);
        } else if (pl.P2 == 8)
            SKA_LAUNCH((part_scatter2_kernel<2, 8, 0>), (unsigned)n_tiles2, SC_T, scatter2_smem(pl.P2), src, (uint64_t)0, sg, f2, pl.P2, pk,
                       (const uint64_t*)nullptr, 0u, cur2, dst, (uint8_t*)nullptr, 0, 0u);
        else if (pl.P2 == 7)
            SKA_LAUNCH((part_scatter2_kernel<2, 7, 0>), (unsigned)n_tiles2, SC_T, scatter2_smem(pl.P2), src, (uint64_t)0, sg, f2, pl.P2, pk,
                       (const uint64_t*)nullptr, 0u, cur2, dst, (uint8_t*)nullptr, 0, 0u);
        else
            SKA_LAUNCH((part_scatter2_kernel<2, 0, 0>), (unsigned)n_tiles2, SC_T, scatter2_smem(pl.P2), src, (uint64_t)0, sg, f2, pl.P2, pk,
                       (const uint64_t*)nullptr, 0u, cur2, dst, (uint8_t*)nullptr, 0, 0u);
    }
    *d_bucket_off = bucket_off;
    timer().end();
}

// both passes on one GPU.  Returns the packed records grouped by the P-bit prefix in `*out` (arena memory) and the
// 2^P + 1 bucket offsets on the device in `*d_off`.
void partition_records(const uint64_t* rec, uint64_t n_pos, const uint64_t* d_sample_off, uint32_t n_samples_local, uint32_t sid_base,
                       const PartPlan& pl, uint64_t* scratch /* n_pos u64 for the second pass, may be rec */, uint64_t** out, uint64_t** d_off,
                       uint64_t** first_pass_buf, const unsigned long long* k1_hist) {
    Context& c = ctx();
    const uint32_t D1 = 1u << pl.P1;
    uint64_t *a = nullptr, *off1 = nullptr, *tile_pref = nullptr;
    uint8_t* next_digit = nullptr;
    partition_coarse(rec, n_pos, d_sample_off, n_samples_local, sid_base, pl, &a, &off1, &tile_pref, k1_hist, &next_digit);
    if (first_pass_buf) *first_pass_buf = a;
    if (pl.P2 == 0) { *out = a; *d_off = off1; return; }
    uint64_t n_tiles2 = 0;
    read_back(&n_tiles2, tile_pref + D1, 8);
    read_back_sync();
    uint64_t* b = scratch ? scratch : arena().take_n<uint64_t>(std::max<uint64_t>(n_pos, 1) * (pl.wide ? 2 : 1));
    const Segs sg{tile_pref, off1, off1 + 1, nullptr, D1};
    partition_fine(a, sg, n_tiles2, D1, pl, b, d_off, next_digit);
    *out = b;
}

// pass 2 at the owner of regions [region_lo, region_lo + nr): `n_runs` received runs, run w holding its pieces of
// those regions in order; h_run_off[w * (nr + 1) + j] = offset in `src` of region (region_lo + j) of run w.
// -> dst grouped by the full P-bit prefix (one run), bucket offsets [nr * 2^P2 + 1] on the device.
void partition_runs(const uint64_t* src, int n_runs, const uint64_t* h_run_off, uint32_t nr, const PartPlan& pl, uint64_t* dst,
                    uint64_t** d_bucket_off) {
    Context& c = ctx();
    std::vector<uint64_t> beg, end, pref;
    std::vector<uint32_t> region;
    uint64_t tiles = 0;
    for (uint32_t r = 0; r < nr; r++)                       // region-major: a CTA's run of tiles changes region rarely
        for (int w = 0; w < n_runs; w++) {
            const uint64_t b0 = h_run_off[(size_t)w * (nr + 1) + r], e0 = h_run_off[(size_t)w * (nr + 1) + r + 1];
            if (e0 < b0) throw Error("run offsets must ascend");
            if (e0 == b0) continue;
            beg.push_back(b0); end.push_back(e0); region.push_back(r); pref.push_back(tiles);
            tiles += (e0 - b0 + PT_TILE - 1) / PT_TILE;
        }
    pref.push_back(tiles);
    const uint32_t ns = (uint32_t)beg.size();
    uint64_t* d_beg = arena().take_n<uint64_t>(std::max<uint32_t>(ns, 1));
    uint64_t* d_end = arena().take_n<uint64_t>(std::max<uint32_t>(ns, 1));
    uint64_t* d_pref = arena().take_n<uint64_t>(ns + 1);
    uint32_t* d_region = arena().take_n<uint32_t>(std::max<uint32_t>(ns, 1));
    if (ns) {
        SKA_CUDA(cudaMemcpyAsync(d_beg, beg.data(), ns * 8, cudaMemcpyHostToDevice, c.stream));
        SKA_CUDA(cudaMemcpyAsync(d_end, end.data(), ns * 8, cudaMemcpyHostToDevice, c.stream));
        SKA_CUDA(cudaMemcpyAsync(d_region, region.data(), ns * 4, cudaMemcpyHostToDevice, c.stream));
    }
    SKA_CUDA(cudaMemcpyAsync(d_pref, pref.data(), (ns + 1) * 8, cudaMemcpyHostToDevice, c.stream));
    const Segs sg{d_pref, d_beg, d_end, d_region, std::max<uint32_t>(ns, 1)};
    partition_fine(src, sg, tiles, nr, pl, dst, d_bucket_off);
    SKA_CUDA(cudaStreamSynchronize(c.stream));              // the host vectors above must outlive their copies
}

// ---------------------------------------------------------------------------------------------- plain grouping of 16-byte records
// Groups n records (two words each, compared as 128-bit integers) by the P1 + P2 bits below `top_bit` (everything above is
// equal in all of them); all-ones records are dropped.  buf_a / buf_b: n records of room each (src may be buf_b).  -> *out (one of the two
// buffers), *d_off [2^(P1+P2) + 1] bucket offsets on the device.  The first half of a sort: sort128.cu finishes every bucket.
void group_wide_records(const uint64_t* src, uint64_t n, int top_bit, int P1, int P2, uint64_t* buf_a, uint64_t* buf_b, uint64_t** out,
                        uint64_t** d_off) {
    Context& c = ctx();
    if (P1 < 0 || P2 < 0 || P1 > PT_MAX_BITS || P2 > PT_MAX_BITS || top_bit - P1 - P2 < 0) throw Error("internal: bad grouping plan");
    set_scatter_attrs();
    const uint32_t D1 = 1u << P1, D2 = 1u << P2;
    unsigned long long* hist1 = arena().take_n<unsigned long long>(D1);
    unsigned long long* cur1 = arena().take_n<unsigned long long>(D1);
    uint64_t* off1 = arena().take_n<uint64_t>(D1 + 1);
    uint64_t* tile_pref = arena().take_n<uint64_t>(D1 + 1);
    dev_fill(hist1, 0, D1 * 8);
    const uint64_t n_tiles1 = (n + PT_TILE - 1) / PT_TILE;
    const DigitFn f1{3, top_bit - P1, D1 - 1};
    const Segs none{nullptr, nullptr, nullptr, nullptr, 1u};
    const PackWide pw{0, 0, 0};
    const unsigned hist_grid = (unsigned)std::min<uint64_t>(std::max<uint64_t>(n_tiles1, 1), (uint64_t)c.sm_count * 4);
    if (n) SKA_LAUNCH(part_hist_wide_kernel, hist_grid, PT_THREADS, D1 * 4, src, n, n_tiles1, none, f1, P1, hist1);
    launch_part_scan((const unsigned long long*)hist1, D1, off1, cur1, tile_pref, 1);
    const bool direct = scatter_direct_enabled();
    if (n && direct) SKA_LAUNCH(part_scatter_wide_direct_kernel, (unsigned)n_tiles1, SCD_T, (size_t)D1 * 12, src, n, none, f1, P1, cur1, buf_a);
    else if (n) SKA_LAUNCH((part_scatter_wide_kernel<2>), (unsigned)n_tiles1, SCW_T, scatter_wide_smem(P1), src, n, none, f1, P1, pw,
                           (const uint64_t*)nullptr, 0u, cur1, buf_a);
    if (P2 == 0) { *out = buf_a; *d_off = off1; return; }
    uint64_t n_tiles2 = 0;
    read_back(&n_tiles2, tile_pref + D1, 8);
    read_back_sync();
    const size_t NB = (size_t)D1 * D2;
    unsigned long long* hist2 = arena().take_n<unsigned long long>(NB);
    unsigned long long* cur2 = arena().take_n<unsigned long long>(NB);
    uint64_t* bucket_off = arena().take_n<uint64_t>(NB + 1);
    dev_fill(hist2, 0, NB * 8);
    const DigitFn f2{2, top_bit - P1 - P2, D2 - 1};
    const Segs sg{tile_pref, off1, off1 + 1, nullptr, D1};
    if (n_tiles2) {
        const unsigned g2 = (unsigned)std::min<uint64_t>(n_tiles2, (uint64_t)c.sm_count * 4);
        SKA_LAUNCH(part_hist_wide_kernel, g2, PT_THREADS, D2 * 4, buf_a, (uint64_t)0, n_tiles2, sg, f2, P2, hist2);
    }
    launch_part_scan((const unsigned long long*)hist2, (uint32_t)NB, bucket_off, cur2, (uint64_t*)nullptr, 1);
    if (n_tiles2 && direct)
        SKA_LAUNCH(part_scatter_wide_direct_kernel, (unsigned)n_tiles2, SCD_T, (size_t)D2 * 12, buf_a, (uint64_t)0, sg, f2, P2, cur2, buf_b);
    else if (n_tiles2)
        SKA_LAUNCH((part_scatter_wide_kernel<2>), (unsigned)n_tiles2, SCW_T, scatter_wide_smem(P2), buf_a, (uint64_t)0, sg, f2, P2, pw,
                   (const uint64_t*)nullptr, 0u, cur2, buf_b);
    *out = buf_b; *d_off = bucket_off;
}

}  // namespace ska
