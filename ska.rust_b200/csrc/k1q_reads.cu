// k1q_reads.cu -- the FASTQ path of `ska build` on the device (K1q + K2q): quality masks, ntHash, and an
// EXACT emulation of the reference's sequential count filter (blocked bloom + counts table).
//
// Replaces, for read input: SplitKmer's quality tests (src/ska_dict/split_kmer.rs:66-71, 98-101, 328-339),
// NtHashIterator (src/ska_dict/nthash.rs:12-77), KmerFilter (src/ska_dict/bloom_filter.rs:35-148) and the
// `middle_base_qual() && filter()==Equal` gate of SkaDict::add_file_kmers (src/ska_dict.rs:155-177).
//
// The reference filter is stateful and order dependent: an observation (a window whose middle base passes
// the quality test) is looked up in a 24 MiB blocked bloom filter; a bloom hit moves it to an exact counts
// table, and the split k-mer is inserted at the moment its count reaches min_count.  A k-mer whose FIRST
// observation is a bloom false positive is counted one higher.  Batch form, per sample:
//   observations in stream order (file 1 then file 2 == position order of the batch), hash = ntHash1
//   (canonical), K = cheap_mix(hash) (a bijection, so grouping by K groups by hash AND orders by bloom word,
//   word = (K * 3145728) >> 64);  for each distinct K: c = #observations, t = first position;
//   FP(K) = fingerprint(hash) is covered by the OR of the fingerprints of the distinct hashes of the same
//   word with smaller t;   observation number j (1-based, stream order) of K is inserted iff
//     min_count == 2 :  j >= 2 or FP            min_count >= 3 :  min(j + FP, 65535) == min_count and (j >= 2 or FP)
// Steps: (1) k1q_kernel<OBS>: per window validity (N, optional strict quality, record ends), middle-base
// quality, ntHash by a prefix-XOR scan (the hash is a XOR of rotated per-base constants, so
// fh(s) = rotl(G(s+k)^G(s), s+k-1) for the prefix G of rotr(H[e_j], j); no sequential roll); (2) sort of the 16-byte
// records (K, position) -- sort128.cu: two scatter passes by hash prefix + a warp-level bitonic sort per bucket; (3) runs -> distinct entries (word, t); sort
// by (word, t); one thread walks each bloom word's entries in t order OR-ing fingerprints -> FP; (4) the
// accepted observations set bits in a per-position accept mask; (5) k1q_kernel<EMIT> recomputes the split
// k-mer of every accepted position and writes compacted records, which then take the normal
// partition + merge path (partition.cu, merge.cu).  min_count <= 1 skips (2)-(4): every observation is accepted.
#include "k1_helpers.cuh"

namespace ska {

enum { RQ_MASK = 0, RQ_OBS = 1, RQ_EMIT = 2 };
constexpr uint64_t BLOOM_WORDS = 3145728;                     // bloom_filter.rs:15-20,93-96
constexpr uint64_t MIX_MUL = 0x85D059AA333121CFULL;           // bloom_filter.rs:56-58
constexpr uint64_t MIX_INV = 0x756DE8DBB3C2452FULL;           // MIX_MUL^-1 mod 2^64

__device__ __forceinline__ uint64_t rotl64(uint64_t x, uint32_t r) { r &= 63u; return r ? (x << r) | (x >> (64u - r)) : x; }
__device__ __forceinline__ uint64_t rotr64(uint64_t x, uint32_t r) { r &= 63u; return r ? (x >> r) | (x << (64u - r)) : x; }
__device__ __forceinline__ uint64_t nt_seed(uint32_t e) {     // nthash.rs:12-17, index = 2-bit code A C T G
    return e == 0 ? 0x3c8bfbb395c60474ULL : e == 1 ? 0x3193c18562a02b4cULL : e == 2 ? 0x295549f54be24456ULL : 0x20323ed082572324ULL;
}
__device__ __forceinline__ uint64_t cheap_mix(uint64_t h) { return (h ^ (h >> 31)) * MIX_MUL; }
__device__ __forceinline__ uint64_t cheap_unmix(uint64_t k) { uint64_t x = k * MIX_INV; return x ^ (x >> 31) ^ (x >> 62); }
__device__ __forceinline__ uint32_t bloom_word(uint64_t mixed) { return (uint32_t)__umul64hi(mixed, BLOOM_WORDS); }   // bloom_filter.rs:50-52,72-74
__device__ __forceinline__ uint64_t fingerprint(uint64_t h) {                                                          // bloom_filter.rs:62-68
    return 1ULL << (h & 63) | 1ULL << ((h >> 6) & 63) | 1ULL << ((h >> 12) & 63) | 1ULL << ((h >> 18) & 63) | 1ULL << ((h >> 24) & 63);
}

// One tile of K1_T window-start positions of ONE sample's position range [lo, hi).  mask words are indexed
// from the first tile of the sample (tile0): word (p - tile0*K1_T) / 32.
template <int MODE>
__global__ void __launch_bounds__(K1_THREADS)
k1q_kernel(const uint8_t* __restrict__ seq, const uint8_t* __restrict__ qual, const uint32_t* __restrict__ start_bits,
           uint64_t lo, uint64_t hi, uint64_t tile0, int k, int rc, uint32_t min_qual, int qmode,
           uint32_t* __restrict__ mask, ulonglong2* __restrict__ obs_rec /* x = position, y = mixed hash; all-ones = no observation */,
           unsigned long long* __restrict__ n_obs, const uint64_t* __restrict__ tile_off, uint64_t* __restrict__ rec_c) {
    __shared__ __align__(128) uint8_t s_ascii[K1_BYTES];
    __shared__ __align__(128) uint8_t s_qual[K1_BYTES];
    __shared__ uint32_t s_le[2 * K1_G32 + 2];
    __shared__ uint32_t s_be[2 * K1_G32 + 2];
    __shared__ uint32_t s_bad[K1_G32];
    __shared__ uint32_t s_qbad[K1_G32];
    __shared__ uint32_t s_start[K1_G32];
    __shared__ uint32_t s_ok[K1_T / 32];
    __shared__ uint32_t s_pre[K1_T / 32 + 1];
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ uint64_t s_wx[8], s_wy[8];
    __shared__ uint32_t s_nobs;
    extern __shared__ __align__(16) uint8_t dyn[];                    // OBS: GX[K1_BYTES + 1], GY[K1_BYTES + 1]

    const int tid = threadIdx.x;
    const uint64_t tile = tile0 + blockIdx.x;
    const uint64_t t0 = tile * K1_T;
    if (tid == 0) { mbar_init(&s_mbar, 1); s_nobs = 0; }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&s_mbar, 2 * K1_BYTES);
        tma_load_1d(s_ascii, seq + FRONT_PAD + t0 - K1_LEAD, K1_BYTES, &s_mbar);
        tma_load_1d(s_qual, qual + FRONT_PAD + t0 - K1_LEAD, K1_BYTES, &s_mbar);
    }
    if (tid < K1_G32) s_start[tid] = start_bits[(t0 + FRONT_PAD - K1_LEAD) / 32 + tid];
    mbar_wait(&s_mbar, 0);

    // ---- pack bases and the two masks: invalid base (N; + low quality when strict) and low quality
    for (int g = tid; g < K1_G32; g += K1_THREADS) {
        const uint4* src = reinterpret_cast<const uint4*>(s_ascii + 32 * g);
        const uint4* qs = reinterpret_cast<const uint4*>(s_qual + 32 * g);
        uint4 a = src[0], b = src[1], qa = qs[0], qb = qs[1];
        uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t q[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
        uint32_t le0 = 0, le1 = 0, bad = 0, qbad = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            le0 |= pack4(w[i]) << (8 * i);
            le1 |= pack4(w[4 + i]) << (8 * i);
            bad |= bad4(w[i]) << (4 * i);
            bad |= bad4(w[4 + i]) << (16 + 4 * i);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) qbad |= badq4(q[i], min_qual) << (4 * i);
        if (qmode == 2) bad |= qbad;                                  // QualFilter::Strict (split_kmer.rs:98-101)
        s_le[2 * g] = le0; s_le[2 * g + 1] = le1;
        s_be[2 * g] = le_to_be(le0); s_be[2 * g + 1] = le_to_be(le1);
        s_bad[g] = bad; s_qbad[g] = qbad;
    }
    if (tid < 2) { s_le[2 * K1_G32 + tid] = 0; s_be[2 * K1_G32 + tid] = 0; }
    __syncthreads();

    // ---- window validity + middle-base quality (split_kmer.rs:328-339) + the sample's own position range
    const int h = (k - 1) >> 1;
    if (tid < K1_T / 32) {
        const int j = tid;
        B128 bad = {(uint64_t)s_bad[j] | ((uint64_t)s_bad[j + 1] << 32), (uint64_t)s_bad[j + 2] | ((uint64_t)s_bad[j + 3] << 32)};
        B128 st = {(uint64_t)s_start[j] | ((uint64_t)s_start[j + 1] << 32), (uint64_t)s_start[j + 2] | ((uint64_t)s_start[j + 3] << 32)};
        B128 bad_any = slide_or(bad, k);
        B128 st_any = shr128(slide_or(st, k - 1), 1);
        B128 bad_prev = {bad.lo << 1, (bad.hi << 1) | (bad.lo >> 63)};
        B128 run_start = or128(st, bad_prev);
        B128 rec_end = shr128(st, k);
        uint64_t not_ok = bad_any.lo | st_any.lo | (run_start.lo & rec_end.lo);
        if (qmode != 0) {                                             // Middle or Strict: quality of base s + h
            B128 qb = {(uint64_t)s_qbad[j] | ((uint64_t)s_qbad[j + 1] << 32), (uint64_t)s_qbad[j + 2] | ((uint64_t)s_qbad[j + 3] << 32)};
            not_ok |= shr128(qb, h).lo;
        }
        uint32_t ok = ~(uint32_t)(not_ok >> 32);
        const uint64_t p0 = t0 + 32ull * j;                           // keep positions in [lo, hi)
        if (p0 + 32 <= lo || p0 >= hi) ok = 0;
        else {
            if (p0 < lo) ok &= ~((1u << (uint32_t)(lo - p0)) - 1u);
            if (p0 + 32 > hi) ok &= (1u << (uint32_t)(hi - p0)) - 1u;
        }
        if (MODE == RQ_EMIT) ok &= mask[blockIdx.x * (K1_T / 32) + j];   // accepted observations only
        s_ok[j] = ok;
        s_pre[j] = __popc(ok);
        if (MODE == RQ_MASK && ok) atomicOr(&mask[blockIdx.x * (K1_T / 32) + j], ok);   // tiles at sample borders are shared
    }
    __syncthreads();
    if (MODE == RQ_MASK) return;

    const int lane = tid & 31, warp = tid >> 5;
    if (MODE == RQ_OBS) {
        // ---- ntHash by prefix XOR: X[r] = rotr(H[e_r], r), Y[r] = rotl(H[e_r ^ 2], r) over staged indices r
        uint64_t* GX = reinterpret_cast<uint64_t*>(dyn);
        uint64_t* GY = GX + K1_BYTES + 1;
        constexpr int PER = (K1_BYTES + K1_THREADS - 1) / K1_THREADS;     // 17
        const int r0 = tid * PER;
        uint64_t xs[PER], ys[PER], tx = 0, ty = 0;
#pragma unroll
        for (int i = 0; i < PER; i++) {
            const int r = r0 + i;
            uint64_t x = 0, y = 0;
            if (r < K1_BYTES) {
                const uint32_t e = (s_le[r >> 4] >> (2 * (r & 15))) & 3u;
                x = rotr64(nt_seed(e), (uint32_t)r);
                y = rotl64(nt_seed(e ^ 2u), (uint32_t)r);
            }
            xs[i] = tx; ys[i] = ty;                                   // exclusive prefix inside the thread
            tx ^= x; ty ^= y;
        }
        uint64_t ix = tx, iy = ty;                                    // inclusive XOR scan over the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t ux = __shfl_up_sync(0xFFFFFFFFu, ix, o), uy = __shfl_up_sync(0xFFFFFFFFu, iy, o);
            if (lane >= o) { ix ^= ux; iy ^= uy; }
        }
        if (lane == 31) { s_wx[warp] = ix; s_wy[warp] = iy; }
        __syncthreads();
        uint64_t bx = ix ^ tx, by = iy ^ ty;                          // exclusive prefix of this thread within the warp
        for (int w = 0; w < warp; w++) { bx ^= s_wx[w]; by ^= s_wy[w]; }
#pragma unroll
        for (int i = 0; i < PER; i++) {
            const int r = r0 + i;
            if (r <= K1_BYTES) { GX[r] = bx ^ xs[i]; GY[r] = by ^ ys[i]; }
        }
        __syncthreads();
        uint32_t mine = 0;
        for (int j = warp; j < K1_T / 32; j += K1_THREADS / 32) {
            const uint64_t p = t0 + 32ull * j + lane;
            if (p < lo || p >= hi) continue;
            const int r = K1_LEAD + 32 * j + lane;
            uint64_t key = ~0ULL; uint32_t pos = 0xFFFFFFFFu;
            if ((s_ok[j] >> lane) & 1u) {
                const uint64_t fh = rotl64(GX[r + k] ^ GX[r], (uint32_t)(r + k - 1));   // nthash.rs:35-41
                uint64_t hash = fh;
                if (rc) { const uint64_t rh = rotr64(GY[r + k] ^ GY[r], (uint32_t)r); hash = min(fh, rh); }   // nthash.rs:43-52,70-76
                key = cheap_mix(hash);
                pos = (uint32_t)(p - tile0 * K1_T);
                mine++;
            }
            obs_rec[p - lo] = key == ~0ULL ? make_ulonglong2(~0ULL, ~0ULL) : make_ulonglong2((unsigned long long)pos, key);
        }
        if (mine) atomicAdd(&s_nobs, mine);
        __syncthreads();
        if (tid == 0 && s_nobs) atomicAdd(n_obs, (unsigned long long)s_nobs);
        return;
    }

    // ---- RQ_EMIT: split k-mer of every accepted position, compacted
    if (tid == 0) { uint32_t run = 0; for (int j = 0; j < K1_T / 32; j++) { uint32_t c = s_pre[j]; s_pre[j] = run; run += c; } s_pre[K1_T / 32] = run; }
    __syncthreads();
    const uint64_t mask2k = (1ULL << (2 * k)) - 1ULL;
    const uint64_t lm = (1ULL << (2 * h)) - 1ULL;
    const uint64_t aa = 0xAAAAAAAAAAAAAAAAULL & mask2k;
    const int sh = 2 * (lane & 15);
    const uint32_t* be = s_be + ((K1_LEAD + lane) >> 4);
    const uint32_t* le = s_le + ((K1_LEAD + lane) >> 4);
    uint64_t* out = rec_c + tile_off[blockIdx.x] * (k > 31 ? 2 : 1);
    if (k > 31) {                                                     // two-word records (see k1_extract.cu, WIDE)
        typedef unsigned __int128 u128;
        const u128 one = 1;
        const u128 mask2k_w = (one << (2 * k)) - 1, lm_w = (one << (2 * h)) - 1;
        const u128 aa_w = ((((u128)0xAAAAAAAAAAAAAAAAULL) << 64) | 0xAAAAAAAAAAAAAAAAULL) & mask2k_w;
        for (int j = warp; j < K1_T / 32; j += K1_THREADS / 32) {
            const uint32_t ok = s_ok[j];
            if (!((ok >> lane) & 1u)) continue;
            const uint32_t b0 = be[2 * j], b1 = be[2 * j + 1], b2 = be[2 * j + 2], b3 = be[2 * j + 3], b4 = be[2 * j + 4];
            const uint64_t thi = ((uint64_t)__funnelshift_l(b1, b0, sh) << 32) | __funnelshift_l(b2, b1, sh);
            const uint64_t tlo = ((uint64_t)__funnelshift_l(b3, b2, sh) << 32) | __funnelshift_l(b4, b3, sh);
            const u128 F = ((((u128)thi) << 64) | tlo) >> (128 - 2 * k);
            u128 key = (F & lm_w) | ((F >> 2) & ~lm_w);
            uint32_t mid = (uint32_t)(F >> (2 * h)) & 3u;
            if (rc) {
                const uint32_t l0 = le[2 * j], l1 = le[2 * j + 1], l2 = le[2 * j + 2], l3 = le[2 * j + 3], l4 = le[2 * j + 4];
                const uint64_t vlo = ((uint64_t)__funnelshift_r(l1, l2, sh) << 32) | __funnelshift_r(l0, l1, sh);
                const uint64_t vhi = ((uint64_t)__funnelshift_r(l3, l4, sh) << 32) | __funnelshift_r(l2, l3, sh);
                const u128 R = (((((u128)vhi) << 64) | vlo) & mask2k_w) ^ aa_w;
                const u128 rkey = (R & lm_w) | ((R >> 2) & ~lm_w);
                if (key > rkey) { key = rkey; mid ^= 2u; }
            }
            const u128 rec = (key << 4) | (u128)(1u << mid);
            const uint32_t idx = s_pre[j] + __popc(ok & ((1u << lane) - 1u));
            out[2 * idx] = (uint64_t)rec; out[2 * idx + 1] = (uint64_t)(rec >> 64);
        }
        return;
    }
    for (int j = warp; j < K1_T / 32; j += K1_THREADS / 32) {
        const uint32_t ok = s_ok[j];
        if (!((ok >> lane) & 1u)) continue;
        const uint32_t b0 = be[2 * j], b1 = be[2 * j + 1], b2 = be[2 * j + 2];
        const uint64_t top = ((uint64_t)__funnelshift_l(b1, b0, sh) << 32) | __funnelshift_l(b2, b1, sh);
        const uint64_t F = top >> (64 - 2 * k);
        uint64_t key = (F & lm) | ((F >> 2) & ~lm);
        uint32_t mid = (uint32_t)(F >> (2 * h)) & 3u;
        if (rc) {
            const uint32_t l0 = le[2 * j], l1 = le[2 * j + 1], l2 = le[2 * j + 2];
            const uint64_t V = (((uint64_t)__funnelshift_r(l1, l2, sh) << 32) | __funnelshift_r(l0, l1, sh)) & mask2k;
            const uint64_t R = V ^ aa;
            const uint64_t rkey = (R & lm) | ((R >> 2) & ~lm);
            if (key > rkey) { key = rkey; mid ^= 2u; }
        }
        out[s_pre[j] + __popc(ok & ((1u << lane) - 1u))] = (key << 4) | (1ull << mid);
    }
}

// accepted windows per tile (popcount of the tile's 128 mask words)
__global__ void tile_popc_kernel(const uint32_t* __restrict__ mask, uint32_t n_tiles, uint32_t* __restrict__ cnt) {
    const uint32_t tile = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (tile >= n_tiles) return;
    uint32_t c = 0;
    for (int j = lane; j < K1_T / 32; j += 32) c += __popc(mask[(size_t)tile * (K1_T / 32) + j]);
    c = __reduce_add_sync(0xFFFFFFFFu, c);
    if (lane == 0) cnt[tile] = c;
}

// entries (sort128.cu: one per distinct hash, written by the warp that sorted the hash's bucket) sorted by (word, first
// position): the thread at the head of a word's segment replays that word of the bloom filter (bloom_filter.rs:77-86) in
// first-occurrence order (a word holds the ~30-40 hashes that map to it)
__global__ void q_bloom_kernel(const ulonglong2* __restrict__ drec, const ulonglong2* __restrict__ ent, uint64_t n_distinct,
                               uint8_t* __restrict__ fp_flag) {
    const uint64_t d = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_distinct) return;
    const uint32_t word = (uint32_t)(drec[d].y >> 32);
    if (d > 0 && (uint32_t)(drec[d - 1].y >> 32) == word) return;
    uint64_t state = 0;
    for (uint64_t e = d; e < n_distinct && (uint32_t)(drec[e].y >> 32) == word; e++) {
        const uint64_t f = fingerprint(cheap_unmix(ent[(uint32_t)drec[e].x].x));
        const bool seen = (state & f) == f;
        if (!seen) state |= f;
        fp_flag[e] = seen;
    }
}

// which observations of every distinct k-mer reach the dictionary (bloom_filter.rs:116-148).  3 <= min_count < 65535: exactly
// one, number min_count - (bloom false positive ? 1 : 0), whose position the entry carries.  min_count == 2 / 65535: every
// observation from there on, read from the sorted observations; entries that accept more than Q_LONG of them are left to
// q_accept_long_kernel.
constexpr uint64_t Q_LONG = 64;
__device__ __forceinline__ void accept_range(uint64_t c, uint64_t fp, uint32_t min_count, uint64_t& j_lo, uint64_t& j_hi) {
    if (min_count == 2) { j_lo = fp ? 1 : 2; j_hi = c; }
    else {
        j_lo = j_hi = (uint64_t)min_count - fp;                       // count after observation j is j + fp
        if (min_count == 65535) j_hi = c;                             // saturating_add keeps the count at u16::MAX
        if (j_lo < 2 - fp) j_lo = 2 - fp;
    }
    if (j_hi > c) j_hi = c;
}
__global__ void q_accept_one_kernel(const ulonglong2* __restrict__ drec, const uint8_t* __restrict__ fp_flag, const ulonglong2* __restrict__ ent,
                                    uint64_t n_distinct, uint32_t* __restrict__ mask) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_distinct) return;
    const uint64_t pp = ent[(uint32_t)drec[t].x].y;
    const uint32_t p = fp_flag[t] ? (uint32_t)pp : (uint32_t)(pp >> 32);
    if (p != 0xFFFFFFFFu) atomicOr(&mask[p >> 5], 1u << (p & 31));
}
__global__ void q_accept_kernel(const ulonglong2* __restrict__ drec, const uint8_t* __restrict__ fp_flag, const ulonglong2* __restrict__ obs,
                                const uint64_t* __restrict__ run_start, uint64_t n_distinct, uint32_t min_count, uint32_t* __restrict__ mask,
                                uint32_t* __restrict__ long_list, unsigned long long* __restrict__ n_long) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_distinct) return;
    const uint64_t i0 = run_start[(uint32_t)drec[t].x], c = drec[t].x >> 32;
    uint64_t j_lo, j_hi;                                              // accepted 1-based observation numbers [j_lo, j_hi]
    accept_range(c, fp_flag[t], min_count, j_lo, j_hi);
    if (j_hi >= j_lo && j_hi - j_lo >= Q_LONG) { long_list[atomicAdd(n_long, 1ULL)] = (uint32_t)t; return; }
    for (uint64_t j = j_lo; j <= j_hi; j++) {
        const uint32_t p = (uint32_t)obs[i0 + j - 1].x;
        atomicOr(&mask[p >> 5], 1u << (p & 31));
    }
}
__global__ void __launch_bounds__(256)
q_accept_long_kernel(const ulonglong2* __restrict__ drec, const uint8_t* __restrict__ fp_flag, const ulonglong2* __restrict__ obs,
                     const uint64_t* __restrict__ run_start, const uint32_t* __restrict__ long_list, uint32_t min_count, uint32_t* __restrict__ mask) {
    const uint64_t t = long_list[blockIdx.x];
    const uint64_t i0 = run_start[(uint32_t)drec[t].x], c = drec[t].x >> 32;
    uint64_t j_lo, j_hi;
    accept_range(c, fp_flag[t], min_count, j_lo, j_hi);
    for (uint64_t j = j_lo + threadIdx.x; j <= j_hi; j += 256) {
        const uint32_t p = (uint32_t)obs[i0 + j - 1].x;
        atomicOr(&mask[p >> 5], 1u << (p & 31));
    }
}

// ------------------------------------------------------------------------------------------------ host driver
struct ArenaMark { std::vector<size_t> used; };
static ArenaMark arena_mark() { ArenaMark m; for (auto& s : arena().slabs) m.used.push_back(s.used); return m; }
static void arena_rewind(const ArenaMark& m) {
    auto& sl = arena().slabs;
    for (size_t i = 0; i < sl.size(); i++) sl[i].used = i < m.used.size() ? m.used[i] : 0;
}

// prefix bits that put ~512 of a sample's accepted windows into one bucket of unique_sample_records
static int unique_prefix_bits(uint64_t n, uint32_t k) {
    int P = 0;
    while ((n >> P) > 512 && P < 2 * ((int)k - 1) && P < 22) P++;
    return P;
}

void build_array_from_reads(const SeqBatch& b, uint32_t k, bool rc, uint16_t min_count_reads, uint8_t min_qual_reads, int qual_filter_reads, Array& out) {
    Context& c = ctx();
    const uint64_t N = b.n_samples;
    const int rw = k > 31 ? 2 : 1;                                    // words per record
    const size_t obs_smem = 2 * (size_t)(K1_BYTES + 1) * 8;
    SKA_CUDA(cudaFuncSetAttribute(k1q_kernel<RQ_OBS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)obs_smem));
    std::vector<DevBuf<uint64_t>> kept(N);                            // every sample's distinct records, outside the arena
    std::vector<uint64_t> coff(N + 1, 0);
    unsigned long long* d_counter = arena().take_n<unsigned long long>(2);

    // one sample at a time: the filter's working set is 32 B per base, the batch may hold hundreds of GB of reads
    for (uint64_t s = 0; s < N; s++) {
        const uint64_t lo = b.h_sample_off[s], hi = b.h_sample_off[s + 1];
        if (hi == lo) throw Error("sample " + std::to_string(s) + " has no valid sequence");   // ska_dict.rs:374-376
        // a FASTA sample among FASTQ ones: every valid window is kept, no quality test, no count filter (ska_dict.rs:155,167 `!is_reads ||`)
        const bool is_reads = b.h_is_reads.empty() || b.h_is_reads[s];
        const uint16_t min_count = is_reads ? min_count_reads : (uint16_t)1;
        const uint8_t min_qual = is_reads ? min_qual_reads : (uint8_t)0;
        const int qual_filter = is_reads ? qual_filter_reads : 0;
        const ArenaMark sample_mark = arena_mark();
        const uint64_t tile_lo = lo / K1_T, n_tiles = (hi - 1) / K1_T - tile_lo + 1;
        const size_t n_words = n_tiles * (K1_T / 32);
        uint32_t* mask = arena().take_n<uint32_t>(n_words);
        timer().begin("reads_hash");
        dev_fill(mask, 0, n_words * 4);
        if (min_count <= 1) {                                         // bloom_filter.rs:121: no filtering
            SKA_LAUNCH(k1q_kernel<RQ_MASK>, (unsigned)n_tiles, K1_THREADS, 0, b.seq.p, b.qual.p, b.start_bits.p, lo, hi, tile_lo,
                       (int)k, rc ? 1 : 0, (uint32_t)min_qual, qual_filter, mask, (ulonglong2*)nullptr,
                       (unsigned long long*)nullptr, (const uint64_t*)nullptr, (uint64_t*)nullptr);
        } else {
            const ArenaMark mark = arena_mark();
            const uint64_t n_s = hi - lo;
            if (n_s >= 0xFFFFFFFFull) throw Error("a read set of 2^32 bases or more per sample is beyond the 32-bit positions of the count filter");
            const bool all_later = min_count == 2 || min_count == 65535;      // every observation past the threshold is accepted
            uint64_t* obs_rec = arena().take_n<uint64_t>(2 * n_s);            // 16-byte records (position, mixed hash)
            uint64_t* obs_tmp = arena().take_n<uint64_t>(2 * n_s);
            uint64_t* ent = arena().take_n<uint64_t>(2 * n_s);                // one 16-byte entry per distinct hash (at most n_s)
            uint64_t* run_start = all_later ? arena().take_n<uint64_t>(n_s) : nullptr;
            dev_fill(d_counter, 0, 16);
            SKA_LAUNCH(k1q_kernel<RQ_OBS>, (unsigned)n_tiles, K1_THREADS, obs_smem, b.seq.p, b.qual.p, b.start_bits.p, lo, hi, tile_lo,
                       (int)k, rc ? 1 : 0, (uint32_t)min_qual, qual_filter, mask, reinterpret_cast<ulonglong2*>(obs_rec), d_counter,
                       (const uint64_t*)nullptr, (uint64_t*)nullptr);
            timer().end();
            // sort by (hash, position) and one entry per distinct hash: count, bloom word, first position, candidate positions
            uint64_t n_obs = 0;
            uint64_t* drec = nullptr;
            const ulonglong2* obs = reinterpret_cast<const ulonglong2*>(sort_observations(obs_rec, n_s, obs_tmp, obs_rec, &n_obs, &drec, ent, run_start,
                                                                                          d_counter + 1, (uint32_t)min_count, (uint32_t)BLOOM_WORDS, all_later));
            unsigned long long h_cnt[2];
            SKA_CUDA(cudaMemcpyAsync(h_cnt, d_counter, 16, cudaMemcpyDeviceToHost, c.stream));
            SKA_CUDA(cudaStreamSynchronize(c.stream));
            timer().begin("count_filter");
            const uint64_t n_d = h_cnt[1];
            if (n_d) {
                // entries by (bloom word, first position): the order the sequential filter met them in, word by word
                uint64_t* dtmp = arena().take_n<uint64_t>(2 * n_d);
                uint64_t n_d2 = 0;
                const ulonglong2* dsorted = reinterpret_cast<const ulonglong2*>(sort_records_wide(drec, n_d, 54, dtmp, drec, &n_d2));
                if (n_d2 != n_d) throw Error("internal: distinct entries lost in the sort");
                uint8_t* fp_flag = arena().take_n<uint8_t>(n_d);
                const ulonglong2* ent2 = reinterpret_cast<const ulonglong2*>(ent);
                SKA_LAUNCH(q_bloom_kernel, (unsigned)((n_d + 255) / 256), 256, 0, dsorted, ent2, n_d, fp_flag);
                if (!all_later) {
                    SKA_LAUNCH(q_accept_one_kernel, (unsigned)((n_d + 255) / 256), 256, 0, dsorted, fp_flag, ent2, n_d, mask);
                } else {
                    uint32_t* long_list = arena().take_n<uint32_t>(n_d);
                    unsigned long long* n_long = arena().take_n<unsigned long long>(1);
                    dev_fill(n_long, 0, 8);
                    SKA_LAUNCH(q_accept_kernel, (unsigned)((n_d + 255) / 256), 256, 0, dsorted, fp_flag, obs, (const uint64_t*)run_start, n_d,
                               (uint32_t)min_count, mask, long_list, n_long);
                    unsigned long long h_long = 0;
                    SKA_CUDA(cudaMemcpyAsync(&h_long, n_long, 8, cudaMemcpyDeviceToHost, c.stream));
                    SKA_CUDA(cudaStreamSynchronize(c.stream));
                    if (h_long) SKA_LAUNCH(q_accept_long_kernel, (unsigned)h_long, 256, 0, dsorted, fp_flag, obs, (const uint64_t*)run_start,
                                           (const uint32_t*)long_list, (uint32_t)min_count, mask);
                }
            }
            timer().end();
            SKA_CUDA(cudaStreamSynchronize(c.stream));
            arena_rewind(mark);
            timer().begin("mask_scan");
        }
        // compaction offsets of the accepted windows
        uint32_t* cnt = arena().take_n<uint32_t>(n_tiles);
        uint64_t* tile_off = arena().take_n<uint64_t>(n_tiles + 1);
        uint32_t* scratch = arena().take_n<uint32_t>(1);
        uint64_t accepted = 0;
        SKA_LAUNCH(tile_popc_kernel, (unsigned)((n_tiles + 7) / 8), 256, 0, mask, (uint32_t)n_tiles, cnt);
        launch_scan_counts(cnt, (uint32_t)n_tiles, tile_off, scratch);
        SKA_CUDA(cudaMemcpyAsync(&accepted, tile_off + n_tiles, 8, cudaMemcpyDeviceToHost, c.stream));
        timer().end();
        SKA_CUDA(cudaStreamSynchronize(c.stream));
        if (accepted == 0) throw Error("sample " + std::to_string(s) + " has no valid sequence");     // ska_dict.rs:374-376

        // the accepted windows' K1 records, then the distinct ones among them (dedup.cu)
        const ArenaMark emit_mark = arena_mark();
        int P = unique_prefix_bits(accepted, k);
        for (int attempt = 0;; attempt++) {
            uint64_t* rec_s = arena().take_n<uint64_t>(accepted * rw);
            timer().begin("k1_extract");
            SKA_LAUNCH(k1q_kernel<RQ_EMIT>, (unsigned)n_tiles, K1_THREADS, 0, b.seq.p, b.qual.p, b.start_bits.p, lo, hi, tile_lo, (int)k, rc ? 1 : 0,
                       (uint32_t)min_qual, qual_filter, mask, (ulonglong2*)nullptr, (unsigned long long*)nullptr, (const uint64_t*)tile_off, rec_s);
            timer().end();
            uint64_t n_u = 0;
            const uint64_t* uniq = unique_sample_records(rec_s, accepted, k, P, &n_u);
            if (n_u != ~0ULL) {
                kept[s].alloc(std::max<uint64_t>(n_u, 1) * rw);
                SKA_CUDA(cudaMemcpyAsync(kept[s].p, uniq, n_u * rw * 8, cudaMemcpyDeviceToDevice, c.stream));
                coff[s + 1] = coff[s] + n_u;
                break;
            }
            const int P_max = std::min(22, 2 * ((int)k - 1));
            if (P >= P_max || attempt > 8) throw Error("reads: a key range is too crowded for the per-sample set");
            P = std::min(P_max, P + 2);
            arena_rewind(emit_mark);
        }
        SKA_CUDA(cudaStreamSynchronize(c.stream));
        arena_rewind(sample_mark);
    }

    // all samples' distinct records back to back -> the merge
    uint64_t* rec_c = arena().take_n<uint64_t>(std::max<uint64_t>(coff[N], 1) * rw);
    for (uint64_t s = 0; s < N; s++) {
        if (coff[s + 1] > coff[s])
            SKA_CUDA(cudaMemcpyAsync(rec_c + coff[s] * rw, kept[s].p, (coff[s + 1] - coff[s]) * rw * 8, cudaMemcpyDeviceToDevice, c.stream));
    }
    uint64_t* d_coff = arena().take_n<uint64_t>(N + 1);
    SKA_CUDA(cudaMemcpyAsync(d_coff, coff.data(), (N + 1) * 8, cudaMemcpyHostToDevice, c.stream));
    SKA_CUDA(cudaStreamSynchronize(c.stream));
    kept.clear();
    if (k > 31) build_array_from_records_wide(rec_c, coff[N], d_coff, (uint32_t)N, 0, N, k, rc, out);
    else build_array_from_records(rec_c, coff[N], d_coff, (uint32_t)N, 0, N, k, rc, out);
}

}  // namespace ska
