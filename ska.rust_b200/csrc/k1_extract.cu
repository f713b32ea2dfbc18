// k1_extract.cu -- K1: split-k-mer extraction with canonical strand select (sm_100a).
//
// Replaces the per-base hot loop of the reference: SplitKmer::build / roll_fwd / update_rc /
// get_curr_kmer (src/ska_dict/split_kmer.rs:78-217, 281-295) driven by SkaDict::add_file_kmers
// (src/ska_dict.rs:118-180), with the 2-bit codec of src/ska_dict/bit_encoding.rs:34-54.
//
// Formulation: every window is a pure function of its k bases, so there is
// no sequential roll.  One CTA owns a tile of K1_T consecutive window-start positions of the
// batch's global position space:
//   1. TMA bulk copy (cp.async.bulk -> UBLKCP) of the tile's ASCII bytes + halo into shared memory;
//   2. pack: 32 bases/thread -> 2-bit words in BOTH orientations (LE: base i at bits 2i, BE: first
//      base most significant) + an "invalid base" bit mask (N/n test `b & 0xF == 14`, optionally
//      the strict quality test `qual-33 > min_qual`);
//   3. window-validity words by sliding-OR bit tricks over the invalid mask and the record-start
//      mask (windows may not contain an invalid base nor span a record boundary; a run of exactly
//      k valid bases that ends at its record's end yields nothing -- split_kmer.rs:89,121);
//   4. extract: lane l of a warp takes window 32j+l, so the 8-byte record stores are coalesced;
//      forward k-mer = funnel shift out of the BE stream, reverse complement = funnel shift out of
//      the LE stream XOR 0b10 per base (no bit reversal per window); key = k-mer with the middle
//      base squeezed out; canonical = min(key, rc_key), ties keep forward (split_kmer.rs:290).
// Output: one u64 per position, rec = key << 4 | 1 << middle_code, or REC_SENTINEL.
//
// Roofline: HBM-bound, 1 B read + 8 B written per base (DESIGN.md "K1").
#include "k1_helpers.cuh"

namespace ska {

uint64_t k1_tiles_ready(uint64_t positions_uploaded, int tail) {      // tiles whose staged bytes (tile + tail halo) lie below that position
    return positions_uploaded < (uint64_t)K1_T + tail ? 0 : (positions_uploaded - tail) / K1_T;
}
static_assert(K1_ASCII_TAIL == K1_TAIL && K1_PACKED_TAIL == K1P_TAIL, "tile tails");

bool k1_wide_available() { return true; }

size_t padded_positions(uint64_t n_pos) { return FRONT_PAD + ((n_pos + K1_T - 1) / K1_T) * K1_T + 256; }

// KT > 0: k fixed at compile time (all shifts/masks become immediates); KT == 0: runtime k.
// HIST: also count the records by the top `hist_bits` bits of the key -- the histogram of the first partition pass
// (partition.cu), which then never re-reads the records for it.  Counts go to shared memory first and to one of
// K1_HIST_COPIES global copies at the end of the tile (tiles are many: a single copy would serialise on its counters).
// PACKED: the batch holds host-packed bases instead of ASCII (skagpu_build_packed: 0.375 B per base across PCIe instead
// of 1): `seq` = 2-bit codes, 4 per byte in the kernel's own LE order, `qual` = the invalid-base bit mask; the tile is
// staged straight into the LE stream and the mask words, the pack phase disappears.
// WIDE: k > 31 -- keys of up to 124 bits (UInt for u128, bit_encoding.rs:239-303); a record is two words,
// (key << 4 | one-hot middle) as a 128-bit little-endian integer, all-ones = no window.
template <int KT, bool HAS_QUAL, bool HIST, bool PACKED, bool WIDE = false>
__global__ void __launch_bounds__(K1_THREADS)
k1_extract_kernel(const uint8_t* __restrict__ seq, const uint8_t* __restrict__ qual,
                  const uint32_t* __restrict__ start_bits, const uint64_t* __restrict__ sample_off,
                  uint32_t n_samples, uint64_t n_pos, int k_rt, int rc, uint32_t min_qual,
                  uint64_t* __restrict__ rec_out, uint16_t* __restrict__ sid_out, uint32_t* __restrict__ sample_valid,
                  uint32_t sid_base, unsigned long long* __restrict__ hist, int hist_shift, int hist_bits, uint32_t tile0) {
    __shared__ uint32_t s_hist[HIST ? (1 << K1_HIST_MAX_BITS) : 1];
    __shared__ __align__(128) uint8_t s_ascii[PACKED ? 16 : K1_BYTES];
    __shared__ __align__(128) uint8_t s_qual[HAS_QUAL ? K1_BYTES : 16];
    // packed batches are staged from K1P_LEAD positions before the tile (16-byte aligned for the bulk copy); the kernel's
    // own arrays start K1_LEAD positions before it, i.e. (K1P_LEAD - K1_LEAD) / 16 LE words / 32 mask words further on
    __shared__ __align__(128) uint32_t s_le_raw[PACKED ? K1P_BYTES / 16 : 2 * K1_G32 + 2];
    __shared__ __align__(128) uint32_t s_bad_raw[PACKED ? K1P_BYTES / 32 : K1_G32];
    __shared__ uint32_t s_be[2 * K1_G32 + 2];
    uint32_t* const s_le = s_le_raw + (PACKED ? (K1P_LEAD - K1_LEAD) / 16 : 0);
    uint32_t* const s_bad = s_bad_raw + (PACKED ? (K1P_LEAD - K1_LEAD) / 32 : 0);
    __shared__ uint32_t s_start[K1_G32];
    __shared__ uint32_t s_ok[K1_T / 32];
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ uint32_t s_sid[2];
    __shared__ uint32_t s_nvalid;

    const int k = KT ? KT : k_rt;
    const int tid = threadIdx.x;
    const uint64_t t0 = ((uint64_t)blockIdx.x + tile0) * K1_T;   // first window-start position of this tile
    // staged region covers positions [t0 - K1_LEAD, t0 + K1_T + K1_TAIL); position p lives at seq[FRONT_PAD + p]
    const uint8_t* gsrc = seq + FRONT_PAD + t0 - K1_LEAD;

    if (tid == 0) {
        mbar_init(&s_mbar, 1);
        s_nvalid = 0;
    }
    if (HIST) for (int i = tid; i < (1 << hist_bits); i += K1_THREADS) s_hist[i] = 0;
    __syncthreads();
    if (tid == 0) {
        if (PACKED) {
            mbar_expect_tx(&s_mbar, K1P_BYTES / 4 + K1P_BYTES / 8);
            tma_load_1d(s_le_raw, seq + (FRONT_PAD + t0 - K1P_LEAD) / 4, K1P_BYTES / 4, &s_mbar);
            tma_load_1d(s_bad_raw, qual + (FRONT_PAD + t0 - K1P_LEAD) / 8, K1P_BYTES / 8, &s_mbar);
        } else {
            mbar_expect_tx(&s_mbar, HAS_QUAL ? 2 * K1_BYTES : K1_BYTES);
            tma_load_1d(s_ascii, gsrc, K1_BYTES, &s_mbar);
            if (HAS_QUAL) tma_load_1d(s_qual, qual + FRONT_PAD + t0 - K1_LEAD, K1_BYTES, &s_mbar);
        }
        // sample ids of the first and last window of the tile (binary search over sample_off)
        uint64_t pf = t0, pl = min(t0 + K1_T - 1, n_pos - 1);
        uint32_t lo = 0, hi = n_samples;                    // find s with off[s] <= pf < off[s+1]
        while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (sample_off[m] <= pf) lo = m; else hi = m; }
        s_sid[0] = lo;
        hi = n_samples;
        while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (sample_off[m] <= pl) lo = m; else hi = m; }
        s_sid[1] = lo;
    }
    // record-start bits for the staged region (word w <-> positions t0 - 32 + 32w ...)
    if (tid < K1_G32) s_start[tid] = start_bits[(t0 + FRONT_PAD - K1_LEAD) / 32 + tid];
    mbar_wait(&s_mbar, 0);

    // ---- pack: 32 bases per thread (packed batches: only the BE copy of the stream is made)
    if (PACKED) {
        for (int g = tid; g < 2 * K1_G32; g += K1_THREADS) s_be[g] = le_to_be(s_le[g]);
    } else
    for (int g = tid; g < K1_G32; g += K1_THREADS) {
        const uint4* src = reinterpret_cast<const uint4*>(s_ascii + 32 * g);
        uint4 a = src[0], b = src[1];
        uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t le0 = 0, le1 = 0, bad = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            le0 |= pack4(w[i]) << (8 * i);
            le1 |= pack4(w[4 + i]) << (8 * i);
            bad |= bad4(w[i]) << (4 * i);
            bad |= bad4(w[4 + i]) << (16 + 4 * i);
        }
        if (HAS_QUAL) {
            const uint4* qs = reinterpret_cast<const uint4*>(s_qual + 32 * g);
            uint4 qa = qs[0], qb = qs[1];
            uint32_t q[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
            for (int i = 0; i < 8; i++) bad |= badq4(q[i], min_qual) << (4 * i);
        }
        s_le[2 * g] = le0; s_le[2 * g + 1] = le1;
        s_be[2 * g] = le_to_be(le0); s_be[2 * g + 1] = le_to_be(le1);
        s_bad[g] = bad;
    }
    if (tid < 2) { if (!PACKED) s_le[2 * K1_G32 + tid] = 0; s_be[2 * K1_G32 + tid] = 0; }
    __syncthreads();

    // ---- window validity: thread j -> windows at staged bits [32(j+1), 32(j+1)+32)
    if (tid < K1_T / 32) {
        const int j = tid;
        B128 bad = {(uint64_t)s_bad[j] | ((uint64_t)s_bad[j + 1] << 32), (uint64_t)s_bad[j + 2] | ((uint64_t)s_bad[j + 3] << 32)};
        B128 st = {(uint64_t)s_start[j] | ((uint64_t)s_start[j + 1] << 32), (uint64_t)s_start[j + 2] | ((uint64_t)s_start[j + 3] << 32)};
        uint32_t ok = 0xFFFFFFFFu;
        if ((bad.lo | bad.hi | st.lo | st.hi) != 0) {            // fast path: clean sequence, no boundary nearby
            B128 bad_any = slide_or(bad, k);                          // any invalid base in [n, n+k)
            B128 st_any = shr128(slide_or(st, k - 1), 1);             // any record start in (n, n+k)
            B128 bad_prev = {bad.lo << 1, (bad.hi << 1) | (bad.lo >> 63)};   // bad[n-1]
            B128 run_start = or128(st, bad_prev);                     // a valid run begins at n
            B128 rec_end = shr128(st, k);                             // record ends right after the window
            uint64_t not_ok = bad_any.lo | st_any.lo | (run_start.lo & rec_end.lo);   // split_kmer.rs:89,121
            ok = ~(uint32_t)(not_ok >> 32);
        }
        // windows at or beyond n_pos do not exist
        uint64_t p0 = t0 + 32ull * j;
        if (p0 >= n_pos) ok = 0;
        else if (p0 + 32 > n_pos) ok &= (1u << (uint32_t)(n_pos - p0)) - 1u;
        s_ok[j] = ok;
        if (ok) atomicAdd(&s_nvalid, (uint32_t)__popc(ok));
    }
    __syncthreads();

    // ---- extract: lane l of warp w takes windows 32j + l, j = w, w+8, ...
    const int lane = tid & 31, warp = tid >> 5;
    const int h = (k - 1) >> 1;
    const uint64_t mask2k = (1ULL << (2 * k)) - 1ULL;             // k <= 31
    const uint64_t lm = (1ULL << (2 * h)) - 1ULL;                 // lower-arm mask
    const uint64_t aa = 0xAAAAAAAAAAAAAAAAULL & mask2k;
    const uint32_t sid_first = s_sid[0], sid_last = s_sid[1];
    const bool uniform = sid_first == sid_last;
    const bool full = (t0 + K1_T <= n_pos);
    const int sh = 2 * (lane & 15);
    const uint32_t* be = s_be + ((K1_LEAD + lane) >> 4);          // word holding this lane's first base for j = 0
    const uint32_t* le = s_le + ((K1_LEAD + lane) >> 4);
    uint64_t* out = rec_out + t0 + lane;
    uint16_t* sout = sid_out ? sid_out + t0 + lane : nullptr;
    auto window = [&](int j) -> uint64_t {
        // forward k-mer (first base most significant) out of the BE stream
        const uint32_t b0 = be[2 * j], b1 = be[2 * j + 1], b2 = be[2 * j + 2];
        const uint64_t top = ((uint64_t)__funnelshift_l(b1, b0, sh) << 32) | __funnelshift_l(b2, b1, sh);
        const uint64_t F = top >> (64 - 2 * k);
        uint64_t key = (F & lm) | ((F >> 2) & ~lm);
        uint32_t mid = (uint32_t)(F >> (2 * h)) & 3u;
        if (rc) {
            // LE window (base at bit 0) XOR 0b10 per base = reverse complement, first base most significant
            const uint32_t l0 = le[2 * j], l1 = le[2 * j + 1], l2 = le[2 * j + 2];
            const uint64_t V = (((uint64_t)__funnelshift_r(l1, l2, sh) << 32) | __funnelshift_r(l0, l1, sh)) & mask2k;
            const uint64_t R = V ^ aa;
            const uint64_t rkey = (R & lm) | ((R >> 2) & ~lm);
            if (key > rkey) { key = rkey; mid ^= 2u; }            // ties keep forward (split_kmer.rs:290)
        }
        const uint32_t klo = (uint32_t)key, khi = (uint32_t)(key >> 32);
        uint64_t rec = ((uint64_t)__funnelshift_l(klo, khi, 4) << 32) | ((klo << 4) | (1u << mid));
        if (!((s_ok[j] >> lane) & 1u)) rec = REC_SENTINEL;
        else if (HIST) atomicAdd(&s_hist[(uint32_t)(rec >> hist_shift)], 1u);
        return rec;
    };
    if (WIDE) {
        typedef unsigned __int128 u128;
        const u128 one = 1;
        const u128 mask2k_w = (one << (2 * k)) - 1;                  // k <= 63
        const u128 lm_w = (one << (2 * h)) - 1;
        const u128 aa_w = ((((u128)0xAAAAAAAAAAAAAAAAULL) << 64) | 0xAAAAAAAAAAAAAAAAULL) & mask2k_w;
        ulonglong2* out2 = reinterpret_cast<ulonglong2*>(rec_out) + t0 + lane;
        for (int j = warp; j < K1_T / 32; j += K1_THREADS / 32) {
            const uint64_t p = t0 + 32ull * j + lane;
            if (p >= n_pos) break;
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
                if (key > rkey) { key = rkey; mid ^= 2u; }            // ties keep forward (split_kmer.rs:290)
            }
            const uint64_t klo = (uint64_t)key, khi = (uint64_t)(key >> 64);
            unsigned long long rlo = (klo << 4) | (uint64_t)(1u << mid), rhi = (khi << 4) | (klo >> 60);
            const bool okw = (s_ok[j] >> lane) & 1u;
            if (!okw) { rlo = ~0ULL; rhi = ~0ULL; }
            else if (HIST) {
                const uint32_t d = hist_shift >= 64 ? (uint32_t)(rhi >> (hist_shift - 64))
                                                    : (uint32_t)((rlo >> hist_shift) | (hist_shift ? rhi << (64 - hist_shift) : 0ULL));
                atomicAdd(&s_hist[d], 1u);
            }
            out2[32 * j] = make_ulonglong2(rlo, rhi);
            if (!uniform && sample_valid && okw) {
                uint32_t lo = sid_first, hi = sid_last + 1;
                while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (sample_off[m] <= p) lo = m; else hi = m; }
                atomicAdd(&sample_valid[lo], 1u);
            }
        }
    } else
    if (full && uniform) {                                        // interior tile of one sample: no per-window checks
        if (sout) {
#pragma unroll 4
            for (int j = warp; j < K1_T / 32; j += K1_THREADS / 32) { out[32 * j] = window(j); sout[32 * j] = (uint16_t)(sid_first + sid_base); }
        } else {
#pragma unroll 4
            for (int j = warp; j < K1_T / 32; j += K1_THREADS / 32) out[32 * j] = window(j);
        }
    } else {
        for (int j = warp; j < K1_T / 32; j += K1_THREADS / 32) {
            const uint64_t p = t0 + 32ull * j + lane;
            if (p >= n_pos) break;
            const uint64_t rec = window(j);
            out[32 * j] = rec;
            uint32_t lo = sid_first, hi = sid_last + 1;           // tile straddles samples or the end of the batch
            while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (sample_off[m] <= p) lo = m; else hi = m; }
            if (!uniform && sample_valid) {
                // one atomic per sample and warp step, not per window: small genomes (30 kbp) put a sample boundary in every
                // seventh tile, and 256 counters took a million contended atomics per build
                const uint32_t act = __activemask();
                const uint32_t vmask = __ballot_sync(act, rec != REC_SENTINEL);
                const uint32_t peers = __match_any_sync(act, lo);
                const uint32_t cnt_v = __popc(peers & vmask);
                if (cnt_v && lane == __ffs(peers) - 1) atomicAdd(&sample_valid[lo], cnt_v);
            }
            if (sout) sout[32 * j] = (uint16_t)(lo + sid_base);
        }
    }
    // accepted windows per sample (a sample with none is an error, ska_dict.rs:374-376)
    if (tid == 0 && sample_valid && uniform && s_nvalid) atomicAdd(&sample_valid[sid_first], s_nvalid);
    if (HIST) {
        __syncthreads();
        unsigned long long* h = hist + ((size_t)(blockIdx.x % K1_HIST_COPIES) << hist_bits);
        for (int i = tid; i < (1 << hist_bits); i += K1_THREADS) {
            const uint32_t cnt = s_hist[i];
            if (cnt) atomicAdd(&h[i], (unsigned long long)cnt);
        }
    }
}

template <int KT>
static void launch_k1_kt(const SeqBatch& b, uint32_t k, bool rc, bool strict_qual, uint8_t min_qual, uint64_t* rec_out,
                         uint16_t* sid_out, uint32_t* sample_valid, uint32_t sid_base, unsigned long long* hist, int hist_shift, int hist_bits,
                         uint64_t tile_begin, uint64_t tile_end) {
    const uint64_t n_tiles = (b.n_pos + K1_T - 1) / K1_T;
    tile_end = std::min(tile_end, n_tiles);
    if (tile_begin >= tile_end) return;
    const unsigned grid = (unsigned)(tile_end - tile_begin);
    const uint32_t tile0 = (uint32_t)tile_begin;
    unsigned long long* nohist = nullptr;
    if (k > 31) {
        if (KT != 0 || b.packed || b.has_qual) throw Error("internal: wide extraction takes ASCII assemblies");
        // the histogram-fused instantiation of the wide kernel came back from the B200 with bit 61 of some keys cleared (same
        // source, the plain instantiation is bit-exact against the reference semantics at k = 33..63): callers count wide histograms in the partition
        if (hist) throw Error("internal: the wide extraction kernel does not count the partition histogram");
        if (hist)
            SKA_LAUNCH((k1_extract_kernel<0, false, true, false, true>), grid, K1_THREADS, 0, b.seq.p, (const uint8_t*)nullptr, b.start_bits.p,
                       b.sample_off.p, (uint32_t)b.n_samples, b.n_pos, (int)k, rc ? 1 : 0, 0u, rec_out, sid_out, sample_valid, sid_base, hist,
                       hist_shift, hist_bits, tile0);
        else
            SKA_LAUNCH((k1_extract_kernel<0, false, false, false, true>), grid, K1_THREADS, 0, b.seq.p, (const uint8_t*)nullptr, b.start_bits.p,
                       b.sample_off.p, (uint32_t)b.n_samples, b.n_pos, (int)k, rc ? 1 : 0, 0u, rec_out, sid_out, sample_valid, sid_base, nohist,
                       0, 0, tile0);
    } else if (b.packed) {
        if (hist)
            SKA_LAUNCH((k1_extract_kernel<KT, false, true, true>), grid, K1_THREADS, 0, b.seq.p, b.qual.p, b.start_bits.p, b.sample_off.p,
                       (uint32_t)b.n_samples, b.n_pos, (int)k, rc ? 1 : 0, 0u, rec_out, sid_out, sample_valid, sid_base, hist, hist_shift, hist_bits, tile0);
        else
            SKA_LAUNCH((k1_extract_kernel<KT, false, false, true>), grid, K1_THREADS, 0, b.seq.p, b.qual.p, b.start_bits.p, b.sample_off.p,
                       (uint32_t)b.n_samples, b.n_pos, (int)k, rc ? 1 : 0, 0u, rec_out, sid_out, sample_valid, sid_base, nohist, 0, 0, tile0);
    } else if (b.has_qual && strict_qual) {
        SKA_LAUNCH((k1_extract_kernel<KT, true, false, false>), grid, K1_THREADS, 0, b.seq.p, b.qual.p, b.start_bits.p, b.sample_off.p,
                   (uint32_t)b.n_samples, b.n_pos, (int)k, rc ? 1 : 0, (uint32_t)min_qual, rec_out, sid_out, sample_valid, sid_base, nohist, 0, 0, tile0);
    } else if (hist) {
        SKA_LAUNCH((k1_extract_kernel<KT, false, true, false>), grid, K1_THREADS, 0, b.seq.p, (const uint8_t*)nullptr, b.start_bits.p,
                   b.sample_off.p, (uint32_t)b.n_samples, b.n_pos, (int)k, rc ? 1 : 0, (uint32_t)min_qual, rec_out, sid_out,
                   sample_valid, sid_base, hist, hist_shift, hist_bits, tile0);
    } else {
        SKA_LAUNCH((k1_extract_kernel<KT, false, false, false>), grid, K1_THREADS, 0, b.seq.p, (const uint8_t*)nullptr, b.start_bits.p,
                   b.sample_off.p, (uint32_t)b.n_samples, b.n_pos, (int)k, rc ? 1 : 0, (uint32_t)min_qual, rec_out, sid_out,
                   sample_valid, sid_base, nohist, 0, 0, tile0);
    }
}

void launch_k1(const SeqBatch& b, uint32_t k, bool rc, bool strict_qual, uint8_t min_qual, uint64_t* rec_out,
               uint16_t* sid_out, uint32_t* sample_valid, uint32_t sid_base, unsigned long long* hist, int hist_shift, int hist_bits,
               uint64_t tile_begin, uint64_t tile_end) {
    if (b.n_pos == 0) return;
    if (hist && (hist_bits < 0 || hist_bits > K1_HIST_MAX_BITS)) throw Error("internal: K1 histogram wider than its shared-memory table");
    // the default k of `ska build` (cli.rs:9) gets a compile-time specialisation
    if (k == 31) launch_k1_kt<31>(b, k, rc, strict_qual, min_qual, rec_out, sid_out, sample_valid, sid_base, hist, hist_shift, hist_bits, tile_begin, tile_end);
    else if (k == 15) launch_k1_kt<15>(b, k, rc, strict_qual, min_qual, rec_out, sid_out, sample_valid, sid_base, hist, hist_shift, hist_bits, tile_begin, tile_end);   // BASELINE configs[4]
    else launch_k1_kt<0>(b, k, rc, strict_qual, min_qual, rec_out, sid_out, sample_valid, sid_base, hist, hist_shift, hist_bits, tile_begin, tile_end);
}

}  // namespace ska
