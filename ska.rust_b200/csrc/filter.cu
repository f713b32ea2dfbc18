// filter.cu -- K4: column-wise variant filter + order-preserving row compaction, and the
// (rows x samples) -> (samples x rows) transpose used to write the alignment.
//
// Replaces MergeSkaArray::filter + update_counts (src/merge_ska_array.rs:289-402, 139-163; caller
// apply_filters, src/generic_modes.rs:112-131) and the transpose in write_fasta
// (src/merge_ska_array.rs:499-503).  The reference walks rows on one thread and allocates a
// HashSet per row; here one warp classifies one row with ballots, a scan gives every surviving
// row its destination, and a second kernel moves rows (applying --ambig-mask on the way).
//
// Semantics kept exactly:
//  * filter_ambig_as_missing: recount non-missing as "not '-' and not ambiguous", drop rows with
//    zero count *and their keys* first (update_counts always compacts split_kmers);
//  * row kept iff count >= min_count and
//      NoFilter: always | NoConst: >1 distinct byte values ('-' counts unless ignore_const_gaps)
//      NoAmbig: no byte outside {A,C,G,T,U,-} case-insensitively
//      NoAmbigOrConst: >1 distinct byte values among {ACGTUacgtu} (+ '-' unless ignore_const_gaps)
//  * keys are compacted by the filter loop only if update_kmers; returned count = rows the filter
//    loop removed (rows dropped by update_counts are not counted).
// Roofline: HBM-bound, reads N+4 B per row twice, writes N B per surviving row.
#include "common.cuh"

namespace ska {

constexpr int FL_THREADS = 256;
constexpr int FL_ROWS = 128;   // rows per CTA (16 per warp)

__device__ __forceinline__ bool is_ambig_dev(uint32_t b) {   // bit_encoding.rs:58-61
    b |= 0x20u;
    return !(b == 'a' || b == 'c' || b == 'g' || b == 't' || b == 'u' || b == '-');
}
// bit index among the 11 byte values that matter to NoAmbigOrConst, or 32 if none
__device__ __forceinline__ uint32_t acgtu_bit(uint32_t b) {
    switch (b) {
        case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; case 'U': return 4;
        case 'a': return 5; case 'c': return 6; case 'g': return 7; case 't': return 8; case 'u': return 9;
        case '-': return 10; default: return 32;
    }
}

// flags[r]: bit0 = survives update_counts stage, bit1 = survives the filter as well
__global__ void __launch_bounds__(FL_THREADS)
filter_flags_kernel(const uint8_t* __restrict__ var, const uint32_t* __restrict__ cnt, uint64_t n_rows, uint32_t n_samples,
                    uint64_t min_count, int ambig_missing, int ftype, int ignore_gaps, uint8_t* __restrict__ flags,
                    uint32_t* __restrict__ new_cnt, uint32_t* __restrict__ cta_keep1, uint32_t* __restrict__ cta_keep2) {
    __shared__ uint32_t s_k1, s_k2;
    if (threadIdx.x == 0) { s_k1 = 0; s_k2 = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t r0 = (uint64_t)blockIdx.x * FL_ROWS;
    uint32_t k1 = 0, k2 = 0;
    for (int rr = warp; rr < FL_ROWS; rr += FL_THREADS / 32) {
        const uint64_t r = r0 + rr;
        if (r >= n_rows) break;
        if (!ambig_missing) {       // the stored count decides: rows below the threshold (or NoFilter) are never read
            const uint32_t c0 = cnt[r];
            if ((uint64_t)c0 < min_count || ftype == 0) {
                if (lane == 0) {
                    const bool kp = (uint64_t)c0 >= min_count;
                    flags[r] = (uint8_t)(1 | (kp ? 2 : 0)); new_cnt[r] = c0; k1 += 1; k2 += kp;
                }
                continue;
            }
        }
        const uint8_t* row = var + r * n_samples;
        uint32_t good = 0;          // not '-' and not ambiguous
        uint32_t any_ambig = 0;
        uint32_t setbits = 0;       // NoAmbigOrConst value set
        uint32_t first = 0xFFFFFFFFu;   // first considered byte value seen by this lane (NoConst)
        uint32_t differs = 0;
        for (uint32_t s = lane; s < n_samples; s += 32) {
            uint32_t b = row[s];
            bool amb = is_ambig_dev(b);
            any_ambig |= amb;
            good += (b != '-' && !amb);
            uint32_t bit = acgtu_bit(b);
            if (bit < 32) setbits |= 1u << bit;
            if (!ignore_gaps || b != '-') {
                if (first == 0xFFFFFFFFu) first = b; else differs |= (b != first);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            good += __shfl_xor_sync(0xFFFFFFFFu, good, o);
            setbits |= __shfl_xor_sync(0xFFFFFFFFu, setbits, o);
        }
        any_ambig = __any_sync(0xFFFFFFFFu, any_ambig);
        // NoConst: reference value = first considered byte of the lowest lane that has one
        uint32_t have = __ballot_sync(0xFFFFFFFFu, first != 0xFFFFFFFFu);
        uint32_t ref = have ? __shfl_sync(0xFFFFFFFFu, first, __ffs(have) - 1) : 0xFFFFFFFFu;
        differs |= (first != 0xFFFFFFFFu && first != ref);
        bool nonconst = __any_sync(0xFFFFFFFFu, differs);

        uint32_t count = ambig_missing ? good : cnt[r];
        bool keep1 = !ambig_missing || good > 0;
        bool keep2 = false;
        if (keep1 && (uint64_t)count >= min_count) {
            switch (ftype) {
                case 0: keep2 = true; break;
                case 1: keep2 = nonconst; break;
                case 2: keep2 = !any_ambig; break;
                default: {
                    uint32_t relevant = ignore_gaps ? 0x3FFu : 0x7FFu;
                    keep2 = __popc(setbits & relevant) > 1;
                }
            }
        }
        if (lane == 0) {
            flags[r] = (uint8_t)((keep1 ? 1 : 0) | (keep2 ? 2 : 0));
            new_cnt[r] = count;
            k1 += keep1; k2 += keep2;
        }
    }
    if (lane == 0) { atomicAdd(&s_k1, k1); atomicAdd(&s_k2, k2); }
    __syncthreads();
    if (threadIdx.x == 0) { cta_keep1[blockIdx.x] = s_k1; cta_keep2[blockIdx.x] = s_k2; }
}

// Same classification with ONE THREAD PER ROW and 16-byte loads (rows of n_samples % 16 == 0 bytes):
// ~10x fewer instructions than the warp-per-row kernel because nothing is reduced across lanes.  The
// 32 rows a warp touches are 32 separate 128-byte lines; each line is fetched once and its remaining
// sectors are served from L1 by the following loads of the same thread.
// per byte value: bits 0-10 = one-hot of the NoAmbigOrConst value set, bit 16 = ambiguous, bit 17 = "good"
// (neither '-' nor ambiguous)
__device__ __forceinline__ uint32_t class_entry(uint32_t b) {
    uint32_t bit = acgtu_bit(b);
    bool amb = is_ambig_dev(b);
    return (bit < 32 ? 1u << bit : 0u) | (amb ? 1u << 16 : 0u) | ((b != '-' && !amb) ? 1u << 17 : 0u);
}
// bit 7 of every byte of the result is set iff that byte of x equals the byte c (exact, no carries between bytes)
__device__ __forceinline__ uint32_t eq_bytes(uint32_t x, uint32_t c4) {
    uint32_t t = x ^ c4;
    return ~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t) & 0x80808080u;
}
// Row classification state.  Values are tracked as a SET: bits 0-10 for the eleven byte values that matter to
// the filters (A C G T U a c g t u -), plus the first byte value outside that set and whether a second,
// different one was seen -- enough to answer "more than one distinct value" without per-byte branches.
struct RowState {
    uint32_t good, any_ambig, setbits, other_first, other_differs;
    const uint32_t* lut;
    __device__ __forceinline__ void byte(uint32_t b) {
        uint32_t e = lut[b];
        any_ambig |= e >> 16 & 1u;
        good += e >> 17;
        setbits |= e & 0x7FFu;
        if (!(e & 0x7FFu)) { if (other_first == 0xFFFFFFFFu) other_first = b; else other_differs |= (b != other_first); }
    }
    __device__ __forceinline__ void word(uint32_t w) {
        // branch-free for words made of A, C, G, T and '-' (every lane of the warp is a different row, so a
        // data-dependent branch here would make the whole warp pay for the slowest row)
        const uint32_t a = eq_bytes(w, 0x41414141u), c = eq_bytes(w, 0x43434343u), g = eq_bytes(w, 0x47474747u),
                       t = eq_bytes(w, 0x54545454u), gap = eq_bytes(w, 0x2D2D2D2Du);
        const uint32_t acgt = a | c | g | t;
        if ((acgt | gap) == 0x80808080u) {
            good += __popc(acgt);
            setbits |= (a ? 1u : 0u) | (c ? 2u : 0u) | (g ? 4u : 0u) | (t ? 8u : 0u) | (gap ? 1u << 10 : 0u);
            return;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) byte((w >> (8 * q)) & 0xFFu);     // IUPAC codes, lower case, U: rare
    }
    __device__ __forceinline__ bool nonconst(bool ignore_gaps) const {
        uint32_t n = __popc(setbits & (ignore_gaps ? 0x3FFu : 0x7FFu)) + (other_first != 0xFFFFFFFFu) + (other_differs != 0);
        return n > 1;
    }
};

__global__ void __launch_bounds__(FL_ROWS)
filter_flags_rows_kernel(const uint8_t* __restrict__ var, const uint32_t* __restrict__ cnt, uint64_t n_rows, uint32_t n_samples,
                         uint64_t min_count, int ambig_missing, int ftype, int ignore_gaps, uint8_t* __restrict__ flags,
                         uint32_t* __restrict__ new_cnt, uint32_t* __restrict__ cta_keep1, uint32_t* __restrict__ cta_keep2) {
    __shared__ uint32_t s_k1, s_k2;
    __shared__ uint32_t s_lut[256];
    if (threadIdx.x == 0) { s_k1 = 0; s_k2 = 0; }
    for (uint32_t b = threadIdx.x; b < 256; b += FL_ROWS) s_lut[b] = class_entry(b);
    __syncthreads();
    const uint64_t r = (uint64_t)blockIdx.x * FL_ROWS + threadIdx.x;
    bool keep1 = false, keep2 = false;
    const uint32_t c_stored = r < n_rows ? cnt[r] : 0u;
    // the stored count decides rows below the threshold (and NoFilter): they are never read
    const bool by_count = !ambig_missing && ((uint64_t)c_stored < min_count || ftype == 0);
    if (r < n_rows && by_count) {
        keep1 = true; keep2 = (uint64_t)c_stored >= min_count;
        flags[r] = (uint8_t)(1 | (keep2 ? 2 : 0));
        new_cnt[r] = c_stored;
    } else if (r < n_rows && !ambig_missing && ftype == 1 && !ignore_gaps) {
        // NoConst where '-' counts as a value: "more than one distinct byte" == "some byte differs from the first"
        const uint4* row = reinterpret_cast<const uint4*>(var + r * n_samples);
        const uint32_t n16 = n_samples >> 4;
        uint32_t diff = 0, pat = 0;
        for (uint32_t c0 = 0; c0 < n16; c0 += 8) {
            uint4 v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) if (c0 + j < n16) v[j] = __ldg(&row[c0 + j]);
            if (c0 == 0) pat = (v[0].x & 0xFFu) * 0x01010101u;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (c0 + j >= n16) break;
                diff |= (v[j].x ^ pat) | (v[j].y ^ pat) | (v[j].z ^ pat) | (v[j].w ^ pat);
            }
            if (diff) break;
        }
        keep1 = true; keep2 = diff != 0;
        flags[r] = (uint8_t)(1 | (keep2 ? 2 : 0));
        new_cnt[r] = c_stored;
    } else if (r < n_rows) {
        const uint4* row = reinterpret_cast<const uint4*>(var + r * n_samples);
        RowState st = {0, 0, 0, 0xFFFFFFFFu, 0, s_lut};
        const uint32_t n16 = n_samples >> 4;
        for (uint32_t c0 = 0; c0 < n16; c0 += 8) {
            uint4 v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) if (c0 + j < n16) v[j] = __ldg(&row[c0 + j]);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (c0 + j >= n16) break;
                st.word(v[j].x); st.word(v[j].y); st.word(v[j].z); st.word(v[j].w);
            }
        }
        uint32_t count = ambig_missing ? st.good : c_stored;
        keep1 = !ambig_missing || st.good > 0;
        if (keep1 && (uint64_t)count >= min_count) {
            switch (ftype) {
                case 0: keep2 = true; break;
                case 1: keep2 = st.nonconst(ignore_gaps != 0); break;
                case 2: keep2 = !st.any_ambig; break;
                default: keep2 = __popc(st.setbits & (ignore_gaps ? 0x3FFu : 0x7FFu)) > 1;
            }
        }
        flags[r] = (uint8_t)((keep1 ? 1 : 0) | (keep2 ? 2 : 0));
        new_cnt[r] = count;
    }
    uint32_t b1 = __ballot_sync(0xFFFFFFFFu, keep1), b2 = __ballot_sync(0xFFFFFFFFu, keep2);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_k1, __popc(b1)); atomicAdd(&s_k2, __popc(b2)); }
    __syncthreads();
    if (threadIdx.x == 0) { cta_keep1[blockIdx.x] = s_k1; cta_keep2[blockIdx.x] = s_k2; }
}

// Wide rows (n_samples >= 512, multiple of 16): one WARP per row with 16-byte loads, so a row of N bytes is read in
// fully coalesced 512-byte steps; the per-lane states are merged with shuffles once per row.
__global__ void __launch_bounds__(FL_THREADS)
filter_flags_wide_kernel(const uint8_t* __restrict__ var, const uint32_t* __restrict__ cnt, uint64_t n_rows, uint32_t n_samples,
                         uint64_t min_count, int ambig_missing, int ftype, int ignore_gaps, uint8_t* __restrict__ flags,
                         uint32_t* __restrict__ new_cnt, uint32_t* __restrict__ cta_keep1, uint32_t* __restrict__ cta_keep2) {
    __shared__ uint32_t s_k1, s_k2;
    __shared__ uint32_t s_lut[256];
    if (threadIdx.x == 0) { s_k1 = 0; s_k2 = 0; }
    for (uint32_t b = threadIdx.x; b < 256; b += FL_THREADS) s_lut[b] = class_entry(b);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t r0 = (uint64_t)blockIdx.x * FL_ROWS;
    const uint32_t n16 = n_samples >> 4;
    uint32_t k1 = 0, k2 = 0;
    for (int rr = warp; rr < FL_ROWS; rr += FL_THREADS / 32) {
        const uint64_t r = r0 + rr;
        if (r >= n_rows) break;
        if (!ambig_missing) {       // the stored count decides: rows below the threshold (or NoFilter) are never read
            const uint32_t c0 = cnt[r];
            if ((uint64_t)c0 < min_count || ftype == 0) {
                if (lane == 0) {
                    const bool kp = (uint64_t)c0 >= min_count;
                    flags[r] = (uint8_t)(1 | (kp ? 2 : 0)); new_cnt[r] = c0; k1 += 1; k2 += kp;
                }
                continue;
            }
        }
        const uint4* row = reinterpret_cast<const uint4*>(var + r * n_samples);
        RowState st = {0, 0, 0, 0xFFFFFFFFu, 0, s_lut};
        for (uint32_t c0 = lane; c0 < n16; c0 += 128) {               // 4 loads in flight per lane
            uint4 v[4];
#pragma unroll
            for (int j = 0; j < 4; j++) if (c0 + 32 * j < n16) v[j] = __ldg(&row[c0 + 32 * j]);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (c0 + 32 * j >= n16) break;
                st.word(v[j].x); st.word(v[j].y); st.word(v[j].z); st.word(v[j].w);
            }
        }
        // merge the 32 lane states
        uint32_t good = __reduce_add_sync(0xFFFFFFFFu, st.good);
        uint32_t setbits = __reduce_or_sync(0xFFFFFFFFu, st.setbits);
        uint32_t any_ambig = __any_sync(0xFFFFFFFFu, st.any_ambig != 0);
        uint32_t have = __ballot_sync(0xFFFFFFFFu, st.other_first != 0xFFFFFFFFu);
        uint32_t ref = have ? __shfl_sync(0xFFFFFFFFu, st.other_first, __ffs(have) - 1) : 0xFFFFFFFFu;
        uint32_t odiff = __any_sync(0xFFFFFFFFu, st.other_differs != 0 || (st.other_first != 0xFFFFFFFFu && st.other_first != ref));
        uint32_t count = ambig_missing ? good : cnt[r];
        bool keep1 = !ambig_missing || good > 0;
        bool keep2 = false;
        if (keep1 && (uint64_t)count >= min_count) {
            switch (ftype) {
                case 0: keep2 = true; break;
                case 1: keep2 = (__popc(setbits & (ignore_gaps ? 0x3FFu : 0x7FFu)) + (have != 0) + (odiff != 0)) > 1; break;
                case 2: keep2 = !any_ambig; break;
                default: keep2 = __popc(setbits & (ignore_gaps ? 0x3FFu : 0x7FFu)) > 1;
            }
        }
        if (lane == 0) {
            flags[r] = (uint8_t)((keep1 ? 1 : 0) | (keep2 ? 2 : 0));
            new_cnt[r] = count;
            k1 += keep1; k2 += keep2;
        }
    }
    if (lane == 0) { atomicAdd(&s_k1, k1); atomicAdd(&s_k2, k2); }
    __syncthreads();
    if (threadIdx.x == 0) { cta_keep1[blockIdx.x] = s_k1; cta_keep2[blockIdx.x] = s_k2; }
}

// moves surviving rows/counts/keys to their compacted positions (order preserving)
__global__ void __launch_bounds__(FL_THREADS)
filter_scatter_kernel(const uint8_t* __restrict__ var, const uint32_t* __restrict__ new_cnt, const uint64_t* __restrict__ keys,
                      const uint8_t* __restrict__ flags, const uint64_t* __restrict__ off1, const uint64_t* __restrict__ off2,
                      uint64_t n_rows, uint64_t n_keys_in, uint32_t n_samples, int key_words, int key_mode /*0 none,1 stage1,2 stage2*/,
                      int mask_ambig, uint8_t* __restrict__ out_var, uint32_t* __restrict__ out_cnt,
                      uint64_t* __restrict__ out_keys) {
    __shared__ uint32_t s_pos1[FL_ROWS], s_pos2[FL_ROWS];
    const uint64_t r0 = (uint64_t)blockIdx.x * FL_ROWS;
    // in-CTA exclusive prefix of the two flag bits (FL_ROWS = 128 rows -> 4 warps of ballots)
    __shared__ uint32_t s_w1[4], s_w2[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t f = 0;
    if (threadIdx.x < FL_ROWS && r0 + threadIdx.x < n_rows) f = flags[r0 + threadIdx.x];
    if (threadIdx.x < FL_ROWS) {
        uint32_t b1 = __ballot_sync(0xFFFFFFFFu, f & 1), b2 = __ballot_sync(0xFFFFFFFFu, f & 2);
        uint32_t lt = (1u << lane) - 1u;
        s_pos1[threadIdx.x] = __popc(b1 & lt);
        s_pos2[threadIdx.x] = __popc(b2 & lt);
        if (lane == 0) { s_w1[warp] = __popc(b1); s_w2[warp] = __popc(b2); }
    }
    __syncthreads();
    if (threadIdx.x < FL_ROWS) {
        uint32_t a1 = 0, a2 = 0;
        for (int w = 0; w < warp; w++) { a1 += s_w1[w]; a2 += s_w2[w]; }
        s_pos1[threadIdx.x] += a1; s_pos2[threadIdx.x] += a2;
    }
    __syncthreads();
    const uint64_t base1 = off1[blockIdx.x], base2 = off2[blockIdx.x];
    // keys + counts: one thread per row
    if (threadIdx.x < FL_ROWS) {
        uint64_t r = r0 + threadIdx.x;
        if (r < n_rows) {
            if ((f & 2)) out_cnt[base2 + s_pos2[threadIdx.x]] = new_cnt[r];
            if (out_keys && r < n_keys_in) {
                bool kk = key_mode == 1 ? (f & 1) : (key_mode == 2 ? (f & 2) : false);
                uint64_t dstk = key_mode == 1 ? base1 + s_pos1[threadIdx.x] : base2 + s_pos2[threadIdx.x];
                if (kk) for (int w = 0; w < key_words; w++) out_keys[dstk * key_words + w] = keys[r * key_words + w];
            }
        }
    }
    // rows: the CTA's surviving rows are listed in shared memory and moved as one dense stream of 16-byte pieces
    // (a surviving row's destination follows the previous one's, so the stores are contiguous)
    __shared__ uint16_t s_keep[FL_ROWS];
    __shared__ uint32_t s_nk;
    if (threadIdx.x < FL_ROWS && (f & 2)) s_keep[s_pos2[threadIdx.x]] = (uint16_t)threadIdx.x;
    if (threadIdx.x == FL_ROWS - 1) s_nk = s_pos2[FL_ROWS - 1] + ((f & 2) ? 1u : 0u);
    __syncthreads();
    const uint32_t nk = s_nk;
    if (!mask_ambig && (n_samples & 15u) == 0) {
        const uint32_t n16 = n_samples >> 4, total = nk * n16;
        const uint4* s4 = reinterpret_cast<const uint4*>(var + r0 * n_samples);
        uint4* d4 = reinterpret_cast<uint4*>(out_var + base2 * n_samples);
        for (uint32_t idx = threadIdx.x; idx < total; idx += FL_THREADS) {
            const uint32_t j = idx / n16, cidx = idx - j * n16;
            d4[idx] = __ldg(&s4[(uint32_t)s_keep[j] * n16 + cidx]);
        }
        return;
    }
    const uint32_t total = nk * n_samples;
    const uint8_t* src0 = var + r0 * n_samples;
    uint8_t* dst0 = out_var + base2 * n_samples;
    for (uint32_t idx = threadIdx.x; idx < total; idx += FL_THREADS) {
        const uint32_t j = idx / n_samples, cidx = idx - j * n_samples;
        uint32_t bb = src0[(uint32_t)s_keep[j] * n_samples + cidx];
        if (mask_ambig && is_ambig_dev(bb)) bb = 'N';            // merge_ska_array.rs:388-399
        dst0[idx] = (uint8_t)bb;
    }
}

int32_t filter_array(Array& a, uint64_t min_count, bool filter_ambig_as_missing, int filter_type, bool mask_ambig,
                     bool ignore_const_gaps, bool update_kmers) {
    Context& c = ctx();
    if (filter_type < 0 || filter_type > 3) throw Error("invalid filter type");
    const uint64_t R = a.n_rows;
    const uint32_t N = (uint32_t)a.n_samples;
    const int kw = a.key_bits / 64;
    if (R == 0) return 0;
    const uint32_t n_cta = (uint32_t)((R + FL_ROWS - 1) / FL_ROWS);
    timer().begin("filter");
    DevBuf<uint8_t> flags(R);
    DevBuf<uint32_t> new_cnt(R), keep1(n_cta), keep2(n_cta), scratch(1);
    DevBuf<uint64_t> off1(n_cta + 1), off2(n_cta + 1);
    if ((N & 15u) == 0 && N >= 512)
        SKA_LAUNCH(filter_flags_wide_kernel, n_cta, FL_THREADS, 0, a.variants.p, a.counts.p, R, N, min_count,
                   filter_ambig_as_missing ? 1 : 0, filter_type, ignore_const_gaps ? 1 : 0, flags.p, new_cnt.p, keep1.p, keep2.p);
    else if ((N & 15u) == 0)
        SKA_LAUNCH(filter_flags_rows_kernel, n_cta, FL_ROWS, 0, a.variants.p, a.counts.p, R, N, min_count,
                   filter_ambig_as_missing ? 1 : 0, filter_type, ignore_const_gaps ? 1 : 0, flags.p, new_cnt.p, keep1.p, keep2.p);
    else
        SKA_LAUNCH(filter_flags_kernel, n_cta, FL_THREADS, 0, a.variants.p, a.counts.p, R, N, min_count,
                   filter_ambig_as_missing ? 1 : 0, filter_type, ignore_const_gaps ? 1 : 0, flags.p, new_cnt.p, keep1.p, keep2.p);
    launch_scan_counts(keep1.p, n_cta, off1.p, scratch.p);
    launch_scan_counts(keep2.p, n_cta, off2.p, scratch.p);
    uint64_t n1 = 0, n2 = 0;
    read_back(&n1, off1.p + n_cta, 8);
    read_back(&n2, off2.p + n_cta, 8);
    read_back_sync();

    // which key compaction applies: the filter loop's (update_kmers) wins over update_counts'
    const bool keys_match_rows = a.n_keys == a.n_rows;   // false only after an earlier update_kmers=false filter
    int key_mode = update_kmers ? 2 : (filter_ambig_as_missing ? 1 : 0);
    if (key_mode && !keys_match_rows)
        throw Error("cannot compact split k-mers: keys no longer correspond to rows (earlier filter ran with update_kmers=false)");
    DevBuf<uint8_t> out_var(n2 * N);
    DevBuf<uint32_t> out_cnt(n2);
    DevBuf<uint64_t> out_keys;
    uint64_t n_keys_out = a.n_keys;
    if (key_mode) { n_keys_out = key_mode == 1 ? n1 : n2; out_keys.alloc(n_keys_out * kw); }
    SKA_LAUNCH(filter_scatter_kernel, n_cta, FL_THREADS, 0, a.variants.p, new_cnt.p, a.keys.p, flags.p, off1.p, off2.p, R,
               a.n_keys, N, kw, key_mode, mask_ambig ? 1 : 0, out_var.p, out_cnt.p, key_mode ? out_keys.p : (uint64_t*)nullptr);
    timer().end();
    a.variants = std::move(out_var);
    a.counts = std::move(out_cnt);
    if (key_mode) { a.keys = std::move(out_keys); a.n_keys = n_keys_out; }
    a.n_rows = n2;
    return (int32_t)(n1 - n2);
}

// ---- transpose (rows x samples) -> (samples x rows)
constexpr int TP_TILE = 64;
__global__ void __launch_bounds__(256)
transpose_kernel(const uint8_t* __restrict__ in, uint64_t n_rows, uint32_t n_samples, uint8_t* __restrict__ out) {
    __shared__ uint8_t tile[TP_TILE][TP_TILE + 4];
    const uint64_t r0 = (uint64_t)blockIdx.x * TP_TILE;
    const uint32_t s0 = blockIdx.y * TP_TILE;
    for (int idx = threadIdx.x; idx < TP_TILE * TP_TILE; idx += 256) {
        int r = idx / TP_TILE, s = idx % TP_TILE;
        if (r0 + r < n_rows && s0 + s < n_samples) tile[r][s] = in[(r0 + r) * n_samples + s0 + s];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < TP_TILE * TP_TILE; idx += 256) {
        int s = idx / TP_TILE, r = idx % TP_TILE;
        if (r0 + r < n_rows && s0 + s < n_samples) out[(uint64_t)(s0 + s) * n_rows + r0 + r] = tile[r][s];
    }
}

// Word form for n_samples % 4 == 0: 128 rows x 64 samples per CTA.  Loads are 4-byte words along a row (a warp
// reads two 64-byte row pieces), 4 x 4 byte blocks are transposed in registers with byte permutes, and every
// sample's 128 output bytes leave as aligned 4-byte words (the first and last word of a run byte by byte, because
// sample s starts at byte s * n_rows of the output, which is aligned only by chance).
// TR_COLS = 128: a CTA takes whole 128-byte lines of a row (with 64 columns two CTAs read each line: 1.2x the algorithmic
// traffic in the ncu captures) -- measured slower, see transpose_array.
constexpr int TR_ROWS = 128;
template <int TR_COLS>
__global__ void __launch_bounds__(256)
transpose_words_kernel(const uint8_t* __restrict__ in, uint64_t n_rows, uint32_t n_samples, uint8_t* __restrict__ out) {
    __shared__ uint32_t s_in[TR_ROWS][TR_COLS / 4 + 1];
    __shared__ uint32_t s_t[TR_COLS][TR_ROWS / 4 + 1];             // s_t[s][q] = rows 4q..4q+3 of sample s
    const uint64_t r0 = (uint64_t)blockIdx.x * TR_ROWS;
    const uint32_t s0 = blockIdx.y * TR_COLS;
    const uint32_t nr = (uint32_t)min((uint64_t)TR_ROWS, n_rows - r0), ns = min((uint32_t)TR_COLS, n_samples - s0);
    const uint32_t nw = ns >> 2;
    const uint32_t* in32 = reinterpret_cast<const uint32_t*>(in);
    for (uint32_t idx = threadIdx.x; idx < TR_ROWS * (TR_COLS / 4); idx += 256) {
        const uint32_t r = idx / (TR_COLS / 4), cw = idx % (TR_COLS / 4);
        uint32_t v = 0;
        if (r < nr && cw < nw) v = __ldg(&in32[(((r0 + r) * n_samples + s0) >> 2) + cw]);
        s_in[r][cw] = v;
    }
    __syncthreads();
    for (uint32_t blk = threadIdx.x; blk < (TR_ROWS / 4) * (TR_COLS / 4); blk += 256) {
        const uint32_t r4 = blk % (TR_ROWS / 4), s4 = blk / (TR_ROWS / 4);
        const uint32_t a0 = s_in[4 * r4][s4], a1 = s_in[4 * r4 + 1][s4], a2 = s_in[4 * r4 + 2][s4], a3 = s_in[4 * r4 + 3][s4];
        const uint32_t t0 = __byte_perm(a0, a1, 0x5140), t1 = __byte_perm(a2, a3, 0x5140);    // (x.0 y.0 x.1 y.1)
        const uint32_t t2 = __byte_perm(a0, a1, 0x7362), t3 = __byte_perm(a2, a3, 0x7362);    // (x.2 y.2 x.3 y.3)
        s_t[4 * s4][r4] = __byte_perm(t0, t1, 0x5410);
        s_t[4 * s4 + 1][r4] = __byte_perm(t0, t1, 0x7632);
        s_t[4 * s4 + 2][r4] = __byte_perm(t2, t3, 0x5410);
        s_t[4 * s4 + 3][r4] = __byte_perm(t2, t3, 0x7632);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t sidx = warp; sidx < ns; sidx += 8) {
        const uint8_t* trow = reinterpret_cast<const uint8_t*>(s_t[sidx]);
        const uint64_t d0 = (uint64_t)(s0 + sidx) * n_rows + r0;   // byte offset of this run in the output
        const uint32_t m = (uint32_t)(d0 & 3u);
        uint8_t* a_out = out + (d0 - m);                          // 4-byte aligned
        for (uint32_t wi = lane; 4 * wi < m + nr; wi += 32) {
            const int p0 = (int)(4 * wi) - (int)m;                // tile row of the word's first byte
            if (p0 >= 0 && p0 + 4 <= (int)nr) {
                const uint32_t word = (uint32_t)trow[p0] | ((uint32_t)trow[p0 + 1] << 8) | ((uint32_t)trow[p0 + 2] << 16) |
                                      ((uint32_t)trow[p0 + 3] << 24);
                *reinterpret_cast<uint32_t*>(a_out + 4 * wi) = word;
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) { const int pp = p0 + j; if (pp >= 0 && pp < (int)nr) a_out[4 * wi + j] = trow[pp]; }
            }
        }
    }
}

void transpose_array(const Array& a, uint8_t* d_out) {
    if (a.n_rows == 0 || a.n_samples == 0) return;
    if ((a.n_samples & 3u) == 0 && (reinterpret_cast<uintptr_t>(d_out) & 3u) == 0) {
        // measured on 3.06 M rows x 128 samples: 0.39 ms with 64 columns, 0.44 ms with 128 (34 KB of shared memory per CTA cost
        // more than the second read of each line, which mostly hits L2): 64 stays the default, SKA_TRANSPOSE_COLS=128 selects the other
        const char* e = getenv("SKA_TRANSPOSE_COLS");
        const bool wide_tile = a.n_samples >= 128 && e && atoi(e) == 128;
        const unsigned cols = wide_tile ? 128u : 64u;
        dim3 grid((unsigned)((a.n_rows + TR_ROWS - 1) / TR_ROWS), (unsigned)((a.n_samples + cols - 1) / cols));
        timer().begin("transpose");
        if (wide_tile) SKA_LAUNCH(transpose_words_kernel<128>, grid, 256, 0, a.variants.p, a.n_rows, (uint32_t)a.n_samples, d_out);
        else SKA_LAUNCH(transpose_words_kernel<64>, grid, 256, 0, a.variants.p, a.n_rows, (uint32_t)a.n_samples, d_out);
        timer().end();
        return;
    }
    dim3 grid((unsigned)((a.n_rows + TP_TILE - 1) / TP_TILE), (unsigned)((a.n_samples + TP_TILE - 1) / TP_TILE));
    timer().begin("transpose");
    SKA_LAUNCH(transpose_kernel, grid, 256, 0, a.variants.p, a.n_rows, (uint32_t)a.n_samples, d_out);
    timer().end();
}

}  // namespace ska
