// distance.cu -- K5: pairwise SNP / mismatch distances over the middle-base array.
//
// Replaces MergeSkaArray::distance + variant_dist (src/merge_ska_array.rs:416-438, 587-632), which
// walks two stride-N byte columns per pair.  Here (default mode, ambiguous bases ignored):
//   1. planes kernel: each sample's column becomes four bit-planes over rows, 32 rows per word --
//      P (not '-'), U (unambiguous: A/C/G/T), B0,B1 (2-bit base code where U) -- built with warp
//      ballots from a coalesced shared-memory tile of the row-major array;
//   2. pair kernel: 16x16 sample tiles; per pair and word
//         mismatches += popc(P_i ^ P_j)                       ('-' vs anything but '-')
//         matches    += popc(U_i & U_j)                       (both unambiguous)
//         distance   += popc(U_i & U_j & ((B0_i^B0_j)|(B1_i^B1_j)))
//      tiles are staged in shared memory (bank-skewed), row-words are split over CTAs and the
//      partial sums reduced with 64-bit atomics.
// This kernel is bound by the integer popc/logic issue rate, not by HBM.
// `--allow-ambiguous` (filt_ambig = false) and arrays whose unambiguous symbols are not plain
// upper-case ACGT use an exact byte-wise kernel: per pair and row, overlap of the two IUPAC sets
// is accumulated as the integer 36*|a&b|/(|a||b|) (bit_encoding.rs:65-85: N and unknown -> empty).
#include <cstring>

#include "common.cuh"

namespace ska {

constexpr int DT = 16;            // samples per tile side
constexpr int DW = 32;            // row-words per staged chunk

// ---- planes: out[plane][sample][word]
__global__ void __launch_bounds__(256)
planes_kernel(const uint8_t* __restrict__ var, uint64_t n_rows, uint32_t n_samples, uint64_t n_words,
              uint32_t* __restrict__ planes, uint32_t* __restrict__ odd_alphabet) {
    extern __shared__ uint8_t s_tile[];                       // 32 rows x (n_samples + 4)
    const uint32_t pitch = n_samples + 4;
    const uint64_t w = blockIdx.x;
    const uint64_t r0 = w * 32;
    for (uint32_t idx = threadIdx.x; idx < 32 * n_samples; idx += 256) {
        uint32_t r = idx / n_samples, s = idx - r * n_samples;
        s_tile[r * pitch + s] = (r0 + r < n_rows) ? var[(r0 + r) * n_samples + s] : (uint8_t)'-';
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t plane_stride = (uint64_t)n_samples * n_words;
    for (uint32_t s = warp; s < n_samples; s += 8) {
        uint32_t b = s_tile[lane * pitch + s];
        bool present = b != '-';
        bool unamb = (b == 'A' || b == 'C' || b == 'G' || b == 'T');
        uint32_t lb = b | 0x20u;
        bool ref_unamb = present && (lb == 'a' || lb == 'c' || lb == 'g' || lb == 't' || lb == 'u');   // !is_ambiguous
        if (ref_unamb != unamb) atomicExch(odd_alphabet, 1u);      // lower case / U: take the exact path
        uint32_t code = (b >> 1) & 3u;
        uint32_t mP = __ballot_sync(0xFFFFFFFFu, present);
        uint32_t mU = __ballot_sync(0xFFFFFFFFu, unamb);
        uint32_t m0 = __ballot_sync(0xFFFFFFFFu, unamb && (code & 1u));
        uint32_t m1 = __ballot_sync(0xFFFFFFFFu, unamb && (code & 2u));
        if (lane == 0) {
            uint64_t o = (uint64_t)s * n_words + w;
            planes[o] = mP; planes[plane_stride + o] = mU; planes[2 * plane_stride + o] = m0; planes[3 * plane_stride + o] = m1;
        }
    }
}

// ---- pair kernel: acc[pair_index(i,j)][3] += partial sums;  grid = (tile pairs, word splits)
__global__ void __launch_bounds__(256)
pair_popc_kernel(const uint32_t* __restrict__ planes, uint32_t n_samples, uint64_t n_words, uint32_t n_tiles,
                 uint64_t words_per_split, unsigned long long* __restrict__ acc) {
    __shared__ uint32_t s_i[4][DT][DW];
    __shared__ uint32_t s_j[4][DT][DW];
    // decode tile pair (ti <= tj) from blockIdx.x
    uint32_t t = blockIdx.x, ti = 0;
    while (t >= n_tiles - ti) { t -= n_tiles - ti; ti++; }
    const uint32_t tj = ti + t;
    const uint32_t a = threadIdx.x / DT, b = threadIdx.x % DT;    // pair (ti*DT + a, tj*DT + b)
    const uint32_t si = ti * DT + a, sj = tj * DT + b;
    const bool active = si < n_samples && sj < n_samples && si < sj;
    const uint64_t w_beg = (uint64_t)blockIdx.y * words_per_split;
    const uint64_t w_end = min(n_words, w_beg + words_per_split);
    const uint64_t plane_stride = (uint64_t)n_samples * n_words;
    const uint32_t skew = (b + DT * (a & 1u)) & (DW - 1);
    uint32_t mism = 0, match = 0, dist = 0;
    for (uint64_t w0 = w_beg; w0 < w_end; w0 += DW) {
        // stage 2 x (4 planes x 16 samples x 32 words)
        for (uint32_t idx = threadIdx.x; idx < 4 * DT * DW; idx += 256) {
            uint32_t p = idx / (DT * DW), rem = idx - p * DT * DW, s = rem / DW, w = rem - s * DW;
            uint64_t gw = w0 + w;
            uint32_t gi = ti * DT + s, gj = tj * DT + s;
            s_i[p][s][w] = (gw < w_end && gi < n_samples) ? planes[p * plane_stride + (uint64_t)gi * n_words + gw] : 0u;
            s_j[p][s][w] = (gw < w_end && gj < n_samples) ? planes[p * plane_stride + (uint64_t)gj * n_words + gw] : 0u;
        }
        __syncthreads();
        if (active) {
#pragma unroll 8
            for (uint32_t w = 0; w < DW; w++) {
                uint32_t ww = (w + skew) & (DW - 1);
                uint32_t pi = s_i[0][a][ww], pj = s_j[0][b][ww];
                uint32_t ui = s_i[1][a][ww], uj = s_j[1][b][ww];
                uint32_t both = ui & uj;
                uint32_t diff = (s_i[2][a][ww] ^ s_j[2][b][ww]) | (s_i[3][a][ww] ^ s_j[3][b][ww]);
                mism += __popc(pi ^ pj);
                match += __popc(both);
                dist += __popc(both & diff);
            }
        }
        __syncthreads();
    }
    if (active && (mism | match | dist)) {
        uint64_t pidx = (uint64_t)si * n_samples - (uint64_t)si * (si + 1) / 2 + (sj - si - 1);
        atomicAdd(&acc[3 * pidx + 0], (unsigned long long)mism);
        atomicAdd(&acc[3 * pidx + 1], (unsigned long long)match);
        atomicAdd(&acc[3 * pidx + 2], (unsigned long long)dist);
    }
}

// ---- exact byte-wise kernel (both modes).  acc[pair][4] = {mismatches, matches, dist_int, n_both}
// dist_int: filt_ambig -> #differing unambiguous pairs; else sum of 36*overlap over both-present rows
__device__ __forceinline__ uint32_t iupac_set(uint32_t b) {     // base_to_prob support set, upper case only
    switch (b) {                                                 // bit0=A bit1=C bit2=T bit3=G
        case 'A': return 1; case 'C': return 2; case 'T': case 'U': return 4; case 'G': return 8;
        case 'R': return 9; case 'Y': return 6; case 'S': return 10; case 'W': return 5; case 'K': return 12;
        case 'M': return 3; case 'B': return 14; case 'D': return 13; case 'H': return 7; case 'V': return 11;
        default: return 0;
    }
}
__device__ __forceinline__ bool is_ambig_d(uint32_t b) {
    b |= 0x20u;
    return !(b == 'a' || b == 'c' || b == 'g' || b == 't' || b == 'u' || b == '-');
}
constexpr int BR = 64;   // rows per staged chunk
__global__ void __launch_bounds__(256)
pair_bytes_kernel(const uint8_t* __restrict__ var, uint64_t n_rows, uint32_t n_samples, uint32_t n_tiles, uint64_t rows_per_split,
                  int filt_ambig, unsigned long long* __restrict__ acc) {
    __shared__ uint8_t s_i[BR][DT + 1];
    __shared__ uint8_t s_j[BR][DT + 1];
    uint32_t t = blockIdx.x, ti = 0;
    while (t >= n_tiles - ti) { t -= n_tiles - ti; ti++; }
    const uint32_t tj = ti + t;
    const uint32_t a = threadIdx.x / DT, b = threadIdx.x % DT;
    const uint32_t si = ti * DT + a, sj = tj * DT + b;
    const bool active = si < n_samples && sj < n_samples && si < sj;
    const uint64_t r_beg = (uint64_t)blockIdx.y * rows_per_split;
    const uint64_t r_end = min(n_rows, r_beg + rows_per_split);
    unsigned long long mism = 0, match = 0, dist = 0, both = 0;
    for (uint64_t r0 = r_beg; r0 < r_end; r0 += BR) {
        for (uint32_t idx = threadIdx.x; idx < BR * DT; idx += 256) {
            uint32_t r = idx / DT, s = idx - r * DT;
            uint64_t gr = r0 + r;
            uint32_t gi = ti * DT + s, gj = tj * DT + s;
            s_i[r][s] = (gr < r_end && gi < n_samples) ? var[gr * n_samples + gi] : (uint8_t)'-';
            s_j[r][s] = (gr < r_end && gj < n_samples) ? var[gr * n_samples + gj] : (uint8_t)'-';
        }
        __syncthreads();
        if (active) {
            for (uint32_t r = 0; r < BR; r++) {
                uint32_t v1 = s_i[r][a], v2 = s_j[r][b];
                if (v1 == '-' || v2 == '-') { mism += (v1 != v2); continue; }     // merge_ska_array.rs:600-603
                if (filt_ambig) {
                    if (!is_ambig_d(v1) && !is_ambig_d(v2)) { match++; dist += (v1 != v2); }   // :604-610
                } else {
                    uint32_t m1 = iupac_set(v1), m2 = iupac_set(v2);
                    both++;
                    uint32_t inter = __popc(m1 & m2);
                    if (inter) { match++; dist += 36u * inter / (__popc(m1) * __popc(m2)); }   // :611-619
                }
            }
        }
        __syncthreads();
    }
    if (active) {
        uint64_t pidx = (uint64_t)si * n_samples - (uint64_t)si * (si + 1) / 2 + (sj - si - 1);
        atomicAdd(&acc[4 * pidx + 0], mism);
        atomicAdd(&acc[4 * pidx + 1], match);
        atomicAdd(&acc[4 * pidx + 2], dist);
        atomicAdd(&acc[4 * pidx + 3], both);
    }
}

static void finish(double constant, double distance, uint64_t match_rows, uint64_t mism, DistOut& o) {
    double matches = constant + (double)match_rows;        // merge_ska_array.rs:598
    double mismatches = (double)mism;
    o.distance = distance;
    o.mismatch_prop = (matches + mismatches) == 0.0 ? 0.0 : mismatches / (matches + mismatches);
    o.match_count = (uint64_t)matches;
    o.mismatch_count = (uint64_t)mismatches;
}

void distance_counters(const Array& a, bool filt_ambig, uint64_t* h_counters) {
    Context& c = ctx();
    const uint32_t N = (uint32_t)a.n_samples;
    if (N < 2) return;
    const uint64_t n_pairs = (uint64_t)N * (N - 1) / 2;
    const uint64_t R = a.n_rows;
    const uint32_t n_tiles = (N + DT - 1) / DT;
    const uint32_t n_tile_pairs = n_tiles * (n_tiles + 1) / 2;
    bool use_planes = filt_ambig && R > 0;
    std::vector<unsigned long long> h_acc;
    memset(h_counters, 0, n_pairs * 4 * sizeof(uint64_t));

    if (use_planes) {
        const uint64_t n_words = (R + 31) / 32;
        DevBuf<uint32_t> planes(4 * (uint64_t)N * n_words), odd(1);
        SKA_CUDA(cudaMemsetAsync(odd.p, 0, 4, c.stream));
        size_t smem = 32 * (size_t)(N + 4);
        if (smem > c.smem_optin) use_planes = false;
        else {
            SKA_CUDA(cudaFuncSetAttribute(planes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            SKA_LAUNCH(planes_kernel, (unsigned)n_words, 256, smem, a.variants.p, R, N, n_words, planes.p, odd.p);
            uint32_t h_odd = 0;
            SKA_CUDA(cudaMemcpyAsync(&h_odd, odd.p, 4, cudaMemcpyDeviceToHost, c.stream));
            SKA_CUDA(cudaStreamSynchronize(c.stream));
            if (h_odd) use_planes = false;
        }
        if (use_planes) {
            DevBuf<unsigned long long> acc(3 * n_pairs);
            SKA_CUDA(cudaMemsetAsync(acc.p, 0, acc.bytes(), c.stream));
            uint64_t chunks = (n_words + DW - 1) / DW;
            uint64_t want_splits = std::max<uint64_t>(1, (4ull * c.sm_count + n_tile_pairs - 1) / n_tile_pairs);
            uint64_t splits = std::min<uint64_t>(std::min<uint64_t>(chunks, want_splits), 65535);
            uint64_t words_per_split = ((chunks + splits - 1) / splits) * DW;
            splits = (n_words + words_per_split - 1) / words_per_split;
            dim3 grid(n_tile_pairs, (unsigned)splits);
            SKA_LAUNCH(pair_popc_kernel, grid, 256, 0, planes.p, N, n_words, n_tiles, words_per_split, acc.p);
            h_acc.resize(3 * n_pairs);
            SKA_CUDA(cudaMemcpyAsync(h_acc.data(), acc.p, acc.bytes(), cudaMemcpyDeviceToHost, c.stream));
            SKA_CUDA(cudaStreamSynchronize(c.stream));
            for (uint64_t p = 0; p < n_pairs; p++) {
                h_counters[4 * p + 0] = h_acc[3 * p + 0]; h_counters[4 * p + 1] = h_acc[3 * p + 1];
                h_counters[4 * p + 2] = h_acc[3 * p + 2]; h_counters[4 * p + 3] = 0;
            }
            return;
        }
    }
    // exact byte-wise path
    DevBuf<unsigned long long> acc(4 * n_pairs);
    SKA_CUDA(cudaMemsetAsync(acc.p, 0, acc.bytes(), c.stream));
    if (R > 0) {
        uint64_t chunks = (R + BR - 1) / BR;
        uint64_t want_splits = std::max<uint64_t>(1, (4ull * c.sm_count + n_tile_pairs - 1) / n_tile_pairs);
        uint64_t splits = std::min<uint64_t>(std::min<uint64_t>(chunks, want_splits), 65535);
        uint64_t rows_per_split = ((chunks + splits - 1) / splits) * BR;
        splits = (R + rows_per_split - 1) / rows_per_split;
        dim3 grid(n_tile_pairs, (unsigned)splits);
        SKA_LAUNCH(pair_bytes_kernel, grid, 256, 0, a.variants.p, R, N, n_tiles, rows_per_split, filt_ambig ? 1 : 0, acc.p);
    }
    static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "");
    SKA_CUDA(cudaMemcpyAsync(h_counters, acc.p, acc.bytes(), cudaMemcpyDeviceToHost, c.stream));
    SKA_CUDA(cudaStreamSynchronize(c.stream));
}

void distance_array(const Array& a, double constant, bool filt_ambig, DistOut* h_out) {
    const uint32_t N = (uint32_t)a.n_samples;
    if (N < 2) return;
    const uint64_t n_pairs = (uint64_t)N * (N - 1) / 2;
    std::vector<uint64_t> cnt(4 * n_pairs);
    distance_counters(a, filt_ambig, cnt.data());
    for (uint64_t p = 0; p < n_pairs; p++) {
        double d = filt_ambig ? (double)cnt[4 * p + 2]
                              : (double)cnt[4 * p + 3] - (double)cnt[4 * p + 2] / 36.0;   // sum of (1 - overlap), exact rational
        finish(constant, d, cnt[4 * p + 1], cnt[4 * p + 0], h_out[p]);
    }
}

}  // namespace ska
