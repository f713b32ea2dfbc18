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

constexpr int DT = 16;            // samples per tile side of the exact byte-wise kernel

// ---- planes: out[plane][word][sample], samples padded to a multiple of 64 (SP) so that tiles of the pair kernel never
// bound-check.  One CTA = `wpc` words (32 rows each): their rows are one contiguous 16-byte aligned block of the
// row-major array, staged with 16-byte loads; then one thread per (word, sample) walks the 32 rows (consecutive threads
// read consecutive bytes: no bank conflicts).  A byte is classified by ONE lookup in a 256-entry shared-memory table
// (bits: present, unambiguous, code bit 0, code bit 1, "odd alphabet"); the four flags of row r go to nibble r of a 128-bit
// accumulator and the nibbles are pulled apart into the four plane words once per 32 rows (~10 instructions per byte
// against ~18 for classifying with compares).
__device__ __forceinline__ uint32_t every_4th_bit(uint32_t acc, int i) {     // bits i, i+4, ... of acc -> low 8 bits
    uint32_t x = (acc >> i) & 0x11111111u;
    x = (x | (x >> 3)) & 0x03030303u;
    x = (x | (x >> 6)) & 0x000F000Fu;
    return (x | (x >> 12)) & 0xFFu;
}
__global__ void __launch_bounds__(256)
planes_kernel(const uint8_t* __restrict__ var, uint64_t n_rows, uint32_t n_samples, uint32_t SP, uint64_t n_words, uint32_t wpc,
              uint32_t* __restrict__ planes, uint32_t* __restrict__ odd_alphabet) {
    extern __shared__ __align__(16) uint8_t s_tile[];                       // wpc x 32 rows x n_samples
    __shared__ uint8_t s_lut[256];
    {
        const uint32_t b = threadIdx.x;
        const uint32_t present = b != '-';
        const uint32_t unamb = (b == 'A' || b == 'C' || b == 'G' || b == 'T');
        const uint32_t lb = b | 0x20u;
        const uint32_t ref_unamb = present & (lb == 'a' || lb == 'c' || lb == 'g' || lb == 't' || lb == 'u');   // !is_ambiguous (bit_encoding.rs:58-61)
        const uint32_t code = (b >> 1) & 3u;
        s_lut[b] = (uint8_t)(present | (unamb << 1) | ((unamb & code & 1u) << 2) | ((unamb & (code >> 1)) << 3) | ((ref_unamb != unamb) << 4));
    }
    const uint64_t w0 = (uint64_t)blockIdx.x * wpc;
    const uint64_t r0 = w0 * 32;
    const uint32_t tile_rows = wpc * 32;
    const uint32_t rows = (uint32_t)min((uint64_t)tile_rows, n_rows - r0);
    const uint32_t n16 = (uint32_t)(((uint64_t)rows * n_samples) / 16);
    const uint4* src = reinterpret_cast<const uint4*>(var + r0 * n_samples);
    uint4* dst = reinterpret_cast<uint4*>(s_tile);
    for (uint32_t i = threadIdx.x; i < n16; i += 256) dst[i] = __ldg(src + i);
    for (uint32_t i = n16 * 16 + threadIdx.x; i < tile_rows * n_samples; i += 256)
        s_tile[i] = i < rows * n_samples ? var[r0 * n_samples + i] : (uint8_t)'-';
    __syncthreads();
    const uint64_t plane_stride = n_words * SP;
    uint32_t odd = 0;
    for (uint32_t item = threadIdx.x; item < wpc * SP; item += 256) {
        const uint32_t wq = item / SP, s = item - wq * SP;
        if (w0 + wq >= n_words) break;
        uint32_t acc[4] = {0u, 0u, 0u, 0u};
        if (s < n_samples) {
            const uint8_t* col = s_tile + (size_t)wq * 32 * n_samples + s;
#pragma unroll
            for (int r = 0; r < 32; r++) {
                const uint32_t v = s_lut[col[r * n_samples]];
                odd |= v;
                acc[r >> 3] |= (v & 0xFu) << (4 * (r & 7));
            }
        }
        uint32_t m[4];
#pragma unroll
        for (int i = 0; i < 4; i++)
            m[i] = every_4th_bit(acc[0], i) | (every_4th_bit(acc[1], i) << 8) | (every_4th_bit(acc[2], i) << 16) | (every_4th_bit(acc[3], i) << 24);
        const uint64_t o = (w0 + wq) * SP + s;
        planes[o] = m[0]; planes[plane_stride + o] = m[1]; planes[2 * plane_stride + o] = m[2]; planes[3 * plane_stride + o] = m[3];
    }
    if (odd & 16u) atomicExch(odd_alphabet, 1u);
}

// ---- pair kernel: register tile of 4 x 4 sample pairs per thread over a TS x TS tile of samples (TS = 32: 64 threads,
// TS = 64: 256 threads), DW words staged per step; per pair and word 3 POPC + 5 LOP3, operands by 8 LDS.128 per 16 pairs.
// POPC issues at 15.4 lanes/clk/SM on B200 (profiles/microbench/popc_mb.cu: 4.5e12 popc/s), a quarter of the LOP3 rate:
// the kernel is bound by the popcounts.  grid = (tile pairs, word splits); partial sums by 64-bit atomics.
constexpr int PW = 8;             // words staged per step
template <int TS>
__global__ void __launch_bounds__((TS / 4) * (TS / 4))
pair_popc_kernel(const uint32_t* __restrict__ planes, uint32_t n_samples, uint32_t SP, uint64_t n_words, uint32_t n_tiles,
                 uint64_t words_per_split, unsigned long long* __restrict__ acc) {
    constexpr int Q = TS / 4;                                   // threads per tile side
    constexpr int NT = Q * Q;
    __shared__ __align__(16) uint32_t s_i[4][PW][TS];
    __shared__ __align__(16) uint32_t s_j[4][PW][TS];
    uint32_t t = blockIdx.x, ti = 0;
    while (t >= n_tiles - ti) { t -= n_tiles - ti; ti++; }
    const uint32_t tj = ti + t;
    const uint32_t a = threadIdx.x / Q, b = threadIdx.x % Q;    // pairs (ti*TS + 4a + x, tj*TS + 4b + y)
    const bool skip = (ti == tj) && (a > b);                    // every pair of the thread has i > j
    const uint64_t w_beg = (uint64_t)blockIdx.y * words_per_split;
    const uint64_t w_end = min(n_words, w_beg + words_per_split);
    const uint64_t plane_stride = n_words * SP;
    uint32_t mism[4][4], match[4][4], dist[4][4];
#pragma unroll
    for (int x = 0; x < 4; x++)
#pragma unroll
        for (int y = 0; y < 4; y++) { mism[x][y] = 0; match[x][y] = 0; dist[x][y] = 0; }
    for (uint64_t w0 = w_beg; w0 < w_end; w0 += PW) {
        for (uint32_t idx = threadIdx.x; idx < 4 * PW * Q; idx += NT) {
            const uint32_t p = idx / (PW * Q), rem = idx - p * (PW * Q), w = rem / Q, q = rem - w * Q;
            const uint64_t gw = w0 + w;
            uint4 vi = make_uint4(0u, 0u, 0u, 0u), vj = vi;
            if (gw < w_end) {
                const uint32_t* base = planes + p * plane_stride + gw * SP;
                vi = __ldg(reinterpret_cast<const uint4*>(base + ti * TS + 4 * q));
                vj = __ldg(reinterpret_cast<const uint4*>(base + tj * TS + 4 * q));
            }
            *reinterpret_cast<uint4*>(&s_i[p][w][4 * q]) = vi;
            *reinterpret_cast<uint4*>(&s_j[p][w][4 * q]) = vj;
        }
        __syncthreads();
        if (!skip) {
#pragma unroll 2
            for (int w = 0; w < PW; w++) {
                const uint4 Pi = *reinterpret_cast<const uint4*>(&s_i[0][w][4 * a]), Ui = *reinterpret_cast<const uint4*>(&s_i[1][w][4 * a]);
                const uint4 Xi = *reinterpret_cast<const uint4*>(&s_i[2][w][4 * a]), Yi = *reinterpret_cast<const uint4*>(&s_i[3][w][4 * a]);
                const uint4 Pj = *reinterpret_cast<const uint4*>(&s_j[0][w][4 * b]), Uj = *reinterpret_cast<const uint4*>(&s_j[1][w][4 * b]);
                const uint4 Xj = *reinterpret_cast<const uint4*>(&s_j[2][w][4 * b]), Yj = *reinterpret_cast<const uint4*>(&s_j[3][w][4 * b]);
                const uint32_t pi[4] = {Pi.x, Pi.y, Pi.z, Pi.w}, ui[4] = {Ui.x, Ui.y, Ui.z, Ui.w}, xi[4] = {Xi.x, Xi.y, Xi.z, Xi.w}, yi[4] = {Yi.x, Yi.y, Yi.z, Yi.w};
                const uint32_t pj[4] = {Pj.x, Pj.y, Pj.z, Pj.w}, uj[4] = {Uj.x, Uj.y, Uj.z, Uj.w}, xj[4] = {Xj.x, Xj.y, Xj.z, Xj.w}, yj[4] = {Yj.x, Yj.y, Yj.z, Yj.w};
#pragma unroll
                for (int x = 0; x < 4; x++)
#pragma unroll
                    for (int y = 0; y < 4; y++) {
                        const uint32_t both = ui[x] & uj[y];
                        const uint32_t diff = (xi[x] ^ xj[y]) | (yi[x] ^ yj[y]);
                        mism[x][y] += __popc(pi[x] ^ pj[y]);
                        match[x][y] += __popc(both);
                        dist[x][y] += __popc(both & diff);
                    }
            }
        }
        __syncthreads();
    }
    if (skip) return;
#pragma unroll
    for (int x = 0; x < 4; x++)
#pragma unroll
        for (int y = 0; y < 4; y++) {
            const uint32_t si = ti * TS + 4 * a + x, sj = tj * TS + 4 * b + y;
            if (si < sj && sj < n_samples && (mism[x][y] | match[x][y] | dist[x][y])) {
                const uint64_t pidx = (uint64_t)si * n_samples - (uint64_t)si * (si + 1) / 2 + (sj - si - 1);
                if (mism[x][y]) atomicAdd(&acc[3 * pidx + 0], (unsigned long long)mism[x][y]);
                if (match[x][y]) atomicAdd(&acc[3 * pidx + 1], (unsigned long long)match[x][y]);
                if (dist[x][y]) atomicAdd(&acc[3 * pidx + 2], (unsigned long long)dist[x][y]);
            }
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
    if (timer().stages.size() > 48) timer().clear();             // repeated distance calls without a build in between
    const uint64_t n_pairs = (uint64_t)N * (N - 1) / 2;
    const uint64_t R = a.n_rows;
    const uint32_t n_tiles = (N + DT - 1) / DT;
    const uint32_t n_tile_pairs = n_tiles * (n_tiles + 1) / 2;
    bool use_planes = filt_ambig && R > 0;
    std::vector<unsigned long long> h_acc;
    memset(h_counters, 0, n_pairs * 4 * sizeof(uint64_t));

    if (use_planes) {
        const uint64_t n_words = (R + 31) / 32;
        const uint32_t SP = (N + 63) / 64 * 64;
        DevBuf<uint32_t> planes(4 * (uint64_t)SP * n_words), odd(1);
        dev_fill(odd.p, 0, 4);
        const uint32_t wpc = std::max<uint32_t>(1, 32768u / (32u * N));      // words per CTA: a 32 KB tile
        size_t smem = (size_t)wpc * 32 * N;
        if (smem + 1024 > c.smem_optin) use_planes = false;
        else {
            timer().begin("planes");
            SKA_CUDA(cudaFuncSetAttribute(planes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            SKA_LAUNCH(planes_kernel, (unsigned)((n_words + wpc - 1) / wpc), 256, smem, a.variants.p, R, N, SP, n_words, wpc, planes.p, odd.p);
            timer().end();
            uint32_t h_odd = 0;
            SKA_CUDA(cudaMemcpyAsync(&h_odd, odd.p, 4, cudaMemcpyDeviceToHost, c.stream));
            SKA_CUDA(cudaStreamSynchronize(c.stream));
            if (h_odd) use_planes = false;
        }
        if (use_planes) {
            DevBuf<unsigned long long> acc(3 * n_pairs);
            dev_fill(acc.p, 0, acc.bytes());
            const bool small = N <= 256;                                   // 32 x 32 tiles waste less of the diagonal
            const uint32_t TS = small ? 32 : 64;
            const uint32_t nt = (N + TS - 1) / TS, ntp = nt * (nt + 1) / 2;
            const uint64_t chunks = (n_words + PW - 1) / PW;
            const uint64_t want_ctas = (uint64_t)c.sm_count * (small ? 24 : 6);
            uint64_t splits = std::min<uint64_t>(std::min<uint64_t>(chunks, std::max<uint64_t>(1, (want_ctas + ntp - 1) / ntp)), 65535);
            const uint64_t words_per_split = ((chunks + splits - 1) / splits) * PW;
            splits = (n_words + words_per_split - 1) / words_per_split;
            dim3 grid(ntp, (unsigned)splits);
            timer().begin("pair_popc");
            if (small) SKA_LAUNCH((pair_popc_kernel<32>), grid, 64, 0, planes.p, N, SP, n_words, nt, words_per_split, acc.p);
            else SKA_LAUNCH((pair_popc_kernel<64>), grid, 256, 0, planes.p, N, SP, n_words, nt, words_per_split, acc.p);
            timer().end();
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
    dev_fill(acc.p, 0, acc.bytes());
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

// sharded: the array is one rank's slab of rows; the per-pair counters and the slabs' constant-site counts are summed over the
// ranks of the library's communicator (comm.cu) before the table is formed, so every rank returns the whole job's table
void distance_array(const Array& a, double constant, bool filt_ambig, DistOut* h_out, bool sharded) {
    const uint32_t N = (uint32_t)a.n_samples;
    if (N < 2) return;
    const uint64_t n_pairs = (uint64_t)N * (N - 1) / 2;
    std::vector<uint64_t> cnt(4 * n_pairs + 1);
    distance_counters(a, filt_ambig, cnt.data());
    if (sharded) {
        cnt[4 * n_pairs] = (uint64_t)constant;
        allreduce_u64(cnt.data(), cnt.size());
        constant = (double)cnt[4 * n_pairs];
    }
    for (uint64_t p = 0; p < n_pairs; p++) {
        double d = filt_ambig ? (double)cnt[4 * p + 2]
                              : (double)cnt[4 * p + 3] - (double)cnt[4 * p + 2] / 36.0;   // sum of (1 - overlap), exact rational
        finish(constant, d, cnt[4 * p + 1], cnt[4 * p + 0], h_out[p]);
    }
}

}  // namespace ska
