// setops.cu -- operations on built arrays that join on the split-k-mer key:
//   extend         MergeSkaDict::extend            src/merge_ska_dict.rs:160-193   (`ska merge`)
//   delete_columns MergeSkaArray::delete_samples   src/merge_ska_array.rs:231-271  (`ska delete`)
//   weed_rows      MergeSkaArray::weed             src/merge_ska_array.rs:452-487  (`ska weed`)
//   map_reference  RefSka::map                     src/ska_ref.rs:508-533          (`ska map`: the lookup; writers are host code)
// The reference goes through HashMap<IntT, Vec<u8>> for all of them.  Here the arrays are sorted by key (everything this
// library builds is; arrays loaded from a reference-written .skf are sorted on first use), so every join is a binary search
// and every result is a stream compaction.  All kernels take keys of one or two 64-bit words (k <= 31 / k <= 63).
// Roofline: HBM; one pass over the rows touched plus log2(R) probes of the (L2-resident for R <= 1.5e7) key array per lookup.
#include <algorithm>

#include "common.cuh"

namespace ska {

struct Key128 { uint64_t lo, hi; };
__device__ __forceinline__ Key128 load_key(const uint64_t* __restrict__ keys, int kw, uint64_t i) {
    Key128 k;
    if (kw == 1) { k.lo = keys[i]; k.hi = 0; } else { k.lo = keys[2 * i]; k.hi = keys[2 * i + 1]; }
    return k;
}
__device__ __forceinline__ bool key_less(const Key128& a, const Key128& b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }
__device__ __forceinline__ bool key_eq(const Key128& a, const Key128& b) { return a.hi == b.hi && a.lo == b.lo; }
// first index with keys[i] >= key
__device__ __forceinline__ uint64_t lower_bound_key(const uint64_t* __restrict__ keys, int kw, uint64_t n, const Key128& key) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) { const uint64_t m = (lo + hi) >> 1; if (key_less(load_key(keys, kw, m), key)) lo = m + 1; else hi = m; }
    return lo;
}

// ---------------------------------------------------------------------------------------------- order
__global__ void order_check_kernel(const uint64_t* __restrict__ keys, int kw, uint64_t n, uint32_t* __restrict__ bad) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 >= n) return;
    if (!key_less(load_key(keys, kw, i), load_key(keys, kw, i + 1))) atomicExch(bad, 1u);
}
bool keys_sorted(const Array& a) {
    Context& c = ctx();
    if (a.n_keys < 2) return true;
    DevBuf<uint32_t> bad(1);
    dev_fill(bad.p, 0, 4);
    SKA_LAUNCH(order_check_kernel, (unsigned)((a.n_keys + 255) / 256), 256, 0, a.keys.p, (int)(a.key_bits / 64), a.n_keys, bad.p);
    uint32_t h = 0;
    SKA_CUDA(cudaMemcpyAsync(&h, bad.p, 4, cudaMemcpyDeviceToHost, c.stream));
    SKA_CUDA(cudaStreamSynchronize(c.stream));
    return h == 0;
}

// bitonic sort of row indices by key, in global memory (n padded to a power of two with "infinite" keys = index >= n).
// Only arrays written by the reference arrive unsorted (hashbrown iteration order): once per load, 300 small passes for 1e7 rows.
__global__ void bitonic_step_kernel(uint32_t* __restrict__ idx, const uint64_t* __restrict__ keys, int kw, uint64_t n, uint64_t np2,
                                    uint64_t size, uint64_t stride) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= np2 / 2) return;
    const uint64_t lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
    const bool up = (lo & size) == 0;
    const uint32_t a = idx[lo], b = idx[hi];
    // padding indices (>= n) sort after everything; ties cannot occur among real rows (keys are unique)
    bool a_gt_b;
    if (a >= n || b >= n) a_gt_b = a >= n && b < n ? true : (b >= n && a < n ? false : a > b);
    else a_gt_b = key_less(load_key(keys, kw, b), load_key(keys, kw, a));
    if (a_gt_b == up) { idx[lo] = b; idx[hi] = a; }
}
__global__ void iota_kernel(uint32_t* __restrict__ idx, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) idx[i] = (uint32_t)i;
}
// out row i = in row idx[i]
__global__ void __launch_bounds__(256)
permute_rows_kernel(const uint32_t* __restrict__ idx, const uint64_t* __restrict__ keys, int kw, const uint8_t* __restrict__ var,
                    const uint32_t* __restrict__ cnt, uint64_t n_rows, uint32_t n_samples, uint64_t* __restrict__ out_keys,
                    uint8_t* __restrict__ out_var, uint32_t* __restrict__ out_cnt) {
    const int lane = threadIdx.x & 31;
    const uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n_rows) return;
    const uint64_t s = idx[r];
    if (lane < kw) out_keys[r * kw + lane] = keys[s * kw + lane];
    if (lane == 2) out_cnt[r] = cnt[s];
    for (uint32_t c = lane; c < n_samples; c += 32) out_var[r * n_samples + c] = var[s * n_samples + c];
}

void ensure_sorted(Array& a) {
    Context& c = ctx();
    if (a.n_keys != a.n_rows) throw Error("array keys and rows are out of step (filtered without update_kmers)");
    if (a.sorted || a.n_rows < 2) { a.sorted = true; return; }
    if (keys_sorted(a)) { a.sorted = true; return; }
    if (a.n_rows >= 0xFFFFFFFFull) throw Error("too many rows to sort");
    const int kw = (int)(a.key_bits / 64);
    const uint64_t n = a.n_rows;
    uint64_t np2 = 1; while (np2 < n) np2 <<= 1;
    DevBuf<uint32_t> idx(np2);
    SKA_LAUNCH(iota_kernel, (unsigned)((np2 + 255) / 256), 256, 0, idx.p, np2);
    for (uint64_t size = 2; size <= np2; size <<= 1)
        for (uint64_t stride = size >> 1; stride > 0; stride >>= 1)
            SKA_LAUNCH(bitonic_step_kernel, (unsigned)((np2 / 2 + 255) / 256), 256, 0, idx.p, a.keys.p, kw, n, np2, size, stride);
    DevBuf<uint64_t> nk(n * kw); DevBuf<uint8_t> nv(n * a.n_samples); DevBuf<uint32_t> nc(n);
    SKA_LAUNCH(permute_rows_kernel, (unsigned)((n * 32 + 255) / 256), 256, 0, idx.p, a.keys.p, kw, a.variants.p, a.counts.p, n,
               (uint32_t)a.n_samples, nk.p, nv.p, nc.p);
    SKA_CUDA(cudaStreamSynchronize(c.stream));
    a.keys = std::move(nk); a.variants = std::move(nv); a.counts = std::move(nc);
    a.sorted = true;
}

// ---------------------------------------------------------------------------------------------- compaction by flags
constexpr int CP_ROWS = 256;        // rows per CTA
__global__ void __launch_bounds__(256) keep_count_kernel(const uint8_t* __restrict__ keep, uint64_t n, uint32_t* __restrict__ cta_cnt) {
    const uint64_t r = (uint64_t)blockIdx.x * CP_ROWS + threadIdx.x;
    const int k = r < n && keep[r];
    const int tot = __syncthreads_count(k);
    if (threadIdx.x == 0) cta_cnt[blockIdx.x] = (uint32_t)tot;
}
__global__ void __launch_bounds__(256)
keep_move_kernel(const uint8_t* __restrict__ keep, const uint64_t* __restrict__ cta_off, const uint64_t* __restrict__ keys, int kw,
                 const uint8_t* __restrict__ var, const uint32_t* __restrict__ cnt, uint64_t n, uint32_t n_samples, uint32_t src_stride,
                 uint64_t* __restrict__ out_keys, uint8_t* __restrict__ out_var, uint32_t* __restrict__ out_cnt) {
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_src[CP_ROWS];
    __shared__ uint32_t s_n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t r0 = (uint64_t)blockIdx.x * CP_ROWS, r = r0 + threadIdx.x;
    const bool k = r < n && keep[r];
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, k);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    uint32_t before = 0;
    for (int w = 0; w < warp; w++) before += s_warp[w];
    if (k) s_src[before + __popc(m & ((1u << lane) - 1u))] = threadIdx.x;
    if (threadIdx.x == 255) s_n = before + __popc(m);
    __syncthreads();
    const uint32_t nk = s_n;
    const uint64_t d0 = cta_off[blockIdx.x];
    for (uint32_t i = threadIdx.x; i < nk; i += 256) {
        const uint64_t s = r0 + s_src[i];
        for (int q = 0; q < kw; q++) out_keys[(d0 + i) * kw + q] = keys[s * kw + q];
        out_cnt[d0 + i] = cnt[s];
    }
    // rows: one warp per kept row
    for (uint32_t i = warp; i < nk; i += 8) {
        const uint8_t* src = var + (r0 + s_src[i]) * src_stride;
        uint8_t* dst = out_var + (d0 + i) * n_samples;
        for (uint32_t cidx = lane; cidx < n_samples; cidx += 32) dst[cidx] = src[cidx];
    }
}
// keeps the rows with keep[r] != 0 (keys, variants, counts), in order
void compact_rows(Array& a, const uint8_t* d_keep) {
    Context& c = ctx();
    if (a.n_keys != a.n_rows) throw Error("array keys and rows are out of step (filtered without update_kmers)");
    const uint64_t n = a.n_rows;
    if (n == 0) return;
    const int kw = (int)(a.key_bits / 64);
    const uint32_t n_cta = (uint32_t)((n + CP_ROWS - 1) / CP_ROWS);
    DevBuf<uint32_t> cta_cnt(n_cta), mx(1);
    DevBuf<uint64_t> cta_off(n_cta + 1);
    SKA_LAUNCH(keep_count_kernel, n_cta, 256, 0, d_keep, n, cta_cnt.p);
    launch_scan_counts(cta_cnt.p, n_cta, cta_off.p, mx.p);
    uint64_t kept = 0;
    SKA_CUDA(cudaMemcpyAsync(&kept, cta_off.p + n_cta, 8, cudaMemcpyDeviceToHost, c.stream));
    SKA_CUDA(cudaStreamSynchronize(c.stream));
    DevBuf<uint64_t> nk(kept * kw); DevBuf<uint8_t> nv(kept * a.n_samples); DevBuf<uint32_t> nc(kept);
    if (kept)
        SKA_LAUNCH(keep_move_kernel, n_cta, 256, 0, d_keep, cta_off.p, a.keys.p, kw, a.variants.p, a.counts.p, n, (uint32_t)a.n_samples,
                   (uint32_t)a.n_samples, nk.p, nv.p, nc.p);
    SKA_CUDA(cudaStreamSynchronize(c.stream));
    a.keys = std::move(nk); a.variants = std::move(nv); a.counts = std::move(nc);
    a.n_rows = a.n_keys = kept;
}

// ---------------------------------------------------------------------------------------------- delete_samples
// new row = the columns listed in keep_cols; count = its non-'-' cells; rows left empty are dropped (update_counts(false),
// merge_ska_array.rs:139-163)
__global__ void __launch_bounds__(256)
drop_columns_kernel(const uint8_t* __restrict__ var, uint64_t n_rows, uint32_t n_old, const uint32_t* __restrict__ keep_cols, uint32_t n_new,
                    uint8_t* __restrict__ out_var, uint32_t* __restrict__ out_cnt, uint8_t* __restrict__ keep) {
    const int lane = threadIdx.x & 31;
    const uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n_rows) return;
    uint32_t nz = 0;
    for (uint32_t cidx = lane; cidx < n_new; cidx += 32) {
        const uint8_t b = var[r * n_old + keep_cols[cidx]];
        out_var[r * n_new + cidx] = b;
        nz += b != '-';
    }
    nz = __reduce_add_sync(0xFFFFFFFFu, nz);
    if (lane == 0) { out_cnt[r] = nz; keep[r] = nz > 0; }
}
void delete_columns(Array& a, const std::vector<uint32_t>& del_cols) {
    Context& c = ctx();
    if (del_cols.empty() || del_cols.size() >= a.n_samples) throw Error("Invalid number of samples to remove");   // merge_ska_array.rs:232-234
    std::vector<char> gone(a.n_samples, 0);
    for (uint32_t d : del_cols) { if (d >= a.n_samples) throw Error("sample index out of range"); gone[d] = 1; }
    std::vector<uint32_t> keep_cols;
    for (uint32_t s = 0; s < a.n_samples; s++) if (!gone[s]) keep_cols.push_back(s);
    const uint32_t n_new = (uint32_t)keep_cols.size();
    if (a.n_keys != a.n_rows) throw Error("array keys and rows are out of step (filtered without update_kmers)");
    DevBuf<uint32_t> d_cols(n_new);
    SKA_CUDA(cudaMemcpyAsync(d_cols.p, keep_cols.data(), (size_t)n_new * 4, cudaMemcpyHostToDevice, c.stream));
    DevBuf<uint8_t> nv(a.n_rows * n_new), keep(std::max<uint64_t>(a.n_rows, 1));
    DevBuf<uint32_t> nc(std::max<uint64_t>(a.n_rows, 1));
    if (a.n_rows)
        SKA_LAUNCH(drop_columns_kernel, (unsigned)((a.n_rows * 32 + 255) / 256), 256, 0, a.variants.p, a.n_rows, (uint32_t)a.n_samples,
                   d_cols.p, n_new, nv.p, nc.p, keep.p);
    SKA_CUDA(cudaStreamSynchronize(c.stream));                    // keep_cols (host) is copied
    a.variants = std::move(nv); a.counts = std::move(nc); a.n_samples = n_new;
    compact_rows(a, keep.p);
}

// ---------------------------------------------------------------------------------------------- weed
// K1 records of the weed sequences (key << 4 | one-hot middle; all-ones = no window) -> mark the rows whose key occurs
template <int KW>
__global__ void mark_records_kernel(const uint64_t* __restrict__ rec, uint64_t n_rec, const uint64_t* __restrict__ keys, uint64_t n_keys,
                                    uint8_t* __restrict__ mark, unsigned long long* __restrict__ n_valid) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rec) return;
    Key128 key;
    if (KW == 1) {
        const uint64_t r = rec[i];
        if (r == REC_SENTINEL) return;
        key.lo = r >> 4; key.hi = 0;
    } else {
        const uint64_t lo = rec[2 * i], hi = rec[2 * i + 1];
        if ((lo & hi) == REC_SENTINEL) return;
        key.lo = (lo >> 4) | (hi << 60); key.hi = hi >> 4;
    }
    atomicAdd(n_valid, 1ULL);
    const uint64_t p = lower_bound_key(keys, KW, n_keys, key);
    if (p < n_keys && key_eq(load_key(keys, KW, p), key)) mark[p] = 1;
}
__global__ void flip_kernel(uint8_t* __restrict__ f, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) f[i] = !f[i];
}
// removes (reverse: keeps only) the rows whose split k-mer occurs in the weed sequences; -> number of weed split k-mers
// (windows, as RefSka::ksize counts them)
uint64_t weed_rows(Array& a, const SeqBatch& weed, bool reverse) {
    Context& c = ctx();
    ensure_sorted(a);
    arena().reset();
    const bool wide = a.k > 31;
    if (wide && !k1_wide_available()) throw Error("k > 31 (128-bit split k-mers): the extraction kernel for wide keys is not built");
    const uint64_t n_pos = weed.n_pos;
    uint64_t* rec = arena().take_n<uint64_t>(std::max<uint64_t>(n_pos, 1) * (wide ? 2 : 1));
    launch_k1(weed, a.k, a.rc, false, 0, rec, nullptr, nullptr);
    DevBuf<uint8_t> mark(std::max<uint64_t>(a.n_rows, 1));
    DevBuf<unsigned long long> nv(1);
    dev_fill(mark.p, 0, std::max<uint64_t>(a.n_rows, 1));
    dev_fill(nv.p, 0, 8);
    if (n_pos) {
        if (wide) SKA_LAUNCH((mark_records_kernel<2>), (unsigned)((n_pos + 255) / 256), 256, 0, rec, n_pos, a.keys.p, a.n_keys, mark.p, nv.p);
        else SKA_LAUNCH((mark_records_kernel<1>), (unsigned)((n_pos + 255) / 256), 256, 0, rec, n_pos, a.keys.p, a.n_keys, mark.p, nv.p);
    }
    unsigned long long h_nv = 0;
    SKA_CUDA(cudaMemcpyAsync(&h_nv, nv.p, 8, cudaMemcpyDeviceToHost, c.stream));
    SKA_CUDA(cudaStreamSynchronize(c.stream));
    if (h_nv == 0) throw Error("weed file has no valid sequence");                    // RefSka::new (ska_ref.rs:253-255)
    if (!reverse && a.n_rows) SKA_LAUNCH(flip_kernel, (unsigned)((a.n_rows + 255) / 256), 256, 0, mark.p, a.n_rows);   // keep = not found
    compact_rows(a, mark.p);
    return h_nv;
}

// ---------------------------------------------------------------------------------------------- extend (ska merge)
// union of two sorted key lists without duplicates inside either: with m[j] = (b_j occurs in A) and pm = exclusive prefix sum
// of m, key a_i lands at i + lb_B(a_i) - pm[lb_B(a_i)] and an unmatched b_j at j - pm[j] + lb_A(b_j); a matched b_j joins its a.
__global__ void match_b_kernel(const uint64_t* __restrict__ ka, uint64_t na, const uint64_t* __restrict__ kb, uint64_t nb, int kw,
                               uint32_t* __restrict__ m, uint64_t* __restrict__ lb_a) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nb) return;
    const Key128 key = load_key(kb, kw, j);
    const uint64_t p = lower_bound_key(ka, kw, na, key);
    lb_a[j] = p;
    m[j] = p < na && key_eq(load_key(ka, kw, p), key);
}
__global__ void __launch_bounds__(256)
place_rows_kernel(const uint64_t* __restrict__ keys, int kw, const uint8_t* __restrict__ var, const uint32_t* __restrict__ cnt, uint64_t n,
                  uint32_t cols, uint32_t col_base, int is_b, const uint64_t* __restrict__ other_keys, uint64_t n_other,
                  const uint64_t* __restrict__ pm /* nb + 1 */, const uint64_t* __restrict__ lb_a /* nb */, const uint32_t* __restrict__ m,
                  uint32_t n_total, uint64_t* __restrict__ out_keys, uint8_t* __restrict__ out_var, uint32_t* __restrict__ out_cnt) {
    const int lane = threadIdx.x & 31;
    const uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    uint64_t pos;
    if (!is_b) {
        const uint64_t lb = lower_bound_key(other_keys, kw, n_other, load_key(keys, kw, i));
        pos = i + lb - pm[lb];
    } else {
        pos = m[i] ? lb_a[i] + i - pm[i] : i - pm[i] + lb_a[i];     // matched: where its a went (a's lb_B is i, pm[i] matches before)
    }
    if (lane < kw) out_keys[pos * kw + lane] = keys[i * kw + lane];
    if (lane == 2) atomicAdd(&out_cnt[pos], cnt[i]);
    for (uint32_t cidx = lane; cidx < cols; cidx += 32) out_var[pos * n_total + col_base + cidx] = var[i * cols + cidx];
}
void extend_arrays(Array& a, Array& b, Array& out) {
    Context& c = ctx();
    if (a.k != b.k) throw Error("K-mer lengths do not match: " + std::to_string(b.k) + " " + std::to_string(a.k));   // merge_ska_dict.rs:161-163
    if (a.rc != b.rc) throw Error("Strand use inconsistent");                                                        // :164-166
    if (a.key_bits != b.key_bits) throw Error("inconsistent key widths");
    ensure_sorted(a); ensure_sorted(b);
    const int kw = (int)(a.key_bits / 64);
    const uint64_t na = a.n_rows, nb = b.n_rows;
    const uint32_t nt = (uint32_t)(a.n_samples + b.n_samples);
    DevBuf<uint32_t> m(std::max<uint64_t>(nb, 1)), mx(1);
    DevBuf<uint64_t> lb_a(std::max<uint64_t>(nb, 1)), pm(nb + 1);
    if (nb) SKA_LAUNCH(match_b_kernel, (unsigned)((nb + 255) / 256), 256, 0, a.keys.p, na, b.keys.p, nb, kw, m.p, lb_a.p);
    uint64_t n_match = 0;
    if (nb) {
        if (nb >= 0xFFFFFFFFull) throw Error("too many rows");
        launch_scan_counts(m.p, (uint32_t)nb, pm.p, mx.p);
        SKA_CUDA(cudaMemcpyAsync(&n_match, pm.p + nb, 8, cudaMemcpyDeviceToHost, c.stream));
    } else dev_fill(pm.p, 0, 8);
    SKA_CUDA(cudaStreamSynchronize(c.stream));
    const uint64_t R = na + nb - n_match;
    out.k = a.k; out.rc = a.rc; out.key_bits = a.key_bits; out.n_samples = nt; out.n_rows = out.n_keys = R; out.sorted = true;
    out.keys.alloc(R * kw); out.variants.alloc(R * nt); out.counts.alloc(R);
    if (R) {
        dev_fill(out.variants.p, '-', R * nt);            // 0 -> '-' (merge_ska_array.rs:175)
        dev_fill(out.counts.p, 0, R * 4);
    }
    if (na) SKA_LAUNCH(place_rows_kernel, (unsigned)((na * 32 + 255) / 256), 256, 0, a.keys.p, kw, a.variants.p, a.counts.p, na, (uint32_t)a.n_samples,
                       0u, 0, b.keys.p, nb, pm.p, lb_a.p, m.p, nt, out.keys.p, out.variants.p, out.counts.p);
    if (nb) SKA_LAUNCH(place_rows_kernel, (unsigned)((nb * 32 + 255) / 256), 256, 0, b.keys.p, kw, b.variants.p, b.counts.p, nb, (uint32_t)b.n_samples,
                       (uint32_t)a.n_samples, 1, a.keys.p, na, pm.p, lb_a.p, m.p, nt, out.keys.p, out.variants.p, out.counts.p);
    SKA_CUDA(cudaStreamSynchronize(c.stream));
}

// ---------------------------------------------------------------------------------------------- map (RefSka::map)
// RC_IUPAC (bit_encoding.rs:474-507, the literal 256-entry table): complement of an IUPAC letter, lower case answered in
// upper case; every other byte -- 'U' and '-' included -- gives '-'
__device__ __forceinline__ uint8_t rc_iupac(uint8_t b) {
    switch (b & 0xDF) {                                             // fold case ('-' & 0xDF = 0x0D falls to the default)
        case 'A': return 'T'; case 'B': return 'V'; case 'C': return 'G'; case 'D': return 'H'; case 'G': return 'C';
        case 'H': return 'D'; case 'K': return 'M'; case 'M': return 'K'; case 'N': return 'N'; case 'R': return 'Y';
        case 'S': return 'S'; case 'T': return 'A'; case 'V': return 'B'; case 'W': return 'W'; case 'Y': return 'R';
        default: return '-';
    }
}
// per reference position p (window start): hit[p] = row of its split k-mer + 1 (0: no window / not in the array), bit 63 of the
// word = the reference k-mer was reverse-complemented to its canonical form (the record's middle base differs from the sequence's)
template <int KW>
__global__ void map_lookup_kernel(const uint64_t* __restrict__ rec, uint64_t n_pos, const uint8_t* __restrict__ seq, int h,
                                  const uint64_t* __restrict__ keys, uint64_t n_keys, uint64_t* __restrict__ hit, uint8_t* __restrict__ flag,
                                  unsigned long long* __restrict__ n_valid) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pos) return;
    Key128 key; uint32_t onehot;
    if (KW == 1) {
        const uint64_t r = rec[p];
        if (r == REC_SENTINEL) { hit[p] = 0; flag[p] = 0; return; }
        key.lo = r >> 4; key.hi = 0; onehot = (uint32_t)(r & 15);
    } else {
        const uint64_t lo = rec[2 * p], hi = rec[2 * p + 1];
        if ((lo & hi) == REC_SENTINEL) { hit[p] = 0; flag[p] = 0; return; }
        key.lo = (lo >> 4) | (hi << 60); key.hi = hi >> 4; onehot = (uint32_t)(lo & 15);
    }
    atomicAdd(n_valid, 1ULL);
    const uint32_t mid = (uint32_t)__ffs(onehot) - 1u;
    const uint32_t fwd_mid = (seq[FRONT_PAD + p + h] >> 1) & 3u;
    const uint64_t q = lower_bound_key(keys, KW, n_keys, key);
    const bool found = q < n_keys && key_eq(load_key(keys, KW, q), key);
    hit[p] = found ? ((q + 1) | (mid != fwd_mid ? (1ULL << 63) : 0ULL)) : 0ULL;
    flag[p] = found;
}
// mapped row m (the m-th hit in reference order) = array row, complemented if the reference k-mer is on the reverse strand
__global__ void __launch_bounds__(256)
map_rows_kernel(const uint64_t* __restrict__ hit, const uint64_t* __restrict__ hit_pos /* compacted positions */, uint64_t n_hits,
                const uint8_t* __restrict__ var, uint32_t n_samples, uint8_t* __restrict__ out_var) {
    const int lane = threadIdx.x & 31;
    const uint64_t mrow = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (mrow >= n_hits) return;
    const uint64_t w = hit[hit_pos[mrow]];
    const bool is_rc = w >> 63;
    const uint64_t r = (w & ~(1ULL << 63)) - 1;
    for (uint32_t cidx = lane; cidx < n_samples; cidx += 32) {
        const uint8_t b = var[r * n_samples + cidx];
        out_var[mrow * n_samples + cidx] = is_rc ? rc_iupac(b) : b;
    }
}
__global__ void __launch_bounds__(256)
hit_positions_kernel(const uint8_t* __restrict__ flag, const uint64_t* __restrict__ cta_off, uint64_t n, uint64_t* __restrict__ out_pos) {
    __shared__ uint32_t s_warp[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t p = (uint64_t)blockIdx.x * CP_ROWS + threadIdx.x;
    const bool k = p < n && flag[p];
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, k);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    uint32_t before = 0;
    for (int w = 0; w < warp; w++) before += s_warp[w];
    if (k) out_pos[cta_off[blockIdx.x] + before + __popc(m & ((1u << lane) - 1u))] = p;
}
// ref: the reference's records as one sample (ASCII batch).  -> positions (window starts in the batch's position space) of the
// reference split k-mers found in the array, in reference order, and their (complemented where needed) rows
void map_reference(Array& a, const SeqBatch& ref, std::vector<uint64_t>& h_pos, DevBuf<uint8_t>& mapped, uint64_t& n_ref_kmers,
                   std::vector<uint64_t>* h_records) {
    Context& c = ctx();
    if (ref.packed || ref.has_qual) throw Error("the reference must be an ASCII FASTA batch");
    ensure_sorted(a);
    arena().reset();
    const bool wide = a.k > 31;
    if (wide && !k1_wide_available()) throw Error("k > 31 (128-bit split k-mers): the extraction kernel for wide keys is not built");
    const uint64_t n_pos = ref.n_pos;
    const int h = ((int)a.k - 1) / 2;
    uint64_t* rec = arena().take_n<uint64_t>(std::max<uint64_t>(n_pos, 1) * (wide ? 2 : 1));
    launch_k1(ref, a.k, a.rc, false, 0, rec, nullptr, nullptr);
    DevBuf<uint64_t> hit(std::max<uint64_t>(n_pos, 1));
    DevBuf<uint8_t> flag(std::max<uint64_t>(n_pos, 1));
    DevBuf<unsigned long long> nv(1);
    dev_fill(nv.p, 0, 8);
    if (n_pos) {
        if (wide) SKA_LAUNCH((map_lookup_kernel<2>), (unsigned)((n_pos + 255) / 256), 256, 0, rec, n_pos, ref.seq.p, h, a.keys.p, a.n_keys, hit.p, flag.p, nv.p);
        else SKA_LAUNCH((map_lookup_kernel<1>), (unsigned)((n_pos + 255) / 256), 256, 0, rec, n_pos, ref.seq.p, h, a.keys.p, a.n_keys, hit.p, flag.p, nv.p);
    }
    const uint32_t n_cta = (uint32_t)((std::max<uint64_t>(n_pos, 1) + CP_ROWS - 1) / CP_ROWS);
    DevBuf<uint32_t> cta_cnt(n_cta), mx(1);
    DevBuf<uint64_t> cta_off(n_cta + 1);
    SKA_LAUNCH(keep_count_kernel, n_cta, 256, 0, flag.p, n_pos, cta_cnt.p);
    launch_scan_counts(cta_cnt.p, n_cta, cta_off.p, mx.p);
    uint64_t n_hits = 0; unsigned long long h_nv = 0;
    SKA_CUDA(cudaMemcpyAsync(&n_hits, cta_off.p + n_cta, 8, cudaMemcpyDeviceToHost, c.stream));
    SKA_CUDA(cudaMemcpyAsync(&h_nv, nv.p, 8, cudaMemcpyDeviceToHost, c.stream));
    SKA_CUDA(cudaStreamSynchronize(c.stream));
    n_ref_kmers = h_nv;
    if (h_records) {                                       // the reference's own split k-mers (RefSka::new keeps them: repeat mask, ska weed)
        h_records->resize(n_pos * (wide ? 2 : 1));
        if (n_pos) SKA_CUDA(cudaMemcpyAsync(h_records->data(), rec, h_records->size() * 8, cudaMemcpyDeviceToHost, c.stream));
    }
    DevBuf<uint64_t> pos(std::max<uint64_t>(n_hits, 1));
    mapped.alloc(n_hits * a.n_samples);
    h_pos.resize(n_hits);
    if (n_hits) {
        SKA_LAUNCH(hit_positions_kernel, n_cta, 256, 0, flag.p, cta_off.p, n_pos, pos.p);
        SKA_LAUNCH(map_rows_kernel, (unsigned)((n_hits * 32 + 255) / 256), 256, 0, hit.p, pos.p, n_hits, a.variants.p, (uint32_t)a.n_samples, mapped.p);
        SKA_CUDA(cudaMemcpyAsync(h_pos.data(), pos.p, n_hits * 8, cudaMemcpyDeviceToHost, c.stream));
    }
    SKA_CUDA(cudaStreamSynchronize(c.stream));
}

}  // namespace ska
