// rowmerge.cu -- the owner side of the multi-GPU "partial rows" exchange (the local
// pre-merge variant): every rank builds the array of ITS samples (all rows, its own columns), rows travel to
// the rank that owns their key range, and the owner joins the partial rows of all source ranks on the key.
//
// Replaces MergeSkaDict::merge (src/merge_ska_dict.rs:119-151): two partial dictionaries over disjoint samples
// are joined on the split k-mer, a missing side stays 0 -> '-' (merge_ska_array.rs:175).  Here a "dictionary"
// is an array sorted by key, so the join is
//   1. union of the runs' keys      -> the slab's sorted distinct keys (partition + count pass of merge.cu on
//                                      one-column records; rows come out in ascending key order again)
//   2. rows = '-' everywhere, then every partial row is copied to (row of its key, its column block);
//      a tile of 256 consecutive rows of a run maps into a narrow window of the slab (both are sorted), so the
//      window's keys are staged in shared memory and each row finds its place by binary search there
//   3. variant_count of a row = sum of the partial rows' counts (columns are disjoint).
// Roofline: HBM; per partial row 8 + C + 4 B read and C B written (C = columns of its source rank), plus the
// N-byte '-' fill of every slab row.
#include "common.cuh"

namespace ska {

constexpr int RM_THREADS = 256;
constexpr int RM_WINDOW = 3072;          // slab keys staged per tile (24 KB)

__device__ __forceinline__ uint64_t lower_bound_dev(const uint64_t* __restrict__ a, uint64_t n, uint64_t key) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) { uint64_t m = (lo + hi) >> 1; if (a[m] < key) lo = m + 1; else hi = m; }
    return lo;
}

// cum[b] = number of keys < (b << shift) for b in [0, n_cuts - 1), cum[n_cuts - 1] = n   (keys ascending)
__global__ void key_cumulative_kernel(const uint64_t* __restrict__ keys, uint64_t n, int shift, uint32_t n_cuts, uint64_t* __restrict__ cum) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_cuts) return;
    cum[b] = (b + 1 == n_cuts) ? n : lower_bound_dev(keys, n, (uint64_t)b << shift);
}

void key_cumulative(const Array& a, int region_bits, uint64_t* h_cum) {
    Context& c = ctx();
    if (a.key_bits != 64) throw Error("key ranges are planned on 64-bit keys (k <= 31)");
    if (a.n_keys != a.n_rows) throw Error("array keys and rows are out of step (filtered without update_kmers)");
    const int key_bits = 2 * ((int)a.k - 1);
    if (region_bits < 0 || region_bits > key_bits || region_bits > 24) throw Error("bad region_bits");
    const uint32_t n_cuts = (1u << region_bits) + 1;
    DevBuf<uint64_t> d(n_cuts);
    SKA_LAUNCH(key_cumulative_kernel, (n_cuts + 255) / 256, 256, 0, a.keys.p, a.n_rows, key_bits - region_bits, n_cuts, d.p);
    read_back(h_cum, d.p, (size_t)n_cuts * 8);
    read_back_sync();
}

__global__ void keys_to_records_kernel(const uint64_t* __restrict__ keys, uint64_t n, uint64_t* __restrict__ rec) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rec[i] = (keys[i] << 4) | 1ULL;            // K1 record format: key << 4 | one-hot middle (any base: only the key matters)
}

__global__ void __launch_bounds__(RM_THREADS)
gather_rows_kernel(const uint64_t* __restrict__ run_keys, const uint8_t* __restrict__ run_var, const uint32_t* __restrict__ run_cnt,
                   uint64_t rows, uint32_t cols, uint32_t col_base,
                   const uint64_t* __restrict__ out_keys, uint64_t n_out, uint64_t n_samples,
                   uint8_t* __restrict__ out_var, uint32_t* __restrict__ out_cnt, uint32_t* __restrict__ missing) {
    __shared__ uint64_t s_win[RM_WINDOW];
    __shared__ uint64_t s_row[RM_THREADS];
    __shared__ uint64_t s_lo, s_hi;
    const uint64_t tile = (uint64_t)blockIdx.x * RM_THREADS;
    const uint32_t nt = (uint32_t)min((uint64_t)RM_THREADS, rows - tile);
    const uint32_t t = threadIdx.x;
    if (t == 0) s_lo = lower_bound_dev(out_keys, n_out, run_keys[tile]);
    if (t == 32) s_hi = lower_bound_dev(out_keys, n_out, run_keys[tile + nt - 1]) + 1;
    __syncthreads();
    const uint64_t lo = s_lo, hi = min(s_hi, n_out);
    const uint64_t window = hi > lo ? hi - lo : 0;
    const bool staged = window <= RM_WINDOW;
    if (staged) for (uint32_t i = t; i < window; i += RM_THREADS) s_win[i] = out_keys[lo + i];
    __syncthreads();
    if (t < nt) {
        const uint64_t key = run_keys[tile + t];
        uint64_t r;
        if (staged) {
            uint32_t a = 0, b = (uint32_t)window;
            while (a < b) { uint32_t m = (a + b) >> 1; if (s_win[m] < key) a = m + 1; else b = m; }
            r = lo + a;
        } else {
            r = lo + lower_bound_dev(out_keys + lo, window, key);
        }
        if (r >= n_out || out_keys[r] != key) { atomicAdd(missing, 1u); r = ~0ULL; }    // cannot happen: the slab keys are the union
        else atomicAdd(&out_cnt[r], run_cnt[tile + t]);
        s_row[t] = r;
    }
    __syncthreads();
    const uint8_t* src = run_var + tile * cols;
    const bool words = ((cols | col_base | (uint32_t)n_samples) & 3u) == 0 && ((reinterpret_cast<uintptr_t>(run_var) | reinterpret_cast<uintptr_t>(out_var)) & 3u) == 0;
    if (words) {
        const uint32_t wpr = cols >> 2, total = nt * wpr;
        const uint32_t* src32 = reinterpret_cast<const uint32_t*>(src);
        uint32_t* dst32 = reinterpret_cast<uint32_t*>(out_var);
        const uint64_t nsw = n_samples >> 2, cbw = col_base >> 2;
        for (uint32_t idx = t; idx < total; idx += RM_THREADS) {
            const uint32_t row = idx / wpr, w = idx - row * wpr;
            const uint64_t r = s_row[row];
            if (r != ~0ULL) dst32[r * nsw + cbw + w] = src32[idx];
        }
    } else {
        const uint32_t total = nt * cols;
        for (uint32_t idx = t; idx < total; idx += RM_THREADS) {
            const uint32_t row = idx / cols, cidx = idx - row * cols;
            const uint64_t r = s_row[row];
            if (r != ~0ULL) out_var[r * n_samples + col_base + cidx] = src[idx];
        }
    }
}

// keys / var / cnt: device buffers holding the runs back to back (run w: run_rows[w] keys, run_rows[w] x run_cols[w]
// bytes, run_rows[w] counts); every run ascending in key; column blocks [run_col_base[w], + run_cols[w]) disjoint.
void merge_partial_rows(const uint64_t* keys, const uint8_t* var, const uint32_t* cnt, int n_runs, const uint64_t* run_rows,
                        const uint32_t* run_cols, const uint32_t* run_col_base, uint64_t n_samples, uint32_t k, bool rc, Array& out) {
    Context& c = ctx();
    if (n_runs < 1) throw Error("row shard needs at least one run");
    uint64_t n_in = 0;
    for (int w = 0; w < n_runs; w++) {
        if ((uint64_t)run_col_base[w] + run_cols[w] > n_samples) throw Error("column block outside the slab");
        n_in += run_rows[w];
    }
    out.k = k; out.rc = rc; out.key_bits = 64; out.n_samples = n_samples; out.n_rows = out.n_keys = 0; out.sorted = true;
    if (n_in == 0) return;
    if (!keys || !var || !cnt) throw Error("null run buffers");

    // 1. union of the keys.  The runs are sorted, so key-range buckets are cut out of them by binary search
    //    (union_sorted_runs, merge.cu); the general partition + merge of one-column records is only the fallback
    //    for key ranges too crowded for the bucket tables.
    timer().begin("union_keys");
    std::vector<uint64_t> h_run_off(n_runs + 1, 0);
    for (int w = 0; w < n_runs; w++) h_run_off[w + 1] = h_run_off[w] + run_rows[w];
    Array uni;
    uint64_t n_uni = 0;
    const char* force_general = getenv("SKA_UNION_GENERAL");        // tests: force the fallback
    if ((force_general && atoi(force_general)) || !union_sorted_runs(keys, n_runs, h_run_off.data(), k, uni.keys, n_uni)) {
        uint64_t* rec = arena().take_n<uint64_t>(n_in);
        SKA_LAUNCH(keys_to_records_kernel, (unsigned)((n_in + 255) / 256), 256, 0, keys, n_in, rec);
        uint64_t* d_off = arena().take_n<uint64_t>(2);
        const uint64_t h_off[2] = {0, n_in};
        SKA_CUDA(cudaMemcpyAsync(d_off, h_off, 16, cudaMemcpyHostToDevice, c.stream));
        SKA_CUDA(cudaStreamSynchronize(c.stream));                   // h_off is a stack array
        int pbits = 0;
        while ((n_in >> pbits) > 1024 && pbits < 22) pbits++;
        timer().muted = true;
        try { build_array_from_records(rec, n_in, d_off, 1, 0, 1, k, rc, uni, pbits); }
        catch (...) { timer().muted = false; throw; }
        timer().muted = false;
        n_uni = uni.n_rows;
    }
    timer().end();

    // 2. + 3.
    timer().begin("gather_rows");
    const uint64_t R = n_uni;
    out.n_rows = out.n_keys = R;
    out.keys = std::move(uni.keys);
    out.variants.alloc(R * n_samples);
    out.counts.alloc(R);
    if (R * n_samples) dev_fill(out.variants.p, '-', R * n_samples);
    if (R) dev_fill(out.counts.p, 0, R * 4);
    uint32_t* missing = arena().take_n<uint32_t>(1);
    dev_fill(missing, 0, 4);
    uint64_t row0 = 0, byte0 = 0;
    for (int w = 0; w < n_runs; w++) {
        const uint64_t rows = run_rows[w];
        if (rows && run_cols[w])
            SKA_LAUNCH(gather_rows_kernel, (unsigned)((rows + RM_THREADS - 1) / RM_THREADS), RM_THREADS, 0, keys + row0, var + byte0, cnt + row0,
                       rows, run_cols[w], run_col_base[w], out.keys.p, R, n_samples, out.variants.p, out.counts.p, missing);
        row0 += rows; byte0 += rows * run_cols[w];
    }
    uint32_t h_missing = 0;
    read_back(&h_missing, missing, 4);
    read_back_sync();
    timer().end();
    if (h_missing) throw Error("internal: " + std::to_string(h_missing) + " partial rows found no slab row");
}

}  // namespace ska
