// verify.cu -- self-checks of a device-resident array, used by bench.py and the tests to prove at FULL size what a CPU
// restatement can only check on samples:
//   * an order-independent checksum over the cells (key, column - col_lo, letter) of a column range, '-' cells
//     excluded.  The rows of a rank's local build and the same columns of the joined slabs (all ranks summed) must
//     give the same value: MergeSkaDict::merge (src/merge_ska_dict.rs:119-151) only moves cells, it never changes them;
//   * rows whose stored variant_count differs from the number of non-'-' cells (merge_ska_array.rs:172);
//   * adjacent keys out of ascending order (the canonical row order of this library).
#include <algorithm>

#include "common.cuh"

namespace ska {

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

// one warp per row
__global__ void __launch_bounds__(256)
array_checksum_kernel(const uint64_t* __restrict__ keys, int key_words, const uint8_t* __restrict__ var, const uint32_t* __restrict__ cnt,
                      uint64_t n_rows, uint32_t n_samples, uint32_t col_lo, uint32_t col_hi, unsigned long long* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    unsigned long long sum = 0, rows_present = 0, bad_cnt = 0, bad_order = 0;
    for (uint64_t r = warp0; r < n_rows; r += n_warps) {
        uint64_t key = keys[r * key_words];
        if (key_words == 2) key = mix64(key) ^ keys[r * 2 + 1];
        const uint8_t* row = var + r * n_samples;
        uint32_t present = 0, in_range = 0;
        unsigned long long s = 0;
        for (uint32_t c = lane; c < n_samples; c += 32) {
            const uint32_t b = row[c];
            if (b == '-') continue;
            present++;
            if (c >= col_lo && c < col_hi) { in_range++; s += mix64(key ^ mix64(((uint64_t)(c - col_lo) << 8) | b)); }
        }
        present = __reduce_add_sync(0xFFFFFFFFu, present);
        in_range = __reduce_add_sync(0xFFFFFFFFu, in_range);
        sum += s;
        if (lane == 0) {
            rows_present += in_range != 0;
            bad_cnt += cnt[r] != present;
            if (r > 0) {
                if (key_words == 1) bad_order += keys[r - 1] >= keys[r];
                else {
                    const uint64_t plo = keys[2 * r - 2], phi = keys[2 * r - 1], lo = keys[2 * r], hi = keys[2 * r + 1];
                    bad_order += (phi > hi) || (phi == hi && plo >= lo);
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    if (lane == 0) {
        if (sum) atomicAdd(&out[0], sum);
        if (rows_present) atomicAdd(&out[1], rows_present);
        if (bad_cnt) atomicAdd(&out[2], bad_cnt);
        if (bad_order) atomicAdd(&out[3], bad_order);
    }
}

void array_checksum(const Array& a, uint32_t col_lo, uint32_t col_hi, uint64_t* h_out /* 4 */) {
    Context& c = ctx();
    if (a.n_keys != a.n_rows) throw Error("array keys and rows are out of step (filtered without update_kmers)");
    if (col_hi > a.n_samples || col_lo > col_hi) throw Error("bad column range");
    DevBuf<unsigned long long> d(4);
    dev_fill(d.p, 0, 32);
    if (a.n_rows) {
        const unsigned grid = (unsigned)std::min<uint64_t>((a.n_rows + 7) / 8, (uint64_t)c.sm_count * 16);
        SKA_LAUNCH(array_checksum_kernel, grid, 256, 0, a.keys.p, (int)(a.key_bits / 64), a.variants.p, a.counts.p, a.n_rows,
                   (uint32_t)a.n_samples, col_lo, col_hi, d.p);
    }
    SKA_CUDA(cudaMemcpyAsync(h_out, d.p, 32, cudaMemcpyDeviceToHost, c.stream));
    SKA_CUDA(cudaStreamSynchronize(c.stream));
}

}  // namespace ska
