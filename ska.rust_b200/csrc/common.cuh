// common.cuh -- shared host/device helpers of libskagpu (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

namespace ska {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define SKA_CUDA(call)                                                                                    \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess)                                                                           \
            throw ska::Error(std::string(#call) + " failed: " + cudaGetErrorString(e__) + " (" + __FILE__ + \
                             ":" + std::to_string(__LINE__) + ")");                                       \
    } while (0)

// One context per process (one process per GPU).
struct Context {
    int device = 0;
    bool init = false;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // host -> device copies that overlap the extraction kernel (skagpu_build)
    int sm_count = 148;
    size_t smem_optin = 0;
    uint64_t launches = 0;
    void ensure();
};
Context& ctx();
void bind_thread_to_device(int device);   // the calling host thread works on this device from now on (api.cu)
void*& comm_slot();                       // the multi-GPU communicator of the calling thread's device (comm.cu)

// Every kernel launch goes through this so the launch count claimed in bench.py is real.
#define SKA_LAUNCH(kernel, grid, block, smem, ...)                         \
    do {                                                                   \
        kernel<<<(grid), (block), (smem), ska::ctx().stream>>>(__VA_ARGS__); \
        ska::ctx().launches++;                                             \
        SKA_CUDA(cudaGetLastError());                                      \
    } while (0)

// Device memory of DevBuf: blocks from cudaMalloc kept in a free list for the life of the process (api.cu).  The first
// growth of CUDA's stream-ordered pool costs 0.5 s for 1.4 GB on a B200 (profiles/microbench/ctx_mb.cu) against 7 ms for a
// 13 GB cudaMalloc -- a third of what a one-shot `ska build` process takes; repeated builds hit the free list either way.
// A block is handed out again only to work that is ordered behind its previous user (everything runs on, or is fenced
// against, the library stream).
// small device -> host read-back on the library stream that does not queue behind bulk DMA (api.cu); the bytes are in h_dst
// after the next read_back_sync(), which also synchronises the stream
void dev_fill(void* p, int byte, size_t bytes);   // memset on the library stream, as a kernel (api.cu)
void read_back(void* h_dst, const void* d_src, size_t bytes);
void read_back_sync();
void* dev_alloc(size_t bytes);
void dev_free(void* p);

// RAII device buffer.
template <class T> struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count) {
        release();
        n = count;
        if (count) p = static_cast<T*>(dev_alloc(count * sizeof(T)));
    }
    void release() {
        if (p) dev_free(p);
        p = nullptr; n = 0;
    }
    size_t bytes() const { return n * sizeof(T); }
};

// Grow-only workspace for the temporaries of a build (records, partition double buffers, bucket
// tables).  Allocating ~13 GB through the stream-ordered pool on every build cost up to tens of
// milliseconds at unpredictable moments; the arena is carved with a bump pointer and kept.
struct Arena {
    struct Slab { uint8_t* p; size_t cap, used; };
    std::vector<Slab> slabs;
    void reset();                       // start of a build: everything is free again (coalesces slabs)
    void* take(size_t bytes);
    template <class T> T* take_n(size_t n) { return static_cast<T*>(take(n * sizeof(T))); }
    void release_all();
};
Arena& arena();

// Stage timer: CUDA events on the library stream.
struct StageTimer {
    struct Stage { const char* name; cudaEvent_t beg, end; uint64_t launches; };
    std::vector<Stage> stages;
    uint64_t launch_mark = 0;
    bool muted = false;                 // inner builds (rowmerge.cu's key union) run under the caller's stage
    void clear();
    void begin(const char* name);
    void end();
    ~StageTimer() { clear(); }
};
StageTimer& timer();

static constexpr uint64_t REC_SENTINEL = ~0ULL;   // "no valid window starts here"
static constexpr int FRONT_PAD = 128;             // bytes of 'N' before position 0 of a batch

// Device-resident sequence batch: all samples' records concatenated in one position space.
struct SeqBatch {
    uint64_t n_pos = 0;               // total bases over all samples
    uint64_t n_samples = 0;
    bool has_qual = false;
    bool packed = false;              // host-packed batch (skagpu_build_packed): seq = 2-bit codes (4 positions per byte),
                                      // qual = invalid-base mask (8 positions per byte); samples start on 32-position boundaries
    DevBuf<uint8_t> seq;              // padded_positions(n_pos) bytes, pads are 'N'
    DevBuf<uint8_t> qual;             // same layout when has_qual
    DevBuf<uint32_t> start_bits;      // bit p set <=> a record starts at position p; bit n_pos set
    DevBuf<uint64_t> sample_off;      // n_samples + 1 positions
    std::vector<uint64_t> h_sample_off;
    std::vector<uint8_t> h_is_reads;  // has_qual batches: 1 = the sample's first file is FASTQ (count filter + quality tests, ska_dict.rs:356-366)
    std::vector<uint64_t> h_marks;    // record boundaries (host copy kept while an asynchronous upload reads it)
    cudaEvent_t ready = nullptr;      // asynchronous upload: recorded behind the last copy; builds wait for it
    SeqBatch() {}
    SeqBatch(const SeqBatch&) = delete;
    SeqBatch& operator=(const SeqBatch&) = delete;
    ~SeqBatch() { if (ready) { cudaEventSynchronize(ready); cudaEventDestroy(ready); } }
};

// Device-resident MergeSkaArray.
struct Array {
    uint32_t k = 0;
    bool rc = false;
    uint32_t key_bits = 64;
    uint64_t n_rows = 0, n_keys = 0, n_samples = 0;
    bool sorted = false;              // keys known to ascend (everything built here; loaded arrays are checked / sorted on first need)
    DevBuf<uint64_t> keys;            // n_keys * (key_bits/64) words, (lo,hi) for 128-bit
    DevBuf<uint8_t> variants;         // n_rows * n_samples, row-major
    DevBuf<uint32_t> counts;          // n_rows
};

// --- kernels' host launchers (one .cu each) ------------------------------------------------------
// k1_extract.cu
// hist != NULL (FASTA batches): also accumulate the first partition pass's histogram -- K1_HIST_COPIES copies of
// 2^hist_bits u64 counters (zeroed by the caller), digit = record >> hist_shift
constexpr int K1_HIST_COPIES = 32;
constexpr int K1_HIST_MAX_BITS = 11;
void launch_k1(const SeqBatch& b, uint32_t k, bool rc, bool strict_qual, uint8_t min_qual, uint64_t* rec_out,
               uint16_t* sid_out, uint32_t* sample_valid, uint32_t sid_base = 0, unsigned long long* hist = nullptr,
               int hist_shift = 0, int hist_bits = 0, uint64_t tile_begin = 0, uint64_t tile_end = ~0ULL /* tiles of 4096 positions */);
constexpr int K1_ASCII_TAIL = 96, K1_PACKED_TAIL = 128;   // staged positions after a tile (k1_helpers.cuh)
uint64_t k1_tiles_ready(uint64_t positions_uploaded, int tail);
size_t padded_positions(uint64_t n_pos);
bool k1_wide_available();          // launch_k1 handles k > 31 (two-word records)
// partition.cu: how records are grouped and packed.  P = P1 + P2 prefix bits of the key select the bucket;
// inside a bucket a record is  key_low << (sb+2) | sample_id << 2 | middle_code  (key_low = key without its top P1 bits)
struct PartPlan { int key_bits, sb, P, P1, P2, low_bits; bool wide; /* k > 31: records of two words */ };
PartPlan make_part_plan(uint32_t k, int prefix_bits, uint64_t n_samples_total);
// k1_hist: the first pass's histogram as K1 left it (K1_HIST_COPIES copies), or NULL to count here
void partition_records(const uint64_t* rec, uint64_t n_pos, const uint64_t* d_sample_off, uint32_t n_samples_local, uint32_t sid_base,
                       const PartPlan& pl, uint64_t* scratch, uint64_t** out, uint64_t** d_off, uint64_t** first_pass_buf = nullptr,
                       const unsigned long long* k1_hist = nullptr);
// first pass only (what a rank does before the exchange) / second pass over received runs (what the owner does after it)
void partition_coarse(const uint64_t* rec, uint64_t n_pos, const uint64_t* d_sample_off, uint32_t n_samples_local, uint32_t sid_base,
                      const PartPlan& pl, uint64_t** out, uint64_t** d_off1, uint64_t** d_tile_pref, const unsigned long long* k1_hist = nullptr,
                      uint8_t** d_next_digit = nullptr /* -> one byte per record: its second-pass digit (NULL when not produced) */);
void partition_runs(const uint64_t* src, int n_runs, const uint64_t* h_run_off, uint32_t n_regions, const PartPlan& pl, uint64_t* dst,
                    uint64_t** d_bucket_off);
// plain MSD grouping of 16-byte records by P1 + P2 bits below `top_bit` (all-ones records dropped); sort128.cu finishes the sort
void group_wide_records(const uint64_t* src, uint64_t n, int top_bit, int P1, int P2, uint64_t* buf_a, uint64_t* buf_b, uint64_t** out,
                        uint64_t** d_off);
// sort128.cu: ascending sort of n 16-byte records as 128-bit integers (all-ones records dropped); key_bits = significant bits of the
// HIGH word.  buf_a / buf_b hold n records each (rec may be buf_b).  -> pointer to the sorted records, *n_out = records kept
uint64_t* sort_records_wide(const uint64_t* rec, uint64_t n, int key_bits, uint64_t* buf_a, uint64_t* buf_b, uint64_t* n_out);
// the observation sort of the count filter: sorted (hash, position) records + one entry per distinct hash (see sort128.cu)
uint64_t* sort_observations(const uint64_t* rec, uint64_t n, uint64_t* buf_a, uint64_t* buf_b, uint64_t* n_obs, uint64_t** drec, uint64_t* ent,
                            uint64_t* run_start, unsigned long long* d_n_ent, uint32_t min_count, uint32_t bloom_words, bool write_sorted);
// comm.cu: the library's own multi-GPU transport (NCCL, dlopen'ed)
void comm_unique_id(uint8_t* id);
void comm_init(const uint8_t* id, int world, int rank);
void comm_destroy();
int comm_world();
int comm_rank();
void plan_owners(const uint64_t* counts, uint32_t n_regions, int world, uint64_t* cuts);
void allreduce_u64(uint64_t* h_vals, size_t n);
void exchange_partial_rows(const Array& local, uint64_t sample_base, uint64_t n_samples_total, Array& slab, uint64_t* stats);
// dedup.cu: one sample's K1 records -> its distinct records (see there)
uint64_t* unique_sample_records(uint64_t* rec, uint64_t n, uint32_t k, int prefix_bits, uint64_t* n_out);
// merge.cu
// K1 records (n_rec entries; sample s owns entries [sample_off[s], sample_off[s+1]); REC_SENTINEL entries are
// skipped) -> array; sample ids are sid_base + local index.  `rec` is overwritten (partition scratch).
void build_array_from_records(uint64_t* rec, uint64_t n_rec, const uint64_t* d_sample_off, uint32_t n_samples_local, uint32_t sid_base,
                              uint64_t n_samples_total, uint32_t k, bool rc, Array& out, int prefix_bits = -1 /* < 0: merge_prefix_bits */,
                              const unsigned long long* k1_hist = nullptr /* histogram of the plan's first pass, from K1 */);
// merge128.cu: the same for k > 31 (records and keys of two words)
int merge_prefix_bits_wide(uint64_t n_records_total, uint32_t k);
void build_array_from_records_wide(uint64_t* rec, uint64_t n_rec, const uint64_t* d_sample_off, uint32_t n_samples_local, uint32_t sid_base,
                                   uint64_t n_samples_total, uint32_t k, bool rc, Array& out, int prefix_bits = -1,
                                   const unsigned long long* k1_hist = nullptr);
// k1q_reads.cu: FASTQ samples (quality masks, ntHash, exact count-filter emulation) -> array
void build_array_from_reads(const SeqBatch& b, uint32_t k, bool rc, uint16_t min_count, uint8_t min_qual, int qual_filter, Array& out);
// row shard: `n_runs` received runs of first-pass records covering first-level regions [lo, hi); `prec` is overwritten
void build_array_from_runs(uint64_t* prec, uint64_t n_rec, int n_runs, const uint64_t* run_off, uint32_t region_lo, uint32_t region_hi,
                           const PartPlan& pl, uint64_t n_samples, uint32_t k, bool rc, Array& out);
// sorted distinct keys of `n_runs` ascending key runs lying back to back in `keys` (h_run_off: n_runs + 1 offsets);
// false if a key range is too crowded for the bucket tables (the caller then takes the general merge)
bool union_sorted_runs(const uint64_t* keys, int n_runs, const uint64_t* h_run_off, uint32_t k, DevBuf<uint64_t>& out_keys, uint64_t& n_out);
// prefix bits the merge wants for `n_records_total` records in the whole key space
int merge_prefix_bits(uint64_t n_records_total, uint32_t k, uint64_t n_samples_total);

// exclusive scan of n u32 counts into n+1 u64 offsets (off[n] = total); *max_out = max count
void launch_scan_counts(const uint32_t* cnt, uint32_t n, uint64_t* off, uint32_t* max_out);
// rowmerge.cu: multi-GPU exchange of partial rows (every rank builds its own samples' array; owners join on the key)
void key_cumulative(const Array& a, int region_bits, uint64_t* h_cum /* 2^region_bits + 1 */);
void merge_partial_rows(const uint64_t* keys, const uint8_t* var, const uint32_t* cnt, int n_runs, const uint64_t* run_rows,
                        const uint32_t* run_cols, const uint32_t* run_col_base, uint64_t n_samples, uint32_t k, bool rc, Array& out);
// filter.cu
int32_t filter_array(Array& a, uint64_t min_count, bool filter_ambig_as_missing, int filter_type, bool mask_ambig,
                     bool ignore_const_gaps, bool update_kmers);
void transpose_array(const Array& a, uint8_t* d_out);
// distance.cu
struct DistOut { double distance, mismatch_prop; uint64_t match_count, mismatch_count; };
void distance_array(const Array& a, double constant, bool filt_ambig, DistOut* h_out, bool sharded = false);
// raw per-pair counters {mismatches, matches, dist_int, n_both} (dist_int: differing unambiguous pairs, or
// 36 x summed overlap when !filt_ambig); partial sums over this array's rows, addable across row shards
void distance_counters(const Array& a, bool filt_ambig, uint64_t* h_counters);

// setops.cu: joins on the key (ska merge / delete / weed / map)
bool keys_sorted(const Array& a);
void ensure_sorted(Array& a);
void compact_rows(Array& a, const uint8_t* d_keep);
void delete_columns(Array& a, const std::vector<uint32_t>& del_cols);
uint64_t weed_rows(Array& a, const SeqBatch& weed, bool reverse);
void extend_arrays(Array& a, Array& b, Array& out);
void map_reference(Array& a, const SeqBatch& ref, std::vector<uint64_t>& h_pos, DevBuf<uint8_t>& mapped, uint64_t& n_ref_kmers,
                   std::vector<uint64_t>* h_records);
// cov.cu: distinct split k-mers of a read set by multiplicity (1000 bins), the counting half of `ska cov`
void kmer_count_histogram(const SeqBatch& reads, uint32_t k, bool rc, uint64_t* h_hist);
// verify.cu: {cell checksum over columns [col_lo, col_hi), rows with a cell there, rows with a wrong count, keys out of order}
void array_checksum(const Array& a, uint32_t col_lo, uint32_t col_hi, uint64_t* h_out);

}  // namespace ska
