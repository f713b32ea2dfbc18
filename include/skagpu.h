/* skagpu.h -- C ABI of libskagpu.so, the B200 (sm_100a) split-k-mer engine.
 *
 * Drop-in boundary for the hot path of bacpop/ska.rust v0.5.2.  The
 * reference has no FFI for this path; each entry point below replaces one Rust call site, cited as
 * file:line in /root/reference.  Plain C: a Rust host binds it with an `extern "C"` block
 * (INTEGRATION.md shows the block), this repo's C++ host (ska.rust_b200/host) and Python
 * (ska.rust_b200/api.py, ctypes) bind the same symbols.
 *
 * Conventions (mirroring the reference's ownership/error behaviour):
 *  - every function returns 0 on success, non-zero on failure; the message is then available from
 *    skagpu_last_error() (thread-local).  Nothing aborts; where the reference panics (invalid k,
 *    empty sample, ...) the same condition is reported here as an error.
 *  - ska_array_t is an opaque, device-resident MergeSkaArray (merge_ska_array.rs:108-126):
 *    created by skagpu_build / skagpu_array_from_host, mutated in place by skagpu_filter
 *    (`&mut self`), read by the rest (`&self`), released by skagpu_free.
 *  - all pointer arguments are HOST pointers unless the name says `_dev`; host buffers may be
 *    pageable or pinned (pinned is faster; skagpu_host_alloc returns pinned memory).
 *  - row order of a built array is ascending split-k-mer key (the reference's order is hashbrown
 *    iteration order, i.e. unspecified; its tests compare as sets).
 *  - single calling thread per handle; the library owns its CUDA stream(s).
 */
#ifndef SKAGPU_H
#define SKAGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ska_array ska_array_t;
typedef struct ska_seqbatch ska_seqbatch_t;
typedef struct ska_records ska_records_t;
typedef struct ska_mapping ska_mapping_t;

/* One sample = InputFastx (merge_ska_dict.rs:25) after host-side parsing (needletail's job,
 * ska_dict.rs:131-147): all records of file 1 followed by all records of file 2, newlines
 * stripped, concatenated; rec_off has n_rec+1 entries.  qual == NULL <=> the sample's first file is
 * FASTA (is_reads = false, ska_dict.rs:356-366); FASTA and FASTQ samples may share one build.  A
 * record without a quality string inside a read sample (a FASTA second file) carries quality bytes
 * 0xFF, which every quality test passes (split_kmer.rs:66-71, `None => true`). */
typedef struct {
    const uint8_t* seq;
    const uint8_t* qual;
    const uint64_t* rec_off;
    uint64_t n_rec;
    uint64_t n_rec_file1;
} ska_input_t;

/* VariantDist (merge_ska_array.rs:44-53) */
typedef struct {
    double distance;
    double mismatch_prop;
    uint64_t match_count;
    uint64_t mismatch_count;
} ska_dist_t;

enum { SKA_QUAL_NONE = 0, SKA_QUAL_MIDDLE = 1, SKA_QUAL_STRICT = 2 };            /* cli.rs:112-123 */
enum { SKA_FILTER_NONE = 0, SKA_FILTER_NO_CONST = 1, SKA_FILTER_NO_AMBIG = 2,    /* cli.rs:129-138 */
       SKA_FILTER_NO_AMBIG_OR_CONST = 3 };

const char* skagpu_last_error(void);
/* Selects the CUDA device used by this process (one process per GPU). Default device 0. */
int skagpu_set_device(int device);

/* build_and_merge::<IntT>(&[InputFastx], k, rc, &QualOpts, threads, proportion_reads)
 * (merge_ska_dict.rs:354-417; callers lib.rs:596-603, io_utils.rs:83-90) followed by
 * MergeSkaArray::new(&dict) (merge_ska_array.rs:166-186; caller generic_modes.rs:279).
 * proportion_reads <= 0 means None.  k odd in 5..=63 else error "Invalid k-mer length"
 * (ska_dict.rs:342-344); a sample with no valid split k-mer is an error (ska_dict.rs:374-376). */
int skagpu_build(const ska_input_t* inputs, size_t n_samples, uint32_t k, int rc, uint16_t min_count,
                 uint8_t min_qual, int qual_filter, double proportion_reads, ska_array_t** out);

/* Same computation with the sequence bytes already resident in HBM: upload once, build many times
 * (what bench.py's device-resident `value` times).  The batch is not consumed. */
int skagpu_seqbatch_upload(const ska_input_t* inputs, size_t n_samples, double proportion_reads,
                           ska_seqbatch_t** out);
int skagpu_build_from_batch(const ska_seqbatch_t* batch, uint32_t k, int rc, uint16_t min_count,
                            uint8_t min_qual, int qual_filter, ska_array_t** out);
void skagpu_seqbatch_free(ska_seqbatch_t* batch);

/* The same build from HOST-PACKED assemblies: what crosses PCIe is 0.375 B per base instead of 1.  The FASTA
 * reader (needletail's place, ska_dict.rs:131-147) packs every record as it parses it:
 *   packed2: the 2-bit code (b >> 1) & 3 of encode_base (bit_encoding.rs:34-36), base i of the sample at bits 2(i%4) of byte i/4;
 *   invalid: 1 bit per base, base i at bit i%8 of byte i/8, set iff (b & 0xF) == 14 -- !valid_base (bit_encoding.rs:52-54);
 *            the unused high bits of the last byte MUST be set.
 * Records of a sample lie back to back in base space (rec_off in bases, rec_off[0] == 0).  FASTA input only (no qualities,
 * no --proportion-reads), k <= 31.  skagpu_pack_bases is the host helper that fills both streams for bases
 * [first, first + n) of a sample from ASCII (buffers preset: packed2 anything, invalid 0xFF). */
typedef struct {
    const uint8_t* packed2;
    const uint8_t* invalid;
    const uint64_t* rec_off;
    uint64_t n_rec;
    uint64_t n_rec_file1;
} ska_input_packed_t;
void skagpu_pack_bases(const uint8_t* ascii, uint64_t n, uint64_t first, uint8_t* packed2, uint8_t* invalid);
int skagpu_build_packed(const ska_input_packed_t* inputs, size_t n_samples, uint32_t k, int rc, ska_array_t** out);
int skagpu_seqbatch_upload_packed(const ska_input_packed_t* inputs, size_t n_samples, ska_seqbatch_t** out);
/* Pipelines of builds: the copies are queued on the library's copy stream and the call returns at once, so the batch
 * crosses PCIe under the kernels of builds that are already queued; skagpu_build_from_batch on it waits for the bytes.
 * The host buffers must stay valid (and unchanged) until that build has returned. */
int skagpu_seqbatch_upload_packed_async(const ska_input_packed_t* inputs, size_t n_samples, ska_seqbatch_t** out);

/* MergeSkaArray::<IntT>::load (merge_ska_array.rs:199-204) once the host has decoded the .skf:
 * keys are little-endian u64 (key_bits 64) or (lo,hi) u64 pairs (key_bits 128). n_keys may exceed
 * n_rows only for arrays exported after a filter with update_kmers=false. */
int skagpu_array_from_host(uint32_t k, int rc, uint32_t key_bits, uint64_t n_rows, uint64_t n_samples,
                           const void* keys, const uint8_t* variants_rowmajor, const uint64_t* counts,
                           ska_array_t** out);

int skagpu_dims(const ska_array_t* a, uint64_t* n_rows, uint64_t* n_keys, uint64_t* n_samples, uint32_t* k,
                int* rc, uint32_t* key_bits);

/* MergeSkaArray::filter(&mut self, min_count, filter_ambig_as_missing, &FilterType, mask_ambig,
 * ignore_const_gaps, update_kmers) -> i32   (merge_ska_array.rs:289-402; caller
 * generic_modes.rs:123-130).  *removed = number of rows removed by the filter loop. */
int skagpu_filter(ska_array_t* a, uint64_t min_count, int filter_ambig_as_missing, int filter_type, int mask_ambig,
                  int ignore_const_gaps, int update_kmers, int32_t* removed);

/* MergeSkaArray::distance(&self, constant, filt_ambig) (merge_ska_array.rs:416-438, variant_dist
 * :587-632; caller generic_modes.rs:171).  out has N(N-1)/2 entries, pairs i<j in row-major order. */
int skagpu_distance(const ska_array_t* a, double constant, int filt_ambig, ska_dist_t* out);

/* Row-major export for save()/Display (merge_ska_array.rs:191-197, 649-698): keys (n_keys entries,
 * 8 or 16 bytes each), variants (n_rows x n_samples), counts (n_rows). Any pointer may be NULL. */
int skagpu_export(const ska_array_t* a, void* keys, uint8_t* variants_rowmajor, uint64_t* counts);
/* write_fasta's transpose (merge_ska_array.rs:499-503): (n_samples x n_rows) row-major.
 * samples_by_rows == NULL: the transpose is computed on the device and dropped (device-only timing). */
int skagpu_export_transposed(const ska_array_t* a, uint8_t* samples_by_rows);

/* Pipelined form for a stream of builds (one batch after another, e.g. `ska align` over many batches): the transpose is
 * queued on the library stream and the copy to `samples_by_rows` (pinned host memory, skagpu_host_alloc) on a second
 * stream behind it; the call returns at once, so the next skagpu_build* overlaps its upload and kernels with this copy.
 * skagpu_export_wait blocks until every pending export has reached host memory.  The array may be freed right after
 * the call; the host buffer must stay valid until skagpu_export_wait returns. */
int skagpu_export_transposed_async(const ska_array_t* a, uint8_t* samples_by_rows);
int skagpu_export_wait(void);

void skagpu_free(ska_array_t* a);

/* ---- row-sharded build across GPUs (one process per GPU) -------------------
 * The reference is single-node shared-memory (README.md:124); these entry points split
 * build_and_merge (merge_ska_dict.rs:354-417) at its one exchange point so that the caller -- one
 * process per GPU, torch.distributed/NCCL for the transport -- can co-locate identical split k-mers:
 *   1. skagpu_extract_records: K1 over this rank's samples (global sample ids = sample_id_base + i), then the
 *      FIRST of the two partition passes: records are packed to one u64 each (sample id inside) and grouped by
 *      the top `region_bits` bits of the key ("regions"; region_bits <= prefix_bits is chosen by the library and
 *      reported by skagpu_records_info).  Region offsets tell how many records fall in every region; summed
 *      over ranks they give balanced contiguous key-range owners.  prefix_bits must come from
 *      skagpu_merge_prefix_bits with the GLOBAL record count so that all ranks use the same grid;
 *   2. the caller exchanges, per owner, the contiguous record range of its regions (one all-to-all);
 *   3. skagpu_merge_records_dev: the receive buffer holds `n_runs` runs (one per source rank, in rank order),
 *      each grouped by region: run_offsets[w * (nr + 1) + j] is the offset inside the buffer of region
 *      (region_lo + j) of run w, nr = region_hi - region_lo.  The SECOND partition pass runs here over all runs
 *      at once (it also concatenates them), then the merge emits this rank's slab of rows with all
 *      n_samples_total columns.  The receive buffer is overwritten (scratch).
 * Slabs are key ranges in rank order, so their concatenation is the canonical (ascending key) array.
 * Record buffers stay valid until the next extract/build/merge call on this process. */
int skagpu_merge_prefix_bits(uint64_t n_records_total, uint32_t k, uint64_t n_samples_total);   /* < 0 on error */
int skagpu_extract_records(const ska_seqbatch_t* batch, uint32_t k, int rc, uint32_t sample_id_base, uint64_t n_samples_total,
                           int prefix_bits, ska_records_t** out);
int skagpu_records_info(const ska_records_t* r, uint64_t* n, const void** rec_dev, int* region_bits,
                        uint64_t* region_offsets /* 2^region_bits + 1 entries, may be NULL */);
void skagpu_records_free(ska_records_t* r);
int skagpu_merge_records_dev(void* rec_dev, int n_runs, const uint64_t* run_offsets, uint32_t region_lo, uint32_t region_hi,
                             int prefix_bits, uint64_t n_samples_total, uint32_t k, int rc, ska_array_t** out);
/* ---- row-sharded build, second form: exchange of PARTIAL ROWS (the default of ska.rust_b200/distributed.py) ------
 * MergeSkaDict::merge (merge_ska_dict.rs:119-151) joins two partial dictionaries over disjoint samples on the split
 * k-mer; build_and_merge's join tree (merge_ska_dict.rs:264-326) does that between threads.  Between GPUs:
 *   1. every rank builds the array of ITS samples with skagpu_build / skagpu_build_from_batch (all its rows, its columns);
 *   2. skagpu_key_cumulative: cum[b] = number of this array's rows whose key is below b << (2(k-1) - region_bits),
 *      b = 0 .. 2^region_bits (cum[2^region_bits] = n_rows).  Summed over ranks this gives balanced contiguous
 *      key-range owners, and cum itself is where this rank's rows split between them;
 *   3. the caller sends each owner its row range: keys (u64), variants (rows x local columns), counts (u32), read in
 *      place through skagpu_array_dev (device pointers, valid until the array is filtered or freed);
 *   4. skagpu_merge_partial_rows_dev joins the `n_runs` received runs (run w: run_rows[w] rows of run_cols[w] columns
 *      that go to columns [run_col_base[w], +run_cols[w]) of the slab; the three device buffers hold the runs back to
 *      back; each run ascending in key).  Cells of samples that do not have a k-mer stay '-' (merge_ska_array.rs:175),
 *      variant_count adds up.  The result is this rank's slab, identical to what skagpu_merge_records_dev gives.
 * Arrays must be in ascending key order (everything this library builds is). */
int skagpu_array_dev(const ska_array_t* a, const void** keys_dev, const void** variants_dev, const void** counts_dev);
int skagpu_key_cumulative(const ska_array_t* a, int region_bits, uint64_t* cum /* 2^region_bits + 1 entries */);
int skagpu_merge_partial_rows_dev(const void* keys_dev, const void* variants_dev, const void* counts_dev, int n_runs,
                                  const uint64_t* run_rows, const uint32_t* run_cols, const uint32_t* run_col_base,
                                  uint64_t n_samples_total, uint32_t k, int rc, ska_array_t** out);

/* ---- the same exchange with the library holding the communicator (csrc/comm.cu): no torch, no host transport ---------
 * For hosts that drive several GPUs themselves -- the C++ `ska --gpus N` (one thread per GPU) or a Rust binding with one
 * process per GPU.  A host thread is bound to one device by skagpu_set_device; every call it makes afterwards (builds,
 * filters, exports, these collectives) runs on that device, so N threads of one process drive N GPUs.
 *   skagpu_comm_unique_id   rank 0 draws the NCCL id (SKAGPU_COMM_ID_BYTES bytes) and hands it to the other ranks by
 *                           whatever the host has (shared memory between threads, MPI, a file, torchrun's store);
 *   skagpu_comm_init        every rank, after skagpu_set_device.  NCCL is dlopen'ed here (libnccl.so.2);
 *   skagpu_exchange_rows    collective.  `local` = the array of this rank's samples (columns [sample_base, sample_base +
 *                           its samples) of the whole job), built by any of the build entry points.  Steps 2-4 above inside
 *                           the library: all-gather of the key histograms, owners cut for equal row counts, ONE grouped
 *                           send/recv of every peer's keys + variant rows + counts over NVLink, join on the key.
 *                           -> *slab: this rank's ascending key range with all n_samples_total columns; the slabs of ranks
 *                           0 .. world-1 concatenated are the single-GPU array (MergeSkaDict::merge, merge_ska_dict.rs:119-151).
 *                           stats (may be NULL): rows sent, rows received, bytes sent, rows of the slab;
 *   skagpu_allreduce_u64    sums host integers over the ranks (the distance counters and the constant-site count of
 *                           generic_modes.rs:136-189 on slabs);
 *   skagpu_distance_sharded collective.  skagpu_distance on a slab: the per-pair integer counters and `constant_local` (the
 *                           constant sites the no-const filter removed from THIS slab) are summed over the ranks, every rank
 *                           gets the whole job's table (generic_modes.rs:136-189, merge_ska_array.rs:416-438);
 *   skagpu_plan_owners      the owner cut on its own (host arithmetic, no device): cuts[world + 1] region indices. */
#define SKAGPU_COMM_ID_BYTES 128
int skagpu_comm_unique_id(uint8_t* id);
int skagpu_comm_init(const uint8_t* id, int world, int rank);
int skagpu_comm_destroy(void);
int skagpu_exchange_rows(const ska_array_t* local, uint64_t sample_base, uint64_t n_samples_total, ska_array_t** slab, uint64_t* stats);
int skagpu_allreduce_u64(uint64_t* values, size_t n);
int skagpu_distance_sharded(const ska_array_t* slab, double constant_local, int filt_ambig, ska_dist_t* out);
int skagpu_plan_owners(const uint64_t* region_rows, uint32_t n_regions, int world, uint64_t* cuts);

/* Per-pair partial sums of variant_dist (merge_ska_array.rs:599-620) over this array's rows, 4 x u64 per
 * pair {mismatches, matches, dist_int, n_both}: integers, so row shards add up exactly (all-reduce);
 * distance = dist_int (filt_ambig) or n_both - dist_int/36 (allow-ambiguous), matches += constant. */
int skagpu_distance_counters(const ska_array_t* a, int filt_ambig, uint64_t* counters);

/* ---- the other consumers of a built array: joins on the split-k-mer key ----------------
 * Arrays are taken in ascending key order; one loaded from a reference-written .skf (hashbrown order) is sorted in place on
 * first use -- row order is not part of an array's value (the reference's own tests compare as sets).
 *
 * skagpu_merge_arrays    MergeSkaDict::extend (merge_ska_dict.rs:160-193; `ska merge`, generic_modes.rs:90-106): the samples of b
 *                        are appended to those of a, rows joined on the key, cells a sample lacks are '-'.  Errors like the
 *                        reference's panics: "K-mer lengths do not match: {b.k} {a.k}", "Strand use inconsistent".
 * skagpu_delete_samples  MergeSkaArray::delete_samples (merge_ska_array.rs:231-271; `ska delete`): columns sample_idx[] are
 *                        dropped, counts recomputed, rows left without any base removed.  Name lookup stays with the host.
 * skagpu_weed            RefSka::new(k, weed_fasta, rc, false, false) + MergeSkaArray::weed (ska_ref.rs:189-311,
 *                        merge_ska_array.rs:452-487; `ska weed`, generic_modes.rs:218-266): rows whose split k-mer occurs in the
 *                        weed sequences are removed (reverse: only those are kept).  weed_fasta = ONE parsed FASTA (all records).
 * skagpu_map             RefSka::new + RefSka::map (ska_ref.rs:189-311, 508-533; `ska map`): every split k-mer of the reference,
 *                        in reference order, is looked up in the array; a hit contributes its row, complemented with RC_IUPAC
 *                        (bit_encoding.rs:474-507) when the reference k-mer is on the reverse strand.  skagpu_map_export gives the
 *                        hits' middle-base positions (coordinates of the reference's records laid back to back), their rows
 *                        (n_mapped x n_samples) and, if asked, the K1 record of every reference position (key << 4 | one-hot
 *                        middle, all-ones = no split k-mer starts here; two words per position for k > 31) from which the host
 *                        derives the repeat mask (ska_ref.rs:256-296).  The pseudo-alignment / VCF writers are host code. */
int skagpu_merge_arrays(ska_array_t* a, ska_array_t* b, ska_array_t** out);
int skagpu_delete_samples(ska_array_t* a, const uint32_t* sample_idx, size_t n_delete);
int skagpu_weed(ska_array_t* a, const ska_input_t* weed_fasta, int reverse, uint64_t* n_weed_kmers);
int skagpu_map(ska_array_t* a, const ska_input_t* ref_fasta, ska_mapping_t** out);
int skagpu_map_dims(const ska_mapping_t* m, uint64_t* n_mapped, uint64_t* n_ref_kmers, uint64_t* n_samples, uint64_t* n_positions);
int skagpu_map_export(const ska_mapping_t* m, uint64_t* middle_positions, uint8_t* variants_rowmajor, uint64_t* ref_records);
void skagpu_map_free(ska_mapping_t* m);

/* The counting half of `ska cov` and of `ska build --min-count auto` (CoverageHistogram::new + the histogram loop of
 * fit_histogram, src/coverage.rs:70-135, 156-163): every window of every read of the sample (both files, qualities ignored)
 * counts towards its canonical split k-mer; hist[c - 1] = number of distinct split k-mers seen exactly c times, c = 1..1000.
 * The Poisson-mixture fit on the histogram (coverage.rs:165-213) is host code. */
int skagpu_kmer_count_hist(const ska_input_t* reads, uint32_t k, int rc, uint64_t* hist /* 1000 bins */);

/* Self-check of a device-resident array (no reference counterpart; bench.py's in-run parity check and the tests):
 * out[0] = order-independent checksum over the non-'-' cells (key, column - col_lo, letter) of columns [col_lo, col_hi)
 *          -- MergeSkaDict::merge (merge_ska_dict.rs:119-151) only moves cells, so a rank's local build and the same
 *          columns of the joined slabs, summed over ranks (mod 2^64), must agree;
 * out[1] = rows with at least one such cell; out[2] = rows whose variant_count is not their number of non-'-' cells
 * (merge_ska_array.rs:172); out[3] = adjacent keys that are not strictly ascending. */
int skagpu_array_checksum(const ska_array_t* a, uint32_t col_lo, uint32_t col_hi, uint64_t* out /* 4 */);

/* Pinned host memory for callers that want full-speed PCIe transfers. */
void* skagpu_host_alloc(size_t bytes);
void skagpu_host_free(void* p);

/* ---- instrumentation (no reference counterpart) ------------------------------------------------
 * Stage timings of the most recent build on this thread, measured with CUDA events on the
 * library's own stream.  names[i] is a static string; returns the number of stages written. */
int skagpu_last_timings(const char** names, float* ms, uint64_t* launches, int max_stages);
/* CUDA-event stopwatch on the library's stream: op 0 = synchronise + start, op 1 = stop, *ms = elapsed
 * device time between the two (what bench.py brackets its timed region with). */
int skagpu_stream_timer(int op, float* ms);
/* Number of kernel launches issued by this library since process start (all entry points). */
uint64_t skagpu_launch_count(void);
/* Raw K1 access for kernel-level parity tests and the roofline measurement: runs the extraction
 * kernel `iters` times over a resident batch; writes one u64 record per base position
 * (key<<4 | one-hot(middle), 0xFFFF...F for positions that start no valid window) to rec_out
 * (host, n_positions entries, may be NULL) and reports the mean kernel time. */
int skagpu_k1_extract(const ska_seqbatch_t* batch, uint32_t k, int rc, uint8_t min_qual, int qual_filter, int iters,
                      uint64_t* rec_out, uint64_t* n_positions, float* mean_ms);

#ifdef __cplusplus
}
#endif
#endif /* SKAGPU_H */
