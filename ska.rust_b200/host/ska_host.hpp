// ska_host.hpp -- C++ host-side mirror of the reference's interface for the hot path, on top of the
// C ABI (include/skagpu.h).  Same names and argument meaning as the Rust API it stands in for:
//   MergeSkaArray::{load, save, filter, distance, write_fasta, Display/Debug}   src/merge_ska_array.rs
//   build_and_merge                                                             src/merge_ska_dict.rs:354-417
//   generic_modes::{align, distance, apply_filters, save_skf}                   src/generic_modes.rs
//   io_utils::{read_input_fastas, get_input_list, load_array}                   src/io_utils.rs
// Errors are C++ exceptions carrying the reference's panic text; the CLI turns them into exit code 101.
#pragma once
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/skagpu.h"
#include "skf_codec.hpp"

namespace ska {

struct SkaError : std::runtime_error { using std::runtime_error::runtime_error; };

enum class QualFilter { NoFilter = 0, Middle = 1, Strict = 2 };                      // cli.rs:112-123
enum class FilterType { NoFilter = 0, NoConst = 1, NoAmbig = 2, NoAmbigOrConst = 3 };  // cli.rs:129-138
struct QualOpts { uint16_t min_count = 5; uint8_t min_qual = 20; QualFilter qual_filter = QualFilter::Strict; };  // lib.rs:533-540
typedef std::tuple<std::string, std::string, std::string> InputFastx;                // (name, file1, file2 or "")

class MergeSkaArray {
  public:
    MergeSkaArray() {}
    MergeSkaArray(ska_array_t* h, std::vector<std::string> names, std::string version = "0.5.2")
        : h_(h), names_(std::move(names)), version_(std::move(version)) {}
    MergeSkaArray(MergeSkaArray&& o) noexcept { *this = std::move(o); }
    MergeSkaArray& operator=(MergeSkaArray&& o) noexcept {
        if (this != &o) { reset(); h_ = o.h_; o.h_ = nullptr; names_ = std::move(o.names_); version_ = std::move(o.version_); k_bits_field_ = o.k_bits_field_; }
        return *this;
    }
    MergeSkaArray(const MergeSkaArray&) = delete;
    MergeSkaArray& operator=(const MergeSkaArray&) = delete;
    ~MergeSkaArray() { reset(); }

    static MergeSkaArray load(const std::string& filename);                  // merge_ska_array.rs:199-204
    void save(const std::string& filename, int threads = 1) const;           // merge_ska_array.rs:191-197
    int32_t filter(size_t min_count, bool filter_ambig_as_missing, FilterType filter, bool mask_ambig,
                   bool ignore_const_gaps, bool update_kmers);               // merge_ska_array.rs:289-402
    std::vector<ska_dist_t> distance(double constant, bool filt_ambig) const;  // merge_ska_array.rs:416-438
    void write_fasta(FILE* f) const;                                         // merge_ska_array.rs:499-517
    std::string display(bool full_info) const;                               // merge_ska_array.rs:649-698

    // ---- joins on the key: ska merge / delete / weed
    static MergeSkaArray extend(MergeSkaArray& a, MergeSkaArray& b);         // MergeSkaDict::extend, merge_ska_dict.rs:160-193
    void delete_samples(const std::vector<std::string>& del_names);          // merge_ska_array.rs:231-271
    uint64_t weed(const std::string& weed_fasta, bool reverse);              // RefSka::new + weed, merge_ska_array.rs:452-487
    ska_array_t* handle() const { return h_; }

    size_t nsamples() const;
    size_t ksize() const;
    size_t kmer_len() const;
    bool rc() const;
    uint32_t k_bits() const;
    const std::vector<std::string>& names() const { return names_; }
    skf::Array to_host() const;

  private:
    void reset() { if (h_) skagpu_free(h_); h_ = nullptr; }
    ska_array_t* h_ = nullptr;
    std::vector<std::string> names_;
    std::string version_;
    uint32_t k_bits_field_ = 0;      // what a loaded file said (informational; see skf::Array::k_bits_field)
};

// GPUs that build_and_merge shards the samples over (default 1; `ska --gpus N` / SKA_GPUS): see build_sharded in ska_host.cpp
void set_gpus(int n);
int gpus();

MergeSkaArray build_and_merge(const std::vector<InputFastx>& input_files, size_t k, bool rc, const QualOpts& qual,
                              size_t threads, double proportion_reads /* <= 0: None */);

// RefSka (src/ska_ref.rs): a reference FASTA, its split k-mers, and after map() the middle bases of every sample at the
// reference positions where a split k-mer of the array occurs.  The lookup runs on the device (skagpu_map); the
// pseudo-alignment writer (src/ska_ref/aln_writer.rs) and the VCF writer (ska_ref.rs:636-752) are host code, sample-parallel.
class RefSka {
  public:
    RefSka(size_t k, const std::string& filename, bool rc, bool ambig_mask, bool repeat_mask);      // ska_ref.rs:189-311 (file part)
    void map(MergeSkaArray& ska_array);                                                              // ska_ref.rs:508-533
    void write_aln(FILE* f, size_t threads);                                                         // ska_ref.rs:636-660
    void write_vcf(FILE* f, size_t threads);                                                         // ska_ref.rs:662-752
    size_t ksize() const { return n_ref_kmers_; }

  private:
    std::vector<std::vector<uint8_t>> pseudoalignment(size_t threads) const;                        // ska_ref.rs:555-587
    size_t k_; bool rc_, ambig_mask_, repeat_mask_;
    std::vector<std::string> chrom_names_;
    std::vector<std::vector<uint8_t>> seq_;
    std::vector<uint8_t> flat_;                  // records back to back (what goes to the device)
    std::vector<uint64_t> rec_off_;
    std::vector<size_t> repeat_coors_;
    size_t n_ref_kmers_ = 0;
    std::vector<std::pair<size_t, size_t>> mapped_pos_;   // (chrom, pos)
    std::vector<uint8_t> mapped_variants_;                // rows loci, columns samples
    std::vector<std::string> mapped_names_;
};

// CoverageHistogram (src/coverage.rs): split k-mer multiplicities of one read set and the two-component Poisson mixture
// fitted to them.  Counting runs on the device (skagpu_kmer_count_hist); the fit is the reference's: BFGS with a
// backtracking Armijo line search (argmin's solvers restated), inverse Hessian started at identity, cost tolerance 1e-6,
// at most 20 iterations (coverage.rs:165-213).
class CoverageHistogram {
  public:
    CoverageHistogram(const std::string& fastq1, const std::string& fastq2, size_t k, bool rc, bool verbose);   // coverage.rs:70-135
    explicit CoverageHistogram(std::vector<uint32_t> counts) : counts_(std::move(counts)), from_counts_(true) {}  // the unit test's entry
    size_t fit_histogram();                                                                                      // coverage.rs:151-213
    std::string plot_hist() const;                                                                               // coverage.rs:220-243
    double w0() const { return w0_; }
    double c() const { return c_; }

  private:
    std::vector<uint32_t> counts_;
    std::vector<uint64_t> hist_;
    bool from_counts_ = false, fitted_ = false, verbose_ = false;
    double w0_ = 0.8, c_ = 20.0;         // INIT_W0, INIT_C (coverage.rs:32-33)
    size_t cutoff_ = 0;
};

namespace io_utils {
std::vector<InputFastx> read_input_fastas(const std::vector<std::string>& seq_files);               // io_utils.rs:31-46
std::vector<InputFastx> get_input_list(const std::string& file_list, const std::vector<std::string>& seq_files);  // :116-146
MergeSkaArray load_array(const std::vector<std::string>& input, size_t threads);                    // io_utils.rs:60-93
// --min-count: a number, "auto" (coverage model on the first files of the first two paired samples), or absent (5)   io_utils.rs:175-212
uint16_t kmer_min_cutoff(const std::string& min_count_arg, const std::vector<InputFastx>& files, size_t k, bool rc, bool verbose);
}  // namespace io_utils

namespace generic_modes {
int32_t apply_filters(MergeSkaArray& a, double min_freq, bool filter_ambig_as_missing, FilterType filter, bool ambig_mask,
                      bool ignore_const_gaps);                                                       // generic_modes.rs:112-131
void align(MergeSkaArray& a, const std::string& output /* "" = stdout */, FilterType filter, bool mask_ambig,
           bool ignore_const_gaps, double min_freq, bool filter_ambig_as_missing);                   // generic_modes.rs:22-50
void distance(MergeSkaArray& a, const std::string& output, double min_freq, bool filt_ambig);        // generic_modes.rs:136-189
void save_skf(const MergeSkaArray& a, const std::string& output, int threads);                       // generic_modes.rs:270-283
void map(MergeSkaArray& a, RefSka& ska_ref, const std::string& output, bool vcf, size_t threads);    // generic_modes.rs:53-84
void merge(MergeSkaArray& first, const std::vector<std::string>& skf_files, const std::string& output, int threads);   // :90-106
void delete_(MergeSkaArray& a, const std::vector<std::string>& names, const std::string& out_file, int threads);      // :191-213
void weed(MergeSkaArray& a, const std::string& weed_file, bool reverse, double min_freq, bool filter_ambig_as_missing, FilterType filter,
          bool ambig_mask, bool ignore_const_gaps, const std::string& out_file, int threads);       // :215-266
}  // namespace generic_modes

}  // namespace ska
