#include "ska_host.hpp"

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <fstream>
#include <regex>
#include <sstream>
#include <thread>

#include "fastx.hpp"

namespace ska {

static void check(int rc) { if (rc) throw SkaError(skagpu_last_error()); }

// SKA_TIMING=1: phase times of the host side on stderr (the file-to-file benchmark reads them)
static void lap(const char* what) {
    static const bool on = getenv("SKA_TIMING") != nullptr;
    static const auto t0 = std::chrono::steady_clock::now();
    static auto last = t0;
    if (!on) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[ska] %-28s +%.3f s  (%.3f s)\n", what, std::chrono::duration<double>(now - last).count(),
            std::chrono::duration<double>(now - t0).count());
    last = now;
}

// ------------------------------------------------------------------------------------------ MergeSkaArray
static void dims(ska_array_t* h, uint64_t& rows, uint64_t& keys, uint64_t& samples, uint32_t& k, int& rc, uint32_t& bits) {
    check(skagpu_dims(h, &rows, &keys, &samples, &k, &rc, &bits));
}
size_t MergeSkaArray::nsamples() const { uint64_t r, kk, s; uint32_t k, b; int rc; dims(h_, r, kk, s, k, rc, b); return s; }
size_t MergeSkaArray::ksize() const { uint64_t r, kk, s; uint32_t k, b; int rc; dims(h_, r, kk, s, k, rc, b); return r; }
size_t MergeSkaArray::kmer_len() const { uint64_t r, kk, s; uint32_t k, b; int rc; dims(h_, r, kk, s, k, rc, b); return k; }
bool MergeSkaArray::rc() const { uint64_t r, kk, s; uint32_t k, b; int rc; dims(h_, r, kk, s, k, rc, b); return rc != 0; }
uint32_t MergeSkaArray::k_bits() const { uint64_t r, kk, s; uint32_t k, b; int rc; dims(h_, r, kk, s, k, rc, b); return b; }

MergeSkaArray MergeSkaArray::load(const std::string& filename) {
    lap("start");
    skf::Array a = skf::load(filename);
    lap("skf read + decoded");
    check(skagpu_set_device(0));
    lap("cuda context ready");
    ska_array_t* h = nullptr;
    check(skagpu_array_from_host(a.k, a.rc, a.key_bits, a.n_rows, a.n_samples, a.keys.data(), a.variants.data(), a.counts.data(), &h));
    lap("array on the device");
    return MergeSkaArray(h, a.names, a.ska_version);
}

skf::Array MergeSkaArray::to_host() const {
    skf::Array a;
    uint64_t rows, keys, samples; uint32_t k, bits; int rc;
    dims(h_, rows, keys, samples, k, rc, bits);
    a.k = k; a.rc = rc != 0; a.key_bits = bits; a.n_rows = rows; a.n_samples = samples; a.names = names_; a.ska_version = version_;
    a.keys.resize(keys * (bits / 64)); a.variants.resize(rows * samples); a.counts.resize(rows);
    skf::prefault(a.variants.data(), a.variants.size(), (int)std::thread::hardware_concurrency());
    check(skagpu_export(h_, a.keys.data(), a.variants.data(), a.counts.data()));
    return a;
}

void MergeSkaArray::save(const std::string& filename, int threads) const {
    skf::Array a = to_host();
    lap("array exported to the host");
    skf::save(a, filename, threads);
    lap("skf encoded + written");
}

int32_t MergeSkaArray::filter(size_t min_count, bool filter_ambig_as_missing, FilterType filter, bool mask_ambig,
                              bool ignore_const_gaps, bool update_kmers) {
    if (ignore_const_gaps && (filter == FilterType::NoAmbig || filter == FilterType::NoFilter))
        fprintf(stderr, "WARN --no-gap-only-sites can only be applied when filtering constant bases\n");   // merge_ska_array.rs:298-300
    int32_t removed = 0;
    check(skagpu_filter(h_, min_count, filter_ambig_as_missing, (int)filter, mask_ambig, ignore_const_gaps, update_kmers, &removed));
    return removed;
}

std::vector<ska_dist_t> MergeSkaArray::distance(double constant, bool filt_ambig) const {
    size_t n = nsamples();
    std::vector<ska_dist_t> out(n * (n - 1) / 2);
    ska_dist_t dummy;
    check(skagpu_distance(h_, constant, filt_ambig, out.empty() ? &dummy : out.data()));
    return out;
}

void MergeSkaArray::write_fasta(FILE* f) const {
    size_t n = nsamples(), r = ksize();
    skf::Bytes t;
    t.resize(n * r);
    skf::prefault(t.data(), t.size(), (int)std::thread::hardware_concurrency());
    check(skagpu_export_transposed(h_, t.empty() ? nullptr : t.data()));
    lap("filtered + transposed + copied");
    for (size_t i = 0; i < n; i++) {                       // needletail write_fasta, Unix line ending, no wrapping
        fputc('>', f); fwrite(names_[i].data(), 1, names_[i].size(), f); fputc('\n', f);
        fwrite(t.data() + i * r, 1, r, f); fputc('\n', f);
    }
}

std::string MergeSkaArray::display(bool full_info) const {
    skf::Array a = to_host();
    std::ostringstream o;
    o << "ska_version=" << a.ska_version << "\nk=" << a.k << "\nk_bits=" << a.key_bits << "\nrc=" << (a.rc ? "true" : "false")
      << "\nk-mers=" << a.n_rows << "\nsamples=" << a.n_samples << "\n";
    o << "sample_names=[";
    for (size_t i = 0; i < a.names.size(); i++) o << (i ? ", " : "") << '"' << a.names[i] << '"';   // Rust {:?} of Vec<String>
    o << "]\nsample_kmers=[";
    for (size_t s = 0; s < a.n_samples; s++) {             // merge_ska_array.rs:554-559
        uint64_t c = 0;
        for (size_t r = 0; r < a.n_rows; r++) c += a.variants[r * a.n_samples + s] != '-';
        o << (s ? ", " : "") << c;
    }
    o << "]\n\n";                                          // println!("{ska_array}")  (lib.rs:813)
    if (full_info) {                                       // println!("{ska_array:?}") (lib.rs:816)
        const int h = (a.k - 1) / 2;
        const char* L = "ACTG";
        const int kw = a.key_bits / 64;
        for (size_t r = 0; r < a.n_rows && r < a.n_keys(); r++) {
            unsigned __int128 key = a.keys[r * kw];
            if (kw == 2) key |= (unsigned __int128)a.keys[r * kw + 1] << 64;
            std::string up, lo;
            for (int i = 0; i < h; i++) up += L[(int)((key >> (2 * h + 2 * (h - 1 - i))) & 3)];
            for (int i = 0; i < h; i++) lo += L[(int)((key >> (2 * (h - 1 - i))) & 3)];
            o << up << '\t' << lo << '\t';
            for (size_t s = 0; s < a.n_samples; s++) o << (s ? "," : "") << (char)a.variants[r * a.n_samples + s];
            o << '\n';
        }
        o << '\n';
    }
    return o.str();
}

// ------------------------------------------------------------------------------------------ build_and_merge
MergeSkaArray build_and_merge(const std::vector<InputFastx>& input_files, size_t k, bool rc, const QualOpts& qual,
                              size_t threads, double proportion_reads) {
    if (k < 5 || k > 63 || k % 2 == 0) throw SkaError("Invalid k-mer length");                    // ska_dict.rs:342-344
    const size_t n = input_files.size();
    std::vector<fastx::Records> recs(n);
    std::vector<uint64_t> n_file1(n);
    std::vector<std::string> errors(n);
    std::atomic<size_t> next{0};
    // assemblies at k <= 31 take the packed path: every parsing thread also packs its file to 2-bit codes + invalid mask
    // (skagpu_pack_bases), so 0.375 B per base cross PCIe instead of 1 and the ASCII copy is dropped at once
    const bool try_packed = k <= 31 && !(proportion_reads > 0) && !getenv("SKA_HOST_ASCII");
    std::vector<std::vector<uint8_t>> packed2(n), invalid(n);
    auto work = [&] {                                      // host parsing is per-file independent (needletail's job)
        for (size_t i; (i = next.fetch_add(1)) < n;) {
            try {
                recs[i] = fastx::parse_file(std::get<1>(input_files[i]));
                n_file1[i] = recs[i].n_rec();
                if (!std::get<2>(input_files[i]).empty()) fastx::append(recs[i], fastx::parse_file(std::get<2>(input_files[i])));
                if (try_packed && !recs[i].is_fastq) {
                    const uint64_t nb = recs[i].seq.size();
                    packed2[i].assign((nb + 3) / 4, 0);
                    invalid[i].assign((nb + 7) / 8, 0xFF);
                    skagpu_pack_bases(recs[i].seq.data(), nb, 0, packed2[i].data(), invalid[i].data());
                    std::vector<uint8_t>().swap(recs[i].seq);
                }
            } catch (std::exception& e) { errors[i] = e.what(); }
        }
    };
    size_t nt = std::max<size_t>(1, std::min<size_t>(threads, n));
    lap("start");
    std::vector<std::thread> th;
    for (size_t t = 1; t < nt; t++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    for (auto& e : errors) if (!e.empty()) throw SkaError(e);
    lap("files parsed (+ packed)");
    check(skagpu_set_device(0));                           // the context is created on a helper thread since program start (main.cpp)
    lap("cuda context ready");

    std::vector<std::string> names(n);
    for (size_t i = 0; i < n; i++) names[i] = std::get<0>(input_files[i]);
    bool all_packed = try_packed;
    for (size_t i = 0; i < n; i++) all_packed = all_packed && !recs[i].is_fastq;
    auto fail = [&](std::string msg) {
        // "<file> has no valid sequence" names the file in the reference (ska_dict.rs:375)
        size_t p = msg.find("sample ");
        if (p == 0 && msg.find("has no valid sequence") != std::string::npos) {
            size_t idx = std::stoul(msg.substr(7));
            if (idx < n) msg = std::get<1>(input_files[idx]) + " has no valid sequence";
        }
        throw SkaError(msg);
    };
    if (all_packed) {
        std::vector<ska_input_packed_t> pin(n);
        for (size_t i = 0; i < n; i++) pin[i] = {packed2[i].data(), invalid[i].data(), recs[i].rec_off.data(), recs[i].n_rec(), n_file1[i]};
        ska_array_t* h = nullptr;
        if (skagpu_build_packed(pin.data(), n, (uint32_t)k, rc, &h)) fail(skagpu_last_error());
        lap("built on the device");
        return MergeSkaArray(h, names);
    }
    if (try_packed)                                        // mixed FASTA/FASTQ input: back to bytes for the samples packed above
        for (size_t i = 0; i < n; i++)
            if (!recs[i].is_fastq) {
                const uint64_t nb = recs[i].rec_off.back();
                recs[i].seq.resize(nb);
                for (uint64_t p = 0; p < nb; p++)
                    recs[i].seq[p] = (invalid[i][p >> 3] >> (p & 7)) & 1 ? 'N' : "ACTG"[(packed2[i][p >> 2] >> (2 * (p & 3))) & 3];
            }
    std::vector<ska_input_t> in(n);
    for (size_t i = 0; i < n; i++) {
        in[i].seq = recs[i].seq.data();
        in[i].qual = recs[i].is_fastq ? recs[i].qual.data() : nullptr;
        in[i].rec_off = recs[i].rec_off.data();
        in[i].n_rec = recs[i].n_rec();
        in[i].n_rec_file1 = n_file1[i];
    }
    ska_array_t* h = nullptr;
    if (skagpu_build(in.data(), n, (uint32_t)k, rc, qual.min_count, qual.min_qual, (int)qual.qual_filter, proportion_reads, &h))
        fail(skagpu_last_error());
    return MergeSkaArray(h, names);
}

// ------------------------------------------------------------------------------------------ io_utils
namespace io_utils {
std::vector<InputFastx> read_input_fastas(const std::vector<std::string>& seq_files) {
    static const std::regex re_path(R"(^.+/(.+)\.(fa|fasta|fastq|fastq\.gz)$)", std::regex::icase);
    static const std::regex re_name(R"(^(.+)\.(fa|fasta|fastq|fastq\.gz)$)", std::regex::icase);
    std::vector<InputFastx> out;
    for (auto& f : seq_files) {
        std::smatch m;
        std::string name = f;
        if (std::regex_match(f, m, re_path) || std::regex_match(f, m, re_name)) name = m[1];
        out.emplace_back(name, f, "");
    }
    return out;
}
std::vector<InputFastx> get_input_list(const std::string& file_list, const std::vector<std::string>& seq_files) {
    if (file_list.empty()) return read_input_fastas(seq_files);
    std::ifstream f(file_list);
    if (!f) throw SkaError("Unable to open file_list");
    std::vector<InputFastx> out;
    std::string line;
    while (std::getline(f, line)) {
        std::istringstream ss(line);
        std::vector<std::string> fields;
        for (std::string w; ss >> w;) fields.push_back(w);
        if (fields.size() == 2) out.emplace_back(fields[0], fields[1], "");
        else if (fields.size() == 3) out.emplace_back(fields[0], fields[1], fields[2]);
        else throw SkaError("Unable to parse line in file_list");
    }
    return out;
}
MergeSkaArray load_array(const std::vector<std::string>& input, size_t threads) {
    if (input.size() == 1) return MergeSkaArray::load(input[0]);
    // "Multiple files as input, running ska build with default settings" (io_utils.rs:75-92)
    return build_and_merge(read_input_fastas(input), 31, true, QualOpts(), threads, 0.0);
}
}  // namespace io_utils

// ------------------------------------------------------------------------------------------ generic_modes
namespace generic_modes {
int32_t apply_filters(MergeSkaArray& a, double min_freq, bool filter_ambig_as_missing, FilterType filter, bool ambig_mask,
                      bool ignore_const_gaps) {
    size_t threshold = (size_t)std::ceil((double)a.nsamples() * min_freq);
    return a.filter(threshold, filter_ambig_as_missing, filter, ambig_mask, ignore_const_gaps, false);
}

static FILE* set_ostream(const std::string& output) {      // io_utils.rs:98-107
    if (output.empty()) return stdout;
    FILE* f = fopen(output.c_str(), "wb");
    if (!f) throw SkaError("cannot create " + output);
    return f;
}

void align(MergeSkaArray& a, const std::string& output, FilterType filter, bool mask_ambig, bool ignore_const_gaps,
           double min_freq, bool filter_ambig_as_missing) {
    apply_filters(a, min_freq, filter_ambig_as_missing, filter, mask_ambig, ignore_const_gaps);
    FILE* f = set_ostream(output);
    a.write_fasta(f);
    if (f != stdout) fclose(f); else fflush(f);
    lap("alignment written");
}

void distance(MergeSkaArray& a, const std::string& output, double min_freq, bool filt_ambig) {
    if (min_freq * (double)a.nsamples() >= 1.0) apply_filters(a, min_freq, false, FilterType::NoFilter, false, false);
    int32_t constant = apply_filters(a, 0.0, false, FilterType::NoConst, false, false);
    std::vector<ska_dist_t> d = a.distance((double)constant, filt_ambig);
    FILE* f = set_ostream(output);
    fputs("Sample1\tSample2\tDistance\tMismatches (proportion)\tMatch count\tMismatch count\n", f);
    const auto& names = a.names();
    size_t p = 0;
    for (size_t i = 0; i < names.size(); i++)
        for (size_t j = i + 1; j < names.size(); j++, p++)
            fprintf(f, "%s\t%s\t%.2f\t%.5f\t%llu\t%llu\n", names[i].c_str(), names[j].c_str(), d[p].distance, d[p].mismatch_prop,
                    (unsigned long long)d[p].match_count, (unsigned long long)d[p].mismatch_count);
    if (f != stdout) fclose(f); else fflush(f);
}

void save_skf(const MergeSkaArray& a, const std::string& output, int threads) {
    std::string path = output;
    if (path.size() < 4 || path.compare(path.size() - 4, 4, ".skf") != 0) path += ".skf";   // generic_modes.rs:270-283
    a.save(path, threads);
}
}  // namespace generic_modes

}  // namespace ska
