#include "ska_host.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <fstream>
#include <functional>
#include <limits>
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
    MergeSkaArray out(h, a.names, a.ska_version);
    out.k_bits_field_ = a.k_bits_field;
    return out;
}

skf::Array MergeSkaArray::to_host() const {
    skf::Array a;
    uint64_t rows, keys, samples; uint32_t k, bits; int rc;
    dims(h_, rows, keys, samples, k, rc, bits);
    a.k = k; a.rc = rc != 0; a.key_bits = bits; a.n_rows = rows; a.n_samples = samples; a.names = names_; a.ska_version = version_;
    a.k_bits_field = k_bits_field_ ? k_bits_field_ : (k > 31 ? 128u : 64u);
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
    o << "ska_version=" << a.ska_version << "\nk=" << a.k << "\nk_bits=" << a.k_bits_field << "\nrc=" << (a.rc ? "true" : "false")
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

// ------------------------------------------------------------------------------------------ merge / delete / weed
MergeSkaArray MergeSkaArray::extend(MergeSkaArray& a, MergeSkaArray& b) {
    ska_array_t* h = nullptr;
    check(skagpu_merge_arrays(a.h_, b.h_, &h));
    std::vector<std::string> names = a.names_;
    names.insert(names.end(), b.names_.begin(), b.names_.end());                    // merge_ska_dict.rs:169
    MergeSkaArray out(h, names, a.version_);
    out.k_bits_field_ = a.k_bits_field_;
    return out;
}

void MergeSkaArray::delete_samples(const std::vector<std::string>& del_names) {
    if (del_names.empty() || del_names.size() == nsamples()) throw SkaError("Invalid number of samples to remove");   // merge_ska_array.rs:232-234
    std::vector<std::string> wanted = del_names, kept;
    std::vector<uint32_t> idx;
    for (size_t i = 0; i < names_.size(); i++) {
        auto it = std::find(wanted.begin(), wanted.end(), names_[i]);
        if (it != wanted.end()) { idx.push_back((uint32_t)i); wanted.erase(it); }
        else kept.push_back(names_[i]);
    }
    if (!wanted.empty()) {
        std::string msg = "Could not find sample(s): {";
        for (size_t i = 0; i < wanted.size(); i++) msg += (i ? ", \"" : "\"") + wanted[i] + "\"";
        throw SkaError(msg + "}");                                                  // merge_ska_array.rs:251-253
    }
    check(skagpu_delete_samples(h_, idx.data(), idx.size()));
    names_ = kept;
}

uint64_t MergeSkaArray::weed(const std::string& weed_fasta, bool reverse) {
    fastx::Records r = fastx::parse_file(weed_fasta);
    if (r.is_fastq) throw SkaError("Cannot create reference from FASTQ files");      // ska_ref.rs:204-206
    ska_input_t in = {r.seq.data(), nullptr, r.rec_off.data(), r.n_rec(), r.n_rec()};
    uint64_t n_weed = 0;
    if (skagpu_weed(h_, &in, reverse, &n_weed)) {
        std::string msg = skagpu_last_error();
        if (msg.find("no valid sequence") != std::string::npos) msg = weed_fasta + " has no valid sequence";
        throw SkaError(msg);
    }
    return n_weed;
}

// ------------------------------------------------------------------------------------------ RefSka
RefSka::RefSka(size_t k, const std::string& filename, bool rc, bool ambig_mask, bool repeat_mask)
    : k_(k), rc_(rc), ambig_mask_(ambig_mask), repeat_mask_(repeat_mask) {
    if (k < 5 || k > 63 || k % 2 == 0) throw SkaError("Invalid k-mer length");       // ska_ref.rs:190-192
    fastx::NamedRecords r = fastx::parse_file_named(filename);
    if (r.rec.is_fastq) throw SkaError("Cannot create reference from FASTQ files");  // ska_ref.rs:204-206
    for (size_t i = 0; i < r.rec.n_rec(); i++) {
        seq_.emplace_back(r.rec.seq.begin() + r.rec.rec_off[i], r.rec.seq.begin() + r.rec.rec_off[i + 1]);
        std::string id = r.ids[i];                                                   // first whitespace-separated word (ska_ref.rs:207-211)
        size_t b = id.find_first_not_of(" \t"), e = id.find_first_of(" \t", b == std::string::npos ? 0 : b);
        chrom_names_.push_back(b == std::string::npos ? "" : id.substr(b, e == std::string::npos ? std::string::npos : e - b));
    }
    flat_ = std::move(r.rec.seq);
    rec_off_ = std::move(r.rec.rec_off);
}

void RefSka::map(MergeSkaArray& ska_array) {
    if (k_ != ska_array.kmer_len()) throw SkaError("K-mer sizes do not match ref:" + std::to_string(k_) + " skf:" + std::to_string(ska_array.ksize()));
    ska_input_t in = {flat_.data(), nullptr, rec_off_.data(), rec_off_.size() - 1, rec_off_.size() - 1};
    ska_mapping_t* m = nullptr;
    if (skagpu_map(ska_array.handle(), &in, &m)) throw SkaError(skagpu_last_error());
    uint64_t n_mapped = 0, n_ref = 0, n_samples = 0, n_pos = 0;
    check(skagpu_map_dims(m, &n_mapped, &n_ref, &n_samples, &n_pos));
    n_ref_kmers_ = n_ref;
    std::vector<uint64_t> mid(n_mapped), records;
    mapped_variants_.resize(n_mapped * n_samples);
    const int kw = k_ > 31 ? 2 : 1;
    if (repeat_mask_) records.resize(n_pos * kw);
    check(skagpu_map_export(m, mid.data(), mapped_variants_.data(), repeat_mask_ ? records.data() : nullptr));
    skagpu_map_free(m);
    mapped_names_ = ska_array.names();
    mapped_pos_.resize(n_mapped);
    size_t chrom = 0;
    for (size_t i = 0; i < n_mapped; i++) {                                          // reference order: chromosomes ascend
        while (chrom + 1 < rec_off_.size() - 1 && mid[i] >= rec_off_[chrom + 1]) chrom++;
        mapped_pos_[i] = {chrom, (size_t)(mid[i] - rec_off_[chrom])};
    }
    if (repeat_mask_) {
        // split k-mers seen more than once in the reference (track_repeats, ska_ref.rs:485-496), then their spans (:256-296)
        const size_t half = (k_ - 1) / 2;
        std::vector<std::pair<std::pair<uint64_t, uint64_t>, size_t>> ks;            // (key, position)
        for (size_t p = 0; p < n_pos; p++) {
            uint64_t lo, hi;
            if (kw == 1) { if (records[p] == ~0ULL) continue; lo = records[p] >> 4; hi = 0; }
            else { if ((records[2 * p] & records[2 * p + 1]) == ~0ULL) continue; lo = (records[2 * p] >> 4) | (records[2 * p + 1] << 60); hi = records[2 * p + 1] >> 4; }
            ks.push_back({{hi, lo}, p});
        }
        std::vector<std::pair<std::pair<uint64_t, uint64_t>, size_t>> sorted = ks;
        std::sort(sorted.begin(), sorted.end());
        std::vector<char> is_rep(n_pos, 0);
        size_t n_repeats = 0;
        for (size_t i = 0; i < sorted.size();) {
            size_t j = i;
            while (j < sorted.size() && sorted[j].first == sorted[i].first) j++;
            if (j - i > 1) { n_repeats++; for (size_t q = i; q < j; q++) is_rep[sorted[q].second] = 1; }
            i = j;
        }
        size_t last_end = 0;
        for (auto& e : ks) {                                                          // window start p, middle at p + half (concatenated coordinates)
            if (!is_rep[e.second]) continue;
            const size_t start = e.second, end = e.second + 2 * half;
            const size_t from = (start > last_end || start == 0) ? start : last_end + 1;
            for (size_t pos = from; pos < end + 1; pos++) repeat_coors_.push_back(pos);
            last_end = end;
        }
        (void)n_repeats;
    }
}

namespace {
// AlnWriter (src/ska_ref/aln_writer.rs): the pseudo-alignment of ONE sample against the reference
struct AlnWriter {
    size_t next_pos, curr_chrom = 0, last_mapped = 0, last_written = 0, chrom_offset = 0, half;
    const std::vector<std::vector<uint8_t>>& ref;
    std::vector<uint8_t> out;
    bool mask_ambig;
    std::vector<std::pair<uint8_t, size_t>> middle_out;
    AlnWriter(const std::vector<std::vector<uint8_t>>& r, size_t k, bool mask) : half((k - 1) / 2), ref(r), mask_ambig(mask) {
        size_t total = 0; for (auto& c : r) total += c.size();
        out.assign(total, '-');
        next_pos = half;
    }
    static bool is_ambiguous(uint8_t b) { b |= 0x20; return !(b == 'a' || b == 'c' || b == 'g' || b == 't' || b == 'u' || b == '-'); }   // bit_encoding.rs:58-61
    void fill_fwd_bases(size_t maximum) {                                              // aln_writer.rs:97-111
        if (last_written > 0) {
            const size_t sum = last_mapped + half;
            const size_t overhang = sum > last_written ? sum - last_written : 0;       // saturating_sub
            const size_t start = last_written + 1;
            const size_t end = std::min(start + overhang, maximum);
            if (end > start) {
                std::copy(ref[curr_chrom].begin() + start, ref[curr_chrom].begin() + end, out.begin() + start + chrom_offset);
                last_written = end;
            }
        }
    }
    void fill_contig() {                                                               // aln_writer.rs:114-120
        const size_t len = ref[curr_chrom].size();
        fill_fwd_bases(len);
        chrom_offset += len; curr_chrom += 1; next_pos = half;
    }
    void write_split_kmer(size_t mapped_pos, size_t mapped_chrom, uint8_t base) {      // aln_writer.rs:127-160
        while (mapped_chrom > curr_chrom) fill_contig();
        middle_out.push_back({(uint8_t)((is_ambiguous(base) && mask_ambig) ? 'N' : base), mapped_pos + chrom_offset});
        if (mapped_pos < next_pos) last_mapped = mapped_pos;
        else {
            if (mapped_pos > next_pos) fill_fwd_bases(mapped_pos - half);
            const size_t start = mapped_pos - half, end = mapped_pos;
            std::copy(ref[curr_chrom].begin() + start, ref[curr_chrom].begin() + end, out.begin() + start + chrom_offset);
            next_pos = mapped_pos + half + 1; last_mapped = mapped_pos; last_written = mapped_pos;
        }
    }
    void finalise(const std::vector<size_t>& repeats) {                                // aln_writer.rs:164-181
        while (curr_chrom < ref.size()) fill_contig();
        for (auto& m : middle_out) out[m.second] = m.first;
        for (size_t r : repeats) if (r < out.size() && out[r] != '-') out[r] = 'N';
    }
};
}  // namespace

std::vector<std::vector<uint8_t>> RefSka::pseudoalignment(size_t threads) const {
    if (mapped_pos_.empty()) throw SkaError("No split k-mers mapped to reference");    // ska_ref.rs:556-558
    const size_t n = mapped_names_.size();
    std::vector<std::vector<uint8_t>> seqs(n);
    std::atomic<size_t> next{0};
    auto work = [&] {
        for (size_t s; (s = next.fetch_add(1)) < n;) {
            AlnWriter w(seq_, k_, ambig_mask_);
            for (size_t i = 0; i < mapped_pos_.size(); i++) {
                const uint8_t base = mapped_variants_[i * n + s];
                if (base != '-') w.write_split_kmer(mapped_pos_[i].second, mapped_pos_[i].first, base);
            }
            w.finalise(repeat_coors_);
            seqs[s] = std::move(w.out);
        }
    };
    const size_t nt = std::max<size_t>(1, std::min(threads, n));
    std::vector<std::thread> th;
    for (size_t t = 1; t < nt; t++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    return seqs;
}

void RefSka::write_aln(FILE* f, size_t threads) {
    if (chrom_names_.size() > 1) fprintf(stderr, "WARN Reference contained multiple contigs, in the output they will be concatenated\n");
    std::vector<std::vector<uint8_t>> aln = pseudoalignment(threads);
    for (size_t s = 0; s < aln.size(); s++) {
        fputc('>', f); fwrite(mapped_names_[s].data(), 1, mapped_names_[s].size(), f); fputc('\n', f);
        fwrite(aln[s].data(), 1, aln[s].size(), f); fputc('\n', f);
    }
}

void RefSka::write_vcf(FILE* f, size_t threads) {
    std::vector<std::vector<uint8_t>> aln = pseudoalignment(threads);
    auto base_of = [](uint8_t b) -> char { return (b == 'A' || b == 'C' || b == 'G' || b == 'T') ? (char)b : 'N'; };   // u8_to_base, ska_ref.rs:152-160
    fputs("##fileformat=VCFv4.4\n", f);                                               // noodles-vcf's default header
    for (auto& c : chrom_names_) fprintf(f, "##contig=<ID=%s>\n", c.c_str());
    fputs("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT", f);
    for (auto& n : mapped_names_) fprintf(f, "\t%s", n.c_str());
    fputc('\n', f);
    const size_t n = aln.size();
    size_t idx = 0;
    std::string line;
    for (size_t chrom = 0; chrom < seq_.size(); chrom++)
        for (size_t pos = 0; pos < seq_[chrom].size(); pos++, idx++) {                 // IdxCheck (src/ska_ref/idx_check.rs)
            const uint8_t ref_base = seq_[chrom][pos];
            std::string alts; std::vector<char> alt_bases; bool variant = false;
            line.clear();
            for (size_t s = 0; s < n; s++) {
                const uint8_t b = aln[s][idx];
                line += '\t';
                if (b == ref_base) line += '0';
                else if (b == '-') { variant = true; line += '.'; }
                else {
                    variant = true;
                    const char ab = base_of(b);
                    size_t a = std::find(alt_bases.begin(), alt_bases.end(), ab) - alt_bases.begin();
                    if (a == alt_bases.size()) alt_bases.push_back(ab);
                    line += std::to_string(a + 1);
                }
            }
            if (!variant) continue;
            for (size_t a = 0; a < alt_bases.size(); a++) { if (a) alts += ','; alts += alt_bases[a]; }
            fprintf(f, "%s\t%zu\t.\t%c\t%s\t.\t.\t.\tGT%s\n", chrom_names_[chrom].c_str(), pos + 1, base_of(ref_base),
                    alts.empty() ? "." : alts.c_str(), line.c_str());
        }
}

// ------------------------------------------------------------------------------------------ CoverageHistogram
namespace {
// coverage.rs:277-302
double lse(double a, double b) { const double x = std::max(a, b); return x + std::log(std::exp(a - x) + std::exp(b - x)); }
double ln_dpois(double x, double lambda) { return x * std::log(lambda) - std::lgamma(x + 1.0) - lambda; }
double comp_a(double w0, double i) { return std::log(w0) + ln_dpois(i, 1.0); }
double comp_b(double w0, double c, double i) { return std::log(1.0 - w0) + ln_dpois(i, c); }
double log_likelihood(const double p[2], const std::vector<double>& counts) {             // coverage.rs:305-320
    const double w0 = p[0], c = p[1];
    if (!(w0 >= 0.0 && w0 <= 1.0) || c < 1.0) return -std::numeric_limits<double>::max();   // f64::MIN, the 'soft' bounds
    double ll = 0.0;
    for (size_t i = 0; i < counts.size(); i++) ll += counts[i] * lse(comp_a(w0, (double)i + 1.0), comp_b(w0, c, (double)i + 1.0));
    return ll;
}
void grad_ll(const double p[2], const std::vector<double>& counts, double g[2]) {           // coverage.rs:324-341
    const double w0 = p[0], c = p[1];
    g[0] = g[1] = 0.0;
    for (size_t i = 0; i < counts.size(); i++) {
        const double x = (double)i + 1.0, a = comp_a(w0, x), b = comp_b(w0, c, x);
        const double dlda = 1.0 / (1.0 + std::exp(b - a)), dldb = 1.0 / (1.0 + std::exp(a - b));
        g[0] += counts[i] * (dlda / w0 - dldb / (1.0 - w0));
        g[1] += counts[i] * (dldb * (x / c - 1.0));
    }
}
size_t find_cutoff(const double p[2], size_t max_cutoff) {                                   // coverage.rs:345-360
    size_t cutoff = 1;
    while (cutoff < max_cutoff) {
        if (comp_a(p[0], (double)cutoff) - comp_b(p[0], p[1], (double)cutoff) < 0.0) break;
        cutoff++;
    }
    return cutoff;
}
// Rust's `{:e}` for f64: shortest digits that round-trip, exponent without sign padding
std::string rust_lower_exp(double x) {
    if (x == 0.0) return "0e0";
    char buf[64];
    for (int prec = 0; prec < 18; prec++) {
        snprintf(buf, sizeof buf, "%.*e", prec, x);
        if (strtod(buf, nullptr) == x) break;
    }
    std::string s(buf);
    const size_t e = s.find('e');
    std::string mant = s.substr(0, e);
    int ex = atoi(s.c_str() + e + 1);
    if (mant.find('.') != std::string::npos) { while (mant.back() == '0') mant.pop_back(); if (mant.back() == '.') mant.pop_back(); }
    return mant + "e" + std::to_string(ex);
}
}  // namespace

CoverageHistogram::CoverageHistogram(const std::string& fastq1, const std::string& fastq2, size_t k, bool rc, bool verbose) : verbose_(verbose) {
    if (k < 5 || k > 63 || k % 2 == 0) throw SkaError("Invalid k-mer length");                // coverage.rs:71-73
    fastx::Records r1 = fastx::parse_file(fastq1), r2 = fastx::parse_file(fastq2);
    for (auto* p : {&r1, &r2})
        if (!p->is_fastq)
            throw SkaError((p == &r1 ? fastq1 : fastq2) + " appears to be FASTA.\nCoverage can only be used with FASTQ files, not FASTA.");
    fastx::append(r1, r2);
    ska_input_t in = {r1.seq.data(), nullptr, r1.rec_off.data(), r1.n_rec(), r1.n_rec()};
    hist_.assign(1000, 0);
    check(skagpu_kmer_count_hist(&in, (uint32_t)k, rc, hist_.data()));
}

size_t CoverageHistogram::fit_histogram() {
    if (fitted_) throw SkaError("Model already fitted");                                        // coverage.rs:152-154
    if (!from_counts_) { counts_.assign(1000, 0); for (size_t i = 0; i < 1000; i++) counts_[i] = (uint32_t)std::min<uint64_t>(hist_[i], 0xFFFFFFFFu); }
    while (!counts_.empty() && counts_.back() < 50) counts_.pop_back();                         // MIN_FREQ (coverage.rs:166-173)
    std::vector<double> cf(counts_.begin(), counts_.end());
    // BFGS (argmin::solver::quasinewton::BFGS) on the negative log likelihood, backtracking line search with the Armijo condition
    // c1 = 1e-4 and contraction 0.9 (argmin's BacktrackingLineSearch defaults), inverse Hessian = identity, tolerances: gradient
    // norm sqrt(eps), cost change 1e-6; at most 20 iterations.
    double x[2] = {w0_, c_}, g[2], H[2][2] = {{1.0, 0.0}, {0.0, 1.0}};
    double cost = -log_likelihood(x, cf), prev_cost = std::numeric_limits<double>::infinity();
    grad_ll(x, cf, g); g[0] = -g[0]; g[1] = -g[1];
    const double tol_grad = std::sqrt(std::numeric_limits<double>::epsilon()), tol_cost = 1e-6;
    bool converged = false;
    for (int iter = 0;; iter++) {
        if (std::sqrt(g[0] * g[0] + g[1] * g[1]) < tol_grad || std::fabs(prev_cost - cost) < tol_cost) { converged = true; break; }
        if (iter >= 20) break;                                                                  // max_iters(20)
        const double p[2] = {-(H[0][0] * g[0] + H[0][1] * g[1]), -(H[1][0] * g[0] + H[1][1] * g[1])};
        const double slope = g[0] * p[0] + g[1] * p[1];
        double alpha = 1.0, xn[2], cn = 0.0;
        for (int ls = 0; ls < 2000; ls++) {                                                     // f(x + a p) <= f(x) + c1 a grad.p
            xn[0] = x[0] + alpha * p[0]; xn[1] = x[1] + alpha * p[1];
            cn = -log_likelihood(xn, cf);
            if (cn <= cost + 1e-4 * alpha * slope) break;
            alpha *= 0.9;
        }
        double gn[2];
        grad_ll(xn, cf, gn); gn[0] = -gn[0]; gn[1] = -gn[1];
        const double s[2] = {xn[0] - x[0], xn[1] - x[1]}, y[2] = {gn[0] - g[0], gn[1] - g[1]};
        const double ys = y[0] * s[0] + y[1] * s[1], rho = 1.0 / ys;
        // H <- (I - rho s y^T) H (I - rho y s^T) + rho s s^T
        double A[2][2] = {{1.0 - rho * s[0] * y[0], -rho * s[0] * y[1]}, {-rho * s[1] * y[0], 1.0 - rho * s[1] * y[1]}};
        double B[2][2] = {{1.0 - rho * y[0] * s[0], -rho * y[0] * s[1]}, {-rho * y[1] * s[0], 1.0 - rho * y[1] * s[1]}};
        double T[2][2], Hn[2][2];
        for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) T[i][j] = A[i][0] * H[0][j] + A[i][1] * H[1][j];
        for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) Hn[i][j] = T[i][0] * B[0][j] + T[i][1] * B[1][j] + rho * s[i] * s[j];
        for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) H[i][j] = Hn[i][j];
        prev_cost = cost; cost = cn;
        x[0] = xn[0]; x[1] = xn[1]; g[0] = gn[0]; g[1] = gn[1];
        if (verbose_) fprintf(stderr, "iter %d cost %.6f w0 %.6f c %.6f\n", iter + 1, cost, x[0], x[1]);
    }
    if (!converged) throw SkaError("Couldn't fit coverage model: Optimiser did not converge: Maximum number of iterations reached");
    w0_ = x[0]; c_ = x[1];
    cutoff_ = find_cutoff(x, counts_.size());
    fitted_ = true;
    return cutoff_;
}

std::string CoverageHistogram::plot_hist() const {
    if (!fitted_) throw SkaError("Model has not yet been fitted");                              // coverage.rs:221-223
    std::string out = "Count\tK_mers\tMixture_density\tComponent\n";
    for (size_t idx = 0; idx < counts_.size(); idx++) {
        const double x = (double)idx + 1.0;
        out += std::to_string(idx + 1) + "\t" + std::to_string(counts_[idx]) + "\t" + rust_lower_exp(std::exp(lse(comp_a(w0_, x), comp_b(w0_, c_, x)))) +
               "\t" + ((idx + 1) < cutoff_ ? "Error" : "Coverage") + "\n";
    }
    return out;
}

// ------------------------------------------------------------------------------------------ build_and_merge
static int g_gpus = 1;
void set_gpus(int n) { g_gpus = std::max(1, n); }
int gpus() { return g_gpus; }

// `ska build --gpus N`: one host thread per GPU, each bound to its device (skagpu_set_device).  Rank r builds the array of its
// contiguous share of the samples, the library exchanges partial rows over NVLink (skagpu_exchange_rows) and every rank ends
// with an ascending key range of the whole array; the slabs concatenated in rank order ARE the single-GPU array, which is
// put back on device 0 so that everything downstream (save, align, distance) is unchanged.
// build_local(r, lo, hi, &h): builds samples [lo, hi) on the calling thread's device, returns an error text or "".
static MergeSkaArray build_sharded(size_t n, const std::vector<std::string>& names, int world,
                                   const std::function<std::string(int, size_t, size_t, ska_array_t**)>& build_local) {
    uint8_t id[SKAGPU_COMM_ID_BYTES];
    check(skagpu_set_device(0));
    check(skagpu_comm_unique_id(id));
    std::vector<skf::Array> slabs(world);
    std::vector<std::string> errors(world);
    auto rank_main = [&](int r) {
        auto err = [&](const char* what) { errors[r] = std::string(what) + ": " + skagpu_last_error(); };
        if (skagpu_set_device(r)) return err("device");
        if (skagpu_comm_init(id, world, r)) return err("communicator");
        const size_t lo = n * (size_t)r / world, hi = n * (size_t)(r + 1) / world;
        ska_array_t* local = nullptr;
        std::string e = build_local(r, lo, hi, &local);
        uint64_t failed = e.empty() ? 0 : 1;                          // a rank that failed still joins the collectives
        if (skagpu_allreduce_u64(&failed, 1)) { err("status exchange"); failed = 1; }
        if (failed) { if (!e.empty()) errors[r] = e; if (local) skagpu_free(local); skagpu_comm_destroy(); return; }
        ska_array_t* slab = nullptr;
        uint64_t stats[4];
        if (skagpu_exchange_rows(local, lo, n, &slab, stats)) { err("row exchange"); skagpu_free(local); skagpu_comm_destroy(); return; }
        skagpu_free(local);
        if (getenv("SKA_TIMING"))
            fprintf(stderr, "[gpu %d] samples [%zu, %zu): %llu rows sent, %llu received, slab of %llu rows\n", r, lo, hi,
                    (unsigned long long)stats[0], (unsigned long long)stats[1], (unsigned long long)stats[3]);
        skf::Array& a = slabs[r];
        uint64_t rows, keys, samples; uint32_t k, bits; int rc;
        if (skagpu_dims(slab, &rows, &keys, &samples, &k, &rc, &bits)) { err("dims"); }
        else {
            a.k = k; a.rc = rc != 0; a.key_bits = bits; a.n_rows = rows; a.n_samples = samples;
            a.keys.resize(keys * (bits / 64)); a.variants.resize(rows * samples); a.counts.resize(rows);
            if (skagpu_export(slab, a.keys.data(), a.variants.data(), a.counts.data())) err("export");
        }
        skagpu_free(slab);
        skagpu_comm_destroy();
    };
    std::vector<std::thread> th;
    for (int r = 0; r < world; r++) th.emplace_back(rank_main, r);
    for (auto& t : th) t.join();
    for (auto& e : errors) if (!e.empty()) throw SkaError(e);
    lap("slabs built on the devices and exported");
    skf::Array all;
    all.k = slabs[0].k; all.rc = slabs[0].rc; all.key_bits = slabs[0].key_bits; all.n_samples = n;
    for (auto& a : slabs) all.n_rows += a.n_rows;
    all.keys.reserve(all.n_rows); all.variants.resize(all.n_rows * n); all.counts.reserve(all.n_rows);
    uint64_t row = 0;
    for (auto& a : slabs) {
        all.keys.insert(all.keys.end(), a.keys.begin(), a.keys.end());
        all.counts.insert(all.counts.end(), a.counts.begin(), a.counts.end());
        if (a.n_rows) memcpy(all.variants.data() + row * n, a.variants.data(), a.n_rows * n);
        row += a.n_rows;
        a = skf::Array();
    }
    ska_array_t* h = nullptr;
    check(skagpu_array_from_host(all.k, all.rc, all.key_bits, all.n_rows, n, all.keys.data(), all.variants.data(), all.counts.data(), &h));
    lap("array on device 0");
    return MergeSkaArray(h, names);
}

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
    const int world = (int)std::min<size_t>((size_t)g_gpus, n);
    if (world > 1 && k > 31) throw SkaError("--gpus: k > 31 builds on one GPU only");
    if (all_packed) {
        std::vector<ska_input_packed_t> pin(n);
        for (size_t i = 0; i < n; i++) pin[i] = {packed2[i].data(), invalid[i].data(), recs[i].rec_off.data(), recs[i].n_rec(), n_file1[i]};
        if (world > 1)
            return build_sharded(n, names, world, [&](int, size_t lo, size_t hi, ska_array_t** h) -> std::string {
                if (skagpu_build_packed(pin.data() + lo, hi - lo, (uint32_t)k, rc, h)) {
                    std::string msg = skagpu_last_error();
                    size_t p = msg.find("sample ");
                    if (p == 0 && msg.find("has no valid sequence") != std::string::npos) {
                        size_t idx = lo + std::stoul(msg.substr(7));
                        if (idx < n) msg = std::get<1>(input_files[idx]) + " has no valid sequence";
                    }
                    return msg;
                }
                return "";
            });
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
    if (world > 1)
        return build_sharded(n, names, world, [&](int, size_t lo, size_t hi, ska_array_t** h) -> std::string {
            if (skagpu_build(in.data() + lo, hi - lo, (uint32_t)k, rc, qual.min_count, qual.min_qual, (int)qual.qual_filter, proportion_reads, h)) {
                std::string msg = skagpu_last_error();
                size_t p = msg.find("sample ");
                if (p == 0 && msg.find("has no valid sequence") != std::string::npos) {
                    size_t idx = lo + std::stoul(msg.substr(7));
                    if (idx < n) msg = std::get<1>(input_files[idx]) + " has no valid sequence";
                }
                return msg;
            }
            return "";
        });
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
uint16_t kmer_min_cutoff(const std::string& v, const std::vector<InputFastx>& files, size_t k, bool rc, bool verbose) {
    if (v.empty()) return 5;                                                                    // DEFAULT_MINCOUNT (cli.rs:25)
    if (v != "auto") return (uint16_t)std::stoul(v);                                            // validated by the CLI parser
    std::vector<std::string> fq;                                                               // get_2_fastq_path: file 1 of the first two paired samples
    for (auto& f : files) if (!std::get<2>(f).empty() && fq.size() < 2) fq.push_back(std::get<1>(f));
    if (fq.size() < 2) return 5;                                                               // "Not enough fastq files to fit mixture model"
    CoverageHistogram cov(fq[0], fq[1], k, rc, verbose);
    const size_t cutoff = cov.fit_histogram();
    const std::string table = cov.plot_hist();
    fwrite(table.data(), 1, table.size(), stdout);                                             // cov.plot_hist() (io_utils.rs:201)
    return (uint16_t)cutoff;
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

void map(MergeSkaArray& a, RefSka& ska_ref, const std::string& output, bool vcf, size_t threads) {
    ska_ref.map(a);
    FILE* f = set_ostream(output);
    if (vcf) ska_ref.write_vcf(f, threads); else ska_ref.write_aln(f, threads);
    if (f != stdout) fclose(f); else fflush(f);
}

void merge(MergeSkaArray& first, const std::vector<std::string>& skf_files, const std::string& output, int threads) {
    MergeSkaArray merged = std::move(first);
    for (auto& f : skf_files) {
        MergeSkaArray next = MergeSkaArray::load(f);
        MergeSkaArray joined = MergeSkaArray::extend(merged, next);
        merged = std::move(joined);
    }
    save_skf(merged, output, threads);
}

void delete_(MergeSkaArray& a, const std::vector<std::string>& names, const std::string& out_file, int threads) {
    a.delete_samples(names);
    save_skf(a, out_file, threads);                        // suffix added if missing (generic_modes.rs:200-205)
}

void weed(MergeSkaArray& a, const std::string& weed_file, bool reverse, double min_freq, bool filter_ambig_as_missing, FilterType filter,
          bool ambig_mask, bool ignore_const_gaps, const std::string& out_file, int threads) {
    if (!weed_file.empty()) a.weed(weed_file, reverse);
    const size_t threshold = (size_t)std::floor((double)a.nsamples() * min_freq);        // floor here, ceil in align (generic_modes.rs:249)
    if (threshold > 0 || filter != FilterType::NoFilter || ambig_mask || ignore_const_gaps)
        a.filter(threshold, filter_ambig_as_missing, filter, ambig_mask, ignore_const_gaps, true);
    a.save(out_file, threads);                             // exactly the name given (generic_modes.rs:262-265)
}

void save_skf(const MergeSkaArray& a, const std::string& output, int threads) {
    std::string path = output;
    if (path.size() < 4 || path.compare(path.size() - 4, 4, ".skf") != 0) path += ".skf";   // generic_modes.rs:270-283
    a.save(path, threads);
}
}  // namespace generic_modes

}  // namespace ska
