// ska_ref_cli.cpp -- TEST INFRASTRUCTURE ONLY (bench.py --workload cli, kind "port"): the reference's `ska build` and
// `ska align` commands FILE TO FILE on the host, as the CPU baseline of the command-line measurement.
//
// The algorithm is the oracle's (oracle/ska_oracle.cpp: build_and_merge with the reference's rayon::join thread rule,
// MergeSkaArray::new, filter, write_fasta); file handling follows the reference's shape:
//   * FASTA read per sample inside the build (needletail's parse_fastx_file, src/ska_dict.rs:131-147);
//   * save = ONE thread: CBOR (ciborium) through a Snappy frame encoder into a buffered file (src/merge_ska_array.rs:191-197);
//   * load = ONE thread (src/merge_ska_array.rs:199-204); align = filter + transpose + FASTA writer (generic_modes.rs:22-50).
// The .skf bytes come from this repo's host codec run on one thread (the reference's ciborium + snap are not available
// here); nothing under ska.rust_b200/ links or calls this program.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <strings.h>
#include <algorithm>
#include <vector>

#include "../ska.rust_b200/host/fastx.hpp"
#include "../ska.rust_b200/host/skf_codec.hpp"

extern "C" {
struct ora_sample_t { const uint8_t* seq; const uint8_t* qual; const uint64_t* rec_off; uint64_t n_rec; uint64_t n_rec_file1; };
const char* ora_last_error(void);
void* ora_build(const ora_sample_t* samples, uint64_t n, uint32_t k, int rc, uint16_t min_count, uint8_t min_qual, int qual_filter,
                double proportion_reads, int threads);
void* ora_from_host(uint32_t k, int rc, uint32_t key_bits, uint64_t n_rows, uint64_t n_samples, const uint64_t* keys_lohi,
                    const uint8_t* variants, const uint64_t* counts);
void ora_dims(void* h, uint64_t* n_rows, uint64_t* n_keys, uint64_t* n_samples, uint32_t* k, int* rc, uint32_t* key_bits);
void ora_export(void* h, uint64_t* keys_lohi, uint8_t* variants, uint64_t* counts);
void ora_export_transposed(void* h, uint8_t* out);
int ora_filter(void* h, uint64_t min_count, int filter_ambig_as_missing, int filter_type, int mask_ambig, int ignore_const_gaps,
               int update_kmers, int32_t* removed);
void ora_free(void* h);
}

static std::string sample_name(const std::string& f) {          // io_utils.rs:31-46 (extensions as in the benchmark: .fa)
    size_t s = f.find_last_of('/');
    std::string b = s == std::string::npos ? f : f.substr(s + 1);
    for (const char* ext : {".fastq.gz", ".fastq", ".fasta", ".fa"}) {
        const size_t n = strlen(ext);
        if (b.size() > n && strcasecmp(b.c_str() + b.size() - n, ext) == 0) return b.substr(0, b.size() - n);
    }
    return f;
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: ska_ref_cli build -o OUT -k K --threads T FILES... | align IN.skf -o OUT\n"); return 2; }
    const std::string cmd = argv[1];
    std::string out; int k = 31, threads = 1; std::vector<std::string> pos;
    for (int i = 2; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "-o" && i + 1 < argc) out = argv[++i];
        else if (a == "-k" && i + 1 < argc) k = atoi(argv[++i]);
        else if (a == "--threads" && i + 1 < argc) threads = atoi(argv[++i]);
        else pos.push_back(a);
    }
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        fprintf(stderr, "[ska_ref_cli] %s: %.3f s\n", what, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    };
    try {
        if (cmd == "build") {
            // the reference parses each file inside the (possibly threaded) build; here: parse serially, then build with
            // the reference's thread rule -- parsing 5 MB FASTA files is ~1 % of the build either way
            const size_t n = pos.size();
            std::vector<fastx::Records> recs(n);
            std::vector<ora_sample_t> in(n);
            std::vector<std::string> names(n);
            for (size_t i = 0; i < n; i++) {
                recs[i] = fastx::parse_file(pos[i]);
                in[i] = {recs[i].seq.data(), nullptr, recs[i].rec_off.data(), recs[i].n_rec(), recs[i].n_rec()};
                names[i] = sample_name(pos[i]);
            }
            lap("parsed");
            int m = (int)std::max<size_t>(1, std::min<size_t>((size_t)threads, 1 + n / 10));     // merge_ska_dict.rs:381-385
            int t = 1; while (2 * t <= m) t *= 2;
            void* h = ora_build(in.data(), n, (uint32_t)k, 1, 5, 20, 2, 0.0, t);
            if (!h) throw std::runtime_error(ora_last_error());
            lap("built");
            skf::Array a;
            uint64_t rows, keys, samples; uint32_t kk, bits; int rc;
            ora_dims(h, &rows, &keys, &samples, &kk, &rc, &bits);
            a.k = kk; a.rc = rc != 0; a.key_bits = bits; a.n_rows = rows; a.n_samples = samples; a.names = names;
            std::vector<uint64_t> lohi(2 * keys);
            a.variants.resize(rows * samples); a.counts.resize(rows);
            ora_export(h, lohi.data(), a.variants.data(), a.counts.data());
            ora_free(h);
            if (bits == 64) { a.keys.resize(keys); for (uint64_t i = 0; i < keys; i++) a.keys[i] = lohi[2 * i]; }
            else a.keys = std::move(lohi);
            if (out.size() < 4 || out.compare(out.size() - 4, 4, ".skf") != 0) out += ".skf";
            skf::save(a, out, 1);                               // single-threaded, like ciborium -> snap::FrameEncoder -> BufWriter
            lap("saved");
        } else if (cmd == "align") {
            skf::Array a = skf::decode_file_single_thread(pos.at(0));
            lap("loaded");
            std::vector<uint64_t> lohi(2 * a.n_keys());
            for (uint64_t i = 0; i < a.n_keys(); i++) {
                lohi[2 * i] = a.key_bits == 64 ? a.keys[i] : a.keys[2 * i];
                lohi[2 * i + 1] = a.key_bits == 64 ? 0 : a.keys[2 * i + 1];
            }
            void* h = ora_from_host(a.k, a.rc, a.key_bits, a.n_rows, a.n_samples, lohi.data(), a.variants.data(), a.counts.data());
            if (!h) throw std::runtime_error(ora_last_error());
            int32_t removed = 0;
            const uint64_t thr = (uint64_t)std::ceil((double)a.n_samples * 0.9);                 // generic_modes.rs:121, cli.rs:15
            if (ora_filter(h, thr, 0, 1 /* no-const */, 0, 0, 0, &removed)) throw std::runtime_error(ora_last_error());
            uint64_t rows, keys, samples; uint32_t kk, bits; int rc;
            ora_dims(h, &rows, &keys, &samples, &kk, &rc, &bits);
            std::vector<uint8_t> t(rows * samples);
            ora_export_transposed(h, t.data());
            ora_free(h);
            lap("filtered");
            FILE* f = out.empty() ? stdout : fopen(out.c_str(), "wb");
            if (!f) throw std::runtime_error("cannot create " + out);
            for (uint64_t s = 0; s < samples; s++) {
                fputc('>', f); fwrite(a.names[s].data(), 1, a.names[s].size(), f); fputc('\n', f);
                fwrite(t.data() + s * rows, 1, rows, f); fputc('\n', f);
            }
            if (f != stdout) fclose(f);
            lap("written");
        } else {
            fprintf(stderr, "unknown command\n");
            return 2;
        }
    } catch (std::exception& e) {
        fprintf(stderr, "thread 'main' panicked:\n%s\n", e.what());
        return 101;
    }
    return 0;
}
