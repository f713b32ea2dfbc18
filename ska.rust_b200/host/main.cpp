// main.cpp -- `ska` command line for the hot path: build / align / distance / nk, with the reference's
// flags and defaults (src/cli.rs:153-431, dispatch src/lib.rs:557-892).  stdout carries only results
// (alignment, distance table, nk text); banner and timing go to stderr as in the reference
// (lib.rs:565,886-890).  Any error exits with status 101, the status of a Rust panic.
#include <chrono>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <strings.h>
#include <thread>
#include <vector>

#include "ska_host.hpp"

using namespace ska;

[[noreturn]] static void usage_error(const std::string& msg) {
    fprintf(stderr, "error: %s\n\nUsage: ska <build|align|map|distance|merge|delete|weed|nk> [OPTIONS]\n", msg.c_str());
    exit(2);                                               // clap's exit status for argument errors
}

struct Args {
    std::vector<std::string> pos;
    std::vector<std::pair<std::string, std::string>> opts;
    bool has(const std::string& k) const { for (auto& o : opts) if (o.first == k) return true; return false; }
    std::string get(const std::string& k, const std::string& dflt = "") const { for (auto& o : opts) if (o.first == k) return o.second; return dflt; }
};

static Args parse(int argc, char** argv, int start, const std::vector<std::string>& value_opts, const std::vector<std::string>& flag_opts) {
    Args a;
    for (int i = start; i < argc; i++) {
        std::string s = argv[i];
        if (s == "-v" || s == "--verbose") { a.opts.emplace_back("--verbose", "1"); continue; }
        if (s == "--gpus") {                                  // not in the reference: GPUs of this node to shard a build over
            if (i + 1 >= argc) usage_error("a value is required for '--gpus'");
            int g = atoi(argv[++i]);
            if (g < 1 || g > 64) usage_error("invalid value for '--gpus <N>'");
            ska::set_gpus(g);
            continue;
        }
        bool is_val = false, is_flag = false;
        for (auto& v : value_opts) if (s == v) is_val = true;
        for (auto& f : flag_opts) if (s == f) is_flag = true;
        if (is_val) { if (i + 1 >= argc) usage_error("a value is required for '" + s + "'"); a.opts.emplace_back(s, argv[++i]); }
        else if (is_flag) a.opts.emplace_back(s, "1");
        else if (s.size() > 1 && s[0] == '-' && s != "-") usage_error("unexpected argument '" + s + "'");
        else a.pos.push_back(s);
    }
    return a;
}

static size_t parse_k(const std::string& s) {              // cli.rs:38-47
    char* e; long k = strtol(s.c_str(), &e, 10);
    if (*e || s.empty()) usage_error("`" + s + "` isn't a valid k-mer");
    if (k < 5 || k > 63 || k % 2 == 0) usage_error("K-mer must be an odd number between 5 and 63 (inclusive)");
    return (size_t)k;
}
static double parse_01(const std::string& s, const char* what) {   // cli.rs:50-71
    char* e; double f = strtod(s.c_str(), &e);
    if (*e || s.empty()) usage_error("`" + s + "` isn't a valid " + what);
    if (f < 0.0 || f > 1.0) usage_error("Frequency must be between 0 and 1 (inclusive)");
    return f;
}
static size_t parse_threads(const std::string& s) {        // cli.rs:74-83
    char* e; long t = strtol(s.c_str(), &e, 10);
    if (*e || s.empty()) usage_error("`" + s + "` isn't a valid number of cores");
    if (t < 1) usage_error("Threads must be one or higher");
    if ((unsigned long)t > std::thread::hardware_concurrency() && std::thread::hardware_concurrency())
        fprintf(stderr, "WARN %ld threads is greater than available cores %u\n", t, std::thread::hardware_concurrency());
    return (size_t)t;
}
static FilterType parse_filter(const std::string& s) {
    if (s == "no-filter") return FilterType::NoFilter;
    if (s == "no-const") return FilterType::NoConst;
    if (s == "no-ambig") return FilterType::NoAmbig;
    if (s == "no-ambig-or-const") return FilterType::NoAmbigOrConst;
    usage_error("invalid value '" + s + "' for '--filter <FILTER>'");
}
static QualFilter parse_qual_filter(const std::string& s) {
    if (s == "no-filter") return QualFilter::NoFilter;
    if (s == "middle") return QualFilter::Middle;
    if (s == "strict") return QualFilter::Strict;
    usage_error("invalid value '" + s + "' for '--qual-filter <QUAL_FILTER>'");
}

int main(int argc, char** argv) {
    if (argc < 2) usage_error("a subcommand is required");
    const std::string cmd = argv[1];
    fprintf(stderr, "SKA: Split K-mer Analysis (the alignment-free aligner)\n");
    if (const char* g = getenv("SKA_GPUS")) ska::set_gpus(atoi(g));
    // stdout carries alignments and distances: NCCL's banner must not land there (NCCL honours NCCL_DEBUG_FILE only above VERSION)
    setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
    if (const char* d = getenv("NCCL_DEBUG")) if (!strcasecmp(d, "VERSION")) setenv("NCCL_DEBUG", "WARN", 1);
    auto start = std::chrono::steady_clock::now();
    // CUDA context creation (~0.4 s on a B200) runs beside the host's first phase (file parsing, .skf decoding)
    std::thread cuda_init;
    if (cmd == "build" || cmd == "align" || cmd == "distance" || cmd == "nk" || cmd == "merge" || cmd == "weed" || cmd == "delete" || cmd == "map" || cmd == "cov")
        cuda_init = std::thread([] { skagpu_set_device(0); });
    try {
        if (cmd == "build") {
            Args a = parse(argc, argv, 2, {"-f", "-o", "-k", "--proportion-reads", "--min-count", "--min-qual", "--qual-filter", "--threads"},
                           {"--single-strand"});
            if (!a.has("-o")) usage_error("the following required arguments were not provided: -o <OUTPUT>");
            if (a.has("-f") == !a.pos.empty()) usage_error("exactly one of <SEQ_FILES> or -f <FILE_LIST> is required");
            size_t k = parse_k(a.get("-k", "31"));
            size_t threads = parse_threads(a.get("--threads", "1"));
            QualOpts q;
            q.qual_filter = parse_qual_filter(a.get("--qual-filter", "strict"));
            std::string mc = a.get("--min-count");          // valid_min_kmer (cli.rs:86-108): "auto" or an integer in 1..=65535
            if (!mc.empty() && mc != "auto") {
                char* e; long v = strtol(mc.c_str(), &e, 10);
                if (*e || v < 1 || v > 65535) usage_error("`" + mc + "` isn't a valid minimum kmer count: pass 'auto' or an integer in 1..65535");
            }
            {
                char* e; long v = strtol(a.get("--min-qual", "20").c_str(), &e, 10);
                if (*e || v < 0 || v > 255) usage_error("invalid value '" + a.get("--min-qual") + "' for '--min-qual <MIN_QUAL>'");
                q.min_qual = (uint8_t)v;
            }
            double prop = a.has("--proportion-reads") ? parse_01(a.get("--proportion-reads"), "proportion") : 0.0;
            auto inputs = io_utils::get_input_list(a.get("-f"), a.pos);
            q.min_count = io_utils::kmer_min_cutoff(mc, inputs, k, !a.has("--single-strand"), a.has("--verbose"));   // lib.rs:592-622
            MergeSkaArray arr = build_and_merge(inputs, k, !a.has("--single-strand"), q, threads, prop);
            generic_modes::save_skf(arr, a.get("-o"), (int)threads);
        } else if (cmd == "align") {
            Args a = parse(argc, argv, 2, {"-o", "-m", "--min-freq", "--filter", "--threads"},
                           {"--filter-ambig-as-missing", "--ambig-mask", "--no-gap-only-sites"});
            if (a.pos.empty()) usage_error("the following required arguments were not provided: <INPUT>...");
            double min_freq = parse_01(a.has("-m") ? a.get("-m") : a.get("--min-freq", "0.9"), "frequency");
            size_t threads = parse_threads(a.get("--threads", "1"));
            MergeSkaArray arr = io_utils::load_array(a.pos, threads);
            generic_modes::align(arr, a.get("-o"), parse_filter(a.get("--filter", "no-const")), a.has("--ambig-mask"),
                                 a.has("--no-gap-only-sites"), min_freq, a.has("--filter-ambig-as-missing"));
        } else if (cmd == "distance") {
            Args a = parse(argc, argv, 2, {"-o", "-m", "--min-freq", "--threads"}, {"--allow-ambiguous"});
            if (a.pos.size() != 1) usage_error("the following required arguments were not provided: <SKF_FILE>");
            double min_freq = parse_01(a.has("-m") ? a.get("-m") : a.get("--min-freq", "0"), "frequency");
            parse_threads(a.get("--threads", "1"));        // only sanity-checked by the reference too (lib.rs:714)
            MergeSkaArray arr = MergeSkaArray::load(a.pos[0]);
            generic_modes::distance(arr, a.get("-o"), min_freq, !a.has("--allow-ambiguous"));
        } else if (cmd == "nk") {
            Args a = parse(argc, argv, 2, {}, {"--full-info"});
            if (a.pos.size() != 1) usage_error("the following required arguments were not provided: <SKF_FILE>");
            MergeSkaArray arr = MergeSkaArray::load(a.pos[0]);
            std::string s = arr.display(a.has("--full-info"));
            fwrite(s.data(), 1, s.size(), stdout);
        } else if (cmd == "map") {                         // lib.rs:663-706
            Args a = parse(argc, argv, 2, {"-o", "-f", "--format", "--threads"}, {"--ambig-mask", "--repeat-mask"});
            if (a.pos.size() < 2) usage_error("the following required arguments were not provided: <REFERENCE> <INPUT>...");
            const std::string fmt = a.has("-f") ? a.get("-f") : a.get("--format", "aln");
            if (fmt != "aln" && fmt != "vcf") usage_error("invalid value '" + fmt + "' for '--format <FORMAT>'");
            size_t threads = parse_threads(a.get("--threads", "1"));
            std::vector<std::string> input(a.pos.begin() + 1, a.pos.end());
            MergeSkaArray arr = io_utils::load_array(input, threads);
            RefSka ska_ref(arr.kmer_len(), a.pos[0], arr.rc(), a.has("--ambig-mask"), a.has("--repeat-mask"));
            generic_modes::map(arr, ska_ref, a.get("-o"), fmt == "vcf", threads);
        } else if (cmd == "merge") {                       // lib.rs:728-741
            Args a = parse(argc, argv, 2, {"-o", "--threads"}, {});
            if (!a.has("-o")) usage_error("the following required arguments were not provided: -o <OUTPUT>");
            if (a.pos.size() < 2) throw SkaError("Need at least two files to merge");
            MergeSkaArray first = MergeSkaArray::load(a.pos[0]);
            generic_modes::merge(first, std::vector<std::string>(a.pos.begin() + 1, a.pos.end()), a.get("-o"),
                                 (int)std::thread::hardware_concurrency());
        } else if (cmd == "delete") {                      // lib.rs:742-759
            Args a = parse(argc, argv, 2, {"-s", "--skf-file", "-o", "-f"}, {});
            const std::string skf_file = a.has("-s") ? a.get("-s") : a.get("--skf-file");
            if (skf_file.empty()) usage_error("the following required arguments were not provided: --skf-file <SKF_FILE>");
            if (a.has("-f") == !a.pos.empty()) usage_error("exactly one of <NAMES> or -f <FILE_LIST> is required");
            std::vector<std::string> names;
            for (auto& t : io_utils::get_input_list(a.get("-f"), a.pos)) names.push_back(std::get<0>(t));
            MergeSkaArray arr = MergeSkaArray::load(skf_file);
            generic_modes::delete_(arr, names, a.has("-o") ? a.get("-o") : skf_file, (int)std::thread::hardware_concurrency());
        } else if (cmd == "weed") {                        // lib.rs:760-810
            Args a = parse(argc, argv, 2, {"-o", "-m", "--min-freq", "--filter"},
                           {"--reverse", "--filter-ambig-as-missing", "--ambig-mask", "--no-gap-only-sites"});
            if (a.pos.empty() || a.pos.size() > 2) usage_error("usage: ska weed <SKF_FILE> [WEED_FILE]");
            double min_freq = parse_01(a.has("-m") ? a.get("-m") : a.get("--min-freq", "0.9"), "frequency");
            MergeSkaArray arr = MergeSkaArray::load(a.pos[0]);
            generic_modes::weed(arr, a.pos.size() > 1 ? a.pos[1] : "", a.has("--reverse"), min_freq, a.has("--filter-ambig-as-missing"),
                                parse_filter(a.get("--filter", "no-filter")), a.has("--ambig-mask"), a.has("--no-gap-only-sites"),
                                a.has("-o") ? a.get("-o") : a.pos[0], (int)std::thread::hardware_concurrency());
        } else if (cmd == "cov") {                         // lib.rs:828-851
            Args a = parse(argc, argv, 2, {"-k"}, {"--single-strand"});
            if (a.pos.size() != 2) usage_error("the following required arguments were not provided: <FASTQ_FWD> <FASTQ_REV>");
            size_t k = parse_k(a.get("-k", "31"));
            CoverageHistogram cov(a.pos[0], a.pos[1], k, !a.has("--single-strand"), a.has("--verbose"));
            const size_t cutoff = cov.fit_histogram();
            const std::string table = cov.plot_hist();
            fwrite(table.data(), 1, table.size(), stdout);
            fprintf(stderr, "Estimated cutoff\t%zu\n", cutoff);
        } else if (cmd == "fit-selftest") {
            // host-only helper (no device): the reference's unit test of the mixture fit (coverage.rs:369-415) -> prints the cutoff
            std::vector<uint32_t> counts;
            for (int i = 2; i < argc; i++) counts.push_back((uint32_t)strtoul(argv[i], nullptr, 10));
            CoverageHistogram cov(counts);
            const size_t cutoff = cov.fit_histogram();
            printf("cutoff %zu w0 %.6f c %.4f\n", cutoff, cov.w0(), cov.c());
        } else if (cmd == "recode") {
            // host-only helper (no device): decode a .skf and write it again with this codec
            Args a = parse(argc, argv, 2, {"--threads"}, {});
            if (a.pos.size() != 2) usage_error("usage: ska recode <IN.skf> <OUT.skf>");
            skf::save(skf::load(a.pos[0]), a.pos[1], (int)parse_threads(a.get("--threads", "1")));
        } else if (cmd == "codec-selftest") {
            // host-only helper (no device): encode + decode a synthetic array of ROWS x SAMPLES and report the codec's speed
            Args a = parse(argc, argv, 2, {"--threads"}, {});
            if (a.pos.size() != 2) usage_error("usage: ska codec-selftest <ROWS> <SAMPLES>");
            const size_t rows = (size_t)std::stoull(a.pos[0]), n = (size_t)std::stoull(a.pos[1]);
            const int threads = (int)parse_threads(a.get("--threads", "1"));
            skf::Array arr; arr.k = 31; arr.rc = true; arr.n_rows = rows; arr.n_samples = n;
            for (size_t i = 0; i < n; i++) arr.names.push_back("s" + std::to_string(i));
            arr.keys.resize(rows); arr.counts.resize(rows); arr.variants.resize(rows * n);
            uint64_t x = 0x9E3779B97F4A7C15ull;
            auto next = [&] { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
            const char L[4] = {'A', 'C', 'G', 'T'};
            for (size_t r = 0; r < rows; r++) {
                arr.keys[r] = next() >> 5;
                const bool single = next() % 5 != 0; const char c = L[next() & 3];
                for (size_t s = 0; s < n; s++) arr.variants[r * n + s] = single ? '-' : c;
                if (single) arr.variants[r * n + next() % n] = c;
                arr.counts[r] = single ? 1 : n;
            }
            auto t0 = std::chrono::steady_clock::now();
            std::vector<uint8_t> enc = skf::encode(arr, threads);
            auto t1 = std::chrono::steady_clock::now();
            skf::Array back = skf::decode(enc.data(), enc.size(), threads);
            auto t2 = std::chrono::steady_clock::now();
            if (back.variants != arr.variants || back.keys != arr.keys || back.counts != arr.counts || back.names != arr.names)
                throw std::runtime_error("codec self-test: round trip differs");
            printf("rows %zu samples %zu file_bytes %zu encode_s %.3f decode_s %.3f\n", rows, n, enc.size(),
                   std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(t2 - t1).count());
        } else {
            usage_error("unrecognized subcommand '" + cmd + "' (this engine provides build, align, map, distance, merge, delete, weed, nk)");
        }
    } catch (std::exception& e) {
        fprintf(stderr, "thread 'main' panicked:\n%s\n", e.what());
        if (cuda_init.joinable()) cuda_init.join();
        return 101;
    }
    if (cuda_init.joinable()) cuda_init.join();
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    fprintf(stderr, "SKA done in %gs\n", secs);
    fprintf(stderr, "\xE2\xAC\x9B\xE2\xAC\x9C\xE2\xAC\x9B\xE2\xAC\x9C\xE2\xAC\x9B\xE2\xAC\x9C\xE2\xAC\x9B\n\xE2\xAC\x9C\xE2\xAC\x9B\xE2\xAC\x9C\xE2\xAC\x9B\xE2\xAC\x9C\xE2\xAC\x9B\xE2\xAC\x9C\n");
    // every output is written and closed: leave without tearing down gigabytes of device and host buffers one by one
    fflush(stdout); fflush(stderr);
    _exit(0);
}
