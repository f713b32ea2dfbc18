// ska_oracle.cpp -- CPU restatement of the bacpop/ska.rust v0.5.2 hot path.
//
// *** TEST INFRASTRUCTURE, NOT PRODUCT CODE. ***
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
// load this library.  The product path (ska.rust_b200/, libskagpu.so) never links or calls it.
//
// Why a restatement: the reference is pure Rust and neither this container nor the GPU box has
// cargo/rustc, vendored crates, a Cargo.lock or network, so oracle/_ref cannot be built
// (DESIGN.md "Oracle").  PARITY PINNING: this file is checked (tests/test_oracle_golden.py)
// against the reference's own fixtures -- the four prebuilt .skf files, the eight golden
// `ska distance` stdouts, the single-column `ska align` goldens, the filter matrix of
// tests/align.rs and the FASTQ expectations of tests/fastq_input.rs.
//
// Shape: deliberately the *reference's* shape -- a sequential rolling SplitKmer iterator, one
// hash-map probe per base, per-sample dict -> merged dict of Vec<u8> rows -> dense array --
// so that (a) it is an independent formulation from the CUDA path (which is stateless, sort/
// partition based) and (b) it doubles as the "CPU restatement of reference algorithm" baseline
// that bench.py times on the host cores.
//
// Citations are file:line in /root/reference (bacpop/ska.rust @ 99bd6e0).

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <utility>
#include <vector>

typedef unsigned __int128 u128;

namespace ora {

// ------------------------------------------------------------------------------------------------
// bit_encoding.rs
// ------------------------------------------------------------------------------------------------
static const uint8_t LETTER_CODE[4] = {'A', 'C', 'T', 'G'};                      // bit_encoding.rs:30
static inline uint8_t encode_base(uint8_t b) { return (b >> 1) & 0x3; }          // bit_encoding.rs:34-36
static inline uint8_t decode_base(uint8_t e) { return LETTER_CODE[e]; }          // bit_encoding.rs:40-42
static inline uint8_t rc_base(uint8_t e) { return e ^ 2; }                       // bit_encoding.rs:46-48
static inline bool valid_base(uint8_t b) { return (b & 0xF) != 14; }             // bit_encoding.rs:52-54
static inline bool is_ambiguous(uint8_t b) {                                     // bit_encoding.rs:58-61
    b |= 0x20;
    return !(b == 'a' || b == 'c' || b == 'g' || b == 't' || b == 'u' || b == '-');
}
// [p(A), p(C), p(T), p(G)]                                                       bit_encoding.rs:65-85
static void base_to_prob(uint8_t b, double p[4]) {
    const double t = 1.0 / 3.0;
    p[0] = p[1] = p[2] = p[3] = 0.0;
    switch (b) {
        case 'A': p[0] = 1; break;
        case 'C': p[1] = 1; break;
        case 'G': p[3] = 1; break;
        case 'T': case 'U': p[2] = 1; break;
        case 'R': p[0] = .5; p[3] = .5; break;
        case 'Y': p[1] = .5; p[2] = .5; break;
        case 'S': p[1] = .5; p[3] = .5; break;
        case 'W': p[0] = .5; p[2] = .5; break;
        case 'K': p[2] = .5; p[3] = .5; break;
        case 'M': p[0] = .5; p[1] = .5; break;
        case 'B': p[1] = t; p[2] = t; p[3] = t; break;
        case 'D': p[0] = t; p[2] = t; p[3] = t; break;
        case 'H': p[0] = t; p[1] = t; p[2] = t; break;
        case 'V': p[0] = t; p[1] = t; p[3] = t; break;
        default: break;  // 'N' and anything else: all zero (bit_encoding.rs:82-83)
    }
}

// IUPAC[new_base*256 + existing] (bit_encoding.rs:388-453).  The reference stores a literal
// 1024-entry table; its documented rule (comment block :354-369) is "IUPAC code of
// set(existing) U {new}", for existing in upper or lower case, 0 for every other byte.  The
// table is regenerated here from that rule (and cross-checked against the .skf fixtures that
// contain merged codes: dup_ss/dup_rc goldens K,D,M,H and multidist.skf).
struct IupacTable {
    uint8_t t[1024];
    IupacTable() {
        memset(t, 0, sizeof t);
        // bit0=A bit1=C bit2=T bit3=G  (bit index == 2-bit code)
        const char* letters = "ACMTWYHGRSVKDBN";  // index = mask-1
        auto mask_of = [&](char c) -> int {
            for (int m = 1; m < 16; m++) if (letters[m - 1] == c) return m;
            return 0;
        };
        for (int nb = 0; nb < 4; nb++)
            for (int e = 0; e < 256; e++) {
                int up = (e >= 'a' && e <= 'z') ? e - 32 : e;
                if (up == 'U') continue;  // 'U' rows are 0 in the reference table (never stored)
                int m = mask_of((char)up);
                if (!m || !((e >= 'A' && e <= 'Z') || (e >= 'a' && e <= 'z'))) continue;
                t[nb * 256 + e] = (uint8_t)letters[(m | (1 << nb)) - 1];
            }
    }
};
static const IupacTable IUPAC;

template <class IntT> struct UIntOps;
template <> struct UIntOps<uint64_t> {
    static const unsigned BITS = 64;
    // bit_encoding.rs:182-195
    static uint64_t rev_comp(uint64_t x, size_t k_size) {
        x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
        x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
        x = ((x >> 8) & 0x00FF00FF00FF00FFULL) | ((x & 0x00FF00FF00FF00FFULL) << 8);
        x = ((x >> 16) & 0x0000FFFF0000FFFFULL) | ((x & 0x0000FFFF0000FFFFULL) << 16);
        x = ((x >> 32) & 0x00000000FFFFFFFFULL) | ((x & 0x00000000FFFFFFFFULL) << 32);
        x ^= 0xAAAAAAAAAAAAAAAAULL;
        return x >> (2 * (32 - k_size));
    }
};
template <> struct UIntOps<u128> {
    static const unsigned BITS = 128;
    // bit_encoding.rs:241-261
    static u128 rev_comp(u128 x, size_t k_size) {
        auto rep = [](uint64_t v) { return ((u128)v << 64) | v; };
        x = ((x >> 2) & rep(0x3333333333333333ULL)) | ((x & rep(0x3333333333333333ULL)) << 2);
        x = ((x >> 4) & rep(0x0F0F0F0F0F0F0F0FULL)) | ((x & rep(0x0F0F0F0F0F0F0F0FULL)) << 4);
        x = ((x >> 8) & rep(0x00FF00FF00FF00FFULL)) | ((x & rep(0x00FF00FF00FF00FFULL)) << 8);
        x = ((x >> 16) & rep(0x0000FFFF0000FFFFULL)) | ((x & rep(0x0000FFFF0000FFFFULL)) << 16);
        x = ((x >> 32) & rep(0x00000000FFFFFFFFULL)) | ((x & rep(0x00000000FFFFFFFFULL)) << 32);
        x = ((x >> 64) & (u128)0xFFFFFFFFFFFFFFFFULL) | ((x & (u128)0xFFFFFFFFFFFFFFFFULL) << 64);
        x ^= rep(0xAAAAAAAAAAAAAAAAULL);
        return x >> (2 * (64 - k_size));
    }
};
// bit_encoding.rs:208-213 / :274-279
template <class IntT> static void generate_masks(size_t k, IntT& lower_mask, IntT& upper_mask) {
    size_t half = (k - 1) / 2;
    lower_mask = (((IntT)1) << (half * 2)) - 1;
    upper_mask = lower_mask << (half * 2);
}

// ------------------------------------------------------------------------------------------------
// nthash.rs
// ------------------------------------------------------------------------------------------------
static const uint64_t HASH_LOOKUP[4] = {0x3c8bfbb395c60474ULL, 0x3193c18562a02b4cULL,     // nthash.rs:12-17
                                        0x295549f54be24456ULL, 0x20323ed082572324ULL};
static const uint64_t RC_HASH_LOOKUP[4] = {0x295549f54be24456ULL, 0x20323ed082572324ULL,  // nthash.rs:18-23
                                           0x3c8bfbb395c60474ULL, 0x3193c18562a02b4cULL};
static inline uint64_t rotl(uint64_t x, unsigned r) { r &= 63; return r ? (x << r) | (x >> (64 - r)) : x; }
static inline uint64_t rotr(uint64_t x, unsigned r) { r &= 63; return r ? (x >> r) | (x << (64 - r)) : x; }

struct NtHashIterator {
    size_t k; uint64_t fh; bool has_rh; uint64_t rh;
    NtHashIterator() : k(0), fh(0), has_rh(false), rh(0) {}
    // nthash.rs:35-52
    NtHashIterator(const uint8_t* seq, size_t k_, bool rc) : k(k_), fh(0), has_rh(rc), rh(0) {
        for (size_t i = 0; i < k; i++) fh ^= rotl(HASH_LOOKUP[encode_base(seq[i])], (unsigned)(k - i - 1));
        if (rc)
            for (size_t i = 0; i < k; i++)  // seq[0..k].iter().rev().enumerate()
                rh ^= rotl(RC_HASH_LOOKUP[encode_base(seq[k - 1 - i])], (unsigned)(k - i - 1));
    }
    // nthash.rs:55-67
    void roll_fwd(uint8_t old_base, uint8_t new_base) {
        fh = rotl(fh, 1) ^ rotl(HASH_LOOKUP[old_base], (unsigned)k) ^ HASH_LOOKUP[new_base];
        if (has_rh) rh = rotr(rh, 1) ^ rotr(RC_HASH_LOOKUP[old_base], 1) ^ rotl(RC_HASH_LOOKUP[new_base], (unsigned)k - 1);
    }
    uint64_t curr_hash() const { return has_rh ? std::min(fh, rh) : fh; }  // nthash.rs:70-76
};

enum QualFilter { QF_NONE = 0, QF_MIDDLE = 1, QF_STRICT = 2 };  // cli.rs order: NoFilter, Middle, Strict

// ------------------------------------------------------------------------------------------------
// split_kmer.rs
// ------------------------------------------------------------------------------------------------
template <class IntT> struct SplitKmer {
    size_t k; IntT upper_mask, lower_mask;
    const uint8_t* seq; size_t seq_len; const uint8_t* qual;
    int qual_filter; uint8_t min_qual; size_t index;
    IntT upper, lower; uint8_t middle_base; bool rc;
    IntT rc_upper, rc_lower; uint8_t rc_middle_base;
    bool is_reads; NtHashIterator hash_gen;

    // split_kmer.rs:66-71   (strict '>')
    static bool valid_qual(size_t idx, const uint8_t* qual, uint8_t min_qual) {
        return qual ? (uint8_t)(qual[idx] - 33) > min_qual : true;
    }
    // split_kmer.rs:78-140
    static bool build(const uint8_t* seq, size_t seq_len, const uint8_t* qual, size_t k, size_t& idx, int qual_filter,
                      uint8_t min_qual, bool is_reads, bool rc, IntT& upper, IntT& lower, uint8_t& middle_base,
                      NtHashIterator& hash_gen) {
        if (idx + k >= seq_len) return false;
        upper = 0; lower = 0; middle_base = 0;
        size_t middle_idx = (k + 1) / 2 - 1;
        size_t i = 0;
        while (i < k) {
            if (valid_base(seq[i + idx]) && (qual_filter != QF_STRICT || valid_qual(i + idx, qual, min_qual))) {
                uint8_t next_base = encode_base(seq[i + idx]);
                if (i > middle_idx) { lower <<= 2; lower |= (IntT)next_base; }
                else if (i < middle_idx) { upper <<= 2; upper |= ((IntT)next_base) << (middle_idx * 2); }
                else middle_base = next_base;
                i++;
            } else {
                idx += i + 1;
                if (idx + k >= seq_len) return false;
                upper = 0; lower = 0; middle_base = 0; i = 0;
            }
        }
        if (is_reads) hash_gen = NtHashIterator(seq + idx, k, rc);
        idx += k - 1;
        return true;
    }
    bool self_palindrome() const { return rc && upper == rc_upper && lower == rc_lower; }  // :144-146
    void update_rc() {                                                                     // :149-153
        rc_upper = UIntOps<IntT>::rev_comp(lower, k - 1) & upper_mask;
        rc_middle_base = rc_base(middle_base);
        rc_lower = UIntOps<IntT>::rev_comp(upper, k - 1) & lower_mask;
    }
    // split_kmer.rs:159-217
    bool roll_fwd() {
        index += 1;
        if (index >= seq_len) return false;
        uint8_t base = seq[index];
        if (!valid_base(base) || (qual_filter == QF_STRICT && !valid_qual(index, qual, min_qual))) {
            IntT u, l; uint8_t m; NtHashIterator h;
            if (!build(seq, seq_len, qual, k, index, qual_filter, min_qual, is_reads, rc, u, l, m, h)) return false;
            upper = u; lower = l; middle_base = m; hash_gen = h;
            if (rc) update_rc();
            return true;
        }
        size_t half_k = (k - 1) / 2;
        uint8_t new_base = encode_base(base);
        if (is_reads) {
            uint8_t old_base = (uint8_t)(upper >> ((k - 2) * 2));
            hash_gen.roll_fwd(old_base, new_base);
        }
        upper = ((upper << 2) | (((IntT)middle_base) << (half_k * 2))) & upper_mask;
        middle_base = (uint8_t)(lower >> (2 * (half_k - 1)));
        lower = ((lower << 2) | (IntT)new_base) & lower_mask;
        if (rc) {
            rc_lower = ((rc_lower >> 2) | (((IntT)rc_middle_base) << (2 * (half_k - 1)))) & lower_mask;
            rc_middle_base = rc_base(middle_base);
            rc_upper = ((rc_upper >> 2) | (((IntT)rc_base(new_base)) << (2 * ((half_k * 2) - 1)))) & upper_mask;
        }
        return true;
    }
    // split_kmer.rs:226-276 ; returns false for None
    bool init(const uint8_t* seq_, size_t seq_len_, const uint8_t* qual_, size_t k_, bool rc_, uint8_t min_qual_,
              int qual_filter_, bool is_reads_) {
        k = k_; seq = seq_; seq_len = seq_len_; qual = qual_; rc = rc_; min_qual = min_qual_;
        qual_filter = qual_filter_; is_reads = is_reads_;
        index = 0;
        if (!build(seq, seq_len, qual, k, index, qual_filter, min_qual, is_reads, rc, upper, lower, middle_base, hash_gen))
            return false;
        generate_masks<IntT>(k, lower_mask, upper_mask);
        rc_upper = 0; rc_lower = 0; rc_middle_base = 0;
        if (rc) update_rc();
        return true;
    }
    // split_kmer.rs:281-295
    void get_curr_kmer(IntT& kmer, uint8_t& base, bool& is_rc) const {
        IntT split_kmer = upper | lower;
        if (rc) {
            IntT rc_split = rc_upper | rc_lower;
            if (split_kmer > rc_split) { kmer = rc_split; base = rc_middle_base; is_rc = true; return; }
        }
        kmer = split_kmer; base = middle_base; is_rc = false;
    }
    uint64_t get_hash() const { return hash_gen.curr_hash(); }               // :302-307
    size_t get_middle_pos() const { return index - ((k + 1) / 2 - 1); }      // :322-325
    bool middle_base_qual() const {                                          // :328-339
        if (qual && (qual_filter == QF_MIDDLE || qual_filter == QF_STRICT))
            return valid_qual(get_middle_pos(), qual, min_qual);
        return true;
    }
};

// ------------------------------------------------------------------------------------------------
// A small open-addressing map (stand-in for hashbrown: SwissTable-like load <= 87.5 %, growth by
// doubling, 64-bit multiplicative mixing).  Only map *semantics* matter for parity.
// ------------------------------------------------------------------------------------------------
static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 32; x *= 0xd6e8feb86659fd93ULL; x ^= x >> 32; x *= 0xd6e8feb86659fd93ULL; x ^= x >> 32;
    return x;
}
static inline uint64_t hash_key(uint64_t k) { return mix64(k); }
static inline uint64_t hash_key(u128 k) { return mix64((uint64_t)k ^ mix64((uint64_t)(k >> 64))); }

// Each map instance salts its hash (as hashbrown does through ahash::RandomState): iterating one
// map in slot order while inserting into another with the SAME hash function builds giant probe
// clusters (quadratic time) -- exactly what MergeSkaDict::append does.
static uint64_t next_salt() {
    static std::atomic<uint64_t> ctr{0x243F6A8885A308D3ULL};
    return ctr.fetch_add(0x9E3779B97F4A7C15ULL);
}
template <class K, class V> struct FlatMap {
    struct Entry { K key; V val; bool used; };          // one cache line per probe, like a SwissTable bucket
    std::vector<Entry> e;
    size_t n = 0, cap = 0;
    uint64_t salt = next_salt();
    void reserve_pow2(size_t c) { cap = c; e.assign(c, Entry{K(), V(), false}); n = 0; }
    size_t size() const { return n; }
    void grow() {
        FlatMap bigger; bigger.salt = salt; bigger.reserve_pow2(cap ? cap * 2 : 16);
        for (size_t i = 0; i < cap; i++) if (e[i].used) *bigger.slot(e[i].key) = std::move(e[i].val);
        *this = std::move(bigger);
    }
    // returns pointer to value; *inserted tells whether key was new
    V* slot(const K& k, bool* inserted = nullptr) {
        if (cap == 0 || (n + 1) * 8 > cap * 7) grow();
        size_t m = cap - 1, i = hash_key(k ^ (K)salt) & m;
        while (e[i].used) {
            if (e[i].key == k) { if (inserted) *inserted = false; return &e[i].val; }
            i = (i + 1) & m;
        }
        e[i].used = true; e[i].key = k; n++;
        if (inserted) *inserted = true;
        return &e[i].val;
    }
    // iteration in slot order (the analogue of hashbrown's arbitrary iteration order)
    template <class F> void for_each(F f) { for (size_t i = 0; i < cap; i++) if (e[i].used) f(e[i].key, e[i].val); }
};

// ------------------------------------------------------------------------------------------------
// bloom_filter.rs
// ------------------------------------------------------------------------------------------------
struct KmerFilter {
    uint64_t buf_size; std::vector<uint64_t> buffer; FlatMap<uint64_t, uint16_t> counts; uint16_t min_count;
    explicit KmerFilter(uint16_t mc = 0) : min_count(mc) {
        // BLOOM_WIDTH = 1<<27, BITS_PER_ENTRY = 12   (bloom_filter.rs:15-20, :93-96)
        buf_size = (uint64_t)std::llround((double)(1ULL << 27) * (12.0 / 8.0) / 64.0);
    }
    void init() { if (buffer.empty()) buffer.assign(buf_size, 0); }                              // :107-111
    static uint64_t reduce(uint64_t key, uint64_t range) { return (uint64_t)(((u128)key * (u128)range) >> 64); }  // :50-52
    static uint64_t cheap_mix(uint64_t key) { return (key ^ (key >> 31)) * 0x85D059AA333121CFULL; }               // :56-58
    static uint64_t fingerprint(uint64_t key) {                                                  // :62-68
        return 1ULL << (key & 63) | 1ULL << ((key >> 6) & 63) | 1ULL << ((key >> 12) & 63) | 1ULL << ((key >> 18) & 63) |
               1ULL << ((key >> 24) & 63);
    }
    bool bloom_add_and_check(uint64_t key) {                                                     // :77-86
        uint64_t f = fingerprint(key);
        uint64_t& v = buffer[reduce(cheap_mix(key), buf_size)];
        if ((v & f) == f) return true;
        v |= f;
        return false;
    }
    // returns true iff Ordering::Equal                                                           :116-148
    bool filter(uint64_t kmer_hash) {
        if (min_count <= 1) return true;
        if (min_count == 2) return bloom_add_and_check(kmer_hash);
        if (bloom_add_and_check(kmer_hash)) {
            bool inserted; uint16_t* c = counts.slot(kmer_hash, &inserted);
            uint16_t count = 2;
            if (!inserted) { count = (*c == 65535) ? 65535 : (uint16_t)(*c + 1); }
            *c = count;
            return min_count == count;
        }
        return false;
    }
};

// ------------------------------------------------------------------------------------------------
// ska_dict.rs
// ------------------------------------------------------------------------------------------------
struct SampleInput {            // one sample = file 1 (+ optional file 2), records already parsed on the host
    const uint8_t* seq;         // concatenated record sequences (newlines stripped), file 1 then file 2
    const uint8_t* qual;        // same layout, or nullptr for FASTA
    const uint64_t* rec_off;    // n_rec + 1 offsets into seq/qual
    uint64_t n_rec;
    uint64_t n_rec_file1;       // records [0, n_rec_file1) are file 1; the rest file 2
};

template <class IntT> struct SkaDict {
    size_t k; bool rc; size_t sample_idx;
    FlatMap<IntT, uint8_t> split_kmers;
    KmerFilter kmer_filter;

    void add_to_dict(IntT kmer, uint8_t base) {                                  // ska_dict.rs:76-81
        bool ins; uint8_t* b = split_kmers.slot(kmer, &ins);
        if (ins) *b = decode_base(base); else *b = IUPAC.t[(size_t)base * 256 + *b];
    }
    void add_palindrome_to_dict(IntT kmer, uint8_t base) {                       // ska_dict.rs:85-113
        bool ins; uint8_t* b = split_kmers.slot(kmer, &ins);
        bool at = (base == 0 || base == 2);
        if (ins) { *b = at ? 'W' : 'S'; return; }
        switch (*b) {
            case 'W': *b = at ? 'W' : 'N'; break;
            case 'S': *b = at ? 'N' : 'S'; break;
            case 'N': break;
            default: throw std::runtime_error("Palindrome middle base not W/S");
        }
    }
    // ska_dict.rs:118-180 (record loop of one file)
    void add_file_kmers(const SampleInput& in, uint64_t rec_begin, uint64_t rec_end, bool is_reads, int qual_filter,
                        uint8_t min_qual, double proportion_reads) {
        size_t step = 1;
        if (proportion_reads > 0) step = (size_t)std::llround(1.0 / proportion_reads);
        if (step == 0) step = 1;  // (1/p).round() as usize of a p>2 would be 0 -> Rust panics on %0; never used
        size_t iter_reads = 0;
        for (uint64_t r = rec_begin; r < rec_end; r++) {
            if (iter_reads % step != 0) { iter_reads++; continue; }
            iter_reads++;
            const uint8_t* s = in.seq + in.rec_off[r];
            size_t len = (size_t)(in.rec_off[r + 1] - in.rec_off[r]);
            const uint8_t* q = in.qual ? in.qual + in.rec_off[r] : nullptr;
            SplitKmer<IntT> it;
            if (!it.init(s, len, q, k, rc, min_qual, qual_filter, is_reads)) continue;
            do {
                if (!is_reads || (it.middle_base_qual() && kmer_filter.filter(it.get_hash()))) {
                    IntT kmer; uint8_t base; bool isrc;
                    it.get_curr_kmer(kmer, base, isrc);
                    if (it.self_palindrome()) add_palindrome_to_dict(kmer, base); else add_to_dict(kmer, base);
                }
            } while (it.roll_fwd());
        }
    }
    // ska_dict.rs:333-378
    void build(size_t k_, size_t sample_idx_, const SampleInput& in, bool rc_, uint16_t min_count, uint8_t min_qual,
               int qual_filter, double proportion_reads) {
        if (k_ < 5 || k_ > 63 || k_ % 2 == 0) throw std::runtime_error("Invalid k-mer length");
        k = k_; rc = rc_; sample_idx = sample_idx_;
        kmer_filter = KmerFilter(min_count);
        bool is_reads = in.qual != nullptr;  // first record of file 1 is FASTQ (ska_dict.rs:356-366)
        if (is_reads) kmer_filter.init();
        add_file_kmers(in, 0, in.n_rec_file1, is_reads, qual_filter, min_qual, proportion_reads);
        if (in.n_rec_file1 < in.n_rec)
            add_file_kmers(in, in.n_rec_file1, in.n_rec, is_reads, qual_filter, min_qual, proportion_reads);
        if (split_kmers.size() == 0) throw std::runtime_error("sample has no valid sequence");  // :374-376
    }
};

// ------------------------------------------------------------------------------------------------
// merge_ska_dict.rs
// ------------------------------------------------------------------------------------------------
template <class IntT> struct MergeSkaDict {
    size_t k = 0; bool rc = false; size_t n_samples = 0;
    FlatMap<IntT, std::vector<uint8_t>> split_kmers;
    // merge_ska_dict.rs:77-109
    void append(SkaDict<IntT>& other) {
        if (other.k != k) throw std::runtime_error("K-mer lengths do not match");
        if (other.rc != rc) throw std::runtime_error("Strand use inconsistent");
        other.split_kmers.for_each([&](const IntT& key, uint8_t& base) {
            bool ins; std::vector<uint8_t>* row = split_kmers.slot(key, &ins);
            if (ins) row->assign(n_samples, 0);
            (*row)[other.sample_idx] = base;
        });
    }
    // merge_ska_dict.rs:119-151
    void merge(MergeSkaDict& other) {
        if (other.split_kmers.size() == 0) return;
        if (other.k != k) throw std::runtime_error("K-mer lengths do not match");
        if (other.rc != rc) throw std::runtime_error("Strand use inconsistent");
        if (split_kmers.size() == 0) { split_kmers = std::move(other.split_kmers); return; }
        other.split_kmers.for_each([&](const IntT& key, std::vector<uint8_t>& o) {
            bool ins; std::vector<uint8_t>* row = split_kmers.slot(key, &ins);
            if (ins) *row = std::move(o);
            else for (size_t j = 0; j < n_samples; j++) (*row)[j] |= o[j];
        });
    }
};

struct BuildParams { size_t k; bool rc; uint16_t min_count; uint8_t min_qual; int qual_filter; double proportion_reads; };

// merge_ska_dict.rs:230-256 multi_append
template <class IntT>
static void multi_append(MergeSkaDict<IntT>& d, const SampleInput* in, size_t offset, size_t n, size_t total,
                         const BuildParams& p) {
    d.k = p.k; d.rc = p.rc; d.n_samples = total;
    for (size_t i = 0; i < n; i++) {
        SkaDict<IntT> s;
        s.build(p.k, offset + i, in[i], p.rc, p.min_count, p.min_qual, p.qual_filter, p.proportion_reads);
        d.append(s);
    }
}
// merge_ska_dict.rs:264-326 parallel_append (rayon::join -> std::thread for one branch)
template <class IntT>
static void parallel_append(MergeSkaDict<IntT>& out, size_t depth, size_t offset, const SampleInput* in, size_t n,
                            size_t total, const BuildParams& p, std::string* err) {
    if (depth == 0) { try { multi_append(out, in, offset, n, total, p); } catch (std::exception& e) { *err = e.what(); } return; }
    size_t split = n / 2;
    MergeSkaDict<IntT> b; std::string err_b;
    std::thread t([&] { parallel_append(b, depth - 1, offset + split, in + split, n - split, total, p, &err_b); });
    parallel_append(out, depth - 1, offset, in, split, total, p, err);
    t.join();
    if (!err_b.empty()) { *err = err_b; return; }
    if (!err->empty()) return;
    out.merge(b);
}

// ------------------------------------------------------------------------------------------------
// merge_ska_array.rs
// ------------------------------------------------------------------------------------------------
enum FilterType { NoFilter = 0, NoConst = 1, NoAmbig = 2, NoAmbigOrConst = 3 };  // cli.rs:129-138

struct VariantDist { double distance, mismatch_prop; uint64_t match_count, mismatch_count; };

struct Array {          // key type erased to u128 so one handle type serves k<=31 and k>31
    size_t k = 0; bool rc = false; unsigned key_bits = 64; size_t n_samples = 0;
    std::vector<u128> split_kmers; std::vector<uint8_t> variants; std::vector<uint64_t> variant_count;
    size_t n_rows() const { return n_samples ? variants.size() / n_samples : 0; }

    // merge_ska_array.rs:166-186 (rows emitted in ascending key order: canonical order)
    template <class IntT> void from_dict(MergeSkaDict<IntT>& d) {
        k = d.k; rc = d.rc; n_samples = d.n_samples; key_bits = UIntOps<IntT>::BITS;
        std::vector<std::pair<IntT, std::vector<uint8_t>*>> order;
        order.reserve(d.split_kmers.size());
        d.split_kmers.for_each([&](const IntT& key, std::vector<uint8_t>& row) { order.push_back({key, &row}); });
        std::sort(order.begin(), order.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
        split_kmers.reserve(order.size()); variants.reserve(order.size() * n_samples);
        for (auto& kv : order) {
            std::vector<uint8_t>& bases = *kv.second;
            split_kmers.push_back((u128)kv.first);
            uint64_t c = 0;
            for (uint8_t b : bases) if (b != 0 && b != '-') c++;
            variant_count.push_back(c);
            for (uint8_t b : bases) variants.push_back(std::max<uint8_t>(b, '-'));  // zeros -> missing (:175)
        }
    }
    // merge_ska_array.rs:139-163
    void update_counts(bool filter_ambig_as_missing) {
        std::vector<u128> nk; std::vector<uint8_t> nv; std::vector<uint64_t> nc;
        size_t R = n_rows();
        for (size_t r = 0; r < R; r++) {
            const uint8_t* row = &variants[r * n_samples]; uint64_t c = 0;
            for (size_t s = 0; s < n_samples; s++)
                if (row[s] != '-' && (!filter_ambig_as_missing || !is_ambiguous(row[s]))) c++;
            if (c > 0) { nc.push_back(c); nk.push_back(split_kmers[r]); nv.insert(nv.end(), row, row + n_samples); }
        }
        split_kmers.swap(nk); variants.swap(nv); variant_count.swap(nc);
    }
    // merge_ska_array.rs:289-402
    int32_t filter(size_t min_count, bool filter_ambig_as_missing, int ftype, bool mask_ambig, bool ignore_const_gaps,
                   bool update_kmers) {
        if (filter_ambig_as_missing) update_counts(true);
        std::vector<u128> fk; std::vector<uint8_t> fv; std::vector<uint64_t> fc;
        int32_t removed = 0; size_t R = n_rows();
        for (size_t r = 0; r < R; r++) {
            const uint8_t* row = &variants[r * n_samples];
            bool keep = false;
            if (variant_count[r] >= min_count) {
                switch (ftype) {
                    case NoFilter: keep = true; break;
                    case NoConst: {
                        bool seen[256] = {false}; int types = 0;
                        for (size_t s = 0; s < n_samples && types <= 1; s++)
                            if (!ignore_const_gaps || row[s] != '-') if (!seen[row[s]]) { seen[row[s]] = true; types++; }
                        keep = types > 1; break;
                    }
                    case NoAmbig: {
                        keep = true;
                        for (size_t s = 0; s < n_samples; s++) if (is_ambiguous(row[s])) { keep = false; break; }
                        break;
                    }
                    case NoAmbigOrConst: {
                        bool seen[256] = {false}; int count = 0;
                        for (size_t s = 0; s < n_samples; s++) seen[row[s]] = true;
                        for (int b = 0; b < 256; b++) {
                            if (!seen[b]) continue;
                            int lb = b | 0x20;
                            if (lb == 'a' || lb == 'c' || lb == 'g' || lb == 't' || lb == 'u') count++;
                            else if (lb == '-') count += ignore_const_gaps ? 0 : 1;
                        }
                        keep = count > 1; break;
                    }
                    default: throw std::runtime_error("bad filter type");
                }
            }
            if (keep) {
                fv.insert(fv.end(), row, row + n_samples); fc.push_back(variant_count[r]);
                if (update_kmers) fk.push_back(split_kmers[r]);
            } else removed++;
        }
        variants.swap(fv); variant_count.swap(fc);
        if (update_kmers) split_kmers.swap(fk);
        if (mask_ambig) for (auto& v : variants) if (is_ambiguous(v)) v = 'N';
        return removed;
    }
    // merge_ska_array.rs:587-632
    VariantDist variant_dist(size_t i, size_t j, double constant, bool filt_ambig) const {
        double distance = 0, mismatches = 0, matches = constant; size_t R = n_rows();
        for (size_t r = 0; r < R; r++) {
            uint8_t v1 = variants[r * n_samples + i], v2 = variants[r * n_samples + j];
            if (v1 == '-' || v2 == '-') { if (!(v1 == '-' && v2 == '-')) mismatches += 1.0; }
            else if (filt_ambig) {
                if (!is_ambiguous(v1) && !is_ambiguous(v2)) { matches += 1.0; if (v1 != v2) distance += 1.0; }
            } else {
                double p1[4], p2[4]; base_to_prob(v1, p1); base_to_prob(v2, p2);
                double overlap = 0; for (int q = 0; q < 4; q++) overlap += p1[q] * p2[q];
                if (overlap > 0.0) matches += 1.0;
                distance += 1.0 - overlap;
            }
        }
        VariantDist d;
        d.distance = distance;
        d.mismatch_prop = (matches + mismatches) == 0.0 ? 0.0 : mismatches / (matches + mismatches);
        d.match_count = (uint64_t)matches; d.mismatch_count = (uint64_t)mismatches;
        return d;
    }
    // merge_ska_array.rs:416-438 ; column-parallel over i on `threads` host threads
    void distance(double constant, bool filt_ambig, VariantDist* out, int threads) const {
        size_t N = n_samples;
        auto work = [&](size_t t, size_t T) {
            for (size_t i = t; i < N; i += T) {
                size_t base = i * N - i * (i + 1) / 2;  // index of pair (i, i+1)
                for (size_t j = i + 1; j < N; j++) out[base + (j - i - 1)] = variant_dist(i, j, constant, filt_ambig);
            }
        };
        if (threads <= 1) { work(0, 1); return; }
        std::vector<std::thread> th;
        for (int t = 0; t < threads; t++) th.emplace_back(work, (size_t)t, (size_t)threads);
        for (auto& t : th) t.join();
    }
};

template <class IntT>
static void build_impl(Array& out, const SampleInput* in, size_t n, const BuildParams& p, size_t threads) {
    // merge_ska_dict.rs:381-400: max_threads = max(1, min(threads, 1 + n/10)); depth = floor(log2)
    size_t max_threads = std::max<size_t>(1, std::min<size_t>(threads, 1 + n / 10));
    size_t depth = 0; while ((2ULL << depth) <= max_threads) depth++;
    MergeSkaDict<IntT> merged; merged.k = p.k; merged.rc = p.rc; merged.n_samples = n;
    std::string err;
    if (p.k < 5 || p.k > 63 || p.k % 2 == 0) throw std::runtime_error("Invalid k-mer length");
    parallel_append(merged, depth, 0, in, n, n, p, &err);
    if (!err.empty()) throw std::runtime_error(err);
    out.from_dict(merged);
}

}  // namespace ora

// =================================================================================================
// C ABI (ctypes).  Names are ora_*; nothing here is exported by the product library.
// =================================================================================================
static thread_local std::string g_err;
#define ORA_TRY try {
#define ORA_CATCH(rv) } catch (std::exception & e) { g_err = e.what(); return rv; }

extern "C" {

const char* ora_last_error(void) { return g_err.c_str(); }

struct ora_sample_t {
    const uint8_t* seq; const uint8_t* qual; const uint64_t* rec_off; uint64_t n_rec; uint64_t n_rec_file1;
};

// build_and_merge + MergeSkaArray::new  (merge_ska_dict.rs:354-417, merge_ska_array.rs:166-186)
void* ora_build(const ora_sample_t* samples, uint64_t n, uint32_t k, int rc, uint16_t min_count, uint8_t min_qual,
                int qual_filter, double proportion_reads, int threads) {
    ORA_TRY
    std::vector<ora::SampleInput> in(n);
    for (uint64_t i = 0; i < n; i++) in[i] = {samples[i].seq, samples[i].qual, samples[i].rec_off, samples[i].n_rec, samples[i].n_rec_file1};
    ora::BuildParams p{k, rc != 0, min_count, min_qual, qual_filter, proportion_reads};
    std::unique_ptr<ora::Array> a(new ora::Array());
    if (k <= 31) ora::build_impl<uint64_t>(*a, in.data(), n, p, (size_t)std::max(1, threads));   // lib.rs:592
    else ora::build_impl<u128>(*a, in.data(), n, p, (size_t)std::max(1, threads));
    return a.release();
    ORA_CATCH(nullptr)
}

// keys as (lo, hi) u64 pairs
void* ora_from_host(uint32_t k, int rc, uint32_t key_bits, uint64_t n_rows, uint64_t n_samples, const uint64_t* keys_lohi,
                    const uint8_t* variants, const uint64_t* counts) {
    ORA_TRY
    std::unique_ptr<ora::Array> a(new ora::Array());
    a->k = k; a->rc = rc != 0; a->key_bits = key_bits; a->n_samples = n_samples;
    a->split_kmers.resize(n_rows);
    for (uint64_t r = 0; r < n_rows; r++) a->split_kmers[r] = ((u128)keys_lohi[2 * r + 1] << 64) | keys_lohi[2 * r];
    a->variants.assign(variants, variants + n_rows * n_samples);
    a->variant_count.assign(counts, counts + n_rows);
    return a.release();
    ORA_CATCH(nullptr)
}
void ora_dims(void* h, uint64_t* n_rows, uint64_t* n_keys, uint64_t* n_samples, uint32_t* k, int* rc, uint32_t* key_bits) {
    ora::Array* a = (ora::Array*)h;
    *n_rows = a->n_rows(); *n_keys = a->split_kmers.size(); *n_samples = a->n_samples; *k = (uint32_t)a->k; *rc = a->rc; *key_bits = a->key_bits;
}
void ora_export(void* h, uint64_t* keys_lohi, uint8_t* variants, uint64_t* counts) {
    ora::Array* a = (ora::Array*)h;
    if (keys_lohi) for (size_t r = 0; r < a->split_kmers.size(); r++) { keys_lohi[2 * r] = (uint64_t)a->split_kmers[r]; keys_lohi[2 * r + 1] = (uint64_t)(a->split_kmers[r] >> 64); }
    if (variants) memcpy(variants, a->variants.data(), a->variants.size());
    if (counts) memcpy(counts, a->variant_count.data(), a->variant_count.size() * 8);
}
// write_fasta's transpose (merge_ska_array.rs:499-503): (N, R') row-major
void ora_export_transposed(void* h, uint8_t* out) {
    ora::Array* a = (ora::Array*)h; size_t R = a->n_rows(), N = a->n_samples;
    for (size_t r = 0; r < R; r++) for (size_t s = 0; s < N; s++) out[s * R + r] = a->variants[r * N + s];
}
int ora_filter(void* h, uint64_t min_count, int filter_ambig_as_missing, int filter_type, int mask_ambig, int ignore_const_gaps,
               int update_kmers, int32_t* removed) {
    ORA_TRY
    *removed = ((ora::Array*)h)->filter(min_count, filter_ambig_as_missing, filter_type, mask_ambig, ignore_const_gaps, update_kmers);
    return 0;
    ORA_CATCH(1)
}
int ora_distance(void* h, double constant, int filt_ambig, ora::VariantDist* out, int threads) {
    ORA_TRY
    ((ora::Array*)h)->distance(constant, filt_ambig != 0, out, threads);
    return 0;
    ORA_CATCH(1)
}
void ora_free(void* h) { delete (ora::Array*)h; }

// Enumerate the windows of ONE record exactly as the rolling iterator yields them
// (split_kmer.rs:226-316 driven as in ska_dict.rs:144-178).  Output arrays sized >= seq_len.
// key as (lo,hi); returns number of windows.  hash only meaningful if is_reads.
uint64_t ora_windows(const uint8_t* seq, uint64_t seq_len, const uint8_t* qual, uint32_t k, int rc, uint8_t min_qual,
                     int qual_filter, int is_reads, uint64_t* keys_lohi, uint8_t* mid, uint8_t* is_rc, uint8_t* is_pal,
                     uint64_t* mid_pos, uint8_t* mid_ok, uint64_t* hash) {
    uint64_t n = 0;
    auto run = [&](auto tag) {
        typedef decltype(tag) IntT;
        ora::SplitKmer<IntT> it;
        if (!it.init(seq, seq_len, qual, k, rc != 0, min_qual, qual_filter, is_reads != 0)) return;
        do {
            IntT kmer; uint8_t base; bool r;
            it.get_curr_kmer(kmer, base, r);
            keys_lohi[2 * n] = (uint64_t)kmer; keys_lohi[2 * n + 1] = (uint64_t)((u128)kmer >> 64);
            mid[n] = base; is_rc[n] = r; is_pal[n] = it.self_palindrome(); mid_pos[n] = it.get_middle_pos();
            mid_ok[n] = it.middle_base_qual(); hash[n] = is_reads ? it.get_hash() : 0;
            n++;
        } while (it.roll_fwd());
    };
    if (k <= 31) run((uint64_t)0); else run((u128)0);
    return n;
}

// IUPAC merge of one more base into an existing code (exposed so tests can pin the table)
uint8_t ora_iupac(uint8_t new_base_2bit, uint8_t existing) { return ora::IUPAC.t[(size_t)(new_base_2bit & 3) * 256 + existing]; }

}  // extern "C"
