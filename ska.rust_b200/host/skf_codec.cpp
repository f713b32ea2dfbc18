// skf_codec.cpp -- see skf_codec.hpp.  Format notes:
//   file    = chunk*            chunk = type:u8 len:u24le body
//   0xFF    stream identifier "sNaPpY"; 0x00 compressed data; 0x01 uncompressed data;
//   0xFE / 0x80..0xFD skippable; data body = masked_crc32c(uncompressed):u32le ++ payload (<= 65536 B)
//   payload = one CBOR map(8): k, rc, names, split_kmers, variants{v,dim,data}, variant_count,
//             ska_version, k_bits     (u128 keys >= 2^64 as tag 2 + big-endian byte string)
#include "skf_codec.hpp"

#include <emmintrin.h>
#include <nmmintrin.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <memory>
#include <mutex>
#include <thread>

namespace skf {

void prefault(uint8_t* p, size_t n, int threads) {
    const int nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(1, threads), n >> 24));
    auto part = [&](int t) {
        const size_t b0 = n / nt * t, b1 = t + 1 == nt ? n : n / nt * (t + 1);
        for (size_t i = b0; i < b1; i += 4096) ((volatile uint8_t*)p)[i] = 0;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(part, t);
    part(0);
    for (auto& t : th) t.join();
}

// ------------------------------------------------------------------------------------------ CRC-32C
uint32_t crc32c(const uint8_t* p, size_t n) {
    uint64_t c = 0xFFFFFFFFu;
    while (n >= 8) { uint64_t v; memcpy(&v, p, 8); c = _mm_crc32_u64(c, v); p += 8; n -= 8; }
    while (n--) c = _mm_crc32_u8((uint32_t)c, *p++);
    return (uint32_t)c ^ 0xFFFFFFFFu;
}
static uint32_t masked_crc(const uint8_t* p, size_t n) {
    uint32_t c = crc32c(p, n);
    return ((c >> 15) | (c << 17)) + 0xA282EAD8u;
}

// ------------------------------------------------------------------------------------------ Snappy raw
// uncompressed length announced by a raw Snappy block (varint header); *hdr = bytes of the header
static uint64_t snappy_ulen(const uint8_t* in, size_t n, size_t* hdr) {
    size_t pos = 0; uint64_t ulen = 0; int shift = 0;
    while (true) {
        if (pos >= n || shift > 35) throw std::runtime_error("skf: truncated snappy block");
        uint8_t b = in[pos++];
        ulen |= (uint64_t)(b & 0x7F) << shift;
        if (!(b & 0x80)) break;
        shift += 7;
    }
    *hdr = pos;
    return ulen;
}
// decompresses one raw Snappy block into dst[0, ulen) (the caller sized it from snappy_ulen)
static void snappy_decompress(const uint8_t* in, size_t n, uint8_t* dst, size_t ulen) {
    size_t pos = 0;
    if (snappy_ulen(in, n, &pos) != ulen) throw std::runtime_error("skf: snappy length mismatch");
    size_t o = 0;
    while (pos < n) {
        uint8_t tag = in[pos++];
        uint32_t t = tag & 3;
        if (t == 0) {
            size_t len = tag >> 2;
            if (len >= 60) {
                size_t nb = len - 59; len = 0;
                if (pos + nb > n) throw std::runtime_error("skf: truncated literal length");
                for (size_t i = 0; i < nb; i++) len |= (size_t)in[pos + i] << (8 * i);
                pos += nb;
            }
            len += 1;
            if (pos + len > n) throw std::runtime_error("skf: truncated literal");
            if (o + len > ulen) throw std::runtime_error("skf: snappy length mismatch");
            memcpy(dst + o, in + pos, len);
            o += len; pos += len;
        } else {
            size_t len, off;
            if (t == 1) { if (pos >= n) throw std::runtime_error("skf: truncated copy"); len = 4 + ((tag >> 2) & 7); off = ((size_t)(tag >> 5) << 8) | in[pos]; pos += 1; }
            else if (t == 2) { if (pos + 2 > n) throw std::runtime_error("skf: truncated copy"); len = (tag >> 2) + 1; off = in[pos] | ((size_t)in[pos + 1] << 8); pos += 2; }
            else { if (pos + 4 > n) throw std::runtime_error("skf: truncated copy"); len = (tag >> 2) + 1; off = in[pos] | ((size_t)in[pos + 1] << 8) | ((size_t)in[pos + 2] << 16) | ((size_t)in[pos + 3] << 24); pos += 4; }
            if (off == 0 || off > o) throw std::runtime_error("skf: bad snappy copy offset");
            if (o + len > ulen) throw std::runtime_error("skf: snappy length mismatch");
            if (off >= len) memcpy(dst + o, dst + o - off, len);
            else if (len < 16) for (size_t i = 0; i < len; i++) dst[o + i] = dst[o - off + i];
            else {
                // the copy overlaps its own output: a run of period `off` (2 for a run of one middle base).  What has been
                // written so far repeats with that period, so pieces can be taken from ever further back: off, 2 off, 4 off ...
                uint8_t* q = dst + o;
                size_t rem = len, d = off;
                while (rem) {
                    const size_t piece = rem < d ? rem : d;
                    memcpy(q, q - d, piece);
                    q += piece; rem -= piece;
                    if (d <= 16384) d *= 2;
                }
            }
            o += len;
        }
    }
    if (o != ulen) throw std::runtime_error("skf: snappy length mismatch");
}

// ------------------------------------------------------------------------------------------ Snappy raw (compressor)
static inline void emit_literal(std::vector<uint8_t>& out, const uint8_t* p, size_t len) {
    if (!len) return;
    size_t n = len - 1;
    if (n < 60) out.push_back((uint8_t)(n << 2));
    else if (n < 256) { out.push_back(60 << 2); out.push_back((uint8_t)n); }
    else { out.push_back(61 << 2); out.push_back((uint8_t)(n & 0xFF)); out.push_back((uint8_t)(n >> 8)); }   // block <= 65536
    out.insert(out.end(), p, p + len);
}
static inline void emit_copy(std::vector<uint8_t>& out, size_t off, size_t len) {
    while (len >= 68) { out.push_back((uint8_t)((63 << 2) | 2)); out.push_back((uint8_t)(off & 0xFF)); out.push_back((uint8_t)(off >> 8)); len -= 64; }
    if (len > 64) { out.push_back((uint8_t)((59 << 2) | 2)); out.push_back((uint8_t)(off & 0xFF)); out.push_back((uint8_t)(off >> 8)); len -= 60; }
    if (len >= 12 || off >= 2048) { out.push_back((uint8_t)(((len - 1) << 2) | 2)); out.push_back((uint8_t)(off & 0xFF)); out.push_back((uint8_t)(off >> 8)); }
    else { out.push_back((uint8_t)(((off >> 8) << 5) | ((len - 4) << 2) | 1)); out.push_back((uint8_t)(off & 0xFF)); }
}
// greedy hash-chain-free matcher over one block of <= 65536 bytes
static void snappy_compress(const uint8_t* in, size_t n, std::vector<uint8_t>& out) {
    out.clear();
    size_t v = n;
    while (v >= 0x80) { out.push_back((uint8_t)(v | 0x80)); v >>= 7; }
    out.push_back((uint8_t)v);
    if (n < 16) { emit_literal(out, in, n); return; }
    uint16_t table[1 << 13];
    memset(table, 0, sizeof table);
    size_t lit = 0, i = 0;
    const size_t limit = n - 8;
    while (i < limit) {
        uint32_t w; memcpy(&w, in + i, 4);
        uint32_t h = (w * 0x1E35A7BDu) >> 19;
        size_t cand = table[h];
        table[h] = (uint16_t)i;
        uint32_t cw; memcpy(&cw, in + cand, 4);
        if (i >= 2) {                                           // the data section is `18 xx 18 xx ...`: a run of one letter is a match at distance 2
            uint32_t pw; memcpy(&pw, in + i - 2, 4);
            if (pw == w) { cand = i - 2; cw = w; }
        }
        if (cand < i && cw == w) {
            size_t len = 4;
            bool open_end = true;                               // extend the match 8 bytes at a time (runs of `18 2D` are long)
            while (i + len + 8 <= n) {
                uint64_t x, y; memcpy(&x, in + cand + len, 8); memcpy(&y, in + i + len, 8);
                const uint64_t d = x ^ y;
                if (d) { len += (size_t)(__builtin_ctzll(d) >> 3); open_end = false; break; }
                len += 8;
            }
            if (open_end) while (i + len < n && in[cand + len] == in[i + len]) len++;
            emit_literal(out, in + lit, i - lit);
            emit_copy(out, i - cand, len);
            i += len; lit = i;
        } else {
            i++;
        }
    }
    emit_literal(out, in + lit, n - lit);
}

// ------------------------------------------------------------------------------------------ framing
// Chunks are independent (<= 64 KiB of payload each): a first walk over the headers sizes the payload once, then the
// chunks are decompressed and checksummed in parallel straight into place.  (Appending chunk by chunk with an exact
// reserve() re-copied the whole payload for every chunk: minutes for a 1 GB file.)
struct Payload { std::unique_ptr<uint8_t[]> p; size_t n = 0; };   // uninitialised storage: the pages are first touched by the worker threads
static Payload frame_decode(const uint8_t* d, size_t n, int threads) {
    struct Chunk { uint8_t type; const uint8_t* body; size_t len, ulen, at; };
    std::vector<Chunk> chunks;
    size_t pos = 0, total = 0; bool seen = false;
    while (pos < n) {
        if (pos + 4 > n) throw std::runtime_error("skf: truncated chunk header");
        uint8_t type = d[pos];
        size_t len = d[pos + 1] | ((size_t)d[pos + 2] << 8) | ((size_t)d[pos + 3] << 16);
        pos += 4;
        if (pos + len > n) throw std::runtime_error("skf: truncated chunk");
        const uint8_t* body = d + pos;
        pos += len;
        if (type == 0xFF) {
            if (len != 6 || memcmp(body, "sNaPpY", 6)) throw std::runtime_error("skf: bad stream identifier");
            seen = true;
        } else if (type == 0x00 || type == 0x01) {
            if (!seen) throw std::runtime_error("skf: missing stream identifier");
            if (len < 4) throw std::runtime_error("skf: short data chunk");
            size_t hdr = 0;
            const size_t ulen = type == 0x00 ? (size_t)snappy_ulen(body + 4, len - 4, &hdr) : len - 4;
            if (ulen > (1u << 24)) throw std::runtime_error("skf: oversized chunk");
            chunks.push_back({type, body, len, ulen, total});
            total += ulen;
        } else if (type == 0xFE || (type >= 0x80 && type <= 0xFD)) {
            continue;
        } else {
            throw std::runtime_error("skf: unskippable reserved chunk");
        }
    }
    if (!seen) throw std::runtime_error("skf: not a snappy frame stream");
    Payload out;
    out.p.reset(new uint8_t[total ? total : 1]);
    out.n = total;
    std::atomic<size_t> next{0};
    std::atomic<bool> failed{false};
    std::string error;
    std::mutex mu;
    auto work = [&] {
        try {
            for (size_t c; (c = next.fetch_add(1)) < chunks.size() && !failed.load();) {
                const Chunk& ch = chunks[c];
                uint32_t crc; memcpy(&crc, ch.body, 4);
                uint8_t* dst = out.p.get() + ch.at;
                if (ch.type == 0x00) snappy_decompress(ch.body + 4, ch.len - 4, dst, ch.ulen);
                else memcpy(dst, ch.body + 4, ch.ulen);
                if (masked_crc(dst, ch.ulen) != crc) throw std::runtime_error("skf: checksum mismatch");
            }
        } catch (std::exception& e) {
            std::lock_guard<std::mutex> g(mu);
            if (!failed.exchange(true)) error = e.what();
        }
    };
    threads = std::max(1, std::min<int>(threads, (int)std::max<size_t>(1, chunks.size() / 16)));
    std::vector<std::thread> th;
    for (int t = 1; t < threads; t++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    if (failed.load()) throw std::runtime_error(error);
    return out;
}

// The payload is never held in one piece: `fill(offset, len, dst)` materialises any part of it into a worker's 64 KiB
// buffer (for an array of 3 GB of middle bases the CBOR payload would be 6 GB of `18 xx` pairs).
template <class Fill>
static std::vector<std::vector<uint8_t>> frame_encode_chunks(size_t payload_size, int threads, Fill fill) {
    const size_t CH = 65536;
    const size_t n_chunks = (payload_size + CH - 1) / CH;
    std::vector<std::vector<uint8_t>> comp(n_chunks);
    std::atomic<size_t> next{0};
    auto work = [&] {
        std::vector<uint8_t> buf, piece(CH);
        for (size_t c; (c = next.fetch_add(1)) < n_chunks;) {
            size_t len = std::min(CH, payload_size - c * CH);
            uint8_t* p = piece.data();
            fill(c * CH, len, p);
            snappy_compress(p, len, buf);
            bool raw = buf.size() >= len - len / 8;          // snap's rule of thumb: keep poorly compressible chunks raw
            size_t body = 4 + (raw ? len : buf.size());
            std::vector<uint8_t>& o = comp[c];
            o.resize(4 + body);
            o[0] = raw ? 0x01 : 0x00; o[1] = body & 0xFF; o[2] = (body >> 8) & 0xFF; o[3] = (body >> 16) & 0xFF;
            uint32_t crc = masked_crc(p, len);
            memcpy(&o[4], &crc, 4);
            memcpy(&o[8], raw ? p : buf.data(), raw ? len : buf.size());
        }
    };
    threads = std::max(1, threads);
    std::vector<std::thread> th;
    for (int t = 1; t < threads; t++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    return comp;
}
static const uint8_t STREAM_ID[10] = {0xFF, 0x06, 0x00, 0x00, 's', 'N', 'a', 'P', 'p', 'Y'};
static std::vector<uint8_t> join_chunks(const std::vector<std::vector<uint8_t>>& comp) {
    std::vector<uint8_t> out(STREAM_ID, STREAM_ID + 10);
    size_t total = out.size();
    for (auto& c : comp) total += c.size();
    out.reserve(total);
    for (auto& c : comp) out.insert(out.end(), c.begin(), c.end());
    return out;
}

// ------------------------------------------------------------------------------------------ CBOR
struct Reader {
    const uint8_t* p; size_t n, pos = 0;
    uint8_t peek() const { if (pos >= n) throw std::runtime_error("skf: truncated CBOR"); return p[pos]; }
    void head(int& major, uint64_t& val) {
        uint8_t ib = peek(); pos++;
        major = ib >> 5; int info = ib & 31;
        if (info < 24) { val = info; return; }
        int nb = info == 24 ? 1 : info == 25 ? 2 : info == 26 ? 4 : info == 27 ? 8 : -1;
        if (nb < 0 || pos + nb > n) throw std::runtime_error("skf: unsupported CBOR head");
        val = 0;
        for (int i = 0; i < nb; i++) val = (val << 8) | p[pos + i];
        pos += nb;
    }
    uint64_t uint_() { int m; uint64_t v; head(m, v); if (m != 0) throw std::runtime_error("skf: expected unsigned integer"); return v; }
    std::string text() { int m; uint64_t v; head(m, v); if (m != 3 || pos + v > n) throw std::runtime_error("skf: expected text"); std::string s((const char*)p + pos, v); pos += v; return s; }
    uint64_t array() { int m; uint64_t v; head(m, v); if (m != 4) throw std::runtime_error("skf: expected array"); return v; }
    uint64_t map() { int m; uint64_t v; head(m, v); if (m != 5) throw std::runtime_error("skf: expected map"); return v; }
    bool boolean() { uint8_t b = peek(); pos++; if (b == 0xF5) return true; if (b == 0xF4) return false; throw std::runtime_error("skf: expected bool"); }
    // uint or tag-2 bignum -> (lo, hi)
    void key128(uint64_t& lo, uint64_t& hi) {
        int m; uint64_t v; head(m, v);
        if (m == 0) { lo = v; hi = 0; return; }
        if (m == 6 && v == 2) {
            int m2; uint64_t len; head(m2, len);
            if (m2 != 2 || len > 16 || pos + len > n) throw std::runtime_error("skf: bad bignum");
            lo = hi = 0;
            for (uint64_t i = 0; i < len; i++) { hi = (hi << 8) | (lo >> 56); lo = (lo << 8) | p[pos + i]; }
            pos += len;
            return;
        }
        throw std::runtime_error("skf: expected split k-mer integer");
    }
};

Array decode(const uint8_t* data, size_t len, int threads) {
    // all hardware threads but one: the caller's CUDA context is usually being created on a helper thread meanwhile
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency() > 2 ? std::thread::hardware_concurrency() - 1 : 1u);
    Payload payload = frame_decode(data, len, threads);
    Reader r{payload.p.get(), payload.n};
    Array a;
    std::vector<uint64_t> lo, hi;
    // an array header cannot announce more elements than the payload has bytes left (every element takes at least one)
    auto bounded = [&](uint64_t n) { if (n > r.n - r.pos) throw std::runtime_error("skf: array longer than the payload"); return n; };
    bool wide = false;
    uint64_t fields = r.map();
    for (uint64_t f = 0; f < fields; f++) {
        std::string key = r.text();
        if (key == "k") a.k = (uint32_t)r.uint_();
        else if (key == "rc") a.rc = r.boolean();
        else if (key == "names") { uint64_t n = r.array(); for (uint64_t i = 0; i < n; i++) a.names.push_back(r.text()); }
        else if (key == "split_kmers") {
            uint64_t n = bounded(r.array()); lo.resize(n); hi.resize(n);
            for (uint64_t i = 0; i < n; i++) { r.key128(lo[i], hi[i]); wide |= hi[i] != 0; }
        } else if (key == "variants") {
            uint64_t m = r.map();
            for (uint64_t j = 0; j < m; j++) {
                std::string vk = r.text();
                if (vk == "v") { if (r.uint_() != 1) throw std::runtime_error("skf: unexpected ndarray version"); }
                else if (vk == "dim") { if (r.array() != 2) throw std::runtime_error("skf: bad dim"); a.n_rows = r.uint_(); a.n_samples = r.uint_(); }
                else if (vk == "data") {
                    uint64_t n = bounded(r.array()); a.variants.resize(n);
                    // every middle-base byte written by ska is '-' or a letter (>= 24): exactly `18 xx` each.  Check and
                    // take that form with all threads; anything else goes through the general integer reader.
                    bool fixed = r.pos + 2 * n <= r.n;
                    if (fixed && n) {
                        const uint8_t* src = r.p + r.pos;
                        const int nt = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)threads, n >> 20));
                        std::vector<char> ok(nt, 1);
                        auto part = [&](int t) {
                            const uint64_t b0 = n * t / nt, b1 = n * (t + 1) / nt;
                            uint8_t bad = 0;
                            for (uint64_t i = b0; i < b1; i++) { bad |= (uint8_t)(src[2 * i] ^ 0x18); a.variants[i] = src[2 * i + 1]; }
                            ok[t] = bad == 0;
                        };
                        std::vector<std::thread> th;
                        for (int t = 1; t < nt; t++) th.emplace_back(part, t);
                        part(0);
                        for (auto& t : th) t.join();
                        for (char c : ok) fixed = fixed && c;
                    }
                    if (fixed) r.pos += 2 * n;
                    else for (uint64_t i = 0; i < n; i++) a.variants[i] = (uint8_t)r.uint_();
                } else throw std::runtime_error("skf: unexpected key in variants");
            }
        } else if (key == "variant_count") { uint64_t n = bounded(r.array()); a.counts.resize(n); for (uint64_t i = 0; i < n; i++) a.counts[i] = r.uint_(); }
        else if (key == "ska_version") a.ska_version = r.text();
        else if (key == "k_bits") a.k_bits_field = (uint32_t)r.uint_();   // informational (merge_ska_array.rs:124-125); width follows the values
        else throw std::runtime_error("skf: unexpected field " + key);
    }
    if (a.variants.size() != a.n_rows * a.n_samples || a.counts.size() != a.n_rows || a.names.size() != a.n_samples)
        throw std::runtime_error("skf: inconsistent dimensions");
    // every command that saves an array compacts its split k-mers with the rows (generic_modes.rs:252: update_kmers = true),
    // and the device copy reads n_rows keys: a file with another number of keys is damaged
    if (lo.size() != a.n_rows || (wide && hi.size() != lo.size()))
        throw std::runtime_error("skf: the number of split k-mers differs from the number of variant rows");
    // the reference tries MergeSkaArray::<u64>::load first and falls back to u128 (lib.rs:635-661)
    a.key_bits = wide ? 128 : 64;
    if (wide) { a.keys.resize(2 * lo.size()); for (size_t i = 0; i < lo.size(); i++) { a.keys[2 * i] = lo[i]; a.keys[2 * i + 1] = hi[i]; } }
    else a.keys = std::move(lo);
    return a;
}

static std::vector<uint8_t> read_file(const std::string& path);
Array decode_file_single_thread(const std::string& path) {
    std::vector<uint8_t> buf = read_file(path);
    return decode(buf.data(), buf.size(), 1);
}

Array load(const std::string& path) {
    std::vector<uint8_t> buf = read_file(path);
    return decode(buf.data(), buf.size(), 0);
}

static std::vector<uint8_t> read_file(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("No such file or directory: " + path);
    std::vector<uint8_t> buf;
    if (fseek(f, 0, SEEK_END) == 0) {
        const long sz = ftell(f);
        fseek(f, 0, SEEK_SET);
        if (sz > 0) { buf.resize((size_t)sz); buf.resize(fread(buf.data(), 1, (size_t)sz, f)); }
    }
    uint8_t tmp[1 << 16];
    size_t n;
    while ((n = fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + n);   // streams that cannot seek
    fclose(f);
    return buf;
}

static inline void put_head(std::vector<uint8_t>& o, int major, uint64_t v) {
    if (v < 24) { o.push_back((uint8_t)((major << 5) | v)); return; }
    int nb = v < (1ull << 8) ? 1 : v < (1ull << 16) ? 2 : v < (1ull << 32) ? 4 : 8;
    o.push_back((uint8_t)((major << 5) | (nb == 1 ? 24 : nb == 2 ? 25 : nb == 4 ? 26 : 27)));
    for (int i = nb - 1; i >= 0; i--) o.push_back((uint8_t)(v >> (8 * i)));
}
static inline void put_text(std::vector<uint8_t>& o, const std::string& s) { put_head(o, 3, s.size()); o.insert(o.end(), s.begin(), s.end()); }

// CBOR unsigned integers of v[0, n) appended to `o`, encoded by `threads` workers (10^7 split k-mers are 90 MB of heads)
template <class Put>
static void put_parallel(std::vector<uint8_t>& o, size_t n, int threads, Put put_one) {
    const int nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)threads, n >> 16));
    std::vector<std::vector<uint8_t>> parts(nt);
    auto part = [&](int t) {
        const size_t b0 = n * t / nt, b1 = n * (t + 1) / nt;
        parts[t].reserve((b1 - b0) * 9 + 16);
        for (size_t i = b0; i < b1; i++) put_one(parts[t], i);
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(part, t);
    part(0);
    for (auto& t : th) t.join();
    size_t total = o.size();
    for (auto& p : parts) total += p.size();
    o.reserve(total + 256);
    for (auto& p : parts) o.insert(o.end(), p.begin(), p.end());
}

static std::vector<std::vector<uint8_t>> encode_chunks(const Array& a, int threads) {
    threads = std::max(1, threads);
    // head: everything before the variant bytes; tail: everything after.  The data section between them is `18 xx`
    // per byte (every middle-base byte is '-' (45) or a letter), written by all threads into uninitialised storage.
    std::vector<uint8_t> o;
    o.reserve(64 + a.keys.size() * 9);
    o.push_back(0xA8);
    put_text(o, "k"); put_head(o, 0, a.k);
    put_text(o, "rc"); o.push_back(a.rc ? 0xF5 : 0xF4);
    put_text(o, "names"); put_head(o, 4, a.names.size()); for (auto& s : a.names) put_text(o, s);
    put_text(o, "split_kmers"); put_head(o, 4, a.n_keys());
    if (a.key_bits == 64) put_parallel(o, a.keys.size(), threads, [&a](std::vector<uint8_t>& w, size_t i) { put_head(w, 0, a.keys[i]); });
    else put_parallel(o, a.keys.size() / 2, threads, [&a](std::vector<uint8_t>& w, size_t j) {
        const uint64_t lo = a.keys[2 * j], hi = a.keys[2 * j + 1];
        if (!hi) { put_head(w, 0, lo); return; }
        w.push_back(0xC2);                                  // tag 2: unsigned bignum, minimal big-endian bytes
        int nb = 16; while (nb > 8 && !((hi >> (8 * (nb - 9))) & 0xFF)) nb--;
        put_head(w, 2, nb);
        for (int b = nb - 1; b >= 0; b--) w.push_back((uint8_t)(b >= 8 ? hi >> (8 * (b - 8)) : lo >> (8 * b)));
    });
    put_text(o, "variants"); o.push_back(0xA3);
    put_text(o, "v"); o.push_back(0x01);
    put_text(o, "dim"); o.push_back(0x82); put_head(o, 0, a.n_rows); put_head(o, 0, a.n_samples);
    put_text(o, "data"); put_head(o, 4, a.variants.size());
    std::vector<uint8_t> tail;
    tail.reserve(64 + a.counts.size() * 3);
    put_text(tail, "variant_count"); put_head(tail, 4, a.counts.size());
    put_parallel(tail, a.counts.size(), threads, [&a](std::vector<uint8_t>& w, size_t i) { put_head(w, 0, a.counts[i]); });
    put_text(tail, "ska_version"); put_text(tail, a.ska_version);
    put_text(tail, "k_bits"); put_head(tail, 0, a.k_bits_field ? a.k_bits_field : a.key_bits);

    const size_t nv = a.variants.size();
    const uint8_t* v = a.variants.data();
    const size_t H = o.size(), D = 2 * nv, total = H + D + tail.size();
    const uint8_t* head = o.data();
    const uint8_t* tl = tail.data();
    // a byte below 24 would be a one-byte CBOR integer (not produced by ska): noticed while filling, then the general
    // integer writer redoes the file
    std::atomic<bool> small{false};
    std::atomic<bool>* small_p = &small;
    auto fill = [=](size_t off, size_t len, uint8_t* dst) {
        size_t done = 0;
        while (done < len) {
            const size_t at = off + done;
            if (at < H) {                                    // head
                const size_t n = std::min(len - done, H - at);
                memcpy(dst + done, head + at, n); done += n;
            } else if (at < H + D) {                         // data: byte j of the section is 0x18 (j even) or variant j/2 (j odd)
                size_t j = at - H;
                const size_t n = std::min(len - done, D - j);
                uint8_t* w = dst + done;
                size_t q = 0;
                uint8_t lt24 = 0;
                if (j & 1) { const uint8_t b = v[j >> 1]; lt24 |= (uint8_t)(b < 24); w[q++] = b; j++; }
                {                                                // 16 cells -> 32 bytes per step
                    const __m128i c18 = _mm_set1_epi8(0x18), c24 = _mm_set1_epi8(24);
                    __m128i ok = _mm_set1_epi8((char)0xFF);
                    for (; q + 32 <= n; q += 32, j += 32) {
                        const __m128i x = _mm_loadu_si128((const __m128i*)(v + (j >> 1)));
                        ok = _mm_and_si128(ok, _mm_cmpeq_epi8(_mm_min_epu8(x, c24), c24));
                        _mm_storeu_si128((__m128i*)(w + q), _mm_unpacklo_epi8(c18, x));
                        _mm_storeu_si128((__m128i*)(w + q + 16), _mm_unpackhi_epi8(c18, x));
                    }
                    if (_mm_movemask_epi8(ok) != 0xFFFF) lt24 = 1;
                }
                for (; q + 1 < n; q += 2, j += 2) { const uint8_t b = v[j >> 1]; lt24 |= (uint8_t)(b < 24); w[q] = 0x18; w[q + 1] = b; }
                if (q < n) w[q++] = 0x18;
                if (lt24) small_p->store(true);
                done += n;
            } else {                                         // tail
                const size_t n = std::min(len - done, total - at);
                memcpy(dst + done, tl + (at - H - D), n); done += n;
            }
        }
    };
    std::vector<std::vector<uint8_t>> framed = frame_encode_chunks(total, threads, fill);
    if (!small.load()) return framed;
    for (size_t i = 0; i < nv; i++) put_head(o, 0, v[i]);
    o.insert(o.end(), tail.begin(), tail.end());
    const uint8_t* whole = o.data();
    return frame_encode_chunks(o.size(), threads, [whole](size_t off, size_t len, uint8_t* dst) { memcpy(dst, whole + off, len); });
}

std::vector<uint8_t> encode(const Array& a, int threads) { return join_chunks(encode_chunks(a, threads)); }

void save(const Array& a, const std::string& path, int threads) {
    const std::vector<std::vector<uint8_t>> chunks = encode_chunks(a, threads);   // written one by one: no second copy of the file
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot create " + path);
    bool ok = fwrite(STREAM_ID, 1, 10, f) == 10;
    for (auto& c : chunks) ok = ok && fwrite(c.data(), 1, c.size(), f) == c.size();
    ok = fclose(f) == 0 && ok;
    if (!ok) throw std::runtime_error("short write to " + path);
}

}  // namespace skf
