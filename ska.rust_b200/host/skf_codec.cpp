// skf_codec.cpp -- see skf_codec.hpp.  Format notes:
//   file    = chunk*            chunk = type:u8 len:u24le body
//   0xFF    stream identifier "sNaPpY"; 0x00 compressed data; 0x01 uncompressed data;
//   0xFE / 0x80..0xFD skippable; data body = masked_crc32c(uncompressed):u32le ++ payload (<= 65536 B)
//   payload = one CBOR map(8): k, rc, names, split_kmers, variants{v,dim,data}, variant_count,
//             ska_version, k_bits     (u128 keys >= 2^64 as tag 2 + big-endian byte string)
#include "skf_codec.hpp"

#include <nmmintrin.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <thread>

namespace skf {

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
static void snappy_decompress(const uint8_t* in, size_t n, std::vector<uint8_t>& out) {
    size_t pos = 0; uint64_t ulen = 0; int shift = 0;
    while (true) {
        if (pos >= n) throw std::runtime_error("skf: truncated snappy block");
        uint8_t b = in[pos++];
        ulen |= (uint64_t)(b & 0x7F) << shift;
        if (!(b & 0x80)) break;
        shift += 7;
    }
    const size_t start = out.size();
    out.reserve(start + ulen);
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
            out.insert(out.end(), in + pos, in + pos + len);
            pos += len;
        } else {
            size_t len, off;
            if (t == 1) { if (pos >= n) throw std::runtime_error("skf: truncated copy"); len = 4 + ((tag >> 2) & 7); off = ((size_t)(tag >> 5) << 8) | in[pos]; pos += 1; }
            else if (t == 2) { if (pos + 2 > n) throw std::runtime_error("skf: truncated copy"); len = (tag >> 2) + 1; off = in[pos] | ((size_t)in[pos + 1] << 8); pos += 2; }
            else { if (pos + 4 > n) throw std::runtime_error("skf: truncated copy"); len = (tag >> 2) + 1; off = in[pos] | ((size_t)in[pos + 1] << 8) | ((size_t)in[pos + 2] << 16) | ((size_t)in[pos + 3] << 24); pos += 4; }
            if (off == 0 || off > out.size() - start) throw std::runtime_error("skf: bad snappy copy offset");
            size_t from = out.size() - off;
            for (size_t i = 0; i < len; i++) out.push_back(out[from + i]);   // may overlap its own output
        }
    }
    if (out.size() - start != ulen) throw std::runtime_error("skf: snappy length mismatch");
}

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
        if (cand < i && cw == w) {
            size_t len = 4;
            while (i + len < n && in[cand + len] == in[i + len]) len++;
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
static std::vector<uint8_t> frame_decode(const uint8_t* d, size_t n) {
    std::vector<uint8_t> out;
    size_t pos = 0; bool seen = false;
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
            uint32_t crc; memcpy(&crc, body, 4);
            size_t before = out.size();
            if (type == 0x00) snappy_decompress(body + 4, len - 4, out);
            else out.insert(out.end(), body + 4, body + len);
            if (masked_crc(out.data() + before, out.size() - before) != crc) throw std::runtime_error("skf: checksum mismatch");
        } else if (type == 0xFE || (type >= 0x80 && type <= 0xFD)) {
            continue;
        } else {
            throw std::runtime_error("skf: unskippable reserved chunk");
        }
    }
    if (!seen) throw std::runtime_error("skf: not a snappy frame stream");
    return out;
}

static std::vector<uint8_t> frame_encode(const std::vector<uint8_t>& payload, int threads) {
    const size_t CH = 65536;
    const size_t n_chunks = (payload.size() + CH - 1) / CH;
    std::vector<std::vector<uint8_t>> comp(n_chunks);
    std::atomic<size_t> next{0};
    auto work = [&] {
        std::vector<uint8_t> buf;
        for (size_t c; (c = next.fetch_add(1)) < n_chunks;) {
            const uint8_t* p = payload.data() + c * CH;
            size_t len = std::min(CH, payload.size() - c * CH);
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
    std::vector<uint8_t> out = {0xFF, 0x06, 0x00, 0x00, 's', 'N', 'a', 'P', 'p', 'Y'};
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

Array decode(const uint8_t* data, size_t len) {
    std::vector<uint8_t> payload = frame_decode(data, len);
    Reader r{payload.data(), payload.size()};
    Array a;
    std::vector<uint64_t> lo, hi;
    bool wide = false;
    uint64_t fields = r.map();
    for (uint64_t f = 0; f < fields; f++) {
        std::string key = r.text();
        if (key == "k") a.k = (uint32_t)r.uint_();
        else if (key == "rc") a.rc = r.boolean();
        else if (key == "names") { uint64_t n = r.array(); for (uint64_t i = 0; i < n; i++) a.names.push_back(r.text()); }
        else if (key == "split_kmers") {
            uint64_t n = r.array(); lo.resize(n); hi.resize(n);
            for (uint64_t i = 0; i < n; i++) { r.key128(lo[i], hi[i]); wide |= hi[i] != 0; }
        } else if (key == "variants") {
            uint64_t m = r.map();
            for (uint64_t j = 0; j < m; j++) {
                std::string vk = r.text();
                if (vk == "v") { if (r.uint_() != 1) throw std::runtime_error("skf: unexpected ndarray version"); }
                else if (vk == "dim") { if (r.array() != 2) throw std::runtime_error("skf: bad dim"); a.n_rows = r.uint_(); a.n_samples = r.uint_(); }
                else if (vk == "data") {
                    uint64_t n = r.array(); a.variants.resize(n);
                    for (uint64_t i = 0; i < n; i++) a.variants[i] = (uint8_t)r.uint_();
                } else throw std::runtime_error("skf: unexpected key in variants");
            }
        } else if (key == "variant_count") { uint64_t n = r.array(); a.counts.resize(n); for (uint64_t i = 0; i < n; i++) a.counts[i] = r.uint_(); }
        else if (key == "ska_version") a.ska_version = r.text();
        else if (key == "k_bits") (void)r.uint_();      // informational (merge_ska_array.rs:124-125); width follows the values
        else throw std::runtime_error("skf: unexpected field " + key);
    }
    if (a.variants.size() != a.n_rows * a.n_samples || a.counts.size() != a.n_rows || a.names.size() != a.n_samples)
        throw std::runtime_error("skf: inconsistent dimensions");
    // the reference tries MergeSkaArray::<u64>::load first and falls back to u128 (lib.rs:635-661)
    a.key_bits = wide ? 128 : 64;
    if (wide) { a.keys.resize(2 * lo.size()); for (size_t i = 0; i < lo.size(); i++) { a.keys[2 * i] = lo[i]; a.keys[2 * i + 1] = hi[i]; } }
    else a.keys = std::move(lo);
    return a;
}

Array load(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("No such file or directory: " + path);
    std::vector<uint8_t> buf;
    uint8_t tmp[1 << 16];
    size_t n;
    while ((n = fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    fclose(f);
    return decode(buf.data(), buf.size());
}

static inline void put_head(std::vector<uint8_t>& o, int major, uint64_t v) {
    if (v < 24) { o.push_back((uint8_t)((major << 5) | v)); return; }
    int nb = v < (1ull << 8) ? 1 : v < (1ull << 16) ? 2 : v < (1ull << 32) ? 4 : 8;
    o.push_back((uint8_t)((major << 5) | (nb == 1 ? 24 : nb == 2 ? 25 : nb == 4 ? 26 : 27)));
    for (int i = nb - 1; i >= 0; i--) o.push_back((uint8_t)(v >> (8 * i)));
}
static inline void put_text(std::vector<uint8_t>& o, const std::string& s) { put_head(o, 3, s.size()); o.insert(o.end(), s.begin(), s.end()); }

std::vector<uint8_t> encode(const Array& a, int threads) {
    std::vector<uint8_t> o;
    o.reserve(64 + a.keys.size() * 9 + a.variants.size() * 2 + a.counts.size() * 3);
    o.push_back(0xA8);
    put_text(o, "k"); put_head(o, 0, a.k);
    put_text(o, "rc"); o.push_back(a.rc ? 0xF5 : 0xF4);
    put_text(o, "names"); put_head(o, 4, a.names.size()); for (auto& s : a.names) put_text(o, s);
    put_text(o, "split_kmers"); put_head(o, 4, a.n_keys());
    if (a.key_bits == 64) for (uint64_t v : a.keys) put_head(o, 0, v);
    else for (size_t i = 0; i < a.keys.size(); i += 2) {
        uint64_t lo = a.keys[i], hi = a.keys[i + 1];
        if (!hi) { put_head(o, 0, lo); continue; }
        o.push_back(0xC2);                                  // tag 2: unsigned bignum, minimal big-endian bytes
        int nb = 16; while (nb > 8 && !((hi >> (8 * (nb - 9))) & 0xFF)) nb--;
        put_head(o, 2, nb);
        for (int b = nb - 1; b >= 0; b--) o.push_back((uint8_t)(b >= 8 ? hi >> (8 * (b - 8)) : lo >> (8 * b)));
    }
    put_text(o, "variants"); o.push_back(0xA3);
    put_text(o, "v"); o.push_back(0x01);
    put_text(o, "dim"); o.push_back(0x82); put_head(o, 0, a.n_rows); put_head(o, 0, a.n_samples);
    put_text(o, "data"); put_head(o, 4, a.variants.size());
    {   // every middle-base byte is '-' (45) or a letter: exactly two bytes `18 xx` each
        size_t at = o.size();
        o.resize(at + 2 * a.variants.size());
        uint8_t* w = o.data() + at; size_t extra = 0;
        for (uint8_t b : a.variants) { if (b < 24) { extra++; break; } *w++ = 0x18; *w++ = b; }
        if (extra) { o.resize(at); for (uint8_t b : a.variants) put_head(o, 0, b); }
    }
    put_text(o, "variant_count"); put_head(o, 4, a.counts.size()); for (uint64_t c : a.counts) put_head(o, 0, c);
    put_text(o, "ska_version"); put_text(o, a.ska_version);
    put_text(o, "k_bits"); put_head(o, 0, a.key_bits);
    return frame_encode(o, threads);
}

void save(const Array& a, const std::string& path, int threads) {
    std::vector<uint8_t> bytes = encode(a, threads);
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot create " + path);
    size_t w = fwrite(bytes.data(), 1, bytes.size(), f);
    fclose(f);
    if (w != bytes.size()) throw std::runtime_error("short write to " + path);
}

}  // namespace skf
