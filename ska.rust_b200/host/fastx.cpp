#include "fastx.hpp"

#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace fastx {

static std::vector<uint8_t> read_all(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("Invalid path/file: " + path);
    std::vector<uint8_t> buf;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (sz > 0) {
        buf.resize((size_t)sz);
        size_t got = fread(buf.data(), 1, (size_t)sz, f);
        buf.resize(got);
    }
    fclose(f);
    return buf;
}

static std::vector<uint8_t> gunzip(const std::vector<uint8_t>& in) {
    std::vector<uint8_t> out;
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, 15 + 32) != Z_OK) throw std::runtime_error("Invalid FASTA/Q record");
    // zlib counts its input in 32 bits: feed files of any size in pieces of at most 1 GiB
    const uint8_t* next = in.data();
    size_t left = in.size();
    auto refill = [&] {
        if (zs.avail_in == 0 && left) {
            const size_t n = left < ((size_t)1 << 30) ? left : ((size_t)1 << 30);
            zs.next_in = const_cast<Bytef*>(next);
            zs.avail_in = (uInt)n;
            next += n; left -= n;
        }
    };
    std::vector<uint8_t> chunk(1 << 20);
    int rc;
    do {
        refill();
        zs.next_out = chunk.data();
        zs.avail_out = (uInt)chunk.size();
        rc = inflate(&zs, Z_NO_FLUSH);
        if (rc != Z_OK && rc != Z_STREAM_END && !(rc == Z_BUF_ERROR && (zs.avail_in || left))) {
            inflateEnd(&zs);
            throw std::runtime_error("Invalid FASTA/Q record");
        }
        out.insert(out.end(), chunk.data(), chunk.data() + (chunk.size() - zs.avail_out));
        if (rc == Z_STREAM_END) {
            refill();
            // what follows a gzip member is another member (concatenated gzip) or nothing
            if (zs.avail_in >= 2 && zs.next_in[0] == 0x1f && zs.next_in[1] == 0x8b) inflateReset(&zs);
            else if (zs.avail_in == 0 && left == 0) break;
            else { inflateEnd(&zs); throw std::runtime_error("Invalid FASTA/Q record"); }
        } else if (zs.avail_in == 0 && left == 0 && zs.avail_out != 0) {
            inflateEnd(&zs);
            throw std::runtime_error("Invalid FASTA/Q record");                  // truncated stream
        }
    } while (true);
    inflateEnd(&zs);
    return out;
}

static Records parse_bytes_ids(const uint8_t* d, size_t n, std::vector<std::string>* ids);
Records parse_bytes(const uint8_t* d, size_t n) { return parse_bytes_ids(d, n, nullptr); }

static Records parse_bytes_ids(const uint8_t* d, size_t n, std::vector<std::string>* ids) {
    Records r;
    if (n == 0) throw std::runtime_error("Invalid FASTA/Q record");
    r.rec_off.push_back(0);
    if (d[0] == '>') {
        r.seq.reserve(n);
        size_t i = 0;
        while (i < n) {
            // header line
            const uint8_t* nl = (const uint8_t*)memchr(d + i, '\n', n - i);
            if (ids) {
                size_t e = nl ? (size_t)(nl - d) : n;
                if (e > i + 1 && d[e - 1] == '\r') e--;
                ids->emplace_back((const char*)d + i + 1, e > i + 1 ? e - i - 1 : 0);
            }
            i = nl ? (size_t)(nl - d) + 1 : n;
            // sequence lines until the next '>' at a line start
            while (i < n && d[i] != '>') {
                const uint8_t* e = (const uint8_t*)memchr(d + i, '\n', n - i);
                size_t end = e ? (size_t)(e - d) : n;
                size_t len = end - i;
                if (len && d[end - 1] == '\r') len--;
                r.seq.insert(r.seq.end(), d + i, d + i + len);
                i = e ? end + 1 : n;
            }
            r.rec_off.push_back(r.seq.size());
        }
    } else if (d[0] == '@') {
        r.is_fastq = true;
        size_t i = 0;
        auto line = [&](size_t& b, size_t& len) {
            if (i >= n) return false;
            const uint8_t* e = (const uint8_t*)memchr(d + i, '\n', n - i);
            size_t end = e ? (size_t)(e - d) : n;
            b = i; len = end - i;
            if (len && d[end - 1] == '\r') len--;
            i = e ? end + 1 : n;
            return true;
        };
        while (i < n) {
            size_t hb, hl, sb, sl, pb, pl, qb, ql;
            if (!line(hb, hl)) break;
            if (hl == 0 && i >= n) break;                         // trailing newline
            if (!line(sb, sl) || !line(pb, pl) || !line(qb, ql) || d[hb] != '@' || pl == 0 || d[pb] != '+' || sl != ql)
                throw std::runtime_error("Invalid FASTA/Q record");
            if (ids) ids->emplace_back((const char*)d + hb + 1, hl ? hl - 1 : 0);
            r.seq.insert(r.seq.end(), d + sb, d + sb + sl);
            r.qual.insert(r.qual.end(), d + qb, d + qb + ql);
            r.rec_off.push_back(r.seq.size());
        }
    } else {
        throw std::runtime_error("Invalid FASTA/Q record");
    }
    if (r.n_rec() == 0) throw std::runtime_error("Invalid FASTA/Q record");
    return r;
}

Records parse_file(const std::string& path) {
    std::vector<uint8_t> raw = read_all(path);
    if (raw.size() >= 2 && raw[0] == 0x1f && raw[1] == 0x8b) raw = gunzip(raw);
    return parse_bytes(raw.data(), raw.size());
}

NamedRecords parse_file_named(const std::string& path) {
    std::vector<uint8_t> raw = read_all(path);
    if (raw.size() >= 2 && raw[0] == 0x1f && raw[1] == 0x8b) raw = gunzip(raw);
    NamedRecords out;
    out.rec = parse_bytes_ids(raw.data(), raw.size(), &out.ids);
    return out;
}

void append(Records& a, const Records& b) {
    uint64_t base = a.seq.size();
    a.seq.insert(a.seq.end(), b.seq.begin(), b.seq.end());
    if (a.is_fastq) {
        if (b.is_fastq) a.qual.insert(a.qual.end(), b.qual.begin(), b.qual.end());
        else a.qual.insert(a.qual.end(), b.seq.size(), (uint8_t)0xFF);   // no quality: every quality test passes (split_kmer.rs:66-71)
    }
    for (size_t i = 1; i < b.rec_off.size(); i++) a.rec_off.push_back(base + b.rec_off[i]);
}

}  // namespace fastx
