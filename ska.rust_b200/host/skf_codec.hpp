// skf_codec.hpp -- reader/writer of the .skf on-disk format (host side of MergeSkaArray::save/load,
// src/merge_ska_array.rs:191-204): a Snappy *framing* stream (snap::write::FrameEncoder) whose
// payload is one CBOR map written by ciborium from the serde derive of MergeSkaArray
// (field order src/merge_ska_array.rs:108-126; ndarray's serde layout {v,dim,data}).
// No third-party code: Snappy raw blocks, CRC-32C and the CBOR subset are implemented here.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace skf {

// Byte buffer WITHOUT zero-initialisation: the middle-base matrix is gigabytes that are overwritten at once (by the device
// copy or by the decoder's threads, which thereby also first-touch its pages in parallel).
class Bytes {
  public:
    Bytes() {}
    Bytes(Bytes&& o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr; o.n_ = 0; }
    Bytes& operator=(Bytes&& o) noexcept { if (this != &o) { delete[] p_; p_ = o.p_; n_ = o.n_; o.p_ = nullptr; o.n_ = 0; } return *this; }
    Bytes(const Bytes& o) : p_(o.n_ ? new uint8_t[o.n_] : nullptr), n_(o.n_) { if (n_) std::memcpy(p_, o.p_, n_); }
    Bytes& operator=(const Bytes& o) { if (this != &o) { Bytes t(o); *this = std::move(t); } return *this; }
    ~Bytes() { delete[] p_; }
    void resize(size_t n) { if (n != n_) { delete[] p_; p_ = n ? new uint8_t[n] : nullptr; n_ = n; } }   // contents are NOT kept
    uint8_t* data() { return p_; }
    const uint8_t* data() const { return p_; }
    size_t size() const { return n_; }
    bool empty() const { return n_ == 0; }
    uint8_t& operator[](size_t i) { return p_[i]; }
    const uint8_t& operator[](size_t i) const { return p_[i]; }
    bool operator==(const Bytes& o) const { return n_ == o.n_ && (n_ == 0 || std::memcmp(p_, o.p_, n_) == 0); }
    bool operator!=(const Bytes& o) const { return !(*this == o); }
  private:
    uint8_t* p_ = nullptr;
    size_t n_ = 0;
};
// touches every page of [p, p + n) from `threads` threads (first touch of a fresh allocation is what a single-threaded
// device->host copy would otherwise spend most of its time on)
void prefault(uint8_t* p, size_t n, int threads);

struct Array {                       // host image of MergeSkaArray<IntT>
    uint32_t k = 0;
    bool rc = false;
    std::vector<std::string> names;
    uint32_t key_bits = 64;          // 64: keys holds 1 word per k-mer; 128: (lo, hi) pairs
    std::vector<uint64_t> keys;
    uint64_t n_rows = 0, n_samples = 0;
    Bytes variants;                  // row-major n_rows x n_samples
    std::vector<uint64_t> counts;
    std::string ska_version = "0.5.2";
    uint32_t k_bits_field = 0;       // the file's informational k_bits (merge_ska_array.rs:124-125): a k = 33 array whose keys all fit
                                     // 64 bits loads as u64 in the reference too, yet keeps saying 128; 0 = derive from key_bits
    uint64_t n_keys() const { return keys.size() / (key_bits / 64); }
};

// Throws std::runtime_error on malformed input (the reference returns Err -> main panics, lib.rs:660).
Array load(const std::string& path);
Array decode_file_single_thread(const std::string& path);   // one thread throughout (what the reference's load does)
Array decode(const uint8_t* data, size_t len, int threads = 0 /* 0: all hardware threads */);
// Writes a file the reference loads: CBOR identical in structure to ciborium's output, Snappy frames
// with compressed (type 0x00) chunks of <= 65536 bytes and masked CRC-32C.  `threads` compress chunks in parallel.
void save(const Array& a, const std::string& path, int threads = 1);
std::vector<uint8_t> encode(const Array& a, int threads = 1);

uint32_t crc32c(const uint8_t* p, size_t n);

}  // namespace skf
