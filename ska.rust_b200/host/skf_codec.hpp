// skf_codec.hpp -- reader/writer of the .skf on-disk format (host side of MergeSkaArray::save/load,
// src/merge_ska_array.rs:191-204): a Snappy *framing* stream (snap::write::FrameEncoder) whose
// payload is one CBOR map written by ciborium from the serde derive of MergeSkaArray
// (field order src/merge_ska_array.rs:108-126; ndarray's serde layout {v,dim,data}).
// No third-party code: Snappy raw blocks, CRC-32C and the CBOR subset are implemented here.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace skf {

struct Array {                       // host image of MergeSkaArray<IntT>
    uint32_t k = 0;
    bool rc = false;
    std::vector<std::string> names;
    uint32_t key_bits = 64;          // 64: keys holds 1 word per k-mer; 128: (lo, hi) pairs
    std::vector<uint64_t> keys;
    uint64_t n_rows = 0, n_samples = 0;
    std::vector<uint8_t> variants;   // row-major n_rows x n_samples
    std::vector<uint64_t> counts;
    std::string ska_version = "0.5.2";
    uint64_t n_keys() const { return keys.size() / (key_bits / 64); }
};

// Throws std::runtime_error on malformed input (the reference returns Err -> main panics, lib.rs:660).
Array load(const std::string& path);
Array decode(const uint8_t* data, size_t len, int threads = 0 /* 0: all hardware threads */);
// Writes a file the reference loads: CBOR identical in structure to ciborium's output, Snappy frames
// with compressed (type 0x00) chunks of <= 65536 bytes and masked CRC-32C.  `threads` compress chunks in parallel.
void save(const Array& a, const std::string& path, int threads = 1);
std::vector<uint8_t> encode(const Array& a, int threads = 1);

uint32_t crc32c(const uint8_t* p, size_t n);

}  // namespace skf
