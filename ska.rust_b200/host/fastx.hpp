// fastx.hpp -- FASTA/FASTQ(.gz) record reader standing in for needletail's parse_fastx_file
// (call sites src/ska_dict.rs:131-147, 353-360).  Behaviour relied upon: format sniffed
// from the first byte ('>' FASTA, '@' FASTQ), gzip by magic bytes, FASTA sequences may be wrapped (line
// breaks stripped), FASTQ is strict 4-line records with |qual| == |seq|.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace fastx {

struct Records {                     // all records of one file, concatenated
    bool is_fastq = false;
    std::vector<uint8_t> seq;
    std::vector<uint8_t> qual;       // empty for FASTA
    std::vector<uint64_t> rec_off;   // n_rec + 1
    uint64_t n_rec() const { return rec_off.empty() ? 0 : rec_off.size() - 1; }
};

// Throws std::runtime_error("Invalid path/file: ...") / ("Invalid FASTA/Q record") like the reference's panics.
Records parse_file(const std::string& path);
Records parse_bytes(const uint8_t* data, size_t n);
// With the record ids (header line without '>' / '@'), as RefSka needs chromosome names (ska_ref.rs:207-211).
struct NamedRecords { Records rec; std::vector<std::string> ids; };
NamedRecords parse_file_named(const std::string& path);
// Appends `b` (file 2 of a paired sample) to `a` (file 1).
void append(Records& a, const Records& b);

}  // namespace fastx
