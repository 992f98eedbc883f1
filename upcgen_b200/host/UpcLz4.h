// UpcLz4.h -- the LZ4 block format and the XXH64 checksum, as ROOT's compressed records use them ("L4" blocks:
// core/lz4 of ROOT; the reference asks for them with TFile compression setting 4 * 100 + 9, src/UpcGenerator.cpp:843).
// Written from the published format descriptions (lz4_Block_format.md, xxhash_spec.md); no code of either library.
//
// A ROOT "L4" block: 'L' '4' <lz4 major version> <3 bytes: compressed size, little endian> <3 bytes: uncompressed
// size> <8 bytes: XXH64 (seed 0) of the LZ4 bytes, big endian> <LZ4 block>.  The compressed size counts the checksum.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace upc_lz4
{
uint64_t xxh64(const unsigned char* data, size_t n, uint64_t seed = 0);

// decodes one LZ4 block of n bytes into exactly out_n bytes; false if the stream is malformed or does not fill out
bool decompress_block(const unsigned char* src, size_t n, unsigned char* out, size_t out_n);

// appends one LZ4 block holding src[0, n) to dst (greedy matcher on a 4-byte hash table; any conforming decoder reads
// it).  Returns the number of bytes appended.
size_t compress_block(const unsigned char* src, size_t n, std::vector<unsigned char>& dst);

// ROOT's framing: appends "L4" blocks (at most 0xffffff input bytes each) for src[0, n) to dst; returns false -- and
// leaves dst as it was -- when the result would not be smaller than n (ROOT then stores the buffer as it is)
bool root_zip(const unsigned char* src, size_t n, std::vector<unsigned char>& dst);

// inflates a chain of "L4" blocks into out_n bytes, checking every checksum
bool root_unzip_block(const unsigned char* block, size_t avail, unsigned char* out, size_t out_cap, size_t* consumed,
                      size_t* produced);
}  // namespace upc_lz4
