// Exercises upcgen_b200/host/UpcLz4.cpp for tests/test_root_file.py (no GPU):
//   lz4_check <input file> <out prefix>
// writes <prefix>.lz4 (one LZ4 block of the whole input), <prefix>.l4 (ROOT "L4" framing, empty file when the input does
// not shrink) and prints a JSON line with the XXH64 of the input and the outcome of decoding both again.
#include <cstdio>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

#include "UpcLz4.h"

int main(int argc, char** argv)
{
  if (argc < 3) return 2;
  std::ifstream in(argv[1], std::ios::binary);
  std::vector<unsigned char> data((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  const std::string prefix = argv[2];
  std::vector<unsigned char> blk;
  upc_lz4::compress_block(data.data(), data.size(), blk);
  std::vector<unsigned char> back(data.size());
  const bool ok_block = upc_lz4::decompress_block(blk.data(), blk.size(), back.data(), back.size()) && back == data;
  std::ofstream(prefix + ".lz4", std::ios::binary).write((const char*)blk.data(), (std::streamsize)blk.size());
  std::vector<unsigned char> framed;
  const bool shrunk = upc_lz4::root_zip(data.data(), data.size(), framed);
  bool ok_frames = true;
  if (shrunk) {
    std::vector<unsigned char> out(data.size());
    size_t q = 0, o = 0;
    while (o < out.size() && ok_frames) {
      size_t used = 0, made = 0;
      ok_frames = upc_lz4::root_unzip_block(framed.data() + q, framed.size() - q, out.data() + o, out.size() - o, &used, &made);
      q += used; o += made;
    }
    ok_frames = ok_frames && q == framed.size() && out == data;
    // a flipped payload byte must be caught by the checksum
    if (ok_frames && framed.size() > 20) {
      framed[18] ^= 1;
      size_t used = 0, made = 0;
      ok_frames = !upc_lz4::root_unzip_block(framed.data(), framed.size(), out.data(), out.size(), &used, &made);
      framed[18] ^= 1;
    }
  }
  std::ofstream(prefix + ".l4", std::ios::binary).write((const char*)framed.data(), (std::streamsize)framed.size());
  std::printf("{\"n\": %zu, \"xxh64\": \"%016llx\", \"block_bytes\": %zu, \"block_roundtrip\": %s, \"shrunk\": %s, \"frames_ok\": %s}\n",
              data.size(), (unsigned long long)upc_lz4::xxh64(data.data(), data.size()), blk.size(), ok_block ? "true" : "false",
              shrunk ? "true" : "false", ok_frames ? "true" : "false");
  return 0;
}
