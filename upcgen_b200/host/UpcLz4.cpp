// UpcLz4.cpp -- see UpcLz4.h.
#include "UpcLz4.h"

#include <cstring>

namespace upc_lz4
{
namespace
{
constexpr uint64_t P1 = 11400714785074694791ull, P2 = 14029467366897019727ull, P3 = 1609587929392839161ull,
                   P4 = 9650029242287828579ull, P5 = 2870177450012600261ull;
inline uint64_t rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
inline uint64_t rd64(const unsigned char* p)
{
  uint64_t v = 0;
  for (int i = 7; i >= 0; --i) v = (v << 8) | p[i];
  return v;
}
inline uint32_t rd32(const unsigned char* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
inline uint64_t round64(uint64_t acc, uint64_t in) { return rotl(acc + in * P2, 31) * P1; }
inline uint64_t merge64(uint64_t acc, uint64_t v) { return (acc ^ round64(0, v)) * P1 + P4; }

// token lengths: a nibble, then bytes of 255 while the rest is >= 255, then the remainder (0 included)
inline void put_len(std::vector<unsigned char>& d, size_t rest)
{
  while (rest >= 255) { d.push_back(255); rest -= 255; }
  d.push_back((unsigned char)rest);
}
}  // namespace

uint64_t xxh64(const unsigned char* p, size_t n, uint64_t seed)
{
  const unsigned char* const end = p + n;
  uint64_t h;
  if (n >= 32) {
    uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
    do {
      v1 = round64(v1, rd64(p));
      v2 = round64(v2, rd64(p + 8));
      v3 = round64(v3, rd64(p + 16));
      v4 = round64(v4, rd64(p + 24));
      p += 32;
    } while (p + 32 <= end);
    h = rotl(v1, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18);
    h = merge64(h, v1); h = merge64(h, v2); h = merge64(h, v3); h = merge64(h, v4);
  } else {
    h = seed + P5;
  }
  h += (uint64_t)n;
  while (p + 8 <= end) { h = rotl(h ^ round64(0, rd64(p)), 27) * P1 + P4; p += 8; }
  if (p + 4 <= end) { h = rotl(h ^ ((uint64_t)rd32(p) * P1), 23) * P2 + P3; p += 4; }
  while (p < end) { h = rotl(h ^ ((uint64_t)*p * P5), 11) * P1; ++p; }
  h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
  return h;
}

bool decompress_block(const unsigned char* src, size_t n, unsigned char* out, size_t out_n)
{
  size_t ip = 0, op = 0;
  while (ip < n) {
    const unsigned token = src[ip++];
    size_t lit = token >> 4;
    if (lit == 15) {
      unsigned b;
      do {
        if (ip >= n) return false;
        b = src[ip++];
        lit += b;
      } while (b == 255);
    }
    if (lit > n - ip || lit > out_n - op) return false;
    std::memcpy(out + op, src + ip, lit);
    ip += lit; op += lit;
    if (ip >= n) break;  // the last sequence is literals only
    if (n - ip < 2) return false;
    const size_t off = (size_t)src[ip] | (size_t)src[ip + 1] << 8;
    ip += 2;
    if (off == 0 || off > op) return false;
    size_t ml = token & 15;
    if (ml == 15) {
      unsigned b;
      do {
        if (ip >= n) return false;
        b = src[ip++];
        ml += b;
      } while (b == 255);
    }
    ml += 4;
    if (ml > out_n - op) return false;
    const unsigned char* m = out + op - off;  // may overlap the bytes being written: one at a time
    for (size_t i = 0; i < ml; ++i) out[op + i] = m[i];
    op += ml;
  }
  return op == out_n;
}

size_t compress_block(const unsigned char* src, size_t n, std::vector<unsigned char>& dst)
{
  const size_t start = dst.size();
  constexpr int kHashBits = 16;
  std::vector<uint32_t> table((size_t)1 << kHashBits, 0u);  // position + 1 of the last 4-byte word with this hash
  size_t anchor = 0, i = 0;
  auto emit = [&](size_t lit_from, size_t lit_n, size_t off, size_t ml) {  // ml = 0: final literals
    const size_t mcode = ml ? ml - 4 : 0;
    dst.push_back((unsigned char)((lit_n < 15 ? lit_n : 15) << 4 | (ml ? (mcode < 15 ? mcode : 15) : 0)));
    if (lit_n >= 15) put_len(dst, lit_n - 15);
    dst.insert(dst.end(), src + lit_from, src + lit_from + lit_n);
    if (ml) {
      dst.push_back((unsigned char)(off & 255)); dst.push_back((unsigned char)(off >> 8));
      if (mcode >= 15) put_len(dst, mcode - 15);
    }
  };
  // the format's end conditions: the last match starts at least 12 bytes before the end, the last 5 bytes are literals
  while (i + 12 <= n) {
    const uint32_t word = rd32(src + i);
    const uint32_t h = (word * 2654435761u) >> (32 - kHashBits);
    const uint32_t cand = table[h];
    table[h] = (uint32_t)i + 1;
    if (cand && i - (cand - 1) <= 65535 && rd32(src + cand - 1) == word) {
      const size_t ref = cand - 1;
      size_t ml = 4;
      while (i + ml < n - 5 && src[ref + ml] == src[i + ml]) ++ml;
      emit(anchor, i - anchor, i - ref, ml);
      i += ml;
      anchor = i;
    } else {
      ++i;
    }
  }
  emit(anchor, n - anchor, 0, 0);
  return dst.size() - start;
}

bool root_zip(const unsigned char* src, size_t n, std::vector<unsigned char>& dst)
{
  const size_t start = dst.size();
  constexpr size_t kMaxChunk = 0xffffff;
  for (size_t p = 0; p < n; p += kMaxChunk) {
    const size_t cn = n - p < kMaxChunk ? n - p : kMaxChunk;
    const size_t hdr = dst.size();
    dst.resize(hdr + 17);
    const size_t zn = compress_block(src + p, cn, dst);
    if (17 + zn >= cn) { dst.resize(start); return false; }
    const uint64_t sum = xxh64(dst.data() + hdr + 17, zn);
    unsigned char* h = dst.data() + hdr;
    const size_t csz = zn + 8;
    h[0] = 'L'; h[1] = '4'; h[2] = 1;
    h[3] = (unsigned char)csz; h[4] = (unsigned char)(csz >> 8); h[5] = (unsigned char)(csz >> 16);
    h[6] = (unsigned char)cn; h[7] = (unsigned char)(cn >> 8); h[8] = (unsigned char)(cn >> 16);
    for (int b = 0; b < 8; ++b) h[9 + b] = (unsigned char)(sum >> (56 - 8 * b));
  }
  if (dst.size() - start >= n) { dst.resize(start); return false; }
  return true;
}

bool root_unzip_block(const unsigned char* block, size_t avail, unsigned char* out, size_t out_cap, size_t* consumed,
                      size_t* produced)
{
  if (avail < 17 || block[0] != 'L' || block[1] != '4') return false;
  const size_t csz = (size_t)block[3] | (size_t)block[4] << 8 | (size_t)block[5] << 16;
  const size_t usz = (size_t)block[6] | (size_t)block[7] << 8 | (size_t)block[8] << 16;
  if (csz < 8 || 9 + csz > avail || usz > out_cap) return false;
  uint64_t stored = 0;
  for (int b = 0; b < 8; ++b) stored = (stored << 8) | block[9 + b];
  if (stored != xxh64(block + 17, csz - 8)) return false;
  if (!decompress_block(block + 17, csz - 8, out, usz)) return false;
  *consumed = 9 + csz;
  *produced = usz;
  return true;
}
}  // namespace upc_lz4
