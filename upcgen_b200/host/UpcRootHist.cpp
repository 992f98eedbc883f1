// UpcRootHist.cpp -- see UpcRootHist.h.  The file layout is ROOT's documented one (TFile / TKey / TBuffer streaming);
// nothing here links against or is copied from ROOT.
#include "UpcRootHist.h"
#include "UpcLz4.h"

#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace {

struct Cursor {
  const unsigned char* p;
  size_t n, pos = 0;
  Cursor(const unsigned char* p_, size_t n_) : p(p_), n(n_) {}
  void need(size_t k) const
  {
    if (pos + k > n) throw std::runtime_error("truncated buffer");
  }
  uint8_t u8() { need(1); return p[pos++]; }
  uint16_t u16() { need(2); uint16_t v = (uint16_t)(p[pos] << 8 | p[pos + 1]); pos += 2; return v; }
  uint32_t u32() { need(4); uint32_t v = (uint32_t)p[pos] << 24 | (uint32_t)p[pos + 1] << 16 | (uint32_t)p[pos + 2] << 8 | p[pos + 3]; pos += 4; return v; }
  int32_t i32() { return (int32_t)u32(); }
  uint64_t u64() { uint64_t hi = u32(); return hi << 32 | u32(); }
  double f64() { uint64_t b = u64(); double d; std::memcpy(&d, &b, 8); return d; }
  std::string str()
  {
    size_t len = u8();
    if (len == 255) len = u32();
    need(len);
    std::string s((const char*)p + pos, len);
    pos += len;
    return s;
  }
  // the (byte count | kByteCountMask, version) pair in front of every streamed object; returns the end offset
  size_t object(uint16_t* version = nullptr)
  {
    const uint32_t v = u32();
    if (!(v & 0x40000000u)) throw std::runtime_error("object without byte count");
    const size_t end = pos + (v & 0x3fffffffu);
    if (end > n) throw std::runtime_error("byte count beyond the buffer");
    const uint16_t ver = u16();
    if (version) *version = ver;
    return end;
  }
  void skip_object() { pos = object(); }
};

void read_axis(Cursor& c, UpcRootAxis& a)
{
  const size_t end = c.object();
  c.skip_object();  // TNamed
  c.skip_object();  // TAttAxis
  a.fNbins = c.i32();
  a.fXmin = c.f64();
  a.fXmax = c.f64();
  const int nedges = c.i32();  // TArrayD fXbins
  if (nedges < 0 || (nedges != 0 && nedges != a.fNbins + 1)) throw std::runtime_error("unexpected bin-edge array");
  a.fXbins.resize(nedges);
  for (int i = 0; i < nedges; ++i) a.fXbins[i] = c.f64();
  c.pos = end;  // fFirst, fLast, fBits2, time format, labels
}

void read_th1(Cursor& c, UpcRootHist& h, int& ncells)
{
  const size_t end = c.object();
  c.skip_object();  // TNamed
  c.skip_object();  // TAttLine
  c.skip_object();  // TAttFill
  c.skip_object();  // TAttMarker
  ncells = c.i32();
  read_axis(c, h.fXaxis);
  read_axis(c, h.fYaxis);
  read_axis(c, h.fZaxis);
  c.pos = end;  // bar offsets, statistics, contours, sumw2, option, function list, buffer
}

}  // namespace

int UpcRootAxis::FindBin(double x) const
{
  if (x < fXmin) return 0;
  if (!(x < fXmax)) return fNbins + 1;
  if (fXbins.empty()) return 1 + int(fNbins * (x - fXmin) / (fXmax - fXmin));
  // TMath::BinarySearch: index of the last edge <= x
  return 1 + int(std::upper_bound(fXbins.begin(), fXbins.end(), x) - fXbins.begin()) - 1;
}

bool UpcRootHist::Read(const std::string& path, const std::string& objName, std::string& err)
{
  std::vector<unsigned char> file;
  {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) { err = "cannot open " + path; return false; }
    std::fseek(f, 0, SEEK_END);
    const long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    file.resize(sz > 0 ? (size_t)sz : 0);
    const size_t got = file.empty() ? 0 : std::fread(file.data(), 1, file.size(), f);
    std::fclose(f);
    if (got != file.size() || file.size() < 64) { err = "cannot read " + path; return false; }
  }
  try {
    Cursor top(file.data(), file.size());
    if (std::memcmp(file.data(), "root", 4) != 0) throw std::runtime_error("not a ROOT file");
    top.pos = 4;
    const int32_t fVersion = top.i32();
    const bool big = fVersion >= 1000000;
    uint64_t fBEGIN = top.u32();
    uint64_t fEND = big ? top.u64() : top.u32();
    if (fEND > file.size()) fEND = file.size();
    // the chain of keys
    uint64_t pos = fBEGIN;
    struct Found { int cycle = -1; std::string cls; uint64_t pos = 0; int32_t nbytes = 0, objlen = 0; int keylen = 0; } best;
    while (pos + 18 <= fEND) {
      Cursor k(file.data(), file.size());
      k.pos = pos;
      const int32_t nbytes = k.i32();
      if (nbytes < 0) { pos += (uint64_t)(-(int64_t)nbytes); continue; }  // free segment
      if (nbytes == 0) break;
      const int16_t kver = (int16_t)k.u16();
      const int32_t objlen = k.i32();
      k.u32();  // date / time
      const int keylen = (int16_t)k.u16();
      const int cycle = (int16_t)k.u16();
      k.pos += kver > 1000 ? 16 : 8;  // seek key, seek parent directory
      const std::string cls = k.str(), nm = k.str();
      if (nm == objName && (cls == "TH1D" || cls == "TH2D") && cycle > best.cycle) {
        best.cycle = cycle; best.cls = cls; best.pos = pos; best.nbytes = nbytes; best.objlen = objlen; best.keylen = keylen;
      }
      pos += (uint64_t)nbytes;
    }
    if (best.cycle < 0) throw std::runtime_error("no TH1D/TH2D named " + objName);
    // these files are external input: the key must lie inside the file and hold its own header
    if (best.keylen <= 0 || best.keylen > best.nbytes || best.pos + (uint64_t)best.nbytes > file.size())
      throw std::runtime_error("truncated key of " + objName);
    if (best.objlen < 0 || (uint64_t)best.objlen > ((uint64_t)1 << 31))
      throw std::runtime_error("implausible object length of " + objName);
    // the object buffer, inflated if the key says so
    const unsigned char* raw = file.data() + best.pos + best.keylen;
    const size_t rawlen = (size_t)best.nbytes - best.keylen;
    std::vector<unsigned char> buf;
    if ((size_t)best.objlen > rawlen) {
      buf.resize(best.objlen);
      size_t q = 0, out = 0;
      while (out < buf.size()) {
        if (q + 9 > rawlen) throw std::runtime_error("truncated compressed object");
        if (raw[q] == 'L' && raw[q + 1] == '4') {  // LZ4 (ROOT 6.14-6.16 wrote it by default; the reference asks for it in events.root)
          size_t used = 0, made = 0;
          if (!upc_lz4::root_unzip_block(raw + q, rawlen - q, buf.data() + out, buf.size() - out, &used, &made))
            throw std::runtime_error("LZ4 block: bad sizes, checksum or stream");
          out += made;
          q += used;
          continue;
        }
        if (raw[q] != 'Z' || raw[q + 1] != 'L')
          throw std::runtime_error(std::string("compression '") + (char)raw[q] + (char)raw[q + 1] + "' is not supported (zlib ZL and LZ4 L4 are)");
        const size_t csz = raw[q + 3] | raw[q + 4] << 8 | raw[q + 5] << 16;
        const size_t usz = raw[q + 6] | raw[q + 7] << 8 | raw[q + 8] << 16;
        if (q + 9 + csz > rawlen || out + usz > buf.size()) throw std::runtime_error("inconsistent compressed block");
        uLongf dl = (uLongf)usz;
        if (uncompress(buf.data() + out, &dl, raw + q + 9, (uLong)csz) != Z_OK || dl != usz) throw std::runtime_error("zlib failure");
        out += usz;
        q += 9 + csz;
      }
    } else {
      buf.assign(raw, raw + rawlen);
    }
    // TH1D = TH1 + TArrayD; TH2D = TH2 (TH1 + scale factor and y statistics) + TArrayD
    Cursor c(buf.data(), buf.size());
    const size_t end = c.object();
    int ncells = 0;
    dim = best.cls == "TH2D" ? 2 : 1;
    if (dim == 2) {
      const size_t end_th2 = c.object();
      read_th1(c, *this, ncells);
      c.pos = end_th2;
    } else {
      read_th1(c, *this, ncells);
    }
    const int n = c.i32();
    const long long expect = (long long)(fXaxis.fNbins + 2) * (dim == 2 ? fYaxis.fNbins + 2 : 1);
    if (n != ncells || (long long)n != expect) throw std::runtime_error("cell array does not match the axes");
    fArray.resize(n);
    for (int i = 0; i < n; ++i) fArray[i] = c.f64();
    if (c.pos != end) throw std::runtime_error("trailing bytes after the cell array");
    name = objName;
  } catch (const std::exception& e) {
    err = path + ": " + e.what();
    return false;
  }
  return true;
}

double UpcRootHist::GetBinContent(int bin) const
{
  if (fArray.empty()) return 0.;
  bin = std::max(0, std::min(bin, (int)fArray.size() - 1));
  return fArray[bin];
}

double UpcRootHist::GetBinContent(int binx, int biny) const
{
  const int nx = fXaxis.fNbins + 2, ny = fYaxis.fNbins + 2;
  binx = std::max(0, std::min(binx, nx - 1));
  biny = std::max(0, std::min(biny, ny - 1));
  return GetBinContent(binx + nx * biny);
}

void UpcRootHist::SetBinContent(int bin, double v)
{
  if (bin >= 0 && bin < (int)fArray.size()) fArray[bin] = v;
}

void UpcRootHist::SetBinContent(int binx, int biny, double v)
{
  const int nx = fXaxis.fNbins + 2, ny = fYaxis.fNbins + 2;
  if (binx < 0 || binx >= nx || biny < 0 || biny >= ny) return;
  SetBinContent(binx + nx * biny, v);
}
