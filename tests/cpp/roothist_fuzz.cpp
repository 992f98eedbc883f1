// Reads every file named on the command line with the ROOT-less histogram reader (UpcRootHist.cpp); built with
// -fsanitize=address,undefined by tests/test_root_hist.py::test_reader_survives_corrupted_files, which feeds it
// truncated and bit-flipped copies of uncompressed, zlib and LZ4 files: errors are fine, memory faults are not.
#include "UpcRootHist.h"
#include <cstdio>
int main(int argc, char** argv) {
  int ok = 0, bad = 0;
  for (int i = 1; i < argc; i++) { UpcRootHist h; std::string err; if (h.Read(argv[i], "hD2LDMDY", err)) { ok++; volatile double v = h.GetBinContent(3, 3); (void)v; } else bad++; }
  std::printf("asan pass: ok %d bad %d\n", ok, bad);
}
