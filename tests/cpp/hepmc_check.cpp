// The HepMC writer of the drop-in (upcgen_b200/host/UpcGenerator.h: WriterHepMC, formatted with std::to_chars) against
// the reference's formatting of the same records (include/UpcGenerator.h:214-241: operator<< with setprecision(9)):
// the two files must be byte-identical, special values included.  Prints both timings.
//   hepmc_check <n events> <dir>
#include <charconv>
#include <cstring>
#include <cstdint>
#include <fstream>
#include <iomanip>
#include <string>
#include <vector>
#include <sstream>
#include <iostream>
#include <random>
#include <chrono>
#include <cmath>
#include <map>
#include <algorithm>
#define private public
#include "UpcGenerator.h"
#undef private
#include <chrono>
#include <random>
#include <sstream>
#include <iostream>
int main(int argc, char** argv)
{
  const long n = argc > 1 ? atol(argv[1]) : 1000000;
  const std::string dir = argc > 2 ? argv[2] : ".";
  std::mt19937_64 rng(1);
  std::uniform_real_distribution<double> u(-50, 50);
  auto t0 = std::chrono::steady_clock::now();
  {
    UpcGenerator::WriterHepMC w(dir + "/fast.hepmc");
    for (long i = 0; i < n; i++) {
      w.writeEventInfo(i, 2, 0);
      for (int j = 0; j < 2; j++) {
        double px = u(rng), py = u(rng), pz = u(rng) * 1e3, e = std::fabs(u(rng)) * 1e-7;
        if (i % 1000 == 7) { px = 0.0; py = -0.0; pz = 1e300; e = 5e-324; }
        if (i % 1000 == 8) { px = 123456789.0; py = 1234567890.0; pz = 0.0001; e = 0.00001; }
        if (i % 1000 == 9) { px = INFINITY; py = -INFINITY; pz = NAN; e = 999999999.5; }
        w.writeParticleInfo(j + 1, 0, j ? -13 : 13, px, py, pz, e, 0.1056583745, 23);
      }
    }
  }
  auto t1 = std::chrono::steady_clock::now();
  // the reference's formatting, for comparison
  rng.seed(1);
  {
    std::ofstream o(dir + "/ref.hepmc");
    o << "HepMC::Version 3.02.04\n" << "HepMC::Asciiv3-START_EVENT_LISTING\n";
    for (long i = 0; i < n; i++) {
      o << "E " << i << " " << 0 << " " << 2 << "\n" << "U GEV MM\n";
      for (int j = 0; j < 2; j++) {
        double px = u(rng), py = u(rng), pz = u(rng) * 1e3, e = std::fabs(u(rng)) * 1e-7;
        if (i % 1000 == 7) { px = 0.0; py = -0.0; pz = 1e300; e = 5e-324; }
        if (i % 1000 == 8) { px = 123456789.0; py = 1234567890.0; pz = 0.0001; e = 0.00001; }
        if (i % 1000 == 9) { px = INFINITY; py = -INFINITY; pz = NAN; e = 999999999.5; }
        o << std::setprecision(9) << "P " << j + 1 << " " << 0 << " " << (j ? -13 : 13) << " " << px << " " << py << " "
          << pz << " " << e << " " << 0.1056583745 << " " << 23 << "\n";
      }
    }
    o << "HepMC::Asciiv3-END_EVENT_LISTING\n";
  }
  auto t2 = std::chrono::steady_clock::now();
  std::cout << "fast " << std::chrono::duration<double>(t1 - t0).count() << " s, stream " << std::chrono::duration<double>(t2 - t1).count() << " s\n";
}
