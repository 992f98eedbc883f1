// Tape of the uniforms the reference's code draws (oracle/refshim): when switched on, every gsl_rng_uniform and
// TRandom::Rndm value is appended with the stream it came from, so that a test can replay an event of the
// reference's generateEvent through the oracle with exactly the same uniforms.  TEST INFRASTRUCTURE.
#pragma once
#include <vector>

struct ShimTape {
  bool on = false;
  std::vector<double> v;
  std::vector<int> tag;  // 0: gRandom (TRandom::Rndm), 1: a gsl_rng stream
  void put(double x, int t) { if (on) { v.push_back(x); tag.push_back(t); } }
};
inline ShimTape& shim_tape() { static ShimTape t; return t; }
