// GSL shim (oracle/refshim): gsl_rng with the default generator mt19937, as include/UpcSampler.h uses it.
// GSL's mt19937 is the 2002 Matsumoto-Nishimura generator seeded by the Knuth recurrence
// state[i] = 1812433253 * (state[i-1] ^ (state[i-1] >> 30)) + i from the low 32 bits of the seed (seed 0 -> 4357),
// and gsl_rng_uniform returns get() / 2^32 -- exactly std::mt19937's stream.  TEST INFRASTRUCTURE.
#pragma once
#include <cstdint>
#include <random>

#include "../shim_tape.h"

struct gsl_rng_type { const char* name; };
static const gsl_rng_type gsl_rng_mt19937_shim = {"mt19937"};
static const gsl_rng_type* const gsl_rng_default = &gsl_rng_mt19937_shim;
struct gsl_rng { std::mt19937 eng{4357u}; };
inline gsl_rng* gsl_rng_alloc(const gsl_rng_type*) { return new gsl_rng(); }
inline void gsl_rng_free(gsl_rng* r) { delete r; }
inline void gsl_rng_set(gsl_rng* r, unsigned long s)
{
  uint32_t s32 = (uint32_t)(s & 0xffffffffUL);
  if (s32 == 0) s32 = 4357u;  // GSL: "the default seed is 4357"
  r->eng.seed(s32);
}
inline double gsl_rng_uniform(const gsl_rng* r)
{
  const double u = const_cast<gsl_rng*>(r)->eng() / 4294967296.0;
  shim_tape().put(u, 1);
  return u;
}
