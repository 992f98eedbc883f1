// Interface compatibility notice: the class, member and method names declared in this file reproduce the public
// interface of Upcgen (https://github.com/nburmaso/upcgen), Copyright (C) 2021-2025 Nazar Burmasov, Evgeny Kryshen,
// distributed under the GNU General Public License, version 3 or later (see LICENSE-UPCGEN-NOTICE.md at the
// repository root).  They are kept identical so that code written against the reference compiles against this
// drop-in; the implementation behind them is this project's own.
// UpcSampler1D / UpcSampler2D -- the reference's inverse-CDF samplers (include/UpcSampler.h) with the
// same constructors, operator() and getBinX/getBinY.  The cumulative table is built on the GPU in
// GSL's sequential order (upcgpu_hist_pdf_init) and draws are produced there in blocks from a
// Philox4x32-10 stream keyed by `seed` (the reference's private MT19937 stream is not
// reproduced; distributions are, and bin selection is bit-exact for equal uniforms).
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <vector>

#include "../../include/upcgpu.h"

namespace upc_host
{
// context used by free-standing samplers: the UpcCrossSection's one when registered, else a
// default-parameter context created on first use
upcgpu_ctx* samplerContext();
void registerSamplerContext(upcgpu_ctx* ctx);
// gRandom stand-in of the host-side event code (std::mt19937_64, as TRandomMT64)
void seedHost(uint64_t s);
double hostUniform(double a, double b);
} // namespace upc_host

class UpcSampler1D
{
 public:
  UpcSampler1D(const std::vector<double>& dist, const std::vector<double>& binEdges,
               uint64_t seed = std::numeric_limits<uint64_t>::max())
    : edges(binEdges), seed_(seed == std::numeric_limits<uint64_t>::max() ? 4357 : seed)
  {
    sum.resize(dist.size() + 1);
    int rc = upcgpu_hist_pdf_init(upc_host::samplerContext(), dist.data(), dist.size(), sum.data());
    if (rc) { std::fprintf(stderr, "Could not initialize 1D sampler (%d)\n", rc); std::abort(); }
  }

  double operator()() const
  {
    if (pos == buf.size()) refill();
    return buf[pos++];
  }

  std::vector<double> sum; // hpdf->sum of the reference

 private:
  void refill() const
  {
    const size_t n = 4096;
    std::vector<double> u(2 * n), u1(n);
    upcgpu_philox(seed_, counter, 0, n, u.data());
    for (size_t i = 0; i < n; i++) u1[i] = u[2 * i];
    counter += n;
    buf.resize(n);
    int rc = upcgpu_hist_sample1d(upc_host::samplerContext(), sum.data(), (int)sum.size() - 1, edges.data(), u1.data(),
                                  n, buf.data());
    if (rc) { std::fprintf(stderr, "1D sampler failed (%d)\n", rc); std::abort(); }
    pos = 0;
  }
  std::vector<double> edges;
  uint64_t seed_;
  mutable uint64_t counter{0};
  mutable std::vector<double> buf;
  mutable size_t pos{0};
};

class UpcSampler2D
{
 public:
  UpcSampler2D(const std::vector<std::vector<double>>& dist, const std::vector<double>& binEdgesX,
               const std::vector<double>& binEdgesY, uint64_t seed = std::numeric_limits<uint64_t>::max())
    : seed_(seed == std::numeric_limits<uint64_t>::max() ? 4357 : seed)
  {
    nBinsX = static_cast<int>(dist.size());
    nBinsY = static_cast<int>(dist[0].size());
    edgesX = binEdgesX;
    edgesY = binEdgesY;
    std::vector<double> flat((size_t)nBinsX * nBinsY);
    for (int i = 0; i < nBinsX; ++i)
      for (int j = 0; j < nBinsY; ++j) flat[(size_t)i * nBinsY + j] = dist[i][j]; // row-major, as gsl_histogram2d
    sum.resize(flat.size() + 1);
    int rc = upcgpu_hist_pdf_init(upc_host::samplerContext(), flat.data(), flat.size(), sum.data());
    if (rc) { std::fprintf(stderr, "Could not initialize 2D sampler (%d)\n", rc); std::abort(); }
  }

  void operator()(double& x, double& y) const
  {
    if (pos == bx.size()) refill();
    x = bx[pos];
    y = by[pos];
    ++pos;
  }

  // from 0 to nBinsX -- the reference's integer arithmetic, kept literally
  int getBinX(double x) { return int(nBinsX * (x - edgesX.front())) / (edgesX.back() - edgesX.front()); }
  int getBinY(double y) { return int(nBinsY * (y - edgesY.front())) / (edgesY.back() - edgesY.front()); }

  int nBinsX;
  int nBinsY;
  std::vector<double> edgesX;
  std::vector<double> edgesY;
  std::vector<double> sum; // hpdf->sum of the reference

 private:
  void refill() const
  {
    const size_t n = 4096;
    std::vector<double> u(2 * n);
    upcgpu_philox(seed_, counter, 0, n, u.data());
    counter += n;
    bx.resize(n);
    by.resize(n);
    int rc = upcgpu_hist_sample2d(upc_host::samplerContext(), sum.data(), nBinsX, nBinsY, edgesX.data(), edgesY.data(),
                                  u.data(), n, nullptr, bx.data(), by.data());
    if (rc) { std::fprintf(stderr, "2D sampler failed (%d)\n", rc); std::abort(); }
    pos = 0;
  }
  uint64_t seed_;
  mutable uint64_t counter{0};
  mutable std::vector<double> bx, by;
  mutable size_t pos{0};
};
