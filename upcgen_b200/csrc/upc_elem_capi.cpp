// upc_elem_capi.cpp -- C entry points over the host-side elementary-process plug-ins
// (upcgen_b200/host/UpcTwoPhoton*.cpp), for callers that cannot instantiate the C++ classes
// (the Python tests/bench).  The C++ facade calls the classes directly.
#include <memory>
#include <string>
#include <vector>

#include "../../include/upcgpu.h"
#include "../host/UpcPhysConstants.h"
#include "../host/UpcRootFile.h"
#include "../host/UpcRootHist.h"
#include "../host/UpcTwoPhotonALP.h"
#include "../host/UpcTwoPhotonDilep.h"
#include "../host/UpcTwoPhotonTabulated.h"

static std::unique_ptr<UpcElemProcess> make_process(int proc_id, double a_lep, double alp_mass, double alp_width)
{
  // UpcCrossSection::setElemProcess, reference src/UpcCrossSection.cpp:67-114
  switch (proc_id) {
    case 11:
    case 13:
    case 15: {
      auto p = std::make_unique<UpcTwoPhotonDilep>(proc_id);
      p->aLep = a_lep;
      return p;
    }
    case 51: return std::make_unique<UpcTwoPhotonALP>(alp_mass, alp_width);
    case 22:
    case 111: {
      // histograms from $UPCGEN_CROSS_SEC_DIR/{lbyl,pi0pi0} (no DO_M_CUT workaround through this entry point)
      std::unique_ptr<UpcTwoPhotonTabulated> p;
      if (proc_id == 22) p = std::make_unique<UpcTwoPhotonLbyL>(false, 0., 9999.);
      else p = std::make_unique<UpcTwoPhotonDipion>(false, 0., 9999.);
      if (!p->ok) return nullptr;
      return p;
    }
    default: return nullptr;
  }
}

extern "C" {

int upcgpu_elem_sigma_m(int proc_id, double a_lep, double alp_mass, double alp_width, int which, const double* m,
                        size_t n, double* out)
{
  auto p = make_process(proc_id, a_lep, alp_mass, alp_width);
  if (!p || !m || !out || which < 0 || which > 2) return UPCGPU_EINVAL;
  for (size_t i = 0; i < n; i++)
    out[i] = which == 0 ? p->calcCrossSectionM(m[i]) : which == 1 ? p->calcCrossSectionMPolS(m[i]) : p->calcCrossSectionMPolPS(m[i]);
  return UPCGPU_OK;
}

// UpcCrossSection::fillCrossSectionZM, reference src/UpcCrossSection.cpp:337-362
int upcgpu_elem_fill_cs_zm(int proc_id, double a_lep, double alp_mass, double alp_width, int flag, double zmin,
                           double zmax, int nz, double mmin, double mmax, int nm, double* out)
{
  auto p = make_process(proc_id, a_lep, alp_mass, alp_width);
  if (!p || !out || flag < 0 || flag > 2 || nz < 1 || nm < 1) return UPCGPU_EINVAL;
  constexpr double scalingFactor = phys_consts::hc * phys_consts::hc * 1e7; // to [nb]
  const double dm = (mmax - mmin) / nm;
  const double dz = (zmax - zmin) / nz;
  for (int im = 0; im < nm; ++im) {
    const double m = mmin + dm * im;
    for (int iz = 0; iz < nz; ++iz) {
      const double z = zmin + dz * iz;
      const double cs = flag == 0 ? p->calcCrossSectionZM(z, m) : flag == 1 ? p->calcCrossSectionZMPolS(z, m) : p->calcCrossSectionZMPolPS(z, m);
      out[(size_t)im * nz + iz] = cs * scalingFactor / dm;
    }
  }
  return UPCGPU_OK;
}

int upcgpu_root_hist_read(const char* path, const char* name, int* dim, int* nx, double* xlo, double* xhi, int* ny,
                          double* ylo, double* yhi, double* cells, size_t cap, size_t* n_cells)
{
  if (!path || !name) return UPCGPU_EINVAL;
  UpcRootHist h;
  std::string err;
  if (!h.Read(path, name, err)) return UPCGPU_EINVAL;
  if (dim) *dim = h.dim;
  if (nx) *nx = h.fXaxis.fNbins;
  if (xlo) *xlo = h.fXaxis.fXmin;
  if (xhi) *xhi = h.fXaxis.fXmax;
  if (ny) *ny = h.dim == 2 ? h.fYaxis.fNbins : 0;
  if (ylo) *ylo = h.fYaxis.fXmin;
  if (yhi) *yhi = h.fYaxis.fXmax;
  if (n_cells) *n_cells = h.fArray.size();
  if (cells) {
    if (cap < h.fArray.size()) return UPCGPU_EINVAL;
    for (size_t i = 0; i < h.fArray.size(); ++i) cells[i] = h.fArray[i];
  }
  return UPCGPU_OK;
}

// The writer side (upcgen_b200/host/UpcRootFile.cpp): the luminosity cache twoPhotonLumi[Pol].root as the reference
// writes it (n_hist TH2D objects: names[i], cells[i] = (nx + 2) * (ny + 2) doubles, x fastest) ...
int upcgpu_root_write_th2d(const char* path, int n_hist, const char* const* names, int nx, double xlo, double xhi, int ny,
                           double ylo, double yhi, const double* const* cells)
{
  if (!path || n_hist < 1 || !names || !cells || nx < 1 || ny < 1) return UPCGPU_EINVAL;
  try {
    UpcRootFileWriter w;
    const size_t nc = (size_t)(nx + 2) * (ny + 2);
    for (int i = 0; i < n_hist; ++i) {
      if (!names[i] || !cells[i]) return UPCGPU_EINVAL;
      std::vector<double> c(cells[i], cells[i] + nc);
      w.AddTH2D(names[i], "", nx, xlo, xhi, ny, ylo, yhi, c, (double)nx * ny);
    }
    std::string err;
    return w.Write(path, err) ? UPCGPU_OK : UPCGPU_EINVAL;
  } catch (...) {
    return UPCGPU_EINVAL;
  }
}

// ... a TH1D (the form of the reference's cross_sections/*/cross_section_m.root): cells = nx + 2 values, under-/overflow included
int upcgpu_root_write_th1d(const char* path, const char* name, int nx, double xlo, double xhi, const double* cells, double entries)
{
  if (!path || !name || !cells || nx < 1) return UPCGPU_EINVAL;
  try {
    UpcRootFileWriter w;
    w.AddTH1D(name, "", nx, xlo, xhi, std::vector<double>(cells, cells + nx + 2), entries);
    std::string err;
    return w.Write(path, err) ? UPCGPU_OK : UPCGPU_EINVAL;
  } catch (...) {
    return UPCGPU_EINVAL;
  }
}

// ... and a TTree of flat branches (events.root: tree "particles", src/UpcGenerator.cpp:842-857): n_cols columns of
// n_rows values each, types[i] = 'I' (Int_t) or 'D' (Double_t), integer columns passed as doubles
int upcgpu_root_write_tree(const char* path, const char* tree, const char* title, int n_cols, const char* const* names,
                           const char* types, const double* const* columns, size_t n_rows)
{
  if (!path || !tree || n_cols < 1 || !names || !types || !columns) return UPCGPU_EINVAL;
  try {
    UpcRootFileWriter w;
    std::vector<UpcRootFileWriter::Column> cols(n_cols);
    for (int i = 0; i < n_cols; ++i) {
      if (!names[i] || !columns[i]) return UPCGPU_EINVAL;
      cols[i].name = names[i];
      cols[i].type = types[i];
      cols[i].values.assign(columns[i], columns[i] + n_rows);
    }
    w.AddTree(tree, title ? title : "", cols);
    std::string err;
    return w.Write(path, err) ? UPCGPU_OK : UPCGPU_EINVAL;
  } catch (...) {
    return UPCGPU_EINVAL;
  }
}

int upcgpu_root_set_compression(int setting, int* previous)
{
  if (previous) *previous = UpcRootFileDefaultCompression();
  const int alg = setting / 100, level = setting % 100;
  if (setting != 0 && !((alg == 1 || alg == 4) && level >= 1 && level <= 9)) return UPCGPU_EINVAL;
  UpcRootFileDefaultCompression(setting);
  return UPCGPU_OK;
}

int upcgpu_root_write_sigma_hists(const char* path, int ny, const double* y_edges, int nm, const double* m_edges,
                                  const double* cs)
{
  if (!path || ny < 1 || nm < 1 || !y_edges || !m_edges || !cs) return UPCGPU_EINVAL;
  try {
    UpcRootFileWriter w;
    std::vector<std::vector<double>> table(ny);
    for (int iy = 0; iy < ny; ++iy) table[iy].assign(cs + (size_t)iy * nm, cs + (size_t)(iy + 1) * nm);
    UpcAddSigmaHists(w, std::vector<double>(y_edges, y_edges + ny + 1), std::vector<double>(m_edges, m_edges + nm + 1), table);
    std::string err;
    return w.Write(path, err) ? UPCGPU_OK : UPCGPU_EINVAL;
  } catch (...) {
    return UPCGPU_EINVAL;
  }
}

} // extern "C"
