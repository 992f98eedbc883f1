// Physical constants of the generator (values as the reference's include/UpcPhysConstants.h:26-32).
#pragma once

namespace phys_consts
{
constexpr double alpha{1.0 / 137.035999074}; // fine structure constant
constexpr double hc{0.1973269718};           // GeV fm
constexpr double mProt{0.9382720813};
constexpr double mNeut{0.939565346};
constexpr double mEl{0.000510998946};
constexpr double mMu{0.1056583745};
constexpr double mTau{1.77686};
} // namespace phys_consts
