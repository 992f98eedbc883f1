// Interface compatibility notice: the class, member and method names declared in this file reproduce the public
// interface of Upcgen (https://github.com/nburmaso/upcgen), Copyright (C) 2021-2025 Nazar Burmasov, Evgeny Kryshen,
// distributed under the GNU General Public License, version 3 or later (see LICENSE-UPCGEN-NOTICE.md at the
// repository root).  They are kept identical so that code written against the reference compiles against this
// drop-in; the implementation behind them is this project's own.
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
