// Host-pointer view of Deltaf_Data in the layout of csrc/dftables.cuh (used for the one evaluation at the surface
// averages).
#pragma once

#include <vector>

#include "../csrc/dftables.cuh"
#include "is3d_host.hpp"

namespace is3dhost {
struct HostDfView {
  is3d::DfTables tb;
  std::vector<double> spc[7];
  explicit HostDfView(const Deltaf_Data &d);
};
}  // namespace is3dhost
