// Host-only driver for the OrientationMap / dg::mortar_size shims.  Reads cases from stdin:
//   F d0 s0 d1 s1 d2 s2 direction          -> "neighbor_direction permutation"
//   M d0 s0 d1 s1 d2 s2 dimension  (self: lev idx)x3  (neighbor: lev idx)x3 -> "size_a size_b"
// (d_k s_k: dimension and sign the host's upper-k direction maps to; sizes as C-ABI codes)
#include <cstdio>
#include <iostream>
#include <string>

#include "../../spectre_b200/host/SpectreShims.hpp"

using namespace spectre_b200;

int main() {
  std::string kind;
  while (std::cin >> kind) {
    std::array<Direction3, 3> mapped{};
    for (auto& m : mapped) std::cin >> m.dimension >> m.sign;
    const OrientationMap<3> orientation(mapped);
    if (kind == "F") {
      int d;
      std::cin >> d;
      const auto fo = face_orientation(orientation, Direction3::from_abi(d));
      std::printf("%d %d\n", fo.neighbor_direction, fo.permutation);
    } else {
      size_t dimension;
      std::cin >> dimension;
      std::array<std::pair<size_t, size_t>, 3> self{}, nb{};
      for (auto& s : self) std::cin >> s.first >> s.second;
      for (auto& s : nb) std::cin >> s.first >> s.second;
      const auto sizes = dg::mortar_size(ElementId<3>(0, self), ElementId<3>(1, nb), dimension, orientation);
      std::printf("%d %d\n", Spectral::abi_size_code(sizes[0]), Spectral::abi_size_code(sizes[1]));
    }
  }
  return 0;
}
