// Host-only driver for the OrientationMap / dg::mortar_size shims.  Reads cases from stdin:
//   F d0 s0 d1 s1 d2 s2 direction          -> "neighbor_direction permutation"
//   M d0 s0 d1 s1 d2 s2 dimension  (self: lev idx)x3  (neighbor: lev idx)x3 -> "size_a size_b"
// (d_k s_k: dimension and sign the host's upper-k direction maps to; sizes as C-ABI codes)
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>

#include "../../spectre_b200/host/SpectreShims.hpp"

using namespace spectre_b200;

// "C n" then n lines "(lev idx)x3", then records "N element direction count nb... d0 s0 d1 s1 d2 s2"
// and "E element direction code", closed by "end": prints the four tables of DgConnectivity
static void connectivity() {
  size_t n;
  std::cin >> n;
  std::vector<ElementId<3>> ids;
  for (size_t e = 0; e < n; ++e) {
    std::array<std::pair<size_t, size_t>, 3> seg{};
    for (auto& s : seg) std::cin >> s.first >> s.second;
    ids.emplace_back(0, seg);
  }
  DgConnectivity conn(ids);
  std::string rec;
  while (std::cin >> rec && rec != "end") {
    size_t element;
    int direction;
    std::cin >> element >> direction;
    if (rec == "E") {
      long code;
      std::cin >> code;
      conn.set_external_boundary(element, Direction3::from_abi(direction), static_cast<int32_t>(code));
      continue;
    }
    size_t count;
    std::cin >> count;
    std::vector<size_t> nbs(count);
    for (auto& v : nbs) std::cin >> v;
    std::array<Direction3, 3> mapped{};
    for (auto& m : mapped) std::cin >> m.dimension >> m.sign;
    conn.set_neighbors(element, Direction3::from_abi(direction), nbs, OrientationMap<3>(mapped));
  }
  auto dump = [](const char* name, const std::vector<int32_t>& v) {
    std::printf("%s", name);
    for (auto x : v) std::printf(" %d", x);
    std::printf("\n");
  };
  dump("neighbors", conn.neighbors());
  dump("directions", conn.neighbor_directions());
  dump("permutations", conn.face_permutations());
  dump("mortars", conn.mortars());
  std::printf("aligned %d\n", conn.aligned() ? 1 : 0);
}

// "P n_elements world rank oriented boundary_slots n_mortars", then the global tables
// (neighbors [n][6]; if oriented: directions, permutations; mortar rows): prints DgPartition
static void partition() {
  long ne, nm;
  int world, rank, oriented, slots;
  std::cin >> ne >> world >> rank >> oriented >> slots >> nm;
  auto read = [](size_t count) {
    std::vector<int32_t> v(count);
    for (auto& x : v) {
      long long t;
      std::cin >> t;
      x = static_cast<int32_t>(t);
    }
    return v;
  };
  const auto nbr = read(6 * ne);
  std::vector<int32_t> dirs, perms;
  if (oriented) {
    dirs = read(6 * ne);
    perms = read(6 * ne);
  }
  const auto mortars = read(6 * nm);
  const DgPartition part(nbr, world, rank, oriented ? &dirs : nullptr, oriented ? &perms : nullptr, mortars,
                         slots != 0);
  auto dump = [](const char* name, const auto& v) {
    std::printf("%s", name);
    for (auto x : v) std::printf(" %d", static_cast<int>(x));
    std::printf("\n");
  };
  std::printf("counts %d %d %d %d\n", part.n_local(), part.n_interior(), part.n_recv(), part.n_ghost());
  dump("global_ids", part.global_ids());
  dump("neighbors", part.local_neighbors());
  dump("directions", part.local_neighbor_directions());
  dump("permutations", part.local_face_permutations());
  dump("mortars", part.local_mortars());
  dump("send_map", part.send_map());
  dump("send_counts", part.send_counts());
  dump("recv_counts", part.recv_counts());
  dump("external_faces", part.external_faces());
}

int main() {
  std::string kind;
  while (std::cin >> kind) {
    if (kind == "C") {
      connectivity();
      continue;
    }
    if (kind == "P") {
      partition();
      continue;
    }
    if (kind == "I") {  // "I d0 s0 d1 s1 d2 s2 (lev idx)x3" -> inverse map, mapped segment ids, is_aligned
      std::array<Direction3, 3> mapped{};
      for (auto& m : mapped) std::cin >> m.dimension >> m.sign;
      std::array<SegmentId, 3> seg{};
      for (auto& sg : seg) std::cin >> sg.refinement_level >> sg.index;
      const OrientationMap<3> o(mapped);
      const auto inv = o.inverse_map();
      for (size_t d = 0; d < 3; ++d) {
        const auto m = inv(Direction3{d, 1});
        std::printf("%zu %d ", m.dimension, m.sign);
      }
      for (const auto& sg : o(seg)) std::printf("%zu %zu ", sg.refinement_level, sg.index);
      std::printf("%d\n", o.is_aligned() ? 1 : 0);
      continue;
    }
    if (kind == "Z") {  // "Z (lev idx)x3" -> domain::z_curve_index
      std::array<std::pair<size_t, size_t>, 3> seg{};
      for (auto& sg : seg) std::cin >> sg.first >> sg.second;
      std::printf("%zu\n", domain::z_curve_index(ElementId<3>(0, seg)));
      continue;
    }
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
