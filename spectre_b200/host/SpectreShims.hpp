// SpectreShims.hpp -- C++ host-side mirror of the reference's operator surface
// for the DG right-hand-side path, forwarding to the C-ABI in include/dgrhs.h.
//
// Same names, argument order and meaning as the reference (v2024.09.29):
//   DataVector                         DataStructures/DataVector.hpp:48, VectorImpl.hpp:222-322
//   Scalar, tnsr::{i,I,ii,II,a,ab,aa,iaa,ijaa}   DataStructures/Tensor/TypeAliases.hpp,
//                                      storage order Tensor/Structure.hpp:162-194
//   Mesh<3>                            NumericalAlgorithms/Spectral/Mesh.hpp:52-244
//   ElementId<3>                       Domain/Structure/ElementId.hpp:29-110
//   partial_derivatives                NumericalAlgorithms/LinearOperators/PartialDerivatives.hpp
//   ScalarWave::TimeDerivative<3>      Evolution/Systems/ScalarWave/TimeDerivative.hpp:26-50
//   gh::TimeDerivative<3>              Evolution/Systems/GeneralizedHarmonic/TimeDerivative.hpp:143-192
//   {ScalarWave,gh}::BoundaryCorrections::UpwindPenalty<3>   .../BoundaryCorrections/UpwindPenalty.hpp
//   dg::lift_flux                      NumericalAlgorithms/DiscontinuousGalerkin/LiftFlux.hpp:41-62
//   TimeSteppers::AdamsBashforth       Time/TimeSteppers/AdamsBashforth.hpp (coefficients only)
//
// Error behaviour: the reference ERRORs (aborts with a message); here a non-zero
// C-ABI status becomes a std::runtime_error carrying dgrhs_last_error().
// The header is self-contained (no Blaze/Charm++); a maintainer of the
// reference would keep SpECTRE's own types and only take the bodies, see
// INTEGRATION.md.
#pragma once

#include <algorithm>
#include <array>
#include <cstddef>
#include <cstdint>
#include <map>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/dgrhs.h"

namespace spectre_b200 {

inline void check(int status) {
  if (status != 0) throw std::runtime_error(dgrhs_last_error());
}

// ---- DataVector: owning or non-owning span of doubles ------------------------
class DataVector {
 public:
  DataVector() = default;
  explicit DataVector(size_t n, double v = 0.0) : owned_(n, v), data_(owned_.data()), size_(n) {}
  DataVector(double* p, size_t n) : data_(p), size_(n) {}  // non-owning view
  DataVector(const DataVector& o) : owned_(o.data_, o.data_ + o.size_), data_(owned_.data()), size_(o.size_) {}
  DataVector& operator=(const DataVector& o) {
    if (this == &o) return *this;
    if (is_owning() || data_ == nullptr) {
      owned_.assign(o.data_, o.data_ + o.size_);
      data_ = owned_.data();
      size_ = o.size_;
    } else {  // assignment through a view writes into the viewed memory
      if (size_ != o.size_) throw std::runtime_error("DataVector view size mismatch");
      for (size_t i = 0; i < size_; ++i) data_[i] = o.data_[i];
    }
    return *this;
  }
  void set_data_ref(double* p, size_t n) {
    owned_.clear();
    data_ = p;
    size_ = n;
  }
  bool is_owning() const { return !owned_.empty() && data_ == owned_.data(); }
  size_t size() const { return size_; }
  double* data() { return data_; }
  const double* data() const { return data_; }
  double& operator[](size_t i) { return data_[i]; }
  const double& operator[](size_t i) const { return data_[i]; }

 private:
  std::vector<double> owned_;
  double* data_ = nullptr;
  size_t size_ = 0;
};

// ---- tensors (independent components in the reference's storage order) -------
constexpr size_t sym4(size_t a, size_t b) {
  return a <= b ? a * 4 - a * (a - 1) / 2 + (b - a) : b * 4 - b * (b - 1) / 2 + (a - b);
}
constexpr size_t sym3(size_t a, size_t b) {
  return a <= b ? a * 3 - a * (a - 1) / 2 + (b - a) : b * 3 - b * (b - 1) / 2 + (a - b);
}

template <size_t K>
struct TensorBase {
  std::array<DataVector, K> c;
  TensorBase() = default;
  explicit TensorBase(size_t n, double v = 0.0) {
    for (auto& x : c) x = DataVector(n, v);
  }
  static constexpr size_t size() { return K; }
  DataVector& operator[](size_t s) { return c[s]; }
  const DataVector& operator[](size_t s) const { return c[s]; }
  auto begin() { return c.begin(); }
  auto end() { return c.end(); }
  auto begin() const { return c.begin(); }
  auto end() const { return c.end(); }
  // gather into / scatter from a contiguous [K][n] block (Variables layout)
  std::vector<double> flat() const {
    const size_t n = c[0].size();
    std::vector<double> out(K * n);
    for (size_t s = 0; s < K; ++s)
      for (size_t p = 0; p < n; ++p) out[s * n + p] = c[s][p];
    return out;
  }
  void from_flat(const double* in, size_t n) {
    for (size_t s = 0; s < K; ++s) {
      if (c[s].size() != n) c[s] = DataVector(n);
      for (size_t p = 0; p < n; ++p) c[s][p] = in[s * n + p];
    }
  }
};

struct ScalarDV : TensorBase<1> {
  using TensorBase<1>::TensorBase;
  DataVector& get() { return c[0]; }
  const DataVector& get() const { return c[0]; }
};
template <typename T = DataVector>
using Scalar = ScalarDV;
inline DataVector& get(ScalarDV& s) { return s.get(); }
inline const DataVector& get(const ScalarDV& s) { return s.get(); }

namespace tnsr {
struct i3 : TensorBase<3> {  // tnsr::i / tnsr::I
  using TensorBase<3>::TensorBase;
  DataVector& get(size_t i) { return c[i]; }
  const DataVector& get(size_t i) const { return c[i]; }
};
struct a4 : TensorBase<4> {  // tnsr::a / tnsr::A
  using TensorBase<4>::TensorBase;
  DataVector& get(size_t a) { return c[a]; }
  const DataVector& get(size_t a) const { return c[a]; }
};
struct ij9 : TensorBase<9> {  // tnsr::ij, first index fastest
  using TensorBase<9>::TensorBase;
  DataVector& get(size_t i, size_t j) { return c[i + 3 * j]; }
  const DataVector& get(size_t i, size_t j) const { return c[i + 3 * j]; }
};
struct ii6 : TensorBase<6> {  // tnsr::ii / tnsr::II
  using TensorBase<6>::TensorBase;
  DataVector& get(size_t i, size_t j) { return c[sym3(i, j)]; }
  const DataVector& get(size_t i, size_t j) const { return c[sym3(i, j)]; }
};
struct ab16 : TensorBase<16> {  // tnsr::ab, first index fastest
  using TensorBase<16>::TensorBase;
  DataVector& get(size_t a, size_t b) { return c[a + 4 * b]; }
  const DataVector& get(size_t a, size_t b) const { return c[a + 4 * b]; }
};
struct aa10 : TensorBase<10> {  // tnsr::aa / tnsr::AA
  using TensorBase<10>::TensorBase;
  DataVector& get(size_t a, size_t b) { return c[sym4(a, b)]; }
  const DataVector& get(size_t a, size_t b) const { return c[sym4(a, b)]; }
};
struct iaa30 : TensorBase<30> {  // tnsr::iaa
  using TensorBase<30>::TensorBase;
  DataVector& get(size_t i, size_t a, size_t b) { return c[i + 3 * sym4(a, b)]; }
  const DataVector& get(size_t i, size_t a, size_t b) const { return c[i + 3 * sym4(a, b)]; }
};
struct ijaa90 : TensorBase<90> {  // tnsr::ijaa
  using TensorBase<90>::TensorBase;
  DataVector& get(size_t i, size_t j, size_t a, size_t b) { return c[i + 3 * (j + 3 * sym4(a, b))]; }
  const DataVector& get(size_t i, size_t j, size_t a, size_t b) const {
    return c[i + 3 * (j + 3 * sym4(a, b))];
  }
};
template <typename T = DataVector, size_t Dim = 3, typename Fr = void> using i = i3;
template <typename T = DataVector, size_t Dim = 3, typename Fr = void> using I = i3;
template <typename T = DataVector, size_t Dim = 3, typename Fr = void> using a = a4;
template <typename T = DataVector, size_t Dim = 3, typename Fr = void> using ij = ij9;
template <typename T = DataVector, size_t Dim = 3, typename Fr = void> using ii = ii6;
template <typename T = DataVector, size_t Dim = 3, typename Fr = void> using II = ii6;
template <typename T = DataVector, size_t Dim = 3, typename Fr = void> using ab = ab16;
template <typename T = DataVector, size_t Dim = 3, typename Fr = void> using aa = aa10;
template <typename T = DataVector, size_t Dim = 3, typename Fr = void> using iaa = iaa30;
template <typename T = DataVector, size_t Dim = 3, typename Fr = void> using ijaa = ijaa90;
}  // namespace tnsr

// InverseJacobian<DataVector, 3, ElementLogical, Inertial>: get(ihat, i) at ihat + 3 i
using InverseJacobian3 = tnsr::ij9;

// ---- Mesh<3>, ElementId<3> -------------------------------------------------------
namespace Spectral {
enum class Basis : uint8_t { Legendre = 1 };
enum class Quadrature : uint8_t { GaussLobatto = 2 };
// Spectral::differentiation_matrix / collocation_points / quadrature_weights
inline std::vector<double> differentiation_matrix(size_t n) {
  std::vector<double> D(n * n);
  check(dgrhs_differentiation_matrix(static_cast<int>(n), D.data()));
  return D;  // row-major D[i * n + j]
}
inline std::vector<double> collocation_points(size_t n) {
  std::vector<double> x(n), w(n);
  check(dgrhs_collocation_points_and_weights(static_cast<int>(n), x.data(), w.data()));
  return x;
}
inline std::vector<double> quadrature_weights(size_t n) {
  std::vector<double> x(n), w(n);
  check(dgrhs_collocation_points_and_weights(static_cast<int>(n), x.data(), w.data()));
  return w;
}
// Spectral::ChildSize / MortarSize and the projection matrices of non-conforming
// mortars (NumericalAlgorithms/Spectral/Projection.hpp:28-36, Projection.cpp:57-362),
// row-major [target point][source point]
enum class ChildSize : uint8_t { Uninitialized = 0, Full = 1, UpperHalf = 2, LowerHalf = 3 };
using MortarSize = ChildSize;
inline int abi_size_code(ChildSize s) {
  switch (s) {
    case ChildSize::Full: return 0;
    case ChildSize::LowerHalf: return 1;
    case ChildSize::UpperHalf: return 2;
    default: throw std::runtime_error("Received uninitialized child_size.");
  }
}
inline std::vector<double> projection_matrix_parent_to_child(size_t n, ChildSize size) {
  std::vector<double> M(n * n);
  check(dgrhs_projection_matrix(static_cast<int>(n), 0, abi_size_code(size), M.data()));
  return M;
}
inline std::vector<double> projection_matrix_child_to_parent(size_t n, ChildSize size) {
  std::vector<double> M(n * n);
  check(dgrhs_projection_matrix(static_cast<int>(n), 1, abi_size_code(size), M.data()));
  return M;
}
// meshes with different extents (the reference's (parent_mesh, child_mesh, size)
// arguments reduced to the two point counts): [n_child][n_parent] and [n_parent][n_child]
inline std::vector<double> projection_matrix_parent_to_child(size_t n_parent, size_t n_child, ChildSize size) {
  std::vector<double> M(n_parent * n_child);
  check(dgrhs_projection_matrix_meshes(static_cast<int>(n_parent), static_cast<int>(n_child), 0,
                                       abi_size_code(size), M.data()));
  return M;
}
inline std::vector<double> projection_matrix_child_to_parent(size_t n_parent, size_t n_child, ChildSize size) {
  std::vector<double> M(n_parent * n_child);
  check(dgrhs_projection_matrix_meshes(static_cast<int>(n_parent), static_cast<int>(n_child), 1,
                                       abi_size_code(size), M.data()));
  return M;
}
}  // namespace Spectral

template <size_t Dim>
class Mesh {
 public:
  Mesh() = default;
  Mesh(size_t isotropic_extents, Spectral::Basis b, Spectral::Quadrature q) : basis_(b), quadrature_(q) {
    extents_.fill(static_cast<uint8_t>(isotropic_extents));
  }
  Mesh(const std::array<size_t, Dim>& extents, Spectral::Basis b, Spectral::Quadrature q)
      : basis_(b), quadrature_(q) {
    for (size_t d = 0; d < Dim; ++d) extents_[d] = static_cast<uint8_t>(extents[d]);
  }
  size_t extents(size_t d) const { return extents_[d]; }
  size_t number_of_grid_points() const {
    size_t n = 1;
    for (auto e : extents_) n *= e;
    return n;
  }
  bool is_isotropic() const {
    for (auto e : extents_)
      if (e != extents_[0]) return false;
    return true;
  }

 private:
  std::array<uint8_t, Dim> extents_{};
  Spectral::Basis basis_ = Spectral::Basis::Legendre;
  Spectral::Quadrature quadrature_ = Spectral::Quadrature::GaussLobatto;
};

template <size_t Dim>
class ElementId {
 public:
  ElementId(size_t block, std::array<std::pair<size_t, size_t>, Dim> level_and_index, size_t grid = 0) {
    bits_ = (block & 0xFF) | ((grid & 0xF) << 8);
    size_t shift = 16;
    for (auto [lev, idx] : level_and_index) {
      bits_ |= static_cast<uint64_t>(idx & 0xFFF) << shift;
      bits_ |= static_cast<uint64_t>(lev & 0xF) << (shift + 12);
      shift += 16;
    }
  }
  uint64_t bits() const { return bits_; }
  size_t block_id() const { return bits_ & 0xFF; }
  size_t refinement_level(size_t d) const { return (bits_ >> (16 * (d + 1) + 12)) & 0xF; }
  size_t index(size_t d) const { return (bits_ >> (16 * (d + 1))) & 0xFFF; }

 private:
  uint64_t bits_ = 0;
};

namespace domain {
// domain::z_curve_index (Domain/Structure/ZCurve.cpp:17-80): position of the element on the
// Morton curve of its block; dimensions are interleaved from the least refined one up, bits of
// a dimension stop once its refinement level is used up.  Elements are handed to a context, and
// cut into per-rank chunks by DgPartition, in (block, z_curve_index) order.
inline size_t z_curve_index(const ElementId<3>& id) {
  std::array<std::pair<size_t, size_t>, 3> dims{};   // (refinement level, dimension), ascending
  for (size_t d = 0; d < 3; ++d) dims[d] = {id.refinement_level(d), d};
  std::sort(dims.begin(), dims.end());
  size_t out = 0, leading_gap = 0;
  for (size_t i = 0; i < 3; ++i) {
    const auto [level, dim] = dims[i];
    size_t total_gap = leading_gap;
    if (level > 0) ++leading_gap;
    for (size_t bit = 0; bit < level; ++bit) {
      out |= (id.index(dim) & (size_t{1} << bit)) << total_gap;
      for (size_t j = 0; j < 3; ++j)
        if (i != j && bit + 1 < dims[j].first) ++total_gap;
    }
  }
  return out;
}
}  // namespace domain

// ---- Direction<3>, SegmentId, OrientationMap<3>, dg::mortar_size ----------------------
// What a caller needs to turn Element<3>::neighbors() (ids + OrientationMap per direction,
// Domain/Structure/{Direction,SegmentId,OrientationMap}.hpp) into the tables of
// dgrhs_set_geometry / dgrhs_set_neighbor_orientations / dgrhs_set_mortars.
struct Direction3 {
  size_t dimension = 0;
  int sign = 1;  // +1 Side::Upper, -1 Side::Lower
  Direction3 opposite() const { return {dimension, -sign}; }
  int abi() const { return static_cast<int>(2 * dimension) + (sign > 0 ? 1 : 0); }  // the C-ABI's d
  static Direction3 from_abi(int d) { return {static_cast<size_t>(d / 2), (d & 1) ? 1 : -1}; }
  bool operator==(const Direction3& o) const { return dimension == o.dimension && sign == o.sign; }
};
struct SegmentId {
  size_t refinement_level = 0, index = 0;
  SegmentId id_if_flipped() const { return {refinement_level, (size_t{1} << refinement_level) - 1 - index}; }
};
template <size_t Dim>
class OrientationMap;
template <>
class OrientationMap<3> {
 public:
  OrientationMap() : mapped_{Direction3{0, 1}, Direction3{1, 1}, Direction3{2, 1}} {}
  // mapped_directions[d]: where the host block's upper-d direction points in the neighbour's frame
  explicit OrientationMap(std::array<Direction3, 3> mapped_directions) : mapped_(mapped_directions) {
    bool seen[3] = {false, false, false};
    for (const auto& m : mapped_) {
      if (m.dimension > 2 || seen[m.dimension]) throw std::runtime_error("This OrientationMap fails to map Directions one-to-one.");
      seen[m.dimension] = true;
    }
  }
  bool is_aligned() const {
    for (size_t d = 0; d < 3; ++d)
      if (mapped_[d].dimension != d || mapped_[d].sign < 0) return false;
    return true;
  }
  Direction3 operator()(const Direction3& host) const {
    return {mapped_[host.dimension].dimension, mapped_[host.dimension].sign * host.sign};
  }
  std::array<SegmentId, 3> operator()(const std::array<SegmentId, 3>& host_segments) const {
    std::array<SegmentId, 3> out{};
    for (size_t d = 0; d < 3; ++d)
      out[mapped_[d].dimension] = mapped_[d].sign > 0 ? host_segments[d] : host_segments[d].id_if_flipped();
    return out;
  }
  OrientationMap inverse_map() const {
    std::array<Direction3, 3> inv{};
    for (size_t d = 0; d < 3; ++d) inv[mapped_[d].dimension] = {d, mapped_[d].sign};
    return OrientationMap(inv);
  }

 private:
  std::array<Direction3, 3> mapped_;
};

// The per-face entries of dgrhs_set_neighbor_orientations for the neighbour reached through
// `direction`: the neighbour's direction that points back, and the permutation of the two
// face coordinates (bit 0: swapped; bit 1 / bit 2: first / second neighbour face coordinate
// runs backwards) -- orient_variables_on_slice restricted to the face
// (OrientationMapHelpers.cpp:25-120).  A mortar row between non-aligned blocks carries
// neighbor_direction | permutation << 3 with the orientation seen from the coarse element.
struct FaceOrientation {
  int neighbor_direction;
  int permutation;
};
inline FaceOrientation face_orientation(const OrientationMap<3>& orientation, const Direction3& direction) {
  size_t tangential[2], k = 0;
  for (size_t d = 0; d < 3; ++d)
    if (d != direction.dimension) tangential[k++] = d;
  const Direction3 ma = orientation(Direction3{tangential[0], 1});
  const Direction3 mb = orientation(Direction3{tangential[1], 1});
  const bool swapped = ma.dimension > mb.dimension;
  const bool reverse_first = (swapped ? mb.sign : ma.sign) < 0;
  const bool reverse_second = (swapped ? ma.sign : mb.sign) < 0;
  return {orientation(direction).opposite().abi(),
          (swapped ? 1 : 0) | (reverse_first ? 2 : 0) | (reverse_second ? 4 : 0)};
}

// orient_variables_on_slice (OrientationMapHelpers.cpp:25-120): variables on the face
// perpendicular to `sliced_dim` of this element, [n_comps][slice points], into the frame of the
// neighbour reached through `orientation_of_neighbor`
inline std::vector<double> orient_variables_on_slice(const std::vector<double>& variables_on_slice, size_t n_comps,
                                                     const std::array<size_t, 2>& slice_extents, size_t sliced_dim,
                                                     const OrientationMap<3>& orientation_of_neighbor) {
  const auto fo = face_orientation(orientation_of_neighbor, Direction3{sliced_dim, 1});
  const int ext[2] = {static_cast<int>(slice_extents[0]), static_cast<int>(slice_extents[1])};
  std::vector<double> out(variables_on_slice.size());
  check(dgrhs_orient_variables_on_slice(static_cast<int>(n_comps), ext, fo.permutation, variables_on_slice.data(),
                                        out.data()));
  return out;
}

namespace dg {
// dg::mortar_size (NumericalAlgorithms/DiscontinuousGalerkin/MortarHelpers.cpp:51-77): size of
// the mortar to `neighbor` inside the face of `self` perpendicular to `dimension`, per face
// dimension of `self` (the two size entries of a dgrhs_set_mortars row when `self` is the
// coarse element; Full everywhere when `self` is the fine one)
inline std::array<Spectral::MortarSize, 2> mortar_size(const ElementId<3>& self, const ElementId<3>& neighbor,
                                                       size_t dimension, const OrientationMap<3>& orientation) {
  std::array<SegmentId, 3> theirs{};
  for (size_t d = 0; d < 3; ++d) theirs[d] = {neighbor.refinement_level(d), neighbor.index(d)};
  const auto in_my_frame = orientation.inverse_map()(theirs);
  std::array<Spectral::MortarSize, 2> out{};
  size_t k = 0;
  for (size_t d = 0; d < 3; ++d) {
    if (d == dimension) continue;
    const long diff = static_cast<long>(in_my_frame[d].refinement_level) - static_cast<long>(self.refinement_level(d));
    if (diff <= 0) {
      out[k++] = Spectral::MortarSize::Full;
    } else if (diff == 1) {
      out[k++] = in_my_frame[d].index % 2 == 0 ? Spectral::MortarSize::LowerHalf : Spectral::MortarSize::UpperHalf;
    } else {
      throw std::runtime_error("neighbours may differ by at most one refinement level (2:1 balance)");
    }
  }
  return out;
}
}  // namespace dg

// ---- partial_derivatives ---------------------------------------------------------
// du[3 c + i] = d_i u_c for a block of n_comps components (Variables layout)
inline void partial_derivatives(std::vector<double>* du, const std::vector<double>& u, size_t n_comps,
                                const Mesh<3>& mesh, const InverseJacobian3& inverse_jacobian) {
  if (!mesh.is_isotropic()) throw std::runtime_error("isotropic meshes only on this path");
  const size_t n = mesh.number_of_grid_points();
  du->resize(3 * n_comps * n);
  const auto J = inverse_jacobian.flat();
  check(dgrhs_partial_derivatives(static_cast<int>(mesh.extents(0)), static_cast<int>(n_comps), u.data(),
                                  J.data(), du->data()));
}

namespace dg {
// dg::mortar_mesh (MortarHelpers.cpp:22-49): the larger extents in every face dimension
inline Mesh<2> mortar_mesh(const Mesh<2>& face_mesh1, const Mesh<2>& face_mesh2) {
  return Mesh<2>({std::max(face_mesh1.extents(0), face_mesh2.extents(0)),
                  std::max(face_mesh1.extents(1), face_mesh2.extents(1))},
                 Spectral::Basis::Legendre, Spectral::Quadrature::GaussLobatto);
}
// MortarHelpers.hpp:60-72
inline bool needs_projection(const Mesh<2>& face_mesh, const Mesh<2>& mortar_mesh,
                             const std::array<Spectral::MortarSize, 2>& mortar_size) {
  return face_mesh.extents(0) != mortar_mesh.extents(0) || face_mesh.extents(1) != mortar_mesh.extents(1) ||
         mortar_size[0] != Spectral::MortarSize::Full || mortar_size[1] != Spectral::MortarSize::Full;
}
// dg::project_to_mortar / project_from_mortar (MortarHelpers.hpp:74-129) on a Variables-like
// block [n_comps][points of the face / mortar mesh], first face dimension fastest
inline std::vector<double> project_to_mortar(const std::vector<double>& vars, size_t n_comps,
                                             const Mesh<2>& face_mesh, const Mesh<2>& mortar_mesh,
                                             const std::array<Spectral::MortarSize, 2>& mortar_size) {
  const int fe[2] = {static_cast<int>(face_mesh.extents(0)), static_cast<int>(face_mesh.extents(1))};
  const int me[2] = {static_cast<int>(mortar_mesh.extents(0)), static_cast<int>(mortar_mesh.extents(1))};
  const int sz[2] = {Spectral::abi_size_code(mortar_size[0]), Spectral::abi_size_code(mortar_size[1])};
  std::vector<double> out(n_comps * mortar_mesh.number_of_grid_points());
  check(dgrhs_project_to_mortar(static_cast<int>(n_comps), fe, me, sz, vars.data(), out.data()));
  return out;
}
inline std::vector<double> project_from_mortar(const std::vector<double>& vars, size_t n_comps,
                                               const Mesh<2>& face_mesh, const Mesh<2>& mortar_mesh,
                                               const std::array<Spectral::MortarSize, 2>& mortar_size) {
  const int fe[2] = {static_cast<int>(face_mesh.extents(0)), static_cast<int>(face_mesh.extents(1))};
  const int me[2] = {static_cast<int>(mortar_mesh.extents(0)), static_cast<int>(mortar_mesh.extents(1))};
  const int sz[2] = {Spectral::abi_size_code(mortar_size[0]), Spectral::abi_size_code(mortar_size[1])};
  std::vector<double> out(n_comps * face_mesh.number_of_grid_points());
  check(dgrhs_project_from_mortar(static_cast<int>(n_comps), fe, me, sz, vars.data(), out.data()));
  return out;
}
enum class Formulation { StrongInertial, WeakInertial };
// dg::lift_flux on a Variables-like block [n_comps][f]
inline void lift_flux(std::vector<double>* boundary_correction_terms, size_t n_comps,
                      size_t extent_perpendicular_to_boundary, const ScalarDV& magnitude_of_face_normal) {
  const size_t f = magnitude_of_face_normal.get().size();
  check(dgrhs_lift_flux(static_cast<int>(f), static_cast<int>(n_comps), boundary_correction_terms->data(),
                        static_cast<int>(extent_perpendicular_to_boundary),
                        magnitude_of_face_normal.get().data()));
}
}  // namespace dg

// ---- ScalarWave ------------------------------------------------------------------
namespace ScalarWave {
template <size_t Dim>
struct TimeDerivative {
  static_assert(Dim == 3);
  static void apply(ScalarDV* dt_psi, ScalarDV* dt_pi, tnsr::i3* dt_phi, ScalarDV* result_gamma2,
                    const tnsr::i3& d_psi, const tnsr::i3& d_pi, const tnsr::ij9& d_phi, const ScalarDV& pi,
                    const tnsr::i3& phi, const ScalarDV& gamma2) {
    const size_t n = pi.get().size();
    std::vector<double> u(5 * n, 0.0), du(15 * n), dt(5 * n);
    // psi itself does not enter the ScalarWave RHS
    for (size_t p = 0; p < n; ++p) {
      u[1 * n + p] = pi.get()[p];
      for (size_t d = 0; d < 3; ++d) {
        u[(2 + d) * n + p] = phi.get(d)[p];
        du[(3 * 0 + d) * n + p] = d_psi.get(d)[p];
        du[(3 * 1 + d) * n + p] = d_pi.get(d)[p];
        for (size_t j = 0; j < 3; ++j) du[(3 * (2 + j) + d) * n + p] = d_phi.get(d, j)[p];
      }
    }
    check(dgrhs_sw_time_derivative(static_cast<int>(n), u.data(), du.data(), gamma2.get().data(), dt.data()));
    *result_gamma2 = gamma2;
    dt_psi->from_flat(dt.data(), n);
    dt_pi->from_flat(dt.data() + n, n);
    dt_phi->from_flat(dt.data() + 2 * n, n);
  }
};

namespace BoundaryCorrections {
template <size_t Dim>
class UpwindPenalty {
 public:
  static_assert(Dim == 3);
  // packaged fields as one Variables<dg_package_field_tags> block [16][f]
  double dg_package_data(std::vector<double>* packaged, const ScalarDV& psi, const ScalarDV& pi,
                         const tnsr::i3& phi, const ScalarDV& constraint_gamma2,
                         const tnsr::i3& normal_covector,
                         const std::optional<tnsr::i3>& /*mesh_velocity*/,
                         const std::optional<ScalarDV>& normal_dot_mesh_velocity) const {
    const size_t f = psi.get().size();
    std::vector<double> u(5 * f);
    for (size_t p = 0; p < f; ++p) {
      u[p] = psi.get()[p];
      u[f + p] = pi.get()[p];
      for (size_t d = 0; d < 3; ++d) u[(2 + d) * f + p] = phi.get(d)[p];
    }
    packaged->resize(16 * f);
    double max_speed = 0.0;
    const auto nrm = normal_covector.flat();
    check(dgrhs_sw_package_data_moving(
        static_cast<int>(f), u.data(), constraint_gamma2.get().data(), nrm.data(),
        normal_dot_mesh_velocity.has_value() ? normal_dot_mesh_velocity->get().data() : nullptr,
        packaged->data(), &max_speed));
    return max_speed;
  }
  void dg_boundary_terms(std::vector<double>* boundary_corrections, const std::vector<double>& packaged_int,
                         const std::vector<double>& packaged_ext, dg::Formulation /*formulation*/) const {
    const size_t f = packaged_int.size() / 16;
    boundary_corrections->resize(5 * f);
    check(dgrhs_sw_boundary_terms(static_cast<int>(f), packaged_int.data(), packaged_ext.data(),
                                  boundary_corrections->data()));
  }
};
}  // namespace BoundaryCorrections
}  // namespace ScalarWave

// ---- GeneralizedHarmonic ------------------------------------------------------------
namespace gh {
namespace gauges {
struct GaugeCondition {
  virtual ~GaugeCondition() = default;
  virtual bool is_harmonic() const = 0;
};
struct Harmonic : GaugeCondition {
  bool is_harmonic() const override { return true; }
};
// gauge source function evaluated by the caller (AnalyticChristoffel on a
// static solution, or the output of gauges::dispatch)
struct Fields : GaugeCondition {
  tnsr::a4 gauge_h;
  tnsr::ab16 d4_gauge_h;
  bool is_harmonic() const override { return false; }
};
}  // namespace gauges

template <size_t Dim>
struct TimeDerivative {
  static_assert(Dim == 3);
  // Outputs, then the temporaries that feed the face projection
  // (dg_package_data_temporary_tags: gamma1, gamma2; lapse/shift/inverse spatial
  // metric are recomputed on the faces by the batched path), then derivatives and
  // arguments in the reference's order.  The other scratch temporaries of the
  // reference signature are private to the GPU kernel and not materialised.
  static void apply(tnsr::aa10* dt_spacetime_metric, tnsr::aa10* dt_pi, tnsr::iaa30* dt_phi,
                    ScalarDV* temp_gamma1, ScalarDV* temp_gamma2, const tnsr::iaa30& d_spacetime_metric,
                    const tnsr::iaa30& d_pi, const tnsr::ijaa90& d_phi, const tnsr::aa10& spacetime_metric,
                    const tnsr::aa10& pi, const tnsr::iaa30& phi, const ScalarDV& gamma0,
                    const ScalarDV& gamma1, const ScalarDV& gamma2,
                    const gauges::GaugeCondition& gauge_condition) {
    const size_t n = gamma0.get().size();
    std::vector<double> u(50 * n), du(150 * n), dt(50 * n);
    const auto pack = [n](double* dst, const auto& t) {
      for (size_t s = 0; s < t.size(); ++s)
        for (size_t p = 0; p < n; ++p) dst[s * n + p] = t[s][p];
    };
    pack(u.data(), spacetime_metric);
    pack(u.data() + 10 * n, pi);
    pack(u.data() + 20 * n, phi);
    // derivative tensors prepend the derivative index: component 3 c + i
    pack(du.data(), d_spacetime_metric);
    pack(du.data() + 30 * n, d_pi);
    pack(du.data() + 60 * n, d_phi);
    const bool harmonic = gauge_condition.is_harmonic();
    std::vector<double> H, dH;
    if (!harmonic) {
      const auto& g = dynamic_cast<const gauges::Fields&>(gauge_condition);
      H = g.gauge_h.flat();
      dH = g.d4_gauge_h.flat();
    }
    check(dgrhs_gh_time_derivative(static_cast<int>(n), u.data(), du.data(), gamma0.get().data(),
                                   gamma1.get().data(), gamma2.get().data(), harmonic ? 1 : 0,
                                   harmonic ? nullptr : H.data(), harmonic ? nullptr : dH.data(), dt.data()));
    *temp_gamma1 = gamma1;
    *temp_gamma2 = gamma2;
    dt_spacetime_metric->from_flat(dt.data(), n);
    dt_pi->from_flat(dt.data() + 10 * n, n);
    dt_phi->from_flat(dt.data() + 20 * n, n);
  }
};

namespace BoundaryCorrections {
template <size_t Dim>
class UpwindPenalty {
 public:
  static_assert(Dim == 3);
  // packaged: Variables<dg_package_field_tags> as one block [134][f]
  double dg_package_data(std::vector<double>* packaged, const tnsr::aa10& spacetime_metric,
                         const tnsr::aa10& pi, const tnsr::iaa30& phi, const ScalarDV& constraint_gamma1,
                         const ScalarDV& constraint_gamma2, const ScalarDV& lapse, const tnsr::i3& shift,
                         const tnsr::i3& normal_covector, const tnsr::i3& normal_vector,
                         const std::optional<tnsr::i3>& /*mesh_velocity*/,
                         const std::optional<ScalarDV>& normal_dot_mesh_velocity) const {
    const size_t f = lapse.get().size();
    std::vector<double> u(50 * f);
    const auto pack = [f](double* dst, const auto& t) {
      for (size_t s = 0; s < t.size(); ++s)
        for (size_t p = 0; p < f; ++p) dst[s * f + p] = t[s][p];
    };
    pack(u.data(), spacetime_metric);
    pack(u.data() + 10 * f, pi);
    pack(u.data() + 20 * f, phi);
    packaged->resize(134 * f);
    double max_speed = 0.0;
    const auto sh = shift.flat(), nl = normal_covector.flat(), nu = normal_vector.flat();
    check(dgrhs_gh_package_data_moving(
        static_cast<int>(f), u.data(), constraint_gamma1.get().data(), constraint_gamma2.get().data(),
        lapse.get().data(), sh.data(), nl.data(), nu.data(),
        normal_dot_mesh_velocity.has_value() ? normal_dot_mesh_velocity->get().data() : nullptr,
        packaged->data(), &max_speed));
    return max_speed;
  }
  void dg_boundary_terms(std::vector<double>* boundary_corrections, const std::vector<double>& packaged_int,
                         const std::vector<double>& packaged_ext, dg::Formulation /*formulation*/) const {
    const size_t f = packaged_int.size() / 134;
    boundary_corrections->resize(50 * f);
    check(dgrhs_gh_boundary_terms(static_cast<int>(f), packaged_int.data(), packaged_ext.data(),
                                  boundary_corrections->data()));
  }
};
}  // namespace BoundaryCorrections

namespace BoundaryConditions {
namespace detail {
enum class ConstraintPreservingBjorhusType { ConstraintPreserving, ConstraintPreservingPhysical };
}
// gh::BoundaryConditions::ConstraintPreservingBjorhus<3> (GeneralizedHarmonic/
// BoundaryConditions/Bjorhus.hpp): dg_time_derivative with the reference's argument
// list; returns the reference's optional error message (none on a static mesh)
template <size_t Dim>
class ConstraintPreservingBjorhus {
 public:
  static_assert(Dim == 3);
  explicit ConstraintPreservingBjorhus(detail::ConstraintPreservingBjorhusType type) : type_(type) {}
  std::optional<std::string> dg_time_derivative(
      tnsr::aa10* dt_spacetime_metric_correction, tnsr::aa10* dt_pi_correction,
      tnsr::iaa30* dt_phi_correction, const std::optional<tnsr::i3>& face_mesh_velocity,
      const tnsr::i3& normal_covector, const tnsr::i3& /*normal_vector*/,
      const tnsr::aa10& spacetime_metric, const tnsr::aa10& pi, const tnsr::iaa30& phi,
      const tnsr::i3& coords, const ScalarDV& gamma1, const ScalarDV& gamma2, const ScalarDV& lapse,
      const tnsr::i3& shift, const tnsr::aa10& inverse_spacetime_metric,
      const tnsr::a4& spacetime_unit_normal_vector, const tnsr::iaa30& three_index_constraint,
      const tnsr::a4& gauge_source, const tnsr::ab16& spacetime_deriv_gauge_source,
      const tnsr::aa10& logical_dt_spacetime_metric, const tnsr::aa10& logical_dt_pi,
      const tnsr::iaa30& logical_dt_phi, const tnsr::iaa30& /*d_spacetime_metric*/,
      const tnsr::iaa30& d_pi, const tnsr::ijaa90& d_phi) const {
    if (face_mesh_velocity.has_value()) throw std::runtime_error("moving meshes are out of scope of this path");
    const size_t n = lapse.get().size();
    std::vector<double> og(10 * n), op(10 * n), oph(30 * n);
    check(dgrhs_gh_bjorhus_dg_time_derivative(
        static_cast<int>(n), type_ == detail::ConstraintPreservingBjorhusType::ConstraintPreservingPhysical,
        normal_covector.flat().data(), spacetime_metric.flat().data(), pi.flat().data(), phi.flat().data(),
        coords.flat().data(), gamma1.get().data(), gamma2.get().data(), lapse.get().data(),
        shift.flat().data(), inverse_spacetime_metric.flat().data(),
        spacetime_unit_normal_vector.flat().data(), three_index_constraint.flat().data(),
        gauge_source.flat().data(), spacetime_deriv_gauge_source.flat().data(),
        logical_dt_spacetime_metric.flat().data(), logical_dt_pi.flat().data(),
        logical_dt_phi.flat().data(), d_pi.flat().data(), d_phi.flat().data(), og.data(), op.data(),
        oph.data()));
    dt_spacetime_metric_correction->from_flat(og.data(), n);
    dt_pi_correction->from_flat(op.data(), n);
    dt_phi_correction->from_flat(oph.data(), n);
    return std::nullopt;
  }

 private:
  detail::ConstraintPreservingBjorhusType type_;
};
}  // namespace BoundaryConditions
}  // namespace gh

// ---- TimeSteppers ------------------------------------------------------------------
namespace TimeSteppers {
namespace adams_coefficients {
// coefficients for a step from step_start to step_end given the history times
// (oldest first): AdamsCoefficients.hpp:64-104
inline std::vector<double> coefficients(const std::vector<double>& history_times, double step_start,
                                        double step_end) {
  std::vector<double> c(history_times.size());
  check(dgrhs_adams_bashforth_coefficients(static_cast<int>(history_times.size()), history_times.data(),
                                           step_start, step_end, c.data()));
  return c;
}
}  // namespace adams_coefficients

// The property queries of TimeStepper (TimeStepper.hpp:47-246) for the steppers the
// path runs; stepping itself happens inside the batched context (DgEvolution below,
// set_time_stepper(stepper.id(), stepper.order(), ...)).
class TimeStepper {
 public:
  virtual ~TimeStepper() = default;
  virtual int id() const = 0;
  size_t order() const { return static_cast<size_t>(props().order); }
  uint64_t number_of_substeps() const { return static_cast<uint64_t>(props().substeps); }
  size_t number_of_past_steps() const { return static_cast<size_t>(props().past_steps); }
  double stable_step() const { return props().stable_step; }

 protected:
  struct Props {
    int order, substeps, past_steps;
    double stable_step;
  };
  virtual size_t requested_order() const { return 0; }
  Props props() const {
    Props p{};
    check(dgrhs_stepper_properties(id(), static_cast<int>(requested_order()), &p.order, &p.substeps,
                                   &p.past_steps, &p.stable_step));
    return p;
  }
};
class AdamsBashforth : public TimeStepper {
 public:
  static constexpr size_t maximum_order = 8;  // AdamsBashforth.hpp:199
  explicit AdamsBashforth(size_t order) : order_(order) { (void)props(); }  // throws on a bad order
  int id() const override { return DGRHS_STEPPER_ADAMS_BASHFORTH; }
  // TimeStepper::update_u (TimeStepper.hpp:96-102, AdamsBashforth.cpp:120-135): u holds the value
  // at the newest history time; history = (time, derivative) oldest first, order() entries
  void update_u(std::vector<double>* u, const std::vector<double>& history_times,
                const std::vector<std::vector<double>>& history_derivatives, double time_step) const {
    if (history_times.size() != order_ || history_derivatives.size() != order_)
      throw std::runtime_error("AdamsBashforth::update_u needs order() history entries");
    std::vector<double> flat;
    for (const auto& d : history_derivatives) flat.insert(flat.end(), d.begin(), d.end());
    check(dgrhs_update_u(id(), static_cast<int>(order_), static_cast<long long>(u->size()), u->data(),
                         static_cast<int>(order_), history_times.data(), flat.data(), nullptr, time_step));
  }

 private:
  size_t requested_order() const override { return order_; }
  size_t order_;
};
#define SPECTRE_B200_RK_STEPPER(NAME, ID)              \
  class NAME : public TimeStepper {                    \
   public:                                             \
    int id() const override { return ID; }             \
  }
SPECTRE_B200_RK_STEPPER(Rk3HesthavenSsp, DGRHS_STEPPER_RK3_HESTHAVEN);
SPECTRE_B200_RK_STEPPER(Rk3Owren, DGRHS_STEPPER_RK3_OWREN);
SPECTRE_B200_RK_STEPPER(Rk3Kennedy, DGRHS_STEPPER_RK3_KENNEDY);
SPECTRE_B200_RK_STEPPER(ClassicalRungeKutta4, DGRHS_STEPPER_RK4);
SPECTRE_B200_RK_STEPPER(DormandPrince5, DGRHS_STEPPER_DORMAND_PRINCE5);
#undef SPECTRE_B200_RK_STEPPER
}  // namespace TimeSteppers

// ---- connectivity tables from Element<3>-style neighbour lists ----------------------
// Collects, per element and direction, what Element<3>::neighbors() holds (the ids of the
// neighbours in that direction and the OrientationMap to their block) or an external-boundary
// marker, and produces the flat tables of dgrhs_set_geometry (neighbors),
// dgrhs_set_neighbor_orientations and dgrhs_set_mortars.  Elements are named by their index in
// the context (the order of the Variables blocks).
class DgConnectivity {
 public:
  explicit DgConnectivity(std::vector<ElementId<3>> element_ids)
      : ids_(std::move(element_ids)),
        neighbors_(6 * ids_.size(), -1),
        directions_(6 * ids_.size()),
        permutations_(6 * ids_.size(), 0) {
    for (size_t e = 0; e < ids_.size(); ++e)
      for (int d = 0; d < 6; ++d) directions_[6 * e + d] = d ^ 1;
  }
  // the neighbours of `element` in `direction`: one (conforming, or coarser than the element)
  // or several (finer; 2:1), all in the block reached through `orientation` (finer neighbours
  // may also be handed over in several calls)
  void set_neighbors(size_t element, const Direction3& direction, const std::vector<size_t>& neighbor_elements,
                     const OrientationMap<3>& orientation) {
    if (element >= ids_.size() || neighbor_elements.empty()) throw std::runtime_error("bad neighbour list");
    const size_t slot = 6 * element + static_cast<size_t>(direction.abi());
    const auto fo = face_orientation(orientation, direction);
    bool any_finer = false, any_coarser = false;
    std::vector<std::array<Spectral::MortarSize, 2>> sizes;
    for (size_t nb : neighbor_elements) {
      if (nb >= ids_.size()) throw std::runtime_error("neighbour index out of range");
      sizes.push_back(dg::mortar_size(ids_[element], ids_[nb], direction.dimension, orientation));
      const auto back = dg::mortar_size(ids_[nb], ids_[element], static_cast<size_t>(fo.neighbor_direction / 2),
                                        orientation.inverse_map());
      for (int k = 0; k < 2; ++k) {
        any_finer = any_finer || sizes.back()[k] != Spectral::MortarSize::Full;
        any_coarser = any_coarser || back[k] != Spectral::MortarSize::Full;
      }
    }
    if (any_finer && any_coarser)
      throw std::runtime_error("a mortar smaller than both faces (finer in one face dimension, coarser in "
                               "the other) is not supported");
    if (!any_finer && !any_coarser) {
      if (neighbor_elements.size() != 1) throw std::runtime_error("several conforming neighbours in one direction");
      neighbors_[slot] = static_cast<int32_t>(neighbor_elements[0]);
      directions_[slot] = fo.neighbor_direction;
      permutations_[slot] = fo.permutation;
      return;
    }
    neighbors_[slot] = DGRHS_NEIGHBOR_HANGING;
    if (any_coarser) return;  // the fine side: the coarse element lists the mortar
    for (size_t k = 0; k < neighbor_elements.size(); ++k) {
      mortars_.insert(mortars_.end(),
                      {static_cast<int32_t>(element), direction.abi(), static_cast<int32_t>(neighbor_elements[k]),
                       fo.neighbor_direction | (fo.permutation << 3), Spectral::abi_size_code(sizes[k][0]),
                       Spectral::abi_size_code(sizes[k][1])});
    }
  }
  // an external face: -1 (no boundary correction), -(slot + 2) (ghost boundary condition),
  // DGRHS_NEIGHBOR_BJORHUS / DGRHS_NEIGHBOR_BJORHUS_PHYSICAL
  void set_external_boundary(size_t element, const Direction3& direction, int32_t code) {
    if (element >= ids_.size() || code >= 0) throw std::runtime_error("bad external boundary code");
    neighbors_[6 * element + static_cast<size_t>(direction.abi())] = code;
  }
  const std::vector<int32_t>& neighbors() const { return neighbors_; }                    // [E][6]
  const std::vector<int32_t>& neighbor_directions() const { return directions_; }        // [E][6]
  const std::vector<int32_t>& face_permutations() const { return permutations_; }        // [E][6]
  const std::vector<int32_t>& mortars() const { return mortars_; }                       // [n][6]
  bool aligned() const {
    for (size_t k = 0; k < neighbors_.size(); ++k)
      if (permutations_[k] != 0 || directions_[k] != static_cast<int32_t>((k % 6) ^ 1)) return false;
    return true;
  }
  // geometry + connectivity of a context in one go
  void apply(dgrhs_ctx* ctx, const double* inv_jacobian, const double* coords) const {
    check(dgrhs_set_geometry(ctx, inv_jacobian, coords, neighbors_.data()));
    if (!aligned()) check(dgrhs_set_neighbor_orientations(ctx, directions_.data(), permutations_.data()));
    if (!mortars_.empty()) check(dgrhs_set_mortars(ctx, static_cast<int>(mortars_.size() / 6), mortars_.data()));
  }

 private:
  std::vector<ElementId<3>> ids_;
  std::vector<int32_t> neighbors_, directions_, permutations_, mortars_;
};

// ---- one rank's share of the element list and its halo bookkeeping -------------------
// The C++ twin of spectre_b200/domain.py::Partition.  The global element list (ordered by
// block and Z-curve like BlockZCurveProcDistribution, ElementDistribution.hpp:33-47) is cut
// into `world` contiguous chunks; this rank's elements are ordered interior first (all
// neighbours local), then boundary, so that the interior range can run while the halo is in
// flight.  Ghost slots: first the faces received from other ranks, sorted by (peer, sender's
// global element, sender's direction, receiver's global element, receiver's direction) -- the
// sender packs in the same order, so the per-peer pieces are contiguous in both buffers --
// then, if requested, one slot per external face for a ghost boundary condition.
// Tables are global on input ([n_elements][6] neighbours / directions / permutations as
// DgConnectivity produces them, mortar rows [n][6]) and local on output.
class DgPartition {
 public:
  DgPartition(const std::vector<int32_t>& neighbors, int world, int rank,
              const std::vector<int32_t>* neighbor_directions = nullptr,
              const std::vector<int32_t>* face_permutations = nullptr,
              const std::vector<int32_t>& mortars = {}, bool boundary_slots = false) {
    const long ne = static_cast<long>(neighbors.size() / 6);
    if (world < 1 || rank < 0 || rank >= world || neighbors.size() % 6 != 0 || mortars.size() % 6 != 0)
      throw std::runtime_error("bad partition arguments");
    auto bound = [&](int r) { return ne * r / world; };
    auto owner = [&](long g) {
      int r = static_cast<int>((g * world) / std::max<long>(ne, 1));
      while (r + 1 <= world - 1 && bound(r + 1) <= g) ++r;
      while (r > 0 && bound(r) > g) --r;
      return r;
    };
    auto dir_of = [&](long g, int d) {
      return neighbor_directions ? (*neighbor_directions)[6 * g + d] : (d ^ 1);
    };
    const long lo = bound(rank), hi = bound(rank + 1);
    const size_t nm = mortars.size() / 6;
    std::vector<char> is_boundary(static_cast<size_t>(hi - lo), 0);
    for (long g = lo; g < hi; ++g)
      for (int d = 0; d < 6; ++d) {
        const int32_t v = neighbors[6 * g + d];
        if (v >= 0 && owner(v) != rank) is_boundary[g - lo] = 1;
        if (v == DGRHS_NEIGHBOR_HANGING && nm == 0) throw std::runtime_error("hanging faces without a mortar table");
      }
    // a mortar group (the mortars of one coarse face) with a side on another rank needs
    // halo data: every local participant is a boundary element
    std::map<std::pair<int32_t, int32_t>, bool> group_remote;
    for (size_t m = 0; m < nm; ++m) {
      const int32_t* r = &mortars[6 * m];
      bool& flag = group_remote[{r[0], r[1]}];
      flag = flag || owner(r[0]) != owner(r[2]);
    }
    for (size_t m = 0; m < nm; ++m) {
      const int32_t* r = &mortars[6 * m];
      if (!group_remote[{r[0], r[1]}]) continue;
      for (int32_t g : {r[0], r[2]})
        if (owner(g) == rank) is_boundary[g - lo] = 1;
    }
    for (long g = lo; g < hi; ++g)
      if (!is_boundary[g - lo]) global_ids_.push_back(static_cast<int32_t>(g));
    n_interior_ = static_cast<int>(global_ids_.size());
    for (long g = lo; g < hi; ++g)
      if (is_boundary[g - lo]) global_ids_.push_back(static_cast<int32_t>(g));
    const size_t nl = global_ids_.size();
    std::map<int32_t, int32_t> g2l;
    for (size_t l = 0; l < nl; ++l) g2l[global_ids_[l]] = static_cast<int32_t>(l);
    local_neighbors_.assign(6 * nl, -1);
    local_directions_.resize(6 * nl);
    local_permutations_.assign(6 * nl, 0);
    using Key = std::array<long, 5>;  // peer, sender element, sender direction, receiver element, its direction
    struct Entry {
      Key key;
      long mortar;   // -1: a conforming face
      int32_t local; // receiver: unused; sender: local element
    };
    std::vector<Entry> recv, send;
    for (size_t l = 0; l < nl; ++l) {
      const long g = global_ids_[l];
      for (int d = 0; d < 6; ++d) {
        local_directions_[6 * l + d] = d ^ 1;
        const int32_t v = neighbors[6 * g + d];
        if (v == DGRHS_NEIGHBOR_HANGING) local_neighbors_[6 * l + d] = DGRHS_NEIGHBOR_HANGING;
        if (v < 0) continue;
        local_directions_[6 * l + d] = dir_of(g, d);
        local_permutations_[6 * l + d] = face_permutations ? (*face_permutations)[6 * g + d] : 0;
        if (owner(v) == rank) {
          local_neighbors_[6 * l + d] = g2l[v];
        } else {
          recv.push_back({{owner(v), v, dir_of(g, d), g, d}, -1, 0});
          send.push_back({{owner(v), g, d, v, dir_of(g, d)}, -1, static_cast<int32_t>(l)});
        }
      }
    }
    for (size_t m = 0; m < nm; ++m) {
      const int32_t* r = &mortars[6 * m];
      const long ec = r[0], dc = r[1], ef = r[2], df = r[3] & 7;
      if (owner(ec) == rank && owner(ef) != rank) {
        recv.push_back({{owner(ef), ef, df, ec, dc}, static_cast<long>(m), 0});
        send.push_back({{owner(ef), ec, dc, ef, df}, static_cast<long>(m), g2l[static_cast<int32_t>(ec)]});
      } else if (owner(ef) == rank && owner(ec) != rank) {
        recv.push_back({{owner(ec), ec, dc, ef, df}, static_cast<long>(m), 0});
        send.push_back({{owner(ec), ef, df, ec, dc}, static_cast<long>(m), g2l[static_cast<int32_t>(ef)]});
      }
    }
    auto by_key = [](const Entry& a, const Entry& b) { return a.key < b.key; };
    std::stable_sort(recv.begin(), recv.end(), by_key);
    std::stable_sort(send.begin(), send.end(), by_key);
    recv_counts_.assign(world, 0);
    send_counts_.assign(world, 0);
    std::map<long, int32_t> mortar_slot;
    for (size_t slot = 0; slot < recv.size(); ++slot) {
      const Entry& e = recv[slot];
      if (e.mortar < 0)
        local_neighbors_[6 * g2l[static_cast<int32_t>(e.key[3])] + e.key[4]] = -(static_cast<int32_t>(slot) + 2);
      else
        mortar_slot[e.mortar] = static_cast<int32_t>(slot);
      ++recv_counts_[e.key[0]];
    }
    n_recv_ = static_cast<int>(recv.size());
    for (size_t m = 0; m < nm; ++m) {
      const int32_t* r = &mortars[6 * m];
      const bool coarse_here = owner(r[0]) == rank, fine_here = owner(r[2]) == rank;
      if (!coarse_here && !fine_here) continue;
      const int32_t lc = coarse_here ? g2l[r[0]] : -(mortar_slot[static_cast<long>(m)] + 2);
      const int32_t lf = fine_here ? g2l[r[2]] : -(mortar_slot[static_cast<long>(m)] + 2);
      local_mortars_.insert(local_mortars_.end(), {lc, r[1], lf, r[3], r[4], r[5]});
    }
    if (boundary_slots)
      for (size_t l = 0; l < nl; ++l)
        for (int d = 0; d < 6; ++d)
          if (neighbors[6 * static_cast<long>(global_ids_[l]) + d] == -1) {
            const int32_t slot = n_recv_ + static_cast<int32_t>(external_faces_.size() / 3);
            external_faces_.insert(external_faces_.end(), {static_cast<int32_t>(l), d, slot});
            local_neighbors_[6 * l + d] = -(slot + 2);
          }
    for (const Entry& e : send) {
      send_map_.insert(send_map_.end(), {e.local, static_cast<int32_t>(e.key[2])});
      ++send_counts_[e.key[0]];
    }
  }
  int n_local() const { return static_cast<int>(global_ids_.size()); }
  int n_interior() const { return n_interior_; }                                   // dgrhs_set_interior_count
  int n_recv() const { return n_recv_; }
  int n_ghost() const { return n_recv_ + static_cast<int>(external_faces_.size() / 3); }  // dgrhs_create
  const std::vector<int32_t>& global_ids() const { return global_ids_; }          // local -> global element
  const std::vector<int32_t>& local_neighbors() const { return local_neighbors_; }        // dgrhs_set_geometry
  const std::vector<int32_t>& local_neighbor_directions() const { return local_directions_; }
  const std::vector<int32_t>& local_face_permutations() const { return local_permutations_; }
  const std::vector<int32_t>& local_mortars() const { return local_mortars_; }    // dgrhs_set_mortars
  const std::vector<int32_t>& send_map() const { return send_map_; }              // dgrhs_set_halo_map: (element, direction)
  const std::vector<int>& send_counts() const { return send_counts_; }            // faces per peer, ncclSend
  const std::vector<int>& recv_counts() const { return recv_counts_; }            // faces per peer, ncclRecv
  const std::vector<int32_t>& external_faces() const { return external_faces_; }  // (element, direction, slot)

 private:
  int n_interior_ = 0, n_recv_ = 0;
  std::vector<int32_t> global_ids_, local_neighbors_, local_directions_, local_permutations_, local_mortars_,
      send_map_, external_faces_;
  std::vector<int> send_counts_, recv_counts_;
};

// ---- batched evolution: the replacement of DgElementArray + step_actions ------------
class DgEvolution {
 public:
  DgEvolution(int system, const Mesh<3>& mesh, int n_elements, int device = 0) {
    check(dgrhs_create(&ctx_, system, static_cast<int>(mesh.extents(0)), n_elements, 0, device));
  }
  ~DgEvolution() { dgrhs_destroy(ctx_); }
  DgEvolution(const DgEvolution&) = delete;
  DgEvolution& operator=(const DgEvolution&) = delete;
  void set_geometry(const double* inv_jacobian, const double* coords, const int32_t* neighbors) {
    check(dgrhs_set_geometry(ctx_, inv_jacobian, coords, neighbors));
  }
  void set_static_fields(const double* fields, int ncomp) { check(dgrhs_set_static_fields(ctx_, fields, ncomp)); }
  void set_variables(const double* u) { check(dgrhs_set_state(ctx_, u)); }
  void get_variables(double* u) { check(dgrhs_get_state(ctx_, u)); }
  void set_time_stepper(int stepper, int order, double t0, double dt) {
    check(dgrhs_set_stepper(ctx_, stepper, order, t0, dt));
  }
  void set_time_stepper(const TimeSteppers::TimeStepper& stepper, double t0, double dt) {
    set_time_stepper(stepper.id(), static_cast<int>(stepper.order()), t0, dt);
  }
  void take_steps(int n) { check(dgrhs_take_steps(ctx_, n)); }
  // dg::Actions::Filter<Filters::Exponential<0>> (ExponentialFilter.cpp:45-76): enabled, it runs
  // after every substep update; apply_exponential_filter() applies it once to the resident state
  void set_exponential_filter(double alpha, unsigned half_power) {
    check(dgrhs_set_exponential_filter(ctx_, 1, alpha, static_cast<int>(half_power)));
  }
  void apply_exponential_filter() { check(dgrhs_apply_exponential_filter(ctx_)); }
  double time() const { return dgrhs_time(ctx_); }
  // Adams-Bashforth local time stepping with fixed step sizes dt_coarse / 2^levels[e] (the
  // LTS executables' take_step + ApplyLtsBoundaryCorrections, ApplyBoundaryCorrections.hpp:
  // 1142-1192): past_states[j - 1] holds, for every element, its state at t0 - j * (its step)
  void start_local_time_stepping(size_t order, double t0, double dt_coarse,
                                 const std::vector<int32_t>& levels,
                                 const std::vector<const double*>& past_states) {
    if (past_states.size() + 1 != order)
      throw std::runtime_error("start_local_time_stepping needs order - 1 past states");
    check(dgrhs_lts_init(ctx_, static_cast<int>(order), t0, dt_coarse, levels.data()));
    for (size_t j = 1; j < order; ++j)
      check(dgrhs_lts_set_past_state(ctx_, static_cast<int>(j), past_states[j - 1]));
  }
  void take_lts_coarse_steps(long long n) {
    long long per_step = 0;
    check(dgrhs_lts_ticks_per_coarse_step(ctx_, &per_step));
    check(dgrhs_lts_take_ticks(ctx_, n * per_step));
  }
  double lts_time() const {
    double t = 0.0;
    check(dgrhs_lts_time(ctx_, &t, nullptr));
    return t;
  }
  // moving mesh: inertial mesh velocity of the current time (nullptr: static mesh)
  void set_mesh_velocity(const double* mesh_velocity) { check(dgrhs_set_mesh_velocity(ctx_, mesh_velocity)); }
  dgrhs_ctx* handle() { return ctx_; }

 private:
  dgrhs_ctx* ctx_ = nullptr;
};

}  // namespace spectre_b200
