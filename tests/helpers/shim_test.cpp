// Exercises spectre_b200/host/SpectreShims.hpp on the GPU the way the
// reference's unit tests exercise the operators (Test_TimeDerivative.cpp for
// ScalarWave; layout checks of the tensors).  Prints "SHIM OK" on success.
#include <cmath>
#include <cstdio>
#include <random>

#include "../../spectre_b200/host/SpectreShims.hpp"

using namespace spectre_b200;

int main() {
  const size_t n = 64;
  std::mt19937 gen(7);
  std::uniform_real_distribution<> dist(-1.0, 1.0);
  auto fill = [&](auto& t) {
    for (auto& comp : t)
      for (size_t p = 0; p < comp.size(); ++p) comp[p] = dist(gen);
  };
  // ---- ScalarWave::TimeDerivative<3>::apply vs the closed form ----
  ScalarDV dt_psi(n), dt_pi(n), res_g2(n), pi(n), gamma2(n);
  tnsr::i3 dt_phi(n), d_psi(n), d_pi(n), phi(n);
  tnsr::ij9 d_phi(n);
  fill(pi); fill(gamma2); fill(d_psi); fill(d_pi); fill(phi); fill(d_phi);
  ScalarWave::TimeDerivative<3>::apply(&dt_psi, &dt_pi, &dt_phi, &res_g2, d_psi, d_pi, d_phi, pi, phi, gamma2);
  double err = 0.0;
  for (size_t p = 0; p < n; ++p) {
    err = std::fmax(err, std::fabs(dt_psi.get()[p] + pi.get()[p]));
    double e = -d_phi.get(0, 0)[p];
    e -= d_phi.get(1, 1)[p];
    e -= d_phi.get(2, 2)[p];
    err = std::fmax(err, std::fabs(dt_pi.get()[p] - e));
    for (size_t d = 0; d < 3; ++d)
      err = std::fmax(err, std::fabs(dt_phi.get(d)[p] -
                                     (-d_pi.get(d)[p] + gamma2.get()[p] * (d_psi.get(d)[p] - phi.get(d)[p]))));
    err = std::fmax(err, std::fabs(res_g2.get()[p] - gamma2.get()[p]));
  }
  if (err > 1e-14) { std::printf("ScalarWave shim mismatch %g\n", err); return 1; }
  // ---- gh::TimeDerivative<3>::apply: flat space in harmonic gauge has zero RHS,
  //      and dt g = -lapse Pi when only Pi is non-zero on flat space ----
  tnsr::aa10 g(n), Pi(n), dtg(n), dtPi(n);
  tnsr::iaa30 Phi(n), dg(n), dPi(n), dtPhi(n);
  tnsr::ijaa90 dPhi(n);
  ScalarDV g0(n, 1.0), g1(n, -1.0), g2(n, 1.0), t1(n), t2(n);
  for (size_t p = 0; p < n; ++p) { g.get(0, 0)[p] = -1.0; g.get(1, 1)[p] = g.get(2, 2)[p] = g.get(3, 3)[p] = 1.0; }
  gh::gauges::Harmonic harmonic;
  gh::TimeDerivative<3>::apply(&dtg, &dtPi, &dtPhi, &t1, &t2, dg, dPi, dPhi, g, Pi, Phi, g0, g1, g2, harmonic);
  err = 0.0;
  for (auto& c : dtg) for (size_t p = 0; p < n; ++p) err = std::fmax(err, std::fabs(c[p]));
  for (auto& c : dtPi) for (size_t p = 0; p < n; ++p) err = std::fmax(err, std::fabs(c[p]));
  for (auto& c : dtPhi) for (size_t p = 0; p < n; ++p) err = std::fmax(err, std::fabs(c[p]));
  if (err != 0.0) { std::printf("GH flat-space RHS not zero: %g\n", err); return 1; }
  for (size_t p = 0; p < n; ++p) Pi.get(1, 2)[p] = 1e-3 * dist(gen);
  gh::TimeDerivative<3>::apply(&dtg, &dtPi, &dtPhi, &t1, &t2, dg, dPi, dPhi, g, Pi, Phi, g0, g1, g2, harmonic);
  for (size_t p = 0; p < n; ++p) err = std::fmax(err, std::fabs(dtg.get(2, 1)[p] + Pi.get(1, 2)[p]));
  if (err > 1e-16) { std::printf("GH dt g != -lapse Pi: %g\n", err); return 1; }
  // ---- ConstraintPreservingBjorhus::dg_time_derivative: flat space with no time
  //      derivative and satisfied constraints needs no boundary correction ----
  {
    using Type = gh::BoundaryConditions::detail::ConstraintPreservingBjorhusType;
    const gh::BoundaryConditions::ConstraintPreservingBjorhus<3> bc(Type::ConstraintPreservingPhysical);
    tnsr::aa10 flat(n), zero_aa(n), cg(n), cPi(n);
    tnsr::iaa30 zero_iaa(n), cPhi(n);
    tnsr::ijaa90 zero_ijaa(n);
    tnsr::i3 normal(n), coords(n), shift(n);
    tnsr::a4 t_up(n), H(n);
    tnsr::ab16 dH(n);
    ScalarDV lapse(n, 1.0), gam1(n, -1.0), gam2(n, 1.0);
    for (size_t p = 0; p < n; ++p) {
      flat.get(0, 0)[p] = -1.0; flat.get(1, 1)[p] = flat.get(2, 2)[p] = flat.get(3, 3)[p] = 1.0;
      normal.get(0)[p] = 1.0; coords.get(0)[p] = 10.0; t_up.get(0)[p] = 1.0;
    }
    fill(cg); fill(cPi); fill(cPhi);
    const auto msg = bc.dg_time_derivative(&cg, &cPi, &cPhi, std::nullopt, normal, normal, flat, zero_aa, zero_iaa,
                                           coords, gam1, gam2, lapse, shift, flat, t_up, zero_iaa, H, dH, zero_aa,
                                           zero_aa, zero_iaa, zero_iaa, zero_iaa, zero_ijaa);
    err = 0.0;
    for (auto& c : cg) for (size_t p = 0; p < n; ++p) err = std::fmax(err, std::fabs(c[p]));
    for (auto& c : cPi) for (size_t p = 0; p < n; ++p) err = std::fmax(err, std::fabs(c[p]));
    for (auto& c : cPhi) for (size_t p = 0; p < n; ++p) err = std::fmax(err, std::fabs(c[p]));
    if (msg.has_value() || err != 0.0) { std::printf("Bjorhus flat-space correction not zero: %g\n", err); return 1; }
  }
  // ---- spectral + stepper helpers ----
  const auto D = Spectral::differentiation_matrix(5);
  const auto x = Spectral::collocation_points(5);
  for (size_t i = 0; i < 5; ++i) {  // D differentiates x^3 exactly
    double s = 0.0;
    for (size_t j = 0; j < 5; ++j) s += D[i * 5 + j] * x[j] * x[j] * x[j];
    if (std::fabs(s - 3 * x[i] * x[i]) > 1e-13) { std::printf("D matrix wrong\n"); return 1; }
  }
  {  // restriction of the prolongation of both halves is the identity
    const auto Pl = Spectral::projection_matrix_parent_to_child(5, Spectral::ChildSize::LowerHalf);
    const auto Pu = Spectral::projection_matrix_parent_to_child(5, Spectral::ChildSize::UpperHalf);
    const auto Rl = Spectral::projection_matrix_child_to_parent(5, Spectral::ChildSize::LowerHalf);
    const auto Ru = Spectral::projection_matrix_child_to_parent(5, Spectral::ChildSize::UpperHalf);
    for (size_t i = 0; i < 5; ++i)
      for (size_t j = 0; j < 5; ++j) {
        double s = 0.0;
        for (size_t k = 0; k < 5; ++k) s += Rl[i * 5 + k] * Pl[k * 5 + j] + Ru[i * 5 + k] * Pu[k * 5 + j];
        if (std::fabs(s - (i == j ? 1.0 : 0.0)) > 1e-12) { std::printf("projection matrices wrong\n"); return 1; }
      }
  }
  const auto c = TimeSteppers::adams_coefficients::coefficients({0.0, 1.0, 2.0}, 2.0, 3.0);
  if (std::fabs(c[2] - 23.0 / 12.0) > 1e-15) { std::printf("AB3 coefficients wrong\n"); return 1; }
  // ---- dg::mortar_mesh / project_to_mortar / project_from_mortar: a p-mortar ----
  {
    const Mesh<2> face({4, 5}, Spectral::Basis::Legendre, Spectral::Quadrature::GaussLobatto);
    const Mesh<2> other({6, 3}, Spectral::Basis::Legendre, Spectral::Quadrature::GaussLobatto);
    const Mesh<2> mortar = dg::mortar_mesh(face, other);
    if (mortar.extents(0) != 6 || mortar.extents(1) != 5) { std::printf("mortar_mesh wrong\n"); return 1; }
    const std::array<Spectral::MortarSize, 2> full{Spectral::MortarSize::Full, Spectral::MortarSize::Full};
    if (!dg::needs_projection(face, mortar, full) || dg::needs_projection(mortar, mortar, full)) {
      std::printf("needs_projection wrong\n");
      return 1;
    }
    std::vector<double> v(2 * 20);
    for (auto& x : v) x = dist(gen);
    const auto on_mortar = dg::project_to_mortar(v, 2, face, mortar, full);
    const auto back = dg::project_from_mortar(on_mortar, 2, face, mortar, full);
    if (on_mortar.size() != 2 * 30) { std::printf("project_to_mortar size wrong\n"); return 1; }
    for (size_t k = 0; k < v.size(); ++k)
      if (std::fabs(back[k] - v[k]) > 1e-12) { std::printf("p-mortar round trip wrong\n"); return 1; }
    // the projection to an upper-half mortar interpolates: a linear function stays linear
    const auto xi4 = Spectral::collocation_points(4);
    std::vector<double> lin(20);
    for (size_t b = 0; b < 5; ++b)
      for (size_t a = 0; a < 4; ++a) lin[a + 4 * b] = 2.0 + 3.0 * xi4[a];
    const std::array<Spectral::MortarSize, 2> upper_a{Spectral::MortarSize::UpperHalf, Spectral::MortarSize::Full};
    const auto half = dg::project_to_mortar(lin, 1, face, face, upper_a);
    for (size_t b = 0; b < 5; ++b)
      for (size_t a = 0; a < 4; ++a)
        if (std::fabs(half[a + 4 * b] - (2.0 + 3.0 * 0.5 * (xi4[a] + 1.0))) > 1e-13) {
          std::printf("upper-half projection wrong\n");
          return 1;
        }
  }
  // ---- orient_variables_on_slice: the aligned map is the identity, a flip of the first
  //      face coordinate reverses it ----
  {
    std::vector<double> v(3 * 4);
    for (size_t k = 0; k < v.size(); ++k) v[k] = static_cast<double>(k);
    const auto same = orient_variables_on_slice(v, 1, {3, 4}, 2, OrientationMap<3>{});
    if (same != v) { std::printf("aligned orient_variables_on_slice is not the identity\n"); return 1; }
  }
  // ---- TimeStepper::update_u on a flat span: AB3 with equal steps ----
  {
    const TimeSteppers::AdamsBashforth ab3(3);
    std::vector<double> u{1.0, -2.0}, want = u;
    const std::vector<std::vector<double>> f{{0.5, 1.0}, {-1.0, 2.0}, {0.25, -0.75}};
    const double dt = 0.1;
    ab3.update_u(&u, {0.0, 0.1, 0.2}, f, dt);
    const double ck[3] = {5.0 / 12.0, -4.0 / 3.0, 23.0 / 12.0};
    for (size_t p = 0; p < 2; ++p)
      for (size_t j = 0; j < 3; ++j) want[p] += dt * ck[j] * f[j][p];
    for (size_t p = 0; p < 2; ++p)
      if (std::fabs(u[p] - want[p]) > 1e-15) { std::printf("AdamsBashforth::update_u wrong\n"); return 1; }
  }
  // ---- error behaviour: bad arguments throw with the library's message ----
  bool threw = false;
  try { Mesh<3> m(13, Spectral::Basis::Legendre, Spectral::Quadrature::GaussLobatto); DgEvolution ev(1, m, 8); }
  catch (const std::runtime_error&) { threw = true; }
  if (!threw) { std::printf("expected an error for N = 13\n"); return 1; }
  std::printf("SHIM OK\n");
  return 0;
}
