// Level-2 integration example in C++20: the configuration of the reference's
// tests/InputFiles/ScalarWave/PlaneWave3D.yaml (EvolveScalarWave3D: plane wave
// with wave vector (1,1,1) on the periodic Brick [0, 2 pi]^3, InitialRefinement 1,
// 5 grid points, AdamsBashforth order 3, step 1e-3, completion at t = 0.05)
// driven from C++ through the C-ABI -- geometry and neighbour table as
// domain::creators::Brick + Affine maps provide them (Domain/Creators/
// Rectilinear.cpp, CoordinateMaps/Affine.cpp), initial data as
// ScalarWave::Solutions::PlaneWave (PointwiseFunctions/AnalyticSolutions/
// WaveEquation/PlaneWave.cpp:56-119), errors as ObserveNorms reports them.
//
//   g++ -std=c++20 -O2 evolve_scalar_wave.cpp -L.. -ldgrhs -Wl,-rpath,.. && ./a.out [steps]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "SpectreShims.hpp"

using namespace spectre_b200;

int main(int argc, char** argv) {
  const int steps = argc > 1 ? std::atoi(argv[1]) : 50;
  const size_t N = 5, n = N * N * N, per_dim = 2, n_elem = per_dim * per_dim * per_dim;
  const double two_pi = 6.283185307179586, h = two_pi / per_dim, dt = 1e-3;
  const Mesh<3> mesh(N, Spectral::Basis::Legendre, Spectral::Quadrature::GaussLobatto);
  const auto xi = Spectral::collocation_points(N);
  const double k[3] = {1.0, 1.0, 1.0}, omega = std::sqrt(3.0);

  // Variables layout per element: [component][grid point], xi fastest (Variables.hpp:94-160)
  std::vector<double> coords(n_elem * 3 * n), inv_jac(n_elem * 9 * n, 0.0), gamma2(n_elem * n, 0.0);
  std::vector<int32_t> neighbors(n_elem * 6);
  auto elem = [&](size_t ix, size_t iy, size_t iz) { return ix + per_dim * (iy + per_dim * iz); };
  for (size_t iz = 0; iz < per_dim; ++iz)
    for (size_t iy = 0; iy < per_dim; ++iy)
      for (size_t ix = 0; ix < per_dim; ++ix) {
        const size_t e = elem(ix, iy, iz);
        const size_t idx[3] = {ix, iy, iz};
        for (size_t p = 0; p < n; ++p) {
          const size_t ijk[3] = {p % N, (p / N) % N, p / (N * N)};
          for (size_t d = 0; d < 3; ++d) {
            coords[(e * 3 + d) * n + p] = h * (idx[d] + 0.5 * (xi[ijk[d]] + 1.0));
            inv_jac[(e * 9 + d + 3 * d) * n + p] = 2.0 / h;  // InverseJacobian(jhat, i) at jhat + 3 i
          }
        }
        for (size_t d = 0; d < 3; ++d)
          for (size_t side = 0; side < 2; ++side) {
            size_t j[3] = {ix, iy, iz};
            j[d] = (j[d] + (side ? 1 : per_dim - 1)) % per_dim;  // periodic: the block is its own neighbour
            neighbors[e * 6 + 2 * d + side] = static_cast<int32_t>(elem(j[0], j[1], j[2]));
          }
      }
  auto solution = [&](double t, std::vector<double>* u) {  // (Psi, Pi, Phi_i)
    u->assign(n_elem * 5 * n, 0.0);
    for (size_t e = 0; e < n_elem; ++e)
      for (size_t p = 0; p < n; ++p) {
        double arg = -omega * t;
        for (size_t d = 0; d < 3; ++d) arg += k[d] * coords[(e * 3 + d) * n + p];
        (*u)[(e * 5 + 0) * n + p] = std::sin(arg);
        (*u)[(e * 5 + 1) * n + p] = omega * std::cos(arg);                       // Pi = -dt Psi
        for (size_t d = 0; d < 3; ++d) (*u)[(e * 5 + 2 + d) * n + p] = k[d] * std::cos(arg);
      }
  };

  try {
    DgEvolution evolution(DGRHS_SYSTEM_SCALAR_WAVE, mesh, static_cast<int>(n_elem));
    evolution.set_geometry(inv_jac.data(), coords.data(), neighbors.data());
    evolution.set_static_fields(gamma2.data(), 1);
    std::vector<double> u, exact;
    solution(0.0, &u);
    evolution.set_variables(u.data());
    evolution.set_time_stepper(DGRHS_STEPPER_ADAMS_BASHFORTH, 3, 0.0, dt);
    evolution.take_steps(steps);
    evolution.get_variables(u.data());
    solution(evolution.time(), &exact);
    const char* names[3] = {"Psi", "Pi", "Phi"};
    const size_t lo[3] = {0, 1, 2}, hi[3] = {1, 2, 5};
    std::printf("time %.17g\n", evolution.time());
    for (int b = 0; b < 3; ++b) {
      double s = 0.0;
      for (size_t e = 0; e < n_elem; ++e)
        for (size_t c = lo[b]; c < hi[b]; ++c)
          for (size_t p = 0; p < n; ++p) {
            const double d = u[(e * 5 + c) * n + p] - exact[(e * 5 + c) * n + p];
            s += d * d;
          }
      std::printf("Error(%s) %.17g\n", names[b], std::sqrt(s / (n_elem * n)));
    }
  } catch (const std::runtime_error& err) {
    std::fprintf(stderr, "ERROR: %s\n", err.what());
    return 1;
  }
  return 0;
}
