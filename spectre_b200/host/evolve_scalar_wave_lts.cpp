// Level-2 integration example in C++20, local time stepping: the plane wave of
// tests/InputFiles/ScalarWave/PlaneWave3D.yaml on the periodic Brick [0, 2 pi]^3 (2^3
// elements, 5 grid points, AdamsBashforth order 3), the elements of the upper half in x taking
// two steps per step of the lower half -- what EvolveScalarWave3D built with local time
// stepping does once its step choosers have settled on those steps.  Driven through the C-ABI
// (dgrhs_lts_*); the Adams-Bashforth histories start from the analytic solution at the
// elements' own past step times (TimeStepperTestUtils::initialize_history).
//
//   g++ -std=c++20 -O2 evolve_scalar_wave_lts.cpp -L.. -ldgrhs -Wl,-rpath,.. && ./a.out [coarse steps]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "SpectreShims.hpp"

using namespace spectre_b200;

int main(int argc, char** argv) {
  const int coarse_steps = argc > 1 ? std::atoi(argv[1]) : 10;
  const size_t N = 5, n = N * N * N, per_dim = 2, n_elem = per_dim * per_dim * per_dim;
  const double two_pi = 6.283185307179586, h = two_pi / per_dim, dt_coarse = 2e-3;
  const int order = 3;
  const Mesh<3> mesh(N, Spectral::Basis::Legendre, Spectral::Quadrature::GaussLobatto);
  const auto xi = Spectral::collocation_points(N);
  const double k[3] = {1.0, 1.0, 1.0}, omega = std::sqrt(3.0);

  std::vector<double> coords(n_elem * 3 * n), inv_jac(n_elem * 9 * n, 0.0), gamma2(n_elem * n, 0.0);
  std::vector<int32_t> neighbors(n_elem * 6), levels(n_elem);
  // x slowest: the elements are then ordered by step-size level (largest steps first)
  auto elem = [&](size_t ix, size_t iy, size_t iz) { return iy + per_dim * (iz + per_dim * ix); };
  for (size_t ix = 0; ix < per_dim; ++ix)
    for (size_t iz = 0; iz < per_dim; ++iz)
      for (size_t iy = 0; iy < per_dim; ++iy) {
        const size_t e = elem(ix, iy, iz);
        levels[e] = static_cast<int32_t>(ix);
        const size_t idx[3] = {ix, iy, iz};
        for (size_t p = 0; p < n; ++p) {
          const size_t ijk[3] = {p % N, (p / N) % N, p / (N * N)};
          for (size_t d = 0; d < 3; ++d) {
            coords[(e * 3 + d) * n + p] = h * (idx[d] + 0.5 * (xi[ijk[d]] + 1.0));
            inv_jac[(e * 9 + d + 3 * d) * n + p] = 2.0 / h;
          }
        }
        for (size_t d = 0; d < 3; ++d)
          for (size_t side = 0; side < 2; ++side) {
            size_t j[3] = {ix, iy, iz};
            j[d] = (j[d] + (side ? 1 : per_dim - 1)) % per_dim;
            neighbors[e * 6 + 2 * d + side] = static_cast<int32_t>(elem(j[0], j[1], j[2]));
          }
      }
  // the solution with every element at its own time t0 - back * (its step)
  auto solution = [&](double t0, int back, std::vector<double>* u) {
    u->assign(n_elem * 5 * n, 0.0);
    for (size_t e = 0; e < n_elem; ++e) {
      const double t = t0 - back * dt_coarse / (1 << levels[e]);
      for (size_t p = 0; p < n; ++p) {
        double arg = -omega * t;
        for (size_t d = 0; d < 3; ++d) arg += k[d] * coords[(e * 3 + d) * n + p];
        (*u)[(e * 5 + 0) * n + p] = std::sin(arg);
        (*u)[(e * 5 + 1) * n + p] = omega * std::cos(arg);
        for (size_t d = 0; d < 3; ++d) (*u)[(e * 5 + 2 + d) * n + p] = k[d] * std::cos(arg);
      }
    }
  };

  try {
    DgEvolution evolution(DGRHS_SYSTEM_SCALAR_WAVE, mesh, static_cast<int>(n_elem));
    evolution.set_geometry(inv_jac.data(), coords.data(), neighbors.data());
    evolution.set_static_fields(gamma2.data(), 1);
    std::vector<double> u, exact;
    solution(0.0, 0, &u);
    evolution.set_variables(u.data());
    std::vector<std::vector<double>> past(order - 1);
    std::vector<const double*> past_ptr;
    for (int j = 1; j < order; ++j) {
      solution(0.0, j, &past[j - 1]);
      past_ptr.push_back(past[j - 1].data());
    }
    evolution.start_local_time_stepping(order, 0.0, dt_coarse, levels, past_ptr);
    evolution.take_lts_coarse_steps(coarse_steps);
    const double time = evolution.lts_time();
    evolution.get_variables(u.data());
    solution(time, 0, &exact);
    const char* names[3] = {"Psi", "Pi", "Phi"};
    const size_t lo[3] = {0, 1, 2}, hi[3] = {1, 2, 5};
    std::printf("time %.17g\n", time);
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
