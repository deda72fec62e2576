// Multi-GPU evolution driven from C++20 through the C-ABI alone (no Python, no
// torch): BASELINE.json configs[1] in small -- GeneralizedHarmonic gauge wave
// (GaugeWave3D.yaml: A = 0.1, lambda = 1, gamma0/1/2 = 1/-1/1, harmonic gauge) on the
// periodic Brick [0,1]^3, AdamsBashforth order 3 -- partitioned over `world` GPUs of one
// box, one host thread per GPU.  The library does the halo exchange itself
// (dgrhs_comm_init: NCCL send/recv of the cut mortar faces, the counterpart of
// send_data_for_fluxes / receive_boundary_data_global_time_stepping,
// ComputeTimeDerivative.hpp:652-774, ApplyBoundaryCorrections.hpp:205-380) inside
// dgrhs_take_steps; DgPartition (SpectreShims.hpp) builds the per-rank element order, ghost
// slots and send map the way the reference's ElementDistribution cuts the element list
// (DgElementArray.hpp:53-66).  The gathered state is compared bit for bit with a
// single-GPU evolution of the whole domain, and with the exact solution.
//
//   g++ -std=c++20 -O2 -pthread evolve_gauge_wave_multi_gpu.cpp -L.. -ldgrhs -Wl,-rpath,.. \
//     && ./a.out [world=2] [steps=5] [refine=2] [N=6]
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "SpectreShims.hpp"

using namespace spectre_b200;

namespace {

void check(int rc, const char* what) {
  if (rc != 0) throw std::runtime_error(std::string(what) + ": " + dgrhs_last_error());
}

struct Domain {
  size_t N, n, per_dim, n_elem;
  double h;
  std::vector<double> xi;
  std::vector<int32_t> neighbors;  // [n_elem][6] global
  size_t elem(size_t ix, size_t iy, size_t iz) const { return ix + per_dim * (iy + per_dim * iz); }
};

Domain make_domain(size_t N, size_t refine) {
  Domain d;
  d.N = N;
  d.n = N * N * N;
  d.per_dim = size_t{1} << refine;
  d.n_elem = d.per_dim * d.per_dim * d.per_dim;
  d.h = 1.0 / static_cast<double>(d.per_dim);
  d.xi = Spectral::collocation_points(N);
  d.neighbors.resize(d.n_elem * 6);
  for (size_t iz = 0; iz < d.per_dim; ++iz)
    for (size_t iy = 0; iy < d.per_dim; ++iy)
      for (size_t ix = 0; ix < d.per_dim; ++ix)
        for (size_t dim = 0; dim < 3; ++dim)
          for (size_t side = 0; side < 2; ++side) {
            size_t j[3] = {ix, iy, iz};
            j[dim] = (j[dim] + (side ? 1 : d.per_dim - 1)) % d.per_dim;
            d.neighbors[d.elem(ix, iy, iz) * 6 + 2 * dim + side] =
                static_cast<int32_t>(d.elem(j[0], j[1], j[2]));
          }
  return d;
}

// geometry and fields of the elements `ids` (Variables layout: [element][component][point])
struct ElementData {
  std::vector<double> coords, inv_jac, damping;
};

ElementData element_data(const Domain& d, const std::vector<int32_t>& ids) {
  ElementData e;
  const size_t ne = ids.size(), n = d.n, N = d.N;
  e.coords.resize(ne * 3 * n);
  e.inv_jac.assign(ne * 9 * n, 0.0);
  e.damping.resize(ne * 3 * n);
  for (size_t l = 0; l < ne; ++l) {
    const size_t g = static_cast<size_t>(ids[l]);
    const size_t idx[3] = {g % d.per_dim, (g / d.per_dim) % d.per_dim, g / (d.per_dim * d.per_dim)};
    for (size_t p = 0; p < n; ++p) {
      const size_t ijk[3] = {p % N, (p / N) % N, p / (N * N)};
      for (size_t k = 0; k < 3; ++k) {
        e.coords[(l * 3 + k) * n + p] = d.h * (static_cast<double>(idx[k]) + 0.5 * (d.xi[ijk[k]] + 1.0));
        e.inv_jac[(l * 9 + k + 3 * k) * n + p] = 2.0 / d.h;  // InverseJacobian(jhat, i) at jhat + 3 i
      }
      e.damping[(l * 3 + 0) * n + p] = 1.0;   // gamma0
      e.damping[(l * 3 + 1) * n + p] = -1.0;  // gamma1
      e.damping[(l * 3 + 2) * n + p] = 1.0;   // gamma2
    }
  }
  return e;
}

// gr::Solutions::GaugeWave (GaugeWave.hpp:34-50) as GH variables (Phi.cpp:25-48, Pi.cpp:26-55):
// g_tt = -H, g_xx = H, H = 1 - A sin(2 pi (x - t) / lambda); lapse = sqrt(H), shift = 0
void gauge_wave(const Domain& d, const std::vector<double>& coords, size_t ne, double t,
                std::vector<double>* u) {
  const double A = 0.1, omega = 6.283185307179586;
  const size_t n = d.n;
  u->assign(ne * 50 * n, 0.0);
  for (size_t l = 0; l < ne; ++l)
    for (size_t p = 0; p < n; ++p) {
      const double x = coords[(l * 3 + 0) * n + p];
      const double H = 1.0 - A * std::sin(omega * (x - t));
      const double dH = -omega * A * std::cos(omega * (x - t));  // d_x H = -d_t H
      const double lapse = std::sqrt(H);
      double* ue = u->data() + l * 50 * n + p;
      ue[0 * n] = -H;                 // g_tt   (sym index 0)
      ue[4 * n] = H;                  // g_xx   (sym index 4)
      ue[7 * n] = 1.0;                // g_yy
      ue[9 * n] = 1.0;                // g_zz
      ue[(10 + 0) * n] = -dH / lapse;  // Pi_tt = -d_t g_tt / lapse, d_t g_tt = dH
      ue[(10 + 4) * n] = dH / lapse;   // Pi_xx = -d_t g_xx / lapse, d_t g_xx = -dH
      ue[(20 + 0 + 3 * 0) * n] = -dH;  // Phi_x,tt
      ue[(20 + 0 + 3 * 4) * n] = dH;   // Phi_x,xx
    }
}

struct RankResult {
  std::vector<int32_t> ids;
  std::vector<double> u;
  double time = 0.0;
};

void run_rank(const Domain& d, int world, int rank, int steps, double dt,
              std::shared_future<std::array<char, 128>> unique_id, RankResult* out) {
  const DgPartition part(d.neighbors, world, rank);
  const auto& ids = part.global_ids();
  const size_t ne = ids.size();
  const ElementData e = element_data(d, ids);
  dgrhs_ctx* ctx = nullptr;
  check(dgrhs_create(&ctx, DGRHS_SYSTEM_GH, static_cast<int>(d.N), static_cast<int>(ne), part.n_ghost(),
                     rank),
        "dgrhs_create");
  check(dgrhs_set_geometry(ctx, e.inv_jac.data(), e.coords.data(), part.local_neighbors().data()),
        "dgrhs_set_geometry");
  check(dgrhs_set_static_fields(ctx, e.damping.data(), 3), "dgrhs_set_static_fields");
  std::vector<double> u;
  gauge_wave(d, e.coords, ne, 0.0, &u);
  check(dgrhs_set_state(ctx, u.data()), "dgrhs_set_state");
  check(dgrhs_set_interior_count(ctx, part.n_interior()), "dgrhs_set_interior_count");
  if (world > 1) {
    check(dgrhs_set_halo_map(ctx, part.send_map().data(), static_cast<int>(part.send_map().size() / 2)),
          "dgrhs_set_halo_map");
    const std::array<char, 128> id = unique_id.get();
    check(dgrhs_comm_init(ctx, id.data(), rank, world), "dgrhs_comm_init");
    std::vector<int32_t> sc(part.send_counts().begin(), part.send_counts().end());
    std::vector<int32_t> rc(part.recv_counts().begin(), part.recv_counts().end());
    check(dgrhs_set_halo_peers(ctx, sc.data(), rc.data()), "dgrhs_set_halo_peers");
  }
  check(dgrhs_set_stepper(ctx, DGRHS_STEPPER_ADAMS_BASHFORTH, 3, 0.0, dt), "dgrhs_set_stepper");
  check(dgrhs_take_steps(ctx, steps), "dgrhs_take_steps");
  check(dgrhs_get_state(ctx, u.data()), "dgrhs_get_state");
  out->ids = ids;
  out->u = std::move(u);
  out->time = dgrhs_time(ctx);
  dgrhs_destroy(ctx);
}

}  // namespace

int main(int argc, char** argv) {
  const int world = argc > 1 ? std::atoi(argv[1]) : 2;
  const int steps = argc > 2 ? std::atoi(argv[2]) : 5;
  const size_t refine = argc > 3 ? static_cast<size_t>(std::atoi(argv[3])) : 2;
  const size_t N = argc > 4 ? static_cast<size_t>(std::atoi(argv[4])) : 6;
  const double dt = 2e-4;
  try {
    const Domain d = make_domain(N, refine);
    const size_t per = 50 * d.n;
    // multi-GPU run: one host thread per GPU
    std::promise<std::array<char, 128>> id_promise;
    std::shared_future<std::array<char, 128>> id_future = id_promise.get_future().share();
    std::array<char, 128> id{};
    if (world > 1) check(dgrhs_comm_unique_id(id.data()), "dgrhs_comm_unique_id");
    id_promise.set_value(id);
    std::vector<RankResult> results(world);
    std::vector<std::string> errors(world);
    std::vector<std::thread> threads;
    for (int r = 0; r < world; ++r)
      threads.emplace_back([&, r] {
        try {
          run_rank(d, world, r, steps, dt, id_future, &results[r]);
        } catch (const std::exception& err) {
          errors[r] = err.what();
        }
      });
    for (auto& t : threads) t.join();
    for (int r = 0; r < world; ++r)
      if (!errors[r].empty()) throw std::runtime_error("rank " + std::to_string(r) + ": " + errors[r]);
    std::vector<double> gathered(d.n_elem * per);
    for (const RankResult& rr : results)
      for (size_t l = 0; l < rr.ids.size(); ++l)
        std::memcpy(&gathered[static_cast<size_t>(rr.ids[l]) * per], &rr.u[l * per], per * sizeof(double));
    // the same evolution on one GPU
    RankResult single;
    run_rank(d, 1, 0, steps, dt, id_future, &single);
    std::vector<double> reference(d.n_elem * per);
    for (size_t l = 0; l < single.ids.size(); ++l)
      std::memcpy(&reference[static_cast<size_t>(single.ids[l]) * per], &single.u[l * per], per * sizeof(double));
    const bool identical = std::memcmp(gathered.data(), reference.data(), gathered.size() * sizeof(double)) == 0;
    // error against the exact solution
    std::vector<int32_t> all(d.n_elem);
    for (size_t g = 0; g < d.n_elem; ++g) all[g] = static_cast<int32_t>(g);
    const ElementData e = element_data(d, all);
    std::vector<double> exact;
    gauge_wave(d, e.coords, d.n_elem, results[0].time, &exact);
    double err = 0.0;
    for (size_t i = 0; i < exact.size(); ++i) err = std::max(err, std::abs(gathered[i] - exact[i]));
    std::printf("world %d elements %zu N %zu steps %d time %.17g\n", world, d.n_elem, N, steps, results[0].time);
    std::printf("max |u - exact| %.3e\n", err);
    std::printf("multi-GPU state bit-identical to single-GPU state: %s\n", identical ? "yes" : "NO");
    return identical && err < 1e-4 ? 0 : 2;
  } catch (const std::exception& err) {
    std::fprintf(stderr, "ERROR: %s\n", err.what());
    return 1;
  }
}
