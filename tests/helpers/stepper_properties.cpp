// Host-only check of the TimeStepper property shims (no GPU needed): prints one
// line per stepper "name order substeps past_steps stable_step".
#include <cstdio>

#include "../../spectre_b200/host/SpectreShims.hpp"

using namespace spectre_b200;

static void show(const char* name, const TimeSteppers::TimeStepper& s) {
  std::printf("%s %zu %llu %zu %.17g\n", name, s.order(), static_cast<unsigned long long>(s.number_of_substeps()),
              s.number_of_past_steps(), s.stable_step());
}

int main() {
  for (size_t k = 1; k <= TimeSteppers::AdamsBashforth::maximum_order; ++k) {
    char name[32];
    std::snprintf(name, sizeof name, "AdamsBashforth%zu", k);
    show(name, TimeSteppers::AdamsBashforth(k));
  }
  show("Rk3HesthavenSsp", TimeSteppers::Rk3HesthavenSsp{});
  show("Rk3Owren", TimeSteppers::Rk3Owren{});
  show("Rk3Kennedy", TimeSteppers::Rk3Kennedy{});
  show("ClassicalRungeKutta4", TimeSteppers::ClassicalRungeKutta4{});
  show("DormandPrince5", TimeSteppers::DormandPrince5{});
  bool threw = false;
  try { TimeSteppers::AdamsBashforth bad(9); } catch (const std::runtime_error&) { threw = true; }
  std::printf("bad_order_throws %d\n", threw ? 1 : 0);
  return 0;
}
