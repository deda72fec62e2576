// Known answers of the reference's tests/Unit/Time/Test_Slab.cpp, Test_Time.cpp and
// Test_TimeStepId.cpp re-checked on the value types of host/SpectreTime.hpp, the ids that
// next_time_id produces for every stepper of the path, and (argument "gpu") DgTimeLoop
// driving libdgrhs.so in slab mode next to dgrhs_take_steps.
//   g++ -std=c++20 -I../../include -I../../spectre_b200/host time_types_test.cpp -ldgrhs
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

#include "SpectreShims.hpp"
#include "SpectreTime.hpp"

using namespace spectre_b200;

static int failures = 0;
#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) {                                                         \
      ++failures;                                                          \
      std::printf("FAILED line %d: %s\n", __LINE__, #cond);                \
    }                                                                      \
  } while (0)
#define CHECK_THROWS(expr)                                                 \
  do {                                                                     \
    bool threw_ = false;                                                   \
    try {                                                                  \
      (void)(expr);                                                        \
    } catch (const std::exception&) {                                      \
      threw_ = true;                                                       \
    }                                                                      \
    if (!threw_) {                                                         \
      ++failures;                                                          \
      std::printf("FAILED line %d: no exception from %s\n", __LINE__, #expr); \
    }                                                                      \
  } while (0)

template <typename T>
static std::string str(const T& t) {
  std::ostringstream os;
  os << t;
  return os.str();
}
static bool near(double a, double b) { return std::abs(a - b) <= 1e-13 * (std::abs(a) + std::abs(b) + 1e-300); }
// a < b strictly, all six operators consistent (the reference's check_cmp)
template <typename T>
static bool ordered(const T& a, const T& b) {
  return a < b && !(b < a) && a <= b && !(b <= a) && b > a && !(a > b) && b >= a && !(a >= b) && a != b &&
         !(a == b) && a == a && a <= a && a >= a && !(a < a);
}

static void test_rational() {
  CHECK(Rational(6, -8).numerator() == -3 && Rational(6, -8).denominator() == 4);
  CHECK(Rational(1, 3) + Rational(1, 6) == Rational(1, 2));
  CHECK(Rational(1, 3) * Rational(3, 5) == Rational(1, 5));
  CHECK(Rational(1, 3) / Rational(2, 3) == Rational(1, 2));
  CHECK(Rational(1, 3) - 1 == Rational(-2, 3));
  CHECK(Rational(3, 5).value() == 3.0 / 5.0);
  CHECK(Rational(2, 5) < Rational(3, 7) && Rational(-1, 2) < 0);
  CHECK(str(Rational(3, -5)) == "-3/5");
  CHECK_THROWS(Rational(1, 0));
  CHECK_THROWS(Rational(0, 1).inverse());
  CHECK_THROWS(Rational(1 << 30, 1) * Rational(4, 1));
}

static void test_slab() {
  const double tstart = 0.68138945475734402635, tend = 76.34481744714527451379;
  const double tend2 = tend + 1.234, duration = tend2;
  const Slab slab(tstart, tend);
  CHECK(slab.start().value() == tstart);
  CHECK(slab.end().value() == tend);
  CHECK(near(slab.duration().value(), tend - tstart));
  CHECK(slab.end() - slab.start() == slab.duration());
  CHECK(Slab::with_duration_from_start(tstart, duration).start().value() == tstart);
  CHECK(near(Slab::with_duration_from_start(tstart, duration).duration().value(), duration));
  CHECK(Slab::with_duration_to_end(tend, duration).end().value() == tend);
  const Slab next = slab.advance(), prev = slab.retreat();
  CHECK(next.start() == slab.end());
  CHECK(near(next.duration().value(), slab.duration().value()));
  CHECK(prev.end() == slab.start());
  CHECK(slab.advance_towards(slab.duration()) == next);
  CHECK(slab.advance_towards(-slab.duration()) == prev);
  CHECK(slab.with_duration_from_start(duration).start().value() == tstart);
  CHECK(slab.with_duration_to_end(duration).end().value() == tend);
  CHECK(slab.is_followed_by(slab.advance().with_duration_from_start(3)));
  CHECK(slab.is_preceeded_by(slab.retreat().with_duration_to_end(3)));
  CHECK(slab == slab);
  CHECK(slab != Slab(tstart, tend2));
  CHECK(slab != Slab(tend2 / 2., tend));
  CHECK(slab != Slab(tend, tend2));
  CHECK(slab.overlaps(slab));
  CHECK(!slab.overlaps(slab.advance()) && !slab.advance().overlaps(slab));
  CHECK(!slab.overlaps(slab.advance().advance()));
  CHECK(slab.overlaps(Slab(tstart, tend + 1.0)) && Slab(tstart - 1.0, tend).overlaps(slab));
  CHECK(slab.overlaps(Slab(tstart - 1.0, tend + 1.0)));
  CHECK(ordered(Slab(1, 2), Slab(3, 4)));
  CHECK(ordered(Slab(1, 2), Slab(2, 4)));
  CHECK(str(Slab(0.5, 1.5)) == "Slab[0.5,1.5]");
  CHECK_THROWS(Slab(1., 0.));
  CHECK_THROWS(Slab::with_duration_from_start(0., -1.));
  CHECK_THROWS(slab.advance_towards(0 * slab.duration()));
  CHECK_THROWS(Slab(0., 1.) < Slab(0.1, 0.9));
  CHECK_THROWS(Slab(0., 1.) >= Slab(-0.1, 1.1));
}

static void test_time() {
  using R = Rational;
  const double tstart = 0.68138945475734402635, tend = 76.34481744714527451379;
  CHECK(tstart + (tend - tstart) != tend);  // values that trigger rounding errors
  const Slab slab(tstart, tend);
  CHECK(Time(slab, 0).value() == tstart);
  CHECK(Time(slab, 1).value() == tend);
  CHECK(near(Time(slab, R(3, 5)).value(), 2. / 5. * tstart + 3. / 5. * tend));
  CHECK(Time(slab, 0).is_at_slab_start() && !Time(slab, R(1, 2)).is_at_slab_start() && !Time(slab, 1).is_at_slab_start());
  CHECK(!Time(slab, 0).is_at_slab_end() && Time(slab, 1).is_at_slab_end());
  CHECK(Time(slab, 0).is_at_slab_boundary() && !Time(slab, R(1, 2)).is_at_slab_boundary());
  CHECK(ordered(Time(slab, 0), Time(slab, 1)));
  CHECK(ordered(Time(slab, R(2, 5)), Time(slab, R(3, 5))));
  CHECK(ordered(Time(slab, 1), Time(slab.advance(), 1)));
  CHECK(ordered(Time(slab, 0), Time(slab.advance(), 0)));
  CHECK(ordered(Time(slab, R(3, 5)), Time(slab.advance(), R(2, 5))));
  CHECK(Time(slab, R(2, 3)).with_slab(slab) == Time(slab, R(2, 3)));
  {
    const double other = 2. * slab.duration().value();
    const Time a2(slab, 1);
    const Time b2 = a2.with_slab(slab.advance());
    CHECK(b2.slab() == slab.advance() && b2.fraction() == 0);
    const Time c2 = a2.with_slab(slab.with_duration_to_end(other));
    CHECK(c2.slab() == slab.with_duration_to_end(other) && c2.fraction() == 1);
  }
  CHECK(Time(slab, 0).value() == Time(slab.retreat(), 1).value());
  CHECK(Time(slab, 1).value() == Time(slab.advance(), 0).value());
  CHECK(Time(slab, R(3, 5)) - Time(slab, R(1, 5)) == TimeDelta(slab, R(2, 5)));
  CHECK(Time(slab, R(1, 5)) + TimeDelta(slab, R(2, 5)) == Time(slab, R(3, 5)));
  CHECK(Time(slab, R(3, 5)) - TimeDelta(slab, R(2, 5)) == Time(slab, R(1, 5)));
  // slab boundary arithmetic
  CHECK(Time(slab.advance(), 0) - Time(slab, R(2, 3)) == TimeDelta(slab, R(1, 3)));
  CHECK(Time(slab, R(2, 3)) - Time(slab.advance(), 0) == TimeDelta(slab, -R(1, 3)));
  CHECK(Time(slab, 1) - Time(slab.advance(), R(2, 3)) == TimeDelta(slab.advance(), -R(2, 3)));
  CHECK(Time(slab.advance(), R(2, 3)) - Time(slab, 1) == TimeDelta(slab.advance(), R(2, 3)));
  CHECK(Time(slab, 1) + TimeDelta(slab.advance(), R(1, 3)) == Time(slab.advance(), R(1, 3)));
  CHECK(Time(slab, 1) - TimeDelta(slab.advance(), -R(1, 3)) == Time(slab.advance(), R(1, 3)));
  CHECK(Time(slab, 0) + TimeDelta(slab.retreat(), -R(1, 3)) == Time(slab.retreat(), R(2, 3)));
  CHECK(Time(slab, 0) - TimeDelta(slab.retreat(), R(1, 3)) == Time(slab.retreat(), R(2, 3)));
  CHECK(str(Time(slab, R(3, 5))) == str(slab) + ":3/5");
  CHECK(str(Time(slab, 1)) == str(slab) + ":1/1");
  CHECK_THROWS(Time(Slab(0., 1.), -1));
  CHECK_THROWS(Time(Slab(0., 1.), 2));
  CHECK_THROWS(Time(Slab(0., 1.), R(1, 2)).with_slab(Slab(1., 2.)));
  CHECK_THROWS(Time(Slab(0., 1.), 0).with_slab(Slab(1., 2.)));
  CHECK_THROWS(Time(Slab(0., 1.), 1).with_slab(Slab(-1., 0.)));
  {  // comparisons of the same instant named in four slabs (round-off in the slab ends)
    const double other = 2. * slab.duration().value();
    const Time a(slab, 0);
    const Time b = a.with_slab(slab.retreat());
    CHECK(b.slab() == slab.retreat() && b.fraction() == 1);
    const Time c = a.with_slab(slab.with_duration_from_start(other));
    CHECK(c.fraction() == 0);
    const Time d = a.with_slab(slab.retreat().with_duration_to_end(other));
    CHECK(d.fraction() == 1);
    const std::array<Time, 4> same{{a, b, c, d}};
    for (const Time& t1 : same) {
      for (const Time& t2 : same) CHECK(t1 == t2 && !(t1 != t2) && !(t1 < t2) && !(t1 > t2) && t1 <= t2 && t1 >= t2);
      for (const Time& t2 : {b.slab().start(), d.slab().start()}) CHECK(ordered(t2, t1));
      for (const Time& t2 : {a.slab().end(), c.slab().end()}) CHECK(ordered(t1, t2));
    }
  }
  {  // TimeDelta
    const double length = tend - tstart;
    CHECK(TimeDelta(slab, R(3, 5)).fraction() == R(3, 5));
    CHECK(TimeDelta(slab, 0).value() == 0);
    CHECK(near(TimeDelta(slab, 1).value(), length));
    CHECK(near(TimeDelta(slab, -R(1, 5)).value(), -length / 5));
    CHECK(near(TimeDelta(slab, 2).value(), 2 * length));
    CHECK(TimeDelta(slab, R(1, 2)).is_positive() && !TimeDelta(slab, -R(1, 2)).is_positive() &&
          !TimeDelta(slab, 0).is_positive());
    const Slab slab2(10., 14.);
    CHECK(TimeDelta(slab, R(1, 3)).with_slab(slab2).slab() == slab2);
    CHECK(TimeDelta(slab, R(1, 3)).with_slab(slab2).fraction() == R(1, 3));
    CHECK(ordered(TimeDelta(slab, R(2, 5)), TimeDelta(slab, R(3, 5))));
    CHECK(-TimeDelta(slab, R(3, 5)) == TimeDelta(slab, -R(3, 5)));
    CHECK(TimeDelta(slab, R(3, 5)) / TimeDelta(slab, R(3, 5)) == 1.);
    CHECK(near(TimeDelta(slab, R(3, 5)) / TimeDelta(slab, R(2, 5)), 1.5));
    CHECK(near(TimeDelta(slab.advance().with_duration_from_start(2.345), R(2, 3)) / TimeDelta(slab, R(4, 5)),
               2.345 * 2. / 3. / (length * 4. / 5.)));
    CHECK(TimeDelta(slab, R(2, 5)) + Time(slab, R(1, 5)) == Time(slab, R(3, 5)));
    CHECK(TimeDelta(slab, R(1, 5)) + TimeDelta(slab, R(2, 5)) == TimeDelta(slab, R(3, 5)));
    CHECK(TimeDelta(slab, R(1, 5)) - TimeDelta(slab, R(2, 5)) == TimeDelta(slab, -R(1, 5)));
    CHECK(TimeDelta(slab, R(1, 5)) * R(2, 5) == TimeDelta(slab, R(2, 25)));
    CHECK(TimeDelta(slab, R(1, 5)) / R(2, 5) == TimeDelta(slab, R(1, 2)));
    CHECK(R(2, 5) * TimeDelta(slab, R(1, 5)) == TimeDelta(slab, R(2, 25)));
    CHECK(abs(TimeDelta(slab, -R(1, 5))) == TimeDelta(slab, R(1, 5)));
    CHECK(TimeDelta(slab.advance(), R(1, 3)) + Time(slab, 1) == Time(slab.advance(), R(1, 3)));
    CHECK(TimeDelta(slab.retreat(), -R(1, 3)) + Time(slab, 0) == Time(slab.retreat(), R(2, 3)));
    CHECK(str(TimeDelta(slab, R(3, -5))) == str(slab) + ":-3/5");
    CHECK_THROWS(TimeDelta(slab, 1) + TimeDelta(slab2, 1));
  }
}

static void test_time_step_id(const bool forward) {
  using Hash = std::hash<TimeStepId>;
  const Slab slab(1.25, 3.5);
  const Time start = forward ? slab.start() : slab.end();
  const Time end = forward ? slab.end() : slab.start();
  const TimeDelta step = end - start;
  CHECK(TimeStepId(forward, 4, start + step / 3) ==
        TimeStepId(forward, 4, start + step / 3, 0, step, (start + step / 3).value()));
  CHECK(!TimeStepId(forward, 4, start + step / 3, 2, step, (start + step / 2).value()).is_at_slab_boundary());
  CHECK(!TimeStepId(forward, 4, start, 2, step, start.value()).is_at_slab_boundary());
  CHECK(!TimeStepId(forward, 4, start, 1, step, end.value()).is_at_slab_boundary());
  CHECK(!TimeStepId(forward, 4, start + step / 3).is_at_slab_boundary());
  CHECK(TimeStepId(forward, 4, start).is_at_slab_boundary());
  CHECK(TimeStepId(forward, 4, end).is_at_slab_boundary());
  CHECK(TimeStepId(forward, 5, start).slab_number() == 5);
  CHECK(TimeStepId(forward, 5, start).step_time().slab() == slab);
  CHECK(TimeStepId(forward, 5, end).slab_number() == 6);
  CHECK(TimeStepId(forward, 5, end).step_time().slab() == slab.advance_towards(step));
  CHECK(TimeStepId(forward, 4, start + step / 2).next_step(step / 4) == TimeStepId(forward, 4, start + step * 3 / 4));
  CHECK(TimeStepId(forward, 4, start + step / 2, 1, step / 4, end.value()).next_step(step / 4) ==
        TimeStepId(forward, 4, start + step * 3 / 4));
  CHECK(TimeStepId(forward, 4, start + step / 2, 1, step / 4, end.value()).next_substep(step / 4, 1.0 / 8.0) ==
        TimeStepId(forward, 4, start + step / 2, 2, step / 4, (start + step * 17 / 32).value()));
  CHECK(TimeStepId(forward, 4, start + step / 2, 1, step / 4, end.value())
            .next_substep(step / 4, 1.0 / 8.0)
            .substep_time() == (start + step * 17 / 32).value());
  const TimeStepId id(forward, 4, start + step / 3, 2, step / 2, (start + step / 2).value());
  CHECK(id.step_size() == step / 2);
  CHECK(id == id && !(id != id) && id == TimeStepId(id));
  const auto compare = [&](std::int64_t slab_delta, const TimeDelta& step_time_delta, std::int64_t substep_delta,
                           const TimeDelta& substep_time_delta) {
    const TimeStepId id2(id.time_runs_forward(), id.slab_number() + slab_delta, id.step_time() + step_time_delta,
                         id.substep() + static_cast<std::uint64_t>(substep_delta), id.step_size(),
                         id.substep_time() + substep_time_delta.value());
    CHECK(ordered(id, id2));
    CHECK(Hash{}(id) != Hash{}(id2));
  };
  compare(1, 0 * step, 0, 0 * step);
  compare(0, step / 8, 0, 0 * step);
  compare(0, 0 * step, 1, 0 * step);
  compare(1, -step / 4, 0, 0 * step);
  compare(1, 0 * step, -1, 0 * step);
  compare(1, 0 * step, 0, -step / 8);
  compare(0, step / 8, -1, 0 * step);
  compare(0, step / 16, 0, -step / 16);
  compare(0, 0 * step, 1, -step / 8);
  {
    const TimeStepId id2(id.time_runs_forward(), id.slab_number() + 1, id.step_time());
    CHECK(ordered(id, id2));
    CHECK(Hash{}(id) != Hash{}(id2));
  }
  CHECK(str(id) == "4:" + str(id.step_time()) + ":2:" + str(id.substep_time()));
}

// the ids a stepper walks through over two slabs of two steps
static void test_next_time_id() {
  const Slab slab(0.1, 0.1 + 0.3);
  const TimeDelta step = slab.duration() / 2;
  struct Case {
    int stepper, order;
    std::vector<double> fractions;  // substep times as fractions of the step
  };
  const std::vector<Case> cases{{DGRHS_STEPPER_ADAMS_BASHFORTH, 3, {}},
                                {DGRHS_STEPPER_RK3_HESTHAVEN, 0, {1.0, 0.5}},
                                {DGRHS_STEPPER_RK3_OWREN, 0, {12.0 / 23.0, 4.0 / 5.0}},
                                {DGRHS_STEPPER_RK4, 0, {0.5, 0.5, 1.0}},
                                {DGRHS_STEPPER_DORMAND_PRINCE5, 0, {0.2, 0.3, 0.8, 8.0 / 9.0, 1.0}}};
  for (const Case& c : cases) {
    TimeStepId id(true, 0, slab.start());
    TimeDelta dt = step;
    for (int k = 0; k < 4; ++k) {
      const Slab expected_slab = k < 2 ? slab : slab.advance();
      const Time t_step(expected_slab, Rational(k % 2, 2));
      CHECK(id.substep() == 0 && id.step_time() == t_step && id.slab_number() == k / 2);
      CHECK(id.substep_time() == t_step.value());
      for (size_t s = 0; s < c.fractions.size(); ++s) {
        id = next_time_id(c.stepper, c.order, id, dt);
        const double t_next = (k % 2 == 1 ? expected_slab.end() : Time(expected_slab, Rational(1, 2))).value();
        CHECK(id.substep() == s + 1 && id.step_time() == t_step);
        CHECK(id.substep_time() == (1.0 - c.fractions[s]) * t_step.value() + c.fractions[s] * t_next);
      }
      id = next_time_id(c.stepper, c.order, id, dt);
      dt = dt.with_slab(id.step_time().slab());
    }
    CHECK(id.slab_number() == 2 && id.step_time() == slab.advance().advance().start());
  }
}

// DgTimeLoop next to dgrhs_take_steps (both in slab mode): one periodic ScalarWave element
static int test_time_loop_gpu() {
  const size_t N = 4, n = N * N * N;
  const auto xi = Spectral::collocation_points(N);
  std::vector<double> coords(3 * n), inv_jac(9 * n, 0.0), gamma2(n, 0.0), u0(5 * n);
  const double two_pi = 6.283185307179586;
  for (size_t p = 0; p < n; ++p) {
    const size_t ijk[3] = {p % N, (p / N) % N, p / (N * N)};
    double arg = 0.0;
    for (size_t d = 0; d < 3; ++d) {
      coords[d * n + p] = two_pi * 0.5 * (xi[ijk[d]] + 1.0);
      inv_jac[(d + 3 * d) * n + p] = 2.0 / two_pi;
      arg += coords[d * n + p];
    }
    u0[p] = std::sin(arg);
    u0[n + p] = std::sqrt(3.0) * std::cos(arg);
    for (size_t d = 0; d < 3; ++d) u0[(2 + d) * n + p] = std::cos(arg);
  }
  const std::vector<int32_t> neighbors{0, 0, 0, 0, 0, 0};
  const Mesh<3> mesh(N, Spectral::Basis::Legendre, Spectral::Quadrature::GaussLobatto);
  const Slab slab(0.68138945475734402635, 0.68138945475734402635 + 3e-3);
  const int steps_per_slab = 3, steps = 8;
  const std::array<std::array<int, 2>, 4> steppers{{{DGRHS_STEPPER_ADAMS_BASHFORTH, 3},
                                                    {DGRHS_STEPPER_RK3_HESTHAVEN, 3},
                                                    {DGRHS_STEPPER_DORMAND_PRINCE5, 5},
                                                    {DGRHS_STEPPER_ADAMS_BASHFORTH, 8}}};
  for (const auto& st : steppers) {
    std::vector<double> a(5 * n), b(5 * n);
    double ta = 0.0, tb = 0.0;
    for (int which = 0; which < 2; ++which) {
      DgEvolution evolution(DGRHS_SYSTEM_SCALAR_WAVE, mesh, 1);
      evolution.set_geometry(inv_jac.data(), coords.data(), neighbors.data());
      evolution.set_static_fields(gamma2.data(), 1);
      evolution.set_variables(u0.data());
      if (which == 0) {
        DgTimeLoop loop(evolution.handle(), st[0], st[1], slab, steps_per_slab);
        loop.take_steps(steps);
        CHECK(loop.time_step_id().slab_number() == steps / steps_per_slab);
        CHECK(loop.time_step_id().step_time().fraction() == Rational(steps % steps_per_slab, steps_per_slab));
        CHECK(loop.time_step_id().substep_time() == evolution.time());
        evolution.get_variables(a.data());
        ta = evolution.time();
      } else {
        evolution.set_time_stepper(st[0], st[1], slab.start_value(), 1e-3);
        if (dgrhs_set_slab(evolution.handle(), slab.start_value(), slab.end_value(), steps_per_slab)) return 1;
        evolution.take_steps(steps);
        evolution.get_variables(b.data());
        tb = evolution.time();
      }
    }
    CHECK(ta == tb);
    CHECK(std::memcmp(a.data(), b.data(), a.size() * sizeof(double)) == 0);
    double change = 0.0;
    for (size_t k = 0; k < a.size(); ++k) change = std::max(change, std::abs(a[k] - u0[k]));
    CHECK(change > 1e-4);  // the state did move
  }
  return 0;
}


// tests/Unit/Time/Test_ChooseLtsStepSize.cpp:12-31
static void test_choose_lts_step_size() {
  const Slab slab(1., 4.);
  CHECK(choose_lts_step_size(slab.start(), 4.) == slab.duration());
  CHECK(choose_lts_step_size(slab.start(), 10.) == slab.duration());
  CHECK(choose_lts_step_size(slab.start(), 2.) == slab.duration() / 2);
  CHECK(choose_lts_step_size(slab.start(), 1.4) == slab.duration() / 4);
  CHECK(choose_lts_step_size(slab.start() + slab.duration() / 4, 2.) == slab.duration() / 4);
  CHECK(choose_lts_step_size(slab.end(), -2.) == -slab.duration() / 2);
  CHECK(choose_lts_step_size(slab.start(), std::numeric_limits<double>::infinity()) == slab.duration());
  CHECK_THROWS(choose_lts_step_size(slab.start() + slab.duration() / 3, 1.));
}

// tests/Unit/Time/TimeSteppers/Test_AdamsLts.cpp:260-283 (make_time / make_id) and :556-596
// ("AB 2:1 order 3"), :615-640 ("AB 3:1 order 2") on the shim's argument types
static void test_adams_lts() {
  namespace lts = TimeSteppers::adams_lts;
  const int max_time = 16;
  const Slab slab(-max_time, max_time);
  const auto make_time = [&](int t) { return Time(slab, Rational(t + max_time, 2 * max_time)); };
  const auto ids = [&](std::initializer_list<int> ts) {
    std::vector<TimeStepId> v;
    for (int t : ts) v.emplace_back(true, 0, make_time(t));
    return v;
  };
  const auto find = [&](const lts::LtsCoefficients& c, int a, int b, double* out) {
    for (const auto& term : c)
      if (std::get<0>(term).step_time() == make_time(a) && std::get<1>(term).step_time() == make_time(b)) {
        *out = std::get<2>(term);
        return true;
      }
    return false;
  };
  const lts::AdamsScheme ab3{lts::SchemeType::Explicit, 3}, ab2{lts::SchemeType::Explicit, 2};
  {
    const auto large = ids({-8, -4, 0}), small = ids({-4, -2, 0, 2});
    const auto c = lts::lts_coefficients(large, small, make_time(0), make_time(4), ab3, ab3, ab3);
    const struct { int a, b; double v; } want[] = {
        {0, 2, 115.0 / 16.0},  {0, 0, 7.0 / 6.0},    {0, -2, -11.0 / 16.0}, {-4, 2, -115.0 / 24.0},
        {-4, -2, -11.0 / 8.0}, {-4, -4, 5.0 / 6.0},  {-8, 2, 23.0 / 16.0},  {-8, -2, 11.0 / 48.0}};
    CHECK(c.size() == 8);
    for (const auto& w : want) {
      double v = 0.0;
      CHECK(find(c, w.a, w.b, &v) && near(v, w.v));
    }
    const auto s2 = lts::lts_coefficients(small, large, make_time(2), make_time(4), ab3, ab3, ab3);
    double v = 0.0;
    CHECK(s2.size() == 7 && find(s2, 2, 0, &v) && near(v, 115.0 / 16.0) && find(s2, -2, -8, &v) &&
          near(v, -5.0 / 48.0));
  }
  {
    const auto large = ids({-3, 0}), small = ids({-1, 0, 1, 2});
    const auto c = lts::lts_coefficients(large, small, make_time(0), make_time(3), ab2, ab2, ab2);
    double v = 0.0;
    CHECK(c.size() == 7 && find(c, -3, 2, &v) && near(v, -1.0) && find(c, 0, 2, &v) && near(v, 5.0 / 2.0) &&
          find(c, 0, -1, &v) && near(v, -1.0 / 3.0));
    // equal sides: the GTS coefficients on the diagonal
    const auto g = lts::lts_coefficients(small, small, make_time(2), make_time(3), ab2, ab2, ab2);
    CHECK(g.size() == 2 && find(g, 1, 1, &v) && near(v, -0.5) && find(g, 2, 2, &v) && near(v, 1.5));
  }
  // "AM GTS order 3" (:469-503): steps {0}, {1} and the predictor value {1, 1} of the step
  // from 1 to 2 (make_id with step_size_if_substep, :266-281)
  const lts::AdamsScheme am3{lts::SchemeType::Implicit, 3};
  {
    auto steps = ids({0, 1});
    const TimeDelta unit = slab.duration() / (2 * max_time);
    steps.emplace_back(true, 0, make_time(1), 1, unit, 2.0);
    const auto c = lts::lts_coefficients(steps, steps, make_time(1), make_time(2), am3, am3, am3);
    CHECK(c.size() == 3);
    double v = 0.0;
    CHECK(find(c, 0, 0, &v) && near(v, -1.0 / 12.0));
    bool found_predictor = false, found_step = false;
    for (const auto& term : c) {
      if (std::get<0>(term).step_time() != make_time(1)) continue;
      if (std::get<0>(term).substep() == 1 && std::get<1>(term).substep() == 1)
        found_predictor = near(std::get<2>(term), 5.0 / 12.0);
      if (std::get<0>(term).substep() == 0 && std::get<1>(term).substep() == 0)
        found_step = near(std::get<2>(term), 2.0 / 3.0);
    }
    CHECK(found_predictor && found_step);
  }
  // an implicit scheme needs the predictor value
  CHECK_THROWS(lts::lts_coefficients(ids({0, 1}), ids({0, 1}), make_time(1), make_time(2), am3, am3, am3));
}

int main(int argc, char** argv) {
  try {
    test_rational();
    test_slab();
    test_time();
    test_time_step_id(true);
    test_time_step_id(false);
    test_next_time_id();
    test_adams_lts();
    test_choose_lts_step_size();
    if (argc > 1 && std::string(argv[1]) == "gpu" && test_time_loop_gpu()) {
      std::printf("FAILED: %s\n", dgrhs_last_error());
      return 1;
    }
  } catch (const std::exception& e) {
    std::printf("FAILED: exception %s\n", e.what());
    return 1;
  }
  std::printf(failures ? "%d checks failed\n" : "all checks passed\n", failures);
  return failures ? 1 : 0;
}
